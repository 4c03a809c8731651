"""Oracle for the training-time augmentation (reference ``__main__.py:153-176`` get_loader_for_crop_batch with
``utils.py:242-247`` pad_resize and ``dataset.py:171-193``): the reference's own chain of torchvision / PIL operations,
applied with EXPLICIT parameters instead of random draws.  Test infrastructure only."""
from math import ceil

import numpy as np
import torch
from PIL import Image
from torchvision.transforms import functional as F


def pad_resize(image, width, height):
    """utils.py:242-247."""
    image = F.pad(image, (ceil((width - image.width) / 2), ceil((height - image.height) / 2)), padding_mode='reflect')
    return F.resize(image, (height, width))


def _chain(img, p, target_hw, crop):
    img = pad_resize(img, target_hw[1], target_hw[0])
    ops = [('b', p['brightness']), ('s', p['saturation'])]
    if p['order'] == 1:
        ops.reverse()
    for kind, f in ops:                       # ColorJitter(saturation=.., brightness=..) with its draws made explicit
        if f > 0:
            img = F.adjust_brightness(img, f) if kind == 'b' else F.adjust_saturation(img, f)
    img = F.crop(img, p['y0'], p['x0'], crop, crop)          # RandomCrop
    if p['hflip']:
        img = F.hflip(img)
    if p['vflip']:
        img = F.vflip(img)
    return img


def augment(images, duals, params, crop, target_hw):
    """images: list of u8 [H,W,3] arrays, duals: list of u8 [H,W] (0/127/255); params: list of dicts (src, x0, y0, hflip,
    vflip, order, brightness, saturation).  Returns (u8 [B,crop,crop,3], u8 classes [B,crop,crop]) -- the sample just
    before ToTensor, and the target after ToTensor * 2, round (dataset.py:184-193)."""
    out_i, out_c = [], []
    for p in params:
        s = _chain(Image.fromarray(images[p['src']]), p, target_hw, crop)
        t = _chain(Image.fromarray(duals[p['src']], mode='L'), p, target_hw, crop)
        out_i.append(np.asarray(s))
        tt = F.to_tensor(t)
        tt = tt * 2
        tt.round_()
        out_c.append(tt.long().squeeze().numpy().astype(np.uint8))
    return np.stack(out_i), np.stack(out_c)
