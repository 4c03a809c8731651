"""Seeded synthetic inputs.  Test / bench infrastructure only -- there is no real dataset offline."""
import os
import struct
import numpy as np
import torch
import torch.nn.functional as F

MEAN = np.array([0.7399, 0.6139, 0.4401], dtype=np.float32)
STD = np.array([0.1068, 0.1272, 0.1271], dtype=np.float32)
WOOD_TYPES = ['epinette_gelee', 'epinette_non_gelee', 'sapin']  # dataset.py:50


def texture_u8(h, w, seed, coarse=16):
    """Wood-like correlated texture: u8 RGB [h, w, 3] whose normalised version is ~N(0,1) per channel."""
    g = torch.Generator().manual_seed(int(seed))
    lo = torch.randn(1, 3, max(4, h // coarse), max(4, w // coarse), generator=g)
    f = F.interpolate(lo, size=(h, w), mode='bicubic', align_corners=False)[0]
    f = f + 0.35 * torch.randn(3, h, w, generator=g)
    f = (f - f.mean(dim=(1, 2), keepdim=True)) / f.std(dim=(1, 2), keepdim=True)
    img = (torch.from_numpy(MEAN).view(3, 1, 1) + torch.from_numpy(STD).view(3, 1, 1) * f) * 255.0
    return img.clamp_(0, 255).round_().to(torch.uint8).permute(1, 2, 0).contiguous().numpy()


def raw_image_u8(seed, size=4096, top=None, bottom=None, specks=64):
    """Raw 4096^2 scan: texture with exact-zero dark bands (rows [0,top) and [size-bottom,size)) + black specks."""
    rng = np.random.default_rng(seed)
    if top is None:
        top = int(rng.integers(400, 1201))
    if bottom is None:
        bottom = int(rng.integers(400, 1201))
    img = texture_u8(size, size, seed, coarse=64)
    img[:top] = 0
    img[size - bottom:] = 0
    ys = rng.integers(top, size - bottom, specks)
    xs = rng.integers(0, size, specks)
    img[ys, xs] = 0
    return img, top, bottom


def write_bmp(path, rgb_u8):
    """24-bit bottom-up BGR BMP (what the scanner produces; dataset.py:82-90 reads it through PIL)."""
    h, w, _ = rgb_u8.shape
    row = (w * 3 + 3) & ~3
    pad = row - w * 3
    size = 54 + row * h
    with open(path, 'wb') as f:
        f.write(b'BM' + struct.pack('<IHHI', size, 0, 0, 54))
        f.write(struct.pack('<IiiHHIIiiII', 40, w, h, 1, 24, 0, row * h, 2835, 2835, 0, 0))
        bgr = rgb_u8[::-1, :, ::-1]
        if pad == 0:
            f.write(np.ascontiguousarray(bgr).tobytes())
        else:
            buf = np.zeros((h, row), dtype=np.uint8)
            buf[:, :w * 3] = bgr.reshape(h, w * 3)
            f.write(buf.tobytes())


def class_mask(h, w, seed, freqs=(0.833, 0.164, 0.003)):
    """Blobby 3-class mask with the reference's class prior (utils.py:72-73) from thresholded low-pass noise."""
    g = torch.Generator().manual_seed(int(seed))
    lo = torch.randn(1, 1, max(4, h // 24), max(4, w // 24), generator=g)
    f = F.interpolate(lo, size=(h, w), mode='bicubic', align_corners=False)[0, 0]
    f = f + 0.25 * torch.randn(h, w, generator=g)
    q1 = torch.quantile(f.flatten()[:1 << 20], freqs[0])
    q2 = torch.quantile(f.flatten()[:1 << 20], freqs[0] + freqs[1])
    m = torch.zeros(h, w, dtype=torch.uint8)
    m[f > q1] = 1
    m[f > q2] = 2
    return m.numpy()


def make_raw_folder(root, n_images, size=4096, pool=None, seed0=0):
    """ROOT/samples/<wood>/img_XXXX.bmp; ``pool`` unique images are generated, the rest are hard links."""
    pool = n_images if pool is None else min(pool, n_images)
    made = []
    per = [n_images // 3 + (1 if i < n_images % 3 else 0) for i in range(3)]
    idx = 0
    for wood, n in zip(WOOD_TYPES, per):
        d = os.path.join(root, 'samples', wood)
        os.makedirs(d, exist_ok=True)
        for k in range(n):
            p = os.path.join(d, 'img_%04d.bmp' % idx)
            if idx < pool:
                img, _, _ = raw_image_u8(seed0 + idx, size)
                write_bmp(p, img)
                made.append(p)
            else:
                src = made[idx % pool]
                if os.path.exists(p):
                    os.remove(p)
                os.link(src, p)
            idx += 1
    return root
