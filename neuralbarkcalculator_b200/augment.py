"""Training-time augmentation on the GPU (row N4): the reference's loader chain (``__main__.py:153-176``: pad_resize ->
ColorJitter(saturation=0.2, brightness=0.1) -> RandomCrop -> random flips, the same draw for image and label,
``dataset.py:171-179``) as one fused gather kernel over a device-resident dataset (``in_memory=True`` in the reference).
``draw_params`` makes the random draws with the distributions of the torchvision transforms; ``augment_batch`` applies
them (csrc/augment.cu)."""
import ctypes as C

import numpy as np
import torch

from . import _lib

PARAM_DTYPE = np.dtype([('src', '<i4'), ('x0', '<i4'), ('y0', '<i4'), ('hflip', '<i4'), ('vflip', '<i4'), ('order', '<i4'),
                        ('brightness', '<f4'), ('saturation', '<f4')])


def draw_params(rng, batch, n_sources, crop, target_hw=(1024, 1024), saturation=0.2, brightness=0.1, sources=None):
    """rng: numpy Generator.  ColorJitter draws each factor uniformly in [max(0, 1 - v), 1 + v] and applies the two
    ops in random order; RandomCrop draws the offset uniformly; each flip has probability 0.5 (torchvision transforms)."""
    p = np.zeros(batch, dtype=PARAM_DTYPE)
    p['src'] = rng.integers(0, n_sources, batch) if sources is None else np.asarray(sources)
    p['x0'] = rng.integers(0, target_hw[1] - crop + 1, batch)
    p['y0'] = rng.integers(0, target_hw[0] - crop + 1, batch)
    p['hflip'] = rng.random(batch) < 0.5
    p['vflip'] = rng.random(batch) < 0.5
    p['order'] = rng.integers(0, 2, batch)
    p['brightness'] = rng.uniform(max(0.0, 1 - brightness), 1 + brightness, batch) if brightness > 0 else 0.0
    p['saturation'] = rng.uniform(max(0.0, 1 - saturation), 1 + saturation, batch) if saturation > 0 else 0.0
    return p


def augment_batch(images, duals, params, crop, target_hw=(1024, 1024)):
    """images u8 CUDA [M,Hs,Ws,3]; duals u8 CUDA [M,Hs,Ws] (0/127/255) or None; params: structured array (PARAM_DTYPE).
    Returns (u8 [B,crop,crop,3], u8 class map [B,crop,crop] or None) on the device -- what Trainer.step consumes."""
    lib = _lib.load()
    if not images.is_cuda or images.dtype != torch.uint8 or images.dim() != 4 or images.shape[3] != 3:
        raise RuntimeError('augment_batch: images must be a CUDA u8 tensor [M,Hs,Ws,3]')
    images = images.contiguous()
    M, Hs, Ws, _ = images.shape
    if duals is not None:
        if not duals.is_cuda or duals.dtype != torch.uint8 or tuple(duals.shape) != (M, Hs, Ws):
            raise RuntimeError('augment_batch: duals must be a CUDA u8 tensor [M,Hs,Ws]')
        duals = duals.contiguous()
    params = np.ascontiguousarray(params, dtype=PARAM_DTYPE)
    B = params.shape[0]
    if B == 0 or params['src'].min() < 0 or params['src'].max() >= M:
        raise RuntimeError('augment_batch: source index out of range')
    if (params['x0'].min() < 0 or params['y0'].min() < 0 or params['x0'].max() + crop > target_hw[1]
            or params['y0'].max() + crop > target_hw[0]):
        raise RuntimeError('augment_batch: crop window outside the padded image')
    with torch.cuda.device(images.device):
        pd = torch.from_numpy(params.view(np.uint8).reshape(B, -1).copy()).to(images.device)
        out = torch.empty((B, crop, crop, 3), dtype=torch.uint8, device=images.device)
        cls = torch.empty((B, crop, crop), dtype=torch.uint8, device=images.device) if duals is not None else None
        _lib.check(lib.nbc_augment_batch(C.c_void_p(images.data_ptr()), C.c_void_p(duals.data_ptr() if duals is not None else 0),
                                         M, Hs, Ws, target_hw[0], target_hw[1], crop, C.c_void_p(pd.data_ptr()), B,
                                         C.c_void_p(out.data_ptr()), C.c_void_p(cls.data_ptr() if cls is not None else 0),
                                         C.c_void_p(torch.cuda.current_stream(images.device).cuda_stream)), 'nbc_augment_batch')
    return out, cls
