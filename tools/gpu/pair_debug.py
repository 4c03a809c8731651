"""CTA-pair (cta_group::2) convolution diagnostics (run on the GPU box with NBC_CTA2=15).
Each case is compared with torch's f32 convolution of the same bf16 operands AND with the single-CTA tcgen05 kernel
run in a child process without NBC_CTA2; a failing case prints where the errors sit (which rows / channel halves),
which tells a TMEM-lane / B-half / barrier mistake apart.  Usage: NBC_CTA2=15 python tools/gpu/pair_debug.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from neuralbarkcalculator_b200 import ops  # noqa: E402


def case(Cin, Cout, k, N, H, W, dil=1, res=False, relu=True):
    dev = torch.device('cuda:0')
    g = torch.Generator().manual_seed(Cin + Cout + k + H)
    x = torch.randn(N, H, W, Cin, generator=g).to(torch.bfloat16)
    w = (torch.randn(Cout, k, k, Cin, generator=g) / np.sqrt(Cin * k * k)).to(torch.bfloat16)
    bias = torch.randn(Cout, generator=g)
    pad = dil if k == 3 else 0
    r = torch.randn(N, H, W, Cout, generator=g).to(torch.bfloat16) if res else None
    name = '%d->%d k%d d%d N%d %dx%d res=%d' % (Cin, Cout, k, dil, N, H, W, int(res))
    try:
        y = ops.conv_bf16(x.to(dev), w.to(dev), bias.to(dev), pad=pad, dil=dil, relu=relu,
                          residual=r.to(dev) if res else None, impl=1)
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        print('%-40s EXCEPTION %s' % (name, str(e)[:200]))
        return None
    ref = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), bias, padding=pad, dilation=dil)
    if res:
        ref = ref + r.float().permute(0, 3, 1, 2)
    if relu:
        ref = ref.relu()
    ref = ref.permute(0, 2, 3, 1)
    got = y.float().cpu()
    err = (got - ref).abs()
    tol = 2.0 ** -7 * ref.abs() + 2e-2
    bad = err > tol
    ok = not bool(bad.any())
    print('%-40s max err %.4g  bad %d / %d  %s' % (name, float(err.max()), int(bad.sum()), bad.numel(), 'ok' if ok else 'FAIL'))
    if not ok:
        flat = bad.view(-1, Cout)            # pixel-major
        rows = flat.any(dim=1).nonzero().flatten()
        cols = flat.any(dim=0).nonzero().flatten()
        print('   bad pixels: %d, first %s last %s; (pixel %% 256 < 128): %d;  bad channels: %d, first %s last %s'
              % (len(rows), rows[:4].tolist(), rows[-4:].tolist(), int(((rows % 256) < 128).sum()), len(cols),
                 cols[:4].tolist(), cols[-4:].tolist()))
        print('   got[0,0,0,:4] %s ref %s' % (got[0, 0, 0, :4].tolist(), ref[0, 0, 0, :4].tolist()))
    return ok


if __name__ == '__main__':
    print('NBC_CTA2 =', os.environ.get('NBC_CTA2'))
    allok = True
    # one pair, one K block; then more K, more tiles, N halves, residual, 3x3, odd tile counts (phantom tile)
    for args in [dict(Cin=64, Cout=256, k=1, N=1, H=2, W=128, relu=False),
                 dict(Cin=64, Cout=128, k=1, N=1, H=2, W=128, relu=False),
                 dict(Cin=512, Cout=256, k=1, N=1, H=2, W=128),
                 dict(Cin=256, Cout=512, k=1, N=1, H=4, W=128),
                 dict(Cin=256, Cout=256, k=1, N=1, H=3, W=128),
                 dict(Cin=1024, Cout=256, k=1, N=2, H=77, W=128),
                 dict(Cin=256, Cout=1024, k=1, N=2, H=77, W=128, res=True),
                 dict(Cin=256, Cout=256, k=3, N=1, H=33, W=128, dil=2),
                 dict(Cin=128, Cout=128, k=3, N=2, H=27, W=40),
                 dict(Cin=2048, Cout=512, k=1, N=3, H=64, W=128),
                 dict(Cin=512, Cout=2048, k=1, N=3, H=64, W=128, res=True)]:
        ok = case(**args)
        allok = allok and bool(ok)
        if ok is None:      # the CUDA context is gone after a device-side trap
            break
    print('pair_debug', 'ALL OK' if allok else 'FAILURES')
    sys.exit(0 if allok else 1)
