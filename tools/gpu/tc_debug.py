"""Diagnostics for the tcgen05 convolution (run on the GPU box).  Structured inputs make layout / descriptor
mistakes visible as permutations instead of noise.  Usage: python tools/tc_debug.py [case ...]"""
import sys
import os
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from neuralbarkcalculator_b200 import ops  # noqa: E402


def run(name, x, w, bias, ref_fn, **kw):
    dev = torch.device('cuda:0')
    y_tc = ops.conv_bf16(x.to(dev), w.to(dev), bias.to(dev), impl=1, **kw)
    torch.cuda.synchronize()
    y_mma = ops.conv_bf16(x.to(dev), w.to(dev), bias.to(dev), impl=2, **kw)
    torch.cuda.synchronize()
    a, b = y_tc.float().cpu(), y_mma.float().cpu()
    ref = ref_fn()
    print('%-28s tc-vs-ref max %.4g | mma-vs-ref max %.4g | tc-vs-mma max %.4g | nonzero tc %d / ref %d'
          % (name, (a - ref).abs().max(), (b - ref).abs().max(), (a - b).abs().max(), int((a != 0).sum()), int((ref != 0).sum())))
    return a, b, ref


def case_identity():
    # 1x1 conv, 128 pixels (one tile), Cin=Cout=64: x one-hot per pixel, w[o][c] = small distinct values
    x = torch.zeros(1, 1, 128, 64)
    for i in range(128):
        x[0, 0, i, i % 64] = 1.0
    w = (torch.arange(64 * 64).float().view(64, 1, 1, 64) % 251) / 256.0
    bias = torch.zeros(64)
    xb, wb = x.to(torch.bfloat16), w.to(torch.bfloat16)
    ref = lambda: torch.einsum('nhwc,okjc->nhwo', xb.float(), wb.float())
    a, b, r = run('identity 1x1 64->64', xb, wb, bias, ref)
    if (a - r).abs().max() > 1e-3:
        np.save('gpurun_out/tc_identity.npy', a.numpy())
        print('  row0 got', a[0, 0, 0, :8].tolist())
        print('  row0 ref', r[0, 0, 0, :8].tolist())
        print('  row1 got', a[0, 0, 1, :8].tolist())
        print('  row1 ref', r[0, 0, 1, :8].tolist())


def case_random(Cin, Cout, k, H, W, dil=1, stride=1, N=1):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(N, H, W, Cin, generator=g).to(torch.bfloat16)
    w = (torch.randn(Cout, k, k, Cin, generator=g) / np.sqrt(Cin * k * k)).to(torch.bfloat16)
    bias = torch.randn(Cout, generator=g)
    pad = dil if k == 3 else 0
    ref = lambda: torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), bias, stride=stride,
                                             padding=pad, dilation=dil).permute(0, 2, 3, 1)
    run('rand %d->%d k%d s%d d%d %dx%dx%d' % (Cin, Cout, k, stride, dil, N, H, W), x, w, bias, ref, stride=stride, pad=pad, dil=dil)


if __name__ == '__main__':
    os.makedirs('gpurun_out', exist_ok=True)
    case_identity()
    case_random(64, 64, 1, 1, 128)
    case_random(128, 64, 1, 1, 128)
    case_random(64, 128, 1, 1, 128)
    case_random(64, 256, 1, 1, 128)
    case_random(64, 64, 1, 8, 16)
    case_random(64, 64, 3, 8, 16)
    case_random(256, 256, 3, 16, 128, dil=2)
    case_random(128, 128, 3, 33, 64, stride=2)
    case_random(512, 1024, 1, 64, 128, N=3)
    print('tc_debug done')
