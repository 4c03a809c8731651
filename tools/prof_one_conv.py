"""One convolution layer in a loop (for ncu).  usage: prof_one_conv.py Cin Cout k dil res N H W [iters]"""
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neuralbarkcalculator_b200 import ops  # noqa: E402

Cin, Cout, k, dil, res, N, H, W = [int(v) for v in sys.argv[1:9]]
iters = int(sys.argv[9]) if len(sys.argv) > 9 else 5
dev = torch.device('cuda:0')
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(N, H, W, Cin, generator=g, device=dev).to(torch.bfloat16)
w = (torch.randn(Cout, k, k, Cin, generator=g, device=dev) / (Cin * k * k) ** 0.5).to(torch.bfloat16)
b = torch.randn(Cout, generator=g, device=dev)
r = torch.randn(N, H, W, Cout, generator=g, device=dev).to(torch.bfloat16) if res else None
pad = dil if k == 3 else 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(iters):
    if i == iters - 1:
        e0.record()
    y = ops.conv_bf16(x, w, b, pad=pad, dil=dil, relu=True, residual=r, impl=1)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
fl = 2.0 * N * H * W * Cout * Cin * k * k
by = (x.numel() + y.numel() * (2 if res else 1)) * 2
print('conv %d->%d k%d d%d res=%d N=%d %dx%d: %.3f ms  %.1f TFLOP/s  %.2f TB/s (algorithmic)' % (Cin, Cout, k, dil, res, N, H, W, ms, fl / ms / 1e9, by / ms / 1e9))
