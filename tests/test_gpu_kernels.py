"""GPU parity tests (run on a B200: ``pytest -m gpu``).  Every test calls the CUDA path through the C-ABI
(neuralbarkcalculator_b200.ops -> libnbc.so) and compares with the CPU oracle / golden fixtures.

Bars: bit-exact for byte / integer / index work (resize+trim, region removal, class counts, argmax given logits);
floating point within the tolerance written in each test."""
import os

import numpy as np
import pytest
import torch

from oracle import losses as olosses
from oracle import model as omodel
from oracle import postprocess as opost
from oracle import preprocess as opre
from oracle import synth

pytestmark = pytest.mark.gpu
IMPLS = [int(v) for v in os.environ.get('NBC_TEST_IMPLS', '2,1').split(',')]   # 1 = tcgen05, 2 = mma.sync


def _synth():
    from oracle import synth
    return synth


def _ops():
    from neuralbarkcalculator_b200 import ops
    return ops


# ------------------------------------------------------------------------------------------------- K1 preprocess
def _run_pre(raw_np, dev, bgr=False, bottom_up=False):
    ops = _ops()
    H, W, _ = raw_np.shape
    src = raw_np
    if bgr:
        src = src[:, :, ::-1]
    if bottom_up:
        src = src[::-1]
    t = torch.from_numpy(np.ascontiguousarray(src)).to(dev)
    out, fl = ops.preprocess_4x(t.view(-1), H, W, bgr=bgr, bottom_up=bottom_up)
    first, last = fl.tolist()
    return out[:(last - first) * (W // 4) * 3].view(last - first, W // 4, 3).cpu().numpy(), first, last


@pytest.mark.parametrize('bgr,bottom_up', [(False, False), (True, True), (True, False)])
def test_preprocess_golden(cuda_device, golden_dir, bgr, bottom_up):
    g = np.load(os.path.join(golden_dir, 'preprocess_small.npz'))
    out, first, last = _run_pre(g['raw'], cuda_device, bgr, bottom_up)
    assert (first, last) == (int(g['first']), int(g['last']))
    assert np.array_equal(out, g['out'])


def test_preprocess_full_size_and_edges(cuda_device):
    raw, _, _ = synth.raw_image_u8(seed=3, size=4096, top=801, bottom=1199)
    exp, f, l = opre.preprocess_u8(raw)
    out, first, last = _run_pre(raw, cuda_device, bgr=True, bottom_up=True)
    assert (first, last) == (f, l) and np.array_equal(out, exp)
    # all dark -> nothing kept -> whole image (argmax of all-False is 0)
    z = np.zeros((256, 256, 3), np.uint8)
    out, first, last = _run_pre(z, cuda_device)
    assert (first, last) == (0, 64) and not out.any()
    # min >= 1 -> every pixel non-dark; max < 255 and min > 0 -> the clip to [min, max] is exercised
    t = synth.texture_u8(256, 256, 9)
    t = np.clip(t, 40, 200).astype(np.uint8)
    t[::7, ::5] = 40
    t[3::7, 2::5] = 200
    big = np.ascontiguousarray(np.tile(t, (4, 4, 1))[:1024, :1024])     # 1024 -> 256
    S = opre.resize4x_S(big)
    Sc = np.clip(S, 256 * int(big.min()), 256 * int(big.max()))
    expect = ((Sc + 128) >> 8).astype(np.uint8)
    assert (S < 256 * int(big.min())).any() or (S > 256 * int(big.max())).any()     # clipping really happens
    out, first, last = _run_pre(big, cuda_device)
    assert (first, last) == (0, 256) and np.array_equal(out, expect)
    # non-square (no trim) and a width that is not a multiple of 16 (unaligned path)
    r = synth.texture_u8(64, 40, 2)
    r[:16] = 0
    out, first, last = _run_pre(r, cuda_device)
    S = opre.resize4x_S(r)
    expect = ((np.clip(S, 256 * int(r.min()), 256 * int(r.max())) + 128) >> 8).astype(np.uint8)
    assert (first, last) == (0, 16) and np.array_equal(out, expect)


def test_preprocess_zero_band_span(cuda_device):
    """Dark bands that never cross PCIe: the host scan finds the all-zero rows above / below the content, only the span in
    between is given to K1 (nbc_preprocess_4x_span_u8) and the result -- resized bytes, [first, last), the global
    min / max clip -- is bit-identical to K1 on the whole scan."""
    ops = _ops()
    cases = []
    raw, _, _ = synth.raw_image_u8(seed=11, size=1024, top=203, bottom=118)          # bands that are not multiples of 4
    cases.append(('bands', raw))
    r2 = raw.copy()
    r2[37, 500, 1] = 9                                                               # a stray byte inside the top band
    r2[1024 - 5, 3, 2] = 1                                                           # ... and near the bottom edge
    cases.append(('stray bytes', r2))
    r3 = np.clip(synth.texture_u8(512, 512, 12), 30, 220).astype(np.uint8)           # no zero anywhere: min = 30 clips
    cases.append(('no bands', r3))
    r4 = r3.copy()
    r4[:64] = 0                                                                      # zero band => global min 0 although
    cases.append(('band changes the clip', r4))                                      # the copied rows have min 30
    cases.append(('all zero', np.zeros((256, 256, 3), np.uint8)))
    for name, img in cases:
        for bgr, bottom_up in ((False, False), (True, True)):
            H, W, _ = img.shape
            src = img[:, :, ::-1] if bgr else img
            src = np.ascontiguousarray(src[::-1] if bottom_up else src)
            row0, rows = ops.host_zero_row_span(src, H, W * 3)
            nz = np.flatnonzero(src.reshape(H, -1).any(axis=1))
            if len(nz):
                assert row0 == nz[0] // 4 * 4 and row0 + rows == min(H, -(-(nz[-1] + 1) // 4) * 4), name
            else:
                assert rows == 0
            full = torch.from_numpy(src).to(cuda_device).view(-1)
            a, fla = ops.preprocess_4x(full, H, W, bgr=bgr, bottom_up=bottom_up)
            part = torch.from_numpy(src[row0:row0 + rows].copy()).to(cuda_device).view(-1)
            b, flb = ops.preprocess_4x(part, H, W, bgr=bgr, bottom_up=bottom_up, span=(row0, rows))
            assert fla.tolist() == flb.tolist(), name
            n = (fla[1] - fla[0]).item() * (W // 4) * 3
            assert torch.equal(a[:n], b[:n]), name
    assert rows == 0       # the last case really was the empty span
    # the chunk entry (three launches for n scans, what the engine issues): whole scans, spans and an all-zero scan mixed
    H = W = 512
    imgs = [synth.raw_image_u8(seed=20 + i, size=H, top=t, bottom=b)[0] for i, (t, b) in enumerate(((57, 130), (4, 9), (200, 1)))]
    imgs.append(np.zeros((H, W, 3), np.uint8))
    srcs = [np.ascontiguousarray(im[::-1, :, ::-1]) for im in imgs]               # BMP order
    spans = [ops.host_zero_row_span(s_, H, W * 3) for s_ in srcs]
    spans[1] = (0, H)                                                             # one scan copied whole
    parts = [torch.from_numpy(s_[r0:r0 + n].copy()).to(cuda_device).view(-1) for s_, (r0, n) in zip(srcs, spans)]
    canvas, fl = ops.preprocess_4x_batch(parts, H, W, spans, bgr=True, bottom_up=True)
    for i, s_ in enumerate(srcs):
        a, fla = ops.preprocess_4x(torch.from_numpy(s_).to(cuda_device).view(-1), H, W, bgr=True, bottom_up=True)
        assert fl[i].tolist() == fla.tolist(), i
        n = (fla[1] - fla[0]).item() * (W // 4) * 3
        assert torch.equal(canvas[i].view(-1)[:n], a[:n]), i


def test_preprocess_general_ratio(cuda_device):
    """Any size -> target x target (models.py:194-198) against the f64 restatement, byte for byte: non-integer ratios both
    ways, upscaling of one axis, BGR / bottom-up sources, dark bands that trigger the trim, an input range that clips."""
    ops = _ops()
    for (H, W, target, seed, bgr, bottom_up) in [(300, 500, 128, 1, False, False), (517, 333, 200, 2, True, True),
                                                 (1000, 90, 96, 3, True, False), (130, 129, 128, 4, False, True)]:
        raw = synth.texture_u8(H, W, seed)
        raw[:H // 6] = 0
        raw[H - H // 9:] = 0
        raw = np.clip(raw, 0, 230).astype(np.uint8)
        exp, f, l = opre.preprocess_u8(raw, target)
        src = raw[..., ::-1] if bgr else raw
        src = src[::-1] if bottom_up else src
        t = torch.from_numpy(np.ascontiguousarray(src)).to(cuda_device)
        out, fl = ops.preprocess_general(t.view(-1), H, W, target, bgr=bgr, bottom_up=bottom_up)
        first, last = fl.tolist()
        got = out[:(last - first) * target * 3].view(last - first, target, 3).cpu().numpy()
        assert (first, last) == (f, l), (H, W, target)
        assert np.array_equal(got, exp), (H, W, target, int((got != exp).sum()))
    # at the exact 4x ratio the f64 path agrees with the integer kernel except on exact .5 ties
    raw = synth.texture_u8(512, 512, 7)
    t = torch.from_numpy(raw).to(cuda_device)
    a, fla = ops.preprocess_general(t.view(-1), 512, 512, 128)
    b, flb = ops.preprocess_4x(t.view(-1), 512, 512)
    assert fla.tolist() == flb.tolist()
    d = (a.cpu().numpy().astype(np.int16) - b.cpu().numpy().astype(np.int16))
    assert np.abs(d).max() <= 1 and (d != 0).mean() < 0.01


def test_trim_u8(cuda_device):
    ops = _ops()
    img = synth.texture_u8(96, 96, 4)
    img[:7] = 0
    img[90:] = 0
    img[30, :10] = 0
    exp, f, l = opre.preprocess_u8(img)
    out, fl = ops.trim_u8(torch.from_numpy(img).to(cuda_device))
    first, last = fl.tolist()
    assert (first, last) == (f, l)
    assert np.array_equal(out[:(last - first) * 96 * 3].view(-1, 96, 3).cpu().numpy(), exp)


# ------------------------------------------------------------------------------------------------- K2 convolutions
# (Cin, Cout, k, stride, dil): the 24 tensor-core conv configurations of FCN-ResNet50
CONV_SHAPES = [(64, 64, 1, 1, 1), (64, 64, 3, 1, 1), (64, 256, 1, 1, 1), (256, 64, 1, 1, 1), (256, 128, 1, 1, 1),
               (128, 128, 3, 2, 1), (128, 128, 3, 1, 1), (128, 512, 1, 1, 1), (256, 512, 1, 2, 1), (512, 128, 1, 1, 1),
               (512, 256, 1, 1, 1), (256, 256, 3, 1, 1), (256, 256, 3, 1, 2), (256, 1024, 1, 1, 1), (512, 1024, 1, 1, 1),
               (1024, 256, 1, 1, 1), (1024, 512, 1, 1, 1), (512, 512, 3, 1, 2), (512, 512, 3, 1, 4), (512, 2048, 1, 1, 1),
               (1024, 2048, 1, 1, 1), (2048, 512, 1, 1, 1), (2048, 512, 3, 1, 1)]


def _conv_case(dev, Cin, Cout, k, stride, dil, N, H, W, relu, use_res, impl, seed=0, dtype=torch.bfloat16):
    ops = _ops()
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(N, H, W, Cin, generator=g).to(dtype)
    w = (torch.randn(Cout, Cin, k, k, generator=g) / np.sqrt(Cin * k * k)).to(torch.float32)
    gamma = torch.rand(Cout, generator=g) + 0.5
    beta = torch.randn(Cout, generator=g) * 0.1
    mean = torch.randn(Cout, generator=g) * 0.1
    var = torch.rand(Cout, generator=g) + 0.5
    pad = dil if k == 3 else 0
    wp, bias = ops.fold_bn_pack(w.to(dev), (gamma.to(dev), beta.to(dev), mean.to(dev), var.to(dev)), dtype=dtype)
    Ho = (H + 2 * pad - dil * (k - 1) - 1) // stride + 1
    Wo = (W + 2 * pad - dil * (k - 1) - 1) // stride + 1
    res = torch.randn(N, Ho, Wo, Cout, generator=g).to(dtype) if use_res else None
    y = ops.conv_bf16(x.to(dev), wp, bias, stride=stride, pad=pad, dil=dil, relu=relu,
                      residual=res.to(dev) if use_res else None, impl=impl)
    torch.cuda.synchronize()
    # oracle: the same bf16-rounded operands in f32 on the CPU
    wq = wp.float().cpu().permute(0, 3, 1, 2).contiguous()
    ref = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), wq, bias.cpu(), stride=stride, padding=pad, dilation=dil)
    if use_res:
        ref = ref + res.float().permute(0, 3, 1, 2)
    if relu:
        ref = ref.relu()
    ref = ref.permute(0, 2, 3, 1)
    got = y.float().cpu()
    assert got.shape == ref.shape
    err = (got - ref).abs()
    # output rounding (bf16: 2^-8 rel, fp16: 2^-11 rel) + f32 accumulation-order slack
    tol = (2.0 ** -7 * ref.abs() + 2e-2) if dtype == torch.bfloat16 else (2.0 ** -10 * ref.abs() + 2e-3)
    bad = (err > tol)
    assert not bad.any(), 'max err %.4g at %s (ref %.4g) bad=%d' % (err.max(), np.unravel_index(err.argmax(), err.shape),
                                                                   ref.flatten()[err.argmax()], int(bad.sum()))
    # fold check: BN scale really went into the weights
    scale = gamma / torch.sqrt(var + 1e-5)
    assert torch.allclose(bias.cpu(), beta - mean * scale, atol=1e-6)
    return float(err.max())


@pytest.mark.parametrize('impl', IMPLS)
@pytest.mark.parametrize('shape', CONV_SHAPES)
def test_conv_all_shapes(cuda_device, shape, impl):
    Cin, Cout, k, stride, dil = shape
    # ragged spatial size (not a multiple of any tile), batch 2
    _conv_case(cuda_device, Cin, Cout, k, stride, dil, N=2, H=27, W=40, relu=True, use_res=(k == 1 and Cout >= 256), impl=impl)


@pytest.mark.parametrize('impl', IMPLS)
def test_conv_full_width_rows(cuda_device, impl):
    # the production geometry: 128-wide rows, trimmed height (77 rows), dilation 2 and 4
    _conv_case(cuda_device, 256, 256, 3, 1, 2, N=1, H=77, W=128, relu=True, use_res=False, impl=impl, seed=1)
    _conv_case(cuda_device, 512, 512, 3, 1, 4, N=1, H=16, W=128, relu=False, use_res=False, impl=impl, seed=2)
    _conv_case(cuda_device, 128, 128, 3, 2, 1, N=1, H=33, W=255, relu=True, use_res=False, impl=impl, seed=3)
    _conv_case(cuda_device, 256, 512, 1, 2, 1, N=2, H=31, W=256, relu=False, use_res=False, impl=impl, seed=4)


@pytest.mark.parametrize('impl', IMPLS)
def test_conv_fp16_storage(cuda_device, impl):
    # fp16 operands / outputs (precision='fp16'): same kernels, tighter tolerance
    for (Cin, Cout, k, stride, dil) in [(64, 256, 1, 1, 1), (128, 128, 3, 2, 1), (512, 512, 3, 1, 4), (2048, 512, 3, 1, 1)]:
        _conv_case(cuda_device, Cin, Cout, k, stride, dil, N=2, H=27, W=40, relu=True, use_res=(k == 1), impl=impl,
                   dtype=torch.float16)


def test_conv_tc_many_tiles(cuda_device):
    # more tiles than SMs so the persistent loop, the smem ring wrap and both TMEM buffers are all exercised
    _conv_case(cuda_device, 512, 1024, 1, 1, 1, N=3, H=64, W=128, relu=True, use_res=True, impl=1, seed=5)
    _conv_case(cuda_device, 64, 64, 3, 1, 1, N=2, H=96, W=256, relu=True, use_res=False, impl=1, seed=6)


@pytest.mark.parametrize('shape', [(64, 64, 256, 1), (128, 256, 512, 2), (256, 512, 1024, 1), (512, 1024, 2048, 1)])
def test_conv_dual_source(cuda_device, shape):
    """conv3 + downsample branch of a bottleneck's first block as ONE launch (two operand sources, one accumulator)
    against the two separate f32 convolutions on the same bf16 operands."""
    ops = _ops()
    C1, C2, Cout, stride2 = shape
    g = torch.Generator().manual_seed(C1 + stride2)
    N, H, W = 2, 27, 40
    H2, W2 = (H * 2 - 1, W * 2 - 1) if stride2 == 2 else (H, W)        # odd input size: the lattice has a ragged edge
    x = torch.randn(N, H, W, C1, generator=g).to(torch.bfloat16)
    x2 = torch.randn(N, H2, W2, C2, generator=g).to(torch.bfloat16)
    w = (torch.randn(Cout, 1, 1, C1, generator=g) / np.sqrt(C1)).to(torch.bfloat16)
    w2 = (torch.randn(Cout, 1, 1, C2, generator=g) / np.sqrt(C2)).to(torch.bfloat16)
    bias = torch.randn(Cout, generator=g)
    y = ops.conv_dual_bf16(x.to(cuda_device), w.to(cuda_device), x2.to(cuda_device), w2.to(cuda_device), bias.to(cuda_device),
                           stride2=stride2, relu=True)
    torch.cuda.synchronize()
    F = torch.nn.functional
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2)) \
        + F.conv2d(x2.float().permute(0, 3, 1, 2), w2.float().permute(0, 3, 1, 2), stride=stride2) + bias.view(1, -1, 1, 1)
    ref = ref.relu().permute(0, 2, 3, 1)
    err = (y.float().cpu() - ref).abs()
    tol = 2.0 ** -7 * ref.abs() + 2e-2
    assert not (err > tol).any(), 'max err %.4g bad=%d' % (err.max(), int((err > tol).sum()))


# ------------------------------------------------------------------------------------------------- weight gradient
def _wgrad_case(dev, Cin, Cout, k, stride, dil, N, H, W, impl, seed=0):
    """dW of one conv layer vs torch's own f32 conv weight gradient on the same bf16-rounded operands."""
    ops = _ops()
    g = torch.Generator().manual_seed(seed)
    pad = dil if k == 3 else 0
    Ho = (H + 2 * pad - dil * (k - 1) - 1) // stride + 1
    Wo = (W + 2 * pad - dil * (k - 1) - 1) // stride + 1
    x = torch.randn(N, H, W, Cin, generator=g).to(torch.bfloat16)
    dz = torch.randn(N, Ho, Wo, Cout, generator=g).to(torch.bfloat16)
    got = ops.conv_wgrad_bf16(dz.to(dev), x.to(dev), k, k, stride=stride, pad=pad, dil=dil, impl=impl)
    torch.cuda.synchronize()
    ref = torch.nn.grad.conv2d_weight(x.float().permute(0, 3, 1, 2), (Cout, Cin, k, k), dz.float().permute(0, 3, 1, 2),
                                      stride=stride, padding=pad, dilation=dil)          # OIHW
    ref = ref.permute(0, 2, 3, 1)                                                          # [Cout][kh][kw][Cin]
    err = (got.cpu() - ref).abs()
    # exact products (bf16 x bf16 fits f32), f32 accumulation over N*Ho*Wo terms of unit variance in a different order
    tol = 1e-4 * np.sqrt(N * Ho * Wo) + 1e-5 * ref.abs()
    assert not (err > tol).any(), 'max err %.4g (ref rms %.4g) bad=%d of %d' % (err.max(), ref.pow(2).mean().sqrt(),
                                                                             int((err > tol).sum()), err.numel())


@pytest.mark.parametrize('impl', [1, 2])
@pytest.mark.parametrize('shape', [(64, 64, 1, 1, 1), (64, 64, 3, 1, 1), (256, 64, 1, 1, 1), (64, 256, 1, 1, 1), (128, 128, 3, 2, 1),
                                   (256, 512, 1, 2, 1), (128, 512, 1, 1, 1), (256, 256, 3, 1, 2), (512, 512, 3, 1, 4),
                                   (1024, 256, 1, 1, 1), (2048, 512, 3, 1, 1)])
def test_conv_wgrad(cuda_device, shape, impl):
    Cin, Cout, k, stride, dil = shape
    _wgrad_case(cuda_device, Cin, Cout, k, stride, dil, N=2, H=27, W=40, impl=impl)


def test_conv_wgrad_tc_split_k(cuda_device):
    # production-like geometry: 128-wide rows, many pixel tiles per CTA and several splits; accumulation into a
    # non-zero buffer
    ops = _ops()
    _wgrad_case(cuda_device, 256, 256, 3, 1, 2, N=2, H=64, W=128, impl=1, seed=3)
    _wgrad_case(cuda_device, 64, 64, 3, 1, 1, N=1, H=200, W=256, impl=1, seed=4)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(1, 16, 64, 128, generator=g).to(torch.bfloat16).to(cuda_device)
    dz = torch.randn(1, 16, 64, 64, generator=g).to(torch.bfloat16).to(cuda_device)
    base = torch.full((64, 1, 1, 128), 2.0, device=cuda_device)
    a = ops.conv_wgrad_bf16(dz, x, 1, 1, impl=1)
    b = ops.conv_wgrad_bf16(dz, x, 1, 1, impl=1, out=base.clone())
    assert torch.allclose(b - 2.0, a, atol=1e-3)


# ------------------------------------------------------------------------------------------------- stem / pool / head
def test_stem_maxpool_head(cuda_device):
    ops = _ops()
    dev = cuda_device
    g = torch.Generator().manual_seed(0)
    img = synth.texture_u8(45, 70, 3)
    w = torch.randn(64, 3, 7, 7, generator=g) * 0.1
    bn = [torch.rand(64, generator=g) + 0.5, torch.randn(64, generator=g) * 0.1, torch.randn(64, generator=g) * 0.1,
          torch.rand(64, generator=g) + 0.5]
    x = omodel.normalise_u8(img)
    scale = bn[0] / torch.sqrt(bn[3] + 1e-5)
    ref = torch.nn.functional.conv2d(x, w * scale.view(-1, 1, 1, 1), bn[1] - bn[2] * scale, stride=2, padding=3).relu()
    # f32 fold of the stem weights happens inside nbc_plan_create; here emulate it with torch on the host
    wf = (w * scale.view(-1, 1, 1, 1)).permute(0, 2, 3, 1).contiguous().to(dev)
    bf = (bn[1] - bn[2] * scale).to(dev)
    y = ops.stem_u8(torch.from_numpy(img).unsqueeze(0).to(dev), omodel.DEFAULT_MEAN, omodel.DEFAULT_STD, wf, bf)
    got = y.float().cpu().permute(0, 3, 1, 2)
    assert (got - ref).abs().max() < 2.0 ** -8 * ref.abs().max() + 1e-3
    y2 = ops.stem_f32(x.to(dev), wf, bf)
    assert torch.equal(y2, y)
    # tensor-core stem (bf16 operands): same conv within bf16 operand rounding, u8 and f32 inputs bit-identical,
    # odd sizes and a batch
    for inp in (torch.from_numpy(img).unsqueeze(0).to(dev), x.to(dev)):
        yt = ops.stem_tc(inp, omodel.DEFAULT_MEAN, omodel.DEFAULT_STD, wf, bf)
        gt = yt.float().cpu().permute(0, 3, 1, 2)
        assert gt.shape == ref.shape
        assert (gt - ref).abs().max() < 0.02 * ref.abs().max() + 1e-2, (gt - ref).abs().max()
    yh = ops.stem_tc(x.to(dev), omodel.DEFAULT_MEAN, omodel.DEFAULT_STD, wf, bf, dtype=torch.float16)
    assert yh.dtype == torch.float16
    assert (yh.float().cpu().permute(0, 3, 1, 2) - ref).abs().max() < 0.003 * ref.abs().max() + 2e-3
    ph = ops.maxpool3x3s2(yh)
    assert torch.equal(ph.float().cpu().permute(0, 3, 1, 2), torch.nn.functional.max_pool2d(yh.float().cpu().permute(0, 3, 1, 2), 3, 2, 1))
    big = np.stack([synth.texture_u8(203, 517, 5), synth.texture_u8(203, 517, 6)])
    yb = ops.stem_tc(torch.from_numpy(big).to(dev), omodel.DEFAULT_MEAN, omodel.DEFAULT_STD, wf, bf).float().cpu()
    for i in range(2):
        r = torch.nn.functional.conv2d(omodel.normalise_u8(big[i]), w * scale.view(-1, 1, 1, 1), bn[1] - bn[2] * scale,
                                       stride=2, padding=3).relu()
        assert (yb[i].permute(2, 0, 1) - r[0]).abs().max() < 0.02 * r.abs().max() + 1e-2
    p = ops.maxpool3x3s2(y)
    refp = torch.nn.functional.max_pool2d(y.float().cpu().permute(0, 3, 1, 2), 3, 2, 1)
    assert torch.equal(p.float().cpu().permute(0, 3, 1, 2), refp)
    feats = torch.randn(2, 9, 13, 512, generator=g).to(torch.bfloat16)
    cw = torch.randn(3, 512, generator=g) * 0.05
    cb = torch.randn(3, generator=g)
    lg = ops.head_1x1(feats.to(dev), cw.to(dev), cb.to(dev)).cpu()
    refl = torch.einsum('nhwc,kc->nkhw', feats.float(), cw) + cb.view(1, 3, 1, 1)
    assert (lg - refl).abs().max() < 1e-4


def test_stem_halo_kernel(cuda_device):
    """conv_tc_stem_kernel (overlapping no-swizzle windows out of a 7-row halo) at the production tile geometry (128 x 1
    output tiles): full-width rows, a width that leaves a partial last tile, a batch, fp16 storage."""
    ops = _ops()
    g = torch.Generator().manual_seed(1)
    w = torch.randn(64, 3, 7, 7, generator=g) * 0.1
    bn = [torch.rand(64, generator=g) + 0.5, torch.randn(64, generator=g) * 0.1, torch.randn(64, generator=g) * 0.1,
          torch.rand(64, generator=g) + 0.5]
    scale = bn[0] / torch.sqrt(bn[3] + 1e-5)
    wf = (w * scale.view(-1, 1, 1, 1)).permute(0, 2, 3, 1).contiguous().to(cuda_device)
    bf = (bn[1] - bn[2] * scale).to(cuda_device)
    for (N, H, W, dtype, tol) in [(1, 61, 1024, torch.bfloat16, 0.02), (2, 61, 900, torch.bfloat16, 0.02),
                                  (3, 333, 1024, torch.bfloat16, 0.02), (1, 61, 1024, torch.float16, 0.003)]:
        imgs = np.stack([synth.texture_u8(H, W, 10 + i) for i in range(N)])
        y = ops.stem_tc(torch.from_numpy(imgs).to(cuda_device), omodel.DEFAULT_MEAN, omodel.DEFAULT_STD, wf, bf, dtype=dtype).float().cpu()
        for i in range(N):
            r = torch.nn.functional.conv2d(omodel.normalise_u8(imgs[i]), w * scale.view(-1, 1, 1, 1), bn[1] - bn[2] * scale,
                                           stride=2, padding=3).relu()
            assert y[i].shape == r[0].permute(1, 2, 0).shape
            assert (y[i].permute(2, 0, 1) - r[0]).abs().max() < tol * r.abs().max() + 1e-2, (N, H, W, i)


# ------------------------------------------------------------------------------------------------- K3 upsample + argmax
@pytest.mark.parametrize('shape', [(2, 13, 16, 100, 128), (1, 128, 128, 1024, 1024), (1, 77, 128, 611, 1024)])
def test_upsample_argmax_bit_exact(cuda_device, shape):
    ops = _ops()
    N, h, w, H, W = shape
    rng = np.random.default_rng(h)
    low = rng.standard_normal((N, 3, h, w)).astype(np.float32)
    low[0, :, 0, :4] = 0.25            # exact ties -> lowest index must win
    exp_up = omodel.upsample_bicubic_restated(low, (H, W))
    t = torch.from_numpy(low).to(cuda_device)
    up = ops.upsample_bicubic(t, (H, W)).cpu().numpy()
    assert np.array_equal(up, exp_up), 'upsample differs from the restated f32 specification'
    mask = ops.upsample_argmax(t, (H, W)).cpu().numpy()
    assert np.array_equal(mask, omodel.argmax_lowest(exp_up))
    # against torch's own kernel: float round-off only
    ref = torch.nn.functional.interpolate(torch.from_numpy(low), size=(H, W), mode='bicubic', align_corners=False)
    assert np.abs(up - ref.numpy()).max() < 5e-5
    assert (mask != ref.argmax(1).numpy()).mean() < 1e-4


def test_upsample_golden(cuda_device, golden_dir):
    ops = _ops()
    g = np.load(os.path.join(golden_dir, 'upsample_argmax.npz'))
    mask = ops.upsample_argmax(torch.from_numpy(g['lowres']).to(cuda_device), g['up'].shape[-2:]).cpu().numpy()
    assert (mask != g['mask']).mean() < 1e-4


# ------------------------------------------------------------------------------------------------- K5 region removal
def test_ccl_golden(cuda_device, golden_dir):
    ops = _ops()
    g = np.load(os.path.join(golden_dir, 'ccl_small.npz'))
    m = torch.from_numpy(g['mask']).unsqueeze(0).to(cuda_device)
    out, counts = ops.remove_small_zones_u8(m)
    assert np.array_equal(out[0].cpu().numpy(), g['out'])
    assert counts[0].tolist() == np.bincount(g['out'].ravel(), minlength=3).tolist()


# (rows wider than 1024 pixels take the per-pixel kernels, everything else the run-based ones: both are covered)
@pytest.mark.parametrize('shape,thr', [((3, 200, 333), 150), ((1, 1024, 1024), 150), ((2, 611, 1024), 150), ((4, 64, 64), 9),
                                       ((2, 90, 1100), 150), ((1, 37, 1024), 30)])
def test_ccl_random_vs_oracle(cuda_device, shape, thr):
    ops = _ops()
    N, H, W = shape
    masks = np.stack([synth.class_mask(H, W, 10 + i) for i in range(N)])
    rng = np.random.default_rng(1)
    noise = rng.random(masks.shape) < 0.02               # salt noise: thousands of tiny components and holes
    masks = np.where(noise, (masks + rng.integers(1, 3, masks.shape)) % 3, masks).astype(np.uint8)
    exp = opost.remove_small_zones(masks, thr)
    t = torch.from_numpy(masks).to(cuda_device)
    out, counts = ops.remove_small_zones_u8(t, thr)
    got = out.cpu().numpy()
    assert np.array_equal(got, exp)
    for i in range(N):
        assert counts[i].tolist() == np.bincount(exp[i].ravel(), minlength=3).tolist()
    # exclude_nodes: 2 -> 1 after the removal (models.py:273-276)
    t2 = torch.from_numpy(masks).to(cuda_device)
    out2, counts2 = ops.remove_small_zones_u8(t2, thr, exclude_nodes=True)
    assert np.array_equal(out2.cpu().numpy(), opost.exclude_nodes(exp))
    assert int(counts2[:, 2].sum()) == 0


def test_ccl_edge_cases(cuda_device):
    ops = _ops()
    dev = cuda_device
    for m in (np.zeros((1, 50, 70), np.uint8), np.full((1, 50, 70), 2, np.uint8)):
        out, counts = ops.remove_small_zones_u8(torch.from_numpy(m.copy()).to(dev))
        assert np.array_equal(out.cpu().numpy(), opost.remove_small_zones(m))
    # spiral / snake: one long thin 8-connected component (deep union-find chains)
    m = np.zeros((1, 128, 128), np.uint8)
    for r in range(0, 128, 4):
        m[0, r, :] = 1
        m[0, r:r + 4, 127 if (r // 4) % 2 == 0 else 0] = 1
    m[0, 64, 50:60] = 0
    out, _ = ops.remove_small_zones_u8(torch.from_numpy(m.copy()).to(dev))
    assert np.array_equal(out.cpu().numpy(), opost.remove_small_zones(m))
    # drop-in wrapper: int64 tensor, in place, returns the same object (utils.py:135-148)
    from neuralbarkcalculator_b200 import utils
    mm = synth.class_mask(90, 100, 3).astype(np.int64)
    t = torch.from_numpy(mm.copy()).unsqueeze(0).to(dev)
    r = utils.remove_small_zones(t)
    assert r is t and np.array_equal(t[0].cpu().numpy(), opost.remove_small_zones_2d(mm))


def test_k3_k5_stay_inside_their_buffers(cuda_device):
    """Bounds check of our own (compute-sanitizer is not available on the GPU pool): K3 and K5 work on views in the middle
    of guard-filled buffers -- mask canvas, logits, workspace, counts -- sized EXACTLY as the ABI says; the guards must be
    untouched afterwards and the dead rows of the ragged canvas (below each image's height) unwritten."""
    ops = _ops()
    from neuralbarkcalculator_b200 import _lib
    lib = _lib.load()
    dev = cuda_device
    G = 4096
    for (N, Hc, W, heights) in [(3, 96, 1024, [96, 41, 7]), (2, 50, 300, [50, 1]), (1, 33, 96, [20])]:
        rng = np.random.default_rng(N)
        hl = [(((h - 1) // 2 + 1 - 1) // 2 + 1 - 1) // 2 + 1 for h in heights]
        hc, w = (((Hc - 1) // 2 + 1 - 1) // 2 + 1 - 1) // 2 + 1, (((W - 1) // 2 + 1 - 1) // 2 + 1 - 1) // 2 + 1

        def guarded(nbytes, dtype):
            raw = torch.full((nbytes + 2 * G,), 0xA5, dtype=torch.uint8, device=dev)
            return raw, raw[G:G + nbytes].view(dtype)

        raw_m, mask = guarded(N * Hc * W, torch.uint8)
        mask = mask.view(N, Hc, W)
        raw_l, logits = guarded(N * 3 * hc * w * 4, torch.float32)
        logits = logits.view(N, 3, hc, w)
        logits.copy_(torch.from_numpy(rng.standard_normal((N, 3, hc, w)).astype(np.float32)))
        need = lib.nbc_ccl_workspace_bytes(N, Hc, W)
        raw_w, ws = guarded(need, torch.uint8)
        raw_c, counts = guarded(N * 3 * 4, torch.int32)
        counts = counts.view(N, 3)
        hd = torch.tensor(heights, dtype=torch.int32, device=dev)
        ops.upsample_argmax_ragged(logits, hd, (Hc, W), out=mask)
        ops.remove_small_zones_ragged(mask, hd, 150, True, workspace=ws, counts=counts)
        torch.cuda.synchronize()
        for name, raw, n in (('mask', raw_m, N * Hc * W), ('logits', raw_l, N * 3 * hc * w * 4), ('workspace', raw_w, need),
                             ('counts', raw_c, N * 12)):
            assert bool((raw[:G] == 0xA5).all()) and bool((raw[G + n:] == 0xA5).all()), '%s guard overwritten (%s)' % (name, (N, Hc, W))
        for i, h in enumerate(heights):
            assert bool((mask[i, h:] == 0xA5).all()), 'dead rows of image %d written' % i
            assert int(counts[i].sum()) == h * W
            up = omodel.upsample_bicubic_restated(logits[i:i + 1, :, :hl[i]].cpu().numpy(), (h, W))
            exp = opost.exclude_nodes(opost.remove_small_zones_2d(omodel.argmax_lowest(up)[0]))
            assert np.array_equal(mask[i, :h].cpu().numpy(), exp), (N, Hc, W, i)


def test_ccl_idempotent_full_size(cuda_device):
    ops = _ops()
    m = torch.from_numpy(np.stack([synth.class_mask(1024, 1024, 77 + i) for i in range(4)])).to(cuda_device)
    a, ca = ops.remove_small_zones_u8(m.clone())
    b, cb = ops.remove_small_zones_u8(a.clone())
    assert torch.equal(a, b) and torch.equal(ca, cb)
    assert int(ca.sum()) == 4 * 1024 * 1024


# ------------------------------------------------------------------------------------------------- K4 weighted CE
def test_wce_golden_and_autograd(cuda_device, golden_dir):
    ops = _ops()
    from neuralbarkcalculator_b200 import utils
    g = np.load(os.path.join(golden_dir, 'wce_small.npz'))
    dev = cuda_device
    logits = torch.from_numpy(g['logits']).to(dev)
    w = torch.from_numpy(g['weights']).to(dev)
    for tgt in (torch.from_numpy(g['target']).to(dev), torch.from_numpy(g['target']).long().to(dev)):
        loss, grad = ops.wce_fwd_bwd(logits, tgt, w)
        assert abs(float(loss) - float(g['loss'])) < 1e-5 * max(1.0, abs(float(g['loss'])))       # f32 tolerance
        assert np.abs(grad.cpu().numpy() - g['grad']).max() < 1e-7 + 1e-5 * np.abs(g['grad']).max()
    crit = utils.CustomWeightedCrossEntropy(utils.get_pos_weight())
    p = logits.clone().requires_grad_(True)
    out = crit(p, torch.from_numpy(g['target']).long().to(dev))
    (out * 2).backward()
    assert out.dim() == 0 and np.abs(p.grad.cpu().numpy() - 2 * g['grad']).max() < 1e-6


def test_wce_large_vs_oracle(cuda_device):
    ops = _ops()
    gen = torch.Generator().manual_seed(4)
    logits = torch.randn(2, 3, 512, 768, generator=gen) * 3
    target = torch.from_numpy(np.stack([synth.class_mask(512, 768, s) for s in (5, 6)])).long()
    w = torch.tensor(olosses.DEFAULT_WEIGHTS)
    ref_loss, ref_grad = olosses.custom_weighted_cross_entropy_with_grad(logits, target, w)
    loss, grad = ops.wce_fwd_bwd(logits.to(cuda_device), target.to(torch.uint8).to(cuda_device), w.to(cuda_device))
    assert abs(float(loss) - float(ref_loss)) < 2e-5 * abs(float(ref_loss))
    assert (grad.cpu() - ref_grad).abs().max() < 1e-9 + 2e-5 * ref_grad.abs().max()


# ------------------------------------------------------------------------------------------------- N1 Lovasz / N2 metrics
def _lovasz_check(dev, logits, target, ref_loss, ref_grad, target_dtype=torch.uint8):
    ops = _ops()
    loss, grad = ops.lovasz_softmax_fwd_bwd(logits.to(dev), target.to(target_dtype).to(dev))
    torch.cuda.synchronize()
    assert abs(float(loss) - float(ref_loss)) < 2e-6 + 1e-5 * abs(float(ref_loss)), (float(loss), float(ref_loss))
    gerr = (grad.cpu() - ref_grad).abs().max()
    # the gradient at a pixel is its Lovasz weight at its rank (a difference of two Jaccard values, computed with the
    # reference's own f32 expression) through the softmax Jacobian: agreement to f32 round-off
    assert gerr < 1e-8 + 2e-5 * float(ref_grad.abs().max()), (float(gerr), float(ref_grad.abs().max()))


def test_lovasz_golden(cuda_device, golden_dir):
    g = np.load(os.path.join(golden_dir, 'lovasz_small.npz'))
    logits, target = torch.from_numpy(g['logits']), torch.from_numpy(g['target'])
    _lovasz_check(cuda_device, logits, target, g['loss'], torch.from_numpy(g['grad']))
    _lovasz_check(cuda_device, logits, target, g['loss'], torch.from_numpy(g['grad']), torch.int64)
    two = target.clone()
    two[two == 2] = 1       # class 2 absent: classes='present' drops it from the mean
    _lovasz_check(cuda_device, logits, two, g['loss_two_classes'], torch.from_numpy(g['grad_two_classes']))


def test_lovasz_large_vs_oracle(cuda_device):
    from oracle import lovasz as olovasz
    g = torch.Generator().manual_seed(21)
    logits = torch.randn(3, 3, 200, 333, generator=g) * 2.5
    target = torch.from_numpy(np.stack([_synth().class_mask(200, 333, s) for s in (1, 2, 3)])).long()
    loss, grad = olovasz.lovasz_softmax_with_grad(logits, target)
    _lovasz_check(cuda_device, logits, target, loss, grad)
    # autograd wrapper + MixedLoss mirror (utils.py:185-192)
    from neuralbarkcalculator_b200 import utils as nutils
    from oracle import losses as olosses
    w = torch.tensor(olosses.DEFAULT_WEIGHTS)
    p = logits.to(cuda_device).requires_grad_(True)
    out = nutils.MixedLoss(w)(p, target.to(cuda_device))
    out.backward()
    ref = logits.clone().requires_grad_(True)
    ref_out = olovasz.mixed_loss(ref, target, w)
    ref_out.backward()
    assert abs(float(out) - float(ref_out)) < 1e-5 * abs(float(ref_out)) + 1e-6
    assert (p.grad.cpu() - ref.grad).abs().max() < 1e-8 + 2e-5 * float(ref.grad.abs().max())


def test_metrics_vs_oracle(cuda_device, golden_dir):
    from oracle import lovasz as olovasz
    from neuralbarkcalculator_b200 import lovasz_losses as nl, utils as nutils
    ops = _ops()
    g = np.load(os.path.join(golden_dir, 'lovasz_small.npz'))
    logits, target = torch.from_numpy(g['logits']), torch.from_numpy(g['target']).long()
    cm = ops.confusion_matrix(ops.argmax3_u8(logits.to(cuda_device)), target.to(cuda_device))
    assert np.array_equal(cm.cpu().numpy(), g['confusion'])
    assert np.array_equal(nl.iou(logits.to(cuda_device), target.to(cuda_device)), g['iou'])
    assert nl.miou(logits.to(cuda_device), target.to(cuda_device)) == np.mean(g['iou'])
    # ties resolve to the lowest index, like torch.argmax
    tie = torch.zeros(1, 3, 4, 4)
    tie[0, 1, 0, 0] = tie[0, 2, 0, 0] = 1.0
    am = ops.argmax3_u8(tie.to(cuda_device)).cpu()
    assert am[0, 0, 0] == 1 and int(am.sum()) == 1
    # PixelWiseF1 on a larger batch (argmax -> per-image region removal -> F1), every class_to_watch flavour
    gen = torch.Generator().manual_seed(4)
    big = torch.nn.functional.interpolate(torch.randn(2, 3, 20, 30, generator=gen), size=(160, 240), mode='bilinear') * 3
    lab = torch.from_numpy(np.stack([_synth().class_mask(160, 240, s) for s in (7, 8)])).long()
    for watch in (None, 'loss', 1, 'all'):
        ours = nutils.PixelWiseF1(watch)(big.to(cuda_device), lab.to(cuda_device))
        ref = olovasz.pixelwise_f1(big, lab, watch)
        assert np.allclose(ours, ref, rtol=0, atol=1e-12), (watch, ours, ref)


# ------------------------------------------------------------------------------------------------- N4 augmentation
def test_augment_batch_vs_oracle(cuda_device):
    """The fused crop / flip / reflect-pad / colour-jitter gather == the reference's torchvision-on-PIL chain, bit for bit."""
    from oracle import augment as oaug
    from neuralbarkcalculator_b200 import augment as naug
    synth = _synth()
    rng = np.random.default_rng(5)
    for (Hs, Ws, target, crop) in ((1024, 1024, (1024, 1024), 512), (1000, 1024, (1024, 1024), 384), (96, 80, (100, 84), 64)):
        images = [synth.texture_u8(Hs, Ws, 40 + i) for i in range(3)]
        duals = [(synth.class_mask(Hs, Ws, 50 + i).astype(np.float32) * 127.5).astype(np.uint8) for i in range(3)]
        params = naug.draw_params(rng, 6, 3, crop, target)
        params['brightness'][0], params['saturation'][1] = 0.0, 0.0          # disabled ops
        params['brightness'][2], params['saturation'][2] = 1.0, 1.0          # factor exactly 1 still goes through PIL
        plist = [{k: (float(p[k]) if k in ('brightness', 'saturation') else int(p[k])) for k in naug.PARAM_DTYPE.names} for p in params]
        ref_i, ref_c = oaug.augment(images, duals, plist, crop, target)
        out, cls = naug.augment_batch(torch.from_numpy(np.stack(images)).to(cuda_device), torch.from_numpy(np.stack(duals)).to(cuda_device),
                                      params, crop, target)
        assert np.array_equal(out.cpu().numpy(), ref_i), (Hs, Ws)
        assert np.array_equal(cls.cpu().numpy(), ref_c), (Hs, Ws)
    with pytest.raises(RuntimeError):        # an odd difference would need PIL's resize: rejected, never approximated
        naug.augment_batch(torch.zeros(1, 1001, 1024, 3, dtype=torch.uint8, device=cuda_device), None, naug.draw_params(rng, 1, 1, 256), 256)
    # end to end: an augmented batch feeds the native training step
    from neuralbarkcalculator_b200.train import Trainer
    from oracle import model as omodel
    sd = omodel.synthetic_state_dict(seed=0, head=None)
    images = torch.from_numpy(np.stack([synth.texture_u8(128, 128, 60 + i) for i in range(2)])).to(cuda_device)
    duals = torch.from_numpy(np.stack([(synth.class_mask(128, 128, 70 + i).astype(np.float32) * 127.5).astype(np.uint8) for i in range(2)])).to(cuda_device)
    x, t = naug.augment_batch(images, duals, naug.draw_params(rng, 2, 2, 64, (128, 128)), 64, (128, 128))
    tr = Trainer(sd, 2, 64, 64, device='cuda:0', dropout=0.0)
    assert np.isfinite(float(tr.step(x, t)))
