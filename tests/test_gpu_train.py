"""GPU parity of the native training step (row a14) against the torch CPU oracle (``pytest -m gpu``)."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from oracle import losses as olosses
from oracle import model as omodel
from oracle import synth
from oracle import train as otrain

pytestmark = pytest.mark.gpu


def _batch(N, H, W, seed=0):
    imgs = np.stack([synth.texture_u8(H, W, seed + i) for i in range(N)])
    tgt = np.stack([synth.class_mask(H, W, seed + 100 + i) for i in range(N)])
    x = torch.stack([omodel.normalise_u8(im)[0] for im in imgs])
    return imgs, tgt, x


def _cos(a, b):
    a, b = a.flatten().double(), b.flatten().double()
    return float((a @ b) / (a.norm() * b.norm() + 1e-30))


def test_adam_matches_torch(cuda_device):
    from neuralbarkcalculator_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(0)
    n = 100003
    p0 = torch.randn(n, generator=g)
    ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=5e-4, weight_decay=2e-3)
    p = p0.clone().to(cuda_device)
    m = torch.zeros_like(p)
    v = torch.zeros_like(p)
    for step in range(1, 4):
        grad = torch.randn(n, generator=g) * (10.0 ** float(torch.randint(-4, 2, (1,), generator=g)))
        ref.grad = grad.clone()
        opt.step()
        gd = (grad * 2).to(cuda_device)       # grad_scale 0.5 below undoes the factor (the 1/world of data parallel)
        _lib.check(lib.nbc_train_adam(C.c_void_p(p.data_ptr()), C.c_void_p(gd.data_ptr()), C.c_void_p(m.data_ptr()),
                                      C.c_void_p(v.data_ptr()), n, C.c_float(5e-4), C.c_float(0.9), C.c_float(0.999), C.c_float(1e-8),
                                      C.c_float(2e-3), step, C.c_float(0.5),
                                      C.c_void_p(torch.cuda.current_stream().cuda_stream)), 'nbc_train_adam')
        assert (p.cpu() - ref.detach()).abs().max() < 2e-6


@pytest.mark.parametrize('weights', [(1.0, 1.0, 1.0), tuple(olosses.DEFAULT_WEIGHTS)])
def test_train_step_matches_oracle(cuda_device, synthetic_sd, weights):
    """One step, batch 2 at 64x96, dropout 0: loss, per-layer gradients, BN running statistics, updated weights.

    With the reference's class weights (0.4 / 2.0 / 93.2) the loss is DISCONTINUOUS in the logits -- the weight is
    picked by argmax(pred) -- so the ~1 % bf16 logit error flips a few pixels between weight 2 and 93 and the
    gradients differ from the f32 oracle by much more than rounding.  The machinery is therefore checked strictly
    with uniform weights (continuous loss) and loosely with the reference's weights."""
    from neuralbarkcalculator_b200.train import Trainer
    N, H, W = 2, 64, 96
    strict = weights[0] == weights[2]
    imgs, tgt, x = _batch(N, H, W)
    wt = torch.tensor(weights)
    ref = otrain.train_step(synthetic_sd, x, torch.from_numpy(tgt).long(), weights=wt, dropout=0.0)
    tr = Trainer(synthetic_sd, N, H, W, device='cuda:0', dropout=0.0, class_weights=wt)
    loss = tr.forward_backward(torch.from_numpy(imgs).to(cuda_device), torch.from_numpy(tgt).to(cuda_device), seed=1)
    loss = float(loss)
    print('\nloss ours %.5f oracle %.5f' % (loss, ref['loss']))
    assert abs(loss - ref['loss']) < (0.01 if strict else 0.05) * abs(ref['loss'])
    grads = tr.gradients()
    rows = []
    for k, gref in ref['grads'].items():
        g = grads[k].cpu()
        assert g.shape == gref.shape, k
        rows.append((k, _cos(g, gref), float(g.norm() / (gref.norm() + 1e-30))))
    for k, c, r in rows[::6] + rows[-5:]:
        print('%-45s cos %.4f  norm ratio %.3f' % (k, c, r))
    conv = [(k, c, r) for k, c, r in rows if k.endswith('.weight') and ('conv' in k or 'downsample.0' in k or k in ('classifier.0.weight', 'classifier.4.weight'))]
    # bf16 activations / gradients: direction and size of every conv weight gradient must match the f32 oracle
    print('conv weight gradients: min cos %.4f median %.4f' % (min(c for _, c, _ in conv), np.median([c for _, c, _ in conv])))
    assert all(0.75 < r < 1.33 for _, _, r in conv), [t for t in conv if not 0.75 < t[2] < 1.33]
    if not strict:
        assert conv[-1][1] > 0.99 and conv[-2][1] > 0.9      # classifier gradients
        return
    assert min(c for _, c, _ in conv) > 0.90, min(conv, key=lambda t: t[1])
    assert np.median([c for _, c, _ in conv]) > 0.97
    bn = [(k, c, r) for k, c, r in rows if (k, c, r) not in conv]
    assert np.median([c for _, c, _ in bn]) > 0.95
    # running statistics after the step (momentum 0.1)
    sd = tr.state_dict()
    for k in ('backbone.bn1.running_mean', 'backbone.layer3.2.bn2.running_var', 'classifier.1.running_mean'):
        a, b = sd[k].cpu(), ref['state_dict'][k]
        assert (a - b).abs().max() < 0.02 * b.abs().max() + 2e-3, k
    # optimiser: parameters move by ~lr in the oracle's direction wherever the gradient is not tiny
    tr.optimizer_step()
    new = tr.state_dict()
    agree, total = 0, 0
    for k, gref in ref['grads'].items():
        big = gref.abs() > 0.05 * gref.abs().max()
        d_ours = (new[k].cpu() - synthetic_sd[k])[big]
        d_ref = (ref['state_dict'][k] - synthetic_sd[k])[big]
        agree += int((torch.sign(d_ours) == torch.sign(d_ref)).sum())
        total += int(big.sum())
        assert d_ours.abs().max() < 1.2e-3      # |update| <= lr (+ weight decay) on the first step
    print('update sign agreement on significant gradients: %.4f (%d)' % (agree / total, total))
    assert agree / total > 0.97


def test_train_dropout_and_determinism(cuda_device, synthetic_sd):
    from neuralbarkcalculator_b200.train import Trainer
    N, H, W = 2, 64, 64
    imgs, tgt, _ = _batch(N, H, W, seed=5)
    tr = Trainer(synthetic_sd, N, H, W, device='cuda:0', dropout=0.8)
    xi, ti = torch.from_numpy(imgs).to(cuda_device), torch.from_numpy(tgt).to(cuda_device)
    l1 = float(tr.forward_backward(xi, ti, seed=7, update_stats=False))
    g1 = tr.grads.clone()
    l2 = float(tr.forward_backward(xi, ti, seed=7, update_stats=False))
    l3 = float(tr.forward_backward(xi, ti, seed=8, update_stats=False))
    assert np.isfinite(l1) and l1 == l2 and l1 != l3
    assert torch.isfinite(tr.grads).all() and float((g1 != 0).float().mean()) > 0.5
    # a few full steps reduce the loss on a fixed batch
    tr2 = Trainer(synthetic_sd, N, H, W, device='cuda:0', dropout=0.0)
    first = float(tr2.step(xi, ti))
    for _ in range(7):
        last = float(tr2.step(xi, ti))
    print('\nloss over 8 steps on a fixed batch: %.4f -> %.4f' % (first, last))
    assert last < first
