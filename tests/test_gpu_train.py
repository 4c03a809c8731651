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

    Two things make a naive comparison with the f32 oracle meaningless, and neither is a property of the kernels:
    (1) a random-init train-mode BatchNorm network amplifies perturbations (~1.1x per bottleneck): ANY bf16 forward is
    ~10 % away from the f32 one at layer4 (oracle/train.py reproduces this by rounding at the same points), so the
    reference here is the same-precision oracle (bf16_sim) and the f32 gap is printed beside it;
    (2) with the reference's class weights (0.4 / 2.0 / 93.2) the loss is DISCONTINUOUS in the logits -- the weight
    is picked by argmax(pred) -- so the machinery is checked strictly with uniform weights and loosely with those."""
    from neuralbarkcalculator_b200.train import Trainer
    N, H, W = 2, 64, 96
    strict = weights[0] == weights[2]
    imgs, tgt, x = _batch(N, H, W)
    wt = torch.tensor(weights)
    ref = otrain.train_step(synthetic_sd, x, torch.from_numpy(tgt).long(), weights=wt, dropout=0.0, bf16_sim=True)
    ref32 = otrain.train_step(synthetic_sd, x, torch.from_numpy(tgt).long(), weights=wt, dropout=0.0)
    tr = Trainer(synthetic_sd, N, H, W, device='cuda:0', dropout=0.0, class_weights=wt)
    loss = tr.forward_backward(torch.from_numpy(imgs).to(cuda_device), torch.from_numpy(tgt).to(cuda_device), seed=1)
    loss = float(loss)
    print('\nloss ours %.5f | same-precision oracle %.5f | f32 oracle %.5f' % (loss, ref['loss'], ref32['loss']))
    # the three losses are samples of the same perturbation sensitivity: ours must be as close to f32 as the
    # same-precision torch oracle is (within a factor), never a gross outlier
    gap = abs(ref['loss'] - ref32['loss'])
    assert abs(loss - ref32['loss']) < max(3.0 * gap, 0.02 * abs(ref32['loss'])) + (0.05 * abs(ref32['loss']) if not strict else 0)
    grads = tr.gradients()
    rows = []
    for k, g32 in ref32['grads'].items():
        g = grads[k].cpu()
        assert g.shape == g32.shape, k
        rows.append((k, _cos(g, g32), _cos(ref['grads'][k], g32), float(g.norm() / (g32.norm() + 1e-30))))
    for (k, c, cs, r) in rows[::6] + rows[-5:]:
        print('%-45s cos(ours,f32) %.4f   cos(same-precision oracle,f32) %.4f   norm ratio %.3f' % (k, c, cs, r))
    conv = [t for t in rows if t[0].endswith('.weight') and ('conv' in t[0] or 'downsample.0' in t[0] or t[0] in ('classifier.0.weight', 'classifier.4.weight'))]
    print('conv weight gradients vs f32: ours min %.4f median %.4f | same-precision oracle min %.4f median %.4f'
          % (min(t[1] for t in conv), np.median([t[1] for t in conv]), min(t[2] for t in conv), np.median([t[2] for t in conv])))
    assert all(0.7 < t[3] < 1.4 for t in conv), [t for t in conv if not 0.7 < t[3] < 1.4]
    assert conv[-1][1] > 0.99                               # classifier.4: nothing upstream of it is perturbed in the backward
    if strict:
        # as close to f32 as the same-precision torch implementation, layer by layer
        assert all(t[1] > t[2] - 0.12 for t in conv), [t for t in conv if not t[1] > t[2] - 0.12]
        assert np.median([t[1] for t in conv]) > np.median([t[2] for t in conv]) - 0.05
    ref = ref32
    # running statistics after the step (momentum 0.1)
    sd = tr.state_dict()
    for k in ('backbone.bn1.running_mean', 'backbone.layer3.2.bn2.running_var', 'classifier.1.running_mean'):
        a, b = sd[k].cpu(), ref['state_dict'][k]
        assert (a - b).abs().max() < 0.1 * b.abs().max() + 2e-2, k
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
    assert agree / total > (0.7 if strict else 0.55)


def test_train_forward_every_unit_teacher_forced(cuda_device, synthetic_sd):
    """TIGHT parity of the train-mode forward, unit by unit, at the full depth of the network: every convolution and every
    batch-statistics BatchNorm (+ residual add + ReLU) is recomputed in torch f32 from the kernel's OWN stored input of that
    unit (teacher forcing), so the layer-to-layer error amplification of a random-init train-mode network -- the reason
    the end-to-end comparison above needs loose bounds -- never enters: what remains is one bf16 rounding per stored value."""
    import torch.nn.functional as F
    from neuralbarkcalculator_b200.train import Trainer
    N, H, W = 2, 64, 96
    imgs, tgt, x = _batch(N, H, W, seed=5)
    sd = synthetic_sd
    tr = Trainer(sd, N, H, W, device='cuda:0', dropout=0.0, class_weights=torch.ones(3))
    tr.forward_backward(torch.from_numpy(imgs).to(cuda_device), torch.from_numpy(tgt).to(cuda_device), seed=1)
    torch.cuda.synchronize()

    def nchw(t):      # bf16 NHWC debug view -> f32 NCHW on the CPU
        return t.float().cpu().permute(0, 3, 1, 2).contiguous()

    def close(got, ref, what):      # one bf16 rounding of the stored value (2^-9 relative) + f32 summation-order noise
        err = (got - ref).abs()
        tol = 2.0 ** -7 * ref.abs() + 2e-3 * ref.abs().max()
        assert (err <= tol).all(), '%s: max err %.4g at |ref| %.4g (max |ref| %.4g)' % (
            what, err.max(), ref.flatten()[err.argmax()].abs(), ref.abs().max())

    def check(unit, conv_key, bn_key, xin, stride, pad, dil, relu, residual=None):
        w = sd[conv_key + '.weight'].to(torch.bfloat16).float()
        z = nchw(tr.debug_tensor(0, unit))
        close(z, F.conv2d(xin, w, stride=stride, padding=pad, dilation=dil), '%s z' % conv_key)
        mean, var = z.mean((0, 2, 3), keepdim=True), z.var((0, 2, 3), unbiased=False, keepdim=True)
        y_ref = (z - mean) / torch.sqrt(var + 1e-5) * sd[bn_key + '.weight'].view(1, -1, 1, 1) + sd[bn_key + '.bias'].view(1, -1, 1, 1)
        if residual is not None:
            y_ref = y_ref + residual
        if relu:
            y_ref = y_ref.relu()
        y = nchw(tr.debug_tensor(1, unit))
        close(y, y_ref, '%s y' % bn_key)
        return y

    xn = x.to(torch.bfloat16).float()      # the staging pass stores the normalised image in bf16
    y = check(0, 'backbone.conv1', 'backbone.bn1', xn, 2, 3, 1, True)
    cur = F.max_pool2d(y, 3, 2, 1)
    unit, dil = 1, 1
    for li, (nb, stride, dilate) in enumerate(((3, 1, False), (4, 2, False), (6, 1, True), (3, 1, True)), start=1):
        for b in range(nb):
            pre = 'backbone.layer%d.%d.' % (li, b)
            s_ = stride if (b == 0 and not dilate) else 1
            prev = dil
            if b == 0 and dilate:
                dil *= 2
            d2 = prev if b == 0 else dil
            a1 = check(unit, pre + 'conv1', pre + 'bn1', cur, 1, 0, 1, True)
            a2 = check(unit + 1, pre + 'conv2', pre + 'bn2', a1, s_, d2, d2, True)
            skip = cur
            if b == 0:
                skip = check(unit + 3, pre + 'downsample.0', pre + 'downsample.1', cur, s_, 0, 1, False)
            cur = check(unit + 2, pre + 'conv3', pre + 'bn3', a2, 1, 0, 1, True, residual=skip)
            unit += 4 if b == 0 else 3
    check(unit, 'classifier.0', 'classifier.1', cur, 1, 1, 1, True)
    assert unit + 1 == tr.lib.nbc_train_num_units(C.c_void_p(tr.handle))


def test_train_gradients_strict_on_tamed_network(cuda_device, synthetic_sd):
    """Strict check of every backward kernel.  Shrinking gamma of the last BN of each bottleneck (x0.05) makes the
    residual branches small perturbations of the identity trunk, which removes the layer-to-layer error amplification
    of the random-init network while every kernel (wgrad, dgrad incl. stride 2, BN / ReLU / maxpool / upsample /
    classifier backward) still produces its gradient: all of them must then match the f32 oracle closely."""
    from neuralbarkcalculator_b200.train import Trainer
    sd = {k: v.clone() for k, v in synthetic_sd.items()}
    for k in sd:
        if k.endswith('bn3.weight'):
            sd[k] *= 0.05
    N, H, W = 2, 64, 96
    imgs, tgt, x = _batch(N, H, W, seed=3)
    wt = torch.ones(3)
    ref = otrain.train_step(sd, x, torch.from_numpy(tgt).long(), weights=wt, dropout=0.0)
    tr = Trainer(sd, N, H, W, device='cuda:0', dropout=0.0, class_weights=wt)
    loss = float(tr.forward_backward(torch.from_numpy(imgs).to(cuda_device), torch.from_numpy(tgt).to(cuda_device)))
    print('\n[tamed] loss ours %.5f f32 oracle %.5f' % (loss, ref['loss']))
    assert abs(loss - ref['loss']) < 0.01 * abs(ref['loss'])
    full = tr.debug_tensor(3).cpu()
    assert (full - ref['logits']).abs().max() < 0.05 * ref['logits'].std()
    grads = tr.gradients()
    worst = []
    for k, g32 in ref['grads'].items():
        c = _cos(grads[k].cpu(), g32)
        r = float(grads[k].cpu().norm() / (g32.norm() + 1e-30))
        worst.append((c, r, k))
    worst.sort()
    for c, r, k in worst[:8]:
        print('[tamed] lowest cos: %-45s cos %.4f norm ratio %.3f' % (k, c, r))
    assert worst[0][0] > 0.94, worst[0]
    assert np.median([c for c, _, _ in worst]) > 0.97
    assert all(0.93 < r < 1.07 for _, r, _ in worst), [t for t in worst if not 0.93 < t[1] < 1.07]


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
    # a few full steps reduce the loss on a fixed batch (well-conditioned variant: small residual branches, continuous
    # loss -- the random-init network with the 93x class weight is chaotic enough for 8 Adam steps to go either way)
    sd = {k: (v * 0.05 if k.endswith('bn3.weight') else v.clone()) for k, v in synthetic_sd.items()}
    tr2 = Trainer(sd, N, H, W, device='cuda:0', dropout=0.0, class_weights=torch.ones(3))
    first = float(tr2.step(xi, ti))
    for _ in range(7):
        last = float(tr2.step(xi, ti))
    print('\nloss over 8 steps on a fixed batch: %.4f -> %.4f' % (first, last))
    assert last < first


@pytest.mark.parametrize('loss_kind', ['lovasz', 'mixed'])
def test_train_step_lovasz_and_mixed(cuda_device, synthetic_sd, loss_kind):
    """The step with the loss the reference's main() really uses (LovaszSoftmax, __main__.py:239) and with MixedLoss
    (utils.py:185-192), on the tamed network (see test_train_gradients_strict_on_tamed_network): loss and every
    gradient against the f32 oracle."""
    from neuralbarkcalculator_b200.train import Trainer
    sd = {k: v.clone() for k, v in synthetic_sd.items()}
    for k in sd:
        if k.endswith('bn3.weight'):
            sd[k] *= 0.05
    N, H, W = 2, 64, 96
    imgs, tgt, x = _batch(N, H, W, seed=3)
    wt = torch.ones(3)
    ref = otrain.train_step(sd, x, torch.from_numpy(tgt).long(), weights=wt, dropout=0.0, loss_kind=loss_kind)
    tr = Trainer(sd, N, H, W, device='cuda:0', dropout=0.0, class_weights=wt, loss=loss_kind)
    loss = float(tr.forward_backward(torch.from_numpy(imgs).to(cuda_device), torch.from_numpy(tgt).to(cuda_device)))
    print('\n[%s] loss ours %.5f f32 oracle %.5f' % (loss_kind, loss, ref['loss']))
    assert abs(loss - ref['loss']) < 0.01 * abs(ref['loss'])
    grads = tr.gradients()
    worst = sorted((_cos(grads[k].cpu(), g32), float(grads[k].cpu().norm() / (g32.norm() + 1e-30)), k) for k, g32 in ref['grads'].items())
    for c, r, k in worst[:5]:
        print('[%s] lowest cos: %-45s cos %.4f norm ratio %.3f' % (loss_kind, k, c, r))
    assert worst[0][0] > 0.9, worst[0]
    assert np.median([c for c, _, _ in worst]) > 0.97
    assert all(0.9 < r < 1.1 for _, r, _ in worst), [t for t in worst if not 0.9 < t[1] < 1.1]


def test_two_rank_nccl_gradient_exchange(cuda_device):
    """Data-parallel training over NCCL: on 2 GPUs the Trainer's bucketed, backward-overlapped all-reduce equals
    the sum of the two single-rank gradients bit for bit, equals one all-reduce of the flat buffer, and a step leaves both
    ranks with identical weights.  Skipped on a 1-GPU box (the driver's GPU tier has one GPU; gpurun --gpus 2 runs it)."""
    import socket
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'nccl_grad_worker.py')
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr',
                        '127.0.0.1', '--master-port', str(port), worker], capture_output=True, text=True, timeout=600)
    print(r.stdout[-2000:], r.stderr[-3000:])
    assert r.returncode == 0 and 'NCCL_GRAD_OK world=2' in r.stdout
