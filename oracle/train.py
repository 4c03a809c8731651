"""Oracle for the training step (reference ``__main__.py:231-269`` with ``CustomWeightedCrossEntropy``,
``utils.py:151-165``, as the loss -- north_star's item (4); the Poutyne epoch loop is third-party and out of scope).
Test infrastructure only: a plain torch CPU step -- ``train()``, forward, loss, ``backward``, ``Adam(lr=5e-4,
weight_decay=2e-3).step()`` -- on the reference's own network (oracle/model.py)."""
import torch

from . import losses, model as omodel


def _simulate_bf16(net):
    """Round to bf16 at the points where the B200 path stores 16-bit values: conv weights (not the final 1x1), conv
    outputs, BN outputs that are stored (bn1 / bn2 / downsample / head BN), and the ReLU outputs (the post-residual
    activation of a bottleneck).  The casts are differentiable (identity backward), so autograd still works in f32."""
    def rnd(m, i, o):
        return o.bfloat16().float()
    for name, m in net.named_modules():
        if name == 'classifier.4' or name.endswith('bn3'):
            continue
        if isinstance(m, (torch.nn.Conv2d, torch.nn.BatchNorm2d, torch.nn.ReLU)):
            m.register_forward_hook(rnd)


def train_step(state_dict, x, target, weights=None, dropout=0.0, lr=5e-4, weight_decay=2e-3, steps=1, bf16_sim=False,
               loss_kind='weighted_ce'):
    """x: f32 [N,3,H,W] normalised; target: int64 [N,H,W].  Returns dict(loss, grads {name: tensor}, state_dict).

    bf16_sim=True reproduces the storage precision of the B200 path inside the torch oracle (see _simulate_bf16): a
    random-init train-mode BatchNorm network amplifies perturbations from layer to layer (~1.1x per bottleneck), so the
    f32 and the bf16 forward differ by ~10 % at layer4 for ANY bf16 implementation; parity of the kernels is therefore
    judged against this same-precision oracle, and the f32 gap is reported beside it."""
    net = omodel.fcn_resnet50(dropout=dropout)
    net.load_state_dict(state_dict, strict=True)
    net.train()
    if bf16_sim:
        _simulate_bf16(net)
        x = x.bfloat16().float()
    w = torch.tensor(losses.DEFAULT_WEIGHTS) if weights is None else weights
    opt = torch.optim.Adam(net.parameters(), lr=lr, weight_decay=weight_decay)
    out = {}
    for _ in range(steps):
        opt.zero_grad()
        if bf16_sim:
            with torch.no_grad():
                for n_, p_ in net.named_parameters():
                    if p_.dim() == 4 and not n_.startswith('classifier.4'):
                        p_.data = p_.data.bfloat16().float()   # weights as the tensor cores see them (master copy not kept: 1 step)
        logits = net(x)
        if loss_kind == 'weighted_ce':      # utils.py:151-165 (north_star's loss)
            loss = losses.custom_weighted_cross_entropy(logits, target, w)
        elif loss_kind == 'lovasz':         # __main__.py:239, the loss main() really uses
            from . import lovasz
            loss = lovasz.lovasz_softmax(logits, target)
        else:                               # utils.py:185-192 MixedLoss
            from . import lovasz
            loss = lovasz.mixed_loss(logits, target, w)
        loss.backward()
        out['loss'] = float(loss)
        out['grads'] = {k: p.grad.detach().clone() for k, p in net.named_parameters()}
        opt.step()
    out['state_dict'] = {k: v.detach().clone() for k, v in net.state_dict().items()}
    out['logits'] = logits.detach()
    return out
