"""Oracle for the training step (reference ``__main__.py:231-269`` with ``CustomWeightedCrossEntropy``,
``utils.py:151-165``, as the loss -- north_star's item (4); the Poutyne epoch loop is third-party and out of scope).
Test infrastructure only: a plain torch CPU step -- ``train()``, forward, loss, ``backward``, ``Adam(lr=5e-4,
weight_decay=2e-3).step()`` -- on the reference's own network (oracle/model.py)."""
import torch

from . import losses, model as omodel


def train_step(state_dict, x, target, weights=None, dropout=0.0, lr=5e-4, weight_decay=2e-3, steps=1):
    """x: f32 [N,3,H,W] normalised; target: int64 [N,H,W].  Returns dict(loss, grads {name: tensor}, state_dict)."""
    net = omodel.fcn_resnet50(dropout=dropout)
    net.load_state_dict(state_dict, strict=True)
    net.train()
    w = torch.tensor(losses.DEFAULT_WEIGHTS) if weights is None else weights
    opt = torch.optim.Adam(net.parameters(), lr=lr, weight_decay=weight_decay)
    out = {}
    for _ in range(steps):
        opt.zero_grad()
        logits = net(x)
        loss = losses.custom_weighted_cross_entropy(logits, target, w)
        loss.backward()
        out['loss'] = float(loss)
        out['grads'] = {k: p.grad.detach().clone() for k, p in net.named_parameters()}
        opt.step()
    out['state_dict'] = {k: v.detach().clone() for k, v in net.state_dict().items()}
    out['logits'] = logits.detach()
    return out
