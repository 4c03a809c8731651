"""Oracle for ``CustomWeightedCrossEntropy`` (reference ``utils.py:151-165``).  Test infrastructure only.

Pure torch in the reference, restated verbatim; pinned by ``oracle/make_golden.py`` against the reference class."""
import torch
import torch.nn.functional as F

DEFAULT_WEIGHTS = [0.4004, 2.0334, 93.1921]  # utils.py:72-73


def custom_weighted_cross_entropy(predict, true, weights):
    """utils.py:157-165: CE(none) * weights[max(argmax(predict,1), true)], mean over all pixels."""
    ent = F.cross_entropy(predict, true, reduction='none')
    mc = torch.max(torch.argmax(predict, dim=1), true).flatten()
    w = torch.index_select(weights, 0, mc).view(true.shape)
    return (ent * w).mean()


def custom_weighted_cross_entropy_with_grad(predict, true, weights):
    p = predict.detach().clone().requires_grad_(True)
    loss = custom_weighted_cross_entropy(p, true, weights)
    loss.backward()
    return loss.detach(), p.grad
