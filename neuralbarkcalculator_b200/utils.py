"""Host-side mirror of the hot-path pieces of the reference's ``utils.py``: ``remove_small_zones`` (utils.py:135-148),
``CustomWeightedCrossEntropy`` (utils.py:151-165) and ``get_pos_weight`` (utils.py:72-73).  Same names, arguments
and in-place behaviour; the work is done by the CUDA kernels behind ``libnbc.so``."""
import torch
from torch import nn

import numpy as np

from . import ops
from .lovasz_losses import LovaszSoftmax


def get_pos_weight():
    """utils.py:72-73."""
    return torch.FloatTensor([0.4004, 2.0334, 93.1921])


def remove_small_zones(img, threshold=150):
    """In place on ``img`` (integer class tensor [B,H,W] or [H,W], CUDA) and returns it, like utils.py:135-148.

    Regions smaller than ``threshold`` pixels (8-connected) are flipped: small foreground blobs -> 0, then small
    background islands -> 1.  Labelling is per image (the reference's 3-D structuring element also links
    neighbouring batch entries, a bug that predict -- batch 1 -- never exercises)."""
    if not isinstance(img, torch.Tensor) or not img.is_cuda:
        raise RuntimeError('remove_small_zones: CUDA tensor required (no CPU path in neuralbarkcalculator_b200)')
    view = img if img.dim() == 3 else img.unsqueeze(0)
    if view.dtype == torch.uint8 and view.is_contiguous():
        ops.remove_small_zones_u8(view, threshold)
        return img
    m = view.to(torch.uint8).contiguous()
    ops.remove_small_zones_u8(m, threshold)
    view.copy_(m)
    return img


class _WCEFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, predict, true, weights):
        loss, grad = ops.wce_fwd_bwd(predict, true, weights, need_grad=True)
        ctx.save_for_backward(grad)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        (grad,) = ctx.saved_tensors
        return grad * grad_out, None, None


class CustomWeightedCrossEntropy(nn.Module):
    """utils.py:151-165: per-pixel CE weighted by ``weights[max(argmax(predict), true)]``, mean over all pixels.
    Forward and backward are one fused kernel (the gradient is produced in the forward pass)."""

    def __init__(self, weights):
        super().__init__()
        self.__name__ = "CustomWeightedCrossEntropy"
        self.weights = weights

    def forward(self, predict, true):
        w = self.weights.to(device=predict.device, dtype=torch.float32)
        return _WCEFunction.apply(predict.float(), true, w)


class MixedLoss(nn.Module):
    """utils.py:185-192: CustomWeightedCrossEntropy / 4 + LovaszSoftmax."""

    def __init__(self, cwe_weights):
        super(MixedLoss, self).__init__()
        self.ce = CustomWeightedCrossEntropy(cwe_weights)
        self.lovasz = LovaszSoftmax()

    def forward(self, predict, true):
        return self.ce(predict, true) / 4 + self.lovasz(predict, true)


class PixelWiseF1(nn.Module):
    """utils.py:201-235: per-class F1 of argmax(outputs) after remove_small_zones.  The argmax, the region removal and the
    3x3 confusion matrix run on the GPU (nbc_argmax3_u8, nbc_remove_small_zones, nbc_confusion_matrix); the F1 formula
    (sklearn f1_score(labels=[0,1,2], average=None) in the reference) is evaluated from the 9 counts on the host."""

    def __init__(self, class_to_watch):
        super().__init__()
        self.class_to_watch = class_to_watch
        if self.class_to_watch is None:
            self.__name__ = "PixelWiseF1"
        else:
            self.__name__ = "PixelWiseF1_class_{}".format(self.class_to_watch)

    def forward(self, outputs, labels):
        if not outputs.is_cuda:
            raise RuntimeError('PixelWiseF1: CUDA tensors required (no CPU path in neuralbarkcalculator_b200)')
        pred = ops.argmax3_u8(outputs.float())
        ops.remove_small_zones_u8(pred, 150)
        cm = ops.confusion_matrix(pred, labels.to(outputs.device)).cpu().numpy().astype(np.float64)
        tp = np.diag(cm)
        denom = cm.sum(0) + cm.sum(1)
        scores = np.where(denom > 0, 2 * tp / np.maximum(denom, 1), 0.0)
        targets_count, outputs_count = cm.sum(1), cm.sum(0)
        for i, count_i in enumerate(targets_count):
            if count_i == 0 and outputs_count[i] == 0:
                scores[i] = np.delete(scores, i).mean()
        if self.class_to_watch is None:
            return scores.mean()
        elif self.class_to_watch == 'loss':
            return 1 - scores.mean()
        elif isinstance(self.class_to_watch, int):
            return scores[self.class_to_watch]
        else:
            return scores
