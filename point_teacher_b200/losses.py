"""Student-branch losses whose kernels are mmcv-native in the reference, behind the
reference's registry names and ``forward`` signatures:

  * ``FocalLoss``       HBB_TOD/mmdet/models/losses/focal_loss.py:103-182 (``mmcv.ops.sigmoid_focal_loss``)
  * ``RotatedIoULoss``  OBB_TOD/mmrotate/models/losses/rotated_iou_loss.py:149-227
  * ``DN_IoULoss``      OBB_TOD/mmrotate/models/losses/rotated_iou_loss.py:229-320   (``mmcv.ops.diff_iou_rotated_2d``)

Each forward is ONE kernel that also produces the exact gradient (csrc/losses.cu); the autograd Functions only scale
it.  CUDA tensors only -- there is no CPU fallback."""
import torch
import torch.nn as nn

from . import _lib
from .ops import _p, _stream
from .registry import LOSSES, ROTATED_LOSSES

_MODES = {"log": 0, "linear": 1, "square": 2}


def _need_cuda(t, name):
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (the B200 path has no CPU fallback)")


class _FocalFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, weight, wmode, gamma, alpha, want_elem):
        N, C = pred.shape
        dev = pred.device
        p = pred.detach().float().contiguous()
        need_grad = pred.requires_grad
        elem = torch.empty((N, C), dtype=torch.float32, device=dev) if want_elem else None
        grad = torch.empty((N, C), dtype=torch.float32, device=dev) if need_grad else None
        total = torch.zeros((1,), dtype=torch.float32, device=dev)
        w = None if weight is None else weight.detach().float().contiguous()
        _lib.call("pt_sigmoid_focal_loss", _p(p), _p(target.long().contiguous()), _p(w), wmode, float(gamma),
                  float(alpha), N, C, _p(elem), _p(grad), _p(None if want_elem else total), _stream())
        ctx.want_elem = want_elem
        ctx.save_for_backward(grad)
        return elem if want_elem else total.reshape(())

    @staticmethod
    def backward(ctx, gout):
        (grad,) = ctx.saved_tensors
        return (grad * gout if grad is not None else None), None, None, None, None, None, None


def sigmoid_focal_loss(pred, target, weight=None, gamma=2.0, alpha=0.25, reduction="mean", avg_factor=None):
    """focal_loss.py:59-100: pred (N,C) logits, target (N,) labels in [0, C]; weight (N,) | (N,C) | (N*C,) | None."""
    _need_cuda(pred, "pred")
    if pred.dim() != 2 or target.shape != pred.shape[:1]:
        raise ValueError("pred must be (N, C) and target (N,)")
    N, C = pred.shape
    wmode = 0
    if weight is not None:
        if weight.numel() == N and (weight.dim() == 1 or weight.shape == (N, 1)):
            wmode = 1
        elif weight.numel() == N * C:
            wmode = 2
        else:
            raise AssertionError("weight must have N or N*C elements")
    if N == 0:
        z = pred.sum() * 0
        return pred.new_zeros((0, C)) if reduction == "none" else z
    if reduction == "none":
        if avg_factor is not None:
            pass
        return _FocalFn.apply(pred, target, weight, wmode, gamma, alpha, True)
    total = _FocalFn.apply(pred, target, weight, wmode, gamma, alpha, False)
    if avg_factor is None:
        return total / (N * C) if reduction == "mean" else total
    if reduction == "mean":
        return total / avg_factor
    raise ValueError('avg_factor can not be used with reduction="sum"')


@LOSSES.register_module(name="FocalLoss", force=True)
class FocalLoss(nn.Module):
    def __init__(self, use_sigmoid=True, gamma=2.0, alpha=0.25, reduction="mean", loss_weight=1.0):
        super().__init__()
        assert use_sigmoid is True, "Only sigmoid focal loss supported now."
        self.use_sigmoid, self.gamma, self.alpha = use_sigmoid, gamma, alpha
        self.reduction, self.loss_weight = reduction, loss_weight

    def forward(self, pred, target, weight=None, avg_factor=None, reduction_override=None):
        assert reduction_override in (None, "none", "mean", "sum")
        reduction = reduction_override if reduction_override else self.reduction
        return self.loss_weight * sigmoid_focal_loss(pred, target, weight, gamma=self.gamma, alpha=self.alpha,
                                                     reduction=reduction, avg_factor=avg_factor)


class _RotIoUFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, mode, eps, dn, hyper):
        n = pred.shape[0]
        p = pred.detach().float().contiguous()
        t = target.detach().float().contiguous()
        loss = torch.empty((n,), dtype=torch.float32, device=pred.device)
        grad = torch.empty((n, 5), dtype=torch.float32, device=pred.device) if pred.requires_grad else None
        _lib.call("pt_rotated_iou_loss", _p(p), _p(t), n, _MODES[mode], float(eps), int(dn), float(hyper), _p(loss),
                  _p(grad), _stream())
        ctx.save_for_backward(grad)
        return loss

    @staticmethod
    def backward(ctx, gout):
        (grad,) = ctx.saved_tensors
        return (grad * gout.unsqueeze(1) if grad is not None else None), None, None, None, None, None


def _reduce(loss, weight, reduction, avg_factor):
    """losses/utils.py:25-54."""
    if weight is not None:
        loss = loss * weight
    if avg_factor is None:
        return loss.mean() if reduction == "mean" else (loss.sum() if reduction == "sum" else loss)
    if reduction == "mean":
        return loss.sum() / avg_factor
    if reduction != "none":
        raise ValueError('avg_factor can not be used with reduction="sum"')
    return loss


class _RotatedLossBase(nn.Module):
    _dn = False

    def __init__(self, linear=False, eps=1e-6, reduction="mean", loss_weight=1.0, mode="log", hyper=0.2):
        super().__init__()
        assert mode in ("linear", "square", "log")
        if linear:
            mode = "linear"
        self.mode, self.linear, self.eps = mode, linear, eps
        self.reduction, self.loss_weight, self.hyper = reduction, loss_weight, hyper

    def forward(self, pred, target, weight=None, avg_factor=None, reduction_override=None, **kwargs):
        assert reduction_override in (None, "none", "mean", "sum")
        _need_cuda(pred, "pred")
        reduction = reduction_override if reduction_override else self.reduction
        if (weight is not None) and (not torch.any(weight > 0)) and (reduction != "none"):
            if pred.dim() == weight.dim() + 1:
                weight = weight.unsqueeze(1)
            return (pred * weight).sum()
        if weight is not None and weight.dim() > 1:
            assert weight.shape == pred.shape
            weight = weight.mean(-1)
        if pred.dim() != 2 or pred.shape[1] != 5 or target.shape != pred.shape:
            raise ValueError("pred / target must be (n, 5) rotated boxes (cx, cy, w, h, theta)")
        if pred.shape[0] == 0:
            return pred.sum() * 0 if reduction != "none" else pred.new_zeros((0,))
        elem = _RotIoUFn.apply(pred, target, self.mode, self.eps, self._dn, self.hyper)
        return self.loss_weight * _reduce(elem, weight, reduction, avg_factor)


@ROTATED_LOSSES.register_module(name="RotatedIoULoss", force=True)
class RotatedIoULoss(_RotatedLossBase):
    """rotated_iou_loss.py:149-227."""

    def __init__(self, linear=False, eps=1e-6, reduction="mean", loss_weight=1.0, mode="log"):
        super().__init__(linear, eps, reduction, loss_weight, mode)


@ROTATED_LOSSES.register_module(name="DN_IoULoss", force=True)
class DN_IoULoss(_RotatedLossBase):
    """rotated_iou_loss.py:229-320: element-wise (IoU loss + min over the 3 x 3 size-jittered targets) / 2."""
    _dn = True
