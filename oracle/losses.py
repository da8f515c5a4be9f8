"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the two remaining mmcv-native losses of the student branch
(SURVEY.md section 8f rank 4):

  * ``FocalLoss`` -- HBB_TOD/mmdet/models/losses/focal_loss.py:11-57 (``py_sigmoid_focal_loss``, the path the reference
    itself takes on CPU, :166-168) + ``weight_reduce_loss`` (losses/utils.py:25-54).  Pinned bit-exact against the
    reference's own ``FocalLoss`` under the import shim (oracle/check_oracle_vs_ref.py --losses).
  * ``RotatedIoULoss`` / ``DN_IoULoss`` -- OBB_TOD/mmrotate/models/losses/rotated_iou_loss.py:17-58, 60-147, 149-320.
    Their kernel is ``mmcv.ops.diff_iou_rotated_2d`` (mmcv-full >= 1.5.0, absent from /root/reference; PARITY UNPINNED
    by any reference test): the published algorithm is restated here in PyTorch (box2corners -> 16 edge/edge
    intersections + corner-in-box tests -> 24 candidate vertices -> angular sort (mmcv's ``sort_vertices`` op, restated
    in numpy) -> shoelace area), differentiable through torch autograd exactly like the original.  The loss wrappers
    around it are pinned against the reference's own file with this function injected as ``diff_iou_rotated_2d``.
    Cross-checks in tests/test_oracle.py: theta = 0 == axis-aligned IoU; == the polygon-clipping IoU of
    oracle/rotated.py (a different algorithm) on random pairs.
"""
import numpy as np
import torch
import torch.nn.functional as F

EPSILON = 1e-8


# ------------------------------------------------------------------------------ reduction helpers (losses/utils.py)
def weight_reduce_loss(loss, weight=None, reduction="mean", avg_factor=None):
    if weight is not None:
        loss = loss * weight
    if avg_factor is None:
        if reduction == "mean":
            return loss.mean()
        if reduction == "sum":
            return loss.sum()
        return loss
    if reduction == "mean":
        return loss.sum() / avg_factor
    if reduction != "none":
        raise ValueError('avg_factor can not be used with reduction="sum"')
    return loss


# ------------------------------------------------------------------------------ focal loss
def sigmoid_focal_loss(pred, target, weight=None, gamma=2.0, alpha=0.25, reduction="mean", avg_factor=None):
    """focal_loss.py:11-57 with the label -> one-hot step of FocalLoss.forward (:164-166); target (N,) int64 in [0, C]."""
    C = pred.size(1)
    t = F.one_hot(target, num_classes=C + 1)[:, :C].type_as(pred)
    p = pred.sigmoid()
    pt = (1 - p) * t + p * (1 - t)
    fw = (alpha * t + (1 - alpha) * (1 - t)) * pt.pow(gamma)
    loss = F.binary_cross_entropy_with_logits(pred, t, reduction="none") * fw
    if weight is not None:
        if weight.shape != loss.shape:
            if weight.size(0) == loss.size(0):
                weight = weight.view(-1, 1)
            else:
                assert weight.numel() == loss.numel()
                weight = weight.view(loss.size(0), -1)
        assert weight.ndim == loss.ndim
    return weight_reduce_loss(loss, weight, reduction, avg_factor)


# ------------------------------------------------------------------------------ diff_iou_rotated_2d (mmcv, restated)
def box2corners(box):
    """(B,N,5) -> (B,N,4,2); corners (+,+), (-,+), (-,-), (+,-) of the half extents, rotated by alpha."""
    B = box.size(0)
    x, y, w, h, alpha = box.split([1, 1, 1, 1, 1], dim=-1)
    x4 = box.new_tensor([0.5, -0.5, -0.5, 0.5]) * w
    y4 = box.new_tensor([0.5, 0.5, -0.5, -0.5]) * h
    corners = torch.stack([x4, y4], dim=-1)
    sin, cos = torch.sin(alpha), torch.cos(alpha)
    rot_t = torch.stack([torch.cat([cos, sin], dim=-1), torch.cat([-sin, cos], dim=-1)], dim=-2)
    rotated = torch.bmm(corners.view(-1, 4, 2), rot_t.view(-1, 2, 2)).view(B, -1, 4, 2)
    rotated = rotated + torch.cat([x, y], dim=-1).unsqueeze(2)
    return rotated


def box_intersection(c1, c2):
    l1 = torch.cat([c1, c1[:, :, [1, 2, 3, 0], :]], dim=3).unsqueeze(3)
    l2 = torch.cat([c2, c2[:, :, [1, 2, 3, 0], :]], dim=3).unsqueeze(2)
    x1, y1, x2, y2 = l1.split([1, 1, 1, 1], dim=-1)
    x3, y3, x4, y4 = l2.split([1, 1, 1, 1], dim=-1)
    num = (x1 - x2) * (y3 - y4) - (y1 - y2) * (x3 - x4)
    den_t = (x1 - x3) * (y3 - y4) - (y1 - y3) * (x3 - x4)
    t = den_t / num
    t = torch.where(num == 0.0, torch.full_like(t, -1.0), t)
    mask_t = (t > 0) & (t < 1)
    den_u = (x1 - x2) * (y1 - y3) - (y1 - y2) * (x1 - x3)
    u = -den_u / num
    u = torch.where(num == 0.0, torch.full_like(u, -1.0), u)
    mask = mask_t & (u > 0) & (u < 1)
    t = den_t / (num + EPSILON)
    inter = torch.stack([x1 + t * (x2 - x1), y1 + t * (y2 - y1)], dim=-1) * mask.float().unsqueeze(-1)
    return inter.squeeze(-2) if inter.dim() == 7 else inter, mask


def box1_in_box2(c1, c2):
    a, b, d = c2[:, :, 0:1, :], c2[:, :, 1:2, :], c2[:, :, 3:4, :]
    ab, am, ad = b - a, c1 - a, d - a
    p_ab, n_ab = torch.sum(ab * am, dim=-1), torch.sum(ab * ab, dim=-1)
    p_ad, n_ad = torch.sum(ad * am, dim=-1), torch.sum(ad * ad, dim=-1)
    return ((p_ab / n_ab > -1e-6) & (p_ab / n_ab < 1 + 1e-6)) & ((p_ad / n_ad > -1e-6) & (p_ad / n_ad < 1 + 1e-6))


def _cmp(x1, y1, x2, y2):
    """mmcv ``compare_vertices`` (sort_vert kernel): vertex 1 comes before vertex 2 in the angular order."""
    f = np.float32
    eps = f(EPSILON)
    same = (np.abs(x1 - x2) < eps) & (np.abs(y2 - y1) < eps)
    n1 = (x1 * x1 + y1 * y1 + eps).astype(f)
    n2 = (x2 * x2 + y2 * y2 + eps).astype(f)
    diff = (np.abs(x1) * x1 / n1 - np.abs(x2) * x2 / n2).astype(f)
    res = np.zeros(x1.shape, bool)
    both_pos, both_neg = (y1 > 0) & (y2 > 0), (y1 < 0) & (y2 < 0)
    res = np.where(both_pos, diff > eps, res)
    res = np.where(both_neg, diff < eps, res)
    res = np.where((y1 > 0) & (y2 < 0), True, res)
    res = np.where((y1 < 0) & (y2 > 0), False, res)
    return res & ~same


def sort_vertices(vn, mask, num_valid):
    """mmcv ``sort_vertices_forward``: vn (M,24,2) fp32 centred vertices, mask (M,24) -> (M,9) int64 indices."""
    M, m = mask.shape
    f = np.float32
    vn, idx = vn.astype(f), np.zeros((M, 9), np.int64)
    inv = ~mask[:, 8:]
    pad = np.where(inv.any(1), inv.argmax(1) + 8, 8)
    rows = np.arange(M)
    x_all, y_all = vn[:, :, 0], vn[:, :, 1]
    for j in range(int(num_valid.max()) if M else 0):
        x_min, y_min = np.ones(M, f), np.full(M, -f(EPSILON), f)
        take = np.zeros(M, np.int64)
        if j:
            x2, y2 = x_all[rows, idx[:, j - 1]], y_all[rows, idx[:, j - 1]]
        for k in range(m):
            x, y = x_all[:, k], y_all[:, k]
            c = mask[:, k] & _cmp(x, y, x_min, y_min)
            if j:
                c &= _cmp(x2, y2, x, y)
            x_min, y_min, take = np.where(c, x, x_min), np.where(c, y, y_min), np.where(c, k, take)
        idx[:, j] = np.where(j < num_valid, take, idx[:, j])
    for i in range(M):
        nv = int(num_valid[i])
        if nv < 3:
            idx[i, :] = pad[i]
            continue
        idx[i, nv] = idx[i, 0]
        idx[i, nv + 1:] = pad[i]
        if nv == 8:
            counter = sum(int(idx[i, k] == idx[i, j]) for j in range(4) for k in range(4, 8))
            if counter == 4:
                idx[i, 4] = idx[i, 0]
                idx[i, 5:] = pad[i]
    return idx


def diff_iou_rotated_2d(box1, box2):
    """(B,N,5) x (B,N,5) -> (B,N) IoU, differentiable (mmcv/ops/diff_iou_rotated.py)."""
    c1, c2 = box2corners(box1), box2corners(box2)
    inter, imask = box_intersection(c1, c2)
    c12, c21 = box1_in_box2(c1, c2), box1_in_box2(c2, c1)
    B, N = c1.shape[:2]
    vertices = torch.cat([c1, c2, inter.reshape(B, N, -1, 2)], dim=2)
    mask = torch.cat([c12, c21, imask.reshape(B, N, -1)], dim=2)
    num_valid = mask.int().sum(dim=2)
    mean = torch.sum(vertices * mask.float().unsqueeze(-1), dim=2, keepdim=True) / num_valid.unsqueeze(-1).unsqueeze(-1)
    vn = (vertices - mean).detach()
    idx = sort_vertices(vn.reshape(B * N, 24, 2).numpy(), mask.reshape(B * N, 24).numpy(),
                        num_valid.reshape(-1).numpy())
    idx = torch.from_numpy(idx).view(B, N, 9)
    sel = torch.gather(vertices, 2, idx.unsqueeze(-1).repeat(1, 1, 1, 2))
    total = sel[:, :, 0:-1, 0] * sel[:, :, 1:, 1] - sel[:, :, 0:-1, 1] * sel[:, :, 1:, 0]
    area = torch.abs(total.sum(dim=2)) / 2
    a1, a2 = box1[:, :, 2] * box1[:, :, 3], box2[:, :, 2] * box2[:, :, 3]
    return area / (a1 + a2 - area)


# ------------------------------------------------------------------------------ loss wrappers (rotated_iou_loss.py)
def rotated_iou_loss_elem(pred, target, mode="log", eps=1e-6):
    """:17-58 / :60-101 element-wise part."""
    ious = diff_iou_rotated_2d(pred.unsqueeze(0), target.unsqueeze(0)).squeeze(0).clamp(min=eps)
    if mode == "linear":
        return 1 - ious
    if mode == "square":
        return 1 - ious ** 2
    return -ious.log()


def dn_iou_loss_elem(pred, target, hyper=0.2, mode="log", eps=1e-6):
    """:105-146: element-wise (base + min over the 3 x 3 size-jittered targets) / 2."""
    base = rotated_iou_loss_elem(pred, target, mode, eps)
    anx = hyper / 2
    w, h = target[:, 2], target[:, 3]
    bank = []
    for i in (-1, 0, 1):
        for j in (-1, 0, 1):
            t = target.clone()
            t[:, 2] = t[:, 2] - anx * w * i
            t[:, 3] = t[:, 3] - anx * h * j
            bank.append(rotated_iou_loss_elem(pred, t, mode, eps).reshape(-1, 1))
    return (base + torch.min(torch.cat(bank, dim=1), dim=1)[0]) / 2


def rotated_loss_forward(elem_fn, pred, target, weight=None, avg_factor=None, reduction="mean", loss_weight=1.0, **kw):
    """``RotatedIoULoss.forward`` / ``DN_IoULoss.forward`` (:183-227, :265-320)."""
    if (weight is not None) and (not torch.any(weight > 0)) and (reduction != "none"):
        if pred.dim() == weight.dim() + 1:
            weight = weight.unsqueeze(1)
        return (pred * weight).sum()
    if weight is not None and weight.dim() > 1:
        assert weight.shape == pred.shape
        weight = weight.mean(-1)
    return loss_weight * weight_reduce_loss(elem_fn(pred, target, **kw), weight, reduction, avg_factor)
