"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the dense-head label assignment of Point Teacher
(SURVEY.md section 8 rows a13-a15).  Paths relative to /root/reference/HBB_TOD/mmdet/core/bbox/.

Pinned by ``python -m oracle.check_oracle_vs_ref --assign`` against the reference's own ``TopkAssigner``,
``FUSETopkAssigner``, ``MaxIoUAssigner``, ``BboxDistanceMetric`` and match costs (unmodified files under the import
shim) and by the reference's known answers (tests/test_oracle.py).  ``torch.topk`` / ``Tensor.max`` are used as the
tie rule on purpose: the reference's indices are whatever ATen's CPU kernels return.
"""
import torch


# ------------------------------------------------------------------------------ match costs (match_costs/match_cost.py)
def focal_loss_cost(cls_pred, gt_labels, weight=1.0, alpha=0.25, gamma=2, eps=1e-12):
    """:54-100."""
    p = cls_pred.sigmoid()
    neg = -(1 - p + eps).log() * (1 - alpha) * p.pow(gamma)
    pos = -(p + eps).log() * alpha * (1 - p).pow(gamma)
    return (pos[:, gt_labels] - neg[:, gt_labels]) * weight


def focal_loss_table(cls_pred, alpha=0.25, gamma=2, eps=1e-12):
    """The per-class table the cost above gathers its columns from (P, C)."""
    p = cls_pred.sigmoid()
    neg = -(1 - p + eps).log() * (1 - alpha) * p.pow(gamma)
    pos = -(p + eps).log() * alpha * (1 - p).pow(gamma)
    return pos - neg


def point_cost(points, gts, mode="L1", weight=1.0):
    """:188-214: distance between the first two columns of both arguments."""
    d = points[:, None, :2] - gts[None, :, :2]
    dist = d.abs().sum(2) if mode == "L1" else (d ** 2).sum(2).sqrt()
    return dist * weight


def insider_cost(boxes_cxcywh, gts, weight=1.0):
    """:217-252: 0 where GT point g lies inside predicted box p (inclusive), else 1."""
    b = boxes_cxcywh[:, :4]
    x1, y1 = b[:, 0] - b[:, 2] / 2, b[:, 1] - b[:, 3] / 2
    x2, y2 = b[:, 0] + b[:, 2] / 2, b[:, 1] + b[:, 3] / 2
    gx, gy = gts[:, 0][None, :], gts[:, 1][None, :]
    inside = (gx >= x1[:, None]) & (gx <= x2[:, None]) & (gy >= y1[:, None]) & (gy <= y2[:, None])
    return (~inside).to(boxes_cxcywh.dtype) * weight


# ------------------------------------------------------------------------------ two-stage top-k assignment
def _two_stage(reg_cost, cost2, gt_labels, num_pre, topk):
    """assigners/topk_assigner.py:118-147 == fuse_topk_assigner.py:96-119.  Quirks kept: the second-stage
    ``topk`` runs over ALL G columns of the candidate rows and its flattened result is assigned to GT i; a later
    GT overwrites an earlier one."""
    P, G = reg_cost.shape
    gt_inds = torch.zeros(P, dtype=torch.long)
    labels = torch.full((P,), -1, dtype=torch.long)
    _, pre = torch.topk(reg_cost, num_pre, dim=0, largest=False)          # (num_pre, G), ATen CPU tie rule
    for i in range(G):
        rows = pre[:, i]
        if rows.numel() <= topk:
            sel = rows
        else:
            _, t = torch.topk(cost2[rows, :], topk, dim=0, largest=False)  # (topk, G): every column votes
            sel = rows[t.flatten()]
        gt_inds[sel] = i + 1
        labels[sel] = gt_labels[i]
    return gt_inds, labels


def topk_assign(bbox_pred, cls_pred, gt_bboxes, gt_labels, num_pre=3, topk=3, cls_weight=1.0, reg_weight=1.0,
                reg_mode="L1"):
    """TopkAssigner.assign (assigners/topk_assigner.py:54-147) with PointCost + FocalLossCost."""
    P = bbox_pred.shape[0]
    if gt_bboxes is None or gt_bboxes.shape[0] == 0:
        return torch.zeros(P, dtype=torch.long), torch.full((P,), -1, dtype=torch.long)
    reg = point_cost(bbox_pred, gt_bboxes, reg_mode, reg_weight)
    cls = focal_loss_cost(cls_pred, gt_labels, cls_weight)
    return _two_stage(reg, cls, gt_labels, num_pre, topk)


def fuse_topk_assign(bbox_pred, points, cls_pred, gt_bboxes, gt_labels, num_pre=5, topk=3, cls_weight=1.0,
                     reg_weight=1.0, loc_weight=1.0, reg_mode="L1"):
    """FUSETopkAssigner.assign (assigners/fuse_topk_assigner.py:56-121): stage 1 on the grid points, stage 2 on
    FocalLossCost + InsiderCost of the decoded boxes (cx, cy, w, h)."""
    P = bbox_pred.shape[0]
    if gt_bboxes is None or gt_bboxes.shape[0] == 0:
        return torch.zeros(P, dtype=torch.long), torch.full((P,), -1, dtype=torch.long)
    reg = point_cost(points, gt_bboxes, reg_mode, reg_weight)
    cost2 = focal_loss_cost(cls_pred, gt_labels, cls_weight) + insider_cost(bbox_pred, gt_bboxes, loc_weight)
    return _two_stage(reg, cost2, gt_labels, num_pre, topk)


# ------------------------------------------------------------------------------ metric matrix (metric_calculator.py)
def bbox_metric(b1, b2, mode="iou", eps=1e-6):
    """iou_calculators/metric_calculator.py:44-185 (M x N only; ``is_aligned`` is ignored by the reference too).
    Quirks kept: eps is added into the union AND the union is clamped by eps; 'iof' returns the IoU."""
    rows, cols = b1.shape[0], b2.shape[0]
    if rows * cols == 0:
        return b1.new_zeros((rows, cols))
    a1 = (b1[:, 2] - b1[:, 0]) * (b1[:, 3] - b1[:, 1])
    a2 = (b2[:, 2] - b2[:, 0]) * (b2[:, 3] - b2[:, 1])
    lt = torch.max(b1[:, None, :2], b2[None, :, :2])
    rb = torch.min(b1[:, None, 2:], b2[None, :, 2:])
    wh = (rb - lt).clamp(min=0)
    overlap = wh[..., 0] * wh[..., 1]
    union = a1[:, None] + a2[None, :] - overlap + eps
    e = union.new_tensor([eps])
    union = torch.max(union, e)
    ious = overlap / union
    if mode in ("iou", "iof"):
        return ious
    if mode == "giou":
        elt = torch.min(b1[:, None, :2], b2[None, :, :2])
        erb = torch.max(b1[:, None, 2:], b2[None, :, 2:])
        ewh = (erb - elt).clamp(min=0)
        earea = torch.max(ewh[..., 0] * ewh[..., 1], e)
        return ious - (earea - union) / earea
    c1 = (b1[:, None, :2] + b1[:, None, 2:]) / 2
    c2 = (b2[None, :, :2] + b2[None, :, 2:]) / 2
    d = c1 - c2
    if mode == "center_distance2":
        return d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1] + 1e-6
    w1 = b1[:, None, 2] - b1[:, None, 0] + eps
    h1 = b1[:, None, 3] - b1[:, None, 1] + eps
    w2 = b2[None, :, 2] - b2[None, :, 0] + eps
    h2 = b2[None, :, 3] - b2[None, :, 1] + eps
    if mode == "wd":
        cd = d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1] + eps
        return 1 / (1 + (cd + ((w1 - w2) ** 2 + (h1 - h2) ** 2) / 4))
    kl = (w2 ** 2 / w1 ** 2 + h2 ** 2 / h1 ** 2 + 4 * d[..., 0] ** 2 / w1 ** 2 + 4 * d[..., 1] ** 2 / h1 ** 2
          + torch.log(w1 ** 2 / w2 ** 2) + torch.log(h1 ** 2 / h2 ** 2) - 2) / 2
    if mode == "kl":
        return 1 / (1 + kl)
    if mode == "kl_10":
        return 1 / (10 + kl)
    if mode == "exp_kl":
        return torch.exp(-kl / 10)
    raise ValueError(mode)


# ------------------------------------------------------------------------------ MaxIoUAssigner
def max_iou_assign(overlaps, gt_labels=None, pos_iou_thr=0.5, neg_iou_thr=0.5, min_pos_iou=0.0,
                   gt_max_assign_all=True, match_low_quality=True):
    """assign_wrt_overlaps (assigners/max_iou_assigner.py:127-212); overlaps (G, A)."""
    G, A = overlaps.shape
    gt_inds = torch.full((A,), -1, dtype=torch.long)
    if G == 0 or A == 0:
        if G == 0:
            gt_inds[:] = 0
        lab = None if gt_labels is None else torch.full((A,), -1, dtype=torch.long)
        return gt_inds, overlaps.new_zeros((A,)), lab
    mx, amx = overlaps.max(dim=0)
    gmx, gamx = overlaps.max(dim=1)
    if isinstance(neg_iou_thr, float):
        gt_inds[(mx >= 0) & (mx < neg_iou_thr)] = 0
    else:
        gt_inds[(mx >= neg_iou_thr[0]) & (mx < neg_iou_thr[1])] = 0
    pos = mx >= pos_iou_thr
    gt_inds[pos] = amx[pos] + 1
    if match_low_quality:
        for i in range(G):
            if gmx[i] >= min_pos_iou:
                if gt_max_assign_all:
                    gt_inds[overlaps[i, :] == gmx[i]] = i + 1
                else:
                    gt_inds[gamx[i]] = i + 1
    lab = None
    if gt_labels is not None:
        lab = torch.full((A,), -1, dtype=torch.long)
        p = gt_inds > 0
        lab[p] = gt_labels[gt_inds[p] - 1]
    return gt_inds, mx, lab


# ------------------------------------------------------------------------------ coarse pseudo boxes (section 8f rank 1)
def generate_pseudo_single(gt_points, gt_labels, gt_bboxes, cls_scores, bbox_preds, points, filter_scores, num_pre=5,
                           topk=3):
    """models/dense_heads/fcos_head_p2b_ts.py:736-794 with the FUSE assigner of the shipped config (5, 3)."""
    from . import hbb
    P, G = points.shape[0], gt_labels.shape[0]
    act = cls_scores.detach().sigmoid()
    dec = torch.stack([points[:, 0] - bbox_preds[:, 0], points[:, 1] - bbox_preds[:, 1], points[:, 0] + bbox_preds[:, 2],
                       points[:, 1] + bbox_preds[:, 3]], -1)                                   # distance2bbox
    gt_inds, lab = fuse_topk_assign(hbb.xyxy_to_cxcywh(dec), points, cls_scores, gt_points, gt_labels, num_pre, topk)
    pos = (gt_inds != 0).nonzero().reshape(-1)
    pos = pos[torch.sort(gt_inds[pos] - 1)[1]]
    labels = torch.zeros(P, dtype=torch.long)
    labels[pos] = lab[pos]
    s = act[torch.arange(P), labels]
    A, B, Cw = dec[pos], gt_inds[pos] - 1, s[pos]
    nums = torch.bincount(B, minlength=G)
    boxes = 8 * torch.ones_like(gt_bboxes)
    boxes[:, :2] = gt_points
    boxes = hbb.cxcywh_to_xyxy(boxes)
    scores = torch.zeros(G)
    pts = gt_points.clone()
    onehot = torch.nn.functional.one_hot(B, num_classes=G).float()
    bsum, ssum = onehot.t() @ (A * Cw[:, None]), onehot.t() @ Cw
    has = (nums != 0).nonzero().reshape(-1)
    boxes[has] = bsum[has] / ssum[has][:, None]
    scores[has] = ssum[has] / nums[has]
    pts[has] = hbb.xyxy_to_cxcywh(boxes[has])[:, :2]
    miou = hbb.bbox_overlaps(boxes[has], gt_bboxes[has], "iou", True).mean()
    valid = torch.zeros(G, dtype=torch.bool)
    valid[has] = scores[has] >= filter_scores
    return boxes, pts, gt_labels, miou, valid.nonzero().reshape(-1), scores, nums


def get_target_pseudo_single(points, cls_scores, gt_points, gt_labels, pseudo_bboxes, pseudo_labels, num_classes,
                             a=(1, 1), pa=(3, 3)):
    """models/dense_heads/fcos_head_p2b_ts.py:657-708 with the shipped assigners (1,1) / (3,3)."""
    from . import hbb
    P = points.shape[0]
    gi, lb = topk_assign(points, cls_scores, gt_points, gt_labels, *a)
    labels = torch.full((P,), num_classes, dtype=torch.long)
    labels[gi != 0] = lb[gi != 0]
    gi2, lb2 = topk_assign(points, cls_scores, hbb.xyxy_to_cxcywh(pseudo_bboxes), pseudo_labels, *pa)
    labels_reg = torch.full((P,), num_classes, dtype=torch.long)
    labels_reg[gi2 != 0] = lb2[gi2 != 0]
    idx = torch.where(gi2 != 0, gi2 - 1, torch.zeros_like(gi2))
    b = pseudo_bboxes[idx]
    t = torch.stack([points[:, 0] - b[:, 0], points[:, 1] - b[:, 1], b[:, 2] - points[:, 0], b[:, 3] - points[:, 1]], -1)
    return labels_reg, t, labels, torch.ones(P)


def centerness_target(t):
    """models/dense_heads/fcos_head_p2b_ts.py:1019-1038."""
    lr, tb = t[:, [0, 2]], t[:, [1, 3]]
    return torch.sqrt((lr.min(-1)[0].clamp(min=0.01) / lr.max(-1)[0]) * (tb.min(-1)[0].clamp(min=0.01) / tb.max(-1)[0]))
