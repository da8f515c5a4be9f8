"""Coarse pseudo-box generation behind the FUSE assignment: the step that produces the
boxes the phase-2 MIL path refines.  ``generate_pseudo_single`` keeps the argument list and the return tuple of
``TS_P2BFCOSHead._gnerate_pseudo_single`` (HBB_TOD/mmdet/models/dense_heads/fcos_head_p2b_ts.py:736-794); the
reference's M x G one-hot matmuls, ``bincount`` and ``.tolist()`` set intersection become two small kernels and one
``nonzero`` (the returned ``valid_inds`` is a dynamically sized index list in the reference too)."""
import torch

from . import ops


def generate_pseudo_single(fuse_assigner, gt_points, gt_labels, gt_bboxes, cls_scores, bbox_preds, centernesses,
                           img_metas, img_list, filter_scores, points, num_points_per_lvl=None):
    """-> (pseudo_bboxes (G,4), pseudo_points (G,2), pseudo_labels (G,), mean IoU of the assigned pseudo boxes with
    their GT boxes, valid_inds (int64 indices of GTs that were assigned points AND score >= filter_scores))."""
    num_gts = len(gt_labels)
    if num_gts == 0:
        dev = gt_labels.device
        return torch.empty((0, 4), device=dev), torch.empty((0, 2), device=dev), torch.empty((0, 1), device=dev), 0.0, None
    pts = points.detach().float().contiguous()
    xyxy, cxcywh = ops.decode_ltrb(pts, bbox_preds.detach().float().contiguous())
    cls = cls_scores.detach().float().contiguous()
    res = fuse_assigner.assign(cxcywh, pts, cls, centernesses, gt_points, gt_labels, gt_bboxes_ignore=None)
    # the reference reads labels[p] = assigned label where gt_inds != 0, else class 0 (:755-759)
    labels = torch.where(res.gt_inds != 0, res.labels, torch.zeros_like(res.labels))
    boxes, ppts, scores, nums, valid, iou = ops.pseudo_aggregate(res.gt_inds, labels.contiguous(), cls, xyxy,
                                                                  gt_points.float().contiguous(),
                                                                  gt_bboxes.float().contiguous(), filter_scores)
    mean_iou = iou[0] / iou[1]                       # nan when nothing was assigned, like .mean() of an empty tensor
    valid_inds = valid.nonzero().reshape(-1)
    return boxes, ppts, gt_labels, mean_iou, valid_inds


def get_target_pseudo_single(assigner, pseudo_assigner, num_classes, gt_points, gt_labels, pseudo_points, pseudo_labels,
                             pseudo_bboxes, cls_scores, bbox_preds, centernesses, img_metas, img_list,
                             gt_augument_ignore, points, num_points_per_lvl=None, burn_in_step1=False):
    """``TS_P2BFCOSHead._get_target_pseudo_single`` (fcos_head_p2b_ts.py:657-708), the
    consumer of the refined boxes: classification labels from ``assigner`` on the GT points, regression labels and
    (l, t, r, b) targets from ``pseudo_assigner`` on the refined boxes.  Returns (labels_reg, bbox_targets, labels,
    weights) like the reference; background = ``num_classes``; background points are measured against box 0."""
    pts = points.detach().float().contiguous()
    cls = cls_scores.detach().float().contiguous()
    n = pts.size(0)
    res = assigner.assign(pts, cls, gt_points, gt_labels, gt_bboxes_ignore=None)
    labels = torch.where(res.gt_inds != 0, res.labels, torch.full_like(res.labels, num_classes))
    weights = torch.ones_like(labels).float()
    if len(pseudo_bboxes) == 0:
        return gt_labels.new_full((n,), num_classes), pseudo_bboxes.new_zeros((n, 4)), labels, weights
    pb = pseudo_bboxes.detach().float().contiguous()
    cxcywh = torch.cat([(pb[:, :2] + pb[:, 2:]) / 2, pb[:, 2:] - pb[:, :2]], 1)          # bbox_xyxy_to_cxcywh
    res2 = pseudo_assigner.assign(pts, cls, cxcywh, pseudo_labels, gt_bboxes_ignore=None)
    targets, labels_reg, _ = ops.ltrb_targets(pts, pb, res2.gt_inds, res2.labels, num_classes)
    return labels_reg, targets, labels, weights


def centerness_target(pos_bbox_targets):
    """fcos_head_p2b_ts.py:1019-1038 (elementwise; evaluated inside ``ops.ltrb_targets`` when asked for all points)."""
    t = pos_bbox_targets
    if t.shape[0] == 0:
        return t[:, 0]
    lr_min, lr_max = torch.minimum(t[:, 0], t[:, 2]), torch.maximum(t[:, 0], t[:, 2])
    tb_min, tb_max = torch.minimum(t[:, 1], t[:, 3]), torch.maximum(t[:, 1], t[:, 3])
    return torch.sqrt((lr_min.clamp(min=0.01) / lr_max) * (tb_min.clamp(min=0.01) / tb_max))
