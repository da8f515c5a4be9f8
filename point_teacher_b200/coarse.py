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
