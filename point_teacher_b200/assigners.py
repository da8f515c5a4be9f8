"""Dense-head label assignment behind the reference's registry surface.

Mirrors, with identical constructor keywords, call signatures and result fields:
  AssignResult        HBB_TOD/mmdet/core/bbox/assigners/assign_result.py (fields num_gts, gt_inds, max_overlaps, labels)
  TopkAssigner        HBB_TOD/mmdet/core/bbox/assigners/topk_assigner.py:12-147
  FUSETopkAssigner    HBB_TOD/mmdet/core/bbox/assigners/fuse_topk_assigner.py:12-121
  MaxIoUAssigner      HBB_TOD/mmdet/core/bbox/assigners/max_iou_assigner.py:10-212
  BboxOverlaps2D      HBB_TOD/mmdet/core/bbox/iou_calculators/iou2d_calculator.py:9-71
  BboxDistanceMetric  HBB_TOD/mmdet/core/bbox/iou_calculators/metric_calculator.py:7-41
  FocalLossCost / PointCost / InsiderCost   HBB_TOD/mmdet/core/bbox/match_costs/match_cost.py:54-100, 188-252

All arithmetic runs in csrc/assign.cu; the P x G cost matrices and the G x A overlap matrix of the reference are
never materialised (the calculators still return full matrices when called directly, as the reference's do)."""
import torch

from . import ops
from .registry import BBOX_ASSIGNERS, IOU_CALCULATORS, MATCH_COST, build_iou_calculator, build_match_cost


class AssignResult:
    def __init__(self, num_gts, gt_inds, max_overlaps, labels=None):
        self.num_gts, self.gt_inds, self.max_overlaps, self.labels = num_gts, gt_inds, max_overlaps, labels
        self._extra_properties = {}

    @property
    def num_preds(self):
        return len(self.gt_inds)

    def __repr__(self):
        return f"<AssignResult(num_gts={self.num_gts}, gt_inds.shape={tuple(self.gt_inds.shape)})>"


# ------------------------------------------------------------------------------ match costs (config holders + callables)
@MATCH_COST.register_module(name="FocalLossCost", force=True)
class FocalLossCost:
    def __init__(self, weight=1.0, alpha=0.25, gamma=2, eps=1e-12):
        self.weight, self.alpha, self.gamma, self.eps = weight, alpha, gamma, eps

    def table(self, cls_pred):
        return ops.focal_cost_table(cls_pred.float().contiguous(), self.alpha, self.gamma, self.eps, self.weight)

    def __call__(self, cls_pred, gt_labels):
        return self.table(cls_pred)[:, gt_labels]


@MATCH_COST.register_module(name="PointCost", force=True)
class PointCost:
    def __init__(self, mode="L1", weight=1.0):
        assert mode in ["L1", "L2"]
        self.mode, self.weight = mode, weight


@MATCH_COST.register_module(name="InsiderCost", force=True)
class InsiderCost:
    def __init__(self, weight=1.0):
        self.weight = weight


def _empty_result(bbox_pred, gt_bboxes):
    n = bbox_pred.size(0)
    gt_inds = bbox_pred.new_full((n,), -1, dtype=torch.long)
    labels = bbox_pred.new_full((n,), -1, dtype=torch.long)
    if gt_bboxes is None:
        gt_inds[:] = 0
        return AssignResult(0, gt_inds, None, labels=labels)
    if gt_bboxes.size(0) == 0:
        gt_inds[:] = 0
    return AssignResult(gt_bboxes.size(0), gt_inds, None, labels=labels)


class _TwoStageTopk:
    def _assign(self, stage1_points, cls_pred, gt_bboxes, gt_labels, pred_boxes):
        P = stage1_points.size(0)
        gts = gt_bboxes.float().contiguous()
        pre = ops.topk_pre(stage1_points.float().contiguous(), gts, self.num_pre, self.reg_cost.mode,
                           self.reg_cost.weight)
        table = self.cls_cost.table(cls_pred) if self.num_pre > self.topk else None
        gt_inds, labels = ops.topk_second(pre, self.topk, P, table, gt_labels.long().contiguous(), gts,
                                          pred_boxes, getattr(getattr(self, "location_cost", None), "weight", 1.0))
        return AssignResult(gts.size(0), gt_inds, None, labels=labels)


@BBOX_ASSIGNERS.register_module(name="TopkAssigner", force=True)
class TopkAssigner(_TwoStageTopk):
    def __init__(self, num_pre, topk, cls_cost=dict(type="ClassificationCost", weight=1.), reg_cost=None,
                 iou_cost=None):
        self.num_pre, self.topk = num_pre, topk
        self.cls_cost = build_match_cost(cls_cost)
        self.reg_cost = build_match_cost(reg_cost)
        if not isinstance(self.cls_cost, FocalLossCost) or not isinstance(self.reg_cost, PointCost):
            raise NotImplementedError("the Point Teacher configs use FocalLossCost + PointCost")

    def assign(self, bbox_pred, cls_pred, gt_bboxes, gt_labels, gt_bboxes_ignore=None, eps=1e-7):
        assert gt_bboxes_ignore is None, "Only case when gt_bboxes_ignore is None is supported."
        if gt_bboxes is None or gt_bboxes.size(0) == 0 or bbox_pred.size(0) == 0:
            return _empty_result(bbox_pred, gt_bboxes)
        return self._assign(bbox_pred, cls_pred, gt_bboxes, gt_labels, None)


@BBOX_ASSIGNERS.register_module(name="FUSETopkAssigner", force=True)
class FUSETopkAssigner(_TwoStageTopk):
    def __init__(self, num_pre, topk, cls_cost=dict(type="ClassificationCost", weight=1.), reg_cost=None,
                 location_cost=None, center_cost=None, iou_cost=None):
        self.num_pre, self.topk = num_pre, topk
        self.cls_cost = build_match_cost(cls_cost)
        self.reg_cost = build_match_cost(reg_cost)
        self.location_cost = build_match_cost(location_cost)
        if not isinstance(self.cls_cost, FocalLossCost) or not isinstance(self.reg_cost, PointCost) or \
                not isinstance(self.location_cost, InsiderCost):
            raise NotImplementedError("the Point Teacher configs use FocalLossCost + PointCost + InsiderCost")

    def assign(self, bbox_pred, points, cls_pred, centerness, gt_bboxes, gt_labels, gt_bboxes_ignore=None, eps=1e-7):
        assert gt_bboxes_ignore is None, "Only case when gt_bboxes_ignore is None is supported."
        if gt_bboxes is None or gt_bboxes.size(0) == 0 or bbox_pred.size(0) == 0:
            return _empty_result(bbox_pred, gt_bboxes)
        return self._assign(points, cls_pred, gt_bboxes, gt_labels, bbox_pred.float().contiguous())


# ------------------------------------------------------------------------------ IoU calculators
@IOU_CALCULATORS.register_module(name="BboxOverlaps2D", force=True)
class BboxOverlaps2D:
    calc = 0

    def __init__(self, scale=1., dtype=None):
        self.scale, self.dtype = scale, dtype

    def __call__(self, bboxes1, bboxes2, mode="iou", is_aligned=False):
        assert bboxes1.size(-1) in [0, 4, 5] and bboxes2.size(-1) in [0, 4, 5]
        b1 = bboxes1[..., :4].float().contiguous()
        b2 = bboxes2[..., :4].float().contiguous()
        return ops.bbox_overlaps(b1, b2, mode, is_aligned)


@IOU_CALCULATORS.register_module(name="BboxDistanceMetric", force=True)
class BboxDistanceMetric:
    calc = 1

    def __call__(self, bboxes1, bboxes2, mode="iou", is_aligned=False):
        assert bboxes1.size(-1) in [0, 4, 5] and bboxes2.size(-1) in [0, 4, 5]
        b1 = bboxes1[..., :4].float().contiguous()
        b2 = bboxes2[..., :4].float().contiguous()
        return ops.bbox_metric(b1, b2, mode, calc=1)          # the reference ignores is_aligned as well


@BBOX_ASSIGNERS.register_module(name="MaxIoUAssigner", force=True)
class MaxIoUAssigner:
    def __init__(self, pos_iou_thr, neg_iou_thr, min_pos_iou=.0, gt_max_assign_all=True, ignore_iof_thr=-1,
                 ignore_wrt_candidates=True, match_low_quality=True, gpu_assign_thr=-1,
                 iou_calculator=dict(type="BboxOverlaps2D"), assign_metric="iou"):
        self.pos_iou_thr, self.neg_iou_thr, self.min_pos_iou = pos_iou_thr, neg_iou_thr, min_pos_iou
        self.gt_max_assign_all, self.ignore_iof_thr = gt_max_assign_all, ignore_iof_thr
        self.ignore_wrt_candidates, self.gpu_assign_thr = ignore_wrt_candidates, gpu_assign_thr
        self.match_low_quality = match_low_quality
        self.iou_calculator = build_iou_calculator(iou_calculator)
        self.assign_metric = assign_metric

    def assign(self, bboxes, gt_bboxes, gt_bboxes_ignore=None, gt_labels=None, mode=None):
        """``gpu_assign_thr`` is accepted and ignored: nothing is moved to the CPU, the G x A matrix is never built."""
        if self.ignore_iof_thr > 0 and gt_bboxes_ignore is not None and gt_bboxes_ignore.numel() > 0:
            raise NotImplementedError("ignore regions are not on the Point Teacher path")
        mode = mode or self.assign_metric
        G, A = gt_bboxes.size(0), bboxes.size(0)
        if G == 0 or A == 0:
            gt_inds = bboxes.new_full((A,), -1, dtype=torch.long)
            if G == 0:
                gt_inds[:] = 0
            labels = None if gt_labels is None else bboxes.new_full((A,), -1, dtype=torch.long)
            return AssignResult(G, gt_inds, bboxes.new_zeros((A,)), labels=labels)
        gi, mx, lb = ops.max_iou_assign(gt_bboxes[:, :4].float().contiguous(), bboxes[:, :4].float().contiguous(),
                                        self.iou_calculator.calc, mode, self.pos_iou_thr, self.neg_iou_thr,
                                        self.min_pos_iou, self.gt_max_assign_all, self.match_low_quality,
                                        None if gt_labels is None else gt_labels.long().contiguous())
        return AssignResult(G, gi, mx, labels=lb)
