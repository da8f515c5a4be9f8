"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the DETECTOR-level callers of the MIL head, i.e. the code that sits
directly above the drop-in boundary, so that the B200 classes can be driven exactly the way the reference drives them:

  HBB  TS_P2B_FCOS.forward_mil_head_burn_in_step1   HBB_TOD/mmdet/models/detectors/fcos_p2b_teacher_student.py:365-424
       TS_P2B_FCOS.forward_mil_head_burn_in_step2   .../fcos_p2b_teacher_student.py:425-466
  OBB  RotatedFCOS_TS.forward_mil_head_burn_in_step1  OBB_TOD/mmrotate/models/detectors/rotated_fcos_teacher_student.py:435-492
       RotatedFCOS_TS.forward_mil_head_burn_in_step2  .../rotated_fcos_teacher_student.py:494-535

The two drivers below are literal: the same slicing / cloning, the same order of by-name function calls
(``MIL_gen_proposals_from_cfg`` for the real bags, then for the synthetic bags, then ``gen_negative_proposals`` --
the order matters because the negatives consume the global CPU RNG), the same ``head.MIL_head_burn_in_step{1,2}``
call with the reference's positional argument list, the same ``alpha`` scaling and write-back.  What they call is a
*backend*:  ``det`` (an object with the detector attributes ``student.bbox_head``, ``num_stages``,
``fine_proposal_cfg``, ``fine_proposal_extensive_cfg``, ``alpha``, ``num_training_burninstep1/2``) and ``fns``
(the by-name module functions + box helpers).  Backends:
  * ``oracle_backend``     oracle/hbb.py / oracle/obb.py behind the reference's method surface (``OracleHead``)
  * ``reference_backend``  the reference's OWN functions and head under oracle/ref_shim.py (only where /root/reference
                           is mounted) -- ``python -m oracle.check_oracle_vs_ref --detector`` runs the reference's own
                           detector methods against these drivers (boxes tol 0, losses 1e-6, gradients 1e-5)
  * tests/test_gpu_dropin.py builds the B200 backend (point_teacher_b200 classes and functions) and runs the SAME
    drivers on the GPU.
Pinned: yes for HBB (every function on the path comes from the reference's files); OBB inherits the "parity unpinned"
mmcv kernels (RoIAlignRotated, box_iou_rotated) stated in oracle/rotated.py."""
import types

import torch

from . import hbb, obb


# ----------------------------------------------------------------------------------------------- literal drivers
def forward_mil_head_burn_in_step2(det, fns, num_img, pseudo_bboxes, pseudo_points, pseudo_labels, gt_bboxes,
                                   img_metas, x_ori):
    """fcos_p2b_teacher_student.py:425-466 / rotated_fcos_teacher_student.py:494-535."""
    cap = det.num_training_burninstep2
    losses = {}
    pb, gb, pp, pl, out_b, out_p = [], [], [], [], [], []
    for i in range(num_img):
        pb.append(pseudo_bboxes[i][:cap, :].clone())
        gb.append(gt_bboxes[i][:cap, :].clone())
        pp.append(pseudo_points[i][:cap, :].clone())
        pl.append(pseudo_labels[i][:cap].clone())
        out_b.append(pseudo_bboxes[i].clone())
        out_p.append(pseudo_points[i].clone())
    losses["coarse_bboxes_iou"] = fns.aligned_iou_mean(torch.cat(pb, 0), torch.cat(gb, 0))
    head = det.student.bbox_head
    for stage in range(det.num_stages):
        props, valids, refs, reals = fns.MIL_gen_proposals_from_cfg(pp, pb, det.fine_proposal_cfg[stage], gb, img_metas)
        negs, neg_w = fns.gen_negative_proposals(pp, det.fine_proposal_cfg[stage], props, img_metas)
        mil_loss, pb = head.MIL_head_burn_in_step2(x_ori, img_metas, props, valids, refs, reals, negs, neg_w, pb, pl,
                                                   det.fine_proposal_extensive_cfg[stage], stage)
        losses[f"stage{stage}_refine_bboxes_iou"] = fns.aligned_iou_mean(torch.cat(list(pb), 0), torch.cat(gb, 0))
        mil_loss[f"stage{stage}_loss_mil_bbox"] = mil_loss[f"stage{stage}_loss_mil_bbox"] * det.alpha[0]
        mil_loss[f"stage{stage}_loss_mil_bags"] = mil_loss[f"stage{stage}_loss_mil_bags"] * det.alpha[1]
        losses.update(mil_loss)
    for i in range(num_img):
        out_b[i][:cap, :] = pb[i]
        out_p[i][:cap, :] = fns.box_points(pb[i])
    return out_b, out_p, losses


def forward_mil_head_burn_in_step1(det, fns, num_img, synthetic_bboxes, pseudo_bboxes, pseudo_points, pseudo_labels,
                                   gt_bboxes, img_metas, x_synthetic, x_ori):
    """fcos_p2b_teacher_student.py:365-424 / rotated_fcos_teacher_student.py:435-492."""
    cap = det.num_training_burninstep1
    losses = {}
    if len([i for i in range(num_img) if synthetic_bboxes[i].shape[0] != 0]) != num_img:
        d = fns.box_dim
        return ([pseudo_bboxes[0].new_empty((0, d)) for _ in range(num_img)],
                [pseudo_bboxes[0].new_empty((0, 2)) for _ in range(num_img)], None)
    sb, sp, pb, gb, pp, pl, out_b, out_p = [], [], [], [], [], [], [], []
    for i in range(num_img):
        sb.append(synthetic_bboxes[i][:cap, :])
        sp.append(fns.box_points(sb[-1]))
        pb.append(pseudo_bboxes[i][:cap, :])
        gb.append(gt_bboxes[i][:cap, :])
        pp.append(pseudo_points[i][:cap, :])
        pl.append(pseudo_labels[i][:cap])
        out_b.append(pseudo_bboxes[i].clone())
        out_p.append(pseudo_points[i].clone())
    losses["coarse_bboxes_iou"] = fns.aligned_iou_mean(torch.cat(pb, 0), torch.cat(gb, 0))
    head = det.student.bbox_head
    for stage in range(det.num_stages):
        cfg = det.fine_proposal_cfg[stage]
        props, valids, refs, reals = fns.MIL_gen_proposals_from_cfg(pp, pb, cfg, gb, img_metas)
        s_props, s_valids, s_refs, s_reals = fns.MIL_gen_proposals_from_cfg(sp, sb, cfg, sb, img_metas)
        negs, neg_w = fns.gen_negative_proposals(pp, cfg, props, img_metas)
        mil_loss, pb = head.MIL_head_burn_in_step1(x_ori, x_synthetic, img_metas, props, valids, refs, reals, s_props,
                                                   s_valids, s_refs, s_reals, negs, neg_w, sb, pb, pl,
                                                   det.fine_proposal_extensive_cfg[stage], stage)
        losses[f"stage{stage}_refine_bboxes_iou"] = fns.aligned_iou_mean(torch.cat(list(pb), 0), torch.cat(gb, 0))
        mil_loss[f"stage{stage}_loss_mil_bbox"] = mil_loss[f"stage{stage}_loss_mil_bbox"] * det.alpha[0]
        mil_loss[f"stage{stage}_loss_mil_bags"] = mil_loss[f"stage{stage}_loss_mil_bags"] * det.alpha[1]
        losses.update(mil_loss)
    for i in range(num_img):
        out_b[i][:cap, :] = pb[i]
        out_p[i][:cap, :] = fns.box_points(pb[i])
    return out_b, out_p, losses


def parse_losses(losses):
    """mmdet BaseDetector._parse_losses (HBB_TOD/mmdet/models/detectors/base.py): the scalar that is back-propagated
    is the sum of every entry whose key contains 'loss'."""
    return sum(v.mean() for k, v in losses.items() if "loss" in k)


def make_detector(head, fine_cfg, ext_cfg, num_stages=1, alpha=(0.01, 0.25), cap1=100, cap2=100):
    return types.SimpleNamespace(student=types.SimpleNamespace(bbox_head=head), num_stages=num_stages,
                                 fine_proposal_cfg=fine_cfg, fine_proposal_extensive_cfg=ext_cfg, alpha=list(alpha),
                                 num_training_burninstep1=cap1, num_training_burninstep2=cap2)


# ----------------------------------------------------------------------------------------------- oracle backend
class OracleHead:
    """oracle/hbb.py / oracle/obb.py behind ``TS_P2BFCOSHead`` / ``TS_P2RBRotatedFCOSHead``'s two entry methods
    (fcos_head_p2b_ts.py:1279-1344, rotated_fcos_head_p2rb_ts.py:1384-1453)."""

    def __init__(self, P, strides, topk, beta=0.25, rotated=False):
        self.P, self.strides, self.topk, self.beta, self.rotated = P, strides, topk, beta, rotated
        self.m = obb if rotated else hbb
        self.last = None

    def _fwd(self, num_gt, per_img, x, props, valids, refs, reals, img_metas, cfg, stage, negs=None):
        return self.m.forward_mil_head(self.P, num_gt, per_img, x, self.strides, props, valids, refs, reals, img_metas,
                                       cfg, stage, negs)

    def MIL_head_burn_in_step2(self, x, img_metas, props, valids, refs, reals, negs, neg_w, pseudo_bboxes,
                               pseudo_labels, cfg, stage):
        num_gt = torch.cat(list(pseudo_bboxes)).shape[0]
        per_img = [b.shape[0] for b in pseudo_bboxes]
        R = self._fwd(num_gt, per_img, x, props, valids, refs, reals, img_metas, cfg, stage, negs)
        losses = {f"stage{stage}_loss_mil_bbox": R["loss_mil_bbox"],
                  f"stage{stage}_loss_mil_bags": self.m.mil_bag_training(R, pseudo_labels, neg_w),
                  f"stage{stage}_coarse_bags_iou": R["coarse_bags_iou"],
                  f"stage{stage}_refine_bags_iou": R["refine_bags_iou"]}
        merged, idx, sc = self.m.mil_bag_selection(R, img_metas, list(pseudo_bboxes), pseudo_labels, self.topk, self.beta)
        R["selected_idx"], R["selected_scores"] = idx, sc
        self.last = R
        return losses, merged

    def MIL_head_burn_in_step1(self, x_ori, x_syn, img_metas, props, valids, refs, reals, s_props, s_valids, s_refs,
                               s_reals, negs, neg_w, synthetic_bboxes, pseudo_bboxes, pseudo_labels, cfg, stage):
        n_syn = torch.cat(list(synthetic_bboxes)).shape[0]
        syn_per_img = [b.shape[0] for b in synthetic_bboxes]
        num_gt = torch.cat(list(pseudo_bboxes)).shape[0]
        per_img = [b.shape[0] for b in pseudo_bboxes]
        losses = {}
        S = self._fwd(n_syn, syn_per_img, x_syn, s_props, s_valids, s_refs, s_reals, img_metas, cfg, stage, None)
        losses[f"stage{stage}_loss_mil_bbox"] = S["loss_mil_bbox"]
        R = self._fwd(num_gt, per_img, x_ori, props, valids, refs, reals, img_metas, cfg, stage, negs)
        losses[f"stage{stage}_loss_mil_bags"] = self.m.mil_bag_training(R, pseudo_labels, neg_w)
        losses[f"stage{stage}_coarse_bags_iou"] = R["coarse_bags_iou"]
        losses[f"stage{stage}_refine_bags_iou"] = R["refine_bags_iou"]
        merged, idx, sc = self.m.mil_bag_selection(R, img_metas, list(pseudo_bboxes), pseudo_labels, self.topk, self.beta)
        R["selected_idx"], R["selected_scores"] = idx, sc
        self.last = R
        return losses, merged


def oracle_backend(rotated=False):
    if rotated:
        return types.SimpleNamespace(
            MIL_gen_proposals_from_cfg=obb.mil_gen_proposals, gen_negative_proposals=obb.gen_negative_proposals,
            aligned_iou_mean=lambda a, b: obb.rbbox_overlaps(a, b, is_aligned=True).mean(),
            box_points=lambda b: b[:, :2], box_dim=5)
    return types.SimpleNamespace(
        MIL_gen_proposals_from_cfg=hbb.mil_gen_proposals, gen_negative_proposals=hbb.gen_negative_proposals,
        aligned_iou_mean=lambda a, b: hbb.bbox_overlaps(a, b, is_aligned=True).mean(),
        box_points=lambda b: hbb.xyxy_to_cxcywh(b)[:, :2], box_dim=4)


def reference_backend(rotated=False):
    """The reference's own by-name functions under the import shim (needs /root/reference)."""
    from . import ref_shim
    if rotated:
        o = ref_shim.install_obb()
        return types.SimpleNamespace(
            MIL_gen_proposals_from_cfg=o.syn.MIL_gen_proposals_from_cfg,
            gen_negative_proposals=o.syn.gen_negative_proposals,
            aligned_iou_mean=lambda a, b: o.rbbox_overlaps(a, b, mode="iou", is_aligned=True).mean(),
            box_points=lambda b: b[:, :2], box_dim=5)
    ns = ref_shim.install()
    return types.SimpleNamespace(
        MIL_gen_proposals_from_cfg=ns.syn.MIL_gen_proposals_from_cfg,
        gen_negative_proposals=ns.syn.gen_negative_proposals,
        aligned_iou_mean=lambda a, b: ns.bbox_overlaps(a, b, mode="iou", is_aligned=True).mean(),
        box_points=lambda b: ns.transforms.bbox_xyxy_to_cxcywh(b)[:, :2], box_dim=4)


# ----------------------------------------------------------------------------------------------- synthetic inputs
def synthetic_boxes_like(d, seed, rotated=False, n_range=(5, 9)):
    """Phase-1 "synthetic" boxes (exact boxes of the pasted regions, fcos_p2b_teacher_student.py:486-492): a second
    seeded draw of boxes with the batch's image size, plus a second feature map standing in for the student's
    features of the synthetic image."""
    import math
    from point_teacher_b200 import synth
    g = torch.Generator().manual_seed(7000 + seed)
    h, w, _ = d["img_metas"][0]["img_shape"]
    boxes = []
    for _ in d["gt_boxes"]:
        n = int(torch.randint(n_range[0], n_range[1] + 1, (1,), generator=g))
        xyxy = synth.make_boxes(g, n, (h, w))
        if rotated:
            th = torch.rand(n, generator=g) * math.pi - math.pi / 2
            boxes.append(torch.stack([(xyxy[:, 0] + xyxy[:, 2]) / 2, (xyxy[:, 1] + xyxy[:, 3]) / 2,
                                      xyxy[:, 2] - xyxy[:, 0], xyxy[:, 3] - xyxy[:, 1], th], 1))
        else:
            boxes.append(xyxy)
    feat_syn = torch.randn(d["feat"].shape, generator=g)
    return boxes, feat_syn
