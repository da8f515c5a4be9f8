"""Phase-2 pseudo-box refinement at the detector level.

Mirrors ``TS_P2B_FCOS.forward_mil_head_burn_in_step2``
(HBB_TOD/mmdet/models/detectors/fcos_p2b_teacher_student.py:425-466): cap at
``num_training_burninstep2`` GTs per image, per stage generate base bags + negatives, run the
MIL head, feed the refined boxes to the next stage, write back into full-length clones and
derive the refined points.  ``P2BRefineMixin`` carries the method for a detector that has
``self.student.bbox_head``; :func:`phase2_refine` is the same code for a bare head."""
import torch

from . import ops
from .proposals import MIL_gen_proposals_from_cfg, gen_negative_proposals


def phase2_refine(head, x_ori, img_metas, pseudo_bboxes, pseudo_points, pseudo_labels, gt_bboxes,
                  fine_proposal_cfg, fine_proposal_extensive_cfg, num_stages=1, num_training_burninstep2=100,
                  alpha=(0.01, 0.25), neg_boxes=None):
    """Returns (refined_pseudo_bboxes, refined_pseudo_points, losses) like the reference method.
    ``neg_boxes[stage][img]`` optionally injects the negative boxes the reference samples on the CPU."""
    cap = num_training_burninstep2
    num_img = len(pseudo_bboxes)
    pb = [b[:cap, :].clone().float() for b in pseudo_bboxes]
    gb = [b[:cap, :].clone().float() for b in gt_bboxes]
    pp = [p[:cap, :].clone().float() for p in pseudo_points]
    pl = [l[:cap].clone() for l in pseudo_labels]
    refined_b = [b.clone() for b in pseudo_bboxes]
    refined_p = [p.clone() for p in pseudo_points]
    losses = {"coarse_bboxes_iou": ops.aligned_iou_mean(torch.cat(pb).contiguous(), torch.cat(gb).contiguous())}
    gcat = torch.cat(gb).contiguous()
    for stage in range(num_stages):
        props, valids, refs, reals = MIL_gen_proposals_from_cfg(pp, pb, fine_proposal_cfg[stage], gb, img_metas)
        negs, neg_w = gen_negative_proposals(pp, fine_proposal_cfg[stage], props, img_metas,
                                             None if neg_boxes is None else neg_boxes[stage])
        mil_loss, merged = head.MIL_head_burn_in_step2(x_ori, img_metas, props, valids, refs, reals, negs, neg_w,
                                                       pb, pl, fine_proposal_extensive_cfg[stage], stage,
                                                       loss_scales=alpha)
        pb = list(merged)
        losses[f"stage{stage}_refine_bboxes_iou"] = ops.aligned_iou_mean(torch.cat(pb).contiguous(), gcat)
        losses.update(mil_loss)
    pts = head.last_results["_b200"]["merged_points"]
    for i, (b, p) in enumerate(zip(pb, torch.split(pts, [len(b) for b in pb]))):
        refined_b[i][:cap, :] = b
        refined_p[i][:cap, :] = p
    return refined_b, refined_p, losses


class P2BRefineMixin:
    """Drop-in for the detector method (same name and argument list)."""

    def forward_mil_head_burn_in_step2(self, num_img, pseudo_bboxes, pseudo_points, pseudo_labels, gt_bboxes,
                                       img_metas, x_ori):
        return phase2_refine(self.student.bbox_head, x_ori, img_metas, pseudo_bboxes, pseudo_points,
                             pseudo_labels, gt_bboxes, self.fine_proposal_cfg, self.fine_proposal_extensive_cfg,
                             self.num_stages, self.num_training_burninstep2, self.alpha)
