"""Rotated twins of the proposal-bag functions, with the module-level surface of
OBB_TOD/mmrotate/models/detectors/syn_images_generator_v2.py:
  ``MIL_gen_proposals_from_cfg`` :26-40 (5-d boxes: bags are generated on the horizontal (cx,cy,w,h) box and carry
  the pseudo box's angle), ``gen_negative_proposals`` :129-156 (spans of 200 px, a fifth draw for theta, the sampled
  (x1,y1,x2,y2,theta) passed as-is into the (cx,cy,w,h,theta) slots -- reference quirk, reproduced),
  ``fine_proposals_from_cfg`` :159-220 (identical to the HBB function: xyxy in, xyxy out).
Same names, argument meaning, return structure and error behaviour; geometry in csrc/bag.cu (bit-exact)."""
import torch

from . import ops
from .proposals import (_split, boxes_to_rois, const_tensor, fine_proposals_from_cfg, img_wh_tensor,  # noqa: F401
                        sample_negative_boxes)


def MIL_gen_proposals_from_cfg(pseudo_points, pseudo_boxes_obb, fine_proposal_cfg, gt_boxes_obb, img_meta):
    gen_model = fine_proposal_cfg["gen_mode"]
    if gen_model != "refine":
        raise ValueError(f"gen_mode {gen_model!r}: only 'refine' is used by the Point Teacher configs")
    if fine_proposal_cfg["gen_proposal_mode"] != "fix_gen":
        raise ValueError(f"gen_proposal_mode {fine_proposal_cfg['gen_proposal_mode']!r} is not supported "
                         "(reference implements 'fix_gen')")
    boxes = [b.float() for b in pseudo_boxes_obb]
    dev = boxes[0].device
    rois = boxes_to_rois(boxes, dim=5)
    out, valid = ops.bag_gen(rois, img_wh_tensor(img_meta, dev), fine_proposal_cfg["base_ratios"],
                             fine_proposal_cfg["shake_ratio"], fine_proposal_cfg["min_scale"], rotated=True)
    U = out.shape[0] // rois.shape[0] if rois.shape[0] else 0
    sizes = [b.shape[0] * U for b in boxes]
    proposals_list = [p[:, 1:6] for p in _split(out, sizes)]
    proposals_valid_list = [v.bool().reshape(-1, 1) for v in _split(valid, sizes)]
    num_aug = int(proposals_list[0].shape[0] / pseudo_points[0].shape[0])     # from image 0, like the reference
    ref, real = [], []
    for i in range(len(pseudo_boxes_obb)):
        ref.append(pseudo_boxes_obb[i].unsqueeze(1).repeat(1, num_aug, 1).reshape(-1, 5))
        real.append(gt_boxes_obb[i].unsqueeze(1).repeat(1, num_aug, 1).reshape(-1, 5))
    return proposals_list, proposals_valid_list, ref, real


def gen_negative_proposals(gt_points, proposal_cfg, aug_generate_proposals, img_meta, neg_boxes=None):
    """``neg_boxes`` (optional list of (n,5)) injects the sampled boxes; default draws them on the host generator
    like the reference.  Returns (list of (n,5), list of (n,) bool)."""
    num_neg_gen = proposal_cfg["gen_num_neg"]
    if num_neg_gen == 0:
        return None, None
    dev = gt_points[0].device
    if neg_boxes is None:
        neg_boxes = [sample_negative_boxes(num_neg_gen, img_meta[i]["img_shape"], rotated=True).to(dev)
                     for i in range(len(gt_points))]
    neg_rois = boxes_to_rois([b.float() for b in neg_boxes], dim=5)
    bag_rois = boxes_to_rois([b.float() for b in aug_generate_proposals], dim=5)
    offs = [0]
    for p in aug_generate_proposals:
        offs.append(offs[-1] + p.shape[0])
    w = ops.neg_weight(neg_rois, bag_rois, const_tensor(offs, torch.int32, dev), rotated=True).bool()
    return list(neg_boxes), _split(w, [b.shape[0] for b in neg_boxes])
