"""Proposal-bag construction with the reference's module-level function surface.

Mirrors (HBB_TOD/mmdet/models/detectors/syn_images_generator_v2.py):
  ``fine_proposals_from_cfg`` :262-324, ``MIL_gen_proposals_from_cfg`` :134-145,
  ``gen_negative_proposals`` :234-259 -- same names, argument meaning, return structure and error
  behaviour (``gen_num_neg == 0`` -> ``(None, None)``, zero GTs in image 0 -> ZeroDivisionError).
The geometry runs in csrc/bag.cu and is bit-exact with the reference's fp32 arithmetic.
"""
import torch

from . import ops


_CONST = {}


def const_tensor(values, dtype, device):
    """Small host-side metadata (image sizes, per-image offsets) as a device tensor, uploaded once per distinct
    value: steady-state steps do no H2D copy for it, which keeps the step CUDA-graph capturable."""
    key = (tuple(map(tuple, values)) if values and isinstance(values[0], (list, tuple)) else tuple(values),
           dtype, str(device))
    t = _CONST.get(key)
    if t is None:
        if len(_CONST) > 4096:
            _CONST.clear()
        t = torch.tensor(values, dtype=dtype, device=device)
        _CONST[key] = t
    return t


def img_wh_tensor(img_meta, device):
    """(B,2) float tensor of (w, h) from ``img_meta[i]['img_shape'] = (h, w, c)``."""
    return const_tensor([[float(m["img_shape"][1]), float(m["img_shape"][0])] for m in img_meta],
                        torch.float32, device)


def boxes_to_rois(box_list, dim=None):
    """``bbox2roi`` (core/bbox/transforms.py:58-78): list of (n_i, >=4) xyxy -> (sum n_i, 5); ``rbbox2roi``
    (OBB_TOD/mmrotate/core/bbox/transforms.py:73-92) for 5-column (cx,cy,w,h,theta) boxes -> (sum n_i, 6).
    ``dim`` forces the number of box columns copied (default: 5 when the boxes have exactly 5 columns, else 4)."""
    n = sum(b.size(0) for b in box_list)
    ref = box_list[0]
    d = dim if dim is not None else (5 if ref.size(-1) == 5 else 4)
    if n == 0:
        return ref.new_empty((0, d + 1))
    idx = const_tensor([i for i, b in enumerate(box_list) for _ in range(int(b.size(0)))], torch.int32, ref.device)
    boxes = torch.cat([b[:, :d] for b in box_list]).float().contiguous()
    return ops.make_rois(boxes, idx)


def _split(t, sizes):
    return list(torch.split(t, sizes))


def fine_proposals_from_cfg(pseudo_boxes, fine_proposal_cfg, img_meta):
    gen_mode = fine_proposal_cfg["gen_proposal_mode"]
    if gen_mode != "fix_gen":
        # the reference silently returns nothing usable for other modes; fail loudly instead
        raise ValueError(f"gen_proposal_mode {gen_mode!r} is not supported (reference implements 'fix_gen')")
    n_img = len(img_meta)
    boxes = [pseudo_boxes[i] for i in range(n_img)]
    dev = boxes[0].device
    rois = boxes_to_rois([b.float() for b in boxes])
    out, valid = ops.bag_gen(rois, img_wh_tensor(img_meta, dev), fine_proposal_cfg["base_ratios"],
                             fine_proposal_cfg["shake_ratio"], fine_proposal_cfg["min_scale"])
    U = out.shape[0] // max(rois.shape[0], 1) if rois.shape[0] else 0
    sizes = [b.shape[0] * U for b in boxes]
    proposal_list = [p[:, 1:5] for p in _split(out, sizes)]
    valid_list = [v.bool().reshape(-1, 1) for v in _split(valid, sizes)]
    return proposal_list, valid_list


def MIL_gen_proposals_from_cfg(pseudo_points, pseudo_boxes, fine_proposal_cfg, gt_boxes, img_meta):
    gen_model = fine_proposal_cfg["gen_mode"]
    if gen_model != "refine":
        raise ValueError(f"gen_mode {gen_model!r}: only 'refine' is used by the Point Teacher configs")
    proposals_list, proposals_valid_list = fine_proposals_from_cfg(pseudo_boxes, fine_proposal_cfg, img_meta)
    num_aug = int(proposals_list[0].shape[0] / pseudo_points[0].shape[0])  # from image 0, like the reference
    ref, real = [], []
    for i in range(len(pseudo_boxes)):
        ref.append(pseudo_boxes[i].unsqueeze(1).repeat(1, num_aug, 1).reshape(-1, pseudo_boxes[i].shape[-1]))
        real.append(gt_boxes[i].unsqueeze(1).repeat(1, num_aug, 1).reshape(-1, gt_boxes[i].shape[-1]))
    return proposals_list, proposals_valid_list, ref, real


def sample_negative_boxes(num, img_shape, generator=None, rotated=False):
    """The reference's four CPU ``torch.rand`` draws (:247-250).  Kept on the host generator so a run
    with the same CPU seed sees the same negatives as the reference.  rotated: the OBB variant
    (OBB_TOD/.../syn_images_generator_v2.py:142-148): spans of 200 px and a fifth draw for theta."""
    import math
    h, w, _ = img_shape
    span = 200 if rotated else 100
    x1 = torch.rand(num, generator=generator) * w * 0.8
    y1 = torch.rand(num, generator=generator) * h * 0.8
    x2 = x1 + torch.rand(num, generator=generator) * span
    y2 = y1 + torch.rand(num, generator=generator) * span
    if rotated:
        th = torch.rand(num, generator=generator) * math.pi - math.pi / 2
        return torch.stack([x1, y1, x2, y2, th], dim=1)
    return torch.stack([x1, y1, x2, y2], dim=1)


def gen_negative_proposals(gt_points, proposal_cfg, aug_generate_proposals, img_meta, neg_boxes=None):
    """``neg_boxes`` (optional list of (n,4)) injects the sampled boxes; default draws them like the
    reference.  Returns (list of (n,4), list of (n,) bool)."""
    num_neg_gen = proposal_cfg["gen_num_neg"]
    if num_neg_gen == 0:
        return None, None
    dev = gt_points[0].device
    if neg_boxes is None:
        neg_boxes = [sample_negative_boxes(num_neg_gen, img_meta[i]["img_shape"]).to(dev)
                     for i in range(len(gt_points))]
    neg_rois = boxes_to_rois(neg_boxes)
    bag_rois = boxes_to_rois(aug_generate_proposals)
    offs = [0]
    for p in aug_generate_proposals:
        offs.append(offs[-1] + p.shape[0])
    offsets = const_tensor(offs, torch.int32, dev)
    w = ops.neg_weight(neg_rois, bag_rois, offsets).bool()
    sizes = [b.shape[0] for b in neg_boxes]
    return list(neg_boxes), _split(w, sizes)
