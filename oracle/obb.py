"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the OBB (rotated) twin of the phase-2 MIL path,
SURVEY.md section 8 rows a2/a3/a4/a6-a12 for OBB_TOD.  Paths relative to /root/reference/OBB_TOD/mmrotate/.

Pinned by ``oracle/check_oracle_vs_ref.py --obb`` against the reference's own
``rotated_fcos_head_p2rb_ts.py`` / ``syn_images_generator_v2.py`` under the import shim.  The three mmcv
kernels underneath (RoIAlignRotated, box_iou_rotated, nms_rotated) are un-vendored: both the shim and this
file bind them to oracle/rotated.py (PARITY UNPINNED for those kernels; cross-checks in tests/test_oracle.py).
"""
import math

import torch
import torch.nn.functional as F

from . import hbb, rotated


def rbbox_overlaps(b1, b2, mode="iou", is_aligned=False):
    """core/bbox/iou_calculators/rotate_iou2d_calculator.py:53-89 (clamps w,h >= 1e-3 first)."""
    rows, cols = b1.shape[0], b2.shape[0]
    if rows * cols == 0:
        return b1.new_zeros((rows, 1)) if is_aligned else b1.new_zeros((rows, cols))
    c1, c2 = b1.detach().clone(), b2.detach().clone()
    c1[:, 2:4].clamp_(min=1e-3)
    c2[:, 2:4].clamp_(min=1e-3)
    return rotated.box_iou_rotated(c1, c2, mode, is_aligned)


def rbbox2roi(boxes_per_img):
    """core/bbox/transforms.py:73-92."""
    out = []
    for i, b in enumerate(boxes_per_img):
        out.append(torch.cat([b.new_full((b.shape[0], 1), i), b[:, :5]], -1) if b.shape[0] else b.new_zeros((0, 6)))
    return torch.cat(out, 0)


def mil_gen_proposals(points, boxes_obb, cfg, gt_obb, img_metas):
    """models/detectors/syn_images_generator_v2.py:26-40: bags are generated on the horizontal
    (cx,cy,w,h) box and re-attached to the pseudo box's angle."""
    hb = [hbb.cxcywh_to_xyxy(b[:, :4]) for b in boxes_obb]
    props, valids = hbb.fine_proposals(hb, cfg, img_metas)
    num_aug = int(props[0].shape[0] / points[0].shape[0])
    ref, real, out = [], [], []
    for i in range(len(boxes_obb)):
        ref.append(boxes_obb[i][:, None, :].repeat(1, num_aug, 1).reshape(-1, 5))
        real.append(gt_obb[i][:, None, :].repeat(1, num_aug, 1).reshape(-1, 5))
        ang = boxes_obb[i][:, -1].reshape(-1, 1)[:, None, :].repeat(1, num_aug, 1).reshape(-1, 1)
        out.append(torch.cat([hbb.xyxy_to_cxcywh(props[i]), ang], 1))
    return out, valids, ref, real


def sample_negative_boxes(n, img_shape, generator=None):
    """syn_images_generator_v2.py:142-148: five CPU draws; the box is (x1,y1,x2,y2,theta) and is later
    consumed AS IF it were (cx,cy,w,h,theta) (reference quirk, kept)."""
    h, w, _ = img_shape
    x1 = torch.rand(n, generator=generator) * w * 0.8
    y1 = torch.rand(n, generator=generator) * h * 0.8
    x2 = x1 + torch.rand(n, generator=generator) * 200
    y2 = y1 + torch.rand(n, generator=generator) * 200
    th = torch.rand(n, generator=generator) * math.pi - math.pi / 2
    return torch.stack([x1, y1, x2, y2, th], 1)


def negative_weights(neg, pos_bags):
    """syn_images_generator_v2.py:151-152."""
    iou = rbbox_overlaps(neg, pos_bags)
    return (iou < 0.3).sum(1) == iou.shape[1]


def gen_negative_proposals(points, cfg, pos_bags, img_metas, injected=None, generator=None):
    n = cfg["gen_num_neg"]
    if n == 0:
        return None, None
    negs, ws = [], []
    for i in range(len(points)):
        nb = injected[i] if injected is not None else sample_negative_boxes(n, img_metas[i]["img_shape"], generator)
        negs.append(nb)
        ws.append(negative_weights(nb, pos_bags[i]))
    return negs, ws


DIFFERENTIABLE_ROI = False      # tests that need torch autograd through the extraction flip this


def rotated_roi_extract(feats, rois, strides, out_size=7, sampling_ratio=2, clockwise=True):
    """models/roi_heads/roi_extractors/rotate_single_level_roi_extractor.py:90-148, single level."""
    assert len(feats) == 1
    if rois.shape[0] == 0:
        return feats[0].new_zeros(0, feats[0].shape[1], out_size, out_size)
    if DIFFERENTIABLE_ROI:
        return rotated.roi_align_rotated_torch(feats[0], rois, out_size, 1 / strides[0], sampling_ratio, True, clockwise)
    return rotated.roi_align_rotated(feats[0], rois, out_size, 1 / strides[0], sampling_ratio, True, clockwise)


def mil_bag_extensive(P, x, strides, img_metas, props, valids, refs, reals, num_gt_per_img, cfg, stage):
    """models/dense_heads/rotated_fcos_head_p2rb_ts.py:1285-1343."""
    R = {}
    U1 = int(props[0].shape[0] / num_gt_per_img[0])
    pts = [p[:, :2] for p in props]
    ebags, evalid, _, eref = mil_gen_proposals(pts, props, cfg, refs, img_metas)
    _, _, _, ereal = mil_gen_proposals(pts, props, cfg, reals, img_metas)
    R["base_shaking_num"] = U1
    R["coarse_bags_iou"] = rbbox_overlaps(torch.cat(ebags), torch.cat(ereal), is_aligned=True).mean()
    U2 = int(ebags[0].shape[0] / (num_gt_per_img[0] * U1))
    R["extensive_shaking_num"] = U2
    feats = rotated_roi_extract(x, rbbox2roi(ebags), strides).flatten(1)
    hid = hbb._fcs(feats, P.shared_fcs_reg[stage])
    deltas = F.linear(hid, *P.fc_reg[stage])
    cat = torch.cat(ebags)
    pred = hbb.delta2bbox(hbb.cxcywh_to_xyxy(cat[:, :4]), deltas, max_shape=img_metas[0]["img_shape"])
    pred_d = pred.clone().detach()
    target = hbb.cxcywh_to_xyxy(torch.cat(eref)[:, :4])
    R["loss_mil_bbox"] = hbb.dn_diou_loss(pred, target, torch.cat(evalid).reshape(-1).float(),
                                          avg_factor=pred.shape[0], hyper=0.2)
    refined, idx = [], 0
    for b in ebags:
        n = b.shape[0]
        refined.append(torch.cat([hbb.xyxy_to_cxcywh(pred_d[idx:idx + n]), b[:, -1].reshape(-1, 1)], 1))
        idx += n
    R["extensive_bags"], R["extensive_bags_valid"] = refined, evalid
    R["extensive_bags_reference"], R["extensive_bags_real"] = eref, ereal
    R["coarse_extensive_bags"] = ebags
    R["refine_bags_iou"] = rbbox_overlaps(torch.cat(refined), torch.cat(ereal), is_aligned=True).mean()
    return R


def forward_mil_head(P, num_gt, num_gt_per_img, x, strides, props, valids, refs, reals, img_metas, cfg, stage,
                     negs=None):
    """rotated_fcos_head_p2rb_ts.py:1365-1384 (+ classifier :1347-1363)."""
    R = mil_bag_extensive(P, x, strides, img_metas, props, valids, refs, reals, num_gt_per_img, cfg, stage)
    feats = rotated_roi_extract(x, rbbox2roi(R["extensive_bags"]), strides).flatten(1)
    hid = hbb._fcs(feats, P.shared_fcs_bag[stage])
    U1, U2 = R["base_shaking_num"], R["extensive_shaking_num"]
    R["cls_score"] = F.linear(hid, *P.fc_cls[stage]).view(num_gt, U1, U2, -1)
    R["ins_score"] = F.linear(hid, *P.fc_ins[stage]).view(num_gt, U1, U2, -1)
    if negs is not None:
        f = rotated_roi_extract(x, rbbox2roi(negs), strides).flatten(1)
        R["neg_cls_score"] = F.linear(hbb._fcs(f, P.shared_fcs_bag[stage]), *P.fc_cls[stage])
    return R


def mil_bag_training(R, labels_per_img, neg_weights):
    """rotated_fcos_head_p2rb_ts.py:1252-1283: as HBB but 0.25 * positive + 0.75 * negative."""
    cls, ins = R["cls_score"], R["ins_score"]
    G, U1, U2, C = cls.shape
    labels = torch.cat(labels_per_img).unsqueeze(1).repeat(1, U1).reshape(-1)
    valid = torch.cat(R["extensive_bags_valid"], 0).reshape(G, U1, U2, 1)
    bag = (cls.sigmoid() * hbb._instance_scores(ins, valid)).sum(2).reshape(-1, C)
    lw = (valid.reshape(G * U1, U2, 1).sum(1) > 0).float()
    num_sample = max(torch.sum(lw.sum(-1) > 0).float().item(), 1.)
    loss = 0.25 * (hbb.gfocal_loss(bag, F.one_hot(labels, C).float(), lw).sum() / num_sample)
    if neg_weights is not None:
        p = R["neg_cls_score"].sigmoid()
        nv = torch.cat(neg_weights).reshape(p.shape[0], -1).float()
        loss = loss + 0.75 * (hbb.gfocal_loss(p, torch.zeros_like(p), nv).sum() / num_sample)
    return loss


def mil_bag_selection(R, img_metas, pseudo_boxes, pseudo_labels, topk=3, beta=0.25):
    """rotated_fcos_head_p2rb_ts.py:1198-1250.  Quirk kept: (cx, cy) are BOTH clamped to [0, w] and then
    to [0, h] (:1211-1212); w, h, theta are score-weighted means, unclamped."""
    labels = torch.cat(pseudo_labels)
    cls, ins = R["cls_score"].detach().clone(), R["ins_score"].detach().clone()
    G, U1, U2, C = cls.shape
    valid = torch.cat(R["extensive_bags_valid"], 0).reshape(G, U1, U2, 1)
    bags = torch.cat(R["extensive_bags"], 0).reshape(G, U1 * U2, 5)
    cls = cls.reshape(G, U1 * U2, C).sigmoid()
    ins = hbb._instance_scores(ins, valid).reshape(G, U1 * U2, C)
    ar = torch.arange(G)
    cls, ins = cls[ar, :, labels], ins[ar, :, labels]
    sizes = [len(b) for b in pseudo_boxes]
    merged, all_idx, all_sc = [], [], []
    for c_i, i_i, bag_i, meta, pb in zip(cls.split(sizes), ins.split(sizes), bags.split(sizes), img_metas,
                                         pseudo_boxes):
        s = c_i * i_i
        sc, idx = s.topk(k=topk, dim=1)
        w = sc.unsqueeze(2).repeat(1, 1, 5)
        w = w / (w.sum(1, keepdim=True) + 1e-8)
        box = (bag_i[torch.arange(bag_i.shape[0]).unsqueeze(1), idx] * w).sum(1)
        h, wd, _ = meta["img_shape"]
        box[:, [0, 1]] = box[:, [0, 1]].clamp(0, wd)
        box[:, [0, 1]] = box[:, [0, 1]].clamp(0, h)
        merged.append((1 - beta) * box + beta * pb)
        all_idx.append(idx)
        all_sc.append(sc)
    return merged, torch.cat(all_idx), torch.cat(all_sc)


def phase2_refine(P, x, strides, img_metas, pseudo_boxes, pseudo_points, pseudo_labels, gt_boxes, fine_cfgs,
                  ext_cfgs, num_stages=1, cap=100, alpha=(0.01, 0.25), topk=3, beta=0.25, injected_negs=None):
    """models/detectors/rotated_fcos_teacher_student.py:494-535."""
    pb = [b[:cap].clone() for b in pseudo_boxes]
    gb = [b[:cap].clone() for b in gt_boxes]
    pp = [p[:cap].clone() for p in pseudo_points]
    pl = [l[:cap].clone() for l in pseudo_labels]
    out_b = [b.clone() for b in pseudo_boxes]
    out_p = [p.clone() for p in pseudo_points]
    losses = {"coarse_bboxes_iou": rbbox_overlaps(torch.cat(pb), torch.cat(gb), is_aligned=True).mean()}
    aux = []
    for s in range(num_stages):
        props, valids, refs, reals = mil_gen_proposals(pp, pb, fine_cfgs[s], gb, img_metas)
        negs, nw = gen_negative_proposals(pp, fine_cfgs[s], props, img_metas,
                                          None if injected_negs is None else injected_negs[s])
        num_gt = sum(b.shape[0] for b in pb)
        R = forward_mil_head(P, num_gt, [b.shape[0] for b in pb], x, strides, props, valids, refs, reals,
                             img_metas, ext_cfgs[s], s, negs)
        loss_bags = mil_bag_training(R, pl, nw)
        merged, idx, sc = mil_bag_selection(R, img_metas, pb, pl, topk, beta)
        R["selected_idx"], R["selected_scores"], R["neg_weight"] = idx, sc, nw
        pb = merged
        losses[f"stage{s}_refine_bboxes_iou"] = rbbox_overlaps(torch.cat(pb), torch.cat(gb), is_aligned=True).mean()
        losses[f"stage{s}_loss_mil_bbox"] = R["loss_mil_bbox"] * alpha[0]
        losses[f"stage{s}_loss_mil_bags"] = loss_bags * alpha[1]
        losses[f"stage{s}_coarse_bags_iou"] = R["coarse_bags_iou"]
        losses[f"stage{s}_refine_bags_iou"] = R["refine_bags_iou"]
        aux.append(R)
    for i in range(len(pb)):
        out_b[i][:cap] = pb[i]
        out_p[i][:cap] = pb[i][:, :2]
    return out_b, out_p, losses, aux
