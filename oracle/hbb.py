"""TEST INFRASTRUCTURE ONLY -- standalone CPU restatement (PyTorch fp32, same operation
order as the reference so results are bit-identical on CPU) of Point Teacher's HBB
phase-2 MIL refinement path, SURVEY.md section 8 rows a1-a5, a7-a12.

Pinned: ``oracle/check_oracle_vs_ref.py`` runs the reference's own files (under
``oracle/ref_shim.py``) on the same seeded inputs and requires bit-equality for box
geometry / validity / selection and <=1e-6 for scores and losses; the reference's own
known answers (GIoU vector, delta2bbox docstring) are asserted in tests/test_oracle.py.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference arm may
import this module; the product (point_teacher_b200/) never does.
All paths below are relative to /root/reference/HBB_TOD/mmdet/.
"""
import math

import torch
import torch.nn.functional as F


# ------------------------------------------------------------------ box formats
def xyxy_to_cxcywh(b):
    """core/bbox/transforms.py:250-261."""
    x1, y1, x2, y2 = b.unbind(-1)
    return torch.stack([(x1 + x2) / 2, (y1 + y2) / 2, x2 - x1, y2 - y1], -1)


def cxcywh_to_xyxy(b):
    """core/bbox/transforms.py:236-247."""
    cx, cy, w, h = b.unbind(-1)
    return torch.stack([cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h], -1)


def bbox2roi(boxes_per_img):
    """core/bbox/transforms.py:58-78: prepend the float image index."""
    out = []
    for i, b in enumerate(boxes_per_img):
        if b.shape[0]:
            out.append(torch.cat([b.new_full((b.shape[0], 1), i), b[:, :4]], -1))
        else:
            out.append(b.new_zeros((0, 5)))
    return torch.cat(out, 0)


def bbox_overlaps(b1, b2, mode="iou", is_aligned=False, eps=1e-6):
    """core/bbox/iou_calculators/iou2d_calculator.py:74-260 (2-D inputs only)."""
    assert mode in ("iou", "iof", "giou")
    rows, cols = b1.shape[0], b2.shape[0]
    if rows * cols == 0:
        return b1.new_zeros((rows,)) if is_aligned else b1.new_zeros((rows, cols))
    a1 = (b1[:, 2] - b1[:, 0]) * (b1[:, 3] - b1[:, 1])
    a2 = (b2[:, 2] - b2[:, 0]) * (b2[:, 3] - b2[:, 1])
    if is_aligned:
        lt, rb = torch.max(b1[:, :2], b2[:, :2]), torch.min(b1[:, 2:], b2[:, 2:])
        wh = (rb - lt).clamp(min=0)
        ov = wh[:, 0] * wh[:, 1]
        union = a1 + a2 - ov if mode != "iof" else a1
        if mode == "giou":
            elt, erb = torch.min(b1[:, :2], b2[:, :2]), torch.max(b1[:, 2:], b2[:, 2:])
    else:
        lt = torch.max(b1[:, None, :2], b2[None, :, :2])
        rb = torch.min(b1[:, None, 2:], b2[None, :, 2:])
        wh = (rb - lt).clamp(min=0)
        ov = wh[..., 0] * wh[..., 1]
        union = a1[:, None] + a2[None, :] - ov if mode != "iof" else a1[:, None].expand_as(ov)
        if mode == "giou":
            elt = torch.min(b1[:, None, :2], b2[None, :, :2])
            erb = torch.max(b1[:, None, 2:], b2[None, :, 2:])
    e = union.new_tensor([eps])
    union = torch.max(union, e)
    iou = ov / union
    if mode != "giou":
        return iou
    ewh = (erb - elt).clamp(min=0)
    ea = torch.max(ewh[..., 0] * ewh[..., 1], e)
    return iou - (ea - union) / ea


# ------------------------------------------------------------------ a1-a3 bags
def fine_proposals(boxes_per_img, cfg, img_metas):
    """models/detectors/syn_images_generator_v2.py:262-324 (gen_proposal_mode='fix_gen',
    cut_mode forced to None at :266 => validity = iof(box, image) > 0.7 at :317-319)."""
    assert cfg["gen_proposal_mode"] == "fix_gen"
    ratios, shake, min_scale = cfg["base_ratios"], cfg["shake_ratio"], cfg["min_scale"]
    props, valids = [], []
    for i in range(len(img_metas)):
        base = boxes_per_img[i]
        grid = []
        for rw in ratios:
            for rh in ratios:
                c = xyxy_to_cxcywh(base)
                w = c[:, 2].clamp(min_scale, 1000) * rw
                h = c[:, 3].clamp(min_scale, 1000) * rh
                grid.append(cxcywh_to_xyxy(torch.stack([c[:, 0], c[:, 1], w, h], -1)))
        old = torch.stack(grid, 1)                                   # (G, R*R, 4)
        if shake is not None:
            parts = [old[:, :, None, :]]
            for r in shake:
                c = xyxy_to_cxcywh(old)
                cx, cy, w, h = c.unbind(-1)
                ctr = torch.stack([torch.stack([cx - r * w, cy], -1), torch.stack([cx + r * w, cy], -1),
                                   torch.stack([cx, cy - r * h], -1), torch.stack([cx, cy + r * h], -1)], 2)
                wh = torch.stack([w, h], -1)[:, :, None, :].expand(ctr.shape)
                parts.append(cxcywh_to_xyxy(torch.cat([ctr, wh], -1)))
            new = torch.cat(parts, 2)                                # (G, R*R, 1+4n, 4)
        else:
            new = old
        h_img, w_img, _ = img_metas[i]["img_shape"]
        img_box = new.new_tensor([[0, 0, w_img, h_img]])
        valids.append(bbox_overlaps(new.reshape(-1, 4), img_box, mode="iof") > 0.7)
        props.append(new.reshape(-1, 4))
    return props, valids


def mil_gen_proposals(points, boxes, cfg, gt_boxes, img_metas):
    """models/detectors/syn_images_generator_v2.py:134-145 ('refine' mode).  num_aug is
    derived from image 0 (reference quirk, :140)."""
    assert cfg["gen_mode"] == "refine"
    props, valids = fine_proposals(boxes, cfg, img_metas)
    num_aug = int(props[0].shape[0] / points[0].shape[0])
    ref = [b[:, None, :].repeat(1, num_aug, 1).reshape(-1, 4) for b in boxes]
    real = [g[:, None, :].repeat(1, num_aug, 1).reshape(-1, 4) for g in gt_boxes]
    return props, valids, ref, real


def sample_negative_boxes(n, img_shape, generator=None):
    """The four consecutive CPU ``torch.rand(n)`` draws of gen_negative_proposals
    (syn_images_generator_v2.py:247-251)."""
    h, w, _ = img_shape
    x1 = torch.rand(n, generator=generator) * w * 0.8
    y1 = torch.rand(n, generator=generator) * h * 0.8
    x2 = x1 + torch.rand(n, generator=generator) * 100
    y2 = y1 + torch.rand(n, generator=generator) * 100
    return torch.stack([x1, y1, x2, y2], 1)


def negative_weights(neg_boxes, pos_bags):
    """syn_images_generator_v2.py:254-255: a negative counts iff IoU < 0.3 with EVERY
    base bag of its image."""
    iou = bbox_overlaps(neg_boxes, pos_bags)
    return (iou < 0.3).sum(1) == iou.shape[1]


def gen_negative_proposals(points, cfg, pos_bags, img_metas, injected=None, generator=None):
    """syn_images_generator_v2.py:234-259; ``injected`` replaces the CPU RNG draws."""
    n = cfg["gen_num_neg"]
    if n == 0:
        return None, None
    negs, weights = [], []
    for i in range(len(points)):
        nb = injected[i] if injected is not None else sample_negative_boxes(
            n, img_metas[i]["img_shape"], generator)
        negs.append(nb)
        weights.append(negative_weights(nb, pos_bags[i]))
    return negs, weights


# ------------------------------------------------------------------ a5 RoI extractor
def map_roi_levels(rois, num_levels, finest_scale=56):
    """models/roi_heads/roi_extractors/single_level_roi_extractor.py:35-54."""
    scale = torch.sqrt((rois[:, 3] - rois[:, 1]) * (rois[:, 4] - rois[:, 2]))
    lvl = torch.floor(torch.log2(scale / finest_scale + 1e-6))
    return lvl.clamp(min=0, max=num_levels - 1).long()


def roi_rescale(rois, f):
    """models/roi_heads/roi_extractors/base_roi_extractor.py:61-83."""
    cx, cy = (rois[:, 1] + rois[:, 3]) * 0.5, (rois[:, 2] + rois[:, 4]) * 0.5
    nw, nh = (rois[:, 3] - rois[:, 1]) * f, (rois[:, 4] - rois[:, 2]) * f
    return torch.stack((rois[:, 0], cx - nw * 0.5, cy - nh * 0.5, cx + nw * 0.5, cy + nh * 0.5), -1)


def roi_align(feat, rois, out_size=7, spatial_scale=0.125, sampling_ratio=0, aligned=True):
    """mmcv.ops.RoIAlign(pool_mode='avg') == torchvision.ops.roi_align (same Detectron2
    kernel; SURVEY.md Appendix A.1, verified to 2.7e-5)."""
    import torchvision
    return torchvision.ops.roi_align(feat, rois, (out_size, out_size), spatial_scale,
                                     sampling_ratio, aligned)


def single_roi_extract(feats, rois, strides, out_size=7, finest_scale=56, sampling_ratio=0,
                       aligned=True, roi_scale_factor=None):
    """single_level_roi_extractor.py:56-114."""
    C = feats[0].shape[1]
    out = feats[0].new_zeros(rois.shape[0], C, out_size, out_size)
    if len(feats) == 1:
        if rois.shape[0] == 0:
            return out
        return roi_align(feats[0], rois, out_size, 1 / strides[0], sampling_ratio, aligned)
    lvls = map_roi_levels(rois, len(feats), finest_scale)
    if roi_scale_factor is not None:
        rois = roi_rescale(rois, roi_scale_factor)
    for i in range(len(feats)):
        inds = (lvls == i).nonzero(as_tuple=False).squeeze(1)
        if inds.numel():
            out[inds] = roi_align(feats[i], rois[inds], out_size, 1 / strides[i], sampling_ratio, aligned)
    return out


# ------------------------------------------------------------------ a7 decode + loss
def delta2bbox(rois, deltas, max_shape=None, wh_ratio_clip=16 / 1000):
    """core/bbox/coder/delta_xywh_bbox_coder.py:144-270 with means 0 / stds 1."""
    d = deltas * deltas.new_tensor([1., 1., 1., 1.]) + deltas.new_tensor([0., 0., 0., 0.])
    dx, dy, dw, dh = d.unbind(-1)
    px, py = (rois[:, 0] + rois[:, 2]) * 0.5, (rois[:, 1] + rois[:, 3]) * 0.5
    pw, ph = rois[:, 2] - rois[:, 0], rois[:, 3] - rois[:, 1]
    mr = abs(math.log(wh_ratio_clip))
    dw, dh = dw.clamp(min=-mr, max=mr), dh.clamp(min=-mr, max=mr)
    gw, gh = pw * dw.exp(), ph * dh.exp()
    gx, gy = px + pw * dx, py + ph * dy
    out = torch.stack([gx - gw * 0.5, gy - gh * 0.5, gx + gw * 0.5, gy + gh * 0.5], -1)
    if max_shape is not None:
        hi = out.new_tensor([max_shape[1], max_shape[0], max_shape[1], max_shape[0]])
        out = torch.where(out < 0, out.new_tensor(0), out)
        out = torch.where(out > hi, hi.expand_as(out), out)
    return out


def _diou_elem(pred, target, eps):
    """models/losses/iou_loss.py:139-190 body."""
    lt, rb = torch.max(pred[:, :2], target[:, :2]), torch.min(pred[:, 2:], target[:, 2:])
    wh = (rb - lt).clamp(min=0)
    ov = wh[:, 0] * wh[:, 1]
    ap = (pred[:, 2] - pred[:, 0]) * (pred[:, 3] - pred[:, 1])
    ag = (target[:, 2] - target[:, 0]) * (target[:, 3] - target[:, 1])
    iou = ov / (ap + ag - ov + eps)
    ewh = (torch.max(pred[:, 2:], target[:, 2:]) - torch.min(pred[:, :2], target[:, :2])).clamp(min=0)
    c2 = ewh[:, 0] ** 2 + ewh[:, 1] ** 2 + eps
    left = ((target[:, 0] + target[:, 2]) - (pred[:, 0] + pred[:, 2])) ** 2 / 4
    right = ((target[:, 1] + target[:, 3]) - (pred[:, 1] + pred[:, 3])) ** 2 / 4
    return 1 - (iou - (left + right) / c2)


def dn_diou_loss(pred, target, weight, avg_factor, hyper=0.2, eps=1e-6, loss_weight=1.0):
    """models/losses/iou_loss.py:398-466 + DN_DIoULoss.forward :851-881.  Quirk kept:
    ``base_loss`` is the *mean-reduced scalar* DIoU over all rows (the inner call goes
    through @weighted_loss with the default reduction)."""
    if weight is not None and not torch.any(weight > 0):
        return (pred * weight.unsqueeze(1)).sum()
    base = _diou_elem(pred, target, eps).mean()
    anx = hyper / 2
    w, h = target[:, 2] - target[:, 0], target[:, 3] - target[:, 1]
    bank = []
    for i in (-1, 0, 1):
        for j in (-1, 0, 1):
            t = target.clone()
            t[:, 0] = t[:, 0] - anx * w * i
            t[:, 2] = t[:, 2] + anx * w * j
            t[:, 1] = t[:, 1] - anx * h * i
            t[:, 3] = t[:, 3] + anx * h * j
            bank.append(_diou_elem(pred, t, eps).reshape(-1, 1))
    ev = (base + torch.cat(bank, 1).min(1)[0]) / 2
    if weight is not None:
        ev = ev * weight
    return loss_weight * (ev.sum() / avg_factor)


# ------------------------------------------------------------------ the MIL head
class MilHeadParams:
    """Plain container for the MIL parameters of TS_P2BFCOSHead (names as in
    models/dense_heads/fcos_head_p2b_ts.py:231-263)."""

    def __init__(self, num_classes=8, num_stages=1, in_channels=256, seed=0, std=0.01,
                 reg_out=4):
        g = torch.Generator().manual_seed(seed)

        def lin(i, o):
            return [torch.randn(o, i, generator=g) * std, torch.zeros(o)]
        self.num_classes, self.num_stages = num_classes, num_stages
        self.shared_fcs_reg, self.shared_fcs_bag = [], []
        self.fc_cls, self.fc_ins, self.fc_reg = [], [], []
        for _ in range(num_stages):
            self.shared_fcs_reg.append([lin(in_channels * 49, 1024), lin(1024, 1024)])
            self.shared_fcs_bag.append([lin(in_channels * 49, 1024), lin(1024, 1024)])
            self.fc_cls.append(lin(1024, num_classes))
            self.fc_ins.append(lin(1024, num_classes))
            self.fc_reg.append(lin(1024, reg_out))

    def state_dict(self):
        sd = {}
        for s in range(self.num_stages):
            for nm, fcs in (("shared_fcs_reg", self.shared_fcs_reg), ("shared_fcs_bag", self.shared_fcs_bag)):
                for j in range(2):
                    sd[f"{nm}.{s}.{j}.weight"], sd[f"{nm}.{s}.{j}.bias"] = fcs[s][j]
            for nm, fc in (("fc_cls", self.fc_cls), ("fc_ins", self.fc_ins), ("fc_reg", self.fc_reg)):
                sd[f"{nm}.{s}.weight"], sd[f"{nm}.{s}.bias"] = fc[s]
        return sd

    def requires_grad_(self, flag=True):
        for t in self.state_dict().values():
            t.requires_grad_(flag)
        return self


def _fcs(x, fcs):
    for w, b in fcs:
        x = F.relu(F.linear(x, w, b))
    return x


def gfocal_loss(p, q, w=1.0, eps=1e-6):
    """fcos_head_p2b_ts.py:1074-1078."""
    l1 = (p - q) ** 2
    l2 = q * (p + eps).log() + (1 - q) * (1 - p + eps).log()
    return -(l1 * l2 * w).sum(-1)


def mil_bag_extensive(P, x, strides, img_metas, props, valids, refs, reals, num_gt_per_img, cfg, stage):
    """fcos_head_p2b_ts.py:1182-1236."""
    R = {}
    U1 = int(props[0].shape[0] / num_gt_per_img[0])
    pts = [xyxy_to_cxcywh(p)[:, :2] for p in props]
    ebags, evalid, _, eref = mil_gen_proposals(pts, props, cfg, refs, img_metas)
    _, _, _, ereal = mil_gen_proposals(pts, props, cfg, reals, img_metas)
    R["base_shaking_num"] = U1
    R["coarse_bags_iou"] = bbox_overlaps(torch.cat(ebags), torch.cat(ereal), is_aligned=True).mean()
    U2 = int(ebags[0].shape[0] / (num_gt_per_img[0] * U1))
    R["extensive_shaking_num"] = U2
    rois = bbox2roi(ebags)
    feats = single_roi_extract(x, rois, strides).flatten(1)
    R["reg_roi_feats"] = feats
    hid = _fcs(feats, P.shared_fcs_reg[stage])
    deltas = F.linear(hid, *P.fc_reg[stage])
    R["reg_deltas"] = deltas
    pred = delta2bbox(torch.cat(ebags), deltas, max_shape=img_metas[0]["img_shape"])
    pred_d = pred.clone().detach()
    R["loss_mil_bbox"] = dn_diou_loss(pred, torch.cat(eref), torch.cat(evalid).reshape(-1).float(),
                                      avg_factor=pred.shape[0], hyper=0.2)
    R["refine_bags_iou"] = bbox_overlaps(pred_d, torch.cat(ereal), is_aligned=True).mean()
    R["iou_target"] = bbox_overlaps(pred_d, torch.cat(eref), is_aligned=True).reshape(-1)
    sizes = [b.shape[0] for b in ebags]
    R["extensive_bags"] = list(torch.split(pred_d, sizes))
    R["extensive_bags_valid"], R["extensive_bags_reference"], R["extensive_bags_real"] = evalid, eref, ereal
    R["coarse_extensive_bags"] = ebags
    return R


def mil_bag_classifier(P, num_gt, x, strides, R, stage):
    """fcos_head_p2b_ts.py:1240-1256."""
    rois = bbox2roi(R["extensive_bags"])
    feats = single_roi_extract(x, rois, strides).flatten(1)
    hid = _fcs(feats, P.shared_fcs_bag[stage])
    cls, ins = F.linear(hid, *P.fc_cls[stage]), F.linear(hid, *P.fc_ins[stage])
    U1, U2 = R["base_shaking_num"], R["extensive_shaking_num"]
    R["cls_score"] = cls.view(num_gt, U1, U2, -1)
    R["ins_score"] = ins.view(num_gt, U1, U2, -1)


def forward_mil_head(P, num_gt, num_gt_per_img, x, strides, props, valids, refs, reals, img_metas,
                     cfg, stage, negs=None):
    """fcos_head_p2b_ts.py:1259-1277."""
    R = mil_bag_extensive(P, x, strides, img_metas, props, valids, refs, reals, num_gt_per_img, cfg, stage)
    mil_bag_classifier(P, num_gt, x, strides, R, stage)
    if negs is not None:
        f = single_roi_extract(x, bbox2roi(negs), strides).flatten(1)
        R["neg_cls_score"] = F.linear(_fcs(f, P.shared_fcs_bag[stage]), *P.fc_cls[stage])
    return R


def _instance_scores(ins, valid4):
    """softmax over U2, mask by validity, L1-normalise over U2
    (fcos_head_p2b_ts.py:1128-1130 and :1158-1160)."""
    ins = ins.softmax(dim=2) * valid4
    return F.normalize(ins, dim=2, p=1)


def mil_bag_training(R, labels_per_img, neg_weights):
    """fcos_head_p2b_ts.py:1147-1180."""
    cls, ins = R["cls_score"], R["ins_score"]
    G, U1, U2, C = cls.shape
    labels = torch.cat(labels_per_img).unsqueeze(1).repeat(1, U1).reshape(-1)
    valid = torch.cat(R["extensive_bags_valid"], 0).reshape(G, U1, U2, 1)
    bag = (cls.sigmoid() * _instance_scores(ins, valid)).sum(2).reshape(-1, C)
    lw = (valid.reshape(G * U1, U2, 1).sum(1) > 0).float()
    num_sample = max(torch.sum(lw.sum(-1) > 0).float().item(), 1.)
    onehot = F.one_hot(labels, C).float()
    loss = gfocal_loss(bag, onehot, lw).sum() / num_sample
    if neg_weights is not None:
        p = R["neg_cls_score"].sigmoid()
        nv = torch.cat(neg_weights).reshape(p.shape[0], -1).float()
        loss = loss + gfocal_loss(p, torch.zeros_like(p), nv).sum() / num_sample
    return loss


def mil_bag_selection(R, img_metas, pseudo_boxes, pseudo_labels, topk=1, beta=0.25):
    """fcos_head_p2b_ts.py:1092-1145.  Returns (merged boxes per image, selected
    instance indices (G, topk), selected scores (G, topk))."""
    labels = torch.cat(pseudo_labels)
    cls, ins = R["cls_score"].detach().clone(), R["ins_score"].detach().clone()
    G, U1, U2, C = cls.shape
    valid = torch.cat(R["extensive_bags_valid"], 0).reshape(G, U1, U2, 1)
    bags = torch.cat(R["extensive_bags"], 0).reshape(G, U1 * U2, 4)
    cls = cls.reshape(G, U1 * U2, C).sigmoid()
    ins = _instance_scores(ins, valid).reshape(G, U1 * U2, C)
    ar = torch.arange(G)
    cls, ins = cls[ar, :, labels], ins[ar, :, labels]
    sizes = [len(b) for b in pseudo_boxes]
    merged, all_idx, all_sc = [], [], []
    for c_i, i_i, bag_i, meta, pb in zip(cls.split(sizes), ins.split(sizes), bags.split(sizes),
                                         img_metas, pseudo_boxes):
        s = c_i * i_i
        sc, idx = s.topk(k=topk, dim=1)
        w = sc.unsqueeze(2).repeat(1, 1, 4)
        w = w / (w.sum(1, keepdim=True) + 1e-8)
        picked = bag_i[torch.arange(bag_i.shape[0]).unsqueeze(1), idx]
        box = (picked * w).sum(1)
        h, wd, _ = meta["img_shape"]
        box[:, 0:4:2] = box[:, 0:4:2].clamp(0, wd)
        box[:, 1:4:2] = box[:, 1:4:2].clamp(0, h)
        merged.append((1 - beta) * box + beta * pb)
        all_idx.append(idx)
        all_sc.append(sc)
    return merged, torch.cat(all_idx), torch.cat(all_sc)


def mil_head_burn_in_step2(P, x, strides, img_metas, props, valids, refs, reals, negs, neg_weights,
                           pseudo_boxes, pseudo_labels, cfg, stage, topk=1, beta=0.25):
    """fcos_head_p2b_ts.py:1318-1344.  Returns (losses, merged, aux)."""
    num_gt = torch.cat(pseudo_boxes).shape[0]
    per_img = [b.shape[0] for b in pseudo_boxes]
    R = forward_mil_head(P, num_gt, per_img, x, strides, props, valids, refs, reals, img_metas, cfg,
                         stage, negs)
    losses = {f"stage{stage}_loss_mil_bbox": R["loss_mil_bbox"],
              f"stage{stage}_loss_mil_bags": mil_bag_training(R, pseudo_labels, neg_weights),
              f"stage{stage}_coarse_bags_iou": R["coarse_bags_iou"],
              f"stage{stage}_refine_bags_iou": R["refine_bags_iou"]}
    merged, idx, sc = mil_bag_selection(R, img_metas, pseudo_boxes, pseudo_labels, topk, beta)
    R["selected_idx"], R["selected_scores"] = idx, sc
    return losses, merged, R


def phase2_refine(P, x, strides, img_metas, pseudo_boxes, pseudo_points, pseudo_labels, gt_boxes,
                  fine_cfgs, ext_cfgs, num_stages=1, cap=100, alpha=(0.01, 0.25), topk=1, beta=0.25,
                  injected_negs=None, generator=None):
    """detectors/fcos_p2b_teacher_student.py:425-466 (forward_mil_head_burn_in_step2):
    cap at ``cap`` GTs per image, per-stage bag gen + MIL head, write-back into clones."""
    n = len(pseudo_boxes)
    pb = [b[:cap].clone() for b in pseudo_boxes]
    gb = [b[:cap].clone() for b in gt_boxes]
    pp = [p[:cap].clone() for p in pseudo_points]
    pl = [l[:cap].clone() for l in pseudo_labels]
    out_b = [b.clone() for b in pseudo_boxes]
    out_p = [p.clone() for p in pseudo_points]
    losses = {"coarse_bboxes_iou": bbox_overlaps(torch.cat(pb), torch.cat(gb), is_aligned=True).mean()}
    aux = []
    for s in range(num_stages):
        props, valids, refs, reals = mil_gen_proposals(pp, pb, fine_cfgs[s], gb, img_metas)
        negs, nw = gen_negative_proposals(pp, fine_cfgs[s], props, img_metas,
                                          None if injected_negs is None else injected_negs[s], generator)
        ml, pb, R = mil_head_burn_in_step2(P, x, strides, img_metas, props, valids, refs, reals, negs, nw,
                                           pb, pl, ext_cfgs[s], s, topk, beta)
        losses[f"stage{s}_refine_bboxes_iou"] = bbox_overlaps(torch.cat(pb), torch.cat(gb), is_aligned=True).mean()
        ml[f"stage{s}_loss_mil_bbox"] = ml[f"stage{s}_loss_mil_bbox"] * alpha[0]
        ml[f"stage{s}_loss_mil_bags"] = ml[f"stage{s}_loss_mil_bags"] * alpha[1]
        losses.update(ml)
        aux.append(R)
    for i in range(n):
        out_b[i][:cap] = pb[i]
        out_p[i][:cap] = xyxy_to_cxcywh(pb[i])[:, :2]
    return out_b, out_p, losses, aux
