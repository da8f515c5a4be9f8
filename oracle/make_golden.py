"""TEST INFRASTRUCTURE ONLY (dev container, needs /root/reference) -- generates the committed
fixtures under tests/golden/ by running the reference's OWN files (oracle/ref_shim.py) on seeded
synthetic inputs.  Inputs are not stored: tests regenerate them from the same seeds with
point_teacher_b200.synth + oracle.hbb.MilHeadParams (CPU generators are machine independent).
Run: python -m oracle.make_golden"""
import os

import torch

from oracle import ref_shim
from point_teacher_b200 import synth

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

SMALL = dict(batch=2, img_hw=(256, 256), gt_range=(6, 10), n_neg=20)
OBB_SMALL = dict(batch=2, img_hw=(256, 256), gt_range=(6, 10), n_neg=20)


def hbb_case(seed, stages, topk, tag):
    ns = ref_shim.install()
    d = synth.hbb_batch(seed=seed, num_stages=stages, **SMALL)
    head = ref_shim.build_ref_mil_head(ns, num_stages=stages, top_k=topk, seed=seed)
    fine, ext = synth.HBB_FINE_CFG, synth.HBB_EXT_CFG
    pb, gb = [b.clone() for b in d["pseudo_boxes"]], d["gt_boxes"]
    pp, pl = d["pseudo_points"], d["pseudo_labels"]
    out = dict(seed=seed, stages=stages, topk=topk, small=SMALL, per_stage=[])
    with torch.no_grad():
        for s in range(stages):
            props, valids, refs, reals = ns.syn.MIL_gen_proposals_from_cfg(pp, pb, fine[s], gb, d["img_metas"])
            negs = d["neg_boxes"][s]
            nw = [((ns.bbox_overlaps(negs[i], props[i]) < 0.3).sum(1) == props[i].shape[0]) for i in range(len(negs))]
            num_gt = sum(b.shape[0] for b in pb)
            per_img = [b.shape[0] for b in pb]
            R = head.forward_mil_head(num_gt, per_img, (d["feat"],), props, valids, refs, reals, d["img_metas"],
                                      ext[s], s, negs, nw)
            loss_bags = head.mil_bag_training(R, pl, nw)
            merged = head.mil_bag_selection(R, d["img_metas"], pb, pl)
            # the coarse extensive bags (inputs of the first RoIAlign)
            pts = [ns.transforms.bbox_xyxy_to_cxcywh(p)[:, :2] for p in props]
            ebags, evalid, _, _ = ns.syn.MIL_gen_proposals_from_cfg(pts, props, ext[s], refs, d["img_metas"])
            out["per_stage"].append(dict(
                base_bags=torch.cat(props), base_valid=torch.cat(valids), neg_weight=torch.cat(nw),
                ext_bags=torch.cat(ebags), ext_valid=torch.cat(evalid),
                refined_bags=torch.cat(R["extensive_bags"]), iou_target=R["iou_target"],
                cls_score=R["cls_score"], ins_score=R["ins_score"], neg_cls_score=R["neg_cls_score"],
                loss_mil_bbox=R["loss_mil_bbox"], loss_mil_bags=loss_bags, coarse_bags_iou=R["coarse_bags_iou"],
                refine_bags_iou=R["refine_bags_iou"], merged=torch.cat(merged)))
            pb = list(merged)
    torch.save(out, os.path.join(OUT, f"hbb_phase2_{tag}.pt"))
    print("wrote", tag, {k: tuple(v.shape) for k, v in out["per_stage"][0].items() if hasattr(v, "shape")})


def obb_case(seed, tag):
    """OBB twin: the reference's own OBB_TOD files (rotated kernels bound to oracle/rotated.py by the shim)."""
    from oracle import check_oracle_vs_ref as chk
    o = ref_shim.install_obb()
    d = synth.obb_batch(seed=seed, **OBB_SMALL)
    head = ref_shim.build_ref_obb_mil_head(o, seed=seed)
    with torch.no_grad():
        pb, losses, per_stage = chk.run_ref_obb(o, d, head, 1)
    st = per_stage[0]
    R = st["R"]
    cap = 100
    pbx = [b[:cap] for b in d["pseudo_boxes"]]
    props, _, refs, _ = o.syn.MIL_gen_proposals_from_cfg([b[:cap] for b in d["pseudo_points"]], pbx,
                                                         synth.OBB_FINE_CFG[0], [b[:cap] for b in d["gt_boxes"]],
                                                         d["img_metas"])
    ebags, _, _, _ = o.syn.MIL_gen_proposals_from_cfg([p[:, :2] for p in props], props, synth.OBB_EXT_CFG[0], refs,
                                                      d["img_metas"])
    out = dict(seed=seed, small=OBB_SMALL, topk=3, merged=st["merged"], neg_weight=st["neg_weight"],
               base_bags=st["base_bags"], base_valid=st["base_valid"],
               ext_bags=torch.cat(ebags),
               ext_valid=torch.cat(R["extensive_bags_valid"]), refined_bags=torch.cat(R["extensive_bags"]),
               cls_score=R["cls_score"], ins_score=R["ins_score"], neg_cls_score=R["neg_cls_score"],
               losses={k: v.clone() for k, v in losses.items()})
    torch.save(out, os.path.join(OUT, f"obb_phase2_{tag}.pt"))
    print("wrote obb", tag, {k: tuple(v.shape) for k, v in out.items() if hasattr(v, "shape")})


def assign_case():
    """Label assignment (rows a13-a15): outputs of the reference's own assigner / metric files."""
    ns = ref_shim.install()
    out = dict(cases=[])
    for seed, ties, npre, tk in [(0, True, 3, 3), (1, True, 1, 1), (2, True, 5, 3), (3, False, 5, 3), (4, True, 7, 2)]:
        d = synth.assign_batch(seed, ties=ties)
        a = ns.TopkAssigner(num_pre=npre, topk=tk, cls_cost=dict(type="FocalLossCost", weight=1.0),
                            reg_cost=dict(type="PointCost", mode="L1", weight=1.0))
        r1 = a.assign(d["pred"], d["logits"], d["gt"], d["labels"])
        f = ns.FUSETopkAssigner(num_pre=npre, topk=tk, cls_cost=dict(type="FocalLossCost", weight=1.0),
                                reg_cost=dict(type="PointCost", mode="L1", weight=1.0),
                                location_cost=dict(type="InsiderCost", weight=1.0))
        r2 = f.assign(d["pred"], d["points"], d["logits"], None, d["gt"], d["labels"])
        out["cases"].append(dict(seed=seed, ties=ties, num_pre=npre, topk=tk, topk_gt_inds=r1.gt_inds,
                                 topk_labels=r1.labels, fuse_gt_inds=r2.gt_inds, fuse_labels=r2.labels,
                                 fl_table=ns.match_cost.FocalLossCost()(d["logits"], torch.arange(8))))
    g = torch.Generator().manual_seed(9)
    gts = synth.make_boxes(g, 23, (400, 400))
    anchors = torch.cat([synth.jitter_boxes(g, gts.repeat(6, 1), 3.0, 0.4), synth.make_boxes(g, 300, (400, 400))])
    anchors[5] = gts[5]
    anchors[77] = gts[5]
    labels = torch.randint(0, 8, (23,), generator=g)
    out["gts"], out["anchors"], out["labels"] = gts, anchors, labels
    out["metric"] = {m: ns.BboxDistanceMetric()(gts, anchors, m)
                     for m in ("iou", "iof", "giou", "wd", "kl", "center_distance2", "exp_kl", "kl_10")}
    out["maxiou"] = []
    for kw in (dict(pos_iou_thr=0.5, neg_iou_thr=0.4, min_pos_iou=0.0), dict(pos_iou_thr=0.7, neg_iou_thr=0.3, min_pos_iou=0.3),
               dict(pos_iou_thr=0.5, neg_iou_thr=0.5, min_pos_iou=0.0, gt_max_assign_all=False),
               dict(pos_iou_thr=0.5, neg_iou_thr=0.5, match_low_quality=False)):
        for calc, mode in ((dict(type="BboxOverlaps2D"), "iou"), (dict(type="BboxDistanceMetric"), "wd")):
            a = ns.MaxIoUAssigner(iou_calculator=calc, **kw)
            r = a.assign_wrt_overlaps(a.iou_calculator(gts, anchors, mode), labels)
            out["maxiou"].append(dict(kw=kw, calc=calc["type"], mode=mode, gt_inds=r.gt_inds, max_overlaps=r.max_overlaps,
                                      labels=r.labels))
    torch.save(out, os.path.join(OUT, "assign.pt"))
    print("wrote assign", len(out["cases"]), len(out["maxiou"]))


def mask_case():
    """Row a16: outputs of the reference's own generate_black_paper (seeded torch + numpy RNG)."""
    import numpy as np
    ns = ref_shim.install()
    out = []
    for seed in (0, 3):
        d = synth.mask_batch(seed)
        pattern, prior = ns.syn.load_basic_shape(synth.SHAPE_LIST)
        torch.manual_seed(seed)
        np.random.seed(seed)
        img_ref, bb_ref = ns.syn.generate_black_paper(d["img"].clone(), d["bb_occupied"].clone(), d["img"].clone(), pattern,
                                                      prior, range(int(len(pattern) / 2)), d["imgsize"])
        filled = (img_ref != d["img"]).any(0) | ((img_ref == 255).all(0) & (d["img"] == 255).all(0))
        m = (img_ref == 255).all(0)
        out.append(dict(seed=seed, kept=bb_ref, mask_bits=torch.from_numpy(np.packbits(m.numpy())),
                        n_px=int(m.sum()), changed=int(filled.sum())))
    torch.save(out, os.path.join(OUT, "black_paper.pt"))
    print("wrote black_paper", [(o["kept"].shape[0], o["n_px"]) for o in out])


def pseudo_case():
    """Section 8f rank 1: outputs of the reference's own TS_P2BFCOSHead._gnerate_pseudo_single."""
    ns = ref_shim.install()
    out = []
    for seed, G in ((0, 120), (2, 400)):
        d = synth.pseudo_batch(seed, G=G)
        head = ns.TS_P2BFCOSHead.__new__(ns.TS_P2BFCOSHead)
        torch.nn.Module.__init__(head)
        head.fuse_assigner = ns.FUSETopkAssigner(num_pre=5, topk=3, cls_cost=dict(type="FocalLossCost", weight=1.0),
                                                 reg_cost=dict(type="PointCost", mode="L1", weight=1.0),
                                                 location_cost=dict(type="InsiderCost", weight=1.0))
        rb, rp, rl, rm, rv = head._gnerate_pseudo_single(d["gt_points"], d["labels"], d["gt_boxes"], d["logits"], d["ltrb"],
                                                         None, None, None, 0.1, d["points"], None)
        head.num_classes = 8
        head.assigner = ns.TopkAssigner(num_pre=1, topk=1, cls_cost=dict(type="FocalLossCost", weight=1.0),
                                        reg_cost=dict(type="PointCost", mode="L1", weight=1.0))
        head.pseudo_assigner = ns.TopkAssigner(num_pre=3, topk=3, cls_cost=dict(type="FocalLossCost", weight=1.0),
                                               reg_cost=dict(type="PointCost", mode="L1", weight=1.0))
        t = head._get_target_pseudo_single(d["gt_points"], d["labels"], rp, d["labels"], rb, d["logits"], d["ltrb"], None,
                                           dict(ori_filename="x"), None, None, d["points"], None, False)
        pos = (t[0] != 8).nonzero().reshape(-1)
        out.append(dict(seed=seed, G=G, boxes=rb, points=rp, mean_iou=rm, valid=torch.sort(rv)[0],
                        labels_reg=t[0].to(torch.int16), bbox_targets_pos=t[1][pos], pos=pos.to(torch.int32),
                        labels=t[2].to(torch.int16), centerness_pos=head.centerness_target(t[1][pos]),
                        targets_checksum=t[1].double().sum()))
    torch.save(out, os.path.join(OUT, "pseudo_boxes.pt"))
    print("wrote pseudo_boxes", [(o["G"], o["valid"].numel()) for o in out])


def augment_case():
    """Section 8f rank 3: outputs of the reference's own strong_augmentation (HBB and OBB) with seeded
    ``random`` / ``np.random``; images stored as uint8 (the reference rounds them)."""
    import random

    import numpy as np
    from oracle import augment
    ns = ref_shim.install()
    o = ref_shim.install_obb()
    out = []
    for rot, seed in ((False, 1), (False, 5), (True, 1), (True, 6)):
        d = synth.augment_batch(seed, rotated=rot)
        args = [d["img"].clone()] + [[t.clone() for t in d[k]] for k in
                                     ("gt_points", "gt_labels", "pseudo_points", "pseudo_labels", "pseudo_bboxes")]
        random.seed(seed)
        np.random.seed(seed)
        choices = augment.draw_choices(2, rotated=rot)
        random.seed(seed)
        np.random.seed(seed)
        r = o.syn.strong_augmentation(*args, "le90") if rot else ns.syn.strong_augmentation(*args)
        assert r[0].min() >= 0 and r[0].max() <= 255 and torch.equal(r[0], r[0].round())
        out.append(dict(rotated=rot, seed=seed, choices=choices, images=r[0].to(torch.uint8), gt_points=r[2],
                        gt_labels=r[3], pseudo_points=r[4], pseudo_labels=r[5], pseudo_bboxes=r[6]))
    torch.save(out, os.path.join(OUT, "augment.pt"))
    print("wrote augment", [(c["rotated"], c["choices"], [len(x) for x in c["gt_points"]]) for c in out])


def overlaps_case():
    ns = ref_shim.install()
    g = torch.Generator().manual_seed(7)
    a = synth.make_boxes(g, 37, (800, 800))
    b = synth.jitter_boxes(g, synth.make_boxes(g, 53, (800, 800)))
    out = {}
    for mode in ("iou", "iof", "giou"):
        out[mode] = ns.bbox_overlaps(a, b, mode)
        out[mode + "_aligned"] = ns.bbox_overlaps(a, synth.jitter_boxes(torch.Generator().manual_seed(8), a), mode, True)
    torch.save(out, os.path.join(OUT, "bbox_overlaps.pt"))
    print("wrote bbox_overlaps")


DET_CAP, DET_NEG = 7, 20      # detector-level cases: cap below every image's GT count; 20 negatives per image


def detector_inputs(rotated, seed, stages=1):
    """Seeded inputs of the detector-level cases (shared by this generator and tests/test_gpu_dropin.py)."""
    from oracle import detector as D
    if rotated:
        d = synth.obb_batch(seed=seed, **OBB_SMALL)
        fine, ext = synth.OBB_FINE_CFG * stages, synth.OBB_EXT_CFG * stages
    else:
        d = synth.hbb_batch(seed=seed, num_stages=stages, **SMALL)
        fine, ext = synth.HBB_FINE_CFG, synth.HBB_EXT_CFG
    fine = [dict(c, gen_num_neg=DET_NEG) for c in fine]
    syn_boxes, feat_syn = D.synthetic_boxes_like(d, seed, rotated)
    return d, fine, ext, syn_boxes, feat_syn


def grad_digest(t, n=256):
    """Norm + a fixed pseudo-random sample of entries (the FC1 weight gradient alone is 51 MB)."""
    flat = t.detach().reshape(-1).double()
    g = torch.Generator().manual_seed(flat.numel() % 9973)
    idx = torch.randperm(flat.numel(), generator=g)[:n].clone()
    return dict(norm=float(flat.norm()), idx=idx, val=flat[idx].float().clone())


def detector_case(rotated, step, seed, stages=1):
    """The reference's OWN detector-level callers (fcos_p2b_teacher_student.py:365-466 /
    rotated_fcos_teacher_student.py:435-535, run unbound on a stand-in ``self``) with the reference's own head, the
    negatives drawn by the reference's own ``gen_negative_proposals`` from the seeded global CPU generator, then
    ``_parse_losses(...).backward()``: refined boxes / points, every loss entry and gradient digests."""
    from oracle import detector as D
    dn = ref_shim.install_detectors()
    d, fine, ext, syn_boxes, feat_syn = detector_inputs(rotated, seed, stages)
    if rotated:
        head, cls = ref_shim.build_ref_obb_mil_head(dn.obb, num_stages=stages, seed=seed), dn.RotatedFCOS_TS
    else:
        head, cls = ref_shim.build_ref_mil_head(dn.hbb, num_stages=stages, top_k=1, seed=seed), dn.TS_P2B_FCOS
    det = D.make_detector(head, fine, ext, stages, cap1=DET_CAP, cap2=DET_CAP)
    n = len(d["pseudo_boxes"])
    fo = d["feat"].clone().requires_grad_(not rotated)
    fs = feat_syn.clone().requires_grad_(not rotated)
    torch.manual_seed(100 + seed)
    args = (d["pseudo_boxes"], d["pseudo_points"], d["pseudo_labels"], d["gt_boxes"], d["img_metas"])
    if step == 2:
        b, p, l = cls.forward_mil_head_burn_in_step2(det, n, *args, (fo,))
    elif rotated:
        b, p, l = cls.forward_mil_head_burn_in_step1(det, n, syn_boxes, *args, (fs,), (fo,))
    else:
        b, p, l = cls.forward_mil_head_burn_in_step1(det, n, syn_boxes, *args, (fs,), (fo,), None)
    D.parse_losses(l).backward()
    grads = {k: grad_digest(v.grad) for k, v in head.named_parameters() if v.grad is not None}
    out = dict(rotated=rotated, step=step, seed=seed, stages=stages, cap=DET_CAP, boxes=[t.detach().clone() for t in b],
               points=[t.detach().clone() for t in p], losses={k: v.detach().clone() for k, v in l.items()}, grads=grads,
               feat_grad=None if rotated else grad_digest(fo.grad),
               feat_syn_grad=None if (rotated or step == 2) else grad_digest(fs.grad))
    tag = f"detector_{'obb' if rotated else 'hbb'}_step{step}"
    torch.save(out, os.path.join(OUT, tag + ".pt"))
    print("wrote", tag, {k: round(float(v), 5) for k, v in out["losses"].items()})


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    import sys
    if "--detector" in sys.argv:
        for rot in (False, True):
            for st in (1, 2):
                detector_case(rot, st, seed=3)
        sys.exit(0)
    hbb_case(0, 1, 1, "s1_top1")
    hbb_case(1, 2, 3, "s2_top3")
    overlaps_case()
    obb_case(0, "s1_top3")
    assign_case()
    mask_case()
    pseudo_case()
