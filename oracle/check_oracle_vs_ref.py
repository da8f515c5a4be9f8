"""TEST INFRASTRUCTURE ONLY (dev container, needs /root/reference) -- pins the standalone
restatement in oracle/ against the reference's OWN files executed under oracle/ref_shim.py
on identical seeded inputs.  Run: python -m oracle.check_oracle_vs_ref"""
import sys
import time

import torch

from oracle import hbb, ref_shim
from point_teacher_b200 import synth


def _eq(a, b, name, tol=0.0):
    a, b = torch.as_tensor(a), torch.as_tensor(b)
    assert a.shape == b.shape, (name, a.shape, b.shape)
    if a.dtype == torch.bool or not a.dtype.is_floating_point:
        ok = torch.equal(a, b)
        err = 0 if ok else 1
    else:
        err = (a.double() - b.double()).abs().max().item() if a.numel() else 0.0
        ok = err <= tol
    print(f"  {'OK ' if ok else 'BAD'} {name}: max|diff|={err:.3g} (tol {tol})")
    return ok


def check_hbb(seed=0, ext_cfg=None, fine_cfg=None, topk=1, stages=1):
    ns = ref_shim.install()
    d = synth.hbb_batch(seed=seed, num_stages=stages)
    fine_cfg = fine_cfg or synth.HBB_FINE_CFG
    ext_cfg = ext_cfg or synth.HBB_EXT_CFG
    cap = 100
    ok = True
    # ---- reference
    head = ref_shim.build_ref_mil_head(ns, num_stages=stages, top_k=topk, seed=seed)
    P = hbb.MilHeadParams(num_stages=stages, seed=seed)
    sd = head.state_dict()
    for k, v in P.state_dict().items():
        assert torch.equal(sd[k], v), k
    pb = [b[:cap].clone() for b in d["pseudo_boxes"]]
    gb = [b[:cap].clone() for b in d["gt_boxes"]]
    pp = [b[:cap].clone() for b in d["pseudo_points"]]
    pl = [b[:cap].clone() for b in d["pseudo_labels"]]
    x = (d["feat"],)
    t0 = time.perf_counter()
    ref_losses = {}
    with torch.no_grad():
        rpb = pb
        for s in range(stages):
            props, valids, refs, reals = ns.syn.MIL_gen_proposals_from_cfg(pp, rpb, fine_cfg[s], gb, d["img_metas"])
            negs = d["neg_boxes"][s]
            nw = [((ns.bbox_overlaps(negs[i], props[i]) < 0.3).sum(1) == props[i].shape[0]) for i in range(len(negs))]
            l, rpb = head.MIL_head_burn_in_step2(x, d["img_metas"], props, valids, refs, reals, negs, nw,
                                                 rpb, pl, ext_cfg[s], s)
            rpb = list(rpb)
            ref_losses.update(l)
    t_ref = time.perf_counter() - t0
    t0 = time.perf_counter()
    with torch.no_grad():
        ob, op, ol, aux = hbb.phase2_refine(P, x, [d["stride"]], d["img_metas"], d["pseudo_boxes"],
                                            d["pseudo_points"], d["pseudo_labels"], d["gt_boxes"], fine_cfg,
                                            ext_cfg, num_stages=stages, cap=cap, alpha=(1.0, 1.0), topk=topk,
                                            injected_negs=d["neg_boxes"])
    t_or = time.perf_counter() - t0
    print(f"seed {seed}: reference {t_ref:.2f}s, oracle {t_or:.2f}s")
    for i in range(len(pb)):
        ok &= _eq(ob[i][:cap], rpb[i], f"refined boxes img{i}", 0.0)
    for k, v in ref_losses.items():
        ok &= _eq(ol[k], v, k, 1e-6)
    return ok


def run_ref_obb(o, d, head, stages, cap=100):
    """The reference's OWN OBB files: syn_images_generator_v2 + TS_P2RBRotatedFCOSHead (negatives injected)."""
    pb = [b[:cap].clone() for b in d["pseudo_boxes"]]
    gb = [b[:cap].clone() for b in d["gt_boxes"]]
    pp = [b[:cap].clone() for b in d["pseudo_points"]]
    pl = [b[:cap].clone() for b in d["pseudo_labels"]]
    x = (d["feat"],)
    losses, per_stage = {}, []
    with torch.no_grad():
        for s in range(stages):
            props, valids, refs, reals = o.syn.MIL_gen_proposals_from_cfg(pp, pb, synth.OBB_FINE_CFG[s], gb, d["img_metas"])
            negs = d["neg_boxes"][s]
            nw = [((o.rbbox_overlaps(negs[i], props[i]) < 0.3).sum(1) == props[i].shape[0]) for i in range(len(negs))]
            num_gt = sum(b.shape[0] for b in pb)
            R = head.forward_mil_head(num_gt, [b.shape[0] for b in pb], x, props, valids, refs, reals, d["img_metas"],
                                      synth.OBB_EXT_CFG[s], s, negs, nw)
            lb = head.mil_bag_training(R, pl, nw)
            merged = head.mil_bag_selection(R, d["img_metas"], pb, pl)
            per_stage.append(dict(R=R, loss_mil_bags=lb, merged=torch.cat(merged), neg_weight=torch.cat(nw),
                                  base_bags=torch.cat(props), base_valid=torch.cat(valids)))
            pb = list(merged)
            losses[f"stage{s}_loss_mil_bbox"] = R["loss_mil_bbox"]
            losses[f"stage{s}_loss_mil_bags"] = lb
            losses[f"stage{s}_coarse_bags_iou"] = R["coarse_bags_iou"]
            losses[f"stage{s}_refine_bags_iou"] = R["refine_bags_iou"]
    return pb, losses, per_stage


def check_obb(seed=0, **small):
    from oracle import obb
    o = ref_shim.install_obb()
    d = synth.obb_batch(seed=seed, **small)
    head = ref_shim.build_ref_obb_mil_head(o, seed=seed)
    P = hbb.MilHeadParams(num_classes=9, num_stages=1, seed=seed)
    sd = head.state_dict()
    for k, v in P.state_dict().items():
        assert torch.equal(sd[k], v), k
    t0 = time.perf_counter()
    rpb, rl, _ = run_ref_obb(o, d, head, 1)
    t1 = time.perf_counter()
    with torch.no_grad():
        ob, op, ol, aux = obb.phase2_refine(P, (d["feat"],), [d["stride"]], d["img_metas"], d["pseudo_boxes"],
                                            d["pseudo_points"], d["pseudo_labels"], d["gt_boxes"], synth.OBB_FINE_CFG,
                                            synth.OBB_EXT_CFG, alpha=(1.0, 1.0), injected_negs=d["neg_boxes"])
    print(f"OBB seed {seed}: reference {t1 - t0:.2f}s, oracle {time.perf_counter() - t1:.2f}s")
    ok = True
    for i in range(len(rpb)):
        ok &= _eq(ob[i][:100], rpb[i], f"OBB refined boxes img{i}", 0.0)
    for k, v in rl.items():
        ok &= _eq(ol[k], v, "OBB " + k, 1e-6)
    return ok


def check_assign():
    from oracle import assign
    ns = ref_shim.install()
    ok = True
    for seed, ties, (npre, tk) in [(0, True, (3, 3)), (1, True, (1, 1)), (2, True, (5, 3)), (3, False, (5, 3)),
                                   (4, True, (7, 2))]:
        d = synth.assign_batch(seed, ties=ties)
        ref = ns.TopkAssigner(num_pre=npre, topk=tk, cls_cost=dict(type="FocalLossCost", weight=1.0),
                              reg_cost=dict(type="PointCost", mode="L1", weight=1.0))
        r = ref.assign(d["pred"], d["logits"], d["gt"], d["labels"])
        gi, lb = assign.topk_assign(d["pred"], d["logits"], d["gt"], d["labels"], npre, tk)
        ok &= _eq(gi, r.gt_inds, f"TopkAssigner({npre},{tk}) gt_inds seed{seed}", 0.0)
        ok &= _eq(lb, r.labels, f"TopkAssigner({npre},{tk}) labels seed{seed}", 0.0)
        ref = ns.FUSETopkAssigner(num_pre=npre, topk=tk, cls_cost=dict(type="FocalLossCost", weight=1.0),
                                  reg_cost=dict(type="PointCost", mode="L1", weight=1.0),
                                  location_cost=dict(type="InsiderCost", weight=1.0))
        r = ref.assign(d["pred"], d["points"], d["logits"], None, d["gt"], d["labels"])
        gi, lb = assign.fuse_topk_assign(d["pred"], d["points"], d["logits"], d["gt"], d["labels"], npre, tk)
        ok &= _eq(gi, r.gt_inds, f"FUSETopkAssigner({npre},{tk}) gt_inds seed{seed}", 0.0)
        ok &= _eq(lb, r.labels, f"FUSETopkAssigner({npre},{tk}) labels seed{seed}", 0.0)
    # metric matrix + MaxIoU
    g = torch.Generator().manual_seed(9)
    gts = synth.make_boxes(g, 23, (400, 400))
    anchors = torch.cat([synth.jitter_boxes(g, gts.repeat(6, 1), 3.0, 0.4), synth.make_boxes(g, 300, (400, 400))])
    anchors[5] = gts[5]
    anchors[77] = gts[5]                      # exact duplicate maxima: gt_max_assign_all matters
    for mode in ("iou", "iof", "giou", "wd", "kl", "center_distance2", "exp_kl", "kl_10"):
        m = ns.BboxDistanceMetric()(gts, anchors, mode)
        ok &= _eq(assign.bbox_metric(gts, anchors, mode), m, f"BboxDistanceMetric {mode}", 0.0)
    for kw in (dict(pos_iou_thr=0.5, neg_iou_thr=0.4, min_pos_iou=0.0), dict(pos_iou_thr=0.7, neg_iou_thr=0.3, min_pos_iou=0.3),
               dict(pos_iou_thr=0.5, neg_iou_thr=0.5, min_pos_iou=0.0, gt_max_assign_all=False),
               dict(pos_iou_thr=0.5, neg_iou_thr=0.5, match_low_quality=False)):
        for calc, mode in ((dict(type="BboxOverlaps2D"), "iou"), (dict(type="BboxDistanceMetric"), "wd")):
            ref = ns.MaxIoUAssigner(iou_calculator=calc, **kw)
            labels = torch.randint(0, 8, (23,), generator=g)
            ov = ref.iou_calculator(gts, anchors, mode)
            r = ref.assign_wrt_overlaps(ov, labels)
            ours_ov = hbb.bbox_overlaps(gts, anchors, mode) if calc["type"] == "BboxOverlaps2D" else assign.bbox_metric(gts, anchors, mode)
            gi, mx, lb = assign.max_iou_assign(ours_ov, labels, **kw)
            ok &= _eq(gi, r.gt_inds, f"MaxIoU {calc['type']} {kw} gt_inds", 0.0)
            ok &= _eq(mx, r.max_overlaps, "MaxIoU max_overlaps", 0.0)
            ok &= _eq(lb, r.labels, "MaxIoU labels", 0.0)
    return ok


def check_mask():
    """Row a16: the reference's own generate_black_paper (seeded torch + numpy RNG) against
    sample_candidates + black_paper_from_candidates of oracle/mask.py."""
    import numpy as np
    from oracle import mask
    ns = ref_shim.install()
    ok = True
    for seed in range(4):
        d = synth.mask_batch(seed)
        pattern, prior = ns.syn.load_basic_shape(synth.SHAPE_LIST)
        dense = range(int(len(pattern) / 2))
        torch.manual_seed(seed)
        np.random.seed(seed)
        img_ref, bb_ref = ns.syn.generate_black_paper(d["img"].clone(), d["bb_occupied"].clone(), d["img"].clone(), pattern,
                                                      prior, dense, d["imgsize"])
        torch.manual_seed(seed)
        np.random.seed(seed)
        allb = mask.sample_candidates(d["bb_occupied"], prior, dense, d["imgsize"])
        img_o, bb_o, sel, polys, m = mask.black_paper_from_candidates(d["img"].clone(), allb, d["imgsize"])
        ok &= _eq(bb_o, bb_ref, f"black paper kept boxes seed{seed} ({bb_ref.shape[0]} kept of {allb.shape[0]})", 0.0)
        ok &= _eq(img_o, img_ref, f"black paper image seed{seed} ({int(m.sum())} px)", 0.0)
        img_r, _, _, _, m2 = mask.black_paper_from_candidates(d["img"].clone(), allb, d["imgsize"], use_cv2=False)
        ok &= _eq(torch.from_numpy(m2), torch.from_numpy(m), f"fill replay == cv2.fillPoly seed{seed}", 0.0)
    return ok


def check_pseudo():
    """Section 8f rank 1: the reference's own TS_P2BFCOSHead._gnerate_pseudo_single against oracle/assign.py."""
    from oracle import assign
    ns = ref_shim.install()
    ok = True
    for seed, G in ((0, 120), (1, 37), (2, 400)):
        d = synth.pseudo_batch(seed, G=G)
        head = ns.TS_P2BFCOSHead.__new__(ns.TS_P2BFCOSHead)
        torch.nn.Module.__init__(head)
        head.fuse_assigner = ns.FUSETopkAssigner(num_pre=5, topk=3, cls_cost=dict(type="FocalLossCost", weight=1.0),
                                                 reg_cost=dict(type="PointCost", mode="L1", weight=1.0),
                                                 location_cost=dict(type="InsiderCost", weight=1.0))
        rb, rp, rl, rm, rv = head._gnerate_pseudo_single(d["gt_points"], d["labels"], d["gt_boxes"], d["logits"], d["ltrb"],
                                                         None, None, None, 0.1, d["points"], None)
        ob, op, ol, om, ov, _, _ = assign.generate_pseudo_single(d["gt_points"], d["labels"], d["gt_boxes"], d["logits"],
                                                                 d["ltrb"], d["points"], 0.1)
        ok &= _eq(ob, rb, f"pseudo boxes seed{seed}", 0.0)
        ok &= _eq(op, rp, f"pseudo points seed{seed}", 0.0)
        ok &= _eq(om, rm, f"mean iou seed{seed}", 0.0)
        ok &= _eq(torch.sort(ov)[0], torch.sort(rv)[0], f"valid inds seed{seed} ({ov.numel()} of {G})", 0.0)
        # rank 2: the consumer of the refined boxes
        head.num_classes = 8
        head.assigner = ns.TopkAssigner(num_pre=1, topk=1, cls_cost=dict(type="FocalLossCost", weight=1.0),
                                        reg_cost=dict(type="PointCost", mode="L1", weight=1.0))
        head.pseudo_assigner = ns.TopkAssigner(num_pre=3, topk=3, cls_cost=dict(type="FocalLossCost", weight=1.0),
                                               reg_cost=dict(type="PointCost", mode="L1", weight=1.0))
        r = head._get_target_pseudo_single(d["gt_points"], d["labels"], rp, d["labels"], rb, d["logits"], d["ltrb"], None,
                                           dict(ori_filename="x"), None, None, d["points"], None, False)
        o = assign.get_target_pseudo_single(d["points"], d["logits"], d["gt_points"], d["labels"], rb, d["labels"], 8)
        for nm, a_, b_ in zip(("labels_reg", "bbox_targets", "labels", "weights"), o, r):
            ok &= _eq(a_, b_, f"get_target_pseudo {nm} seed{seed}", 0.0)
        pos = (r[0] != 8).nonzero().reshape(-1)
        ok &= _eq(assign.centerness_target(o[1][pos]), head.centerness_target(r[1][pos]), f"centerness seed{seed}", 0.0)
    return ok


def _aug_args(d):
    return [d["img"].clone()] + [[t.clone() for t in d[k]] for k in
                                 ("gt_points", "gt_labels", "pseudo_points", "pseudo_labels", "pseudo_bboxes")]


def check_augment():
    """Section 8f rank 3: the reference's own strong_augmentation (HBB and OBB, seeded ``random`` + ``np.random``)
    against oracle/augment.py fed with the replayed draws; plus the explicit-arithmetic resamplers against the
    library calls they restate."""
    import random

    import numpy as np
    import torch.nn.functional as F
    import torchvision.transforms.functional as TF
    from oracle import augment
    ns = ref_shim.install()
    o = ref_shim.install_obb()
    ok = True
    names = ("images", "image list", "gt points", "gt labels", "pseudo points", "pseudo labels", "pseudo boxes")
    for rot in (False, True):
        for seed in range(8):
            d = synth.augment_batch(seed, rotated=rot)
            random.seed(seed)
            np.random.seed(seed)
            ref = o.syn.strong_augmentation(*_aug_args(d), "le90") if rot else ns.syn.strong_augmentation(*_aug_args(d))
            random.seed(seed)
            np.random.seed(seed)
            ch = augment.draw_choices(2, rotated=rot)
            fn = augment.strong_augmentation_obb if rot else augment.strong_augmentation_hbb
            got = fn(*_aug_args(d), ch)
            for nm, a_, b_ in zip(names, got, ref):
                if torch.is_tensor(a_):
                    ok &= _eq(a_, b_, f"augment {'obb' if rot else 'hbb'} {nm} seed{seed} {ch}", 0.0)
                else:
                    ok &= len(a_) == len(b_)
                    for x_, y_ in zip(a_, b_):
                        ok &= _eq(x_, y_, f"augment {'obb' if rot else 'hbb'} {nm} seed{seed}", 0.0)
    img = synth.augment_batch(0, img_hw=(200, 168))["img"][0]
    for sf in (0.8, 0.9, 1.1, 1.2):
        oh, ow = int(200 * sf), int(168 * sf)
        ok &= _eq(augment.bilinear_resize_exact(img, oh, ow),
                  F.interpolate(img[None], size=(oh, ow), mode="bilinear", align_corners=False)[0], f"bilinear x{sf}", 0.0)
    for ang in (1, 7, 19):
        ok &= _eq(augment.rotate_nearest_exact(img, ang), TF.rotate(img, ang, fill=0), f"rotate {ang} deg", 0.0)
    return ok


def check_losses():
    """Section 8f rank 4: the reference's own FocalLoss (CPU branch) and RotatedIoULoss / DN_IoULoss wrappers (with
    the restated diff_iou_rotated_2d bound into mmcv.ops) against oracle/losses.py."""
    import math
    from oracle import losses as L
    o = ref_shim.install_obb()
    ok = True
    g = torch.Generator().manual_seed(0)
    n = 300
    c = torch.rand(n, 2, generator=g) * 200 + 20
    wh = torch.rand(n, 2, generator=g) * 40 + 2
    b1 = torch.cat([c, wh, torch.rand(n, 1, generator=g) * math.pi - math.pi / 2], 1)
    b2 = b1.clone()
    b2[:, :2] += torch.randn(n, 2, generator=g) * 6
    b2[:, 2:4] *= torch.exp(torch.randn(n, 2, generator=g) * 0.3)
    b2[:, 4] += torch.randn(n, generator=g) * 0.5
    w = (torch.rand(n, generator=g) > 0.3).float()
    for mode in ("log", "linear", "square"):
        ref = o.riou_loss.RotatedIoULoss(mode=mode, loss_weight=1.5)(b1, b2, weight=w, avg_factor=77.0)
        got = L.rotated_loss_forward(L.rotated_iou_loss_elem, b1, b2, weight=w, avg_factor=77.0, loss_weight=1.5, mode=mode)
        ok &= _eq(got, ref, f"RotatedIoULoss {mode}", 0.0)
        ref = o.riou_loss.DN_IoULoss(mode=mode, hyper=0.3)(b1, b2, weight=w, avg_factor=77.0)
        got = L.rotated_loss_forward(L.dn_iou_loss_elem, b1, b2, weight=w, avg_factor=77.0, hyper=0.3, mode=mode)
        ok &= _eq(got, ref, f"DN_IoULoss {mode}", 0.0)
    P, t, ww = torch.randn(500, 8, generator=g) * 2, torch.randint(0, 9, (500,), generator=g), torch.rand(500, generator=g)
    fl = o.focal.FocalLoss(gamma=2.0, alpha=0.25, loss_weight=1.0)
    ok &= _eq(L.sigmoid_focal_loss(P, t, ww, avg_factor=33.0), fl(P, t, ww, avg_factor=33.0), "FocalLoss weighted", 0.0)
    ok &= _eq(L.sigmoid_focal_loss(P, t), fl(P, t), "FocalLoss mean", 0.0)
    ok &= _eq(L.sigmoid_focal_loss(P, t, reduction="none"), fl(P, t, reduction_override="none"), "FocalLoss none", 0.0)
    return ok


def check_detector(rotated=False, step=2, seed=0, stages=1):
    """The reference's OWN detector-level MIL callers (``forward_mil_head_burn_in_step1/2``, run unbound on a stand-in
    ``self`` whose ``student.bbox_head`` is the reference's own head) against oracle/detector.py's literal drivers over
    the oracle backend: refined boxes / points tol 0, losses 1e-6, and the gradients of the summed losses
    (``_parse_losses``) w.r.t. every MIL parameter (and the feature maps for HBB) 1e-5."""
    from oracle import detector as D
    dn = ref_shim.install_detectors()
    if rotated:
        d = synth.obb_batch(seed=seed, batch=2, img_hw=(256, 256), gt_range=(6, 10), n_neg=20)
        head = ref_shim.build_ref_obb_mil_head(dn.obb, num_stages=stages, seed=seed)
        P = hbb.MilHeadParams(num_classes=9, num_stages=stages, seed=seed).requires_grad_(True)
        fine, ext, topk = synth.OBB_FINE_CFG * stages, synth.OBB_EXT_CFG * stages, 3
        cls = dn.RotatedFCOS_TS
    else:
        d = synth.hbb_batch(seed=seed, batch=2, img_hw=(256, 256), gt_range=(6, 10), n_neg=20, num_stages=stages)
        head = ref_shim.build_ref_mil_head(dn.hbb, num_stages=stages, top_k=1, seed=seed)
        P = hbb.MilHeadParams(num_stages=stages, seed=seed).requires_grad_(True)
        fine, ext, topk = synth.HBB_FINE_CFG, synth.HBB_EXT_CFG, 1
        cls = dn.TS_P2B_FCOS
    fine = [dict(c, gen_num_neg=20) for c in fine]
    syn_boxes, feat_syn = D.synthetic_boxes_like(d, seed, rotated)
    n = len(d["pseudo_boxes"])
    cap = 7                                   # fewer than the GTs of every image: the tail must come back untouched

    def run(which):
        fo = d["feat"].clone().requires_grad_(not rotated)
        fs = feat_syn.clone().requires_grad_(not rotated)
        if which == "ref":
            det = D.make_detector(head, fine, ext, stages, cap1=cap, cap2=cap)
            head.zero_grad()
            torch.manual_seed(100 + seed)
            if step == 2:
                out = cls.forward_mil_head_burn_in_step2(det, n, d["pseudo_boxes"], d["pseudo_points"],
                                                         d["pseudo_labels"], d["gt_boxes"], d["img_metas"], (fo,))
            elif rotated:
                out = cls.forward_mil_head_burn_in_step1(det, n, syn_boxes, d["pseudo_boxes"], d["pseudo_points"],
                                                         d["pseudo_labels"], d["gt_boxes"], d["img_metas"], (fs,), (fo,))
            else:
                out = cls.forward_mil_head_burn_in_step1(det, n, syn_boxes, d["pseudo_boxes"], d["pseudo_points"],
                                                         d["pseudo_labels"], d["gt_boxes"], d["img_metas"], (fs,), (fo,), None)
            params = dict(head.named_parameters())
        else:
            for t in P.state_dict().values():
                t.grad = None
            det = D.make_detector(D.OracleHead(P, [d["stride"]], topk, rotated=rotated), fine, ext, stages, cap1=cap, cap2=cap)
            fns = D.oracle_backend(rotated)
            torch.manual_seed(100 + seed)
            if step == 2:
                out = D.forward_mil_head_burn_in_step2(det, fns, n, d["pseudo_boxes"], d["pseudo_points"],
                                                       d["pseudo_labels"], d["gt_boxes"], d["img_metas"], (fo,))
            else:
                out = D.forward_mil_head_burn_in_step1(det, fns, n, syn_boxes, d["pseudo_boxes"], d["pseudo_points"],
                                                       d["pseudo_labels"], d["gt_boxes"], d["img_metas"], (fs,), (fo,))
            params = P.state_dict()
        D.parse_losses(out[2]).backward()
        grads = {k: (None if params[k].grad is None else params[k].grad.clone()) for k in P.state_dict()}
        return out, grads, fo.grad, fs.grad

    (rb, rp, rl), rg, rfo, rfs = run("ref")
    (ob, op, ol), og, ofo, ofs = run("oracle")
    tag = f"{'OBB' if rotated else 'HBB'} step{step}"
    ok = True
    for i in range(n):
        ok &= _eq(ob[i], rb[i], f"{tag} boxes img{i}", 0.0)
        ok &= _eq(op[i], rp[i], f"{tag} points img{i}", 0.0)
        ok &= bool(torch.equal(rb[i][cap:], d["pseudo_boxes"][i][cap:]))
    assert set(rl) == set(ol), (sorted(rl), sorted(ol))
    for k in rl:
        ok &= _eq(ol[k].detach(), rl[k].detach(), f"{tag} {k}", 1e-6)
    for k in rg:
        if rg[k] is None or og[k] is None:
            ok &= rg[k] is None and og[k] is None
            continue
        ok &= _eq(og[k], rg[k], f"{tag} grad {k}", 1e-5)
    if not rotated:
        ok &= _eq(ofo, rfo, f"{tag} grad feat_ori", 1e-5)
        if step == 1:
            ok &= _eq(ofs, rfs, f"{tag} grad feat_syn", 1e-5)
    return ok


if __name__ == "__main__":
    if "--detector" in sys.argv:
        good = True
        for rot in (False, True):
            for st in (2, 1):
                good &= check_detector(rot, st, seed=3)
        good &= check_detector(False, 1, seed=4, stages=2)
        print("ALL OK" if good else "MISMATCH")
        sys.exit(0 if good else 1)
    if "--losses" in sys.argv:
        good = check_losses()
        print("ALL OK" if good else "MISMATCH")
        sys.exit(0 if good else 1)
    if "--augment" in sys.argv:
        good = check_augment()
        print("ALL OK" if good else "MISMATCH")
        sys.exit(0 if good else 1)
    if "--pseudo" in sys.argv:
        good = check_pseudo()
        print("ALL OK" if good else "MISMATCH")
        sys.exit(0 if good else 1)
    if "--mask" in sys.argv:
        good = check_mask()
        print("ALL OK" if good else "MISMATCH")
        sys.exit(0 if good else 1)
    if "--assign" in sys.argv:
        good = check_assign()
        print("ALL OK" if good else "MISMATCH")
        sys.exit(0 if good else 1)
    if "--obb" in sys.argv:
        good = check_obb(0, batch=2, img_hw=(512, 512), gt_range=(20, 30), n_neg=40)
        print("ALL OK" if good else "MISMATCH")
        sys.exit(0 if good else 1)
    good = check_hbb(0)
    good &= check_hbb(1, topk=3)
    good &= check_hbb(2, stages=2)
    print("ALL OK" if good else "MISMATCH")
    sys.exit(0 if good else 1)
