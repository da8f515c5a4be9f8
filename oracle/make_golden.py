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


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    hbb_case(0, 1, 1, "s1_top1")
    hbb_case(1, 2, 3, "s2_top3")
    overlaps_case()
    obb_case(0, "s1_top3")
