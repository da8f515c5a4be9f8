"""Label assignment kernels on a B200: indices bit-exact against the CPU oracle and
the golden vectors produced by the reference's own files, including the ATen CPU top-k tie rule."""
import os

import pytest
import torch

from oracle import assign, hbb
from point_teacher_b200 import synth

pytestmark = pytest.mark.gpu


def _assigners():
    from point_teacher_b200 import assigners
    return assigners


def _mk_topk(num_pre, topk):
    return _assigners().TopkAssigner(num_pre=num_pre, topk=topk, cls_cost=dict(type="FocalLossCost", weight=1.0),
                                     reg_cost=dict(type="PointCost", mode="L1", weight=1.0))


def _mk_fuse(num_pre, topk, mode="L1"):
    return _assigners().FUSETopkAssigner(num_pre=num_pre, topk=topk, cls_cost=dict(type="FocalLossCost", weight=1.0),
                                         reg_cost=dict(type="PointCost", mode=mode, weight=1.0),
                                         location_cost=dict(type="InsiderCost", weight=1.0))


def test_assigners_against_reference_golden(cuda, golden_dir):
    g = torch.load(os.path.join(golden_dir, "assign.pt"))
    for c in g["cases"]:
        d = {k: v.to(cuda) for k, v in synth.assign_batch(c["seed"], ties=c["ties"]).items()}
        r = _mk_topk(c["num_pre"], c["topk"]).assign(d["pred"], d["logits"], d["gt"], d["labels"])
        assert torch.equal(r.gt_inds.cpu(), c["topk_gt_inds"]), c["seed"]
        assert torch.equal(r.labels.cpu(), c["topk_labels"])
        r = _mk_fuse(c["num_pre"], c["topk"]).assign(d["pred"], d["points"], d["logits"], None, d["gt"], d["labels"])
        assert torch.equal(r.gt_inds.cpu(), c["fuse_gt_inds"]), c["seed"]
        assert torch.equal(r.labels.cpu(), c["fuse_labels"])
        assert r.num_gts == d["gt"].shape[0] and r.max_overlaps is None
        from point_teacher_b200 import ops
        t = ops.focal_cost_table(d["logits"]).cpu()
        assert (t - c["fl_table"]).abs().max() <= 5e-6 * c["fl_table"].abs().max()      # expf/logf: ulp-level
    A = _assigners()
    gts, anchors, labels = g["gts"].to(cuda), g["anchors"].to(cuda), g["labels"].to(cuda)
    for m, ref in g["metric"].items():
        got = A.BboxDistanceMetric()(gts, anchors, m).cpu()
        if m in ("kl", "exp_kl", "kl_10"):                      # logf / expf: ulp-level differences
            assert (got - ref).abs().max() <= 1e-5 * ref.abs().max(), m
        else:
            assert torch.equal(got, ref), m
    for c in g["maxiou"]:
        a = A.MaxIoUAssigner(iou_calculator=dict(type=c["calc"]), **c["kw"])
        r = a.assign(anchors, gts, gt_labels=labels, mode=c["mode"])
        assert torch.equal(r.gt_inds.cpu(), c["gt_inds"]), c
        assert torch.equal(r.max_overlaps.cpu(), c["max_overlaps"])
        assert torch.equal(r.labels.cpu(), c["labels"])


@pytest.mark.parametrize("P_hw,G,num_pre,topk,ties,mode", [
    ((100, 100), 150, 5, 3, True, "L1"),          # shipped FUSE setting on the 800x800 grid, tie-heavy integer GTs
    ((100, 100), 150, 3, 3, True, "L1"),          # shipped pseudo / syn assigners: second stage never ranks
    ((100, 100), 150, 1, 1, False, "L1"),
    ((128, 128), 400, 7, 2, True, "L2"),
    ((20, 20), 30, 9, 4, True, "L1"),             # num_pre*64 > P: ATen's nth_element path on the whole column
    ((100, 100), 1500, 5, 3, True, "L1"),         # stress config #4
])
def test_two_stage_topk_vs_oracle(cuda, P_hw, G, num_pre, topk, ties, mode):
    from point_teacher_b200 import ops
    d = synth.assign_batch(100 + G + num_pre, P_hw=P_hw, G=G, ties=ties)
    dc = {k: v.to(cuda) for k, v in d.items()}
    # stage 1 indices, column by column, against torch.topk on the CPU (the reference's tie rule)
    cost = assign.point_cost(d["points"], d["gt"], mode)
    ref_pre = torch.topk(cost, num_pre, dim=0, largest=False).indices
    pre = ops.topk_pre(dc["points"], dc["gt"], num_pre, mode)
    assert torch.equal(pre.cpu().long(), ref_pre)
    tie_cols = (torch.sort(cost, 0).values[num_pre - 1] == torch.sort(cost, 0).values[num_pre]).float().mean().item()
    if ties and mode == "L1":
        assert tie_cols > 0.1               # the test really exercises boundary ties
    # full assignment; the device-computed focal table can differ from the CPU one by an ulp, so feed the oracle
    # the device table (identical fp32 costs) and require bit-exact indices
    table = ops.focal_cost_table(dc["logits"])
    cost2 = table.cpu()[:, d["labels"]] + assign.insider_cost(d["pred"], d["gt"])
    gi_ref, lb_ref = assign._two_stage(cost, cost2, d["labels"], num_pre, topk)
    fuse = _assigners().FUSETopkAssigner(num_pre=num_pre, topk=topk, cls_cost=dict(type="FocalLossCost", weight=1.0),
                                         reg_cost=dict(type="PointCost", mode=mode, weight=1.0),
                                         location_cost=dict(type="InsiderCost", weight=1.0))
    r = fuse.assign(dc["pred"], dc["points"], dc["logits"], None, dc["gt"], dc["labels"])
    assert torch.equal(r.gt_inds.cpu(), gi_ref)
    assert torch.equal(r.labels.cpu(), lb_ref)
    # and with the CPU-computed table the agreement must still be >= 99.9 % of the points
    gi_cpu, _ = assign.fuse_topk_assign(d["pred"], d["points"], d["logits"], d["gt"], d["labels"], num_pre, topk,
                                        reg_mode=mode)
    assert (r.gt_inds.cpu() == gi_cpu).float().mean().item() >= 0.999


def test_topk_assigner_edge_cases(cuda):
    A = _assigners()
    d = {k: v.to(cuda) for k, v in synth.assign_batch(5, P_hw=(20, 20), G=6).items()}
    a = _mk_topk(3, 3)
    r = a.assign(d["pred"], d["logits"], d["gt"][:0], d["labels"][:0])          # no GT: all background
    assert r.num_gts == 0 and (r.gt_inds == 0).all() and (r.labels == -1).all()
    r = a.assign(d["pred"], d["logits"], None, None)
    assert r.num_gts == 0 and (r.gt_inds == 0).all()
    with pytest.raises(RuntimeError):                                            # torch.topk: k out of range
        _mk_topk(16, 3).assign(d["pred"][:10].contiguous(), d["logits"][:10].contiguous(), d["gt"], d["labels"])
    with pytest.raises(ValueError):
        a.assign(d["pred"].cpu(), d["logits"].cpu(), d["gt"].cpu(), d["labels"].cpu())
    m = A.MaxIoUAssigner(0.5, 0.5)
    r = m.assign(d["pred"], d["gt"][:0])
    assert (r.gt_inds == 0).all() and r.labels is None
    # known answer of the reference's own test (HBB_TOD/tests/test_utils/test_assigner.py:15-63)
    bboxes = torch.tensor([[0, 0, 10, 10], [10, 10, 20, 20], [5, 5, 15, 15], [32, 32, 38, 42]], dtype=torch.float32, device=cuda)
    gts = torch.tensor([[0, 0, 10, 9], [0, 10, 10, 19]], dtype=torch.float32, device=cuda)
    r = m.assign(bboxes, gts, gt_labels=torch.tensor([2, 3], device=cuda))
    assert r.gt_inds.tolist() == [1, 0, 2, 0] and r.labels.tolist() == [2, -1, 3, -1]


@pytest.mark.parametrize("calc,mode", [("BboxOverlaps2D", "iou"), ("BboxOverlaps2D", "giou"), ("BboxDistanceMetric", "wd"),
                                       ("BboxDistanceMetric", "iou")])
@pytest.mark.parametrize("G,A", [(100, 10000), (1500, 16384)])
def test_max_iou_assign_vs_oracle_dense(cuda, calc, mode, G, A):
    """Config #4 sweep shapes: every decision (argmax with first-index ties, thresholds, low-quality equality
    matching) bit-exact against the oracle that materialises the G x A matrix."""
    g = torch.Generator().manual_seed(G + A)
    gts = synth.make_boxes(g, G, (800, 800))
    rep = synth.jitter_boxes(g, gts.repeat(4, 1), 2.0, 0.3)
    anchors = torch.cat([rep, synth.make_boxes(g, A - rep.shape[0], (800, 800), median=16)])[:A].contiguous()
    anchors[:50] = gts[:50]                              # exact-duplicate maxima
    anchors[50:100] = gts[:50]
    labels = torch.randint(0, 8, (G,), generator=g)
    ov = hbb.bbox_overlaps(gts, anchors, mode) if calc == "BboxOverlaps2D" else assign.bbox_metric(gts, anchors, mode)
    Amod = _assigners()
    for kw in (dict(pos_iou_thr=0.5, neg_iou_thr=0.4, min_pos_iou=0.2), dict(pos_iou_thr=0.6, neg_iou_thr=(0.1, 0.4),
                                                                           gt_max_assign_all=False)):
        gi, mx, lb = assign.max_iou_assign(ov, labels, **kw)
        a = Amod.MaxIoUAssigner(iou_calculator=dict(type=calc), **kw)
        r = a.assign(anchors.to(cuda), gts.to(cuda), gt_labels=labels.to(cuda), mode=mode)
        assert torch.equal(r.gt_inds.cpu(), gi)
        assert torch.equal(r.max_overlaps.cpu(), mx)
        assert torch.equal(r.labels.cpu(), lb)
    # the matrix itself, when the caller asks for it
    m = Amod.BboxOverlaps2D()(gts.to(cuda), anchors.to(cuda), mode) if calc == "BboxOverlaps2D" else \
        Amod.BboxDistanceMetric()(gts.to(cuda), anchors.to(cuda), mode)
    assert torch.equal(m.cpu(), ov)


@pytest.mark.parametrize("assign_all", [True, False])
def test_max_iou_ties_across_gt_tiles(cuda, assign_all):
    """The device path tiles the GTs by 256 and merges tiles with integer atomics: duplicated GTs in DIFFERENT tiles
    must still resolve like the reference (per-anchor argmax = FIRST index among equal maxima; low-quality matching =
    LAST matching GT wins), and anchors equal to a GT give exact-equality maxima."""
    g = torch.Generator().manual_seed(91)
    G, A = 700, 3000
    gts = synth.make_boxes(g, G, (800, 800))
    gts[300:340] = gts[10:50]                    # duplicates one tile later
    gts[600:620] = gts[10:30]                    # and two tiles later
    anchors = torch.cat([synth.jitter_boxes(g, gts.repeat(4, 1), 2.0, 0.3), synth.make_boxes(g, A - 4 * G, (800, 800))])
    anchors[:60] = gts[:60]                      # IoU == 1 with three different GT indices
    labels = torch.randint(0, 8, (G,), generator=g)
    for calc, mode in (("BboxOverlaps2D", "iou"), ("BboxOverlaps2D", "giou"), ("BboxDistanceMetric", "wd")):
        ov = hbb.bbox_overlaps(gts, anchors, mode) if calc == "BboxOverlaps2D" else assign.bbox_metric(gts, anchors, mode)
        kw = dict(pos_iou_thr=0.5, neg_iou_thr=0.4, min_pos_iou=0.0, gt_max_assign_all=assign_all)
        gi, mx, lb = assign.max_iou_assign(ov, labels, **kw)
        r = _assigners().MaxIoUAssigner(iou_calculator=dict(type=calc), **kw).assign(
            anchors.to(cuda), gts.to(cuda), gt_labels=labels.to(cuda), mode=mode)
        assert torch.equal(r.gt_inds.cpu(), gi), (calc, mode)
        assert torch.equal(r.max_overlaps.cpu(), mx)
        assert torch.equal(r.labels.cpu(), lb)


def test_coarse_pseudo_boxes_vs_reference_golden_and_oracle(cuda, golden_dir):
    """Section 8f rank 1 (_gnerate_pseudo_single): FUSE assignment + score-weighted box aggregation."""
    from point_teacher_b200 import coarse
    fuse = _mk_fuse(5, 3)
    for c in torch.load(os.path.join(golden_dir, "pseudo_boxes.pt")):
        d = synth.pseudo_batch(c["seed"], G=c["G"])
        dc = {k: v.to(cuda) for k, v in d.items()}
        b, p, lab, miou, valid = coarse.generate_pseudo_single(fuse, dc["gt_points"], dc["labels"], dc["gt_boxes"],
                                                               dc["logits"], dc["ltrb"], None, None, None, 0.1,
                                                               dc["points"], None)
        # sums of <= 5 score-weighted boxes per GT: fp32 summation order is the only difference
        assert (b.cpu() - c["boxes"]).abs().max() <= 1e-5 * c["boxes"].abs().max()
        assert (p.cpu() - c["points"]).abs().max() <= 1e-5 * c["points"].abs().max()
        assert abs(float(miou) - float(c["mean_iou"])) < 1e-5
        assert torch.equal(valid.cpu(), c["valid"])
        assert torch.equal(lab, dc["labels"])
    # no GT: the reference's empty return
    b, p, lab, miou, valid = coarse.generate_pseudo_single(fuse, dc["gt_points"][:0], dc["labels"][:0], dc["gt_boxes"][:0],
                                                           dc["logits"], dc["ltrb"], None, None, None, 0.1, dc["points"])
    assert b.shape == (0, 4) and p.shape == (0, 2) and valid is None and miou == 0.0
    # a GT far away from every point keeps the 8 x 8 box at its point and is never valid
    d = synth.pseudo_batch(3, G=10)
    dc = {k: v.to(cuda) for k, v in d.items()}
    b, p, _, _, valid = coarse.generate_pseudo_single(fuse, dc["gt_points"], dc["labels"], dc["gt_boxes"], dc["logits"],
                                                      dc["ltrb"], None, None, None, 2.0, dc["points"])
    assert valid.numel() == 0                                   # score filter above any sigmoid


def test_target_pseudo_vs_reference_golden(cuda, golden_dir):
    """Section 8f rank 2 (_get_target_pseudo_single + centerness_target): labels bit-exact, ltrb targets bit-exact."""
    from point_teacher_b200 import coarse, ops
    a, pa = _mk_topk(1, 1), _mk_topk(3, 3)
    for c in torch.load(os.path.join(golden_dir, "pseudo_boxes.pt")):
        d = {k: v.to(cuda) for k, v in synth.pseudo_batch(c["seed"], G=c["G"]).items()}
        boxes = c["boxes"].to(cuda)
        lr, t, lb, w = coarse.get_target_pseudo_single(a, pa, 8, d["gt_points"], d["labels"], c["points"].to(cuda),
                                                       d["labels"], boxes, d["logits"], d["ltrb"], None, None, None, None,
                                                       d["points"])
        pos = c["pos"].long()
        assert torch.equal(lr.cpu(), c["labels_reg"].long()) and torch.equal(lb.cpu(), c["labels"].long())
        assert torch.equal(t.cpu()[pos], c["bbox_targets_pos"])
        assert float(t.cpu().double().sum()) == float(c["targets_checksum"])
        assert (w == 1).all()
        cen = coarse.centerness_target(t[pos.to(cuda)])
        assert (cen.cpu() - c["centerness_pos"]).abs().max() < 1e-6
        # fused variant: centerness for every point from the same kernel
        res = pa.assign(d["points"], d["logits"], torch.cat([(boxes[:, :2] + boxes[:, 2:]) / 2, boxes[:, 2:] - boxes[:, :2]], 1),
                        d["labels"])
        _, _, cen_all = ops.ltrb_targets(d["points"], boxes, res.gt_inds, res.labels, 8, want_centerness=True)
        assert (cen_all.cpu()[pos] - c["centerness_pos"]).abs().max() < 1e-6
    # no pseudo boxes: the reference's early return
    lr, t, lb, w = coarse.get_target_pseudo_single(a, pa, 8, d["gt_points"], d["labels"], d["gt_points"][:0],
                                                   d["labels"][:0], boxes[:0], d["logits"], d["ltrb"], None, None, None,
                                                   None, d["points"])
    assert (lr == 8).all() and torch.count_nonzero(t) == 0
