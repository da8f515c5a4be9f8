"""Parity ON THE CONFIGURATIONS THE BENCH TIMES (BASELINE.json configs[0..3]), through the captured CUDA graph
(``CapturedPhase2.replay()``: fp16 NHWC feature map, 160-tile FC1 with the split-K tail, graph replay) and in both
precisions, against the CPU oracle on identical seeded inputs:
  * bag geometry and validity flags bit-exact,
  * RoI-refined bags, MIL scores, refined boxes / points and every loss entry within 1e-3 (fp32) / 2e-2 (bf16),
  * selected-instance agreement >= 99.9 % in fp32 precision; the bf16 figure is MEASURED (94-95 % top-1 HBB, 90 %
    ordered top-3 OBB with random-init heads), asserted against a floor, every pick asserted tolerance-consistent
    (the oracle's own score of the picked instance is within 2e-2 of its best), and written to
    gpurun_out/parity_r02.json (copied to profiles/ and quoted in DESIGN.md and the bench line).
In bf16 a top-k pick that flips on a 1e-3 score difference moves that GT's merged box by pixels; refined boxes are
therefore compared on the GTs whose selection agrees and the flips are counted, not hidden."""
import json
import os

import pytest
import torch

from oracle import hbb, obb
from point_teacher_b200 import synth

pytestmark = pytest.mark.gpu
TOL = {"fp32": 1e-3, "bf16": 2e-2}
RECORD = {}


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def _record(key, **kv):
    RECORD[key] = kv
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    os.makedirs(out, exist_ok=True)
    path = os.path.join(out, "parity_r02.json")
    cur = json.load(open(path)) if os.path.exists(path) else {}
    cur[key] = kv
    json.dump(cur, open(path, "w"), indent=1, sort_keys=True)


def _captured(cuda, head, d, fine, ext, cap):
    from point_teacher_b200.refine import CapturedPhase2
    to = lambda l: [t.to(cuda) for t in l]  # noqa: E731
    inputs = dict(feat=d["feat"].to(cuda), pseudo_boxes=to(d["pseudo_boxes"]), pseudo_points=to(d["pseudo_points"]),
                  pseudo_labels=to(d["pseudo_labels"]), gt_boxes=to(d["gt_boxes"]),
                  neg_boxes=[to(n) for n in d["neg_boxes"]] if d.get("neg_boxes") else None)
    c = CapturedPhase2(head, inputs, d["img_metas"], fine, ext, num_stages=1, cap=cap, refresh_weights=True)
    for _ in range(2):
        out = c.replay()
    torch.cuda.synchronize()
    return out, c


def _compare(tag, precision, head, out, oracle_out, box_dim, cap, labels):
    gb, gp, gl = out
    ob, op, ol, aux = oracle_out
    tol = TOL[precision]
    R, ref = head.last_results, aux[-1]
    assert torch.equal(R["_b200"]["coarse"][:, 1:1 + box_dim].cpu(), torch.cat(ref["coarse_extensive_bags"]))   # bit-exact
    assert torch.equal(R["_b200"]["evalid"].bool().cpu().reshape(-1, 1), torch.cat(ref["extensive_bags_valid"]))
    errs = dict(cls=_rel(R["cls_score"], ref["cls_score"]), ins=_rel(R["ins_score"], ref["ins_score"]),
                neg_cls=_rel(R["neg_cls_score"], ref["neg_cls_score"]),
                refined_bags=_rel(torch.cat(R["extensive_bags"]), torch.cat(ref["extensive_bags"])))
    for k, e in errs.items():
        assert e < tol, (tag, precision, k, e)
    for k in ol:
        assert abs(float(gl[k]) - float(ol[k])) <= tol * max(abs(float(ol[k])), 1e-3), (tag, k, float(gl[k]), float(ol[k]))
    sel, sel_ref = R["_b200"]["sel_idx"].cpu().long(), ref["selected_idx"]
    same = (sel == sel_ref).all(1)
    agree = same.float().mean().item()
    # refined boxes / points: all of them in fp32; in bf16 the GTs whose pick agrees (flips are counted, see above)
    m_g = torch.cat([b[:cap] for b in gb]).cpu()
    m_o = torch.cat([b[:cap] for b in ob])
    p_g, p_o = torch.cat([p[:cap] for p in gp]).cpu(), torch.cat([p[:cap] for p in op])
    mask = same if precision == "bf16" else torch.ones_like(same)
    scale = m_o.abs().max()
    box_err = ((m_g - m_o).abs().max(1).values[mask].max() / scale).item()
    pt_err = ((p_g - p_o).abs().max(1).values[mask].max() / scale).item()
    assert box_err < tol and pt_err < tol, (tag, precision, box_err, pt_err)
    for i in range(len(gb)):                                    # untouched tail beyond the cap: bit-exact
        assert torch.equal(gb[i][cap:].cpu(), ob[i][cap:])
    # every pick -- flipped or not -- must be tolerance-consistent: the ORACLE's score of the instance the GPU picked
    # is within ``tol`` (relative to the bag's best score) of the oracle's own j-th best.  A flip is then by
    # construction a choice between instances the reference itself separates by less than the stated tolerance
    # (random-init heads score the 25 heavily overlapping instances of a bag almost identically).
    G, U1, U2, C = ref["cls_score"].shape
    valid = torch.cat(ref["extensive_bags_valid"], 0).reshape(G, U1, U2, 1)
    s_or = (ref["cls_score"].sigmoid() * hbb._instance_scores(ref["ins_score"], valid)).reshape(G, U1 * U2, C)
    s_or = s_or[torch.arange(G), :, labels]
    best = s_or.topk(sel.shape[1], dim=1).values
    gap = ((best - s_or.gather(1, sel)) / best[:, :1].clamp_min(1e-12)).abs().max().item()
    assert gap < tol, (tag, precision, gap)
    same_set = (sel.sort(1).values == sel_ref.sort(1).values).all(1).float().mean().item()
    if precision == "fp32":
        assert agree >= 0.999, (tag, agree)
    else:
        assert agree >= 0.85, (tag, agree)
    flipped_shift = float((m_g - m_o).abs().max(1).values[~same].max()) if (~same).any() else 0.0
    _record(f"{tag}/{precision}", selected_instance_agreement=agree, gts=int(same.numel()), flips=int((~same).sum()),
            max_box_shift_px_on_flips=flipped_shift, selected_set_agreement=same_set,
            max_oracle_score_gap_of_any_pick=gap, refined_box_rel_err=box_err, **{f"{k}_rel_err": v for k, v in errs.items()},
            losses_rel_err=max(abs(float(gl[k]) - float(ol[k])) / max(abs(float(ol[k])), 1e-3) for k in ol))
    return agree


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("seed", [0, 1, 2, 3, 4])
def test_cfg1_hbb_800_through_captured_graph(cuda, precision, seed):
    """BASELINE config #1 (the bench workload): 2 images 800x800, 200-600 GT/img capped at 100, K = 5000 + 400 neg."""
    from point_teacher_b200.mil_head import MILHead
    d = synth.hbb_batch(seed=seed)
    P = hbb.MilHeadParams(num_stages=1, seed=seed)
    head = MILHead(num_classes=8, num_stages=1, top_k=1, precision=precision).to(cuda)
    head.load_state_dict(P.state_dict(), strict=False)
    out, cap = _captured(cuda, head, d, synth.HBB_FINE_CFG, synth.HBB_EXT_CFG, 100)
    with torch.no_grad():
        ref = hbb.phase2_refine(P, (d["feat"],), [d["stride"]], d["img_metas"], d["pseudo_boxes"], d["pseudo_points"],
                                d["pseudo_labels"], d["gt_boxes"], synth.HBB_FINE_CFG, synth.HBB_EXT_CFG, num_stages=1,
                                cap=100, topk=1, injected_negs=d["neg_boxes"])
    assert head.last_results["_b200"]["K"] == 5000
    _compare(f"cfg1_hbb_800_seed{seed}", precision, head, out, ref, 4, 100,
             torch.cat([l[:100] for l in d["pseudo_labels"]]))


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_cfg3_obb_1024_through_captured_graph(cuda, precision):
    """BASELINE config #3: OBB, 2 images 1024x1024, 100 GT/img, RoIAlignRotated(2, clockwise), top-3 merge."""
    from point_teacher_b200.mil_head import RotatedMILHead
    d = synth.obb_batch(seed=0)
    P = hbb.MilHeadParams(num_classes=9, num_stages=1, seed=0)
    head = RotatedMILHead(num_classes=9, num_stages=1, top_k=3, precision=precision).to(cuda)
    head.load_state_dict(P.state_dict(), strict=False)
    out, cap = _captured(cuda, head, d, synth.OBB_FINE_CFG, synth.OBB_EXT_CFG, 100)
    with torch.no_grad():
        ref = obb.phase2_refine(P, (d["feat"],), [d["stride"]], d["img_metas"], d["pseudo_boxes"], d["pseudo_points"],
                                d["pseudo_labels"], d["gt_boxes"], synth.OBB_FINE_CFG, synth.OBB_EXT_CFG,
                                injected_negs=d["neg_boxes"])
    _compare("cfg3_obb_1024_seed0", precision, head, out, ref, 5, 100,
             torch.cat([l[:100] for l in d["pseudo_labels"]]))


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_cfg4_stress_1500x64_subsample_vs_oracle(cuda, precision):
    """BASELINE config #4: 1500 GT/image x bag of 64 (192 000 RoIs per pass over 2 images) on the GPU through the
    captured graph; bags are independent per GT, so the oracle refines a 50-GT subsample (25 per image, spread over
    the whole index range incl. first and last) from the same feature map and every per-GT output is compared."""
    from point_teacher_b200.mil_head import MILHead
    G = 1500
    d = synth.hbb_batch(seed=6, gt_range=(G, G))
    ext = synth.stress_ext_cfg(8)
    fine = [dict(synth.HBB_FINE_CFG[0], gen_num_neg=0)]
    d["neg_boxes"] = None
    P = hbb.MilHeadParams(num_stages=1, seed=6)
    head = MILHead(num_classes=8, num_stages=1, top_k=1, precision=precision).to(cuda)
    head.load_state_dict(P.state_dict(), strict=False)
    (gb, gp, gl), cap = _captured(cuda, head, d, fine, ext, G)
    b = head.last_results["_b200"]
    assert b["U2"] == 64 and b["K"] == 2 * G * 64
    pick = torch.linspace(0, G - 1, 25).round().long()
    sub = lambda l: [t[pick] for t in l]  # noqa: E731
    with torch.no_grad():
        ob, op, ol, aux = hbb.phase2_refine(P, (d["feat"],), [d["stride"]], d["img_metas"], sub(d["pseudo_boxes"]),
                                            sub(d["pseudo_points"]), sub(d["pseudo_labels"]), sub(d["gt_boxes"]), fine,
                                            ext, num_stages=1, cap=G, topk=1)
    tol = TOL[precision]
    ref = aux[-1]
    rows = torch.cat([pick, pick + G])                             # GT rows of the packed GPU result
    inst = (rows[:, None] * 64 + torch.arange(64)[None]).reshape(-1)
    assert torch.equal(b["coarse"][:, 1:5].cpu()[inst], torch.cat(ref["coarse_extensive_bags"]))
    assert torch.equal(b["evalid"].bool().cpu()[inst].reshape(-1, 1), torch.cat(ref["extensive_bags_valid"]))
    R = head.last_results
    assert _rel(R["cls_score"].cpu()[rows], ref["cls_score"]) < tol
    assert _rel(R["ins_score"].cpu()[rows], ref["ins_score"]) < tol
    assert _rel(b["refined"][:, 1:5].cpu()[inst], torch.cat(ref["extensive_bags"])) < tol
    sel = b["sel_idx"].cpu().long()[rows]
    same = (sel == ref["selected_idx"]).all(1)
    agree = same.float().mean().item()
    m_g = torch.cat([gb[0].cpu()[pick], gb[1].cpu()[pick]])
    m_o = torch.cat(ob)
    mask = same if precision == "bf16" else torch.ones_like(same)
    err = ((m_g - m_o).abs().max(1).values[mask].max() / m_o.abs().max()).item()
    assert err < tol, err
    assert agree >= (0.98 if precision == "fp32" else 0.9), agree          # 50 GTs: one flip = 2 %
    _record(f"cfg4_stress_1500x64_sub50/{precision}", selected_instance_agreement=agree, gts=50, flips=int((~same).sum()),
            refined_box_rel_err=err)
    for k, v in gl.items():
        assert torch.isfinite(v).all(), k


def test_feature_cache_is_keyed_on_tensor_identity(cuda):
    """two successive steps with DIFFERENT backbone outputs, the first one freed so that the
    caching allocator hands the same address to the second -- the eager path must pool from the new features."""
    from point_teacher_b200.mil_head import MILHead
    from point_teacher_b200.refine import phase2_refine
    small = dict(batch=2, img_hw=(256, 256), gt_range=(6, 10), n_neg=20)
    d = synth.hbb_batch(seed=21, **small)
    d2 = synth.hbb_batch(seed=22, **small)
    P = hbb.MilHeadParams(num_stages=1, seed=21)
    head = MILHead(num_classes=8, num_stages=1, top_k=1, precision="fp32").to(cuda)
    head.load_state_dict(P.state_dict(), strict=False)
    to = lambda l: [t.to(cuda) for t in l]  # noqa: E731

    def run(feat_cpu):
        x = feat_cpu.to(cuda)
        ptr = x.data_ptr()
        with torch.no_grad():
            out = phase2_refine(head, (x,), d["img_metas"], to(d["pseudo_boxes"]), to(d["pseudo_points"]),
                                to(d["pseudo_labels"]), to(d["gt_boxes"]), synth.HBB_FINE_CFG, synth.HBB_EXT_CFG,
                                neg_boxes=[to(d["neg_boxes"][0])])
        torch.cuda.synchronize()
        return [b.clone() for b in out[0]], ptr
    b1, ptr1 = run(d["feat"])
    b2, ptr2 = run(d2["feat"])                       # x of the first call is dead: same shape, usually the same address
    with torch.no_grad():
        o2 = hbb.phase2_refine(P, (d2["feat"],), [8], d["img_metas"], d["pseudo_boxes"], d["pseudo_points"],
                               d["pseudo_labels"], d["gt_boxes"], synth.HBB_FINE_CFG, synth.HBB_EXT_CFG,
                               injected_negs=d["neg_boxes"])[0]
    assert not torch.equal(b1[0], b2[0])
    for a, b in zip(b2, o2):
        assert _rel(a, b) < 1e-3
    if ptr1 != ptr2:
        pytest.skip("allocator did not reuse the address this time (results still checked against the oracle)")


def test_fp16_feature_map_saturation_is_reported(cuda):
    """Values beyond the fp16 range are counted by the transpose kernel; the eager path falls back to bf16 maps with a
    warning, a captured step raises (it cannot switch dtype)."""
    import warnings
    from point_teacher_b200.roi_extractors import FeatureRangeError, RoIAlign
    layer = RoIAlign(7, spatial_scale=0.125)
    x = torch.randn(1, 64, 16, 16, device=cuda)
    f = layer.nhwc(x, torch.float16)
    assert f.dtype == torch.float16 and layer._cache.check(sync=True) == 0
    y = x.clone()
    y[0, 3, 2, 5] = 7.0e4
    y[0, 9, 0, 0] = -float("inf")
    f = layer.nhwc(y, torch.float16)                    # saturates two values; reported on the NEXT call
    assert float(f[0, 2, 5, 3]) == 65504.0
    torch.cuda.synchronize()
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        z = x.clone()
        f = layer.nhwc(z, torch.float16)
    assert any("fp16 range" in str(i.message) for i in w)
    assert f.dtype == torch.bfloat16 and layer._cache.saturated_total == 2
    layer2 = RoIAlign(7, spatial_scale=0.125)
    layer2._cache.on_saturation = "raise"
    layer2.nhwc(y, torch.float16)
    with pytest.raises(FeatureRangeError):
        layer2._cache.check(sync=True)
