"""The drop-in boundary, driven the way the reference drives it: oracle/detector.py's literal restatement of the
detectors' ``forward_mil_head_burn_in_step1/2`` (pinned bit-for-bit, gradients included, against the reference's own
detector methods by ``python -m oracle.check_oracle_vs_ref --detector``) calls the B200 classes -- built by the
reference's type names through the registry -- and the B200 by-name module functions, on the GPU, and is compared with
  * the golden fixtures produced by the reference's OWN detector + head + proposal functions (tests/golden/detector_*.pt,
    oracle/make_golden.py --detector), and
  * the same driver over the CPU oracle backend, run live, including the parameter / feature gradients of
    ``_parse_losses(losses).backward()``.
HBB (fcos_p2b_teacher_student.py:365-466) and OBB (rotated_fcos_teacher_student.py:435-535), phase 1 and phase 2."""
import os
import types

import pytest
import torch

from oracle import detector as D
from oracle import hbb, obb
from oracle.make_golden import DET_CAP, detector_inputs

pytestmark = pytest.mark.gpu


def _b200_backend(rotated):
    from point_teacher_b200 import ops, proposals, proposals_obb
    m = proposals_obb if rotated else proposals
    pts = (lambda b: b[:, :2]) if rotated else (lambda b: torch.stack([(b[:, 0] + b[:, 2]) / 2, (b[:, 1] + b[:, 3]) / 2], 1))
    return types.SimpleNamespace(
        MIL_gen_proposals_from_cfg=m.MIL_gen_proposals_from_cfg, gen_negative_proposals=m.gen_negative_proposals,
        aligned_iou_mean=lambda a, b: ops.aligned_iou_mean(a.float().contiguous(), b.float().contiguous(), rotated),
        box_points=pts, box_dim=5 if rotated else 4)


def _b200_head(cuda, rotated, P, stages, topk, precision):
    """Built through the registry under the REFERENCE's type name, with the reference config's keywords."""
    from point_teacher_b200 import registry
    import point_teacher_b200
    point_teacher_b200.install()
    if rotated:
        cfg = dict(type="TS_P2RBRotatedFCOSHead", num_classes=P.num_classes, in_channels=256, num_stages=stages,
                   top_k=topk, beta=0.25,
                   bbox_roi_extractor=dict(type="RotatedSingleRoIExtractor",
                                           roi_layer=dict(type="RoIAlignRotated", out_size=7, sample_num=2, clockwise=True),
                                           out_channels=256, featmap_strides=[8]),
                   loss_bbox_denosing=dict(type="DN_DIoULoss", loss_weight=1.0, hyper=0.2), precision=precision)
        head = registry.ROTATED_HEADS.build(cfg)
    else:
        cfg = dict(type="TS_P2BFCOSHead", num_classes=P.num_classes, in_channels=256, num_stages=stages, top_k=topk,
                   beta=0.25, mil_stack_conv=0, strides=[8], center_sampling=True, norm_on_bbox=True,   # FCOS keywords
                   bbox_roi_extractor=dict(type="SingleRoIExtractor", roi_layer=dict(type="RoIAlign", output_size=7),
                                           out_channels=256, featmap_strides=[8]),
                   loss_bbox_denosing=dict(type="DN_DIoULoss", loss_weight=1.0, hyper=0.2), precision=precision)
        head = registry.build_head(cfg)
    head = head.to(cuda)
    missing, unexpected = head.load_state_dict(P.state_dict(), strict=False)
    assert not unexpected
    return head


def _rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def _drive(det, fns, step, d, syn_boxes, fo, fs, dev=None):
    to = (lambda l: [t.to(dev) for t in l]) if dev is not None else (lambda l: l)
    args = (to(d["pseudo_boxes"]), to(d["pseudo_points"]), to(d["pseudo_labels"]), to(d["gt_boxes"]), d["img_metas"])
    n = len(d["pseudo_boxes"])
    torch.manual_seed(100 + d["_seed"])           # the negatives come from the global CPU generator, like the reference
    if step == 2:
        return D.forward_mil_head_burn_in_step2(det, fns, n, *args, (fo,))
    return D.forward_mil_head_burn_in_step1(det, fns, n, to(syn_boxes), *args, (fs,), (fo,))


@pytest.mark.parametrize("rotated", [False, True], ids=["hbb", "obb"])
@pytest.mark.parametrize("step", [2, 1])
def test_reference_driver_over_b200_classes(cuda, golden_dir, rotated, step):
    seed, stages = 3, 1
    g = torch.load(os.path.join(golden_dir, f"detector_{'obb' if rotated else 'hbb'}_step{step}.pt"))
    d, fine, ext, syn_boxes, feat_syn = detector_inputs(rotated, seed, stages)
    d["_seed"] = seed
    topk = 3 if rotated else 1
    P = hbb.MilHeadParams(num_classes=9 if rotated else 8, num_stages=stages, seed=seed)
    fns = _b200_backend(rotated)

    # ---- forward parity in fp32-emulation precision (forward only): the reference's own numbers at 1e-3
    head32 = _b200_head(cuda, rotated, P, stages, topk, "fp32")
    det = D.make_detector(head32, fine, ext, stages, cap1=DET_CAP, cap2=DET_CAP)
    with torch.no_grad():
        b32, p32, l32 = _drive(det, fns, step, d, syn_boxes, d["feat"].to(cuda), feat_syn.to(cuda), cuda)
    assert set(l32) == set(g["losses"])
    for i in range(len(b32)):
        assert _rel(b32[i], g["boxes"][i]) < 1e-3 and _rel(p32[i], g["points"][i]) < 1e-3
        assert torch.equal(b32[i][DET_CAP:].cpu(), d["pseudo_boxes"][i][DET_CAP:])          # untouched tail
    for k, v in g["losses"].items():
        assert abs(float(l32[k]) - float(v)) <= 1e-3 * max(abs(float(v)), 1e-3), (k, float(l32[k]), float(v))
    # fp32 precision has no backward: asking for gradients must fail loudly, not detach silently
    with pytest.raises(NotImplementedError):
        _drive(det, fns, step, d, syn_boxes, d["feat"].to(cuda), feat_syn.to(cuda), cuda)

    # ---- bf16 precision under autograd: losses carry a grad_fn, _parse_losses(...).backward() trains the head
    head = _b200_head(cuda, rotated, P, stages, topk, "bf16")
    det = D.make_detector(head, fine, ext, stages, cap1=DET_CAP, cap2=DET_CAP)
    fo = d["feat"].to(cuda).requires_grad_(True)
    fs = feat_syn.to(cuda).requires_grad_(True)
    b16, p16, l16 = _drive(det, fns, step, d, syn_boxes, fo, fs, cuda)
    for k in (f"stage0_loss_mil_bbox", f"stage0_loss_mil_bags"):
        assert l16[k].grad_fn is not None, k
    for k, v in g["losses"].items():
        assert abs(float(l16[k].detach()) - float(v)) <= 2e-2 * max(abs(float(v)), 1e-3), (k, float(l16[k].detach()), float(v))
    D.parse_losses(l16).backward()
    torch.cuda.synchronize()

    # ---- the same driver over the CPU oracle backend, gradients by torch autograd
    Pg = hbb.MilHeadParams(num_classes=9 if rotated else 8, num_stages=stages, seed=seed).requires_grad_(True)
    obb.DIFFERENTIABLE_ROI = True
    try:
        deto = D.make_detector(D.OracleHead(Pg, [d["stride"]], topk, rotated=rotated), fine, ext, stages, cap1=DET_CAP,
                               cap2=DET_CAP)
        ofo, ofs = d["feat"].clone().requires_grad_(True), feat_syn.clone().requires_grad_(True)
        ob, op, ol = _drive(deto, D.oracle_backend(rotated), step, d, syn_boxes, ofo, ofs)
        D.parse_losses(ol).backward()
    finally:
        obb.DIFFERENTIABLE_ROI = False
    for k in ol:
        assert abs(float(l16[k].detach()) - float(ol[k].detach())) <= 2e-2 * max(abs(float(ol[k].detach())), 1e-3), k
    sd = Pg.state_dict()
    named = dict(head.named_parameters())
    checked = 0
    for k, t in sd.items():
        ref = t.grad
        got = named[k].grad
        assert (ref is None) == (got is None), k
        if ref is None:
            continue
        got = got.float().cpu()
        checked += 1
        if ref.abs().max() < 1e-9:        # mathematically zero (softmax shift invariance of fc_ins.bias)
            assert got.abs().max() < 1e-6, k
            continue
        cos = torch.nn.functional.cosine_similarity(got.reshape(1, -1).double(), ref.reshape(1, -1).double()).item()
        assert cos >= 0.995, (k, cos)
        assert (got - ref).norm() <= 0.1 * ref.norm() + 1e-12, (k, float((got - ref).norm() / ref.norm()))
        # the reference's own gradient digest (norm + sampled entries)
        dg = g["grads"][k]
        assert abs(float(got.double().norm()) - dg["norm"]) <= 0.1 * dg["norm"] + 1e-12, k
        samp = got.reshape(-1)[dg["idx"]]
        assert (samp - dg["val"]).norm() <= 0.15 * dg["val"].norm() + 1e-3 * dg["norm"], k
    assert checked == 14
    # every used parameter received a gradient; the constructed-but-unused reference modules did not
    for n_, p_ in head.named_parameters():
        assert (p_.grad is None) == n_.startswith(("shared_fcs.", "shared_fcs_refine.", "fc_iou.")), n_
    for got, ref, dg in ((fo.grad, ofo.grad, g["feat_grad"]), (fs.grad, ofs.grad, g["feat_syn_grad"])):
        if ref is None or step == 2 and got is fs.grad:
            continue
        got = got.float().cpu()
        cos = torch.nn.functional.cosine_similarity(got.reshape(1, -1).double(), ref.reshape(1, -1).double()).item()
        assert cos >= 0.995, cos
        if dg is not None:
            assert abs(float(got.double().norm()) - dg["norm"]) <= 0.1 * dg["norm"]
    if step == 2:
        assert fs.grad is None                      # the synthetic features are not part of a phase-2 step


def test_fine_grained_methods_refuse_to_drop_gradients(cuda):
    """forward_mil_head / mil_bag_training return detached numbers: under autograd with trainable parameters they
    raise instead of silently training nothing."""
    from point_teacher_b200 import proposals, synth
    d, fine, ext, _, _ = detector_inputs(False, 5, 1)
    P = hbb.MilHeadParams(num_stages=1, seed=5)
    head = _b200_head(cuda, False, P, 1, 1, "bf16")
    to = lambda l: [t.to(cuda) for t in l]  # noqa: E731
    pb, pp, gb = to(d["pseudo_boxes"]), to(d["pseudo_points"]), to(d["gt_boxes"])
    props, valids, refs, reals = proposals.MIL_gen_proposals_from_cfg(pp, pb, fine[0], gb, d["img_metas"])
    x = (d["feat"].to(cuda),)
    n = sum(b.shape[0] for b in pb)
    with pytest.raises(RuntimeError, match="detached"):
        head.forward_mil_head(n, [b.shape[0] for b in pb], x, props, valids, refs, reals, d["img_metas"], ext[0], 0)
    with torch.no_grad():
        R = head.forward_mil_head(n, [b.shape[0] for b in pb], x, props, valids, refs, reals, d["img_metas"], ext[0], 0)
        loss = head.mil_bag_training(R, to(d["pseudo_labels"]), None)
        merged = head.mil_bag_selection(R, d["img_metas"], pb, to(d["pseudo_labels"]))
    assert torch.isfinite(loss) and len(merged) == len(pb)
    assert R["cls_score"].shape[:3] == (n, 1, 25) and R["extensive_bags"][0].shape[1] == 4
