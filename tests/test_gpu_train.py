"""Backward of the MIL head on a B200 against torch autograd through the CPU oracle (fp32): every backward kernel on
its own, then the whole training step (parameter gradients + feature-map gradient).  Tolerance class: bf16 (2e-2 of
the largest reference entry) where bf16 operands are involved, 1e-4 for the fp32 loss-gradient kernels."""
import pytest
import torch
import torch.nn.functional as F

from oracle import hbb
from point_teacher_b200 import synth

pytestmark = pytest.mark.gpu

SMALL = dict(batch=2, img_hw=(256, 256), gt_range=(6, 10), n_neg=20)


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def test_reg_loss_grad_vs_autograd(cuda):
    from point_teacher_b200 import ops
    g = torch.Generator().manual_seed(1)
    G, U = 40, 25
    K = G * U
    ref = synth.make_boxes(g, G, (256, 256))
    bags = synth.jitter_boxes(g, ref.repeat_interleave(U, 0), 3.0, 0.3)
    bags[:5, :2] -= 40.0                                    # decoded boxes leave the image: clipped coordinates
    deltas = torch.randn(K, 4, generator=g) * 0.3
    deltas[7, 2] = 9.0                                      # beyond wh_ratio_clip: clamped, zero gradient
    valid = torch.rand(K, generator=g) > 0.2
    d = deltas.clone().requires_grad_(True)
    pred = hbb.delta2bbox(bags, d, max_shape=(256, 256, 3))
    loss = hbb.dn_diou_loss(pred, ref.repeat_interleave(U, 0), valid.float(), avg_factor=K, hyper=0.2)
    (loss * 0.7).backward()
    rois = torch.cat([torch.zeros(K, 1), bags], 1).to(cuda)
    sums = torch.zeros(8, device=cuda)
    sums[1] = float(valid.sum())
    got = ops.reg_loss_grad(deltas.to(cuda), rois, valid.to(torch.uint8).to(cuda), ref.to(cuda), U, (256, 256), sums,
                            torch.tensor([0.7], device=cuda), 1.0)
    assert _rel(got, d.grad) < 1e-4
    assert float(got[7, 2]) == 0.0


@pytest.mark.parametrize("U1,U2,C,n_neg", [(1, 25, 8, 40), (2, 25, 8, 0), (1, 64, 9, 17)])
def test_bag_loss_grad_vs_autograd(cuda, U1, U2, C, n_neg):
    from point_teacher_b200 import ops
    g = torch.Generator().manual_seed(U2 + C)
    G = 30
    K = G * U1 * U2
    cls = (torch.randn(K + n_neg, C, generator=g) * 1.5).requires_grad_(True)
    ins = (torch.randn(K, C, generator=g) * 1.5).requires_grad_(True)
    valid = torch.rand(K, generator=g) > 0.15
    valid[:U1 * U2] = False                                 # an all-invalid bag (weight 0, clamped normaliser)
    labels = torch.randint(0, C, (G,), generator=g)
    negw = torch.rand(n_neg, generator=g) > 0.3
    R = dict(cls_score=cls[:K].view(G, U1, U2, C), ins_score=ins.view(G, U1, U2, C),
             extensive_bags_valid=[valid.reshape(-1, 1)], neg_cls_score=cls[K:] if n_neg else None)
    loss = hbb.mil_bag_training(R, [labels], [negw] if n_neg else None)
    (loss * 0.25).backward()
    lw = valid.view(G * U1, U2).any(1)
    sums = torch.zeros(8, device=cuda)
    sums[6] = max(float(lw.sum()), 1.0)
    insp = torch.cat([ins.detach(), torch.zeros(n_neg, C)]).to(cuda)
    got = ops.bag_loss_grad(cls.detach().to(cuda), insp, valid.to(torch.uint8).to(cuda), labels.to(cuda), G, U1, U2,
                            negw.to(torch.uint8).to(cuda) if n_neg else None, n_neg, sums,
                            torch.tensor([0.25], device=cuda), 1.0, 1.0)
    assert _rel(got[:, :C], cls.grad) < 1e-4
    assert _rel(got[:K, C:], ins.grad) < 1e-4
    assert torch.count_nonzero(got[K:, C:]) == 0


def test_head_bwd_transposes_and_colsum(cuda):
    from point_teacher_b200 import ops
    g = torch.Generator().manual_seed(3)
    M, D, NO = 777, 1024, 16
    H = (torch.randn(M, D, generator=g)).relu().to(torch.bfloat16)
    W = torch.randn(NO, D, generator=g) * 0.05
    gr = torch.randn(M, NO, generator=g)
    dW, db = torch.zeros(NO, D, device=cuda), torch.zeros(NO, device=cuda)
    dZ = ops.head_bwd(gr.to(cuda), H.to(cuda), W.to(cuda), dW, db)
    ref_dZ = (gr @ W) * (H.float() > 0)
    assert _rel(dZ.float(), ref_dZ) < 1e-2
    assert _rel(dW, gr.t() @ H.float()) < 1e-4
    assert _rel(db, gr.sum(0)) < 1e-4
    t = ops.transpose_pad(H.to(cuda), rows=700)
    assert t.shape == (D, 704)
    assert torch.equal(t[:, :700].cpu(), H[:700].t()) and torch.count_nonzero(t[:, 700:]) == 0
    cs = ops.colsum_bf16(H.to(cuda), torch.zeros(D, device=cuda), M=700)
    assert _rel(cs, H[:700].float().sum(0)) < 1e-4
    # masked dgrad GEMM
    A = torch.randn(M, 1024, generator=g).to(torch.bfloat16)
    B = (torch.randn(1024, 1024, generator=g) * 0.05).to(torch.bfloat16)
    out = ops.fc_gemm_masked(A.to(cuda), B.to(cuda), H.to(cuda))
    ref = (A.float() @ B.float().t()) * (H.float() > 0)
    assert _rel(out.float(), ref) < 2e-2
    # FC1 weight-gradient column order
    dwp = torch.randn(8, 49 * 16, generator=g)
    got = ops.unpermute_dw1(dwp.to(cuda), 16, 49, torch.empty(8, 49 * 16, device=cuda), False).cpu()
    assert torch.equal(got, dwp.view(8, 49, 16).permute(0, 2, 1).reshape(8, -1))
    x = torch.randn(2, 5, 7, 16, generator=g)
    assert torch.equal(ops.nhwc_to_nchw_f32(x.to(cuda)).cpu(), x.permute(0, 3, 1, 2).contiguous())


def test_roi_align_backward_vs_torchvision_autograd(cuda):
    import torchvision
    from point_teacher_b200 import ops
    g = torch.Generator().manual_seed(4)
    x = torch.randn(2, 256, 40, 40, generator=g, requires_grad=True)
    boxes = synth.make_boxes(g, 200, (320, 320), median=14, hi=120)
    boxes[:4] += torch.tensor([-30., -30., -30., -30.])
    boxes[4] = torch.tensor([0., 0., 320., 320.])
    rois = torch.cat([torch.randint(0, 2, (200, 1), generator=g).float(), boxes], 1)
    out = torchvision.ops.roi_align(x, rois, (7, 7), 0.125, 0, True)              # (K, C, 7, 7)
    gout = torch.randn(out.shape, generator=g)
    out.backward(gout)
    dA = gout.permute(0, 2, 3, 1).reshape(200, -1).to(torch.bfloat16)             # bin-major operand layout
    dfeat = ops.roi_align_backward(dA.to(cuda), rois.to(cuda), (2, 40, 40, 256), 0.125)
    got = ops.nhwc_to_nchw_f32(dfeat).cpu()
    # reference with the same bf16-rounded upstream gradient
    x2 = x.detach().clone().requires_grad_(True)
    torchvision.ops.roi_align(x2, rois, (7, 7), 0.125, 0, True).backward(dA.float().view(200, 7, 7, 256).permute(0, 3, 1, 2))
    assert _rel(got, x2.grad) < 1e-4
    assert _rel(got, x.grad) < 2e-2


@pytest.mark.parametrize("seed,alpha", [(0, (1.0, 1.0)), (2, (0.01, 0.25))])
def test_training_step_gradients_vs_oracle(cuda, seed, alpha):
    from point_teacher_b200.mil_head import MILHead
    from point_teacher_b200.refine import phase2_refine
    d = synth.hbb_batch(seed=seed, **SMALL)
    P = hbb.MilHeadParams(num_stages=1, seed=seed).requires_grad_(True)
    feat = d["feat"].clone().requires_grad_(True)
    ob, op, ol, aux = hbb.phase2_refine(P, (feat,), [d["stride"]], d["img_metas"], d["pseudo_boxes"], d["pseudo_points"],
                                        d["pseudo_labels"], d["gt_boxes"], synth.HBB_FINE_CFG, synth.HBB_EXT_CFG,
                                        num_stages=1, alpha=alpha, topk=1, injected_negs=d["neg_boxes"])
    (ol["stage0_loss_mil_bbox"] + ol["stage0_loss_mil_bags"]).backward()
    ref = {k: v.grad for k, v in P.state_dict().items()}
    head = MILHead(num_classes=8, num_stages=1, top_k=1, precision="bf16").to(cuda)
    head.load_state_dict({k: v.detach() for k, v in P.state_dict().items()}, strict=False)
    x = d["feat"].to(cuda).requires_grad_(True)
    to = lambda l: [t.to(cuda) for t in l]  # noqa: E731
    gb, gp, gl = phase2_refine(head, (x,), d["img_metas"], to(d["pseudo_boxes"]), to(d["pseudo_points"]),
                               to(d["pseudo_labels"]), to(d["gt_boxes"]), synth.HBB_FINE_CFG, synth.HBB_EXT_CFG,
                               num_stages=1, alpha=alpha, neg_boxes=[to(d["neg_boxes"][0])], train=True)
    for k in ("stage0_loss_mil_bbox", "stage0_loss_mil_bags"):
        assert gl[k].requires_grad
        assert abs(float(gl[k].detach()) - float(ol[k].detach())) <= 2e-2 * max(abs(float(ol[k].detach())), 1e-3)
    (gl["stage0_loss_mil_bbox"] + gl["stage0_loss_mil_bags"]).backward()
    got = dict(head.named_parameters())

    def close(a, b, name):
        # mixed-precision gradients: every layer input / activation gradient is rounded to bf16 and the sums cancel
        # heavily, so the bound is on direction and on the Frobenius error, not on the worst single entry:
        # cosine >= 0.998 and ||got - ref|| <= 8e-2 ||ref||  (fp32 accumulation everywhere)
        a, b = a.double().cpu().flatten(), b.double().flatten()
        if b.abs().max() < 1e-9:                          # mathematically zero (softmax shift invariance of fc_ins.bias)
            assert a.abs().max() < 1e-6, name
            return
        cos = F.cosine_similarity(a, b, 0).item()
        fro = ((a - b).norm() / b.norm()).item()
        assert cos >= 0.998 and fro <= 8e-2, (name, cos, fro)

    for k, r in ref.items():
        assert got[k].grad is not None, k
        close(got[k].grad, r, k)
    for k in ("fc_cls.0.weight", "fc_cls.0.bias", "fc_reg.0.weight", "fc_reg.0.bias"):    # one layer deep: the bf16 class
        # (2e-2 of the largest entry: the hidden activation that multiplies the loss gradient is bf16, and which way its
        # 2^-9 roundings fall depends on the GEMM's summation order -- 0.95e-2 .. 1.1e-2 measured across schedules)
        assert _rel(got[k].grad, ref[k]) < 2e-2, (k, _rel(got[k].grad, ref[k]))
    for n, p in head.named_parameters():                   # constructed-but-unused modules never get a gradient
        if n.split(".")[0] in ("shared_fcs", "shared_fcs_refine", "fc_iou"):
            assert p.grad is None
    close(x.grad, feat.grad, "feature map")
    for a, b in zip(gb, ob):                               # the forward results are unchanged by train=True
        assert _rel(a.detach(), b) < 2e-2


def test_phase2_trainer_single_process(cuda):
    """Trainer step == plain backward when there is one rank; two stages chain through detached boxes."""
    from point_teacher_b200.mil_head import MILHead
    from point_teacher_b200.refine import phase2_refine
    from point_teacher_b200.train import Phase2Trainer
    d = synth.hbb_batch(seed=5, num_stages=2, **SMALL)
    torch.manual_seed(0)
    head = MILHead(num_classes=8, num_stages=2, top_k=3, precision="bf16").to(cuda)
    to = lambda l: [t.to(cuda) for t in l]  # noqa: E731
    args = (d["img_metas"], to(d["pseudo_boxes"]), to(d["pseudo_points"]), to(d["pseudo_labels"]), to(d["gt_boxes"]))
    negs = [to(n) for n in d["neg_boxes"]]
    x = d["feat"].to(cuda).requires_grad_(True)
    tr = Phase2Trainer(head, synth.HBB_FINE_CFG, synth.HBB_EXT_CFG, num_stages=2)
    boxes, pts, losses = tr.step((x,), *args, neg_boxes=negs, use_autograd=True)
    ga = {n: p.grad.clone() for n, p in head.named_parameters() if p.grad is not None}
    gxa = x.grad.clone()
    x.grad = None
    boxes, pts, losses = tr.step((x,), *args, neg_boxes=negs)            # direct (engine-free) backward: same gradients
    for n, p in head.named_parameters():
        if p.grad is not None:
            assert torch.allclose(p.grad, ga[n], rtol=0, atol=1e-5 * float(ga[n].abs().max()) + 1e-9), n
    assert torch.allclose(x.grad, gxa, rtol=0, atol=1e-5 * float(gxa.abs().max()))
    g1 = {n: p.grad.clone() for n, p in head.named_parameters() if p.grad is not None}
    gx1 = x.grad.clone()
    assert set(k.split(".")[0] for k in g1) == {"shared_fcs_reg", "shared_fcs_bag", "fc_cls", "fc_ins", "fc_reg"}
    assert all(k.split(".")[1] in ("0", "1") for k in g1) and len(g1) == 28          # both stages got gradients
    for p in head.parameters():
        p.grad = None
    x.grad = None
    b2, p2, l2 = phase2_refine(head, (x,), *args, synth.HBB_FINE_CFG, synth.HBB_EXT_CFG, num_stages=2, neg_boxes=negs,
                               train=True)
    sum(v for k, v in l2.items() if "loss" in k).backward()
    for n, p in head.named_parameters():
        if p.grad is not None:
            # atomics (small-head dW, split-K GEMM tails) make the summation order run-dependent
            assert torch.allclose(p.grad, g1[n], rtol=0, atol=1e-5 * float(g1[n].abs().max()) + 1e-9), n
    assert torch.allclose(x.grad, gx1, rtol=0, atol=1e-6 * float(gx1.abs().max()))      # red.add order only
    with torch.no_grad():
        b3, _, l3 = phase2_refine(head, (x.detach(),), *args, synth.HBB_FINE_CFG, synth.HBB_EXT_CFG, num_stages=2,
                                  neg_boxes=negs)
    for a, b in zip(boxes, b3):
        assert torch.equal(a, b)
    assert all(torch.isfinite(v).all() for v in losses.values())


def test_captured_train_step_matches_eager(cuda):
    from point_teacher_b200.mil_head import MILHead
    from point_teacher_b200.train import CapturedTrainStep, Phase2Trainer
    d = synth.hbb_batch(seed=6, **SMALL)
    torch.manual_seed(1)
    head = MILHead(num_classes=8, num_stages=1, top_k=1, precision="bf16").to(cuda)
    to = lambda l: [t.to(cuda) for t in l]  # noqa: E731
    inputs = dict(feat=d["feat"].to(cuda), pseudo_boxes=to(d["pseudo_boxes"]), pseudo_points=to(d["pseudo_points"]),
                  pseudo_labels=to(d["pseudo_labels"]), gt_boxes=to(d["gt_boxes"]), neg_boxes=[to(d["neg_boxes"][0])])
    x = inputs["feat"].clone().requires_grad_(True)
    tr = Phase2Trainer(head, synth.HBB_FINE_CFG, synth.HBB_EXT_CFG)
    eb, ep, el = tr.step((x,), d["img_metas"], inputs["pseudo_boxes"], inputs["pseudo_points"], inputs["pseudo_labels"],
                         inputs["gt_boxes"], neg_boxes=inputs["neg_boxes"])
    g_eager = {n: p.grad.clone() for n, p in tr.bucket.named}
    gx_eager = x.grad.clone()
    cap = CapturedTrainStep(head, inputs, d["img_metas"], synth.HBB_FINE_CFG, synth.HBB_EXT_CFG)
    for _ in range(2):
        boxes, pts, losses = cap.replay()
    torch.cuda.synchronize()
    for a, b in zip(eb, boxes):
        assert torch.equal(a, b)
    for n, p in tr.bucket.named:
        pg = dict(head.named_parameters())[n].grad
        assert torch.allclose(pg, g_eager[n], rtol=0, atol=1e-5 * float(g_eager[n].abs().max()) + 1e-9), n
    assert torch.allclose(cap.x.grad, gx_eager, rtol=0, atol=1e-5 * float(gx_eager.abs().max()))
    # new data through the static buffers changes the gradients
    inputs["feat"].mul_(0.5)
    cap.replay()
    torch.cuda.synchronize()
    assert not torch.allclose(cap.x.grad, gx_eager, rtol=0, atol=1e-5 * float(gx_eager.abs().max()))


# ------------------------------------------------------------------------------ OBB twin
def test_roi_align_rotated_backward_vs_autograd(cuda):
    import math
    from oracle import rotated
    from point_teacher_b200 import ops
    g = torch.Generator().manual_seed(7)
    K = 240
    x = torch.randn(2, 256, 40, 40, generator=g, requires_grad=True)
    rois = torch.stack([torch.randint(0, 2, (K,), generator=g).float(), torch.rand(K, generator=g) * 340 - 10,
                        torch.rand(K, generator=g) * 340 - 10, torch.rand(K, generator=g) * 40 + 2,
                        torch.rand(K, generator=g) * 40 + 2, torch.rand(K, generator=g) * math.pi - math.pi / 2], 1)
    rois[0, 3:5] = torch.tensor([300., 200.])                     # large RoI: many 4x4-pixel chunks
    rois[1, 1:3] = torch.tensor([-40., -40.])                     # entirely outside: no gradient
    for cw in (True, False):
        x.grad = None
        out = rotated.roi_align_rotated_torch(x, rois, 7, 0.125, 2, True, cw)
        gout = torch.randn(out.shape, generator=g).to(torch.bfloat16).float()
        out.backward(gout)
        dA = gout.permute(0, 2, 3, 1).reshape(K, -1).to(torch.bfloat16)
        dfeat = ops.roi_align_backward(dA.to(cuda), rois.to(cuda), (2, 40, 40, 256), 0.125, sampling_ratio=2,
                                       rotated=True, clockwise=cw)
        assert _rel(ops.nhwc_to_nchw_f32(dfeat), x.grad) < 1e-4


@pytest.mark.parametrize("seed,alpha", [(0, (1.0, 1.0)), (3, (0.01, 0.25))])
def test_obb_training_step_gradients_vs_oracle(cuda, seed, alpha):
    """OBB twin of the training step (rotated bags, RoIAlignRotated backward, 0.25/0.75 bag loss) against torch
    autograd through oracle/obb.py with the differentiable RoIAlignRotated twin."""
    from oracle import obb
    from point_teacher_b200.mil_head import RotatedMILHead
    from point_teacher_b200.train import Phase2Trainer
    d = synth.obb_batch(seed=seed, **SMALL)
    P = hbb.MilHeadParams(num_classes=9, num_stages=1, seed=seed).requires_grad_(True)
    feat = d["feat"].clone().requires_grad_(True)
    obb.DIFFERENTIABLE_ROI = True
    try:
        ob, op, ol, aux = obb.phase2_refine(P, (feat,), [d["stride"]], d["img_metas"], d["pseudo_boxes"],
                                            d["pseudo_points"], d["pseudo_labels"], d["gt_boxes"], synth.OBB_FINE_CFG,
                                            synth.OBB_EXT_CFG, alpha=alpha, injected_negs=d["neg_boxes"])
    finally:
        obb.DIFFERENTIABLE_ROI = False
    (ol["stage0_loss_mil_bbox"] + ol["stage0_loss_mil_bags"]).backward()
    ref = {k: v.grad for k, v in P.state_dict().items()}
    head = RotatedMILHead(num_classes=9, num_stages=1, top_k=3, precision="bf16").to(cuda)
    head.load_state_dict({k: v.detach() for k, v in P.state_dict().items()}, strict=False)
    x = d["feat"].to(cuda).requires_grad_(True)
    to = lambda l: [t.to(cuda) for t in l]  # noqa: E731
    tr = Phase2Trainer(head, synth.OBB_FINE_CFG, synth.OBB_EXT_CFG, num_stages=1, alpha=alpha)
    gb, gp, gl = tr.step((x,), d["img_metas"], to(d["pseudo_boxes"]), to(d["pseudo_points"]), to(d["pseudo_labels"]),
                         to(d["gt_boxes"]), neg_boxes=[to(d["neg_boxes"][0])])
    for k in ("stage0_loss_mil_bbox", "stage0_loss_mil_bags"):
        assert abs(float(gl[k]) - float(ol[k].detach())) <= 2e-2 * max(abs(float(ol[k].detach())), 1e-3), k
    got = dict(head.named_parameters())

    def close(a, b, name):
        a, b = a.double().cpu().flatten(), b.double().flatten()
        if b.abs().max() < 1e-9:
            assert a.abs().max() < 1e-6, name
            return
        cos = F.cosine_similarity(a, b, 0).item()
        fro = ((a - b).norm() / b.norm()).item()
        # bf16 operands at every layer, fp16 interpolation weights in the tensor-core RoIAlignRotated (measured:
        # cos 0.9980 / fro 0.063 on the FC1 weight gradient, the worst tensor)
        assert cos >= 0.997 and fro <= 8e-2, (name, cos, fro)

    for k, r in ref.items():
        assert got[k].grad is not None, k
        close(got[k].grad, r, k)
    close(x.grad, feat.grad, "feature map")
    for a, b in zip(gb, ob):
        assert a.shape[1] == 5 and _rel(a.detach(), b.detach()) < 2e-2


def _multilevel_batch(seed, n_gt=12):
    """A 512x512 batch whose boxes span three FPN levels under finest_scale = 56 (sqrt(wh) < 112 / < 224 / above)."""
    g = torch.Generator().manual_seed(900 + seed)
    d = synth.hbb_batch(seed=seed, batch=2, img_hw=(512, 512), gt_range=(n_gt, n_gt), n_neg=20)
    gts = [synth.make_boxes(g, n_gt, (512, 512), median=110.0, sigma=0.8, lo=8.0, hi=420.0) for _ in range(2)]
    d["gt_boxes"] = gts
    d["pseudo_boxes"] = [synth.jitter_boxes(g, b, ctr_sigma=3.0, log_sigma=0.15) for b in gts]
    d["pseudo_points"] = [torch.stack([(b[:, 0] + b[:, 2]) / 2, (b[:, 1] + b[:, 3]) / 2], 1) for b in d["pseudo_boxes"]]
    feats = [d["feat"], torch.randn(2, 256, 32, 32, generator=g), torch.randn(2, 256, 16, 16, generator=g)]
    return d, feats


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-3), ("bf16", 2e-2)])
def test_multilevel_fused_mil_path_vs_oracle(cuda, precision, tol):
    """north_star: "multi-level FPN RoIAlign over the bags" (single_level_roi_extractor.py:35-54, 98-104) on the FUSED
    path: every RoI pooled from its mapped level straight into the FC1 operand; forward in both precisions and, in
    bf16, the gradients w.r.t. all three feature maps and the parameters against torch autograd through the oracle."""
    from point_teacher_b200.mil_head import MILHead
    from point_teacher_b200.refine import phase2_refine
    strides = [8, 16, 32]
    d, feats = _multilevel_batch(3)
    P = hbb.MilHeadParams(num_stages=1, seed=3).requires_grad_(precision == "bf16")
    fo = [f.clone().requires_grad_(precision == "bf16") for f in feats]
    ctx = torch.enable_grad() if precision == "bf16" else torch.no_grad()
    with ctx:
        ob, op, ol, aux = hbb.phase2_refine(P, tuple(fo), strides, d["img_metas"], d["pseudo_boxes"], d["pseudo_points"],
                                            d["pseudo_labels"], d["gt_boxes"], synth.HBB_FINE_CFG, synth.HBB_EXT_CFG,
                                            num_stages=1, alpha=(1.0, 1.0), topk=1, injected_negs=d["neg_boxes"])
    lv = hbb.map_roi_levels(hbb.bbox2roi(aux[-1]["coarse_extensive_bags"]), 3)
    assert all(int((lv == i).sum()) > 10 for i in range(3)), "the batch must exercise every level"
    head = MILHead(num_classes=8, num_stages=1, top_k=1, precision=precision,
                   bbox_roi_extractor=dict(type="SingleRoIExtractor", roi_layer=dict(type="RoIAlign", output_size=7),
                                           out_channels=256, featmap_strides=strides)).to(cuda)
    head.load_state_dict({k: v.detach() for k, v in P.state_dict().items()}, strict=False)
    to = lambda l: [t.to(cuda) for t in l]  # noqa: E731
    x = tuple(f.to(cuda).requires_grad_(precision == "bf16") for f in feats)
    with ctx:
        gb, gp, gl = phase2_refine(head, x, d["img_metas"], to(d["pseudo_boxes"]), to(d["pseudo_points"]),
                                   to(d["pseudo_labels"]), to(d["gt_boxes"]), synth.HBB_FINE_CFG, synth.HBB_EXT_CFG,
                                   num_stages=1, alpha=(1.0, 1.0), neg_boxes=[to(d["neg_boxes"][0])],
                                   train=precision == "bf16")
    R, ref = head.last_results, aux[-1]
    assert _rel(R["cls_score"], ref["cls_score"].detach()) < tol
    assert _rel(R["ins_score"], ref["ins_score"].detach()) < tol
    assert _rel(torch.cat(R["extensive_bags"]), torch.cat(ref["extensive_bags"])) < tol
    for k in ol:
        assert abs(float(gl[k]) - float(ol[k])) <= tol * max(abs(float(ol[k])), 1e-3), k
    for a, b in zip(gb, ob):
        assert _rel(a.detach(), b.detach()) < tol
    if precision != "bf16":
        return
    (ol["stage0_loss_mil_bbox"] + ol["stage0_loss_mil_bags"]).backward()
    (gl["stage0_loss_mil_bbox"] + gl["stage0_loss_mil_bags"]).backward()
    for i in range(3):
        a, b = x[i].grad.double().cpu().flatten(), fo[i].grad.double().flatten()
        cos = F.cosine_similarity(a, b, 0).item()
        assert cos >= 0.995 and ((a - b).norm() / b.norm()).item() <= 0.1, (i, cos)
    got = dict(head.named_parameters())
    for k, v in P.state_dict().items():
        a, b = got[k].grad.double().cpu().flatten(), v.grad.double().flatten()
        if b.abs().max() < 1e-9:
            continue
        assert F.cosine_similarity(a, b, 0).item() >= 0.995, k
