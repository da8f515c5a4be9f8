"""Kernel-level parity (B200): every CUDA kernel against the CPU oracle on the same seeded inputs.
Bit-exact for geometry / validity / indices; 1e-3 relative (fp32) or 2e-2 (bf16) for features."""
import math
import os

import pytest
import torch

from oracle import hbb, rotated
from point_teacher_b200 import synth

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


# ------------------------------------------------------------------------------ bags (bit-exact)
@pytest.mark.parametrize("cfg_i,seed", [(0, 0), (1, 1)])
def test_bag_gen_bit_exact(cuda, cfg_i, seed):
    from point_teacher_b200 import proposals
    d = synth.hbb_batch(seed=seed, gt_range=(40, 90))
    for cfg in (synth.HBB_FINE_CFG[cfg_i], synth.HBB_EXT_CFG[cfg_i]):
        ref_p, ref_v = hbb.fine_proposals(d["pseudo_boxes"], cfg, d["img_metas"])
        got_p, got_v = proposals.fine_proposals_from_cfg([b.to(cuda) for b in d["pseudo_boxes"]], cfg, d["img_metas"])
        for rp, rv, gp, gv in zip(ref_p, ref_v, got_p, got_v):
            assert torch.equal(gp.cpu(), rp)
            assert torch.equal(gv.cpu(), rv)


def test_bag_gen_edge_boxes_and_validity(cuda):
    from point_teacher_b200 import proposals
    metas = [dict(img_shape=(800, 800, 3))]
    boxes = [torch.tensor([[-30., -30., 10., 10.], [780., 790., 900., 830.], [5., 5., 5., 5.], [0., 0., 800., 800.],
                           [100., 100., 3000., 3000.], [400., 400., 401., 403.]])]
    cfg = synth.HBB_EXT_CFG[1]
    rp, rv = hbb.fine_proposals(boxes, cfg, metas)
    gp, gv = proposals.fine_proposals_from_cfg([boxes[0].to(cuda)], cfg, metas)
    assert torch.equal(gp[0].cpu(), rp[0]) and torch.equal(gv[0].cpu(), rv[0])
    assert 0 < int(rv[0].sum()) < rv[0].numel()


def test_mil_gen_proposals_replication_and_neg_weights(cuda):
    from point_teacher_b200 import proposals
    d = synth.hbb_batch(seed=2, gt_range=(30, 60))
    cfg = synth.HBB_FINE_CFG[0]
    to = lambda l: [t.to(cuda) for t in l]  # noqa: E731
    rp, rv, rr, rl = hbb.mil_gen_proposals(d["pseudo_points"], d["pseudo_boxes"], cfg, d["gt_boxes"], d["img_metas"])
    gp, gv, gr, gl = proposals.MIL_gen_proposals_from_cfg(to(d["pseudo_points"]), to(d["pseudo_boxes"]), cfg,
                                                          to(d["gt_boxes"]), d["img_metas"])
    for a, b in zip(rr + rl, gr + gl):
        assert torch.equal(b.cpu(), a)
    negs = [n.clone() for n in d["neg_boxes"][0]]
    for i in range(len(negs)):                       # make a third of the negatives collide with base bags
        k = min(negs[i].shape[0] // 3, rp[i].shape[0])
        negs[i][:k] = rp[i][:k] + torch.tensor([1.5, -1.0, 2.0, 0.5])
    rn, rw = hbb.gen_negative_proposals(d["pseudo_points"], cfg, rp, d["img_metas"], injected=negs)
    gn, gw = proposals.gen_negative_proposals(to(d["pseudo_points"]), cfg, gp, d["img_metas"], neg_boxes=to(negs))
    for a, b in zip(rw, gw):
        assert torch.equal(b.cpu(), a)
    assert 0 < int(torch.cat(rw).sum()) < torch.cat(rw).numel()


def test_bbox_overlaps_bit_exact_and_golden(cuda, golden_dir):
    from point_teacher_b200 import ops
    g = torch.load(os.path.join(golden_dir, "bbox_overlaps.pt"))
    gen = torch.Generator().manual_seed(7)
    a = synth.make_boxes(gen, 37, (800, 800))
    b = synth.jitter_boxes(gen, synth.make_boxes(gen, 53, (800, 800)))
    a2 = synth.jitter_boxes(torch.Generator().manual_seed(8), a)
    for mode in ("iou", "iof", "giou"):
        assert torch.equal(ops.bbox_overlaps(a.to(cuda), b.to(cuda), mode).cpu(), g[mode])
        assert torch.equal(ops.bbox_overlaps(a.to(cuda), a2.to(cuda), mode, True).cpu(), g[mode + "_aligned"])
    assert ops.bbox_overlaps(a.to(cuda), b[:0].to(cuda)).shape == (37, 0)
    kat1 = torch.tensor([[0., 0, 10, 10], [10, 10, 20, 20], [32, 32, 38, 42]], device=cuda)
    kat2 = torch.tensor([[0., 0, 10, 20], [0, 10, 10, 19], [10, 10, 20, 20]], device=cuda)
    assert torch.allclose(ops.bbox_overlaps(kat1, kat2, "giou", True).cpu(),
                          torch.tensor([0.5, -0.05, -0.8214]), atol=1e-4)


# ------------------------------------------------------------------------------ RoIAlign
def _roi_inputs(seed, n, hw=(320, 320), C=256, B=2):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, C, hw[0] // 8, hw[1] // 8, generator=g)
    boxes = synth.make_boxes(g, n, hw, median=14, hi=120)
    boxes[:4] += torch.tensor([-30., -30., -30., -30.])
    boxes[4:8] += torch.tensor([300., 290., 310., 300.])
    boxes[8] = torch.tensor([50., 60., 40., 50.])            # negative width/height -> zeros
    boxes[9] = torch.tensor([100., 100., 100., 100.])        # degenerate
    boxes[10] = torch.tensor([0., 0., 320., 320.])           # whole image: 6x6 sampling grid
    rois = torch.cat([torch.randint(0, B, (n, 1), generator=g).float(), boxes], 1)
    return x, rois


def test_roi_align_fp32_nchw_vs_torchvision_semantics(cuda):
    from point_teacher_b200.roi_extractors import SingleRoIExtractor
    x, rois = _roi_inputs(11, 300)
    ext = SingleRoIExtractor(dict(type="RoIAlign", output_size=7), 256, [8]).to(cuda)
    got = ext((x.to(cuda),), rois.to(cuda))
    ref = hbb.single_roi_extract((x,), rois, [8])
    assert got.shape == ref.shape == (300, 256, 7, 7)
    assert _rel(got, ref) < 1e-3
    assert (got.cpu() - ref).abs().max() < 2e-5 * 8
    assert torch.count_nonzero(got[8]) == 0
    # empty RoIs -> zeros (0, C, 7, 7), like single_level_roi_extractor.py:75-77
    assert ext((x.to(cuda),), rois[:0].to(cuda)).shape == (0, 256, 7, 7)


def test_roi_align_sampling_ratio2_and_small_channel_count(cuda):
    from point_teacher_b200 import ops
    x, rois = _roi_inputs(12, 64, C=16)
    feat = ops.nchw_to_nhwc(x.to(cuda))
    assert torch.equal(feat.cpu(), x.permute(0, 2, 3, 1).contiguous())
    got = ops.roi_align_forward(feat, rois.to(cuda), ops.OUT_F32_NCHW, 0.125, sampling_ratio=2)
    ref = rotated.roi_align(x, rois, 7, 0.125, 2, True)
    assert _rel(got, ref) < 1e-3


def test_roi_align_bf16_operand_layout(cuda):
    from point_teacher_b200 import ops
    x, rois = _roi_inputs(13, 200)
    feat = ops.nchw_to_nhwc(x.to(cuda))
    ref = hbb.single_roi_extract((x,), rois, [8])                       # (K, C, 7, 7)
    ref_binmajor = ref.permute(0, 2, 3, 1).reshape(200, -1)             # k' = (ph*7+pw)*C + c
    got = ops.roi_align_forward(feat, rois.to(cuda), ops.OUT_BF16_BINMAJOR, 0.125).float().cpu()
    assert (got - ref_binmajor).abs().max() <= 2e-2 * ref_binmajor.abs().max()
    x3 = ops.roi_align_forward(feat, rois.to(cuda), ops.OUT_BF16X3_BINMAJOR, 0.125).float().cpu()
    hi, lo, hi2 = x3[:, :12544], x3[:, 12544:25088], x3[:, 25088:]
    assert torch.equal(hi, hi2)
    assert _rel(hi + lo, ref_binmajor) < 1e-3
    # bf16 feature map input (2e-2 class)
    featb = ops.nchw_to_nhwc(x.to(cuda), torch.bfloat16)
    gotb = ops.roi_align_forward(featb, rois.to(cuda), ops.OUT_BF16_BINMAJOR, 0.125).float().cpu()
    assert (gotb - ref_binmajor).abs().max() <= 2e-2 * ref_binmajor.abs().max()


@pytest.mark.parametrize("fdt", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("sr,C,hw", [(0, 256, (320, 320)), (2, 256, (320, 320)), (0, 128, (256, 512)), (0, 64, (800, 800))])
def test_roi_align_mma_path_vs_oracle(cuda, sr, C, hw, fdt):
    """TMA + mma.sync RoIAlign (bf16 / fp16 NHWC in, bf16 bin-major out): tiny, large, whole-image, border, outside,
    degenerate and negative-size RoIs; adaptive and fixed sampling; multi-level skip mask."""
    from point_teacher_b200 import ops
    x, rois = _roi_inputs(40 + sr + C, 400, hw=hw, C=C)
    g = torch.Generator().manual_seed(C)
    big = synth.make_boxes(g, 30, hw, median=150, sigma=0.6, lo=60, hi=700)      # many 4x4-pixel chunks per RoI
    rois[20:50, 1:] = big
    ref = rotated.roi_align(x, rois, 7, 0.125, sr, True).permute(0, 2, 3, 1).reshape(400, -1)
    featb = ops.nchw_to_nhwc(x.to(cuda), fdt)
    assert featb.dtype == fdt
    got = ops.roi_align_forward(featb, rois.to(cuda), ops.OUT_BF16_BINMAJOR, 0.125, sampling_ratio=sr).float().cpu()
    err = (got - ref).abs()
    assert err.max() <= 2e-2 * ref.abs().max(), err.max()
    # mean error: bf16 rounding of feature, weights and output; with fp16 operands only the output rounding is left
    assert err.mean() <= (4e-3 if fdt == torch.bfloat16 else 2e-3) * ref.abs().mean() + 1e-6
    assert torch.isfinite(got).all()
    if sr == 0:
        assert torch.count_nonzero(got[8]) == 0          # negative size: empty adaptive grid -> zeros
    # rows of another FPN level are left untouched
    lv = (torch.arange(400) % 3 == 0).int()
    out = torch.full((400, 49 * C), 7.0, dtype=torch.bfloat16, device=cuda)
    ops.roi_align_forward(featb, rois.to(cuda), ops.OUT_BF16_BINMAJOR, 0.125, sampling_ratio=sr, out=out,
                          roi_level=lv.to(cuda), level=1)
    o = out.float().cpu()
    assert torch.equal(o[lv == 0], torch.full_like(o[lv == 0], 7.0))
    assert torch.equal(o[lv == 1], got[lv == 1])


def test_nhwc_fp16_saturates(cuda):
    from point_teacher_b200 import ops
    x = torch.tensor([1e6, -1e6, 3.0, 65504.0, 7e4, -7e4, 0.5, -0.25]).reshape(1, 8, 1, 1).repeat(1, 1, 2, 2).contiguous()
    f = ops.nchw_to_nhwc(x.to(cuda), torch.float16).float().cpu()
    assert torch.isfinite(f).all()
    assert torch.equal(f[0, 0, 0], torch.tensor([65504., -65504., 3., 65504., 65504., -65504., 0.5, -0.25]))


def test_roi_align_rotated_vs_oracle(cuda):
    from point_teacher_b200.roi_extractors import RotatedSingleRoIExtractor
    g = torch.Generator().manual_seed(14)
    x = torch.randn(2, 256, 64, 64, generator=g)
    boxes = synth.make_boxes(g, 150, (512, 512), median=20, hi=100)
    c = hbb.xyxy_to_cxcywh(boxes)
    th = torch.rand(150, 1, generator=g) * math.pi - math.pi / 2
    rois = torch.cat([torch.randint(0, 2, (150, 1), generator=g).float(), c, th], 1)
    rois[:3, 1:3] = torch.tensor([[2., 2.], [510., 5.], [256., 511.]])     # straddling the border
    ext = RotatedSingleRoIExtractor(dict(type="RoIAlignRotated", out_size=7, sample_num=2, clockwise=True), 256,
                                    [8]).to(cuda)
    got = ext((x.to(cuda),), rois.to(cuda))
    ref = rotated.roi_align_rotated(x, rois, 7, 0.125, 2, True, True)
    assert _rel(got, ref) < 1e-3


@pytest.mark.parametrize("fdt", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("sr,clockwise,C,hw", [(2, True, 256, (512, 512)), (1, False, 128, (256, 384)), (2, False, 64, (1024, 1024))])
def test_roi_align_rotated_mma_path_vs_oracle(cuda, sr, clockwise, C, hw, fdt):
    """Rotated twin of the TMA + mma.sync RoIAlign (bf16 / fp16 NHWC in, bf16 bin-major out): tiny, large (many 4x4
    chunks), border-straddling, fully-outside and zero-size rotated RoIs; 1 and 2 samples per axis; both angle
    conventions; multi-level skip mask."""
    from point_teacher_b200 import ops
    g = torch.Generator().manual_seed(77 + sr + C)
    n = 300
    x = torch.randn(2, C, hw[0] // 8, hw[1] // 8, generator=g)
    boxes = synth.make_boxes(g, n, hw, median=16, hi=90)
    boxes[20:50] = synth.make_boxes(g, 30, hw, median=150, sigma=0.6, lo=60, hi=min(hw) * 0.8)
    c = hbb.xyxy_to_cxcywh(boxes)
    th = torch.rand(n, 1, generator=g) * math.pi - math.pi / 2
    th[50:60] = 0.0
    th[60:65] = math.pi / 2
    rois = torch.cat([torch.randint(0, 2, (n, 1), generator=g).float(), c, th], 1)
    rois[:3, 1:3] = torch.tensor([[2., 2.], [hw[1] - 2., 5.], [hw[1] / 2, hw[0] - 1.]])     # straddling the border
    rois[3, 1:3] = torch.tensor([-400., -400.])                                                # fully outside
    rois[4, 3:5] = 0.0                                                                         # zero size: one point
    rois[5, 1:5] = torch.tensor([hw[1] + 3.0, hw[0] + 3.0, 30.0, 30.0])                      # beyond the far corner
    ref = rotated.roi_align_rotated(x, rois, 7, 0.125, sr, True, clockwise).permute(0, 2, 3, 1).reshape(n, -1)
    feat = ops.nchw_to_nhwc(x.to(cuda), fdt)
    got = ops.roi_align_forward(feat, rois.to(cuda), ops.OUT_BF16_BINMAJOR, 0.125, sampling_ratio=sr, rotated=True,
                                clockwise=clockwise).float().cpu()
    err = (got - ref).abs()
    assert torch.isfinite(got).all()
    assert err.max() <= 2e-2 * ref.abs().max(), err.max()
    assert err.mean() <= (4e-3 if fdt == torch.bfloat16 else 2e-3) * ref.abs().mean() + 1e-6
    assert torch.count_nonzero(got[3]) == 0
    # agrees with the direct (fp32-exact) rotated kernel on the same bf16/fp16 feature map to output rounding
    direct = ops.roi_align_forward(feat, rois.to(cuda), ops.OUT_F32_NCHW, 0.125, sampling_ratio=sr, rotated=True,
                                   clockwise=clockwise).permute(0, 2, 3, 1).reshape(n, -1).cpu()
    assert (got - direct).abs().max() <= 1.2e-2 * direct.abs().max()
    lv = (torch.arange(n) % 3 == 0).int()
    out = torch.full((n, 49 * C), 7.0, dtype=torch.bfloat16, device=cuda)
    ops.roi_align_forward(feat, rois.to(cuda), ops.OUT_BF16_BINMAJOR, 0.125, sampling_ratio=sr, rotated=True,
                          clockwise=clockwise, out=out, roi_level=lv.to(cuda), level=1)
    o = out.float().cpu()
    assert torch.equal(o[lv == 0], torch.full_like(o[lv == 0], 7.0))
    assert torch.equal(o[lv == 1], got[lv == 1])


def test_multi_level_extractor_matches_oracle(cuda):
    from point_teacher_b200.roi_extractors import SingleRoIExtractor
    g = torch.Generator().manual_seed(15)
    feats = [torch.randn(2, 256, 64 >> i, 64 >> i, generator=g) for i in range(3)]
    boxes = synth.make_boxes(g, 120, (512, 512), median=90, sigma=1.0, lo=8, hi=500)
    rois = torch.cat([torch.randint(0, 2, (120, 1), generator=g).float(), boxes], 1)
    lv = hbb.map_roi_levels(rois, 3)
    assert len(set(lv.tolist())) == 3
    ext = SingleRoIExtractor(dict(type="RoIAlign", output_size=7), 256, [8, 16, 32]).to(cuda)
    assert torch.equal(ext.map_roi_levels(rois.to(cuda), 3).cpu(), lv)
    got = ext(tuple(f.to(cuda) for f in feats), rois.to(cuda))
    ref = hbb.single_roi_extract(tuple(feats), rois, [8, 16, 32])
    assert _rel(got, ref) < 1e-3
    got2 = ext(tuple(f.to(cuda) for f in feats), rois.to(cuda), roi_scale_factor=1.3)
    ref2 = hbb.single_roi_extract(tuple(feats), rois, [8, 16, 32], roi_scale_factor=1.3)
    assert _rel(got2, ref2) < 1e-3


# ------------------------------------------------------------------------------ tcgen05 GEMM
@pytest.mark.parametrize("M,N,K,relu,f32,split", [
    (128, 256, 64, False, True, False), (300, 256, 128, True, False, False), (1000, 1024, 1024, True, False, True),
    (5400, 1024, 12544, True, False, True), (5400, 1024, 12544, True, False, False),
    (4736, 1024, 1024, False, True, True), (77, 512, 3072, True, True, True),
    # short contractions with a tail round: the tail tiles are cut into 4 / 4 / 2 column slices (no K-split)
    (5000, 1024, 1024, True, False, True), (20000, 512, 512, False, True, True), (6600, 1024, 512, True, True, True)])
def test_fc_gemm_vs_fp32_reference(cuda, M, N, K, relu, f32, split):
    from point_teacher_b200 import ops
    g = torch.Generator().manual_seed(M + N + K)
    A = (torch.randn(M, K, generator=g)).to(torch.bfloat16)
    B = (torch.randn(N, K, generator=g) * 0.05).to(torch.bfloat16)
    bias = torch.randn(N, generator=g)
    ref = A.float().to(cuda) @ B.float().to(cuda).t() + bias.to(cuda)     # fp32 reference of the same op
    if relu:
        ref = ref.relu()
    for _ in range(2):   # second call checks the self-re-zeroing split-K workspace
        out = ops.fc_gemm(A.to(cuda), B.to(cuda), bias.to(cuda), relu=relu,
                          out_dtype=torch.float32 if f32 else torch.bfloat16, allow_split=split)
        tol = 1e-3 if f32 else 2e-2                      # fp32 out: accumulation-order noise only
        err = (out.float() - ref).abs().max().item()
        assert err <= tol * ref.abs().max().item(), (err, ref.abs().max().item())
    # the split-K arrival / completion counters (head of the workspace) re-arm themselves
    assert torch.count_nonzero(ops.gemm_workspace(cuda)[:4096]) == 0


def test_fc_gemm_paired_cta_group2_variant(cuda):
    """The opt-in cta_group::2 kernel (PTB200_GEMM_PAIR=1) in a fresh process: same numerics as the fp32 reference,
    including the tail-wave split-K reduction at pair granularity."""
    import subprocess
    import sys
    code = (
        "import torch\n"
        "from point_teacher_b200 import ops\n"
        "g = torch.Generator().manual_seed(1)\n"
        "for M, N, K in [(5400, 1024, 12544), (5000, 1024, 4096), (700, 512, 8192)]:\n"
        "    A = torch.randn(M, K, generator=g).to(torch.bfloat16).cuda()\n"
        "    B = (torch.randn(N, K, generator=g) * 0.05).to(torch.bfloat16).cuda()\n"
        "    bias = torch.randn(N, generator=g).cuda()\n"
        "    ref = (A.float() @ B.float().t() + bias).relu()\n"
        "    for _ in range(2):\n"
        "        out = ops.fc_gemm(A, B, bias, relu=True, out_dtype=torch.float32)\n"
        "        err = (out - ref).abs().max().item() / ref.abs().max().item()\n"
        "        assert err < 1e-3, (M, N, K, err)\n"
        "torch.cuda.synchronize()\n"
        "print('paired ok')\n")
    env = dict(os.environ, PTB200_GEMM_PAIR="1")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "paired ok" in r.stdout, r.stderr[-2000:]


@pytest.mark.parametrize("a_mn,b_mn,M,N,K", [
    (False, True, 700, 1024, 1024),      # dgrad dX = dY W           (W stored [out, in] = [K, N])
    (True, True, 1024, 1024, 5000),      # wgrad dW = dY^T X          (ragged K = RoIs: TMA zero-fill)
    (True, True, 1000, 12544, 777),      # FC1 wgrad shape class, ragged M and K, split-K tail
    (True, False, 296, 256, 192),        # mixed
    (False, True, 5000, 12544, 1024)])   # FC1 dgrad
def test_fc_gemm_mn_major_operands(cuda, a_mn, b_mn, M, N, K):
    """Operands stored contraction-index-major ([K, M] / [K, N]) feed tcgen05 as MN-major tiles: same result as the
    fp32 reference of the transposed-copy formulation, with padded trailing rows ignored."""
    from point_teacher_b200 import ops
    g = torch.Generator().manual_seed(M + 3 * N + K)
    A = torch.randn(M, K, generator=g).to(torch.bfloat16)
    B = (torch.randn(N, K, generator=g) * 0.05).to(torch.bfloat16)
    ref = A.float().to(cuda) @ B.float().to(cuda).t()
    junk = lambda r, c: torch.full((r, c), float("nan"), dtype=torch.bfloat16)  # noqa: E731
    As = torch.cat([A.t().contiguous(), junk(13, M)], 0) if a_mn else torch.cat([A, junk(5, K)], 0)
    Bs = torch.cat([B.t().contiguous(), junk(9, N)], 0) if b_mn else B
    for split in (True, False):
        out = ops.fc_gemm_mn(As.to(cuda), Bs.to(cuda), a_mn=a_mn, b_mn=b_mn, out_dtype=torch.float32, M=M, K=K,
                             allow_split=split)
        err = (out[:M] - ref).abs().max().item()
        assert err <= 1e-3 * ref.abs().max().item(), (err, ref.abs().max().item())
    assert torch.count_nonzero(ops.gemm_workspace(cuda)[:4096]) == 0


def test_fc_gemm_rejects_bad_shapes(cuda):
    from point_teacher_b200 import ops
    from point_teacher_b200._lib import PTB200Error
    A = torch.zeros(8, 64, dtype=torch.bfloat16, device=cuda)
    with pytest.raises(PTB200Error):
        ops.fc_gemm(A, torch.zeros(100, 64, dtype=torch.bfloat16, device=cuda))
    with pytest.raises(ValueError):
        ops.fc_gemm(A, torch.zeros(256, 128, dtype=torch.bfloat16, device=cuda))


def test_weight_prep_layouts(cuda):
    from point_teacher_b200 import ops
    g = torch.Generator().manual_seed(21)
    w = torch.randn(256, 256 * 49, generator=g) * 0.01
    p = ops.prep_fc1_weight(w.to(cuda), 256).float().cpu()
    ref = w.view(256, 256, 49).permute(0, 2, 1).reshape(256, -1)
    assert torch.equal(p, ref.to(torch.bfloat16).float())
    p3 = ops.prep_fc1_weight(w.to(cuda), 256, x3=True).float().cpu()
    K = 12544
    assert torch.equal(p3[:, :K], p3[:, K:2 * K])
    assert _rel(p3[:, :K] + p3[:, 2 * K:], ref) < 1e-4
    c3 = ops.cast_weight(w[:, :1024].contiguous().to(cuda), x3=True).float().cpu()
    assert _rel(c3[:, :1024] + c3[:, 2048:], w[:, :1024]) < 1e-4


# ------------------------------------------------------------------------------ rotated IoU / OBB bags
def _rboxes(g, n, img=512, median=20.0, hi=120.0):
    b = hbb.xyxy_to_cxcywh(synth.make_boxes(g, n, (img, img), median=median, hi=hi))
    th = torch.rand(n, 1, generator=g) * math.pi - math.pi / 2
    return torch.cat([b, th], 1)


@pytest.mark.parametrize("mode", ["iou", "iof"])
def test_box_iou_rotated_vs_oracle(cuda, mode):
    from point_teacher_b200 import ops
    g = torch.Generator().manual_seed(31)
    a = _rboxes(g, 90)
    b = _rboxes(g, 70)
    b[:40, :2] = a[:40, :2] + torch.randn(40, 2, generator=g) * 4        # many true overlaps
    b[40:45] = a[40:45]                                                  # identical boxes
    b[45, 2:4] = 0.0                                                     # degenerate
    ref = rotated.box_iou_rotated(a, b, mode)
    got = ops.box_iou_rotated(a.to(cuda), b.to(cuda), mode).cpu()
    assert (got - ref).abs().max() < 1e-3, (got - ref).abs().max()
    ref_al = rotated.box_iou_rotated(a[:70], b, mode, True)
    got_al = ops.box_iou_rotated(a[:70].contiguous().to(cuda), b.to(cuda), mode, aligned=True).cpu()
    assert (got_al - ref_al.reshape(-1)).abs().max() < 1e-3
    # threshold decisions the path consumes (neg weights: iou < 0.3)
    dis = ((got < 0.3) != (ref < 0.3)).float().mean().item()
    assert dis < 1e-3, dis
    # theta = 0 degenerates to bbox_overlaps
    a0, b0 = a.clone(), b.clone()
    a0[:, 4] = 0
    b0[:, 4] = 0
    if mode == "iou":
        ax = hbb.bbox_overlaps(hbb.cxcywh_to_xyxy(a0[:, :4]), hbb.cxcywh_to_xyxy(b0[:, :4]))
        assert (ops.box_iou_rotated(a0.to(cuda), b0.to(cuda)).cpu() - ax).abs().max() < 1e-3


def test_box_iou_rotated_known_answer_from_reference_tests(cuda):
    """OBB_TOD/tests/test_utils/test_overlaps.py:7-15 (vanishing / astronomically large boxes -> IoU 0 at atol 1e-3),
    through ``rbbox_overlaps``' clamp of w, h to >= 1e-3 (mmrotate/core/bbox/iou_calculators/rotate_iou2d_calculator.py):
    without it a box whose four vertices collapse onto its centre in fp32 makes the published algorithm itself return
    garbage, which is why the reference clamps."""
    from oracle import obb
    from point_teacher_b200 import ops
    from test_oracle import REF_RBBOX_GT, REF_RBBOX_PREDICT      # same directory (rootdir-relative test modules)
    a, b = torch.tensor(REF_RBBOX_PREDICT), torch.tensor(REF_RBBOX_GT)
    got = ops.box_iou_rotated(a.to(cuda), b.to(cuda), clamp_wh=True).cpu()
    assert got.shape == (3, 4)
    assert torch.allclose(got, torch.zeros(3, 4), atol=1e-3), got
    assert torch.allclose(got, obb.rbbox_overlaps(a, b), atol=1e-3)


def test_obb_bag_gen_bit_exact_and_neg_weight(cuda):
    from oracle import obb
    from point_teacher_b200 import ops
    g = torch.Generator().manual_seed(32)
    metas = [dict(img_shape=(512, 512, 3))] * 2
    boxes = [_rboxes(g, 11), _rboxes(g, 17)]
    boxes[0][0, :2] = torch.tensor([3., 3.])           # mostly outside -> invalid instances
    pts = [b[:, :2] for b in boxes]
    cfg = synth.OBB_EXT_CFG[0]
    ref, rvalid, _, _ = obb.mil_gen_proposals(pts, boxes, cfg, boxes, metas)
    rois = obb.rbbox2roi(boxes).to(cuda)
    wh = torch.tensor([[512., 512.], [512., 512.]], device=cuda)
    out, valid = ops.bag_gen(rois, wh, cfg["base_ratios"], cfg["shake_ratio"], cfg["min_scale"], rotated=True)
    assert torch.equal(out[:, 1:].cpu(), torch.cat(ref))
    assert torch.equal(valid.bool().cpu().reshape(-1, 1), torch.cat(rvalid))
    # negatives: the reference passes (x1,y1,x2,y2,theta) AS (cx,cy,w,h,theta)
    negs = [obb.sample_negative_boxes(50, (512, 512, 3), g) for _ in range(2)]
    rw = torch.cat([obb.negative_weights(negs[i], ref[i]) for i in range(2)])
    nrois = obb.rbbox2roi(negs).to(cuda)
    offs = torch.tensor([0, ref[0].shape[0], ref[0].shape[0] + ref[1].shape[0]], dtype=torch.int32, device=cuda)
    w = ops.neg_weight(nrois, out, offs, rotated=True).bool().cpu()
    # The decision is ``all(IoU < 0.3)``.  Rotated IoU is a polygon-clipping result whose last bits depend on the
    # evaluation order (oracle/c/rotated.c vs rotated_iou.cuh agree to ~1e-6), so a decision may only differ where
    # some IoU of that negative sits within 1e-4 of the threshold: enumerate that borderline set and demand EXACT
    # equality everywhere else (no percentage allowance).
    border = torch.zeros_like(rw)
    o = 0
    for i in range(2):
        iou = obb.rbbox_overlaps(negs[i], ref[i])
        border[o:o + negs[i].shape[0]] = ((iou - 0.3).abs() < 1e-4).any(1)
        o += negs[i].shape[0]
    assert torch.equal(w[~border], rw[~border]), (w != rw).nonzero().flatten().tolist()
    assert int(border.sum()) <= 2, "the borderline set must stay a handful, otherwise the test says nothing"
    # many more negatives, this time AROUND the bags (perturbed copies of bag boxes) so that the IoUs cover the
    # whole range and roughly half of the decisions are rejections: the same statement on 2 x 2000 boxes
    negs2 = []
    for i in range(2):
        pick = ref[i][torch.randint(0, ref[i].shape[0], (2000,), generator=g)].clone()
        pick[:, :2] += torch.randn(2000, 2, generator=g) * 4
        pick[:, 2:4] *= (torch.randn(2000, 2, generator=g) * 0.4).exp()
        pick[:, 4] += torch.randn(2000, generator=g) * 0.3
        negs2.append(pick)
    rw2 = torch.cat([obb.negative_weights(negs2[i], ref[i]) for i in range(2)])
    w2 = ops.neg_weight(obb.rbbox2roi(negs2).to(cuda), out, offs, rotated=True).bool().cpu()
    border2 = torch.cat([((obb.rbbox_overlaps(negs2[i], ref[i]) - 0.3).abs() < 1e-4).any(1) for i in range(2)])
    assert torch.equal(w2[~border2], rw2[~border2])
    assert int(border2.sum()) <= 40 and 0.02 < rw2.float().mean() < 0.98
