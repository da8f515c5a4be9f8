"""CPU tests that pin the oracle: the reference's own known answers, the committed golden
vectors generated from the reference's files (oracle/make_golden.py), and cross-checks of the
C restatements of the un-vendored mmcv kernels."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import hbb, rotated
from point_teacher_b200 import synth


def test_giou_known_answer_from_reference_tests():
    # HBB_TOD/tests/test_metrics/test_box_overlap.py:86-99
    b1 = torch.FloatTensor([[0, 0, 10, 10], [10, 10, 20, 20], [32, 32, 38, 42]])
    b2 = torch.FloatTensor([[0, 0, 10, 20], [0, 10, 10, 19], [10, 10, 20, 20]])
    g = hbb.bbox_overlaps(b1, b2, "giou", is_aligned=True)
    assert torch.allclose(g, torch.tensor([0.5000, -0.0500, -0.8214]), atol=1e-4)


def test_delta2bbox_docstring_example():
    # HBB_TOD/mmdet/core/bbox/coder/delta_xywh_bbox_coder.py delta2bbox docstring
    rois = torch.Tensor([[0., 0., 1., 1.], [0., 0., 1., 1.], [0., 0., 1., 1.], [5., 5., 5., 5.]])
    deltas = torch.Tensor([[0., 0., 0., 0.], [1., 1., 1., 1.], [0., 0., 2., -1.], [0.7, -1.9, -0.5, 0.3]])
    out = hbb.delta2bbox(rois, deltas, max_shape=(32, 32, 3))
    exp = torch.tensor([[0.0000, 0.0000, 1.0000, 1.0000], [0.1409, 0.1409, 2.8591, 2.8591],
                        [0.0000, 0.3161, 4.1945, 0.6839], [5.0000, 5.0000, 5.0000, 5.0000]])
    assert torch.allclose(out, exp, atol=1e-4)


def test_bbox_overlaps_golden(golden_dir):
    g = torch.load(os.path.join(golden_dir, "bbox_overlaps.pt"))
    gen = torch.Generator().manual_seed(7)
    a = synth.make_boxes(gen, 37, (800, 800))
    b = synth.jitter_boxes(gen, synth.make_boxes(gen, 53, (800, 800)))
    a2 = synth.jitter_boxes(torch.Generator().manual_seed(8), a)
    for mode in ("iou", "iof", "giou"):
        assert torch.equal(hbb.bbox_overlaps(a, b, mode), g[mode])
        assert torch.equal(hbb.bbox_overlaps(a, a2, mode, True), g[mode + "_aligned"])


def _replay_golden(path):
    """Re-run the oracle on the regenerated inputs and compare with the reference's outputs."""
    g = torch.load(path)
    d = synth.hbb_batch(seed=g["seed"], num_stages=g["stages"], **g["small"])
    P = hbb.MilHeadParams(num_stages=g["stages"], seed=g["seed"])
    with torch.no_grad():
        _, _, losses, aux = hbb.phase2_refine(P, (d["feat"],), [d["stride"]], d["img_metas"], d["pseudo_boxes"],
                                              d["pseudo_points"], d["pseudo_labels"], d["gt_boxes"],
                                              synth.HBB_FINE_CFG, synth.HBB_EXT_CFG, num_stages=g["stages"],
                                              alpha=(1.0, 1.0), topk=g["topk"], injected_negs=d["neg_boxes"])
    return g, losses, aux


@pytest.mark.parametrize("tag", ["s1_top1", "s2_top3"])
def test_oracle_reproduces_reference_golden(golden_dir, tag):
    g, losses, aux = _replay_golden(os.path.join(golden_dir, f"hbb_phase2_{tag}.pt"))
    for s, (ref, R) in enumerate(zip(g["per_stage"], aux)):
        assert torch.equal(torch.cat(R["coarse_extensive_bags"]), ref["ext_bags"])       # bag geometry: bit-exact
        assert torch.equal(torch.cat(R["extensive_bags_valid"]), ref["ext_valid"])
        assert torch.equal(torch.cat(R["extensive_bags"]), ref["refined_bags"])
        assert torch.equal(R["cls_score"], ref["cls_score"])
        assert torch.equal(R["ins_score"], ref["ins_score"])
        assert torch.equal(R["neg_cls_score"], ref["neg_cls_score"])
        assert torch.allclose(losses[f"stage{s}_loss_mil_bbox"], ref["loss_mil_bbox"], rtol=1e-6)
        assert torch.allclose(losses[f"stage{s}_loss_mil_bags"], ref["loss_mil_bags"], rtol=1e-6)
        assert torch.allclose(R["coarse_bags_iou"], ref["coarse_bags_iou"], rtol=1e-6)


def test_c_roi_align_matches_torchvision():
    import torchvision
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 16, 40, 40, generator=g)
    boxes = synth.make_boxes(g, 60, (320, 320), median=30, hi=200)
    boxes[:5] += torch.tensor([-40., -40., -40., -40.])      # partly outside the image
    boxes[5:8] += torch.tensor([300., 300., 300., 300.])
    rois = torch.cat([torch.randint(0, 2, (60, 1), generator=g).float(), boxes], 1)
    for sr in (0, 2):
        ref = torchvision.ops.roi_align(x, rois, (7, 7), 0.125, sr, True)
        out = rotated.roi_align(x, rois, 7, 0.125, sr, True)
        assert torch.allclose(out, ref, atol=2e-5), (out - ref).abs().max()


def test_rotated_roi_align_theta0_equals_horizontal():
    g = torch.Generator().manual_seed(4)
    x = torch.randn(2, 8, 32, 32, generator=g)
    boxes = synth.make_boxes(g, 40, (256, 256), median=24, hi=120)
    b = torch.randint(0, 2, (40, 1), generator=g).float()
    rois5 = torch.cat([b, boxes], 1)
    c = hbb.xyxy_to_cxcywh(boxes)
    rois6 = torch.cat([b, c, torch.zeros(40, 1)], 1)
    ref = rotated.roi_align(x, rois5, 7, 0.125, 2, True)
    out = rotated.roi_align_rotated(x, rois6, 7, 0.125, 2, True, True)
    assert torch.allclose(out, ref, atol=2e-5)


def test_rotated_roi_align_torch_twin_matches_c_restatement():
    """The autograd-capable PyTorch twin (used for the OBB gradient tests) == the C restatement, incl. RoIs that
    leave the map and both rotation directions."""
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 16, 32, 32, generator=g)
    K = 300
    rois = torch.stack([torch.randint(0, 2, (K,), generator=g).float(), torch.rand(K, generator=g) * 280 - 10,
                        torch.rand(K, generator=g) * 280 - 10, torch.rand(K, generator=g) * 60 + 1,
                        torch.rand(K, generator=g) * 60 + 1, torch.rand(K, generator=g) * math.pi - math.pi / 2], 1)
    for cw in (True, False):
        a = rotated.roi_align_rotated(x, rois, 7, 0.125, 2, True, cw)
        b = rotated.roi_align_rotated_torch(x, rois, 7, 0.125, 2, True, cw)
        assert torch.allclose(a, b, atol=2e-5)


def test_rotated_roi_align_orientation_kat():
    # square RoI, theta=+pi/2, clockwise=True == theta=0 transposed and flipped
    g = torch.Generator().manual_seed(5)
    x = torch.randn(1, 4, 32, 32, generator=g)
    r0 = torch.tensor([[0., 128., 120., 64., 64., 0.]])
    r90 = torch.tensor([[0., 128., 120., 64., 64., math.pi / 2]])
    a = rotated.roi_align_rotated(x, r0, 7, 0.125, 2, True, True)
    b = rotated.roi_align_rotated(x, r90, 7, 0.125, 2, True, True)
    assert torch.allclose(b, a.transpose(2, 3).flip(2), atol=1e-5)


def test_rotated_iou_theta0_equals_axis_aligned_and_cv2():
    import cv2
    g = torch.Generator().manual_seed(6)
    a = synth.make_boxes(g, 50, (400, 400), median=40, hi=150)
    b = synth.jitter_boxes(g, a, ctr_sigma=10)
    a5 = torch.cat([hbb.xyxy_to_cxcywh(a), torch.zeros(50, 1)], 1)
    b5 = torch.cat([hbb.xyxy_to_cxcywh(b), torch.zeros(50, 1)], 1)
    assert torch.allclose(rotated.box_iou_rotated(a5, b5), hbb.bbox_overlaps(a, b), atol=1e-4)
    # random angles against OpenCV's polygon intersection (Appendix A.7)
    a5[:, 4] = torch.rand(50, generator=g) * math.pi - math.pi / 2
    b5[:, 4] = torch.rand(50, generator=g) * math.pi - math.pi / 2
    got = rotated.box_iou_rotated(a5, b5, aligned=True)
    for i in range(50):
        ra = ((float(a5[i, 0]), float(a5[i, 1])), (float(a5[i, 2]), float(a5[i, 3])), math.degrees(float(a5[i, 4])))
        rb = ((float(b5[i, 0]), float(b5[i, 1])), (float(b5[i, 2]), float(b5[i, 3])), math.degrees(float(b5[i, 4])))
        ret, pts = cv2.rotatedRectangleIntersection(ra, rb)
        inter = cv2.contourArea(cv2.convexHull(pts)) if ret != 0 and pts is not None else 0.0
        union = float(a5[i, 2] * a5[i, 3] + b5[i, 2] * b5[i, 3]) - inter
        assert abs(float(got[i]) - inter / union) < 2e-3
    ident = rotated.box_iou_rotated(a5, a5, aligned=True)
    assert torch.allclose(ident, torch.ones(50), atol=1e-4)


# The one known-answer vector the reference itself holds for the rotated overlaps
# (OBB_TOD/tests/test_utils/test_overlaps.py:7-15: vanishing and astronomically large predictions against four
# ordinary boxes; expected IoU 0 everywhere at atol 1e-3).
REF_RBBOX_PREDICT = [[903.34, 1034.4, 1.81e-7, 1e-7, -0.312], [903.34, 1034.4, 1e-7, 1e-3, -0.312],
                     [903.34, 1034.4, 1.81e7, 1e7, -0.312]]
REF_RBBOX_GT = [[2.1525e+02, 7.5750e+01, 3.3204e+01, 1.2649e+01, 3.2175e-01],
                [3.0013e+02, 7.7144e+02, 4.9222e+02, 3.1368e+02, -1.3978e+00],
                [8.4887e+02, 6.9989e+02, 4.6854e+02, 3.0743e+02, -1.4008e+00],
                [8.5250e+02, 7.0250e+02, 7.6181e+02, 3.8200e+02, -1.3984e+00]]


def test_rbbox_overlaps_known_answer_from_reference_tests():
    from oracle import obb
    ious = obb.rbbox_overlaps(torch.tensor(REF_RBBOX_PREDICT), torch.tensor(REF_RBBOX_GT))
    assert ious.shape == (3, 4)
    assert torch.allclose(ious, torch.zeros(3, 4), atol=1e-3), ious


def test_rotated_nms_keeps_descending_score_order():
    dets = torch.tensor([[50., 50., 20., 20., 0.], [52., 50., 20., 20., 0.1], [150., 150., 20., 10., 0.7],
                         [50., 51., 20., 20., 0.]])
    scores = torch.tensor([0.9, 0.8, 0.5, 0.95])
    out, keep = rotated.nms_rotated(dets, scores, 0.05)
    assert keep.tolist() == [3, 2]
    assert out.shape == (2, 6)


def test_obb_oracle_reproduces_reference_golden(golden_dir):
    """OBB twin: oracle/obb.py against outputs of the reference's own OBB_TOD files (oracle/make_golden.py)."""
    from oracle import obb
    g = torch.load(os.path.join(golden_dir, "obb_phase2_s1_top3.pt"))
    d = synth.obb_batch(seed=g["seed"], **g["small"])
    P = hbb.MilHeadParams(num_classes=9, num_stages=1, seed=g["seed"])
    with torch.no_grad():
        ob, _, losses, aux = obb.phase2_refine(P, (d["feat"],), [d["stride"]], d["img_metas"], d["pseudo_boxes"],
                                               d["pseudo_points"], d["pseudo_labels"], d["gt_boxes"],
                                               synth.OBB_FINE_CFG, synth.OBB_EXT_CFG, alpha=(1.0, 1.0),
                                               topk=g["topk"], injected_negs=d["neg_boxes"])
    R = aux[0]
    assert torch.equal(torch.cat(R["coarse_extensive_bags"]), g["ext_bags"])            # bag geometry: bit-exact
    assert torch.equal(torch.cat(R["extensive_bags_valid"]), g["ext_valid"])
    assert torch.equal(torch.cat(R["extensive_bags"]), g["refined_bags"])
    assert torch.equal(R["cls_score"], g["cls_score"])
    assert torch.equal(R["ins_score"], g["ins_score"])
    assert torch.equal(torch.cat(R["neg_weight"]), g["neg_weight"])
    assert torch.equal(torch.cat([b[:100] for b in ob]), g["merged"])
    for k, v in g["losses"].items():
        assert torch.allclose(losses[k], v, rtol=1e-6), k


def test_assign_oracle_reproduces_reference_golden(golden_dir):
    """Rows a13-a15: oracle/assign.py against outputs of the reference's own assigner / metric files."""
    from oracle import assign
    g = torch.load(os.path.join(golden_dir, "assign.pt"))
    for c in g["cases"]:
        d = synth.assign_batch(c["seed"], ties=c["ties"])
        gi, lb = assign.topk_assign(d["pred"], d["logits"], d["gt"], d["labels"], c["num_pre"], c["topk"])
        assert torch.equal(gi, c["topk_gt_inds"]) and torch.equal(lb, c["topk_labels"])
        gi, lb = assign.fuse_topk_assign(d["pred"], d["points"], d["logits"], d["gt"], d["labels"], c["num_pre"], c["topk"])
        assert torch.equal(gi, c["fuse_gt_inds"]) and torch.equal(lb, c["fuse_labels"])
        assert torch.equal(assign.focal_loss_table(d["logits"]), c["fl_table"])
    for m, ref in g["metric"].items():
        assert torch.equal(assign.bbox_metric(g["gts"], g["anchors"], m), ref), m
    for c in g["maxiou"]:
        ov = hbb.bbox_overlaps(g["gts"], g["anchors"], c["mode"]) if c["calc"] == "BboxOverlaps2D" else \
            assign.bbox_metric(g["gts"], g["anchors"], c["mode"])
        gi, mx, lb = assign.max_iou_assign(ov, g["labels"], **c["kw"])
        assert torch.equal(gi, c["gt_inds"]) and torch.equal(mx, c["max_overlaps"]) and torch.equal(lb, c["labels"])


def test_max_iou_known_answers_from_reference_tests():
    # HBB_TOD/tests/test_utils/test_assigner.py:15-63
    from oracle import assign
    bboxes = torch.FloatTensor([[0, 0, 10, 10], [10, 10, 20, 20], [5, 5, 15, 15], [32, 32, 38, 42]])
    gts = torch.FloatTensor([[0, 0, 10, 9], [0, 10, 10, 19]])
    gi, _, lb = assign.max_iou_assign(hbb.bbox_overlaps(gts, bboxes), torch.LongTensor([2, 3]), 0.5, 0.5)
    assert gi.tolist() == [1, 0, 2, 0]
    assert lb.tolist() == [2, -1, 3, -1]


def test_focal_loss_cost_docstring_shape_and_sign():
    # match_cost.py:66-74: costs are negative for confident positives; table column gather == direct call
    from oracle import assign
    g = torch.Generator().manual_seed(0)
    x = torch.rand(4, 3, generator=g)
    lab = torch.tensor([0, 1, 2])
    c = assign.focal_loss_cost(x, lab)
    assert c.shape == (4, 3) and (c < 0).all()
    assert torch.equal(c, assign.focal_loss_table(x)[:, lab])


def _random_quads(n, seed, size=300):
    from oracle import mask as M
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < n:
        cx, cy = rng.uniform(40, size - 40, 2)
        w, h = rng.uniform(0.3, 80, 2)
        a = rng.uniform(-np.pi / 2, np.pi / 2)
        if len(out) % 10 == 0:
            a = 0.0
        if len(out) % 17 == 0:
            a = float(np.pi / 2 * rng.integers(-1, 2))
        p = M.obb2poly_le90(torch.tensor([[cx, cy, w, h, a]], dtype=torch.float32)).view(4, 2).numpy().astype(np.int32)
        if p.min() >= 0 and p.max() < size:
            out.append(p)
    return out


def test_fill_poly_replay_equals_cv2():
    """The rasteriser restatement against the installed OpenCV (the reference calls cv2.fillPoly)."""
    import cv2
    from oracle import mask as M
    for p in _random_quads(400, 3):
        ref = np.zeros((300, 300), np.uint8)
        cv2.fillPoly(ref, [p], 1)
        got = np.zeros((300, 300), np.uint8)
        M.fill_poly_replay(got, p)
        assert np.array_equal(got, ref), p.tolist()


def test_black_paper_oracle_reproduces_reference_golden(golden_dir):
    from oracle import mask as M
    from point_teacher_b200 import masking
    g = torch.load(os.path.join(golden_dir, "black_paper.pt"))
    for c in g:
        d = synth.mask_batch(c["seed"])
        pattern, prior = masking.load_basic_shape(synth.SHAPE_LIST)
        dense = range(int(len(pattern) / 2))
        torch.manual_seed(c["seed"])
        np.random.seed(c["seed"])
        allb = M.sample_candidates(d["bb_occupied"], prior, dense, d["imgsize"])
        torch.manual_seed(c["seed"])
        np.random.seed(c["seed"])
        allb2 = masking.sample_black_paper_candidates(d["bb_occupied"], prior, dense, d["imgsize"])
        assert torch.equal(allb, allb2)                       # the product's host-side draw == the oracle's
        img, bb, sel, polys, m = M.black_paper_from_candidates(d["img"].clone(), allb, d["imgsize"])
        assert torch.equal(bb, c["kept"])
        assert np.array_equal(np.packbits(m.astype(bool)), c["mask_bits"].numpy())
        assert int(m.sum()) == c["n_px"]
        _, _, _, _, m2 = M.black_paper_from_candidates(d["img"].clone(), allb, d["imgsize"], use_cv2=False)
        assert np.array_equal(m, m2)


def test_pseudo_box_oracle_reproduces_reference_golden(golden_dir):
    from oracle import assign
    for c in torch.load(os.path.join(golden_dir, "pseudo_boxes.pt")):
        d = synth.pseudo_batch(c["seed"], G=c["G"])
        b, p, _, m, v, _, _ = assign.generate_pseudo_single(d["gt_points"], d["labels"], d["gt_boxes"], d["logits"],
                                                            d["ltrb"], d["points"], 0.1)
        assert torch.equal(b, c["boxes"]) and torch.equal(p, c["points"])
        assert torch.equal(m, c["mean_iou"]) and torch.equal(torch.sort(v)[0], c["valid"])


def test_target_pseudo_oracle_reproduces_reference_golden(golden_dir):
    from oracle import assign
    for c in torch.load(os.path.join(golden_dir, "pseudo_boxes.pt")):
        d = synth.pseudo_batch(c["seed"], G=c["G"])
        lr, t, lb, w = assign.get_target_pseudo_single(d["points"], d["logits"], d["gt_points"], d["labels"], c["boxes"],
                                                       d["labels"], 8)
        pos = c["pos"].long()
        assert torch.equal(lr, c["labels_reg"].long()) and torch.equal(lb, c["labels"].long())
        assert torch.equal(t[pos], c["bbox_targets_pos"]) and float(t.double().sum()) == float(c["targets_checksum"])
        assert torch.equal(assign.centerness_target(t[pos]), c["centerness_pos"])


# ------------------------------------------------------------------------------ strong_augmentation (8f rank 3)
def _aug_args(d):
    return [d["img"].clone()] + [[t.clone() for t in d[k]] for k in
                                 ("gt_points", "gt_labels", "pseudo_points", "pseudo_labels", "pseudo_bboxes")]


def test_augment_oracle_reproduces_reference_golden(golden_dir):
    """Outputs of the reference's own strong_augmentation (HBB and OBB) from the injected draws: kept sets, labels
    and coordinates bit-exact; the rounded image bit-exact (same ATen / torchvision CPU kernels)."""
    import random
    from oracle import augment
    for c in torch.load(os.path.join(golden_dir, "augment.pt")):
        d = synth.augment_batch(c["seed"], rotated=c["rotated"])
        random.seed(c["seed"])
        np.random.seed(c["seed"])
        assert augment.draw_choices(2, rotated=c["rotated"]) == [tuple(x) for x in c["choices"]]     # RNG call order
        fn = augment.strong_augmentation_obb if c["rotated"] else augment.strong_augmentation_hbb
        out = fn(*_aug_args(d), c["choices"])
        assert torch.equal(out[0], c["images"].float())
        for key, got in zip(("gt_points", "gt_labels", "pseudo_points", "pseudo_labels", "pseudo_bboxes"), out[2:]):
            for a, b in zip(got, c[key]):
                assert a.shape == b.shape and torch.equal(a, b), key


def test_augment_explicit_resamplers_match_library_kernels():
    """The explicit fp32 formulae the device kernel replays == F.interpolate / torchvision rotate on this CPU
    (the generic ATen upsample kernel is the one used for H + W > 128 with more than one thread)."""
    import torch.nn.functional as F
    import torchvision.transforms.functional as TF
    from oracle import augment
    if torch.get_num_threads() == 1:
        pytest.skip("single-threaded ATen takes the other (vectorised) bilinear kernel for 3-channel images")
    img = synth.augment_batch(0, img_hw=(200, 168))["img"][0]
    for sf in (0.8, 0.9, 1.0, 1.1, 1.2):
        oh, ow = int(200 * sf), int(168 * sf)
        ref = F.interpolate(img[None], size=(oh, ow), mode="bilinear", align_corners=False)[0]
        assert torch.equal(augment.bilinear_resize_exact(img, oh, ow), ref), sf
    for ang in (1, 7, 19):
        assert torch.equal(augment.rotate_nearest_exact(img, ang), TF.rotate(img, ang, fill=0)), ang


# ------------------------------------------------------------------------------ losses (8f rank 4)
def test_diff_iou_rotated_restatement_cross_checks():
    """mmcv's diff_iou_rotated_2d is un-vendored (parity unpinned): the restatement agrees with the polygon-clipping
    IoU (a different algorithm, oracle/rotated.py) and, at theta = 0, with the axis-aligned IoU."""
    from oracle import losses as L
    g = torch.Generator().manual_seed(0)
    n = 400
    c = torch.rand(n, 2, generator=g) * 200 + 20
    wh = torch.rand(n, 2, generator=g) * 40 + 2
    a = torch.rand(n, 1, generator=g) * math.pi - math.pi / 2
    b1 = torch.cat([c, wh, a], 1)
    b2 = b1.clone()
    b2[:, :2] += torch.randn(n, 2, generator=g) * 6
    b2[:, 2:4] *= torch.exp(torch.randn(n, 2, generator=g) * 0.3)
    b2[:, 4] += torch.randn(n, generator=g) * 0.5
    iou = L.diff_iou_rotated_2d(b1[None], b2[None])[0]
    assert (iou - rotated.box_iou_rotated(b1, b2, aligned=True)).abs().max().item() < 2e-4
    b1[:, 4] = 0
    b2[:, 4] = 0
    ref = hbb.bbox_overlaps(hbb.cxcywh_to_xyxy(b1[:, :4]), hbb.cxcywh_to_xyxy(b2[:, :4]), is_aligned=True)
    assert (L.diff_iou_rotated_2d(b1[None], b2[None])[0] - ref).abs().max().item() < 2e-4


def test_focal_loss_matches_binary_cross_entropy_limit():
    """gamma = 0, alpha = 0.5 reduces the focal loss to half the BCE-with-logits (focal_loss.py:11-57)."""
    import torch.nn.functional as F
    from oracle import losses as L
    g = torch.Generator().manual_seed(1)
    p, t = torch.randn(50, 6, generator=g), torch.randint(0, 7, (50,), generator=g)
    oh = F.one_hot(t, 7)[:, :6].float()
    assert torch.allclose(L.sigmoid_focal_loss(p, t, gamma=0.0, alpha=0.5, reduction="sum"),
                          0.5 * F.binary_cross_entropy_with_logits(p, oh, reduction="sum"), rtol=1e-6)
