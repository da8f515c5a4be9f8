"""CPU checks of the drop-in boundary: the C-ABI library loads and exports exactly what
include/ptb200.h declares; the Python surface mirrors the reference's names; the product never
imports the oracle."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "ptb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pt_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from point_teacher_b200 import _lib, build
    build.build()
    lib = ctypes.CDLL(_lib.SO_PATH)
    syms = _header_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/ptb200.h but not exported"
    assert set(_lib.SIGNATURES) == set(syms), "ctypes table and header disagree"
    lib.pt_build_arch.restype = ctypes.c_char_p
    assert lib.pt_build_arch() == b"sm_100a"
    assert lib.pt_abi_version() == 1


def test_signatures_have_no_torch_types():
    src = open(os.path.join(ROOT, "include", "ptb200.h")).read()
    assert "at::" not in src and "torch" not in src.lower().replace("pytorch", "")


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "point_teacher_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), f
                assert not re.search(r"^\s*(from|import)\s+torchvision", txt, flags=re.M), f


def test_reference_surface_names():
    from point_teacher_b200 import mil_head, proposals, refine, registry, roi_extractors
    for n in ("fine_proposals_from_cfg", "MIL_gen_proposals_from_cfg", "gen_negative_proposals"):
        assert callable(getattr(proposals, n))
    for n in ("forward_mil", "forward_mil_head", "mil_bag_training", "mil_bag_selection",
              "MIL_head_burn_in_step1", "MIL_head_burn_in_step2"):
        assert callable(getattr(mil_head.MILHeadMixin, n))
    assert registry.ROI_EXTRACTORS.get("SingleRoIExtractor") is roi_extractors.SingleRoIExtractor
    assert registry.ROI_EXTRACTORS.get("RotatedSingleRoIExtractor") is roi_extractors.RotatedSingleRoIExtractor
    assert callable(refine.P2BRefineMixin.forward_mil_head_burn_in_step2)
    from point_teacher_b200 import assigners
    for n in ("TopkAssigner", "FUSETopkAssigner", "MaxIoUAssigner"):
        assert registry.BBOX_ASSIGNERS.get(n) is getattr(assigners, n)
    for n in ("BboxOverlaps2D", "BboxDistanceMetric"):
        assert registry.IOU_CALCULATORS.get(n) is getattr(assigners, n)
    a = registry.build_assigner(dict(type="FUSETopkAssigner", num_pre=5, topk=3,
                                     cls_cost=dict(type="FocalLossCost", weight=1.0),
                                     reg_cost=dict(type="PointCost", mode="L1", weight=1.0),
                                     location_cost=dict(type="InsiderCost", weight=1.0)))
    assert a.num_pre == 5 and a.topk == 3 and callable(a.assign)


def test_head_parameter_names_match_reference_checkpoints():
    from point_teacher_b200.mil_head import MILHead
    h = MILHead(num_classes=8, num_stages=2, top_k=1)
    names = set(dict(h.named_parameters()))
    for s in range(2):
        for n in (f"shared_fcs_reg.{s}.0.weight", f"shared_fcs_reg.{s}.1.bias", f"shared_fcs_bag.{s}.0.weight",
                  f"shared_fcs_bag.{s}.1.weight", f"fc_cls.{s}.weight", f"fc_ins.{s}.bias", f"fc_reg.{s}.weight",
                  f"fc_iou.{s}.weight"):
            assert n in names, n
    assert h.shared_fcs_reg[0][0].weight.shape == (1024, 12544)
    ext = h.bbox_roi_extractor
    assert ext.num_inputs == 1 and ext.roi_layers[0].output_size == (7, 7)
    assert abs(ext.roi_layers[0].spatial_scale - 0.125) < 1e-12


def test_rotated_layer_accepts_deprecated_aliases():
    from point_teacher_b200.registry import build_roi_extractor
    ext = build_roi_extractor(dict(type="RotatedSingleRoIExtractor",
                                   roi_layer=dict(type="RoIAlignRotated", out_size=7, sample_num=2, clockwise=True),
                                   out_channels=256, featmap_strides=[8]))
    l = ext.roi_layers[0]
    assert l.output_size == (7, 7) and l.sampling_ratio == 2 and l.clockwise is True


def test_cpu_tensors_are_rejected_not_silently_computed():
    import torch
    from point_teacher_b200 import ops
    with pytest.raises(ValueError):
        ops.bbox_overlaps(torch.zeros(3, 4), torch.zeros(2, 4))
    with pytest.raises(ValueError):
        ops.bag_gen(torch.zeros(3, 5), torch.zeros(1, 2), [1.0], None, 0)


def test_gen_num_neg_zero_returns_none_pair():
    import torch
    from point_teacher_b200.proposals import gen_negative_proposals
    assert gen_negative_proposals([torch.zeros(1, 2)], dict(gen_num_neg=0), None, None) == (None, None)


def test_bench_reference_arm_prints_one_contract_line():
    """``bench.py --impl reference`` (the CPU arm the driver runs beside ours): exactly one JSON line on stdout with
    the contract's keys; needs no GPU."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "imgs/s" and d["higher_is_better"] is True and d["value"] > 0
    from oracle import ref_shim
    # the reference's own files under the import shim where /root/reference is mounted, the pinned port elsewhere
    assert d["cpu_baseline"]["kind"] == ("reference" if ref_shim.available() else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "imgs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("HBB cfg#1") and d["vs_baseline"] is None
