"""CPU: the registration branch for an OpenMMLab environment.  ``oracle/ref_shim`` provides ``mmcv`` / ``mmdet`` /
``mmrotate`` shells whose leaf modules are the reference's own files; under them ``point_teacher_b200.install()`` must
re-register the reference's head classes as (B200 mix-in, reference class) subclasses under the reference's own type
names and replace the RoI extractors and assigners in the reference's registries.  Runs in a subprocess so that the
stub packages never leak into the test session.  Skipped where /root/reference is not mounted (GPU box)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

SCRIPT = r"""
import sys
sys.path.insert(0, %r)
from oracle import ref_shim
dn = ref_shim.install_detectors()                     # mmdet.models.builder etc. now resolve to the shim
import point_teacher_b200
from point_teacher_b200 import registry, mil_head, roi_extractors
assert registry.USING_OPENMMLAB, "the OpenMMLab branch of registry.py did not run"
import mmdet.models.builder as mb
assert registry.HEADS is mb.HEADS
out = point_teacher_b200.install()
H = mb.HEADS.get("TS_P2BFCOSHead")
R = mb.HEADS.get("TS_P2RBRotatedFCOSHead")
assert H is out["TS_P2BFCOSHead"] and R is out["TS_P2RBRotatedFCOSHead"]
assert issubclass(H, mil_head.MILHeadMixin) and issubclass(H, dn.hbb.TS_P2BFCOSHead) and H.__name__ == "TS_P2BFCOSHead"
assert issubclass(R, mil_head.RotatedMILHeadMixin) and issubclass(R, dn.obb.TS_P2RBRotatedFCOSHead)
for n in ("MIL_head_burn_in_step1", "MIL_head_burn_in_step2", "forward_mil_head", "mil_bag_training", "mil_bag_selection"):
    assert getattr(H, n) is getattr(mil_head.MILHeadMixin, n), n          # the mix-in wins the MRO
    assert getattr(R, n) is getattr(mil_head.MILHeadMixin, n), n
assert H.loss is dn.hbb.TS_P2BFCOSHead.loss                                # the FCOS part stays the reference's
assert R.bag_loss_pos_scale == 0.25 and R.bag_loss_neg_scale == 0.75
assert point_teacher_b200.install()["TS_P2BFCOSHead"] is H                 # idempotent
# the reference's own build_roi_extractor now yields the B200 extractor, and the head built the reference's way
# (``__new__`` + the MIL attributes, as oracle/ref_shim.build_ref_mil_head does) carries it
ext = mb.build_roi_extractor(dict(type="SingleRoIExtractor", roi_layer=dict(type="RoIAlign", output_size=7),
                                  out_channels=256, featmap_strides=[8]))
assert type(ext) is roi_extractors.SingleRoIExtractor and ext.roi_layers[0].output_size == (7, 7)
dn.hbb.TS_P2BFCOSHead = H
head = ref_shim.build_ref_mil_head(dn.hbb, num_stages=2, top_k=1, seed=0)
assert isinstance(head, mil_head.MILHeadMixin) and isinstance(head, dn.hbb.TS_P2BFCOSHead._b200_wrapped_reference)
head.bbox_roi_extractor = ext                 # what the reference's __init__ does through build_roi_extractor (:177)
assert head._dn_hyper() == 0.2 and head._feat_dtype().__str__() == "torch.float16" and head.roi_feat_area == 49
from mmdet.core.bbox.builder import BBOX_ASSIGNERS
from point_teacher_b200 import assigners
assert BBOX_ASSIGNERS.get("TopkAssigner") is assigners.TopkAssigner
assert BBOX_ASSIGNERS.get("FUSETopkAssigner") is assigners.FUSETopkAssigner
print("REGISTRY-OK")
"""


def test_install_under_openmmlab_shells():
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("reference tree not mounted")
    r = subprocess.run([sys.executable, "-W", "ignore", "-c", SCRIPT % ROOT], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "REGISTRY-OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_standalone_names_registered():
    import point_teacher_b200
    from point_teacher_b200 import mil_head, registry
    out = point_teacher_b200.install()
    if not registry.USING_OPENMMLAB:
        assert out["TS_P2BFCOSHead"] is mil_head.MILHead and out["TS_P2RBRotatedFCOSHead"] is mil_head.RotatedMILHead
    assert registry.HEADS.get("TS_P2BFCOSHead") is out["TS_P2BFCOSHead"]
