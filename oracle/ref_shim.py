"""TEST INFRASTRUCTURE ONLY -- import shim that lets the *unmodified* reference
files under /root/reference run in the dev container (no mmcv / mmdet wheels).

It is used by ``oracle/make_golden.py`` to generate the committed fixtures in
``tests/golden/`` and by ``oracle/check_oracle_vs_ref.py`` to pin the
standalone restatement (``oracle/*.py``) bit-for-bit against the reference.
It never travels to the GPU box in any meaningful sense (``/root/reference``
does not exist there) and nothing under ``point_teacher_b200/`` may import it.

Recipe:
  * empty ``types.ModuleType`` shells with ``__path__`` set to the reference
    directories, so that the heavy ``__init__.py`` files never execute but the
    leaf files import unmodified;
  * a stub ``mmcv`` (Registry/build_from_cfg, jit/force_fp32 identity
    decorators, ``mmcv.ops.RoIAlign`` -> ``torchvision.ops.roi_align``);
  * ``mmcv.ops.nms_rotated`` / ``box_iou_rotated`` / ``RoIAlignRotated`` are
    NOT available from the reference (un-vendored mmcv kernels): they are
    bound to the restatement in ``oracle/rotated.py`` (parity unpinned for
    those three).
"""
import importlib
import inspect
import os
import sys
import types

import torch
import torch.nn as nn

REF_ROOT = os.environ.get("POINT_TEACHER_REF", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "HBB_TOD", "mmdet"))


# ----------------------------------------------------------------------------
# minimal mmcv
# ----------------------------------------------------------------------------
class Registry:
    def __init__(self, name, build_func=None, parent=None, scope=None):
        self.name = name
        self._module_dict = {}

    def get(self, key):
        return self._module_dict.get(key)

    def register_module(self, name=None, force=False, module=None):
        def _reg(cls):
            self._module_dict[name or cls.__name__] = cls
            return cls
        if module is not None:
            return _reg(module)
        return _reg

    def build(self, cfg, **kw):
        return build_from_cfg(cfg, self, kw or None)


def build_from_cfg(cfg, registry, default_args=None):
    args = dict(cfg)
    if default_args:
        for k, v in default_args.items():
            args.setdefault(k, v)
    typ = args.pop("type")
    cls = registry.get(typ) if isinstance(typ, str) else typ
    if cls is None:
        raise KeyError(f"{typ} is not in the {registry.name} registry")
    return cls(**args)


def _identity_decorator(*dargs, **dkw):
    if len(dargs) == 1 and callable(dargs[0]) and not dkw:
        return dargs[0]

    def deco(f):
        return f
    return deco


class _BaseModule(nn.Module):
    def __init__(self, init_cfg=None):
        super().__init__()
        self.init_cfg = init_cfg


class _Scale(nn.Module):
    def __init__(self, scale=1.0):
        super().__init__()
        self.scale = nn.Parameter(torch.tensor(scale, dtype=torch.float))

    def forward(self, x):
        return x * self.scale


class _RoIAlign(nn.Module):
    """mmcv.ops.RoIAlign stand-in: torchvision shares the Detectron2 avg-pool
    kernel semantics (mmcv's own ``use_torchvision=True`` switch treats them
    as interchangeable)."""

    def __init__(self, output_size, spatial_scale=1.0, sampling_ratio=0,
                 pool_mode="avg", aligned=True, use_torchvision=False):
        super().__init__()
        if isinstance(output_size, int):
            output_size = (output_size, output_size)
        self.output_size = tuple(output_size)
        self.spatial_scale = float(spatial_scale)
        self.sampling_ratio = int(sampling_ratio)
        self.pool_mode = pool_mode
        self.aligned = aligned
        assert pool_mode == "avg"

    def forward(self, x, rois):
        import torchvision
        return torchvision.ops.roi_align(x, rois, self.output_size, self.spatial_scale,
                                         self.sampling_ratio, self.aligned)


def _mk(name, path=None, **attrs):
    m = types.ModuleType(name)
    if path is not None:
        m.__path__ = [path]
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def _install_mmcv():
    from . import rotated as _rot

    class _RoIAlignRotated(nn.Module):
        def __init__(self, output_size=None, spatial_scale=1.0, sampling_ratio=0,
                     aligned=True, clockwise=False, out_size=None, sample_num=None):
            super().__init__()
            output_size = output_size if output_size is not None else out_size
            sampling_ratio = sample_num if sample_num is not None else sampling_ratio
            if isinstance(output_size, int):
                output_size = (output_size, output_size)
            self.output_size = tuple(output_size)
            self.spatial_scale = float(spatial_scale)
            self.sampling_ratio = int(sampling_ratio)
            self.aligned = aligned
            self.clockwise = clockwise

        def forward(self, x, rois):
            return _rot.roi_align_rotated(x, rois, self.output_size[0], self.spatial_scale,
                                          self.sampling_ratio, self.aligned, self.clockwise)

    mmcv = _mk("mmcv", __version__="1.7.0", jit=_identity_decorator)
    _mk("mmcv.utils", Registry=Registry, build_from_cfg=build_from_cfg)
    _mk("mmcv.cnn", Scale=_Scale, ConvModule=nn.Module)
    _mk("mmcv.runner", force_fp32=_identity_decorator, auto_fp16=_identity_decorator,
        BaseModule=_BaseModule)
    ops = _mk("mmcv.ops", RoIAlign=_RoIAlign, RoIAlignRotated=_RoIAlignRotated,
              nms_rotated=_rot.nms_rotated, box_iou_rotated=_rot.box_iou_rotated)
    mmcv.ops = ops
    mmcv.utils = sys.modules["mmcv.utils"]
    for n in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.collections"):
        _mk(n, Polygon=object, PatchCollection=object)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]


class NiceRepr:
    def __repr__(self):
        return f"<{self.__class__.__name__}>"


def multi_apply(func, *args, **kwargs):
    from functools import partial
    pfunc = partial(func, **kwargs) if kwargs else func
    map_results = map(pfunc, *args)
    return tuple(map(list, zip(*map_results)))


def _imp(name):
    return importlib.import_module(name)


_INSTALLED = None


def install(flavor="hbb"):
    """Install the shim and return a namespace with the reference callables.

    flavor: 'hbb' -> HBB_TOD/mmdet only; 'obb' -> additionally OBB_TOD/mmrotate
    leaf files (with HBB_TOD's mmdet serving as the ``mmdet`` the OBB tree
    imports).
    """
    global _INSTALLED
    if not available():
        raise RuntimeError("reference tree not present; golden fixtures are the travelling oracle")
    if _INSTALLED is not None:
        return _INSTALLED
    _install_mmcv()
    hbb = os.path.join(REF_ROOT, "HBB_TOD", "mmdet")
    p = lambda *a: os.path.join(hbb, *a)  # noqa: E731
    mmdet = _mk("mmdet", hbb, __version__="2.13.0")
    _mk("mmdet.utils", p("utils"))
    _mk("mmdet.utils.util_mixins", NiceRepr=NiceRepr)
    sys.modules["mmdet.utils"].util_mixins = sys.modules["mmdet.utils.util_mixins"]
    core = _mk("mmdet.core", p("core"))
    _mk("mmdet.core.bbox", p("core", "bbox"))
    for sub in ("assigners", "coder", "iou_calculators", "match_costs"):
        _mk(f"mmdet.core.bbox.{sub}", p("core", "bbox", sub))
    _mk("mmdet.core.utils", p("core", "utils"))
    _mk("mmdet.models", p("models"))
    for sub in ("detectors", "dense_heads", "roi_heads", "losses", "utils"):
        _mk(f"mmdet.models.{sub}", p("models", sub))
    _mk("mmdet.models.roi_heads.roi_extractors", p("models", "roi_heads", "roi_extractors"))

    # --- leaf files, in dependency order -----------------------------------
    bb_builder = _imp("mmdet.core.bbox.builder")
    iou_b = _imp("mmdet.core.bbox.iou_calculators.builder")
    iou2d = _imp("mmdet.core.bbox.iou_calculators.iou2d_calculator")
    metric = _imp("mmdet.core.bbox.iou_calculators.metric_calculator")
    ic = sys.modules["mmdet.core.bbox.iou_calculators"]
    ic.build_iou_calculator = iou_b.build_iou_calculator
    ic.bbox_overlaps = iou2d.bbox_overlaps
    ic.BboxOverlaps2D = iou2d.BboxOverlaps2D
    ic.BboxDistanceMetric = metric.BboxDistanceMetric
    tr = _imp("mmdet.core.bbox.transforms")
    mc_b = _imp("mmdet.core.bbox.match_costs.builder")
    mcost = _imp("mmdet.core.bbox.match_costs.match_cost")
    sys.modules["mmdet.core.bbox.match_costs"].build_match_cost = mc_b.build_match_cost
    ar = _imp("mmdet.core.bbox.assigners.assign_result")
    _imp("mmdet.core.bbox.assigners.base_assigner")
    topk = _imp("mmdet.core.bbox.assigners.topk_assigner")
    fuse = _imp("mmdet.core.bbox.assigners.fuse_topk_assigner")
    maxiou = _imp("mmdet.core.bbox.assigners.max_iou_assigner")
    _imp("mmdet.core.bbox.coder.base_bbox_coder")
    coder = _imp("mmdet.core.bbox.coder.delta_xywh_bbox_coder")

    for n in ("bbox_cxcywh_to_xyxy", "bbox_xyxy_to_cxcywh", "distance2bbox", "bbox2distance",
              "bbox2roi"):
        setattr(core, n, getattr(tr, n))
    core.bbox_overlaps = iou2d.bbox_overlaps
    core.multi_apply = multi_apply
    core.multiclass_nms = None
    core.reduce_mean = lambda t: t
    core.build_assigner = bb_builder.build_assigner
    core.build_sampler = bb_builder.build_sampler
    core.build_bbox_coder = bb_builder.build_bbox_coder
    core.bbox = sys.modules["mmdet.core.bbox"]
    mmdet.core = core

    # --- models ------------------------------------------------------------
    MODELS = Registry("models")
    mb = _mk("mmdet.models.builder", MODELS=MODELS, HEADS=MODELS, LOSSES=MODELS,
             ROI_EXTRACTORS=MODELS, DETECTORS=MODELS, NECKS=MODELS, BACKBONES=MODELS,
             SHARED_HEADS=MODELS,
             build_loss=lambda cfg: build_from_cfg(cfg, MODELS),
             build_roi_extractor=lambda cfg: build_from_cfg(cfg, MODELS),
             build_head=lambda cfg: build_from_cfg(cfg, MODELS))
    sys.modules["mmdet.models"].builder = mb
    mu = sys.modules["mmdet.models.utils"]
    mu.build_linear_layer = lambda cfg, *a, **k: nn.Linear(*a, **k)
    lutils = _imp("mmdet.models.losses.utils")
    _mk("mmdet.models.losses.yolo_loss", IoU_Cal=object)
    iou_loss = _imp("mmdet.models.losses.iou_loss")
    ce = _imp("mmdet.models.losses.cross_entropy_loss")
    _imp("mmdet.models.roi_heads.roi_extractors.base_roi_extractor")
    sre = _imp("mmdet.models.roi_heads.roi_extractors.single_level_roi_extractor")
    _mk("mmdet.models.dense_heads.anchor_free_head", AnchorFreeHead=nn.Module)
    syn = _imp("mmdet.models.detectors.syn_images_generator_v2")
    head = _imp("mmdet.models.dense_heads.fcos_head_p2b_ts")

    ns = types.SimpleNamespace(
        bbox_overlaps=iou2d.bbox_overlaps, BboxOverlaps2D=iou2d.BboxOverlaps2D,
        BboxDistanceMetric=metric.BboxDistanceMetric, transforms=tr, match_cost=mcost,
        AssignResult=ar.AssignResult, TopkAssigner=topk.TopkAssigner,
        FUSETopkAssigner=fuse.FUSETopkAssigner, MaxIoUAssigner=maxiou.MaxIoUAssigner,
        DeltaXYWHBBoxCoder=coder.DeltaXYWHBBoxCoder, delta2bbox=coder.delta2bbox,
        SingleRoIExtractor=sre.SingleRoIExtractor, iou_loss=iou_loss,
        DN_DIoULoss=iou_loss.DN_DIoULoss, weight_reduce_loss=lutils.weight_reduce_loss,
        expand_onehot=ce._expand_onehot_labels, syn=syn, head_mod=head,
        TS_P2BFCOSHead=head.TS_P2BFCOSHead, build_assigner=bb_builder.build_assigner,
    )
    _INSTALLED = ns
    return ns


def build_ref_mil_head(ns, num_classes=8, num_stages=1, top_k=1, beta=0.25, hyper=0.2,
                       in_channels=256, stride=8, seed=0):
    """``TS_P2BFCOSHead`` with only the MIL attributes populated (the FCOS tower
    is off the hot path); weights N(0, 0.01), biases 0 as mmdet's Normal init."""
    H = ns.TS_P2BFCOSHead
    head = H.__new__(H)
    nn.Module.__init__(head)
    head.beta, head.topk, head.num_classes, head.num_stages = beta, top_k, num_classes, num_stages
    head.in_channels = in_channels
    head.bbox_roi_extractor = ns.SingleRoIExtractor(
        roi_layer=dict(type="RoIAlign", output_size=7), out_channels=in_channels,
        featmap_strides=[stride])
    head.mil_bbox_decoder = ns.DeltaXYWHBBoxCoder(target_means=[.0, .0, .0, .0],
                                                  target_stds=[1., 1., 1., 1.])
    head.loss_bbox_denosing = ns.DN_DIoULoss(loss_weight=1.0, hyper=hyper)
    head.relu = nn.ReLU(inplace=True)
    g = torch.Generator().manual_seed(seed)

    def lin(i, o):
        l = nn.Linear(i, o)
        with torch.no_grad():
            l.weight.copy_(torch.randn(o, i, generator=g) * 0.01)
            l.bias.zero_()
        return l
    head.shared_fcs_reg, head.shared_fcs_bag = nn.ModuleList(), nn.ModuleList()
    head.fc_cls, head.fc_ins, head.fc_reg = nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
    for _ in range(num_stages):
        head.shared_fcs_reg.append(nn.ModuleList([lin(in_channels * 49, 1024), lin(1024, 1024)]))
        head.shared_fcs_bag.append(nn.ModuleList([lin(in_channels * 49, 1024), lin(1024, 1024)]))
        head.fc_cls.append(lin(1024, num_classes))
        head.fc_ins.append(lin(1024, num_classes))
        head.fc_reg.append(lin(1024, 4))
    torch.cuda.empty_cache = lambda: None
    return head


_DET = None


def install_detectors():
    """The reference's own teacher-student detector modules (only the class objects: their MIL caller methods
    ``forward_mil_head_burn_in_step1/2`` are run unbound on a stand-in ``self``, see oracle/detector.py)."""
    global _DET
    if _DET is not None:
        return _DET
    ns = install()
    _mk("mmdet.models.detectors.single_stage", SingleStageDetector=nn.Module)
    _mk("mmdet.models.detectors.base", BaseDetector=nn.Module)
    mb = sys.modules["mmdet.models.builder"]
    for n in ("build_backbone", "build_neck", "build_detector"):
        setattr(mb, n, lambda cfg, **k: None)
    sys.modules["mmdet.core"].bbox2result = None
    hbb_det = _imp("mmdet.models.detectors.fcos_p2b_teacher_student")
    o = install_obb()
    _mk("mmrotate.models.detectors.single_stage", RotatedSingleStageDetector=nn.Module)
    _mk("mmrotate.models.detectors.base", BaseDetector=nn.Module)
    rb = sys.modules["mmrotate.models.builder"]
    rb.ROTATED_DETECTORS = sys.modules["mmdet.models.builder"].MODELS
    rb.build_detector = lambda cfg, **k: None
    sys.modules["mmcv"].ConfigDict = dict
    core = sys.modules["mmrotate.core"]
    for n in ("rbbox2result", "build_assigner", "build_sampler", "obb2xyxy"):
        if not hasattr(core, n):
            setattr(core, n, None)
    obb_det = None
    try:
        obb_det = _imp("mmrotate.models.detectors.rotated_fcos_teacher_student")
    except Exception as e:  # pragma: no cover - reported by the caller
        print("OBB detector module did not import under the shim:", repr(e))
    _DET = types.SimpleNamespace(hbb=ns, obb=o, TS_P2B_FCOS=hbb_det.TS_P2B_FCOS,
                                 RotatedFCOS_TS=getattr(obb_det, "RotatedFCOS_TS", None))
    return _DET


_OBB = None


def install_obb():
    """OBB flavour: additionally import the unmodified OBB_TOD/mmrotate leaf files for the rotated MIL path
    (HBB_TOD's mmdet serves as the ``mmdet`` the OBB tree imports)."""
    global _OBB
    if _OBB is not None:
        return _OBB
    ns = install()
    from . import rotated as _rot
    obb = os.path.join(REF_ROOT, "OBB_TOD", "mmrotate")
    p = lambda *a: os.path.join(obb, *a)  # noqa: E731
    mmr = _mk("mmrotate", obb, __version__="0.3.3")
    core = _mk("mmrotate.core", p("core"))
    _mk("mmrotate.core.bbox", p("core", "bbox"))
    _mk("mmrotate.core.bbox.iou_calculators", p("core", "bbox", "iou_calculators"))
    _mk("mmrotate.core.visualization", p("core", "visualization"))
    _mk("mmrotate.core.visualization.palette", get_palette=lambda *a, **k: None)
    _mk("mmrotate.models", p("models"))
    for sub in ("detectors", "dense_heads", "roi_heads"):
        _mk(f"mmrotate.models.{sub}", p("models", sub))
    _mk("mmrotate.models.roi_heads.roi_extractors", p("models", "roi_heads", "roi_extractors"))
    # extra mmdet / mmcv names the OBB leaf files import
    _mk("mmdet.core.visualization", palette_val=lambda *a, **k: None)
    _mk("mmdet.core.visualization.image", draw_labels=None, draw_masks=None)
    _mk("mmdet.models.roi_heads.bbox_heads", None)
    _mk("mmdet.models.roi_heads.bbox_heads.bbox_head", BBoxHead=nn.Module)
    sys.modules["mmdet.models.losses"].accuracy = None
    sys.modules["mmcv.utils"].to_2tuple = lambda v: (v, v) if not isinstance(v, (tuple, list)) else tuple(v)
    sys.modules["mmcv"].digit_version = lambda v: tuple(int(x) for x in v.split(".")[:3])
    mmr.digit_version = sys.modules["mmcv"].digit_version
    mmr.mmcv_version = (1, 7, 0)
    import mmcv.ops as mops  # the stub
    mops.RiRoIAlignRotated = type("RiRoIAlignRotated", (), {})
    sys.modules["mmcv"].ops = mops
    iou_b = _imp("mmdet.core.bbox.iou_calculators.builder")
    MODELS = sys.modules["mmdet.models.builder"].MODELS
    _mk("mmrotate.models.builder", ROTATED_HEADS=MODELS, ROTATED_ROI_EXTRACTORS=MODELS, ROTATED_LOSSES=MODELS,
        build_loss=lambda cfg: build_from_cfg(cfg, MODELS), build_roi_extractor=lambda cfg: build_from_cfg(cfg, MODELS))
    sys.modules["mmrotate.models"].builder = sys.modules["mmrotate.models.builder"]
    rb = _imp("mmrotate.core.bbox.iou_calculators.builder")
    riou = _imp("mmrotate.core.bbox.iou_calculators.rotate_iou2d_calculator")
    ic = sys.modules["mmrotate.core.bbox.iou_calculators"]
    ic.build_iou_calculator, ic.rbbox_overlaps = rb.build_iou_calculator, riou.rbbox_overlaps
    tr = _imp("mmrotate.core.bbox.transforms")
    core.build_bbox_coder = sys.modules["mmdet.core"].build_bbox_coder
    core.multiclass_nms_rotated = None
    core.rbbox2roi = tr.rbbox2roi
    core.obb2poly_np = getattr(tr, "obb2poly_np", None)
    sys.modules["mmdet.core"].bbox_overlaps = ns.bbox_overlaps
    _mk("mmrotate.models.dense_heads.rotated_anchor_free_head", RotatedAnchorFreeHead=nn.Module)
    rext = _imp("mmrotate.models.roi_heads.roi_extractors.rotate_single_level_roi_extractor")
    syn = _imp("mmrotate.models.detectors.syn_images_generator_v2")
    head = _imp("mmrotate.models.dense_heads.rotated_fcos_head_p2rb_ts")
    # OBB_TOD targets the un-vendored mmdet >= 2.2x whose _expand_onehot_labels takes ``ignore_index`` and
    # returns a 3-tuple; HBB_TOD's vendored 2.13 version is wrapped to that signature.
    _old = head._expand_onehot_labels
    head._expand_onehot_labels = lambda labels, w, ch, ignore_index=None: _old(labels, w, ch) + (None,)
    # section 8f rank 4: the rotated IoU losses; their mmcv kernel (diff_iou_rotated_2d) is un-vendored and is bound to
    # the restatement in oracle/losses.py (PARITY UNPINNED for that kernel; the wrappers are the reference's own)
    from . import losses as _losses
    _mk("mmrotate.models.losses", p("models", "losses"))
    sys.modules["mmcv.ops"].diff_iou_rotated_2d = _losses.diff_iou_rotated_2d
    riou_loss = _imp("mmrotate.models.losses.rotated_iou_loss")
    focal = None
    try:
        sys.modules["mmcv.ops"].sigmoid_focal_loss = None      # only the CUDA branch uses it
        focal = _imp("mmdet.models.losses.focal_loss")
    except Exception:  # pragma: no cover
        pass
    _OBB = types.SimpleNamespace(hbb=ns, rbbox_overlaps=riou.rbbox_overlaps, transforms=tr, syn=syn,
                                 riou_loss=riou_loss, focal=focal,
                                 RotatedSingleRoIExtractor=rext.RotatedSingleRoIExtractor, head_mod=head,
                                 TS_P2RBRotatedFCOSHead=head.TS_P2RBRotatedFCOSHead)
    return _OBB


def build_ref_obb_mil_head(o, num_classes=9, num_stages=1, top_k=3, beta=0.25, hyper=0.2, in_channels=256,
                           stride=8, seed=0):
    """``TS_P2RBRotatedFCOSHead`` with only the MIL attributes populated."""
    H = o.TS_P2RBRotatedFCOSHead
    head = H.__new__(H)
    nn.Module.__init__(head)
    head.beta, head.topk, head.num_classes, head.num_stages = beta, top_k, num_classes, num_stages
    head.in_channels = in_channels
    head.bbox_roi_extractor = o.RotatedSingleRoIExtractor(
        roi_layer=dict(type="RoIAlignRotated", out_size=7, sample_num=2, clockwise=True), out_channels=in_channels,
        featmap_strides=[stride])
    head.mil_bbox_decoder = o.hbb.DeltaXYWHBBoxCoder(target_means=[.0, .0, .0, .0], target_stds=[1., 1., 1., 1.])
    head.loss_bbox_denosing = o.hbb.DN_DIoULoss(loss_weight=1.0, hyper=hyper)   # OBB does not ship DN_DIoULoss
    head.relu = nn.ReLU(inplace=True)
    g = torch.Generator().manual_seed(seed)

    def lin(i, oo):
        l = nn.Linear(i, oo)
        with torch.no_grad():
            l.weight.copy_(torch.randn(oo, i, generator=g) * 0.01)
            l.bias.zero_()
        return l
    head.shared_fcs_reg, head.shared_fcs_bag = nn.ModuleList(), nn.ModuleList()
    head.fc_cls, head.fc_ins, head.fc_reg = nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
    for _ in range(num_stages):
        head.shared_fcs_reg.append(nn.ModuleList([lin(in_channels * 49, 1024), lin(1024, 1024)]))
        head.shared_fcs_bag.append(nn.ModuleList([lin(in_channels * 49, 1024), lin(1024, 1024)]))
        head.fc_cls.append(lin(1024, num_classes))
        head.fc_ins.append(lin(1024, num_classes))
        head.fc_reg.append(lin(1024, 4))
    torch.cuda.empty_cache = lambda: None
    return head
