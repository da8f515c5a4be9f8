"""Registration behind the reference's registries.

The reference builds every component from config dicts through ``mmcv.utils.Registry``
(HBB_TOD/mmdet/models/builder.py:6-14, core/bbox/builder.py:3-5, iou_calculators/builder.py:3,
match_costs/builder.py:3).  When mmdet / mmrotate are importable the B200 classes are registered
into those registries under the reference's own type names (``force=True`` so they replace the
stock classes and the teacher-student detectors pick them up unchanged); otherwise a local
registry with the same ``register_module`` / ``build`` surface is used so the path is usable and
testable without OpenMMLab installed."""


class Registry:
    def __init__(self, name):
        self.name = name
        self._module_dict = {}

    def get(self, key):
        return self._module_dict.get(key)

    def register_module(self, name=None, force=False, module=None):
        def _register(cls):
            key = name or cls.__name__
            if key in self._module_dict and not force:
                raise KeyError(f"{key} is already registered in {self.name}")
            self._module_dict[key] = cls
            return cls
        if module is not None:
            return _register(module)
        return _register

    def build(self, cfg, **default_args):
        return build_from_cfg(cfg, self, default_args or None)


def build_from_cfg(cfg, registry, default_args=None):
    if not isinstance(cfg, dict) or "type" not in cfg:
        raise KeyError('cfg must be a dict containing the key "type"')
    args = dict(cfg)
    for k, v in (default_args or {}).items():
        args.setdefault(k, v)
    typ = args.pop("type")
    cls = registry.get(typ) if isinstance(typ, str) else typ
    if cls is None:
        raise KeyError(f"{typ} is not in the {registry.name} registry")
    return cls(**args)


def _try_import_registries():
    """Covered by tests/test_registry_shim.py, which runs this under oracle/ref_shim's mmcv / mmdet shells."""
    regs = {}
    try:
        from mmdet.models.builder import HEADS, ROI_EXTRACTORS, LOSSES
        from mmdet.core.bbox.builder import BBOX_ASSIGNERS, BBOX_CODERS
        from mmdet.core.bbox.iou_calculators.builder import IOU_CALCULATORS
        from mmdet.core.bbox.match_costs.builder import MATCH_COST
        regs.update(HEADS=HEADS, ROI_EXTRACTORS=ROI_EXTRACTORS, LOSSES=LOSSES, BBOX_ASSIGNERS=BBOX_ASSIGNERS,
                    BBOX_CODERS=BBOX_CODERS, IOU_CALCULATORS=IOU_CALCULATORS, MATCH_COST=MATCH_COST)
        try:
            from mmrotate.models.builder import ROTATED_ROI_EXTRACTORS, ROTATED_HEADS, ROTATED_LOSSES
            regs.update(ROTATED_ROI_EXTRACTORS=ROTATED_ROI_EXTRACTORS, ROTATED_HEADS=ROTATED_HEADS,
                        ROTATED_LOSSES=ROTATED_LOSSES)
        except Exception:
            pass
    except Exception:
        pass
    return regs


_EXTERNAL = _try_import_registries()
USING_OPENMMLAB = bool(_EXTERNAL)


def _get(name):
    return _EXTERNAL.get(name) or Registry(name)


HEADS = _get("HEADS")
ROI_EXTRACTORS = _get("ROI_EXTRACTORS")
ROTATED_ROI_EXTRACTORS = _EXTERNAL.get("ROTATED_ROI_EXTRACTORS") or ROI_EXTRACTORS
ROTATED_HEADS = _EXTERNAL.get("ROTATED_HEADS") or HEADS
LOSSES = _get("LOSSES")
ROTATED_LOSSES = _EXTERNAL.get("ROTATED_LOSSES") or LOSSES
BBOX_ASSIGNERS = _get("BBOX_ASSIGNERS")
BBOX_CODERS = _get("BBOX_CODERS")
IOU_CALCULATORS = _get("IOU_CALCULATORS")
MATCH_COST = _get("MATCH_COST")


def build_roi_extractor(cfg):
    return ROI_EXTRACTORS.build(cfg)


def build_assigner(cfg, **kw):
    return BBOX_ASSIGNERS.build(cfg, **kw)


def build_iou_calculator(cfg, default_args=None):
    return build_from_cfg(cfg, IOU_CALCULATORS, default_args)


def build_match_cost(cfg, default_args=None):
    return build_from_cfg(cfg, MATCH_COST, default_args)


def build_head(cfg):
    return HEADS.build(cfg)


# reference type name -> (registry, mix-in class name in mil_head, standalone class name in mil_head)
_REFERENCE_HEADS = {"TS_P2BFCOSHead": ("HEADS", "MILHeadMixin", "MILHead"),
                    "TS_P2RBRotatedFCOSHead": ("ROTATED_HEADS", "RotatedMILHeadMixin", "RotatedMILHead")}


def install_reference_heads(registries=None):
    """Put the B200 MIL path behind the reference's own head type names, so that the teacher-student detectors'
    configs (``bbox_head=dict(type='TS_P2BFCOSHead', ...)`` / ``'TS_P2RBRotatedFCOSHead'``, built at
    HBB_TOD/mmdet/models/detectors/fcos_p2b_teacher_student.py:40-47 through HBB_TOD/mmdet/models/builder.py:47-59)
    pick it up unchanged:

      * the reference class is in the registry (mmdet / mmrotate imported): it is re-registered (``force=True``) as a
        subclass ``type(name, (MILHeadMixin, ReferenceClass), {})`` -- the FCOS tower, losses and target code stay the
        reference's, every MIL method resolves to the mix-in first in the MRO; parameter names are the reference's own;
      * otherwise the standalone ``MILHead`` / ``RotatedMILHead`` (MIL part only) is registered under that name.

    Call again after ``import mmdet.models`` if this package was imported first.  Returns {name: class}."""
    from . import mil_head
    regs = registries or {"HEADS": HEADS, "ROTATED_HEADS": ROTATED_HEADS}
    out = {}
    for name, (reg_name, mixin_name, standalone_name) in _REFERENCE_HEADS.items():
        reg = regs.get(reg_name) or regs.get("HEADS")
        mixin, standalone = getattr(mil_head, mixin_name), getattr(mil_head, standalone_name)
        cur = reg.get(name)
        if cur is None or cur is standalone:
            cls = standalone
        elif issubclass(cur, mil_head.MILHeadMixin):
            cls = cur                                   # already wrapped
        else:
            cls = type(name, (mixin, cur), {"__module__": cur.__module__, "__doc__": cur.__doc__,
                                            "_b200_wrapped_reference": cur})
        reg.register_module(name=name, force=True, module=cls)
        out[name] = cls
    return out
