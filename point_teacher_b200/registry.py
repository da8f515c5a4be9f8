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
    regs = {}
    try:  # pragma: no cover - OpenMMLab is not installed in the build image
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
