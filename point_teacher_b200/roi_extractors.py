"""Drop-in RoI layers and extractors with the reference's constructor / attribute / forward surface.

Reference interfaces mirrored (paths under /root/reference):
  * ``mmcv.ops.RoIAlign`` / ``mmcv.ops.RoIAlignRotated`` as constructed by
    HBB_TOD/mmdet/models/roi_heads/roi_extractors/base_roi_extractor.py:50-59
    (``layer_cls(spatial_scale=1/s, **cfg)``; deprecated aliases ``out_size`` / ``sample_num``)
  * ``SingleRoIExtractor``  HBB_TOD/.../single_level_roi_extractor.py:9-114
  * ``RotatedSingleRoIExtractor``  OBB_TOD/mmrotate/models/roi_heads/roi_extractors/
    rotate_single_level_roi_extractor.py:13-167
Forward is a hand-written sm_100a kernel (csrc/roi_align.cu); there is no CPU path.
"""
import torch
import torch.nn as nn

from . import ops
from .registry import ROI_EXTRACTORS, ROTATED_ROI_EXTRACTORS


def _pair(v):
    return (int(v), int(v)) if isinstance(v, int) else tuple(int(t) for t in v)


class _NHWCCache:
    """The feature map is transposed to NHWC once per (tensor, version) and reused by every
    RoIAlign call of the step (three per MIL stage in the reference)."""

    def __init__(self):
        self._key, self._val = None, None

    def get(self, x, dtype):
        key = (x.data_ptr(), x._version, tuple(x.shape), dtype, x.device)
        if key != self._key:
            self._val = ops.nchw_to_nhwc(x.contiguous(), dtype)
            self._key = key
        return self._val

    def clear(self):
        self._key, self._val = None, None


class RoIAlign(nn.Module):
    """mmcv.ops.RoIAlign surface: ``output_size`` 2-tuple, ``spatial_scale``, ``sampling_ratio``,
    ``pool_mode``, ``aligned``.  forward(input NCHW fp32, rois (K,5)) -> (K,C,7,7) fp32."""

    def __init__(self, output_size, spatial_scale=1.0, sampling_ratio=0, pool_mode="avg", aligned=True,
                 use_torchvision=False):
        super().__init__()
        self.output_size = _pair(output_size)
        self.spatial_scale = float(spatial_scale)
        self.sampling_ratio = int(sampling_ratio)
        self.pool_mode = pool_mode
        self.aligned = aligned
        self.use_torchvision = use_torchvision
        if pool_mode != "avg":
            raise NotImplementedError("only pool_mode='avg' is on the Point Teacher path")
        if self.output_size != (7, 7):
            raise NotImplementedError("the sm_100a kernel is built for output_size=7 (the shipped config)")
        self._cache = _NHWCCache()
        self.rotated = False

    def nhwc(self, x, dtype=torch.float32):
        return self._cache.get(x, dtype)

    def forward(self, input, rois, roi_level=None, level=0):
        if rois.dim() != 2 or rois.size(1) != 5:
            raise ValueError("RoI must be (idx, x1, y1, x2, y2)!")
        feat = self.nhwc(input)
        return ops.roi_align_forward(feat, rois.contiguous().float(), ops.OUT_F32_NCHW, self.spatial_scale,
                                     self.sampling_ratio, self.aligned, roi_level=roi_level, level=level)

    def __repr__(self):
        return (f"{self.__class__.__name__}(output_size={self.output_size}, spatial_scale={self.spatial_scale}, "
                f"sampling_ratio={self.sampling_ratio}, pool_mode={self.pool_mode}, aligned={self.aligned})")


class RoIAlignRotated(nn.Module):
    """mmcv.ops.RoIAlignRotated surface (``out_size`` / ``sample_num`` accepted as deprecated aliases of
    ``output_size`` / ``sampling_ratio``).  rois (K,6) = (idx, cx, cy, w, h, theta[rad])."""

    def __init__(self, output_size=None, spatial_scale=1.0, sampling_ratio=0, aligned=True, clockwise=False,
                 out_size=None, sample_num=None):
        super().__init__()
        if output_size is None:
            output_size = out_size
        if sample_num is not None:
            sampling_ratio = sample_num
        if output_size is None:
            raise ValueError("output_size (or the deprecated out_size) is required")
        self.output_size = _pair(output_size)
        self.out_size = self.output_size
        self.spatial_scale = float(spatial_scale)
        self.sampling_ratio = int(sampling_ratio)
        self.sample_num = self.sampling_ratio
        self.aligned = aligned
        self.clockwise = clockwise
        if self.output_size != (7, 7):
            raise NotImplementedError("the sm_100a kernel is built for output_size=7 (the shipped config)")
        self._cache = _NHWCCache()
        self.rotated = True

    def nhwc(self, x, dtype=torch.float32):
        return self._cache.get(x, dtype)

    def forward(self, input, rois, roi_level=None, level=0):
        if rois.dim() != 2 or rois.size(1) != 6:
            raise ValueError("RoI must be (idx, cx, cy, w, h, theta)!")
        feat = self.nhwc(input)
        return ops.roi_align_forward(feat, rois.contiguous().float(), ops.OUT_F32_NCHW, self.spatial_scale,
                                     self.sampling_ratio, self.aligned, rotated=True, clockwise=self.clockwise,
                                     roi_level=roi_level, level=level)


_LAYERS = {"RoIAlign": RoIAlign, "RoIAlignRotated": RoIAlignRotated}


class BaseRoIExtractor(nn.Module):
    """base_roi_extractor.py:9-87."""
    rotated = False

    def __init__(self, roi_layer, out_channels, featmap_strides, init_cfg=None):
        super().__init__()
        self.init_cfg = init_cfg
        self.roi_layers = self.build_roi_layers(roi_layer, featmap_strides)
        self.out_channels = out_channels
        self.featmap_strides = featmap_strides
        self.fp16_enabled = False

    @property
    def num_inputs(self):
        return len(self.featmap_strides)

    def build_roi_layers(self, layer_cfg, featmap_strides):
        cfg = dict(layer_cfg)
        layer_type = cfg.pop("type")
        if layer_type not in _LAYERS:
            raise KeyError(f"RoI layer {layer_type} is not built for sm_100a (have {sorted(_LAYERS)})")
        return nn.ModuleList([_LAYERS[layer_type](spatial_scale=1 / s, **cfg) for s in featmap_strides])

    def roi_rescale(self, rois, scale_factor):
        return ops.roi_rescale(rois.contiguous(), scale_factor, rotated=self.rotated)

    def map_roi_levels(self, rois, num_levels):
        return ops.map_roi_levels(rois.contiguous(), num_levels, self.finest_scale, rotated=self.rotated).long()

    def forward(self, feats, rois, roi_scale_factor=None):
        """single_level_roi_extractor.py:56-114 / rotate_single_level_roi_extractor.py:90-148: one level ->
        that level's layer; several -> each RoI is pooled from its mapped level (the kernel skips the RoIs of
        other levels and writes straight into the shared output: no nonzero() sync, no gather/scatter)."""
        layer0 = self.roi_layers[0]
        out_size = layer0.output_size
        num_levels = len(feats)
        rois = rois.contiguous().float()
        if rois.size(0) == 0:
            return feats[0].new_zeros(0, self.out_channels, *out_size)
        if num_levels == 1:
            return layer0(feats[0], rois)
        lvls = ops.map_roi_levels(rois, num_levels, self.finest_scale, rotated=self.rotated)
        if roi_scale_factor is not None:
            rois = self.roi_rescale(rois, roi_scale_factor)
        out = torch.empty((rois.size(0), self.out_channels, *out_size), dtype=torch.float32, device=rois.device)
        for i in range(num_levels):
            layer = self.roi_layers[i]
            ops.roi_align_forward(layer.nhwc(feats[i]), rois, ops.OUT_F32_NCHW, layer.spatial_scale,
                                  layer.sampling_ratio, layer.aligned, rotated=self.rotated,
                                  clockwise=getattr(layer, "clockwise", True), out=out, roi_level=lvls, level=i)
        return out


@ROI_EXTRACTORS.register_module(name="SingleRoIExtractor", force=True)
class SingleRoIExtractor(BaseRoIExtractor):
    def __init__(self, roi_layer, out_channels, featmap_strides, finest_scale=56, init_cfg=None):
        super().__init__(roi_layer, out_channels, featmap_strides, init_cfg)
        self.finest_scale = finest_scale


@ROTATED_ROI_EXTRACTORS.register_module(name="RotatedSingleRoIExtractor", force=True)
class RotatedSingleRoIExtractor(BaseRoIExtractor):
    rotated = True

    def __init__(self, roi_layer, out_channels, featmap_strides, finest_scale=56, init_cfg=None):
        super().__init__(roi_layer, out_channels, featmap_strides, init_cfg)
        self.finest_scale = finest_scale
