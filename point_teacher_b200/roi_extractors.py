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
import weakref

import torch
import torch.nn as nn

from . import ops
from .registry import ROI_EXTRACTORS, ROTATED_ROI_EXTRACTORS


def _pair(v):
    return (int(v), int(v)) if isinstance(v, int) else tuple(int(t) for t in v)


class FeatureRangeError(RuntimeError):
    """The fp16 NHWC feature map saturated (|value| > 65504): the fp32 reference has no such limit."""


class _NHWCCache:
    """The feature map is transposed to NHWC once per step and reused by every RoIAlign call of the step (three per
    MIL stage in the reference).

    The hit test is TENSOR IDENTITY (a weak reference to the source tensor + its version counter), never the data
    pointer: a training loop produces a fresh backbone output every step, and the caching allocator usually hands
    the same address back, so a pointer key would silently pool step N+1 from step N's features.

    fp16 maps (the tensor-core RoIAlign operand) saturate at +-65504.  Every fp16 transpose adds the number of values
    it clipped to a word in pinned host memory (written by the kernel itself, only when something clipped); the NEXT
    call looks at it (no synchronisation on the hot path) and, if anything was clipped, switches this layer to bf16 maps for good
    with a warning -- or raises ``FeatureRangeError`` when ``on_saturation == 'raise'`` (captured steps, which cannot
    switch dtype, always raise from ``CapturedPhase2.replay``)."""

    on_saturation = "fallback"          # 'fallback' (bf16 maps from now on, warn once) | 'raise'

    def __init__(self):
        self._ref, self._key, self._val = None, None, None
        self._sat_host = None
        self.force_bf16 = False
        self.saturated_total = 0

    def _poll(self, captured=False):
        """Non-blocking look at the saturation word (it may lag the device by the transposes still in flight)."""
        if self._sat_host is None:
            return
        n = int(self._sat_host[0])
        if n > self.saturated_total:
            new, self.saturated_total = n - self.saturated_total, n
            msg = (f"{new} feature values exceeded the fp16 range (+-65504) and were clipped in the NHWC map that "
                   "feeds the tensor-core RoIAlign; the fp32 reference does not clip")
            if self.on_saturation == "raise" or captured:
                raise FeatureRangeError(msg + "; build the head with feat_dtype=torch.bfloat16 (or precision='fp32')")
            import warnings
            warnings.warn(msg + "; switching this RoI layer to bf16 feature maps", RuntimeWarning, stacklevel=3)
            self.force_bf16 = True

    def check(self, sync=False, captured=False):
        """Explicit check (``sync=True`` waits for the device first; ``captured``: the caller replays a CUDA graph
        whose dtype is frozen, so a clipped value always raises)."""
        if sync:
            torch.cuda.synchronize()
        self._poll(captured)
        return self.saturated_total

    def get(self, x, dtype):
        if dtype == torch.float16:
            self._poll()
            if self.force_bf16:
                dtype = torch.bfloat16
        key = (x._version, tuple(x.shape), dtype, x.device)
        if self._ref is None or self._ref() is not x or key != self._key:
            sat = None
            if dtype == torch.float16:
                if self._sat_host is None:
                    # one int32 in PINNED HOST memory, incremented by the kernel itself through the unified address
                    # space -- and only when a value actually saturates, so the normal step pays nothing: no
                    # device-to-host copy, no event, nothing extra inside a captured graph
                    self._sat_host = torch.zeros((1,), dtype=torch.int32).pin_memory()
                sat = self._sat_host
            self._val = ops.nchw_to_nhwc(x.contiguous(), dtype, sat_count=sat)
            self._ref, self._key = weakref.ref(x), key
        return self._val

    def clear(self):
        self._ref, self._key, self._val = None, None, None


class _RoIAlignFn(torch.autograd.Function):
    """Autograd for the public (K,C,7,7) fp32 extractor output, so that the ``force=True`` replacement of
    ``SingleRoIExtractor`` stays usable by every other differentiable user of the registry.  The backward reuses
    the MIL path's scatter kernel, which consumes the output gradient as bf16 in bin-major order: the incoming fp32
    gradient is rounded to bf16 once (relative 2^-9) -- stated here because mmcv's backward is fp32 throughout."""

    @staticmethod
    def forward(ctx, input, rois, layer, roi_level, level, out):
        feat = layer.nhwc(input)
        rot = layer.rotated
        res = ops.roi_align_forward(feat, rois, ops.OUT_F32_NCHW, layer.spatial_scale, layer.sampling_ratio,
                                    layer.aligned, rotated=rot, clockwise=getattr(layer, "clockwise", True), out=out,
                                    roi_level=roi_level, level=level)
        ctx.layer, ctx.level, ctx.shape = layer, level, tuple(input.shape)
        ctx.save_for_backward(rois, roi_level if roi_level is not None else rois.new_empty(0))
        ctx.has_lvl = roi_level is not None
        if out is not None:
            ctx.mark_dirty(out)
        return res

    @staticmethod
    def backward(ctx, g):
        rois, lvl = ctx.saved_tensors
        layer = ctx.layer
        B, C, H, W = ctx.shape
        K = rois.shape[0]
        dA = g.reshape(K, C, -1).permute(0, 2, 1).to(torch.bfloat16).reshape(K, -1).contiguous()   # bin-major bf16
        dn = ops.roi_align_backward(dA, rois, (B, H, W, C), layer.spatial_scale, layer.sampling_ratio, layer.aligned,
                                    rotated=layer.rotated, clockwise=getattr(layer, "clockwise", True),
                                    roi_level=lvl if ctx.has_lvl else None, level=ctx.level)
        return ops.nhwc_to_nchw_f32(dn), None, None, None, None, None


def _roi_layer_forward(layer, input, rois, roi_level, level):
    if torch.is_grad_enabled() and input.requires_grad:
        out = None
        if roi_level is not None:      # rows of other levels are not written by the kernel: start from zeros
            out = torch.zeros((rois.shape[0], input.shape[1], *layer.output_size), dtype=torch.float32,
                              device=rois.device)
        return _RoIAlignFn.apply(input, rois, layer, roi_level, level, out)
    feat = layer.nhwc(input)
    return ops.roi_align_forward(feat, rois, ops.OUT_F32_NCHW, layer.spatial_scale, layer.sampling_ratio,
                                 layer.aligned, rotated=layer.rotated, clockwise=getattr(layer, "clockwise", True),
                                 roi_level=roi_level, level=level)


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
        return _roi_layer_forward(self, input, rois.contiguous().float(), roi_level, level)

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
        return _roi_layer_forward(self, input, rois.contiguous().float(), roi_level, level)


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
        if torch.is_grad_enabled() and any(f.requires_grad for f in feats[:num_levels]):
            # differentiable (rare) path: one zero-initialised output per level, summed (every RoI lives in one level)
            return sum(self.roi_layers[i](feats[i], rois, roi_level=lvls, level=i) for i in range(num_levels))
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
