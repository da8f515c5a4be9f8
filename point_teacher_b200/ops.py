"""Tensor-level wrappers over the C-ABI (include/ptb200.h).  PyTorch is used for device memory
and streams only; every computation below is a hand-written sm_100a kernel.  Shape / dtype /
device violations raise ``ValueError`` before anything is launched; a missing or failing
library raises ``PTB200Error`` -- there is no CPU or eager fallback."""
import ctypes

import torch

from . import _lib

_f32, _bf16, _u8, _i32, _i64 = torch.float32, torch.bfloat16, torch.uint8, torch.int32, torch.int64
_f16 = torch.float16
_FEAT_CODE = {_f32: 0, _bf16: 1, _f16: 2}     # feature-map dtype code of the C-ABI


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _chk(t, name, dtype=None, dim=None, last=None):
    if not isinstance(t, torch.Tensor):
        raise ValueError(f"{name}: expected a tensor")
    if not t.is_cuda:
        raise ValueError(f"{name}: expected a CUDA tensor (this path has no CPU fallback)")
    if dtype is not None and t.dtype != dtype:
        raise ValueError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if dim is not None and t.dim() != dim:
        raise ValueError(f"{name}: expected {dim} dims, got shape {tuple(t.shape)}")
    if last is not None and t.shape[-1] != last:
        raise ValueError(f"{name}: expected last dim {last}, got shape {tuple(t.shape)}")
    if not t.is_contiguous():
        raise ValueError(f"{name}: expected a contiguous tensor")
    return t


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _farr(vals):
    vals = [float(v) for v in vals]
    return (ctypes.c_float * max(len(vals), 1))(*vals), len(vals)


_SIDE = {}


class fork:
    """``with ops.fork() as f: <launches>`` runs the launches on a side stream forked from the current stream;
    ``f.join()`` makes the current stream wait for them.  Used for work that is off the critical path of the step
    (logged IoU means); graph-capture safe as long as every fork is joined before the capture ends."""

    def __init__(self, lane=1):
        self.lane = lane          # independent side streams: work forked on different lanes does not queue up

    def __enter__(self):
        main = torch.cuda.current_stream()
        key = (main.device.index, self.lane)
        if key not in _SIDE:
            _SIDE[key] = torch.cuda.Stream(device=main.device)
        self.side = _SIDE[key]
        self.side.wait_stream(main)
        self.ctx = torch.cuda.stream(self.side)
        self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        self.event = torch.cuda.Event()
        self.event.record(self.side)
        return self.ctx.__exit__(*exc)

    def join(self):
        torch.cuda.current_stream().wait_event(self.event)


def num_sms():
    return torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count


# SMs the persistent GEMM leaves free (its grid is num_sms() - reserve CTAs).  The trainer sets it while a gradient
# all-reduce is in flight on NCCL's stream: the GEMM assigns tiles to CTAs statically, so CTAs that cannot become
# resident beside NCCL's would otherwise hold their tiles until the collective has finished.
_SM_RESERVE = {"n": 0}


def set_gemm_sm_reserve(n):
    _SM_RESERVE["n"] = max(0, int(n))


def gemm_sms():
    return max(8, num_sms() - _SM_RESERVE["n"])


# ------------------------------------------------------------------------------ bags
def bag_gen(in_rois, img_wh, base_ratios, shake_ratio, min_scale, rotated=False):
    """in_rois (G,5) -> (rois (G*U,5), valid (G*U,) uint8); rotated: 6-column RoIs (img,cx,cy,w,h,theta)."""
    _chk(in_rois, "in_rois", _f32, 2, 6 if rotated else 5)
    _chk(img_wh, "img_wh", _f32, 2, 2)
    ratios, nr = _farr(base_ratios)
    shake, ns = _farr(shake_ratio or [])
    U = nr * nr * (1 + 4 * ns)
    G = in_rois.shape[0]
    out = torch.empty((G * U, in_rois.shape[1]), dtype=_f32, device=in_rois.device)
    valid = torch.empty((G * U,), dtype=_u8, device=in_rois.device)
    _lib.call("pt_bag_gen", _p(in_rois), G, _p(img_wh), img_wh.shape[0], ratios, nr, shake, ns, float(min_scale),
              _p(out), _p(valid), int(rotated), _stream())
    return out, valid


def make_rois(boxes, img_idx, out=None):
    """boxes (n,4|5) + img_idx (n,) int32 -> rois (n,5|6); ``out`` may be a row-slice of a larger RoI buffer."""
    _chk(boxes, "boxes", _f32, 2)
    _chk(img_idx, "img_idx", _i32, 1)
    n, d = boxes.shape
    if out is None:
        out = torch.empty((n, d + 1), dtype=_f32, device=boxes.device)
    _lib.call("pt_make_rois", _p(boxes), d, _p(img_idx), n, d, _p(out), _stream())
    return out


def neg_weight(neg_rois, bag_rois, bag_offsets, rotated=False):
    _chk(neg_rois, "neg_rois", _f32, 2, 6 if rotated else 5)
    _chk(bag_rois, "bag_rois", _f32, 2, 6 if rotated else 5)
    _chk(bag_offsets, "bag_offsets", _i32, 1)
    w = torch.empty((neg_rois.shape[0],), dtype=_u8, device=neg_rois.device)
    _lib.call("pt_neg_weight", _p(neg_rois), neg_rois.shape[0], _p(bag_rois), _p(bag_offsets),
              bag_offsets.shape[0] - 1, _p(w), int(rotated), _stream())
    return w


def box_iou_rotated(b1, b2, mode="iou", aligned=False, clamp_wh=False):
    """(M,5) x (N,5) (cx,cy,w,h,theta) -> (M,N) | aligned (M,)."""
    if mode not in ("iou", "iof"):
        raise ValueError(f"Unsupported mode {mode}")
    _chk(b1, "bboxes1", _f32, 2)
    _chk(b2, "bboxes2", _f32, 2)
    M, N = b1.shape[0], b2.shape[0]
    if aligned and M != N:
        raise ValueError("aligned requires the same number of boxes")
    out = torch.empty((M,) if aligned else (M, N), dtype=_f32, device=b1.device)
    if M * N == 0:
        return out
    if b1.shape[1] < 5 or b2.shape[1] < 5:
        raise ValueError("rotated boxes need 5 columns")
    _lib.call("pt_box_iou_rotated", _p(b1), b1.shape[1], _p(b2), b2.shape[1], M, N, 0 if mode == "iou" else 1,
              int(aligned), int(clamp_wh), _p(out), _stream())
    return out


_MODES = {"iou": 0, "iof": 1, "giou": 2}


def bbox_overlaps(b1, b2, mode="iou", is_aligned=False, eps=1e-6):
    if mode not in _MODES:
        raise ValueError(f"Unsupported mode {mode}")
    _chk(b1, "bboxes1", _f32, 2)
    _chk(b2, "bboxes2", _f32, 2)
    M, N = b1.shape[0], b2.shape[0]
    if is_aligned and M != N:
        raise ValueError("is_aligned requires the same number of boxes")
    out = torch.empty((M,) if is_aligned else (M, N), dtype=_f32, device=b1.device)
    if M * N == 0:
        return out
    if b1.shape[1] < 4 or b2.shape[1] < 4:
        raise ValueError("boxes must have at least 4 columns")
    _lib.call("pt_bbox_overlaps", _p(b1), b1.shape[1], _p(b2), b2.shape[1], M, N, _MODES[mode], int(is_aligned),
              float(eps), _p(out), _stream())
    return out


# ------------------------------------------------------------------------------ RoIAlign
def nchw_to_nhwc(x, out_dtype=_f32, sat_count=None):
    """``sat_count``: optional (1,) int32 device counter; an fp16 output adds the number of saturated values."""
    _chk(x, "feat", _f32, 4)
    B, C, H, W = x.shape
    out = torch.empty((B, H, W, C), dtype=out_dtype, device=x.device)
    if out_dtype not in _FEAT_CODE:
        raise ValueError("NHWC feature map must be fp32, bf16 or fp16")
    if sat_count is not None:      # device int32, or PINNED host int32 (reached through the unified address space)
        if sat_count.dtype != _i32 or sat_count.numel() != 1 or not (sat_count.is_cuda or sat_count.is_pinned()):
            raise ValueError("sat_count: expected one int32 in device or pinned host memory")
    _lib.call("pt_nchw_to_nhwc_ex", _p(x), _p(out), B, C, H, W, _FEAT_CODE[out_dtype], _p(sat_count), _stream())
    return out


OUT_BF16_BINMAJOR, OUT_F32_NCHW, OUT_BF16X3_BINMAJOR = 0, 1, 2


def roi_align_forward(feat_nhwc, rois, out_mode, spatial_scale, sampling_ratio=0, aligned=True, rotated=False,
                      clockwise=True, pooled=7, out=None, rows=None, roi_level=None, level=0):
    """feat_nhwc (B,H,W,C) fp32|bf16; rois (K,5) or rotated (K,6).  ``out``/``rows`` let the caller
    hand in a (row-padded) GEMM operand buffer."""
    if feat_nhwc.dtype not in _FEAT_CODE:
        raise ValueError("feat must be fp32, bf16 or fp16")
    _chk(feat_nhwc, "feat_nhwc", None, 4)
    _chk(rois, "rois", _f32, 2, 6 if rotated else 5)
    B, H, W, C = feat_nhwc.shape
    K = rois.shape[0]
    kcols = pooled * pooled * C
    if out is None:
        if out_mode == OUT_F32_NCHW:
            out = torch.empty((K, C, pooled, pooled), dtype=_f32, device=rois.device)
        else:
            out = torch.empty((rows or K, kcols * (3 if out_mode == OUT_BF16X3_BINMAJOR else 1)), dtype=_bf16,
                              device=rois.device)
    ld = out.shape[1] if out_mode != OUT_F32_NCHW else 0
    if PROFILE["on"]:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    _roi_call(feat_nhwc, rois, out, ld, out_mode, K, B, C, H, W, pooled, spatial_scale, sampling_ratio, aligned,
              rotated, clockwise, roi_level, level)
    if PROFILE["on"]:
        e1.record()
        esz = 4 if out_mode == OUT_F32_NCHW else (6 if out_mode == OUT_BF16X3_BINMAJOR else 2)
        nbytes = K * kcols * esz + feat_nhwc.numel() * feat_nhwc.element_size() + K * rois.shape[1] * 4
        PROFILE["events"].append(("roi_align", e0, e1, float(nbytes), (K, C, out_mode)))
    return out


def _roi_call(feat_nhwc, rois, out, ld, out_mode, K, B, C, H, W, pooled, spatial_scale, sampling_ratio, aligned,
              rotated, clockwise, roi_level, level):
    _lib.call("pt_roi_align_forward", _p(feat_nhwc), _FEAT_CODE[feat_nhwc.dtype], _p(rois), _p(out), ld, out_mode,
              K, B, C, H, W, pooled, float(spatial_scale), int(sampling_ratio), int(aligned), int(rotated),
              int(clockwise), _p(roi_level), int(level), _stream())


def map_roi_levels(rois, num_levels, finest_scale=56, rotated=False):
    _chk(rois, "rois", _f32, 2, 6 if rotated else 5)
    lv = torch.empty((rois.shape[0],), dtype=_i32, device=rois.device)
    _lib.call("pt_map_roi_levels", _p(rois), rois.shape[0], int(rotated), float(finest_scale), int(num_levels),
              _p(lv), _stream())
    return lv


def roi_rescale(rois, factor, rotated=False):
    _chk(rois, "rois", _f32, 2, 6 if rotated else 5)
    fh, fw = (factor, factor) if not isinstance(factor, (tuple, list)) else factor
    out = torch.empty_like(rois)
    _lib.call("pt_roi_rescale", _p(rois), rois.shape[0], int(rotated), float(fh), float(fw), _p(out), _stream())
    return out


# ------------------------------------------------------------------------------ FC GEMM
_WS = {}
# optional per-launch timing hook for bench.py: {"events": [(tag, start, stop, flops)], "on": bool}
PROFILE = {"on": False, "events": []}


def gemm_workspace(device):
    """Split-K workspace (partial slots + counters that must be zero between launches).  Two GEMMs running
    concurrently on different streams must not share it, so eager launches get one workspace per (device, stream);
    launches recorded into CUDA graphs share one capture workspace per device (graph replays are stream-ordered by
    their callers), allocated ahead of the capture so that no allocation / memset is recorded into a graph."""
    dev = device.index if device.index is not None else torch.cuda.current_device()
    n = None
    cap_key = (dev, "capture")
    if cap_key not in _WS and not torch.cuda.is_current_stream_capturing():
        n = _lib.load().pt_fc_gemm_workspace_bytes(num_sms())
        _WS[cap_key] = torch.zeros((n,), dtype=_u8, device=device)
    if torch.cuda.is_current_stream_capturing():
        key = cap_key
    else:
        key = (dev, torch.cuda.current_stream(device).cuda_stream)
    if key not in _WS:
        n = n or _lib.load().pt_fc_gemm_workspace_bytes(num_sms())
        _WS[key] = torch.zeros((n,), dtype=_u8, device=device)  # zeroed once; the kernel leaves it zeroed
    return _WS[key]


def fc_gemm(A, B, bias=None, relu=False, out_dtype=_bf16, M=None, out=None, allow_split=True):
    """C[M,N] = act(A[M,K] @ B[N,K]^T + bias) on tcgen05 tensor cores; A, B bf16."""
    _chk(A, "A", _bf16, 2)
    _chk(B, "B", _bf16, 2)
    if A.shape[1] != B.shape[1]:
        raise ValueError(f"K mismatch: A {tuple(A.shape)} vs B {tuple(B.shape)}")
    M = A.shape[0] if M is None else M
    N, K = B.shape
    if bias is not None:
        _chk(bias, "bias", _f32, 1)
        if bias.shape[0] != N:
            raise ValueError("bias length must equal N")
    if out is None:
        out = torch.empty((A.shape[0], N), dtype=out_dtype, device=A.device)
    ws = gemm_workspace(A.device)
    if PROFILE["on"]:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    _lib.call("pt_fc_gemm_bf16", _p(A), A.shape[1], _p(B), B.shape[1], _p(bias), _p(out), out.shape[1], M, N, K,
              int(relu), int(out.dtype == _f32), _p(ws), ws.numel(), gemm_sms(), int(allow_split), _stream())
    if PROFILE["on"]:
        e1.record()
        PROFILE["events"].append(("fc_gemm", e0, e1, 2.0 * M * N * K, (M, N, K)))
    return out


def prep_fc1_weight(w, C, bins=49, x3=False):
    _chk(w, "fc1 weight", _f32, 2)
    N, K = w.shape
    if K != C * bins:
        raise ValueError(f"fc1 weight has {K} inputs, expected {C}*{bins}")
    out = torch.empty((N, K * (3 if x3 else 1)), dtype=_bf16, device=w.device)
    _lib.call("pt_prep_fc1_weight", _p(w), _p(out), N, C, bins, out.shape[1], int(x3), _stream())
    return out


def cast_weight(w, x3=False):
    _chk(w, "weight", _f32, 2)
    N, K = w.shape
    out = torch.empty((N, K * (3 if x3 else 1)), dtype=_bf16, device=w.device)
    _lib.call("pt_cast_weight_bf16", _p(w), _p(out), N, K, out.shape[1], int(x3), _stream())
    return out


# ------------------------------------------------------------------------------ head tails
def reg_decode(H, Wreg, breg, bag_rois, valid, ref_boxes, real_boxes, U, max_wh, sums, K=None, hyper=0.2,
               eps=1e-6, wh_ratio_clip=16 / 1000, want_deltas=False, out_rois=None, rotated=False):
    if H.dtype not in (_f32, _bf16):
        raise ValueError("hidden must be fp32 or bf16")
    _chk(H, "hidden", None, 2)
    _chk(Wreg, "fc_reg.weight", _f32, 2)
    rs = 6 if rotated else 5
    _chk(bag_rois, "bag_rois", _f32, 2, rs)
    K = bag_rois.shape[0] if K is None else K
    dev = H.device
    if out_rois is None:
        out_rois = torch.empty((K, rs), dtype=_f32, device=dev)
    else:
        _chk(out_rois, "out_rois", _f32, 2, rs)
        if out_rois.shape[0] < K:
            raise ValueError("out_rois has fewer rows than K")
    deltas = torch.empty((K, 4), dtype=_f32, device=dev) if want_deltas else None
    iou_t = torch.empty((K,), dtype=_f32, device=dev)
    deltas_in = None
    if H.dtype == _bf16 and H.shape[1] % 256 == 0:         # H . Wreg^T on tensor cores; the kernel below only decodes
        deltas_in = torch.empty((K, 4), dtype=_f32, device=dev)
        _lib.call("pt_small_heads_bf16", _p(H), H.shape[1], Wreg.shape[1], _p(Wreg), 4, _p(None), _p(None), 0, _p(None),
                  K, _p(deltas_in), _p(None), _stream())
    _lib.call("pt_reg_decode", _p(H), int(H.dtype == _f32), H.shape[1], Wreg.shape[1], _p(Wreg), _p(breg),
              _p(bag_rois), _p(valid), _p(ref_boxes), _p(real_boxes), U, K, float(max_wh[0]), float(max_wh[1]),
              float(wh_ratio_clip), float(hyper), float(eps), _p(out_rois), _p(deltas), _p(iou_t), _p(sums),
              int(rotated), _p(deltas_in), _stream())
    return out_rois, deltas, iou_t


def cls_ins_heads(H, Wcls, bcls, Wins, bins, M=None):
    _chk(H, "hidden", None, 2)
    M = H.shape[0] if M is None else M
    C = Wcls.shape[0]
    cls = torch.empty((M, C), dtype=_f32, device=H.device)
    ins = torch.empty((M, C), dtype=_f32, device=H.device)
    if H.dtype == _bf16 and H.shape[1] % 256 == 0 and 2 * C <= 32:       # tensor-core path (heads_mma.cu)
        _lib.call("pt_small_heads_bf16", _p(H), H.shape[1], Wcls.shape[1], _p(Wcls), C, _p(bcls), _p(Wins), C, _p(bins), M,
                  _p(cls), _p(ins), _stream())
        return cls, ins
    _lib.call("pt_cls_ins_heads", _p(H), int(H.dtype == _f32), H.shape[1], Wcls.shape[1], _p(Wcls), _p(bcls),
              _p(Wins), _p(bins), C, M, _p(cls), _p(ins), _stream())
    return cls, ins


def score_select(cls, ins, valid, bag_rois, labels, pseudo, img_wh, G, U1, U2, topk, beta, sums, rotated=False):
    """Returns (merged (G,4), merged centres (G,2), selected idx (G,topk) int32, scores (G,topk))."""
    _chk(cls, "cls", _f32, 2)
    _chk(ins, "ins", _f32, 2)
    _chk(labels, "labels", _i64, 1)
    bd = 5 if rotated else 4
    if pseudo is not None:
        _chk(pseudo, "pseudo_boxes", _f32, 2, bd)
    C = cls.shape[1]
    dev = cls.device
    merged = torch.empty((G, bd), dtype=_f32, device=dev)
    pts = torch.empty((G, 2), dtype=_f32, device=dev)
    idx = torch.empty((G, topk), dtype=_i32, device=dev)
    sc = torch.empty((G, topk), dtype=_f32, device=dev)
    _lib.call("pt_score_select", _p(cls), _p(ins), _p(valid), _p(bag_rois), _p(labels), _p(pseudo), _p(img_wh),
              img_wh.shape[0], G, U1, U2, C, topk, float(beta), _p(merged), _p(pts), _p(idx), _p(sc), _p(sums),
              int(rotated), _stream())
    return merged, pts, idx, sc


def neg_loss(neg_cls, weight, sums, n=None):
    n = neg_cls.shape[0] if n is None else n
    _lib.call("pt_neg_loss", _p(neg_cls), _p(weight), n, neg_cls.shape[1], _p(sums), _stream())


def finalize_losses(sums, K, has_neg, scale_bbox=1.0, scale_bags=1.0, pos_w=1.0, neg_w=1.0):
    out = torch.empty((5,), dtype=_f32, device=sums.device)
    _lib.call("pt_finalize_losses", _p(sums), K, int(has_neg), float(scale_bbox), float(scale_bags), float(pos_w),
              float(neg_w), _p(out), _stream())
    return out


def split_bf16x3(x):
    _chk(x, "x", _f32, 2)
    out = torch.empty((x.shape[0], 3 * x.shape[1]), dtype=_bf16, device=x.device)
    _lib.call("pt_split_bf16x3", _p(x), _p(out), x.shape[0], x.shape[1], _stream())
    return out


def aligned_iou_mean(a, b, rotated=False):
    _chk(a, "a", _f32, 2)
    _chk(b, "b", _f32, 2)
    if a.shape[0] != b.shape[0]:
        raise ValueError("aligned IoU needs the same number of boxes")
    out = torch.empty((1,), dtype=_f32, device=a.device)
    _lib.call("pt_aligned_iou_mean", _p(a), a.shape[1], _p(b), b.shape[1], a.shape[0], int(rotated), _p(out),
              _stream())
    return out[0]


# ------------------------------------------------------------------------------ dense-head label assignment
METRIC_MODES = {"iou": 0, "iof": 1, "giou": 2, "wd": 3, "kl": 4, "center_distance2": 5, "exp_kl": 6, "kl_10": 7}


def focal_cost_table(logits, alpha=0.25, gamma=2.0, eps=1e-12, weight=1.0):
    """(P, C) logits -> (P, C) FocalLossCost table (pos - neg) * weight."""
    _chk(logits, "cls_pred", _f32, 2)
    out = torch.empty_like(logits)
    _lib.call("pt_focal_cost_table", _p(logits), logits.numel(), float(alpha), float(gamma), float(eps), float(weight),
              _p(out), _stream())
    return out


def topk_pre(points, gts, num_pre, mode="L1", weight=1.0):
    """torch.topk(PointCost(points, gts), num_pre, dim=0, largest=False).indices as (num_pre, G) int32."""
    _chk(points, "points", _f32, 2)
    _chk(gts, "gt_bboxes", _f32, 2)
    if points.shape[1] < 2 or gts.shape[1] < 2:
        raise ValueError("points / gt_bboxes need at least 2 columns")
    P, G = points.shape[0], gts.shape[0]
    if num_pre > P:
        raise RuntimeError("selected index k out of range")          # what torch.topk raises
    pre = torch.empty((num_pre, G), dtype=_i32, device=points.device)
    sv = si = None
    if num_pre * 64 > P:
        sv = torch.empty((G, P), dtype=_f32, device=points.device)
        si = torch.empty((G, P), dtype=_i32, device=points.device)
    _lib.call("pt_topk_pre", _p(points), points.shape[1], P, _p(gts), gts.shape[1], G, int(mode == "L2"), float(weight),
              int(num_pre), _p(pre), _p(sv), _p(si), _stream())
    return pre


def topk_second(pre_idx, topk, P, fl_table, gt_labels, gts, pred=None, loc_weight=1.0):
    """-> (gt_inds (P,) int64, labels (P,) int64)."""
    num_pre, G = pre_idx.shape
    _chk(gt_labels, "gt_labels", _i64, 1)
    dev = pre_idx.device
    ws = torch.empty((P,), dtype=_i32, device=dev)
    gt_inds = torch.empty((P,), dtype=_i64, device=dev)
    labels = torch.empty((P,), dtype=_i64, device=dev)
    C = fl_table.shape[1] if fl_table is not None else 0
    _lib.call("pt_topk_second", _p(pre_idx), num_pre, int(topk), G, P, _p(fl_table), C, _p(gt_labels), _p(pred),
              pred.shape[1] if pred is not None else 0, _p(gts), gts.shape[1], float(loc_weight), _p(ws), _p(gt_inds),
              _p(labels), _stream())
    return gt_inds, labels


def bbox_metric(b1, b2, mode="iou", calc=1, eps=1e-6):
    """(M,4) x (N,4) -> (M,N): calc 0 BboxOverlaps2D, calc 1 BboxDistanceMetric."""
    if mode not in METRIC_MODES or (calc == 0 and METRIC_MODES[mode] > 2):
        raise AssertionError(f"Unsupported mode {mode}")
    _chk(b1, "bboxes1", _f32, 2)
    _chk(b2, "bboxes2", _f32, 2)
    M, N = b1.shape[0], b2.shape[0]
    out = torch.empty((M, N), dtype=_f32, device=b1.device)
    if M * N == 0:
        return out
    if b1.shape[1] < 4 or b2.shape[1] < 4:
        raise ValueError("boxes must have at least 4 columns")
    _lib.call("pt_bbox_metric", _p(b1), b1.shape[1], _p(b2), b2.shape[1], M, N, int(calc), METRIC_MODES[mode], float(eps),
              _p(out), _stream())
    return out


def max_iou_assign(gts, anchors, calc, mode, pos_iou_thr, neg_iou_thr, min_pos_iou, gt_max_assign_all,
                   match_low_quality, gt_labels=None, eps=1e-6):
    """-> (gt_inds (A,) int64, max_overlaps (A,) fp32, labels (A,) int64 | None); G > 0 and A > 0."""
    _chk(gts, "gt_bboxes", _f32, 2)
    _chk(anchors, "bboxes", _f32, 2)
    G, A = gts.shape[0], anchors.shape[0]
    dev = anchors.device
    gt_inds = torch.empty((A,), dtype=_i64, device=dev)
    mx = torch.empty((A,), dtype=_f32, device=dev)
    labels = torch.empty((A,), dtype=_i64, device=dev) if gt_labels is not None else None
    if gt_labels is not None:
        _chk(gt_labels, "gt_labels", _i64, 1)
    amx = torch.empty((A,), dtype=_i32, device=dev)
    ws = torch.empty((2 * G,), dtype=_i32, device=dev)
    if isinstance(neg_iou_thr, (tuple, list)):
        lo, hi = float(neg_iou_thr[0]), float(neg_iou_thr[1])
    else:
        lo, hi = 0.0, float(neg_iou_thr)
    _lib.call("pt_max_iou_assign", _p(gts), gts.shape[1], G, _p(anchors), anchors.shape[1], A, int(calc),
              METRIC_MODES[mode], float(eps), float(pos_iou_thr), lo, hi, float(min_pos_iou), int(gt_max_assign_all),
              int(match_low_quality), _p(gt_labels), _p(gt_inds), _p(mx), _p(labels), _p(amx), _p(ws), _stream())
    return gt_inds, mx, labels


# ------------------------------------------------------------------------------ phase-1 region masking
def nms_rotated(dets, scores, iou_threshold):
    """mmcv.ops.nms_rotated decision list: (order (N,) int32 by descending score, keep_sorted (N,) uint8)."""
    _chk(dets, "dets", _f32, 2)
    if dets.shape[1] < 5:
        raise ValueError("dets need (cx, cy, w, h, theta)")
    if scores.dtype != _f32 or scores.dim() != 1 or not scores.is_cuda or scores.shape[0] != dets.shape[0]:
        raise ValueError("scores must be a CUDA fp32 vector with one entry per box")
    N = dets.shape[0]
    dev = dets.device
    order = torch.empty((N,), dtype=_i32, device=dev)
    keep = torch.empty((N,), dtype=_u8, device=dev)
    if N == 0:
        return order, keep
    nbytes = _lib.load().pt_nms_rotated_workspace_bytes(N)
    ws = torch.empty((nbytes,), dtype=_u8, device=dev)
    _lib.call("pt_nms_rotated", _p(dets), dets.stride(0), _p(scores), scores.stride(0), N, float(iou_threshold),
              _p(order), _p(keep), _p(ws), nbytes, _stream())
    return order, keep


def black_paper_select(bb, order, keep_sorted, imgsize, trig=None):
    """-> (kept boxes (N,7) padded, sel (N,) int32 padded, polys (N,4,2) int32 padded, count (1,) int32).
    ``trig`` (N,2) fp32 = host-computed (sin, cos) of column 4 (bit-exact corners); None = device sincos."""
    _chk(bb, "bb", _f32, 2, 7)
    if trig is not None:
        _chk(trig, "trig", _f32, 2, 2)
        if trig.shape[0] != bb.shape[0]:
            raise ValueError("trig needs one (sin, cos) row per box")
    N, dev = bb.shape[0], bb.device
    out = torch.zeros((N, 7), dtype=_f32, device=dev)
    sel = torch.zeros((N,), dtype=_i32, device=dev)
    polys = torch.zeros((N, 4, 2), dtype=_i32, device=dev)
    count = torch.zeros((1,), dtype=_i32, device=dev)
    _lib.call("pt_black_paper_select_ex", _p(bb), N, _p(order), _p(keep_sorted), float(imgsize), _p(out), _p(sel),
              _p(polys), _p(count), _p(trig), _stream())
    return out, sel, polys, count


def fill_polys(polys, img=None, mask=None, value=255.0, count=None):
    """cv2.fillPoly of (M,4,2) int32 quadrilaterals into img (C,H,W) fp32 (<- value) and/or mask (H,W) uint8."""
    _chk(polys, "polygons", _i32, 3, 2)
    if polys.shape[1] != 4:
        raise ValueError("only quadrilaterals are on the Point Teacher path")
    if img is None and mask is None:
        raise ValueError("nothing to draw into")
    if img is not None:
        _chk(img, "img", _f32, 3)
        C, H, W = img.shape
    if mask is not None:
        _chk(mask, "mask", _u8, 2)
        if img is not None and tuple(mask.shape) != (H, W):
            raise ValueError("mask and img disagree on (H, W)")
        C, (H, W) = (C if img is not None else 0), mask.shape
    _lib.call("pt_fill_polys", _p(polys), _p(count), polys.shape[0], _p(img), _p(mask), C, H, W, float(value),
              _stream())


# ------------------------------------------------------------------------------ backward of the MIL head
def fc_gemm_masked(A, B, mask, out_dtype=_bf16, M=None, allow_split=True):
    """C = (A @ B^T) * (mask > 0): dgrad through a ReLU layer, mask = the saved bf16 forward activation."""
    _chk(A, "A", _bf16, 2)
    _chk(B, "B", _bf16, 2)
    _chk(mask, "mask", _bf16, 2)
    M = A.shape[0] if M is None else M
    N, K = B.shape
    if A.shape[1] != K or mask.shape[1] != N or mask.shape[0] < M:
        raise ValueError("fc_gemm_masked: shape mismatch")
    out = torch.empty((A.shape[0], N), dtype=out_dtype, device=A.device)
    ws = gemm_workspace(A.device)
    _lib.call("pt_fc_gemm_bf16_ex", _p(A), A.shape[1], _p(B), B.shape[1], _p(None), _p(out), out.shape[1], M, N, K, 0,
              int(out.dtype == _f32), _p(mask), mask.shape[1], _p(ws), ws.numel(), gemm_sms(), int(allow_split), _stream())
    return out


def fc_gemm_mn(A, B, a_mn=False, b_mn=False, mask=None, out_dtype=_bf16, M=None, K=None, allow_split=True, out=None):
    """C[M,N] = A' @ B'^T with operands as they are stored: ``a_mn`` -> A is [K, M] (else [M, K]); ``b_mn`` -> B is
    [K, N] (else [N, K]).  ``M`` / ``K`` restrict to the leading rows of padded buffers.  Optional ReLU-backward
    ``mask`` [>=M, N] bf16.  nn.Linear autograd (dW = dY^T X, dX = dY W) with no transposed copies."""
    _chk(A, "A", _bf16, 2)
    _chk(B, "B", _bf16, 2)
    Ka, Ma = (A.shape[0], A.shape[1]) if a_mn else (A.shape[1], A.shape[0])
    Kb, N = (B.shape[0], B.shape[1]) if b_mn else (B.shape[1], B.shape[0])
    if a_mn:
        K = min(Ka, Kb) if K is None else K
        M = Ma if M is None else M
    else:
        M = Ma if M is None else M
        K = Ka if K is None else K
    if K > Ka or K > Kb or M > Ma:
        raise ValueError(f"fc_gemm_mn: M={M} K={K} exceed the operands A {tuple(A.shape)} B {tuple(B.shape)}")
    rows_out = M if a_mn else A.shape[0]
    if mask is not None:
        _chk(mask, "mask", _bf16, 2)
        if mask.shape[1] != N or mask.shape[0] < M:
            raise ValueError("fc_gemm_mn: mask shape mismatch")
    if out is None:
        out = torch.empty((rows_out, N), dtype=out_dtype, device=A.device)
    else:
        if out.dim() != 2 or out.shape[0] < rows_out or out.shape[1] != N or out.stride(1) != 1 or \
                out.dtype not in (_bf16, _f32) or (out.data_ptr() & 15) or (out.stride(0) * out.element_size()) & 15:
            raise ValueError("fc_gemm_mn: out must be a 16-byte aligned row-major [>=M, N] bf16 / fp32 tensor")
    ws = gemm_workspace(A.device)
    _lib.call("pt_fc_gemm_bf16_mn", _p(A), A.shape[1], int(a_mn), _p(B), B.shape[1], int(b_mn), _p(None), _p(out),
              out.stride(0), M, N, K, 0, int(out.dtype == _f32), _p(mask), 0 if mask is None else mask.shape[1], _p(ws),
              ws.numel(), gemm_sms(), int(allow_split), _stream())
    return out


def reg_loss_grad(deltas, bag_rois, valid, ref_boxes, U, max_wh, sums, gscale, scale, hyper=0.2, eps=1e-6,
                  wh_ratio_clip=16 / 1000, rotated=False):
    _chk(deltas, "deltas", _f32, 2, 4)
    _chk(bag_rois, "bag_rois", _f32, 2, 6 if rotated else 5)
    _chk(ref_boxes, "ref_boxes", _f32, 2, 5 if rotated else 4)
    K = deltas.shape[0]
    g = torch.empty((K, 4), dtype=_f32, device=deltas.device)
    _lib.call("pt_reg_loss_grad_ex", _p(deltas), _p(bag_rois), _p(valid), _p(ref_boxes), int(U), K, float(max_wh[0]),
              float(max_wh[1]), float(wh_ratio_clip), float(hyper), float(eps), _p(sums), _p(gscale), float(scale), _p(g),
              int(rotated), _stream())
    return g


def bag_loss_grad(cls, ins, valid, labels, G, U1, U2, neg_weight, n_neg, sums, gscale, pos_scale, neg_scale):
    _chk(cls, "cls", _f32, 2)
    _chk(ins, "ins", _f32, 2)
    C = cls.shape[1]
    M = G * U1 * U2 + n_neg
    if cls.shape[0] < M:
        raise ValueError("cls has fewer rows than bags + negatives")
    g = torch.zeros((M, 2 * C), dtype=_f32, device=cls.device)
    _lib.call("pt_bag_loss_grad", _p(cls), _p(ins), _p(valid), _p(labels), G, U1, U2, C, _p(neg_weight), int(n_neg),
              _p(sums), _p(gscale), float(pos_scale), float(neg_scale), _p(g), _stream())
    return g


def head_bwd(g, H, W, dW, db, M=None):
    """g [M,nout] fp32, H [M,D] bf16, W [nout,D] fp32 -> dZ [M,D] bf16; dW / db accumulated in place."""
    _chk(g, "g", _f32, 2)
    _chk(H, "hidden", _bf16, 2)
    _chk(W, "weight", _f32, 2)
    M = g.shape[0] if M is None else M
    dZ = torch.empty((H.shape[0], H.shape[1]), dtype=_bf16, device=H.device)
    _lib.call("pt_head_bwd", _p(g), g.shape[1], _p(H), H.shape[1], H.shape[1], _p(W), M, _p(dZ), dZ.shape[1], _p(dW),
              _p(db), _stream())
    return dZ


def transpose_pad(x, rows=None, pad_to=64):
    """bf16 [R, C] -> [C, Rp] with Rp = R rounded up to ``pad_to``, zero padded."""
    _chk(x, "x", _bf16, 2)
    R = x.shape[0] if rows is None else rows
    Rp = (R + pad_to - 1) // pad_to * pad_to
    out = torch.empty((x.shape[1], Rp), dtype=_bf16, device=x.device)
    _lib.call("pt_transpose_pad_bf16", _p(x), x.shape[1], R, x.shape[1], _p(out), Rp, _stream())
    return out


def unpermute_dw1(dw_binmajor, C, bins, grad, accumulate):
    _chk(dw_binmajor, "dW1", _f32, 2)
    _lib.call("pt_unpermute_dw1", _p(dw_binmajor), dw_binmajor.shape[0], C, bins, _p(grad), int(accumulate), _stream())
    return grad


def colsum_bf16(dZ, db, M=None):
    _chk(dZ, "dZ", _bf16, 2)
    _lib.call("pt_colsum_bf16", _p(dZ), dZ.shape[1], dZ.shape[0] if M is None else M, dZ.shape[1], _p(db), _stream())
    return db


def nhwc_to_nchw_f32(x, out=None, accumulate=False):
    _chk(x, "x", _f32, 4)
    B, H, W, C = x.shape
    if out is None:
        out = torch.empty((B, C, H, W), dtype=_f32, device=x.device)
    _lib.call("pt_nhwc_to_nchw_f32", _p(x), _p(out), B, C, H, W, int(accumulate), _stream())
    return out


def roi_align_backward(dA, rois, feat_shape_nhwc, spatial_scale, sampling_ratio=0, aligned=True, dfeat=None, K=None,
                       rotated=False, clockwise=True, roi_level=None, level=0):
    """dA bf16 [K, 49*C] bin-major -> dfeat NHWC fp32 (accumulated into ``dfeat`` when given, else zero-initialised).
    ``rotated``: rois (K,6) [b,cx,cy,w,h,theta], RoIAlignRotated semantics."""
    _chk(dA, "dA", _bf16, 2)
    _chk(rois, "rois", _f32, 2, 6 if rotated else 5)
    B, H, W, C = feat_shape_nhwc
    if dfeat is None:
        dfeat = torch.zeros((B, H, W, C), dtype=_f32, device=dA.device)
    K = rois.shape[0] if K is None else K
    if rotated:
        _lib.call("pt_roi_align_rotated_backward_ex", _p(dA), dA.shape[1], _p(rois), K, B, C, H, W, float(spatial_scale),
                  int(sampling_ratio), int(aligned), int(clockwise), _p(dfeat), _p(roi_level), int(level), _stream())
        return dfeat
    _lib.call("pt_roi_align_backward_ex", _p(dA), dA.shape[1], _p(rois), K, B, C, H, W, float(spatial_scale),
              int(sampling_ratio), int(aligned), _p(dfeat), _p(roi_level), int(level), _stream())
    return dfeat


# ------------------------------------------------------------------------------ coarse pseudo boxes (section 8f rank 1)
def decode_ltrb(points, ltrb):
    """distance2bbox + bbox_xyxy_to_cxcywh: -> (xyxy (P,4), cxcywh (P,4))."""
    _chk(points, "points", _f32, 2, 2)
    _chk(ltrb, "bbox_preds", _f32, 2, 4)
    P = points.shape[0]
    xyxy = torch.empty((P, 4), dtype=_f32, device=points.device)
    cxcywh = torch.empty((P, 4), dtype=_f32, device=points.device)
    _lib.call("pt_decode_ltrb", _p(points), _p(ltrb), P, _p(xyxy), _p(cxcywh), _stream())
    return xyxy, cxcywh


def pseudo_aggregate(gt_inds, labels, cls_scores, xyxy, gt_points, gt_bboxes, filter_score):
    _chk(gt_inds, "gt_inds", _i64, 1)
    _chk(labels, "labels", _i64, 1)
    _chk(cls_scores, "cls_scores", _f32, 2)
    _chk(gt_points, "gt_points", _f32, 2, 2)
    _chk(gt_bboxes, "gt_bboxes", _f32, 2, 4)
    G, P, dev = gt_points.shape[0], gt_inds.shape[0], gt_inds.device
    ws = torch.empty((G * 6,), dtype=_f32, device=dev)
    boxes = torch.empty((G, 4), dtype=_f32, device=dev)
    pts = torch.empty((G, 2), dtype=_f32, device=dev)
    scores = torch.empty((G,), dtype=_f32, device=dev)
    nums = torch.empty((G,), dtype=_i64, device=dev)
    valid = torch.empty((G,), dtype=_u8, device=dev)
    iou = torch.empty((2,), dtype=_f32, device=dev)
    _lib.call("pt_pseudo_aggregate", _p(gt_inds), _p(labels), _p(cls_scores), cls_scores.shape[1], _p(xyxy), P,
              _p(gt_points), _p(gt_bboxes), G, float(filter_score), _p(ws), _p(boxes), _p(pts), _p(scores), _p(nums),
              _p(valid), _p(iou), _stream())
    return boxes, pts, scores, nums, valid, iou


def ltrb_targets(points, boxes, gt_inds, assigned_labels, num_classes, want_centerness=False):
    """-> (targets (P,4), labels (P,) int64 with num_classes = background, centerness (P,) | None)."""
    _chk(points, "points", _f32, 2, 2)
    _chk(boxes, "pseudo_bboxes", _f32, 2, 4)
    _chk(gt_inds, "gt_inds", _i64, 1)
    _chk(assigned_labels, "labels", _i64, 1)
    P, dev = points.shape[0], points.device
    t = torch.empty((P, 4), dtype=_f32, device=dev)
    lab = torch.empty((P,), dtype=_i64, device=dev)
    cen = torch.empty((P,), dtype=_f32, device=dev) if want_centerness else None
    _lib.call("pt_ltrb_targets", _p(points), _p(boxes), _p(gt_inds), _p(assigned_labels), P, int(num_classes), _p(t),
              _p(lab), _p(cen), _stream())
    return t, lab, cen
