"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the three un-vendored mmcv
rotated kernels the OBB path and phase-1 masking reach:

  * ``mmcv.ops.RoIAlignRotated``   (call site OBB_TOD/mmrotate/models/roi_heads/
    roi_extractors/rotate_single_level_roi_extractor.py:90-167)
  * ``mmcv.ops.box_iou_rotated``   (OBB_TOD/mmrotate/core/bbox/iou_calculators/
    rotate_iou2d_calculator.py:53-89)
  * ``mmcv.ops.nms_rotated``       (HBB_TOD/mmdet/models/detectors/
    syn_images_generator_v2.py:667)

mmcv-full 1.x is a dependency that is absent from /root/reference (accepted
range 1.3.2..1.7.2 HBB / 1.5.3..1.8.0 OBB), so these follow the published
Detectron2/mmcv algorithm.  PARITY UNPINNED by any
reference test; cross-checks used instead: theta=0 == torchvision roi_align /
axis-aligned bbox_overlaps, cv2.rotatedRectangleIntersection (Appendix A.7).

The per-pair polygon clipping lives in C (oracle/c/rotated.c -> oracle/_build/
liboracle.so) because pure-Python loops are too slow for M x N matrices.
"""
import ctypes
import math

import numpy as np
import torch

from . import cbuild


def _lib():
    return cbuild.load()


def _f32(t):
    return np.ascontiguousarray(t.detach().cpu().numpy().astype(np.float32))


def box_iou_rotated(bboxes1, bboxes2, mode="iou", aligned=False, clockwise=True):
    """(M,5)x(N,5) [cx,cy,w,h,theta rad] -> (M,N) or aligned (M,).  fp32 in/out."""
    assert mode in ("iou", "iof")
    a, b = _f32(bboxes1), _f32(bboxes2)
    m, n = a.shape[0], b.shape[0]
    if not clockwise:
        a, b = a.copy(), b.copy()
        a[:, 4] *= -1
        b[:, 4] *= -1
    if aligned:
        assert m == n
        out = np.zeros((m,), np.float32)
    else:
        out = np.zeros((m, n), np.float32)
    if m * n:
        _lib().oracle_box_iou_rotated(
            a.ctypes.data_as(ctypes.c_void_p), b.ctypes.data_as(ctypes.c_void_p),
            out.ctypes.data_as(ctypes.c_void_p), m, n, int(aligned), 0 if mode == "iou" else 1)
    return torch.from_numpy(out).to(bboxes1.device)


def nms_rotated(dets, scores, iou_threshold, labels=None, clockwise=True):
    """Greedy rotated NMS; returns (cat(dets, scores)[keep], keep) like mmcv."""
    if dets.shape[0] == 0:
        return dets, None
    d = _f32(dets[:, :5])
    if not clockwise:
        d = d.copy()
        d[:, 4] *= -1
    # mmcv's wrapper calls scores.sort(0, descending=True) WITHOUT stable=True: the order of equal scores is then
    # implementation-defined (ATen dispatches to an ISA-specific vectorised quicksort on recent builds), so it
    # cannot be part of a parity contract.  Oracle and product both fix it to the stable order.
    order = torch.sort(scores.detach().cpu(), stable=True, dim=0, descending=True)[1].numpy().astype(np.int64)
    n = d.shape[0]
    keep_mask = np.zeros((n,), np.uint8)
    _lib().oracle_nms_rotated(d.ctypes.data_as(ctypes.c_void_p),
                              order.ctypes.data_as(ctypes.c_void_p),
                              keep_mask.ctypes.data_as(ctypes.c_void_p), n, float(iou_threshold))
    # mmcv returns kept indices in descending-score order
    keep = torch.from_numpy(order[keep_mask[order] == 1]).to(dets.device)
    out = torch.cat((dets[keep], scores[keep].reshape(-1, 1)), dim=1)
    return out, keep


def roi_align_rotated(x, rois, out_size, spatial_scale, sampling_ratio, aligned, clockwise):
    """x (B,C,H,W) fp32, rois (K,6) [b,cx,cy,w,h,theta] -> (K,C,P,P).  Appendix A.2."""
    x_np = _f32(x)
    r = _f32(rois)
    B, C, H, W = x_np.shape
    K = r.shape[0]
    out = np.zeros((K, C, out_size, out_size), np.float32)
    if K:
        _lib().oracle_roi_align_rotated(
            x_np.ctypes.data_as(ctypes.c_void_p), r.ctypes.data_as(ctypes.c_void_p),
            out.ctypes.data_as(ctypes.c_void_p), B, C, H, W, K, out_size,
            ctypes.c_float(spatial_scale), int(sampling_ratio), int(aligned), int(clockwise))
    return torch.from_numpy(out).to(x.device)


def roi_align_rotated_torch(x, rois, out_size, spatial_scale, sampling_ratio, aligned, clockwise):
    """Differentiable (w.r.t. ``x``) PyTorch twin of ``roi_align_rotated`` for a fixed sampling grid
    (``sampling_ratio > 0``), same formulae in the same fp32 order (Appendix A.2); used only where a test needs
    torch autograd through the extraction.  Checked against the C restatement in tests/test_oracle.py."""
    assert sampling_ratio > 0
    K, P, g = rois.shape[0], out_size, sampling_ratio
    B, C, H, W = x.shape
    off = 0.5 if aligned else 0.0
    b = rois[:, 0].long()
    cx, cy = rois[:, 1] * spatial_scale - off, rois[:, 2] * spatial_scale - off
    rw, rh = rois[:, 3] * spatial_scale, rois[:, 4] * spatial_scale
    th = -rois[:, 5] if clockwise else rois[:, 5]
    if not aligned:
        rw, rh = rw.clamp(min=1.0), rh.clamp(min=1.0)
    bh, bw = rh / P, rw / P
    sh, sw = -rh / 2.0, -rw / 2.0
    ct, st = torch.cos(th), torch.sin(th)
    pi = torch.arange(P, dtype=torch.float32)
    si = torch.arange(g, dtype=torch.float32) + 0.5
    # yy[k, ph, iy], xx[k, pw, ix]
    yy = (sh[:, None, None] + pi[None, :, None] * bh[:, None, None]) + si[None, None, :] * bh[:, None, None] / g
    xx = (sw[:, None, None] + pi[None, :, None] * bw[:, None, None]) + si[None, None, :] * bw[:, None, None] / g
    yy = yy[:, :, None, :, None]                                     # (K, P, 1, g, 1)
    xx = xx[:, None, :, None, :]                                     # (K, 1, P, 1, g)
    c4, s4 = ct[:, None, None, None, None], st[:, None, None, None, None]
    y = (yy * c4 - xx * s4) + cy[:, None, None, None, None]
    xs = (yy * s4 + xx * c4) + cx[:, None, None, None, None]
    ok = ~((y < -1.0) | (y > H) | (xs < -1.0) | (xs > W))
    y, xs = y.clamp(min=0.0), xs.clamp(min=0.0)
    yl, xl = y.floor().long(), xs.floor().long()
    ycap, xcap = yl >= H - 1, xl >= W - 1
    yl, xl = torch.where(ycap, torch.full_like(yl, H - 1), yl), torch.where(xcap, torch.full_like(xl, W - 1), xl)
    yh, xh = torch.where(ycap, yl, yl + 1), torch.where(xcap, xl, xl + 1)
    y, xs = torch.where(ycap, yl.float(), y), torch.where(xcap, xl.float(), xs)
    ly, lx = y - yl.float(), xs - xl.float()
    hy, hx = 1.0 - ly, 1.0 - lx
    xn = x.permute(0, 2, 3, 1)                                        # (B, H, W, C)
    bb = b[:, None, None, None, None].expand_as(yl)
    okf = ok.float()
    val = ((hy * hx * okf)[..., None] * xn[bb, yl, xl] + (hy * lx * okf)[..., None] * xn[bb, yl, xh]
           + (ly * hx * okf)[..., None] * xn[bb, yh, xl] + (ly * lx * okf)[..., None] * xn[bb, yh, xh])
    out = val.sum(dim=(3, 4)) / float(max(g * g, 1))                  # (K, P, P, C)
    return out.permute(0, 3, 1, 2).contiguous()


def roi_align(x, rois, out_size, spatial_scale, sampling_ratio, aligned):
    """Horizontal RoIAlign (Appendix A.1), C restatement; pinned against
    torchvision.ops.roi_align in tests/test_oracle.py."""
    x_np = _f32(x)
    r = _f32(rois)
    B, C, H, W = x_np.shape
    K = r.shape[0]
    out = np.zeros((K, C, out_size, out_size), np.float32)
    if K:
        _lib().oracle_roi_align(
            x_np.ctypes.data_as(ctypes.c_void_p), r.ctypes.data_as(ctypes.c_void_p),
            out.ctypes.data_as(ctypes.c_void_p), B, C, H, W, K, out_size,
            ctypes.c_float(spatial_scale), int(sampling_ratio), int(aligned))
    return torch.from_numpy(out).to(x.device)
