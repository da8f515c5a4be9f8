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
