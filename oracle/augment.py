"""TEST INFRASTRUCTURE ONLY -- CPU restatement of ``strong_augmentation``:

  * HBB: HBB_TOD/mmdet/models/detectors/syn_images_generator_v2.py:24-132
        (random flip -> random rescale 0.8..1.2 with centre crop / zero pad -> round -> box re-normalisation)
  * OBB: OBB_TOD/mmrotate/models/detectors/syn_images_generator_v2.py:223-357
        (boxes as polygons; random flip -> random rotation 1..19 deg (torchvision ``TF.rotate``, nearest, fill 0) with
        the in-image filter -> the same rescale -> ``poly2obb``)

The reference draws its random numbers from CPU generators in a fixed order per image (``random.choice`` for the flip,
``np.random.randint(1, 20)`` for the OBB angle, ``np.random.uniform(0.8, 1.2)`` rounded to one decimal for the scale);
``draw_choices`` replays exactly those calls, so with the same ``random`` / ``np.random`` seeds it returns the
reference's own draws, and every function below takes the draws as an explicit argument (the harness injects them).

Pinned bit-exact against the reference's own functions by ``oracle/check_oracle_vs_ref.py --augment`` (HBB and OBB
under the import shim); golden vectors in tests/golden/augment.pt.  The image resampling itself is
``F.interpolate(bilinear, align_corners=False)`` and torchvision's ``rotate`` on CPU -- the same library calls the
reference makes; their arithmetic (ATen's generic CPU upsample kernel; affine_grid + grid_sample nearest) is restated
explicitly in ``bilinear_resize_exact`` / ``rotate_nearest_exact`` so that the device kernels have a formula to match.
"""
import math
import random

import numpy as np
import torch
import torch.nn.functional as F

FLIPS = ("horizontal", "vertical", "diagonal", "None")


def draw_choices(batch, rotated=False):
    """The reference's RNG calls, in its order: per image flip, [angle], scale factor."""
    out = []
    for _ in range(batch):
        flip = random.choice(list(FLIPS))
        angle = int(np.random.randint(1, 20)) if rotated else 0
        sf = float(np.around(np.random.uniform(0.8, 1.2), 1))
        out.append((flip, angle, sf))
    return out


def _flip(img, flip, W, H, xs, ys):
    """xs / ys: lists of tensors whose columns are x / y coordinates to mirror (in place)."""
    if flip == "horizontal":
        img = torch.flip(img, dims=[2])
    elif flip == "vertical":
        img = torch.flip(img, dims=[1])
    elif flip == "diagonal":
        img = torch.flip(img, dims=[1, 2])
    if flip in ("horizontal", "diagonal"):
        for t in xs:
            t.copy_(W - t)
    if flip in ("vertical", "diagonal"):
        for t in ys:
            t.copy_(H - t)
    return img


def _rescale(img, sf, H, W, gt_points, gt_labels, pts, labels, boxes):
    """:64-113 (HBB) == :300-349 (OBB): scale coordinates, crop (sf >= 1) or pad (sf < 1), resample, round."""
    scale_H, scale_W = int(H * sf), int(W * sf)
    if sf < 1.0:
        blank_h, blank_w = int((H - scale_H) / 2), int((W - scale_W) / 2)
    else:
        blank_h, blank_w = int((scale_H - H) / 2), int((scale_W - W) / 2)
    boxes, pts, gt_points = boxes * sf, pts * sf, gt_points * sf
    if sf >= 1.0:
        keep = ((gt_points[:, 0] >= blank_w) & (gt_points[:, 0] < (W + blank_w)) & (gt_points[:, 1] >= blank_h)
                & (gt_points[:, 1] < (H + blank_h))).nonzero().reshape(-1)
        gt_points, gt_labels = gt_points[keep, :], gt_labels[keep]
        gt_points[:, 0] -= blank_w
        gt_points[:, 1] -= blank_h
        keep = ((pts[:, 0] >= blank_w) & (pts[:, 0] < (W + blank_w)) & (pts[:, 1] >= blank_h)
                & (pts[:, 1] < (H + blank_h))).nonzero().reshape(-1)
        boxes, pts, labels = boxes[keep, :], pts[keep, :], labels[keep]
        pts[:, 0] -= blank_w
        pts[:, 1] -= blank_h
        boxes[:, 0::2] -= blank_w
        boxes[:, 1::2] -= blank_h
    else:
        gt_points[:, 0] += blank_w
        gt_points[:, 1] += blank_h
        pts[:, 0] += blank_w
        pts[:, 1] += blank_h
        boxes[:, 0::2] += blank_w
        boxes[:, 1::2] += blank_h
    rescaled = F.interpolate(img.unsqueeze(0), size=(scale_H, scale_W), mode="bilinear", align_corners=False).squeeze(0)
    out = torch.zeros_like(img)
    if sf < 1.0:
        sy, sx = (H - scale_H) // 2, (W - scale_W) // 2
        out[:, sy:sy + scale_H, sx:sx + scale_W] = rescaled
    else:
        sy, sx = (scale_H - H) // 2, (scale_W - W) // 2
        out = rescaled[:, sy:sy + H, sx:sx + W]
    return torch.round(out), gt_points, gt_labels, pts, labels, boxes


def strong_augmentation_hbb(img, gt_points, gt_labels, pseudo_points, pseudo_labels, pseudo_bboxes, choices):
    """HBB_TOD/mmdet/models/detectors/syn_images_generator_v2.py:24-132 with the random draws injected."""
    B, C, H, W = img.shape
    outs = [[] for _ in range(6)]
    for i in range(B):
        flip, _, sf = choices[i]
        im = img[i].clone()
        gp, gl = gt_points[i].clone(), gt_labels[i].clone()
        bx, pl, pp = pseudo_bboxes[i].clone(), pseudo_labels[i].clone(), pseudo_points[i].clone()
        im = _flip(im, flip, W, H, [bx[:, 0::2], pp[:, 0], gp[:, 0]], [bx[:, 1::2], pp[:, 1], gp[:, 1]])
        im, gp, gl, pp, pl, bx = _rescale(im, sf, H, W, gp, gl, pp, pl, bx)
        if len(bx) != 0:       # :115-121
            w = (bx[:, 0] - bx[:, 2]).abs().reshape(-1, 1)
            h = (bx[:, 1] - bx[:, 3]).abs().reshape(-1, 1)
            x = torch.min(bx[:, [0, 2]], dim=1)[0].reshape(-1, 1)
            y = torch.min(bx[:, [1, 3]], dim=1)[0].reshape(-1, 1)
            c = torch.cat([x + w / 2, y + h / 2, w, h], dim=1)
            bx = torch.cat([c[:, :2] - 0.5 * c[:, 2:], c[:, :2] + 0.5 * c[:, 2:]], dim=-1)   # bbox_cxcywh_to_xyxy
        for lst, v in zip(outs, (im, gp, gl, pp, pl, bx)):
            lst.append(v)
    return (torch.stack(outs[0], 0), *outs)


# ------------------------------------------------------------------------------ OBB
def obb2poly_le90(r):
    """OBB_TOD/mmrotate/core/bbox/transforms.py:474-499."""
    N = r.shape[0]
    if N == 0:
        return r.new_zeros((0, 8))
    x, y, w, h, a = r[:, 0], r[:, 1], r[:, 2], r[:, 3], r[:, 4]
    tlx, tly, brx, bry = -w * 0.5, -h * 0.5, w * 0.5, h * 0.5
    rects = torch.stack([tlx, brx, brx, tlx, tly, tly, bry, bry], dim=0).reshape(2, 4, N).permute(2, 0, 1)
    sin, cos = torch.sin(a), torch.cos(a)
    M = torch.stack([cos, -sin, sin, cos], dim=0).reshape(2, 2, N).permute(2, 0, 1)
    polys = M.matmul(rects).permute(2, 1, 0).reshape(-1, N).transpose(1, 0)
    polys[:, ::2] += x.unsqueeze(1)
    polys[:, 1::2] += y.unsqueeze(1)
    return polys.contiguous()


def poly2obb_le90(polys):
    """transforms.py:301-331."""
    polys = polys.reshape(-1, 8)
    p1, p2, p3, p4 = polys.chunk(4, 1)
    e1 = torch.sqrt(torch.pow(p1[..., 0] - p2[..., 0], 2) + torch.pow(p1[..., 1] - p2[..., 1], 2))
    e2 = torch.sqrt(torch.pow(p2[..., 0] - p3[..., 0], 2) + torch.pow(p2[..., 1] - p3[..., 1], 2))
    a1 = torch.atan2(p2[..., 1] - p1[..., 1], p2[..., 0] - p1[..., 0])
    a2 = torch.atan2(p4[..., 1] - p1[..., 1], p4[..., 0] - p1[..., 0])
    ang = polys.new_zeros(polys.shape[0])
    ang[e1 > e2] = a1[e1 > e2]
    ang[e1 <= e2] = a2[e1 <= e2]
    ang = (ang + np.pi / 2) % np.pi - np.pi / 2          # norm_angle 'le90' (:864-865)
    xc, yc = (p1[..., 0] + p3[..., 0]) / 2.0, (p1[..., 1] + p3[..., 1]) / 2.0
    e = torch.stack([e1, e2], dim=1)
    return torch.stack([xc, yc, e.max(1)[0], e.min(1)[0], ang], 1)


def strong_augmentation_obb(img, gt_points, gt_labels, pseudo_points, pseudo_labels, pseudo_bboxes, choices,
                            angle_version="le90"):
    """OBB_TOD/mmrotate/models/detectors/syn_images_generator_v2.py:223-357 with the random draws injected."""
    import torchvision.transforms.functional as TF
    assert angle_version == "le90"
    B, C, H, W = img.shape
    outs = [[] for _ in range(6)]
    for i in range(B):
        flip, angle, sf = choices[i]
        im = img[i].clone()
        gp, gl = gt_points[i].clone(), gt_labels[i].clone()
        bx = obb2poly_le90(pseudo_bboxes[i].clone())
        pl, pp = pseudo_labels[i].clone(), pseudo_points[i].clone()
        im = _flip(im, flip, W, H, [bx[:, 0::2], pp[:, 0], gp[:, 0]], [bx[:, 1::2], pp[:, 1], gp[:, 1]])
        # random rotate (:265-298)
        cx, cy = W / 2, H / 2
        im = TF.rotate(im, angle, fill=0)
        rad = np.deg2rad(-angle)
        ca, sa = np.cos(rad), np.sin(rad)
        tb, tp, tg = bx.clone(), pp.clone(), gp.clone()
        bx[:, 0::2] = ca * (tb[:, 0::2] - cx) - sa * (tb[:, 1::2] - cy) + cx
        bx[:, 1::2] = sa * (tb[:, 0::2] - cx) + ca * (tb[:, 1::2] - cy) + cy
        pp[:, 0] = ca * (tp[:, 0] - cx) - sa * (tp[:, 1] - cy) + cx
        pp[:, 1] = sa * (tp[:, 0] - cx) + ca * (tp[:, 1] - cy) + cy
        gp[:, 0] = ca * (tg[:, 0] - cx) - sa * (tg[:, 1] - cy) + cx
        gp[:, 1] = sa * (tg[:, 0] - cx) + ca * (tg[:, 1] - cy) + cy
        keep = (((0 <= gp[:, 0]) & (gp[:, 0] < W)) & ((0 <= gp[:, 1]) & (gp[:, 1] < H))).nonzero().reshape(-1)
        gp, gl = gp[keep, :], gl[keep]
        keep = (((0 <= pp[:, 0]) & (pp[:, 0] < W)) & ((0 <= pp[:, 1]) & (pp[:, 1] < H))).nonzero().reshape(-1)
        pp, pl, bx = pp[keep, :], pl[keep], bx[keep, :]
        im, gp, gl, pp, pl, bx = _rescale(im, sf, H, W, gp, gl, pp, pl, bx)
        bx = poly2obb_le90(bx) if len(bx) != 0 else torch.empty(0, 5, dtype=gt_points[0].dtype)
        for lst, v in zip(outs, (im, gp, gl, pp, pl, bx)):
            lst.append(v)
    return (torch.stack(outs[0], 0), *outs)


# ------------------------------------------------------------------------------ explicit arithmetic of the resamplers
def _f32(x):
    return np.asarray(x, dtype=np.float32)


def _fma(a, b, c):
    """fp32 fused multiply-add emulated in float64 (the product of two fp32 is exact in fp64)."""
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(np.float32)


def bilinear_axis(in_size, out_size):
    """ATen ``compute_source_index_and_lambda`` (aten/src/ATen/native/UpSample.h), align_corners=False, fp32, as the
    CPU build evaluates it: src = fma(scale, dst + 0.5, -0.5) clamped at 0; i0 = min(floor(src), in-1);
    l1 = clamp(src - i0, 0, 1); l0 = 1 - l1; i1 = i0 + (i0 < in-1).  out_size == in_size copies."""
    if in_size == out_size:
        i = np.arange(out_size)
        return i, i, np.ones(out_size, np.float32), np.zeros(out_size, np.float32)
    scale = np.float32(in_size) / np.float32(out_size)
    d = np.arange(out_size, dtype=np.float32)
    src = _fma(scale, d + np.float32(0.5), -np.float32(0.5))
    src = np.where(src < 0, np.float32(0), src).astype(np.float32)
    i0 = np.minimum(np.floor(src).astype(np.int64), in_size - 1)
    l1 = np.clip(src - i0.astype(np.float32), 0, 1).astype(np.float32)
    i1 = i0 + (i0 < in_size - 1)
    return i0, i1, (np.float32(1) - l1).astype(np.float32), l1


def bilinear_resize_exact(img, out_h, out_w):
    """ATen's generic CPU bilinear kernel (UpSampleKernel.cpp ``Interpolate<2>``) in explicit fp32:
    row(y) = fma(v[y][x0], wx0, v[y][x1] * wx1);  out = fma(row(y0), wy0, row(y1) * wy1).  Bit-identical to
    ``F.interpolate`` on the CPU builds checked (tests/test_oracle.py)."""
    a = img.numpy()
    y0, y1, wy0, wy1 = bilinear_axis(a.shape[1], out_h)
    x0, x1, wx0, wx1 = bilinear_axis(a.shape[2], out_w)

    def row(yy):
        r = a[:, yy, :]
        return _fma(r[:, :, x0], wx0[None, None, :], (r[:, :, x1] * wx1[None, None, :]).astype(np.float32))
    r0, r1 = row(y0), row(y1)
    return torch.from_numpy(_fma(r0, wy0[None, :, None], (r1 * wy1[None, :, None]).astype(np.float32)))


def rotate_matrix(angle):
    """torchvision ``_get_inverse_affine_matrix(center=[0,0], angle=-angle, translate=[0,0], scale=1, shear=[0,0])``
    (torchvision/transforms/functional.py), python float64 arithmetic -> 6 floats."""
    rot = math.radians(-angle)
    sx = sy = 0.0
    a = math.cos(rot - sy) / math.cos(sy)
    b = -math.cos(rot - sy) * math.tan(sx) / math.cos(sy) - math.sin(rot)
    c = math.sin(rot - sy) / math.cos(sy)
    d = -math.sin(rot - sy) * math.tan(sx) / math.cos(sy) + math.cos(rot)
    m = [d, -b, 0.0, -c, a, 0.0]
    m = [x / 1.0 for x in m]
    m[2] += m[0] * 0.0 + m[1] * 0.0
    m[5] += m[3] * 0.0 + m[4] * 0.0
    return m


def rotate_source_index(angle, H, W):
    """Nearest source pixel (ix, iy) of every output pixel of ``TF.rotate(img, angle)`` (tensor path:
    ``_gen_affine_grid`` + ``grid_sample(nearest, zeros, align_corners=False)``), explicit fp32:
    gx = x * r00 + y * r10 with r = theta^T / (0.5 w, 0.5 h) and half-integer base coordinates;
    ix = nearbyint(((gx + 1) * (W / 2)) - 0.5).  Returns int arrays (H, W); out-of-range indices mean fill."""
    m = np.asarray(rotate_matrix(angle), np.float32).reshape(2, 3)
    r = (m.T / np.asarray([0.5 * W, 0.5 * H], np.float32)).astype(np.float32)      # (3, 2)
    xs = (np.arange(W, dtype=np.float32) + np.float32(-W * 0.5 + 0.5))[None, :]
    ys = (np.arange(H, dtype=np.float32) + np.float32(-H * 0.5 + 0.5))[:, None]
    gx = _fma(ys, r[1, 0], (xs * r[0, 0]).astype(np.float32))
    gy = _fma(ys, r[1, 1], (xs * r[0, 1]).astype(np.float32))
    ix = (((gx + np.float32(1)) * np.float32(W / 2)).astype(np.float32) - np.float32(0.5)).astype(np.float32)
    iy = (((gy + np.float32(1)) * np.float32(H / 2)).astype(np.float32) - np.float32(0.5)).astype(np.float32)
    return np.rint(ix).astype(np.int64), np.rint(iy).astype(np.int64)


def rotate_nearest_exact(img, angle):
    C, H, W = img.shape
    ix, iy = rotate_source_index(angle, H, W)
    ok = (ix >= 0) & (ix < W) & (iy >= 0) & (iy < H)
    a = img.numpy()
    out = np.zeros_like(a)
    out[:, ok] = a[:, iy[ok], ix[ok]]
    return torch.from_numpy(out)
