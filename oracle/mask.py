"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the phase-1 random region masking.
Paths relative to /root/reference/HBB_TOD/mmdet/models/detectors/.

  * ``black_paper_from_candidates``  the deterministic tail of ``generate_black_paper``
    (syn_images_generator_v2.py:664-690): rotated NMS at 0.05, score < 1 filter, inside-image filter,
    ``obb2poly_le90`` (data_augument_bank.py:516-541), int32-truncated corners, ``cv2.fillPoly``, pixels = 255.
  * ``fill_poly_replay``  OpenCV's integer polygon rasteriser restated (drawing.cpp: CollectPolyEdges -> Line via
    LineIterator (8-connected, left-to-right) + FillEdgeCollection in 16.16 fixed point, x1 = ceil, x2 = floor).
    Pinned against ``cv2.fillPoly`` of the installed OpenCV (4.13) on thousands of random rotated rectangles
    (tests/test_oracle.py); ``cv2`` itself is the reference implementation for this step.
  * ``sample_candidates``  the random part (:596-663), same draw order from the torch / numpy global generators.

Pinned end to end by ``python -m oracle.check_oracle_vs_ref --mask`` against the reference's own
``generate_black_paper`` under the import shim (mmcv's nms_rotated bound to oracle/rotated.py: PARITY UNPINNED for
that kernel, see oracle/rotated.py)."""
import math

import numpy as np
import torch

from . import rotated

XY_SHIFT = 16
XY_ONE = 1 << XY_SHIFT


def obb2xyxy(rb):
    """The generator's own helper, syn_images_generator_v2.py:382-396 (|cos|, |sin|)."""
    w, h, a = rb[:, 2], rb[:, 3], rb[:, 4]
    cosa, sina = torch.cos(a).abs(), torch.sin(a).abs()
    bw, bh = cosa * w + sina * h, sina * w + cosa * h
    return torch.stack((rb[:, 0] - bw / 2, rb[:, 1] - bh / 2, rb[:, 0] + bw / 2, rb[:, 1] + bh / 2), -1)


def obb2poly_le90(rb):
    """data_augument_bank.py:516-541; the 2x2 @ 2x4 bmm is evaluated as separate multiplies and one add."""
    if rb.shape[0] == 0:
        return rb.new_zeros((0, 8))
    x, y, w, h, a = rb[:, 0], rb[:, 1], rb[:, 2], rb[:, 3], rb[:, 4]
    tlx, tly, brx, bry = -w * 0.5, -h * 0.5, w * 0.5, h * 0.5
    sin, cos = torch.sin(a), torch.cos(a)
    xs = [tlx, brx, brx, tlx]
    ys = [tly, tly, bry, bry]
    out = []
    for px, py in zip(xs, ys):
        out.append(cos * px + (-sin) * py + x)
        out.append(sin * px + cos * py + y)
    return torch.stack(out, 1)


def _line8(mask, p0, p1):
    h, w = mask.shape
    x0, y0 = p0
    x1, y1 = p1
    dx, dy = x1 - x0, y1 - y0
    if dx < 0:                      # LineIterator(leftToRight=true)
        x0, y0, dx, dy = x1, y1, -dx, -dy
    sy = 1 if dy >= 0 else -1
    ady = abs(dy)
    x, y = x0, y0
    if ady > dx:                    # steep
        err, plus, minus = ady - 2 * dx, 2 * ady, -2 * dx
        for _ in range(ady + 1):
            if 0 <= x < w and 0 <= y < h:
                mask[y, x] = 1
            neg = err < 0
            err += minus + (plus if neg else 0)
            y += sy
            x += 1 if neg else 0
    else:
        err, plus, minus = dx - 2 * ady, 2 * dx, -2 * ady
        for _ in range(dx + 1):
            if 0 <= x < w and 0 <= y < h:
                mask[y, x] = 1
            neg = err < 0
            err += minus + (plus if neg else 0)
            x += 1
            y += sy if neg else 0


def fill_poly_replay(mask, pts):
    """cv2.fillPoly(mask, [pts], 1) for one polygon with integer vertices lying inside the image."""
    h, w = mask.shape
    pts = [(int(p[0]), int(p[1])) for p in pts]
    edges = []
    p0 = pts[-1]
    for p1 in pts:
        _line8(mask, p0, p1)
        if p0[1] != p1[1]:
            num, den = (p1[0] - p0[0]) << XY_SHIFT, p1[1] - p0[1]
            dx = (abs(num) // abs(den)) * (1 if (num >= 0) == (den > 0) else -1)      # C++ int64 division
            lo, hi = (p0, p1) if p0[1] < p1[1] else (p1, p0)
            edges.append((lo[1], hi[1], lo[0] << XY_SHIFT, dx))
        p0 = p1
    if len(edges) < 2:
        return
    y_min, y_max = min(e[0] for e in edges), min(max(e[1] for e in edges), h)
    for y in range(max(y_min, 0), y_max):
        xs = sorted(e[2] + (y - e[0]) * e[3] for e in edges if e[0] <= y < e[1])
        for k in range(0, len(xs) - 1, 2):
            x1, x2 = (xs[k] + XY_ONE - 1) >> XY_SHIFT, xs[k + 1] >> XY_SHIFT
            if x1 < w and x2 >= 0 and x2 >= x1:
                mask[y, max(x1, 0):min(x2, w - 1) + 1] = 1


def black_paper_from_candidates(img_syn, bb_all, imgsize, use_cv2=True):
    """syn_images_generator_v2.py:664-690.  bb_all (N,7) = cat(bb_occupied, candidates): (x, y, w, h, a, score, cls).
    Returns (img_syn, kept boxes (M,7), keep indices into bb_all (M,), int32 polygons (M,4,2), mask (H,W) uint8)."""
    _, keep = rotated.nms_rotated(bb_all[:, 0:5], bb_all[:, 5], 0.05)
    sel = keep[bb_all[keep, 5] < 1]
    bb = bb_all[sel]
    xyxy = obb2xyxy(bb)
    if bb.shape[0]:
        ok = torch.logical_and(xyxy.min(-1)[0] >= 0, xyxy.max(-1)[0] <= imgsize - 1)
    else:
        ok = torch.zeros((0,), dtype=torch.bool)
    bb, sel = bb[ok], sel[ok]
    polys = obb2poly_le90(bb[:, :5]).view(-1, 4, 2).numpy().astype(np.int32)
    H, W = img_syn.shape[-2:]
    mask = np.zeros((H, W), dtype=np.uint8)
    if use_cv2:
        import cv2
        for p in polys:
            cv2.fillPoly(mask, [p], 1)
    else:
        for p in polys:
            fill_poly_replay(mask, p)
    m = torch.from_numpy(mask).bool()
    img_syn[:, m] = 255
    return img_syn, bb, sel, polys, mask


def sample_candidates(bb_occupied, prior_size, dense_cls, imgsize):
    """syn_images_generator_v2.py:596-663 without the (unused downstream) palette: the same torch / numpy global
    RNG draws in the same order.  Returns cat(bb_occupied', candidates) (N,7)."""
    cen = [50, imgsize - 50]
    scale_vary = torch.rand(bb_occupied.shape[0]) * 2.0 + 0.5
    occ = bb_occupied.clone()
    occ[:, 2] = prior_size[occ[:, 6].long(), 0] * 0.7
    occ[:, 3] = prior_size[occ[:, 6].long(), 0] * 0.7
    occ[:, 4] = 0
    bb, adjboost = [], 2
    for n, b in enumerate(occ):
        base = scale_vary[n]
        c = b[6].long()
        x, y = torch.rand(2) * (cen[1] - cen[0]) + cen[0]
        w = base * torch.exp((torch.randn(1) * 0.4).clamp(-1, 1) * prior_size[c, 2])
        r = (torch.randn(1) * 0.4).clamp(-1, 1) * prior_size[c, 3]
        h = w * torch.exp(r)
        w = w * prior_size[c, 0]
        h = h * prior_size[c, 1]
        a = torch.rand(1) * torch.pi - torch.pi / 2
        x = x.clip(0.71 * w, imgsize - 1 - 0.71 * w)
        y = y.clip(0.71 * h, imgsize - 1 - 0.71 * h)
        bb.append([x, y, w, h, a, (w * h) / imgsize / imgsize + 0.1, b[6]])
        if np.random.random() < 0.2 and adjboost > 0:
            adjboost -= 1
            if c in dense_cls:
                itv, dev, reps = torch.rand(1) * 4 + 2, torch.rand(1) * 8 - 4, 6
            else:
                itv, dev, reps = torch.rand(1) * 40 + 10, torch.rand(1) * 0, 4
            ofx = (h + itv) * torch.sin(-a) + dev * torch.cos(a)
            ofy = (h + itv) * torch.cos(a) + dev * torch.sin(a)
            for k in range(1, reps):
                bb.append([x + k * ofx, y + k * ofy, w, h, a, (w * h) / imgsize / imgsize + 0.1 - 0.001 * k, b[6]])
    cand = torch.tensor(bb) if bb else torch.zeros((0, 7))
    return torch.cat((occ, cand), 0)
