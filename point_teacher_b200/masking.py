"""Phase-1 random region masking behind the reference's function surface.

``generate_black_paper(img, bb_occupied, img_syn, pattern, prior_size, dense_cls, imgsize)`` keeps the signature of
HBB_TOD/mmdet/models/detectors/syn_images_generator_v2.py:591-690 (imported by name at
fcos_p2b_teacher_student.py:13).  The candidate boxes are drawn exactly like the reference does -- from the torch and
numpy GLOBAL CPU generators, in the same order (:596-663), because no device generator can reproduce that stream --
and everything after that (rotated NMS, filters, polygon construction, cv2.fillPoly-exact rasterisation, pixel
write) runs on the GPU on the image where it already lives: no D2H / H2D of full images."""
import math

import numpy as np
import torch

from . import ops


def load_basic_shape(shape_list):
    """:578-586."""
    torch.pi = math.pi
    prior_size = torch.Tensor(shape_list).float()
    pattern = [[torch.zeros(s[:2]).float()] for s in shape_list]
    return pattern, prior_size


def sample_black_paper_candidates(bb_occupied, prior_size, dense_cls, imgsize):
    """The random half of generate_black_paper (:596-663) on the host generators; returns cat(bb_occupied', candidates)
    (N,7) on the CPU.  The colour palette the reference also collects there is never used and is not computed."""
    bb_occupied, prior_size = bb_occupied.detach().cpu().float(), prior_size.detach().cpu().float()
    lo, hi = 50, imgsize - 50
    scale_vary = torch.rand(bb_occupied.shape[0]) * 2.0 + 0.5
    occ = bb_occupied.clone()
    side = prior_size[occ[:, 6].long(), 0] * 0.7
    occ[:, 2], occ[:, 3], occ[:, 4] = side, side, 0
    rows, boost = [], 2
    for n in range(occ.shape[0]):
        cls = occ[n, 6]
        ci = cls.long()
        xy = torch.rand(2) * (hi - lo) + lo
        w = scale_vary[n] * torch.exp((torch.randn(1) * 0.4).clamp(-1, 1) * prior_size[ci, 2])
        h = w * torch.exp((torch.randn(1) * 0.4).clamp(-1, 1) * prior_size[ci, 3])
        w, h = w * prior_size[ci, 0], h * prior_size[ci, 1]
        a = torch.rand(1) * math.pi - math.pi / 2
        x = xy[0].clip(0.71 * w, imgsize - 1 - 0.71 * w)
        y = xy[1].clip(0.71 * h, imgsize - 1 - 0.71 * h)
        score = (w * h) / imgsize / imgsize + 0.1
        rows.append([x, y, w, h, a, score, cls])
        if np.random.random() < 0.2 and boost > 0:
            boost -= 1
            if ci in dense_cls:
                itv, dev, last = torch.rand(1) * 4 + 2, torch.rand(1) * 8 - 4, 5
            else:
                itv, dev, last = torch.rand(1) * 40 + 10, torch.rand(1) * 0, 3
            ofx = (h + itv) * torch.sin(-a) + dev * torch.cos(a)
            ofy = (h + itv) * torch.cos(a) + dev * torch.sin(a)
            for k in range(1, last + 1):
                rows.append([x + k * ofx, y + k * ofy, w, h, a, score - 0.001 * k, cls])
    cand = torch.tensor(rows) if rows else torch.zeros((0, 7))
    return torch.cat((occ, cand), 0)


def sample_black_paper_candidates_fast(bb_occupied, prior_size, dense_cls, imgsize, rng=None):
    """Vectorised draw of the same candidate DISTRIBUTION as ``sample_black_paper_candidates`` (:596-663) from a numpy
    ``Generator``: one bulk draw per quantity instead of ~15 scalar tensor ops per GT (the reference's loop costs tens
    of milliseconds per image at AI-TOD GT counts).  It does NOT reproduce the reference's random stream -- use it
    when throughput matters and the literal function when the stream does.  Returns (N,7) on the CPU."""
    rng = rng if rng is not None else np.random.default_rng()
    occ = bb_occupied.detach().cpu().float().numpy().copy()
    prior = prior_size.detach().cpu().float().numpy()
    G = occ.shape[0]
    ci = occ[:, 6].astype(np.int64)
    side = prior[ci, 0] * 0.7
    occ[:, 2], occ[:, 3], occ[:, 4] = side, side, 0
    if G == 0:
        return torch.from_numpy(occ)
    lo, hi = 50, imgsize - 50
    scale = rng.random(G, dtype=np.float32) * 2.0 + 0.5
    xy = rng.random((G, 2), dtype=np.float32) * (hi - lo) + lo
    w = scale * np.exp(np.clip(rng.standard_normal(G).astype(np.float32) * 0.4, -1, 1) * prior[ci, 2])
    h = w * np.exp(np.clip(rng.standard_normal(G).astype(np.float32) * 0.4, -1, 1) * prior[ci, 3])
    w, h = w * prior[ci, 0], h * prior[ci, 1]
    a = rng.random(G, dtype=np.float32) * np.float32(math.pi) - np.float32(math.pi / 2)
    x = np.clip(xy[:, 0], 0.71 * w, imgsize - 1 - 0.71 * w)
    y = np.clip(xy[:, 1], 0.71 * h, imgsize - 1 - 0.71 * h)
    score = (w * h) / imgsize / imgsize + 0.1
    rows = [np.stack([x, y, w, h, a, score, occ[:, 6]], 1)]
    dense = set(int(c) for c in dense_cls)
    for n in np.nonzero(rng.random(G) < 0.2)[0][:2]:                 # at most two neighbour runs (adjboost = 2)
        if int(ci[n]) in dense:
            itv, dev, last = rng.random() * 4 + 2, rng.random() * 8 - 4, 5
        else:
            itv, dev, last = rng.random() * 40 + 10, 0.0, 3
        ofx = (h[n] + itv) * math.sin(-a[n]) + dev * math.cos(a[n])
        ofy = (h[n] + itv) * math.cos(a[n]) + dev * math.sin(a[n])
        k = np.arange(1, last + 1, dtype=np.float32)
        rows.append(np.stack([x[n] + k * ofx, y[n] + k * ofy, np.full_like(k, w[n]), np.full_like(k, h[n]),
                              np.full_like(k, a[n]), score[n] - 0.001 * k, np.full_like(k, occ[n, 6])], 1))
    return torch.from_numpy(np.concatenate([occ] + rows, 0).astype(np.float32))


def host_trig(bb_all_cpu):
    """(N,2) fp32 (sin, cos) of the angle column, evaluated the way the reference evaluates them: ``torch.sin`` /
    ``torch.cos`` on the CPU over the STRIDED column view of a (.,7) box tensor (``obb2xyxy``,
    syn_images_generator_v2.py:382-396; ``obb2poly_le90``, data_augument_bank.py:516-541).  ATen's CPU kernels take
    the scalar libm loop for such a view, so the value of an element does not depend on where it sits in the tensor."""
    a = bb_all_cpu.float().contiguous()[:, 4]
    return torch.stack([torch.sin(a), torch.cos(a)], 1).contiguous()


def black_paper_from_candidates(img_syn, bb_all, imgsize, return_debug=False, trig=None):
    """Device tail (:664-690).  img_syn (C,H,W) fp32 CUDA, modified in place; bb_all (N,7).
    Returns (img_syn, kept boxes (M,7)); one host read of the survivor count (the reference returns a
    dynamically sized tensor too).

    Bit-exactness of the filled pixel set: the polygon corners (and the inside-image filter) go through sin / cos of
    the box angle.  When ``bb_all`` arrives on the HOST -- as it does from ``generate_black_paper``, whose candidates
    are drawn from the CPU generators -- the (N,2) trig table is computed there with the reference's own torch ops
    (``host_trig``) and uploaded with the boxes: corners, filter decisions and pixels are then identical to the
    reference's.  Device-resident candidates without ``trig`` use the device's sincos (rounded once from double):
    a corner sitting on an integer boundary may truncate differently (<= 1 px on < 1 % of the corners)."""
    if not img_syn.is_cuda:
        raise ValueError("img_syn: expected a CUDA tensor (this path has no CPU fallback)")
    if trig is None and not bb_all.is_cuda:
        trig = host_trig(bb_all)
    if trig is not None:
        trig = trig.to(img_syn.device, non_blocking=True).float().contiguous()
    bb_all = bb_all.to(img_syn.device).float().contiguous()
    order, keep = ops.nms_rotated(bb_all, bb_all[:, 5], 0.05)
    out, sel, polys, count = ops.black_paper_select(bb_all, order, keep, imgsize, trig)
    ops.fill_polys(polys, img=img_syn, value=255.0, count=count)
    m = int(count.item())
    if return_debug:
        return img_syn, out[:m], dict(sel=sel[:m], polys=polys[:m], order=order, keep_sorted=keep)
    return img_syn, out[:m]


def generate_black_paper(img, bb_occupied, img_syn, pattern, prior_size, dense_cls, imgsize, candidates=None):
    """Reference signature (:591-592) plus ``candidates`` to inject a pre-drawn (N,7) box list.  ``img`` / ``pattern``
    are accepted for compatibility (the reference only uses them for the unused palette)."""
    if candidates is None:
        candidates = sample_black_paper_candidates(bb_occupied, prior_size, dense_cls, imgsize)
    if img_syn.dtype != torch.float32:
        raise ValueError("img_syn must be fp32")
    if not img_syn.is_contiguous():
        raise ValueError("img_syn must be contiguous")
    return black_paper_from_candidates(img_syn, candidates, imgsize)
