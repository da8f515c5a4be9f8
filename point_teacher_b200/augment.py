"""``strong_augmentation`` of the teacher-student detectors on the device.

Drop-in for the module-level function the detectors import by name
(HBB_TOD/mmdet/models/detectors/fcos_p2b_teacher_student.py:13,196,237 -> syn_images_generator_v2.py:24-132;
OBB_TOD/mmrotate/models/detectors/rotated_fcos_teacher_student.py:24,231,290 -> syn_images_generator_v2.py:223-357):
same positional arguments, same 7-tuple back.  The random draws are made on the host with the reference's own calls in
the reference's order (``random.choice`` flip, ``np.random.randint(1, 20)`` angle for OBB, ``np.random.uniform(0.8,
1.2)`` rounded to one decimal), so seeding ``random`` / ``np.random`` reproduces the reference's augmentation; pass
``choices`` to inject them instead.  Everything else is two kernels (csrc/augment.cu) and ONE device->host read of the
kept-element counts (the reference synchronises at every ``nonzero``)."""
import math
import random

import numpy as np
import torch

from . import _lib
from .ops import _p, _stream

FLIPS = ("horizontal", "vertical", "diagonal", "None")


def draw_choices(batch, rotated=False):
    """[(flip, angle_deg, scale_factor)] per image, drawn exactly like the reference draws them."""
    out = []
    for _ in range(batch):
        flip = random.choice(list(FLIPS))
        angle = int(np.random.randint(1, 20)) if rotated else 0
        sf = float(np.around(np.random.uniform(0.8, 1.2), 1))
        out.append((flip, angle, sf))
    return out


def _rotate_matrix(angle):
    """torchvision ``_get_inverse_affine_matrix([0, 0], -angle, [0, 0], 1.0, [0, 0])`` (what ``TF.rotate`` builds)."""
    rot = math.radians(-angle)
    a, b, c, d = math.cos(rot), -math.sin(rot), math.sin(rot), math.cos(rot)
    return [d, -b, 0.0, -c, a, 0.0]


def image_params(choices, H, W, rotated):
    """Per-image parameter block of csrc/augment.cu (fp32; the host-side scalar arithmetic of the reference)."""
    stride = _lib.load().pt_augment_param_stride()
    P = np.zeros((len(choices), stride), np.float32)
    for i, (flip, angle, sf) in enumerate(choices):
        if flip not in FLIPS:
            raise ValueError(f"unknown flip {flip!r}")
        sH, sW = int(H * sf), int(W * sf)
        pad = sf < 1.0
        if pad:
            blank_h, blank_w = int((H - sH) / 2), int((W - sW) / 2)
            start_y, start_x = (H - sH) // 2, (W - sW) // 2
        else:
            blank_h, blank_w = int((sH - H) / 2), int((sW - W) / 2)
            start_y, start_x = (sH - H) // 2, (sW - W) // 2
        P[i, 0] = flip in ("horizontal", "diagonal")
        P[i, 1] = flip in ("vertical", "diagonal")
        if rotated:
            m = np.asarray(_rotate_matrix(angle), np.float32).reshape(2, 3)
            r = (m.T / np.asarray([0.5 * W, 0.5 * H], np.float32)).astype(np.float32)
            rad = np.deg2rad(-angle)
            P[i, 2] = 1.0
            P[i, 3], P[i, 4], P[i, 5], P[i, 6] = r[0, 0], r[0, 1], r[1, 0], r[1, 1]
            P[i, 13], P[i, 14] = np.float32(np.cos(rad)), np.float32(np.sin(rad))
        P[i, 7], P[i, 8], P[i, 9], P[i, 10] = sH, sW, start_y, start_x
        P[i, 11] = pad
        P[i, 12] = np.float32(sf)
        P[i, 15], P[i, 16] = blank_w, blank_h
    return P


def _pack(lists, cols, dtype, dev):
    off = [0]
    for t in lists:
        off.append(off[-1] + t.shape[0])
    if off[-1] == 0:
        flat = torch.zeros((0, cols) if cols else (0,), dtype=dtype, device=dev)
    else:
        flat = torch.cat([t.reshape(-1, cols) if cols else t.reshape(-1) for t in lists]).to(dtype).contiguous()
    return flat, off


def strong_augmentation(img, gt_points, gt_labels, pseudo_points, pseudo_labels, pseudo_bboxes, angle_version=None,
                        choices=None):
    """-> (aug_images (B,C,H,W), [image_i], [gt_points_i], [gt_labels_i], [pseudo_points_i], [pseudo_labels_i],
    [pseudo_bboxes_i]).  ``angle_version=None``: HBB (xyxy boxes); ``'le90'``: OBB (cx,cy,w,h,theta)."""
    if angle_version not in (None, "le90"):
        raise NotImplementedError("the Point Teacher OBB configs use angle_version='le90'")
    if img.dim() != 4 or img.dtype != torch.float32 or not img.is_cuda:
        raise ValueError("img must be a CUDA fp32 (B,C,H,W) tensor")
    rotated = angle_version is not None
    B, C, H, W = img.shape
    dev = img.device
    if choices is None:
        choices = draw_choices(B, rotated)
    if len(choices) != B:
        raise ValueError("one (flip, angle, scale) triple per image")
    bd = 5 if rotated else 4
    params = torch.from_numpy(image_params(choices, H, W, rotated)).to(dev, non_blocking=True)
    out = torch.empty_like(img)
    _lib.call("pt_augment_image", _p(img.contiguous()), _p(out), _p(params), B, C, H, W, _stream())
    gp, goff = _pack(gt_points, 2, torch.float32, dev)
    gl, _ = _pack(gt_labels, 0, torch.int64, dev)
    pp, poff = _pack(pseudo_points, 2, torch.float32, dev)
    pl, _ = _pack(pseudo_labels, 0, torch.int64, dev)
    pb, _ = _pack(pseudo_bboxes, bd, torch.float32, dev)
    if pb.shape[0] != pp.shape[0] or gl.shape[0] != gp.shape[0] or pl.shape[0] != pp.shape[0]:
        raise ValueError("points / labels / boxes of an image must have the same length")
    goff_t = torch.tensor(goff, dtype=torch.int32).to(dev, non_blocking=True)
    poff_t = torch.tensor(poff, dtype=torch.int32).to(dev, non_blocking=True)
    ogp, ogl, opp, opl, opb = (torch.empty_like(t) for t in (gp, gl, pp, pl, pb))
    counts = torch.empty((B, 2), dtype=torch.int32, device=dev)
    _lib.call("pt_augment_coords", _p(gp), _p(gl), _p(goff_t), _p(pp), _p(pl), _p(pb), _p(poff_t), bd, _p(params), B,
              H, W, _p(ogp), _p(ogl), _p(opp), _p(opl), _p(opb), _p(counts), _stream())
    cnt = counts.tolist()                                  # the one host read: lengths of the returned lists
    ldt, pdt = gt_labels[0].dtype if len(gt_labels) else torch.int64, gt_points[0].dtype if len(gt_points) else torch.float32
    imgs, gps, gls, pps, pls, pbs = [], [], [], [], [], []
    for i in range(B):
        ng, npz = cnt[i]
        imgs.append(out[i])
        gps.append(ogp[goff[i]:goff[i] + ng].to(pdt))
        gls.append(ogl[goff[i]:goff[i] + ng].to(ldt))
        pps.append(opp[poff[i]:poff[i] + npz].to(pdt))
        pls.append(opl[poff[i]:poff[i] + npz].to(pseudo_labels[i].dtype))
        pbs.append(opb[poff[i]:poff[i] + npz].to(pdt))
    return out, imgs, gps, gls, pps, pls, pbs
