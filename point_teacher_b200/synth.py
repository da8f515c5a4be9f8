"""Synthetic AI-TOD-v2 / SODA-A shaped batches.  Everything is
drawn from a seeded CPU ``torch.Generator`` and then copied to the device, so the CPU
oracle and the CUDA path see bit-identical inputs."""
import math

import torch

# shipped stage-0 proposal dicts, HBB_TOD/configs/point_teacher/aitodv2_point_teacher_0%.py:127-166
HBB_FINE_CFG = [dict(gen_mode="refine", gen_proposal_mode="fix_gen", cut_mode=None, shake_ratio=None,
                     base_ratios=[1.0], min_scale=0, pos_iou_thr=0.3, neg_iou_thr=0.3, gen_num_neg=200),
                dict(gen_mode="refine", gen_proposal_mode="fix_gen", cut_mode=None, shake_ratio=None,
                     base_ratios=[1.0], min_scale=4, pos_iou_thr=0.3, neg_iou_thr=0.3, gen_num_neg=200)]
HBB_EXT_CFG = [dict(gen_mode="refine", gen_proposal_mode="fix_gen", cut_mode=None, shake_ratio=None,
                    base_ratios=[1.0, 1.2, 1.3, 0.8, 0.7], min_scale=4, pos_iou_thr=0.3, neg_iou_thr=0.3,
                    gen_num_neg=0),
               dict(gen_mode="refine", gen_proposal_mode="fix_gen", cut_mode=None, shake_ratio=[0.1],
                    base_ratios=[1.0, 1.2, 1.3, 0.8, 0.7], min_scale=16, pos_iou_thr=0.3, neg_iou_thr=0.3,
                    gen_num_neg=0)]
# OBB_TOD/configs/point teacher/sodaa_fcos_pointteacher_1x.py:134-175
OBB_FINE_CFG = [dict(gen_mode="refine", gen_proposal_mode="fix_gen", cut_mode=None, shake_ratio=None,
                     base_ratios=[1.0], min_scale=0, pos_iou_thr=0.3, neg_iou_thr=0.3, gen_num_neg=200)]
OBB_EXT_CFG = [dict(gen_mode="refine", gen_proposal_mode="fix_gen", cut_mode=None, shake_ratio=None,
                    base_ratios=[1.0, 1.2, 1.3, 0.8, 0.6], min_scale=4, pos_iou_thr=0.3, neg_iou_thr=0.3,
                    gen_num_neg=0)]


def stress_ext_cfg(n_ratios=8):
    """Config #4: bag of n_ratios^2 instances (64 at the default)."""
    ratios = [round(0.6 + 0.1 * i, 2) for i in range(n_ratios)]
    return [dict(gen_mode="refine", gen_proposal_mode="fix_gen", cut_mode=None, shake_ratio=None,
                 base_ratios=ratios, min_scale=4, pos_iou_thr=0.3, neg_iou_thr=0.3, gen_num_neg=0)]


def make_boxes(g, n, img_hw, median=12.0, sigma=0.5, lo=2.0, hi=64.0):
    h, w = img_hw
    cx = torch.rand(n, generator=g) * (w - 16) + 8
    cy = torch.rand(n, generator=g) * (h - 16) + 8
    bw = (torch.randn(n, generator=g) * sigma + math.log(median)).exp().clamp(lo, hi)
    bh = (torch.randn(n, generator=g) * sigma + math.log(median)).exp().clamp(lo, hi)
    return torch.stack([cx - bw / 2, cy - bh / 2, cx + bw / 2, cy + bh / 2], 1)


def jitter_boxes(g, boxes, ctr_sigma=2.0, log_sigma=0.3):
    n = boxes.shape[0]
    cx = (boxes[:, 0] + boxes[:, 2]) / 2 + torch.randn(n, generator=g) * ctr_sigma
    cy = (boxes[:, 1] + boxes[:, 3]) / 2 + torch.randn(n, generator=g) * ctr_sigma
    bw = (boxes[:, 2] - boxes[:, 0]) * (torch.randn(n, generator=g) * log_sigma).exp()
    bh = (boxes[:, 3] - boxes[:, 1]) * (torch.randn(n, generator=g) * log_sigma).exp()
    return torch.stack([cx - bw / 2, cy - bh / 2, cx + bw / 2, cy + bh / 2], 1)


def hbb_batch(seed=0, batch=2, img_hw=(800, 800), stride=8, channels=256, num_classes=8,
              gt_range=(200, 600), n_neg=200, num_stages=1):
    """Config #1 inputs: feature map ~N(0,1), G_i ~ U{gt_range} log-normal boxes, coarse
    pseudo boxes = jittered GT, injected negative boxes (the reference draws them from
    the CPU RNG at syn_images_generator_v2.py:247-250, which no device can reproduce)."""
    g = torch.Generator().manual_seed(seed)
    h, w = img_hw
    feat = torch.randn(batch, channels, h // stride, w // stride, generator=g)
    gts, pseudo, labels, negs = [], [], [], []
    for _ in range(batch):
        n = int(torch.randint(gt_range[0], gt_range[1] + 1, (1,), generator=g))
        gt = make_boxes(g, n, img_hw)
        gts.append(gt)
        pseudo.append(jitter_boxes(g, gt))
        labels.append(torch.randint(0, num_classes, (n,), generator=g))
    for _ in range(num_stages):
        per_img = []
        for _ in range(batch):
            x1 = torch.rand(n_neg, generator=g) * w * 0.8
            y1 = torch.rand(n_neg, generator=g) * h * 0.8
            x2 = x1 + torch.rand(n_neg, generator=g) * 100
            y2 = y1 + torch.rand(n_neg, generator=g) * 100
            per_img.append(torch.stack([x1, y1, x2, y2], 1))
        negs.append(per_img)
    metas = [dict(img_shape=(h, w, 3)) for _ in range(batch)]
    points = [torch.stack([(b[:, 0] + b[:, 2]) / 2, (b[:, 1] + b[:, 3]) / 2], 1) for b in pseudo]
    return dict(feat=feat, gt_boxes=gts, pseudo_boxes=pseudo, pseudo_points=points,
                pseudo_labels=labels, neg_boxes=negs, img_metas=metas, stride=stride,
                num_classes=num_classes)


def obb_batch(seed=0, batch=2, img_hw=(1024, 1024), stride=8, channels=256, num_classes=9,
              gt_range=(200, 600), n_neg=200):
    """Config #3 inputs: 5-d boxes (cx, cy, w, h, theta), theta ~ U[-pi/2, pi/2)."""
    g = torch.Generator().manual_seed(seed)
    h, w = img_hw
    feat = torch.randn(batch, channels, h // stride, w // stride, generator=g)
    gts, pseudo, labels, negs = [], [], [], []
    for _ in range(batch):
        n = int(torch.randint(gt_range[0], gt_range[1] + 1, (1,), generator=g))
        xyxy = make_boxes(g, n, img_hw, median=14.0)
        th = torch.rand(n, generator=g) * math.pi - math.pi / 2
        jit = jitter_boxes(g, xyxy)

        def to5(b):
            return torch.stack([(b[:, 0] + b[:, 2]) / 2, (b[:, 1] + b[:, 3]) / 2, b[:, 2] - b[:, 0],
                                b[:, 3] - b[:, 1], th], 1)
        gts.append(to5(xyxy))
        pseudo.append(to5(jit))
        labels.append(torch.randint(0, num_classes, (n,), generator=g))
    per_img = []
    for _ in range(batch):
        x1 = torch.rand(n_neg, generator=g) * w * 0.8
        y1 = torch.rand(n_neg, generator=g) * h * 0.8
        x2 = x1 + torch.rand(n_neg, generator=g) * 200
        y2 = y1 + torch.rand(n_neg, generator=g) * 200
        a = torch.rand(n_neg, generator=g) * math.pi - math.pi / 2
        per_img.append(torch.stack([x1, y1, x2, y2, a], 1))
    negs.append(per_img)
    metas = [dict(img_shape=(h, w, 3)) for _ in range(batch)]
    return dict(feat=feat, gt_boxes=gts, pseudo_boxes=pseudo, pseudo_points=[b[:, :2] for b in pseudo],
                pseudo_labels=labels, neg_boxes=negs, img_metas=metas, stride=stride,
                num_classes=num_classes)


def assign_batch(seed, P_hw=(40, 40), G=37, C=8, stride=8, ties=True):
    """Grid points of a stride-8 map, GT points (integer-valued when ``ties`` so that L1 ties are everywhere),
    logits, decoded boxes."""
    g = torch.Generator().manual_seed(seed)
    h, w = P_hw
    ys, xs = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    points = torch.stack([xs.reshape(-1), ys.reshape(-1)], 1).float() * stride + stride // 2
    P = points.shape[0]
    ctr = torch.rand(G, 2, generator=g) * torch.tensor([w * stride - 16.0, h * stride - 16.0]) + 8
    if ties:
        ctr = ctr.round()
    wh = (torch.randn(G, 2, generator=g) * 0.5 + 2.5).exp().clamp(2, 64)
    gt_cxcywh = torch.cat([ctr, wh], 1)
    labels = torch.randint(0, C, (G,), generator=g)
    logits = torch.randn(P, C, generator=g) * 2 - 2
    pred_wh = (torch.randn(P, 2, generator=g) * 0.6 + 2.8).exp()
    pred = torch.cat([points + torch.randn(P, 2, generator=g) * 4, pred_wh], 1)       # (cx, cy, w, h)
    return dict(points=points, gt=gt_cxcywh, labels=labels, logits=logits, pred=pred)


# HBB_TOD/configs/point_teacher/aitodv2_point_teacher_0%.py: shape_list of the phase-1 masking priors
SHAPE_LIST = [[20, 20, 0.5, 0.5], [30, 120, 0.5, 0.5], [10, 20, 0.5, 0.5], [20, 50, 0.5, 0.5], [30, 20, 0.5, 0.5]]


def mask_batch(seed, imgsize=800, n_gt=(150, 300), n_shapes=5):
    """Phase-1 masking inputs: one uint8-valued float image (3,H,W) and bb_occupied (G,7) =
    (cx, cy, w, h, 0, 1, shape class) as built by fcos_p2b_teacher_student.py:470-479."""
    g = torch.Generator().manual_seed(1000 + seed)
    n = int(torch.randint(n_gt[0], n_gt[1] + 1, (1,), generator=g))
    img = torch.randint(0, 256, (3, imgsize, imgsize), generator=g).float()
    xyxy = make_boxes(g, n, (imgsize, imgsize))
    cxcywh = torch.stack([(xyxy[:, 0] + xyxy[:, 2]) / 2, (xyxy[:, 1] + xyxy[:, 3]) / 2, xyxy[:, 2] - xyxy[:, 0],
                          xyxy[:, 3] - xyxy[:, 1]], 1)
    cls = torch.randint(0, n_shapes, (n, 1), generator=g).float()
    bb = torch.cat([cxcywh, torch.zeros(n, 1), torch.ones(n, 1), cls], 1)
    return dict(img=img, bb_occupied=bb, imgsize=imgsize)


def pseudo_batch(seed, P_hw=(100, 100), G=120, C=8, stride=8):
    """Inputs of the coarse pseudo-box step: grid points, FCOS-style (l, t, r, b) distances and logits, GT points
    (integer valued: L1 ties), GT boxes (xyxy) and labels."""
    d = assign_batch(seed, P_hw=P_hw, G=G, C=C, stride=stride, ties=True)
    g = torch.Generator().manual_seed(5000 + seed)
    P = d["points"].shape[0]
    ltrb = (torch.randn(P, 4, generator=g) * 0.5 + 2.0).exp()
    ctr, wh = d["gt"][:, :2], d["gt"][:, 2:]
    gt_boxes = torch.cat([ctr - wh / 2, ctr + wh / 2], 1)
    logits = d["logits"].clone()
    logits[torch.randperm(P, generator=g)[:P // 3]] += 3.0          # some confident points: scores above the filter
    return dict(points=d["points"], ltrb=ltrb, logits=logits, gt_points=ctr.clone(), gt_boxes=gt_boxes,
                labels=d["labels"])


def augment_batch(seed=0, batch=2, img_hw=(160, 176), n=30, rotated=False, num_classes=8):
    """Inputs of ``strong_augmentation`` (section 8f rank 3): uint8-valued fp32 images, GT points, pseudo points /
    labels / boxes (xyxy, or (cx,cy,w,h,theta) when ``rotated``), a few of them outside the image."""
    g = torch.Generator().manual_seed(seed)
    h, w = img_hw
    img = torch.randint(0, 256, (batch, 3, h, w), generator=g).float()
    wh_t = torch.tensor([w, h], dtype=torch.float32)
    gtp = [torch.rand(n, 2, generator=g) * wh_t for _ in range(batch)]
    gtl = [torch.randint(0, num_classes, (n,), generator=g) for _ in range(batch)]
    pp = [p + torch.randn(n, 2, generator=g) * 2 for p in gtp]
    pl = [l.clone() for l in gtl]
    sz = [torch.rand(n, 2, generator=g) * 30 + 4 for _ in range(batch)]
    if rotated:
        pb = [torch.cat([p, s, torch.rand(n, 1, generator=g) * math.pi - math.pi / 2], 1) for p, s in zip(pp, sz)]
    else:
        pb = [torch.cat([p - s / 2, p + s / 2], 1) for p, s in zip(pp, sz)]
    return dict(img=img, gt_points=gtp, gt_labels=gtl, pseudo_points=pp, pseudo_labels=pl, pseudo_bboxes=pb)
