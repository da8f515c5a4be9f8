"""BASELINE.json config #2: "Point Teacher HBB AI-TOD-v2 full teacher-student training step (phase-1 masking +
phase-2 MIL), synthetic batch, random-init R50-FPN on 1 B200" -- ``python bench.py --config train``.

The step follows ``TS_P2B_FCOS.forward_train`` (HBB_TOD/mmdet/models/detectors/fcos_p2b_teacher_student.py:116-253)
stage by stage.  Everything SURVEY.md section 8 puts on the path runs on this repo's kernels through the reference's
surface; what section 8 marks out of scope is STOCK PYTORCH and is named as such in the JSON line:

  ours   phase-1 region masking ``generate_black_paper`` (a16), coarse pseudo boxes ``generate_pseudo_single`` behind
         ``FUSETopkAssigner`` (a14, f1), ``forward_mil_head_burn_in_step1/2`` = ``refine.phase1_refine`` /
         ``phase2_refine`` with autograd (a1-a12, differentiable), ``strong_augmentation`` (f3), dense targets
         ``get_target_pseudo_single`` behind ``TopkAssigner`` x2 (a13, f2), ``FocalLoss`` (f4)
  stock  ResNet-50 (torchvision, random init) + FPN + a PSAGG-style aggregation neck to ONE stride-8 256-channel map
         (HBB_TOD/mmdet/models/necks/ps_fpn.py:56-75), the FCOS tower (4 x conv3x3 + GN + ReLU per branch), GIoU and
         centerness BCE terms of the dense loss, EMA teacher update, SGD(momentum) step, fp32 (TF32 convolutions as
         torch defaults allow; the reference uses no AMP)

One "step" = one phase-1 iteration followed by one phase-2 iteration (the shipped recipe runs 4 000 of the former, then
the latter); both are timed separately as well.  2 images 800x800, GT counts / boxes / labels of config #1."""
import json
import math
import os
import time

import torch
import torch.nn as nn
import torch.nn.functional as F


class Backbone(nn.Module):
    """Stock: torchvision ResNet-50 -> FPN (C3..C5 + two extra levels) -> PSAGG-like top-down aggregation into the
    stride-8 level (ps_fpn.py: lateral conv on the coarsest input, then repeatedly upsample + add + 3x3 conv)."""

    def __init__(self):
        super().__init__()
        import torchvision
        r = torchvision.models.resnet50(weights=None)
        self.stem = nn.Sequential(r.conv1, r.bn1, r.relu, r.maxpool)
        self.layers = nn.ModuleList([r.layer1, r.layer2, r.layer3, r.layer4])
        self.lateral = nn.ModuleList([nn.Conv2d(c, 256, 1) for c in (512, 1024, 2048)])
        self.fpn_out = nn.ModuleList([nn.Conv2d(256, 256, 3, padding=1) for _ in range(3)])
        self.extra = nn.ModuleList([nn.Conv2d(256, 256, 3, stride=2, padding=1) for _ in range(2)])
        self.agg = nn.ModuleList([nn.Conv2d(256, 256, 3, padding=1) for _ in range(5)])

    def forward(self, img):
        x = self.stem(img)
        feats = []
        for i, l in enumerate(self.layers):
            x = l(x)
            if i >= 1:
                feats.append(x)
        lat = [l(f) for l, f in zip(self.lateral, feats)]
        for i in (1, 0):
            lat[i] = lat[i] + F.interpolate(lat[i + 1], size=lat[i].shape[2:], mode="nearest")
        outs = [c(x) for c, x in zip(self.fpn_out, lat)]
        outs.append(self.extra[0](outs[-1]))
        outs.append(self.extra[1](F.relu(outs[-1])))
        outs[-1] = self.agg[0](outs[-1])
        for i in range(4, 0, -1):
            outs[i - 1] = self.agg[5 - i](outs[i - 1] + F.interpolate(outs[i], size=outs[i - 1].shape[2:], mode="nearest"))
        return (outs[0],)                       # ONE stride-8 map, like the shipped config (featmap_strides=[8])


class FCOSTower(nn.Module):
    """Stock: the dense part of TS_P2BFCOSHead (anchor_free_head towers + conv_cls / conv_reg / conv_centerness)."""

    def __init__(self, num_classes=8, stride=8):
        super().__init__()
        def tower():
            return nn.Sequential(*[m for _ in range(4) for m in (nn.Conv2d(256, 256, 3, padding=1, bias=False),
                                                                 nn.GroupNorm(32, 256), nn.ReLU(inplace=True))])
        self.cls_tower, self.reg_tower = tower(), tower()
        self.conv_cls = nn.Conv2d(256, num_classes, 3, padding=1)
        self.conv_reg = nn.Conv2d(256, 4, 3, padding=1)
        self.conv_ctr = nn.Conv2d(256, 1, 3, padding=1)
        self.scale = nn.Parameter(torch.tensor(1.0))
        self.stride = stride
        nn.init.constant_(self.conv_cls.bias, -math.log(99))

    def forward(self, x):
        c, r = self.cls_tower(x), self.reg_tower(x)
        B = x.shape[0]
        cls = self.conv_cls(c).permute(0, 2, 3, 1).reshape(B, -1, self.conv_cls.out_channels)
        reg = (self.scale * self.conv_reg(r)).float().exp().permute(0, 2, 3, 1).reshape(B, -1, 4) * self.stride
        ctr = self.conv_ctr(r).permute(0, 2, 3, 1).reshape(B, -1)
        return cls, reg, ctr


def grid_points(h, w, stride, dev):
    ys, xs = torch.meshgrid(torch.arange(h, device=dev), torch.arange(w, device=dev), indexing="ij")
    return (torch.stack([xs.reshape(-1), ys.reshape(-1)], 1).float() * stride + stride // 2).contiguous()


def giou_loss(pred, target, eps=1e-7):
    """Stock torch (mmdet GIoULoss on decoded boxes)."""
    lt, rb = torch.max(pred[:, :2], target[:, :2]), torch.min(pred[:, 2:], target[:, 2:])
    wh = (rb - lt).clamp(min=0)
    inter = wh[:, 0] * wh[:, 1]
    ap = (pred[:, 2] - pred[:, 0]) * (pred[:, 3] - pred[:, 1])
    at = (target[:, 2] - target[:, 0]) * (target[:, 3] - target[:, 1])
    union = ap + at - inter + eps
    elt, erb = torch.min(pred[:, :2], target[:, :2]), torch.max(pred[:, 2:], target[:, 2:])
    ewh = (erb - elt).clamp(min=0)
    earea = ewh[:, 0] * ewh[:, 1] + eps
    return 1 - (inter / union - (earea - union) / earea)


def dense_loss(tower_out, points, labels_cls, labels_reg, targets, num_classes, focal):
    """FCOS ``loss`` / ``loss_pseudo`` tail (fcos_head_p2b_ts.py:392-534) from per-image targets: focal (ours),
    GIoU + centerness BCE (stock)."""
    from point_teacher_b200.coarse import centerness_target
    cls, reg, ctr = tower_out
    cls_f, reg_f, ctr_f = cls.reshape(-1, num_classes), reg.reshape(-1, 4), ctr.reshape(-1)
    lc, lr, tg = torch.cat(labels_cls), torch.cat(labels_reg), torch.cat(targets)
    num_pos = max(int((lc < num_classes).sum()), 1)
    loss_cls = focal(cls_f.float().contiguous(), lc.contiguous(), avg_factor=num_pos)
    pos = (lr < num_classes).nonzero().reshape(-1)
    if pos.numel() == 0:
        z = reg_f.sum() * 0
        return loss_cls, z, z
    pts = points.repeat(cls.shape[0], 1)[pos]
    pt, pp = tg[pos], reg_f[pos]
    cen = centerness_target(pt)
    dec = lambda p, d: torch.stack([p[:, 0] - d[:, 0], p[:, 1] - d[:, 1], p[:, 0] + d[:, 2], p[:, 1] + d[:, 3]], 1)  # noqa: E731
    loss_box = (giou_loss(dec(pts, pp), dec(pts, pt)) * cen).sum() / cen.sum().clamp(min=1e-6)
    loss_ctr = F.binary_cross_entropy_with_logits(ctr_f[pos], cen)
    return loss_cls, loss_box, loss_ctr


class Detector(nn.Module):
    def __init__(self, dev):
        super().__init__()
        import point_teacher_b200
        from point_teacher_b200 import registry
        point_teacher_b200.install()
        self.backbone, self.tower = Backbone(), FCOSTower()
        self.bbox_head = registry.build_head(dict(type="TS_P2BFCOSHead", num_classes=8, in_channels=256, num_stages=1,
                                                  top_k=1, beta=0.25, precision="bf16"))
        self.to(dev)


def run(args, ClockSampler, peaks):
    import numpy as np
    from point_teacher_b200 import _lib, augment, coarse, masking, registry, synth
    from point_teacher_b200.losses import FocalLoss
    from point_teacher_b200.refine import phase1_refine, phase2_refine
    dev = torch.device("cuda", 0)
    _lib.load()
    torch.manual_seed(0)
    np.random.seed(0)
    d = synth.hbb_batch(seed=0)
    C, stride, cap = 8, 8, 100
    student, teacher = Detector(dev), Detector(dev)
    teacher.load_state_dict(student.state_dict())
    for p in teacher.parameters():
        p.requires_grad_(False)
    for n, p in student.bbox_head.named_parameters():        # constructed-but-unused reference modules: no gradient
        if n.startswith(("shared_fcs.", "shared_fcs_refine.", "fc_iou.")):
            p.requires_grad_(False)
    params = [p for p in student.parameters() if p.requires_grad]
    opt = torch.optim.SGD(params, lr=0.005, momentum=0.9, weight_decay=1e-4)
    kw = dict(cls_cost=dict(type="FocalLossCost", weight=1.0), reg_cost=dict(type="PointCost", mode="L1", weight=1.0))
    assigner = registry.build_assigner(dict(type="TopkAssigner", num_pre=1, topk=1, **kw))
    pseudo_assigner = registry.build_assigner(dict(type="TopkAssigner", num_pre=3, topk=3, **kw))
    syn_assigner = registry.build_assigner(dict(type="TopkAssigner", num_pre=3, topk=3, **kw))
    fuse_assigner = registry.build_assigner(dict(type="FUSETopkAssigner", num_pre=5, topk=3,
                                                 location_cost=dict(type="InsiderCost", weight=1.0), **kw))
    focal = FocalLoss(use_sigmoid=True, gamma=2.0, alpha=0.25, loss_weight=1.0)
    to = lambda l: [t.to(dev) for t in l]  # noqa: E731
    img = torch.randint(0, 256, (2, 3, 800, 800), generator=torch.Generator().manual_seed(1)).float().to(dev)
    gt_boxes, gt_labels, metas = to(d["gt_boxes"]), to(d["pseudo_labels"]), d["img_metas"]
    gt_points = [torch.stack([(b[:, 0] + b[:, 2]) / 2, (b[:, 1] + b[:, 3]) / 2], 1) for b in gt_boxes]
    points = grid_points(100, 100, stride, dev)
    pattern, prior = masking.load_basic_shape(synth.SHAPE_LIST)
    n_shapes = len(synth.SHAPE_LIST)
    fine, ext = synth.HBB_FINE_CFG, synth.HBB_EXT_CFG
    head = student.bbox_head
    rng = np.random.default_rng(0)

    def ema():
        with torch.no_grad():
            sp, tp = list(student.parameters()), list(teacher.parameters())
            torch._foreach_mul_(tp, 0.999)
            torch._foreach_add_(tp, sp, alpha=0.001)

    def coarse_boxes():
        """Teacher forward + coarse pseudo boxes (get_pseudo_bbox, fcos_head_p2b_ts.py:710-794)."""
        with torch.no_grad():
            cls, reg, ctr = teacher.tower(teacher.backbone(img)[0])
            pb, pp, pl = [], [], []
            for i in range(2):
                b, p, l, _, _ = coarse.generate_pseudo_single(fuse_assigner, gt_points[i], gt_labels[i], gt_boxes[i],
                                                              cls[i], reg[i], ctr[i], metas[i], None, 0.0, points)
                pb.append(b), pp.append(p), pl.append(l)
        return pb, pp, pl

    def targets_pseudo(tower_out, gp, gl, pp, pl, pb, burn1):
        lcs, lrs, tgs = [], [], []
        for i in range(2):
            lr, tg, lc, _ = coarse.get_target_pseudo_single(assigner, pseudo_assigner, C, gp[i], gl[i], pp[i], pl[i], pb[i],
                                                            tower_out[0][i].detach(), tower_out[1][i].detach(),
                                                            tower_out[2][i].detach(), metas[i], None, None, points,
                                                            burn_in_step1=burn1)
            lcs.append(lc), lrs.append(lr), tgs.append(tg)
        return lcs, lrs, tgs

    def second_half(losses, pb, pp, pl, burn1):
        """strong_augmentation -> student on the augmented image -> loss_pseudo (:196-209 / :237-250)."""
        img_aug, _, gp, gl, pps, pls, pbs = augment.strong_augmentation(img, gt_points, gt_labels, pp, pl, pb)
        out = student.tower(student.backbone(img_aug)[0])
        lcs, lrs, tgs = targets_pseudo(out, gp, gl, pps, pls, pbs, burn1)
        lc, lb, lct = dense_loss(out, points, lcs, lrs, tgs, C, focal)
        losses["loss_cls"] = lc
        if not burn1:
            losses["loss_bbox"], losses["loss_centerness"] = lb, lct
        total = sum(v for k, v in losses.items() if "loss" in k)          # _parse_losses
        opt.zero_grad(set_to_none=True)
        total.backward()
        opt.step()
        return total.detach()

    def phase1():
        ema()
        # genrate_syn (:469-502): masking on the device, synthetic boxes = enclosing xyxy of the kept regions
        img_syn = img.clone()
        syn = []
        for i in range(2):
            bb = torch.cat([gt_points[i], torch.zeros((gt_points[i].shape[0], 2), device=dev),
                            torch.zeros((gt_points[i].shape[0], 1), device=dev), torch.ones((gt_points[i].shape[0], 1), device=dev),
                            (gt_labels[i] % n_shapes).float()[:, None]], 1)
            cand = masking.sample_black_paper_candidates_fast(bb, prior, range(n_shapes // 2), 800, rng)
            _, kept = masking.generate_black_paper(img[i], bb, img_syn[i], pattern, prior, range(n_shapes // 2), 800,
                                                   candidates=cand)
            ca, sa = kept[:, 4].cos().abs(), kept[:, 4].sin().abs()
            bw, bh = ca * kept[:, 2] + sa * kept[:, 3], sa * kept[:, 2] + ca * kept[:, 3]
            syn.append(torch.stack([kept[:, 0] - bw / 2, kept[:, 1] - bh / 2, kept[:, 0] + bw / 2, kept[:, 1] + bh / 2], 1))
        x_all = student.backbone(torch.cat([img_syn, img]))[0]
        x_syn, x_ori = x_all[:2], x_all[2:]
        out_syn = student.tower(x_syn)
        # loss on the synthetic boxes (:168-169): syn_assigner targets + dense loss
        lcs, lrs, tgs = [], [], []
        for i in range(2):
            sl = torch.zeros((syn[i].shape[0],), dtype=torch.long, device=dev)
            cx = torch.cat([(syn[i][:, :2] + syn[i][:, 2:]) / 2, syn[i][:, 2:] - syn[i][:, :2]], 1)
            res = syn_assigner.assign(points, out_syn[0][i].detach().contiguous(), cx, sl)
            from point_teacher_b200 import ops
            tg, lr, _ = ops.ltrb_targets(points, syn[i].contiguous(), res.gt_inds, res.labels, C)
            lcs.append(lr), lrs.append(lr), tgs.append(tg)
        _, l_box, l_ctr = dense_loss(out_syn, points, lcs, lrs, tgs, C, focal)
        pb, pp, pl = coarse_boxes()
        _, _, mil = phase1_refine(head, (x_syn,), (x_ori,), metas, syn, pb, pp, pl, gt_boxes, fine, ext,
                                  num_stages=1, num_training_burninstep1=cap, train=True)
        losses = dict(mil or {})
        losses["loss_bbox"], losses["loss_centerness"] = l_box, l_ctr
        return second_half(losses, pb, pp, pl, True)           # phase 1 keeps the coarse boxes (:187)

    def phase2():
        ema()
        pb, pp, pl = coarse_boxes()
        x = student.backbone(img)
        rb, rp, mil = phase2_refine(head, x, metas, pb, pp, pl, gt_boxes, fine, ext, num_stages=1,
                                    num_training_burninstep2=cap, train=True)
        return second_half(dict(mil), rb, rp, pl, False)

    def timed(fn, n, warm):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            last = fn()
        e1.record()
        torch.cuda.synchronize()
        assert torch.isfinite(last)
        return e0.elapsed_time(e1) / n

    W, K = max(args.warmup, 3), max(min(args.steps, 30), 3)
    sampler = ClockSampler(0)
    sampler.start()
    sampler.region(True)
    ms1 = timed(phase1, K, W)
    ms2 = timed(phase2, K, W)
    c1 = _lib.LAUNCHES["count"]
    both = timed(lambda: (phase1(), phase2())[1], K, 1)
    launches = (_lib.LAUNCHES["count"] - c1) // (K + 1)
    sampler.region(False)
    # what the MIL part costs inside these steps: the same two entry points alone, same inputs, with autograd
    x_fix = tuple(t.detach() for t in student.backbone(img))
    pb, pp, pl = coarse_boxes()

    def mil_only():
        opt.zero_grad(set_to_none=True)
        _, _, l = phase2_refine(head, x_fix, metas, pb, pp, pl, gt_boxes, fine, ext, num_stages=1,
                                num_training_burninstep2=cap, train=True)
        t = sum(v for k, v in l.items() if "loss" in k)
        t.backward()
        return t.detach()
    ms_mil = timed(mil_only, K, W)
    hbm_peak, tf_peak, peak_src = peaks
    fc1_flops = 2.0 * (5000 + 5400) * 12544 * 1024 * 3          # forward + dgrad + wgrad of both FC1s
    return {
        "metric": "full teacher-student training step imgs/s (phase-1 iteration + phase-2 iteration)",
        "value": 4e3 / both, "unit": "imgs/s", "n_gpus": 1, "steps": K, "warmup": W, "ms_per_step": both,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "MIL path bf16 (fp32 accumulate); stock backbone / FCOS tower fp32 (TF32 convolutions)",
        "data": "synthetic",
        "config": {"workload": "cfg#2: Point Teacher HBB full teacher-student training step, 2 imgs 800x800, random-init "
                               "R50-FPN + PSAGG-style neck, 200-600 GT/img (MIL capped at 100/img); one step = one "
                               "phase-1 (burn-in: region masking + synthetic-image pass + MIL_head_burn_in_step1) "
                               "iteration + one phase-2 (MIL refinement) iteration, each with EMA teacher, teacher "
                               "forward + coarse pseudo boxes, strong_augmentation, second student pass, dense loss, "
                               "backward and SGD step",
                   "launch": "eager (PyTorch autograd drives the backward)", "l2": "working set (R50 activations of 4-6 "
                   "images at 800x800) far exceeds the 126 MB L2",
                   "ours": "generate_black_paper, FUSETopkAssigner + generate_pseudo_single, phase1_refine / phase2_refine "
                           "(differentiable), strong_augmentation, TopkAssigner x2 + get_target_pseudo_single, FocalLoss",
                   "stock_pytorch": "ResNet-50 + FPN + aggregation neck, FCOS tower, GIoU + centerness BCE, EMA, SGD"},
        "phases": {"phase1_ms": ms1, "phase2_ms": ms2,
                   "mil_phase2_fwd_bwd_alone_ms": ms_mil,
                   "note": "mil_phase2_fwd_bwd_alone_ms = phase2_refine(train=True) + backward on a fixed feature map: the "
                           "share of the step this repo's hot path accounts for; the rest is the stock backbone"},
        "e2e": None, "gpu_launches": launches * K, "clocks": sampler.summary(),
        "roofline": {"kernel": "fc_gemm_kernel (FC1 forward + dgrad + wgrad inside the MIL part of the step)", "bound": "tensor",
                     "achieved": fc1_flops / (ms_mil * 1e-3) / 1e12, "peak": tf_peak, "unit": "TFLOP/s",
                     "frac": fc1_flops / (ms_mil * 1e-3) / 1e12 / tf_peak, "traffic": None, "peak_source": peak_src,
                     "note": "lower bound: FC1 flops over the WHOLE MIL forward+backward time (RoIAlign, FC2, heads, losses "
                             "and their backward included)"},
        "cpu_baseline": None}
