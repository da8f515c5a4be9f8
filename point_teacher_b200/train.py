"""Training step of the MIL head: forward (``MILHeadMixin.mil_stage_packed``) + hand-written backward, exposed as a
``torch.autograd.Function`` so that the reference's ``loss.backward()`` flow keeps working
(HBB_TOD/mmdet/models/detectors/fcos_p2b_teacher_student.py:425-466 returns the loss dict that
``BaseDetector._parse_losses`` sums and back-propagates; HBB_TOD/mmdet/models/dense_heads/fcos_head_p2b_ts.py:1318-1344).

Differentiable outputs: ``stage{s}_loss_mil_bbox`` and ``stage{s}_loss_mil_bags``.  Differentiable inputs: the
feature map (gradient returned in NCHW fp32, through RoIAlign backward) and the 14 parameter tensors of the stage
(``shared_fcs_reg.{s}.{0,1}``, ``shared_fcs_bag.{s}.{0,1}``, ``fc_reg.{s}``, ``fc_cls.{s}``, ``fc_ins.{s}``).  The
refined boxes are detached exactly where the reference detaches them (``pred.clone().detach()`` :1212; scores
``.detach()`` :1126-1127), so no gradient flows through box coordinates.

The big contractions run on the tcgen05 GEMM: dgrad = GEMM against a transposed bf16 weight copy with the ReLU mask
fused in the epilogue; wgrad = GEMM over transposed (row-padded) copies of the activation gradient and the layer
input.  bf16 operands, fp32 accumulation, fp32 parameter gradients."""
import contextlib
import os

import torch

from . import ops


def _early_roi_bwd():
    """PTB200_EARLY_ROI_BWD=0|1 overrides; default: on for a single rank, off when a gradient all-reduce shares the SMs."""
    e = os.environ.get("PTB200_EARLY_ROI_BWD")
    if e is not None:
        return e != "0"
    from .dist import world
    return world() == 1


def stage_params(head, stage):
    fr, fb = head.shared_fcs_reg[stage], head.shared_fcs_bag[stage]
    mods = [fr[0], fr[1], fb[0], fb[1], head.fc_reg[stage], head.fc_cls[stage], head.fc_ins[stage]]
    return [t for m in mods for t in (m.weight, m.bias)]


def _fc_branch_backward(head, k, dZ2, need_dA, out=None, after_wgrad=None):
    """k: what ``_fc_stack`` kept (A, H1, H2, W1 (bin-major bf16), W2 (bf16), M).  dZ2 bf16 [rows, 1024] = gradient
    at the pre-activation of the second FC (already ReLU-masked).  Returns (dW1, db1, dW2, db2, dA | None).
    ``out`` (``dist.MILGradBucket.targets``): the kernels write into the bucket's views; dW1 then STAYS in the
    operand's bin-major column order (``MILGradBucket.finish_`` un-permutes it after the all-reduce); without
    ``out`` fresh tensors are returned and dW1 is in the parameter's column order.  ``after_wgrad()`` is called once the
    branch's last PARAMETER-gradient kernel is enqueued, i.e. before the FC1 dgrad (which only feeds the feature map):
    the gradient all-reduce of the branch starts there and the dgrad runs under it."""
    M, dev = k["M"], dZ2.device
    N1 = k["W1"].shape[0]
    o = out or {}
    # every operand is consumed in the layout the forward left it in (MN-major tcgen05 tiles): no transposed copies
    dZ1 = ops.fc_gemm_mn(dZ2, k["W2"], b_mn=True, mask=k["H1"], M=M)                                # (dZ2 @ W2) * (H1 > 0)
    db2 = o["b2"] if out else torch.zeros((dZ2.shape[1],), dtype=torch.float32, device=dev)
    db1 = o["b1"] if out else torch.zeros((N1,), dtype=torch.float32, device=dev)
    with ops.fork() as f:             # the two bias gradients (10 MB column sums) run beside the wgrad GEMMs
        ops.colsum_bf16(dZ2, db2, M=M)
        ops.colsum_bf16(dZ1, db1, M=M)
    dW2 = ops.fc_gemm_mn(dZ2, k["H1"], a_mn=True, b_mn=True, out_dtype=torch.float32, K=M, out=o.get("W2"))   # dZ2^T @ H1
    dW1p = ops.fc_gemm_mn(dZ1, k["A"], a_mn=True, b_mn=True, out_dtype=torch.float32, K=M, out=o.get("W1p"))  # bin-major
    if out:
        dW1 = dW1p
    else:
        dW1 = torch.empty_like(dW1p)
        ops.unpermute_dw1(dW1p, head.in_channels, head.roi_feat_area, dW1, accumulate=False)
    f.join()
    if after_wgrad is not None:
        after_wgrad()
    dA = None
    if need_dA:
        dA = ops.fc_gemm_mn(dZ1, k["W1"], b_mn=True, out_dtype=torch.bfloat16, M=M)                 # dZ1 @ W1
    return dW1, db1, dW2, db2, dA


def mil_stage_backward(head, keep, x, g_bbox, g_bags, need_feat_grad=True, targets=None, branch_done=None):
    """Gradients of (g_bbox * loss_mil_bbox + g_bags * loss_mil_bags) for one stage.  g_* are 1-element fp32 CUDA
    tensors (no host read); ``None`` skips that branch altogether (phase 1: the synthetic pass has no bag loss and the
    real pass no regression loss, fcos_head_p2b_ts.py:1279-1316).  ``x``: the feature map or the tuple of maps the
    forward pooled from.  Returns (dfeat NCHW fp32 | None -- a list when ``x`` is a tuple --, [14 parameter gradients in
    ``stage_params`` order, ``None`` for the skipped branch]).
    ``targets`` (``MILGradBucket.targets(stage)``): parameter gradients are written in place into the bucket (FC1 weight
    gradients bin-major, small-head / bias slots must have been zeroed); ``branch_done(stage, 'reg'|'bag')`` is
    called the moment a branch's last parameter-gradient kernel is enqueued (the trainer launches that group's
    all-reduce there, so that it runs under the rest of the backward)."""
    feats = list(x) if isinstance(x, (tuple, list)) else [x]
    feats = feats[:head.bbox_roi_extractor.num_inputs]
    stage, dev = keep["stage"], keep["ebags"].device
    K, G, U1, U2, n_neg = keep["K"], keep["G"], keep["U1"], keep["U2"], keep["n_neg"]
    s_bbox, s_bags = keep["loss_scales"]
    fr, fc, fi = head.fc_reg[stage], head.fc_cls[stage], head.fc_ins[stage]
    C = fc.weight.shape[0]
    ext = head.bbox_roi_extractor
    rot = ext.rotated
    do_reg = g_bbox is not None
    do_bag = g_bags is not None and keep["cls"] is not None
    r = b = None
    dWreg = dbreg = dWci = dbci = None
    # ---- regression branch: DN-DIoU -> delta2bbox -> fc_reg -> FC2 -> FC1 -> RoIAlign
    # The bag branch's front end (loss gradient -> fc_cls / fc_ins backward, ~50 us of small HBM-bound kernels) does not
    # depend on the regression branch: it runs on the forked stream under the regression branch's GEMMs.
    fb = None
    if do_bag:
        with (ops.fork() if do_reg else contextlib.nullcontext()) as fb:
            g16 = ops.bag_loss_grad(keep["cls"], keep["ins"], keep["evalid"], keep["labels"], G, U1, U2, keep["neg_w"],
                                    n_neg, keep["sums"], g_bags, s_bags * head.bag_loss_pos_scale,
                                    s_bags * head.bag_loss_neg_scale)
            Wci = torch.cat([fc.weight.detach(), fi.weight.detach()], 0).contiguous()
            tgb = targets["bag"] if targets else None
            dWci, dbci = (tgb["Wh"], tgb["bh"]) if tgb else \
                (torch.zeros_like(Wci), torch.zeros((2 * C,), dtype=torch.float32, device=dev))
            dZ2b = ops.head_bwd(g16, keep["bag"]["H2"], Wci, dWci, dbci, M=K + n_neg)
    if do_reg:
        g4 = ops.reg_loss_grad(keep["deltas"], keep["ebags"], keep["evalid"], keep["ref"], U1 * U2, keep["max_wh"],
                               keep["sums"], g_bbox, s_bbox, hyper=head._dn_hyper(), rotated=rot)
        tg = targets["reg"] if targets else None
        dWreg, dbreg = (tg["Wh"], tg["bh"]) if tg else (torch.zeros_like(fr.weight), torch.zeros_like(fr.bias))
        dZ2 = ops.head_bwd(g4, keep["reg"]["H2"], fr.weight.detach(), dWreg, dbreg, M=K)
        r = _fc_branch_backward(head, keep["reg"], dZ2, need_feat_grad, tg,
                                None if branch_done is None else (lambda: branch_done(stage, "reg")))
    # ---- bag branch: gfocal -> bag score -> (sigmoid, softmax x valid x L1) -> fc_cls / fc_ins -> FC2 -> FC1
    # the regression branch's RoIAlign backward only needs that branch's dA: forked under the bag branch's GEMMs
    # (it fills the SMs their tail waves leave idle)
    early = None
    if do_reg and do_bag and need_feat_grad and len(feats) == 1 and _early_roi_bwd():
        layer = ext.roi_layers[0]
        Bn, Cf, H, W = feats[0].shape
        with ops.fork(lane=2) as fe:     # its own lane: the bag branch's column sums must not queue behind it
            early = ops.roi_align_backward(r[4], keep["ebags"], (Bn, H, W, Cf), layer.spatial_scale, layer.sampling_ratio,
                                           layer.aligned, K=K, roi_level=keep["reg"].get("lvls"), rotated=rot,
                                           clockwise=getattr(layer, "clockwise", True), level=0)
    if do_bag:
        if fb is not None:
            fb.join()
        b = _fc_branch_backward(head, keep["bag"], dZ2b, need_feat_grad, tgb,
                                None if branch_done is None else (lambda: branch_done(stage, "bag")))
    dfeat = None
    if need_feat_grad:
        dfeat = []
        for lvl, f in enumerate(feats):
            layer = ext.roi_layers[lvl]
            Bn, Cf, H, W = f.shape
            shape = (Bn, H, W, Cf)
            rk = dict(rotated=rot, clockwise=getattr(layer, "clockwise", True), level=lvl)
            dn = None
            if early is not None:
                fe.join()
                dn = early
            elif do_reg:
                dn = ops.roi_align_backward(r[4], keep["ebags"], shape, layer.spatial_scale, layer.sampling_ratio,
                                            layer.aligned, K=K, roi_level=keep["reg"].get("lvls"), **rk)
            if do_bag:
                dn = ops.roi_align_backward(b[4], keep["rois2"], shape, layer.spatial_scale, layer.sampling_ratio,
                                            layer.aligned, dfeat=dn, K=K + n_neg, roi_level=keep["bag"].get("lvls"), **rk)
            dfeat.append(None if dn is None else ops.nhwc_to_nchw_f32(dn))
        if not isinstance(x, (tuple, list)):
            dfeat = dfeat[0]
    r4 = list(r[:4]) if do_reg else [None] * 4
    b4 = list(b[:4]) if do_bag else [None] * 4
    grads = r4 + b4 + [dWreg, dbreg] + ([dWci[:C], dbci[:C], dWci[C:], dbci[C:]] if do_bag else [None] * 4)
    return dfeat, grads


class _MILStageFn(torch.autograd.Function):
    """forward(head, args, kwargs, n_feat, *feats, *stage_params) -> (loss_mil_bbox, loss_mil_bags, merged, pts)."""

    @staticmethod
    def forward(ctx, head, args, kwargs, n_feat, *tensors):
        feats = tuple(tensors[:n_feat])
        keep = {}
        with torch.no_grad():
            merged, pts, losses = head.mil_stage_packed(feats, *args, keep=keep, **kwargs)
        ctx.head, ctx.keep, ctx.n_feat = head, keep, n_feat
        ctx.feat_needs = any(f.requires_grad for f in feats)
        ctx.n_params = len(tensors) - n_feat
        ctx.save_for_backward(*feats)
        ctx.set_materialize_grads(False)          # an unused loss arrives as None and its branch is skipped
        s = keep["stage"]
        dev = feats[0].device
        if merged is None:                        # 'reg_only': nothing is selected
            merged, pts = torch.empty((0,), device=dev), torch.empty((0,), device=dev)
        ctx.mark_non_differentiable(merged, pts)
        return losses[f"stage{s}_loss_mil_bbox"].reshape(()), losses[f"stage{s}_loss_mil_bags"].reshape(()), merged, pts

    @staticmethod
    def backward(ctx, g_bbox, g_bags, _gm, _gp):
        feats = ctx.saved_tensors
        one = lambda g: None if g is None else g.detach().float().reshape(1).contiguous()  # noqa: E731
        dfeat, grads = mil_stage_backward(ctx.head, ctx.keep, tuple(feats), one(g_bbox), one(g_bags), ctx.feat_needs)
        ctx.keep = None
        dfeat = dfeat if dfeat is not None else [None] * ctx.n_feat
        return (None, None, None, None, *dfeat, *grads)


def mil_stage_train(head, x, img_metas, img_wh, base_rois, U1, ref, real, neg_boxes, neg_img_idx, bag_offsets, labels,
                    pseudo, cfg, stage, loss_scales=(1.0, 1.0), neg_w=None, mode="full", clear_weights=True):
    """Differentiable twin of ``mil_stage_packed``: same arguments and return value, but the two loss entries carry a
    ``grad_fn`` (feature maps + the stage's parameters)."""
    if getattr(head, "precision", "bf16") != "bf16":
        raise NotImplementedError("the backward runs in bf16 precision (precision='fp32' is forward-only: wrap the "
                                  "call in torch.no_grad())")
    if clear_weights:
        head._weights().clear()                        # parameters change every iteration
    args = (img_metas, img_wh, base_rois, U1, ref, real, neg_boxes, neg_img_idx, bag_offsets, labels, pseudo, cfg, stage)
    kwargs = dict(loss_scales=loss_scales, neg_w=neg_w, mode=mode)
    feats = tuple(x[:head.bbox_roi_extractor.num_inputs])
    lb, lg, merged, pts = _MILStageFn.apply(head, args, kwargs, len(feats), *feats, *stage_params(head, stage))
    losses = dict(head.last_losses)                    # detached logs (bag IoUs) + the two differentiable losses
    losses[f"stage{stage}_loss_mil_bbox"], losses[f"stage{stage}_loss_mil_bags"] = lb, lg
    if mode != "full":
        merged = pts = None
    return merged, pts, losses


class Phase2Trainer:
    """One data-parallel training step of the phase-2 MIL path: forward + backward on this rank's images, then ONE
    flat-bucket all-reduce (average) of the MIL-head gradients over NCCL / NVLink (``dist.MILGradBucket``), and the
    rank-mean of the logged scalars -- the only two exchange steps of the path.  Loss
    denominators stay per rank, as under the reference's DDP."""

    def __init__(self, head, fine_cfg, ext_cfg, num_stages=1, cap=100, alpha=(0.01, 0.25)):
        from .dist import MILGradBucket
        self.head, self.kw = head, dict(fine_proposal_cfg=fine_cfg, fine_proposal_extensive_cfg=ext_cfg,
                                        num_stages=num_stages, num_training_burninstep2=cap, alpha=alpha)
        self.bucket = MILGradBucket(head)

    def step(self, x, img_metas, pseudo_bboxes, pseudo_points, pseudo_labels, gt_bboxes, neg_boxes=None,
             reduce_logs=True, use_autograd=False):
        """x: tuple with the (B,C,H,W) fp32 feature map (``requires_grad`` decides whether its gradient is produced).
        Returns (refined boxes, refined points, losses dict); parameter gradients are left in ``param.grad``
        (already averaged over ranks), the feature gradient in ``x[0].grad``.

        By default the backward kernels are driven directly (every loss has upstream gradient 1, as when
        ``_parse_losses`` sums the dict) -- no autograd engine, hence capturable in a CUDA graph without
        AccumulateGrad stream bookkeeping.  ``use_autograd=True`` goes through ``loss.backward()`` instead."""
        from .dist import reduce_mean_losses
        from .refine import phase2_refine
        head = self.head
        for _, p in self.bucket.named:
            p.grad = None
        if use_autograd:
            boxes, pts, losses = phase2_refine(head, x, img_metas, pseudo_bboxes, pseudo_points, pseudo_labels,
                                               gt_bboxes, neg_boxes=neg_boxes, train=True, **self.kw)
            total = sum(v for k, v in losses.items() if "loss" in k)       # BaseDetector._parse_losses
            total.backward()
        else:
            head._train_keeps = []
            with torch.no_grad():
                boxes, pts, losses = phase2_refine(head, (x[0].detach(),), img_metas, pseudo_bboxes, pseudo_points,
                                                   pseudo_labels, gt_bboxes, neg_boxes=neg_boxes, train="manual",
                                                   **self.kw)
                one = torch.ones((1,), dtype=torch.float32, device=x[0].device)
                need_x = x[0].requires_grad
                dx = None
                self.bucket.zero_small_()
                for keep in head._train_keeps:
                    dfeat, _ = mil_stage_backward(head, keep, x[0], one, one, need_x,
                                                  targets=self.bucket.targets(keep["stage"]),
                                                  branch_done=self.bucket.reduce_group_)
                    if need_x:
                        dx = dfeat if dx is None else dx + dfeat
                head._train_keeps = []
                if need_x:
                    x[0].grad = dx if x[0].grad is None else x[0].grad + dx
                self.bucket.finish_()
            return boxes, pts, (reduce_mean_losses(losses) if reduce_logs else losses)
        self.bucket.all_reduce_()
        return boxes, pts, (reduce_mean_losses(losses) if reduce_logs else losses)


class CapturedTrainStep:
    """``Phase2Trainer.step`` (forward + backward + gradient all-reduce) captured once into a CUDA graph: the
    training step is ~90 launches of mostly short kernels, so eager Python launch overhead is larger than the device
    time.  Static inputs live in ``self.inputs`` (copy new data in, ``replay()``); results in ``self.outputs``
    (boxes, points, losses), parameter gradients in ``param.grad`` and the feature gradient in ``self.x.grad`` --
    all at fixed addresses that every replay overwrites."""

    def __init__(self, head, inputs, img_metas, fine_cfg, ext_cfg, num_stages=1, cap=100, alpha=(0.01, 0.25),
                 warmup=3, feat_grad=True):
        self.trainer = Phase2Trainer(head, fine_cfg, ext_cfg, num_stages, cap, alpha)
        self.inputs, self.img_metas = inputs, img_metas
        self.x = inputs["feat"].detach().requires_grad_(feat_grad)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._step()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.outputs = self._step()

    def _step(self):
        i = self.inputs
        self.x.grad = None
        for layer in self.trainer.head.bbox_roi_extractor.roi_layers:
            layer._cache.clear()
        return self.trainer.step((self.x,), self.img_metas, i["pseudo_boxes"], i["pseudo_points"], i["pseudo_labels"],
                                 i["gt_boxes"], neg_boxes=i.get("neg_boxes"), reduce_logs=False)

    def replay(self):
        for layer in self.trainer.head.bbox_roi_extractor.roi_layers:
            layer._cache.check(captured=True)          # fp16 feature-map range guard (word of earlier replays)
        self.graph.replay()
        return self.outputs
