"""Phase-2 pseudo-box refinement at the detector level.

Mirrors ``TS_P2B_FCOS.forward_mil_head_burn_in_step2``
(HBB_TOD/mmdet/models/detectors/fcos_p2b_teacher_student.py:425-466): cap at
``num_training_burninstep2`` GTs per image, per stage generate base bags + negatives, run the
MIL head, feed the refined boxes to the next stage, write back into full-length clones and
derive the refined points.  ``P2BRefineMixin`` carries the method for a detector that has
``self.student.bbox_head``; :func:`phase2_refine` is the same code for a bare head."""
import torch

from . import ops
from .proposals import (MIL_gen_proposals_from_cfg, const_tensor, gen_negative_proposals, img_wh_tensor,
                        sample_negative_boxes)


def phase2_refine(head, x_ori, img_metas, pseudo_bboxes, pseudo_points, pseudo_labels, gt_bboxes,
                  fine_proposal_cfg, fine_proposal_extensive_cfg, num_stages=1, num_training_burninstep2=100,
                  alpha=(0.01, 0.25), neg_boxes=None, train=False, x_synthetic=None, synthetic_bboxes=None):
    """Returns (refined_pseudo_bboxes, refined_pseudo_points, losses) like the reference method.
    ``neg_boxes[stage][img]`` optionally injects the negative boxes the reference samples on the CPU.

    Everything between the input lists and the output lists runs on packed tensors: one ``torch.cat`` per
    input list, per-image bookkeeping as cached constant index tensors, ~20 kernel launches per stage and
    no host synchronisation, so the whole call is CUDA-graph capturable (``CapturedPhase2``).

    ``train``: False -> forward only; True -> the MIL losses carry a ``grad_fn`` (train.mil_stage_train);
    'manual' -> forward only with the intermediates kept for an explicit backward (train.Phase2Trainer).
    ``x_synthetic`` + ``synthetic_bboxes``: the phase-1 variant (:func:`phase1_refine`)."""
    cap = num_training_burninstep2
    dev = pseudo_bboxes[0].device
    rot = head.bbox_roi_extractor.rotated       # OBB twin: rotated_fcos_teacher_student.py:494-535
    if train == "manual" or train:
        head._weights().clear()                 # parameters change every iteration
    counts = [min(int(b.shape[0]), cap) for b in pseudo_bboxes]
    pb = torch.cat([b[:cap, :] for b in pseudo_bboxes]).float().contiguous()
    img_idx = const_tensor([i for i, c in enumerate(counts) for _ in range(c)], torch.int32, dev)
    img_wh = img_wh_tensor(img_metas, dev)
    phase1 = synthetic_bboxes is not None
    if phase1:
        s_counts = [min(int(b.shape[0]), cap) for b in synthetic_bboxes]
        sb = torch.cat([b[:cap, :] for b in synthetic_bboxes]).float().contiguous()
        s_idx = const_tensor([i for i, c in enumerate(s_counts) for _ in range(c)], torch.int32, dev)
    forks = []
    # the packing of everything the first bag generation does not read (GT boxes, labels, the first stage's negatives)
    # and the logged scalars run beside the data path; joined right before the first stage needs them
    nl0 = neg_boxes[0] if (neg_boxes is not None and fine_proposal_cfg[0]["gen_num_neg"]) else None
    negs0 = None
    with ops.fork() as f0:
        gb = torch.cat([b[:cap, :] for b in gt_bboxes]).float().contiguous()
        labels = torch.cat([l[:cap] for l in pseudo_labels]).long().contiguous()
        if nl0 is not None:
            negs0 = torch.cat(list(nl0)).float().contiguous()
        losses = {"coarse_bboxes_iou": ops.aligned_iou_mean(pb, gb, rot)}
    pts = None

    def run_stage(x, *args, **kw):
        if train == "manual":     # forward only, intermediates kept for an explicit backward (train.Phase2Trainer)
            keep = {}
            out = head.mil_stage_packed(x, *args, keep=keep, **kw)
            head._train_keeps = getattr(head, "_train_keeps", []) + [keep]
            return out
        if train:                 # the two MIL losses carry a grad_fn (feature map + head parameters), see train.py
            from .train import mil_stage_train
            return mil_stage_train(head, x, *args, clear_weights=False, **kw)
        return head.mil_stage_packed(x, *args, **kw)

    for stage in range(num_stages):
        cfg = fine_proposal_cfg[stage]
        base_rois, _ = ops.bag_gen(ops.make_rois(pb, img_idx), img_wh, cfg["base_ratios"], cfg["shake_ratio"],
                                   cfg["min_scale"], rot)
        U1 = base_rois.shape[0] // max(pb.shape[0], 1)
        negs = neg_idx = offs = None
        n_neg = cfg["gen_num_neg"]
        if n_neg:
            nl = neg_boxes[stage] if neg_boxes is not None else \
                [sample_negative_boxes(n_neg, m["img_shape"], rotated=rot).to(dev) for m in img_metas]
            negs = negs0 if (stage == 0 and negs0 is not None) else torch.cat(list(nl)).float().contiguous()
            neg_idx = const_tensor([i for i, t in enumerate(nl) for _ in range(int(t.shape[0]))], torch.int32, dev)
            o = [0]
            for c in counts:
                o.append(o[-1] + c * U1)
            offs = const_tensor(o, torch.int32, dev)
        syn_loss = None
        if phase1:   # regression loss from the synthetic image's bags (fcos_p2b_teacher_student.py:395-397, head :1300-1304)
            s_rois, _ = ops.bag_gen(ops.make_rois(sb, s_idx), img_wh, cfg["base_ratios"], cfg["shake_ratio"],
                                    cfg["min_scale"], rot)
            _, _, syn_loss = run_stage(x_synthetic, img_metas, img_wh, s_rois, U1, sb, sb, None, None, None, None, None,
                                       fine_proposal_extensive_cfg[stage], stage, loss_scales=alpha, mode="reg_only")
        if stage == 0:
            f0.join()
        pb_new, pts, mil_loss = run_stage(x_ori, img_metas, img_wh, base_rois, U1, pb, gb, negs, neg_idx, offs, labels, pb,
                                          fine_proposal_extensive_cfg[stage], stage, loss_scales=alpha)
        if syn_loss is not None:
            mil_loss = dict(mil_loss)
            mil_loss[f"stage{stage}_loss_mil_bbox"] = syn_loss[f"stage{stage}_loss_mil_bbox"]
        pb = pb_new
        with ops.fork() as f:
            losses[f"stage{stage}_refine_bboxes_iou"] = ops.aligned_iou_mean(pb, gb, rot)
        forks.append(f)
        losses.update(mil_loss)
    # write-back (:463-465): refined head + untouched tail, one concat for boxes and one for points
    mb, mp = torch.split(pb, counts), torch.split(pts, counts)
    sizes = [int(b.shape[0]) for b in pseudo_bboxes]
    box_parts, pt_parts = [], []
    for i in range(len(pseudo_bboxes)):
        box_parts += [mb[i].to(pseudo_bboxes[i].dtype), pseudo_bboxes[i][cap:, :]]
        pt_parts += [mp[i].to(pseudo_points[i].dtype), pseudo_points[i][cap:, :]]
    with ops.fork() as f:
        refined_p = list(torch.split(torch.cat(pt_parts), sizes))
    forks.append(f)
    refined_b = list(torch.split(torch.cat(box_parts), sizes))
    for f in forks:
        f.join()
    return refined_b, refined_p, losses


def phase1_refine(head, x_synthetic, x_ori, img_metas, synthetic_bboxes, pseudo_bboxes, pseudo_points, pseudo_labels,
                  gt_bboxes, fine_proposal_cfg, fine_proposal_extensive_cfg, num_stages=1, num_training_burninstep1=100,
                  alpha=(0.01, 0.25), neg_boxes=None, train=False):
    """Phase-1 (burn-in) MIL training step at the detector level: ``forward_mil_head_burn_in_step1``
    (HBB_TOD/mmdet/models/detectors/fcos_p2b_teacher_student.py:365-424; OBB rotated_fcos_teacher_student.py:435-492).
    The regression loss is taken on the bags around the SYNTHETIC boxes pooled from the synthetic image's features,
    the bag loss / logs / selection on the real image.  Returns the same triple as the reference; when an image has
    no synthetic box the reference returns empty lists and ``None`` (:369-373) -- so does this."""
    if any(int(b.shape[0]) == 0 for b in synthetic_bboxes):
        d = 5 if head.bbox_roi_extractor.rotated else 4
        e = pseudo_bboxes[0]
        return [e.new_empty((0, d)) for _ in pseudo_bboxes], [e.new_empty((0, 2)) for _ in pseudo_bboxes], None
    return phase2_refine(head, x_ori, img_metas, pseudo_bboxes, pseudo_points, pseudo_labels, gt_bboxes,
                         fine_proposal_cfg, fine_proposal_extensive_cfg, num_stages, num_training_burninstep1, alpha,
                         neg_boxes, train, x_synthetic=x_synthetic, synthetic_bboxes=synthetic_bboxes)


def phase2_refine_lists(head, x_ori, img_metas, pseudo_bboxes, pseudo_points, pseudo_labels, gt_bboxes,
                        fine_proposal_cfg, fine_proposal_extensive_cfg, num_stages=1, num_training_burninstep2=100,
                        alpha=(0.01, 0.25), neg_boxes=None):
    """The same step spelled exactly like the reference method, through the list-based module functions
    (``MIL_gen_proposals_from_cfg`` / ``gen_negative_proposals`` / ``head.MIL_head_burn_in_step2``).  Slower
    (dozens of tiny glue launches); kept as the literal drop-in and cross-checked against the packed path."""
    cap = num_training_burninstep2
    pb = [b[:cap, :].clone().float() for b in pseudo_bboxes]
    gb = [b[:cap, :].clone().float() for b in gt_bboxes]
    pp = [p[:cap, :].clone().float() for p in pseudo_points]
    pl = [l[:cap].clone() for l in pseudo_labels]
    refined_b = [b.clone() for b in pseudo_bboxes]
    refined_p = [p.clone() for p in pseudo_points]
    gcat = torch.cat(gb).contiguous()
    losses = {"coarse_bboxes_iou": ops.aligned_iou_mean(torch.cat(pb).contiguous(), gcat)}
    for stage in range(num_stages):
        props, valids, refs, reals = MIL_gen_proposals_from_cfg(pp, pb, fine_proposal_cfg[stage], gb, img_metas)
        negs, neg_w = gen_negative_proposals(pp, fine_proposal_cfg[stage], props, img_metas,
                                             None if neg_boxes is None else neg_boxes[stage])
        mil_loss, merged = head.MIL_head_burn_in_step2(x_ori, img_metas, props, valids, refs, reals, negs, neg_w,
                                                       pb, pl, fine_proposal_extensive_cfg[stage], stage,
                                                       loss_scales=alpha)
        pb = list(merged)
        losses[f"stage{stage}_refine_bboxes_iou"] = ops.aligned_iou_mean(torch.cat(pb).contiguous(), gcat)
        losses.update(mil_loss)
    pts = head.last_results["_b200"]["merged_points"]
    for i, (b, p) in enumerate(zip(pb, torch.split(pts, [len(b) for b in pb]))):
        refined_b[i][:cap, :] = b
        refined_p[i][:cap, :] = p
    return refined_b, refined_p, losses


class P2BRefineMixin:
    """Drop-ins for the detector methods (same names and argument lists; HBB ``TS_P2B_FCOS`` and, without the trailing
    ``img`` argument, OBB ``RotatedFCOS_TS``).  Differentiable whenever autograd is on and something requires a
    gradient, like the reference's own methods."""

    def _mil_train_flag(self, x):
        return self.student.bbox_head._grad_wanted(x)

    def forward_mil_head_burn_in_step2(self, num_img, pseudo_bboxes, pseudo_points, pseudo_labels, gt_bboxes,
                                       img_metas, x_ori):
        return phase2_refine(self.student.bbox_head, x_ori, img_metas, pseudo_bboxes, pseudo_points,
                             pseudo_labels, gt_bboxes, self.fine_proposal_cfg, self.fine_proposal_extensive_cfg,
                             self.num_stages, self.num_training_burninstep2, self.alpha,
                             train=self._mil_train_flag(x_ori))

    def forward_mil_head_burn_in_step1(self, num_img, synthetic_bboxes, pseudo_bboxes, pseudo_points, pseudo_labels,
                                       gt_bboxes, img_metas, x_synthetic, x_ori, img=None):
        return phase1_refine(self.student.bbox_head, x_synthetic, x_ori, img_metas, synthetic_bboxes, pseudo_bboxes,
                             pseudo_points, pseudo_labels, gt_bboxes, self.fine_proposal_cfg,
                             self.fine_proposal_extensive_cfg, self.num_stages, self.num_training_burninstep1,
                             self.alpha, train=self._mil_train_flag(x_ori) or self._mil_train_flag(x_synthetic))


class CapturedPhase2:
    """The whole phase-2 refinement step captured once into a CUDA graph and replayed (static shapes).

    At the shipped config the path is latency-bound (~25 launches, <0.5 ms of device work), so host
    launch overhead is removed by graph replay instead of a tracing compiler.  The inputs live in
    static device buffers (``self.inputs``); copy new data into them, ``replay()``, read ``self.outputs``.
    ``refresh_weights=True`` keeps the fp32 -> bf16 (+ FC1 column permutation) weight preparation inside
    the captured step, which is what a training loop (weights change every iteration) needs."""

    def __init__(self, head, inputs, img_metas, fine_cfg, ext_cfg, num_stages=1, cap=100, alpha=(0.01, 0.25),
                 refresh_weights=True, warmup=3, gather_outputs=False):
        self.head, self.inputs, self.img_metas = head, inputs, img_metas
        self.kw = dict(fine_proposal_cfg=fine_cfg, fine_proposal_extensive_cfg=ext_cfg, num_stages=num_stages,
                       num_training_burninstep2=cap, alpha=alpha)
        self.refresh_weights = refresh_weights
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s), torch.no_grad():
            for _ in range(warmup):
                self._step()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.outputs = self._step()
            self.loss_keys = sorted(self.outputs[2])
            self.loss_vec = torch.stack([self.outputs[2][k].reshape(()).float() for k in self.loss_keys])
            self.out_arena = None
            if gather_outputs:      # the step's results gathered into ONE device buffer by the graph's last nodes
                tree = dict(boxes=list(self.outputs[0]), points=list(self.outputs[1]), losses=self.loss_vec)
                self.out_arena = _Arena(tree, self.loss_vec.device)
                self.out_arena.fill(tree)

    def _step(self):
        i = self.inputs
        if self.refresh_weights:
            self.head._weights().clear()
        # the NHWC feature-map cache is keyed on (pointer, version): inside a captured step it must MISS, otherwise
        # the capture would bake in the tensor transposed during warm-up and replays would read stale features
        for layer in self.head.bbox_roi_extractor.roi_layers:
            layer._cache.clear()
        return phase2_refine(self.head, (i["feat"],), self.img_metas, i["pseudo_boxes"], i["pseudo_points"],
                             i["pseudo_labels"], i["gt_boxes"], neg_boxes=i.get("neg_boxes"), **self.kw)

    def replay(self):
        # fp16 feature-map range guard: the word written by EARLIER replays (no synchronisation here)
        for layer in self.head.bbox_roi_extractor.roi_layers:
            layer._cache.check(captured=True)
        self.graph.replay()
        return self.outputs


class _Arena:
    """ONE contiguous byte buffer (device, or pinned host) carrying a nested structure of tensors as 256-byte aligned
    typed views, so that the whole structure crosses PCIe as a single ``cudaMemcpyAsync``."""

    def __init__(self, example, device=None):
        self.spec, off = [], 0

        def walk(v):
            nonlocal off
            if v is None:
                return None
            if isinstance(v, dict):
                return {k: walk(v[k]) for k in sorted(v)}
            if isinstance(v, (list, tuple)):
                return [walk(x) for x in v]
            off = (off + 255) // 256 * 256
            rec = (off, tuple(v.shape), v.dtype)
            off += v.numel() * v.element_size()
            self.spec.append(rec)
            return rec
        tree = walk(example)
        self.nbytes = (off + 255) // 256 * 256
        if device is None:
            self.buf = torch.empty((self.nbytes,), dtype=torch.uint8)
            if torch.cuda.is_available():
                self.buf = self.buf.pin_memory()
        else:
            self.buf = torch.empty((self.nbytes,), dtype=torch.uint8, device=device)
        self.payload = sum(int(torch.tensor([], dtype=d).element_size()) * int(torch.Size(sh).numel()) for _, sh, d in self.spec)

        def build(t):
            if t is None:
                return None
            if isinstance(t, dict):
                return {k: build(x) for k, x in t.items()}
            if isinstance(t, list):
                return [build(x) for x in t]
            o, shape, dtype = t
            n = int(torch.Size(shape).numel()) * torch.tensor([], dtype=dtype).element_size()
            return self.buf[o:o + n].view(dtype).view(shape)
        self.views = build(tree)

    @staticmethod
    def flat(v):
        if v is None:
            return []
        if isinstance(v, dict):
            return [t for k in sorted(v) for t in _Arena.flat(v[k])]
        if isinstance(v, (list, tuple)):
            return [t for x in v for t in _Arena.flat(x)]
        return [v]

    def fill(self, values):
        for dst, src in zip(self.flat(self.views), self.flat(values)):
            dst.copy_(src)


class Phase2Pipeline:
    """Host-facing throughput API: phase-2 refinement of a stream of batches whose inputs live in HOST memory.

    ``depth`` captured steps (``CapturedPhase2``) are used round-robin.  Every slot owns ONE pinned host staging
    buffer and ONE device buffer for its inputs (and one pair for its results) with identical layouts
    (``_Arena``): a step is exactly one ``cudaMemcpyAsync`` host -> device on a copy stream (overlapping the previous
    step's compute), one graph replay (whose last nodes gather boxes / points / losses into the result buffer) and
    one ``cudaMemcpyAsync`` device -> host.  The producer writes the next batch straight into
    ``host_inputs(slot)`` (views of the pinned staging buffer); ``submit(values)`` with loose host tensors is also
    accepted and first copies them into the staging buffer on the host.
    ``submit`` never blocks the host; ``result(ticket)`` waits for that batch.

    example_inputs: dict(feat (B,C,H,W) fp32, pseudo_boxes, pseudo_points, pseudo_labels, gt_boxes: lists of
    per-image tensors, neg_boxes: [per-stage][per-image] injected negatives or None) on the device."""

    def __init__(self, head, example_inputs, img_metas, fine_cfg, ext_cfg, num_stages=1, cap=100,
                 alpha=(0.01, 0.25), depth=3, refresh_weights=True):
        dev = example_inputs["feat"].device
        self.depth = depth
        self.dev_in = [_Arena(example_inputs, dev) for _ in range(depth)]
        self.host_in = [_Arena(example_inputs) for _ in range(depth)]
        for a in self.dev_in:
            a.fill(example_inputs)
        for a in self.host_in:
            a.fill(example_inputs)
        self.slots, self.dev_out, self.host_out = [], [], []
        for s in range(depth):
            c = CapturedPhase2(head, self.dev_in[s].views, img_metas, fine_cfg, ext_cfg, num_stages, cap, alpha,
                               refresh_weights, gather_outputs=True)
            self.slots.append(c)
            self.dev_out.append(c.out_arena)
            self.host_out.append(_Arena(c.out_arena.views))
        self.copy_stream = torch.cuda.Stream()
        self.h2d_done = [torch.cuda.Event() for _ in range(depth)]
        self.done = [torch.cuda.Event() for _ in range(depth)]
        self.loss_keys = self.slots[0].loss_keys
        self.n = 0
        self.h2d_bytes = self.host_in[0].payload          # what the step consumes (the copy moves nbytes incl. padding)
        self.d2h_bytes = self.host_out[0].payload
        self.h2d_copies_per_step = self.d2h_copies_per_step = 1

    def host_inputs(self, slot=None):
        """Pinned staging views of the slot the NEXT ``submit()`` will use: fill them, then ``submit()``."""
        return self.host_in[self.n % self.depth if slot is None else slot].views

    def submit(self, host_inputs=None):
        s = self.n % self.depth
        slot, cur = self.slots[s], torch.cuda.current_stream()
        if host_inputs is not None:
            self.done[s].synchronize()                         # the staging buffer's previous H2D has long finished
            self.host_in[s].fill(host_inputs)
        self.copy_stream.wait_event(self.done[s])              # slot inputs are free once its last step finished
        with torch.cuda.stream(self.copy_stream):
            self.dev_in[s].buf.copy_(self.host_in[s].buf, non_blocking=True)       # the ONE H2D of this step
            self.h2d_done[s].record(self.copy_stream)
        cur.wait_event(self.h2d_done[s])
        slot.replay()
        self.host_out[s].buf.copy_(self.dev_out[s].buf, non_blocking=True)         # the ONE D2H of this step
        self.done[s].record(cur)
        self.n += 1
        return s

    def result(self, ticket):
        self.done[ticket].synchronize()
        ho = self.host_out[ticket].views
        return ho["boxes"], ho["points"], dict(zip(self.loss_keys, ho["losses"].tolist()))
