"""The MIL two-stream head of Point Teacher behind the reference's method surface.

Mirrors ``TS_P2BFCOSHead``'s MIL methods (HBB_TOD/mmdet/models/dense_heads/fcos_head_p2b_ts.py):
  _init_layers (MIL part) :212-263, forward_mil :1080-1090, mil_bag_selection :1112-1145,
  mil_bag_training :1147-1180, forward_mil_head :1259-1277, MIL_head_burn_in_step1 :1279-1316,
  MIL_head_burn_in_step2 :1318-1344, inference_mil_head :1346-1390
and the rotated twin ``TS_P2RBRotatedFCOSHead`` (OBB_TOD/mmrotate/models/dense_heads/rotated_fcos_head_p2rb_ts.py
:1186-1453) with identical argument lists, loss keys and parameter names (``shared_fcs_reg.{s}.{0,1}``,
``shared_fcs_bag.{s}.{0,1}``, ``fc_cls.{s}``, ``fc_ins.{s}``, ``fc_reg.{s}``, ``fc_iou.{s}``) so reference
checkpoints load unchanged.  All compute is hand-written sm_100a CUDA (see csrc/): NHWC RoIAlign
writing the bf16 GEMM operand directly, tcgen05/TMA GEMMs for the bag FCs, fused decode / loss /
score / select kernels.

Differentiability.  ``MIL_head_burn_in_step1`` / ``MIL_head_burn_in_step2`` -- the two methods the reference's
detectors call -- return losses that carry a ``grad_fn`` (feature maps + the stage's 14 parameter tensors) whenever
autograd is enabled and something requires a gradient, through ``train._MILStageFn`` (hand-written backward kernels),
so the reference's ``_parse_losses`` -> ``loss.backward()`` flow trains the head.  The finer-grained methods
(``forward_mil_head``, ``mil_bag_training``, ``mil_bag_selection``, ``inference_mil_head``) are forward-only and
REFUSE to run when their detached losses would silently drop a gradient.

``precision``:
  'bf16'   one bf16 tensor-core pass per FC (2e-2 tolerance class)
  'fp32'   bf16x3 split emulation (A_hi*W_hi + A_lo*W_hi + A_hi*W_lo, fp32 accumulate) on the same
           tcgen05 kernel: 3x the MMA work, ~1e-5 agreement with the fp32 reference (forward only)
"""
import torch
import torch.nn as nn

from . import ops
from .proposals import boxes_to_rois, const_tensor, img_wh_tensor
from .registry import HEADS, ROTATED_HEADS, build_roi_extractor
from . import roi_extractors  # noqa: F401  (registers the extractors)


class MILHeadMixin:
    """Mix into a head that defines ``num_classes``, ``in_channels``, ``num_stages``, ``beta``, ``topk`` and the MIL
    layers (``_init_mil_layers`` here, or the reference head's own ``_init_layers``)."""

    bag_loss_pos_scale = 1.0   # OBB scales 0.25 * pos + 0.75 * neg (rotated_fcos_head_p2rb_ts.py:1272,1282)
    bag_loss_neg_scale = 1.0
    bag_loss_bbox_scale = 1.0
    reg_dim = 4
    precision = "bf16"         # class defaults: the mix-in must work on top of the reference head's own __init__
    feat_dtype = None          # None -> fp16 NHWC map for 'bf16' precision, fp32 map for 'fp32'
    roi_feat_area = 49

    # ------------------------------------------------------------------ construction
    def _init_mil_layers(self, roi_feat_size=7, fc_out_channels=1024, std=0.01):
        self.num_shared_fcs = 2
        self.fc_out_channels = fc_out_channels
        self.roi_feat_area = roi_feat_size * roi_feat_size
        self.relu = nn.ReLU(inplace=True)
        self.fc_cls, self.fc_ins = nn.ModuleList(), nn.ModuleList()
        self.fc_reg, self.fc_iou = nn.ModuleList(), nn.ModuleList()
        self.shared_fcs_bag, self.shared_fcs_reg = nn.ModuleList(), nn.ModuleList()

        def branch():
            return nn.ModuleList([nn.Linear(self.in_channels * self.roi_feat_area, fc_out_channels),
                                  nn.Linear(fc_out_channels, fc_out_channels)])
        # constructed-but-unused modules of the reference (kept for state_dict compatibility)
        self.shared_fcs, self.shared_fcs_refine = branch(), branch()
        self.cls_fcs, self.ins_fcs = nn.ModuleList(), nn.ModuleList()
        for _ in range(self.num_stages):
            self.shared_fcs_bag.append(branch())
            self.shared_fcs_reg.append(branch())
            self.fc_cls.append(nn.Linear(fc_out_channels, self.num_classes))
            self.fc_ins.append(nn.Linear(fc_out_channels, self.num_classes))
            self.fc_reg.append(nn.Linear(fc_out_channels, self.reg_dim))
            self.fc_iou.append(nn.Linear(fc_out_channels, 1))
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.normal_(m.weight, 0, std)
                nn.init.constant_(m.bias, 0)
        self._wcache = {}

    # ------------------------------------------------------------------ helpers
    def _x3(self):
        return self.precision == "fp32"

    def _feat_dtype(self):
        if self.feat_dtype is not None:
            return self.feat_dtype
        return torch.float16 if self.precision == "bf16" else torch.float32

    def _dn_hyper(self):
        """``hyper`` of DN_DIoULoss: MILHead stores the number; on top of the reference head it is read from the built
        loss module (fcos_head_p2b_ts.py:207)."""
        h = self.__dict__.get("loss_bbox_denosing_hyper")
        if h is not None:
            return h
        mod = getattr(self, "loss_bbox_denosing", None)
        return float(getattr(mod, "hyper", 0.2))

    def _weights(self):
        if "_wcache" not in self.__dict__:
            self.__dict__["_wcache"] = {}
        return self.__dict__["_wcache"]

    def _weight(self, lin, first):
        """bf16 (optionally [hi|hi|lo]) GEMM operand of a Linear, rebuilt only when the parameter changes."""
        w = lin.weight
        key = (id(lin), first, self._x3())
        tag = (w.data_ptr(), w._version)
        cache = self._weights()
        hit = cache.get(key)
        if hit is None or hit[0] != tag:
            wd = w.detach().contiguous()
            op = ops.prep_fc1_weight(wd, self.in_channels, self.roi_feat_area, self._x3()) if first \
                else ops.cast_weight(wd, self._x3())
            cache[key] = (tag, op)
            hit = cache[key]
        return hit[1]

    def _side_work(self, stage, between=None, stacks=(0, 1), after=None):
        """Fork/join: work that the data path does not need immediately runs on a side stream -- the fp32 -> bf16
        rebuild of both FC stacks' operands when they are stale (training: every step; 2 x 77 MB weight streams that
        overlap bag generation and the first RoIAlign instead of sitting in front of the GEMMs) and ``between()``
        (the negatives' RoIs and weights).  Returns the events (first stack's weights ready, everything done); the
        caller joins with ``wait_event``.  (Measured and NOT kept: starting the first stack's preparation even earlier,
        at the top of ``phase2_refine`` -- the weight stream then competes with the first RoIAlign's stores for HBM:
        0.453 -> 0.460 ms.  The preparation costs ~35 us of the step however it is scheduled -- 0.420 ms with frozen
        operands, ``bench.py --frozen-weights`` -- because the step's first 60 us are HBM-bound either way.)
        Works the same inside a CUDA-graph capture (the side stream is forked
        from, and joined back into, the capturing stream).  (Measured: moving the NCHW -> NHWC transpose here as
        well gains nothing -- it then competes with the weight stream for HBM in front of the first RoIAlign.)"""
        main = torch.cuda.current_stream()
        side = self.__dict__.get("_side_stream")
        if side is None or side.device != main.device:
            side = self.__dict__["_side_stream"] = torch.cuda.Stream(device=main.device)
        side.wait_stream(main)
        evs = []
        with torch.cuda.stream(side):
            for i, fcs in enumerate((self.shared_fcs_reg[stage], self.shared_fcs_bag[stage])):
                if i == 1 and between is not None:
                    between()
                if i in stacks:
                    self._weight(fcs[0], True)
                    self._weight(fcs[1], False)
                if i == 1 and after is not None:
                    after()
                ev = torch.cuda.Event()
                ev.record(side)
                evs.append(ev)
        return evs

    def _fc_stack(self, A, fcs, M, keep=None):
        w1, w2 = self._weight(fcs[0], True), self._weight(fcs[1], False)
        b1, b2 = fcs[0].bias.detach(), fcs[1].bias.detach()
        if not self._x3():
            h1 = ops.fc_gemm(A, w1, b1, relu=True, out_dtype=torch.bfloat16, M=M)
            h2 = ops.fc_gemm(h1, w2, b2, relu=True, out_dtype=torch.bfloat16, M=M)
            if keep is not None:                      # training: the backward needs both activations and operands
                keep.update(A=A, H1=h1, H2=h2, W1=w1, W2=w2, M=M)
            return h2
        if keep is not None:
            raise NotImplementedError("the backward runs in bf16 precision (precision='bf16')")
        h1 = ops.fc_gemm(A, w1, b1, relu=True, out_dtype=torch.float32, M=M)
        return ops.fc_gemm(ops.split_bf16x3(h1), w2, b2, relu=True, out_dtype=torch.float32, M=M)

    def _nhwc_maps(self, x):
        """NHWC maps of the pooled levels (transposed once per step, cached on tensor identity)."""
        ext = self.bbox_roi_extractor
        fd = self._feat_dtype()
        return [ext.roi_layers[i].nhwc(f, fd) for i, f in enumerate(x[:ext.num_inputs])]

    def _roi_operand(self, x, rois, out=None):
        """RoIAlign straight into the FC1 operand layout (bf16, bin-major columns); ``out``: rows of a preallocated
        operand.  Several feature levels (single_level_roi_extractor.py:35-54, 98-104): each RoI is pooled from its
        mapped level -- one launch per level, every launch skipping the other levels' RoIs and writing into the
        shared operand (no ``nonzero()`` synchronisation, no gather / scatter copies).  Returns (operand, levels)."""
        ext = self.bbox_roi_extractor
        maps = self._nhwc_maps(x)
        mode = ops.OUT_BF16X3_BINMAJOR if self._x3() else ops.OUT_BF16_BINMAJOR
        lvls = ops.map_roi_levels(rois, len(maps), ext.finest_scale, rotated=ext.rotated) if len(maps) > 1 else None
        for i, m in enumerate(maps):
            layer = ext.roi_layers[i]
            out = ops.roi_align_forward(m, rois, mode, layer.spatial_scale, layer.sampling_ratio, layer.aligned,
                                        rotated=ext.rotated, clockwise=getattr(layer, "clockwise", True), out=out,
                                        roi_level=lvls, level=i)
        return out, lvls

    def _grad_wanted(self, x):
        if not torch.is_grad_enabled():
            return False
        if any(t.requires_grad for t in x[:self.bbox_roi_extractor.num_inputs]):
            return True
        from .train import stage_params
        return any(p.requires_grad for s in range(len(self.fc_reg)) for p in stage_params(self, s))

    def _refuse_silent_detach(self, x, what):
        if self._grad_wanted(x):
            raise RuntimeError(
                f"{what} returns detached losses; under autograd with parameters that require a gradient this would "
                "silently train nothing.  Call MIL_head_burn_in_step1 / MIL_head_burn_in_step2 (differentiable), or wrap "
                "the call in torch.no_grad().")

    # ------------------------------------------------------------------ reference surface
    def forward_mil(self, feats):
        """:1080-1090 with ``mil_stack_conv = 0`` (both shipped configs): identity."""
        if len(getattr(self, "conv_mil", [])) != 0:
            raise NotImplementedError("mil_stack_conv > 0 is not used by the Point Teacher configs")
        return list(feats)

    def _pack_lists(self, x, img_metas, num_gt_pre_image, proposals_list, proposals_reference_list, proposals_real_list,
                    neg_proposal_list, neg_weight_list):
        """List arguments of the reference surface -> the packed tensors ``mil_stage_packed`` runs on.  The reference
        replicates the pseudo ("reference") and GT ("real") boxes U1 times per GT (syn_images_generator_v2.py:142-144);
        the packed path indexes instance k -> GT k // (U1*U2) instead, so every U1-th row is taken back."""
        dev = x[0].device
        U1 = int(proposals_list[0].shape[0] / num_gt_pre_image[0])      # image 0, like the reference (:1185)
        base_rois = boxes_to_rois([p.float() for p in proposals_list])
        ref = torch.cat(proposals_reference_list)[::max(U1, 1)].float().contiguous()
        real = torch.cat(proposals_real_list)[::max(U1, 1)].float().contiguous()
        negs = neg_idx = neg_w = None
        if neg_proposal_list is not None and sum(int(p.shape[0]) for p in neg_proposal_list) > 0:
            negs = torch.cat(list(neg_proposal_list)).float().contiguous()
            neg_idx = const_tensor([i for i, t in enumerate(neg_proposal_list) for _ in range(int(t.shape[0]))],
                                   torch.int32, dev)
            if neg_weight_list is not None:
                neg_w = torch.cat(list(neg_weight_list)).reshape(-1).to(torch.uint8).contiguous()
            else:   # reference: neg_cls_score is computed but mil_bag_training skips the term (:1169)
                neg_w = torch.zeros((negs.shape[0],), dtype=torch.uint8, device=dev)
        return dict(img_wh=img_wh_tensor(img_metas, dev), base_rois=base_rois, U1=U1, ref=ref, real=real, negs=negs,
                    neg_idx=neg_idx, neg_w=neg_w)

    def _results_dict(self, num_gt, proposals_list, proposals_valid_list, proposals_reference_list,
                      proposals_real_list, losses, stage):
        """``bbox_results`` of the reference (:1182-1277) from what the packed stage left in ``last_results``."""
        L = self.last_results
        b = L["_b200"]
        U1, U2, K = b["U1"], b["U2"], b["K"]
        rs = 6 if self.bbox_roi_extractor.rotated else 5
        sizes = [int(p.shape[0]) * U2 for p in proposals_list]
        R = dict(base_shaking_num=U1, extensive_shaking_num=U2, base_bags=proposals_list,
                 base_bags_valid=proposals_valid_list, iou_target=L["iou_target"],
                 extensive_bags=[r[:, 1:rs] for r in torch.split(b["refined"], sizes)],
                 extensive_bags_valid=[v.bool().reshape(-1, 1) for v in torch.split(b["evalid"], sizes)],
                 extensive_bags_reference=[r.unsqueeze(1).repeat(1, U2, 1).reshape(-1, r.shape[-1])
                                           for r in proposals_reference_list],
                 extensive_bags_real=[r.unsqueeze(1).repeat(1, U2, 1).reshape(-1, r.shape[-1])
                                      for r in proposals_real_list],
                 loss_mil_bbox=losses[f"stage{stage}_loss_mil_bbox"],
                 coarse_bags_iou=losses[f"stage{stage}_coarse_bags_iou"],
                 refine_bags_iou=losses[f"stage{stage}_refine_bags_iou"], _b200=b)
        if L.get("cls_score") is not None:
            R.update(cls_score=L["cls_score"], ins_score=L["ins_score"])
            if b["n_neg"]:
                R["neg_cls_score"] = L["neg_cls_score"]
        return R

    def forward_mil_head(self, num_gt, num_gt_pre_image, x, proposals_list, proposals_valid_list,
                         proposals_reference_list, proposals_real_list, img_metas, fine_proposal_cfg, stage,
                         neg_proposal_list=None, neg_weight_list=None):
        """:1259-1277 (= mil_bag_extensive :1182-1236 + mil_bag_classifier :1240-1256 + negatives); OBB :1285-1382.
        Forward only (``loss_mil_bbox`` is detached)."""
        self._refuse_silent_detach(x, "forward_mil_head")
        p = self._pack_lists(x, img_metas, num_gt_pre_image, proposals_list, proposals_reference_list,
                             proposals_real_list, neg_proposal_list, neg_weight_list)
        with torch.no_grad():
            _, _, losses = self.mil_stage_packed(x, img_metas, p["img_wh"], p["base_rois"], p["U1"], p["ref"], p["real"],
                                                 p["negs"], p["neg_idx"], None, None, None, fine_proposal_cfg, stage,
                                                 neg_w=p["neg_w"], mode="no_select")
        return self._results_dict(num_gt, proposals_list, proposals_valid_list, proposals_reference_list,
                                  proposals_real_list, losses, stage)

    def _score_select(self, R, labels, pseudo, with_loss):
        b = R["_b200"]
        G = labels.shape[0]
        return ops.score_select(b["cls"], b["ins"], b["evalid"], b["refined"], labels, pseudo, b["img_wh"], G,
                                b["U1"], b["U2"], self.topk, self.beta, b["sums"] if with_loss else None,
                                self.bbox_roi_extractor.rotated)

    def mil_bag_training(self, bbox_results, gt_labels, neg_weight_list):
        """:1147-1180 (OBB :1252-1283) -> scalar loss_mil_bags (detached)."""
        b = bbox_results["_b200"]
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.fc_cls.parameters()):
            raise RuntimeError("mil_bag_training returns a detached loss; call MIL_head_burn_in_step1/2 for a "
                               "differentiable loss, or wrap the call in torch.no_grad()")
        labels = torch.cat(gt_labels).long().contiguous()
        b["sums"][5:8] = 0
        self._score_select(bbox_results, labels, None, True)
        has_neg = neg_weight_list is not None and b["n_neg"] > 0
        if has_neg:
            w = torch.cat(list(neg_weight_list)).reshape(-1).to(torch.uint8).contiguous()
            ops.neg_loss(b["cls"][b["K"]:], w, b["sums"])
        out = ops.finalize_losses(b["sums"], b["K"], has_neg, 1.0, 1.0, self.bag_loss_pos_scale,
                                  self.bag_loss_neg_scale)
        return out[1]

    def mil_bag_selection(self, bbox_results, img_metas, pseudo_bboxes, pseudo_labels):
        """:1112-1145 (OBB :1218-1250) -> tuple of merged (G_i, 4|5) boxes."""
        labels = torch.cat(pseudo_labels).long().contiguous()
        pseudo = torch.cat(pseudo_bboxes).float().contiguous()
        merged, pts, idx, sc = self._score_select(bbox_results, labels, pseudo, False)
        bbox_results["_b200"].update(sel_idx=idx, sel_score=sc, merged_points=pts)
        return tuple(torch.split(merged, [len(p) for p in pseudo_bboxes]))

    def _stage_from_lists(self, x, img_metas, proposals_list, proposals_valid_list, proposals_reference_list,
                          proposals_real_list, neg_proposal_list, neg_weight_list, boxes, labels_list, cfg, stage,
                          loss_scales, mode):
        """One stage through the packed path from the reference's list arguments; differentiable when autograd asks
        for it.  ``boxes``: the per-image (G_i, 4|5) pseudo boxes (synthetic boxes for the phase-1 regression pass)."""
        per_img = [int(b.shape[0]) for b in boxes]
        p = self._pack_lists(x, img_metas, per_img, proposals_list, proposals_reference_list, proposals_real_list,
                             neg_proposal_list, neg_weight_list)
        pseudo = torch.cat(list(boxes)).float().contiguous()
        labels = None if labels_list is None else torch.cat(list(labels_list)).long().contiguous()
        args = (img_metas, p["img_wh"], p["base_rois"], p["U1"], p["ref"], p["real"], p["negs"], p["neg_idx"], None,
                labels, pseudo, cfg, stage)
        if self._grad_wanted(x):
            from .train import mil_stage_train
            merged, pts, losses = mil_stage_train(self, x, *args, loss_scales=loss_scales, neg_w=p["neg_w"], mode=mode)
        else:
            with torch.no_grad():
                merged, pts, losses = self.mil_stage_packed(x, *args, loss_scales=loss_scales, neg_w=p["neg_w"],
                                                            mode=mode)
        return merged, pts, losses, per_img

    def MIL_head_burn_in_step2(self, x, img_metas, proposals_list, proposals_valid_list, proposals_reference_list,
                               proposals_real_list, neg_proposal_list, neg_weight_list, pseudo_bboxes,
                               pseudo_labels, fine_proposal_cfg, stage, loss_scales=(1.0, 1.0)):
        """:1318-1344 (OBB :1426-1453), fused: one score/select launch produces the bag loss AND the merged boxes.
        ``stage{s}_loss_mil_bbox`` / ``stage{s}_loss_mil_bags`` carry a ``grad_fn`` under autograd."""
        merged, pts, losses, per_img = self._stage_from_lists(
            x, img_metas, proposals_list, proposals_valid_list, proposals_reference_list, proposals_real_list,
            neg_proposal_list, neg_weight_list, pseudo_bboxes, pseudo_labels, fine_proposal_cfg, stage, loss_scales,
            "full")
        return losses, tuple(torch.split(merged, per_img))

    def MIL_head_burn_in_step1(self, x_ori, x_synethic, img_metas, proposals_list, proposals_valid_list,
                               proposals_reference_list, proposals_real_list, syn_proposals_list,
                               syn_proposals_valid_list, syn_proposals_reference_list, syn_proposals_real_list,
                               neg_proposal_list, neg_weight_list, synthetic_bboxes, pseudo_bboxes, pseudo_labels,
                               fine_proposal_cfg, stage, loss_scales=(1.0, 1.0)):
        """:1279-1316 (OBB :1384-1424): the regression loss comes from the SYNTHETIC image's bags (exact boxes are known
        there), the bag loss, the logged bag IoUs and the selection from the real image.  The reference also runs the
        classifier on the synthetic bags and throws the result away (:1300-1304); that pass is skipped here, and the
        regression branch of the real image runs forward only (its loss is not part of phase 1), so the backward costs
        one regression branch + one bag branch like a phase-2 step."""
        _, _, syn_losses, _ = self._stage_from_lists(
            x_synethic, img_metas, syn_proposals_list, syn_proposals_valid_list, syn_proposals_reference_list,
            syn_proposals_real_list, None, None, synthetic_bboxes, None, fine_proposal_cfg, stage, loss_scales,
            "reg_only")
        merged, pts, losses, per_img = self._stage_from_lists(
            x_ori, img_metas, proposals_list, proposals_valid_list, proposals_reference_list, proposals_real_list,
            neg_proposal_list, neg_weight_list, pseudo_bboxes, pseudo_labels, fine_proposal_cfg, stage, loss_scales,
            "full")
        losses = dict(losses)
        losses[f"stage{stage}_loss_mil_bbox"] = syn_losses[f"stage{stage}_loss_mil_bbox"]
        return losses, tuple(torch.split(merged, per_img))

    def inference_mil_head(self, x, img_metas, proposals_list, proposals_valid_list, proposals_reference_list,
                           proposals_real_list, pseudo_bboxes, pseudo_labels, fine_proposal_cfg, stage):
        """:1346-1390: refinement without losses (no negatives) -> (list of merged (G_i, 4|5) boxes, IoU logs).  No
        detector of the reference calls it; kept so that the head's MIL surface is complete.  Its
        ``fine_proposal_cfg is None`` branch calls ``self.fc_cls(...)`` on a ModuleList in the reference (:1373) and
        cannot run there either."""
        if fine_proposal_cfg is None:
            raise NotImplementedError("inference_mil_head(fine_proposal_cfg=None): the reference branch calls a "
                                      "ModuleList (fcos_head_p2b_ts.py:1373) and never runs")
        with torch.no_grad():
            merged, pts, losses, per_img = self._stage_from_lists(
                x, img_metas, proposals_list, proposals_valid_list, proposals_reference_list, proposals_real_list,
                None, None, pseudo_bboxes, pseudo_labels, fine_proposal_cfg, stage, (1.0, 1.0), "full")
        logs = {f"stage{stage}_coarse_bags_iou": losses[f"stage{stage}_coarse_bags_iou"],
                f"stage{stage}_refine_bags_iou": losses[f"stage{stage}_refine_bags_iou"]}
        return list(torch.split(merged, per_img)), logs

    def mil_stage_packed(self, x, img_metas, img_wh, base_rois, U1, ref, real, neg_boxes, neg_img_idx, bag_offsets,
                         labels, pseudo, cfg, stage, loss_scales=(1.0, 1.0), keep=None, neg_w=None, mode="full"):
        """One MIL stage on packed tensors (the fast path behind ``phase2_refine`` and the list surface): no per-image
        lists, no replicated reference/real boxes (instance k belongs to GT k // (U1*U2)), negatives appended to the
        classification pass.  base_rois (G*U1,5|6); ref/real/pseudo (G,4|5); labels (G,) int64;
        neg_boxes (Nn,4|5)|None with neg_img_idx (Nn,) int32; bag_offsets (B+1,) int32 into base_rois (only needed
        when the negatives' weights are computed here, i.e. ``neg_w is None``).
        mode: 'full'      regression + classification + loss + selection
              'reg_only'  regression branch and its loss only (phase-1 synthetic bags, :1300-1304)
              'no_select' everything but the score/select kernel (the list surface's ``forward_mil_head``)
        Returns (merged (G,4|5), merged centres (G,2), losses dict); merged / centres are None unless mode == 'full'."""
        dev = x[0].device
        rot = self.bbox_roi_extractor.rotated
        rs = 6 if rot else 5
        n_neg = 0 if (neg_boxes is None or mode == "reg_only") else neg_boxes.shape[0]
        U2 = len(cfg["base_ratios"]) ** 2 * (1 + 4 * len(cfg["shake_ratio"] or []))
        K = base_rois.shape[0] * U2
        G = base_rois.shape[0] // max(U1, 1)
        rois2 = torch.empty((K + n_neg, rs), dtype=torch.float32, device=dev)
        side = {"neg_w": neg_w, "lv": None}
        A2 = None
        if mode != "reg_only":
            cols = self.in_channels * self.roi_feat_area * (3 if self._x3() else 1)
            A2 = torch.empty((K + n_neg, cols), dtype=torch.bfloat16, device=dev)

        def negatives():
            if n_neg:
                ops.make_rois(neg_boxes, neg_img_idx, out=rois2[K:])
                if neg_w is None:
                    side["neg_w"] = ops.neg_weight(rois2[K:], base_rois, bag_offsets, rot)

        # (Measured and NOT kept: pooling the negatives' rows of the classification operand on the side stream, under
        # the first FC stack -- they do not depend on the regression pass.  The RoIAlign CTAs (113 / 220 KB of shared
        # memory) cannot share an SM with the persistent GEMM's 200 KB CTAs, so the GEMM lost SMs for its first wave:
        # HBB step 0.460 -> 0.468 ms, OBB 0.586 -> 0.604 ms.  Second attempt: only the rotated negatives' large-RoI GATHER
        # pass -- no shared memory -- on the side stream: its three 224-thread CTAs per SM hold the registers the
        # cooperative FC1 launch needs for co-residency, the GEMM waits for them: OBB 0.576 -> 0.598 ms.)
        ev_reg, ev_all = self._side_work(stage, negatives, stacks=(0,) if mode == "reg_only" else (0, 1))
        neg_w = side["neg_w"]
        ebags, evalid = ops.bag_gen(base_rois, img_wh, cfg["base_ratios"], cfg["shake_ratio"], cfg["min_scale"], rot)
        assert ebags.shape[0] == K
        sums = torch.zeros((8,), dtype=torch.float32, device=dev)
        kreg = {} if keep is not None else None
        kbag = {} if keep is not None else None
        A, lv = self._roi_operand(x, ebags)
        if kreg is not None and lv is not None:
            kreg["lvls"] = lv
        torch.cuda.current_stream().wait_event(ev_reg)
        H = self._fc_stack(A, self.shared_fcs_reg[stage], K, kreg)
        h0, w0, _ = img_metas[0]["img_shape"]            # decode clips to image 0 (reference quirk, :1211)
        fr = self.fc_reg[stage]
        _, deltas, iou_t = ops.reg_decode(H, fr.weight.detach(), fr.bias.detach(), ebags, evalid, ref, real, U1 * U2,
                                          (w0, h0), sums, K=K, hyper=self._dn_hyper(), out_rois=rois2,
                                          rotated=rot, want_deltas=keep is not None)
        del A, H
        cls = ins = merged = pts = idx = sc = None
        if mode != "reg_only":
            torch.cuda.current_stream().wait_event(ev_all)      # the negatives' RoIs + the bag stack's operands
            _, lv = self._roi_operand(x, rois2, out=A2)
            if kbag is not None and lv is not None:
                kbag["lvls"] = lv
            H2 = self._fc_stack(A2, self.shared_fcs_bag[stage], K + n_neg, kbag)
            fc, fi = self.fc_cls[stage], self.fc_ins[stage]
            cls, ins = ops.cls_ins_heads(H2, fc.weight.detach(), fc.bias.detach(), fi.weight.detach(),
                                         fi.bias.detach(), M=K + n_neg)
        else:
            torch.cuda.current_stream().wait_event(ev_all)
        if mode == "full":
            if n_neg:
                with ops.fork() as fneg:                      # beside score_select (both reduce into ``sums``)
                    ops.neg_loss(cls[K:], neg_w, sums)
            merged, pts, idx, sc = ops.score_select(cls, ins, evalid, rois2, labels, pseudo, img_wh, G, U1, U2,
                                                    self.topk, self.beta, sums, rot)
            if n_neg:
                fneg.join()
        out = ops.finalize_losses(sums, K, bool(n_neg) and mode == "full", loss_scales[0], loss_scales[1],
                                  self.bag_loss_pos_scale, self.bag_loss_neg_scale)
        losses = {f"stage{stage}_loss_mil_bbox": out[0], f"stage{stage}_loss_mil_bags": out[1],
                  f"stage{stage}_coarse_bags_iou": out[2], f"stage{stage}_refine_bags_iou": out[3]}
        self.last_losses = losses
        if keep is not None:
            keep.update(reg=kreg, bag=kbag, deltas=deltas, ebags=ebags, evalid=evalid, ref=ref, rois2=rois2, cls=cls,
                        ins=ins, neg_w=neg_w, n_neg=n_neg, labels=labels, sums=sums, K=K, G=G, U1=U1, U2=U2,
                        max_wh=(w0, h0), stage=stage, loss_scales=loss_scales, mode=mode)
        self.last_results = dict(
            cls_score=None if cls is None else cls[:K].view(G, U1, U2, -1),
            ins_score=None if ins is None else ins[:K].view(G, U1, U2, -1),
            neg_cls_score=cls[K:] if (n_neg and cls is not None) else None, neg_weight=neg_w, iou_target=iou_t,
            extensive_bags=[rois2[:K, 1:rs]], base_shaking_num=U1, extensive_shaking_num=U2,
            _b200=dict(sums=sums, K=K, evalid=evalid, refined=rois2[:K], coarse=ebags, img_wh=img_wh, cls=cls,
                       ins=ins, n_neg=n_neg, U1=U1, U2=U2, sel_idx=idx, sel_score=sc, merged_points=pts))
        return merged, pts, losses


class RotatedMILHeadMixin(MILHeadMixin):
    """MIL part of ``TS_P2RBRotatedFCOSHead`` (OBB_TOD/mmrotate/models/dense_heads/rotated_fcos_head_p2rb_ts.py
    :1186-1453): 5-d boxes, RoIAlignRotated, regression on the horizontal (cx,cy,w,h) box with the angle
    carried through (:1314-1334), rotated-IoU logs, bag loss 0.25 * pos + 0.75 * neg (:1272,1282) and the
    top-k score-weighted merge with the (cx,cy) clamp quirk (:1198-1216)."""
    bag_loss_pos_scale = 0.25
    bag_loss_neg_scale = 0.75


class MILHead(nn.Module, MILHeadMixin):
    """Standalone MIL head (the FCOS tower of ``TS_P2BFCOSHead`` is off the hot path).  Constructor keywords
    follow the reference head (fcos_head_p2b_ts.py:80-146); the FCOS-only ones are accepted and ignored.  Registered
    as ``B200MILHead`` and -- when the reference's own class is not in the registry to be wrapped, see
    ``registry.install_reference_heads`` -- as ``TS_P2BFCOSHead``."""

    def __init__(self, num_classes, in_channels=256, beta=0.25, top_k=3, num_stages=2,
                 bbox_roi_extractor=dict(type="SingleRoIExtractor", roi_layer=dict(type="RoIAlign", output_size=7),
                                         out_channels=256, featmap_strides=[8]),
                 loss_bbox_denosing=dict(type="DN_DIoULoss", loss_weight=1.0, hyper=0.2), precision="bf16",
                 feat_dtype=None, mil_stack_conv=0, **kwargs):
        super().__init__()
        self.num_classes, self.in_channels = num_classes, in_channels
        self.beta, self.topk, self.num_stages = beta, top_k, num_stages
        # bf16 precision: fp16 NHWC feature map (TMA + tensor-core RoIAlign; fp16 keeps 3 more mantissa bits than
        # bf16 through the interpolation; values beyond +-65504 are counted and refused, roi_extractors._NHWCCache);
        # fp32 precision: fp32 feature map
        if feat_dtype is None:
            feat_dtype = torch.float16 if precision == "bf16" else torch.float32
        self.precision, self.feat_dtype = precision, feat_dtype
        if loss_bbox_denosing.get("type") != "DN_DIoULoss" or loss_bbox_denosing.get("loss_weight", 1.0) != 1.0:
            raise NotImplementedError("the fused decode kernel implements DN_DIoULoss(loss_weight=1.0)")
        if mil_stack_conv:
            raise NotImplementedError("mil_stack_conv > 0 is not used by the Point Teacher configs")
        self.loss_bbox_denosing_hyper = loss_bbox_denosing.get("hyper", 0.2)
        self.bbox_roi_extractor = build_roi_extractor(bbox_roi_extractor)
        self.conv_mil = nn.ModuleList()
        self._init_mil_layers()


class RotatedMILHead(MILHead, RotatedMILHeadMixin):
    bag_loss_pos_scale = 0.25
    bag_loss_neg_scale = 0.75

    def __init__(self, num_classes, in_channels=256, beta=0.25, top_k=3, num_stages=2,
                 bbox_roi_extractor=dict(type="RotatedSingleRoIExtractor",
                                         roi_layer=dict(type="RoIAlignRotated", out_size=7, sample_num=2,
                                                        clockwise=True),
                                         out_channels=256, featmap_strides=[8]), **kwargs):
        super().__init__(num_classes, in_channels=in_channels, beta=beta, top_k=top_k, num_stages=num_stages,
                         bbox_roi_extractor=bbox_roi_extractor, **kwargs)
        if not self.bbox_roi_extractor.rotated:
            raise ValueError("RotatedMILHead needs a RotatedSingleRoIExtractor")


HEADS.register_module(name="B200MILHead", force=True, module=MILHead)
ROTATED_HEADS.register_module(name="B200RotatedMILHead", force=True, module=RotatedMILHead)

from .registry import install_reference_heads as _install_reference_heads  # noqa: E402

_install_reference_heads()       # 'TS_P2BFCOSHead' / 'TS_P2RBRotatedFCOSHead': wrap the reference class or stand in
