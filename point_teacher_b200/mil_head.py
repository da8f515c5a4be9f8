"""The MIL two-stream head of Point Teacher behind the reference's method surface.

Mirrors ``TS_P2BFCOSHead``'s MIL methods (HBB_TOD/mmdet/models/dense_heads/fcos_head_p2b_ts.py):
  _init_layers (MIL part) :212-263, forward_mil :1080-1090, mil_bag_selection :1112-1145,
  mil_bag_training :1147-1180, forward_mil_head :1259-1277, MIL_head_burn_in_step1 :1279-1316,
  MIL_head_burn_in_step2 :1318-1344, inference_mil_head :1346-1390
with identical argument lists, loss keys and parameter names (``shared_fcs_reg.{s}.{0,1}``,
``shared_fcs_bag.{s}.{0,1}``, ``fc_cls.{s}``, ``fc_ins.{s}``, ``fc_reg.{s}``, ``fc_iou.{s}``) so reference
checkpoints load unchanged.  All compute is hand-written sm_100a CUDA (see csrc/): NHWC RoIAlign
writing the bf16 GEMM operand directly, tcgen05/TMA GEMMs for the bag FCs, fused decode / loss /
score / select kernels.  Forward only in this round (the losses are returned detached).

``precision``:
  'bf16'   one bf16 tensor-core pass per FC (2e-2 tolerance class)
  'fp32'   bf16x3 split emulation (A_hi*W_hi + A_lo*W_hi + A_hi*W_lo, fp32 accumulate) on the same
           tcgen05 kernel: 3x the MMA work, ~1e-5 agreement with the fp32 reference
"""
import torch
import torch.nn as nn

from . import ops
from .proposals import boxes_to_rois, img_wh_tensor
from .registry import HEADS, build_roi_extractor
from . import roi_extractors  # noqa: F401  (registers the extractors)


class MILHeadMixin:
    """Mix into a head that defines ``num_classes``, ``in_channels``, ``num_stages``, ``beta``, ``topk``."""

    bag_loss_pos_scale = 1.0   # OBB scales 0.25 * pos + 0.75 * neg (rotated_fcos_head_p2rb_ts.py:1272,1282)
    bag_loss_neg_scale = 1.0
    bag_loss_bbox_scale = 1.0
    reg_dim = 4

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
        return getattr(self, "precision", "bf16") == "fp32"

    def _weight(self, lin, first):
        """bf16 (optionally [hi|hi|lo]) GEMM operand of a Linear, rebuilt only when the parameter changes."""
        w = lin.weight
        key = (id(lin), first, self._x3())
        tag = (w.data_ptr(), w._version)
        hit = self._wcache.get(key)
        if hit is None or hit[0] != tag:
            wd = w.detach().contiguous()
            op = ops.prep_fc1_weight(wd, self.in_channels, self.roi_feat_area, self._x3()) if first \
                else ops.cast_weight(wd, self._x3())
            self._wcache[key] = (tag, op)
            hit = self._wcache[key]
        return hit[1]

    def _side_work(self, stage, between=None):
        """Fork/join: work that the data path does not need immediately runs on a side stream -- the fp32 -> bf16
        rebuild of both FC stacks' operands when they are stale (training: every step; 2 x 77 MB weight streams that
        overlap bag generation and the first RoIAlign instead of sitting in front of the GEMMs) and ``between()``
        (the negatives' RoIs and weights).  Returns the events (first stack's weights ready, everything done); the
        caller joins with ``wait_event``.  Works the same inside a CUDA-graph capture (the side stream is forked
        from, and joined back into, the capturing stream).  (Measured: moving the NCHW -> NHWC transpose here as
        well gains nothing -- it then competes with the weight stream for HBM in front of the first RoIAlign.)"""
        main = torch.cuda.current_stream()
        side = getattr(self, "_side_stream", None)
        if side is None or side.device != main.device:
            side = self._side_stream = torch.cuda.Stream(device=main.device)
        side.wait_stream(main)
        evs = []
        with torch.cuda.stream(side):
            for i, fcs in enumerate((self.shared_fcs_reg[stage], self.shared_fcs_bag[stage])):
                if i == 1 and between is not None:
                    between()
                self._weight(fcs[0], True)
                self._weight(fcs[1], False)
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

    def _roi_operand(self, x, rois):
        """RoIAlign straight into the FC1 operand layout (bf16, bin-major columns)."""
        ext = self.bbox_roi_extractor
        if len(x[:ext.num_inputs]) != 1:
            raise NotImplementedError("the fused MIL path runs on a single feature level (both shipped configs); "
                                      "use bbox_roi_extractor(...) for multi-level extraction")
        layer = ext.roi_layers[0]
        feat = layer.nhwc(x[0], getattr(self, "feat_dtype", torch.float32))
        mode = ops.OUT_BF16X3_BINMAJOR if self._x3() else ops.OUT_BF16_BINMAJOR
        return ops.roi_align_forward(feat, rois, mode, layer.spatial_scale, layer.sampling_ratio, layer.aligned,
                                     rotated=ext.rotated, clockwise=getattr(layer, "clockwise", True))

    # ------------------------------------------------------------------ reference surface
    def forward_mil(self, feats):
        """:1080-1090 with ``mil_stack_conv = 0`` (both shipped configs): identity."""
        if len(getattr(self, "conv_mil", [])) != 0:
            raise NotImplementedError("mil_stack_conv > 0 is not used by the Point Teacher configs")
        return list(feats)

    def forward_mil_head(self, num_gt, num_gt_pre_image, x, proposals_list, proposals_valid_list,
                         proposals_reference_list, proposals_real_list, img_metas, fine_proposal_cfg, stage,
                         neg_proposal_list=None, neg_weight_list=None):
        """:1259-1277 (= mil_bag_extensive :1182-1236 + mil_bag_classifier :1240-1256 + negatives)."""
        if self.bbox_roi_extractor.rotated:
            raise NotImplementedError("the list-based MIL methods are the HBB surface (4-d boxes); the rotated head runs "
                                      "through refine.phase2_refine / mil_stage_packed (5-d boxes, rotated bags)")
        dev = x[0].device
        R = {}
        U1 = int(proposals_list[0].shape[0] / num_gt_pre_image[0])
        base_rois = boxes_to_rois([p.float() for p in proposals_list])
        img_wh = img_wh_tensor(img_metas, dev)
        ebags, evalid = ops.bag_gen(base_rois, img_wh, fine_proposal_cfg["base_ratios"],
                                    fine_proposal_cfg["shake_ratio"], fine_proposal_cfg["min_scale"])
        K = ebags.shape[0]
        U2 = int((K // max(base_rois.shape[0], 1)))
        ref = torch.cat(proposals_reference_list).float().contiguous()
        real = torch.cat(proposals_real_list).float().contiguous()
        sums = torch.zeros((8,), dtype=torch.float32, device=dev)
        # --- regression branch
        A = self._roi_operand(x, ebags)
        H = self._fc_stack(A, self.shared_fcs_reg[stage], K)
        h0, w0, _ = img_metas[0]["img_shape"]           # decode clips to image 0 (reference quirk, :1211)
        n_neg = 0 if neg_proposal_list is None else sum(p.shape[0] for p in neg_proposal_list)
        rois2 = torch.empty((K + n_neg, 5), dtype=torch.float32, device=dev)
        if n_neg:
            rois2[K:] = boxes_to_rois([p.float() for p in neg_proposal_list])
        fr = self.fc_reg[stage]
        _, deltas, iou_t = ops.reg_decode(H, fr.weight.detach(), fr.bias.detach(), ebags, evalid, ref, real, U2,
                                          (w0, h0), sums, K=K, hyper=self.loss_bbox_denosing_hyper,
                                          want_deltas=True, out_rois=rois2)
        del A, H
        # --- classification branch on the refined bags (+ negatives through the same GEMMs)
        A2 = self._roi_operand(x, rois2)
        H2 = self._fc_stack(A2, self.shared_fcs_bag[stage], K + n_neg)
        fc, fi = self.fc_cls[stage], self.fc_ins[stage]
        cls, ins = ops.cls_ins_heads(H2, fc.weight.detach(), fc.bias.detach(), fi.weight.detach(),
                                     fi.bias.detach(), M=K + n_neg)
        sizes = [p.shape[0] * U2 for p in proposals_list]
        refined = rois2[:K]
        R.update(base_shaking_num=U1, extensive_shaking_num=U2, base_bags=proposals_list,
                 base_bags_valid=proposals_valid_list, iou_target=iou_t, bbox_deltas=deltas,
                 extensive_bags=[r[:, 1:5] for r in torch.split(refined, sizes)],
                 extensive_bags_valid=[v.bool().reshape(-1, 1) for v in torch.split(evalid, sizes)],
                 extensive_bags_reference=[r.unsqueeze(1).repeat(1, U2, 1).reshape(-1, 4)
                                           for r in proposals_reference_list],
                 extensive_bags_real=[r.unsqueeze(1).repeat(1, U2, 1).reshape(-1, 4) for r in proposals_real_list],
                 cls_score=cls[:K].view(num_gt, U1, U2, -1), ins_score=ins[:K].view(num_gt, U1, U2, -1))
        if n_neg:
            R["neg_cls_score"] = cls[K:]
        R["_b200"] = dict(sums=sums, K=K, evalid=evalid, refined=refined, coarse=ebags, img_wh=img_wh,
                          cls=cls, ins=ins, n_neg=n_neg, U1=U1, U2=U2)
        part = ops.finalize_losses(sums, K, False)
        R["loss_mil_bbox"], R["coarse_bags_iou"], R["refine_bags_iou"] = part[0], part[2], part[3]
        return R

    def _score_select(self, R, labels, pseudo, with_loss):
        b = R["_b200"]
        K, G = b["K"], labels.shape[0]
        return ops.score_select(b["cls"], b["ins"], b["evalid"], b["refined"], labels, pseudo, b["img_wh"], G,
                                b["U1"], b["U2"], self.topk, self.beta, b["sums"] if with_loss else None)

    def mil_bag_training(self, bbox_results, gt_labels, neg_weight_list):
        """:1147-1180 -> scalar loss_mil_bags."""
        b = bbox_results["_b200"]
        labels = torch.cat(gt_labels).long().contiguous()
        b["sums"][5:8] = 0
        self._score_select(bbox_results, labels, None, True)
        has_neg = neg_weight_list is not None and b["n_neg"] > 0
        if has_neg:
            w = torch.cat(neg_weight_list).to(torch.uint8).contiguous()
            ops.neg_loss(b["cls"][b["K"]:], w, b["sums"])
        out = ops.finalize_losses(b["sums"], b["K"], has_neg)
        return out[1]

    def mil_bag_selection(self, bbox_results, img_metas, pseudo_bboxes, pseudo_labels):
        """:1112-1145 -> tuple of merged (G_i, 4) boxes."""
        labels = torch.cat(pseudo_labels).long().contiguous()
        pseudo = torch.cat(pseudo_bboxes).float().contiguous()
        merged, pts, idx, sc = self._score_select(bbox_results, labels, pseudo, False)
        bbox_results["_b200"].update(sel_idx=idx, sel_score=sc, merged_points=pts)
        return tuple(torch.split(merged, [len(p) for p in pseudo_bboxes]))

    def MIL_head_burn_in_step2(self, x, img_metas, proposals_list, proposals_valid_list, proposals_reference_list,
                               proposals_real_list, neg_proposal_list, neg_weight_list, pseudo_bboxes,
                               pseudo_labels, fine_proposal_cfg, stage, loss_scales=(1.0, 1.0)):
        """:1318-1344, fused: one score/select launch produces the bag loss AND the merged boxes."""
        num_gt = sum(p.shape[0] for p in pseudo_bboxes)
        per_img = [p.shape[0] for p in pseudo_bboxes]
        R = self.forward_mil_head(num_gt, per_img, x, proposals_list, proposals_valid_list,
                                  proposals_reference_list, proposals_real_list, img_metas, fine_proposal_cfg,
                                  stage, neg_proposal_list, neg_weight_list)
        b = R["_b200"]
        labels = torch.cat(pseudo_labels).long().contiguous()
        pseudo = torch.cat(pseudo_bboxes).float().contiguous()
        merged, pts, idx, sc = ops.score_select(b["cls"], b["ins"], b["evalid"], b["refined"], labels, pseudo,
                                                b["img_wh"], num_gt, b["U1"], b["U2"], self.topk, self.beta,
                                                b["sums"])
        has_neg = neg_weight_list is not None and b["n_neg"] > 0
        if has_neg:
            ops.neg_loss(b["cls"][b["K"]:], torch.cat(neg_weight_list).to(torch.uint8).contiguous(), b["sums"])
        out = ops.finalize_losses(b["sums"], b["K"], has_neg, loss_scales[0], loss_scales[1])
        losses = {f"stage{stage}_loss_mil_bbox": out[0], f"stage{stage}_loss_mil_bags": out[1],
                  f"stage{stage}_coarse_bags_iou": out[2], f"stage{stage}_refine_bags_iou": out[3]}
        b.update(sel_idx=idx, sel_score=sc, merged_points=pts)
        self.last_results = R
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
        num_gt = sum(p.shape[0] for p in pseudo_bboxes)
        per_img = [p.shape[0] for p in pseudo_bboxes]
        R = self.forward_mil_head(num_gt, per_img, x, proposals_list, proposals_valid_list, proposals_reference_list,
                                  proposals_real_list, img_metas, fine_proposal_cfg, stage)
        losses = {f"stage{stage}_coarse_bags_iou": R["coarse_bags_iou"],
                  f"stage{stage}_refine_bags_iou": R["refine_bags_iou"]}
        merged = self.mil_bag_selection(R, img_metas, pseudo_bboxes, pseudo_labels)
        self.last_results = R
        return list(merged), losses

    def mil_stage_packed(self, x, img_metas, img_wh, base_rois, U1, ref, real, neg_boxes, neg_img_idx, bag_offsets,
                         labels, pseudo, cfg, stage, loss_scales=(1.0, 1.0), keep=None):
        """One MIL stage on packed tensors (the fast path behind ``phase2_refine``): no per-image lists, no
        replicated reference/real boxes (instance k belongs to GT k // (U1*U2)), negatives appended to the
        classification pass.  base_rois (G*U1,5|6); ref/real/pseudo (G,4|5); labels (G,) int64;
        neg_boxes (Nn,4|5)|None with neg_img_idx (Nn,) int32; bag_offsets (B+1,) int32 into base_rois.
        Returns (merged (G,4|5), merged centres (G,2), losses dict)."""
        dev = x[0].device
        rot = self.bbox_roi_extractor.rotated
        rs = 6 if rot else 5
        n_neg = 0 if neg_boxes is None else neg_boxes.shape[0]
        U2 = len(cfg["base_ratios"]) ** 2 * (1 + 4 * len(cfg["shake_ratio"] or []))
        K, G = base_rois.shape[0] * U2, pseudo.shape[0]
        rois2 = torch.empty((K + n_neg, rs), dtype=torch.float32, device=dev)
        side = {}

        def negatives():
            if n_neg:
                ops.make_rois(neg_boxes, neg_img_idx, out=rois2[K:])
                side["neg_w"] = ops.neg_weight(rois2[K:], base_rois, bag_offsets, rot)
        ev_reg, ev_all = self._side_work(stage, negatives)
        neg_w = side.get("neg_w")
        ebags, evalid = ops.bag_gen(base_rois, img_wh, cfg["base_ratios"], cfg["shake_ratio"], cfg["min_scale"], rot)
        assert ebags.shape[0] == K
        sums = torch.zeros((8,), dtype=torch.float32, device=dev)
        kreg = {} if keep is not None else None
        kbag = {} if keep is not None else None
        A = self._roi_operand(x, ebags)
        torch.cuda.current_stream().wait_event(ev_reg)
        H = self._fc_stack(A, self.shared_fcs_reg[stage], K, kreg)
        h0, w0, _ = img_metas[0]["img_shape"]
        fr = self.fc_reg[stage]
        _, deltas, iou_t = ops.reg_decode(H, fr.weight.detach(), fr.bias.detach(), ebags, evalid, ref, real, U1 * U2,
                                          (w0, h0), sums, K=K, hyper=self.loss_bbox_denosing_hyper, out_rois=rois2,
                                          rotated=rot, want_deltas=keep is not None)
        del A, H
        torch.cuda.current_stream().wait_event(ev_all)      # negatives' RoIs + the bag stack's operands
        A2 = self._roi_operand(x, rois2)
        H2 = self._fc_stack(A2, self.shared_fcs_bag[stage], K + n_neg, kbag)
        fc, fi = self.fc_cls[stage], self.fc_ins[stage]
        cls, ins = ops.cls_ins_heads(H2, fc.weight.detach(), fc.bias.detach(), fi.weight.detach(),
                                     fi.bias.detach(), M=K + n_neg)
        if n_neg:
            with ops.fork() as fneg:                      # beside score_select (both reduce into ``sums``)
                ops.neg_loss(cls[K:], neg_w, sums)
        merged, pts, idx, sc = ops.score_select(cls, ins, evalid, rois2, labels, pseudo, img_wh, G, U1, U2,
                                                self.topk, self.beta, sums, rot)
        if n_neg:
            fneg.join()
        out = ops.finalize_losses(sums, K, bool(n_neg), loss_scales[0], loss_scales[1], self.bag_loss_pos_scale,
                                  self.bag_loss_neg_scale)
        losses = {f"stage{stage}_loss_mil_bbox": out[0], f"stage{stage}_loss_mil_bags": out[1],
                  f"stage{stage}_coarse_bags_iou": out[2], f"stage{stage}_refine_bags_iou": out[3]}
        self.last_losses = losses
        if keep is not None:
            keep.update(reg=kreg, bag=kbag, deltas=deltas, ebags=ebags, evalid=evalid, ref=ref, rois2=rois2, cls=cls,
                        ins=ins, neg_w=neg_w, n_neg=n_neg, labels=labels, sums=sums, K=K, G=G, U1=U1, U2=U2,
                        max_wh=(w0, h0), stage=stage, loss_scales=loss_scales)
        self.last_results = dict(
            cls_score=cls[:K].view(G, U1, U2, -1), ins_score=ins[:K].view(G, U1, U2, -1),
            neg_cls_score=cls[K:] if n_neg else None, neg_weight=neg_w, iou_target=iou_t,
            extensive_bags=[rois2[:K, 1:rs]], base_shaking_num=U1, extensive_shaking_num=U2,
            _b200=dict(sums=sums, K=K, evalid=evalid, refined=rois2[:K], coarse=ebags, img_wh=img_wh, cls=cls,
                       ins=ins, n_neg=n_neg, U1=U1, U2=U2, sel_idx=idx, sel_score=sc, merged_points=pts))
        return merged, pts, losses

    def MIL_head_burn_in_step1(self, x_ori, x_synethic, img_metas, proposals_list, proposals_valid_list,
                               proposals_reference_list, proposals_real_list, syn_proposals_list,
                               syn_proposals_valid_list, syn_proposals_reference_list, syn_proposals_real_list,
                               neg_proposal_list, neg_weight_list, synthetic_bboxes, pseudo_bboxes, pseudo_labels,
                               fine_proposal_cfg, stage):
        """:1279-1316: regression loss from the synthetic image's bags, bag loss + selection from the real one."""
        n_syn = sum(b.shape[0] for b in synthetic_bboxes)
        syn_per_img = [b.shape[0] for b in synthetic_bboxes]
        syn = self.forward_mil_head(n_syn, syn_per_img, x_synethic, syn_proposals_list, syn_proposals_valid_list,
                                    syn_proposals_reference_list, syn_proposals_real_list, img_metas,
                                    fine_proposal_cfg, stage)
        losses, merged = self.MIL_head_burn_in_step2(x_ori, img_metas, proposals_list, proposals_valid_list,
                                                     proposals_reference_list, proposals_real_list,
                                                     neg_proposal_list, neg_weight_list, pseudo_bboxes,
                                                     pseudo_labels, fine_proposal_cfg, stage)
        losses[f"stage{stage}_loss_mil_bbox"] = syn["loss_mil_bbox"]
        return losses, merged


@HEADS.register_module(name="B200MILHead", force=True)
class MILHead(nn.Module, MILHeadMixin):
    """Standalone MIL head (the FCOS tower of ``TS_P2BFCOSHead`` is off the hot path).  Constructor keywords
    follow the reference head (fcos_head_p2b_ts.py:80-146)."""

    def __init__(self, num_classes, in_channels=256, beta=0.25, top_k=3, num_stages=2,
                 bbox_roi_extractor=dict(type="SingleRoIExtractor", roi_layer=dict(type="RoIAlign", output_size=7),
                                         out_channels=256, featmap_strides=[8]),
                 loss_bbox_denosing=dict(type="DN_DIoULoss", loss_weight=1.0, hyper=0.2), precision="bf16",
                 feat_dtype=None, **kwargs):
        super().__init__()
        self.num_classes, self.in_channels = num_classes, in_channels
        self.beta, self.topk, self.num_stages = beta, top_k, num_stages
        # bf16 precision: fp16 NHWC feature map (TMA + tensor-core RoIAlign; fp16 keeps 3 more mantissa bits than
        # bf16 through the interpolation, saturating at +-65504); fp32 precision: fp32 feature map
        if feat_dtype is None:
            feat_dtype = torch.float16 if precision == "bf16" else torch.float32
        self.precision, self.feat_dtype = precision, feat_dtype
        if loss_bbox_denosing.get("type") != "DN_DIoULoss" or loss_bbox_denosing.get("loss_weight", 1.0) != 1.0:
            raise NotImplementedError("the fused decode kernel implements DN_DIoULoss(loss_weight=1.0)")
        self.loss_bbox_denosing_hyper = loss_bbox_denosing.get("hyper", 0.2)
        self.bbox_roi_extractor = build_roi_extractor(bbox_roi_extractor)
        self.conv_mil = nn.ModuleList()
        self._init_mil_layers()


@HEADS.register_module(name="B200RotatedMILHead", force=True)
class RotatedMILHead(MILHead):
    """MIL part of ``TS_P2RBRotatedFCOSHead`` (OBB_TOD/mmrotate/models/dense_heads/rotated_fcos_head_p2rb_ts.py
    :1186-1453): 5-d boxes, RoIAlignRotated, regression on the horizontal (cx,cy,w,h) box with the angle
    carried through (:1314-1334), rotated-IoU logs, bag loss 0.25 * pos + 0.75 * neg (:1272,1282) and the
    top-k score-weighted merge with the (cx,cy) clamp quirk (:1198-1216)."""
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
