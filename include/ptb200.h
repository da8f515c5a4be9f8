/* ptb200.h -- C-ABI of the B200-native Point Teacher hot path (libptb200.so, sm_100a).
 *
 * Conventions: every entry point returns 0 on success or a negative PT_ERR_* code and never throws;
 * pt_last_error() returns the thread-local message.  All pointers are DEVICE pointers unless the
 * parameter name ends in _host.  The caller owns all memory (outputs pre-allocated), kernels are
 * enqueued on `stream` (a cudaStream_t passed as void*), there are no internal syncs and no global
 * state, so calls are re-entrant per stream and CUDA-graph capturable.
 *
 * Each function names the reference interface it replaces (paths relative to /root/reference).
 */
#ifndef PTB200_H_
#define PTB200_H_
#ifdef __cplusplus
extern "C" {
#endif

#define PT_OK 0
#define PT_ERR_ARG (-1)
#define PT_ERR_CUDA (-2)
#define PT_ERR_UNSUPPORTED (-3)
#define PT_ERR_DRIVER (-4)

const char* pt_last_error(void);
int pt_abi_version(void);
const char* pt_build_arch(void);

/* ---- bag construction ----------------------------------------------------------------------
 * fine_proposals_from_cfg + bbox2roi:
 *   HBB_TOD/mmdet/models/detectors/syn_images_generator_v2.py:262-324, core/bbox/transforms.py:58-78.
 * in_rois [G,5] (img,x1,y1,x2,y2) -> out_rois [G*U,5], valid [G*U] (iof with the image > 0.7),
 * U = n_ratios^2 * (1 + 4*n_shake); order GT-major, ratio_w outer, ratio_h inner, shake minor.
 * ratios_host / shake_host are HOST arrays.  img_wh [B,2] = (w,h) per image. */
int pt_bag_gen(const float* in_rois, long long G, const float* img_wh, int B, const float* ratios_host,
               int n_ratios, const float* shake_host, int n_shake, float min_scale, float* out_rois,
               unsigned char* valid, int rotated, void* stream);
/* rotated != 0: the OBB twin (OBB_TOD/mmrotate/models/detectors/syn_images_generator_v2.py:26-40): RoIs are
 * [G,6] (img,cx,cy,w,h,theta); bags are generated on cxcywh_to_xyxy(box[:4]) and returned as cxcywh + theta. */

/* bbox2roi / rbbox2roi (HBB_TOD/mmdet/core/bbox/transforms.py:58-78, OBB_TOD/mmrotate/core/bbox/transforms.py:73-92):
 * boxes [n, ldb] + img_idx [n] int32 -> out_rois [n, box_dim+1] = (img, box...). */
int pt_make_rois(const float* boxes, int ldb, const int* img_idx, int n, int box_dim, float* out_rois, void* stream);

/* gen_negative_proposals' weight test (syn_images_generator_v2.py:254-255):
 * weight[n] = all(IoU(neg n, base bags of the same image) < 0.3).  bag_rois sorted by image,
 * bag_offsets [B+1] int32. */
int pt_neg_weight(const float* neg_rois, int n_neg, const float* bag_rois, const int* bag_offsets, int B,
                  unsigned char* weight, int rotated, void* stream);
/* rotated: rbbox_overlaps of [*,6] RoIs (OBB_TOD/.../syn_images_generator_v2.py:146-152). */

/* mmcv.ops.box_iou_rotated behind rbbox_overlaps
 * (OBB_TOD/mmrotate/core/bbox/iou_calculators/rotate_iou2d_calculator.py:53-89): a [M,5], b [N,5]
 * (cx,cy,w,h,theta rad); mode 0 iou / 1 iof; clamp_wh != 0 applies rbbox_overlaps' w,h >= 1e-3 clamp. */
int pt_box_iou_rotated(const float* a, int lda, const float* b, int ldb, long long M, long long N, int mode,
                       int aligned, int clamp_wh, float* out, void* stream);

/* bbox_overlaps (HBB_TOD/mmdet/core/bbox/iou_calculators/iou2d_calculator.py:74-260):
 * mode 0 iou / 1 iof / 2 giou; a [M,4] row stride lda floats, b [N,4] stride ldb; out (M,N) or,
 * aligned, (M,). */
int pt_bbox_overlaps(const float* a, int lda, const float* b, int ldb, long long M, long long N, int mode,
                     int aligned, float eps, float* out, void* stream);

/* ---- RoIAlign --------------------------------------------------------------------------------
 * mmcv.ops.RoIAlign / RoIAlignRotated forward as built by
 *   HBB_TOD/mmdet/models/roi_heads/roi_extractors/base_roi_extractor.py:50-59 and called from
 *   single_level_roi_extractor.py:56-114 / OBB_TOD/.../rotate_single_level_roi_extractor.py:90-167.
 * pt_nchw_to_nhwc: [B,C,H,W] fp32 -> [B,H,W,C] fp32 (out_bf16 = 0), bf16 (1) or saturating fp16 (2); once per step.
 * feat_bf16 of pt_roi_align_forward uses the same dtype code.  bf16/fp16 feature maps with out_mode 0 take the
 * TMA + mma.sync throughput kernel (roi_align_mma.cu) when C % 64 == 0 and C <= 256.
 * pt_roi_align_forward: feat NHWC; rois [K,5] or rotated [K,6] (img,cx,cy,w,h,theta rad).
 *   out_mode 0: bf16 [K, ld_out], column (ph*7+pw)*C + c (the FC1 operand layout)
 *   out_mode 1: fp32 [K,C,7,7] (the extractor's public layout)
 *   out_mode 2: bf16 [K, ld_out] = [hi | lo | hi] segments of 49*C (fp32-emulation operand) */
int pt_nchw_to_nhwc(const float* in, void* out, int B, int C, int H, int W, int out_bf16, void* stream);
/* same, and (fp16 output only) adds to *sat_count (device int32, may be NULL) the number of values the saturating
 * conversion changed (|v| > 65504): the fp32 reference has no such limit, so the host layer refuses / falls back. */
int pt_nchw_to_nhwc_ex(const float* in, void* out, int B, int C, int H, int W, int out_bf16, int* sat_count,
                       void* stream);
int pt_roi_align_forward(const void* feat, int feat_bf16, const float* rois, void* out, long long ld_out,
                         int out_mode, int K, int B, int C, int H, int W, int pooled, float spatial_scale,
                         int sampling_ratio, int aligned, int rotated, int clockwise, const int* roi_level,
                         int level, void* stream);
/* multi-level FPN helpers (single_level_roi_extractor.py:35-54, base_roi_extractor.py:61-83):
 * levels[k] = clamp(floor(log2(sqrt(w*h)/finest_scale + 1e-6)), 0, L-1); pt_roi_align_forward then skips
 * RoIs whose roi_level != level (roi_level may be NULL: single level). */
int pt_map_roi_levels(const float* rois, int K, int rotated, float finest_scale, int num_levels, int* levels,
                      void* stream);
int pt_roi_rescale(const float* rois, int K, int rotated, float factor_h, float factor_w, float* out, void* stream);

/* ---- bag-feature FC (tcgen05 / TMA bf16 GEMM) --------------------------------------------------
 * nn.Linear (+ReLU) of shared_fcs_reg / shared_fcs_bag,
 *   HBB_TOD/mmdet/models/dense_heads/fcos_head_p2b_ts.py:1205-1206, 1246-1247, 1271-1272.
 * C[M,N] = act(A[M,K] * B[N,K]^T + bias); A, B bf16 row-major (K contiguous); C bf16 or fp32.
 * N % 256 == 0; a ragged K is zero-filled by TMA.  workspace: pt_fc_gemm_workspace_bytes(), ZERO-filled once by the
 * caller (the kernel leaves it zeroed); may be NULL (no split-K tail balancing). */
long long pt_fc_gemm_workspace_bytes(int num_sms);
int pt_fc_gemm_bf16(const void* A, long long lda, const void* B, long long ldb, const float* bias, void* C,
                    long long ldc, int M, int N, int K, int relu, int out_f32, void* workspace,
                    long long workspace_bytes, int num_sms, int allow_split, void* stream);

/* weight preparation for the GEMM (FC1 column permutation c*49+bin -> bin*C+c, bf16 cast;
 * x3: [hi | hi | lo] K segments for the fp32-emulation mode) */
int pt_prep_fc1_weight(const float* w, void* out_bf16, int N, int C, int bins, long long ld, int x3, void* stream);
int pt_cast_weight_bf16(const float* w, void* out_bf16, int N, int K, long long ld, int x3, void* stream);

/* ---- MIL head tails ------------------------------------------------------------------------------
 * pt_reg_decode: fc_reg + DeltaXYWHBBoxCoder.decode + DN_DIoULoss partial sums + IoU logs
 *   (fcos_head_p2b_ts.py:1207-1223; core/bbox/coder/delta_xywh_bbox_coder.py:144-270;
 *    models/losses/iou_loss.py:398-466).  sums: float[8], zeroed by the caller per stage.
 * pt_cls_ins_heads: fc_cls / fc_ins (fcos_head_p2b_ts.py:1249, :1273).
 * pt_score_select: mil_bag_training's positive term + mil_bag_selection(_single)
 *   (fcos_head_p2b_ts.py:1092-1180); top-k replays ATen's CPU tie rule.
 *   pseudo / merged_pts / sums may be NULL (no beta blend / no centres / no loss accumulation).
 * pt_neg_loss: negative-bag term (:1169-1179).  pt_finalize_losses -> out[5] =
 *   {scale_bbox*loss_mil_bbox, scale_bags*loss_mil_bags, coarse_bags_iou, refine_bags_iou, num_sample};
 *   the scales are the detector's alpha (fcos_p2b_teacher_student.py:460-461). */
int pt_reg_decode(const void* H, int h_f32, long long ldh, int D, const float* Wreg, const float* breg,
                  const float* bag_rois, const unsigned char* valid, const float* ref_boxes,
                  const float* real_boxes, int U, int K, float max_w, float max_h, float wh_ratio_clip,
                  float hyper, float eps, float* out_rois, float* out_deltas, float* iou_target, float* sums,
                  int rotated, const float* deltas_in, void* stream);
/* deltas_in != NULL: [K,4] = H . Wreg^T (no bias) already computed by pt_small_heads_bf16; only the decode runs.
 * pt_small_heads_bf16: out0 [M,n0] = H W0^T + b0, out1 [M,n1] = H W1^T + b1 on mma.sync (H bf16; W1/b0/b1 may be
 * NULL): fc_reg (fcos_head_p2b_ts.py:1207) and fc_cls + fc_ins (:1250-1251). */
int pt_small_heads_bf16(const void* H, long long ldh, int D, const float* W0, int n0, const float* b0, const float* W1,
                        int n1, const float* b1, int M, float* out0, float* out1, void* stream);
int pt_cls_ins_heads(const void* H, int h_f32, long long ldh, int D, const float* Wcls, const float* bcls,
                     const float* Wins, const float* bins, int C, int M, float* cls, float* ins, void* stream);
int pt_score_select(const float* cls, const float* ins, const unsigned char* valid, const float* bag_rois,
                    const long long* labels, const float* pseudo, const float* img_wh, int B, int G, int U1,
                    int U2, int C, int topk, float beta, float* merged, float* merged_pts, int* sel_idx,
                    float* sel_score, float* sums, int rotated, void* stream);
int pt_neg_loss(const float* neg_cls, const unsigned char* weight, int n, int C, float* sums, void* stream);
int pt_finalize_losses(const float* sums, int K, int has_neg, float scale_bbox, float scale_bags, float pos_w,
                       float neg_w, float* out, void* stream);
/* rotated != 0 selects the OBB head's semantics (OBB_TOD/mmrotate/models/dense_heads/rotated_fcos_head_p2rb_ts.py
 * :1198-1343): 6-column RoIs, decode on the horizontal box with theta carried, rotated-IoU logs, (cx,cy) clamped by
 * w then h, 5-d merged boxes; pos_w / neg_w are the 0.25 / 0.75 bag-loss weights of :1272,1282 (1 / 1 for HBB). */
/* fp32 [M,N] -> bf16 [M,3N] = [hi | lo | hi]: activation operand of the fp32-emulation (bf16x3) GEMM */
int pt_split_bf16x3(const float* in, void* out_bf16, long long M, int N, void* stream);
/* mean aligned IoU of n pairs: the coarse_bboxes_iou / stage{s}_refine_bboxes_iou logs of
 * HBB_TOD/mmdet/models/detectors/fcos_p2b_teacher_student.py:436-438, :457-459 */
int pt_aligned_iou_mean(const float* a, int lda, const float* b, int ldb, int n, int rotated, float* out,
                        void* stream);

/* ---- dense-head label assignment (rows a13-a15) -------------------------------------------------------------
 * FocalLossCost table (HBB_TOD/mmdet/core/bbox/match_costs/match_cost.py:54-100): out[p,c] = (pos - neg) * weight. */
int pt_focal_cost_table(const float* logits, long long n, float alpha, float gamma, float eps, float weight,
                        float* out, void* stream);
/* Stage 1 of TopkAssigner / FUSETopkAssigner (assigners/topk_assigner.py:118-125, fuse_topk_assigner.py:96-101):
 * pre_idx [num_pre, G] int32 = indices of the num_pre smallest PointCost(pts, gts) per GT column, with the
 * ATen CPU tie rule.  pts [P, ldp] (x, y first), gts [G, ldg] (cx, cy first); l2 selects PointCost mode 'L2'.
 * scratch_v / scratch_i [G, P] are required only when num_pre * 64 > P (ATen's nth_element path). */
int pt_topk_pre(const float* pts, int ldp, int P, const float* gts, int ldg, int G, int l2, float weight, int num_pre,
                int* pre_idx, float* scratch_v, int* scratch_i, void* stream);
/* Stage 2 (topk_assigner.py:128-145, fuse_topk_assigner.py:104-119): second cost = fl_table[p, label_c]
 * (+ InsiderCost of pred [P, ldb] = (cx, cy, w, h) vs GT point c, times loc_weight, when pred != NULL); the
 * all-columns top-k quirk and last-GT-wins overwrite are reproduced.  gt_inds / out_labels [P] int64. */
int pt_topk_second(const int* pre_idx, int num_pre, int topk, int G, int P, const float* fl_table, int C,
                   const long long* labels, const float* pred, int ldb, const float* gts, int ldg, float loc_weight,
                   int* assigned_ws, long long* gt_inds, long long* out_labels, void* stream);
/* BboxOverlaps2D (calc 0; iou2d_calculator.py:74-260) / BboxDistanceMetric (calc 1; metric_calculator.py:44-185)
 * matrix [M, N]; modes 0 iou, 1 iof, 2 giou, 3 wd, 4 kl, 5 center_distance2, 6 exp_kl, 7 kl_10 (3.. calc 1 only). */
int pt_bbox_metric(const float* a, int lda, const float* b, int ldb, long long M, long long N, int calc, int mode,
                   float eps, float* out, void* stream);
/* MaxIoUAssigner.assign (assigners/max_iou_assigner.py:60-212) on calc(gts, anchors, mode) without materialising
 * the G x A matrix.  argmax_ws [A] int32, gt_ws [2*G] 32-bit words. */
int pt_max_iou_assign(const float* gts, int ldg, int G, const float* anchors, int lda, int A, int calc, int mode,
                      float eps, float pos_thr, float neg_lo, float neg_hi, float min_pos, int gt_max_assign_all,
                      int match_low_quality, const long long* gt_labels, long long* gt_inds, float* max_overlaps,
                      long long* labels, int* argmax_ws, unsigned* gt_ws, void* stream);

/* ---- coarse pseudo boxes behind the FUSE assignment (section 8f rank 1) ------------------------------------
 * pt_decode_ltrb: distance2bbox + bbox_xyxy_to_cxcywh (HBB_TOD/mmdet/core/bbox/transforms.py:134-166, 249-261).
 * pt_pseudo_aggregate: _gnerate_pseudo_single (HBB_TOD/mmdet/models/dense_heads/fcos_head_p2b_ts.py:762-790):
 *   per GT the sigmoid-score-weighted mean of the decoded boxes of its assigned points (8 x 8 box at the GT point
 *   when none), mean score, assigned count, valid = assigned AND score >= filter, and the IoU-vs-GT sum / count. */
int pt_decode_ltrb(const float* points, const float* ltrb, int P, float* xyxy, float* cxcywh, void* stream);
int pt_pseudo_aggregate(const long long* gt_inds, const long long* labels, const float* cls, int C, const float* xyxy,
                        int P, const float* gt_points, const float* gt_bboxes, int G, float filter_score, float* acc_ws,
                        float* boxes, float* points, float* scores, long long* assign_nums, unsigned char* valid,
                        float* iou_sum, void* stream);

/* pt_ltrb_targets (section 8f rank 2): _get_target_pseudo_single's regression targets + centerness_target
 * (fcos_head_p2b_ts.py:657-708, 1019-1038) without the (P, G, 4) broadcast. */
int pt_ltrb_targets(const float* points, const float* boxes, const long long* gt_inds, const long long* assigned_labels,
                    int P, int num_classes, float* targets, long long* labels, float* centerness, void* stream);

/* ---- phase-1 random region masking (row a16) --------------------------------------------------------------
 * The deterministic tail of generate_black_paper (HBB_TOD/mmdet/models/detectors/syn_images_generator_v2.py:664-690).
 * pt_nms_rotated: mmcv.ops.nms_rotated(dets [N, ld>=5], scores (stride lds), thr): order [N] int32 = box indices by
 *   descending score (stable), keep_sorted [N] uint8 = survivor flags in that order.
 * pt_black_paper_select: score < 1 and inside-image filters (:668-675), obb2poly_le90 + int32 truncation (:678-682);
 *   bb [N,7]; the first *count rows of out_bb [N,7] / out_sel [N] / polys [N,8] are valid (score order).
 * pt_fill_polys: cv2.fillPoly of integer quadrilaterals into img [C,H,W] fp32 (<- value) and / or mask [H,W] uint8. */
long long pt_nms_rotated_workspace_bytes(int N);
int pt_nms_rotated(const float* dets, int ld, const float* scores, int lds, int N, float thr, int* order,
                   unsigned char* keep_sorted, void* workspace, long long workspace_bytes, void* stream);
int pt_black_paper_select(const float* bb, int N, const int* order, const unsigned char* keep_sorted, float imgsize,
                          float* out_bb, int* out_sel, int* polys, int* count, void* stream);
/* same with trig = [N,2] (sin, cos) of every candidate's angle computed on the HOST (the CPU libm sin / cos the reference's tensor ops
 * reach on the strided angle column, as obb2xyxy / obb2poly_le90 of the reference do, syn_images_generator_v2.py:382-396,
 * data_augument_bank.py:516-541): corners then truncate exactly like the reference's; NULL = device sincos. */
int pt_black_paper_select_ex(const float* bb, int N, const int* order, const unsigned char* keep_sorted, float imgsize,
                             float* out_bb, int* out_sel, int* polys, int* count, const float* trig, void* stream);
int pt_fill_polys(const int* polys, const int* count, int max_polys, float* img, unsigned char* mask, int C, int H,
                  int W, float value, void* stream);

/* ---- backward of the MIL head (training step; gradients all-reduced by point_teacher_b200/dist.py) -----------
 * Autograd graph of fcos_head_p2b_ts.py:1147-1236 + delta_xywh_bbox_coder.py:144-250 + iou_loss.py:139-190,398-466.
 * pt_fc_gemm_bf16_ex: pt_fc_gemm_bf16 with an optional bf16 mask [M, ldmask]: out = mask > 0 ? out : 0 (ReLU
 *   backward by the saved forward activation), used for the dgrad GEMMs.
 * pt_reg_loss_grad: g [K,4] = gscale[0] * scale * d loss_mil_bbox / d deltas (sums = the forward's loss sums).
 * pt_bag_loss_grad: g [K + n_neg, 2C] = d loss_mil_bags / d (cls logits | ins logits), positives scaled by
 *   gscale[0] * pos_scale, negatives by gscale[0] * neg_scale.
 * pt_head_bwd: small heads (nout = 4 | 2C): dZ [M,D] bf16 = (g W) masked by H > 0, dW [nout,D] += g^T H, db += sum g.
 * pt_transpose_pad_bf16: [R, C] -> [C, ldout >= R] zero padded (wgrad operands).
 * pt_unpermute_dw1: fp32 [N, bins*C] bin-major FC1 weight gradient -> the parameter's c*bins + bin column order.
 * pt_colsum_bf16: db [N] += column sums of bf16 [M, N].   pt_nhwc_to_nchw_f32: feature-gradient layout change.
 * pt_roi_align_backward: mmcv RoIAlign backward (aligned avg pooling): dfeat NHWC fp32 (zeroed by the caller) +=
 *   scatter of dA bf16 [K, ld] (bin-major columns). */
int pt_fc_gemm_bf16_ex(const void* A, long long lda, const void* B, long long ldb, const float* bias, void* C,
                       long long ldc, int M, int N, int K, int relu, int out_f32, const void* mask, long long ldmask,
                       void* workspace, long long workspace_bytes, int num_sms, int allow_split, void* stream);
/* pt_fc_gemm_bf16_mn: the same GEMM with operands that may be stored contraction-index-major (a_mn: A as [K, M],
 *   b_mn: B as [K, N]; ld = elements per stored row) and are fed to tcgen05 as MN-major shared-memory tiles: the
 *   autograd of nn.Linear (dW = dY^T X, dX = dY W) without materialising a transposed copy. */
int pt_fc_gemm_bf16_mn(const void* A, long long lda, int a_mn, const void* B, long long ldb, int b_mn, const float* bias,
                       void* C, long long ldc, int M, int N, int K, int relu, int out_f32, const void* mask,
                       long long ldmask, void* workspace, long long workspace_bytes, int num_sms, int allow_split,
                       void* stream);
int pt_reg_loss_grad(const float* deltas, const float* bag_rois, const unsigned char* valid, const float* ref_boxes,
                     int U, int K, float max_w, float max_h, float wh_ratio_clip, float hyper, float eps,
                     const float* sums, const float* gscale, float scale, float* g, void* stream);
/* OBB twin (rotated = 1): bag_rois [K,6] (b,cx,cy,w,h,theta), ref_boxes [G,5]; horizontal DN-DIoU on the (cx,cy,w,h)
 *   parts, OBB_TOD/mmrotate/models/dense_heads/rotated_fcos_head_p2rb_ts.py:1314-1322. */
int pt_reg_loss_grad_ex(const float* deltas, const float* bag_rois, const unsigned char* valid, const float* ref_boxes,
                        int U, int K, float max_w, float max_h, float wh_ratio_clip, float hyper, float eps,
                        const float* sums, const float* gscale, float scale, float* g, int rotated, void* stream);
/* mmcv RoIAlignRotated backward (call site OBB_TOD/mmrotate/models/roi_heads/roi_extractors/
 *   rotate_single_level_roi_extractor.py:90-167): rois [K,6]; dfeat NHWC fp32 (zeroed by the caller) += scatter. */
int pt_roi_align_rotated_backward(const void* dA_bf16, long long ld, const float* rois, int K, int B, int C, int H,
                                  int W, float spatial_scale, int sampling_ratio, int aligned, int clockwise,
                                  float* dfeat, void* stream);
int pt_bag_loss_grad(const float* cls, const float* ins, const unsigned char* valid, const long long* labels, int G,
                     int U1, int U2, int C, const unsigned char* neg_weight, int n_neg, const float* sums,
                     const float* gscale, float pos_scale, float neg_scale, float* g, void* stream);
int pt_head_bwd(const float* g, int nout, const void* H_bf16, long long ldh, int D, const float* W, int M,
                void* dZ_bf16, long long ldz, float* dW, float* db, void* stream);
int pt_transpose_pad_bf16(const void* in, long long ldin, int R, int C, void* out, long long ldout, void* stream);
int pt_unpermute_dw1(const float* dw_binmajor, int N, int C, int bins, float* grad, int accumulate, void* stream);
int pt_colsum_bf16(const void* dZ, long long ld, int M, int N, float* db, void* stream);
int pt_nhwc_to_nchw_f32(const float* in, float* out, int B, int C, int H, int W, int accumulate, void* stream);
int pt_roi_align_backward(const void* dA_bf16, long long ld, const float* rois, int K, int B, int C, int H, int W,
                          float spatial_scale, int sampling_ratio, int aligned, float* dfeat, void* stream);
/* multi-level FPN twins (single_level_roi_extractor.py:98-104): RoIs whose roi_level[k] != level are skipped, so one
 * call per level scatters each RoI's gradient into its own level's map.  roi_level may be NULL (= all RoIs). */
int pt_roi_align_backward_ex(const void* dA_bf16, long long ld, const float* rois, int K, int B, int C, int H, int W,
                             float spatial_scale, int sampling_ratio, int aligned, float* dfeat, const int* roi_level,
                             int level, void* stream);
int pt_roi_align_rotated_backward_ex(const void* dA_bf16, long long ld, const float* rois, int K, int B, int C, int H,
                                     int W, float spatial_scale, int sampling_ratio, int aligned, int clockwise,
                                     float* dfeat, const int* roi_level, int level, void* stream);

/* ---- strong_augmentation ---------------------------------------------------------
 * Replaces HBB_TOD/mmdet/models/detectors/syn_images_generator_v2.py:24-132 and
 * OBB_TOD/mmrotate/models/detectors/syn_images_generator_v2.py:223-357.
 * params: fp32 [B, pt_augment_param_stride()], one block per image:
 *   [0] flip x  [1] flip y  [2] rotate flag  [3..6] r00 r01 r10 r11 (affine sampling grid theta^T / (w/2, h/2))
 *   [7] scale_H [8] scale_W [9] start_y [10] start_x [11] zero-pad flag (scale < 1)  [12] scale factor
 *   [13] cos(-angle) [14] sin(-angle) [15] blank_w [16] blank_h.
 * pt_augment_image: flip -> (rotate nearest, fill 0) -> bilinear resize (align_corners = false) -> crop / pad ->
 *   round, one pass, [B,C,H,W] fp32 -> [B,C,H,W] fp32.
 * pt_augment_coords: the same transform on the packed GT-point / pseudo-point / pseudo-box lists (int32 offsets
 *   [B+1]) with the reference's in-image filters and a stable compaction; box_dim 4 = xyxy, 5 = (cx,cy,w,h,theta)
 *   le90 via obb2poly / poly2obb; counts int32 [B,2] = kept (GT points, pseudo entries) per image. */
int pt_augment_param_stride(void);
int pt_augment_image(const float* img, float* out, const float* params, int B, int C, int H, int W, void* stream);
int pt_augment_coords(const float* gt_pts, const long long* gt_lab, const int* gt_off, const float* ps_pts,
                      const long long* ps_lab, const float* ps_box, const int* ps_off, int box_dim, const float* params,
                      int B, int H, int W, float* o_gt_pts, long long* o_gt_lab, float* o_ps_pts, long long* o_ps_lab,
                      float* o_ps_box, int* counts, void* stream);

/* ---- student-branch losses that are mmcv-native in the reference -----------------
 * pt_sigmoid_focal_loss: HBB_TOD/mmdet/models/losses/focal_loss.py:11-100 (mmcv.ops.sigmoid_focal_loss).
 *   pred [N,C] logits, target [N] int64 in [0,C] (C = background); weight_mode 0 none | 1 [N] | 2 [N*C];
 *   loss_elem / grad_elem [N,C] (either may be NULL): weighted element loss and its derivative w.r.t. pred;
 *   sum (may be NULL): += total weighted loss.
 * pt_rotated_iou_loss: OBB_TOD/mmrotate/models/losses/rotated_iou_loss.py:17-147 on mmcv.ops.diff_iou_rotated_2d.
 *   pred / target [n,5] (cx,cy,w,h,theta); mode 0 log | 1 linear | 2 square; dn = 1: DN_iou_loss with `hyper`;
 *   loss [n] element losses, grad [n,5] (may be NULL) = d loss / d pred. */
int pt_sigmoid_focal_loss(const float* pred, const long long* target, const float* weight, int weight_mode, float gamma,
                          float alpha, int N, int C, float* loss_elem, float* grad_elem, float* sum, void* stream);
int pt_rotated_iou_loss(const float* pred, const float* target, int n, int mode, float eps, int dn, float hyper,
                        float* loss, float* grad, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PTB200_H_ */
