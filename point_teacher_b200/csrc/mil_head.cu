// MIL two-stream head: weight preparation, small-N heads fused with box decode / losses, and the fused
// sigmoid x instance-softmax score + top-k instance selection + pseudo-box write-back.
//
// Replaces (paths under /root/reference/HBB_TOD/mmdet/):
//   fc_reg + DeltaXYWHBBoxCoder.decode + DN_DIoULoss + IoU logs   models/dense_heads/fcos_head_p2b_ts.py:1207-1223,
//                                   core/bbox/coder/delta_xywh_bbox_coder.py:144-270, models/losses/iou_loss.py:398-466
//   fc_cls / fc_ins                 fcos_head_p2b_ts.py:1249, :1273
//   mil_bag_training + gfocal_loss  fcos_head_p2b_ts.py:1147-1180, :1074-1078
//   mil_bag_selection(_single)      fcos_head_p2b_ts.py:1092-1145
// Reductions use warp shuffles; the top-k replays ATen's CPU tie rule (libstdc++ partial_sort / nth_element on
// (value, index) pairs with a value-only comparator) so selected indices match the
// reference's CPU path even on exact ties.
#include "common.cuh"
#include "rotated_iou.cuh"
#include "topk_replay.cuh"

namespace ptb {

// --------------------------------------------------------------------------------- weight preparation
// FC1 weight [N, C*49] with k = c*49 + bin  ->  bf16 [N, ld] with k' = bin*C + c (the order RoIAlign emits).
// x3: three K segments [hi | hi | lo] pairing with the activation's [hi | lo | hi].
__global__ void prep_fc1_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int C, int bins,
                                       long long ld, int x3) {
  // transposed staging [bin][C+4]: rows stay 16-byte aligned so the read phase is two conflict-free LDS.128 per
  // 8 outputs.  (Measured alternatives: [C+1] with 8 scalar reads per output group, 8-way conflicted: +10 us per
  // step; a two-stage scalar-load variant with conflict-free stores on both sides: +27 us -- the 4-byte global
  // loads cost more than the 16-way store conflicts of this layout.)
  extern __shared__ __align__(16) float srow[];
  const int n = blockIdx.x, K = C * bins, S = C + 4;
  const float* src = w + (size_t)n * K;
  // coalesced 16-byte reads (k = c*bins + bin), 4 independent loads in flight per thread: the kernel is a pure
  // HBM stream (77 MB per FC1 weight) and needs ~32 KB in flight per SM to reach the copy bandwidth
  if ((K & 3) == 0) {
    const float4* src4 = reinterpret_cast<const float4*>(src);
    const int K4 = K >> 2;
    for (int i0 = threadIdx.x; i0 < K4; i0 += blockDim.x * 4) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int i = i0 + u * blockDim.x;
        v[u] = i < K4 ? __ldg(src4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int i = i0 + u * blockDim.x;
        if (i < K4) {
          const float e[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
          const int c0 = (4 * i) / bins, bin0 = 4 * i - c0 * bins;
          // lane-rotated element order: consecutive lanes hit bins 4 apart (banks 16 apart with the [C+4] rows);
          // rotating which of the 4 elements each lane stores per instruction spreads a warp over 8 banks
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const int jj = (j + threadIdx.x) & 3;
            const float val = jj == 0 ? e[0] : (jj == 1 ? e[1] : (jj == 2 ? e[2] : e[3]));
            int bin = bin0 + jj, c = c0;
            if (bin >= bins) { bin -= bins; c++; }
            srow[bin * S + c] = val;
          }
        }
      }
    }
  } else {
    for (int i = threadIdx.x; i < K; i += blockDim.x) {
      const int c = i / bins, bin = i - c * bins;
      srow[bin * S + c] = src[i];
    }
  }
  __syncthreads();
  __nv_bfloat16* dst = out + (size_t)n * ld;
  for (int kp = threadIdx.x * 8; kp < K; kp += blockDim.x * 8) {   // 16-byte stores, k' = bin*C + c
    const int bin = kp / C, c = kp - bin * C;
    float v[8], l[8];
    {
      const float4 a = *reinterpret_cast<const float4*>(srow + bin * S + c);
      const float4 b = *reinterpret_cast<const float4*>(srow + bin * S + c + 4);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
#pragma unroll
    for (int j = 0; j < 8; j++) l[j] = v[j] - __bfloat162float(__float2bfloat16_rn(v[j]));
    const uint4 hi = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
    *reinterpret_cast<uint4*>(dst + kp) = hi;
    if (x3) {
      *reinterpret_cast<uint4*>(dst + K + kp) = hi;
      *reinterpret_cast<uint4*>(dst + 2 * K + kp) =
          make_uint4(pack_bf16(l[0], l[1]), pack_bf16(l[2], l[3]), pack_bf16(l[4], l[5]), pack_bf16(l[6], l[7]));
    }
  }
}

// bf16 (non-x3) fast path, whole row in shared memory in PARAMETER order: the global read is one contiguous
// 16-byte-vector stream and lands with conflict-free 16-byte shared stores; the permuted read k' = bin*C + c ->
// c*bins + bin walks shared memory at stride `bins` (49: odd, bank-conflict-free scalar loads); 16-byte global stores.
__global__ void __launch_bounds__(256)
prep_fc1_weight_row_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int C, int bins, long long ld) {
  extern __shared__ __align__(16) float prow[];
  const int n = blockIdx.x, K = C * bins;
  const float4* src4 = reinterpret_cast<const float4*>(w + (size_t)n * K);
  float4* p4 = reinterpret_cast<float4*>(prow);
  for (int i = threadIdx.x; i < K / 4; i += blockDim.x) p4[i] = __ldcs(src4 + i);     // streamed: read exactly once
  __syncthreads();
  __nv_bfloat16* dst = out + (size_t)n * ld;
  for (int kp = threadIdx.x * 8; kp < K; kp += blockDim.x * 8) {
    const int bin = kp / C, c = kp - bin * C;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; j++) v[j] = prow[(c + j) * bins + bin];
    *reinterpret_cast<uint4*>(dst + kp) =
        make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
  }
}

// bf16 (non-x3) cast with 16-byte loads / 8-byte stores, no per-element division (K % 4 == 0, ld % 4 == 0)
__global__ void __launch_bounds__(256)
cast_weight_vec_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int N, int K4, long long ld) {
  const long long total = (long long)N * K4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / K4), k4 = (int)(i - (long long)n * K4);
    const float4 v = __ldcs(reinterpret_cast<const float4*>(w) + i);
    *reinterpret_cast<uint2*>(out + (size_t)n * ld + 4 * k4) = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
  }
}

__global__ void cast_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int N, int K,
                                   long long ld, int x3) {
  const long long total = (long long)N * K;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / K, k = i - n * K;
    const float v = w[i];
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    out[n * ld + k] = hi;
    if (x3) {
      out[n * ld + K + k] = hi;
      out[n * ld + 2 * K + k] = __float2bfloat16_rn(v - __bfloat162float(hi));
    }
  }
}

// --------------------------------------------------------------------------------- row . small-N weight
template <typename TH> struct RowLoader;
template <> struct RowLoader<__nv_bfloat16> {
  // 8 consecutive hidden units per lane per step
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float* v) {
    uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; i++) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
  }
};
template <> struct RowLoader<float> {
  static __device__ __forceinline__ void load(const float* p, float* v) {
    float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
};

// out[r][o] = sum_k h_r[k] * Wsm[o*D + k] for ROWS consecutive rows: every weight vector fetched from shared
// memory is reused by ROWS rows (the small-N heads are shared-memory-bandwidth bound otherwise).  All lanes
// return the reduced values.  nrows <= ROWS rows are valid.
constexpr int HEAD_ROWS = 4;
template <typename TH, int NOUT>
__device__ __forceinline__ void rows_dot(const TH* __restrict__ h, long long ldh, int nrows,
                                         const float* __restrict__ wsm, int D, int lane,
                                         float (&out)[HEAD_ROWS][NOUT]) {
#pragma unroll
  for (int r = 0; r < HEAD_ROWS; r++)
#pragma unroll
    for (int o = 0; o < NOUT; o++) out[r][o] = 0.f;
  for (int k0 = lane * 8; k0 < D; k0 += 256) {
    float v[HEAD_ROWS][8];
#pragma unroll
    for (int r = 0; r < HEAD_ROWS; r++) {
      if (r < nrows) RowLoader<TH>::load(h + (size_t)r * ldh + k0, v[r]);
      else {
#pragma unroll
        for (int j = 0; j < 8; j++) v[r][j] = 0.f;
      }
    }
#pragma unroll
    for (int o = 0; o < NOUT; o++) {
      const float4* wp = reinterpret_cast<const float4*>(wsm + (size_t)o * D + k0);
      const float4 w0 = wp[0], w1 = wp[1];
#pragma unroll
      for (int r = 0; r < HEAD_ROWS; r++)
        out[r][o] += v[r][0] * w0.x + v[r][1] * w0.y + v[r][2] * w0.z + v[r][3] * w0.w + v[r][4] * w1.x +
                     v[r][5] * w1.y + v[r][6] * w1.z + v[r][7] * w1.w;
    }
  }
#pragma unroll
  for (int r = 0; r < HEAD_ROWS; r++)
#pragma unroll
    for (int o = 0; o < NOUT; o++) out[r][o] = warp_sum(out[r][o]);
}

__device__ __forceinline__ float aligned_iou(const float* a, const float* b) {
  const float a1 = fmul(fsub(a[2], a[0]), fsub(a[3], a[1])), a2 = fmul(fsub(b[2], b[0]), fsub(b[3], b[1]));
  const float ow = fmaxf(fsub(fminf(a[2], b[2]), fmaxf(a[0], b[0])), 0.f);
  const float oh = fmaxf(fsub(fminf(a[3], b[3]), fmaxf(a[1], b[1])), 0.f);
  const float ov = fmul(ow, oh);
  return fdiv(ov, fmaxf(fsub(fadd(a1, a2), ov), 1e-6f));
}

// DIoU element (models/losses/iou_loss.py:139-190 body), eps inside the denominators
__device__ __forceinline__ float diou_elem(const float* p, const float* t, float eps) {
  const float ow = fmaxf(fminf(p[2], t[2]) - fmaxf(p[0], t[0]), 0.f);
  const float oh = fmaxf(fminf(p[3], t[3]) - fmaxf(p[1], t[1]), 0.f);
  const float ov = ow * oh;
  const float ap = (p[2] - p[0]) * (p[3] - p[1]), ag = (t[2] - t[0]) * (t[3] - t[1]);
  const float iou = ov / (ap + ag - ov + eps);
  const float cw = fmaxf(fmaxf(p[2], t[2]) - fminf(p[0], t[0]), 0.f);
  const float ch = fmaxf(fmaxf(p[3], t[3]) - fminf(p[1], t[1]), 0.f);
  const float c2 = cw * cw + ch * ch + eps;
  const float dx = (t[0] + t[2]) - (p[0] + p[2]), dy = (t[1] + t[3]) - (p[1] + p[3]);
  const float rho2 = dx * dx / 4.f + dy * dy / 4.f;
  return 1.f - (iou - rho2 / c2);
}

// sums[]: 0 S_base_diou, 1 S_weight, 2 S_weight*min_bank, 3 S_refine_iou(vs real), 4 S_coarse_iou(vs real),
//         5 S_pos_bag_loss, 6 num_sample, 7 S_neg_bag_loss
enum { S_BASE = 0, S_W = 1, S_WMIN = 2, S_REFINE = 3, S_COARSE = 4, S_POS = 5, S_NSAMPLE = 6, S_NEG = 7, S_COUNT = 8 };

__device__ __forceinline__ void block_accumulate(float* sums, const float* vals, int n, float* red) {
  // vals: per-warp partials already reduced to lane 0; red: smem [warps][n]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (lane == 0)
    for (int i = 0; i < n; i++) red[warp * n + i] = vals[i];
  __syncthreads();
  if (threadIdx.x < n) {
    float s = 0.f;
    for (int w = 0; w < nw; w++) s += red[w * n + threadIdx.x];
    atomicAdd(sums + threadIdx.x, s);
  }
}

// Warp per bag instance k: deltas = H[k]·Wreg^T + b  ->  decode against the coarse bag  ->  refined RoI,
// IoU logs and DN-DIoU partial sums.  ref_boxes / real_boxes are [G,4]; instance k belongs to GT k / U.
template <typename TH>
__global__ void reg_decode_kernel(const TH* __restrict__ H, long long ldh, int D, const float* __restrict__ Wreg,
                                  const float* __restrict__ breg, const float* __restrict__ bag_rois,
                                  const uint8_t* __restrict__ valid, const float* __restrict__ ref_boxes,
                                  const float* __restrict__ real_boxes, int U, int K, float max_w, float max_h,
                                  float max_ratio, float hyper, float eps, float* __restrict__ out_rois,
                                  float* __restrict__ out_deltas, float* __restrict__ iou_target,
                                  float* __restrict__ sums, int rotated, const float* __restrict__ deltas_in) {
  extern __shared__ float wsm[];  // [4][D] + reduction scratch
  // deltas_in != NULL: H . Wreg^T was already computed on tensor cores (heads_mma.cu); only the decode runs here
  if (deltas_in == nullptr)
    for (int i = threadIdx.x; i < 4 * D; i += blockDim.x) wsm[i] = Wreg[i];
  __syncthreads();
  float* red = wsm + 4 * D;
  const int lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  float part[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  for (int kb = (blockIdx.x * nw + (threadIdx.x >> 5)) * HEAD_ROWS; kb < K; kb += gridDim.x * nw * HEAD_ROWS) {
    float dd[HEAD_ROWS][4];
    const int nrows = K - kb < HEAD_ROWS ? K - kb : HEAD_ROWS;
    float d[4] = {0.f, 0.f, 0.f, 0.f};
    if (deltas_in != nullptr) {
      if (lane < nrows) {
        const float4 v = *reinterpret_cast<const float4*>(deltas_in + (size_t)(kb + lane) * 4);
        d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
      }
    } else {
      rows_dot<TH, 4>(H + (size_t)kb * ldh, ldh, nrows, wsm, D, lane, dd);
#pragma unroll
      for (int r = 0; r < HEAD_ROWS; r++)
#pragma unroll
        for (int j = 0; j < 4; j++) d[j] = (lane == r) ? dd[r][j] : d[j];
    }
    const int k = kb + lane;          // lane r decodes row kb + r
    if (lane < nrows) {
      // rotated (OBB_TOD/.../rotated_fcos_head_p2rb_ts.py:1314-1343): decode on cxcywh_to_xyxy(bag[:, :4]),
      // DN-DIoU against cxcywh_to_xyxy(reference[:, :4]), refined bag = xyxy_to_cxcywh(pred) + the bag's angle
      const int rs = rotated ? 6 : 5, bs = rotated ? 5 : 4;
      const float* r = bag_rois + (size_t)k * rs;
      float dx = d[0] + breg[0], dy = d[1] + breg[1], dw = d[2] + breg[2], dh = d[3] + breg[3];
      if (out_deltas) { float* od = out_deltas + (size_t)k * 4; od[0] = dx; od[1] = dy; od[2] = dw; od[3] = dh; }
      float rx1 = r[1], ry1 = r[2], rx2 = r[3], ry2 = r[4];
      if (rotated) {
        rx1 = fsub(r[1], fmul(0.5f, r[3])); ry1 = fsub(r[2], fmul(0.5f, r[4]));
        rx2 = fadd(r[1], fmul(0.5f, r[3])); ry2 = fadd(r[2], fmul(0.5f, r[4]));
      }
      // delta2bbox, means 0 / stds 1 (delta_xywh_bbox_coder.py:209-247)
      const float px = fmul(fadd(rx1, rx2), 0.5f), py = fmul(fadd(ry1, ry2), 0.5f);
      const float pw = fsub(rx2, rx1), ph = fsub(ry2, ry1);
      dw = fminf(fmaxf(dw, -max_ratio), max_ratio);
      dh = fminf(fmaxf(dh, -max_ratio), max_ratio);
      const float gw = fmul(pw, expf(dw)), gh = fmul(ph, expf(dh));
      const float gx = fadd(px, fmul(pw, dx)), gy = fadd(py, fmul(ph, dy));
      float b[4] = {fsub(gx, fmul(gw, 0.5f)), fsub(gy, fmul(gh, 0.5f)), fadd(gx, fmul(gw, 0.5f)),
                    fadd(gy, fmul(gh, 0.5f))};
      b[0] = fminf(fmaxf(b[0], 0.f), max_w); b[2] = fminf(fmaxf(b[2], 0.f), max_w);
      b[1] = fminf(fmaxf(b[1], 0.f), max_h); b[3] = fminf(fmaxf(b[3], 0.f), max_h);
      float* o = out_rois + (size_t)k * rs;
      const float* refp = ref_boxes + (size_t)(k / U) * bs;
      const float* real = real_boxes + (size_t)(k / U) * bs;
      float ref[4];
      if (rotated) {
        const float ob[5] = {fdiv(fadd(b[0], b[2]), 2.f), fdiv(fadd(b[1], b[3]), 2.f), fsub(b[2], b[0]),
                             fsub(b[3], b[1]), r[5]};
        o[0] = r[0]; o[1] = ob[0]; o[2] = ob[1]; o[3] = ob[2]; o[4] = ob[3]; o[5] = ob[4];
        ref[0] = fsub(refp[0], fmul(0.5f, refp[2])); ref[1] = fsub(refp[1], fmul(0.5f, refp[3]));
        ref[2] = fadd(refp[0], fmul(0.5f, refp[2])); ref[3] = fadd(refp[1], fmul(0.5f, refp[3]));
        if (iou_target) iou_target[k] = aligned_iou(b, ref);
        part[3] += riou::clamped_iou(ob, real, 0);
        part[4] += riou::clamped_iou(r + 1, real, 0);
      } else {
        o[0] = r[0]; o[1] = b[0]; o[2] = b[1]; o[3] = b[2]; o[4] = b[3];
        ref[0] = refp[0]; ref[1] = refp[1]; ref[2] = refp[2]; ref[3] = refp[3];
        if (iou_target) iou_target[k] = aligned_iou(b, ref);
        part[3] += aligned_iou(b, real);
        part[4] += aligned_iou(r + 1, real);
      }
      // DN-DIoU (iou_loss.py:398-466): min over the 3x3 noisy targets, plus the scalar mean base loss
      part[0] += diou_elem(b, ref, eps);
      const float anx = hyper / 2.f, tw = ref[2] - ref[0], th = ref[3] - ref[1];
      float best = 3.0e38f;
#pragma unroll
      for (int i = -1; i <= 1; i++)
#pragma unroll
        for (int j = -1; j <= 1; j++) {
          const float t[4] = {ref[0] - anx * tw * (float)i, ref[1] - anx * th * (float)i,
                              ref[2] + anx * tw * (float)j, ref[3] + anx * th * (float)j};
          best = fminf(best, diou_elem(b, t, eps));
        }
      const float wv = valid[k] ? 1.f : 0.f;
      part[1] += wv;
      part[2] += wv * best;
    }
  }
#pragma unroll
  for (int i = 0; i < 5; i++) part[i] = warp_sum(part[i]);
  block_accumulate(sums, part, 5, red);
}

// Warp per row: cls = H·Wcls^T + b, ins = H·Wins^T + b.  Weights staged in smem as [2C][D].
template <typename TH, int C2>
__global__ void cls_ins_kernel(const TH* __restrict__ H, long long ldh, int D, const float* __restrict__ Wcls,
                               const float* __restrict__ bcls, const float* __restrict__ Wins,
                               const float* __restrict__ bins, int M, float* __restrict__ cls,
                               float* __restrict__ ins) {
  extern __shared__ float wsm[];
  constexpr int C = C2 / 2;
  for (int i = threadIdx.x; i < C * D; i += blockDim.x) { wsm[i] = Wcls[i]; wsm[C * D + i] = Wins[i]; }
  __syncthreads();
  const int lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int kb = (blockIdx.x * nw + (threadIdx.x >> 5)) * HEAD_ROWS; kb < M; kb += gridDim.x * nw * HEAD_ROWS) {
    float o[HEAD_ROWS][C2];
    const int nrows = M - kb < HEAD_ROWS ? M - kb : HEAD_ROWS;
    rows_dot<TH, C2>(H + (size_t)kb * ldh, ldh, nrows, wsm, D, lane, o);
#pragma unroll
    for (int r = 0; r < HEAD_ROWS; r++) {
      // dynamic register indexing is avoided with a select chain
      float v = 0.f, u = 0.f;
#pragma unroll
      for (int c = 0; c < C; c++) { v = (lane == c) ? o[r][C + c] : v; u = (lane == c) ? o[r][c] : u; }
      if (lane < C && r < nrows) {
        cls[(size_t)(kb + r) * C + lane] = u + bcls[lane];
        ins[(size_t)(kb + r) * C + lane] = v + bins[lane];
      }
    }
  }
}

// --------------------------------------------------------------------------------- score + select
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// One CTA per GT.  cls / ins [G*U1*U2, C]; valid [G*U1*U2]; bag_rois [G*U1*U2, 5] (refined bags);
// labels [G] int64; pseudo [G,4].  The GT's score tiles are staged in shared memory with coalesced loads, the
// (U1 group, class) softmax columns are spread over the warps, thread 0 replays the CPU top-k and merges.
constexpr int SS_WARPS = 4;
__global__ void __launch_bounds__(SS_WARPS * 32)
score_select_kernel(const float* __restrict__ cls, const float* __restrict__ ins, const uint8_t* __restrict__ valid,
                    const float* __restrict__ bag_rois, const long long* __restrict__ labels,
                    const float* __restrict__ pseudo, const float* __restrict__ img_wh, int B, int G, int U1, int U2,
                    int C, int topk, float beta, float* __restrict__ merged, float* __restrict__ merged_pts,
                    int* __restrict__ sel_idx, float* __restrict__ sel_score, float* __restrict__ sums, int rotated) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int U = U1 * U2, g = blockIdx.x;
  float* cls_s = sm;                                 // [U][C]
  float* ins_s = cls_s + (size_t)U * C;              // [U][C]
  float* sc = ins_s + (size_t)U * C;                 // [U] selection scores
  int* si = reinterpret_cast<int*>(sc + U);          // [U] companion indices for the replay
  float* red = reinterpret_cast<float*>(si + U);     // [SS_WARPS][2]
  uint8_t* val_s = reinterpret_cast<uint8_t*>(red + 2 * SS_WARPS);   // [U]
  const size_t row0 = (size_t)g * U;
  for (int i = threadIdx.x; i < U * C; i += blockDim.x) { cls_s[i] = cls[row0 * C + i]; ins_s[i] = ins[row0 * C + i]; }
  for (int i = threadIdx.x; i < U; i += blockDim.x) val_s[i] = valid[row0 + i];
  __syncthreads();
  const int label = (int)labels[g];
  float part[2] = {0.f, 0.f};
  for (int p = warp; p < U1 * C; p += SS_WARPS) {
    const int u1 = p / C, c = p - u1 * C;
    const int base = u1 * U2;
    bool any_valid = false;
    for (int u = lane; u < U2; u += 32) any_valid |= val_s[base + u] != 0;
    any_valid = __any_sync(0xffffffffu, any_valid);
    // softmax over the U2 instances of this class column, masked by validity, L1-normalised
    float mx = -3.0e38f;
    for (int u = lane; u < U2; u += 32) mx = fmaxf(mx, ins_s[(base + u) * C + c]);
    mx = warp_max(mx);
    float se = 0.f;
    for (int u = lane; u < U2; u += 32) se += expf(ins_s[(base + u) * C + c] - mx);
    se = warp_sum(se);
    float sv = 0.f;
    for (int u = lane; u < U2; u += 32) sv += val_s[base + u] ? expf(ins_s[(base + u) * C + c] - mx) / se : 0.f;
    sv = warp_sum(sv);
    const float den = fmaxf(sv, 1e-12f);
    float bag = 0.f;
    for (int u = lane; u < U2; u += 32) {
      const float insn = (val_s[base + u] ? expf(ins_s[(base + u) * C + c] - mx) / se : 0.f) / den;
      const float pr = sigmoidf_(cls_s[(base + u) * C + c]);
      bag += pr * insn;
      if (c == label) { sc[base + u] = pr * insn; si[base + u] = base + u; }
    }
    bag = warp_sum(bag);
    // gfocal (fcos_head_p2b_ts.py:1074-1078) against the one-hot label, weight = bag has a valid instance
    const float q = (c == label) ? 1.f : 0.f;
    const float l1 = (bag - q) * (bag - q);
    const float l2 = q * logf(bag + 1e-6f) + (1.f - q) * logf(1.f - bag + 1e-6f);
    if (lane == 0 && any_valid) { part[0] += -(l1 * l2); if (c == 0) part[1] += 1.f; }
  }
  if (lane == 0) { red[warp * 2] = part[0]; red[warp * 2 + 1] = part[1]; }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (sums != nullptr) {
      float a = 0.f, b = 0.f;
      for (int w = 0; w < SS_WARPS; w++) { a += red[w * 2]; b += red[w * 2 + 1]; }
      atomicAdd(sums + S_POS, a);
      atomicAdd(sums + S_NSAMPLE, b);
    }
    // replay of torch.topk(largest=True) on CPU over the flattened U1*U2 axis
    cpu_topk_replay(sc, si, U, topk, /*largest=*/true);
    float wsum = 0.f;
    for (int t = 0; t < topk; t++) wsum += sc[t];
    wsum += 1e-8f;
    const int bd = rotated ? 5 : 4, rs = bd + 1;
    float bx[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    for (int t = 0; t < topk; t++) {
      const float w = sc[t] / wsum;
      const float* r = bag_rois + ((size_t)g * U + si[t]) * rs;
      for (int j = 0; j < bd; j++) bx[j] += r[1 + j] * w;
      sel_idx[(size_t)g * topk + t] = si[t];
      sel_score[(size_t)g * topk + t] = sc[t];
    }
    int bi = (int)bag_rois[(size_t)g * U * rs];
    bi = bi < 0 ? 0 : (bi >= B ? B - 1 : bi);
    const float iw = img_wh[2 * bi], ih = img_wh[2 * bi + 1];
    if (rotated) {
      // rotated_fcos_head_p2rb_ts.py:1211-1212: (cx, cy) both clamped to [0, w] and then to [0, h]
      bx[0] = fminf(fmaxf(fminf(fmaxf(bx[0], 0.f), iw), 0.f), ih);
      bx[1] = fminf(fmaxf(fminf(fmaxf(bx[1], 0.f), iw), 0.f), ih);
    } else {
      bx[0] = fminf(fmaxf(bx[0], 0.f), iw); bx[2] = fminf(fmaxf(bx[2], 0.f), iw);
      bx[1] = fminf(fmaxf(bx[1], 0.f), ih); bx[3] = fminf(fmaxf(bx[3], 0.f), ih);
    }
    if (pseudo != nullptr) {   // (1-beta)*box + beta*coarse   (fcos_head_p2b_ts.py:1109)
      const float* pb = pseudo + (size_t)g * bd;
      for (int j = 0; j < bd; j++) bx[j] = fadd(fmul(1.f - beta, bx[j]), fmul(beta, pb[j]));
    }
    for (int j = 0; j < bd; j++) merged[(size_t)g * bd + j] = bx[j];
    if (merged_pts != nullptr) {  // refined points = box centres (fcos_p2b_teacher_student.py:465; OBB: box[:, :2])
      merged_pts[(size_t)g * 2] = rotated ? bx[0] : fdiv(fadd(bx[0], bx[2]), 2.f);
      merged_pts[(size_t)g * 2 + 1] = rotated ? bx[1] : fdiv(fadd(bx[1], bx[3]), 2.f);
    }
  }
}

// negatives: gfocal(sigmoid(neg_cls), 0, weight)   (fcos_head_p2b_ts.py:1169-1179)
__global__ void neg_loss_kernel(const float* __restrict__ neg_cls, const uint8_t* __restrict__ weight, int n, int C,
                                float* __restrict__ sums) {
  __shared__ float red[32];
  float acc = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n * C; i += gridDim.x * blockDim.x) {
    const int r = i / C;
    if (weight[r]) {
      const float p = sigmoidf_(neg_cls[i]);
      acc += -(p * p * logf(1.f - p + 1e-6f));
    }
  }
  acc = warp_sum(acc);
  float part[1] = {acc};
  block_accumulate(sums + S_NEG, part, 1, red);
}

// out[0] loss_mil_bbox, out[1] loss_mil_bags, out[2] coarse_bags_iou, out[3] refine_bags_iou, out[4] num_sample
__global__ void finalize_losses_kernel(const float* __restrict__ sums, int K, int has_neg, float scale_bbox,
                                       float scale_bags, float pos_w, float neg_w, float* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const float Kf = (float)K;
  const float base_mean = sums[S_BASE] / Kf;
  // DN_DIoULoss: (base_mean + min_bank_k)/2 * w_k summed / avg_factor(K); all-zero weights -> 0
  out[0] = scale_bbox * (sums[S_W] > 0.f ? (base_mean * sums[S_W] + sums[S_WMIN]) / 2.f / Kf : 0.f);
  const float ns = fmaxf(sums[S_NSAMPLE], 1.f);
  out[1] = scale_bags * (pos_w * (sums[S_POS] / ns) + (has_neg ? neg_w * (sums[S_NEG] / ns) : 0.f));
  out[2] = sums[S_COARSE] / Kf;
  out[3] = sums[S_REFINE] / Kf;
  out[4] = ns;
}

// fp32 [M,N] -> bf16 [M,3N] = [hi | lo | hi] (activation operand of the fp32-emulation GEMM)
__global__ void split_bf16x3_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long M, int N) {
  const long long total = M * N;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / N; const int n = (int)(i - m * N);
    const float v = in[i];
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    __nv_bfloat16* row = out + m * 3 * N;
    row[n] = hi; row[N + n] = __float2bfloat16_rn(v - __bfloat162float(hi)); row[2 * N + n] = hi;
  }
}

}  // namespace ptb

using namespace ptb;

extern "C" int pt_split_bf16x3(const float* in, void* out_bf16, long long M, int N, void* stream) {
  if (M <= 0) return PT_OK;
  split_bf16x3_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(in, (__nv_bfloat16*)out_bf16, M, N);
  return check_launch("split_bf16x3_kernel");
}

extern "C" int pt_prep_fc1_weight(const float* w, void* out_bf16, int N, int C, int bins, long long ld, int x3,
                                  void* stream) {
  const long long K = (long long)C * bins;
  if (ld < (x3 ? 3 : 1) * K) { set_error("pt_prep_fc1_weight: ld too small"); return PT_ERR_ARG; }
  if (C % 8 != 0 || ld % 8 != 0) { set_error("pt_prep_fc1_weight: C and ld must be multiples of 8"); return PT_ERR_ARG; }
  const size_t smem = (size_t)(C + 4) * bins * sizeof(float);
  if (smem > 200 * 1024) { set_error("pt_prep_fc1_weight: row of %lld floats does not fit shared memory", K); return PT_ERR_UNSUPPORTED; }
  if (!x3 && (K & 3) == 0 && (size_t)K * sizeof(float) <= 64 * 1024 && (((uintptr_t)w | (uintptr_t)out_bf16) & 15) == 0) {
    static bool done = false;
    if (!done) { cudaFuncSetAttribute(prep_fc1_weight_row_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024); done = true; }
    prep_fc1_weight_row_kernel<<<N, 256, (size_t)K * sizeof(float), (cudaStream_t)stream>>>(w, (__nv_bfloat16*)out_bf16, C, bins, ld);
    return check_launch("prep_fc1_weight_row_kernel");
  }
  cudaFuncSetAttribute(prep_fc1_weight_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  prep_fc1_weight_kernel<<<N, 512, smem, (cudaStream_t)stream>>>(w, (__nv_bfloat16*)out_bf16, C, bins, ld, x3);
  return check_launch("prep_fc1_weight_kernel");
}

extern "C" int pt_cast_weight_bf16(const float* w, void* out_bf16, int N, int K, long long ld, int x3, void* stream) {
  if (ld < (x3 ? 3LL : 1LL) * K) { set_error("pt_cast_weight_bf16: ld too small"); return PT_ERR_ARG; }
  if (!x3 && (K & 3) == 0 && (ld & 3) == 0 && (((uintptr_t)w & 15) | ((uintptr_t)out_bf16 & 7)) == 0) {
    cast_weight_vec_kernel<<<148 * 4, 256, 0, (cudaStream_t)stream>>>(w, (__nv_bfloat16*)out_bf16, N, K / 4, ld);
    return check_launch("cast_weight_vec_kernel");
  }
  cast_weight_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(w, (__nv_bfloat16*)out_bf16, N, K, ld, x3);
  return check_launch("cast_weight_kernel");
}

extern "C" int pt_reg_decode(const void* H, int h_f32, long long ldh, int D, const float* Wreg, const float* breg,
                             const float* bag_rois, const unsigned char* valid, const float* ref_boxes,
                             const float* real_boxes, int U, int K, float max_w, float max_h, float wh_ratio_clip,
                             float hyper, float eps, float* out_rois, float* out_deltas, float* iou_target,
                             float* sums, int rotated, const float* deltas_in, void* stream) {
  if (K <= 0) return PT_OK;
  if (D % 256 != 0) { set_error("pt_reg_decode: hidden width must be a multiple of 256 (got %d)", D); return PT_ERR_ARG; }
  const int threads = 256;
  const size_t smem = (size_t)(4 * D + 8 * 8) * sizeof(float);
  const float max_ratio = fabsf(logf(wh_ratio_clip));
  int grid = (K + 8 * HEAD_ROWS - 1) / (8 * HEAD_ROWS);
  if (grid > 148 * 2) grid = 148 * 2;
  if (h_f32) {
    cudaFuncSetAttribute(reg_decode_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    reg_decode_kernel<float><<<grid, threads, smem, (cudaStream_t)stream>>>(
        (const float*)H, ldh, D, Wreg, breg, bag_rois, valid, ref_boxes, real_boxes, U, K, max_w, max_h, max_ratio,
        hyper, eps, out_rois, out_deltas, iou_target, sums, rotated, deltas_in);
  } else {
    cudaFuncSetAttribute(reg_decode_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    reg_decode_kernel<__nv_bfloat16><<<grid, threads, smem, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)H, ldh, D, Wreg, breg, bag_rois, valid, ref_boxes, real_boxes, U, K, max_w, max_h,
        max_ratio, hyper, eps, out_rois, out_deltas, iou_target, sums, rotated, deltas_in);
  }
  return check_launch("reg_decode_kernel");
}

template <typename TH, int C2>
static int launch_cls_ins(const void* H, long long ldh, int D, const float* Wcls, const float* bcls,
                          const float* Wins, const float* bins, int M, float* cls, float* ins, cudaStream_t s) {
  const size_t smem = (size_t)C2 * D * sizeof(float);
  auto kern = cls_ins_kernel<TH, C2>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int grid = (M + 8 * HEAD_ROWS - 1) / (8 * HEAD_ROWS);
  if (grid > 148) grid = 148;
  kern<<<grid, 256, smem, s>>>((const TH*)H, ldh, D, Wcls, bcls, Wins, bins, M, cls, ins);
  return check_launch("cls_ins_kernel");
}

extern "C" int pt_cls_ins_heads(const void* H, int h_f32, long long ldh, int D, const float* Wcls, const float* bcls,
                                const float* Wins, const float* bins, int C, int M, float* cls, float* ins,
                                void* stream) {
  if (M <= 0) return PT_OK;
  if (D % 256 != 0) { set_error("pt_cls_ins_heads: hidden width must be a multiple of 256"); return PT_ERR_ARG; }
  cudaStream_t s = (cudaStream_t)stream;
#define PT_CASE(CC)                                                                                              \
  if (C == CC)                                                                                                   \
    return h_f32 ? launch_cls_ins<float, 2 * CC>(H, ldh, D, Wcls, bcls, Wins, bins, M, cls, ins, s)              \
                 : launch_cls_ins<__nv_bfloat16, 2 * CC>(H, ldh, D, Wcls, bcls, Wins, bins, M, cls, ins, s)
  PT_CASE(1); PT_CASE(2); PT_CASE(4); PT_CASE(8); PT_CASE(9); PT_CASE(10); PT_CASE(12);
#undef PT_CASE
  set_error("pt_cls_ins_heads: num_classes %d not instantiated (1,2,4,8,9,10,12)", C);
  return PT_ERR_UNSUPPORTED;
}

extern "C" int pt_score_select(const float* cls, const float* ins, const unsigned char* valid, const float* bag_rois,
                               const long long* labels, const float* pseudo, const float* img_wh, int B, int G,
                               int U1, int U2, int C, int topk, float beta, float* merged, float* merged_pts,
                               int* sel_idx, float* sel_score, float* sums, int rotated, void* stream) {
  if (G <= 0) return PT_OK;
  const int U = U1 * U2;
  if (topk < 1 || topk > 16 || topk > U) { set_error("pt_score_select: topk must be in [1, min(16, U)]"); return PT_ERR_ARG; }
  const size_t smem = ((size_t)2 * U * C + 2 * U + 2 * SS_WARPS) * sizeof(float) + ((U + 15) / 16) * 16;
  if (smem > 200 * 1024) { set_error("pt_score_select: bag of %d instances x %d classes does not fit shared memory", U, C); return PT_ERR_UNSUPPORTED; }
  cudaFuncSetAttribute(score_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  score_select_kernel<<<G, SS_WARPS * 32, smem, (cudaStream_t)stream>>>(cls, ins, valid, bag_rois, labels, pseudo,
                                                                        img_wh, B, G, U1, U2, C, topk, beta, merged,
                                                                        merged_pts, sel_idx, sel_score, sums, rotated);
  return check_launch("score_select_kernel");
}

extern "C" int pt_neg_loss(const float* neg_cls, const unsigned char* weight, int n, int C, float* sums, void* stream) {
  if (n <= 0) return PT_OK;
  neg_loss_kernel<<<8, 256, 0, (cudaStream_t)stream>>>(neg_cls, weight, n, C, sums);
  return check_launch("neg_loss_kernel");
}

extern "C" int pt_finalize_losses(const float* sums, int K, int has_neg, float scale_bbox, float scale_bags,
                                  float pos_w, float neg_w, float* out, void* stream) {
  finalize_losses_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(sums, K, has_neg, scale_bbox, scale_bags, pos_w, neg_w,
                                                             out);
  return check_launch("finalize_losses_kernel");
}
