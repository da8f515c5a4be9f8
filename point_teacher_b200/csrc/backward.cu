// Backward of the MIL two-stream head for sm_100a: everything between the two loss scalars and the gradients of
// the head parameters / of the RoI feature operand.  Follows the autograd graph of
//   HBB_TOD/mmdet/models/dense_heads/fcos_head_p2b_ts.py:1147-1236 (mil_bag_training, mil_bag_extensive),
//   HBB_TOD/mmdet/core/bbox/coder/delta_xywh_bbox_coder.py:144-250 (delta2bbox),
//   HBB_TOD/mmdet/models/losses/iou_loss.py:139-190, 398-466 (DIoU, DN-DIoU).
// The big contractions (dgrad / wgrad of the two FCs) reuse the tcgen05 GEMM (gemm_tcgen05.cu) on transposed bf16
// copies made here; the loss gradients are tiny and latency-bound.
//
//   reg_loss_grad_kernel   d loss_mil_bbox / d deltas (K,4): forward-mode dual numbers through delta2bbox ->
//                          clip -> DIoU -> "mean base + min over the 3x3 noisy targets" (argmin carries the gradient)
//   bag_loss_grad_kernel   d loss_mil_bags / d (cls logits, ins logits): sigmoid x softmax(U2) x valid x L1-norm ->
//                          bag score -> gfocal; negatives: sigmoid -> gfocal against 0
//   head_bwd_kernel<NOUT>  small heads (fc_reg: 4, fc_cls + fc_ins: 2C): dZ = (g . W) masked by the ReLU of the
//                          hidden activation (bf16, the dgrad GEMM operand), dW += g^T H, db += sum g
//   transpose_pad_bf16     [R, C] -> [C, Rpad] (zero padded): turns wgrad into the K-major x K-major form
//   unpermute_dw1          bin-major FC1 weight gradient -> the reference's (c*49 + bin) column order
//   colsum_bf16            bias gradients of the two FCs
//   roi_align_bwd_kernel   d feature map (NHWC fp32, red.global.add.v4.f32) from d RoI features (bf16 bin-major)
#include "common.cuh"

namespace ptb {

namespace ramma {   // roi_align_mma.cu: tensor-core RoIAlign backward
bool bwd_supported(int C, int H, int W, long long ld);
int launch_bwd(const void* dA_bf16, long long ld, const float* rois, int K, int B, int C, int H, int W, float scale,
               int sampling_ratio, int aligned, float* dfeat, const int* roi_level, int level, cudaStream_t stream);
}  // namespace ramma

// ---------------------------------------------------------------------------------------- dual numbers (4 seeds)
struct D4 {
  float v, d[4];
};
__device__ __forceinline__ D4 dconst(float v) { D4 r; r.v = v; r.d[0] = r.d[1] = r.d[2] = r.d[3] = 0.f; return r; }
__device__ __forceinline__ D4 dvar(float v, int i) { D4 r = dconst(v); r.d[i] = 1.f; return r; }
__device__ __forceinline__ D4 operator+(const D4& a, const D4& b) { D4 r; r.v = a.v + b.v; for (int i = 0; i < 4; i++) r.d[i] = a.d[i] + b.d[i]; return r; }
__device__ __forceinline__ D4 operator-(const D4& a, const D4& b) { D4 r; r.v = a.v - b.v; for (int i = 0; i < 4; i++) r.d[i] = a.d[i] - b.d[i]; return r; }
__device__ __forceinline__ D4 operator*(const D4& a, const D4& b) { D4 r; r.v = a.v * b.v; for (int i = 0; i < 4; i++) r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
__device__ __forceinline__ D4 operator/(const D4& a, const D4& b) {
  D4 r; r.v = a.v / b.v;
  const float inv = 1.f / b.v;
  for (int i = 0; i < 4; i++) r.d[i] = (a.d[i] - r.v * b.d[i]) * inv;
  return r;
}
__device__ __forceinline__ D4 operator*(const D4& a, float s) { D4 r; r.v = a.v * s; for (int i = 0; i < 4; i++) r.d[i] = a.d[i] * s; return r; }
__device__ __forceinline__ D4 operator+(const D4& a, float s) { D4 r = a; r.v += s; return r; }
__device__ __forceinline__ D4 dmax(const D4& a, const D4& b) { return a.v >= b.v ? a : b; }
__device__ __forceinline__ D4 dmin(const D4& a, const D4& b) { return a.v <= b.v ? a : b; }
__device__ __forceinline__ D4 dmaxc(const D4& a, float c) { return a.v >= c ? a : dconst(c); }   // clamp(min=c) / where
__device__ __forceinline__ D4 dminc(const D4& a, float c) { return a.v <= c ? a : dconst(c); }

__device__ __forceinline__ D4 diou_dual(const D4* p, const float* t, float eps) {
  const D4 ow = dmaxc(dminc(p[2], t[2]) - dmaxc(p[0], t[0]), 0.f);
  const D4 oh = dmaxc(dminc(p[3], t[3]) - dmaxc(p[1], t[1]), 0.f);
  const D4 ov = ow * oh;
  const D4 ap = (p[2] - p[0]) * (p[3] - p[1]);
  const float ag = (t[2] - t[0]) * (t[3] - t[1]);
  const D4 iou = ov / (ap + ag - ov + eps);
  const D4 cw = dmaxc(dmaxc(p[2], t[2]) - dminc(p[0], t[0]), 0.f);
  const D4 ch = dmaxc(dmaxc(p[3], t[3]) - dminc(p[1], t[1]), 0.f);
  const D4 c2 = cw * cw + ch * ch + eps;
  const D4 dx = dconst(t[0] + t[2]) - (p[0] + p[2]), dy = dconst(t[1] + t[3]) - (p[1] + p[3]);
  const D4 rho2 = dx * dx * 0.25f + dy * dy * 0.25f;
  return dconst(1.f) - (iou - rho2 / c2);
}

// sums layout of the forward (mil_head.cu): 1 = sum of weights, 6 = num_sample
// g[k][0..3] = gscale * scale * d [ sum_j w_j (base + min_j) / 2 / K ] / d deltas_k
__global__ void reg_loss_grad_kernel(const float* __restrict__ deltas, const float* __restrict__ bag_rois,
                                     const uint8_t* __restrict__ valid, const float* __restrict__ ref_boxes, int U,
                                     int K, float max_w, float max_h, float max_ratio, float hyper, float eps,
                                     const float* __restrict__ sums, const float* __restrict__ gscale, float scale,
                                     float* __restrict__ g, int rotated) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  const float* r = bag_rois + (size_t)k * (rotated ? 6 : 5);
  const float* dl = deltas + (size_t)k * 4;
  // rotated (rotated_fcos_head_p2rb_ts.py:1314-1320): the decode runs on cxcywh_to_xyxy(bag[:, :4])
  float rx1 = r[1], ry1 = r[2], rx2 = r[3], ry2 = r[4];
  if (rotated) {
    rx1 = fsub(r[1], fmul(0.5f, r[3])); ry1 = fsub(r[2], fmul(0.5f, r[4]));
    rx2 = fadd(r[1], fmul(0.5f, r[3])); ry2 = fadd(r[2], fmul(0.5f, r[4]));
  }
  const float px = (rx1 + rx2) * 0.5f, py = (ry1 + ry2) * 0.5f, pw = rx2 - rx1, ph = ry2 - ry1;
  const D4 dx = dvar(dl[0], 0), dy = dvar(dl[1], 1);
  D4 dw = dvar(dl[2], 2), dh = dvar(dl[3], 3);
  // clamp(min=-mr, max=mr): zero gradient outside
  if (dw.v < -max_ratio || dw.v > max_ratio) dw = dconst(fminf(fmaxf(dw.v, -max_ratio), max_ratio));
  if (dh.v < -max_ratio || dh.v > max_ratio) dh = dconst(fminf(fmaxf(dh.v, -max_ratio), max_ratio));
  D4 ew = dw, eh = dh;
  ew.v = expf(dw.v); for (int i = 0; i < 4; i++) ew.d[i] = dw.d[i] * ew.v;
  eh.v = expf(dh.v); for (int i = 0; i < 4; i++) eh.d[i] = dh.d[i] * eh.v;
  const D4 gw = ew * pw, gh = eh * ph;
  const D4 gx = dx * pw + px, gy = dy * ph + py;
  D4 b[4] = {gx - gw * 0.5f, gy - gh * 0.5f, gx + gw * 0.5f, gy + gh * 0.5f};
  // torch.where(out < 0, 0, out); torch.where(out > hi, hi, out): clipped coordinates carry no gradient
  const float hi[4] = {max_w, max_h, max_w, max_h};
  for (int i = 0; i < 4; i++) {
    if (b[i].v < 0.f) b[i] = dconst(0.f);
    if (b[i].v > hi[i]) b[i] = dconst(hi[i]);
  }
  const float* refp = ref_boxes + (size_t)(k / U) * (rotated ? 5 : 4);
  float ref[4] = {refp[0], refp[1], refp[2], refp[3]};
  if (rotated) {
    ref[0] = fsub(refp[0], fmul(0.5f, refp[2])); ref[1] = fsub(refp[1], fmul(0.5f, refp[3]));
    ref[2] = fadd(refp[0], fmul(0.5f, refp[2])); ref[3] = fadd(refp[1], fmul(0.5f, refp[3]));
  }
  const D4 base = diou_dual(b, ref, eps);
  const float anx = hyper / 2.f, tw = ref[2] - ref[0], th = ref[3] - ref[1];
  D4 best = dconst(3.0e38f);
  for (int i = -1; i <= 1; i++)
    for (int j = -1; j <= 1; j++) {
      const float t[4] = {ref[0] - anx * tw * (float)i, ref[1] - anx * th * (float)i, ref[2] + anx * tw * (float)j,
                          ref[3] + anx * th * (float)j};
      const D4 e = diou_dual(b, t, eps);
      if (e.v < best.v) best = e;                // torch.min(dim) returns (and differentiates) the first minimum
    }
  const float wk = valid[k] ? 1.f : 0.f;
  const float sw = sums[1];
  const float c = gscale[0] * scale / (2.f * (float)K);
  for (int i = 0; i < 4; i++) g[(size_t)k * 4 + i] = c * (sw / (float)K * base.d[i] + wk * best.d[i]);
}

__device__ __forceinline__ float sigm(float x) { return 1.f / (1.f + expf(-x)); }

// One CTA per (GT, U1 group) bag; thread = (instance u, class c) pairs strided.  g [M, 2C]: cls grads then ins grads.
__global__ void bag_loss_grad_kernel(const float* __restrict__ cls, const float* __restrict__ ins,
                                     const uint8_t* __restrict__ valid, const long long* __restrict__ labels, int U1,
                                     int U2, int C, const float* __restrict__ sums, const float* __restrict__ gscale,
                                     float scale, float eps, float* __restrict__ g) {
  extern __shared__ float sm[];           // [U2*C] e (softmax), [U2*C] s (sigmoid), [C] Z, [C] b, [C] gb, [C] dot
  float* e = sm;
  float* s = e + U2 * C;
  float* Zc = s + U2 * C;
  float* bc = Zc + C;
  float* gb = bc + C;
  float* dot = gb + C;
  __shared__ int any_valid;
  const int bag = blockIdx.x;             // g * U1 + u1
  const size_t row0 = (size_t)bag * U2;
  const int lab = (int)labels[bag / U1];
  if (threadIdx.x == 0) any_valid = 0;
  __syncthreads();
  // per class: softmax over U2
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float mx = -3.0e38f;
    for (int u = 0; u < U2; u++) mx = fmaxf(mx, ins[(row0 + u) * C + c]);
    float se = 0.f;
    for (int u = 0; u < U2; u++) { const float t = expf(ins[(row0 + u) * C + c] - mx); e[u * C + c] = t; se += t; }
    float z = 0.f, b = 0.f;
    for (int u = 0; u < U2; u++) {
      const float ev = e[u * C + c] / se;
      e[u * C + c] = ev;
      const float sv = sigm(cls[(row0 + u) * C + c]);
      s[u * C + c] = sv;
      const float m = valid[row0 + u] ? ev : 0.f;
      z += m;
    }
    const float Z = fmaxf(z, 1e-12f);
    for (int u = 0; u < U2; u++) b += s[u * C + c] * ((valid[row0 + u] ? e[u * C + c] : 0.f) / Z);
    Zc[c] = Z; bc[c] = b;
    if (z > 0.f) any_valid = 1;
  }
  for (int u = threadIdx.x; u < U2; u += blockDim.x)
    if (valid[row0 + u]) any_valid = 1;
  __syncthreads();
  const float w = any_valid ? 1.f : 0.f;
  const float ns = fmaxf(sums[6], 1.f);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float b = bc[c], q = (c == lab) ? 1.f : 0.f;
    const float l2 = q * logf(b + eps) + (1.f - q) * logf(1.f - b + eps);
    const float dl = 2.f * (b - q) * l2 + (b - q) * (b - q) * (q / (b + eps) - (1.f - q) / (1.f - b + eps));
    gb[c] = -w / ns * dl * gscale[0] * scale;
    // sum_v e_v t_v with t_v = valid_v (s_v - b) / Z
    float d = 0.f;
    for (int u = 0; u < U2; u++) d += e[u * C + c] * (valid[row0 + u] ? (s[u * C + c] - b) / Zc[c] : 0.f);
    dot[c] = d;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < U2 * C; i += blockDim.x) {
    const int u = i / C, c = i - u * C;
    const bool v = valid[row0 + u] != 0;
    const float sv = s[i], ev = e[i];
    const float n = v ? ev / Zc[c] : 0.f;
    const float t = v ? (sv - bc[c]) / Zc[c] : 0.f;
    g[(row0 + u) * 2 * C + c] = gb[c] * n * sv * (1.f - sv);
    g[(row0 + u) * 2 * C + C + c] = gb[c] * ev * (t - dot[c]);
  }
}

// negatives: rows [K, K+n): p = sigmoid(z), loss = -sum p^2 log(1 - p + eps) * w / num_sample
__global__ void neg_loss_grad_kernel(const float* __restrict__ neg_cls, const uint8_t* __restrict__ weight, int n, int C,
                                     const float* __restrict__ sums, const float* __restrict__ gscale, float scale,
                                     float eps, float* __restrict__ g) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * C) return;
  const int r = i / C, c = i - r * C;
  const float p = sigm(neg_cls[i]);
  const float wv = weight[r] ? 1.f : 0.f;
  const float ns = fmaxf(sums[6], 1.f);
  const float dl = 2.f * p * logf(1.f - p + eps) - p * p / (1.f - p + eps);
  g[(size_t)r * 2 * C + c] = -wv / ns * dl * p * (1.f - p) * gscale[0] * scale;
  g[(size_t)r * 2 * C + C + c] = 0.f;
}

// ---------------------------------------------------------------------------------------- small-head backward
// g [M, NOUT] fp32, H [M, D] bf16 (post-ReLU hidden), W [NOUT, D] fp32 -> dZ [M, D] bf16 = (g W) * (H > 0);
// dW [NOUT, D] += g^T H; db [NOUT] += sum_m g.  Block = 256 threads, thread owns 4 consecutive hidden units of a
// 1024-wide column tile; a block walks a contiguous range of rows.
template <int NOUT>
__global__ void __launch_bounds__(256, 2)
head_bwd_kernel(const float* __restrict__ g, const __nv_bfloat16* __restrict__ H, long long ldh, int D,
                const float* __restrict__ W, int M, __nv_bfloat16* __restrict__ dZ, long long ldz,
                float* __restrict__ dW, float* __restrict__ db) {
  // thread = 2 consecutive hidden units of a 512-wide column slab (blockIdx.y), block = a contiguous range of rows;
  // two rows per iteration so that two independent H loads are in flight.  (Round 1 owned 4 units per thread:
  // 167 registers, one CTA per SM, 296 CTAs = two serial waves of latency-bound row walks: 34 us at 5400 x 16.)
  const int rows_per = (M + gridDim.x - 1) / gridDim.x;
  const int r0 = blockIdx.x * rows_per, r1 = min(M, r0 + rows_per);
  if (r0 >= r1) return;
  for (int c0 = blockIdx.y * 512 + threadIdx.x * 2; c0 < D; c0 += gridDim.y * 512) {
    float w[NOUT][2], acc[NOUT][2];
#pragma unroll
    for (int o = 0; o < NOUT; o++) {
      const float2 t = *reinterpret_cast<const float2*>(W + (size_t)o * D + c0);
      w[o][0] = t.x; w[o][1] = t.y;
      acc[o][0] = 0.f; acc[o][1] = 0.f;
    }
    for (int r = r0; r < r1; r += 2) {
      const bool two = r + 1 < r1;
      const uint32_t hv0 = *reinterpret_cast<const uint32_t*>(H + (size_t)r * ldh + c0);
      const uint32_t hv1 = two ? *reinterpret_cast<const uint32_t*>(H + (size_t)(r + 1) * ldh + c0) : 0u;
      const float h0[2] = {__uint_as_float(hv0 << 16), __uint_as_float(hv0 & 0xffff0000u)};
      const float h1[2] = {__uint_as_float(hv1 << 16), __uint_as_float(hv1 & 0xffff0000u)};
      float dz0[2] = {0.f, 0.f}, dz1[2] = {0.f, 0.f};
#pragma unroll
      for (int o = 0; o < NOUT; o++) {
        const float g0 = __ldg(g + (size_t)r * NOUT + o);
        const float g1 = two ? __ldg(g + (size_t)(r + 1) * NOUT + o) : 0.f;
#pragma unroll
        for (int j = 0; j < 2; j++) {
          dz0[j] += g0 * w[o][j]; dz1[j] += g1 * w[o][j];
          acc[o][j] += g0 * h0[j] + g1 * h1[j];
        }
      }
      *reinterpret_cast<uint32_t*>(dZ + (size_t)r * ldz + c0) = pack_bf16(h0[0] > 0.f ? dz0[0] : 0.f, h0[1] > 0.f ? dz0[1] : 0.f);
      if (two)
        *reinterpret_cast<uint32_t*>(dZ + (size_t)(r + 1) * ldz + c0) =
            pack_bf16(h1[0] > 0.f ? dz1[0] : 0.f, h1[1] > 0.f ? dz1[1] : 0.f);
    }
#pragma unroll
    for (int o = 0; o < NOUT; o++)
      asm volatile("red.relaxed.gpu.global.add.v2.f32 [%0], {%1, %2};" ::"l"(dW + (size_t)o * D + c0), "f"(acc[o][0]),
                   "f"(acc[o][1]) : "memory");
  }
  if (blockIdx.y == 0 && threadIdx.x < NOUT) {
    float s = 0.f;
    for (int r = r0; r < r1; r++) s += g[(size_t)r * NOUT + threadIdx.x];
    atomicAdd(db + threadIdx.x, s);
  }
}

// ---------------------------------------------------------------------------------------- layout helpers
// in [R, ldin] bf16 (C columns used) -> out [C, ldout] bf16 with out[c][r] = in[r][c]; columns r in [R, ldout) = 0
__global__ void __launch_bounds__(256)
transpose_pad_bf16_kernel(const __nv_bfloat16* __restrict__ in, long long ldin, int R, int C,
                          __nv_bfloat16* __restrict__ out, long long ldout) {
  __shared__ __nv_bfloat16 tile[64][66];
  const int r0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;     // 64 x 4
  for (int i = ty; i < 64; i += 4) {
    const int r = r0 + i, c = c0 + tx;
    tile[i][tx] = (r < R && c < C) ? in[(size_t)r * ldin + c] : __float2bfloat16_rn(0.f);
  }
  __syncthreads();
  for (int i = ty; i < 64; i += 4) {
    const int c = c0 + i, r = r0 + tx;
    if (c < C && r < ldout) out[(size_t)c * ldout + r] = tile[tx][i];
  }
}

// dW1 (bin-major columns bin*C + c, fp32 [N, C*bins]) -> the parameter's order c*bins + bin, accumulated into grad.
// Block = (row n, slab of UNP_CH channels): reads `bins` segments of UNP_CH floats, writes UNP_CH*bins contiguous floats.
constexpr int UNP_CH = 64;
__global__ void __launch_bounds__(256)
unpermute_dw1_kernel(const float* __restrict__ dwp, int C, int bins, float* __restrict__ grad, int accumulate) {
  extern __shared__ float slab[];                       // [bins][UNP_CH + 1]
  const int n = blockIdx.x, c0 = blockIdx.y * UNP_CH;
  const int nc = min(UNP_CH, C - c0);
  const float* src = dwp + (size_t)n * C * bins + c0;
  for (int i = threadIdx.x; i < bins * UNP_CH; i += blockDim.x) {
    const int b = i / UNP_CH, c = i - b * UNP_CH;
    if (c < nc) slab[b * (UNP_CH + 1) + c] = src[(size_t)b * C + c];
  }
  __syncthreads();
  float* dst = grad + (size_t)n * C * bins + (size_t)c0 * bins;
  for (int k = threadIdx.x; k < nc * bins; k += blockDim.x) {
    const int c = k / bins, b = k - c * bins;
    const float v = slab[b * (UNP_CH + 1) + c];
    dst[k] = accumulate ? dst[k] + v : v;
  }
}

// Whole-row variant (C * bins * 4 B <= 64 KB, i.e. the shipped 256 x 49 = 50 KB): the bin-major row is one contiguous
// 16-byte-vector read, lands in shared memory at its parameter position (stride `bins` between consecutive channels:
// odd -> bank-conflict-free scalar stores) and leaves as one contiguous 16-byte-vector write.  4 CTAs / SM.
__global__ void __launch_bounds__(256)
unpermute_dw1_row_kernel(const float* __restrict__ dwp, int C, int bins, float* __restrict__ grad, int accumulate) {
  extern __shared__ __align__(16) float row[];          // [C * bins] in parameter order
  const int n = blockIdx.x, L = C * bins;
  const float4* src = reinterpret_cast<const float4*>(dwp + (size_t)n * L);
  for (int i = threadIdx.x; i < L / 4; i += blockDim.x) {
    const float4 v = __ldcs(src + i);                   // streamed: read exactly once
    const int j = i * 4, b = j / C, c = j - b * C;      // C % 4 == 0: the four values share the bin
    row[(c + 0) * bins + b] = v.x; row[(c + 1) * bins + b] = v.y;
    row[(c + 2) * bins + b] = v.z; row[(c + 3) * bins + b] = v.w;
  }
  __syncthreads();
  float4* dst = reinterpret_cast<float4*>(grad + (size_t)n * L);
  const float4* r4 = reinterpret_cast<const float4*>(row);
  for (int i = threadIdx.x; i < L / 4; i += blockDim.x) {
    float4 v = r4[i];
    if (accumulate) { const float4 o = dst[i]; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
    dst[i] = v;
  }
}

// db[n] (+)= sum_m dZ[m][n], dZ bf16 [M, ld].  Block = 32 column groups (8 columns = one 16-byte load each) x 8 row
// lanes; a block covers 256 columns and a contiguous slab of rows; partials meet in shared memory, one atomic per column.
__global__ void __launch_bounds__(256)
colsum_bf16_kernel(const __nv_bfloat16* __restrict__ dZ, long long ld, int M, int N, float* __restrict__ db) {
  __shared__ float part[8][256 + 8];
  const int cg = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int n0 = blockIdx.x * 256 + cg * 8;
  const int rows_per = (M + gridDim.y - 1) / gridDim.y;
  const int r0 = blockIdx.y * rows_per, r1 = min(M, r0 + rows_per);
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (n0 < N) {
#pragma unroll 4
    for (int r = r0 + rl; r < r1; r += 8) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(dZ + (size_t)r * ld + n0));
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int q = 0; q < 4; q++) { s[2 * q] += __uint_as_float(w[q] << 16); s[2 * q + 1] += __uint_as_float(w[q] & 0xffff0000u); }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; j++) part[rl][cg * 8 + j] = s[j];
  __syncthreads();
  const int n = blockIdx.x * 256 + threadIdx.x;
  if (n < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) t += part[i][threadIdx.x];
    atomicAdd(db + n, t);
  }
}

__global__ void __launch_bounds__(256)
nhwc_to_nchw_f32_kernel(const float* __restrict__ in, float* __restrict__ out, int C, int HW, int accumulate) {
  __shared__ float tile[64][65];
  const int b = blockIdx.z, p0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
  const float* src = in + (size_t)b * C * HW;
  float* dst = out + (size_t)b * C * HW;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  for (int i = ty; i < 64; i += 4) {
    const int p = p0 + i, c = c0 + tx;
    tile[i][tx] = (p < HW && c < C) ? src[(size_t)p * C + c] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 64; i += 4) {
    const int c = c0 + i, p = p0 + tx;
    if (c < C && p < HW) {
      const size_t o = (size_t)c * HW + p;
      dst[o] = accumulate ? dst[o] + tile[tx][i] : tile[tx][i];
    }
  }
}

// ---------------------------------------------------------------------------------------- RoIAlign backward
// One warp per RoI, lane = 8 channels (C = 256).  The patch gradient dP[pix] = sum_bins Wmat[bin][pix] dA[bin] is
// accumulated in registers for one 4x4-pixel chunk at a time with the same separable weights as the forward
// (one axis of the Detectron2 bilinear rule per table), then added to the NHWC fp32 map with 16-byte reductions.
__device__ __forceinline__ bool axis_setup_b(float v, int size, int& lo, int& hi, float& l, float& h) {
  if (v < -1.0f || v > (float)size) return false;
  if (v <= 0.f) v = 0.f;
  lo = (int)v;
  if (lo >= size - 1) { hi = lo = size - 1; v = (float)lo; } else hi = lo + 1;
  l = fsub(v, (float)lo);
  h = fsub(1.0f, l);
  return true;
}

constexpr int RB_WARPS = 4;

__global__ void __launch_bounds__(RB_WARPS * 32)
roi_align_bwd_kernel(const __nv_bfloat16* __restrict__ dA, long long ld, const float* __restrict__ rois, int K, int B,
                     int C, int H, int W, float scale, int sampling_ratio, int aligned, float* __restrict__ dfeat,
                     const int* __restrict__ roi_level, int level) {
  extern __shared__ float tabs[];     // per warp: wx[(W+4)*8] , wy[(H+4)*8]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* wx = tabs + (size_t)warp * ((W + 4) + (H + 4)) * 8;
  float* wy = wx + (W + 4) * 8;
  const float off = aligned ? 0.5f : 0.f;
  for (int roi = blockIdx.x * RB_WARPS + warp; roi < K; roi += gridDim.x * RB_WARPS) {
    const float* r = rois + (size_t)roi * 5;
    const int b = (int)r[0];
    if (b < 0 || b >= B) continue;
    if (roi_level != nullptr && roi_level[roi] != level) continue;     // multi-level FPN: another level's RoI
    const float x1 = fsub(fmul(r[1], scale), off), y1 = fsub(fmul(r[2], scale), off);
    const float x2 = fsub(fmul(r[3], scale), off), y2 = fsub(fmul(r[4], scale), off);
    float rw = fsub(x2, x1), rh = fsub(y2, y1);
    if (!aligned) { rw = fmaxf(rw, 1.f); rh = fmaxf(rh, 1.f); }
    const bool isx = (lane & 8) == 0;
    const int bi = lane & 7;
    const float bin = fdiv(isx ? rw : rh, 7.f);
    const float bin_o = __shfl_xor_sync(0xffffffffu, bin, 8);
    const int gs = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(bin);
    const int gs_o = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(bin_o);
    const int cnt = gs * gs_o > 1 ? gs * gs_o : 1;
    const float inv_count = 1.0f / (float)cnt;
    const float start = isx ? x1 : y1;
    const int size = isx ? W : H;
    float* tab = isx ? wx : wy;
    const float base = fadd(start, fmul((float)bi, bin));
    const bool owner = lane < 16 && bi < 7;
    int lo = 1 << 30, hi = -1;
    if (owner)
      for (int i = 0; i < gs; i++) {
        const float v = fadd(base, fdiv(fmul((float)i + .5f, bin), (float)gs));
        int l, h; float fl, fh;
        if (axis_setup_b(v, size, l, h, fl, fh)) { lo = min(lo, l); hi = max(hi, h); }
      }
    int glo = lo, ghi = hi;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      glo = min(glo, __shfl_xor_sync(0xffffffffu, glo, o));
      ghi = max(ghi, __shfl_xor_sync(0xffffffffu, ghi, o));
    }
    if (ghi < 0) { glo = 0; ghi = -1; }
    __syncwarp();
    if (owner) {
      const int npad = (ghi - glo + 4) & ~3;
      for (int c = 0; c < npad; c++) tab[c * 8 + bi] = 0.f;
      for (int i = 0; i < gs; i++) {
        const float v = fadd(base, fdiv(fmul((float)i + .5f, bin), (float)gs));
        int l, h; float fl, fh;
        if (axis_setup_b(v, size, l, h, fl, fh)) { tab[(l - glo) * 8 + bi] += fh; tab[(h - glo) * 8 + bi] += fl; }
      }
    }
    __syncwarp();
    const int xmin = __shfl_sync(0xffffffffu, glo, 0), xmax = __shfl_sync(0xffffffffu, ghi, 0);
    const int ymin = __shfl_sync(0xffffffffu, glo, 8), ymax = __shfl_sync(0xffffffffu, ghi, 8);
    if (xmax < xmin || ymax < ymin) continue;
    const __nv_bfloat16* drow = dA + (size_t)roi * ld;
    float* fb = dfeat + (size_t)b * H * W * C;
    for (int c0 = lane * 8; c0 < C; c0 += 256) {
      for (int cy = ymin; cy <= ymax; cy += 4) {
        for (int cx = xmin; cx <= xmax; cx += 4) {
          float acc[16][8];
#pragma unroll
          for (int p = 0; p < 16; p++)
#pragma unroll
            for (int j = 0; j < 8; j++) acc[p][j] = 0.f;
          for (int ph = 0; ph < 7; ph++) {
            float wyv[4];
#pragma unroll
            for (int q = 0; q < 4; q++) wyv[q] = wy[(cy - ymin + q) * 8 + ph] * inv_count;
            if (wyv[0] == 0.f && wyv[1] == 0.f && wyv[2] == 0.f && wyv[3] == 0.f) continue;
            // the 7 bins of this output row are fetched together (7 x 16 B in flight per lane) before any is used
            uint4 u7[7];
#pragma unroll
            for (int pw = 0; pw < 7; pw++)
              u7[pw] = __ldg(reinterpret_cast<const uint4*>(drow + (size_t)(ph * 7 + pw) * C + c0));
            // separable accumulation: T[px] = sum_pw wx[px][pw] * g[pw] (one output row), then acc[py][px] += wy[py] * T[px]
            // -- ~180 FMAs per output row instead of 49 x (16 tests + 16 products + ~32 FMAs) in the per-pixel form
            float T[4][8];
#pragma unroll
            for (int px = 0; px < 4; px++)
#pragma unroll
              for (int j = 0; j < 8; j++) T[px][j] = 0.f;
#pragma unroll
            for (int pw = 0; pw < 7; pw++) {
              float wxv[4];
#pragma unroll
              for (int q = 0; q < 4; q++) wxv[q] = wx[(cx - xmin + q) * 8 + pw];
              if (wxv[0] == 0.f && wxv[1] == 0.f && wxv[2] == 0.f && wxv[3] == 0.f) continue;
              const uint32_t wv[4] = {u7[pw].x, u7[pw].y, u7[pw].z, u7[pw].w};
              float d[8];
#pragma unroll
              for (int q = 0; q < 4; q++) { d[2 * q] = __uint_as_float(wv[q] << 16); d[2 * q + 1] = __uint_as_float(wv[q] & 0xffff0000u); }
#pragma unroll
              for (int px = 0; px < 4; px++) {
                if (wxv[px] == 0.f) continue;        // warp-uniform: a bin of a tiny RoI touches 2 of the 4 columns
#pragma unroll
                for (int j = 0; j < 8; j++) T[px][j] = fmaf(wxv[px], d[j], T[px][j]);
              }
            }
#pragma unroll
            for (int py = 0; py < 4; py++) {
              if (wyv[py] == 0.f) continue;
#pragma unroll
              for (int px = 0; px < 4; px++)
#pragma unroll
                for (int j = 0; j < 8; j++) acc[py * 4 + px][j] = fmaf(wyv[py], T[px][j], acc[py * 4 + px][j]);
            }
          }
#pragma unroll
          for (int py = 0; py < 4; py++)
#pragma unroll
            for (int px = 0; px < 4; px++) {
              const int yy = cy + py, xx = cx + px;
              if (yy > ymax || xx > xmax) continue;
              float* dst = fb + ((size_t)yy * W + xx) * C + c0;
              const float* a = acc[py * 4 + px];
              asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(a[0]), "f"(a[1]),
                           "f"(a[2]), "f"(a[3]) : "memory");
              asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4), "f"(a[4]), "f"(a[5]),
                           "f"(a[6]), "f"(a[7]) : "memory");
            }
        }
      }
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------- RoIAlignRotated backward
// One warp per RoI, lane = 8 channels.  The samples of a (tiny) rotated RoI fall in a small pixel patch: per 4x4-pixel
// chunk the warp first builds Wmat[bin][pixel] (sum of the bilinear tap weights of that bin's sampling grid that land
// on the pixel; lane b owns bins b, b+32 -> no conflicts) in shared memory with the forward's exact coordinate
// arithmetic (roi_align.cu: roi_align_rotated_fwd_kernel), then accumulates dP[pixel] = sum_bins Wmat[bin][pixel] dA[bin]
// in registers and adds it to the NHWC fp32 map with 16-byte reductions.
__global__ void __launch_bounds__(RB_WARPS * 32)
roi_align_rotated_bwd_kernel(const __nv_bfloat16* __restrict__ dA, long long ld, const float* __restrict__ rois, int K,
                             int B, int C, int H, int W, float scale, int sampling_ratio, int aligned, int clockwise,
                             float* __restrict__ dfeat, const int* __restrict__ roi_level, int level) {
  __shared__ __align__(16) float wmat_all[RB_WARPS][49 * 16];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* wmat = wmat_all[warp];
  const float off = aligned ? 0.5f : 0.f;
  for (int roi = blockIdx.x * RB_WARPS + warp; roi < K; roi += gridDim.x * RB_WARPS) {
    const float* r = rois + (size_t)roi * 6;
    const int b = (int)__ldg(r);
    if (b < 0 || b >= B) continue;
    if (roi_level != nullptr && roi_level[roi] != level) continue;
    const float cx = fsub(fmul(__ldg(r + 1), scale), off), cy = fsub(fmul(__ldg(r + 2), scale), off);
    float rw = fmul(__ldg(r + 3), scale), rh = fmul(__ldg(r + 4), scale);
    float theta = __ldg(r + 5);
    if (clockwise) theta = -theta;
    if (!aligned) { rw = fmaxf(rw, 1.f); rh = fmaxf(rh, 1.f); }
    const float bh = fdiv(rh, 7.f), bw = fdiv(rw, 7.f);
    const int gh = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(bh);
    const int gw = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(bw);
    const float sh = fdiv(-rh, 2.0f), sw = fdiv(-rw, 2.0f);
    const float ct = cosf(theta), st = sinf(theta);
    const int cnt = gh * gw > 1 ? gh * gw : 1;
    const float inv_count = 1.0f / (float)cnt;
    // pixel bounding box of all taps: the sample coordinates are affine in (yy, xx), extremes at the four corner samples
    float xlo = 3.0e38f, xhi = -3.0e38f, ylo = 3.0e38f, yhi = -3.0e38f;
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int ph = (q & 1) ? 6 : 0, iy = (q & 1) ? gh - 1 : 0, pw = (q & 2) ? 6 : 0, ix = (q & 2) ? gw - 1 : 0;
      const float yy = fadd(fadd(sh, fmul((float)ph, bh)), fdiv(fmul((float)iy + .5f, bh), (float)gh));
      const float xx = fadd(fadd(sw, fmul((float)pw, bw)), fdiv(fmul((float)ix + .5f, bw), (float)gw));
      const float y = fadd(fsub(fmul(yy, ct), fmul(xx, st)), cy);
      const float x = fadd(fadd(fmul(yy, st), fmul(xx, ct)), cx);
      xlo = fminf(xlo, x); xhi = fmaxf(xhi, x); ylo = fminf(ylo, y); yhi = fmaxf(yhi, y);
    }
    // one pixel of slack on each side absorbs the rounding of the interior samples relative to the corners
    const int xmin = max(0, (int)floorf(xlo) - 1), xmax = min(W - 1, (int)floorf(xhi) + 2);
    const int ymin = max(0, (int)floorf(ylo) - 1), ymax = min(H - 1, (int)floorf(yhi) + 2);
    if (xmax < xmin || ymax < ymin || gh <= 0 || gw <= 0) continue;
    const __nv_bfloat16* drow = dA + (size_t)roi * ld;
    float* fb = dfeat + (size_t)b * H * W * C;
    for (int cy0 = ymin; cy0 <= ymax; cy0 += 4) {
      for (int cx0 = xmin; cx0 <= xmax; cx0 += 4) {
        __syncwarp();
        int any = 0;
        for (int bin = lane; bin < 49; bin += 32) {
          float wr[16];
#pragma unroll
          for (int i = 0; i < 16; i++) wr[i] = 0.f;
          const int ph = bin / 7, pw = bin - ph * 7;
          const float xb = fadd(sw, fmul((float)pw, bw));
          for (int iy = 0; iy < gh; iy++) {
            const float yy = fadd(fadd(sh, fmul((float)ph, bh)), fdiv(fmul((float)iy + .5f, bh), (float)gh));
            for (int ix = 0; ix < gw; ix++) {
              const float xx = fadd(xb, fdiv(fmul((float)ix + .5f, bw), (float)gw));
              const float y = fadd(fsub(fmul(yy, ct), fmul(xx, st)), cy);
              const float x = fadd(fadd(fmul(yy, st), fmul(xx, ct)), cx);
              int yl, yh, xl, xh; float ly, hy, lx, hx;
              if (y < -1.0f || y > (float)H || x < -1.0f || x > (float)W) continue;
              axis_setup_b(y, H, yl, yh, ly, hy);
              axis_setup_b(x, W, xl, xh, lx, hx);
              const int ty[2] = {yl - cy0, yh - cy0}, tx[2] = {xl - cx0, xh - cx0};
              const float wyv[2] = {hy, ly}, wxv[2] = {hx, lx};
#pragma unroll
              for (int a = 0; a < 2; a++)
#pragma unroll
                for (int c = 0; c < 2; c++) {
                  if ((unsigned)ty[a] < 4u && (unsigned)tx[c] < 4u) {
                    const int idx = ty[a] * 4 + tx[c];
                    const float wv = wyv[a] * wxv[c] * inv_count;
#pragma unroll
                    for (int i = 0; i < 16; i++) wr[i] += (i == idx) ? wv : 0.f;
                    any = 1;
                  }
                }
            }
          }
#pragma unroll
          for (int i = 0; i < 4; i++)
            *reinterpret_cast<float4*>(wmat + bin * 16 + i * 4) = make_float4(wr[4 * i], wr[4 * i + 1], wr[4 * i + 2], wr[4 * i + 3]);
        }
        any = __any_sync(0xffffffffu, any);
        __syncwarp();
        if (!any) continue;
        for (int c0 = lane * 8; c0 < C; c0 += 256) {
          float acc[16][8];
#pragma unroll
          for (int p = 0; p < 16; p++)
#pragma unroll
            for (int j = 0; j < 8; j++) acc[p][j] = 0.f;
          for (int ph = 0; ph < 7; ph++) {
            uint4 u7[7];
#pragma unroll
            for (int pw = 0; pw < 7; pw++)
              u7[pw] = __ldg(reinterpret_cast<const uint4*>(drow + (size_t)(ph * 7 + pw) * C + c0));
#pragma unroll
            for (int pw = 0; pw < 7; pw++) {
              const uint32_t wv[4] = {u7[pw].x, u7[pw].y, u7[pw].z, u7[pw].w};
              float d[8];
#pragma unroll
              for (int q = 0; q < 4; q++) { d[2 * q] = __uint_as_float(wv[q] << 16); d[2 * q + 1] = __uint_as_float(wv[q] & 0xffff0000u); }
#pragma unroll
              for (int q4 = 0; q4 < 4; q4++) {
                const float4 w4 = *reinterpret_cast<const float4*>(wmat + (ph * 7 + pw) * 16 + q4 * 4);
                const float wq[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                for (int e = 0; e < 4; e++) {
                  if (wq[e] == 0.f) continue;                 // warp-uniform
#pragma unroll
                  for (int j = 0; j < 8; j++) acc[q4 * 4 + e][j] = fmaf(wq[e], d[j], acc[q4 * 4 + e][j]);
                }
              }
            }
          }
#pragma unroll
          for (int py = 0; py < 4; py++)
#pragma unroll
            for (int px = 0; px < 4; px++) {
              const int yy = cy0 + py, xx = cx0 + px;
              if (yy > ymax || xx > xmax) continue;
              float* dst = fb + ((size_t)yy * W + xx) * C + c0;
              const float* a = acc[py * 4 + px];
              asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(a[0]), "f"(a[1]),
                           "f"(a[2]), "f"(a[3]) : "memory");
              asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4), "f"(a[4]), "f"(a[5]),
                           "f"(a[6]), "f"(a[7]) : "memory");
            }
        }
      }
    }
  }
}

}  // namespace ptb

using namespace ptb;

extern "C" int pt_reg_loss_grad_ex(const float* deltas, const float* bag_rois, const unsigned char* valid,
                                   const float* ref_boxes, int U, int K, float max_w, float max_h, float wh_ratio_clip,
                                   float hyper, float eps, const float* sums, const float* gscale, float scale, float* g,
                                   int rotated, void* stream);

extern "C" int pt_reg_loss_grad(const float* deltas, const float* bag_rois, const unsigned char* valid,
                                const float* ref_boxes, int U, int K, float max_w, float max_h, float wh_ratio_clip,
                                float hyper, float eps, const float* sums, const float* gscale, float scale, float* g,
                                void* stream) {
  return pt_reg_loss_grad_ex(deltas, bag_rois, valid, ref_boxes, U, K, max_w, max_h, wh_ratio_clip, hyper, eps, sums,
                             gscale, scale, g, 0, stream);
}

// rotated = 1: bag_rois [K,6] (b,cx,cy,w,h,theta), ref_boxes [G,5]; the loss itself is the horizontal DN-DIoU on the
// (cx,cy,w,h) parts (OBB_TOD/mmrotate/models/dense_heads/rotated_fcos_head_p2rb_ts.py:1314-1322)
extern "C" int pt_reg_loss_grad_ex(const float* deltas, const float* bag_rois, const unsigned char* valid,
                                   const float* ref_boxes, int U, int K, float max_w, float max_h, float wh_ratio_clip,
                                   float hyper, float eps, const float* sums, const float* gscale, float scale, float* g,
                                   int rotated, void* stream) {
  if (K <= 0) return PT_OK;
  reg_loss_grad_kernel<<<(K + 127) / 128, 128, 0, (cudaStream_t)stream>>>(deltas, bag_rois, valid, ref_boxes, U, K, max_w,
                                                                         max_h, fabsf(logf(wh_ratio_clip)), hyper, eps,
                                                                         sums, gscale, scale, g, rotated);
  return check_launch("reg_loss_grad_kernel");
}

// g [K + n_neg, 2C]; positives: K = G*U1*U2 rows; negatives appended (neg_cls = cls + K*C).
extern "C" int pt_bag_loss_grad(const float* cls, const float* ins, const unsigned char* valid, const long long* labels,
                                int G, int U1, int U2, int C, const unsigned char* neg_weight, int n_neg,
                                const float* sums, const float* gscale, float pos_scale, float neg_scale, float* g,
                                void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  const float eps = 1e-6f;
  if (G > 0) {
    const size_t smem = ((size_t)2 * U2 * C + 4 * C) * sizeof(float);
    if (smem > 48 * 1024) { set_error("pt_bag_loss_grad: bag of %d x %d classes exceeds shared memory", U2, C); return PT_ERR_UNSUPPORTED; }
    bag_loss_grad_kernel<<<G * U1, 128, smem, s>>>(cls, ins, valid, labels, U1, U2, C, sums, gscale, pos_scale, eps, g);
    int rc = check_launch("bag_loss_grad_kernel");
    if (rc != PT_OK) return rc;
  }
  if (n_neg > 0) {
    const long long K = (long long)G * U1 * U2;
    neg_loss_grad_kernel<<<(n_neg * C + 255) / 256, 256, 0, s>>>(cls + K * C, neg_weight, n_neg, C, sums, gscale, neg_scale,
                                                                 eps, g + K * 2 * C);
    return check_launch("neg_loss_grad_kernel");
  }
  return PT_OK;
}

extern "C" int pt_head_bwd(const float* g, int nout, const void* H_bf16, long long ldh, int D, const float* W, int M,
                           void* dZ_bf16, long long ldz, float* dW, float* db, void* stream) {
  if (M <= 0) return PT_OK;
  if (D % 512 != 0 || (ldh % 4) || (ldz % 4)) { set_error("pt_head_bwd: D must be a multiple of 512, leading dimensions of 4"); return PT_ERR_ARG; }
  cudaStream_t s = (cudaStream_t)stream;
  const dim3 grid(M < 296 ? M : 296, D >= 1024 ? 2 : 1);   // 592 CTAs, 2 resident per SM; dW: grid.x x nout x D / 2 reds
  const __nv_bfloat16* H = reinterpret_cast<const __nv_bfloat16*>(H_bf16);
  __nv_bfloat16* dZ = reinterpret_cast<__nv_bfloat16*>(dZ_bf16);
  switch (nout) {
    case 4: head_bwd_kernel<4><<<grid, 256, 0, s>>>(g, H, ldh, D, W, M, dZ, ldz, dW, db); break;
    case 16: head_bwd_kernel<16><<<grid, 256, 0, s>>>(g, H, ldh, D, W, M, dZ, ldz, dW, db); break;
    case 18: head_bwd_kernel<18><<<grid, 256, 0, s>>>(g, H, ldh, D, W, M, dZ, ldz, dW, db); break;
    default: set_error("pt_head_bwd: unsupported head width %d (built: 4, 16, 18)", nout); return PT_ERR_UNSUPPORTED;
  }
  return check_launch("head_bwd_kernel");
}

extern "C" int pt_transpose_pad_bf16(const void* in, long long ldin, int R, int C, void* out, long long ldout,
                                     void* stream) {
  if (R <= 0 || C <= 0) return PT_OK;
  if (ldout < R) { set_error("pt_transpose_pad_bf16: ldout < R"); return PT_ERR_ARG; }
  dim3 grid((unsigned)((ldout + 63) / 64), (C + 63) / 64);
  transpose_pad_bf16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(in), ldin, R, C,
                                                                   reinterpret_cast<__nv_bfloat16*>(out), ldout);
  return check_launch("transpose_pad_bf16_kernel");
}

static int unpermute_row(const float* dw_binmajor, int N, int C, int bins, float* grad, int accumulate, cudaStream_t st) {
  const size_t smem = (size_t)C * bins * sizeof(float);
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(ptb::unpermute_dw1_row_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    if (e != cudaSuccess) { ptb::set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return PT_ERR_CUDA; }
    attr_done = true;
  }
  ptb::unpermute_dw1_row_kernel<<<N, 256, smem, st>>>(dw_binmajor, C, bins, grad, accumulate);
  return ptb::check_launch("unpermute_dw1_row_kernel");
}

extern "C" int pt_unpermute_dw1(const float* dw_binmajor, int N, int C, int bins, float* grad, int accumulate,
                                void* stream) {
  if (N <= 0) return PT_OK;
  if (C % 4 == 0 && (size_t)C * bins * sizeof(float) <= 64 * 1024 && (((uintptr_t)dw_binmajor | (uintptr_t)grad) & 15) == 0)
    return unpermute_row(dw_binmajor, N, C, bins, grad, accumulate, (cudaStream_t)stream);
  const size_t smem = (size_t)(UNP_CH + 1) * bins * sizeof(float);
  if (smem > 48 * 1024) { set_error("pt_unpermute_dw1: %d bins exceed shared memory", bins); return PT_ERR_UNSUPPORTED; }
  dim3 grid(N, (C + UNP_CH - 1) / UNP_CH);
  unpermute_dw1_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(dw_binmajor, C, bins, grad, accumulate);
  return check_launch("unpermute_dw1_kernel");
}

extern "C" int pt_colsum_bf16(const void* dZ, long long ld, int M, int N, float* db, void* stream) {
  if (M <= 0 || N <= 0) return PT_OK;
  if ((N % 8) || (ld % 8) || ((uintptr_t)dZ & 15)) { set_error("pt_colsum_bf16: N / ld must be multiples of 8, 16-byte aligned"); return PT_ERR_ARG; }
  const int col_blocks = (N + 255) / 256;
  int row_blocks = (M + 31) / 32;
  if (row_blocks * col_blocks > 592) row_blocks = (592 + col_blocks - 1) / col_blocks;
  dim3 grid(col_blocks, row_blocks);
  colsum_bf16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(dZ), ld, M, N, db);
  return check_launch("colsum_bf16_kernel");
}

extern "C" int pt_nhwc_to_nchw_f32(const float* in, float* out, int B, int C, int H, int W, int accumulate, void* stream) {
  if (B <= 0) return PT_OK;
  const int HW = H * W;
  dim3 grid((HW + 63) / 64, (C + 63) / 64, B);
  nhwc_to_nchw_f32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in, out, C, HW, accumulate);
  return check_launch("nhwc_to_nchw_f32_kernel");
}

// dfeat NHWC fp32 [B,H,W,C] must be zeroed by the caller; dA bf16 [K, ld] bin-major.
extern "C" int pt_roi_align_backward_ex(const void* dA_bf16, long long ld, const float* rois, int K, int B, int C,
                                        int H, int W, float spatial_scale, int sampling_ratio, int aligned,
                                        float* dfeat, const int* roi_level, int level, void* stream) {
  if (K <= 0) return PT_OK;
  if (C % 8 != 0) { set_error("pt_roi_align_backward: C must be a multiple of 8"); return PT_ERR_ARG; }
  // tensor-core path (roi_align_mma.cu); PTB200_RA_BWD_MMA=0 keeps the register formulation below for A/B measurements
  static const bool use_mma = [] { const char* e = getenv("PTB200_RA_BWD_MMA"); return e == nullptr || e[0] != '0'; }();
  if (use_mma && ramma::bwd_supported(C, H, W, ld))
    return ramma::launch_bwd(dA_bf16, ld, rois, K, B, C, H, W, spatial_scale, sampling_ratio, aligned, dfeat, roi_level,
                             level, (cudaStream_t)stream);
  const size_t smem = (size_t)RB_WARPS * ((W + 4) + (H + 4)) * 8 * sizeof(float);
  if (smem > 200 * 1024) { set_error("pt_roi_align_backward: feature map too large for the shared weight tables"); return PT_ERR_UNSUPPORTED; }
  cudaError_t e = cudaFuncSetAttribute(roi_align_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return PT_ERR_CUDA; }
  const int blocks = (K + RB_WARPS - 1) / RB_WARPS;
  roi_align_bwd_kernel<<<blocks < 148 * 8 ? blocks : 148 * 8, RB_WARPS * 32, smem, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(dA_bf16), ld, rois, K, B, C, H, W, spatial_scale, sampling_ratio, aligned,
      dfeat, roi_level, level);
  return check_launch("roi_align_bwd_kernel");
}
extern "C" int pt_roi_align_backward(const void* dA_bf16, long long ld, const float* rois, int K, int B, int C, int H,
                                     int W, float spatial_scale, int sampling_ratio, int aligned, float* dfeat,
                                     void* stream) {
  return pt_roi_align_backward_ex(dA_bf16, ld, rois, K, B, C, H, W, spatial_scale, sampling_ratio, aligned, dfeat,
                                  nullptr, 0, stream);
}

// RoIAlignRotated backward (mmcv roi_align_rotated, fixed or adaptive sampling grid): rois [K,6] (b,cx,cy,w,h,theta).
extern "C" int pt_roi_align_rotated_backward_ex(const void* dA_bf16, long long ld, const float* rois, int K, int B,
                                                int C, int H, int W, float spatial_scale, int sampling_ratio,
                                                int aligned, int clockwise, float* dfeat, const int* roi_level,
                                                int level, void* stream) {
  if (K <= 0) return PT_OK;
  if (C % 8 != 0) { set_error("pt_roi_align_rotated_backward: C must be a multiple of 8"); return PT_ERR_ARG; }
  const int blocks = (K + RB_WARPS - 1) / RB_WARPS;
  roi_align_rotated_bwd_kernel<<<blocks < 148 * 8 ? blocks : 148 * 8, RB_WARPS * 32, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(dA_bf16), ld, rois, K, B, C, H, W, spatial_scale, sampling_ratio, aligned,
      clockwise, dfeat, roi_level, level);
  return check_launch("roi_align_rotated_bwd_kernel");
}
extern "C" int pt_roi_align_rotated_backward(const void* dA_bf16, long long ld, const float* rois, int K, int B, int C,
                                             int H, int W, float spatial_scale, int sampling_ratio, int aligned,
                                             int clockwise, float* dfeat, void* stream) {
  return pt_roi_align_rotated_backward_ex(dA_bf16, ld, rois, K, B, C, H, W, spatial_scale, sampling_ratio, aligned,
                                          clockwise, dfeat, nullptr, 0, stream);
}
