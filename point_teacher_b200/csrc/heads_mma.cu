// Small output heads of the MIL stack on warp-level tensor cores: out[M, n0 + n1] = H[M, D] . [W0 ; W1]^T (+ bias).
//   fc_reg (4 outputs)            HBB_TOD/mmdet/models/dense_heads/fcos_head_p2b_ts.py:1207
//   fc_cls + fc_ins (C + C)       HBB_TOD/mmdet/models/dense_heads/fcos_head_p2b_ts.py:1250-1251, 1273
// H is the bf16 hidden activation written by the FC GEMM; the fp32 weights are split into bf16 hi + bf16 lo while
// they are staged in shared memory (two mma per fragment), so the heads keep fp32-class weight precision: box
// deltas and the loss gradients behind them are sensitive to it.  The FFMA formulation
// (mil_head.cu: cls_ins_kernel, 36 us for 5400 x 1024 x 16) is latency- and shuffle-bound; here one warp computes a
// 16-row tile over half of K with mma.sync.m16n8k16 from 16-byte loads.  The K order inside a 32-wide block is
// permuted identically for A and B (lane t owns k = 32 s + 8 t .. + 7), which turns every fragment load into one
// LDG.128 / LDS.128.
#include "common.cuh"

namespace ptb {

constexpr int SH_KS = 8;        // K slices per row tile: D/8 = 128 hidden units = 4 fully unrolled 32-wide steps per warp
constexpr int SH_WARPS = 2 * SH_KS;   // 2 row tiles x 8 K slices per CTA: one round of load latency per tile

// Two CTAs per SM (<= 64 registers, 2 x 82 KB of shared memory): the bench's 169 row tiles then run as ONE wave of
// the 296 slots; with one CTA per SM (69 registers) the last 21 tiles formed a second wave and doubled the kernel time.
template <int NT>
__global__ void __launch_bounds__(SH_WARPS * 32, 2)
small_head_mma_kernel(const __nv_bfloat16* __restrict__ H, long long ldh, int D, const float* __restrict__ W0, int n0,
                      const float* __restrict__ b0, const float* __restrict__ W1, int n1, const float* __restrict__ b1,
                      int M, float* __restrict__ out0, float* __restrict__ out1) {
  extern __shared__ __align__(16) uint8_t smem_h[];
  const int row_bytes = D * 2 + 16;                               // padded: conflict-free LDS.128 across 8 rows
  // weights as bf16 hi + bf16 lo (w ~= hi + lo to 16 mantissa bits): [2][NT*8][D + 8]
  const size_t lo_off = (size_t)NT * 8 * row_bytes;
  float* part = reinterpret_cast<float*>(smem_h + 2 * lo_off);    // [2 tiles][SH_KS][32 lanes][NT*4]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  {
    const int total = NT * 8 * (D / 4);
    for (int i0 = threadIdx.x; i0 < total; i0 += blockDim.x * 8) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; u++) {                               // 8 independent 16-byte loads in flight per thread
        const int i = i0 + u * blockDim.x;
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < total) {
          const int n = i / (D / 4), k = (i - n * (D / 4)) * 4;
          if (n < n0) v[u] = __ldg(reinterpret_cast<const float4*>(W0 + (size_t)n * D + k));
          else if (n < n0 + n1) v[u] = __ldg(reinterpret_cast<const float4*>(W1 + (size_t)(n - n0) * D + k));
        }
      }
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const int i = i0 + u * blockDim.x;
        if (i < total) {
          const int n = i / (D / 4), k = (i - n * (D / 4)) * 4;
          const float e[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
          float l[4];
#pragma unroll
          for (int j = 0; j < 4; j++) l[j] = e[j] - __bfloat162float(__float2bfloat16_rn(e[j]));
          uint8_t* dst = smem_h + (size_t)n * row_bytes + k * 2;
          *reinterpret_cast<uint2*>(dst) = make_uint2(pack_bf16(e[0], e[1]), pack_bf16(e[2], e[3]));
          *reinterpret_cast<uint2*>(dst + lo_off) = make_uint2(pack_bf16(l[0], l[1]), pack_bf16(l[2], l[3]));
        }
      }
    }
  }
  __syncthreads();
  const int g = lane >> 2, t = lane & 3;
  const int tile = warp / SH_KS, ks = warp % SH_KS;
  const int kbeg = ks * (D / SH_KS), kend = kbeg + D / SH_KS;
  for (int rp = blockIdx.x * 32; rp < M; rp += gridDim.x * 32) {    // trip count uniform over the CTA (barriers inside)
    const int r0 = rp + tile * 16;
    const int ra = min(r0 + g, M - 1), rb = min(r0 + g + 8, M - 1);
    const __nv_bfloat16* pa = H + (size_t)ra * ldh + 8 * t;
    const __nv_bfloat16* pb = H + (size_t)rb * ldh + 8 * t;
    float acc[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; nt++)
#pragma unroll
      for (int e = 0; e < 4; e++) acc[nt][e] = 0.f;
#pragma unroll 4
    for (int k0 = kbeg; k0 < kend; k0 += 32) {
      const uint4 ua = __ldg(reinterpret_cast<const uint4*>(pa + k0));
      const uint4 ub = __ldg(reinterpret_cast<const uint4*>(pb + k0));
#pragma unroll
      for (int nt = 0; nt < NT; nt++) {
        const uint8_t* wp = smem_h + (size_t)(nt * 8 + g) * row_bytes + (k0 + 8 * t) * 2;
#pragma unroll
        for (int part_i = 0; part_i < 2; part_i++) {              // hi, then lo
          const uint4 wb = *reinterpret_cast<const uint4*>(wp + part_i * lo_off);
          asm volatile(
              "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
              : "+f"(acc[nt][0]), "+f"(acc[nt][1]), "+f"(acc[nt][2]), "+f"(acc[nt][3])
              : "r"(ua.x), "r"(ub.x), "r"(ua.y), "r"(ub.y), "r"(wb.x), "r"(wb.y));
          asm volatile(
              "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
              : "+f"(acc[nt][0]), "+f"(acc[nt][1]), "+f"(acc[nt][2]), "+f"(acc[nt][3])
              : "r"(ua.z), "r"(ub.z), "r"(ua.w), "r"(ub.w), "r"(wb.z), "r"(wb.w));
        }
      }
    }
    // reduce the K slices through shared memory, then bias + store (fp32)
    __syncthreads();
#pragma unroll
    for (int nt = 0; nt < NT; nt++)
#pragma unroll
      for (int e = 0; e < 4; e++) part[(((size_t)tile * SH_KS + ks) * 32 + lane) * NT * 4 + nt * 4 + e] = acc[nt][e];
    __syncthreads();
    if (ks == 0) {
#pragma unroll
      for (int nt = 0; nt < NT; nt++) {
#pragma unroll
        for (int e = 0; e < 4; e++) {
          float v = 0.f;
#pragma unroll
          for (int q = 0; q < SH_KS; q++) v += part[(((size_t)tile * SH_KS + q) * 32 + lane) * NT * 4 + nt * 4 + e];
          const int row = r0 + g + (e >> 1) * 8, col = nt * 8 + 2 * t + (e & 1);
          if (row < M) {
            if (col < n0) out0[(size_t)row * n0 + col] = v + (b0 != nullptr ? b0[col] : 0.f);
            else if (col < n0 + n1) out1[(size_t)row * n1 + (col - n0)] = v + (b1 != nullptr ? b1[col - n0] : 0.f);
          }
        }
      }
    }
  }
}

template <int NT>
static int launch_small_head(const void* H, long long ldh, int D, const float* W0, int n0, const float* b0,
                             const float* W1, int n1, const float* b1, int M, float* out0, float* out1, cudaStream_t s) {
  const size_t smem = (size_t)2 * NT * 8 * (D * 2 + 16) + (size_t)2 * SH_KS * 32 * NT * 4 * sizeof(float);
  auto kern = small_head_mma_kernel<NT>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return PT_ERR_CUDA; }
  int grid = (M + 31) / 32;
  if (grid > 148 * 2) grid = 148 * 2;           // resident slots; beyond that the persistent row loop takes over
  kern<<<grid, SH_WARPS * 32, smem, s>>>(reinterpret_cast<const __nv_bfloat16*>(H), ldh, D, W0, n0, b0, W1, n1, b1, M, out0,
                                         out1);
  return check_launch("small_head_mma_kernel");
}

}  // namespace ptb

using namespace ptb;

// out0 [M, n0] = H W0^T + b0, out1 [M, n1] = H W1^T + b1 (W1 / b0 / b1 may be NULL); H bf16 [M, ldh], D % 64 == 0.
extern "C" int pt_small_heads_bf16(const void* H, long long ldh, int D, const float* W0, int n0, const float* b0,
                                   const float* W1, int n1, const float* b1, int M, float* out0, float* out1,
                                   void* stream) {
  if (M <= 0) return PT_OK;
  if (D % (32 * SH_KS) != 0 || (ldh % 8) != 0 || ((uintptr_t)H & 15)) { set_error("pt_small_heads_bf16: D %% 256, ldh %% 8 and 16-byte alignment required"); return PT_ERR_ARG; }
  if (W1 == nullptr) n1 = 0;
  const int n = n0 + n1;
  cudaStream_t s = (cudaStream_t)stream;
  if (n <= 8) return launch_small_head<1>(H, ldh, D, W0, n0, b0, W1, n1, b1, M, out0, out1, s);
  if (n <= 16) return launch_small_head<2>(H, ldh, D, W0, n0, b0, W1, n1, b1, M, out0, out1, s);
  if (n <= 24) return launch_small_head<3>(H, ldh, D, W0, n0, b0, W1, n1, b1, M, out0, out1, s);
  if (n <= 32) return launch_small_head<4>(H, ldh, D, W0, n0, b0, W1, n1, b1, M, out0, out1, s);
  set_error("pt_small_heads_bf16: %d outputs exceed the built 32", n);
  return PT_ERR_UNSUPPORTED;
}
