// RoIAlign / RoIAlignRotated forward for sm_100a: NHWC gathers, one warp per (RoI, output row).
//
// Replaces mmcv.ops.RoIAlign(output_size=7, sampling_ratio=0, pool_mode='avg', aligned=True) and
// mmcv.ops.RoIAlignRotated(output_size=7, sampling_ratio=2, aligned=True, clockwise=True) as reached through
//   HBB_TOD/mmdet/models/roi_heads/roi_extractors/base_roi_extractor.py:50-59 (layer construction)
//   HBB_TOD/mmdet/models/roi_heads/roi_extractors/single_level_roi_extractor.py:56-114
//   OBB_TOD/mmrotate/models/roi_heads/roi_extractors/rotate_single_level_roi_extractor.py:90-167
//
// Access pattern: the feature map is transposed once per step to NHWC so that the 256 channels of one pixel
// are one contiguous 1 KB (fp32) / 512 B (bf16) line; a lane owns 8 consecutive channels (16/32-byte vector
// loads, 16-byte bf16 stores).  For the horizontal op a warp sweeps one output row left to right, keeps the
// two current feature columns (already interpolated in y) in registers and only loads a new column when the
// sample crosses a pixel boundary -- tiny objects (2x2 feature pixels) touch ~3 columns for 7 bins.
// Sample coordinates use non-contracted fp32 arithmetic in the reference's operation order so the adaptive
// grid size and the border rule are decided identically.
#include <cuda_fp16.h>

#include "common.cuh"

namespace ptb {

namespace ramma {   // roi_align_mma.cu: TMA + mma.sync bf16 throughput path
bool supported(int C, int H, int W, bool rot, int sampling_ratio);
int launch(const void* feat_nhwc, int feat_f16, const float* rois, void* out, long long ld_out, int K, int B, int C,
           int H, int W, float scale, int sampling_ratio, int aligned, const int* roi_level, int level,
           bool rot, int clockwise, cudaStream_t stream);
}  // namespace ramma

constexpr int P7 = 7;
constexpr int GW_CHUNK = 16;   // x samples per bin held in the shared sample table at a time

// --------------------------------------------------------------------------------- NCHW -> NHWC
// 64 channels x 64 positions per CTA: 16-byte loads along HW, 16-/8-byte stores along C.
template <typename T> struct IS_HALF { static constexpr bool value = false; };
template <> struct IS_HALF<__half> { static constexpr bool value = true; };
// fp16 feature maps saturate to +-65504 instead of overflowing to inf
__device__ __forceinline__ uint32_t pack_f16_sat(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
template <typename TOut> __device__ __forceinline__ TOut cvt_out(float v) { return (TOut)v; }
template <> __device__ __forceinline__ __half cvt_out<__half>(float v) {
  return __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
}
template <typename TOut>
__global__ void __launch_bounds__(256)
nchw_to_nhwc_kernel(const float* __restrict__ in, TOut* __restrict__ out, int C, int HW, int* __restrict__ sat_count) {
  __shared__ float tile[64][65];
  int n_sat = 0;   // fp16 only: values the saturating conversion changes (|v| > 65504, inf) -- a deviation from the
                   // fp32 reference that the host must hear about (roi_extractors._NHWCCache)
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * 64, p0 = blockIdx.x * 64;
  const float* src = in + (size_t)b * C * HW;
  TOut* dst = out + (size_t)b * C * HW;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;     // 16 x 16
  const bool vec_in = (HW & 3) == 0;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int c = c0 + ty + 16 * i, p = p0 + tx * 4;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (c < C) {
      if (vec_in && p + 3 < HW) {
        const float4 f = *reinterpret_cast<const float4*>(src + (size_t)c * HW + p);
        v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
      } else {
#pragma unroll
        for (int j = 0; j < 4; j++) if (p + j < HW) v[j] = src[(size_t)c * HW + p + j];
      }
    }
#pragma unroll
    for (int j = 0; j < 4; j++) tile[ty + 16 * i][tx * 4 + j] = v[j];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int p = p0 + ty + 16 * i, c = c0 + tx * 4;
    if (p < HW && c < C) {
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; j++) v[j] = tile[tx * 4 + j][ty + 16 * i];
      if (IS_HALF<TOut>::value) {
#pragma unroll
        for (int j = 0; j < 4; j++) n_sat += (c + j < C && fabsf(v[j]) > 65504.f) ? 1 : 0;
      }
      TOut* o = dst + (size_t)p * C + c;
      if ((C & 3) == 0) {
        if (sizeof(TOut) == 4) *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
        else if (IS_HALF<TOut>::value) *reinterpret_cast<uint2*>(o) = make_uint2(pack_f16_sat(v[0], v[1]), pack_f16_sat(v[2], v[3]));
        else *reinterpret_cast<uint2*>(o) = make_uint2(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]));
      } else {
#pragma unroll
        for (int j = 0; j < 4; j++) if (c + j < C) o[j] = cvt_out<TOut>(v[j]);
      }
    }
  }
  if (IS_HALF<TOut>::value && sat_count != nullptr && __any_sync(0xffffffffu, n_sat != 0)) {
    n_sat = __reduce_add_sync(0xffffffffu, n_sat);
    if ((threadIdx.x & 31) == 0) atomicAdd(sat_count, n_sat);
  }
}

// --------------------------------------------------------------------------------- helpers
struct F8 { float v[8]; };

__device__ __forceinline__ F8 load8(const float* p) {
  F8 r;
  float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ F8 load8(const __nv_bfloat16* p) {
  F8 r;
  uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; i++) {
    r.v[2 * i] = __uint_as_float(w[i] << 16);
    r.v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
  return r;
}

// one axis of the Detectron2 bilinear rule; returns false when the sample is outside [-1, size]
__device__ __forceinline__ bool axis_setup(float v, int size, int& lo, int& hi, float& l, float& h) {
  if (v < -1.0f || v > (float)size) return false;
  if (v <= 0.f) v = 0.f;
  lo = (int)v;
  if (lo >= size - 1) { hi = lo = size - 1; v = (float)lo; } else hi = lo + 1;
  l = fsub(v, (float)lo);
  h = fsub(1.0f, l);
  return true;
}

__device__ __forceinline__ F8 load8(const __half* p) {
  F8 r;
  uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
    r.v[2 * i] = f.x;
    r.v[2 * i + 1] = f.y;
  }
  return r;
}

// raw 8-channel loads of the 16-bit feature maps (kept unconverted while a batch of gathers is in flight)
__device__ __forceinline__ uint4 raw8(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ uint4 raw8(const __half* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ F8 cvt8(const uint4& u, const __nv_bfloat16*) {
  F8 r;
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; i++) { r.v[2 * i] = __uint_as_float(w[i] << 16); r.v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
  return r;
}
__device__ __forceinline__ F8 cvt8(const uint4& u, const __half*) {
  F8 r;
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
    r.v[2 * i] = f.x; r.v[2 * i + 1] = f.y;
  }
  return r;
}

enum { OUT_BF16_BINMAJOR = 0, OUT_F32_NCHW = 1, OUT_BF16X3_BINMAJOR = 2 };

template <int MODE>
__device__ __forceinline__ void store_bins(void* out, long long ld_out, int roi, int ph, int C, int c0,
                                           const float (&acc)[P7][8], float inv_count, float* stage) {
  if (MODE == OUT_F32_NCHW) {
    // transposed staging: element (c, bin) at bin*(C+1) + (c%8)*(C/8) + c/8 (conflict-free both ways for C=256)
#pragma unroll
    for (int pw = 0; pw < P7; pw++) {
      const int bin = ph * P7 + pw;
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const int c = c0 + j;
        stage[bin * (C + 1) + (c & 7) * (C >> 3) + (c >> 3)] = acc[pw][j] * inv_count;
      }
    }
  } else {
    __nv_bfloat16* row = reinterpret_cast<__nv_bfloat16*>(out) + (size_t)roi * ld_out;
    const int seg = P7 * P7 * C;
#pragma unroll
    for (int pw = 0; pw < P7; pw++) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; j++) v[j] = acc[pw][j] * inv_count;
      const size_t o = (size_t)(ph * P7 + pw) * C + c0;
      uint4 hi = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
      *reinterpret_cast<uint4*>(row + o) = hi;
      if (MODE == OUT_BF16X3_BINMAJOR) {
        // fp32 emulation operand [hi | lo | hi]: A*W ~= hi*Whi + lo*Whi + hi*Wlo
        float l[8];
#pragma unroll
        for (int j = 0; j < 8; j++) l[j] = v[j] - __bfloat162float(__float2bfloat16_rn(v[j]));
        uint4 lo = make_uint4(pack_bf16(l[0], l[1]), pack_bf16(l[2], l[3]), pack_bf16(l[4], l[5]), pack_bf16(l[6], l[7]));
        *reinterpret_cast<uint4*>(row + seg + o) = lo;
        *reinterpret_cast<uint4*>(row + 2 * seg + o) = hi;
      }
    }
  }
}

template <int MODE>
__device__ __forceinline__ void flush_stage(float* out, int roi, int C, const float* stage) {
  // coalesced write of one RoI's (C, 7, 7) block from the transposed smem staging
  const int n = C * P7 * P7;
  float* dst = out + (size_t)roi * n;
  for (int o = threadIdx.x * 4; o < n; o += P7 * 32 * 4) {
    float v[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int c = (o + i) / (P7 * P7), bin = (o + i) - c * (P7 * P7);
      v[i] = stage[bin * (C + 1) + (c & 7) * (C >> 3) + (c >> 3)];
    }
    *reinterpret_cast<float4*>(dst + o) = make_float4(v[0], v[1], v[2], v[3]);
  }
}

// --------------------------------------------------------------------------------- horizontal RoIAlign
// Separable formulation.  Bilinear average pooling is linear in the feature map and the sample grid is a
// Cartesian product, so for one RoI
//     out[ph][pw][c] = 1/count * sum_row sum_col  wy[ph][row] * wx[pw][col] * feat[row][col][c]
// where wx[pw][col] (wy[ph][row]) accumulates the low/high bilinear weights of every x (y) sample of bin pw
// (ph) that lands on that column (row); samples rejected by the border rule simply add no weight.
// CTA = 7 compute warps (one output row each) + 1 builder warp that runs ONE RoI AHEAD: it loads the RoI,
// builds the two tiny weight tables with the reference's exact coordinate arithmetic (double-buffered in
// shared memory, handed over with named barriers) and prefetches the RoI's feature patch into L1, so the
// compute warps never see the RoI-load / table-build latency and their pixel loads hit L1.
struct RoiMeta {
  int xmin, xmax, ymin, skip;
  float inv_count;
  int img;
  int ylo[8], yhi[8];       // row range touched by output row ph
};

constexpr int RA_THREADS = (P7 + 1) * 32;
enum { BAR_FULL0 = 1, BAR_EMPTY0 = 3, BAR_COMPUTE = 5 };

__device__ __forceinline__ void nbar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void nbar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

template <typename TIn, int MODE>
__global__ void __launch_bounds__(RA_THREADS, 2)
roi_align_fwd_kernel(const TIn* __restrict__ feat, const float* __restrict__ rois, void* __restrict__ out,
                     long long ld_out, int K, int B, int C, int H, int W, float scale, int sampling_ratio,
                     int aligned, const int* __restrict__ roi_level, int level) {
  extern __shared__ float smem_dyn[];
  // [2 buffers][ (W+1)*8 wx | (H+1)*8 wy ] then the NCHW staging area
  const int tab_floats = (W + 4) * 8 + (H + 1) * 8;
  float* stage = smem_dyn + 2 * tab_floats;
  __shared__ RoiMeta meta[2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_iter = blockIdx.x < K ? (K - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (warp == P7) {
    // ------------------------------------------------------------------ builder warp
    const float off = aligned ? 0.5f : 0.f;
    for (int it = 0; it < n_iter; it++) {
      const int roi = blockIdx.x + it * gridDim.x, buf = it & 1;
      if (it >= 2) nbar_sync(BAR_EMPTY0 + buf, RA_THREADS);
      float* wx = smem_dyn + buf * tab_floats;       // wx[col - xmin][pw]
      float* wy = wx + (W + 4) * 8;                  // wy[row - ymin][ph]
      RoiMeta& mt = meta[buf];
      const bool skip = roi_level != nullptr && roi_level[roi] != level;   // multi-level FPN: other level's RoI
      const float* r = rois + (size_t)roi * 5;
      const int b = (int)__ldg(r);
      const float x1 = fsub(fmul(__ldg(r + 1), scale), off), y1 = fsub(fmul(__ldg(r + 2), scale), off);
      const float x2 = fsub(fmul(__ldg(r + 3), scale), off), y2 = fsub(fmul(__ldg(r + 4), scale), off);
      float rw = fsub(x2, x1), rh = fsub(y2, y1);
      if (!aligned) { rw = fmaxf(rw, 1.f); rh = fmaxf(rh, 1.f); }
      const float bh = fdiv(rh, (float)P7), bw = fdiv(rw, (float)P7);
      const int gh = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(bh);
      const int gw = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(bw);
      const int cnt = gh * gw > 1 ? gh * gw : 1;
      const bool b_ok = b >= 0 && b < B && !skip;
      // lanes 0..6 own the x bins, lanes 8..14 the y bins (shuffle groups of 8)
      const bool isx = lane < 8;
      const int bi = lane & 7;
      const float start = isx ? x1 : y1, bin = isx ? bw : bh;
      const int g = isx ? gw : gh, size = isx ? W : H;
      float* tab = isx ? wx : wy;
      const float base = fadd(start, fmul((float)bi, bin));
      int lo = 1 << 30, hi = -1;
      const bool owner = lane < 16 && bi < P7;
      if (owner && b_ok) {
        for (int i = 0; i < g; i++) {
          const float v = fadd(base, fdiv(fmul((float)i + .5f, bin), (float)g));
          int l, h; float fl, fh;
          if (axis_setup(v, size, l, h, fl, fh)) { lo = min(lo, l); hi = max(hi, h); }
        }
      }
      int glo = lo, ghi = hi;
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) {
        glo = min(glo, __shfl_xor_sync(0xffffffffu, glo, o));
        ghi = max(ghi, __shfl_xor_sync(0xffffffffu, ghi, o));
      }
      if (ghi < 0) { glo = 0; ghi = -1; }
      if (owner) {
        const int npad = isx ? ((ghi - glo + 4) & ~3) : (ghi - glo + 1);   // x table zero-padded to 4 columns
        for (int c = 0; c < npad; c++) tab[c * 8 + bi] = 0.f;
        if (b_ok) {
          for (int i = 0; i < g; i++) {
            const float v = fadd(base, fdiv(fmul((float)i + .5f, bin), (float)g));
            int l, h; float fl, fh;
            if (axis_setup(v, size, l, h, fl, fh)) { tab[(l - glo) * 8 + bi] += fh; tab[(h - glo) * 8 + bi] += fl; }
          }
        }
        if (!isx) { mt.ylo[bi] = lo; mt.yhi[bi] = hi; }
      }
      const int xmin = __shfl_sync(0xffffffffu, glo, 0), xmax = __shfl_sync(0xffffffffu, ghi, 0);
      const int ymin = __shfl_sync(0xffffffffu, glo, 8), ymax = __shfl_sync(0xffffffffu, ghi, 8);
      if (lane == 0) {
        mt.xmin = xmin; mt.xmax = xmax; mt.ymin = ymin; mt.skip = skip ? 1 : 0;
        mt.inv_count = 1.0f / (float)cnt; mt.img = b_ok ? b : 0;
      }
      nbar_arrive(BAR_FULL0 + buf, RA_THREADS);
      // pull the RoI's feature patch into L1 one iteration before the compute warps ask for it
      if (b_ok && xmax >= xmin && ymax >= ymin) {
        const int row_lines = ((xmax - xmin + 1) * C * (int)sizeof(TIn) + 127) / 128;
        const int nrows = ymax - ymin + 1;
        int total = row_lines * nrows;
        if (total > 256) total = 256;
        const char* pbase = reinterpret_cast<const char*>(feat + ((size_t)b * H * W + (size_t)ymin * W + xmin) * C);
        for (int i = lane; i < total; i += 32) {
          const int rr = i / row_lines, ll = i - rr * row_lines;
          asm volatile("prefetch.global.L1 [%0];" ::"l"(pbase + ((size_t)rr * W * C) * sizeof(TIn) + (size_t)ll * 128));
        }
      }
    }
    return;
  }

  // -------------------------------------------------------------------- compute warps (warp = output row ph)
  const int ph = warp;
  for (int it = 0; it < n_iter; it++) {
    const int roi = blockIdx.x + it * gridDim.x, buf = it & 1;
    nbar_sync(BAR_FULL0 + buf, RA_THREADS);
    const float* wx = smem_dyn + buf * tab_floats;
    const float* wy = wx + (W + 4) * 8;
    const RoiMeta& mt = meta[buf];
    const int xmin = mt.xmin, xmax = mt.xmax, ymin = mt.ymin;
    const int ylo = mt.ylo[ph], yhi = mt.yhi[ph];
    const float inv_count = mt.inv_count;
    const bool skip = mt.skip != 0;
    const TIn* fb = feat + (size_t)mt.img * H * W * C;

    if (!skip) {
      for (int c0 = lane * 8; c0 - lane * 8 < C; c0 += 256) {
        float acc[P7][8];
#pragma unroll
        for (int pw = 0; pw < P7; pw++)
#pragma unroll
          for (int j = 0; j < 8; j++) acc[pw][j] = 0.f;
        if (c0 < C) {
          for (int cc = xmin; cc <= xmax; cc += 4) {
            float t[4][8];
#pragma unroll
            for (int q = 0; q < 4; q++)
#pragma unroll
              for (int j = 0; j < 8; j++) t[q][j] = 0.f;
            // columns past xmax are clamped onto xmax: their (zero-padded) weights discard them
            int coff[4];
#pragma unroll
            for (int q = 0; q < 4; q++) coff[q] = (min(cc + q, xmax) - cc) * C;
#pragma unroll 2
            for (int row = ylo; row <= yhi; row++) {
              const float wyv = wy[(row - ymin) * 8 + ph];
              const TIn* prow = fb + ((size_t)row * W + cc) * C + c0;
#pragma unroll
              for (int q = 0; q < 4; q++) {
                const F8 v = load8(prow + coff[q]);
#pragma unroll
                for (int j = 0; j < 8; j++) t[q][j] = fmaf(wyv, v.v[j], t[q][j]);
              }
            }
#pragma unroll
            for (int q = 0; q < 4; q++) {
              const float4* wp = reinterpret_cast<const float4*>(wx + (cc + q - xmin) * 8);
              const float4 wa = wp[0], wb = wp[1];
              const float wv[P7] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z};
#pragma unroll
              for (int pw = 0; pw < P7; pw++)
#pragma unroll
                for (int j = 0; j < 8; j++) acc[pw][j] = fmaf(wv[pw], t[q][j], acc[pw][j]);
            }
          }
          store_bins<MODE>(out, ld_out, roi, ph, C, c0, acc, inv_count, stage);
        }
      }
    }
    if (it + 2 < n_iter) nbar_arrive(BAR_EMPTY0 + buf, RA_THREADS);
    if (MODE == OUT_F32_NCHW && !skip) {
      nbar_sync(BAR_COMPUTE, P7 * 32);
      flush_stage<MODE>(reinterpret_cast<float*>(out), roi, C, stage);
      nbar_sync(BAR_COMPUTE, P7 * 32);
    }
  }
}

// --------------------------------------------------------------------------------- rotated RoIAlign
template <typename TIn, int MODE>
__global__ void __launch_bounds__(P7 * 32, 2)
roi_align_rotated_fwd_kernel(const TIn* __restrict__ feat, const float* __restrict__ rois, void* __restrict__ out,
                             long long ld_out, int K, int B, int C, int H, int W, float scale,
                             int sampling_ratio, int aligned, int clockwise,
                             const int* __restrict__ roi_level, int level) {
  extern __shared__ float stage[];
  const int ph = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float off = aligned ? 0.5f : 0.f;
  for (int roi = blockIdx.x; roi < K; roi += gridDim.x) {
    if (roi_level != nullptr && roi_level[roi] != level) continue;
    const float* r = rois + (size_t)roi * 6;
    const int b = (int)__ldg(r);
    const float cx = fsub(fmul(__ldg(r + 1), scale), off), cy = fsub(fmul(__ldg(r + 2), scale), off);
    float rw = fmul(__ldg(r + 3), scale), rh = fmul(__ldg(r + 4), scale);
    float theta = __ldg(r + 5);
    if (clockwise) theta = -theta;
    if (!aligned) { rw = fmaxf(rw, 1.f); rh = fmaxf(rh, 1.f); }
    const float bh = fdiv(rh, (float)P7), bw = fdiv(rw, (float)P7);
    const int gh = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(bh);
    const int gw = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(bw);
    const float sh = fdiv(-rh, 2.0f), sw = fdiv(-rw, 2.0f);
    const float ct = cosf(theta), st = sinf(theta);
    const int cnt = gh * gw > 1 ? gh * gw : 1;
    const float inv_count = 1.0f / (float)cnt;
    const bool b_ok = b >= 0 && b < B;
    const TIn* fb = feat + (size_t)(b_ok ? b : 0) * H * W * C;

    for (int c0 = lane * 8; c0 < C; c0 += 256) {
      float acc[P7][8];
#pragma unroll
      for (int pw = 0; pw < P7; pw++)
#pragma unroll
        for (int j = 0; j < 8; j++) acc[pw][j] = 0.f;
      for (int iy = 0; iy < gh && b_ok; iy++) {
        const float yy = fadd(fadd(sh, fmul((float)ph, bh)), fdiv(fmul((float)iy + .5f, bh), (float)gh));
#pragma unroll
        for (int pw = 0; pw < P7; pw++) {
          const float xb = fadd(sw, fmul((float)pw, bw));
          for (int ix = 0; ix < gw; ix++) {
            const float xx = fadd(xb, fdiv(fmul((float)ix + .5f, bw), (float)gw));
            const float y = fadd(fsub(fmul(yy, ct), fmul(xx, st)), cy);
            const float x = fadd(fadd(fmul(yy, st), fmul(xx, ct)), cx);
            int yl, yh, xl, xh; float ly, hy, lx, hx;
            if (y < -1.0f || y > (float)H || x < -1.0f || x > (float)W) continue;
            axis_setup(y, H, yl, yh, ly, hy);
            axis_setup(x, W, xl, xh, lx, hx);
            const float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
            F8 f1 = load8(fb + ((size_t)yl * W + xl) * C + c0), f2 = load8(fb + ((size_t)yl * W + xh) * C + c0);
            F8 f3 = load8(fb + ((size_t)yh * W + xl) * C + c0), f4 = load8(fb + ((size_t)yh * W + xh) * C + c0);
#pragma unroll
            for (int j = 0; j < 8; j++) acc[pw][j] += w1 * f1.v[j] + w2 * f2.v[j] + w3 * f3.v[j] + w4 * f4.v[j];
          }
        }
      }
      store_bins<MODE>(out, ld_out, roi, ph, C, c0, acc, inv_count, stage);
    }
    if (MODE == OUT_F32_NCHW) {
      __syncthreads();
      flush_stage<MODE>(reinterpret_cast<float*>(out), roi, C, stage);
      __syncthreads();
    }
  }
}

// "Large RoIs only" pass behind the tensor-core kernel (roi_align_mma.cu leaves out every rotated RoI whose longer
// side exceeds ROT_BIG_THRESHOLD feature pixels; sampling_ratio is 1 or 2 there).  Few RoIs, so the kernel is
// organised for LATENCY:
//   * CTA (block b, column pw): the 32 lanes of every warp test 32 RoIs at once (b, b + NB, b + 2 NB, ...: strided,
//     so that the 400 consecutive negatives spread over all blocks) -- one L2 round trip to find the large ones;
//   * one work item = one output column (pw) of one large RoI, a warp per output row: the <= 4 samples of the bin are
//     set up first and their 16 gathers issued back to back.
// Same per-sample arithmetic and accumulation order (iy outer, ix inner) as roi_align_rotated_fwd_kernel.
// __launch_bounds__(224, 3): <= 96 registers, three CTAs per SM (at 162 registers / one CTA per SM the 8 waves of this
// grid ran back to back: 54 us for the 480 large RoIs of the classification pass).
template <typename TIn>
__global__ void __launch_bounds__(P7 * 32, 3)
roi_align_rotated_big_kernel(const TIn* __restrict__ feat, const float* __restrict__ rois, __nv_bfloat16* __restrict__ out,
                             long long ld_out, int K, int B, int C, int H, int W, float scale, int gs, int aligned,
                             int clockwise, const int* __restrict__ roi_level, int level, float only_larger_than, int NB) {
  const int ph = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int blk = blockIdx.x / P7, pw = blockIdx.x - blk * P7;
  const float off = aligned ? 0.5f : 0.f;
  const int cand = blk + lane * NB;
  bool big = false;
  if (cand < K && (roi_level == nullptr || roi_level[cand] == level)) {
    float rw = fmul(__ldg(rois + (size_t)cand * 6 + 3), scale), rh = fmul(__ldg(rois + (size_t)cand * 6 + 4), scale);
    if (!aligned) { rw = fmaxf(rw, 1.f); rh = fmaxf(rh, 1.f); }
    big = fmaxf(rw, rh) > only_larger_than;
  }
  unsigned todo = __ballot_sync(0xffffffffu, big);
  const int cnt = gs * gs;
  const float inv_count = 1.0f / (float)cnt;
  while (todo != 0u) {
    const int roi = blk + (__ffs((int)todo) - 1) * NB;
    todo &= todo - 1u;
    const float* r = rois + (size_t)roi * 6;
    const int b = (int)__ldg(r);
    const float cx = fsub(fmul(__ldg(r + 1), scale), off), cy = fsub(fmul(__ldg(r + 2), scale), off);
    float rw = fmul(__ldg(r + 3), scale), rh = fmul(__ldg(r + 4), scale);
    float theta = __ldg(r + 5);
    if (clockwise) theta = -theta;
    if (!aligned) { rw = fmaxf(rw, 1.f); rh = fmaxf(rh, 1.f); }
    const float bh = fdiv(rh, (float)P7), bw = fdiv(rw, (float)P7);
    const float sh = fdiv(-rh, 2.0f), sw = fdiv(-rw, 2.0f);
    const float ct = cosf(theta), st = sinf(theta);
    const bool b_ok = b >= 0 && b < B;
    const TIn* fb = feat + (size_t)(b_ok ? b : 0) * H * W * C;
    const float yb = fadd(sh, fmul((float)ph, bh)), xb = fadd(sw, fmul((float)pw, bw));
    float wt[4][4];
    uint32_t po[4][4];                             // element offsets inside this image's map (H * W * C < 2^32)
    bool ok[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int iy = gs == 2 ? (k >> 1) : 0, ix = gs == 2 ? (k & 1) : 0;
      const float yy = fadd(yb, fdiv(fmul((float)iy + .5f, bh), (float)gs));
      const float xx = fadd(xb, fdiv(fmul((float)ix + .5f, bw), (float)gs));
      const float y = fadd(fsub(fmul(yy, ct), fmul(xx, st)), cy);
      const float x = fadd(fadd(fmul(yy, st), fmul(xx, ct)), cx);
      ok[k] = b_ok && k < cnt && !(y < -1.0f || y > (float)H || x < -1.0f || x > (float)W);
      int yl = 0, yh = 0, xl = 0, xh = 0; float ly = 0.f, hy = 0.f, lx = 0.f, hx = 0.f;
      if (ok[k]) { axis_setup(y, H, yl, yh, ly, hy); axis_setup(x, W, xl, xh, lx, hx); }
      wt[k][0] = hy * hx; wt[k][1] = hy * lx; wt[k][2] = ly * hx; wt[k][3] = ly * lx;
      po[k][0] = (uint32_t)((yl * W + xl) * C); po[k][1] = (uint32_t)((yl * W + xh) * C);
      po[k][2] = (uint32_t)((yh * W + xl) * C); po[k][3] = (uint32_t)((yh * W + xh) * C);
    }
    for (int c0 = lane * 8; c0 < C; c0 += 256) {
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; j++) acc[j] = 0.f;
#pragma unroll
      for (int k0 = 0; k0 < 4; k0 += 2) {          // two samples = 8 gathers in flight per lane
        uint4 raw[2][4];
#pragma unroll
        for (int k = 0; k < 2; k++)
#pragma unroll
          for (int q = 0; q < 4; q++) raw[k][q] = raw8(fb + po[k0 + k][q] + c0);   // invalid samples read pixel (0, 0): harmless
#pragma unroll
        for (int k = 0; k < 2; k++) {
          if (ok[k0 + k]) {
            const F8 f1 = cvt8(raw[k][0], fb), f2 = cvt8(raw[k][1], fb), f3 = cvt8(raw[k][2], fb), f4 = cvt8(raw[k][3], fb);
#pragma unroll
            for (int j = 0; j < 8; j++)
              acc[j] += wt[k0 + k][0] * f1.v[j] + wt[k0 + k][1] * f2.v[j] + wt[k0 + k][2] * f3.v[j] + wt[k0 + k][3] * f4.v[j];
          }
        }
      }
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; j++) v[j] = acc[j] * inv_count;
      *reinterpret_cast<uint4*>(out + (size_t)roi * ld_out + (size_t)(ph * P7 + pw) * C + c0) =
          make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
    }
  }
}

// FPN level of each RoI: clamp(floor(log2(sqrt(w*h)/finest + 1e-6)), 0, L-1)
// (single_level_roi_extractor.py:35-54; rotated: sqrt(w*h) of columns 3,4, rotate_single_level_roi_extractor.py:84)
__global__ void map_roi_levels_kernel(const float* __restrict__ rois, int K, int rotated, float finest, int L,
                                      int* __restrict__ lvl) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K) return;
  const float* r = rois + (size_t)i * (rotated ? 6 : 5);
  const float area = rotated ? fmul(r[3], r[4]) : fmul(fsub(r[3], r[1]), fsub(r[4], r[2]));
  const float scale = sqrtf(area);
  float t = floorf(log2f(fadd(fdiv(scale, finest), 1e-6f)));
  t = fminf(fmaxf(t, 0.f), (float)(L - 1));   // NaN (negative area) clamps like torch: stays NaN -> long cast
  lvl[i] = (t == t) ? (int)t : 0;
}

// roi_rescale (base_roi_extractor.py:61-83; rotated: rotate_single_level_roi_extractor.py:150-167)
__global__ void roi_rescale_kernel(const float* __restrict__ rois, int K, int rotated, float fh, float fw,
                                   float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K) return;
  if (rotated) {
    const float* r = rois + (size_t)i * 6; float* o = out + (size_t)i * 6;
    o[0] = r[0]; o[1] = r[1]; o[2] = r[2]; o[3] = fmul(fw, r[3]); o[4] = fmul(fh, r[4]); o[5] = r[5];
  } else {
    const float* r = rois + (size_t)i * 5; float* o = out + (size_t)i * 5;
    const float cx = fmul(fadd(r[1], r[3]), 0.5f), cy = fmul(fadd(r[2], r[4]), 0.5f);
    const float nw = fmul(fsub(r[3], r[1]), fw), nh = fmul(fsub(r[4], r[2]), fh);
    o[0] = r[0]; o[1] = fsub(cx, fmul(nw, 0.5f)); o[2] = fsub(cy, fmul(nh, 0.5f));
    o[3] = fadd(cx, fmul(nw, 0.5f)); o[4] = fadd(cy, fmul(nh, 0.5f));
  }
}

template <typename TIn, int MODE>
static int launch_fwd(bool rotated, const void* feat, const float* rois, void* out, long long ld_out, int K, int B,
                      int C, int H, int W, float scale, int sampling_ratio, int aligned, int clockwise,
                      const int* roi_level, int level, cudaStream_t stream) {
  const size_t stage_bytes = MODE == OUT_F32_NCHW ? (size_t)P7 * P7 * (C + 1) * sizeof(float) : 0;
  const size_t tab_bytes = rotated ? 0 : 2 * ((size_t)(W + 4) * 8 + (size_t)(H + 1) * 8) * sizeof(float);
  size_t smem = stage_bytes + tab_bytes;
  if (smem > 100 * 1024) { set_error("roi_align: feature map %dx%d too large for the shared weight tables", H, W); return PT_ERR_UNSUPPORTED; }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int grid = K < sms * 8 ? K : sms * 8;
  if (rotated) {
    auto kern = roi_align_rotated_fwd_kernel<TIn, MODE>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<grid, P7 * 32, smem, stream>>>(reinterpret_cast<const TIn*>(feat), rois, out, ld_out, K, B, C, H, W,
                                           scale, sampling_ratio, aligned, clockwise, roi_level, level);
  } else {
    auto kern = roi_align_fwd_kernel<TIn, MODE>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<grid, RA_THREADS, smem, stream>>>(reinterpret_cast<const TIn*>(feat), rois, out, ld_out, K, B, C, H, W,
                                              scale, sampling_ratio, aligned, roi_level, level);
  }
  return check_launch(rotated ? "roi_align_rotated_fwd_kernel" : "roi_align_fwd_kernel");
}

}  // namespace ptb

using namespace ptb;

extern "C" int pt_nchw_to_nhwc_ex(const float* in, void* out, int B, int C, int H, int W, int out_bf16,
                                  int* sat_count, void* stream) {
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return PT_OK;
  const int HW = H * W;
  dim3 grid((HW + 63) / 64, (C + 63) / 64, B), block(256);
  if (out_bf16 == 2)
    nchw_to_nhwc_kernel<__half><<<grid, block, 0, (cudaStream_t)stream>>>(in, (__half*)out, C, HW, sat_count);
  else if (out_bf16)
    nchw_to_nhwc_kernel<__nv_bfloat16><<<grid, block, 0, (cudaStream_t)stream>>>(in, (__nv_bfloat16*)out, C, HW, nullptr);
  else
    nchw_to_nhwc_kernel<float><<<grid, block, 0, (cudaStream_t)stream>>>(in, (float*)out, C, HW, nullptr);
  return check_launch("nchw_to_nhwc_kernel");
}
extern "C" int pt_nchw_to_nhwc(const float* in, void* out, int B, int C, int H, int W, int out_bf16, void* stream) {
  return pt_nchw_to_nhwc_ex(in, out, B, C, H, W, out_bf16, nullptr, stream);
}

// feat: NHWC [B,H,W,C] fp32 (feat_bf16=0), bf16 (=1) or fp16 (=2).  rois: [K,5] (b,x1,y1,x2,y2) or, rotated, [K,6]
// (b,cx,cy,w,h,theta).  out_mode 0: bf16 [K, ld_out] with k = (ph*7+pw)*C + c;  1: fp32 [K,C,7,7];
// 2: bf16 [K, ld_out] three segments [hi | lo | hi] of 49*C each.
extern "C" int pt_roi_align_forward(const void* feat, int feat_bf16, const float* rois, void* out, long long ld_out,
                                    int out_mode, int K, int B, int C, int H, int W, int pooled, float spatial_scale,
                                    int sampling_ratio, int aligned, int rotated, int clockwise,
                                    const int* roi_level, int level, void* stream) {
  if (K <= 0) return PT_OK;
  if (pooled != P7) { set_error("pt_roi_align_forward: only output_size=7 is built (got %d)", pooled); return PT_ERR_UNSUPPORTED; }
  if (C % 8 != 0) { set_error("pt_roi_align_forward: C must be a multiple of 8 (got %d)", C); return PT_ERR_ARG; }
  if (out_mode == OUT_F32_NCHW && C % 8 != 0) return PT_ERR_ARG;
  if (out_mode != OUT_F32_NCHW) {
    const long long need = (out_mode == OUT_BF16X3_BINMAJOR ? 3LL : 1LL) * P7 * P7 * C;
    if (ld_out < need || (ld_out % 8) != 0) { set_error("pt_roi_align_forward: ld_out %lld too small / unaligned", ld_out); return PT_ERR_ARG; }
  }
  cudaStream_t s = (cudaStream_t)stream;
  const bool rot = rotated != 0;
  if (feat_bf16 && out_mode == OUT_BF16_BINMAJOR && ramma::supported(C, H, W, rot, sampling_ratio)) {
    const int rc = ramma::launch(feat, feat_bf16 == 2, rois, out, ld_out, K, B, C, H, W, spatial_scale, sampling_ratio,
                                 aligned, roi_level, level, rot, clockwise, s);
    if (!rot || rc != PT_OK) return rc;
    // rotated: RoIs larger than ROT_BIG_THRESHOLD feature pixels were left out above (sparse in the chunked
    // formulation); the direct gather kernel fills exactly those rows
    const int NB = (K + 31) / 32;
    if (feat_bf16 == 2)
      roi_align_rotated_big_kernel<__half><<<NB * P7, P7 * 32, 0, s>>>(
          reinterpret_cast<const __half*>(feat), rois, reinterpret_cast<__nv_bfloat16*>(out), ld_out, K, B, C, H, W,
          spatial_scale, sampling_ratio, aligned, clockwise, roi_level, level, ROT_BIG_THRESHOLD, NB);
    else
      roi_align_rotated_big_kernel<__nv_bfloat16><<<NB * P7, P7 * 32, 0, s>>>(
          reinterpret_cast<const __nv_bfloat16*>(feat), rois, reinterpret_cast<__nv_bfloat16*>(out), ld_out, K, B, C, H, W,
          spatial_scale, sampling_ratio, aligned, clockwise, roi_level, level, ROT_BIG_THRESHOLD, NB);
    return check_launch("roi_align_rotated_big_kernel");
  }
#define PT_DISPATCH(T, M) return launch_fwd<T, M>(rot, feat, rois, out, ld_out, K, B, C, H, W, spatial_scale, sampling_ratio, aligned, clockwise, roi_level, level, s)
  if (feat_bf16 == 2) {
    if (out_mode == OUT_BF16_BINMAJOR) PT_DISPATCH(__half, OUT_BF16_BINMAJOR);
    if (out_mode == OUT_F32_NCHW) PT_DISPATCH(__half, OUT_F32_NCHW);
    if (out_mode == OUT_BF16X3_BINMAJOR) PT_DISPATCH(__half, OUT_BF16X3_BINMAJOR);
  } else if (feat_bf16) {
    if (out_mode == OUT_BF16_BINMAJOR) PT_DISPATCH(__nv_bfloat16, OUT_BF16_BINMAJOR);
    if (out_mode == OUT_F32_NCHW) PT_DISPATCH(__nv_bfloat16, OUT_F32_NCHW);
    if (out_mode == OUT_BF16X3_BINMAJOR) PT_DISPATCH(__nv_bfloat16, OUT_BF16X3_BINMAJOR);
  } else {
    if (out_mode == OUT_BF16_BINMAJOR) PT_DISPATCH(float, OUT_BF16_BINMAJOR);
    if (out_mode == OUT_F32_NCHW) PT_DISPATCH(float, OUT_F32_NCHW);
    if (out_mode == OUT_BF16X3_BINMAJOR) PT_DISPATCH(float, OUT_BF16X3_BINMAJOR);
  }
#undef PT_DISPATCH
  set_error("pt_roi_align_forward: bad out_mode %d", out_mode);
  return PT_ERR_ARG;
}

extern "C" int pt_map_roi_levels(const float* rois, int K, int rotated, float finest_scale, int num_levels,
                                 int* levels, void* stream) {
  if (K <= 0) return PT_OK;
  map_roi_levels_kernel<<<(K + 255) / 256, 256, 0, (cudaStream_t)stream>>>(rois, K, rotated, finest_scale,
                                                                           num_levels, levels);
  return check_launch("map_roi_levels_kernel");
}

extern "C" int pt_roi_rescale(const float* rois, int K, int rotated, float factor_h, float factor_w, float* out,
                              void* stream) {
  if (K <= 0) return PT_OK;
  roi_rescale_kernel<<<(K + 255) / 256, 256, 0, (cudaStream_t)stream>>>(rois, K, rotated, factor_h, factor_w, out);
  return check_launch("roi_rescale_kernel");
}
