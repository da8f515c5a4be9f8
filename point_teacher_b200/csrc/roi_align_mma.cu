// RoIAlign forward, bf16 throughput path for sm_100a: TMA patch loads + warp-level tensor-core interpolation.
//
// Same operator as roi_align.cu's roi_align_fwd_kernel (mmcv.ops.RoIAlign(output_size=7, sampling_ratio=0|n,
// pool_mode='avg', aligned) reached through
//   HBB_TOD/mmdet/models/roi_heads/roi_extractors/single_level_roi_extractor.py:56-114),
// specialised for the MIL path's operand: bf16 NHWC feature map in, bf16 bin-major (K, 49*C) FC1 operand out.
//
// Why tensor cores on an HBM-bound op: bilinear average pooling of one RoI is the small contraction
//     out[bin, c] = sum_pix Wmat[bin, pix] * patch[pix, c]          (49 x npix) * (npix x 256)
// with Wmat[bin=(ph,pw), pix=(row,col)] = wy[ph][row] * wx[pw][col] / count (the separable tables of the
// reference's sample grid).  Tiny objects touch a 3x3..4x4 pixel patch, so one m16n8k16 k-step covers a whole
// RoI.  Doing the contraction with mma.sync cuts the issue slots per RoI ~7x against the FFMA formulation
// (6 000 -> ~900 warp instructions), which is what lets the SMs saturate the 25 KB/RoI store stream.  The
// tensor pipe itself stays almost idle; the bound is the HBM/L2 write path.
//
// CTA = 8 MMA warps (32 channels each, all 49 bins -> 64 fp32 accumulators/thread) + 2 builder warps that take
// alternate RoIs (a single builder warp is latency-bound at ~4 700 cycles per RoI and starves the MMA warps).
//   builder : reads the RoI (prefetched one iteration ahead), builds wx/wy with the reference's exact (non-contracted) coordinate arithmetic,
//             walks the patch in 4x4-pixel chunks; per chunk it issues ONE 5-D TMA box load
//             (64 ch x 4 cols x 4 rows x 1 img x 4 channel-quarters = 8 KB, SWIZZLE_128B, OOB zero fill) into a
//             its own 2-deep mbarrier ring and writes the chunk's A fragments (Wmat in mma register order) next to it.
//   MMA     : per chunk 4 x LDS.128 (A) + 2 x ldmatrix.x4.trans (B) + 16 x mma.m16n8k16; after the RoI's last
//             chunk: bf16 pack -> stmatrix into a warp-private SWIZZLE_64B staging -> one TMA tensor store of the
//             warp's 49 bins x 32 channels (no CTA-wide barrier in the steady state; the output never touches
//             the LSU pipe, which the ncu captures showed to be the limiter of the LDS+STG copy-out).
//
// Rotated twin (ROT = true; mmcv.ops.RoIAlignRotated(sampling_ratio = 1|2, clockwise) reached through
//   OBB_TOD/mmrotate/models/roi_heads/roi_extractors/rotate_single_level_roi_extractor.py:90-167):
// the sample grid is not axis-separable, so the builder warp first writes one 16-byte record per sample
// (x_low|x_high, y_low|y_high, lx, ly -- the reference's exact coordinate arithmetic and border rule) into shared
// memory, then builds each 4x4-pixel chunk's Wmat fragment by summing the <= 4 samples of every bin it owns.  The
// MMA / store side is unchanged: only the weights differ.
#include <cuda.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace ptb {
namespace ramma {

constexpr int P7 = 7, NBIN = 49;
constexpr int BUILDERS = 2, DEPTH = 2, STAGES = BUILDERS * DEPTH;   // every builder warp owns a DEPTH-deep ring
constexpr int MMA_WARPS = 8;
constexpr int THREADS = (MMA_WARPS + BUILDERS) * 32;
// Rotated variant: the builder does ~4x the work per RoI (non-separable weights), so it runs ONE CTA per SM with 6
// builder warps (12 stages, ~208 KB of shared memory) instead of two CTAs with 2 builders each.  The ncu source view
// of the 2-builder variant showed the 8 MMA warps spinning on the full barriers 80 % of the time.  (8 builders with
// single-buffered output staging measured the same step time: beyond 6 the builders are no longer the limiter.)
constexpr int ROT_BUILDERS = 8, ROT_DEPTH = 13;    // rotated: builder warps, stages of the shared ring
constexpr int ROT_AHEAD = 4;                        // chunks whose TMA loads a builder fires ahead of its weight build
constexpr int QUARTER_BYTES = 16 * 128;           // 16 pixels x 64 channels bf16
constexpr int PATCH_BYTES = 4 * QUARTER_BYTES;    // 8 KB
constexpr int AFRAG_BYTES = 4 * 32 * 16;          // 4 m-tiles x 32 lanes x uint4
// warp-private output staging, double buffered: 49 bins x 32 channels (64 B rows) in the TMA SWIZZLE_64B layout
// (16 B chunk index ^= (row >> 1) & 3), which also makes the stmatrix writes bank-conflict free
constexpr int STG_ROW_BYTES = 64;
constexpr int STG_BYTES = ((NBIN * STG_ROW_BYTES + 511) / 512) * 512;   // per MMA warp and buffer
constexpr int ROT_SAMPLES = 4;                                          // rotated: sampling_ratio^2 <= 4 samples per bin
constexpr int ROT_MAP_WORDS = 32;                                       // occupancy bitmap: up to 1024 chunks per RoI
constexpr int ROT_W_STRIDE = 20;                                        // floats per W row (16 pixels, 80-byte pitch)
constexpr int ROT_TAB_FLOATS = NBIN * ROT_W_STRIDE + ROT_MAP_WORDS;     // per builder warp: W[49][20] + bitmap
enum { F_LAST = 2, F_ZERO = 4, F_SKIP = 8 };

__device__ __forceinline__ void tma_load_5d(uint32_t dst, const void* desc, uint32_t bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(desc), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void stsm_x4(uint32_t addr, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3) {
  asm volatile("stmatrix.sync.aligned.m8n8.x4.shared.b16 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(r0), "r"(r1),
               "r"(r2), "r"(r3)
               : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t pack_f16(float lo, float hi) {
  __half2 v = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void mma_f16_z(float (&d)[4], const uint4& a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%10, %10, %10, %10};"
      : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
      : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1), "f"(0.f));
}
__device__ __forceinline__ void mma_f16(float (&d)[4], const uint4& a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}
// D = A * B (no accumulator input: the first chunk of a RoI starts from zero without clearing registers)
__device__ __forceinline__ void mma_bf16_z(float (&d)[4], const uint4& a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%10, %10, %10, %10};"
      : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
      : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1), "f"(0.f));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint4& a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}

// one axis of the Detectron2 bilinear rule (same as roi_align.cu)
__device__ __forceinline__ bool axis_setup(float v, int size, int& lo, int& hi, float& l, float& h) {
  if (v < -1.0f || v > (float)size) return false;
  if (v <= 0.f) v = 0.f;
  lo = (int)v;
  if (lo >= size - 1) { hi = lo = size - 1; v = (float)lo; } else hi = lo + 1;
  l = fsub(v, (float)lo);
  h = fsub(1.0f, l);
  return true;
}

// F16: the feature map (and therefore the interpolation weights) are fp16 instead of bf16 -- 3 more mantissa
// bits on both mma operands, so the interpolation itself adds no visible error on top of the bf16 output.
template <bool F16, bool ROT, int NB, int DP>      // NB builder warps, each owning a DP-deep stage ring
__global__ void __launch_bounds__((MMA_WARPS + NB) * 32, ROT ? 1 : 2)
roi_align_mma_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap omap,
                     const float* __restrict__ rois, int K, int B, int C, int H, int W,
                     float scale, int sampling_ratio, int aligned, const int* __restrict__ roi_level, int level,
                     int stg_bufs, int clockwise) {
  extern __shared__ uint8_t smem_raw[];
  // horizontal: every builder warp owns a DP-deep ring (RoIs of 1..2 chunks).  rotated: ONE ring of DP stages shared by
  // all builders and consumed strictly in order -- a RoI's chunks occupy consecutive stages starting at the prefix sum
  // of the chunk counts of all earlier RoIs of this CTA (published builder to builder through a sequence word), so a
  // many-chunk RoI can use the whole ring and eight builders fit beside it in shared memory.
  constexpr int NSTAGES = ROT ? DP : NB * DP;
  // shared-window byte addresses (explicit .shared accesses below; generic pointers would cost LD/ST.E)
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t s_patch = sbase;                                    // [STAGES][PATCH_BYTES], 1 KB aligned
  const uint32_t s_afrag = s_patch + NSTAGES * PATCH_BYTES;          // [STAGES][AFRAG_BYTES]
  const uint32_t s_stg = s_afrag + NSTAGES * AFRAG_BYTES;            // [MMA_WARPS][2][STG_BYTES]
  const uint32_t s_tab = s_stg + MMA_WARPS * stg_bufs * STG_BYTES;                      // [BUILDERS][(W+4)*8 + (H+4)*8] floats
  const int tab_floats = ROT ? ROT_TAB_FLOATS : (W + 4) * 8 + (H + 4) * 8;
  const uint32_t s_full = s_tab + NB * tab_floats * 4;               // [STAGES] mbarriers
  const uint32_t s_empty = s_full + NSTAGES * 8;
  const uint32_t s_meta = s_empty + NSTAGES * 8;                     // [STAGES] {roi, flags}
  uint8_t* gen = smem_raw + (sbase - smem_u32(smem_raw));            // generic view of the same window
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(gen + (s_full - sbase));
  uint64_t* empty_bar = reinterpret_cast<uint64_t*>(gen + (s_empty - sbase));
  volatile int* seq_g = reinterpret_cast<volatile int*>(gen + (s_meta + NSTAGES * 8 - sbase));   // {RoIs published, stages}

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap);
    tma_prefetch_desc(&omap);
    // rotated: a stage is full after TWO arrivals -- the TMA issue (expect_tx, as soon as the stage is acquired) and
    // the fragment write (later): the patch flies while the builder still computes the chunk's weights
    for (int i = 0; i < NSTAGES; i++) { mbar_init(full_bar + i, ROT ? 2 : 1); mbar_init(empty_bar + i, MMA_WARPS); }
    seq_g[0] = 0; seq_g[1] = 0; seq_g[2] = 0;
    fence_barrier_init();
  }
  __syncthreads();
  const int n_iter = blockIdx.x < K ? (K - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int g = lane >> 2, t = lane & 3;

  if (warp >= MMA_WARPS) {
    // ---------------------------------------------------------------------------------- builder warps
    const int bw_id = warp - MMA_WARPS;                 // RoIs bw_id, bw_id + BUILDERS, ... of this CTA
    if constexpr (ROT) {
      // Lane-owned bins: lane owns bins `lane` and `lane + 32` (< 49).  Its <= 8 sample records live in registers;
      // per occupied 4x4-pixel chunk the lane accumulates the 16 pixel weights of each owned bin in registers
      // (row weights x column weights, fully unrolled: no shared-memory latency in the chain), writes the two rows
      // into a small shared W[bin][16] table, and the warp re-reads that table in mma fragment order.
      const uint32_t s_w = s_tab + (uint32_t)(bw_id * tab_floats) * 4u;            // [NBIN][ROT_W_STRIDE] floats
      // chunk-occupancy bitmap: a rotated RoI's bounding box may span many chunks that no sample touches
      unsigned int* map_g = reinterpret_cast<unsigned int*>(gen + (s_w + NBIN * ROT_W_STRIDE * 4 - sbase));
      const float off = aligned ? 0.5f : 0.f;
      const uint32_t tx_bytes = (uint32_t)(C / 64) * QUARTER_BYTES;
      const int gs = sampling_ratio, cnt = gs * gs;        // host guarantees 1 <= sampling_ratio <= 2
      const float inv_count = 1.0f / (float)cnt;
      float rnext = 0.f;
      if (bw_id < n_iter && lane < 6) rnext = __ldg(rois + (size_t)(blockIdx.x + bw_id * gridDim.x) * 6 + lane);
      for (int it = bw_id; it < n_iter; it += NB) {
        const int roi = blockIdx.x + it * gridDim.x;
        const float rcur = rnext;
        if (it + NB < n_iter && lane < 6)
          rnext = __ldg(rois + (size_t)(blockIdx.x + (it + NB) * gridDim.x) * 6 + lane);
        bool skip = roi_level != nullptr && roi_level[roi] != level;
        const int b = (int)__shfl_sync(0xffffffffu, rcur, 0);
        const float cxr = fsub(fmul(__shfl_sync(0xffffffffu, rcur, 1), scale), off);
        const float cyr = fsub(fmul(__shfl_sync(0xffffffffu, rcur, 2), scale), off);
        float rw = fmul(__shfl_sync(0xffffffffu, rcur, 3), scale), rh = fmul(__shfl_sync(0xffffffffu, rcur, 4), scale);
        float theta = __shfl_sync(0xffffffffu, rcur, 5);
        if (clockwise) theta = -theta;
        if (!aligned) { rw = fmaxf(rw, 1.f); rh = fmaxf(rh, 1.f); }
        // large RoIs (the 200-px random negatives) are sparse in this formulation -- 196 samples scattered over
        // hundreds of chunks; they are left to roi_align.cu's direct gather kernel, launched right after this one
        // with the same test
        if (fmaxf(rw, rh) > ROT_BIG_THRESHOLD) skip = true;
        const float bh = fdiv(rh, (float)P7), bw = fdiv(rw, (float)P7);
        const float sh = fdiv(-rh, 2.0f), sw = fdiv(-rw, 2.0f);
        const float ct = cosf(theta), st = sinf(theta);
        const bool b_ok = b >= 0 && b < B && !skip;
        // (i + .5) * bin / gs for i = 0, 1: the same values the per-sample expression of the reference produces
        const float ys0 = fdiv(fmul(.5f, bh), (float)gs), ys1 = fdiv(fmul(1.5f, bh), (float)gs);
        const float xs0 = fdiv(fmul(.5f, bw), (float)gs), xs1 = fdiv(fmul(1.5f, bw), (float)gs);
        int xlo = 1 << 30, xhi = -1, ylo = 1 << 30, yhi = -1;
        // records: (x_low | x_high << 16, y_low | y_high << 16, lx, ly); 0x7fff never matches a pixel coordinate
        uint4 rec[2 * ROT_SAMPLES];
#pragma unroll
        for (int j = 0; j < 2; j++) {
          const int bin = lane + 32 * j;
          const int ph = bin / P7, pw = bin - ph * P7;
          const float yb = fadd(sh, fmul((float)ph, bh)), xb = fadd(sw, fmul((float)pw, bw));
#pragma unroll
          for (int sub = 0; sub < ROT_SAMPLES; sub++) {
            // branch-free (selects only) so that the 8 unrolled samples of a lane interleave: the builder is a single
            // warp and lives on instruction-level parallelism.  Same arithmetic and border rule as axis_setup().
            const int iy = gs == 2 ? (sub >> 1) : 0, ix = gs == 2 ? (sub & 1) : 0;
            const float yy = fadd(yb, iy ? ys1 : ys0);
            const float xx = fadd(xb, ix ? xs1 : xs0);
            const float y = fadd(fsub(fmul(yy, ct), fmul(xx, st)), cyr);
            const float x = fadd(fadd(fmul(yy, st), fmul(xx, ct)), cxr);
            const bool ok = b_ok && bin < NBIN && sub < cnt && !(y < -1.0f || y > (float)H || x < -1.0f || x > (float)W);
            float yc = y <= 0.f ? 0.f : y, xc = x <= 0.f ? 0.f : x;
            int yl = (int)yc, xl = (int)xc;
            const bool ytop = yl >= H - 1, xtop = xl >= W - 1;
            yl = ytop ? H - 1 : yl; xl = xtop ? W - 1 : xl;
            const int yh = ytop ? yl : yl + 1, xh = xtop ? xl : xl + 1;
            yc = ytop ? (float)yl : yc; xc = xtop ? (float)xl : xc;
            const float ly = fsub(yc, (float)yl), lx = fsub(xc, (float)xl);
            uint4 r;
            r.x = ok ? ((uint32_t)xl | ((uint32_t)xh << 16)) : 0x7fff7fffu;      // 0x7fff never matches a pixel coordinate
            r.y = ok ? ((uint32_t)yl | ((uint32_t)yh << 16)) : 0x7fff7fffu;
            r.z = __float_as_uint(lx); r.w = __float_as_uint(ly);
            xlo = ok ? min(xlo, xl) : xlo; xhi = ok ? max(xhi, xh) : xhi;
            ylo = ok ? min(ylo, yl) : ylo; yhi = ok ? max(yhi, yh) : yhi;
            rec[j * ROT_SAMPLES + sub] = r;
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          xlo = min(xlo, __shfl_xor_sync(0xffffffffu, xlo, o)); xhi = max(xhi, __shfl_xor_sync(0xffffffffu, xhi, o));
          ylo = min(ylo, __shfl_xor_sync(0xffffffffu, ylo, o)); yhi = max(yhi, __shfl_xor_sync(0xffffffffu, yhi, o));
        }
        const bool empty = !b_ok || xhi < xlo || yhi < ylo;
        const int ncx = empty ? 1 : (xhi - xlo) / 4 + 1, ncy = empty ? 1 : (yhi - ylo) / 4 + 1;
        const int nch = ncx * ncy;
        // few chunks (the common case): walk them all -- an untouched chunk only costs a zero-weight pass
        const bool use_map = !empty && nch > 6 && nch <= ROT_MAP_WORDS * 32;
        const int nwords = (nch + 31) >> 5;
        int total = nch;
        if (use_map) {
          __syncwarp();   // the previous RoI's chunk walk is done reading the bitmap
          for (int w = lane; w < nwords; w += 32) map_g[w] = 0u;
          __syncwarp();
#pragma unroll
          for (int q = 0; q < 2 * ROT_SAMPLES; q++) {
            const uint4 r = rec[q];
            if (r.x != 0x7fff7fffu) {
              const int cxl = ((int)(r.x & 0xffffu) - xlo) >> 2, cxh = ((int)(r.x >> 16) - xlo) >> 2;
              const int cyl = ((int)(r.y & 0xffffu) - ylo) >> 2, cyh = ((int)(r.y >> 16) - ylo) >> 2;
              int ch = cyl * ncx + cxl;
              atomicOr(map_g + (ch >> 5), 1u << (ch & 31));
              if (cxh != cxl) { ch = cyl * ncx + cxh; atomicOr(map_g + (ch >> 5), 1u << (ch & 31)); }
              if (cyh != cyl) {
                ch = cyh * ncx + cxl; atomicOr(map_g + (ch >> 5), 1u << (ch & 31));
                if (cxh != cxl) { ch = cyh * ncx + cxh; atomicOr(map_g + (ch >> 5), 1u << (ch & 31)); }
              }
            }
          }
          __syncwarp();
          int pc = 0;
          for (int w = lane; w < nwords; w += 32) pc += __popc(map_g[w]);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) pc += __shfl_xor_sync(0xffffffffu, pc, o);
          total = pc;
        }
        // Two cursors walk the RoI's chunk list: the ISSUE cursor acquires a stage and fires the chunk's TMA load at
        // once (the box coordinates are all it needs), up to DP chunks ahead of the BUILD cursor, which computes the
        // chunk's weights and writes its fragments into the already acquired stage.  The MMA warps consume RoIs in
        // order, so what they used to wait for on every chunk -- fragment build, then TMA issue, then the load's L2
        // latency -- is now off their path for every RoI of up to DP chunks.
        auto word = [&](int w) -> uint32_t { return use_map ? map_g[w] : 0xffffffffu; };
        auto next = [&](int& w, uint32_t& bits) -> int {
          for (;;) {
            if (w >= nwords) return -1;
            if (bits != 0u) {
              const int chn = w * 32 + __ffs((int)bits) - 1;
              bits &= bits - 1u;
              if (chn >= nch) { w = nwords; return -1; }
              return chn;
            }
            if (++w < nwords) bits = word(w);
          }
        };
        // this RoI's first stage in the shared ring = stages used by all earlier RoIs of the CTA
        int base_j = 0;
        if (lane == 0) {
          while (seq_g[0] != it) __nanosleep(32);
          base_j = seq_g[1];
          seq_g[1] = base_j + total;
          __threadfence_block();
          seq_g[0] = it + 1;
        }
        base_j = __shfl_sync(0xffffffffu, base_j, 0);
        // mbarrier waits only see a phase PARITY: a builder two laps ahead of the consumers would take the parity of
        // lap k-3 for lap k-1.  Builders of many-chunk RoIs can get that far ahead, so the acquire first waits until MMA
        // warp 0 has consumed stage j - DP (a plain shared word), which pins the barrier to lap k-1 or k.
        auto acquire = [&](int j) {
          const int stage = j % DP;
          if (j >= DP) {
            if (lane == 0) { while (seq_g[2] < j - DP + 1) __nanosleep(64); }
            __syncwarp();
          }
          mbar_wait(empty_bar + stage, (uint32_t)(((j / DP) & 1) ^ 1));
          return stage;
        };
        if (empty) {
          const int stage = acquire(base_j);
          if (lane == 0) {
            asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(s_meta + stage * 8), "r"(roi), "r"(F_LAST | (skip ? F_SKIP : F_ZERO)) : "memory");
            mbar_arrive(full_bar + stage);
            mbar_arrive(full_bar + stage);
          }
          continue;
        }
        int wi = 0, wb = 0;
        uint32_t bits_i = word(0), bits_b = bits_i;
        int issued = 0, built = 0;
        auto issue = [&]() {
          const int chn = next(wi, bits_i);
          const int cy = chn / ncx, cx = chn - cy * ncx;
          const int stage = acquire(base_j + issued);
          if (lane == 0) {
            mbar_expect_tx(full_bar + stage, tx_bytes);
            tma_load_5d(s_patch + stage * PATCH_BYTES, &tmap, s_full + stage * 8, 0, xlo + cx * 4, ylo + cy * 4, b, 0);
          }
          issued++;
        };
        while (issued < total && issued < ROT_AHEAD) issue();
        while (built < total) {
          const int chn = next(wb, bits_b);
          const int cy = chn / ncx, cx = chn - cy * ncx;
          const int x0 = xlo + cx * 4, y0 = ylo + cy * 4;
          __syncwarp();   // the previous chunk's fragment gather is done reading W
#pragma unroll
          for (int j = 0; j < 2; j++) {
            const int bin = lane + 32 * j;
            if (bin < NBIN) {
              float wv[16];
#pragma unroll
              for (int i = 0; i < 16; i++) wv[i] = 0.f;
#pragma unroll
              for (int sub = 0; sub < ROT_SAMPLES; sub++) {
                const uint4 r = rec[j * ROT_SAMPLES + sub];
                const int dxl = (int)(r.x & 0xffffu) - x0, dxh = (int)(r.x >> 16) - x0;
                const int dyl = (int)(r.y & 0xffffu) - y0, dyh = (int)(r.y >> 16) - y0;
                const float lx = __uint_as_float(r.z), ly = __uint_as_float(r.w);
                const float hx = fsub(1.0f, lx), hy = fsub(1.0f, ly);
                float cw[4], rwt[4];
#pragma unroll
                for (int c = 0; c < 4; c++) {
                  cw[c] = (dxl == c ? hx : 0.f) + (dxh == c ? lx : 0.f);
                  rwt[c] = (dyl == c ? hy : 0.f) + (dyh == c ? ly : 0.f);
                }
#pragma unroll
                for (int rr = 0; rr < 4; rr++)
#pragma unroll
                  for (int c = 0; c < 4; c++) wv[rr * 4 + c] += rwt[rr] * cw[c];
              }
              const uint32_t wrow = s_w + (uint32_t)bin * (ROT_W_STRIDE * 4);
#pragma unroll
              for (int i = 0; i < 4; i++)
                sts128(wrow + i * 16, make_uint4(__float_as_uint(wv[4 * i] * inv_count), __float_as_uint(wv[4 * i + 1] * inv_count),
                                                 __float_as_uint(wv[4 * i + 2] * inv_count), __float_as_uint(wv[4 * i + 3] * inv_count)));
            }
          }
          __syncwarp();
          const int stage = (base_j + built) % DP;
          {
            // fragment order: rows g / g + 8 of m-tile mt, pixels k = 2t, 2t+1 (chunk row t>>1) and 2t+8, 2t+9 (row +2)
            const uint32_t dst = s_afrag + stage * AFRAG_BYTES + lane * 16;
            const uint32_t pix = (uint32_t)(((t >> 1) * 4 + 2 * (t & 1)) * 4);
#pragma unroll
            for (int mt = 0; mt < 4; mt++) {
              uint32_t a[4];
#pragma unroll
              for (int hl = 0; hl < 2; hl++) {
                const int bin = mt * 16 + g + hl * 8;
                float2 lo = make_float2(0.f, 0.f), hi = make_float2(0.f, 0.f);
                if (bin < NBIN) {
                  const uint2 u0 = lds64(s_w + (uint32_t)bin * (ROT_W_STRIDE * 4) + pix);
                  const uint2 u1 = lds64(s_w + (uint32_t)bin * (ROT_W_STRIDE * 4) + pix + 32);
                  lo = make_float2(__uint_as_float(u0.x), __uint_as_float(u0.y));
                  hi = make_float2(__uint_as_float(u1.x), __uint_as_float(u1.y));
                }
                a[hl] = F16 ? pack_f16(lo.x, lo.y) : pack_bf16(lo.x, lo.y);
                a[2 + hl] = F16 ? pack_f16(hi.x, hi.y) : pack_bf16(hi.x, hi.y);
              }
              sts128(dst + mt * 512, make_uint4(a[0], a[1], a[2], a[3]));
            }
          }
          built++;
          if (lane == 0) asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(s_meta + stage * 8), "r"(roi), "r"(built == total ? (int)F_LAST : 0) : "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(full_bar + stage);        // second arrival: fragments + meta are in place
          if (issued < total) issue();
        }
      }
      return;
    }
    float* wx = reinterpret_cast<float*>(gen + (s_tab - sbase)) + bw_id * tab_floats;   // wx[col - xmin][pw]
    float* wy = wx + (W + 4) * 8;                                                        // wy[row - ymin][ph]
    const float off = aligned ? 0.5f : 0.f;
    const uint32_t tx_bytes = (uint32_t)(C / 64) * QUARTER_BYTES;
    int slot = 0; uint32_t phase = 0;
    // bins owned by this lane in the A fragments: m-tile mt -> rows mt*16+g (lo) and +8 (hi)
    int phs[8], pws[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const int b = (i >> 1) * 16 + g + (i & 1) * 8;
      phs[i] = b < NBIN ? b / P7 : -1;
      pws[i] = b < NBIN ? b % P7 : 0;
    }
    // software prefetch of the RoI record one iteration ahead (lanes 0..4 hold the 5 floats)
    float rnext = 0.f;
    if (bw_id < n_iter && lane < 5) rnext = __ldg(rois + (size_t)(blockIdx.x + bw_id * gridDim.x) * 5 + lane);
    for (int it = bw_id; it < n_iter; it += NB) {
      const int roi = blockIdx.x + it * gridDim.x;
      const float rcur = rnext;
      if (it + NB < n_iter && lane < 5)
        rnext = __ldg(rois + (size_t)(blockIdx.x + (it + NB) * gridDim.x) * 5 + lane);
      const bool skip = roi_level != nullptr && roi_level[roi] != level;   // another FPN level owns this RoI
      const int b = (int)__shfl_sync(0xffffffffu, rcur, 0);
      const float x1 = fsub(fmul(__shfl_sync(0xffffffffu, rcur, 1), scale), off);
      const float y1 = fsub(fmul(__shfl_sync(0xffffffffu, rcur, 2), scale), off);
      const float x2 = fsub(fmul(__shfl_sync(0xffffffffu, rcur, 3), scale), off);
      const float y2 = fsub(fmul(__shfl_sync(0xffffffffu, rcur, 4), scale), off);
      float rw = fsub(x2, x1), rh = fsub(y2, y1);
      if (!aligned) { rw = fmaxf(rw, 1.f); rh = fmaxf(rh, 1.f); }
      // lanes 0..6 own the x bins, lanes 8..14 the y bins (lanes 16..31 mirror them so that every lane sees
      // both bin sizes through one xor-8 shuffle)
      const bool isx = (lane & 8) == 0;
      const int bi = lane & 7;
      const float bin = fdiv(isx ? rw : rh, (float)P7);
      const float bin_o = __shfl_xor_sync(0xffffffffu, bin, 8);       // the other axis' bin size
      const int gs = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(bin);
      const int gs_o = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(bin_o);
      const int cnt = gs * gs_o > 1 ? gs * gs_o : 1;
      const float inv_count = 1.0f / (float)cnt;
      const bool b_ok = b >= 0 && b < B && !skip;
      const float start = isx ? x1 : y1;
      const int size = isx ? W : H;
      float* tab = isx ? wx : wy;
      const float base = fadd(start, fmul((float)bi, bin));
      int lo = 1 << 30, hi = -1;
      const bool owner = lane < 16 && bi < P7;
      // gs == 1 (every RoI up to 7 feature pixels wide): the single sample is kept in registers
      int l1 = 0, h1 = 0; float fl1 = 0.f, fh1 = 0.f; bool ok1 = false;
      if (owner && b_ok) {
        if (gs == 1) {
          ok1 = axis_setup(fadd(base, fmul(.5f, bin)), size, l1, h1, fl1, fh1);   // x/1.0f is exact: no division
          if (ok1) { lo = l1; hi = h1; }
        } else {
          for (int i = 0; i < gs; i++) {
            const float v = fadd(base, fdiv(fmul((float)i + .5f, bin), (float)gs));
            int l, h; float fl, fh;
            if (axis_setup(v, size, l, h, fl, fh)) { lo = min(lo, l); hi = max(hi, h); }
          }
        }
      }
      int glo = lo, ghi = hi;
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) {
        glo = min(glo, __shfl_xor_sync(0xffffffffu, glo, o));
        ghi = max(ghi, __shfl_xor_sync(0xffffffffu, ghi, o));
      }
      if (ghi < 0) { glo = 0; ghi = -1; }
      __syncwarp();   // the previous RoI's fragment builds are done reading the tables
      if (owner) {
        const int npad = (ghi - glo + 4) & ~3;   // zero-padded to whole 4-pixel chunks
        for (int c = 0; c < npad; c++) tab[c * 8 + bi] = 0.f;
        if (gs == 1) {
          if (ok1) { tab[(l1 - glo) * 8 + bi] += fh1; tab[(h1 - glo) * 8 + bi] += fl1; }
        } else if (b_ok) {
          for (int i = 0; i < gs; i++) {
            const float v = fadd(base, fdiv(fmul((float)i + .5f, bin), (float)gs));
            int l, h; float fl, fh;
            if (axis_setup(v, size, l, h, fl, fh)) { tab[(l - glo) * 8 + bi] += fh; tab[(h - glo) * 8 + bi] += fl; }
          }
        }
      }
      __syncwarp();
      const int xmin = __shfl_sync(0xffffffffu, glo, 0), xmax = __shfl_sync(0xffffffffu, ghi, 0);
      const int ymin = __shfl_sync(0xffffffffu, glo, 8), ymax = __shfl_sync(0xffffffffu, ghi, 8);
      const bool empty = !b_ok || xmax < xmin || ymax < ymin;
      const int ncx = empty ? 1 : (xmax - xmin) / 4 + 1, ncy = empty ? 1 : (ymax - ymin) / 4 + 1;
      for (int cy = 0; cy < ncy; cy++) {
        for (int cx = 0; cx < ncx; cx++) {
          const int stage = bw_id * DP + slot;
          mbar_wait(empty_bar + stage, phase ^ 1);
          int flags = (cy == ncy - 1 && cx == ncx - 1 ? F_LAST : 0);
          if (empty) {
            flags |= skip ? F_SKIP : F_ZERO;
          } else {
            const int r0 = cy * 4 + (t >> 1), c0 = cx * 4 + 2 * (t & 1);
            const uint32_t dst = s_afrag + stage * AFRAG_BYTES + lane * 16;
#pragma unroll
            for (int mt = 0; mt < 4; mt++) {
              uint32_t a[4];
#pragma unroll
              for (int hl = 0; hl < 2; hl++) {
                const int i = mt * 2 + hl;
                float wy0 = 0.f, wy2 = 0.f, wxa = 0.f, wxb = 0.f;
                if (phs[i] >= 0) {
                  wy0 = wy[r0 * 8 + phs[i]] * inv_count;
                  wy2 = wy[(r0 + 2) * 8 + phs[i]] * inv_count;
                  wxa = wx[c0 * 8 + pws[i]];
                  wxb = wx[(c0 + 1) * 8 + pws[i]];
                }
                a[hl] = F16 ? pack_f16(wy0 * wxa, wy0 * wxb) : pack_bf16(wy0 * wxa, wy0 * wxb);          // k = 2t, 2t+1
                a[2 + hl] = F16 ? pack_f16(wy2 * wxa, wy2 * wxb) : pack_bf16(wy2 * wxa, wy2 * wxb);      // k = 2t+8, +9
              }
              sts128(dst + mt * 512, make_uint4(a[0], a[1], a[2], a[3]));
            }
          }
          if (lane == 0) asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(s_meta + stage * 8), "r"(roi), "r"(flags) : "memory");
          __syncwarp();
          if (lane == 0) {
            if (empty) {
              mbar_arrive(full_bar + stage);
            } else {
              mbar_expect_tx(full_bar + stage, tx_bytes);
              tma_load_5d(s_patch + stage * PATCH_BYTES, &tmap, s_full + stage * 8, 0, xmin + cx * 4, ymin + cy * 4, b, 0);
            }
          }
          if (++slot == DP) { slot = 0; phase ^= 1; }
        }
      }
    }
    return;
  }

  // ------------------------------------------------------------------------------------ MMA warps
  const int cb = warp * 32;                      // first channel of this warp
  const bool active = cb < C;
  const int q = warp >> 1, j0 = (warp & 1) * 4;  // channel quarter, first 16 B chunk inside the quarter's 128 B row
  uint32_t boff[2];
  {
    const int i = lane >> 3, rr = lane & 7;
    const int p = (i & 1) * 8 + rr;
#pragma unroll
    for (int np = 0; np < 2; np++) {
      const int jj = j0 + np * 2 + (i >> 1);
      boff[np] = (uint32_t)(q * QUARTER_BYTES + p * 128 + ((jj ^ rr) << 4));
    }
  }
  // stmatrix row address of this lane inside the warp's staging: matrix (lane>>3) = n-tile = 16 B chunk,
  // row (lane&7) = bin; SWIZZLE_64B: chunk ^= (row >> 1) & 3 (buffers are 512 B aligned)
  const uint32_t sg0 = s_stg + warp * stg_bufs * STG_BYTES;
  const uint32_t stsm_off = (uint32_t)((lane & 7) * STG_ROW_BYTES + (((lane >> 3) ^ ((lane & 7) >> 1)) << 4));
  int sbuf = 0;
  float acc[4][4][4];

  uint32_t slotpk = 0, phasebits = 0;            // per-builder ring position (4 bits each) / parity (bit = builder id)
  int jring = 0;                                 // rotated: position in the shared ring
  for (int it = 0; it < n_iter; it++) {
    const int bw_id = it % NB;
    int roi, flags;
    bool first = true;
    do {
      const int slot = ROT ? 0 : (int)((slotpk >> (4 * bw_id)) & 15u);
      const int stage = ROT ? jring % DP : bw_id * DP + slot;
      if (ROT) mbar_wait_backoff(full_bar + stage, (uint32_t)((jring / DP) & 1), 64);   // builders need the issue slots
      else mbar_wait(full_bar + stage, (phasebits >> bw_id) & 1);
      {
        const uint2 m = lds64(s_meta + stage * 8);
        roi = (int)m.x; flags = (int)m.y;
      }
      if (!(flags & (F_ZERO | F_SKIP)) && active) {
        const uint32_t af = s_afrag + stage * AFRAG_BYTES + lane * 16;
        uint4 a[4];
#pragma unroll
        for (int mt = 0; mt < 4; mt++) a[mt] = lds128(af + mt * 512);
        const uint32_t pb = s_patch + stage * PATCH_BYTES;
        uint32_t b01[4], b23[4];
        ldsm_x4_trans(pb + boff[0], b01);
        ldsm_x4_trans(pb + boff[1], b23);
        if (first) {
#pragma unroll
          for (int mt = 0; mt < 4; mt++) {
            if (F16) {
              mma_f16_z(acc[mt][0], a[mt], b01[0], b01[1]); mma_f16_z(acc[mt][1], a[mt], b01[2], b01[3]);
              mma_f16_z(acc[mt][2], a[mt], b23[0], b23[1]); mma_f16_z(acc[mt][3], a[mt], b23[2], b23[3]);
            } else {
              mma_bf16_z(acc[mt][0], a[mt], b01[0], b01[1]); mma_bf16_z(acc[mt][1], a[mt], b01[2], b01[3]);
              mma_bf16_z(acc[mt][2], a[mt], b23[0], b23[1]); mma_bf16_z(acc[mt][3], a[mt], b23[2], b23[3]);
            }
          }
        } else {
#pragma unroll
          for (int mt = 0; mt < 4; mt++) {
            if (F16) {
              mma_f16(acc[mt][0], a[mt], b01[0], b01[1]); mma_f16(acc[mt][1], a[mt], b01[2], b01[3]);
              mma_f16(acc[mt][2], a[mt], b23[0], b23[1]); mma_f16(acc[mt][3], a[mt], b23[2], b23[3]);
            } else {
              mma_bf16(acc[mt][0], a[mt], b01[0], b01[1]); mma_bf16(acc[mt][1], a[mt], b01[2], b01[3]);
              mma_bf16(acc[mt][2], a[mt], b23[0], b23[1]); mma_bf16(acc[mt][3], a[mt], b23[2], b23[3]);
            }
          }
        }
      } else if (first) {
#pragma unroll
        for (int mt = 0; mt < 4; mt++)
#pragma unroll
          for (int nt = 0; nt < 4; nt++)
#pragma unroll
            for (int e = 0; e < 4; e++) acc[mt][nt][e] = 0.f;
      }
      first = false;
      __syncwarp();
      if (lane == 0) mbar_arrive(empty_bar + stage);
      if (ROT) {
        jring++;
        if (warp == 0 && lane == 0) seq_g[2] = jring;     // after this warp's arrive on the stage's empty barrier
      } else {
        const int nslot = slot + 1 == DP ? 0 : slot + 1;
        if (nslot == 0) phasebits ^= 1u << bw_id;
        slotpk = (slotpk & ~(15u << (4 * bw_id))) | ((uint32_t)nslot << (4 * bw_id));
      }
    } while (!(flags & F_LAST));
    if ((flags & F_SKIP) || !active) continue;

    // epilogue, warp-private (no CTA barrier): bf16 pack -> stmatrix into the swizzled staging -> ONE TMA tensor
    // store of the warp's 49 x 32-channel slice (the LSU pipe never sees the 25 KB/RoI output again)
    const uint32_t sg = sg0 + sbuf * STG_BYTES;
    if (lane == 0) {   // buffer free: the store issued stg_bufs RoIs ago has finished reading it
      if (stg_bufs == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    __syncwarp();
#pragma unroll
    for (int mt = 0; mt < 3; mt++) {
      stsm_x4(sg + mt * 16 * STG_ROW_BYTES + stsm_off,
              pack_bf16(acc[mt][0][0], acc[mt][0][1]), pack_bf16(acc[mt][1][0], acc[mt][1][1]),
              pack_bf16(acc[mt][2][0], acc[mt][2][1]), pack_bf16(acc[mt][3][0], acc[mt][3][1]));
      stsm_x4(sg + (mt * 16 + 8) * STG_ROW_BYTES + stsm_off,
              pack_bf16(acc[mt][0][2], acc[mt][0][3]), pack_bf16(acc[mt][1][2], acc[mt][1][3]),
              pack_bf16(acc[mt][2][2], acc[mt][2][3]), pack_bf16(acc[mt][3][2], acc[mt][3][3]));
    }
    if (g == 0) {   // bin 48: row 0 of m-tile 3 ((48 >> 1) & 3 == 0: unswizzled)
#pragma unroll
      for (int nt = 0; nt < 4; nt++)
        sts32(sg + 48 * STG_ROW_BYTES + nt * 16 + t * 4, pack_bf16(acc[3][nt][0], acc[3][nt][1]));
    }
    fence_proxy_async();      // generic-proxy smem writes -> visible to the async proxy (TMA)
    __syncwarp();
    if (lane == 0) {
      asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(&omap),
                   "r"(cb), "r"(0), "r"(roi), "r"(sg)
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    sbuf ^= stg_bufs - 1;
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // smem must outlive the reads
}

// ------------------------------------------------------------------------------------------------ backward
// mmcv RoIAlign backward (HBB_TOD/mmdet/models/roi_heads/roi_extractors/single_level_roi_extractor.py:56-114 under
// autograd) with the forward's machinery turned around: per 4x4-pixel chunk
//     dP[pix, c] = sum_bin Wmat[bin, pix] * dA[bin, c]                  (16 x 49) * (49 x 256)
// is ONE m16n8k16 row of tiles per warp (M = the chunk's 16 pixels, K = 49 bins padded to 64, N = the warp's 32
// channels) instead of ~3 000 FFMAs per lane in backward.cu's register formulation.  dA of a RoI is fetched once per
// MMA warp by TMA (the forward's output box: 49 bins x 32 channels, SWIZZLE_64B) and held as B fragments for all
// chunks of the RoI; the builder warps produce Wmat in A-fragment order, split into a bf16 head and a bf16 remainder
// (two mma passes), so the weights carry ~16 mantissa bits and the result matches the fp32-weight kernel to the
// rounding of the fp32 accumulation.  The chunk's gradient goes to the NHWC fp32 map with red.global.add.v4.f32.
constexpr int BW_BUILDERS = 2, BW_DEPTH = 2, BW_STAGES = BW_BUILDERS * BW_DEPTH;
constexpr int BW_STG_BYTES = 64 * STG_ROW_BYTES;     // 64 bin rows (49 loaded by TMA, 15 kept zero) x 32 channels
constexpr int BW_META_BYTES = 32;                    // {roi, flags, x0, y0, image, xmax, ymax, -}

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void* desc, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(desc), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}
// bf16 head / remainder of a weight pair
__device__ __forceinline__ void split_bf16(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  const float2 hf = __bfloat1622float2(h);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = pack_bf16(a - hf.x, b - hf.y);
}

__global__ void __launch_bounds__((MMA_WARPS + BW_BUILDERS) * 32, 2)
roi_align_bwd_mma_kernel(const __grid_constant__ CUtensorMap gmap, const float* __restrict__ rois, int K, int B, int C,
                         int H, int W, float scale, int sampling_ratio, int aligned, float* __restrict__ dfeat,
                         const int* __restrict__ roi_level, int level) {
  extern __shared__ uint8_t smem_raw[];
  constexpr int NB = BW_BUILDERS, DP = BW_DEPTH, NSTAGES = BW_STAGES;
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t s_stg = sbase;                                           // [MMA_WARPS][2][BW_STG_BYTES], 512 B aligned
  const uint32_t s_afrag = s_stg + MMA_WARPS * 2 * BW_STG_BYTES;          // [STAGES][hi, lo][AFRAG_BYTES]
  const uint32_t s_tab = s_afrag + NSTAGES * 2 * AFRAG_BYTES;             // [BUILDERS] wx | wy
  const int tab_floats = (W + 4) * 8 + (H + 4) * 8;
  const uint32_t s_full = s_tab + NB * tab_floats * 4;
  const uint32_t s_empty = s_full + NSTAGES * 8;
  const uint32_t s_gbar = s_empty + NSTAGES * 8;                          // [MMA_WARPS][2]: a warp's dA slice has landed
  const uint32_t s_meta = s_gbar + MMA_WARPS * 2 * 8;                     // [STAGES][BW_META_BYTES]
  uint8_t* gen = smem_raw + (sbase - smem_u32(smem_raw));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(gen + (s_full - sbase));
  uint64_t* empty_bar = reinterpret_cast<uint64_t*>(gen + (s_empty - sbase));
  uint64_t* g_bar = reinterpret_cast<uint64_t*>(gen + (s_gbar - sbase));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&gmap);
    for (int i = 0; i < NSTAGES; i++) { mbar_init(full_bar + i, 1); mbar_init(empty_bar + i, MMA_WARPS); }
    for (int i = 0; i < MMA_WARPS * 2; i++) mbar_init(g_bar + i, 1);
    fence_barrier_init();
  }
  // bin rows 49..63 of every staging buffer: read by the last k-step, never written by TMA
  for (int i = threadIdx.x; i < MMA_WARPS * 2 * (64 - NBIN) * (STG_ROW_BYTES / 16); i += blockDim.x) {
    const int bufi = i / ((64 - NBIN) * (STG_ROW_BYTES / 16)), r = i % ((64 - NBIN) * (STG_ROW_BYTES / 16));
    sts128(s_stg + bufi * BW_STG_BYTES + NBIN * STG_ROW_BYTES + r * 16, make_uint4(0u, 0u, 0u, 0u));
  }
  __syncthreads();
  const int n_iter = blockIdx.x < K ? (K - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int g = lane >> 2, t = lane & 3;

  if (warp >= MMA_WARPS) {
    // ---------------------------------------------------------------------------------- builder warps
    const int bw_id = warp - MMA_WARPS;
    float* wx = reinterpret_cast<float*>(gen + (s_tab - sbase)) + bw_id * tab_floats;   // wx[col - xmin][pw]
    float* wy = wx + (W + 4) * 8;                                                        // wy[row - ymin][ph]
    const float off = aligned ? 0.5f : 0.f;
    int slot = 0; uint32_t phase = 0;
    // A fragment of Wmat^T: rows = pixels g (rows 0..1 of the chunk) and g + 8 (rows 2..3), k = bins
    // 16 ks + 2t + {0, 1, 8, 9}
    int phs[16], pws[16];
#pragma unroll
    for (int i = 0; i < 16; i++) {
      const int bq = (i >> 2) * 16 + 2 * t + (i & 1) + ((i >> 1) & 1) * 8;
      phs[i] = bq < NBIN ? bq / P7 : -1;
      pws[i] = bq < NBIN ? bq % P7 : 0;
    }
    float rnext = 0.f;
    if (bw_id < n_iter && lane < 5) rnext = __ldg(rois + (size_t)(blockIdx.x + bw_id * gridDim.x) * 5 + lane);
    for (int it = bw_id; it < n_iter; it += NB) {
      const int roi = blockIdx.x + it * gridDim.x;
      const float rcur = rnext;
      if (it + NB < n_iter && lane < 5)
        rnext = __ldg(rois + (size_t)(blockIdx.x + (it + NB) * gridDim.x) * 5 + lane);
      const bool skip = roi_level != nullptr && roi_level[roi] != level;
      const int b = (int)__shfl_sync(0xffffffffu, rcur, 0);
      const float x1 = fsub(fmul(__shfl_sync(0xffffffffu, rcur, 1), scale), off);
      const float y1 = fsub(fmul(__shfl_sync(0xffffffffu, rcur, 2), scale), off);
      const float x2 = fsub(fmul(__shfl_sync(0xffffffffu, rcur, 3), scale), off);
      const float y2 = fsub(fmul(__shfl_sync(0xffffffffu, rcur, 4), scale), off);
      float rw = fsub(x2, x1), rh = fsub(y2, y1);
      if (!aligned) { rw = fmaxf(rw, 1.f); rh = fmaxf(rh, 1.f); }
      const bool isx = (lane & 8) == 0;
      const int bi = lane & 7;
      const float bin = fdiv(isx ? rw : rh, (float)P7);
      const float bin_o = __shfl_xor_sync(0xffffffffu, bin, 8);
      const int gs = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(bin);
      const int gs_o = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(bin_o);
      const int cnt = gs * gs_o > 1 ? gs * gs_o : 1;
      const float inv_count = 1.0f / (float)cnt;
      const bool b_ok = b >= 0 && b < B && !skip;
      const float start = isx ? x1 : y1;
      const int size = isx ? W : H;
      float* tab = isx ? wx : wy;
      const float base = fadd(start, fmul((float)bi, bin));
      int lo = 1 << 30, hi = -1;
      const bool owner = lane < 16 && bi < P7;
      if (owner && b_ok) {
        for (int i = 0; i < gs; i++) {
          const float v = fadd(base, fdiv(fmul((float)i + .5f, bin), (float)gs));
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
      __syncwarp();   // the previous RoI's fragment builds are done reading the tables
      if (owner) {
        const int npad = (ghi - glo + 4) & ~3;
        for (int c = 0; c < npad; c++) tab[c * 8 + bi] = 0.f;
        if (b_ok) {
          for (int i = 0; i < gs; i++) {
            const float v = fadd(base, fdiv(fmul((float)i + .5f, bin), (float)gs));
            int l, h; float fl, fh;
            if (axis_setup(v, size, l, h, fl, fh)) { tab[(l - glo) * 8 + bi] += fh; tab[(h - glo) * 8 + bi] += fl; }
          }
        }
      }
      __syncwarp();
      const int xmin = __shfl_sync(0xffffffffu, glo, 0), xmax = __shfl_sync(0xffffffffu, ghi, 0);
      const int ymin = __shfl_sync(0xffffffffu, glo, 8), ymax = __shfl_sync(0xffffffffu, ghi, 8);
      const bool empty = !b_ok || xmax < xmin || ymax < ymin;
      const int ncx = empty ? 1 : (xmax - xmin) / 4 + 1, ncy = empty ? 1 : (ymax - ymin) / 4 + 1;
      for (int cy = 0; cy < ncy; cy++) {
        for (int cx = 0; cx < ncx; cx++) {
          const int stage = bw_id * DP + slot;
          mbar_wait(empty_bar + stage, phase ^ 1);
          int flags = (cy == ncy - 1 && cx == ncx - 1 ? F_LAST : 0);
          if (empty) {
            flags |= F_SKIP;
          } else {
            const int rA = cy * 4 + (g >> 2), cc = cx * 4 + (g & 3);
            const uint32_t dst = s_afrag + stage * 2 * AFRAG_BYTES + lane * 16;
#pragma unroll
            for (int ks = 0; ks < 4; ks++) {
              uint32_t ah[4], al[4];
#pragma unroll
              for (int hb = 0; hb < 2; hb++) {          // bins 2t, 2t+1 (hb = 0) and 2t+8, 2t+9 (hb = 1) of this k-step
                float vA[2] = {0.f, 0.f}, vB[2] = {0.f, 0.f};
#pragma unroll
                for (int e = 0; e < 2; e++) {
                  const int i = ks * 4 + hb * 2 + e;
                  if (phs[i] >= 0) {
                    const float wxv = wx[cc * 8 + pws[i]];
                    vA[e] = wy[rA * 8 + phs[i]] * inv_count * wxv;
                    vB[e] = wy[(rA + 2) * 8 + phs[i]] * inv_count * wxv;
                  }
                }
                split_bf16(vA[0], vA[1], ah[hb * 2], al[hb * 2]);              // a0 / a2: pixel row g
                split_bf16(vB[0], vB[1], ah[hb * 2 + 1], al[hb * 2 + 1]);      // a1 / a3: pixel row g + 8
              }
              sts128(dst + ks * 512, make_uint4(ah[0], ah[1], ah[2], ah[3]));
              sts128(dst + AFRAG_BYTES + ks * 512, make_uint4(al[0], al[1], al[2], al[3]));
            }
          }
          if (lane == 0) {
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(s_meta + stage * BW_META_BYTES), "r"(roi), "r"(flags),
                         "r"(xmin + cx * 4), "r"(ymin + cy * 4) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(s_meta + stage * BW_META_BYTES + 16), "r"(b), "r"(xmax),
                         "r"(ymax), "r"(0) : "memory");
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(full_bar + stage);
          if (++slot == DP) { slot = 0; phase ^= 1; }
        }
      }
    }
    return;
  }

  // ------------------------------------------------------------------------------------ MMA warps
  const int cb = warp * 32;
  const bool active = cb < C;
  const uint32_t sg0 = s_stg + warp * 2 * BW_STG_BYTES;
  const uint32_t gb0 = s_gbar + warp * 16;
  // ldmatrix.x4.trans row address inside the warp's [bin][32 ch] staging (SWIZZLE_64B: chunk ^= (row >> 1) & 3):
  // matrix i = lane >> 3: bins (i & 1) * 8 + (lane & 7) of the k-step, channels 8 * (2 np + (i >> 1))
  uint32_t boff[2];
  {
    const int i = lane >> 3, rr = lane & 7;
    const int p = (i & 1) * 8 + rr;                         // + 16 ks: (p >> 1) & 3 is unchanged by multiples of 8
#pragma unroll
    for (int np = 0; np < 2; np++) {
      const int jj = np * 2 + (i >> 1);
      boff[np] = (uint32_t)(p * STG_ROW_BYTES + ((jj ^ ((p >> 1) & 3)) << 4));
    }
  }
  const uint32_t tx_bytes = NBIN * STG_ROW_BYTES;
  if (active && lane == 0 && n_iter > 0) {
    mbar_expect_tx(g_bar + warp * 2, tx_bytes);
    tma_load_3d(sg0, &gmap, gb0, cb, 0, blockIdx.x);
  }
  uint32_t slotpk = 0, phasebits = 0;
  for (int it = 0; it < n_iter; it++) {
    const int bw_id = it % NB;
    const int buf = it & 1;
    uint32_t bf[4][2][4];
    if (active) {
      // the other buffer was last read (ldmatrix, values long consumed) two RoIs ago: refill it for the next RoI
      if (lane == 0 && it + 1 < n_iter) {
        mbar_expect_tx(g_bar + warp * 2 + (buf ^ 1), tx_bytes);
        tma_load_3d(sg0 + (buf ^ 1) * BW_STG_BYTES, &gmap, gb0 + (buf ^ 1) * 8, cb, 0, blockIdx.x + (it + 1) * gridDim.x);
      }
      mbar_wait(g_bar + warp * 2 + buf, (uint32_t)((it >> 1) & 1));
      const uint32_t sg = sg0 + buf * BW_STG_BYTES;
#pragma unroll
      for (int ks = 0; ks < 4; ks++) {
        ldsm_x4_trans(sg + ks * 16 * STG_ROW_BYTES + boff[0], bf[ks][0]);
        ldsm_x4_trans(sg + ks * 16 * STG_ROW_BYTES + boff[1], bf[ks][1]);
      }
    }
    int flags;
    do {
      const int slot = (int)((slotpk >> (4 * bw_id)) & 15u);
      const int stage = bw_id * DP + slot;
      mbar_wait(full_bar + stage, (phasebits >> bw_id) & 1);
      const uint4 m0 = lds128(s_meta + stage * BW_META_BYTES);
      flags = (int)m0.y;
      if (!(flags & F_SKIP) && active) {
        const uint4 m1 = lds128(s_meta + stage * BW_META_BYTES + 16);
        const uint32_t af = s_afrag + stage * 2 * AFRAG_BYTES + lane * 16;
        float acc[4][4];
#pragma unroll
        for (int ks = 0; ks < 4; ks++) {
          const uint4 ah = lds128(af + ks * 512);
          const uint4 al = lds128(af + AFRAG_BYTES + ks * 512);
#pragma unroll
          for (int nt = 0; nt < 4; nt++) {
            const uint32_t b0 = bf[ks][nt >> 1][(nt & 1) * 2], b1 = bf[ks][nt >> 1][(nt & 1) * 2 + 1];
            if (ks == 0) mma_bf16_z(acc[nt], ah, b0, b1);
            else mma_bf16(acc[nt], ah, b0, b1);
            mma_bf16(acc[nt], al, b0, b1);
          }
        }
        // lane (g, t) holds pixel g (acc[.][0..1]) and pixel g + 8 (acc[.][2..3]), channels 8 nt + 2t, +1: the lane pair
        // (t, t ^ 1) trades halves so that the even lane owns 4 consecutive channels of pixel g, the odd lane of g + 8
        const int x = (int)m0.z + (g & 3);
        const int y = (int)m0.w + (g >> 2) + ((t & 1) ? 2 : 0);
        const bool ok = x <= (int)m1.y && y <= (int)m1.z;
        float* dst = dfeat + (((size_t)(int)m1.x * H + y) * W + x) * C + cb + 2 * (t & ~1);
#pragma unroll
        for (int nt = 0; nt < 4; nt++) {
          const float s0 = (t & 1) ? acc[nt][0] : acc[nt][2], s1 = (t & 1) ? acc[nt][1] : acc[nt][3];
          const float r0 = __shfl_xor_sync(0xffffffffu, s0, 1), r1 = __shfl_xor_sync(0xffffffffu, s1, 1);
          if (ok) {
            if (t & 1) red_add_v4(dst + nt * 8, r0, r1, acc[nt][2], acc[nt][3]);
            else red_add_v4(dst + nt * 8, acc[nt][0], acc[nt][1], r0, r1);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty_bar + stage);
      const int nslot = slot + 1 == DP ? 0 : slot + 1;
      if (nslot == 0) phasebits ^= 1u << bw_id;
      slotpk = (slotpk & ~(15u << (4 * bw_id))) | ((uint32_t)nslot << (4 * bw_id));
    } while (!(flags & F_LAST));
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || ptr == nullptr) return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

constexpr size_t SMEM_LIMIT = 113 * 1024;       // two CTAs per SM
constexpr size_t SMEM_LIMIT_ROT = 226 * 1024;   // rotated: one CTA per SM

// rotated pipeline shape: builder warps, stages of the shared ring (PTB200_ROT_CFG=<nb*100+stages> picks another
// instantiated shape for measurements: 813, 810, 614, 610, 415, 410)
static void rot_cfg(int& nb, int& dp) {
  static int v = -1;
  if (v < 0) { const char* e = getenv("PTB200_ROT_CFG"); v = (e != nullptr && e[0] && e[1]) ? atoi(e) : ROT_BUILDERS * 100 + ROT_DEPTH; }
  nb = v / 100; dp = v % 100;
}

size_t smem_bytes(int H, int W, int stg_bufs, bool rot) {
  const size_t tab = rot ? (size_t)ROT_TAB_FLOATS : (size_t)(W + 4) * 8 + (size_t)(H + 4) * 8;
  int rnb = 0, rdp = 0;
  rot_cfg(rnb, rdp);
  const size_t nb = rot ? rnb : BUILDERS, stages = rot ? (size_t)rdp : nb * DEPTH;
  return 1024 + stages * (PATCH_BYTES + AFRAG_BYTES) + MMA_WARPS * stg_bufs * (size_t)STG_BYTES +
         nb * tab * sizeof(float) + 3 * stages * 8 + 64 + 32;
}

// rotated: fixed sampling grids of 1 or 2 samples per axis (the shipped sampling_ratio = 2); the adaptive grid
// (sampling_ratio = 0) stays on roi_align.cu's direct kernel.  Coordinates are packed into 16 bits.
bool supported(int C, int H, int W, bool rot, int sampling_ratio) {
  if (rot && (sampling_ratio < 1 || sampling_ratio * sampling_ratio > ROT_SAMPLES || H > 32000 || W > 32000)) return false;
  return C % 64 == 0 && C >= 64 && C <= 256 && smem_bytes(H, W, 1, rot) <= (rot ? SMEM_LIMIT_ROT : SMEM_LIMIT);
}

int launch(const void* feat_bf16_nhwc, int feat_f16, const float* rois, void* out, long long ld_out, int K, int B, int C,
           int H, int W, float scale, int sampling_ratio, int aligned, const int* roi_level, int level,
           bool rot, int clockwise, cudaStream_t stream) {
  EncodeTiledFn enc = get_encode();
  if (enc == nullptr) { set_error("cuTensorMapEncodeTiled entry point not found"); return PT_ERR_DRIVER; }
  if ((uintptr_t)feat_bf16_nhwc & 15) { set_error("roi_align_mma: feature map must be 16-byte aligned"); return PT_ERR_ARG; }
  CUtensorMap map;
  // NHWC bf16 viewed as (64 ch, W, H, B, C/64): the channel quarter is the slowest box dimension so that each
  // quarter lands as a [16 pixels][128 B] SWIZZLE_128B tile, the ldmatrix-friendly layout
  cuuint64_t dims[5] = {64, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B, (cuuint64_t)(C / 64)};
  cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2, 128};
  cuuint32_t box[5] = {64, 4, 4, 1, (cuuint32_t)(C / 64)};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(&map, feat_f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(feat_bf16_nhwc), dims, strides, box,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("roi_align_mma: cuTensorMapEncodeTiled failed: CUresult %d", (int)r); return PT_ERR_DRIVER; }
  // output (K, 49, C) bf16, row pitch ld_out: one box = one warp's slice of one RoI (32 ch x 49 bins), SWIZZLE_64B
  CUtensorMap omap;
  {
    cuuint64_t od[3] = {(cuuint64_t)C, (cuuint64_t)NBIN, (cuuint64_t)K};
    cuuint64_t os[2] = {(cuuint64_t)C * 2, (cuuint64_t)ld_out * 2};
    cuuint32_t ob[3] = {32, (cuuint32_t)NBIN, 1};
    cuuint32_t oe[3] = {1, 1, 1};
    if ((uintptr_t)out & 15) { set_error("roi_align_mma: output must be 16-byte aligned"); return PT_ERR_ARG; }
    r = enc(&omap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, out, od, os, ob, oe, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("roi_align_mma: output cuTensorMapEncodeTiled failed: CUresult %d", (int)r); return PT_ERR_DRIVER; }
  }
  const int stg_bufs = smem_bytes(H, W, 2, rot) <= (rot ? SMEM_LIMIT_ROT : SMEM_LIMIT) ? 2 : 1;   // large maps: single-buffered staging
  const size_t smem = smem_bytes(H, W, stg_bufs, rot);
  int rnb = 0, rdp = 0;
  rot_cfg(rnb, rdp);
  using KernT = void (*)(const CUtensorMap, const CUtensorMap, const float*, int, int, int, int, int, float, int, int,
                         const int*, int, int, int);
  KernT kern = nullptr;
  int nbld = BUILDERS;
  if (!rot) {
    kern = feat_f16 ? roi_align_mma_kernel<true, false, BUILDERS, DEPTH> : roi_align_mma_kernel<false, false, BUILDERS, DEPTH>;
  } else {
    nbld = rnb;
#define PT_ROT(NBV, DPV) if (rnb == NBV && rdp == DPV) kern = feat_f16 ? roi_align_mma_kernel<true, true, NBV, DPV> : roi_align_mma_kernel<false, true, NBV, DPV>
    PT_ROT(8, 13); PT_ROT(8, 10); PT_ROT(6, 14); PT_ROT(6, 10); PT_ROT(4, 15); PT_ROT(4, 10);
#undef PT_ROT
    if (kern == nullptr) { set_error("roi_align_mma: PTB200_ROT_CFG=%d%d is not an instantiated pipeline shape", rnb, rdp); return PT_ERR_ARG; }
  }
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return PT_ERR_CUDA; }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int per_sm = rot ? 1 : 2;
  const int grid = K < per_sm * sms ? K : per_sm * sms;
  kern<<<grid, (MMA_WARPS + nbld) * 32, smem, stream>>>(map, omap, rois, K, B, C, H, W, scale, sampling_ratio, aligned, roi_level, level,
                                        stg_bufs, clockwise);
  return check_launch("roi_align_mma_kernel");
}

size_t bwd_smem_bytes(int H, int W) {
  return 1024 + MMA_WARPS * 2 * (size_t)BW_STG_BYTES + BW_STAGES * 2 * (size_t)AFRAG_BYTES +
         BW_BUILDERS * ((size_t)(W + 4) * 8 + (size_t)(H + 4) * 8) * sizeof(float) + 2 * BW_STAGES * 8 +
         MMA_WARPS * 2 * 8 + BW_STAGES * BW_META_BYTES + 64;
}

bool bwd_supported(int C, int H, int W, long long ld) {
  return C % 32 == 0 && C >= 32 && C <= 256 && ld % 8 == 0 && bwd_smem_bytes(H, W) <= SMEM_LIMIT;
}

// dfeat (B, H, W, C) fp32, zeroed by the caller, += the RoIAlign gradient of dA (K rows of pitch ld, bf16 bin-major)
int launch_bwd(const void* dA_bf16, long long ld, const float* rois, int K, int B, int C, int H, int W, float scale,
               int sampling_ratio, int aligned, float* dfeat, const int* roi_level, int level, cudaStream_t stream) {
  EncodeTiledFn enc = get_encode();
  if (enc == nullptr) { set_error("cuTensorMapEncodeTiled entry point not found"); return PT_ERR_DRIVER; }
  if ((uintptr_t)dA_bf16 & 15) { set_error("roi_align_bwd_mma: dA must be 16-byte aligned"); return PT_ERR_ARG; }
  CUtensorMap gmap;
  cuuint64_t gd[3] = {(cuuint64_t)C, (cuuint64_t)NBIN, (cuuint64_t)K};
  cuuint64_t gs[2] = {(cuuint64_t)C * 2, (cuuint64_t)ld * 2};
  cuuint32_t gb[3] = {32, (cuuint32_t)NBIN, 1};
  cuuint32_t ge[3] = {1, 1, 1};
  CUresult r = enc(&gmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(dA_bf16), gd, gs, gb, ge,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("roi_align_bwd_mma: cuTensorMapEncodeTiled failed: CUresult %d", (int)r); return PT_ERR_DRIVER; }
  const size_t smem = bwd_smem_bytes(H, W);
  cudaError_t e = cudaFuncSetAttribute(roi_align_bwd_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return PT_ERR_CUDA; }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = K < 2 * sms ? K : 2 * sms;
  roi_align_bwd_mma_kernel<<<grid, (MMA_WARPS + BW_BUILDERS) * 32, smem, stream>>>(
      gmap, rois, K, B, C, H, W, scale, sampling_ratio, aligned, dfeat, roi_level, level);
  return check_launch("roi_align_bwd_mma_kernel");
}

}  // namespace ramma
}  // namespace ptb
