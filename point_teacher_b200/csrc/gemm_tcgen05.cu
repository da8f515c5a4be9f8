// Bag-feature FC as a hand-written tcgen05 / TMA bf16 GEMM for sm_100a.
//
//   C[M,N] = act(A[M,K] * B[N,K]^T + bias[N])        A, B bf16 K-major, fp32 accumulate in TMEM
//
// Replaces the cuBLAS SGEMM behind nn.Linear at the reference call sites
// HBB_TOD/mmdet/models/dense_heads/fcos_head_p2b_ts.py:1205-1207, 1246-1249, 1271-1273
// (shared_fcs_reg / shared_fcs_bag: 12544->1024->1024, + ReLU).
//
// Design (one CTA per SM, persistent, warp specialised, 192 threads):
//   warp 0      TMA producer: 128x64 A tile + 256x64 B tile per stage, SWIZZLE_128B, 4-stage mbarrier ring
//   warp 1      MMA issuer (one elected lane): 4 x tcgen05.mma 128x256x16 per stage, accumulators in TMEM,
//               two 256-column accumulator stages so the epilogue of tile i overlaps the MMAs of tile i+1
//   warps 2..5  epilogue: tcgen05.ld 32x32b -> +bias -> ReLU -> bf16 pack (or fp32) -> 16-byte global stores
// Tail balancing: when tiles % CTAs != 0 the last partial wave is split along K across the idle CTAs; the S
// partials of a tile are parked in private fp32 workspace slots and reduced all-to-all (each CTA finalises 1/S of
// the rows), which is deterministic and free of same-address atomics.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace ptb {
namespace gemm {

constexpr int BM = 128, BN = 256, BK = 64, UMMA_K = 16, ACC_STAGES = 2;
constexpr int A_BYTES = BM * BK * 2;
constexpr int NUM_THREADS = 192;
constexpr int TMEM_COLS = ACC_STAGES * BN;  // 512
// TWO = false: one CTA per tile (128 x 256), B tile 256 rows, 4 stages of 48 KB.
// TWO = true : CTA pair (cluster of 2, tcgen05 cta_group::2) per 256 x 256 tile: every CTA stages its own 128 A rows
//              and HALF of the B tile (128 rows), 6 stages of 32 KB.  The single-CTA mainloop moves 96 KB through
//              shared memory per k-block (48 KB written by TMA + 48 KB read by the MMA) in the ~0.28 us the tensor
//              pipe needs -- more than the 128 B/clk the SM's shared memory delivers; pairing cuts that to 64 KB.
template <bool TWO> struct Cfg {
  static constexpr int STAGES = TWO ? 6 : 4;
  static constexpr int B_ROWS = TWO ? BN / 2 : BN;
  static constexpr int B_BYTES = B_ROWS * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 256 + 1024;
};
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;   // shared::cluster address of the same offset in CTA 0 of the pair

struct Params {
  const float* bias;  // [N] or nullptr
  void* C;            // bf16 or fp32 [M, ldc]
  float* ws;          // split-K workspace: [tail_tiles][S][BM][BN] private fp32 partial slots (no initialisation needed)
  int* counters;      // [tail_tiles][2] (arrived, finished), zero on entry / exit
  int M, N, K, ldc;
  int relu, out_f32;
  const __nv_bfloat16* mask;  // optional [M, ldmask]: out = mask > 0 ? out : 0 (ReLU backward by the forward activation)
  int ldmask;
  int tiles_n, tiles_total, split;  // split = S (>=2) when the tail wave is K-split, else 0
  int full_rounds;
  int a_mn, b_mn;                   // operand stored [K, M] / [K, N] (MN-major): wgrad / dgrad without a transposed copy
  int diag;                         // measurement only (PTB200_GEMM_DIAG): 1 = no TMA loads after the first ring fill
                                    // (MMA + shared-memory operand reads alone), 2 = no MMAs (TMA feed alone)
  int nsplit;                       // > 1: the tiles of the tail round are cut into nsplit column slices (short K)
  int b_box;                        // rows of one K-major B TMA box (BN, or 64 when nsplit is active)
};

struct Unit {
  int m_blk, n_blk, kb0, kb1, tail_idx;  // tail_idx >= 0: split unit
  int n_off, n_w;                        // column slice [n_off, n_off + n_w) of the tile (whole tile: 0, BN)
};

// G / c: number and index of the scheduling units (CTAs, or CTA pairs); rank: CTA inside the pair (0 when unpaired).
// Paired: a "tile" is 256 x 256 and this CTA owns rows [128 * (2 m + rank), +128) of it.
__device__ __forceinline__ bool get_unit(const Params& p, int it, Unit& u, int G, int c, int pair, int rank) {
  const int kb_total = (p.K + BK - 1) / BK;   // a ragged last k-block is zero-filled by TMA (out-of-bounds rows / columns)
  int tile;
  u.n_off = 0; u.n_w = BN;
  if (it < p.full_rounds) {
    tile = it * G + c;
    u.kb0 = 0; u.kb1 = kb_total; u.tail_idx = -1;
  } else if (it == p.full_rounds) {
    const int tail = p.tiles_total - p.full_rounds * G;
    if (tail == 0) return false;
    if (p.split >= 2) {
      const int t = c / p.split, s = c % p.split;
      if (t >= tail) return false;
      tile = p.full_rounds * G + t;
      u.kb0 = (int)(((long long)s * kb_total) / p.split);
      u.kb1 = (int)(((long long)(s + 1) * kb_total) / p.split);
      u.tail_idx = t;
    } else if (p.nsplit > 1) {
      // short contraction: no K-split (the reduction would cost more than it saves); the tail tiles are cut into
      // column slices instead -- no reduction at all, every CTA writes its own output columns
      const int t = c / p.nsplit, q = c % p.nsplit;
      if (t >= tail) return false;
      tile = p.full_rounds * G + t;
      u.n_w = BN / p.nsplit; u.n_off = q * u.n_w;
      u.kb0 = 0; u.kb1 = kb_total; u.tail_idx = -1;
    } else {
      if (c >= tail) return false;
      tile = p.full_rounds * G + c;
      u.kb0 = 0; u.kb1 = kb_total; u.tail_idx = -1;
    }
  } else {
    return false;
  }
  u.m_blk = tile / p.tiles_n;  // n fastest: the N-tiles of one M-tile run concurrently -> A read from HBM once
  u.n_blk = tile % p.tiles_n;
  if (pair) {
    u.m_blk = u.m_blk * 2 + rank;
    if (u.tail_idx >= 0) u.tail_idx = u.tail_idx * 2 + rank;
  }
  return true;
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const void* desc, uint64_t* bar, int c0, int c1) {
  // executed by both CTAs of the pair; the transaction bytes update CTA 0's barrier
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(desc), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_mma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar) {   // arrives on the same barrier offset in both CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_cta0(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & PEER_BIT_MASK) : "memory");
}

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  // K-major, SWIZZLE_128B canonical layout: LBO = 1 (16 B units, ignored), SBO = 8 rows * 128 B = 1024 B,
  // version = 1 (Blackwell), layout_type = 2 (SWIZZLE_128B).
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr) {
  // MN-major, SWIZZLE_128B canonical layout ((64 mn, n_atoms), (8 k, k_groups)): one TMA box = 64 k-rows of 128 B
  // (64 MN elements); SBO = 8 k-rows * 128 B = 1024 B between k-groups, LBO = 64 k-rows * 128 B = 8192 B between
  // 64-element MN atoms (consecutive TMA boxes).
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(8192 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

constexpr uint32_t IDESC = (1u << 4)      // D format: fp32
                           | (1u << 7)    // A format: bf16
                           | (1u << 10)   // B format: bf16
                           | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);  // K-major A and B
constexpr uint32_t IDESC_PAIR = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) |
                                ((uint32_t)((2 * BM) >> 4) << 24);                        // M = 256 across the pair

__device__ __forceinline__ void store_chunk(const Params& p, int row, int col0, const uint32_t* r, bool row_ok) {
  // r: 32 fp32 accumulators of columns col0..col0+31 of this thread's row
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; j++) v[j] = __uint_as_float(r[j]);
  if (p.bias != nullptr) {
    const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
    for (int j = 0; j < 8; j++) {
      float4 b = __ldg(b4 + j);
      v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
    }
  }
  if (p.relu) {
#pragma unroll
    for (int j = 0; j < 32; j++) v[j] = fmaxf(v[j], 0.f);
  }
  if (!row_ok) return;
  if (p.mask != nullptr) {
    const uint4* m4 = reinterpret_cast<const uint4*>(p.mask + (size_t)row * p.ldmask + col0);
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const uint4 m = __ldg(m4 + j);
      const uint32_t w[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
      for (int q = 0; q < 4; q++) {
        // bf16 > 0  <=>  sign bit clear and magnitude non-zero
        if (!((w[q] & 0x7fffu) != 0 && (w[q] & 0x8000u) == 0)) v[8 * j + 2 * q] = 0.f;
        if (!((w[q] & 0x7fff0000u) != 0 && (w[q] & 0x80000000u) == 0)) v[8 * j + 2 * q + 1] = 0.f;
      }
    }
  }
  if (p.out_f32) {
    float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.C) + (size_t)row * p.ldc + col0);
#pragma unroll
    for (int j = 0; j < 8; j++) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  } else {
    uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.C) + (size_t)row * p.ldc + col0);
#pragma unroll
    for (int j = 0; j < 4; j++)
      dst[j] = make_uint4(pack_bf16(v[8 * j], v[8 * j + 1]), pack_bf16(v[8 * j + 2], v[8 * j + 3]),
                          pack_bf16(v[8 * j + 4], v[8 * j + 5]), pack_bf16(v[8 * j + 6], v[8 * j + 7]));
  }
}

// AMN / BMN: operand staged as an MN-major tile (compile-time, so the K-major forward kernel is unchanged)
template <bool TWO, bool AMN = false, bool BMN = false>
__global__ void __launch_bounds__(NUM_THREADS, 1)
fc_gemm_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
               const Params p) {
  constexpr int STAGES = Cfg<TWO>::STAGES, STAGE_BYTES = Cfg<TWO>::STAGE_BYTES, B_ROWS = Cfg<TWO>::B_ROWS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int rank = TWO ? (int)cluster_ctarank() : 0;
  const int sched_G = TWO ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int sched_c = TWO ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + ACC_STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty_bar + ACC_STAGES);
  int* fin_flag = reinterpret_cast<int*>(tmem_ptr + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    for (int i = 0; i < STAGES; i++) { mbar_init(full_bar + i, 1); mbar_init(empty_bar + i, 1); }
    for (int i = 0; i < ACC_STAGES; i++) { mbar_init(tfull_bar + i, 1); mbar_init(tempty_bar + i, TWO ? 8 : 4); }
    fence_barrier_init();
  }
  if (TWO) cluster_sync_all();          // the peer's barriers exist before anything can signal them
  if (warp == 1) {
    if (TWO) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)),
                   "r"((uint32_t)TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      tmem_alloc(tmem_ptr, TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (TWO) cluster_sync_all();          // both CTAs own their accumulator columns before the first paired MMA
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      Unit u;
      for (int it = 0; get_unit(p, it, u, sched_G, sched_c, TWO, rank); it++) {
        for (int kb = u.kb0; kb < u.kb1; kb++) {
          mbar_wait(empty_bar + stage, phase ^ 1);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          if (p.diag == 1 && (it > 0 || kb - u.kb0 >= STAGES)) {      // diagnostic: stale operands, timing only
            mbar_arrive(full_bar + stage);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
            continue;
          }
          if (TWO) {
            // CTA 0's barrier collects the bytes of both CTAs' loads (its own arrive arms 2 x STAGE_BYTES)
            if (rank == 0) mbar_expect_tx(full_bar + stage, 2 * STAGE_BYTES);
            tma_load_2d_pair(sa, &tma_a, full_bar + stage, kb * BK, u.m_blk * BM);
            tma_load_2d_pair(sa + A_BYTES, &tma_b, full_bar + stage, kb * BK, u.n_blk * BN + rank * B_ROWS);
          } else {
            mbar_expect_tx(full_bar + stage, (uint32_t)(A_BYTES + u.n_w * BK * 2));
            if (AMN) {
#pragma unroll
              for (int j = 0; j < BM / 64; j++)
                tma_load_2d(sa + j * (64 * BK * 2), &tma_a, full_bar + stage, u.m_blk * BM + j * 64, kb * BK);
            } else {
              tma_load_2d(sa, &tma_a, full_bar + stage, kb * BK, u.m_blk * BM);
            }
            const int ncol0 = u.n_blk * BN + u.n_off;
            if (BMN) {
              for (int j = 0; j < u.n_w / 64; j++)
                tma_load_2d(sa + A_BYTES + j * (64 * BK * 2), &tma_b, full_bar + stage, ncol0 + j * 64, kb * BK);
            } else {
              for (int j = 0; j < u.n_w / p.b_box; j++)      // one box for a whole tile unless the launch slices its tail
                tma_load_2d(sa + A_BYTES + j * (p.b_box * BK * 2), &tma_b, full_bar + stage, kb * BK, ncol0 + j * p.b_box);
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (paired: CTA 0 only)
    if (!TWO || rank == 0) {
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      Unit u;
      for (int it = 0; get_unit(p, it, u, sched_G, sched_c, TWO, rank); it++) {
        mbar_wait(tempty_bar + acc, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = u.kb0; kb < u.kb1; kb++) {
          mbar_wait(full_bar + stage, phase);
          tc_fence_after();
          if (lane == 0) {
            const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
            constexpr bool amn = !TWO && AMN, bmn = !TWO && BMN;
            const uint64_t adesc = amn ? make_desc_mn(sa) : make_desc(sa);
            const uint64_t bdesc = bmn ? make_desc_mn(sa + A_BYTES) : make_desc(sa + A_BYTES);
            // K-major: advance 16 bf16 = 32 B along K inside the 128 B swizzle row: +2 in 16 B units;
            // MN-major: 16 k-rows of 128 B = 2048 B: +128
            constexpr uint32_t astep = amn ? 128u : 2u, bstep = bmn ? 128u : 2u;
            // N of the instruction follows the unit's column slice (256, or 128 / 64 for a sliced tail tile)
            const uint32_t idesc = ((IDESC & ~(0x3Fu << 17)) | ((uint32_t)(u.n_w >> 3) << 17)) |
                                   (amn ? (1u << 15) : 0u) | (bmn ? (1u << 16) : 0u);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; k++) {
              if (p.diag == 2 && !(kb == u.kb0 && k == 0)) continue;   // diagnostic: one MMA per tile, timing only
              if (TWO) tc_mma_bf16_pair(tmem_d, adesc + 2 * k, bdesc + 2 * k, IDESC_PAIR, (kb > u.kb0 || k > 0) ? 1u : 0u);
              else tc_mma_bf16(tmem_d, adesc + astep * k, bdesc + bstep * k, idesc, (kb > u.kb0 || k > 0) ? 1u : 0u);
            }
            // frees the smem slot (in both CTAs when paired) when these MMAs retire
            if (TWO) tc_commit_pair(empty_bar + stage); else tc_commit(empty_bar + stage);
            if (kb == u.kb1 - 1) { if (TWO) tc_commit_pair(tfull_bar + acc); else tc_commit(tfull_bar + acc); }
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5)
    const int q = warp & 3;  // TMEM lane quarter this warp may touch
    int acc = 0; uint32_t acc_phase = 0;
    Unit u;
    for (int it = 0; get_unit(p, it, u, sched_G, sched_c, TWO, rank); it++) {
      mbar_wait(tfull_bar + acc, acc_phase);
      tc_fence_after();
      const int row = u.m_blk * BM + q * 32 + lane;
      const bool row_ok = row < p.M;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
      if (u.tail_idx < 0) {
#pragma unroll 1
        for (int ch = 0; ch < u.n_w / 32; ch++) {
          uint32_t r[32];
          tmem_ld_32x32(taddr + ch * 32, r);
          tmem_ld_wait();
          store_chunk(p, row, u.n_blk * BN + u.n_off + ch * 32, r, row_ok);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (TWO) mbar_arrive_cta0(tempty_bar + acc); else mbar_arrive(tempty_bar + acc); }
      } else {
        // split-K partial (tail wave): the S CTAs of a tile each park their fp32 partial in a private workspace
        // slot (plain 16-byte stores), meet at a counter, and then each CTA reduces and finalises 1/S of the tile's
        // rows from all S slots -- a deterministic all-to-all reduction, 2 x 128 KB of L2 traffic per CTA.  (The
        // earlier red.global.add version serialised S-way on every address and cost ~25 us per FC1 launch.)
        const int s_idx = sched_c % p.split;
        float* slot = p.ws + ((size_t)u.tail_idx * p.split + s_idx) * (BM * BN);
        float* wrow = slot + (size_t)(q * 32 + lane) * BN;
#pragma unroll 1
        for (int ch = 0; ch < BN / 32; ch++) {
          uint32_t r[32];
          tmem_ld_32x32(taddr + ch * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; j++)
            __stcg(reinterpret_cast<float4*>(wrow + ch * 32) + j,
                   make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                               __uint_as_float(r[4 * j + 3])));
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (TWO) mbar_arrive_cta0(tempty_bar + acc); else mbar_arrive(tempty_bar + acc); }
        __threadfence();
        asm volatile("bar.sync 1, 128;" ::: "memory");  // the 4 epilogue warps
        int* cnt = p.counters + 2 * u.tail_idx;
        if (threadIdx.x == 64) {
          atomicAdd(cnt, 1);
          // all S CTAs of the tile are co-resident (grid <= #SMs, one CTA per SM): bounded spin
          long long t0 = clock64();
          while (*reinterpret_cast<volatile int*>(cnt) < p.split) {
            if (clock64() - t0 > 4000000000LL) __trap();
          }
          __threadfence();
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        {
          const int r_lo = (int)(((long long)s_idx * BM) / p.split), r_hi = (int)(((long long)(s_idx + 1) * BM) / p.split);
          const float* tile_ws = p.ws + (size_t)u.tail_idx * p.split * (BM * BN);
          const int items = (r_hi - r_lo) * (BN / 4);
          for (int idx = threadIdx.x - 64; idx < items; idx += 128) {
            const int rr = r_lo + idx / (BN / 4), c4 = idx % (BN / 4);
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int sp = 0; sp < p.split; sp++) {
              const float4 t = __ldcg(reinterpret_cast<const float4*>(tile_ws + ((size_t)sp * BM + rr) * BN) + c4);
              a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
            }
            const int grow = u.m_blk * BM + rr, gcol = u.n_blk * BN + c4 * 4;
            if (grow < p.M) {
              float v[4] = {a.x, a.y, a.z, a.w};
              if (p.bias != nullptr) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + gcol));
                v[0] += b.x; v[1] += b.y; v[2] += b.z; v[3] += b.w;
              }
              if (p.relu) {
#pragma unroll
                for (int j = 0; j < 4; j++) v[j] = fmaxf(v[j], 0.f);
              }
              if (p.mask != nullptr) {
                const uint2 m = __ldg(reinterpret_cast<const uint2*>(p.mask + (size_t)grow * p.ldmask + gcol));
                if (!((m.x & 0x7fffu) != 0 && (m.x & 0x8000u) == 0)) v[0] = 0.f;
                if (!((m.x & 0x7fff0000u) != 0 && (m.x & 0x80000000u) == 0)) v[1] = 0.f;
                if (!((m.y & 0x7fffu) != 0 && (m.y & 0x8000u) == 0)) v[2] = 0.f;
                if (!((m.y & 0x7fff0000u) != 0 && (m.y & 0x80000000u) == 0)) v[3] = 0.f;
              }
              if (p.out_f32)
                *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.C) + (size_t)grow * p.ldc + gcol) =
                    make_float4(v[0], v[1], v[2], v[3]);
              else
                *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.C) + (size_t)grow * p.ldc + gcol) =
                    make_uint2(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]));
            }
          }
        }
        // the last CTA to finish its slice re-arms the counters for the next launch
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (threadIdx.x == 64) {
          const int done = atomicAdd(cnt + 1, 1);
          if (done == p.split - 1) { cnt[0] = 0; cnt[1] = 0; __threadfence(); }
        }
      }
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (TWO) cluster_sync_all();          // the peer may still be reading accumulators the pair allocated together
  if (warp == 1) {
    tc_fence_after();
    if (TWO) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
    else tmem_dealloc(tmem_base, TMEM_COLS);
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

// PTB200_GEMM_PAIR=1 selects the paired (cta_group::2) kernel for the long contractions.  Measured on B200
// (tools/bench_gemm_ab.py): wave-exact 18944x1024x12544 1490 vs 1467 TFLOP/s, 5000 rows 105 vs 111 us alone, but no
// gain inside the captured step (0.495 vs 0.494 ms) -- the single-CTA mainloop already runs at ~90 % of the measured
// cuBLAS peak and the FC1 launches are bound by wave quantisation, so the simpler kernel stays the default.
static bool pair_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("PTB200_GEMM_PAIR"); v = (e != nullptr && e[0] == '1') ? 1 : 0; }
  return v == 1;
}

static int make_map(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (enc == nullptr) { set_error("cuTensorMapEncodeTiled entry point not found"); return PT_ERR_DRIVER; }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed: CUresult %d", (int)r); return PT_ERR_DRIVER; }
  return PT_OK;
}

// MN-major operand stored [kdim rows, mndim columns] (ld elements per row): boxes of 64 columns (128 B) x BK rows.
static int make_map_mn(CUtensorMap* map, const void* base, long long kdim, long long mndim, long long ld) {
  EncodeTiledFn enc = get_encode();
  if (enc == nullptr) { set_error("cuTensorMapEncodeTiled entry point not found"); return PT_ERR_DRIVER; }
  cuuint64_t dims[2] = {(cuuint64_t)mndim, (cuuint64_t)kdim};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)BK};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (MN-major) failed: CUresult %d", (int)r); return PT_ERR_DRIVER; }
  return PT_OK;
}

}  // namespace gemm
}  // namespace ptb

using namespace ptb;

extern "C" long long pt_fc_gemm_workspace_bytes(int num_sms) {
  // tail tiles x splits <= num_sms private fp32 BMxBN slots, + two int counters per tail tile (zero on entry / exit)
  return (long long)num_sms * ((long long)gemm::BM * gemm::BN * 4) + 4096;
}

extern "C" int pt_fc_gemm_bf16_ex(const void* A, long long lda, const void* B, long long ldb, const float* bias,
                                  void* C, long long ldc, int M, int N, int K, int relu, int out_f32, const void* mask,
                                  long long ldmask, void* workspace, long long workspace_bytes, int num_sms,
                                  int allow_split, void* stream);
extern "C" int pt_fc_gemm_bf16_mn(const void* A, long long lda, int a_mn, const void* B, long long ldb, int b_mn,
                                  const float* bias, void* C, long long ldc, int M, int N, int K, int relu, int out_f32,
                                  const void* mask, long long ldmask, void* workspace, long long workspace_bytes,
                                  int num_sms, int allow_split, void* stream);

extern "C" int pt_fc_gemm_bf16(const void* A, long long lda, const void* B, long long ldb, const float* bias,
                               void* C, long long ldc, int M, int N, int K, int relu, int out_f32, void* workspace,
                               long long workspace_bytes, int num_sms, int allow_split, void* stream) {
  return pt_fc_gemm_bf16_ex(A, lda, B, ldb, bias, C, ldc, M, N, K, relu, out_f32, nullptr, 0, workspace, workspace_bytes,
                            num_sms, allow_split, stream);
}

extern "C" int pt_fc_gemm_bf16_ex(const void* A, long long lda, const void* B, long long ldb, const float* bias,
                                  void* C, long long ldc, int M, int N, int K, int relu, int out_f32, const void* mask,
                                  long long ldmask, void* workspace, long long workspace_bytes, int num_sms,
                                  int allow_split, void* stream) {
  return pt_fc_gemm_bf16_mn(A, lda, 0, B, ldb, 0, bias, C, ldc, M, N, K, relu, out_f32, mask, ldmask, workspace,
                            workspace_bytes, num_sms, allow_split, stream);
}

// a_mn / b_mn = 1: the operand is stored with the contraction index as its ROW index (A as [K, M], B as [K, N], ld =
// elements per row) -- the layouts wgrad (dW = dY^T X) and dgrad (dX = dY W) meet -- and is fed to the tensor core as
// an MN-major shared-memory tile, so no transposed copy is ever materialised.  K may then be ragged (TMA zero-fills).
extern "C" int pt_fc_gemm_bf16_mn(const void* A, long long lda, int a_mn, const void* B, long long ldb, int b_mn,
                                  const float* bias, void* C, long long ldc, int M, int N, int K, int relu, int out_f32,
                                  const void* mask, long long ldmask, void* workspace, long long workspace_bytes,
                                  int num_sms, int allow_split, void* stream) {
  using namespace gemm;
  if (mask != nullptr && (((uintptr_t)mask & 15) || (ldmask % 8))) {
    set_error("pt_fc_gemm_bf16_ex: mask must be 16-byte aligned with ldmask a multiple of 8");
    return PT_ERR_ARG;
  }
  if (M <= 0) return PT_OK;
  if (N % BN != 0 || K <= 0) {   // a ragged K is fine: the tensor maps carry the true extents and TMA zero-fills
    set_error("pt_fc_gemm_bf16: N must be a multiple of %d and K positive (got N=%d K=%d)", BN, N, K);
    return PT_ERR_ARG;
  }
  if (((uintptr_t)A & 15) || ((uintptr_t)B & 15) || ((uintptr_t)C & 15) || (lda % 8) || (ldb % 8) || (ldc % 8)) {
    set_error("pt_fc_gemm_bf16: operands must be 16-byte aligned with leading dimensions multiple of 8");
    return PT_ERR_ARG;
  }
  if (num_sms <= 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  // paired (cta_group::2) kernel for the long contractions; single-CTA kernel otherwise
  const bool two = (K >= 4096) && (num_sms % 2 == 0) && (M > 2 * BM) && pair_enabled() && !a_mn && !b_mn;
  CUtensorMap ma, mb;
  int rc = a_mn ? make_map_mn(&ma, A, K, M, lda) : make_map(&ma, A, M, K, lda, BM);
  if (rc != PT_OK) return rc;

  Params p;
  p.mask = reinterpret_cast<const __nv_bfloat16*>(mask); p.ldmask = (int)ldmask;
  p.a_mn = a_mn ? 1 : 0; p.b_mn = b_mn ? 1 : 0;
  {
    static int diag = -1;
    if (diag < 0) { const char* e = getenv("PTB200_GEMM_DIAG"); diag = e != nullptr ? atoi(e) : 0; }
    p.diag = diag;
  }
  p.bias = bias; p.C = C; p.M = M; p.N = N; p.K = K; p.ldc = (int)ldc; p.relu = relu; p.out_f32 = out_f32;
  p.tiles_n = N / BN;
  const int tile_rows = two ? 2 * BM : BM;                 // rows of one scheduling tile
  const int tiles_m = (M + tile_rows - 1) / tile_rows;
  p.tiles_total = tiles_m * p.tiles_n;
  const int units = two ? num_sms / 2 : num_sms;           // scheduling units: CTA pairs or CTAs
  int grid_u = p.tiles_total < units ? p.tiles_total : units;
  const int kb_total = (K + BK - 1) / BK;
  const int sub = two ? 2 : 1;                             // 128-row output sub-tiles per scheduling tile
  p.split = 0;
  p.ws = nullptr; p.counters = nullptr;
  int tail;
  // Few tiles, long contraction (wgrad of the 1024 x 1024 layer: 32 tiles, K = 5000 RoIs): EVERY tile is split along K
  // over units / tiles CTAs -- the whole launch is then one "tail wave" of the scheme below (full_rounds = 0).
  int S_all = (allow_split && workspace != nullptr && p.tiles_total * 2 <= units && kb_total >= 64) ? units / p.tiles_total : 0;
  if (S_all > kb_total / 8) S_all = kb_total / 8;
  if (S_all >= 2 && p.tiles_total * sub * 8 <= 4096 &&
      (long long)p.tiles_total * sub * S_all * BM * BN * 4 + 4096 <= workspace_bytes) {
    grid_u = S_all * p.tiles_total;
    p.full_rounds = 0;
    tail = p.tiles_total;
    p.split = S_all;
    p.counters = reinterpret_cast<int*>(workspace);
    p.ws = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + 4096);
  } else {
    p.full_rounds = p.tiles_total / grid_u;
    tail = p.tiles_total - p.full_rounds * grid_u;
  }
  if (p.split == 0 && allow_split && tail > 0 && tail * sub * 8 <= 4096 && workspace != nullptr) {
    int S = grid_u / tail;
    // every split must keep >= 8 k-blocks of MMA work, otherwise the fp32 reduction costs more than it saves; short
    // contractions (K < 4096: FC2, measured 43 us split vs 27 us unsplit at 5000x1024x1024) are never split
    if (S > kb_total / 8) S = kb_total / 8;
    if (kb_total < 64) S = 0;
    const long long need = (long long)tail * sub * S * BM * BN * 4 + 4096;
    if (S >= 2 && need <= workspace_bytes) {
      p.split = S;
      // counters first (they must stay zero between launches; the partial slots need no initialisation)
      p.counters = reinterpret_cast<int*>(workspace);
      p.ws = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + 4096);   // counters live in the first 4 KB
    }
  }
  // Short contractions (FC2, the dgrads: K = 1024) are never K-split; their tail round (160 tiles on 148 CTAs: 12 tiles
  // in the second round) is cut into 2 or 4 column slices per tile instead, one CTA each.  PTB200_GEMM_NSPLIT=0: off.
  p.nsplit = 0; p.b_box = BN;
  {
    static int ns_on = -1;
    if (ns_on < 0) { const char* e = getenv("PTB200_GEMM_NSPLIT"); ns_on = (e != nullptr && e[0] == '0') ? 0 : 1; }
    if (ns_on && !two && p.split == 0 && tail > 0 && p.full_rounds > 0 && kb_total < 64) {
      const int ns = grid_u / tail;
      p.nsplit = ns >= 4 ? 4 : (ns >= 2 ? 2 : 0);
      if (p.nsplit) p.b_box = 64;
    }
  }
  rc = b_mn ? make_map_mn(&mb, B, K, N, ldb) : make_map(&mb, B, N, K, ldb, two ? BN / 2 : p.b_box);
  if (rc != PT_OK) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(fc_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<false>::SMEM_BYTES);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(fc_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<true>::SMEM_BYTES);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(fc_gemm_kernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<false>::SMEM_BYTES);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(fc_gemm_kernel<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<false>::SMEM_BYTES);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(fc_gemm_kernel<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<false>::SMEM_BYTES);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return PT_ERR_CUDA; }
    attr_set = true;
  }
  if (two) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * grid_u); cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = Cfg<true>::SMEM_BYTES; cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, fc_gemm_kernel<true>, ma, mb, p);
    if (e != cudaSuccess) { set_error("fc_gemm_kernel<pair>: launch failed: %s", cudaGetErrorString(e)); return PT_ERR_CUDA; }
    return check_launch("fc_gemm_kernel<pair>");
  }
  const size_t sm = Cfg<false>::SMEM_BYTES;
  cudaStream_t st = (cudaStream_t)stream;
  // The split-K tail meets at a counter, so every CTA of the grid must be resident at the same time.  That is a
  // CHECKED launch precondition, not an assumption: (1) the occupancy query must allow grid_u CTAs on num_sms SMs,
  // else the launch falls back to the unsplit schedule; (2) split launches carry the cooperative attribute, with
  // which the driver refuses a grid that cannot be co-resident (cudaErrorCooperativeLaunchTooLarge) and the
  // hardware schedules the grid all-or-nothing even when other kernels (weight casts on the side stream, NCCL's
  // all-reduce under the backward) hold part of the GPU.
  static int occ_per_sm = -1;
  if (occ_per_sm < 0) {
    int o = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, fc_gemm_kernel<false>, NUM_THREADS, sm);
    if (e != cudaSuccess) { set_error("cudaOccupancyMaxActiveBlocksPerMultiprocessor: %s", cudaGetErrorString(e)); return PT_ERR_CUDA; }
    occ_per_sm = o;
  }
  if (p.split > 0 && (long long)grid_u > (long long)occ_per_sm * num_sms) { p.split = 0; p.ws = nullptr; p.counters = nullptr; }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid_u); cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = sm; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative;
  at[0].val.cooperative = 1;
  static int no_coop = -1;       // PTB200_GEMM_NO_COOP=1: measurement switch only (drops the co-scheduling guarantee)
  if (no_coop < 0) { const char* e = getenv("PTB200_GEMM_NO_COOP"); no_coop = (e != nullptr && e[0] == '1') ? 1 : 0; }
  cfg.attrs = at; cfg.numAttrs = (p.split > 0 && !no_coop) ? 1 : 0;
  cudaError_t le;
  if (a_mn && b_mn) le = cudaLaunchKernelEx(&cfg, fc_gemm_kernel<false, true, true>, ma, mb, p);
  else if (a_mn) le = cudaLaunchKernelEx(&cfg, fc_gemm_kernel<false, true, false>, ma, mb, p);
  else if (b_mn) le = cudaLaunchKernelEx(&cfg, fc_gemm_kernel<false, false, true>, ma, mb, p);
  else le = cudaLaunchKernelEx(&cfg, fc_gemm_kernel<false>, ma, mb, p);
  if (le != cudaSuccess) {
    set_error("fc_gemm_kernel: launch failed (split %d, grid %d): %s", p.split, grid_u, cudaGetErrorString(le));
    return PT_ERR_CUDA;
  }
  return check_launch("fc_gemm_kernel");
}
