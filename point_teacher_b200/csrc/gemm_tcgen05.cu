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

#include "common.cuh"

namespace ptb {
namespace gemm {

constexpr int BM = 128, BN = 256, BK = 64, UMMA_K = 16, STAGES = 4, ACC_STAGES = 2;
constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int NUM_THREADS = 192;
constexpr int TMEM_COLS = ACC_STAGES * BN;  // 512
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 256 + 1024;

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
};

struct Unit {
  int m_blk, n_blk, kb0, kb1, tail_idx;  // tail_idx >= 0: split unit
};

__device__ __forceinline__ bool get_unit(const Params& p, int it, Unit& u) {
  const int G = gridDim.x, c = blockIdx.x;
  const int kb_total = p.K / BK;
  int tile;
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
  return true;
}

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  // K-major, SWIZZLE_128B canonical layout: LBO = 1 (16 B units, ignored), SBO = 8 rows * 128 B = 1024 B,
  // version = 1 (Blackwell), layout_type = 2 (SWIZZLE_128B).
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

constexpr uint32_t IDESC = (1u << 4)      // D format: fp32
                           | (1u << 7)    // A format: bf16
                           | (1u << 10)   // B format: bf16
                           | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);  // K-major A and B

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

__global__ void __launch_bounds__(NUM_THREADS, 1)
fc_gemm_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
               const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
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
    for (int i = 0; i < ACC_STAGES; i++) { mbar_init(tfull_bar + i, 1); mbar_init(tempty_bar + i, 4); }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      Unit u;
      for (int it = 0; get_unit(p, it, u); it++) {
        for (int kb = u.kb0; kb < u.kb1; kb++) {
          mbar_wait(empty_bar + stage, phase ^ 1);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          mbar_expect_tx(full_bar + stage, STAGE_BYTES);
          tma_load_2d(sa, &tma_a, full_bar + stage, kb * BK, u.m_blk * BM);
          tma_load_2d(sa + A_BYTES, &tma_b, full_bar + stage, kb * BK, u.n_blk * BN);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    int stage = 0; uint32_t phase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    Unit u;
    for (int it = 0; get_unit(p, it, u); it++) {
      mbar_wait(tempty_bar + acc, acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BN;
      for (int kb = u.kb0; kb < u.kb1; kb++) {
        mbar_wait(full_bar + stage, phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const uint64_t adesc = make_desc(sa), bdesc = make_desc(sa + A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; k++) {
            // advance 16 bf16 = 32 B along K inside the 128 B swizzle row: +2 in 16 B units
            tc_mma_bf16(tmem_d, adesc + 2 * k, bdesc + 2 * k, IDESC, (kb > u.kb0 || k > 0) ? 1u : 0u);
          }
          tc_commit(empty_bar + stage);  // frees the smem slot when these MMAs retire
          if (kb == u.kb1 - 1) tc_commit(tfull_bar + acc);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5)
    const int q = warp & 3;  // TMEM lane quarter this warp may touch
    int acc = 0; uint32_t acc_phase = 0;
    Unit u;
    for (int it = 0; get_unit(p, it, u); it++) {
      mbar_wait(tfull_bar + acc, acc_phase);
      tc_fence_after();
      const int row = u.m_blk * BM + q * 32 + lane;
      const bool row_ok = row < p.M;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
      if (u.tail_idx < 0) {
#pragma unroll 1
        for (int ch = 0; ch < BN / 32; ch++) {
          uint32_t r[32];
          tmem_ld_32x32(taddr + ch * 32, r);
          tmem_ld_wait();
          store_chunk(p, row, u.n_blk * BN + ch * 32, r, row_ok);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar + acc);
      } else {
        // split-K partial (tail wave): the S CTAs of a tile each park their fp32 partial in a private workspace
        // slot (plain 16-byte stores), meet at a counter, and then each CTA reduces and finalises 1/S of the tile's
        // rows from all S slots -- a deterministic all-to-all reduction, 2 x 128 KB of L2 traffic per CTA.  (The
        // earlier red.global.add version serialised S-way on every address and cost ~25 us per FC1 launch.)
        const int s_idx = blockIdx.x % p.split;
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
        if (lane == 0) mbar_arrive(tempty_bar + acc);
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
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
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
  using namespace gemm;
  if (mask != nullptr && (((uintptr_t)mask & 15) || (ldmask % 8))) {
    set_error("pt_fc_gemm_bf16_ex: mask must be 16-byte aligned with ldmask a multiple of 8");
    return PT_ERR_ARG;
  }
  if (M <= 0) return PT_OK;
  if (N % BN != 0 || K % BK != 0 || K <= 0) {
    set_error("pt_fc_gemm_bf16: N must be a multiple of %d and K of %d (got N=%d K=%d)", BN, BK, N, K);
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
  CUtensorMap ma, mb;
  int rc = make_map(&ma, A, M, K, lda, BM);
  if (rc != PT_OK) return rc;
  rc = make_map(&mb, B, N, K, ldb, BN);
  if (rc != PT_OK) return rc;

  Params p;
  p.mask = reinterpret_cast<const __nv_bfloat16*>(mask); p.ldmask = (int)ldmask;
  p.bias = bias; p.C = C; p.M = M; p.N = N; p.K = K; p.ldc = (int)ldc; p.relu = relu; p.out_f32 = out_f32;
  p.tiles_n = N / BN;
  const int tiles_m = (M + BM - 1) / BM;
  p.tiles_total = tiles_m * p.tiles_n;
  int grid = p.tiles_total < num_sms ? p.tiles_total : num_sms;
  p.full_rounds = p.tiles_total / grid;
  const int tail = p.tiles_total - p.full_rounds * grid;
  p.split = 0;
  p.ws = nullptr; p.counters = nullptr;
  if (allow_split && tail > 0 && tail * 8 <= 4096 && workspace != nullptr) {
    int S = grid / tail;
    const int kb_total = K / BK;
    // every split must keep >= 8 k-blocks of MMA work, otherwise the fp32 reduction costs more than it saves; short
    // contractions (K < 4096: FC2, measured 43 us split vs 27 us unsplit at 5000x1024x1024) are never split
    if (S > kb_total / 8) S = kb_total / 8;
    if (kb_total < 64) S = 0;
    const long long need = (long long)tail * S * BM * BN * 4 + 4096;
    if (S >= 2 && need <= workspace_bytes) {
      p.split = S;
      // counters first (they must stay zero between launches; the partial slots need no initialisation)
      p.counters = reinterpret_cast<int*>(workspace);
      p.ws = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + 4096);   // counters live in the first 4 KB
    }
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(fc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return PT_ERR_CUDA; }
    attr_set = true;
  }
  fc_gemm_kernel<<<grid, NUM_THREADS, SMEM_BYTES, (cudaStream_t)stream>>>(ma, mb, p);
  return check_launch("fc_gemm_kernel");
}
