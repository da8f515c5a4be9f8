// Dense-head label assignment of Point Teacher for sm_100a, without ever
// materialising the P x G cost / overlap matrices the reference builds.
//
//   TopkAssigner      HBB_TOD/mmdet/core/bbox/assigners/topk_assigner.py:54-147
//   FUSETopkAssigner  HBB_TOD/mmdet/core/bbox/assigners/fuse_topk_assigner.py:56-121
//   match costs       HBB_TOD/mmdet/core/bbox/match_costs/match_cost.py:54-100 (FocalLossCost), :188-214 (PointCost),
//                     :217-252 (InsiderCost)
//   MaxIoUAssigner    HBB_TOD/mmdet/core/bbox/assigners/max_iou_assigner.py:60-212
//   BboxDistanceMetric / BboxOverlaps2D matrices: overlap_metric.cuh
//
// The assignment indices are "whatever ATen's CPU torch.topk / Tensor.max return" (tie rule):
//   stage 1: per GT column the num_pre smallest point costs.  For k*64 <= P ATen runs std::partial_sort = a
//            k-element max-heap streamed over the P points in index order.  One WARP per GT replays exactly those
//            heap moves: 32 costs are evaluated per step, a ballot finds the lanes that beat the current heap top
//            and only those are pushed (in index order, re-checked against the updated top) by lane 0.
//            For k*64 > P ATen uses nth_element on the whole column: replayed by one thread per GT on a scratch
//            column (tiny P only).
//   stage 2: per GT i the reference ranks its num_pre candidate rows under EVERY column of the second cost
//            (quirk) and assigns the union of the per-column top-`topk` rows to i; later GTs overwrite earlier
//            ones.  One CTA per GT, one thread per column replaying the 5-element nth_element; winners are OR-ed
//            into a bit mask and written with atomicMax(i + 1) (sequential overwrite == largest i wins).
#include "overlap_metric.cuh"
#include "topk_replay.cuh"

namespace ptb {

constexpr int MAXK = 16;

__device__ __forceinline__ float sigmoid_ref(float x) { return fdiv(1.f, fadd(1.f, expf(-x))); }

// FocalLossCost table (P, C): pos - neg at every class, times weight
__global__ void focal_cost_table_kernel(const float* __restrict__ logits, long long n, float alpha, float gamma,
                                        float eps, float weight, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float p = sigmoid_ref(logits[i]);
  const float q = fsub(1.f, p);
  const float pg = gamma == 2.f ? fmul(p, p) : powf(p, gamma);
  const float qg = gamma == 2.f ? fmul(q, q) : powf(q, gamma);
  const float neg = fmul(fmul(-logf(fadd(q, eps)), fsub(1.f, alpha)), pg);
  const float pos = fmul(fmul(-logf(fadd(p, eps)), alpha), qg);
  out[i] = fmul(fsub(pos, neg), weight);
}

__device__ __forceinline__ float point_cost(const float* __restrict__ pts, int ldp, int p, float gx, float gy, int l2,
                                            float weight) {
  const float dx = fsub(pts[(size_t)p * ldp], gx), dy = fsub(pts[(size_t)p * ldp + 1], gy);
  const float d = l2 ? sqrtf(fadd(fmul(dx, dx), fmul(dy, dy))) : fadd(fabsf(dx), fabsf(dy));
  return fmul(d, weight);
}

// ---- stage 1, partial_sort path: one warp per GT
__global__ void __launch_bounds__(128)
topk_pre_heap_kernel(const float* __restrict__ pts, int ldp, int P, const float* __restrict__ gts, int ldg, int G,
                     int l2, float weight, int k, int* __restrict__ pre_idx) {
  __shared__ float s_v[4][MAXK];
  __shared__ int s_i[4][MAXK];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = blockIdx.x * 4 + w;
  if (g >= G) return;
  const float gx = gts[(size_t)g * ldg], gy = gts[(size_t)g * ldg + 1];
  PairArray a{s_v[w], s_i[w]};
  if (lane < k) { s_v[w][lane] = point_cost(pts, ldp, lane, gx, gy, l2, weight); s_i[w][lane] = lane; }
  __syncwarp();
  if (lane == 0) tk_make_heap<false>(a, 0, k);
  __syncwarp();
  for (int base = k; base < P; base += 32) {
    const int p = base + lane;
    VI mine; mine.i = p; mine.v = 0.f;
    bool pass = false;
    if (p < P) {
      mine.v = point_cost(pts, ldp, p, gx, gy, l2, weight);
      pass = tk_comp<false>(mine, a.get(0));
    }
    unsigned mask = __ballot_sync(0xffffffffu, pass);
    while (mask) {
      const int L = __ffs(mask) - 1;
      VI cand;
      cand.v = __shfl_sync(0xffffffffu, mine.v, L);
      cand.i = base + L;
      if (lane == 0 && tk_comp<false>(cand, a.get(0))) tk_adjust_heap<false>(a, 0, 0, k, cand);   // __pop_heap
      __syncwarp();
      pass = pass && lane > L && tk_comp<false>(mine, a.get(0));
      mask = __ballot_sync(0xffffffffu, pass);
    }
  }
  if (lane == 0) {
    tk_sort_heap<false>(a, 0, k);
    for (int j = 0; j < k; j++) pre_idx[(size_t)j * G + g] = s_i[w][j];
  }
}

// ---- stage 1, nth_element path (k*64 > P): one thread per GT on a scratch column
__global__ void topk_pre_full_kernel(const float* __restrict__ pts, int ldp, int P, const float* __restrict__ gts,
                                     int ldg, int G, int l2, float weight, int k, float* __restrict__ sc_v,
                                     int* __restrict__ sc_i, int* __restrict__ pre_idx) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= G) return;
  const float gx = gts[(size_t)g * ldg], gy = gts[(size_t)g * ldg + 1];
  float* v = sc_v + (size_t)g * P;
  int* ix = sc_i + (size_t)g * P;
  for (int p = 0; p < P; p++) { v[p] = point_cost(pts, ldp, p, gx, gy, l2, weight); ix[p] = p; }
  cpu_topk_replay(v, ix, P, k, false);
  for (int j = 0; j < k; j++) pre_idx[(size_t)j * G + g] = ix[j];
}

// ---- stage 2: one CTA per GT i
__global__ void __launch_bounds__(128)
topk_second_kernel(const int* __restrict__ pre_idx, int num_pre, int topk, int G, const float* __restrict__ fl_table,
                   int C, const long long* __restrict__ labels, const float* __restrict__ pred, int ldb,
                   const float* __restrict__ gts, int ldg, float loc_weight, int* __restrict__ assigned) {
  __shared__ int rows[MAXK];
  __shared__ float bx1[MAXK], by1[MAXK], bx2[MAXK], by2[MAXK];
  __shared__ unsigned win;
  const int i = blockIdx.x;
  if (threadIdx.x < num_pre) {
    const int r = pre_idx[(size_t)threadIdx.x * G + i];
    rows[threadIdx.x] = r;
    if (pred != nullptr) {
      const float* b = pred + (size_t)r * ldb;
      const float hw = fdiv(b[2], 2.f), hh = fdiv(b[3], 2.f);
      bx1[threadIdx.x] = fsub(b[0], hw); by1[threadIdx.x] = fsub(b[1], hh);
      bx2[threadIdx.x] = fadd(b[0], hw); by2[threadIdx.x] = fadd(b[1], hh);
    }
  }
  if (threadIdx.x == 0) win = 0;
  __syncthreads();
  if (num_pre <= topk) {
    if (threadIdx.x < num_pre) atomicMax(assigned + rows[threadIdx.x], i + 1);
    return;
  }
  unsigned mine = 0;
  for (int c = threadIdx.x; c < G; c += blockDim.x) {
    float v[MAXK]; int ix[MAXK];
    const int lab = (int)labels[c];
    const float gx = gts[(size_t)c * ldg], gy = gts[(size_t)c * ldg + 1];
    for (int r = 0; r < num_pre; r++) {
      float cost = fl_table[(size_t)rows[r] * C + lab];
      if (pred != nullptr) {
        const bool inside = gx >= bx1[r] && gx <= bx2[r] && gy >= by1[r] && gy <= by2[r];
        cost = fadd(cost, fmul(inside ? 0.f : 1.f, loc_weight));
      }
      v[r] = cost; ix[r] = r;
    }
    cpu_topk_replay(v, ix, num_pre, topk, false);
    for (int j = 0; j < topk; j++) mine |= 1u << ix[j];
  }
  if (mine) atomicOr(&win, mine);
  __syncthreads();
  if (threadIdx.x < num_pre && ((win >> threadIdx.x) & 1u)) atomicMax(assigned + rows[threadIdx.x], i + 1);
}

__global__ void assign_finalize_kernel(const int* __restrict__ assigned, const long long* __restrict__ gt_labels,
                                       int P, long long* __restrict__ gt_inds, long long* __restrict__ labels) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const int a = assigned[p];
  gt_inds[p] = a;
  if (labels != nullptr) labels[p] = (a > 0 && gt_labels != nullptr) ? gt_labels[a - 1] : -1;
}

// ---- metric matrix
__global__ void bbox_metric_kernel(const float* __restrict__ a, int lda, const float* __restrict__ b, int ldb,
                                   long long M, long long N, int calc, int mode, float eps, float* __restrict__ out) {
  const long long total = M * N;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long i = idx / N, j = idx - i * N;
    const float* pa = a + i * lda;
    const float* pb = b + j * ldb;
    out[idx] = pair_metric(calc, mode, pa[0], pa[1], pa[2], pa[3], pb[0], pb[1], pb[2], pb[3], eps);
  }
}

// Tiled variant: the matrix is a pure 4 B/element output stream (60 MB at G = 1500, A = 10^4), so the kernel is
// organised around the store: a thread owns 4 consecutive columns (its 4 boxes stay in registers) and walks 8 rows
// whose box is a broadcast load; one 16-byte streaming store per row.  The element arithmetic is pair_metric()
// unchanged, so every entry is bit-identical to the one-thread-per-element kernel.
constexpr int MM_ROWS = 8, MM_THREADS = 256;
__global__ void __launch_bounds__(MM_THREADS)
metric_matrix_kernel(const float* __restrict__ a, int lda, const float* __restrict__ b, int ldb, int M, int N, int calc,
                     int mode, float eps, float* __restrict__ out, int vec_ok) {
  const int j0 = (blockIdx.x * MM_THREADS + threadIdx.x) * 4;
  if (j0 >= N) return;
  float bx[4][4];
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const float* p = b + (size_t)min(j0 + q, N - 1) * ldb;
    bx[q][0] = __ldg(p); bx[q][1] = __ldg(p + 1); bx[q][2] = __ldg(p + 2); bx[q][3] = __ldg(p + 3);
  }
  const int i0 = blockIdx.y * MM_ROWS;
#pragma unroll 2
  for (int r = 0; r < MM_ROWS; r++) {
    const int i = i0 + r;
    if (i >= M) break;
    const float* pa = a + (size_t)i * lda;
    const float a0 = __ldg(pa), a1 = __ldg(pa + 1), a2 = __ldg(pa + 2), a3 = __ldg(pa + 3);
    float v[4];
#pragma unroll
    for (int q = 0; q < 4; q++) v[q] = pair_metric(calc, mode, a0, a1, a2, a3, bx[q][0], bx[q][1], bx[q][2], bx[q][3], eps);
    float* o = out + (size_t)i * N + j0;
    if (vec_ok && j0 + 3 < N) {
      __stcs(reinterpret_cast<float4*>(o), make_float4(v[0], v[1], v[2], v[3]));
    } else {
#pragma unroll
      for (int q = 0; q < 4; q++)
        if (j0 + q < N) o[q] = v[q];
    }
  }
}

int launch_metric_matrix(const float* a, int lda, const float* b, int ldb, long long M, long long N, int calc, int mode,
                         float eps, float* out, cudaStream_t stream) {
  const long long gy = (M + MM_ROWS - 1) / MM_ROWS, gx = (N + MM_THREADS * 4 - 1) / (MM_THREADS * 4);
  if (gy > 65535 || gx > 2147483647LL || M > 2147483647LL || N > 2147483647LL) return PT_ERR_UNSUPPORTED;
  const int vec_ok = (N % 4 == 0) && (((uintptr_t)out & 15) == 0);
  metric_matrix_kernel<<<dim3((unsigned)gx, (unsigned)gy), MM_THREADS, 0, stream>>>(a, lda, b, ldb, (int)M, (int)N, calc, mode,
                                                                                    eps, out, vec_ok);
  return check_launch("metric_matrix_kernel");
}

// ---- MaxIoUAssigner without the G x A matrix
// order-preserving float <-> uint32 so that atomicMax works on signed floats (GIoU is negative)
__device__ __forceinline__ unsigned enc_f(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float dec_f(unsigned e) {
  return __uint_as_float((e & 0x80000000u) ? (e & 0x7fffffffu) : ~e);
}

constexpr int GT_TILE = 256;

// The G x A work is tiled over BOTH axes (grid = anchor blocks x GT tiles): at G = 1500, A = 10^4 an anchors-only grid
// is 40 CTAs with a 1500-iteration serial loop per thread (0.72 ms); the 2-D grid runs 240 CTAs of 256 iterations.
// Cross-tile reductions are order-independent integer atomics, so the result does not depend on scheduling:
//   per anchor (max, FIRST argmax)  -> 64-bit atomicMax of  enc(v) << 32 | ~g       (key buffer = the gt_inds output)
//   per GT max over anchors         -> 32-bit atomicMax of  enc(v)
//   low-quality matches (last GT wins) -> 32-bit atomicMax of g + 1                 (buffer = argmax_ws, re-used)

// pass 1: per anchor max / first argmax over this GT tile; per GT max over this anchor block
__global__ void __launch_bounds__(256)
max_iou_pass1_kernel(const float* __restrict__ gts, int ldg, int G, const float* __restrict__ anchors, int lda, int A,
                     int calc, int mode, float eps, unsigned long long* __restrict__ key, unsigned* __restrict__ gt_max) {
  __shared__ float4 sg[GT_TILE];
  __shared__ unsigned smax[GT_TILE];
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  const int g0 = blockIdx.y * GT_TILE, n = min(GT_TILE, G - g0);
  float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
  if (a < A) { const float* p = anchors + (size_t)a * lda; bx = make_float4(p[0], p[1], p[2], p[3]); }
  if (threadIdx.x < n) {
    const float* p = gts + (size_t)(g0 + threadIdx.x) * ldg;
    sg[threadIdx.x] = make_float4(p[0], p[1], p[2], p[3]);
    smax[threadIdx.x] = 0u;   // below enc_f of every float
  }
  __syncthreads();
  if (a < A) {
    unsigned best = 0u; int bi = 0;
    for (int j = 0; j < n; j++) {
      const float4 gb = sg[j];
      const unsigned e = enc_f(pair_metric(calc, mode, gb.x, gb.y, gb.z, gb.w, bx.x, bx.y, bx.z, bx.w, eps));
      if (e > best) { best = e; bi = g0 + j; }                          // first index among equal maxima
      // the running maximum settles after a few anchors: a broadcast read filters almost every atomic
      if (e > smax[j]) atomicMax(&smax[j], e);
    }
    atomicMax(key + a, ((unsigned long long)best << 32) | (unsigned long long)(0xffffffffu - (unsigned)bi));
  }
  __syncthreads();
  if (threadIdx.x < n && smax[threadIdx.x] != 0u) atomicMax(gt_max + g0 + threadIdx.x, smax[threadIdx.x]);
}

// pass 1b (gt_max_assign_all == False): first anchor index achieving each GT's max
__global__ void __launch_bounds__(256)
max_iou_argmax_kernel(const float* __restrict__ gts, int ldg, int G, const float* __restrict__ anchors, int lda,
                      int A, int calc, int mode, float eps, const unsigned* __restrict__ gt_max,
                      int* __restrict__ gt_argmax) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= A) return;
  const float* p = anchors + (size_t)a * lda;
  const int g0 = blockIdx.y * GT_TILE, g1 = min(G, g0 + GT_TILE);
  for (int g = g0; g < g1; g++) {
    const float* q = gts + (size_t)g * ldg;
    const float v = pair_metric(calc, mode, q[0], q[1], q[2], q[3], p[0], p[1], p[2], p[3], eps);
    if (v == dec_f(gt_max[g])) atomicMin(gt_argmax + g, a);
  }
}

// between the passes: decode the key, apply the thresholds; the key buffer becomes gt_inds, lowq is cleared
__global__ void max_iou_mid_kernel(int A, float pos_thr, float neg_lo, float neg_hi, long long* __restrict__ gt_inds,
                                   float* __restrict__ max_ov, int* __restrict__ lowq) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= A) return;
  const unsigned long long k = reinterpret_cast<const unsigned long long*>(gt_inds)[a];
  const float m = dec_f((unsigned)(k >> 32));
  const int bi = (int)(0xffffffffu - (unsigned)(k & 0xffffffffull));
  long long asg = -1;
  if (m >= neg_lo && m < neg_hi) asg = 0;
  if (m >= pos_thr) asg = bi + 1;
  max_ov[a] = m;
  gt_inds[a] = asg;
  lowq[a] = 0;
}

// pass 2: low-quality matching (sequential "for i in range(num_gts)" == largest matching i wins)
__global__ void __launch_bounds__(256)
max_iou_pass2_kernel(const float* __restrict__ gts, int ldg, int G, const float* __restrict__ anchors, int lda, int A,
                     int calc, int mode, float eps, const unsigned* __restrict__ gt_max,
                     const int* __restrict__ gt_argmax, float min_pos, int assign_all, int* __restrict__ lowq) {
  __shared__ float4 sg[GT_TILE];
  __shared__ float sm[GT_TILE];
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  const int g0 = blockIdx.y * GT_TILE, n = min(GT_TILE, G - g0);
  if (threadIdx.x < n) {
    const float* p = gts + (size_t)(g0 + threadIdx.x) * ldg;
    sg[threadIdx.x] = make_float4(p[0], p[1], p[2], p[3]);
    sm[threadIdx.x] = dec_f(gt_max[g0 + threadIdx.x]);
  }
  __syncthreads();
  if (a >= A) return;
  const float* p = anchors + (size_t)a * lda;
  const float4 bx = make_float4(p[0], p[1], p[2], p[3]);
  int hit = 0;
  for (int j = 0; j < n; j++) {
    const float gm = sm[j];
    if (!(gm >= min_pos)) continue;
    if (assign_all) {
      const float4 gb = sg[j];
      const float v = pair_metric(calc, mode, gb.x, gb.y, gb.z, gb.w, bx.x, bx.y, bx.z, bx.w, eps);
      if (v == gm) hit = g0 + j + 1;
    } else if (gt_argmax[g0 + j] == a) {
      hit = g0 + j + 1;
    }
  }
  if (hit > 0) atomicMax(lowq + a, hit);
}

__global__ void max_iou_final_kernel(int A, int low_quality, const int* __restrict__ lowq,
                                     const long long* __restrict__ gt_labels, long long* __restrict__ gt_inds,
                                     long long* __restrict__ labels) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= A) return;
  long long asg = gt_inds[a];
  if (low_quality && lowq[a] > 0) asg = lowq[a];
  gt_inds[a] = asg;
  if (labels != nullptr) labels[a] = asg > 0 ? gt_labels[asg - 1] : -1;
}

// ---- coarse pseudo-box aggregation behind the FUSE assignment
//   _gnerate_pseudo_single   HBB_TOD/mmdet/models/dense_heads/fcos_head_p2b_ts.py:736-794
// The reference builds an M x G one-hot matrix and two matmuls; here every assigned point adds its
// score-weighted decoded box to its GT with six float atomics, and one thread per GT finishes.
__global__ void decode_ltrb_kernel(const float* __restrict__ pts, const float* __restrict__ ltrb, int P,
                                   float* __restrict__ xyxy, float* __restrict__ cxcywh) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const float px = pts[2 * p], py = pts[2 * p + 1];
  const float* d = ltrb + (size_t)p * 4;
  const float x1 = fsub(px, d[0]), y1 = fsub(py, d[1]), x2 = fadd(px, d[2]), y2 = fadd(py, d[3]);   // distance2bbox
  float* o = xyxy + (size_t)p * 4;
  o[0] = x1; o[1] = y1; o[2] = x2; o[3] = y2;
  float* c = cxcywh + (size_t)p * 4;                                                                  // bbox_xyxy_to_cxcywh
  c[0] = fdiv(fadd(x1, x2), 2.f); c[1] = fdiv(fadd(y1, y2), 2.f); c[2] = fsub(x2, x1); c[3] = fsub(y2, y1);
}

__global__ void pseudo_accumulate_kernel(const long long* __restrict__ gt_inds, const long long* __restrict__ labels,
                                         const float* __restrict__ cls, int C, const float* __restrict__ xyxy, int P,
                                         float* __restrict__ acc /* [G][6]: sum(box*s) x4, sum(s), count */) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const long long g1 = gt_inds[p];
  if (g1 == 0) return;
  const int g = (int)g1 - 1;
  const float s = sigmoid_ref(cls[(size_t)p * C + (int)labels[p]]);
  const float* b = xyxy + (size_t)p * 4;
  float* a = acc + (size_t)g * 6;
#pragma unroll
  for (int j = 0; j < 4; j++) atomicAdd(a + j, fmul(b[j], s));
  atomicAdd(a + 4, s);
  atomicAdd(a + 5, 1.f);
}

__global__ void pseudo_finalize_kernel(const float* __restrict__ acc, const float* __restrict__ gt_points,
                                       const float* __restrict__ gt_bboxes, int G, float filter_score,
                                       float* __restrict__ boxes, float* __restrict__ points,
                                       float* __restrict__ scores, long long* __restrict__ nums,
                                       unsigned char* __restrict__ valid, float* __restrict__ iou_sum) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  float iou = 0.f, one = 0.f;
  if (g < G) {
    const float* a = acc + (size_t)g * 6;
    const float gx = gt_points[2 * g], gy = gt_points[2 * g + 1];
    float b[4] = {fsub(gx, 4.f), fsub(gy, 4.f), fadd(gx, 4.f), fadd(gy, 4.f)};   // bbox_cxcywh_to_xyxy of an 8 x 8 box
    float px = gx, py = gy, sc = 0.f;
    const bool has = a[5] > 0.f;
    if (has) {
#pragma unroll
      for (int j = 0; j < 4; j++) b[j] = fdiv(a[j], a[4]);
      sc = fdiv(a[4], a[5]);
      px = fdiv(fadd(b[0], b[2]), 2.f); py = fdiv(fadd(b[1], b[3]), 2.f);
      const float* t = gt_bboxes + (size_t)g * 4;
      iou = pair_metric(0, METRIC_IOU, b[0], b[1], b[2], b[3], t[0], t[1], t[2], t[3], 1e-6f);
      one = 1.f;
    }
#pragma unroll
    for (int j = 0; j < 4; j++) boxes[(size_t)g * 4 + j] = b[j];
    points[2 * g] = px; points[2 * g + 1] = py;
    scores[g] = sc;
    nums[g] = (long long)a[5];
    valid[g] = (has && sc >= filter_score) ? 1 : 0;
  }
  iou = warp_sum(iou); one = warp_sum(one);
  if ((threadIdx.x & 31) == 0 && one > 0.f) { atomicAdd(iou_sum, iou); atomicAdd(iou_sum + 1, one); }
}

// ---- dense regression targets after the refinement
//   _get_target_pseudo_single   HBB_TOD/mmdet/models/dense_heads/fcos_head_p2b_ts.py:657-708
//   centerness_target           HBB_TOD/mmdet/models/dense_heads/fcos_head_p2b_ts.py:1019-1038
// The reference broadcasts a (P, G, 4) ltrb tensor and gathers one column; here each point reads its own box.
__global__ void ltrb_targets_kernel(const float* __restrict__ pts, const float* __restrict__ boxes,
                                    const long long* __restrict__ gt_inds, const long long* __restrict__ asg_labels,
                                    int P, int num_classes, float* __restrict__ targets, long long* __restrict__ labels,
                                    float* __restrict__ centerness) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const long long g1 = gt_inds[p];
  const int g = g1 != 0 ? (int)g1 - 1 : 0;          // background points are measured against box 0 (reference quirk)
  const float x = pts[2 * p], y = pts[2 * p + 1];
  const float* b = boxes + (size_t)g * 4;
  const float l = fsub(x, b[0]), t = fsub(y, b[1]), r = fsub(b[2], x), bt = fsub(b[3], y);
  float* o = targets + (size_t)p * 4;
  o[0] = l; o[1] = t; o[2] = r; o[3] = bt;
  labels[p] = g1 != 0 ? asg_labels[p] : (long long)num_classes;
  if (centerness != nullptr) {
    const float lr = fdiv(fmaxf(fminf(l, r), 0.01f), fmaxf(l, r));
    const float tb = fdiv(fmaxf(fminf(t, bt), 0.01f), fmaxf(t, bt));
    centerness[p] = sqrtf(fmul(lr, tb));
  }
}

}  // namespace ptb

using namespace ptb;

extern "C" int pt_focal_cost_table(const float* logits, long long n, float alpha, float gamma, float eps,
                                   float weight, float* out, void* stream) {
  if (n <= 0) return PT_OK;
  focal_cost_table_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(logits, n, alpha, gamma, eps,
                                                                                        weight, out);
  return check_launch("focal_cost_table_kernel");
}

// Stage 1 of Topk/FUSETopkAssigner.  pts [P, ldp] (x, y first), gts [G, ldg] (cx, cy first); pre_idx [num_pre, G]
// int32 = torch.topk(cost, num_pre, dim=0, largest=False).indices.  scratch_v / scratch_i ([G, P] each) are only
// needed (and only touched) when num_pre * 64 > P.
extern "C" int pt_topk_pre(const float* pts, int ldp, int P, const float* gts, int ldg, int G, int l2, float weight,
                           int num_pre, int* pre_idx, float* scratch_v, int* scratch_i, void* stream) {
  if (G <= 0) return PT_OK;
  if (num_pre < 1 || num_pre > MAXK) { set_error("pt_topk_pre: num_pre must be in 1..%d (got %d)", MAXK, num_pre); return PT_ERR_UNSUPPORTED; }
  if (num_pre > P) { set_error("pt_topk_pre: selected index k out of range (num_pre %d > %d points)", num_pre, P); return PT_ERR_ARG; }
  cudaStream_t s = (cudaStream_t)stream;
  if ((long long)num_pre * 64 <= P) {
    topk_pre_heap_kernel<<<(G + 3) / 4, 128, 0, s>>>(pts, ldp, P, gts, ldg, G, l2, weight, num_pre, pre_idx);
    return check_launch("topk_pre_heap_kernel");
  }
  if (scratch_v == nullptr || scratch_i == nullptr) { set_error("pt_topk_pre: scratch columns required when num_pre*64 > P"); return PT_ERR_ARG; }
  topk_pre_full_kernel<<<(G + 63) / 64, 64, 0, s>>>(pts, ldp, P, gts, ldg, G, l2, weight, num_pre, scratch_v, scratch_i,
                                                    pre_idx);
  return check_launch("topk_pre_full_kernel");
}

// Stage 2 + write-out.  fl_table [P, C] (may be NULL when num_pre <= topk), labels [G] int64, pred [P, ldb]
// (cx, cy, w, h) or NULL (TopkAssigner), assigned_ws [P] int32 scratch; gt_inds / out_labels [P] int64.
extern "C" int pt_topk_second(const int* pre_idx, int num_pre, int topk, int G, int P, const float* fl_table, int C,
                              const long long* labels, const float* pred, int ldb, const float* gts, int ldg,
                              float loc_weight, int* assigned_ws, long long* gt_inds, long long* out_labels,
                              void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (P <= 0) return PT_OK;
  cudaMemsetAsync(assigned_ws, 0, (size_t)P * sizeof(int), s);
  if (G > 0) {
    if (num_pre < 1 || num_pre > MAXK || topk < 1) { set_error("pt_topk_second: bad num_pre/topk"); return PT_ERR_ARG; }
    if (num_pre > topk && fl_table == nullptr) { set_error("pt_topk_second: cost table required when num_pre > topk"); return PT_ERR_ARG; }
    topk_second_kernel<<<G, 128, 0, s>>>(pre_idx, num_pre, topk, G, fl_table, C, labels, pred, ldb, gts, ldg, loc_weight,
                                         assigned_ws);
    int rc = check_launch("topk_second_kernel");
    if (rc != PT_OK) return rc;
  }
  assign_finalize_kernel<<<(P + 255) / 256, 256, 0, s>>>(assigned_ws, labels, P, gt_inds, out_labels);
  return check_launch("assign_finalize_kernel");
}

extern "C" int pt_bbox_metric(const float* a, int lda, const float* b, int ldb, long long M, long long N, int calc,
                              int mode, float eps, float* out, void* stream) {
  if (M * N <= 0) return PT_OK;
  if (calc < 0 || calc > 1 || mode < 0 || mode > METRIC_KL10 || (calc == 0 && mode > METRIC_GIOU)) {
    set_error("pt_bbox_metric: unsupported calculator %d / mode %d", calc, mode);
    return PT_ERR_ARG;
  }
  const int rc = launch_metric_matrix(a, lda, b, ldb, M, N, calc, mode, eps, out, (cudaStream_t)stream);
  if (rc != PT_ERR_UNSUPPORTED) return rc;
  const long long total = M * N;                      // > 524 280 rows: one thread per element
  const int blocks = (int)((total + 255) / 256 < 148LL * 32 ? (total + 255) / 256 : 148LL * 32);
  bbox_metric_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(a, lda, b, ldb, M, N, calc, mode, eps, out);
  return check_launch("bbox_metric_kernel");
}

// MaxIoUAssigner.assign (overlaps = calc(gts, anchors, mode)); ws: G uint32 (gt max) + G int32 (gt argmax).
extern "C" int pt_max_iou_assign(const float* gts, int ldg, int G, const float* anchors, int lda, int A, int calc,
                                 int mode, float eps, float pos_thr, float neg_lo, float neg_hi, float min_pos,
                                 int gt_max_assign_all, int match_low_quality, const long long* gt_labels,
                                 long long* gt_inds, float* max_overlaps, long long* labels, int* argmax_ws,
                                 unsigned* gt_ws, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (A <= 0) return PT_OK;
  if (G <= 0) { set_error("pt_max_iou_assign: G must be > 0 (the host wrapper handles the empty case)"); return PT_ERR_ARG; }
  if (calc < 0 || calc > 1 || mode < 0 || mode > METRIC_KL10 || (calc == 0 && mode > METRIC_GIOU)) {
    set_error("pt_max_iou_assign: unsupported calculator %d / mode %d", calc, mode);
    return PT_ERR_ARG;
  }
  int* gt_argmax = reinterpret_cast<int*>(gt_ws + G);
  cudaMemsetAsync(gt_ws, 0, (size_t)G * sizeof(unsigned), s);
  cudaMemsetAsync(gt_argmax, 0x7f, (size_t)G * sizeof(int), s);
  cudaMemsetAsync(gt_inds, 0, (size_t)A * sizeof(long long), s);       // the per-anchor (max, argmax) keys
  const int blocks = (A + 255) / 256;
  const dim3 grid2(blocks, (G + GT_TILE - 1) / GT_TILE);
  if (grid2.y > 65535) { set_error("pt_max_iou_assign: more than 16 M GTs"); return PT_ERR_UNSUPPORTED; }
  max_iou_pass1_kernel<<<grid2, 256, 0, s>>>(gts, ldg, G, anchors, lda, A, calc, mode, eps,
                                             reinterpret_cast<unsigned long long*>(gt_inds), gt_ws);
  int rc = check_launch("max_iou_pass1_kernel");
  if (rc != PT_OK) return rc;
  if (match_low_quality && !gt_max_assign_all) {
    max_iou_argmax_kernel<<<grid2, 256, 0, s>>>(gts, ldg, G, anchors, lda, A, calc, mode, eps, gt_ws, gt_argmax);
    rc = check_launch("max_iou_argmax_kernel");
    if (rc != PT_OK) return rc;
  }
  max_iou_mid_kernel<<<blocks, 256, 0, s>>>(A, pos_thr, neg_lo, neg_hi, gt_inds, max_overlaps, argmax_ws);
  if (match_low_quality) {
    max_iou_pass2_kernel<<<grid2, 256, 0, s>>>(gts, ldg, G, anchors, lda, A, calc, mode, eps, gt_ws, gt_argmax, min_pos,
                                               gt_max_assign_all, argmax_ws);
    rc = check_launch("max_iou_pass2_kernel");
    if (rc != PT_OK) return rc;
  }
  max_iou_final_kernel<<<blocks, 256, 0, s>>>(A, match_low_quality, argmax_ws, gt_labels, gt_inds, labels);
  return check_launch("max_iou_final_kernel");
}

// distance2bbox + bbox_xyxy_to_cxcywh (HBB_TOD/mmdet/core/bbox/transforms.py:134-166, 249-261)
extern "C" int pt_decode_ltrb(const float* points, const float* ltrb, int P, float* xyxy, float* cxcywh, void* stream) {
  if (P <= 0) return PT_OK;
  decode_ltrb_kernel<<<(P + 255) / 256, 256, 0, (cudaStream_t)stream>>>(points, ltrb, P, xyxy, cxcywh);
  return check_launch("decode_ltrb_kernel");
}

// _gnerate_pseudo_single aggregation (fcos_head_p2b_ts.py:762-790).  acc_ws [G*6] + iou_sum [2] are scratch (zeroed
// here); outputs: boxes [G,4], points [G,2], scores [G], assign_nums [G] int64, valid [G] uint8 (assigned AND
// score >= filter), iou_sum = (sum of IoU(pseudo, gt) over assigned GTs, their count).
extern "C" int pt_pseudo_aggregate(const long long* gt_inds, const long long* labels, const float* cls, int C,
                                   const float* xyxy, int P, const float* gt_points, const float* gt_bboxes, int G,
                                   float filter_score, float* acc_ws, float* boxes, float* points, float* scores,
                                   long long* assign_nums, unsigned char* valid, float* iou_sum, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (G <= 0) return PT_OK;
  cudaMemsetAsync(acc_ws, 0, (size_t)G * 6 * sizeof(float), s);
  cudaMemsetAsync(iou_sum, 0, 2 * sizeof(float), s);
  if (P > 0) {
    pseudo_accumulate_kernel<<<(P + 255) / 256, 256, 0, s>>>(gt_inds, labels, cls, C, xyxy, P, acc_ws);
    int rc = check_launch("pseudo_accumulate_kernel");
    if (rc != PT_OK) return rc;
  }
  pseudo_finalize_kernel<<<(G + 127) / 128, 128, 0, s>>>(acc_ws, gt_points, gt_bboxes, G, filter_score, boxes, points,
                                                          scores, assign_nums, valid, iou_sum);
  return check_launch("pseudo_finalize_kernel");
}

// targets [P,4] = (l, t, r, b) of every point against its assigned pseudo box (box 0 for background), labels [P]
// (num_classes for background), optional centerness [P] = centerness_target(targets).
extern "C" int pt_ltrb_targets(const float* points, const float* boxes, const long long* gt_inds,
                               const long long* assigned_labels, int P, int num_classes, float* targets,
                               long long* labels, float* centerness, void* stream) {
  if (P <= 0) return PT_OK;
  ltrb_targets_kernel<<<(P + 255) / 256, 256, 0, (cudaStream_t)stream>>>(points, boxes, gt_inds, assigned_labels, P,
                                                                        num_classes, targets, labels, centerness);
  return check_launch("ltrb_targets_kernel");
}
