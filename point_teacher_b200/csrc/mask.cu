// Phase-1 random region masking for sm_100a: the deterministic tail of
//   generate_black_paper   HBB_TOD/mmdet/models/detectors/syn_images_generator_v2.py:664-690
// i.e. rotated NMS at 0.05 (mmcv.ops.nms_rotated), score < 1 filter, inside-image filter (local obb2xyxy :382-396),
// obb2poly_le90 (data_augument_bank.py:516-541), int32-truncated corners, cv2.fillPoly, pixels = 255 -- on the device,
// so the image never makes the reference's GPU -> CPU -> GPU round trip.
//
// Rasterisation parity: cv2.fillPoly is "edge lines (8-connected LineIterator, drawn left to right) UNION scanline
// spans between edge pairs in 16.16 fixed point with x1 = ceil, x2 = floor" (drawing.cpp: CollectPolyEdges +
// FillEdgeCollection).  Both are replayed with the same integer arithmetic: one warp per polygon, lanes 0..3 walk
// the four edges, all 32 lanes take scanlines (edge x at scanline y is x0 + (y - y0) * dx exactly, so scanlines are
// independent).  Pinned bit-exact against cv2 4.13 in tests.
#include "rotated_iou.cuh"

namespace ptb {

// order[rank] = i for the stable descending sort of the scores (ATen's CPU sort is stable)
__global__ void nms_rank_kernel(const float* __restrict__ scores, int lds, int N, int* __restrict__ order) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const float s = scores[(size_t)i * lds];
  int r = 0;
  for (int j = 0; j < N; j++) {
    const float t = scores[(size_t)j * lds];
    r += (t > s || (t == s && j < i)) ? 1 : 0;
  }
  order[r] = i;
}

// sup[i][w]: bit j of word w set when sorted box 32*w+j (> i) is suppressed by sorted box i (IoU >= thr)
__global__ void __launch_bounds__(128)
nms_mask_kernel(const float* __restrict__ dets, int ld, const int* __restrict__ order, int N, int nw, float thr,
                unsigned* __restrict__ sup) {
  const int i = blockIdx.x;
  const float* a = dets + (size_t)order[i] * ld;
  const float ra = 0.5f * sqrtf(a[2] * a[2] + a[3] * a[3]);
  for (int w = threadIdx.x >> 5; w < nw; w += blockDim.x >> 5) {
    const int j = w * 32 + (threadIdx.x & 31);
    bool hit = false;
    if (j > i && j < N) {
      const float* b = dets + (size_t)order[j] * ld;
      const float rb = 0.5f * sqrtf(b[2] * b[2] + b[3] * b[3]);
      const float dx = a[0] - b[0], dy = a[1] - b[1];
      const float reach = (ra + rb) * 1.0001f + 1e-3f;
      if (dx * dx + dy * dy <= reach * reach)            // disjoint circumcircles: IoU is exactly 0
        hit = riou::single_iou(a, b, 0) >= thr;
    }
    const unsigned bits = __ballot_sync(0xffffffffu, hit);
    if ((threadIdx.x & 31) == 0) sup[(size_t)i * nw + w] = bits;
  }
}

// greedy scan in score order by one warp; keep_sorted[i] = 1 for survivors
__global__ void nms_scan_kernel(const unsigned* __restrict__ sup, int N, int nw, unsigned char* __restrict__ keep_sorted) {
  extern __shared__ unsigned removed[];
  const int lane = threadIdx.x;
  for (int w = lane; w < nw; w += 32) removed[w] = 0u;
  __syncwarp();
  for (int i = 0; i < N; i++) {
    const bool dead = (removed[i >> 5] >> (i & 31)) & 1u;
    if (lane == 0) keep_sorted[i] = dead ? 0 : 1;
    if (!dead)
      for (int w = lane; w < nw; w += 32) removed[w] |= sup[(size_t)i * nw + w];
    __syncwarp();
  }
}

__device__ __forceinline__ void sincos_ref(float a, float& s, float& c) {
  double sd, cd;
  sincos((double)a, &sd, &cd);      // rounded once to fp32: the closest stand-in for the CPU libm result
  s = (float)sd; c = (float)cd;
}

// score < 1, inside-image test, compaction in score order, integer polygons.  One CTA (N <= a few thousand).
__global__ void __launch_bounds__(1024)
black_paper_select_kernel(const float* __restrict__ bb, int N, const int* __restrict__ order,
                          const unsigned char* __restrict__ keep_sorted, float imgsize, float* __restrict__ out_bb,
                          int* __restrict__ out_sel, int* __restrict__ polys, int* __restrict__ count,
                          const float* __restrict__ trig) {
  __shared__ int warp_sums[32];
  __shared__ int base;
  if (threadIdx.x == 0) base = 0;
  __syncthreads();
  for (int i0 = 0; i0 < N; i0 += blockDim.x) {
    const int i = i0 + threadIdx.x;
    bool ok = false;
    const float* b = nullptr;
    float s = 0.f, c = 1.f;
    if (i < N && keep_sorted[i]) {
      b = bb + (size_t)order[i] * 7;
      if (b[5] < 1.0f) {
        // trig != NULL: (sin, cos) of every candidate's angle as the HOST's torch.sin / torch.cos produced them -- the
        // only way to truncate a corner that sits on an integer boundary exactly like the reference's CPU libm does
        if (trig != nullptr) { s = trig[2 * (size_t)order[i]]; c = trig[2 * (size_t)order[i] + 1]; }
        else sincos_ref(b[4], s, c);
        const float ca = fabsf(c), sa = fabsf(s);
        const float dw = fadd(fmul(ca, b[2]), fmul(sa, b[3])), dh = fadd(fmul(sa, b[2]), fmul(ca, b[3]));
        const float x1 = fsub(b[0], fdiv(dw, 2.f)), y1 = fsub(b[1], fdiv(dh, 2.f));
        const float x2 = fadd(b[0], fdiv(dw, 2.f)), y2 = fadd(b[1], fdiv(dh, 2.f));
        const float lo = fminf(fminf(x1, y1), fminf(x2, y2)), hi = fmaxf(fmaxf(x1, y1), fmaxf(x2, y2));
        ok = lo >= 0.f && hi <= fsub(imgsize, 1.f);
      }
    }
    const unsigned m = __ballot_sync(0xffffffffu, ok);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) warp_sums[wid] = __popc(m);
    __syncthreads();
    int before = base;
    for (int w = 0; w < wid; w++) before += warp_sums[w];
    const int pos = before + __popc(m & ((1u << lane) - 1u));
    if (ok) {
      for (int k = 0; k < 7; k++) out_bb[(size_t)pos * 7 + k] = b[k];
      out_sel[pos] = order[i];
      const float hw = fmul(b[2], 0.5f), hh = fmul(b[3], 0.5f);
      const float px[4] = {-hw, hw, hw, -hw}, py[4] = {-hh, -hh, hh, hh};
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const float x = fadd(fadd(fmul(c, px[k]), fmul(-s, py[k])), b[0]);
        const float y = fadd(fadd(fmul(s, px[k]), fmul(c, py[k])), b[1]);
        polys[(size_t)pos * 8 + 2 * k] = (int)x;        // numpy astype(int32): truncation toward zero
        polys[(size_t)pos * 8 + 2 * k + 1] = (int)y;
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
      for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += warp_sums[w];
      base += t;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *count = base;
}

__device__ __forceinline__ void put_px(float* __restrict__ img, unsigned char* __restrict__ mask, int C, int H, int W,
                                       int x, int y, float value) {
  if ((unsigned)x >= (unsigned)W || (unsigned)y >= (unsigned)H) return;
  const size_t o = (size_t)y * W + x;
  if (img != nullptr)
    for (int c = 0; c < C; c++) img[(size_t)c * H * W + o] = value;
  if (mask != nullptr) mask[o] = 1;
}

// cv::Line(LINE_8) == LineIterator(pt1, pt2, 8, leftToRight = true)
__device__ inline void line8(float* img, unsigned char* mask, int C, int H, int W, int x0, int y0, int x1, int y1,
                             float value) {
  int dx = x1 - x0, dy = y1 - y0;
  if (dx < 0) { x0 = x1; y0 = y1; dx = -dx; dy = -dy; }
  const int sy = dy >= 0 ? 1 : -1;
  const int ady = dy >= 0 ? dy : -dy;
  int x = x0, y = y0;
  if (ady > dx) {
    int err = ady - 2 * dx;
    for (int i = 0; i <= ady; i++) {
      put_px(img, mask, C, H, W, x, y, value);
      const bool neg = err < 0;
      err += -2 * dx + (neg ? 2 * ady : 0);
      y += sy;
      x += neg ? 1 : 0;
    }
  } else {
    int err = dx - 2 * ady;
    for (int i = 0; i <= dx; i++) {
      put_px(img, mask, C, H, W, x, y, value);
      const bool neg = err < 0;
      err += -2 * ady + (neg ? 2 * dx : 0);
      x += 1;
      y += neg ? sy : 0;
    }
  }
}

// one warp per polygon (4 integer vertices)
__global__ void __launch_bounds__(128)
fill_polys_kernel(const int* __restrict__ polys, const int* __restrict__ count, int max_polys, float* __restrict__ img,
                  unsigned char* __restrict__ mask, int C, int H, int W, float value) {
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int n = count != nullptr ? min(*count, max_polys) : max_polys;
  if (wid >= n) return;
  const int* p = polys + (size_t)wid * 8;
  int vx[4], vy[4];
#pragma unroll
  for (int k = 0; k < 4; k++) { vx[k] = p[2 * k]; vy[k] = p[2 * k + 1]; }
  // edges k: v[k-1] -> v[k]
  int ey0[4], ey1[4];
  long long ex[4], edx[4];
  int ymin = 1 << 30, ymax = -(1 << 30), n_edges = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int ax = vx[(k + 3) & 3], ay = vy[(k + 3) & 3], bx = vx[k], by = vy[k];
    if (lane == k) line8(img, mask, C, H, W, ax, ay, bx, by, value);
    ey0[k] = 0; ey1[k] = 0; ex[k] = 0; edx[k] = 0;          // y0 == y1: inactive everywhere
    if (ay != by) {
      const long long num = ((long long)(bx - ax)) << 16;
      edx[k] = num / (long long)(by - ay);                   // C++ int64 division: truncation toward zero
      if (ay < by) { ey0[k] = ay; ey1[k] = by; ex[k] = ((long long)ax) << 16; }
      else { ey0[k] = by; ey1[k] = ay; ex[k] = ((long long)bx) << 16; }
      ymin = min(ymin, ey0[k]); ymax = max(ymax, ey1[k]);
      n_edges++;
    }
  }
  if (n_edges < 2) return;
  ymax = min(ymax, H);
  for (int y = max(ymin, 0) + lane; y < ymax; y += 32) {
    long long xs[4];
    int na = 0;
#pragma unroll
    for (int k = 0; k < 4; k++)
      if (ey0[k] <= y && y < ey1[k]) {
        // insertion into the sorted list of active crossings
        const long long x = ex[k] + (long long)(y - ey0[k]) * edx[k];
        int q = na++;
        while (q > 0 && xs[q - 1] > x) { xs[q] = xs[q - 1]; q--; }
        xs[q] = x;
      }
    for (int k = 0; k + 1 < na; k += 2) {
      int x1 = (int)((xs[k] + 65535) >> 16), x2 = (int)(xs[k + 1] >> 16);
      if (x1 < W && x2 >= 0) {
        x1 = max(x1, 0); x2 = min(x2, W - 1);
        for (int x = x1; x <= x2; x++) put_px(img, mask, C, H, W, x, y, value);
      }
    }
  }
}

}  // namespace ptb

using namespace ptb;

extern "C" long long pt_nms_rotated_workspace_bytes(int N) {
  const long long nw = (N + 31) / 32;
  return (long long)N * 4 + (long long)N * nw * 4 + 256;      // order + suppression bit matrix
}

// mmcv.ops.nms_rotated(dets[:, :5], scores, thr) as used at syn_images_generator_v2.py:667: order [N] int32 =
// box indices by descending score (stable), keep_sorted [N] uint8 = survivor flags in that order.
extern "C" int pt_nms_rotated(const float* dets, int ld, const float* scores, int lds, int N, float thr, int* order,
                              unsigned char* keep_sorted, void* workspace, long long workspace_bytes, void* stream) {
  if (N <= 0) return PT_OK;
  const int nw = (N + 31) / 32;
  if (N > 16384) { set_error("pt_nms_rotated: N = %d exceeds the supported 16384 boxes", N); return PT_ERR_UNSUPPORTED; }
  if (workspace == nullptr || workspace_bytes < (long long)N * nw * 4) { set_error("pt_nms_rotated: workspace too small"); return PT_ERR_ARG; }
  cudaStream_t s = (cudaStream_t)stream;
  unsigned* sup = reinterpret_cast<unsigned*>(workspace);
  nms_rank_kernel<<<(N + 127) / 128, 128, 0, s>>>(scores, lds, N, order);
  int rc = check_launch("nms_rank_kernel");
  if (rc != PT_OK) return rc;
  nms_mask_kernel<<<N, 128, 0, s>>>(dets, ld, order, N, nw, thr, sup);
  rc = check_launch("nms_mask_kernel");
  if (rc != PT_OK) return rc;
  nms_scan_kernel<<<1, 32, nw * sizeof(unsigned), s>>>(sup, N, nw, keep_sorted);
  return check_launch("nms_scan_kernel");
}

// The filters + polygon construction of generate_black_paper (:668-683).  bb [N,7]; out_bb [N,7], out_sel [N],
// polys [N,8] int32 are filled for the first *count rows (score order).
extern "C" int pt_black_paper_select_ex(const float* bb, int N, const int* order, const unsigned char* keep_sorted,
                                        float imgsize, float* out_bb, int* out_sel, int* polys, int* count,
                                        const float* trig, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (N <= 0) { cudaMemsetAsync(count, 0, sizeof(int), s); return PT_OK; }
  black_paper_select_kernel<<<1, 1024, 0, s>>>(bb, N, order, keep_sorted, imgsize, out_bb, out_sel, polys, count, trig);
  return check_launch("black_paper_select_kernel");
}
extern "C" int pt_black_paper_select(const float* bb, int N, const int* order, const unsigned char* keep_sorted,
                                     float imgsize, float* out_bb, int* out_sel, int* polys, int* count, void* stream) {
  return pt_black_paper_select_ex(bb, N, order, keep_sorted, imgsize, out_bb, out_sel, polys, count, nullptr, stream);
}

// cv2.fillPoly of max_polys (or *count, when count != NULL) integer quadrilaterals: img [C,H,W] fp32 <- value,
// mask [H,W] uint8 <- 1 (either may be NULL).
extern "C" int pt_fill_polys(const int* polys, const int* count, int max_polys, float* img, unsigned char* mask, int C,
                             int H, int W, float value, void* stream) {
  if (max_polys <= 0) return PT_OK;
  fill_polys_kernel<<<(max_polys + 3) / 4, 128, 0, (cudaStream_t)stream>>>(polys, count, max_polys, img, mask, C, H, W,
                                                                           value);
  return check_launch("fill_polys_kernel");
}
