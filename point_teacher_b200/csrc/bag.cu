// Dynamic proposal-bag construction, negative-bag weights and axis-aligned overlap kernels.
//
// Replaces (all paths under /root/reference):
//   fine_proposals_from_cfg      HBB_TOD/mmdet/models/detectors/syn_images_generator_v2.py:262-324
//   MIL_gen_proposals_from_cfg   same file :134-145 (the replication of reference / real boxes is index math here)
//   gen_negative_proposals       same file :234-259 (the IoU<0.3-against-every-base-bag test; the RNG draws are
//                                injected by the caller)
//   bbox2roi                     HBB_TOD/mmdet/core/bbox/transforms.py:58-78 (bags are emitted as RoIs)
//   bbox_overlaps                HBB_TOD/mmdet/core/bbox/iou_calculators/iou2d_calculator.py:74-260
// Bit-exact contract: every fp32 operation is issued in the reference's order with non-contracted intrinsics.
#include "common.cuh"
#include "overlap_metric.cuh"
#include "rotated_iou.cuh"

namespace ptb {

struct BagCfg {
  float ratios[16];
  float shake[8];
  int n_ratios, n_shake;
  float min_scale;
};

// iof(box, [0,0,w,h]) > 0.7   (syn_images_generator_v2.py:317-319)
__device__ __forceinline__ bool inside_image(float x1, float y1, float x2, float y2, float w, float h) {
  const float area = fmul(fsub(x2, x1), fsub(y2, y1));
  const float lx = fmaxf(x1, 0.f), ly = fmaxf(y1, 0.f), rx = fminf(x2, w), ry = fminf(y2, h);
  const float ow = fmaxf(fsub(rx, lx), 0.f), oh = fmaxf(fsub(ry, ly), 0.f);
  const float iof = fdiv(fmul(ow, oh), fmaxf(area, 1e-6f));
  return iof > 0.7f;
}

// One thread per generated proposal.  in_rois [G,5] = (img, x1,y1,x2,y2); out_rois [G*U,5]; valid [G*U].
__global__ void bag_gen_kernel(const float* __restrict__ in_rois, const float* __restrict__ img_wh, int B,
                               long long total, int U, const __grid_constant__ BagCfg cfg,
                               float* __restrict__ out_rois,
                               uint8_t* __restrict__ valid, int rotated) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int S = 1 + 4 * cfg.n_shake;
  const long long g = idx / U;
  const int u = (int)(idx - g * U);
  const int rr = u / S, s = u - rr * S;
  const float rw = cfg.ratios[rr / cfg.n_ratios], rh = cfg.ratios[rr % cfg.n_ratios];
  // rotated (OBB_TOD/.../syn_images_generator_v2.py:26-40): bags are generated on the horizontal box
  // cxcywh_to_xyxy(obb[:, :4]) and come back as (cx, cy, w, h) with the pseudo box's angle re-attached
  const float* r = in_rois + g * (rotated ? 6 : 5);
  const float bimg = r[0];
  float bx1 = r[1], by1 = r[2], bx2 = r[3], by2 = r[4];
  if (rotated) {
    const float ocx = r[1], ocy = r[2], ow = r[3], oh = r[4];
    bx1 = fsub(ocx, fmul(0.5f, ow)); by1 = fsub(ocy, fmul(0.5f, oh));
    bx2 = fadd(ocx, fmul(0.5f, ow)); by2 = fadd(ocy, fmul(0.5f, oh));
  }
  // xyxy -> cxcywh, clamp, scale, back (:277-283)
  float cx = fdiv(fadd(bx1, bx2), 2.f), cy = fdiv(fadd(by1, by2), 2.f);
  float w = fsub(bx2, bx1), h = fsub(by2, by1);
  w = fmul(fminf(fmaxf(w, cfg.min_scale), 1000.f), rw);
  h = fmul(fminf(fmaxf(h, cfg.min_scale), 1000.f), rh);
  float x1 = fsub(cx, fmul(0.5f, w)), y1 = fsub(cy, fmul(0.5f, h));
  float x2 = fadd(cx, fmul(0.5f, w)), y2 = fadd(cy, fmul(0.5f, h));
  if (s > 0) {
    // centre shake (:286-306): recomputed from the xyxy box exactly as the reference does
    const float ratio = cfg.shake[(s - 1) >> 2];
    const int dir = (s - 1) & 3;
    float pcx = fdiv(fadd(x1, x2), 2.f), pcy = fdiv(fadd(y1, y2), 2.f);
    const float pw = fsub(x2, x1), phh = fsub(y2, y1);
    if (dir == 0) pcx = fsub(pcx, fmul(ratio, pw));
    else if (dir == 1) pcx = fadd(pcx, fmul(ratio, pw));
    else if (dir == 2) pcy = fsub(pcy, fmul(ratio, phh));
    else pcy = fadd(pcy, fmul(ratio, phh));
    x1 = fsub(pcx, fmul(0.5f, pw)); y1 = fsub(pcy, fmul(0.5f, phh));
    x2 = fadd(pcx, fmul(0.5f, pw)); y2 = fadd(pcy, fmul(0.5f, phh));
  }
  int bi = (int)bimg;
  bi = bi < 0 ? 0 : (bi >= B ? B - 1 : bi);
  const float iw = img_wh[2 * bi], ih = img_wh[2 * bi + 1];
  if (rotated) {
    float* o = out_rois + idx * 6;
    o[0] = bimg; o[1] = fdiv(fadd(x1, x2), 2.f); o[2] = fdiv(fadd(y1, y2), 2.f);
    o[3] = fsub(x2, x1); o[4] = fsub(y2, y1); o[5] = r[5];
  } else {
    float* o = out_rois + idx * 5;
    o[0] = bimg; o[1] = x1; o[2] = y1; o[3] = x2; o[4] = y2;
  }
  valid[idx] = inside_image(x1, y1, x2, y2, iw, ih) ? 1 : 0;
}

__device__ __forceinline__ float iou_xyxy(float ax1, float ay1, float ax2, float ay2, float bx1, float by1,
                                          float bx2, float by2) {
  const float a1 = fmul(fsub(ax2, ax1), fsub(ay2, ay1)), a2 = fmul(fsub(bx2, bx1), fsub(by2, by1));
  const float ow = fmaxf(fsub(fminf(ax2, bx2), fmaxf(ax1, bx1)), 0.f);
  const float oh = fmaxf(fsub(fminf(ay2, by2), fmaxf(ay1, by1)), 0.f);
  const float ov = fmul(ow, oh);
  const float uni = fmaxf(fsub(fadd(a1, a2), ov), 1e-6f);
  return fdiv(ov, uni);
}

// One warp per negative: weight = all(IoU(neg, every base bag of the same image) < 0.3).
// bag_rois [Kb,5] sorted by image, bag_offsets [B+1].
__global__ void neg_weight_kernel(const float* __restrict__ neg_rois, int n_neg, const float* __restrict__ bag_rois,
                                  const int* __restrict__ bag_offsets, int B, uint8_t* __restrict__ weight,
                                  int rotated) {
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (wid >= n_neg) return;
  const int st = rotated ? 6 : 5;
  const float* n = neg_rois + (size_t)wid * st;
  int b = (int)n[0];
  b = b < 0 ? 0 : (b >= B ? B - 1 : b);
  const float x1 = n[1], y1 = n[2], x2 = n[3], y2 = n[4];
  bool ok = true;
  for (int i = bag_offsets[b] + lane; i < bag_offsets[b + 1]; i += 32) {
    const float* p = bag_rois + (size_t)i * st;
    // rotated: rbbox_overlaps(neg, bags) with the negative's (x1,y1,x2,y2,theta) read as (cx,cy,w,h,theta)
    // (OBB_TOD/.../syn_images_generator_v2.py:146-152, reference quirk)
    const float v = rotated ? riou::clamped_iou(n + 1, p + 1, 0) : iou_xyxy(x1, y1, x2, y2, p[1], p[2], p[3], p[4]);
    if (!(v < 0.3f)) ok = false;
  }
  ok = __all_sync(0xffffffffu, ok);
  if (lane == 0) weight[wid] = ok ? 1 : 0;
}

// Generic overlap matrix / aligned vector: mode 0 iou, 1 iof, 2 giou; a [M,4], b [N,4] (row strides in floats).
__global__ void bbox_overlaps_kernel(const float* __restrict__ a, int lda, const float* __restrict__ b, int ldb,
                                     long long M, long long N, int mode, int aligned, float eps,
                                     float* __restrict__ out) {
  const long long total = aligned ? M : M * N;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long i = aligned ? idx : idx / N, j = aligned ? idx : idx - i * N;
    const float* pa = a + i * lda;
    const float* pb = b + j * ldb;
    const float ax1 = pa[0], ay1 = pa[1], ax2 = pa[2], ay2 = pa[3];
    const float bx1 = pb[0], by1 = pb[1], bx2 = pb[2], by2 = pb[3];
    const float a1 = fmul(fsub(ax2, ax1), fsub(ay2, ay1)), a2 = fmul(fsub(bx2, bx1), fsub(by2, by1));
    const float ow = fmaxf(fsub(fminf(ax2, bx2), fmaxf(ax1, bx1)), 0.f);
    const float oh = fmaxf(fsub(fminf(ay2, by2), fmaxf(ay1, by1)), 0.f);
    const float ov = fmul(ow, oh);
    float uni = mode == 1 ? a1 : fsub(fadd(a1, a2), ov);
    uni = fmaxf(uni, eps);
    float v = fdiv(ov, uni);
    if (mode == 2) {
      const float ew = fmaxf(fsub(fmaxf(ax2, bx2), fminf(ax1, bx1)), 0.f);
      const float eh = fmaxf(fsub(fmaxf(ay2, by2), fminf(ay1, by1)), 0.f);
      const float ea = fmaxf(fmul(ew, eh), eps);
      v = fsub(v, fdiv(fsub(ea, uni), ea));
    }
    out[idx] = v;
  }
}

// rbbox_overlaps (rotate_iou2d_calculator.py:53-89): a [M,5], b [N,5] (cx,cy,w,h,theta), mode 0 iou / 1 iof
__global__ void box_iou_rotated_kernel(const float* __restrict__ a, int lda, const float* __restrict__ b, int ldb,
                                       long long M, long long N, int mode, int aligned, int clamp_wh,
                                       float* __restrict__ out) {
  const long long total = aligned ? M : M * N;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long i = aligned ? idx : idx / N, j = aligned ? idx : idx - i * N;
    out[idx] = clamp_wh ? riou::clamped_iou(a + i * lda, b + j * ldb, mode) : riou::single_iou(a + i * lda, b + j * ldb, mode);
  }
}

// mean aligned IoU of n box pairs (the coarse_bboxes_iou / refine_bboxes_iou logs,
// fcos_p2b_teacher_student.py:436-438, :457-459); single block, fixed summation order.
__global__ void aligned_iou_mean_kernel(const float* __restrict__ a, int lda, const float* __restrict__ b, int ldb,
                                        int n, float* __restrict__ out, int rotated) {
  __shared__ float red[32];
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float* pa = a + (size_t)i * lda; const float* pb = b + (size_t)i * ldb;
    acc += rotated ? riou::clamped_iou(pa, pb, 0) : iou_xyxy(pa[0], pa[1], pa[2], pa[3], pb[0], pb[1], pb[2], pb[3]);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) out[0] = n > 0 ? v / (float)n : 0.f;
  }
}

}  // namespace ptb

using namespace ptb;

extern "C" int pt_aligned_iou_mean(const float* a, int lda, const float* b, int ldb, int n, int rotated, float* out,
                                   void* stream) {
  aligned_iou_mean_kernel<<<1, rotated ? 256 : 1024, 0, (cudaStream_t)stream>>>(a, lda, b, ldb, n, out, rotated);
  return check_launch("aligned_iou_mean_kernel");
}

extern "C" int pt_box_iou_rotated(const float* a, int lda, const float* b, int ldb, long long M, long long N, int mode,
                                  int aligned, int clamp_wh, float* out, void* stream) {
  if (mode < 0 || mode > 1) { set_error("pt_box_iou_rotated: mode must be 0 iou / 1 iof"); return PT_ERR_ARG; }
  if (aligned && M != N) { set_error("pt_box_iou_rotated: aligned needs M == N"); return PT_ERR_ARG; }
  const long long total = aligned ? M : M * N;
  if (total <= 0) return PT_OK;
  long long blocks = (total + 127) / 128;
  if (blocks > 148LL * 16) blocks = 148LL * 16;
  box_iou_rotated_kernel<<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(a, lda, b, ldb, M, N, mode, aligned,
                                                                             clamp_wh, out);
  return check_launch("box_iou_rotated_kernel");
}

extern "C" int pt_bag_gen(const float* in_rois, long long G, const float* img_wh, int B, const float* ratios,
                          int n_ratios, const float* shake, int n_shake, float min_scale, float* out_rois,
                          unsigned char* valid, int rotated, void* stream) {
  if (n_ratios <= 0 || n_ratios > 16 || n_shake < 0 || n_shake > 8) {
    set_error("pt_bag_gen: need 1..16 base ratios and 0..8 shake ratios (got %d, %d)", n_ratios, n_shake);
    return PT_ERR_ARG;
  }
  BagCfg cfg;
  for (int i = 0; i < 16; i++) cfg.ratios[i] = i < n_ratios ? ratios[i] : 1.f;
  for (int i = 0; i < 8; i++) cfg.shake[i] = i < n_shake ? shake[i] : 0.f;
  cfg.n_ratios = n_ratios; cfg.n_shake = n_shake; cfg.min_scale = min_scale;
  const int U = n_ratios * n_ratios * (1 + 4 * n_shake);
  const long long total = G * U;
  if (total <= 0) return PT_OK;
  const int threads = 256;
  const long long blocks = (total + threads - 1) / threads;
  bag_gen_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(in_rois, img_wh, B, total, U, cfg, out_rois,
                                                                         valid, rotated);
  return check_launch("bag_gen_kernel");
}

extern "C" int pt_neg_weight(const float* neg_rois, int n_neg, const float* bag_rois, const int* bag_offsets, int B,
                             unsigned char* weight, int rotated, void* stream) {
  if (n_neg <= 0) return PT_OK;
  const int threads = 256, wpb = threads / 32;
  neg_weight_kernel<<<(n_neg + wpb - 1) / wpb, threads, 0, (cudaStream_t)stream>>>(neg_rois, n_neg, bag_rois,
                                                                                   bag_offsets, B, weight, rotated);
  return check_launch("neg_weight_kernel");
}

extern "C" int pt_bbox_overlaps(const float* a, int lda, const float* b, int ldb, long long M, long long N, int mode,
                                int aligned, float eps, float* out, void* stream) {
  if (mode < 0 || mode > 2) { set_error("pt_bbox_overlaps: mode must be 0 iou / 1 iof / 2 giou"); return PT_ERR_ARG; }
  if (aligned && M != N) { set_error("pt_bbox_overlaps: aligned needs M == N"); return PT_ERR_ARG; }
  const long long total = aligned ? M : M * N;
  if (total <= 0) return PT_OK;
  if (!aligned) {   // the M x N matrix: tiled kernel of assign.cu (same element arithmetic, calc 0)
    const int rc = launch_metric_matrix(a, lda, b, ldb, M, N, 0, mode, eps, out, (cudaStream_t)stream);
    if (rc != PT_ERR_UNSUPPORTED) return rc;
  }
  const int threads = 256;
  long long blocks = (total + threads - 1) / threads;
  if (blocks > 148LL * 32) blocks = 148LL * 32;
  bbox_overlaps_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(a, lda, b, ldb, M, N, mode, aligned,
                                                                               eps, out);
  return check_launch("bbox_overlaps_kernel");
}

namespace ptb {
// boxes [n, ldb>=4(+1 angle)] + per-box image index -> RoIs [n, 5|6] (bbox2roi / rbbox2roi without host loops)
__global__ void make_rois_kernel(const float* __restrict__ boxes, int ldb, const int* __restrict__ img, int n,
                                 int nbox, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float* o = out + (size_t)i * (nbox + 1);
  o[0] = (float)img[i];
  for (int j = 0; j < nbox; j++) o[1 + j] = boxes[(size_t)i * ldb + j];
}
}  // namespace ptb

extern "C" int pt_make_rois(const float* boxes, int ldb, const int* img_idx, int n, int box_dim, float* out_rois,
                            void* stream) {
  if (n <= 0) return PT_OK;
  if (box_dim != 4 && box_dim != 5) { ptb::set_error("pt_make_rois: box_dim must be 4 or 5"); return PT_ERR_ARG; }
  ptb::make_rois_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(boxes, ldb, img_idx, n, box_dim, out_rois);
  return ptb::check_launch("make_rois_kernel");
}
