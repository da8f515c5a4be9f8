// Pairwise box metrics of the dense-head assigners, in the reference's fp32 operation order (never contracted)
// so that matrix entries -- and every threshold / equality test made on them -- match the CPU reference.
//   calc 0  BboxOverlaps2D      HBB_TOD/mmdet/core/bbox/iou_calculators/iou2d_calculator.py:74-260
//           modes 0 iou, 1 iof, 2 giou; union = max(union, eps)
//   calc 1  BboxDistanceMetric  HBB_TOD/mmdet/core/bbox/iou_calculators/metric_calculator.py:44-185
//           modes 0 iou, 1 iof (== iou: reference quirk), 2 giou, 3 wd (normalised Wasserstein), 4 kl,
//           5 center_distance2, 6 exp_kl, 7 kl_10; eps is added INTO the union and the union is clamped by eps
#pragma once
#include "common.cuh"

namespace ptb {

enum { METRIC_IOU = 0, METRIC_IOF = 1, METRIC_GIOU = 2, METRIC_WD = 3, METRIC_KL = 4, METRIC_CD2 = 5,
       METRIC_EXP_KL = 6, METRIC_KL10 = 7 };

__device__ __forceinline__ float pair_metric(int calc, int mode, float ax1, float ay1, float ax2, float ay2,
                                             float bx1, float by1, float bx2, float by2, float eps) {
  const float a1 = fmul(fsub(ax2, ax1), fsub(ay2, ay1)), a2 = fmul(fsub(bx2, bx1), fsub(by2, by1));
  const float ow = fmaxf(fsub(fminf(ax2, bx2), fmaxf(ax1, bx1)), 0.f);
  const float oh = fmaxf(fsub(fminf(ay2, by2), fmaxf(ay1, by1)), 0.f);
  const float ov = fmul(ow, oh);
  if (calc == 0) {
    float uni = mode == METRIC_IOF ? a1 : fsub(fadd(a1, a2), ov);
    uni = fmaxf(uni, eps);
    float v = fdiv(ov, uni);
    if (mode == METRIC_GIOU) {
      const float ew = fmaxf(fsub(fmaxf(ax2, bx2), fminf(ax1, bx1)), 0.f);
      const float eh = fmaxf(fsub(fmaxf(ay2, by2), fminf(ay1, by1)), 0.f);
      const float ea = fmaxf(fmul(ew, eh), eps);
      v = fsub(v, fdiv(fsub(ea, uni), ea));
    }
    return v;
  }
  if (mode <= METRIC_GIOU) {
    const float uni = fmaxf(fadd(fsub(fadd(a1, a2), ov), eps), eps);
    float v = fdiv(ov, uni);
    if (mode == METRIC_GIOU) {
      const float ew = fmaxf(fsub(fmaxf(ax2, bx2), fminf(ax1, bx1)), 0.f);
      const float eh = fmaxf(fsub(fmaxf(ay2, by2), fminf(ay1, by1)), 0.f);
      const float ea = fmaxf(fmul(ew, eh), eps);
      v = fsub(v, fdiv(fsub(ea, uni), ea));
    }
    return v;
  }
  const float dx = fsub(fdiv(fadd(ax1, ax2), 2.f), fdiv(fadd(bx1, bx2), 2.f));
  const float dy = fsub(fdiv(fadd(ay1, ay2), 2.f), fdiv(fadd(by1, by2), 2.f));
  if (mode == METRIC_CD2) return fadd(fadd(fmul(dx, dx), fmul(dy, dy)), 1e-6f);
  const float w1 = fadd(fsub(ax2, ax1), eps), h1 = fadd(fsub(ay2, ay1), eps);
  const float w2 = fadd(fsub(bx2, bx1), eps), h2 = fadd(fsub(by2, by1), eps);
  if (mode == METRIC_WD) {
    const float cd = fadd(fadd(fmul(dx, dx), fmul(dy, dy)), eps);
    const float dw = fsub(w1, w2), dh = fsub(h1, h2);
    const float whd = fdiv(fadd(fmul(dw, dw), fmul(dh, dh)), 4.f);
    return fdiv(1.f, fadd(1.f, fadd(cd, whd)));
  }
  const float w1s = fmul(w1, w1), h1s = fmul(h1, h1), w2s = fmul(w2, w2), h2s = fmul(h2, h2);
  float kl = fadd(fdiv(w2s, w1s), fdiv(h2s, h1s));
  kl = fadd(kl, fdiv(fmul(4.f, fmul(dx, dx)), w1s));
  kl = fadd(kl, fdiv(fmul(4.f, fmul(dy, dy)), h1s));
  kl = fadd(kl, logf(fdiv(w1s, w2s)));
  kl = fadd(kl, logf(fdiv(h1s, h2s)));
  kl = fdiv(fsub(kl, 2.f), 2.f);
  if (mode == METRIC_KL) return fdiv(1.f, fadd(1.f, kl));
  if (mode == METRIC_KL10) return fdiv(1.f, fadd(10.f, kl));
  return expf(fdiv(-kl, 10.f));
}

// Tiled G x A matrix of pair_metric (assign.cu): 8 rows x 4 columns per thread, 16-byte streaming stores, no 64-bit
// index division.  Returns PT_OK or a negative code; falls back to nothing -- callers keep their aligned kernels.
int launch_metric_matrix(const float* a, int lda, const float* b, int ldb, long long M, long long N, int calc, int mode,
                         float eps, float* out, cudaStream_t stream);

}  // namespace ptb
