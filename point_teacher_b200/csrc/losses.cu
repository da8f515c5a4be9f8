// The two remaining mmcv-native losses of the student branch, sm_100a.
//   * sigmoid focal loss   HBB_TOD/mmdet/models/losses/focal_loss.py:11-100 (mmcv.ops.sigmoid_focal_loss on CUDA,
//                          py_sigmoid_focal_loss on CPU -- the formula followed here), forward + gradient in one pass.
//   * rotated IoU losses   OBB_TOD/mmrotate/models/losses/rotated_iou_loss.py:17-147 (RotatedIoULoss, DN_IoULoss) on
//                          mmcv.ops.diff_iou_rotated_2d (mmcv/ops/diff_iou_rotated.py + sort_vertices kernel).
//     mmcv builds the intersection polygon from 24 candidate vertices (8 corners + 16 edge/edge intersections), sorts
//     the valid ones by angle and takes the shoelace area, back-propagating through the gathered vertices with
//     autograd (~60 tiny launches + one custom op per call).  Here one thread evaluates a box pair with forward-mode
//     dual numbers over the 5 predicted parameters, so the loss and its exact gradient come out of ONE launch; the
//     DN variant's 1 + 9 evaluations run in the same thread and the argmin carries the gradient (torch.min).
// Both are latency-bound element-wise kernels (tens of KB of traffic); the win is the launch count.
#include <math.h>

#include "common.cuh"

namespace ptb {

// ---------------------------------------------------------------------------------------- focal loss
// pred [N,C] logits, target [N] int64 in [0, C] (C = background).  wmode: 0 none, 1 per row [N], 2 per element [N*C].
// loss_elem / grad_elem (either may be null): unreduced weighted loss and d(weighted loss)/d pred; sum += total.
__global__ void focal_loss_kernel(const float* __restrict__ pred, const long long* __restrict__ target,
                                  const float* __restrict__ weight, int wmode, float gamma, float alpha, int N, int C,
                                  float* __restrict__ loss_elem, float* __restrict__ grad_elem, float* __restrict__ sum) {
  __shared__ float red[8];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float l = 0.f;
  if (i < N * C) {
    const int r = i / C, c = i - r * C;
    const float x = pred[i];
    const bool pos = target[r] == (long long)c;
    const float p = 1.f / (1.f + expf(-x));
    const float pt = pos ? 1.f - p : p;
    const float mod = powf(pt, gamma);
    // binary_cross_entropy_with_logits: max(x, 0) - x t + log1p(exp(-|x|))
    const float bce = fmaxf(x, 0.f) - (pos ? x : 0.f) + log1pf(expf(-fabsf(x)));
    const float a = pos ? alpha : 1.f - alpha;
    const float w = wmode == 0 ? 1.f : (wmode == 1 ? weight[r] : weight[i]);
    l = bce * (a * mod) * w;
    if (loss_elem) loss_elem[i] = l;
    if (grad_elem) {
      // d/dx [a pt^g bce]:  pos: -a pt^g (g p bce + (1 - p));   neg: a pt^g (g (1 - p) bce + p)
      // pt^g / pt * ... is avoided (pt may underflow): pt^(g-1) * p (1-p) = pt^g * (pos ? p : 1-p)
      const float g = pos ? -a * mod * (gamma * p * bce + (1.f - p)) : a * mod * (gamma * (1.f - p) * bce + p);
      grad_elem[i] = g * w;
    }
  }
  if (sum) {
    l = warp_sum(l);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = l;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int k = 0; k < (int)(blockDim.x >> 5); k++) t += red[k];
      atomicAdd(sum, t);
    }
  }
}

// ---------------------------------------------------------------------------------------- dual numbers, 5 partials
struct D5 {
  float v, d[5];
};
__device__ __forceinline__ D5 dc(float v) { D5 r; r.v = v; for (int i = 0; i < 5; i++) r.d[i] = 0.f; return r; }
__device__ __forceinline__ D5 dv(float v, int k) { D5 r = dc(v); r.d[k] = 1.f; return r; }
__device__ __forceinline__ D5 operator+(const D5& a, const D5& b) { D5 r; r.v = a.v + b.v; for (int i = 0; i < 5; i++) r.d[i] = a.d[i] + b.d[i]; return r; }
__device__ __forceinline__ D5 operator-(const D5& a, const D5& b) { D5 r; r.v = a.v - b.v; for (int i = 0; i < 5; i++) r.d[i] = a.d[i] - b.d[i]; return r; }
__device__ __forceinline__ D5 operator*(const D5& a, const D5& b) { D5 r; r.v = a.v * b.v; for (int i = 0; i < 5; i++) r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
__device__ __forceinline__ D5 operator/(const D5& a, const D5& b) {
  D5 r; r.v = a.v / b.v;
  const float ib = 1.f / b.v;
  for (int i = 0; i < 5; i++) r.d[i] = (a.d[i] - r.v * b.d[i]) * ib;
  return r;
}
__device__ __forceinline__ D5 operator*(const D5& a, float s) { D5 r; r.v = a.v * s; for (int i = 0; i < 5; i++) r.d[i] = a.d[i] * s; return r; }
__device__ __forceinline__ D5 operator+(const D5& a, float s) { D5 r = a; r.v += s; return r; }
__device__ __forceinline__ D5 operator-(const D5& a, float s) { D5 r = a; r.v -= s; return r; }
__device__ __forceinline__ D5 dneg(const D5& a) { D5 r; r.v = -a.v; for (int i = 0; i < 5; i++) r.d[i] = -a.d[i]; return r; }

constexpr float DIOU_EPS = 1e-8f;

// mmcv sort_vertices kernel: compare_vertices
__device__ __forceinline__ bool cmp_vert(float x1, float y1, float x2, float y2) {
  if (fabsf(x1 - x2) < DIOU_EPS && fabsf(y2 - y1) < DIOU_EPS) return false;
  if (y1 > 0.f && y2 < 0.f) return true;
  if (y1 < 0.f && y2 > 0.f) return false;
  const float n1 = x1 * x1 + y1 * y1 + DIOU_EPS, n2 = x2 * x2 + y2 * y2 + DIOU_EPS;
  const float diff = fabsf(x1) * x1 / n1 - fabsf(x2) * x2 / n2;
  if (y1 > 0.f && y2 > 0.f) return diff > DIOU_EPS;
  if (y1 < 0.f && y2 < 0.f) return diff < DIOU_EPS;
  return false;
}

__device__ __forceinline__ void corners_const(const float* b, float* cx, float* cy) {
  const float sn = sinf(b[4]), cs = cosf(b[4]);
  const float sx[4] = {0.5f, -0.5f, -0.5f, 0.5f}, sy[4] = {0.5f, 0.5f, -0.5f, -0.5f};
  for (int k = 0; k < 4; k++) {
    const float x4 = sx[k] * b[2], y4 = sy[k] * b[3];
    cx[k] = x4 * cs + y4 * (-sn) + b[0];
    cy[k] = x4 * sn + y4 * cs + b[1];
  }
}

// c1 in box c2 (mmcv box1_in_box2), values only
__device__ __forceinline__ bool in_box(float mx, float my, const float* cx, const float* cy) {
  const float abx = cx[1] - cx[0], aby = cy[1] - cy[0], adx = cx[3] - cx[0], ady = cy[3] - cy[0];
  const float amx = mx - cx[0], amy = my - cy[0];
  const float pab = abx * amx + aby * amy, nab = abx * abx + aby * aby;
  const float pad = adx * amx + ady * amy, nad = adx * adx + ady * ady;
  const float ra = pab / nab, rd = pad / nad;
  return (ra > -1e-6f) && (ra < 1.f + 1e-6f) && (rd > -1e-6f) && (rd < 1.f + 1e-6f);
}

// IoU(pred, target) with derivatives w.r.t. pred (x, y, w, h, alpha): mmcv diff_iou_rotated_2d
__device__ D5 diff_iou(const float* p, const float* t) {
  D5 vx[24], vy[24];
  bool mask[24];
  // corners of the predicted box (dual) and of the target (constant)
  {
    const D5 X = dv(p[0], 0), Y = dv(p[1], 1), Wd = dv(p[2], 2), Hd = dv(p[3], 3);
    D5 sn = dc(sinf(p[4])), cs = dc(cosf(p[4]));
    sn.d[4] = cs.v; cs.d[4] = -sn.v;
    const float sx[4] = {0.5f, -0.5f, -0.5f, 0.5f}, sy[4] = {0.5f, 0.5f, -0.5f, -0.5f};
    for (int k = 0; k < 4; k++) {
      const D5 x4 = Wd * sx[k], y4 = Hd * sy[k];
      vx[k] = x4 * cs + y4 * dneg(sn) + X;
      vy[k] = x4 * sn + y4 * cs + Y;
    }
  }
  float c2x[4], c2y[4], c1x[4], c1y[4];
  corners_const(t, c2x, c2y);
  for (int k = 0; k < 4; k++) { vx[4 + k] = dc(c2x[k]); vy[4 + k] = dc(c2y[k]); c1x[k] = vx[k].v; c1y[k] = vy[k].v; }
  // 16 edge / edge intersections
  for (int i = 0; i < 4; i++) {
    const D5 x1 = vx[i], y1 = vy[i], x2 = vx[(i + 1) & 3], y2 = vy[(i + 1) & 3];
    for (int j = 0; j < 4; j++) {
      const float x3 = c2x[j], y3 = c2y[j], x4 = c2x[(j + 1) & 3], y4 = c2y[(j + 1) & 3];
      const D5 num = (x1 - x2) * (y3 - y4) - (y1 - y2) * (x3 - x4);
      const D5 den_t = (x1 - x3) * (y3 - y4) - (y1 - y3) * (x3 - x4);
      const float den_u = (x1.v - x2.v) * (y1.v - y3) - (y1.v - y2.v) * (x1.v - x3);
      float tv = den_t.v / num.v, uv = -den_u / num.v;
      if (num.v == 0.f) { tv = -1.f; uv = -1.f; }
      const bool m = (tv > 0.f) && (tv < 1.f) && (uv > 0.f) && (uv < 1.f);
      const int k = 8 + i * 4 + j;
      mask[k] = m;
      if (m) {
        const D5 tt = den_t / (num + DIOU_EPS);
        vx[k] = x1 + tt * (x2 - x1);
        vy[k] = y1 + tt * (y2 - y1);
      } else {
        vx[k] = dc(0.f); vy[k] = dc(0.f);
      }
    }
  }
  for (int k = 0; k < 4; k++) {
    mask[k] = in_box(c1x[k], c1y[k], c2x, c2y);
    mask[4 + k] = in_box(c2x[k], c2y[k], c1x, c1y);
  }
  int nv = 0, pad = 8;
  float mx = 0.f, my = 0.f;
  for (int k = 0; k < 24; k++)
    if (mask[k]) { nv++; mx += vx[k].v; my += vy[k].v; }
  for (int k = 8; k < 24; k++)
    if (!mask[k]) { pad = k; break; }
  mx /= (float)nv; my /= (float)nv;
  int idx[9];
  if (nv < 3) {
    for (int j = 0; j < 9; j++) idx[j] = pad;
  } else {
    for (int j = 0; j < nv && j < 8; j++) {
      float xm = 1.f, ym = -DIOU_EPS;
      int take = 0;
      float x2 = 0.f, y2 = 0.f;
      if (j != 0) { x2 = vx[idx[j - 1]].v - mx; y2 = vy[idx[j - 1]].v - my; }
      for (int k = 0; k < 24; k++) {
        const float x = vx[k].v - mx, y = vy[k].v - my;
        if (mask[k] && cmp_vert(x, y, xm, ym)) {
          if (j == 0 || cmp_vert(x2, y2, x, y)) { xm = x; ym = y; take = k; }
        }
      }
      idx[j] = take;
    }
    const int nvc = nv < 8 ? nv : 8;
    idx[nvc] = idx[0];
    for (int j = nvc + 1; j < 9; j++) idx[j] = pad;
    if (nv == 8) {
      int counter = 0;
      for (int j = 0; j < 4; j++)
        for (int k = 4; k < 8; k++) counter += idx[k] == idx[j] ? 1 : 0;
      if (counter == 4) { idx[4] = idx[0]; for (int j = 5; j < 9; j++) idx[j] = pad; }
    }
  }
  D5 total = dc(0.f);
  for (int k = 0; k < 8; k++)
    total = total + (vx[idx[k]] * vy[idx[k + 1]] - vy[idx[k]] * vx[idx[k + 1]]);
  D5 area = total.v < 0.f ? dneg(total) : total;
  area = area * 0.5f;
  const D5 a1 = dv(p[2], 2) * dv(p[3], 3);
  const float a2 = t[2] * t[3];
  return area / (a1 + a2 - area);
}

__device__ __forceinline__ D5 iou_loss_from(D5 iou, int mode, float eps) {
  if (iou.v < eps) iou = dc(eps);                  // clamp(min=eps): no gradient below
  if (mode == 1) return dneg(iou) + 1.f;           // linear
  if (mode == 2) return dneg(iou * iou) + 1.f;     // square
  D5 r; r.v = -logf(iou.v);
  for (int i = 0; i < 5; i++) r.d[i] = -iou.d[i] / iou.v;
  return r;
}

// mode: 0 log, 1 linear, 2 square.  dn != 0: DN_iou_loss (base + min over the 3 x 3 size-jittered targets) / 2.
__global__ void rotated_iou_loss_kernel(const float* __restrict__ pred, const float* __restrict__ target, int n, int mode,
                                        float eps, int dn, float hyper, float* __restrict__ loss,
                                        float* __restrict__ grad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float p[5], t[5];
  for (int k = 0; k < 5; k++) { p[k] = pred[(size_t)i * 5 + k]; t[k] = target[(size_t)i * 5 + k]; }
  D5 out = iou_loss_from(diff_iou(p, t), mode, eps);
  if (dn) {
    const float anx = hyper / 2.f, w = t[2], h = t[3];
    D5 best = dc(3.0e38f);
    for (int a = -1; a <= 1; a++)
      for (int b = -1; b <= 1; b++) {
        float tt[5] = {t[0], t[1], w - anx * w * (float)a, h - anx * h * (float)b, t[4]};
        const D5 e = iou_loss_from(diff_iou(p, tt), mode, eps);
        if (e.v < best.v) best = e;                // torch.min(dim): first minimum carries the gradient
      }
    out = (out + best) * 0.5f;
  }
  loss[i] = out.v;
  if (grad)
    for (int k = 0; k < 5; k++) grad[(size_t)i * 5 + k] = out.d[k];
}

}  // namespace ptb

using namespace ptb;

extern "C" int pt_sigmoid_focal_loss(const float* pred, const long long* target, const float* weight, int weight_mode,
                                     float gamma, float alpha, int N, int C, float* loss_elem, float* grad_elem,
                                     float* sum, void* stream) {
  if (N <= 0 || C <= 0) return PT_OK;
  if (weight_mode < 0 || weight_mode > 2 || (weight_mode != 0 && weight == nullptr)) {
    set_error("pt_sigmoid_focal_loss: weight_mode 0 (none) | 1 (per row) | 2 (per element)");
    return PT_ERR_ARG;
  }
  const long long total = (long long)N * C;
  if (total >= (1ll << 31)) { set_error("pt_sigmoid_focal_loss: N * C too large"); return PT_ERR_ARG; }
  focal_loss_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(pred, target, weight, weight_mode,
                                                                                      gamma, alpha, N, C, loss_elem,
                                                                                      grad_elem, sum);
  return check_launch("focal_loss_kernel");
}

extern "C" int pt_rotated_iou_loss(const float* pred, const float* target, int n, int mode, float eps, int dn, float hyper,
                                   float* loss, float* grad, void* stream) {
  if (n <= 0) return PT_OK;
  if (mode < 0 || mode > 2) { set_error("pt_rotated_iou_loss: mode 0 (log) | 1 (linear) | 2 (square)"); return PT_ERR_ARG; }
  rotated_iou_loss_kernel<<<(n + 63) / 64, 64, 0, (cudaStream_t)stream>>>(pred, target, n, mode, eps, dn, hyper, loss, grad);
  return check_launch("rotated_iou_loss_kernel");
}
