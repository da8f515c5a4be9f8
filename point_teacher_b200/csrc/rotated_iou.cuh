// Rotated-rectangle IoU on the device: the published Detectron2 / mmcv `box_iou_rotated` algorithm
// (vertices -> 16 edge/edge intersections + contained vertices -> Graham scan -> shoelace), fp32.
// Replaces mmcv.ops.box_iou_rotated as reached through
//   OBB_TOD/mmrotate/core/bbox/iou_calculators/rotate_iou2d_calculator.py:53-89 (rbbox_overlaps, w/h clamped >= 1e-3)
// and the IoU test inside mmcv.ops.nms_rotated (HBB_TOD/mmdet/models/detectors/syn_images_generator_v2.py:667).
#pragma once
#include "common.cuh"

namespace ptb {
namespace riou {

struct P2 { float x, y; };
__device__ __forceinline__ float dot2(P2 a, P2 b) { return a.x * b.x + a.y * b.y; }
__device__ __forceinline__ float cross2(P2 a, P2 b) { return a.x * b.y - b.x * a.y; }
__device__ __forceinline__ P2 sub(P2 a, P2 b) { return P2{a.x - b.x, a.y - b.y}; }

__device__ __forceinline__ void vertices(const float* b, P2* p) {
  const float c2 = cosf(b[4]) * 0.5f, s2 = sinf(b[4]) * 0.5f;
  p[0].x = b[0] - s2 * b[3] - c2 * b[2];
  p[0].y = b[1] + c2 * b[3] - s2 * b[2];
  p[1].x = b[0] + s2 * b[3] - c2 * b[2];
  p[1].y = b[1] - c2 * b[3] - s2 * b[2];
  p[2].x = 2 * b[0] - p[0].x; p[2].y = 2 * b[1] - p[0].y;
  p[3].x = 2 * b[0] - p[1].x; p[3].y = 2 * b[1] - p[1].y;
}

__device__ inline int intersections(const P2* p1, const P2* p2, P2* out) {
  P2 v1[4], v2[4];
#pragma unroll
  for (int i = 0; i < 4; i++) { v1[i] = sub(p1[(i + 1) & 3], p1[i]); v2[i] = sub(p2[(i + 1) & 3], p2[i]); }
  int n = 0;
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) {
      const float det = cross2(v2[j], v1[i]);
      if (fabsf(det) <= 1e-14f) continue;
      const P2 v12 = sub(p2[j], p1[i]);
      const float t1 = cross2(v2[j], v12) / det, t2 = cross2(v1[i], v12) / det;
      if (t1 >= 0.0f && t1 <= 1.0f && t2 >= 0.0f && t2 <= 1.0f) { out[n].x = p1[i].x + v1[i].x * t1; out[n].y = p1[i].y + v1[i].y * t1; n++; }
    }
  {
    const P2 AB = v2[0], DA = v2[3];
    const float ABAB = dot2(AB, AB), ADAD = dot2(DA, DA);
    for (int i = 0; i < 4; i++) {
      const P2 AP = sub(p1[i], p2[0]);
      const float a = dot2(AP, AB), d = -dot2(AP, DA);
      if (a >= 0 && d >= 0 && a <= ABAB && d <= ADAD) out[n++] = p1[i];
    }
  }
  {
    const P2 AB = v1[0], DA = v1[3];
    const float ABAB = dot2(AB, AB), ADAD = dot2(DA, DA);
    for (int i = 0; i < 4; i++) {
      const P2 AP = sub(p2[i], p1[0]);
      const float a = dot2(AP, AB), d = -dot2(AP, DA);
      if (a >= 0 && d >= 0 && a <= ABAB && d <= ADAD) out[n++] = p2[i];
    }
  }
  return n;
}

__device__ inline int convex_hull(const P2* p, int n, P2* q) {
  int t = 0;
  for (int i = 1; i < n; i++) if (p[i].y < p[t].y || (p[i].y == p[t].y && p[i].x < p[t].x)) t = i;
  const P2 start = p[t];
  for (int i = 0; i < n; i++) q[i] = sub(p[i], start);
  P2 tmp = q[0]; q[0] = q[t]; q[t] = tmp;
  float dist[24];
  for (int i = 0; i < n; i++) dist[i] = dot2(q[i], q[i]);
  for (int i = 1; i < n - 1; i++)
    for (int j = i + 1; j < n; j++) {
      const float cp = cross2(q[i], q[j]);
      if (cp < -1e-6f || (fabsf(cp) < 1e-6f && dist[i] > dist[j])) {
        P2 qt = q[i]; q[i] = q[j]; q[j] = qt;
        float dt = dist[i]; dist[i] = dist[j]; dist[j] = dt;
      }
    }
  int k;
  for (k = 1; k < n; k++) if (dist[k] > 1e-8f) break;
  if (k == n) { q[0] = p[t]; return 1; }
  q[1] = q[k];
  int m = 2;
  for (int i = k + 1; i < n; i++) {
    while (m > 1 && cross2(sub(q[i], q[m - 2]), sub(q[m - 1], q[m - 2])) >= 0) m--;
    q[m++] = q[i];
  }
  return m;
}

// mode 0: IoU, 1: IoF (intersection / area of box 1).  Boxes (cx, cy, w, h, theta[rad]).
__device__ inline float single_iou(const float* r1, const float* r2, int mode) {
  const float sx = (r1[0] + r2[0]) / 2.0f, sy = (r1[1] + r2[1]) / 2.0f;
  const float b1[5] = {r1[0] - sx, r1[1] - sy, r1[2], r1[3], r1[4]};
  const float b2[5] = {r2[0] - sx, r2[1] - sy, r2[2], r2[3], r2[4]};
  const float a1 = b1[2] * b1[3], a2 = b2[2] * b2[3];
  if (a1 < 1e-14f || a2 < 1e-14f) return 0.f;
  P2 p1[4], p2[4], ip[24], hull[24];
  vertices(b1, p1);
  vertices(b2, p2);
  const int n = intersections(p1, p2, ip);
  float inter = 0.f;
  if (n > 2) {
    const int m = convex_hull(ip, n, hull);
    if (m > 2) {
      for (int i = 1; i < m - 1; i++) inter += fabsf(cross2(sub(hull[i], hull[0]), sub(hull[i + 1], hull[0])));
      inter /= 2.0f;
    }
  }
  const float base = mode == 0 ? (a1 + a2 - inter) : a1;
  return inter / base;
}

// rbbox_overlaps' pre-clamp of w, h to >= 1e-3
__device__ __forceinline__ float clamped_iou(const float* a, const float* b, int mode) {
  const float x[5] = {a[0], a[1], fmaxf(a[2], 1e-3f), fmaxf(a[3], 1e-3f), a[4]};
  const float y[5] = {b[0], b[1], fmaxf(b[2], 1e-3f), fmaxf(b[3], 1e-3f), b[4]};
  return single_iou(x, y, mode);
}

}  // namespace riou
}  // namespace ptb
