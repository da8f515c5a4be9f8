/* TEST INFRASTRUCTURE ONLY -- plain-C restatement of the mmcv-full 1.x kernels the
 * Point Teacher hot path reaches but that are NOT vendored under /root/reference
 * (SURVEY.md section 2.2, Appendix A.1-A.3).  Follows the published
 * Detectron2-lineage algorithm:
 *   - RoIAlign (avg, aligned)        <- call site HBB_TOD/mmdet/models/roi_heads/
 *                                       roi_extractors/base_roi_extractor.py:54-58
 *   - RoIAlignRotated                <- OBB_TOD/mmrotate/models/roi_heads/roi_extractors/
 *                                       rotate_single_level_roi_extractor.py:90-167
 *   - box_iou_rotated / nms_rotated  <- OBB_TOD/mmrotate/core/bbox/iou_calculators/
 *                                       rotate_iou2d_calculator.py:89 and HBB_TOD/mmdet/
 *                                       models/detectors/syn_images_generator_v2.py:667
 * RoIAlign is pinned against torchvision.ops.roi_align (tests/test_oracle.py); the rotated
 * ops are PARITY UNPINNED (no runnable reference), cross-checked at theta=0 and against
 * cv2.rotatedRectangleIntersection.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may call this. */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float x, y; } pt;

static inline float dot2(pt a, pt b) { return a.x * b.x + a.y * b.y; }
static inline float cross2(pt a, pt b) { return a.x * b.y - b.x * a.y; }
static inline pt sub(pt a, pt b) { pt r = {a.x - b.x, a.y - b.y}; return r; }

static void vertices(const float* b, pt* p) {
  double theta = b[4];
  float c2 = (float)cos(theta) * 0.5f, s2 = (float)sin(theta) * 0.5f;
  p[0].x = b[0] - s2 * b[3] - c2 * b[2];
  p[0].y = b[1] + c2 * b[3] - s2 * b[2];
  p[1].x = b[0] + s2 * b[3] - c2 * b[2];
  p[1].y = b[1] - c2 * b[3] - s2 * b[2];
  p[2].x = 2 * b[0] - p[0].x;
  p[2].y = 2 * b[1] - p[0].y;
  p[3].x = 2 * b[0] - p[1].x;
  p[3].y = 2 * b[1] - p[1].y;
}

static int intersections(const pt* p1, const pt* p2, pt* out) {
  pt v1[4], v2[4];
  for (int i = 0; i < 4; i++) {
    v1[i] = sub(p1[(i + 1) % 4], p1[i]);
    v2[i] = sub(p2[(i + 1) % 4], p2[i]);
  }
  int n = 0;
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) {
      float det = cross2(v2[j], v1[i]);
      if (fabsf(det) <= 1e-14f) continue;
      pt v12 = sub(p2[j], p1[i]);
      float t1 = cross2(v2[j], v12) / det;
      float t2 = cross2(v1[i], v12) / det;
      if (t1 >= 0.0f && t1 <= 1.0f && t2 >= 0.0f && t2 <= 1.0f) {
        out[n].x = p1[i].x + v1[i].x * t1;
        out[n].y = p1[i].y + v1[i].y * t1;
        n++;
      }
    }
  {
    pt AB = v2[0], DA = v2[3];
    float ABdotAB = dot2(AB, AB), ADdotAD = dot2(DA, DA);
    for (int i = 0; i < 4; i++) {
      pt AP = sub(p1[i], p2[0]);
      float APdotAB = dot2(AP, AB), APdotAD = -dot2(AP, DA);
      if (APdotAB >= 0 && APdotAD >= 0 && APdotAB <= ABdotAB && APdotAD <= ADdotAD) out[n++] = p1[i];
    }
  }
  {
    pt AB = v1[0], DA = v1[3];
    float ABdotAB = dot2(AB, AB), ADdotAD = dot2(DA, DA);
    for (int i = 0; i < 4; i++) {
      pt AP = sub(p2[i], p1[0]);
      float APdotAB = dot2(AP, AB), APdotAD = -dot2(AP, DA);
      if (APdotAB >= 0 && APdotAD >= 0 && APdotAB <= ABdotAB && APdotAD <= ADdotAD) out[n++] = p2[i];
    }
  }
  return n;
}

/* Graham scan, the device-side variant of the published kernel (selection-style sort on
 * the cross product with the 1e-6 collinearity band, ties by distance). */
static int convex_hull(const pt* p, int n, pt* q) {
  int t = 0;
  for (int i = 1; i < n; i++)
    if (p[i].y < p[t].y || (p[i].y == p[t].y && p[i].x < p[t].x)) t = i;
  pt start = p[t];
  for (int i = 0; i < n; i++) q[i] = sub(p[i], start);
  pt tmp = q[0]; q[0] = q[t]; q[t] = tmp;
  float dist[24];
  for (int i = 0; i < n; i++) dist[i] = dot2(q[i], q[i]);
  for (int i = 1; i < n - 1; i++)
    for (int j = i + 1; j < n; j++) {
      float cp = cross2(q[i], q[j]);
      if (cp < -1e-6f || (fabsf(cp) < 1e-6f && dist[i] > dist[j])) {
        pt qt = q[i]; q[i] = q[j]; q[j] = qt;
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
  return m; /* shift_to_zero = true: area is translation invariant */
}

static float poly_area(const pt* q, int m) {
  if (m <= 2) return 0.f;
  float a = 0.f;
  for (int i = 1; i < m - 1; i++) a += fabsf(cross2(sub(q[i], q[0]), sub(q[i + 1], q[0])));
  return a / 2.0f;
}

float oracle_single_iou_rotated(const float* r1, const float* r2, int mode) {
  float sx = (r1[0] + r2[0]) / 2.0f, sy = (r1[1] + r2[1]) / 2.0f;
  float b1[5] = {r1[0] - sx, r1[1] - sy, r1[2], r1[3], r1[4]};
  float b2[5] = {r2[0] - sx, r2[1] - sy, r2[2], r2[3], r2[4]};
  float a1 = b1[2] * b1[3], a2 = b2[2] * b2[3];
  if (a1 < 1e-14f || a2 < 1e-14f) return 0.f;
  pt p1[4], p2[4], ip[24], hull[24];
  vertices(b1, p1);
  vertices(b2, p2);
  int n = intersections(p1, p2, ip);
  float inter = 0.f;
  if (n > 2) {
    int m = convex_hull(ip, n, hull);
    inter = poly_area(hull, m);
  }
  float base = mode == 0 ? (a1 + a2 - inter) : a1;
  return inter / base;
}

void oracle_box_iou_rotated(const float* a, const float* b, float* out, long m, long n,
                            int aligned, int mode) {
  if (aligned) {
    for (long i = 0; i < m; i++) out[i] = oracle_single_iou_rotated(a + 5 * i, b + 5 * i, mode);
  } else {
    for (long i = 0; i < m; i++)
      for (long j = 0; j < n; j++)
        out[i * n + j] = oracle_single_iou_rotated(a + 5 * i, b + 5 * j, mode);
  }
}

/* nms_rotated CPU path: order = indices sorted by descending score (made by the caller);
 * keep[i]=1 for survivors; suppression test is `iou >= thr` (CPU kernel of the published
 * implementation; the CUDA one uses `>`; the two differ only at exact equality). */
void oracle_nms_rotated(const float* dets, const int64_t* order, uint8_t* keep, long n, float thr) {
  uint8_t* sup = (uint8_t*)calloc(n, 1);
  for (long _i = 0; _i < n; _i++) {
    long i = order[_i];
    if (sup[i]) continue;
    keep[i] = 1;
    for (long _j = _i + 1; _j < n; _j++) {
      long j = order[_j];
      if (sup[j]) continue;
      float ovr = oracle_single_iou_rotated(dets + 5 * i, dets + 5 * j, 0);
      if (ovr >= thr) sup[j] = 1;
    }
  }
  free(sup);
}

/* ------------------------------------------------------------------ RoIAlign */
static float bilinear(const float* f, int H, int W, float y, float x) {
  if (y < -1.0f || y > (float)H || x < -1.0f || x > (float)W) return 0.f;
  if (y <= 0) y = 0;
  if (x <= 0) x = 0;
  int yl = (int)y, xl = (int)x, yh, xh;
  if (yl >= H - 1) { yh = yl = H - 1; y = (float)yl; } else yh = yl + 1;
  if (xl >= W - 1) { xh = xl = W - 1; x = (float)xl; } else xh = xl + 1;
  float ly = y - yl, lx = x - xl, hy = 1.f - ly, hx = 1.f - lx;
  float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
  return w1 * f[yl * W + xl] + w2 * f[yl * W + xh] + w3 * f[yh * W + xl] + w4 * f[yh * W + xh];
}

void oracle_roi_align(const float* x, const float* rois, float* out, int B, int C, int H, int W,
                      long K, int P, float scale, int sampling_ratio, int aligned) {
  (void)B;
  float off = aligned ? 0.5f : 0.f;
  for (long k = 0; k < K; k++) {
    const float* r = rois + 5 * k;
    int b = (int)r[0];
    float x1 = r[1] * scale - off, y1 = r[2] * scale - off;
    float x2 = r[3] * scale - off, y2 = r[4] * scale - off;
    float rw = x2 - x1, rh = y2 - y1;
    if (!aligned) { rw = fmaxf(rw, 1.f); rh = fmaxf(rh, 1.f); }
    float bh = rh / (float)P, bw = rw / (float)P;
    int gh = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(rh / (float)P);
    int gw = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(rw / (float)P);
    float count = (float)(gh * gw > 1 ? gh * gw : 1);
    for (int c = 0; c < C; c++) {
      const float* f = x + ((long)b * C + c) * H * W;
      for (int ph = 0; ph < P; ph++)
        for (int pw = 0; pw < P; pw++) {
          float acc = 0.f;
          for (int iy = 0; iy < gh; iy++) {
            float y = y1 + ph * bh + (iy + .5f) * bh / (float)gh;
            for (int ix = 0; ix < gw; ix++) {
              float xx = x1 + pw * bw + (ix + .5f) * bw / (float)gw;
              acc += bilinear(f, H, W, y, xx);
            }
          }
          out[((k * C + c) * P + ph) * P + pw] = acc / count;
        }
    }
  }
}

void oracle_roi_align_rotated(const float* x, const float* rois, float* out, int B, int C, int H,
                              int W, long K, int P, float scale, int sampling_ratio, int aligned,
                              int clockwise) {
  (void)B;
  float off = aligned ? 0.5f : 0.f;
  for (long k = 0; k < K; k++) {
    const float* r = rois + 6 * k;
    int b = (int)r[0];
    float cx = r[1] * scale - off, cy = r[2] * scale - off;
    float rw = r[3] * scale, rh = r[4] * scale;
    float theta = r[5];
    if (clockwise) theta = -theta;
    if (!aligned) { rw = fmaxf(rw, 1.f); rh = fmaxf(rh, 1.f); }
    float bh = rh / (float)P, bw = rw / (float)P;
    int gh = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(rh / (float)P);
    int gw = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(rw / (float)P);
    float sh = -rh / 2.0f, sw = -rw / 2.0f;
    float ct = cosf(theta), st = sinf(theta);
    float count = (float)(gh * gw > 1 ? gh * gw : 1);
    for (int c = 0; c < C; c++) {
      const float* f = x + ((long)b * C + c) * H * W;
      for (int ph = 0; ph < P; ph++)
        for (int pw = 0; pw < P; pw++) {
          float acc = 0.f;
          for (int iy = 0; iy < gh; iy++) {
            float yy = sh + ph * bh + (iy + .5f) * bh / (float)gh;
            for (int ix = 0; ix < gw; ix++) {
              float xx = sw + pw * bw + (ix + .5f) * bw / (float)gw;
              float y = yy * ct - xx * st + cy;
              float xr = yy * st + xx * ct + cx;
              acc += bilinear(f, H, W, y, xr);
            }
          }
          out[((k * C + c) * P + ph) * P + pw] = acc / count;
        }
    }
  }
}
