// strong_augmentation on the device, sm_100a.
//   HBB: HBB_TOD/mmdet/models/detectors/syn_images_generator_v2.py:24-132
//   OBB: OBB_TOD/mmrotate/models/detectors/syn_images_generator_v2.py:223-357
// The reference flips, (OBB) rotates with torchvision's TF.rotate (nearest, fill 0), resizes with
// F.interpolate(bilinear, align_corners=False), crops / zero-pads back to H x W and rounds -- four full-image passes
// plus a dozen tiny elementwise launches per image on the coordinate lists.  Here:
//   * augment_image_kernel: ONE pass.  Every output pixel pulls its four bilinear taps straight from the source image;
//     a tap at integer (y, x) of the "rotated, flipped" image is the nearest-neighbour source pixel of the rotation
//     grid, mirrored by the flip.  The arithmetic replays ATen's CPU kernels in explicit fp32 (oracle/augment.py:
//     bilinear_axis / bilinear_resize_exact / rotate_source_index), so the rounded image is bit-identical.
//     HBM-bound: 2 x B x C x H x W x 4 bytes.
//   * augment_coords_kernel: one CTA per (image, list): flip / rotate / in-image filter / scale / crop filter / shift
//     and the box re-normalisation (HBB: xyxy re-ordering; OBB: poly2obb_le90) with a stable in-CTA compaction, the
//     reference's op order in non-contracted fp32 -> kept sets and coordinates bit-exact (OBB box parameters go
//     through sin / cos / atan2 whose device implementations differ from the host's in the last ulp).
#include <math.h>

#include "common.cuh"

namespace ptb {

// per-image parameter block (floats; integers are exact)
enum {
  AP_FLIPX = 0, AP_FLIPY, AP_ROT, AP_R00, AP_R01, AP_R10, AP_R11, AP_SCALE_H, AP_SCALE_W, AP_START_Y, AP_START_X,
  AP_PAD, AP_SF, AP_CA, AP_SA, AP_BLANK_W, AP_BLANK_H, AP_STRIDE = 20
};

struct Axis { int i0, i1; float w0, w1; };

// ATen compute_source_index_and_lambda, align_corners = false (as compiled for the CPU: the source index is one FMA)
__device__ __forceinline__ Axis bilinear_axis(int dst, int in_size, int out_size) {
  Axis a;
  if (in_size == out_size) { a.i0 = a.i1 = dst; a.w0 = 1.f; a.w1 = 0.f; return a; }
  const float scale = fdiv((float)in_size, (float)out_size);
  float src = __fmaf_rn(scale, fadd((float)dst, 0.5f), -0.5f);
  if (src < 0.f) src = 0.f;
  int i0 = (int)floorf(src);
  if (i0 > in_size - 1) i0 = in_size - 1;
  float l1 = fsub(src, (float)i0);
  l1 = fminf(fmaxf(l1, 0.f), 1.f);
  a.i0 = i0; a.i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  a.w1 = l1; a.w0 = fsub(1.f, l1);
  return a;
}

// offset of the source pixel behind integer position (yy, xx) of the flipped (+ rotated) image, or -1 (fill)
__device__ __forceinline__ int tap_offset(int yy, int xx, int H, int W, const float* p) {
  if (p[AP_ROT] != 0.f) {
    // torchvision _gen_affine_grid + grid_sample(nearest, zeros, align_corners=False)
    const float xs = fadd((float)xx, fadd(fmul(-(float)W, 0.5f), 0.5f));
    const float ys = fadd((float)yy, fadd(fmul(-(float)H, 0.5f), 0.5f));
    const float gx = __fmaf_rn(ys, p[AP_R10], fmul(xs, p[AP_R00]));
    const float gy = __fmaf_rn(ys, p[AP_R11], fmul(xs, p[AP_R01]));
    const float fx = fsub(fmul(fadd(gx, 1.f), fdiv((float)W, 2.f)), 0.5f);
    const float fy = fsub(fmul(fadd(gy, 1.f), fdiv((float)H, 2.f)), 0.5f);
    const float rx = rintf(fx), ry = rintf(fy);
    if (!(rx >= 0.f && rx < (float)W && ry >= 0.f && ry < (float)H)) return -1;
    xx = (int)rx; yy = (int)ry;
  }
  if (p[AP_FLIPX] != 0.f) xx = W - 1 - xx;
  if (p[AP_FLIPY] != 0.f) yy = H - 1 - yy;
  return yy * W + xx;
}

__global__ void __launch_bounds__(256)
augment_image_kernel(const float* __restrict__ in, float* __restrict__ out, const float* __restrict__ params, int C,
                     int H, int W) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, b = blockIdx.z;
  if (x >= W) return;
  const float* p = params + (size_t)b * AP_STRIDE;
  const int sH = (int)p[AP_SCALE_H], sW = (int)p[AP_SCALE_W], sy = (int)p[AP_START_Y], sx = (int)p[AP_START_X];
  int ry, rx;
  bool inside = true;
  if (p[AP_PAD] != 0.f) { ry = y - sy; rx = x - sx; inside = ry >= 0 && ry < sH && rx >= 0 && rx < sW; }
  else { ry = y + sy; rx = x + sx; }
  const size_t plane = (size_t)H * W;
  const float* src = in + (size_t)b * C * plane;
  float* dst = out + (size_t)b * C * plane + (size_t)y * W + x;
  if (!inside) {
    for (int c = 0; c < C; c++) dst[c * plane] = 0.f;
    return;
  }
  const Axis ay = bilinear_axis(ry, H, sH), ax = bilinear_axis(rx, W, sW);
  const int o00 = tap_offset(ay.i0, ax.i0, H, W, p), o01 = tap_offset(ay.i0, ax.i1, H, W, p);
  const int o10 = tap_offset(ay.i1, ax.i0, H, W, p), o11 = tap_offset(ay.i1, ax.i1, H, W, p);
  for (int c = 0; c < C; c++) {
    const float* s = src + c * plane;
    const float v00 = o00 >= 0 ? __ldg(s + o00) : 0.f, v01 = o01 >= 0 ? __ldg(s + o01) : 0.f;
    const float v10 = o10 >= 0 ? __ldg(s + o10) : 0.f, v11 = o11 >= 0 ? __ldg(s + o11) : 0.f;
    // ATen Interpolate<2>: row = fma(v0, wx0, v1 * wx1); out = fma(row0, wy0, row1 * wy1); then torch.round
    const float r0 = __fmaf_rn(v00, ax.w0, fmul(v01, ax.w1));
    const float r1 = __fmaf_rn(v10, ax.w0, fmul(v11, ax.w1));
    dst[c * plane] = rintf(__fmaf_rn(r0, ay.w0, fmul(r1, ay.w1)));
  }
}

// ---------------------------------------------------------------------------------------- coordinate lists
// One CTA per (image, list); list 0 = GT points [n,2] (+ labels), list 1 = pseudo points [n,2] + boxes [n,BD] (+ labels).
// BD = 4 (HBB xyxy) or 5 (OBB cx,cy,w,h,theta).  Outputs are compacted to the front of each image's segment; counts[b][list].
__device__ __forceinline__ void flip_xy(float& x, float& y, const float* p, float W, float H) {
  if (p[AP_FLIPX] != 0.f) x = fsub(W, x);
  if (p[AP_FLIPY] != 0.f) y = fsub(H, y);
}
__device__ __forceinline__ void rot_xy(float& x, float& y, const float* p, float cx, float cy) {
  // ca * (x - cx) - sa * (y - cy) + cx ; sa * (x - cx) + ca * (y - cy) + cy    (separate torch ops)
  const float dx = fsub(x, cx), dy = fsub(y, cy);
  const float nx = fadd(fsub(fmul(p[AP_CA], dx), fmul(p[AP_SA], dy)), cx);
  const float ny = fadd(fadd(fmul(p[AP_SA], dx), fmul(p[AP_CA], dy)), cy);
  x = nx; y = ny;
}

__global__ void __launch_bounds__(256)
augment_coords_kernel(const float* __restrict__ gt_pts, const long long* __restrict__ gt_lab, const int* __restrict__ gt_off,
                      const float* __restrict__ ps_pts, const long long* __restrict__ ps_lab,
                      const float* __restrict__ ps_box, const int* __restrict__ ps_off, int BD,
                      const float* __restrict__ params, int H, int W, float* __restrict__ o_gt_pts,
                      long long* __restrict__ o_gt_lab, float* __restrict__ o_ps_pts, long long* __restrict__ o_ps_lab,
                      float* __restrict__ o_ps_box, int* __restrict__ counts) {
  __shared__ int warp_tot[8];
  __shared__ int base_s;
  const int b = blockIdx.x, list = blockIdx.y;
  const float* p = params + (size_t)b * AP_STRIDE;
  const int* off = list == 0 ? gt_off : ps_off;
  const int n0 = off[b], n = off[b + 1] - n0;
  const float* pts = (list == 0 ? gt_pts : ps_pts) + (size_t)n0 * 2;
  const long long* lab = (list == 0 ? gt_lab : ps_lab) + n0;
  float* opts = (list == 0 ? o_gt_pts : o_ps_pts) + (size_t)n0 * 2;
  long long* olab = (list == 0 ? o_gt_lab : o_ps_lab) + n0;
  const float Wf = (float)W, Hf = (float)H, cx = fdiv(Wf, 2.f), cy = fdiv(Hf, 2.f);
  const bool rot = p[AP_ROT] != 0.f, rotated_boxes = BD == 5;
  const float sf = p[AP_SF], bw = p[AP_BLANK_W], bh = p[AP_BLANK_H];
  const bool pad = p[AP_PAD] != 0.f;
  if (threadIdx.x == 0) base_s = 0;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i0 = 0; i0 < n; i0 += blockDim.x) {
    const int i = i0 + threadIdx.x;
    bool keep = false;
    float x = 0.f, y = 0.f, poly[8], box[4];
    if (i < n) {
      keep = true;
      x = pts[2 * i]; y = pts[2 * i + 1];
      flip_xy(x, y, p, Wf, Hf);
      if (list == 1) {
        const float* bx = ps_box + (size_t)(n0 + i) * BD;
        if (rotated_boxes) {
          // obb2poly_le90 (OBB_TOD/mmrotate/core/bbox/transforms.py:474-499): corners tl, tr, br, bl rotated by theta
          const float hw = fmul(bx[2], 0.5f), hh = fmul(bx[3], 0.5f);
          const float sn = sinf(bx[4]), cs = cosf(bx[4]);
          const float rxs[4] = {-hw, hw, hw, -hw}, rys[4] = {-hh, -hh, hh, hh};
#pragma unroll
          for (int k = 0; k < 4; k++) {
            poly[2 * k] = fadd(fadd(fmul(cs, rxs[k]), fmul(-sn, rys[k])), bx[0]);
            poly[2 * k + 1] = fadd(fadd(fmul(sn, rxs[k]), fmul(cs, rys[k])), bx[1]);
            flip_xy(poly[2 * k], poly[2 * k + 1], p, Wf, Hf);
          }
        } else {
          box[0] = bx[0]; box[1] = bx[1]; box[2] = bx[2]; box[3] = bx[3];
          flip_xy(box[0], box[1], p, Wf, Hf);
          flip_xy(box[2], box[3], p, Wf, Hf);
        }
      }
      if (rot) {
        rot_xy(x, y, p, cx, cy);
        if (list == 1) {
#pragma unroll
          for (int k = 0; k < 4; k++) rot_xy(poly[2 * k], poly[2 * k + 1], p, cx, cy);
        }
        keep = (0.f <= x) && (x < Wf) && (0.f <= y) && (y < Hf);
      }
      x = fmul(x, sf); y = fmul(y, sf);
      if (!pad) {
        keep = keep && (x >= bw) && (x < fadd(Wf, bw)) && (y >= bh) && (y < fadd(Hf, bh));
        x = fsub(x, bw); y = fsub(y, bh);
      } else {
        x = fadd(x, bw); y = fadd(y, bh);
      }
      if (list == 1) {
        const float sgn_w = pad ? bw : -bw, sgn_h = pad ? bh : -bh;
        if (rotated_boxes) {
#pragma unroll
          for (int k = 0; k < 4; k++) {
            poly[2 * k] = fadd(fmul(poly[2 * k], sf), sgn_w);
            poly[2 * k + 1] = fadd(fmul(poly[2 * k + 1], sf), sgn_h);
          }
        } else {
          box[0] = fadd(fmul(box[0], sf), sgn_w); box[2] = fadd(fmul(box[2], sf), sgn_w);
          box[1] = fadd(fmul(box[1], sf), sgn_h); box[3] = fadd(fmul(box[3], sf), sgn_h);
        }
      }
    }
    // stable compaction
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) warp_tot[warp] = __popc(m);
    __syncthreads();
    int pos = base_s + __popc(m & ((1u << lane) - 1u));
    for (int w = 0; w < warp; w++) pos += warp_tot[w];
    if (keep) {
      opts[2 * pos] = x; opts[2 * pos + 1] = y;
      olab[pos] = lab[i];
      if (list == 1) {
        float* ob = o_ps_box + (size_t)(n0 + pos) * BD;
        if (rotated_boxes) {
          // poly2obb_le90 (transforms.py:301-331)
          const float d12x = fsub(poly[0], poly[2]), d12y = fsub(poly[1], poly[3]);
          const float d23x = fsub(poly[2], poly[4]), d23y = fsub(poly[3], poly[5]);
          const float e1 = sqrtf(fadd(fmul(d12x, d12x), fmul(d12y, d12y)));
          const float e2 = sqrtf(fadd(fmul(d23x, d23x), fmul(d23y, d23y)));
          float ang = e1 > e2 ? atan2f(fsub(poly[3], poly[1]), fsub(poly[2], poly[0]))
                              : atan2f(fsub(poly[7], poly[1]), fsub(poly[6], poly[0]));
          // norm_angle 'le90': (a + pi/2) % pi - pi/2 with python-float constants applied to an fp32 tensor
          const float hpi = 1.5707963267948966f, pi = 3.141592653589793f;
          float t = fadd(ang, hpi);
          float r = fmodf(t, pi);
          if (r != 0.f && (r < 0.f)) r = fadd(r, pi);
          ang = fsub(r, hpi);
          ob[0] = fdiv(fadd(poly[0], poly[4]), 2.f);
          ob[1] = fdiv(fadd(poly[1], poly[5]), 2.f);
          ob[2] = fmaxf(e1, e2); ob[3] = fminf(e1, e2); ob[4] = ang;
        } else {
          // :115-121: w = |x0 - x2|, h = |y0 - y2|, (x, y) = mins, cxcywh -> xyxy
          const float w = fabsf(fsub(box[0], box[2])), h = fabsf(fsub(box[1], box[3]));
          const float x0 = fminf(box[0], box[2]), y0 = fminf(box[1], box[3]);
          const float ccx = fadd(x0, fdiv(w, 2.f)), ccy = fadd(y0, fdiv(h, 2.f));
          ob[0] = fsub(ccx, fmul(0.5f, w)); ob[1] = fsub(ccy, fmul(0.5f, h));
          ob[2] = fadd(ccx, fmul(0.5f, w)); ob[3] = fadd(ccy, fmul(0.5f, h));
        }
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < 8; w++) t += warp_tot[w]; base_s += t; }
    __syncthreads();
  }
  if (threadIdx.x == 0) counts[b * 2 + list] = base_s;
}

}  // namespace ptb

using namespace ptb;

extern "C" int pt_augment_param_stride(void) { return AP_STRIDE; }

// img / out: [B, C, H, W] fp32; params: [B, pt_augment_param_stride()] fp32 (see point_teacher_b200/augment.py)
extern "C" int pt_augment_image(const float* img, float* out, const float* params, int B, int C, int H, int W,
                                void* stream) {
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return PT_OK;
  if ((long long)H * W >= (1ll << 31) || H > 65535 || B > 65535) { set_error("pt_augment_image: image too large"); return PT_ERR_ARG; }
  dim3 grid((W + 255) / 256, H, B);
  augment_image_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(img, out, params, C, H, W);
  return check_launch("augment_image_kernel");
}

// packed lists with int32 offsets [B+1]; box_dim 4 (HBB) | 5 (OBB, angle version le90); counts: int32 [B, 2]
extern "C" int pt_augment_coords(const float* gt_pts, const long long* gt_lab, const int* gt_off, const float* ps_pts,
                                 const long long* ps_lab, const float* ps_box, const int* ps_off, int box_dim,
                                 const float* params, int B, int H, int W, float* o_gt_pts, long long* o_gt_lab,
                                 float* o_ps_pts, long long* o_ps_lab, float* o_ps_box, int* counts, void* stream) {
  if (B <= 0) return PT_OK;
  if (box_dim != 4 && box_dim != 5) { set_error("pt_augment_coords: box_dim must be 4 or 5"); return PT_ERR_ARG; }
  dim3 grid(B, 2);
  augment_coords_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(gt_pts, gt_lab, gt_off, ps_pts, ps_lab, ps_box, ps_off,
                                                               box_dim, params, H, W, o_gt_pts, o_gt_lab, o_ps_pts,
                                                               o_ps_lab, o_ps_box, counts);
  return check_launch("augment_coords_kernel");
}
