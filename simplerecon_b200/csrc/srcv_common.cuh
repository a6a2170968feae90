// Shared device-side geometry for the plane-sweep kernels (sm_100a).
//
// The arithmetic restated here is the reference's
//   BackprojectDepth.forward   utils/geometry_utils.py:51-59  (+0.5 pixel centres :34-44)
//   Project3D.forward          utils/geometry_utils.py:72-89  (eps 1e-8)
//   uv normalisation           modules/cost_volume.py:199, :587
//   F.grid_sample(bilinear, zeros, align_corners=False)  modules/cost_volume.py:201-212
// re-associated the B200 way: the depth-invariant part of the projection,
//   Hm = (K E)[:3,:3] invK[:3,:3]   and   t = (K E)[:3,3],
// is folded once per (frame, view) by the prep kernel (fp64, rounded once), so a
// plane hypothesis d at pixel centre p projects with three FMAs,  c = d (Hm p) + t,
// i.e. the plane-induced homography of the sweep.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace srcv {

constexpr float kEpsProj = 1e-8f;   // utils/geometry_utils.py:66
constexpr float kEpsNorm = 1e-12f;  // F.normalize default
constexpr float kEpsCos = 1e-5f;    // modules/cost_volume.py:687
constexpr float kLeaky = 0.01f;     // nn.LeakyReLU default slope

// Per (frame b, view k) constants, written by prep_kernel.  32 floats = 128 B.
struct ViewParams {
  float Hm[9];     // (K E)[:3,:3] @ invK[:3,:3], row-major
  float t[3];      // (K E)[:3,3]
  float centre[3]; // src_poses[:3,3]: source camera centre in the reference frame
  float comb;      // pose_distance: sqrt(t_meas^2 + r_meas^2)
  float rmeas;     // sqrt(2 (1 - min(3, tr R)/3))
  float tmeas;     // |src_poses[:3,3]|
  float pad[14];
};
static_assert(sizeof(ViewParams) == 128, "ViewParams must be 128 bytes");

// Per frame constants: invK[:3,:3] (row-major) for the ray r = invK3 p.
struct FrameParams {
  float invK[9];
  float pad[7];
};
static_assert(sizeof(FrameParams) == 64, "FrameParams must be 64 bytes");

// Bilinear footprint of one projected sample.
struct Taps {
  int x0, y0;        // top-left texel (may be out of range)
  float w[4];        // nw, ne, sw, se weights, already zeroed for invalid taps
                     // and for points behind the camera when `fold_mask`
  unsigned valid;    // bit0 nw, bit1 ne, bit2 sw, bit3 se in-bounds
  float zp;          // z' = z + eps   (the depth channel the reference returns)
  float px, py;      // projected pixel coordinates
};

// c = d * a + t, guarded divide (geometry_utils.py:83-89), grid normalisation
// (cost_volume.py:199) and ATen's unnormalise for align_corners=False.
__device__ __forceinline__ void project_point(float d, float ax, float ay, float az,
                                              float tx, float ty, float tz,
                                              float& px, float& py, float& zp) {
  const float cx = fmaf(d, ax, tx);
  const float cy = fmaf(d, ay, ty);
  const float z = fmaf(d, az, tz);
  zp = __fadd_rn(z, kEpsProj);
  const float s = (fabsf(z) > kEpsProj) ? __frcp_rn(zp) : 1.0f;
  px = __fmul_rn(cx, s);
  py = __fmul_rn(cy, s);
}

__device__ __forceinline__ void bilinear_taps(float px, float py, int W, int H,
                                              float inv_w, float inv_h, Taps& tp) {
  // g = 2*p*(1/size) - 1 ; i = ((g + 1)*size - 1)/2   (separate roundings, as torch)
  const float gx = __fadd_rn(__fmul_rn(__fmul_rn(2.0f, px), inv_w), -1.0f);
  const float gy = __fadd_rn(__fmul_rn(__fmul_rn(2.0f, py), inv_h), -1.0f);
  const float ix = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(gx, 1.0f), (float)W), -1.0f), 0.5f);
  const float iy = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(gy, 1.0f), (float)H), -1.0f), 0.5f);
  const float x0f = floorf(ix), y0f = floorf(iy);
  const float x1f = x0f + 1.0f, y1f = y0f + 1.0f;
  const float wx1 = ix - x0f, wx0 = x1f - ix;
  const float wy1 = iy - y0f, wy0 = y1f - iy;
  // in-bounds tests in float first: NaN and huge coordinates fail all of them
  const bool xa = (x0f >= 0.0f) && (x0f <= (float)(W - 1));
  const bool xb = (x1f >= 0.0f) && (x1f <= (float)(W - 1));
  const bool ya = (y0f >= 0.0f) && (y0f <= (float)(H - 1));
  const bool yb = (y1f >= 0.0f) && (y1f <= (float)(H - 1));
  const bool any = (xa || xb) && (ya || yb);
  // when no tap is valid the integer coordinates are never used; pin them
  tp.x0 = any ? (int)x0f : 0;
  tp.y0 = any ? (int)y0f : 0;
  tp.valid = (unsigned)(xa && ya) | ((unsigned)(xb && ya) << 1) |
             ((unsigned)(xa && yb) << 2) | ((unsigned)(xb && yb) << 3);
  tp.w[0] = (xa && ya) ? wx0 * wy0 : 0.0f;
  tp.w[1] = (xb && ya) ? wx1 * wy0 : 0.0f;
  tp.w[2] = (xa && yb) ? wx0 * wy1 : 0.0f;
  tp.w[3] = (xb && yb) ? wx1 * wy1 : 0.0f;
  tp.px = px;
  tp.py = py;
}

__device__ __forceinline__ float leaky(float x) { return x > 0.0f ? x : kLeaky * x; }

// argmax update with torch.argmax semantics: first index wins ties, NaN is max.
__device__ __forceinline__ void argmax_update(float v, float dval, float& best, float& best_d,
                                              bool first) {
  const bool take = first || (v > best) || ((v != v) && (best == best));
  if (take) { best = v; best_d = dval; }
}

}  // namespace srcv
