/*
 * Plain-C restatement of SimpleRecon's plane-sweep cost-volume path.
 *
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.  Built by oracle/Makefile into
 * oracle/_build/libcvoracle.so; only tests/ and bench.py's cpu_baseline leg load it.
 *
 * It is a second, independent restatement next to oracle/costvolume_oracle.py (which
 * leans on ATen ops like the reference does): here every step is spelled out as scalar
 * loops, in the reference's own operation order, so it (a) cross-checks the torch oracle
 * and (b) gives a CPU baseline that is not dominated by ATen dispatch overhead (OpenMP
 * over pixels).  REAL selects fp32 (cvo_*_f32) or fp64 (cvo_*_f64).
 *
 * Reference lines restated (relative to the reference tree):
 *   BackprojectDepth.forward         utils/geometry_utils.py:51-59  (+0.5 centres :34-44)
 *   Project3D.forward                utils/geometry_utils.py:72-89  (eps 1e-8)
 *   uv normalisation                 modules/cost_volume.py:199, :587
 *   grid_sample bilinear/zeros/ac=F  call sites :201-212, :590-601 (ATen-CUDA unnormalise order)
 *   depth mask, dot, view sum        modules/cost_volume.py:231-232, :322-333
 *   argmax -> depth                  modules/cost_volume.py:338-342, :374-378
 *   rays, ray angle, concat order    modules/cost_volume.py:641-723, utils/geometry_utils.py:168-173
 *   pose_distance                    utils/geometry_utils.py:178-191
 *   MLP (LeakyReLU 0.01)             modules/networks.py:129-147
 *   overall mask (last plane)        modules/cost_volume.py:625-637, bounds :90-95
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef REAL
#define REAL float
#endif
#ifndef SUFFIX
#define SUFFIX f32
#endif
#define CAT_(a, b) a##_##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUFFIX)

typedef struct {
  REAL px, py, zp; /* projected pixel coordinates, z' = z + eps */
} proj_t;

/* P = K @ E (4x4, row-major), fp arithmetic of REAL */
static void matmul4(const REAL* K, const REAL* E, REAL* P) {
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      REAL a = 0;
      for (int l = 0; l < 4; ++l) a += K[i * 4 + l] * E[l * 4 + j];
      P[i * 4 + j] = a;
    }
}

static proj_t project(const REAL* P, REAL X, REAL Y, REAL Z) {
  const REAL eps = (REAL)1e-8;
  const REAL cx = P[0] * X + P[1] * Y + P[2] * Z + P[3];
  const REAL cy = P[4] * X + P[5] * Y + P[6] * Z + P[7];
  const REAL z = P[8] * X + P[9] * Y + P[10] * Z + P[11];
  proj_t r;
  r.zp = z + eps;
  const REAL s = (fabs((double)z) > 1e-8) ? (REAL)1 / r.zp : (REAL)1;
  r.px = cx * s;
  r.py = cy * s;
  return r;
}

/* bilinear, zeros padding, align_corners=False at pixel coords (px, py): out[c], c < C */
static void sample(const REAL* img, int C, int H, int W, REAL px, REAL py, REAL* out) {
  const REAL gx = (REAL)2 * px * ((REAL)1 / (REAL)W) - (REAL)1;
  const REAL gy = (REAL)2 * py * ((REAL)1 / (REAL)H) - (REAL)1;
  const REAL ix = ((gx + (REAL)1) * (REAL)W - (REAL)1) / (REAL)2;
  const REAL iy = ((gy + (REAL)1) * (REAL)H - (REAL)1) / (REAL)2;
  for (int c = 0; c < C; ++c) out[c] = 0;
  if (!(fabs((double)ix) < 1e9) || !(fabs((double)iy) < 1e9)) return; /* NaN / inf: all padding */
  const REAL x0 = (REAL)floor((double)ix), y0 = (REAL)floor((double)iy);
  const REAL x1 = x0 + 1, y1 = y0 + 1;
  const REAL w[4] = {(x1 - ix) * (y1 - iy), (ix - x0) * (y1 - iy), (x1 - ix) * (iy - y0), (ix - x0) * (iy - y0)};
  const REAL xs[4] = {x0, x1, x0, x1}, ys[4] = {y0, y0, y1, y1};
  for (int t = 0; t < 4; ++t) {
    if (xs[t] < 0 || xs[t] > (REAL)(W - 1) || ys[t] < 0 || ys[t] > (REAL)(H - 1)) continue;
    const long off = (long)ys[t] * W + (long)xs[t];
    for (int c = 0; c < C; ++c) out[c] += w[t] * img[(long)c * H * W + off];
  }
}

static REAL leaky(REAL x) { return x > 0 ? x : (REAL)0.01 * x; }

/* planes: per_pixel ? (B,D,H,W) : (B,D).  cost (B,D,H,W), lowest (B,H,W) (may be NULL). */
void FN(cvo_dot)(int B, int K, int C, int H, int W, int D, const REAL* cur, const REAL* src,
                 const REAL* E, const REAL* Ks, const REAL* invK, const REAL* planes, int per_pixel,
                 REAL* cost, REAL* lowest) {
  const long HW = (long)H * W;
  REAL* Pall = (REAL*)malloc(sizeof(REAL) * 16 * (size_t)B * K); /* P = K @ E per (frame, view) */
  for (long i = 0; i < (long)B * K; ++i) matmul4(Ks + i * 16, E + i * 16, Pall + i * 16);
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < B; ++b)
    for (long p = 0; p < HW; ++p) {
      REAL warped[1024];
      const REAL u = (REAL)(p % W) + (REAL)0.5, v = (REAL)(p / W) + (REAL)0.5;
      const REAL* iK = invK + (long)b * 16;
      const REAL rx = iK[0] * u + iK[1] * v + iK[2];
      const REAL ry = iK[4] * u + iK[5] * v + iK[6];
      const REAL rz = iK[8] * u + iK[9] * v + iK[10];
      REAL best = 0, best_d = 0;
      for (int d = 0; d < D; ++d) {
        const REAL dv = per_pixel ? planes[((long)b * D + d) * HW + p] : planes[(long)b * D + d];
        const REAL X = dv * rx, Y = dv * ry, Z = dv * rz;
        REAL acc = 0;
        for (int k = 0; k < K; ++k) {
          const proj_t pr = project(Pall + ((long)b * K + k) * 16, X, Y, Z);
          sample(src + ((long)b * K + k) * C * HW, C, H, W, pr.px, pr.py, warped);
          REAL dot = 0;
          for (int c = 0; c < C; ++c) dot += warped[c] * cur[((long)b * C + c) * HW + p];
          acc += dot * (pr.zp > 0 ? (REAL)1 : (REAL)0);
        }
        cost[((long)b * D + d) * HW + p] = acc;
        if (d == 0 || acc > best || (acc != acc && best == best)) { best = acc; best_d = dv; }
      }
      if (lowest) lowest[(long)b * HW + p] = best_d;
    }
  free(Pall);
}

/* weights: nn.Linear layout w1 (H1,F), b1, w2 (H2,H1), b2, w3 (1,H2), b3.  mask (B,H,W) u8 or NULL. */
void FN(cvo_mlp)(int B, int K, int C, int H, int W, int D, const REAL* cur, const REAL* src,
                 const REAL* E, const REAL* poses, const REAL* Ks, const REAL* invK,
                 const REAL* planes, int per_pixel, const REAL* w1, const REAL* b1, const REAL* w2,
                 const REAL* b2, const REAL* w3, const REAL* b3, int H1, int H2, REAL* cost,
                 REAL* lowest, uint8_t* mask) {
  const long HW = (long)H * W;
  const int F = C * (K + 1) + 10 * K + 4;
  const int o_cur = K * C, o_mask = o_cur + C, o_z = o_mask + K, o_depth = o_z + K, o_dot = o_depth + 1,
            o_ang = o_dot + K, o_ncur = o_ang + K, o_nsrc = o_ncur + 3, o_comb = o_nsrc + 3 * K,
            o_r = o_comb + K, o_t = o_r + K;
  REAL* Pall = (REAL*)malloc(sizeof(REAL) * 16 * (size_t)B * K);
  for (long i = 0; i < (long)B * K; ++i) matmul4(Ks + i * 16, E + i * 16, Pall + i * 16);
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < B; ++b)
    for (long p = 0; p < HW; ++p) {
      REAL* f = (REAL*)malloc(sizeof(REAL) * (size_t)(F + H1 + H2));
      REAL* h1 = f + F;
      REAL* h2 = h1 + H1;
      const REAL u = (REAL)(p % W) + (REAL)0.5, v = (REAL)(p / W) + (REAL)0.5;
      const REAL* iK = invK + (long)b * 16;
      const REAL rx = iK[0] * u + iK[1] * v + iK[2];
      const REAL ry = iK[4] * u + iK[5] * v + iK[6];
      const REAL rz = iK[8] * u + iK[9] * v + iK[10];
      REAL best = 0, best_d = 0;
      for (int d = 0; d < D; ++d) {
        const REAL dv = per_pixel ? planes[((long)b * D + d) * HW + p] : planes[(long)b * D + d];
        const REAL X = dv * rx, Y = dv * ry, Z = dv * rz;
        const REAL nc = (REAL)fmax(sqrt((double)(X * X + Y * Y + Z * Z)), 1e-12);
        const REAL cx = X / nc, cy = Y / nc, cz = Z / nc;
        const REAL n1 = (REAL)fmax(sqrt((double)(cx * cx + cy * cy + cz * cz)), 1e-5);
        int any_depth = 0, any_bounds = 0;
        for (int k = 0; k < K; ++k) {
          const REAL* pose = poses + ((long)b * K + k) * 16;
          const proj_t pr = project(Pall + ((long)b * K + k) * 16, X, Y, Z);
          sample(src + ((long)b * K + k) * C * HW, C, H, W, pr.px, pr.py, f + k * C);
          const REAL m = pr.zp > 0 ? (REAL)1 : (REAL)0;
          REAL dot = 0;
          for (int c = 0; c < C; ++c) dot += f[k * C + c] * cur[((long)b * C + c) * HW + p];
          const REAL sx0 = X - pose[3], sy0 = Y - pose[7], sz0 = Z - pose[11];
          const REAL ns = (REAL)fmax(sqrt((double)(sx0 * sx0 + sy0 * sy0 + sz0 * sz0)), 1e-12);
          const REAL sx = sx0 / ns, sy = sy0 / ns, sz = sz0 / ns;
          const REAL n2 = (REAL)fmax(sqrt((double)(sx * sx + sy * sy + sz * sz)), 1e-5);
          const REAL tr = pose[0] + pose[5] + pose[10];
          const REAL rm = (REAL)sqrt((double)((REAL)2 * ((REAL)1 - (tr < 3 ? tr : (REAL)3) / (REAL)3)));
          const REAL tm = (REAL)sqrt((double)(pose[3] * pose[3] + pose[7] * pose[7] + pose[11] * pose[11]));
          f[o_mask + k] = m;
          f[o_z + k] = pr.zp;
          f[o_dot + k] = dot * m;
          f[o_ang + k] = (cx / n1) * (sx / n2) + (cy / n1) * (sy / n2) + (cz / n1) * (sz / n2);
          f[o_nsrc + 3 * k] = sx; f[o_nsrc + 3 * k + 1] = sy; f[o_nsrc + 3 * k + 2] = sz;
          f[o_comb + k] = (REAL)sqrt((double)(tm * tm + rm * rm));
          f[o_r + k] = rm;
          f[o_t + k] = tm;
          if (pr.zp > 0) any_depth = 1;
          if (pr.px > 2 && pr.px < (REAL)(W - 2) && pr.py > 2 && pr.py < (REAL)(H - 2)) any_bounds = 1;
        }
        for (int c = 0; c < C; ++c) f[o_cur + c] = cur[((long)b * C + c) * HW + p];
        f[o_depth] = dv;
        f[o_ncur] = cx; f[o_ncur + 1] = cy; f[o_ncur + 2] = cz;
        for (int n = 0; n < H1; ++n) {
          REAL a = b1[n];
          for (int i = 0; i < F; ++i) a += w1[(long)n * F + i] * f[i];
          h1[n] = leaky(a);
        }
        for (int n = 0; n < H2; ++n) {
          REAL a = b2[n];
          for (int i = 0; i < H1; ++i) a += w2[(long)n * H1 + i] * h1[i];
          h2[n] = leaky(a);
        }
        REAL out = b3[0];
        for (int i = 0; i < H2; ++i) out += w3[i] * h2[i];
        cost[((long)b * D + d) * HW + p] = out;
        if (d == 0 || out > best || (out != out && best == best)) { best = out; best_d = dv; }
        if (mask && d == D - 1) mask[(long)b * HW + p] = (uint8_t)(any_depth && any_bounds);
      }
      if (lowest) lowest[(long)b * HW + p] = best_d;
      free(f);
    }
  free(Pall);
}
