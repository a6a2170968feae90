// Multi-view depth-consistency fusion — the 3DVNet-style point-cloud fuser of the reference's
// pc_fusion.py (reference tools/torch_point_cloud_fusion.py:12-97, process_depth).
//
// For every pixel of a reference depth map: un-project, re-project into every OTHER frame of the
// scan, nearest-sample that frame's depth, count the frames whose depth agrees within z_thresh,
// back-project the sampled depths and average the consistent points.  The reference materialises
// (n_src, 3, H*W) tensors several times over (bmm, grid_sample, masks) per batch of 100 sources;
// here one thread walks all sources of its pixel in registers: the only memory traffic is the
// nearest depth sample per (pixel, source) and 17 output bytes per pixel.
//
// Same geometry front end as the sweeps (un-project -> rigid transform -> project -> sample), in
// the reference's fp32 operation order: matrix products as k-ascending FMA chains, a separate add
// for the translation, true divisions, align_corners=True un-normalisation of grid_sample.
#include "srcv_kernels.h"

namespace srcv {

namespace {

struct MvsFrame {      // per frame, staged by the prep kernel: 3x3 K, K^-1, R, t of world->camera
  float K[9], Kinv[9], R[9], t[3];
};

__global__ void mvs_prep_kernel(const float* __restrict__ K, const float* __restrict__ Kinv,
                                const float* __restrict__ P, int n, MvsFrame* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  MvsFrame f;
#pragma unroll
  for (int j = 0; j < 9; ++j) { f.K[j] = K[i * 9 + j]; f.Kinv[j] = Kinv[i * 9 + j]; }
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int c = 0; c < 3; ++c) f.R[r * 3 + c] = P[i * 16 + r * 4 + c];
    f.t[r] = P[i * 16 + r * 4 + 3];
  }
  out[i] = f;
}

// y = M x as the k-ascending FMA chain of a fp32 GEMM
__device__ __forceinline__ void mat3(const float* __restrict__ M, float x, float y, float z, float& ox, float& oy,
                                     float& oz) {
  ox = __fmaf_rn(M[2], z, __fmaf_rn(M[1], y, __fmul_rn(M[0], x)));
  oy = __fmaf_rn(M[5], z, __fmaf_rn(M[4], y, __fmul_rn(M[3], x)));
  oz = __fmaf_rn(M[8], z, __fmaf_rn(M[7], y, __fmul_rn(M[6], x)));
}
// y = M^T x
__device__ __forceinline__ void mat3t(const float* __restrict__ M, float x, float y, float z, float& ox, float& oy,
                                      float& oz) {
  ox = __fmaf_rn(M[6], z, __fmaf_rn(M[3], y, __fmul_rn(M[0], x)));
  oy = __fmaf_rn(M[7], z, __fmaf_rn(M[4], y, __fmul_rn(M[1], x)));
  oz = __fmaf_rn(M[8], z, __fmaf_rn(M[5], y, __fmul_rn(M[2], x)));
}

__global__ void __launch_bounds__(128)
mvs_consistency_kernel(const float* __restrict__ depths, const MvsFrame* __restrict__ frames,
                       const float* __restrict__ ref_Pinv, int n, int H, int W, int ref, float z_thresh,
                       int n_consistent, float* __restrict__ pts_avg, int* __restrict__ n_valid_out,
                       uint8_t* __restrict__ valid_out) {
  SRCV_DYNAMIC_SMEM(float, s_frames_raw);
  MvsFrame* sf = reinterpret_cast<MvsFrame*>(s_frames_raw);
  const int HW = H * W;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = p < HW;
  const int px = live ? p % W : 0, py = live ? p / W : 0;
  // reference point: P_ref^-1[:3,:3] (K_ref^-1 (x d, y d, d)) + P_ref^-1[:3,3]   (:33-35)
  const MvsFrame& rf = frames[ref];
  const float d = live ? __ldg(depths + (size_t)ref * HW + p) : 0.f;
  float cx, cy, cz, X, Y, Z;
  mat3(rf.Kinv, __fmul_rn((float)px, d), __fmul_rn((float)py, d), d, cx, cy, cz);
  {
    float rx, ry, rz;
    const float Ri[9] = {ref_Pinv[0], ref_Pinv[1], ref_Pinv[2], ref_Pinv[4], ref_Pinv[5], ref_Pinv[6],
                         ref_Pinv[8], ref_Pinv[9], ref_Pinv[10]};
    mat3(Ri, cx, cy, cz, rx, ry, rz);
    X = __fadd_rn(rx, ref_Pinv[3]); Y = __fadd_rn(ry, ref_Pinv[7]); Z = __fadd_rn(rz, ref_Pinv[11]);
  }
  float ax = X, ay = Y, az = Z;
  int nv = 0;
  const float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
  constexpr int kChunk = 32;            // source frames staged in shared memory at a time
  for (int i0 = 0; i0 < n; i0 += kChunk) {
    const int cnt = min(kChunk, n - i0);
    __syncthreads();
    for (int j = threadIdx.x; j < cnt * (int)(sizeof(MvsFrame) / 4); j += blockDim.x)
      s_frames_raw[j] = reinterpret_cast<const float*>(frames + i0)[j];
    __syncthreads();
    if (!live) continue;
    for (int j = 0; j < cnt; ++j) {
      const int i = i0 + j;
      if (i == ref) continue;
      const MvsFrame& f = sf[j];
      // re-project: K (R X + t), perspective divide of all three rows (:51-55)
      float qx, qy, qz, u, v, z;
      mat3(f.R, X, Y, Z, qx, qy, qz);
      qx = __fadd_rn(qx, f.t[0]); qy = __fadd_rn(qy, f.t[1]); qz = __fadd_rn(qz, f.t[2]);
      mat3(f.K, qx, qy, qz, u, v, z);
      const float un = __fdiv_rn(u, z), vn = __fdiv_rn(v, z), on = __fdiv_rn(z, z);
      // nearest sample, align_corners=True: g = u / (w-1) * 2 - 1;  i = ((g + 1) / 2) (w-1)   (:61-66)
      const float gx = __fadd_rn(__fmul_rn(__fdiv_rn(un, wm1), 2.0f), -1.0f);
      const float gy = __fadd_rn(__fmul_rn(__fdiv_rn(vn, hm1), 2.0f), -1.0f);
      const float sx = rintf(__fmul_rn(__fdiv_rn(__fadd_rn(gx, 1.0f), 2.0f), wm1));
      const float sy = rintf(__fmul_rn(__fdiv_rn(__fadd_rn(gy, 1.0f), 2.0f), hm1));
      float zs = 0.f;
      if (sx >= 0.f && sx <= wm1 && sy >= 0.f && sy <= hm1)
        zs = __ldg(depths + (size_t)i * HW + (int)sy * W + (int)sx);
      const bool ok = (fabsf(__fadd_rn(z, -zs)) < z_thresh) && (un >= 0.f) && (un <= wm1) && (vn >= 0.f) &&
                      (vn <= hm1) && (z > 1e-4f);                       // :68-71
      if (ok) ++nv;
      // back-project the SAMPLED depth: R^T (K^-1 (u zs, v zs, 1 zs) - t)   (:75-77)
      float bx, by, bz, wx, wy, wz;
      mat3(f.Kinv, __fmul_rn(un, zs), __fmul_rn(vn, zs), __fmul_rn(on, zs), bx, by, bz);
      mat3t(f.R, __fadd_rn(bx, -f.t[0]), __fadd_rn(by, -f.t[1]), __fadd_rn(bz, -f.t[2]), wx, wy, wz);
      const bool nan = (wx != wx) || (wy != wy) || (wz != wz);          // :88-91
      if (ok && !nan) { ax = __fadd_rn(ax, wx); ay = __fadd_rn(ay, wy); az = __fadd_rn(az, wz); }
    }
  }
  if (!live) return;
  const float den = (float)(nv + 1);
  pts_avg[(size_t)p * 3 + 0] = __fdiv_rn(ax, den);
  pts_avg[(size_t)p * 3 + 1] = __fdiv_rn(ay, den);
  pts_avg[(size_t)p * 3 + 2] = __fdiv_rn(az, den);
  n_valid_out[p] = nv;
  valid_out[p] = nv >= n_consistent ? 1 : 0;
}

}  // namespace

size_t mvs_workspace_bytes(int n) { return sizeof(MvsFrame) * (size_t)n + 256; }

cudaError_t launch_mvs_consistency(const srcv_mvs_scan& s, int ref, float z_thresh, int n_consistent,
                                   float* pts_avg, int* n_valid, uint8_t* valid, void* workspace,
                                   bool frames_ready, cudaStream_t stream) {
  MvsFrame* frames = reinterpret_cast<MvsFrame*>(workspace);
  if (!frames_ready) {
    SRCV_LAUNCH(mvs_prep_kernel, (s.N + 127) / 128, 128, 0, stream, s.K, s.K_inv, s.cam_T_world, s.N, frames);
    note_launch();
  }
  const int HW = s.H * s.W;
  const size_t smem = sizeof(MvsFrame) * 32;
  SRCV_LAUNCH(mvs_consistency_kernel, (HW + 127) / 128, 128, smem, stream, s.depths, frames,
              s.world_T_cam + (size_t)ref * 16, s.N, s.H, s.W, ref, z_thresh, n_consistent, pts_avg, n_valid, valid);
  note_launch();
  return cudaGetLastError();
}

}  // namespace srcv
