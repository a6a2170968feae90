// Dot-product plane-sweep volume — replaces CostVolumeManager.build_cost_volume
// (+ the argmax of CostVolumeManager.forward), reference
// modules/cost_volume.py:237-335, :345-380.
//
//   cost[b,d,p] = sum_k [z'_k > 0] * sum_c cur[b,c,p] * bilinear(src[b,k,c], proj_k(d, p))
//
// The (B,K,D,C,H,W) warped tensor of the reference is never formed: every
// (plane, view, pixel) sample is projected, gathered and reduced in registers.
//
// Two variants:
//  * generic — any K, C; one thread per pixel, planar (NCHW) scalar gathers.
//  * fast    — C == 16; gathers from the channel-last copy made by the prep pass.
//    A warp owns an 8x4 block of pixels (2-D so the bilinear footprints of its
//    pixels overlap in both directions).  Lane l does the projection of pixel
//    (l%8, l/8) ("owner" role); for the gather the warp re-partitions itself into 8
//    groups of 4 lanes: in round r group g serves pixel (g, r) and lane j of the
//    group loads the 16-byte channel chunk j of each tap, so one warp-wide LDG.128
//    moves 8 whole 64-byte texels that are contiguous in memory whenever the 8
//    pixels' taps are (they are for the near-translational homographies of a
//    keyframe sweep): full 128 B/clk L1 wavefronts instead of 32 scattered 16-byte
//    pieces.  The sample's footprint travels owner->group by three shuffles and
//    the partial dots come back the same way.  Loop order is
//    plane-chunk -> view -> plane-in-chunk, so consecutive samples walk along one
//    epipolar line of one view and re-hit the lines they just pulled into L1.
#include "srcv_kernels.h"

namespace srcv {

namespace {

// --------------------------------------------------------------------------- //
// generic                                                                     //
// --------------------------------------------------------------------------- //
template <bool PER_PIXEL>
__global__ void __launch_bounds__(128)
dot_generic_kernel(srcv_shape s, const float* __restrict__ cur, const float* __restrict__ src,
                   const ViewParams* __restrict__ views, const float* __restrict__ planes,
                   float* __restrict__ cost, float* __restrict__ lowest) {
  extern __shared__ float sview[];  // K * 12: a0, hx, hy, t
  const int b = blockIdx.y;
  const int HW = s.H * s.W;
  for (int i = threadIdx.x; i < s.K * kViewFloats; i += blockDim.x)
    sview[i] = reinterpret_cast<const float*>(views + b * s.K + i / kViewFloats)[i % kViewFloats];
  __syncthreads();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  const Centre ctr(s.W, s.H);
  const float dx = ((float)(p % s.W) + 0.5f) - ctr.half_w, dy = ((float)(p / s.W) + 0.5f) - ctr.half_h;
  const float* curp = cur + (size_t)b * s.C * HW + p;
  float best = 0.f, best_d = 0.f;
  for (int d = 0; d < s.D; ++d) {
    const float dval = PER_PIXEL ? __ldg(planes + ((size_t)b * s.D + d) * HW + p)
                                 : __ldg(planes + b * s.D + d);
    float acc = 0.f;
    for (int k = 0; k < s.K; ++k) {
      const float* vp = sview + k * kViewFloats;
      float ax, ay, az, px, py, zp;
      homography_point(vp, dx, dy, ax, ay, az);
      project_point(dval, ax, ay, az, vp[9], vp[10], vp[11], px, py, zp);
      Taps tp;
      bilinear_taps(px, py, s.W, s.H, ctr, tp);
      if (!(zp > 0.0f) || tp.valid == 0u) continue;  // mask == 0 or all taps padded
      const float w00 = (1.0f - tp.fx) * (1.0f - tp.fy), w01 = tp.fx * (1.0f - tp.fy);
      const float w10 = (1.0f - tp.fx) * tp.fy, w11 = tp.fx * tp.fy;
      const float* sp = src + ((size_t)(b * s.K + k) * s.C) * HW + (tp.y0 * s.W + tp.x0);
      float dot = 0.f;
      for (int c = 0; c < s.C; ++c) {
        const float* q = sp + (size_t)c * HW;
        float v = 0.f;
        if (tp.valid & 1u) v = w00 * __ldg(q);
        if (tp.valid & 2u) v = fmaf(w01, __ldg(q + 1), v);
        if (tp.valid & 4u) v = fmaf(w10, __ldg(q + s.W), v);
        if (tp.valid & 8u) v = fmaf(w11, __ldg(q + s.W + 1), v);
        dot = fmaf(v, __ldg(curp + (size_t)c * HW), dot);
      }
      acc += dot;
    }
    cost[((size_t)b * s.D + d) * HW + p] = acc;
    argmax_update(acc, dval, best, best_d, d == 0);
  }
  if (lowest) lowest[(size_t)b * HW + p] = best_d;
}

// --------------------------------------------------------------------------- //
// fast: C = 16, channel-last gathers, 4 lanes per texel                       //
// --------------------------------------------------------------------------- //
constexpr int kFastC = 16;
constexpr int kTileW = 8, kTileH = 4;   // pixels per warp
constexpr int kFastWarps = 2;           // CTA = 8 x 8 pixels
constexpr int kDC = 4;                  // planes per inner chunk
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ float dot4(const float4& v, const float4& c) {
  return fmaf(v.x, c.x, fmaf(v.y, c.y, fmaf(v.z, c.z, v.w * c.w)));
}

template <bool PER_PIXEL>
__global__ void __launch_bounds__(kFastWarps * 32)
dot_fast_kernel(srcv_shape s, const float* __restrict__ cur, const float4* __restrict__ src4,
                const ViewParams* __restrict__ views, const float* __restrict__ planes,
                float* __restrict__ cost, float* __restrict__ lowest) {
  extern __shared__ float sview[];  // K * 12
  const int b = blockIdx.y;
  const int W = s.W, H = s.H, HW = W * H, K = s.K;
  for (int i = threadIdx.x; i < K * kViewFloats; i += blockDim.x)
    sview[i] = reinterpret_cast<const float*>(views + b * K + i / kViewFloats)[i % kViewFloats];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int tiles_x = (W + kTileW - 1) / kTileW;
  const int x_base = (blockIdx.x % tiles_x) * kTileW;
  const int y_base = ((blockIdx.x / tiles_x) * kFastWarps + (threadIdx.x >> 5)) * kTileH;
  if (y_base >= H) return;  // warp-uniform
  // planes handled by this CTA: [d_begin, d_end)
  const int dper = (s.D + gridDim.z - 1) / gridDim.z;
  const int d_begin = blockIdx.z * dper;
  const int d_end = min(s.D, d_begin + dper);
  const bool fuse_argmax = (gridDim.z == 1) && (lowest != nullptr);
  const Centre ctr(W, H);

  // ---- owner role: pixel (lane % 8, lane / 8) of the block -------------------
  const int ox = x_base + (lane & 7), oy = y_base + (lane >> 3);
  const bool active = ox < W && oy < H;
  const int p = min(oy, H - 1) * W + min(ox, W - 1);
  const float dx = ((float)min(ox, W - 1) + 0.5f) - ctr.half_w;
  const float dy = ((float)min(oy, H - 1) + 0.5f) - ctr.half_h;
  // ---- helper role: channel chunk j of pixel (g, r) in round r ----------------
  const int g = lane >> 2, j = lane & 3;
  float4 cur4[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int q = min(y_base + r, H - 1) * W + min(x_base + g, W - 1);
    const float* cp = cur + ((size_t)b * kFastC + 4 * j) * HW + q;
    cur4[r] = make_float4(__ldg(cp), __ldg(cp + HW), __ldg(cp + 2 * (size_t)HW), __ldg(cp + 3 * (size_t)HW));
  }
  float best = 0.f, best_d = 0.f;

  for (int d0 = d_begin; d0 < d_end; d0 += kDC) {
    float dval[kDC];
#pragma unroll
    for (int dd = 0; dd < kDC; ++dd) {
      const int d = min(d0 + dd, d_end - 1);
      dval[dd] = PER_PIXEL ? __ldg(planes + ((size_t)b * s.D + d) * HW + p)
                           : __ldg(planes + b * s.D + d);
    }
    float acc[kDC][4];
#pragma unroll
    for (int dd = 0; dd < kDC; ++dd)
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[dd][r] = 0.f;

#pragma unroll 1
    for (int k = 0; k < K; ++k) {
      const float* vp = sview + k * kViewFloats;
      float ax, ay, az;
      homography_point(vp, dx, dy, ax, ay, az);
      const float tx = vp[9], ty = vp[10], tz = vp[11];
      int packed[kDC];
      float fx[kDC], fy[kDC];
#pragma unroll
      for (int dd = 0; dd < kDC; ++dd) {
        float px, py, zp;
        project_point(dval[dd], ax, ay, az, tx, ty, tz, px, py, zp);
        Taps tp;
        bilinear_taps(px, py, W, H, ctr, tp);
        const unsigned valid = (zp > 0.0f) ? tp.valid : 0u;  // depth mask folded into the footprint
        packed[dd] = (tp.y0 * W + tp.x0) * 16 + (int)valid;
        fx[dd] = tp.fx;
        fy[dd] = tp.fy;
      }
      const float4* view4 = src4 + (size_t)(b * K + k) * HW * 4 + j;
#pragma unroll
      for (int dd = 0; dd < kDC; ++dd) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int src_lane = 8 * r + g;
          const int pk = __shfl_sync(kFull, packed[dd], src_lane);
          const float hfx = __shfl_sync(kFull, fx[dd], src_lane);
          const float hfy = __shfl_sync(kFull, fy[dd], src_lane);
          const float gx = 1.0f - hfx, gy = 1.0f - hfy;
          const float4* t4 = view4 + (ptrdiff_t)(pk >> 4) * 4;
          const float4 c4 = cur4[r];
          float a = acc[dd][r];
          if (pk & 1) a = fmaf(gx * gy, dot4(__ldg(t4), c4), a);
          if (pk & 2) a = fmaf(hfx * gy, dot4(__ldg(t4 + 4), c4), a);
          if (pk & 4) a = fmaf(gx * hfy, dot4(__ldg(t4 + 4 * W), c4), a);
          if (pk & 8) a = fmaf(hfx * hfy, dot4(__ldg(t4 + 4 * W + 4), c4), a);
          acc[dd][r] = a;
        }
      }
    }
    // reduce the 4 channel chunks of each group, then hand pixel (g, r)'s sum to its owner
#pragma unroll
    for (int dd = 0; dd < kDC; ++dd) {
      float mine = 0.f;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        float a = acc[dd][r];
        a += __shfl_xor_sync(kFull, a, 1);
        a += __shfl_xor_sync(kFull, a, 2);
        const float t = __shfl_sync(kFull, a, 4 * (lane & 7));
        if ((lane >> 3) == r) mine = t;
      }
      const int d = d0 + dd;
      if (d < d_end) {
        if (active) cost[((size_t)b * s.D + d) * HW + p] = mine;
        argmax_update(mine, dval[dd], best, best_d, d == d_begin);
      }
    }
  }
  if (fuse_argmax && active) lowest[(size_t)b * HW + p] = best_d;
}

}  // namespace

cudaError_t launch_dot_generic(const srcv_shape& s, const float* cur, const float* src,
                               const Workspace& ws, const float* planes, bool per_pixel,
                               float* cost, float* lowest, cudaStream_t stream) {
  const int HW = s.H * s.W;
  dim3 grid((HW + 127) / 128, s.B), block(128);
  const size_t smem = sizeof(float) * kViewFloats * s.K;
  if (per_pixel)
    dot_generic_kernel<true><<<grid, block, smem, stream>>>(s, cur, src, ws.views, planes, cost, lowest);
  else
    dot_generic_kernel<false><<<grid, block, smem, stream>>>(s, cur, src, ws.views, planes, cost, lowest);
  note_launch();
  return cudaGetLastError();
}

bool dot_fast_supported(const srcv_shape& s) {
  // (texel index * 16) must fit an int; 16-byte texel chunks need C == 16
  return s.C == kFastC && (long long)s.H * s.W < (1ll << 26) && s.K <= 512;
}

cudaError_t launch_dot_fast(const srcv_shape& s, const float* cur, const Workspace& ws,
                            const float* planes, bool per_pixel, float* cost, float* lowest,
                            cudaStream_t stream) {
  // Split the plane loop across CTAs only when the batch alone cannot fill the
  // machine (about 16 warps per SM wanted); a split sweep cannot fuse the argmax.
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int tiles_x = (s.W + kTileW - 1) / kTileW;
  const int tiles_y = (s.H + kTileH * kFastWarps - 1) / (kTileH * kFastWarps);
  const long long warps = (long long)s.B * tiles_x * tiles_y * kFastWarps;
  int d_split = 1;
  while (d_split < 8 && warps * d_split < 16ll * sms && s.D / (d_split * 2) >= 2 * kDC) d_split *= 2;
  dim3 grid(tiles_x * tiles_y, s.B, d_split), block(kFastWarps * 32);
  const size_t smem = sizeof(float) * kViewFloats * s.K;
  const float4* src4 = reinterpret_cast<const float4*>(ws.src_nhwc);
  if (per_pixel)
    dot_fast_kernel<true><<<grid, block, smem, stream>>>(s, cur, src4, ws.views, planes, cost, lowest);
  else
    dot_fast_kernel<false><<<grid, block, smem, stream>>>(s, cur, src4, ws.views, planes, cost, lowest);
  note_launch();
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return err;
  if (d_split > 1 && lowest) err = launch_argmax(s, cost, planes, per_pixel, lowest, stream);
  return err;
}

}  // namespace srcv
