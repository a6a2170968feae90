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
//  * fast    — C == 16, K <= 8; gathers from the channel-last copy made by the
//    prep pass.  A warp owns 32 consecutive pixels.  Lane l does the projection of
//    pixel l ("owner" role); for the gather the warp re-partitions itself into 8
//    groups of 4 lanes: in round r group g serves pixel 8r+g and lane j of the
//    group loads the 16-byte channel chunk j of each tap, so one warp-wide
//    LDG.128 moves 8 whole 64-byte texels that are contiguous in memory whenever
//    the 8 pixels' taps are (they are for the near-translational homographies of
//    a keyframe sweep).  That is the access shape that saturates the 128 B/clk
//    L1 path; the sample's footprint/weights travel owner->group by shuffles and
//    the partial dots come back the same way.
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
  extern __shared__ float sview[];  // K * 12: Hm(9), t(3)
  const int b = blockIdx.y;
  const int HW = s.H * s.W;
  for (int i = threadIdx.x; i < s.K * 12; i += blockDim.x) {
    const int k = i / 12, j = i - k * 12;
    const ViewParams& vp = views[b * s.K + k];
    sview[i] = j < 9 ? vp.Hm[j] : vp.t[j - 9];
  }
  __syncthreads();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  const float pxc = (float)(p % s.W) + 0.5f, pyc = (float)(p / s.W) + 0.5f;
  const float inv_w = 1.0f / (float)s.W, inv_h = 1.0f / (float)s.H;
  const float* curp = cur + (size_t)b * s.C * HW + p;
  float best = 0.f, best_d = 0.f;
  for (int d = 0; d < s.D; ++d) {
    const float dval = PER_PIXEL ? __ldg(planes + ((size_t)b * s.D + d) * HW + p)
                                 : __ldg(planes + b * s.D + d);
    float acc = 0.f;
    for (int k = 0; k < s.K; ++k) {
      const float* h = sview + k * 12;
      const float ax = fmaf(h[0], pxc, fmaf(h[1], pyc, h[2]));
      const float ay = fmaf(h[3], pxc, fmaf(h[4], pyc, h[5]));
      const float az = fmaf(h[6], pxc, fmaf(h[7], pyc, h[8]));
      float px, py, zp;
      project_point(dval, ax, ay, az, h[9], h[10], h[11], px, py, zp);
      Taps tp;
      bilinear_taps(px, py, s.W, s.H, inv_w, inv_h, tp);
      if (!(zp > 0.0f) || tp.valid == 0u) continue;  // mask == 0 or all taps padded
      const float* sp = src + ((size_t)(b * s.K + k) * s.C) * HW + (tp.y0 * s.W + tp.x0);
      float dot = 0.f;
      for (int c = 0; c < s.C; ++c) {
        const float* q = sp + (size_t)c * HW;
        float v = 0.f;
        if (tp.valid & 1u) v = tp.w[0] * __ldg(q);
        if (tp.valid & 2u) v = fmaf(tp.w[1], __ldg(q + 1), v);
        if (tp.valid & 4u) v = fmaf(tp.w[2], __ldg(q + s.W), v);
        if (tp.valid & 8u) v = fmaf(tp.w[3], __ldg(q + s.W + 1), v);
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
constexpr int kFastWarps = 2;
constexpr unsigned kFull = 0xffffffffu;

template <int K, bool PER_PIXEL>
__global__ void __launch_bounds__(kFastWarps * 32)
dot_fast_kernel(srcv_shape s, const float* __restrict__ cur, const float4* __restrict__ src4,
                const ViewParams* __restrict__ views, const float* __restrict__ planes,
                float* __restrict__ cost, float* __restrict__ lowest) {
  __shared__ float sview[K * 12];
  const int b = blockIdx.y;
  const int HW = s.H * s.W;
  for (int i = threadIdx.x; i < K * 12; i += blockDim.x) {
    const int k = i / 12, j = i - k * 12;
    const ViewParams& vp = views[b * K + k];
    sview[i] = j < 9 ? vp.Hm[j] : vp.t[j - 9];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warp_base = (blockIdx.x * kFastWarps + (threadIdx.x >> 5)) * 32;
  if (warp_base >= HW) return;  // whole warp out of range (warp-uniform)
  // planes handled by this CTA: [d_begin, d_end)
  const int dper = (s.D + gridDim.z - 1) / gridDim.z;
  const int d_begin = blockIdx.z * dper;
  const int d_end = min(s.D, d_begin + dper);
  const bool fuse_argmax = (gridDim.z == 1) && (lowest != nullptr);

  // ---- owner role: pixel `p`, depth-invariant a_k = Hm_k p ------------------
  const int p_raw = warp_base + lane;
  const bool active = p_raw < HW;
  const int p = active ? p_raw : HW - 1;
  const float pxc = (float)(p % s.W) + 0.5f, pyc = (float)(p / s.W) + 0.5f;
  float ax[K], ay[K], az[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const float* h = sview + k * 12;
    ax[k] = fmaf(h[0], pxc, fmaf(h[1], pyc, h[2]));
    ay[k] = fmaf(h[3], pxc, fmaf(h[4], pyc, h[5]));
    az[k] = fmaf(h[6], pxc, fmaf(h[7], pyc, h[8]));
  }
  // ---- helper role: chunk j of pixels 8r + g --------------------------------
  const int g = lane >> 2, j = lane & 3;
  float4 cur4[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int q = min(warp_base + 8 * r + g, HW - 1);
    const float* cp = cur + ((size_t)b * kFastC + 4 * j) * HW + q;
    cur4[r] = make_float4(__ldg(cp), __ldg(cp + HW), __ldg(cp + 2 * (size_t)HW), __ldg(cp + 3 * (size_t)HW));
  }
  const float inv_w = 1.0f / (float)s.W, inv_h = 1.0f / (float)s.H;
  const int W = s.W;
  float best = 0.f, best_d = 0.f;

  for (int d = d_begin; d < d_end; ++d) {
    const float dval = PER_PIXEL ? __ldg(planes + ((size_t)b * s.D + d) * HW + p)
                                 : __ldg(planes + b * s.D + d);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < K; ++k) {
      float px, py, zp;
      project_point(dval, ax[k], ay[k], az[k], sview[k * 12 + 9], sview[k * 12 + 10],
                    sview[k * 12 + 11], px, py, zp);
      Taps tp;
      bilinear_taps(px, py, W, s.H, inv_w, inv_h, tp);
      const unsigned valid = (zp > 0.0f) ? tp.valid : 0u;  // depth mask folded into the footprint
      const int packed = (tp.y0 * W + tp.x0) * 16 + (int)valid;
      const float4* view4 = src4 + (size_t)(b * K + k) * HW * 4 + j;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int src_lane = 8 * r + g;
        const int pk = __shfl_sync(kFull, packed, src_lane);
        const float w0 = __shfl_sync(kFull, tp.w[0], src_lane);
        const float w1 = __shfl_sync(kFull, tp.w[1], src_lane);
        const float w2 = __shfl_sync(kFull, tp.w[2], src_lane);
        const float w3 = __shfl_sync(kFull, tp.w[3], src_lane);
        const float4* t4 = view4 + (ptrdiff_t)(pk >> 4) * 4;
        const float4 c4 = cur4[r];
        float a = acc[r];
        if (pk & 1) { const float4 v = __ldg(t4);               a = fmaf(w0, fmaf(v.x, c4.x, fmaf(v.y, c4.y, fmaf(v.z, c4.z, v.w * c4.w))), a); }
        if (pk & 2) { const float4 v = __ldg(t4 + 4);           a = fmaf(w1, fmaf(v.x, c4.x, fmaf(v.y, c4.y, fmaf(v.z, c4.z, v.w * c4.w))), a); }
        if (pk & 4) { const float4 v = __ldg(t4 + 4 * W);       a = fmaf(w2, fmaf(v.x, c4.x, fmaf(v.y, c4.y, fmaf(v.z, c4.z, v.w * c4.w))), a); }
        if (pk & 8) { const float4 v = __ldg(t4 + 4 * W + 4);   a = fmaf(w3, fmaf(v.x, c4.x, fmaf(v.y, c4.y, fmaf(v.z, c4.z, v.w * c4.w))), a); }
        acc[r] = a;
      }
    }
    // reduce the 4 channel chunks of each group, then hand pixel 8r+g's sum to its owner
    float mine = 0.f;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      float a = acc[r];
      a += __shfl_xor_sync(kFull, a, 1);
      a += __shfl_xor_sync(kFull, a, 2);
      const float t = __shfl_sync(kFull, a, 4 * (lane & 7));
      if ((lane >> 3) == r) mine = t;
    }
    if (active) cost[((size_t)b * s.D + d) * HW + p] = mine;
    argmax_update(mine, dval, best, best_d, d == d_begin);
  }
  if (fuse_argmax && active) lowest[(size_t)b * HW + p] = best_d;
}

template <int K>
cudaError_t launch_fast_k(const srcv_shape& s, const float* cur, const Workspace& ws,
                          const float* planes, bool per_pixel, float* cost, float* lowest,
                          int d_split, cudaStream_t stream) {
  const int HW = s.H * s.W;
  dim3 grid((HW + kFastWarps * 32 - 1) / (kFastWarps * 32), s.B, d_split);
  dim3 block(kFastWarps * 32);
  const float4* src4 = reinterpret_cast<const float4*>(ws.src_nhwc);
  if (per_pixel)
    dot_fast_kernel<K, true><<<grid, block, 0, stream>>>(s, cur, src4, ws.views, planes, cost, lowest);
  else
    dot_fast_kernel<K, false><<<grid, block, 0, stream>>>(s, cur, src4, ws.views, planes, cost, lowest);
  note_launch();
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_dot_generic(const srcv_shape& s, const float* cur, const float* src,
                               const Workspace& ws, const float* planes, bool per_pixel,
                               float* cost, float* lowest, cudaStream_t stream) {
  const int HW = s.H * s.W;
  dim3 grid((HW + 127) / 128, s.B), block(128);
  const size_t smem = sizeof(float) * 12 * s.K;
  if (per_pixel)
    dot_generic_kernel<true><<<grid, block, smem, stream>>>(s, cur, src, ws.views, planes, cost, lowest);
  else
    dot_generic_kernel<false><<<grid, block, smem, stream>>>(s, cur, src, ws.views, planes, cost, lowest);
  note_launch();
  return cudaGetLastError();
}

bool dot_fast_supported(const srcv_shape& s) {
  // (texel index << 4) must fit an int; 16-byte texel chunks need C == 16
  return s.C == kFastC && s.K >= 1 && s.K <= 8 && (long long)s.H * s.W < (1ll << 26);
}

cudaError_t launch_dot_fast(const srcv_shape& s, const float* cur, const Workspace& ws,
                            const float* planes, bool per_pixel, float* cost, float* lowest,
                            cudaStream_t stream) {
  // Split the plane loop across CTAs only when the batch alone cannot fill the
  // machine (about 16 warps per SM wanted); a split sweep cannot fuse the argmax.
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long warps = (long long)s.B * ((s.H * s.W + 31) / 32);
  int d_split = 1;
  while (d_split < 8 && warps * d_split < 16ll * sms && s.D / (d_split * 2) >= 8) d_split *= 2;
  cudaError_t err;
  switch (s.K) {
#define SRCV_CASE(KK) \
  case KK: err = launch_fast_k<KK>(s, cur, ws, planes, per_pixel, cost, lowest, d_split, stream); break;
    SRCV_CASE(1) SRCV_CASE(2) SRCV_CASE(3) SRCV_CASE(4)
    SRCV_CASE(5) SRCV_CASE(6) SRCV_CASE(7) SRCV_CASE(8)
#undef SRCV_CASE
    default: return cudaErrorInvalidValue;
  }
  if (err != cudaSuccess) return err;
  if (d_split > 1 && lowest) err = launch_argmax(s, cost, planes, per_pixel, lowest, stream);
  return err;
}

}  // namespace srcv
