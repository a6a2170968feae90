// Metadata-MLP plane-sweep volume on the 5th-generation tensor cores (tcgen05).
//
// Replaces FeatureVolumeManager / FastFeatureVolumeManager.build_cost_volume
// (reference modules/cost_volume.py:451-736, :967-1164) for the hero layout
// K = 7 source views, C = 16 channels, MLP 202 -> 128 -> 128 -> 1.
//
// One persistent CTA per SM walks over row tiles; a tile = 128 rows = a 16 x 2 patch of
// pixels at four consecutive depth planes.  Warp roles (640 threads):
//   * 12 BUILDER warps, three threads per row: project, gather (chunk-planar copies of the
//     features, csrc/srcv_prep.cu), blend on packed fp32 pairs (FFMA2) and build the row's metadata
//     channels in registers, split every value into an fp16 (hi, lo) pair and write them straight
//     into TENSOR MEMORY as the A operand (tcgen05.st) — the (B,F,H,W) / (B*D,H,W,F) tensors the
//     reference materialises exist only as 192 TMEM columns per row.  Two whole views per thread,
//     then half of view 2 (slots 0, 1) or the tail (slot 2);
//   * 1 MMA-issuing thread (elect.sync): layer 1 = 12 x 3 tcgen05.mma (A from TMEM, weights from
//     shared memory, fp32 accumulator in TMEM):  A_hi W_hi + A_hi W_lo + A_lo W_hi, issued as two
//     64-column halves; the WHOLE layer-1 bias rides in the MMA — a constant-one K position against
//     a weight row holding 16 (b1 + the frame's pose-measure contribution), which this warp rewrites
//     in shared memory when its tiles move on to another frame; layer 2 = 8 x 3 tcgen05.mma in two
//     K halves, accumulating into the first 128 columns of the tile's own (by then dead) A1 buffer;
//   * 4 EPILOGUE warps, one thread per row: after each layer-1 half they read the accumulator,
//     apply LeakyReLU, split to fp16 (hi, lo) and write the layer-2 A operand back IN PLACE (a
//     chunk's 32 fp32 columns become its 16 hi + 16 lo operand columns, so nothing else is
//     touched); after layer 2 they apply bias + LeakyReLU and the 128 -> 1 layer as one dot
//     product per row (packed FFMA2 chains) and store the cost.
// Stages are chained by mbarriers (tcgen05.commit for MMA completion).  The roles overlap: with two
// A1 buffers the builders gather tile t+1 while the epilogue warps post-process tile t and the
// tensor pipe runs its MMAs; the layer-1 MMAs of tile t+1 start while the layer-2 epilogue of tile
// t is still running; within a tile the epilogue of one half overlaps the MMAs of the other.
// Weights (both layers, hi and lo, 164 KB) stay resident in shared memory for the CTA's lifetime,
// laid out as K-major no-swizzle core matrices by the pack kernel.
//
// The K order of layer 1 is OURS (the pack kernel permutes W1's columns to match), 8 blocks x 24:
//   per view k:  16 warped | mask | z' | dot | ray angle | n_src (3) | pad
//     (view 2, built by two threads:  8 warped | dot | mask | z' | angle || 8 warped | dot | n_src (3))
//   tail:        16 reference features | plane depth | n_cur (3) | ONE (bias) | pad
// The 21 pose measures (comb, r, t per view) are constant per (frame, view) and enter through the
// bias row instead of 21 K positions.
//
// Scaling: W1 and W2 are stored x16 (exact) so the lo halves of typical |w| ~ 0.05 weights
// stay out of the fp16 subnormals.  Layer 1 therefore yields 16 (W1 x + b1); LeakyReLU
// commutes with the positive scale, so the layer-2 operand is 16 a and its accumulator is
// 256 (W2 a): the single 1/256 is folded into the layer-2 bias FMA.
#include "srcv_kernels.h"
#include "srcv_tc.cuh"

namespace srcv {

namespace {

using namespace tc;

constexpr int kC = 16, kViews = 7;
constexpr int kRows = 128;               // rows (TMEM lanes) per tile
// A tile is a 16 x 2 pixel patch at FOUR consecutive depth planes (32 pixels x 4 planes =
// 128 rows): warp w of a lane group works plane w of the same patch, so the four warps'
// gathers for one source view walk along the same epipolar lines and share L1 lines.
constexpr int kTileW = 16, kTileH = 2, kTileD = 4;
constexpr int kN = 128;                  // layer widths
constexpr int kBlk = 24;                 // K positions per block (23 per view + 1 pad; tail 21 + 3 pad)
constexpr int kBlkCols = kBlk / 2;       // 12 packed columns per block and half
constexpr int kK1 = 192;                 // 8 blocks x 24  (12 k-steps of 16)
constexpr int kK2 = 128;
constexpr int kF = kC * (kViews + 1) + 10 * kViews + 4;  // 202
constexpr int kBiasPos = kViews * kBlk + kC + 4;         // K position of the constant one

// TMEM columns (32-bit): two K elements per column.  TWO A1 buffers (tile t uses buffer t & 1), so
// the builders run a whole tile ahead of the tensor pipe; the layer-2 accumulator D2 of tile t
// lands in the first 128 columns of tile t's own A1 buffer, which is dead once its layer-1 MMAs
// are done.   A1[0]: hi 0..95 | lo 96..191    A1[1]: 192..383    DA: 384..511
constexpr uint32_t kColA1 = 0, kA1Stride = kK1, kA1LoOff = kK1 / 2, kColDA = 2 * kK1, kTmemCols = 512;
static_assert(kColDA + kN <= kTmemCols, "TMEM budget");
static_assert(kN <= kK1, "D2 must fit in an A1 buffer");

constexpr float kWScale = 16.0f;
constexpr float kUnscale2 = 1.0f / (kWScale * kWScale);

// Warps 0-11 build, 12-15 run the epilogues, 16 issues the MMAs; 17-19 complete its warpgroup
// (setmaxnreg works on aligned groups of four warps): the kernel is launched at 96 registers per
// thread, the MMA group shrinks to 24, builders grow to 112 (all 16 vector loads of a bilinear
// footprint in flight) and the epilogue group to 120.
// setmaxnreg only REDISTRIBUTES the CTA's launch-time allocation (640 x 96 registers): an .inc
// blocks until the pool holds enough, so the totals below must fit it or the CTA deadlocks.
#ifdef SRCV_TC_W24
// 24 warps: SIXTEEN builder warps (four threads per row, two K blocks each: views 2s, 2s+1; slot 3 the
// last view and the tail), four epilogue warps, the MMA warp and its three idle group mates.  Registers
// are allocated per four warps, so 768 threads launch at 80; the MMA group's 56 x 128 freed registers go
// to the epilogue group (and, in the B split, to the builders).
constexpr int kBuildWarps = 16, kEpiWarps = 4, kMmaWarp = 20;
constexpr int kThreads = 24 * 32;
constexpr int kSlots = 4;
constexpr bool kSetmaxnreg = true;
constexpr int kRegsLaunch = 80;
#ifdef SRCV_TC_W24B
constexpr int kRegsBuild = 88, kRegsEpi = 104, kRegsMma = 24;
constexpr int kEpiStep = 32;
#else
constexpr int kRegsBuild = 80, kRegsEpi = 120, kRegsMma = 24;
constexpr int kEpiStep = 64;
#endif
constexpr int kSplitView = -1;           // no view is split: 2 / 2 / 2 / 1 + tail blocks per slot
#else
constexpr int kBuildWarps = 12, kEpiWarps = 4, kMmaWarp = 16;
constexpr int kThreads = 20 * 32;
constexpr int kSlots = 3;
#ifdef SRCV_TC_NO_SETMAXNREG
constexpr bool kSetmaxnreg = false;
#else
constexpr bool kSetmaxnreg = true;
#endif
constexpr int kRegsLaunch = 96;          // what ptxas allocates under __launch_bounds__(640, 1)
constexpr int kRegsBuild = 112, kRegsEpi = 120, kRegsMma = 24;
constexpr int kEpiStep = 64;
// The 7 view blocks + the tail do not divide by the three builder threads of a row: view kSplitView is
// built as TWO half-view units (channels 8h..8h+7 each, see unit_issue) by slots 0 and 1, so the slots
// carry 2.5 / 2.5 / 2 views + tail instead of 3 / 3 / 1 + tail.
constexpr int kSplitView = 2;
#endif
constexpr int kBuilders = kBuildWarps * 32, kEpis = kEpiWarps * 32;
static_assert(!kSetmaxnreg || kBuilders * kRegsBuild + kEpis * kRegsEpi + (kThreads - kBuilders - kEpis) * kRegsMma <= kThreads * kRegsLaunch,
              "setmaxnreg budget: increases must be covered by the decreases within the CTA's launch allocation");
static_assert(kRegsBuild % 8 == 0 && kRegsEpi % 8 == 0 && kRegsMma % 8 == 0, "setmaxnreg takes multiples of 8");

// shared memory image (bytes)
constexpr uint32_t kW1Bytes = kN * kK1 * 2, kW2Bytes = kN * kK2 * 2;   // one of (hi, lo)
constexpr uint32_t kOffW1Hi = 0, kOffW1Lo = kW1Bytes, kOffW2Hi = 2 * kW1Bytes,
                   kOffW2Lo = 2 * kW1Bytes + kW2Bytes, kOffVec = 2 * kW1Bytes + 2 * kW2Bytes;
constexpr uint32_t kVecFloats = 3 * kN + 4;   // b2 | 0.505 w3 | 0.495 w3 | b3
constexpr uint32_t kOffBar = kOffVec + kVecFloats * 4;
constexpr uint32_t kOffFlag = kOffBar + 12 * 8;           // 12 mbarrier slots (10 used)
constexpr uint32_t kSmemBytes = kOffFlag + 2 * kSlots * kRows;   // mask bits [parity][builder slot][row]
// image = [W1hi | W1lo | W2hi | W2lo | vec] exactly as it sits in shared memory
constexpr uint32_t kImageBytes = kOffBar;

// K-major no-swizzle core-matrix offset (in halves) of element (n, k) of an N x Kp operand
__host__ __device__ inline uint32_t core_offset(int n, int k, int N) {
  return (uint32_t)(k >> 3) * (uint32_t)(N * 8) + (uint32_t)(n >> 3) * 64u + (uint32_t)(n & 7) * 8u + (uint32_t)(k & 7);
}

// reference channel index (modules/cost_volume.py:698-723 order) of OUR layer-1 K position;
// -2 = the bias position (constant one in A), -1 = padding.  The 21 pose measures (comb, r, t per
// view) are NOT in K: they are constant per (frame, view), so their layer-1 contribution is a
// per-frame bias vector (tc_frame_bias_kernel) added in the layer-1 epilogue.
__host__ __device__ inline int ref_channel(int kk) {
  if (kSplitView >= 0 && kk / kBlk == kSplitView) {
    // the split view's block is two units of 12 K positions, unit h =
    //   8 warped features (channels 8h..8h+7) | partial dot over those channels | three measures:
    //   h = 0: mask, z', ray angle      h = 1: n_src (3)
    // The dot appears twice (both partial sums meet the same W1 column; the MMA adds them).
    const int k = kk / kBlk, j = kk - k * kBlk, h = j / (kBlk / 2), i = j - h * (kBlk / 2);
    const int base = kC * (kViews + 1);
    if (i < 8) return k * kC + 8 * h + i;
    if (i == 8) return base + 2 * kViews + 1 + k;          // dot
    if (h == 0) return i == 9 ? base + k : (i == 10 ? base + kViews + k : base + 3 * kViews + 1 + k);
    return base + 4 * kViews + 4 + 3 * k + (i - 9);        // n_src
  }
  if (kk < kViews * kBlk) {
    const int k = kk / kBlk, j = kk - k * kBlk;
    if (j < kC) return k * kC + j;                       // warped features
    const int base = kC * (kViews + 1);                  // 128
    switch (j - kC) {
      case 0: return base + k;                           // mask
      case 1: return base + kViews + k;                  // z'
      case 2: return base + 2 * kViews + 1 + k;          // dot
      case 3: return base + 3 * kViews + 1 + k;          // ray angle
      case 4: case 5: case 6: return base + 4 * kViews + 4 + 3 * k + (j - kC - 4);  // n_src
      default: return -1;                                // pad
    }
  }
  const int j = kk - kViews * kBlk;
  if (j < kC) return kViews * kC + j;                    // reference-frame features
  if (j == kC) return kC * (kViews + 1) + 2 * kViews;    // plane depth
  if (j < kC + 4) return kC * (kViews + 1) + 4 * kViews + 1 + (j - kC - 1);  // n_cur
  if (kk == kBiasPos) return -2;                         // bias
  return -1;                                             // pad
}
// reference channels of the pose measures of view k: comb, r, t (:718-720)
__host__ __device__ inline int pose_channel(int which, int k) {
  return kC * (kViews + 1) + (7 + which) * kViews + 4 + k;
}

// Builds the shared-memory image from nn.Linear weights: fp16 (hi, lo) core matrices.
__global__ void __launch_bounds__(256)
tc_pack_kernel(srcv_mlp_weights w, uint8_t* __restrict__ image) {
  __half* w1hi = reinterpret_cast<__half*>(image + kOffW1Hi);
  __half* w1lo = reinterpret_cast<__half*>(image + kOffW1Lo);
  __half* w2hi = reinterpret_cast<__half*>(image + kOffW2Hi);
  __half* w2lo = reinterpret_cast<__half*>(image + kOffW2Lo);
  float* vec = reinterpret_cast<float*>(image + kOffVec);
  const int n1 = kN * kK1, n2 = kN * kK2;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n1 + n2 + (int)kVecFloats;
       i += gridDim.x * blockDim.x) {
    if (i < n1) {
      const int n = i / kK1, kk = i - n * kK1;
      const int f = ref_channel(kk);
      const float v = kWScale * (f >= 0 ? w.w1[(size_t)n * kF + f] : (f == -2 ? w.b1[n] : 0.f));
      const __half h = __float2half_rn(v);
      w1hi[core_offset(n, kk, kN)] = h;
      w1lo[core_offset(n, kk, kN)] = __float2half_rn(v - __half2float(h));
    } else if (i < n1 + n2) {
      const int q = i - n1, n = q / kK2, kk = q - n * kK2;
      const float v = kWScale * w.w2[(size_t)n * kK2 + kk];
      const __half h = __float2half_rn(v);
      w2hi[core_offset(n, kk, kN)] = h;
      w2lo[core_offset(n, kk, kN)] = __float2half_rn(v - __half2float(h));
    } else {
      // LeakyReLU(h) w3 = (0.505 w3) h + (0.495 w3) |h|  (slope 0.01): two FMAs per column
      const int q = i - n1 - n2;
      float v = 0.f;
      if (q < kN) v = w.b2[q];
      else if (q < 2 * kN) v = (0.5f * (1.0f + kLeaky)) * w.w3[q - kN];
      else if (q < 3 * kN) v = (0.5f * (1.0f - kLeaky)) * w.w3[q - 2 * kN];
      else if (q == 3 * kN) v = w.b3[0];
      vec[q] = v;
    }
  }
}

// pb[b][n] = 16 * ( b1[n] + sum_k ( W1[n, comb_k] comb(b,k) + W1[n, r_k] r(b,k) + W1[n, t_k] t(b,k) ) ): the
// layer-1 bias of frame b = b1 + the contribution of the 21 pose measures, fp64 accumulation.
__global__ void __launch_bounds__(kN)
tc_frame_bias_kernel(srcv_mlp_weights w, const ViewParams* __restrict__ views, float* __restrict__ pb) {
  const int b = blockIdx.x, n = threadIdx.x;
  double acc = 0.0;
  for (int k = 0; k < kViews; ++k) {
    const ViewParams& vp = views[b * kViews + k];
    acc += (double)w.w1[(size_t)n * kF + pose_channel(0, k)] * (double)vp.comb;
    acc += (double)w.w1[(size_t)n * kF + pose_channel(1, k)] * (double)vp.rmeas;
    acc += (double)w.w1[(size_t)n * kF + pose_channel(2, k)] * (double)vp.tmeas;
  }
  // + b1: the MMA issuer writes this vector, split into fp16 (hi, lo), into W1's bias row whenever its
  // CTA moves on to another frame, so the whole layer-1 bias rides in the MMA (constant-one K position)
  pb[(size_t)b * kN + n] = (float)((double)kWScale * (acc + (double)w.b1[n]));
}

// 12 packed columns of one K block -> TMEM (hi and lo regions of the tile's A1 buffer)
__device__ __forceinline__ void store_block(uint32_t a1_lane, uint32_t col, const uint32_t (&hi)[kBlkCols],
                                            const uint32_t (&lo)[kBlkCols]) {
  st_x8(a1_lane + col, hi);
  st_x4(a1_lane + col + 8, hi + 8);
  st_x8(a1_lane + kA1LoOff + col, lo);
  st_x4(a1_lane + kA1LoOff + col + 8, lo + 8);
}

// What a worker thread keeps about the tile it builds the A operand for.
struct RowCtx {
  float dval, dxc, dyc;      // plane depth, centred pixel
  float X, Y, Z;             // back-projected point
  float cx, cy, cz;          // n_cur
  float4 cur4[4];            // reference-frame features of the pixel
  int b;                     // frame
  long long out;             // offset of the row's cost element, < 0 for rows outside the volume
  bool last_plane;           // the row's plane is D - 1 (the overall mask is taken there)
};

// The depth-invariant part of one (frame, view): four 16-byte loads (ViewParams is 128-byte
// aligned; the struct order a0 hx hy t | centre comb is what they read).
struct ViewRegs {
  float a0[3], hx[3], hy[3], t[3], centre[3];
};
__device__ __forceinline__ void load_view(const ViewParams* __restrict__ vp, ViewRegs& v) {
  const float4* p = reinterpret_cast<const float4*>(vp);
  const float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2), d = __ldg(p + 3);
  v.a0[0] = a.x; v.a0[1] = a.y; v.a0[2] = a.z; v.hx[0] = a.w;
  v.hx[1] = b.x; v.hx[2] = b.y; v.hy[0] = b.z; v.hy[1] = b.w;
  v.hy[2] = c.x; v.t[0] = c.y; v.t[1] = c.z; v.t[2] = c.w;
  v.centre[0] = d.x; v.centre[1] = d.y; v.centre[2] = d.z;
}

// One source view of one row: project, gather, metadata -> the 12 (hi, lo) column pairs of
// the K block.  HWC != 0: compile-time map size, every gather address is base + immediate.
// WANT_BITS (warp-uniform): also return the depth-valid / in-bounds bits of the overall mask.
template <int TW, int HWC>
__device__ __forceinline__ unsigned view_block(const RowCtx& rc, const ViewParams* __restrict__ vpp,
                                               const float4* __restrict__ view4, int Wrt, int H, int HWrt,
                                               const Centre& ctr, bool want_bits, uint32_t (&hi)[kBlkCols],
                                               uint32_t (&lo)[kBlkCols]) {
  const int W = TW ? TW : Wrt, HW = HWC ? HWC : HWrt;
  ViewRegs vr;
  load_view(vpp, vr);
  const float ax = fmaf(vr.hx[0], rc.dxc, fmaf(vr.hy[0], rc.dyc, vr.a0[0]));
  const float ay = fmaf(vr.hx[1], rc.dxc, fmaf(vr.hy[1], rc.dyc, vr.a0[1]));
  const float az = fmaf(vr.hx[2], rc.dxc, fmaf(vr.hy[2], rc.dyc, vr.a0[2]));
  float px, py, zp;
  project_point(rc.dval, ax, ay, az, vr.t[0], vr.t[1], vr.t[2], px, py, zp);
  // bilinear footprint.  Warp-uniform fast path: every lane's 2 x 2 footprint lies inside the
  // map (float compares: NaN / inf coordinates fail them and take the general path).
  const float x0f = floorf(px), y0f = floorf(py);
  const float fx = px - x0f, fy = py - y0f;
  const bool inside = x0f >= -(float)ctr.nx && x0f <= (float)(W - 2 - ctr.nx) &&
                      y0f >= -(float)ctr.ny && y0f <= (float)(H - 2 - ctr.ny);
  const bool interior = __all_sync(0xffffffffu, inside);
  const float gx = 1.0f - fx, gy = 1.0f - fy;
  float w00 = gx * gy, w01 = fx * gy, w10 = gx * fy, w11 = fx * fy;
  int o00, o01, o10, o11;
  if (interior) {
    o00 = ((int)y0f + ctr.ny) * W + ((int)x0f + ctr.nx);
    o01 = o00 + 1; o10 = o00 + W; o11 = o00 + W + 1;
  } else {
    // border patches, branch-free: every tap is loaded from an in-range (clamped) texel and
    // padding taps get a zero weight (zeros padding of grid_sample).
    Taps tp;
    bilinear_taps(px, py, W, H, ctr, tp);
    const int cxa = min(max(tp.x0, 0), W - 1), cxb = min(max(tp.x0 + 1, 0), W - 1);
    const int cya = min(max(tp.y0, 0), H - 1) * W, cyb = min(max(tp.y0 + 1, 0), H - 1) * W;
    o00 = cya + cxa; o01 = cya + cxb; o10 = cyb + cxa; o11 = cyb + cxb;
    w00 = (tp.valid & 1u) ? w00 : 0.f; w01 = (tp.valid & 2u) ? w01 : 0.f;
    w10 = (tp.valid & 4u) ? w10 : 0.f; w11 = (tp.valid & 8u) ? w11 : 0.f;
  }
  // features are sampled even for points behind the camera (only the dot is masked,
  // reference modules/cost_volume.py:590-623)
  // blend and dot on packed pairs (FFMA2 / FMUL2: two IEEE fp32 operations per issue slot — the
  // builders are issue-bound, not FMA-pipe-bound)
  float2 v2[kC / 2];
  {
    const float4 *q0 = view4 + o00, *q1 = view4 + o01, *q2 = view4 + o10, *q3 = view4 + o11;
    float4 f[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      f[0][j] = __ldg(q0 + (size_t)j * HW);
      f[1][j] = __ldg(q1 + (size_t)j * HW);
      f[2][j] = __ldg(q2 + (size_t)j * HW);
      f[3][j] = __ldg(q3 + (size_t)j * HW);
    }
    const float2 p00 = make_float2(w00, w00), p01 = make_float2(w01, w01), p10 = make_float2(w10, w10),
                 p11 = make_float2(w11, w11);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v2[2 * j] = fma2(p11, make_float2(f[3][j].x, f[3][j].y), fma2(p10, make_float2(f[2][j].x, f[2][j].y),
                  fma2(p01, make_float2(f[1][j].x, f[1][j].y), mul2(p00, make_float2(f[0][j].x, f[0][j].y)))));
      v2[2 * j + 1] = fma2(p11, make_float2(f[3][j].z, f[3][j].w), fma2(p10, make_float2(f[2][j].z, f[2][j].w),
                      fma2(p01, make_float2(f[1][j].z, f[1][j].w), mul2(p00, make_float2(f[0][j].z, f[0][j].w)))));
    }
  }
  float2 d2 = make_float2(0.f, 0.f);       // even / odd channel partial sums
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    d2 = fma2(v2[2 * j], make_float2(rc.cur4[j].x, rc.cur4[j].y), d2);
    d2 = fma2(v2[2 * j + 1], make_float2(rc.cur4[j].z, rc.cur4[j].w), d2);
  }
  const float dot = d2.x + d2.y;
  const float mk = zp > 0.0f ? 1.0f : 0.0f;
  // n_src = (X - centre_k)/max(|.|, 1e-12) ; ray angle = cosine_similarity(n_cur, n_src, eps 1e-5)
  const float sx0 = rc.X - vr.centre[0], sy0 = rc.Y - vr.centre[1], sz0 = rc.Z - vr.centre[2];
  const float is = inv_norm(fmaf(sx0, sx0, fmaf(sy0, sy0, sz0 * sz0)), kEpsNorm);
  const float sx = sx0 * is, sy = sy0 * is, sz = sz0 * is;
  // cosine_similarity divides each operand by max(|.|, 1e-5) again (:683-688): both are unit vectors to
  // an ulp here (or exactly zero, which stays zero), so the plain dot product is the same to ~1e-7
  const float ang = fmaf(rc.cx, sx, fmaf(rc.cy, sy, rc.cz * sz));
#pragma unroll
  for (int i = 0; i < 8; ++i) split_pack(v2[i].x, v2[i].y, hi[i], lo[i]);
  split_pack(mk, zp, hi[8], lo[8]);
  split_pack(dot * mk, ang, hi[9], lo[9]);
  split_pack(sx, sy, hi[10], lo[10]);
  split_pack(sz, 0.0f, hi[11], lo[11]);
  unsigned bits = 0;
  if (want_bits) {
    if (zp > 0.0f) bits |= 1u;
    if (in_mask_bounds(px, py, W, H, ctr)) bits |= 2u;
  }
  return bits;
}

// The view-independent tail block: reference features | plane depth | n_cur | one | zeros
__device__ __forceinline__ void tail_block(const RowCtx& rc, uint32_t (&hi)[kBlkCols], uint32_t (&lo)[kBlkCols]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    split_pack(rc.cur4[j].x, rc.cur4[j].y, hi[2 * j], lo[2 * j]);
    split_pack(rc.cur4[j].z, rc.cur4[j].w, hi[2 * j + 1], lo[2 * j + 1]);
  }
  split_pack(rc.dval, rc.cx, hi[8], lo[8]);
  split_pack(rc.cy, rc.cz, hi[9], lo[9]);
  hi[10] = 0x00003C00u; lo[10] = 0u;      // (1.0, 0): the bias position
  hi[11] = 0u; lo[11] = 0u;
}

// ---- half-view unit: how view kSplitView is built ------------------------------------------------
// A unit = one source view, channels 8h..8h+7: eight vector gathers, eight warped channels, the
// partial dot over those channels and three of the view's six per-sample measures = 6 packed
// columns.  The projection is done by both threads that share the view (~70 instructions each);
// what moves off the critical slots is half of the gathers, the blend, the split and the stores.
constexpr int kUnitCols = kBlkCols / 2;     // 6

// what the conversion of a unit needs once its gathers are under way
struct UnitCtx {
  float w00, w01, w10, w11;   // bilinear weights (zeros-padding taps: 0)
  float mk, e0, e1, e2;       // depth validity of the view; the unit's three measures
};

// project, set the footprint up and ISSUE the unit's eight gathers (f: [tap][chunk of the half])
template <int TW, int HWC>
__device__ __forceinline__ unsigned unit_issue(const RowCtx& rc, int k, int half, const float4* __restrict__ src4,
                                               const ViewParams* __restrict__ views, int Wrt, int H, int HWrt,
                                               const Centre& ctr, bool want_bits, UnitCtx& uc, float4 (&f)[4][2]) {
  const int W = TW ? TW : Wrt, HW = HWC ? HWC : HWrt;
  const ViewParams* vpp = views + (rc.b * kViews + k);
  const float4* view4 = src4 + ((size_t)(rc.b * kViews + k) * 4 + 2 * half) * HW;
  ViewRegs vr;
  load_view(vpp, vr);
  const float ax = fmaf(vr.hx[0], rc.dxc, fmaf(vr.hy[0], rc.dyc, vr.a0[0]));
  const float ay = fmaf(vr.hx[1], rc.dxc, fmaf(vr.hy[1], rc.dyc, vr.a0[1]));
  const float az = fmaf(vr.hx[2], rc.dxc, fmaf(vr.hy[2], rc.dyc, vr.a0[2]));
  float px, py, zp;
  project_point(rc.dval, ax, ay, az, vr.t[0], vr.t[1], vr.t[2], px, py, zp);
  const float x0f = floorf(px), y0f = floorf(py);
  const float fx = px - x0f, fy = py - y0f;
  const bool inside = x0f >= -(float)ctr.nx && x0f <= (float)(W - 2 - ctr.nx) &&
                      y0f >= -(float)ctr.ny && y0f <= (float)(H - 2 - ctr.ny);
  const bool interior = __all_sync(0xffffffffu, inside);
  const float gx = 1.0f - fx, gy = 1.0f - fy;
  uc.w00 = gx * gy; uc.w01 = fx * gy; uc.w10 = gx * fy; uc.w11 = fx * fy;
  int o00, o01, o10, o11;
  if (interior) {
    o00 = ((int)y0f + ctr.ny) * W + ((int)x0f + ctr.nx);
    o01 = o00 + 1; o10 = o00 + W; o11 = o00 + W + 1;
  } else {
    // border patches, branch-free: clamped (in-range) addresses, zero weights for padding taps
    Taps tp;
    bilinear_taps(px, py, W, H, ctr, tp);
    const int cxa = min(max(tp.x0, 0), W - 1), cxb = min(max(tp.x0 + 1, 0), W - 1);
    const int cya = min(max(tp.y0, 0), H - 1) * W, cyb = min(max(tp.y0 + 1, 0), H - 1) * W;
    o00 = cya + cxa; o01 = cya + cxb; o10 = cyb + cxa; o11 = cyb + cxb;
    uc.w00 = (tp.valid & 1u) ? uc.w00 : 0.f; uc.w01 = (tp.valid & 2u) ? uc.w01 : 0.f;
    uc.w10 = (tp.valid & 4u) ? uc.w10 : 0.f; uc.w11 = (tp.valid & 8u) ? uc.w11 : 0.f;
  }
  {
    const float4 *q0 = view4 + o00, *q1 = view4 + o01, *q2 = view4 + o10, *q3 = view4 + o11;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
#ifdef SRCV_TC_ABL_NOGATHER   // ablation build (wrong results): how much of a tile is gather latency?
      f[0][j] = f[1][j] = f[2][j] = f[3][j] = make_float4(uc.w00, uc.w01, (float)(o00 + o11), (float)(o01 + o10 + (int)(size_t)q0));
      (void)q1; (void)q2; (void)q3;
#else
      f[0][j] = __ldg(q0 + (size_t)j * HW);
      f[1][j] = __ldg(q1 + (size_t)j * HW);
      f[2][j] = __ldg(q2 + (size_t)j * HW);
      f[3][j] = __ldg(q3 + (size_t)j * HW);
#endif
    }
  }
  // the view's measures (independent of the gathers): mask, z', ray angle | n_src
  const float mk = zp > 0.0f ? 1.0f : 0.0f;
  const float sx0 = rc.X - vr.centre[0], sy0 = rc.Y - vr.centre[1], sz0 = rc.Z - vr.centre[2];
  const float is = inv_norm(fmaf(sx0, sx0, fmaf(sy0, sy0, sz0 * sz0)), kEpsNorm);
  const float sx = sx0 * is, sy = sy0 * is, sz = sz0 * is;
  // cosine_similarity divides each operand by max(|.|, 1e-5) again (:683-688): both are unit vectors to
  // an ulp here (or exactly zero, which stays zero), so the plain dot product is the same to ~1e-7
  const float ang = fmaf(rc.cx, sx, fmaf(rc.cy, sy, rc.cz * sz));
  uc.mk = mk;
  uc.e0 = half ? sx : mk;
  uc.e1 = half ? sy : zp;
  uc.e2 = half ? sz : ang;
  unsigned bits = 0;
  if (want_bits) {
    if (zp > 0.0f) bits |= 1u;
    if (in_mask_bounds(px, py, W, H, ctr)) bits |= 2u;
  }
  return bits;
}

// bilinear blend, partial dot, (hi, lo) split and the unit's 6 + 6 packed columns -> TMEM
__device__ __forceinline__ void unit_convert(const RowCtx& rc, int k, int half, const UnitCtx& uc,
                                             const float4 (&f)[4][2], uint32_t a1_lane) {
  const float2 p00 = make_float2(uc.w00, uc.w00), p01 = make_float2(uc.w01, uc.w01),
               p10 = make_float2(uc.w10, uc.w10), p11 = make_float2(uc.w11, uc.w11);
  float2 v2[4], d2 = make_float2(0.f, 0.f);
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    v2[2 * j] = fma2(p11, make_float2(f[3][j].x, f[3][j].y), fma2(p10, make_float2(f[2][j].x, f[2][j].y),
                fma2(p01, make_float2(f[1][j].x, f[1][j].y), mul2(p00, make_float2(f[0][j].x, f[0][j].y)))));
    v2[2 * j + 1] = fma2(p11, make_float2(f[3][j].z, f[3][j].w), fma2(p10, make_float2(f[2][j].z, f[2][j].w),
                    fma2(p01, make_float2(f[1][j].z, f[1][j].w), mul2(p00, make_float2(f[0][j].z, f[0][j].w)))));
    const float4 c = half ? rc.cur4[2 + j] : rc.cur4[j];
    d2 = fma2(v2[2 * j], make_float2(c.x, c.y), d2);
    d2 = fma2(v2[2 * j + 1], make_float2(c.z, c.w), d2);
  }
  uint32_t hi[kUnitCols], lo[kUnitCols];
#pragma unroll
  for (int i = 0; i < 4; ++i) split_pack(v2[i].x, v2[i].y, hi[i], lo[i]);
  split_pack((d2.x + d2.y) * uc.mk, uc.e0, hi[4], lo[4]);
  split_pack(uc.e1, uc.e2, hi[5], lo[5]);
  const uint32_t col = (uint32_t)(kBlkCols * k + kUnitCols * half);
  st_x4(a1_lane + col, hi);
  st_x2(a1_lane + col + 4, hi + 4);
  st_x4(a1_lane + kA1LoOff + col, lo);
  st_x2(a1_lane + kA1LoOff + col + 4, lo + 4);
}

// Optional per-tile timeline of CTA 0 (experiment builds only, -DSRCV_TC_TIMELINE): one lane of
// each role stamps %globaltimer-free SM clocks at its phase boundaries; read back with
// srcv_debug_read_timeline.  Not compiled into the shipped library.
#ifdef SRCV_TC_TIMELINE
constexpr int kTlTiles = 48, kTlEvents = 16;
__device__ long long g_timeline[kTlTiles * kTlEvents];
#define SRCV_TL(it, ev)                                                                            \
  do {                                                                                             \
    if (blockIdx.x == 0 && lane == 0 && (warp & 3) == 0 && (it) < (unsigned)kTlTiles)                \
      g_timeline[(it) * kTlEvents + (ev)] = clock64();                                             \
  } while (0)
#else
#define SRCV_TL(it, ev) do { } while (0)
#endif

// issue D (+)= A_hi W_hi + A_hi W_lo + A_lo W_hi over the K steps [KS0, KS1) of 16, for the N
// accumulator columns whose weight rows start at smem_hi / smem_lo (row n of a K chunk sits 16 n
// bytes into it, so a column half is a start-address offset; the K-chunk stride stays kN * 16).
// LAYER2: the operand sits where the layer-1 accumulator was — k-step ks reads the hi columns
// 32 (ks / 2) + 8 (ks % 2) and the lo columns 16 further (see the layer-1 epilogue).
#ifdef SRCV_TC_NO_SPLIT_HALVES   // experiment switch: whole-layer MMAs, epilogue and tensor pipe take turns
constexpr bool kSplitHalves = false;
#else
constexpr bool kSplitHalves = true;
#endif

template <int N, int KS0, int KS1, bool LAYER2>
__device__ __forceinline__ void issue_steps(uint32_t d, uint32_t a_base, uint32_t a_lo_off,
                                            uint32_t smem_hi, uint32_t smem_lo) {
  constexpr uint32_t idesc = idesc_f16_f32(kRows, N);
  constexpr uint32_t kLbo = kN * 16, kSbo = 128, kStepBytes = 2 * kLbo;  // two 8-wide K chunks per MMA
  // only the 14-bit start-address field changes from step to step
  uint64_t bhi = smem_desc(smem_hi + KS0 * kStepBytes, kLbo, kSbo), blo = smem_desc(smem_lo + KS0 * kStepBytes, kLbo, kSbo);
#pragma unroll 1
  for (int ks = KS0; ks < KS1; ++ks) {
    const uint32_t ahi = LAYER2 ? a_base + 32u * (uint32_t)(ks >> 1) + 8u * (uint32_t)(ks & 1)
                                : a_base + 8u * (uint32_t)ks;
    const uint32_t alo = ahi + a_lo_off;
    mma_ts(d, ahi, bhi, idesc, ks > 0 ? 1u : 0u);
    mma_ts(d, ahi, blo, idesc, 1u);
    mma_ts(d, alo, bhi, idesc, 1u);
    bhi += kStepBytes >> 4; blo += kStepBytes >> 4;
  }
}

// tile id runs plane-chunk fastest, then pixel patch, then frame (32-bit: the launcher
// refuses >= 2^31 tiles); row -> (plane-in-chunk = row / 32, pixel of the 16 x 2 patch = row % 32)
// id / nd for the runtime plane-chunk count nd: multiply-high by ceil(2^32 / nd), exact while
// id * nd < 2^32 (the launcher refuses larger volumes) — 2 instructions instead of the ~20 of a
// generic 32-bit division, in every warp of every tile.
struct DivNd {
  unsigned nd, magic;
  __device__ __forceinline__ explicit DivNd(unsigned n) : nd(n), magic((unsigned)((0x100000000ull + n - 1) / n)) {}
  __device__ __forceinline__ unsigned div(unsigned id) const { return nd == 1u ? id : __umulhi(id, magic); }
};

template <bool PER_PIXEL>
__device__ __forceinline__ void make_row(unsigned id, int row, int W, int H, int HW, int D, const DivNd& nd,
                                         unsigned tiles_x, unsigned tiles_xy, const Centre& ctr,
                                         const float4* __restrict__ cur4g,
                                         const FrameParams* __restrict__ frames,
                                         const float* __restrict__ planes, RowCtx& rc) {
  const unsigned r = nd.div(id);
  const int d0 = (int)(id - r * nd.nd) * kTileD;
  const unsigned txy = r % tiles_xy;
  const int b = (int)(r / tiles_xy);
  const int x0 = (int)(txy % tiles_x) * kTileW, y0 = (int)(txy / tiles_x) * kTileH;
  const int rx = row & (kTileW - 1), ry = (row >> 4) & (kTileH - 1), dd = row >> 5;
  const int d = min(d0 + dd, D - 1);
  const bool active = (x0 + rx < W) && (y0 + ry < H) && (d0 + dd < D);
  const int ox = min(x0 + rx, W - 1), oy = min(y0 + ry, H - 1);
  const int p = oy * W + ox;
  rc.b = b;
  rc.out = active ? ((long long)b * D + d) * HW + p : -1;
  rc.last_plane = (d == D - 1);
  const float pxc = (float)ox + 0.5f, pyc = (float)oy + 0.5f;
  rc.dval = PER_PIXEL ? __ldg(planes + ((size_t)b * D + d) * HW + p) : __ldg(planes + b * D + d);
  rc.dxc = pxc - ctr.half_w;
  rc.dyc = pyc - ctr.half_h;
#pragma unroll
  for (int j = 0; j < 4; ++j) rc.cur4[j] = __ldg(cur4g + ((size_t)b * 4 + j) * HW + p);
  // rays: X = d * (invK3 p); n_cur = X / max(|X|, 1e-12)
  const float4* fp = reinterpret_cast<const float4*>(frames + b);
  const float4 k0 = __ldg(fp), k1 = __ldg(fp + 1);
  const float k8 = __ldg(reinterpret_cast<const float*>(fp + 2));
  const float rxv = fmaf(k0.x, pxc, fmaf(k0.y, pyc, k0.z));
  const float ryv = fmaf(k0.w, pxc, fmaf(k1.x, pyc, k1.y));
  const float rzv = fmaf(k1.z, pxc, fmaf(k1.w, pyc, k8));
  rc.X = rc.dval * rxv; rc.Y = rc.dval * ryv; rc.Z = rc.dval * rzv;
  const float ic = inv_norm(fmaf(rc.X, rc.X, fmaf(rc.Y, rc.Y, rc.Z * rc.Z)), kEpsNorm);
  rc.cx = rc.X * ic; rc.cy = rc.Y * ic; rc.cz = rc.Z * ic;
}

// K block `blk` (0..6: source view, 7: view-independent tail) of a row -> packed (hi, lo)
template <int TW, int HWC>
__device__ __forceinline__ unsigned build_block(const RowCtx& rc, int blk, const float4* __restrict__ src4,
                                                const ViewParams* __restrict__ views, int W, int H, int HW,
                                                const Centre& ctr, bool want_bits, uint32_t (&hi)[kBlkCols],
                                                uint32_t (&lo)[kBlkCols]) {
  if (blk < kViews)
    return view_block<TW, HWC>(rc, views + (rc.b * kViews + blk),
                               src4 + ((size_t)(rc.b * kViews + blk) * 4) * HW, W, H, HW, ctr, want_bits, hi, lo);
  tail_block(rc, hi, lo);
  return 0u;
}

template <bool PER_PIXEL, int TW, int TH>
__global__ void __launch_bounds__(kThreads, 1)
mlp_tc_kernel(srcv_shape s, const float4* __restrict__ cur4g, const float4* __restrict__ src4,
              const ViewParams* __restrict__ views, const FrameParams* __restrict__ frames,
              const float* __restrict__ planes, const uint8_t* __restrict__ image,
              const float* __restrict__ frame_bias, float* __restrict__ cost, uint8_t* __restrict__ mask_out,
              unsigned num_tiles) {
  SRCV_DYNAMIC_SMEM_ALIGNED(uint8_t, smem, 1024);
  __shared__ uint32_t s_tmem_base;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  // Barriers of a BUFFER (a1_full, d2_free) exist once per A1 buffer and complete every second
  // tile: their waiters can lag two tiles, which a single phase bit could not tell apart.
  uint64_t* bar_a1_full = bars + 0;   // [2] builders -> MMA: A1 of a tile is in TMEM          (384 arrivals)
  uint64_t* bar_d2_free = bars + 2;   // [2] epilogues done reading D2: the buffer is free     (128 arrivals)
  // Layer 1 runs as two 64-column halves so that the epilogue of the first half overlaps the MMAs
  // of the second, and layer 2 as two K halves so that it starts on the first half's operand while
  // the epilogue still converts the second: the tensor pipe and the epilogue warps stop taking turns.
  uint64_t* bar_mma1 = bars + 4;      // [2] layer-1 MMAs of a column half done: DA half holds D1 (commit)
  uint64_t* bar_a2_full = bars + 6;   // [2] epilogues -> MMA: A2 half written over its D1 half  (128 arrivals)
  uint64_t* bar_mma2 = bars + 8;      // layer-2 MMAs done: D2 ready, DA free                 (commit)
  uint64_t* bar_img = bars + 9;       // weight image landed in shared memory (bulk copies)
  uint8_t* sflag = smem + kOffFlag;
  const float* svec = reinterpret_cast<const float*>(smem + kOffVec);

  // The warp index goes through a shuffle broadcast so that ptxas knows the role branches are
  // warp-uniform and keeps the global-memory descriptors in uniform registers.
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const int W = TW ? TW : s.W, H = TH ? TH : s.H, HW = W * H, D = s.D;
  constexpr int HWC = TW * TH;
  const unsigned tiles_x = (unsigned)(W + kTileW - 1) / kTileW;
  const unsigned tiles_xy = tiles_x * ((unsigned)(H + kTileH - 1) / kTileH);
  const DivNd nd((unsigned)(D + kTileD - 1) / kTileD);

  // ---- one-time setup ----------------------------------------------------------------
  if (tid == 0) {
    mbar_init(bar_img, 1);
    mbar_init(bar_a1_full, kBuilders);
    mbar_init(bar_a1_full + 1, kBuilders);
    mbar_init(bar_d2_free, kEpis);
    mbar_init(bar_d2_free + 1, kEpis);
    mbar_init(bar_mma1, 1);
    mbar_init(bar_mma1 + 1, 1);
    mbar_init(bar_a2_full, kEpis);
    mbar_init(bar_a2_full + 1, kEpis);
    mbar_init(bar_mma2, 1);
    mbar_fence_init();
  }
  if (warp == kMmaWarp) tmem_alloc(&s_tmem_base, kTmemCols);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  // The 170 KB weight image (fp16 core matrices + biases) is pulled in by the bulk-copy
  // (TMA) engine straight into shared memory — written by the async proxy, which is also
  // the proxy the tensor core reads it through — while the threads finish their setup.
  if (tid == 0) {
    mbar_expect_tx(bar_img, kImageBytes);
    constexpr uint32_t kPiece = 32768;
    for (uint32_t off = 0; off < kImageBytes; off += kPiece)
      bulk_g2s(smem + off, image + off, (kImageBytes - off < kPiece) ? (kImageBytes - off) : kPiece, bar_img);
  }
  mbar_wait(bar_img, 0);
  const uint32_t tmem_base = s_tmem_base;
  const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
  // Work split: CTA c takes tiles c, c + grid, c + 2 grid, ... of the plane-fastest tile
  // order, i.e. at any moment the 148 CTAs sweep ~2 neighbouring pixel blocks x 64 planes of
  // ONE frame, whose source features (39 MB) stay L2-resident, and every CTA sees the same mix
  // of interior and border patches.
#ifdef SRCV_TC_CONTIG
  // experiment switch: one contiguous tile range per CTA (consecutive plane chunks of a patch)
  const unsigned t_lo = (unsigned)(((unsigned long long)num_tiles * blockIdx.x) / gridDim.x);
  const unsigned t_hi = (unsigned)(((unsigned long long)num_tiles * (blockIdx.x + 1)) / gridDim.x);
  const unsigned n_local = t_hi - t_lo;
  auto tile_id = [&](unsigned j) { return t_lo + j; };
#else
  const unsigned n_local = (num_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
  auto tile_id = [&](unsigned j) { return blockIdx.x + j * gridDim.x; };
#endif

  if (warp < kBuildWarps) {
    // =============================== builders ============================================
    // Twelve warps: three threads per row (slot = warp / 4): two whole source views each, then half of
    // view kSplitView (slots 0, 1) or the tail (slot 2).  With
    // two A1 buffers they run up to a whole tile ahead of the tensor pipe: buffer t & 1 is free
    // again once the layer-2 epilogue of tile t - 2 has read its accumulator out of it.
    if (kSetmaxnreg && kRegsBuild > kRegsLaunch) reg_inc<kRegsBuild>();
    const int row = (warp & 3) * 32 + lane, slot = warp >> 2;
    const Centre ctr(W, H);
    // slot 0: views 0, 1 + the first half of view kSplitView; slot 1: views 3, 4 + its second half;
    // slot 2: views 5, 6 + the tail
    const int blk_first = kSlots == 4 ? 2 * slot : ((slot < 2) ? 3 * slot : 5);
    const bool masks = mask_out != nullptr;
    RowCtx rc;
    uint32_t hi[kBlkCols], lo[kBlkCols];
    for (unsigned it = 0; it < n_local; ++it) {
      const uint32_t buf = it & 1u, use = it >> 1;       // use-th tile of this buffer
      SRCV_TL(it, 0);
      make_row<PER_PIXEL>(tile_id(it), row, W, H, HW, D, nd, tiles_x, tiles_xy, ctr, cur4g, frames, planes, rc);
      const bool wb = masks && rc.last_plane;
      const uint32_t a1_lane = lane_base + kColA1 + buf * kA1Stride;
      unsigned bits = build_block<TW, HWC>(rc, blk_first, src4, views, W, H, HW, ctr, wb, hi, lo);
      SRCV_TL(it, 1);
      if (use > 0) {
        mbar_wait(bar_d2_free + buf, (use - 1) & 1u);    // tile it - 2 is completely out of this buffer
        fence_after_sync();
      }
      SRCV_TL(it, 2);
      store_block(a1_lane, (uint32_t)(kBlkCols * blk_first), hi, lo);
      bits |= build_block<TW, HWC>(rc, blk_first + 1, src4, views, W, H, HW, ctr, wb, hi, lo);   // slot 3 of 4: the tail
      store_block(a1_lane, (uint32_t)(kBlkCols * (blk_first + 1)), hi, lo);
      if (kSlots == 4) {
        // two blocks per thread: done
      } else if (slot < 2) {
        float4 f[4][2];
        UnitCtx uc;
        bits |= unit_issue<TW, HWC>(rc, kSplitView, slot, src4, views, W, H, HW, ctr, wb, uc, f);
        unit_convert(rc, kSplitView, slot, uc, f, a1_lane);
      } else {
        tail_block(rc, hi, lo);
        store_block(a1_lane, (uint32_t)(kBlkCols * kViews), hi, lo);
      }
      // Mask bits of this tile live with its buffer: the epilogue warp reads them before its
      // bar_d2_free arrival for this tile, and the next store into the slot (tile it + 2) follows
      // this thread's wait on exactly that barrier phase.
      if (wb) sflag[(buf * kSlots + slot) * kRows + row] = (uint8_t)bits;
      wait_st();
      fence_before_sync();
      mbar_arrive(bar_a1_full + buf);
      SRCV_TL(it, 3);
    }
  } else if (warp < kBuildWarps + kEpiWarps) {
    // =============================== epilogues ===========================================
    // Four warps, one thread per row.  Layer-1 epilogue: per-frame pose bias, LeakyReLU, (hi, lo)
    // split, layer-2 operand written IN PLACE over the accumulator, 32 columns at a time (a
    // chunk's 32 fp32 columns become its 16 hi + 16 lo operand columns).  Layer-2 epilogue: bias,
    // LeakyReLU and the 128 -> 1 layer as one dot product per row — no cross-thread reduction.
    if (kSetmaxnreg) reg_inc<kRegsEpi>();
    const int row = (warp & 3) * 32 + lane;
    const bool masks = mask_out != nullptr;
    for (unsigned it = 0; it < n_local; ++it) {
      const uint32_t par = it & 1u, buf = it & 1u;
      // where this row's cost goes (integer part of make_row)
      long long out;
      bool last_plane;
      int b;
      {
        const unsigned id = tile_id(it);
        const unsigned r = nd.div(id), txy = r % tiles_xy;
        const int d0 = (int)(id - r * nd.nd) * kTileD;
        b = (int)(r / tiles_xy);
        const int x0 = (int)(txy % tiles_x) * kTileW, y0 = (int)(txy / tiles_x) * kTileH;
        const int rx = row & (kTileW - 1), ry = (row >> 4) & (kTileH - 1), dd = row >> 5;
        const int d = d0 + dd;
        const bool active = (x0 + rx < W) && (y0 + ry < H) && (d < D);
        out = active ? ((long long)b * D + d) * HW + ((y0 + ry) * W + (x0 + rx)) : -1;
        last_plane = (d == D - 1);
      }
      // kEpiStep columns per step (64: two TMEM loads in flight per wait, twice the independent
      // work per dependent chain; 32 where the register budget of this group is 96)
#pragma unroll 1
      for (int c = 0; c < kN; c += kEpiStep) {
        if ((c & 63) == 0 && (kSplitHalves || c == 0)) {
          mbar_wait(bar_mma1 + (c >> 6), par);   // this column half of D1 is complete
          fence_after_sync();
          SRCV_TL(it, c ? 14 : 4);
        }
        uint32_t r[kEpiStep];
#pragma unroll
        for (int h = 0; h < kEpiStep / 32; ++h) ld_x32(lane_base + kColDA + c + 32 * h, r + 32 * h);
        wait_ld();
#pragma unroll
        for (int h = 0; h < kEpiStep / 32; ++h) {
          uint32_t ehi[16], elo[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            // LeakyReLU = max(x, 0.01 x), the multiply as one packed FMUL2 per column pair (the
            // bias, incl. the frame's pose measures, came out of the MMA)
            const float2 x = make_float2(__uint_as_float(r[32 * h + 2 * j]), __uint_as_float(r[32 * h + 2 * j + 1]));
            const float2 m = mul2(x, make_float2(kLeaky, kLeaky));
            split_pack(fmaxf(x.x, m.x), fmaxf(x.y, m.y), ehi[j], elo[j]);
          }
          st_x16(lane_base + kColDA + c + 32 * h, ehi);
          st_x16(lane_base + kColDA + c + 32 * h + 16, elo);
        }
        if (((c + kEpiStep) & 63) == 0 && (kSplitHalves || c + kEpiStep == kN)) {
          wait_st();
          fence_before_sync();
          mbar_arrive(bar_a2_full + (c >> 6));   // k-steps 4 (c / 64) .. + 3 of the layer-2 operand
          SRCV_TL(it, (c >> 6) ? 5 : 13);
        }
      }
      // mask bits of the builders: visible since the bar_mma1 waits above (a1_full -> MMA -> commit)
      unsigned tile_bits = 0;
      if (masks && last_plane) {
        const uint8_t* fl = sflag + buf * kSlots * kRows + row;
#pragma unroll
        for (int q = 0; q < kSlots; ++q) tile_bits |= fl[q * kRows];
      }
      // ---- layer-2 epilogue: the accumulator sits in the first 128 columns of this tile's A1 buffer
      mbar_wait(bar_mma2, par);
      fence_after_sync();
      SRCV_TL(it, 6);
      const uint32_t d2_lane = lane_base + kColA1 + buf * kA1Stride;
      float2 acc2 = make_float2(0.f, 0.f);     // even / odd columns, as packed FFMA2 chains
#pragma unroll 1
      for (int c = 0; c < kN; c += kEpiStep) {
        uint32_t r[kEpiStep];
#pragma unroll
        for (int h = 0; h < kEpiStep / 32; ++h) ld_x32(d2_lane + c + 32 * h, r + 32 * h);
        wait_ld();
        if (c == kN - kEpiStep) {
          fence_before_sync();
          mbar_arrive(bar_d2_free + buf);   // the buffer is free: the builders may write tile it + 2 into it
          SRCV_TL(it, 7);
        }
#pragma unroll
        for (int j = 0; j < kEpiStep; j += 4) {
          const float4 bb = *reinterpret_cast<const float4*>(svec + c + j);
          const float4 wa = *reinterpret_cast<const float4*>(svec + kN + c + j);
          const float4 wn = *reinterpret_cast<const float4*>(svec + 2 * kN + c + j);
          const float2 us = make_float2(kUnscale2, kUnscale2);
          const float2 h01 = fma2(make_float2(__uint_as_float(r[j + 0]), __uint_as_float(r[j + 1])), us, make_float2(bb.x, bb.y));
          const float2 h23 = fma2(make_float2(__uint_as_float(r[j + 2]), __uint_as_float(r[j + 3])), us, make_float2(bb.z, bb.w));
          acc2 = fma2(make_float2(wn.x, wn.y), make_float2(fabsf(h01.x), fabsf(h01.y)), fma2(make_float2(wa.x, wa.y), h01, acc2));
          acc2 = fma2(make_float2(wn.z, wn.w), make_float2(fabsf(h23.x), fabsf(h23.y)), fma2(make_float2(wa.z, wa.w), h23, acc2));
        }
      }
      if (out >= 0) {
        cost[out] = (acc2.x + acc2.y) + svec[3 * kN];
        if (masks && last_plane) {
          // overall mask of the LAST plane (reference :625-637): any view in front AND any view
          // inside the 2-pixel border, independently
          mask_out[(size_t)b * HW + (out - ((long long)b * D + (D - 1)) * HW)] = (tile_bits == 3u) ? 1 : 0;
        }
      }
      SRCV_TL(it, 8);
    }
  } else {
    if (kSetmaxnreg) reg_dec<kRegsMma>();
    if (warp == kMmaWarp) {
      // =============================== MMA issuer ==========================================
      const uint32_t sbase = smem_u32(smem);
      __half* w1hi = reinterpret_cast<__half*>(smem + kOffW1Hi);
      __half* w1lo = reinterpret_cast<__half*>(smem + kOffW1Lo);
      int bias_frame = -1;
      for (unsigned it = 0; it < n_local; ++it) {
        const uint32_t par = it & 1u, buf = it & 1u, use = it >> 1;
        const uint32_t a1 = tmem_base + kColA1 + buf * kA1Stride;
        // The layer-1 bias of the tile's frame (b1 + its pose measures) is W1's row at the constant-one
        // K position: rewritten (fp16 hi, lo) when the CTA's tiles move on to another frame.  Its
        // readers — the layer-1 MMAs of earlier tiles — completed before this warp issued the
        // previous tile's last layer-2 MMAs (bar_mma1 -> epilogue -> bar_a2_full -> here).
        const int fb = (int)(nd.div(tile_id(it)) / tiles_xy);
        if (fb != bias_frame) {
          bias_frame = fb;
#pragma unroll
          for (int i = 0; i < kN / 32; ++i) {
            const int n = lane + 32 * i;
            const float v = __ldg(frame_bias + (size_t)fb * kN + n);
            const __half h = __float2half_rn(v);
            w1hi[core_offset(n, kBiasPos, kN)] = h;
            w1lo[core_offset(n, kBiasPos, kN)] = __float2half_rn(v - __half2float(h));
          }
          fence_proxy_async_smem();                // generic-proxy stores -> visible to the tensor core
          __syncwarp();
        }
        mbar_wait(bar_a1_full + buf, use & 1u);   // A1 of this tile is in TMEM
        fence_after_sync();
        SRCV_TL(it, 9);
        // DA is free: the layer-2 MMAs of the previous tile (its last readers) were issued before
        // these and the tensor pipe executes this thread's MMAs in order.
        if (kSplitHalves) {
          if (elect_one()) {
            issue_steps<kN / 2, 0, kK1 / 16, false>(tmem_base + kColDA, a1, kA1LoOff, sbase + kOffW1Hi, sbase + kOffW1Lo);
            mma_commit(bar_mma1);
            issue_steps<kN / 2, 0, kK1 / 16, false>(tmem_base + kColDA + kN / 2, a1, kA1LoOff,
                                                    sbase + kOffW1Hi + (kN / 2) * 16, sbase + kOffW1Lo + (kN / 2) * 16);
            mma_commit(bar_mma1 + 1);
          }
          __syncwarp();
          SRCV_TL(it, 10);
          mbar_wait(bar_a2_full, par);             // first half of A2 written over D1
          fence_after_sync();
          SRCV_TL(it, 11);
          // D2 goes into this tile's own A1 buffer: its layer-1 MMAs (just above, same in-order
          // pipe) are the last readers, and the builders do not touch it before bar_d2_free.
          if (elect_one())
            issue_steps<kN, 0, kK2 / 32, true>(a1, tmem_base + kColDA, 16u, sbase + kOffW2Hi, sbase + kOffW2Lo);
          __syncwarp();
          mbar_wait(bar_a2_full + 1, par);         // second half
          fence_after_sync();
          SRCV_TL(it, 15);
          if (elect_one()) {
            issue_steps<kN, kK2 / 32, kK2 / 16, true>(a1, tmem_base + kColDA, 16u, sbase + kOffW2Hi, sbase + kOffW2Lo);
            mma_commit(bar_mma2);
          }
        } else {
          if (elect_one()) {
            issue_steps<kN, 0, kK1 / 16, false>(tmem_base + kColDA, a1, kA1LoOff, sbase + kOffW1Hi, sbase + kOffW1Lo);
            mma_commit(bar_mma1);
          }
          __syncwarp();
          SRCV_TL(it, 10);
          mbar_wait(bar_a2_full + 1, par);         // A2 written over D1
          fence_after_sync();
          SRCV_TL(it, 11);
          if (elect_one()) {
            issue_steps<kN, 0, kK2 / 16, true>(a1, tmem_base + kColDA, 16u, sbase + kOffW2Hi, sbase + kOffW2Lo);
            mma_commit(bar_mma2);
          }
        }
        __syncwarp();
        SRCV_TL(it, 12);
      }
    }
  }
  // ---- teardown ---------------------------------------------------------------------------
  fence_before_sync();
  __syncthreads();
  if (warp == kMmaWarp) {
    fence_after_sync();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------
// self-test: D[128,128] = A[128,Kp] W[128,Kp]^T through the same TMEM / descriptor / barrier
// machinery (A split and written to TMEM by the row threads, W packed to core matrices).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
tc_selftest_pack(const float* __restrict__ Wm, int Kp, __half* __restrict__ hi, __half* __restrict__ lo) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < kN * Kp; i += gridDim.x * blockDim.x) {
    const int n = i / Kp, k = i - n * Kp;
    const float v = Wm[i];
    const __half h = __float2half_rn(v);
    hi[core_offset(n, k, kN)] = h;
    lo[core_offset(n, k, kN)] = __float2half_rn(v - __half2float(h));
  }
}

__global__ void __launch_bounds__(160, 1)
tc_selftest_kernel(const float* __restrict__ A, const __half* __restrict__ whi,
                   const __half* __restrict__ wlo, int Kp, float* __restrict__ Dout) {
  SRCV_DYNAMIC_SMEM_ALIGNED(uint8_t, smem, 1024);
  __shared__ uint32_t s_tmem_base;
  __shared__ uint64_t bar_a, bar_d;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t wbytes = (uint32_t)kN * Kp * 2;
  for (uint32_t i = tid; i < wbytes / 16; i += blockDim.x) {
    reinterpret_cast<uint4*>(smem)[i] = __ldg(reinterpret_cast<const uint4*>(whi) + i);
    reinterpret_cast<uint4*>(smem + wbytes)[i] = __ldg(reinterpret_cast<const uint4*>(wlo) + i);
  }
  if (tid == 0) { mbar_init(&bar_a, 128); mbar_init(&bar_d, 1); mbar_fence_init(); }
  if (warp == 4) tmem_alloc(&s_tmem_base, kTmemCols);
  fence_proxy_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = s_tmem_base;
  const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
  const uint32_t col_lo = Kp / 2, col_d = 256;
  if (warp < 4) {
    const float* a = A + (size_t)tid * Kp;
    for (int c = 0; c < Kp / 2; c += 8) {
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) split_pack(a[2 * (c + j)], a[2 * (c + j) + 1], hi[j], lo[j]);
      st_x8(lane_base + c, hi);
      st_x8(lane_base + col_lo + c, lo);
    }
    wait_st();
    fence_before_sync();
    mbar_arrive(&bar_a);
    mbar_wait(&bar_d, 0);
    fence_after_sync();
    for (int c0 = 0; c0 < kN; c0 += 32) {
      uint32_t r[32];
      ld_x32(lane_base + col_d + c0, r);
      wait_ld();
#pragma unroll
      for (int j = 0; j < 32; ++j) Dout[(size_t)tid * kN + c0 + j] = __uint_as_float(r[j]);
    }
  } else {
    mbar_wait(&bar_a, 0);
    fence_after_sync();
    if (lane == 0) {
      constexpr uint32_t idesc = idesc_f16_f32(kRows, kN);
      constexpr uint32_t kLbo = kN * 16, kSbo = 128;
      const uint32_t sb = smem_u32(smem);
      for (int ks = 0; ks < Kp / 16; ++ks) {
        const uint64_t bhi = smem_desc(sb + ks * 2 * kLbo, kLbo, kSbo);
        const uint64_t blo = smem_desc(sb + wbytes + ks * 2 * kLbo, kLbo, kSbo);
        mma_ts(tmem_base + col_d, tmem_base + ks * 8, bhi, idesc, ks > 0 ? 1u : 0u);
        mma_ts(tmem_base + col_d, tmem_base + ks * 8, blo, idesc, 1u);
        mma_ts(tmem_base + col_d, tmem_base + col_lo + ks * 8, bhi, idesc, 1u);
      }
      mma_commit(&bar_d);
    }
    __syncwarp();
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 4) { fence_after_sync(); tmem_dealloc(tmem_base, kTmemCols); }
}

}  // namespace

bool mlp_tc_supported(const srcv_shape& s, const srcv_mlp_weights& w) {
  return s.K == kViews && s.C == kC && w.hidden1 == kN && w.hidden2 == kN &&
         (long long)s.H * s.W < (1ll << 26);
}

static size_t image_bytes_aligned() { return (kImageBytes + 255) & ~(size_t)255; }
// workspace tail of the tensor-core variant: [weight image | per-frame pose bias (B x 128 floats)]
size_t mlp_tc_extra_bytes(const srcv_shape& s) { return image_bytes_aligned() + (((size_t)s.B * kN * 4 + 255) & ~(size_t)255); }
size_t mlp_tc_image_bytes() { return kImageBytes; }

cudaError_t launch_mlp_tc_pack(const srcv_mlp_weights& w, void* image, cudaStream_t stream) {
  SRCV_LAUNCH(tc_pack_kernel, 64, 256, 0, stream, w, reinterpret_cast<uint8_t*>(image));
  note_launch();
  return cudaGetLastError();
}

cudaError_t launch_mlp_tc(const srcv_shape& s, const float* cur, const Workspace& ws,
                          const float* planes, bool per_pixel, const srcv_mlp_weights& w, float* cost,
                          float* lowest, uint8_t* mask, cudaStream_t stream) {
  // the caller may hand over an image packed earlier (srcv_mlp_pack_weights): nothing to do per call
  const uint8_t* image = reinterpret_cast<const uint8_t*>(w.packed_image);
  cudaError_t err = cudaSuccess;
  if (image == nullptr) {
    err = launch_mlp_tc_pack(w, ws.extra, stream);
    if (err != cudaSuccess) return err;
    image = reinterpret_cast<const uint8_t*>(ws.extra);
  }
  // the 21 pose measures enter layer 1 as a per-frame bias (views were written by the prep pass)
  float* frame_bias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ws.extra) + image_bytes_aligned());
  SRCV_LAUNCH(tc_frame_bias_kernel, s.B, kN, 0, stream, w, ws.views, frame_bias);
  note_launch();
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int tiles_x = (s.W + kTileW - 1) / kTileW, tiles_y = (s.H + kTileH - 1) / kTileH;
  const long long num_tiles = (long long)s.B * ((s.D + kTileD - 1) / kTileD) * tiles_x * tiles_y;
  // tile ids are 32-bit, and the kernel's multiply-high division by the plane-chunk count is exact below 2^32 / nd
  if (num_tiles >= (1ll << 31) || num_tiles * ((s.D + kTileD - 1) / kTileD) >= (1ll << 32)) return cudaErrorInvalidValue;
  const int grid = (int)(num_tiles < sms ? num_tiles : sms);
  const float4* src4 = reinterpret_cast<const float4*>(ws.src_c4);
  const float4* cur4 = reinterpret_cast<const float4*>(ws.cur_c4);
  (void)cur;
#define SRCV_TC_LAUNCH(PP, TW_, TH_)                                                                  \
  do {                                                                                                \
    err = cudaFuncSetAttribute(mlp_tc_kernel<PP, TW_, TH_>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                               (int)kSmemBytes);                                                      \
    if (err != cudaSuccess) return err;                                                               \
    SRCV_LAUNCH((mlp_tc_kernel<PP, TW_, TH_>), grid, kThreads, kSmemBytes, stream,                    \
                s, cur4, src4, ws.views, ws.frames, planes, image, frame_bias, cost, mask,            \
                (unsigned)num_tiles);                                                                 \
  } while (0)
#define SRCV_TC_SIZES(PP)                                                   \
  if (s.W == 160 && s.H == 120) SRCV_TC_LAUNCH(PP, 160, 120);               \
  else if (s.W == 128 && s.H == 96) SRCV_TC_LAUNCH(PP, 128, 96);            \
  else SRCV_TC_LAUNCH(PP, 0, 0)
  if (per_pixel) { SRCV_TC_SIZES(true); } else { SRCV_TC_SIZES(false); }
#undef SRCV_TC_SIZES
#undef SRCV_TC_LAUNCH
  note_launch();
  err = cudaGetLastError();
  if (err != cudaSuccess) return err;
  if (lowest) err = launch_argmax(s, cost, planes, per_pixel, lowest, stream);
  return err;
}

#ifdef SRCV_TC_TIMELINE
}  // namespace srcv
extern "C" int32_t srcv_debug_read_timeline(long long* host, int32_t n) {
  const size_t bytes = sizeof(long long) * (size_t)(n < srcv::kTlTiles * srcv::kTlEvents ? n : srcv::kTlTiles * srcv::kTlEvents);
  return (int32_t)cudaMemcpyFromSymbol(host, srcv::g_timeline, bytes);
}
namespace srcv {
#endif

// D (128 x 128) = A (128 x Kp) W^T (128 x Kp), Kp a multiple of 16 and <= 256; `scratch`
// needs 2 * 128 * Kp halves.  Device pointers; test hook for tests/test_gpu_tc.py.
cudaError_t launch_tc_selftest(const float* A, const float* Wm, int Kp, float* Dout, void* scratch,
                               cudaStream_t stream) {
  if (Kp % 16 != 0 || Kp <= 0 || Kp > 256) return cudaErrorInvalidValue;
  __half* hi = reinterpret_cast<__half*>(scratch);
  __half* lo = hi + (size_t)kN * Kp;
  SRCV_LAUNCH(tc_selftest_pack, 32, 256, 0, stream, Wm, Kp, hi, lo);
  note_launch();
  const size_t smem = (size_t)2 * kN * Kp * 2;
  cudaError_t err = cudaFuncSetAttribute(tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) return err;
  SRCV_LAUNCH(tc_selftest_kernel, 1, 160, smem, stream, A, hi, lo, Kp, Dout);
  note_launch();
  return cudaGetLastError();
}

}  // namespace srcv
