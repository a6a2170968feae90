// Metadata-MLP plane-sweep volume on the 5th-generation tensor cores (tcgen05).
//
// Replaces FeatureVolumeManager / FastFeatureVolumeManager.build_cost_volume
// (reference modules/cost_volume.py:451-736, :967-1164) for the hero layout
// K = 7 source views, C = 16 channels, MLP 202 -> 128 -> 128 -> 1.
//
// One persistent CTA per SM walks over row tiles; a tile = 128 rows = a 16 x 2 patch of
// pixels at four consecutive depth planes.  Sixteen "row worker" warps (four threads per row) and one
// MMA-issuing thread run this per-tile pipeline:
//   1. the workers project, gather (chunk-planar copies of the features, csrc/srcv_prep.cu)
//      and build the row's 202 metadata channels in registers, split every value into an
//      fp16 (hi, lo) pair and write them straight into TENSOR MEMORY as the A operand
//      (tcgen05.st) — the (B,F,H,W) / (B*D,H,W,F) tensors the reference materialises exist
//      only as 208 TMEM columns per row;
//   2. layer 1 = 13 x 3 tcgen05.mma (A from TMEM, weights from shared memory, fp32
//      accumulator in TMEM):  A_hi W_hi + A_hi W_lo + A_lo W_hi;
//   3. the workers read the accumulator (each its 32 columns), add bias, LeakyReLU, split to
//      fp16 (hi, lo) again and write the layer-2 A operand back to TMEM;
//   4. layer 2 = 8 x 3 tcgen05.mma into the same accumulator columns;
//   5. the workers apply bias + LeakyReLU and the 128 -> 1 layer as partial dots, reduced in
//      a fixed order through shared memory, and store the cost.
// Stages are chained by mbarriers (tcgen05.commit for MMA completion).  A worker builds the
// first K block of tile t+1 while tile t's layer-1 MMAs run and the second one while its
// layer-2 MMAs run, so the tensor pipe and the SIMT pipes overlap.
// Weights (both layers, hi and lo, 168 KB) stay resident in shared memory for the
// CTA's lifetime, laid out as K-major no-swizzle core matrices by the pack kernel.
//
// The K order of layer 1 is OURS (the pack kernel permutes W1's columns to match):
//   per view k (26 channels): 16 warped | mask | z' | dot | ray angle | n_src (3) | comb | r | t
//   tail (20 + 6 pad):        16 reference features | plane depth | n_cur (3) | zeros
// i.e. every worker thread writes 2 x 26 = 52 consecutive K positions; the first of its
// two blocks is computed before it waits for the A columns to be free, so producing tile
// t+1 overlaps the layer-1 MMAs of tile t.
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
constexpr int kBlk = kC + 10;            // channels per view block (26)
constexpr int kK1 = 208;                 // 7*26 + 20 + 6 pad  (13 k-steps of 16)
constexpr int kK2 = 128;
constexpr int kF = kC * (kViews + 1) + 10 * kViews + 4;  // 202

// TMEM columns (32-bit): two K elements per column
constexpr uint32_t kColA1Hi = 0, kColA1Lo = kK1 / 2, kColD = kK1, kColA2Hi = kK1 + kN,
                   kColA2Lo = kK1 + kN + kK2 / 2, kTmemCols = 512;
static_assert(kColA2Lo + kK2 / 2 <= kTmemCols, "TMEM budget");

// Weights are stored x16 (exact) so that the lo halves of typical |w| ~ 0.05 weights stay
// out of the fp16 subnormal range; the epilogues fold the 1/16 into their bias FMA.
constexpr float kWScale = 16.0f, kWUnscale = 1.0f / 16.0f;

// 16 "row worker" warps: four threads per row (quarter q = warp / 4).  Every worker is
// producer AND epilogue of its rows — it builds K blocks {2q, 2q+1} of the A operand and
// post-processes accumulator columns [32q, 32q+32) — so the epilogues, which sit on the
// per-tile critical path MMA1 -> epi1 -> MMA2 -> epi2, get all 16 warps' issue slots
// instead of competing with separate producer warps.  Warp 16 issues the MMAs; warps
// 17-19 only pad the block to a 4-warp allocation unit (20 warps x 96 registers).
constexpr int kWorkWarps = 16, kMmaWarp = 16;
constexpr int kThreads = 20 * 32;
constexpr int kWorkers = kWorkWarps * 32;

// shared memory image (bytes)
constexpr uint32_t kW1Bytes = kN * kK1 * 2, kW2Bytes = kN * kK2 * 2;   // one of (hi, lo)
constexpr uint32_t kOffW1Hi = 0, kOffW1Lo = kW1Bytes, kOffW2Hi = 2 * kW1Bytes,
                   kOffW2Lo = 2 * kW1Bytes + kW2Bytes, kOffVec = 2 * kW1Bytes + 2 * kW2Bytes;
constexpr uint32_t kVecFloats = 3 * kN + 4;   // b1 | b2 | w3 | b3
constexpr uint32_t kOffBar = kOffVec + kVecFloats * 4;
constexpr uint32_t kOffFlag = kOffBar + 8 * 8;            // 8 mbarrier slots
constexpr uint32_t kOffPart = kOffFlag + 2 * 4 * kRows;   // mask bits [parity][quarter][row]
constexpr uint32_t kSmemBytes = kOffPart + 2 * 4 * kRows * 4;  // layer-3 partial dots [parity][quarter][row]
// image = [W1hi | W1lo | W2hi | W2lo | b1 b2 w3 b3] exactly as it sits in shared memory
constexpr uint32_t kImageBytes = kOffBar;

// K-major no-swizzle core-matrix offset (in halves) of element (n, k) of an N x Kp operand
__host__ __device__ inline uint32_t core_offset(int n, int k, int N) {
  return (uint32_t)(k >> 3) * (uint32_t)(N * 8) + (uint32_t)(n >> 3) * 64u + (uint32_t)(n & 7) * 8u + (uint32_t)(k & 7);
}

// reference channel index (modules/cost_volume.py:698-723 order) of OUR layer-1 K position
__host__ __device__ inline int ref_channel(int kk) {
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
      case 7: return base + 7 * kViews + 4 + k;          // comb
      case 8: return base + 8 * kViews + 4 + k;          // r
      default: return base + 9 * kViews + 4 + k;         // t
    }
  }
  const int j = kk - kViews * kBlk;
  if (j < kC) return kViews * kC + j;                    // reference-frame features
  if (j == kC) return kC * (kViews + 1) + 2 * kViews;    // plane depth
  if (j < kC + 4) return kC * (kViews + 1) + 4 * kViews + 1 + (j - kC - 1);  // n_cur
  return -1;                                             // pad
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
      const float v = kWScale * (f >= 0 ? w.w1[(size_t)n * kF + f] : 0.f);
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
      const int q = i - n1 - n2;
      float v = 0.f;
      if (q < kN) v = w.b1[q];
      else if (q < 2 * kN) v = w.b2[q - kN];
      else if (q < 3 * kN) v = w.w3[q - 2 * kN];
      else if (q == 3 * kN) v = w.b3[0];
      vec[q] = v;
    }
  }
}

// 26 values of one K block -> 13 packed (hi, lo) column pairs
__device__ __forceinline__ void split_block(const float (&v)[kBlk], uint32_t (&hi)[13], uint32_t (&lo)[13]) {
#pragma unroll
  for (int i = 0; i < 13; ++i) split_pack(v[2 * i], v[2 * i + 1], hi[i], lo[i]);
}
// 13 packed columns of one K block -> TMEM (hi and lo regions)
__device__ __forceinline__ void store_block(uint32_t tbase_lane, uint32_t col, const uint32_t (&hi)[13],
                                            const uint32_t (&lo)[13]) {
  st_x8(tbase_lane + kColA1Hi + col, hi);
  st_x4(tbase_lane + kColA1Hi + col + 8, hi + 8);
  st_x1(tbase_lane + kColA1Hi + col + 12, hi[12]);
  st_x8(tbase_lane + kColA1Lo + col, lo);
  st_x4(tbase_lane + kColA1Lo + col + 8, lo + 8);
  st_x1(tbase_lane + kColA1Lo + col + 12, lo[12]);
}

// Per-row quantities shared by the view blocks of one tile.
struct RowCtx {
  float dval, dxc, dyc;      // plane depth, centred pixel
  float X, Y, Z;             // back-projected point
  float cxn, cyn, czn;       // n_cur / max(|n_cur|, eps_cos)  (cosine_similarity operand)
  float4 cur4[4];            // reference-frame features of the pixel
};

// One source view of one row: project, gather, metadata (channel order of the K block).
// HWC != 0: compile-time map size, every gather address is base + immediate.
template <int TW, int HWC>
__device__ __forceinline__ unsigned view_block(const RowCtx& rc, const ViewParams& vp,
                                               const float4* __restrict__ view4, int Wrt, int H, int HWrt,
                                               const Centre& ctr, float (&v)[kBlk]) {
  const int W = TW ? TW : Wrt, HW = HWC ? HWC : HWrt;
  float ax, ay, az, px, py, zp;
  homography_point(vp.a0, rc.dxc, rc.dyc, ax, ay, az);
  project_point(rc.dval, ax, ay, az, vp.t[0], vp.t[1], vp.t[2], px, py, zp);
  Taps tp;
  bilinear_taps(px, py, W, H, ctr, tp);
  const float gx = 1.0f - tp.fx, gy = 1.0f - tp.fy;
  const float wgt[4] = {gx * gy, tp.fx * gy, gx * tp.fy, tp.fx * tp.fy};
  const int off[4] = {0, 1, W, W + 1};
  const float4* q = view4 + (tp.y0 * W + tp.x0);
  // features are sampled even for points behind the camera (only the dot is masked,
  // reference modules/cost_volume.py:590-623); padding taps contribute zeros
  // two taps (8 vector loads) in flight at a time keeps the worker inside its register budget
  const bool interior = __all_sync(0xffffffffu, tp.valid == 15u);
#ifdef SRCV_TC_CLAMPED_TAPS
  const int cxa = min(max(tp.x0, 0), W - 1), cxb = min(max(tp.x0 + 1, 0), W - 1);
  const int cya = min(max(tp.y0, 0), H - 1) * W, cyb = min(max(tp.y0 + 1, 0), H - 1) * W;
  const int coff[4] = {cya + cxa, cya + cxb, cyb + cxa, cyb + cxb};
#endif
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    float4 f[2][4];
#pragma unroll
    for (int tt = 0; tt < 2; ++tt) {
      const int tap = 2 * half + tt;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (interior) f[tt][j] = __ldg(q + off[tap] + (size_t)j * HW);
#ifdef SRCV_TC_CLAMPED_TAPS
        else {
          // Round-2 experiment (off by default): branch-free borders.  Every tap is loaded from an
          // in-range (clamped) texel and padding taps are zeroed by selects on the VALUES (not on
          // the weights: a non-finite texel must not leak through 0 * inf) — instead of a
          // BSSY / BRA / CS2R region around every conditional load.
          const float4 tv = __ldg(view4 + coff[tap] + (size_t)j * HW);
          const bool on = ((tp.valid >> tap) & 1u) != 0u;
          f[tt][j] = make_float4(on ? tv.x : 0.f, on ? tv.y : 0.f, on ? tv.z : 0.f, on ? tv.w : 0.f);
        }
#else
        else f[tt][j] = ((tp.valid >> tap) & 1u) ? __ldg(q + off[tap] + (size_t)j * HW)
                                                 : make_float4(0.f, 0.f, 0.f, 0.f);
#endif
      }
    }
    const float wa = wgt[2 * half], wb = wgt[2 * half + 1];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (half == 0) {
        v[4 * j + 0] = fmaf(wa, f[0][j].x, wb * f[1][j].x);
        v[4 * j + 1] = fmaf(wa, f[0][j].y, wb * f[1][j].y);
        v[4 * j + 2] = fmaf(wa, f[0][j].z, wb * f[1][j].z);
        v[4 * j + 3] = fmaf(wa, f[0][j].w, wb * f[1][j].w);
      } else {
        v[4 * j + 0] = fmaf(wa, f[0][j].x, fmaf(wb, f[1][j].x, v[4 * j + 0]));
        v[4 * j + 1] = fmaf(wa, f[0][j].y, fmaf(wb, f[1][j].y, v[4 * j + 1]));
        v[4 * j + 2] = fmaf(wa, f[0][j].z, fmaf(wb, f[1][j].z, v[4 * j + 2]));
        v[4 * j + 3] = fmaf(wa, f[0][j].w, fmaf(wb, f[1][j].w, v[4 * j + 3]));
      }
    }
  }
  float dot = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    dot = fmaf(v[4 * j], rc.cur4[j].x, fmaf(v[4 * j + 1], rc.cur4[j].y,
          fmaf(v[4 * j + 2], rc.cur4[j].z, fmaf(v[4 * j + 3], rc.cur4[j].w, dot))));
  const float mk = zp > 0.0f ? 1.0f : 0.0f;
  // n_src = (X - centre_k)/max(|.|, 1e-12) ; ray angle = cosine_similarity(n_cur, n_src, eps 1e-5)
  const float sx0 = rc.X - vp.centre[0], sy0 = rc.Y - vp.centre[1], sz0 = rc.Z - vp.centre[2];
  const float is = inv_norm(fmaf(sx0, sx0, fmaf(sy0, sy0, sz0 * sz0)), kEpsNorm);
  const float sx = sx0 * is, sy = sy0 * is, sz = sz0 * is;
  const float i2 = inv_norm(fmaf(sx, sx, fmaf(sy, sy, sz * sz)), kEpsCos);
  v[kC + 0] = mk;
  v[kC + 1] = zp;
  v[kC + 2] = dot * mk;
  v[kC + 3] = fmaf(rc.cxn, sx * i2, fmaf(rc.cyn, sy * i2, rc.czn * (sz * i2)));
  v[kC + 4] = sx; v[kC + 5] = sy; v[kC + 6] = sz;
  v[kC + 7] = vp.comb; v[kC + 8] = vp.rmeas; v[kC + 9] = vp.tmeas;
  unsigned bits = 0;
  if (zp > 0.0f) bits |= 1u;
  if (in_mask_bounds(px, py, W, H, ctr)) bits |= 2u;
  return bits;
}

// issue D (+)= A_hi W_hi + A_hi W_lo + A_lo W_hi over KSTEPS K steps of 16
template <int KSTEPS>
__device__ __forceinline__ void issue_layer(uint32_t tmem_base, uint32_t col_hi, uint32_t col_lo,
                                            uint32_t smem_hi, uint32_t smem_lo) {
  constexpr uint32_t idesc = idesc_f16_f32(kRows, kN);
  constexpr uint32_t kLbo = kN * 16, kSbo = 128, kStepBytes = 2 * kLbo;  // two 8-wide K chunks per MMA
  // only the 14-bit start-address field changes from step to step
  const uint64_t bhi0 = smem_desc(smem_hi, kLbo, kSbo), blo0 = smem_desc(smem_lo, kLbo, kSbo);
  const uint32_t d = tmem_base + kColD;
  uint64_t bhi = bhi0, blo = blo0;
  uint32_t ahi = tmem_base + col_hi, alo = tmem_base + col_lo;
#pragma unroll 1
  for (int ks = 0; ks < KSTEPS; ++ks) {
    mma_ts(d, ahi, bhi, idesc, ks > 0 ? 1u : 0u);
    mma_ts(d, ahi, blo, idesc, 1u);
    mma_ts(d, alo, bhi, idesc, 1u);
    bhi += kStepBytes >> 4; blo += kStepBytes >> 4;
    ahi += 8; alo += 8;
  }
}

struct TileCoord { int b, d0, x0, y0; };

// tile id runs plane-chunk fastest, then pixel patch, then frame
__device__ __forceinline__ TileCoord tile_coord(long long id, int D, int tiles_x, int tiles_xy) {
  TileCoord t;
  const int nd = (D + kTileD - 1) / kTileD;
#ifdef SRCV_TC_TILE32
  // Round-2 experiment (off by default): tile ids fit 32 bits for every supported shape (the
  // launcher refuses >= 2^31 tiles), and the 64-bit div/mod pairs cost ~120 instructions per
  // thread and tile (two make_tile_row per tile) — 7 % of the worker's instruction budget.
  const unsigned uid = (unsigned)id, und = (unsigned)nd, uxy = (unsigned)tiles_xy;
  t.d0 = (int)(uid % und) * kTileD;
  const unsigned r = uid / und;
  const int txy = (int)(r % uxy);
  t.b = (int)(r / uxy);
#else
  t.d0 = (int)(id % nd) * kTileD;
  const long long r = id / nd;
  const int txy = (int)(r % tiles_xy);
  t.b = (int)(r / tiles_xy);
#endif
  t.x0 = (txy % tiles_x) * kTileW;
  t.y0 = (txy / tiles_x) * kTileH;
  return t;
}

// Row context of one tile for one worker thread (what the K blocks and the final store need).
struct TileRow {
  RowCtx rc;
  float cx, cy, cz;   // n_cur
  int b, d, p;        // frame, plane, pixel index (clamped into the map)
  bool active;        // row maps to a real pixel (partial tiles at the right/bottom border)
};

template <bool PER_PIXEL>
__device__ __forceinline__ void make_tile_row(long long id, int row, int W, int H, int HW, int D,
                                              int tiles_x, int tiles_xy, const Centre& ctr,
                                              const FrameParams* __restrict__ frames,
                                              const float* __restrict__ planes, TileRow& tr) {
  const TileCoord t = tile_coord(id, D, tiles_x, tiles_xy);
  // row -> (plane-in-chunk = row / 32, pixel of the 16 x 2 patch = row % 32)
  const int rx = row & (kTileW - 1), ry = (row >> 4) & (kTileH - 1), dd = row >> 5;
  tr.b = t.b;
  tr.d = min(t.d0 + dd, D - 1);
  tr.active = (t.x0 + rx < W) && (t.y0 + ry < H) && (t.d0 + dd < D);
  const int ox = min(t.x0 + rx, W - 1), oy = min(t.y0 + ry, H - 1);
  tr.p = oy * W + ox;
  const float pxc = (float)ox + 0.5f, pyc = (float)oy + 0.5f;
  RowCtx& rc = tr.rc;
  rc.dval = PER_PIXEL ? __ldg(planes + ((size_t)t.b * D + tr.d) * HW + tr.p) : __ldg(planes + t.b * D + tr.d);
  rc.dxc = pxc - ctr.half_w;
  rc.dyc = pyc - ctr.half_h;
  // rays: X = d * (invK3 p); n_cur = X / max(|X|, 1e-12)
  const FrameParams& fp = frames[t.b];
  const float rxv = fmaf(fp.invK[0], pxc, fmaf(fp.invK[1], pyc, fp.invK[2]));
  const float ryv = fmaf(fp.invK[3], pxc, fmaf(fp.invK[4], pyc, fp.invK[5]));
  const float rzv = fmaf(fp.invK[6], pxc, fmaf(fp.invK[7], pyc, fp.invK[8]));
  rc.X = rc.dval * rxv; rc.Y = rc.dval * ryv; rc.Z = rc.dval * rzv;
  const float ic = inv_norm(fmaf(rc.X, rc.X, fmaf(rc.Y, rc.Y, rc.Z * rc.Z)), kEpsNorm);
  tr.cx = rc.X * ic; tr.cy = rc.Y * ic; tr.cz = rc.Z * ic;
  const float i1 = inv_norm(fmaf(tr.cx, tr.cx, fmaf(tr.cy, tr.cy, tr.cz * tr.cz)), kEpsCos);
  rc.cxn = tr.cx * i1; rc.cyn = tr.cy * i1; rc.czn = tr.cz * i1;
}

// K block `blk` (0..6: source view, 7: view-independent tail) of a row -> packed (hi, lo)
template <int TW, int HWC>
__device__ __forceinline__ unsigned build_block(TileRow& tr, int blk, const float4* __restrict__ cur4g,
                                                const float4* __restrict__ src4,
                                                const ViewParams* __restrict__ views, int W, int H, int HW,
                                                const Centre& ctr, uint32_t (&hi)[13], uint32_t (&lo)[13]) {
  float v[kBlk];
  unsigned bits = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) tr.rc.cur4[j] = __ldg(cur4g + ((size_t)tr.b * 4 + j) * HW + tr.p);
  if (blk < kViews) {
    bits = view_block<TW, HWC>(tr.rc, views[tr.b * kViews + blk],
                               src4 + ((size_t)(tr.b * kViews + blk) * 4) * HW, W, H, HW, ctr, v);
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[4 * j] = tr.rc.cur4[j].x; v[4 * j + 1] = tr.rc.cur4[j].y;
      v[4 * j + 2] = tr.rc.cur4[j].z; v[4 * j + 3] = tr.rc.cur4[j].w;
    }
    v[kC] = tr.rc.dval; v[kC + 1] = tr.cx; v[kC + 2] = tr.cy; v[kC + 3] = tr.cz;
#pragma unroll
    for (int j = kC + 4; j < kBlk; ++j) v[j] = 0.f;
  }
  split_block(v, hi, lo);
  return bits;
}

template <bool PER_PIXEL, int TW, int TH>
__global__ void __launch_bounds__(kThreads, 1)
mlp_tc_kernel(srcv_shape s, const float4* __restrict__ cur4g, const float4* __restrict__ src4,
              const ViewParams* __restrict__ views, const FrameParams* __restrict__ frames,
              const float* __restrict__ planes, const uint8_t* __restrict__ image,
              float* __restrict__ cost, uint8_t* __restrict__ mask_out, long long num_tiles) {
  SRCV_DYNAMIC_SMEM_ALIGNED(uint8_t, smem, 1024);
  __shared__ uint32_t s_tmem_base;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* bar_a1_full = bars + 0;   // workers -> MMA: A1 of a tile is in TMEM      (512 arrivals)
  uint64_t* bar_mma1 = bars + 1;      // layer-1 MMAs done                           (commit)
  uint64_t* bar_a2_full = bars + 2;   // workers -> MMA: A2 written, D1 consumed      (512 arrivals)
  uint64_t* bar_mma2 = bars + 3;      // layer-2 MMAs done                           (commit)
  uint64_t* bar_d_free = bars + 4;    // workers done reading the accumulator        (512 arrivals)
  uint64_t* bar_e2 = bars + 5;        // layer-3 partial dots are in shared memory   (512 arrivals)
  uint8_t* sflag = smem + kOffFlag;
  float* spart = reinterpret_cast<float*>(smem + kOffPart);
  const float* svec = reinterpret_cast<const float*>(smem + kOffVec);

  // SRCV_TC_UNIFORM_WARP (round-2 experiment, off by default): taking the warp index through a
  // shuffle broadcast tells ptxas the role branches below are warp-uniform, so the global-memory
  // descriptor stays in uniform registers instead of being re-materialised with two R2UR before
  // every LDG of the worker loop (static SASS of the 160x120 kernel: 441 -> 59 R2UR, 3784 -> 3464
  // instructions, spill stores 82 -> 54 bytes).  Same value on every lane either way; not yet
  // run on a GPU, hence not the default.
#ifdef SRCV_TC_UNIFORM_WARP
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
#else
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#endif
  const int W = TW ? TW : s.W, H = TH ? TH : s.H, HW = W * H, D = s.D;
  constexpr int HWC = TW * TH;
  const int tiles_x = (W + kTileW - 1) / kTileW, tiles_xy = tiles_x * ((H + kTileH - 1) / kTileH);

  // ---- one-time setup ----------------------------------------------------------------
  uint64_t* bar_img = bars + 6;       // weight image landed in shared memory (bulk copies)
  if (tid == 0) {
    mbar_init(bar_img, 1);
    mbar_init(bar_a1_full, kWorkers);
    mbar_init(bar_mma1, 1);
    mbar_init(bar_a2_full, kWorkers);
    mbar_init(bar_mma2, 1);
    mbar_init(bar_d_free, kWorkers);
    mbar_init(bar_e2, kWorkers);
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
  // ONE frame, whose source features (39 MB) stay L2-resident.  (Measured alternatives: one
  // contiguous range per CTA over the whole batch 2.98 ms, per frame 3.02 ms vs 2.71 ms for
  // this interleaving at B = 8.)  Local index j -> tile id.
  const long long tile_begin = 0;
  const long long tile_end = (num_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
  auto tile_id = [&](long long j) { return (long long)blockIdx.x + j * gridDim.x; };

  if (warp < kWorkWarps) {
    // =============================== row workers =========================================
    const int row = tid & (kRows - 1), q = tid >> 7;
    const Centre ctr(W, H);
    const int blk0 = 2 * q, blk1 = 2 * q + 1;   // q = 3: view 6 and the tail block (7)
    TileRow cur_row, nxt_row;
    uint32_t hi[13], lo[13];
    long long id = tile_begin;
    if (id < tile_end) {
      // prologue: the first tile's A operand (its columns are free)
      make_tile_row<PER_PIXEL>(tile_id(id), row, W, H, HW, D, tiles_x, tiles_xy, ctr, frames, planes, cur_row);
      unsigned bits = build_block<TW, HWC>(cur_row, blk0, cur4g, src4, views, W, H, HW, ctr, hi, lo);
      store_block(lane_base, (uint32_t)(13 * blk0), hi, lo);
      bits |= build_block<TW, HWC>(cur_row, blk1, cur4g, src4, views, W, H, HW, ctr, hi, lo);
      store_block(lane_base, (uint32_t)(13 * blk1), hi, lo);
      sflag[(0 * 4 + q) * kRows + row] = (uint8_t)bits;
      wait_st();
      fence_before_sync();
      mbar_arrive(bar_a1_full);
    }
    int it = 0;
    for (; id < tile_end; ++id, ++it) {
      const long long nid = id + 1;
      const bool has_next = nid < tile_end;
      unsigned nbits = 0;
      if (has_next) {
        // first K block of the NEXT tile, built while this tile's layer-1 MMAs run
        make_tile_row<PER_PIXEL>(tile_id(nid), row, W, H, HW, D, tiles_x, tiles_xy, ctr, frames, planes, nxt_row);
        nbits = build_block<TW, HWC>(nxt_row, blk0, cur4g, src4, views, W, H, HW, ctr, hi, lo);
      }
      // ---- layer-1 epilogue on columns [32q, 32q+32): bias + LeakyReLU, (hi, lo), A2 -> TMEM
      mbar_wait(bar_mma1, it & 1);
      fence_after_sync();
#pragma unroll
      for (int c0 = 32 * q; c0 < 32 * q + 32; c0 += 16) {
        uint32_t r[16];
        ld_x16(lane_base + kColD + c0, r);
        wait_ld();
        uint32_t ehi[8], elo[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float2 bb = *reinterpret_cast<const float2*>(svec + c0 + 2 * j);
          const float a = leaky(fmaf(__uint_as_float(r[2 * j]), kWUnscale, bb.x));
          const float b = leaky(fmaf(__uint_as_float(r[2 * j + 1]), kWUnscale, bb.y));
          split_pack(a, b, ehi[j], elo[j]);
        }
        st_x8(lane_base + kColA2Hi + c0 / 2, ehi);
        st_x8(lane_base + kColA2Lo + c0 / 2, elo);
      }
      wait_st();
      fence_before_sync();
      mbar_arrive(bar_a2_full);
      // ---- the rest of the next tile's A operand (A1 is free: layer 1 of this tile is done)
      if (has_next) {
        store_block(lane_base, (uint32_t)(13 * blk0), hi, lo);
        nbits |= build_block<TW, HWC>(nxt_row, blk1, cur4g, src4, views, W, H, HW, ctr, hi, lo);
        store_block(lane_base, (uint32_t)(13 * blk1), hi, lo);
        sflag[(((it + 1) & 1) * 4 + q) * kRows + row] = (uint8_t)nbits;
        wait_st();
        fence_before_sync();
        mbar_arrive(bar_a1_full);
      }
      // ---- layer-2 epilogue on columns [32q, 32q+32): bias + LeakyReLU, 128 -> 1 partial dot
      mbar_wait(bar_mma2, it & 1);
      fence_after_sync();
      float acc = 0.f;
#pragma unroll
      for (int c0 = 32 * q; c0 < 32 * q + 32; c0 += 16) {
        uint32_t r[16];
        ld_x16(lane_base + kColD + c0, r);
        wait_ld();
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const float4 bb = *reinterpret_cast<const float4*>(svec + kN + c0 + j);
          const float4 ww = *reinterpret_cast<const float4*>(svec + 2 * kN + c0 + j);
          acc = fmaf(leaky(fmaf(__uint_as_float(r[j + 0]), kWUnscale, bb.x)), ww.x, acc);
          acc = fmaf(leaky(fmaf(__uint_as_float(r[j + 1]), kWUnscale, bb.y)), ww.y, acc);
          acc = fmaf(leaky(fmaf(__uint_as_float(r[j + 2]), kWUnscale, bb.z)), ww.z, acc);
          acc = fmaf(leaky(fmaf(__uint_as_float(r[j + 3]), kWUnscale, bb.w)), ww.w, acc);
        }
      }
#ifdef SRCV_TC_EARLY_FLAGS
      // Mask bits of THIS tile, read before the bar_d_free arrival.  The four quarters wrote them
      // before their bar_a1_full arrival, so they are visible since the bar_mma1 wait above; read
      // here, the read is ordered (d_free -> MMA issue -> bar_mma1 of the next tile) before the
      // write of the tile after next into the same parity slot.  Read late (below, the default
      // until this variant has run on a GPU) that write is unordered with the read: found by
      // ThreadSanitizer on the host emulation; on hardware the reader would have to stall for a
      // whole MMA + epilogue (> 3 k clocks) between two adjacent instructions to lose the race.
      unsigned tile_bits = 0;
      if (q == 0 && mask_out != nullptr && cur_row.d == D - 1) {
        const uint8_t* fl = sflag + (it & 1) * 4 * kRows + row;
        tile_bits = fl[0] | fl[kRows] | fl[2 * kRows] | fl[3 * kRows];
      }
#endif
      fence_before_sync();
      mbar_arrive(bar_d_free);
      spart[((it & 1) * 4 + q) * kRows + row] = acc;
      mbar_arrive(bar_e2);
      if (q == 0) {
        // fixed-order reduction of the four quarters (deterministic), bias, store
        mbar_wait(bar_e2, it & 1);
        const float* sp = spart + (it & 1) * 4 * kRows + row;
        const float total = ((sp[0] + sp[kRows]) + sp[2 * kRows]) + sp[3 * kRows];
        if (cur_row.active) {
          cost[((size_t)cur_row.b * D + cur_row.d) * HW + cur_row.p] = total + svec[3 * kN];
          if (mask_out != nullptr && cur_row.d == D - 1) {
#ifdef SRCV_TC_EARLY_FLAGS
            const unsigned bits = tile_bits;
#else
            const uint8_t* fl = sflag + (it & 1) * 4 * kRows + row;
            const unsigned bits = fl[0] | fl[kRows] | fl[2 * kRows] | fl[3 * kRows];
#endif
            mask_out[(size_t)cur_row.b * HW + cur_row.p] = (bits == 3u) ? 1 : 0;
          }
        }
      }
      cur_row = nxt_row;
    }
  } else if (warp == kMmaWarp) {
    // =============================== MMA issuer ==========================================
    const uint32_t sbase = smem_u32(smem);
    int it = 0;
    for (long long id = tile_begin; id < tile_end; ++id, ++it) {
      mbar_wait(bar_a1_full, it & 1);          // A1 of this tile is in TMEM
      mbar_wait(bar_d_free, (it & 1) ^ 1);     // previous tile's accumulator has been read
      fence_after_sync();
      if (lane == 0) {
        issue_layer<kK1 / 16>(tmem_base, kColA1Hi, kColA1Lo, sbase + kOffW1Hi, sbase + kOffW1Lo);
        mma_commit(bar_mma1);
      }
      __syncwarp();
      mbar_wait(bar_a2_full, it & 1);          // A2 written, accumulator columns consumed
      fence_after_sync();
      if (lane == 0) {
        issue_layer<kK2 / 16>(tmem_base, kColA2Hi, kColA2Lo, sbase + kOffW2Hi, sbase + kOffW2Lo);
        mma_commit(bar_mma2);
      }
      __syncwarp();
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

size_t mlp_tc_extra_bytes() { return (kImageBytes + 255) & ~(size_t)255; }

cudaError_t launch_mlp_tc(const srcv_shape& s, const float* cur, const Workspace& ws,
                          const float* planes, bool per_pixel, const srcv_mlp_weights& w, float* cost,
                          float* lowest, uint8_t* mask, cudaStream_t stream) {
  uint8_t* image = reinterpret_cast<uint8_t*>(ws.extra);
  SRCV_LAUNCH(tc_pack_kernel, 64, 256, 0, stream, w, image);
  note_launch();
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return err;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int tiles_x = (s.W + kTileW - 1) / kTileW, tiles_y = (s.H + kTileH - 1) / kTileH;
  const long long num_tiles = (long long)s.B * ((s.D + kTileD - 1) / kTileD) * tiles_x * tiles_y;
  if (num_tiles >= (1ll << 31)) return cudaErrorInvalidValue;   // tile ids are kept 32-bit-safe
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
                s, cur4, src4, ws.views, ws.frames, planes, image, cost, mask, num_tiles);            \
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
