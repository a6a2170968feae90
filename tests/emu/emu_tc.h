// Host emulation of the tcgen05 / TMEM / mbarrier / bulk-copy layer of csrc/srcv_tc.cuh —
// TEST INFRASTRUCTURE (see emu_cuda.h).  Same function names and argument meaning as the
// inline-PTX wrappers, with the FUNCTIONAL semantics the GPU-validated kernel relies on:
//
//   * tensor memory: 128 lanes x 512 32-bit columns per CTA (NaN-poisoned); address =
//     (lane << 16) | column; a .32x32b access of thread i touches lane (addr.lane + i % 32);
//   * tcgen05.mma.kind::f16, M = 128: D[m][n] (+)= sum_{k<16} A[m][k] B[n][k], A from TMEM (two
//     K elements per column, even K in the low half) or from shared memory, B from shared
//     memory through a K-major no-swizzle descriptor (start, LBO, SBO in 16-byte units);
//     products are exact, accumulation is fp32 in k order (the hardware's order is not
//     specified: results agree to rounding, not bit for bit);
//   * MMAs are queued by the issuing thread and EXECUTED AT tcgen05.commit, just before the
//     commit's mbarrier arrival — i.e. as late as the hardware is allowed to run them, so that
//     an operand overwritten between issue and commit corrupts the result here too;
//   * mbarrier: phase bit | pending arrivals | arrival count | pending transaction bytes packed
//     in the 64-bit word, updated with CAS; waits block on the word (std::atomic_ref::wait).
//     Acquire/release on the word gives ThreadSanitizer the happens-before edges of the real
//     protocol: a TMEM / shared-memory hand-off that is not ordered by a barrier is a data race.
//   * tcgen05.fence / wait::ld / wait::st / proxy fences: no-ops (program order on the host).
// Not modelled: asynchrony of tcgen05.st / ld (so a missing tcgen05.wait is not detected),
// cta_group::2, swizzled layouts, multicast.
#pragma once
#include "emu_cuda.h"

#include <atomic>
#include <cstdio>
#include <vector>

struct __half { _Float16 v; };
struct __half2 { __half x, y; };
inline __half __float2half_rn(float f) { return __half{(_Float16)f}; }
inline float __half2float(__half h) { return (float)h.v; }
inline float2 __half22float2(__half2 h) { return make_float2((float)h.x.v, (float)h.y.v); }
inline __half2 __floats2half2_rn(float a, float b) { return __half2{__half{(_Float16)a}, __half{(_Float16)b}}; }

namespace srcv {
namespace tc {

inline uint32_t smem_u32(const void* p) {
  return (uint32_t)(reinterpret_cast<const char*>(p) - reinterpret_cast<const char*>(::emu::dynamic_smem()));
}

// ---- mbarrier ---------------------------------------------------------------------------
namespace detail {
struct Word { uint32_t phase, pending, init; int32_t tx; };
inline Word unpack(uint64_t w) {
  return Word{(uint32_t)(w >> 63), (uint32_t)((w >> 32) & 0x7FFF), (uint32_t)((w >> 47) & 0x7FFF), (int32_t)(uint32_t)w};
}
inline uint64_t pack(const Word& s) {
  return ((uint64_t)s.phase << 63) | ((uint64_t)s.init << 47) | ((uint64_t)s.pending << 32) | (uint32_t)s.tx;
}
inline void update(uint64_t* bar, uint32_t arrivals, int32_t dtx) {
  std::atomic_ref<uint64_t> a(*bar);
  uint64_t old = a.load(std::memory_order_relaxed), nw;
  bool flipped;
  do {
    Word s = unpack(old);
    if (s.pending < arrivals) { std::fprintf(stderr, "emu: mbarrier over-arrival\n"); std::abort(); }
    s.pending -= arrivals;
    s.tx += dtx;
    flipped = (s.pending == 0 && s.tx == 0);
    if (flipped) { s.phase ^= 1u; s.pending = s.init; }
    nw = pack(s);
  } while (!a.compare_exchange_weak(old, nw, std::memory_order_acq_rel, std::memory_order_relaxed));
  if (flipped) a.notify_all();
}
}  // namespace detail

inline void mbar_init(uint64_t* bar, uint32_t count) {
  std::atomic_ref<uint64_t>(*bar).store(detail::pack(detail::Word{0, count, count, 0}), std::memory_order_release);
}
inline void mbar_fence_init() {}
inline void mbar_arrive(uint64_t* bar) { detail::update(bar, 1, 0); }
inline void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { detail::update(bar, 1, (int32_t)bytes); }
inline void mbar_wait(uint64_t* bar, uint32_t parity) {
  std::atomic_ref<uint64_t> a(*bar);
  for (;;) {
    const uint64_t w = a.load(std::memory_order_acquire);
    if ((uint32_t)(w >> 63) != (parity & 1u)) return;
    a.wait(w, std::memory_order_acquire);      // blocks until the word changes
  }
}
inline void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  std::memcpy(smem_dst, gmem_src, bytes);
  detail::update(bar, 0, -(int32_t)bytes);
}
inline void fence_proxy_async_smem() {}

// ---- tensor memory --------------------------------------------------------------------------
constexpr int kTmemLanes = 128, kTmemColumns = 512;
inline uint32_t* tmem_cell(uint32_t taddr, int lane_off, int n) {
  const uint32_t lane = (taddr >> 16) + (uint32_t)lane_off, col = taddr & 0xFFFFu;
  if (lane >= (uint32_t)kTmemLanes || col + (uint32_t)n > (uint32_t)kTmemColumns) {
    std::fprintf(stderr, "emu: TMEM access out of range (lane %u, columns %u..%u)\n", lane, col, col + n);
    std::abort();
  }
  return ::emu::t_tmem + (size_t)lane * kTmemColumns + col;
}
inline void tmem_alloc(uint32_t* smem_dst, uint32_t) { if ((threadIdx.x & 31) == 0) *smem_dst = 0u; }
inline void tmem_dealloc(uint32_t, uint32_t) {}
inline void fence_before_sync() {}
inline void fence_after_sync() {}
inline void wait_ld() {}
inline void wait_st() {}

template <int N> inline void st_n(uint32_t taddr, const uint32_t* r) {
  uint32_t* c = tmem_cell(taddr, threadIdx.x & 31, N);
  for (int j = 0; j < N; ++j) c[j] = r[j];
}
template <int N> inline void ld_n(uint32_t taddr, uint32_t* r) {
  const uint32_t* c = tmem_cell(taddr, threadIdx.x & 31, N);
  for (int j = 0; j < N; ++j) r[j] = c[j];
}
inline void st_x1(uint32_t taddr, uint32_t r0) { st_n<1>(taddr, &r0); }
inline void st_x2(uint32_t taddr, const uint32_t* r) { st_n<2>(taddr, r); }
inline void st_x4(uint32_t taddr, const uint32_t* r) { st_n<4>(taddr, r); }
inline void st_x8(uint32_t taddr, const uint32_t* r) { st_n<8>(taddr, r); }
inline void st_x16(uint32_t taddr, const uint32_t* r) { st_n<16>(taddr, r); }
inline void ld_x16(uint32_t taddr, uint32_t* r) { ld_n<16>(taddr, r); }
inline void ld_x32(uint32_t taddr, uint32_t* r) { ld_n<32>(taddr, r); }

// ---- descriptors (bit-identical to the device versions) -------------------------------------
inline uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;
}
constexpr uint32_t idesc_f16_f32(int M, int N) {
  return (1u << 4) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

// ---- MMA ------------------------------------------------------------------------------------
namespace detail {
struct Mma { uint32_t d; bool a_in_tmem; uint32_t a_tmem; uint64_t a_desc, b_desc; uint32_t idesc, acc; };
inline thread_local std::vector<Mma> t_queue;
inline float half_bits(uint16_t h) { _Float16 f; std::memcpy(&f, &h, 2); return (float)f; }
inline const char* operand(uint64_t desc, int row, int k) {       // K-major, no swizzle, 8 x 16-byte core matrices
  const uint32_t start = (uint32_t)(desc & 0x3FFFu) << 4, lbo = (uint32_t)((desc >> 16) & 0x3FFFu) << 4,
                 sbo = (uint32_t)((desc >> 32) & 0x3FFFu) << 4;
  return reinterpret_cast<const char*>(::emu::dynamic_smem()) + start + (uint32_t)(k >> 3) * lbo +
         (uint32_t)(row >> 3) * sbo + (uint32_t)(row & 7) * 16u + (uint32_t)(k & 7) * 2u;
}
inline void execute(const Mma& m) {
  const int M = (int)((m.idesc >> 24) & 0x1Fu) << 4, N = (int)((m.idesc >> 17) & 0x3Fu) << 3;
  if (M != 128) { std::fprintf(stderr, "emu: only M = 128 MMAs are modelled\n"); std::abort(); }
  for (int r = 0; r < M; ++r) {
    float a[16];
    for (int k = 0; k < 16; ++k) {
      uint16_t h;
      if (m.a_in_tmem) {
        const uint32_t w = *tmem_cell(m.a_tmem + (uint32_t)(k >> 1), r, 1);
        h = (uint16_t)(w >> (16 * (k & 1)));
      } else {
        std::memcpy(&h, operand(m.a_desc, r, k), 2);
      }
      a[k] = half_bits(h);
    }
    uint32_t* drow = tmem_cell(m.d, r, N);
    for (int n = 0; n < N; ++n) {
      float s;
      if (m.acc) std::memcpy(&s, &drow[n], 4); else s = 0.f;
      for (int k = 0; k < 16; ++k) {
        uint16_t h;
        std::memcpy(&h, operand(m.b_desc, n, k), 2);
        s += a[k] * half_bits(h);
      }
      std::memcpy(&drow[n], &s, 4);
    }
  }
}
}  // namespace detail

inline void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  detail::t_queue.push_back(detail::Mma{d_tmem, true, a_tmem, 0, b_desc, idesc, accumulate});
}
inline void mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  detail::t_queue.push_back(detail::Mma{d_tmem, false, 0, a_desc, b_desc, idesc, accumulate});
}
inline void mma_commit(uint64_t* bar) {
  for (const auto& m : detail::t_queue) detail::execute(m);
  detail::t_queue.clear();
  detail::update(bar, 1, 0);
}

// elect.sync of a converged warp: the model always picks lane 0
inline bool elect_one() { return (threadIdx.x & 31) == 0; }

template <int N> inline void reg_inc() {}
template <int N> inline void reg_dec() {}

// ---- fp32 -> (hi, lo) fp16 pair split (cvt.rn.satfinite.f16x2.f32) ------------------------------
inline uint16_t f16_sat_bits(float x) {
  if (x != x) return 0x7FFFu;
  const _Float16 h = (_Float16)fminf(fmaxf(x, -65504.f), 65504.f);
  uint16_t b;
  std::memcpy(&b, &h, 2);
  return b;
}
inline uint32_t pack_f16x2_sat(float even, float odd) {
  return (uint32_t)f16_sat_bits(even) | ((uint32_t)f16_sat_bits(odd) << 16);
}
inline void split_pack(float even, float odd, uint32_t& hi, uint32_t& lo) {
  hi = pack_f16x2_sat(even, odd);
  const float be = detail::half_bits((uint16_t)hi), bo = detail::half_bits((uint16_t)(hi >> 16));
  lo = pack_f16x2_sat(even - be, odd - bo);
}

}  // namespace tc
}  // namespace srcv
