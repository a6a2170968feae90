// Tensor-pipe probe (measurement tool, not part of the library): how long does one tcgen05.mma
// kind::f16 M=128 x N x K=16 take on this part as a function of N, of where A lives (TMEM / shared
// memory) and of concurrent tcgen05.ld / tcgen05.st traffic from other warps?  The hero kernel's
// timeline (profiles/r02_hero_timeline.json) shows ~87 clk per MMA for N=128 AND for N=64; this
// separates "A-operand read" from "tensor-pipe floor" from "TMEM port contention".
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I simplerecon_b200/csrc scripts/mma_probe.cu -o scripts/_bin/mma_probe
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "srcv_tc.cuh"

using namespace srcv::tc;

constexpr int kThreads = 32 * 17;   // 16 traffic warps + 1 MMA warp

// mode: 0 = A in TMEM, 1 = A in shared memory.  traffic: bit0 = 4 warps loop tcgen05.ld x32 pairs,
// bit1 = 12 warps loop tcgen05.st x16.  guard: how the issuing thread is selected (see below).
template <int guard>
__global__ void __launch_bounds__(kThreads, 1)
probe(int mode, int N, int iters, int traffic, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t s_base;
  __shared__ uint64_t bar;
  __shared__ volatile int stop;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 160 * 1024 / 4; i += kThreads) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;  // ones
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); stop = 0; }
  if (warp == 16) tmem_alloc(&s_base, 512);
  fence_proxy_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t base = s_base;
  const uint32_t lane_base = base + ((uint32_t)((warp & 3) * 32) << 16);
  // fill TMEM with something finite
  if (warp < 4) {
    uint32_t r[16];
    for (int j = 0; j < 16; ++j) r[j] = 0x3C003C00u;
    for (int c = 0; c < 512; c += 16) st_x16(lane_base + c, r);
    wait_st();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  long long ld_ops = 0;
  if (warp == 16) {
    // guard: 0 = `lane == 0` branch (ptxas wraps every MMA in an ELECT / branch loop),
    //        1 = elect.sync predicate (bare UTCHMMA)
    bool me;
    if (guard) me = elect_one(); else me = (lane == 0);
    if (me) {
      const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t sb = smem_u32(smem);
      const uint32_t lbo = (uint32_t)N * 16;
      const uint64_t bdesc0 = smem_desc(sb, lbo, 128);
      const uint64_t adesc0 = smem_desc(sb + 96 * 1024, 2048, 128);   // A tile region (SS mode)
      const uint32_t d = base + 256;
      const uint64_t bstep = (uint64_t)((2 * lbo) >> 4);
      const long long t0 = clock64();
      if (mode == 0) {
        for (int rep = 0; rep < iters / 36; ++rep) {
          uint64_t bd = bdesc0;
#pragma unroll 1
          for (int ks = 0; ks < 12; ++ks) {   // the hero's layer-1 loop: (hi, W), (hi, W'), (lo, W)
            const uint32_t a = base + 8u * (uint32_t)ks;
            mma_ts(d, a, bd, idesc, 1u);
            mma_ts(d, a, bd + 1, idesc, 1u);
            mma_ts(d, a + 96u, bd, idesc, 1u);
            bd += bstep;
          }
        }
      } else {
        for (int rep = 0; rep < iters / 36; ++rep) {
          uint64_t bd = bdesc0, ad = adesc0;
#pragma unroll 1
          for (int ks = 0; ks < 12; ++ks) {
            mma_ss(d, ad, bd, idesc, 1u);
            mma_ss(d, ad, bd + 1, idesc, 1u);
            mma_ss(d, ad + 1, bd, idesc, 1u);
            bd += bstep; ad += (ks == 5) ? (uint64_t)0 - 5 * 256 : 256;
          }
        }
      }
      const long long t1 = clock64();
      mma_commit(&bar);
      mbar_wait(&bar, 0);
      const long long t2 = clock64();
      out[0] = t1 - t0;
      out[1] = t2 - t0;
      stop = 1;
    }
    __syncwarp();
  } else if (warp < 4 && (traffic & 1)) {
    uint32_t r[64];
    uint32_t acc = 0;
    while (!stop) {
      ld_x32(lane_base + 256, r);
      ld_x32(lane_base + 288, r + 32);
      wait_ld();
#pragma unroll
      for (int j = 0; j < 64; ++j) acc ^= r[j];
      ld_ops += 2;
    }
    if (acc == 0x12345678u) out[7] = acc;
    if (tid == 0) out[2] = ld_ops;
  } else if (warp >= 4 && warp < 16 && (traffic & 2)) {
    uint32_t r[16];
    for (int j = 0; j < 16; ++j) r[j] = 0x3C003C00u + j;
    long long st_ops = 0;
    while (!stop) {
      st_x16(lane_base + 192 + 16 * ((warp >> 2) - 1), r);   // cols 192..239: not an operand of the MMAs
      wait_st();
      ++st_ops;
    }
    if (tid == 128) out[3] = st_ops;
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 16) { fence_after_sync(); tmem_dealloc(base, 512); }
}

// tcgen05.ld alone: 4 warps (one per lane quadrant) x `iters` x32 loads, clocks per warp-load
__global__ void __launch_bounds__(128, 1) ld_probe(int iters, int nwarps_active, long long* out) {
  __shared__ uint32_t s_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(&s_base, 512);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t lane_base = s_base + ((uint32_t)(warp * 32) << 16);
  uint32_t r[32], acc = 0;
  if (warp < nwarps_active) {
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      ld_x32(lane_base + 32 * (i & 7), r);
      wait_ld();
#pragma unroll
      for (int j = 0; j < 32; ++j) acc ^= r[j];
    }
    const long long t1 = clock64();
    if ((tid & 31) == 0) out[warp] = t1 - t0;
    if (acc == 0x12345678u) out[7] = acc;
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) { fence_after_sync(); tmem_dealloc(s_base, 512); }
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  const int smem_bytes = 160 * 1024;
  cudaFuncSetAttribute(probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  const int iters = 720;
  for (int mode = 0; mode < 2; ++mode)
    for (int N : {32, 64, 128, 256})
      for (int traffic : {0, 1, 2, 3})
        for (int pattern : {0, 1}) {   // = guard
          long long h[8] = {0};
          cudaMemset(d, 0, 64);
          if (pattern) probe<1><<<1, kThreads, smem_bytes>>>(mode, N, iters, traffic, d);
          else probe<0><<<1, kThreads, smem_bytes>>>(mode, N, iters, traffic, d);
          cudaError_t e = cudaDeviceSynchronize();
          cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
          printf("{\"probe\": \"mma\", \"a_in\": \"%s\", \"N\": %d, \"traffic\": \"%s%s\", \"issue_guard\": \"%s\", \"mmas\": %d, "
                 "\"clk_per_mma_issue\": %.1f, \"clk_per_mma_done\": %.1f, \"floor_clk\": %.1f, \"ld_x32_per_warp\": %lld, "
                 "\"st_x16_per_warp\": %lld, \"err\": \"%s\"}\n",
                 mode ? "smem" : "tmem", N, (traffic & 1) ? "ld" : "", (traffic & 2) ? "st" : "", pattern ? "elect.sync" : "lane==0",
                 iters, (double)h[0] / iters, (double)h[1] / iters, 128.0 * N / 256.0, h[2], h[3], cudaGetErrorString(e));
          if (e != cudaSuccess) return 1;
        }
  for (int nw : {1, 2, 4}) {
    long long h[8] = {0};
    cudaMemset(d, 0, 64);
    ld_probe<<<1, 128>>>(2000, nw, d);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
    printf("{\"probe\": \"tcgen05.ld.32x32b.x32 + wait\", \"warps\": %d, \"clk_per_load\": %.1f, \"bytes_per_clk_all_warps\": %.1f, \"err\": \"%s\"}\n",
           nw, (double)h[0] / 2000, nw * 4096.0 / ((double)h[0] / 2000), cudaGetErrorString(e));
  }
  return 0;
}
