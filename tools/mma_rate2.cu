// Microbenchmark for the CTA-pair (tcgen05 cta_group::2, M = 256) MMA shapes of the fused kernels:
// cycles per MMA alone, and with the epilogue warps of both CTAs streaming tcgen05.ld from other
// TMEM columns at the same time (does the TMEM read port limit the .ts A operand?).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tools/mma_rate2 tools/mma_rate2.cu
#include <cstdio>
#include <cstdlib>
#include "../deepbayes_b200/csrc/tc05.cuh"
using namespace tc05;

extern __shared__ __align__(1024) uint8_t smem_raw[];

// mode: 0 = SS (A smem K-major SW128, B smem MN-major SW128), 1 = TS (A TMEM)
// ld_iters: each epilogue warp issues this many tcgen05.ld.32x32b.x32 (4 KB each) while the MMAs run
__global__ void __launch_bounds__(192) rate2_kernel(int mode, int N, int iters, int ld_iters, long long* out, int d_col, int a_col, int ep_col, int ep_mode, int commit_every) {
  __shared__ __align__(8) uint64_t bar;
  __shared__ __align__(8) uint64_t bar2;
  __shared__ uint32_t slot;
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t s0 = smem_u32(smem) & 0x3FFFF;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_init(smem_u32(&bar2), 1000000); fence_mbar_init(); }
  if (warp == 0) tmem_alloc2<512>(smem_u32(&slot));
  fence_before_sync(); __syncthreads(); cluster_sync(); fence_after_sync();
  const uint32_t tmem = slot;
  if (ep_mode >= 2) {   // realistic operands: pseudo-random bf16 in smem and in the TMEM A columns, zeroed accumulators
    uint32_t x = 0x9E3779B9u * (threadIdx.x + 1) + blockIdx.x * 7919u;
    for (int i = threadIdx.x; i < 4 * 32768 / 4; i += blockDim.x) {
      x = x * 1664525u + 1013904223u;
      const uint32_t lo = 0x3C00u | ((x >> 9) & 0x83FFu), hi = 0x3C00u | ((x >> 20) & 0x83FFu);   // |v| in [0.0078, 0.0156), random sign+mantissa
      reinterpret_cast<uint32_t*>(smem)[i] = lo | (hi << 16);
    }
    if (warp >= 2) {
      const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
      for (int c = 0; c < 512; c += 16) {
        uint32_t w[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) { x = x * 1664525u + 1013904223u; w[j] = (c >= a_col && c < a_col + 64) ? (0x3C003C00u | (x & 0x83FF83FFu)) : 0u; }
        tmem_st16(tmem + lane_base + c, w);
      }
      wait_st();
    }
    fence_proxy_async_smem();
    fence_before_sync(); __syncthreads(); cluster_sync(); fence_after_sync();
  }
  if (warp == 1) {
    if (rank == 0) {
      const uint32_t idesc = make_idesc_bf16(256, N, 0, 1);
      constexpr uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
      long long t0 = 0, t1 = 0;
      if (elect_one()) {
        t0 = clock64();
        for (int i = 0; i < iters; ++i) {
          const uint32_t base = s0 + (uint32_t)(i % 4) * 32768;
          const uint32_t a_lo = (base >> 4) | (1u << 16);
          const uint32_t b_lo = ((base + 16384) >> 4) | ((8192u >> 4) << 16);
#pragma unroll
          if (mode == 2) {   // the fused kernel's layer-2 pattern: 8 MMAs per stage, 4 stages per chunk, D alternates per chunk
            const int sg = (i >> 1) & 3, j = i >> 3, half = i & 1;
            const uint32_t dbuf = tmem + (j & 1) * 256 + 128;
            const uint32_t a0 = tmem + (uint32_t)(sg >> 1) * 256 + (uint32_t)(sg & 1) * 64 + half * 32;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              mma2_ts_lh(dbuf, a0 + kk * 8, b_lo + kk * 128, hi, idesc, (sg | half | kk) > 0);
          } else
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            if (mode == 1) mma2_ts_lh(tmem + d_col, tmem + a_col + (i & 1) * 32 + kk * 8, b_lo + kk * 128, hi, idesc, 1);
            else mma2_ss_lh(tmem + d_col, a_lo + kk * 2, hi, b_lo + kk * 128, hi, idesc, 1);
          }
          if (commit_every && (i % commit_every) == commit_every - 1) mma2_commit_mc(smem_u32(&bar2), 0x3);   // like a per-stage `empty` commit
          if (ep_mode >= 20 && (mode == 0 || (i & 1))) {   // issue-thread stall of ep_mode cycles after every 8 MMAs: how deep is the MMA queue?
            const long long t = clock64();
            while (clock64() - t < ep_mode) {}
          }
        }
        mma2_commit_mc(smem_u32(&bar), 0x3);
      }
      __syncwarp();
      mbar_wait(smem_u32(&bar), 0);
      if (elect_one()) { t1 = clock64(); if (blockIdx.x == 0) out[0] = t1 - t0; }
    }
  } else if (warp >= 2) {
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    long long t0 = clock64();
    uint32_t acc = 0;
    int done_iters = 0;
    for (int i = 0; i < ld_iters; ++i, ++done_iters) {
      uint32_t ok;
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
      if (ok) break;    // the MMAs have retired (commit is multicast to both CTAs)
      uint32_t v[32];
      tmem_ld32(tmem + lane_base + ep_col + (i & 3) * 32, v);
      wait_ld();
#pragma unroll
      for (int j = 0; j < 32; ++j) acc ^= v[j];
      if (ep_mode == 1) {   // pack-like: write 16 columns back
        uint32_t w[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) w[j] = v[2 * j] + v[2 * j + 1];
        tmem_st16(tmem + lane_base + ep_col + (i & 3) * 16, w);
        wait_st();
      }
    }
    long long t1 = clock64();
    if (lane == 0 && warp == 2 && blockIdx.x == 0) { out[1] = t1 - t0; out[3] = done_iters; }
    if (acc == 0x12345678u) out[2] = acc;
  }
  fence_before_sync(); __syncthreads(); cluster_sync();
  if (warp == 0) tmem_dealloc2<512>(tmem);
}

int main() {
  long long* d; cudaMalloc(&d, 32);
  const int smem = 4 * 32768 + 2048;
  cudaFuncSetAttribute(rate2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const char* names[2] = {"pair SS", "pair TS"};
  struct Case { int mode, N, ld, d_col, a_col, ep_col, ep_mode; const char* what; int commit_every; int clusters; };
  const Case cases[] = {
    {0, 256, 0, 0, 0, 384, 0, "SS N=256 D=[0,256) alone"},
    {0, 256, 1, 0, 0, 384, 0, "SS N=256 D=[0,256) + ld [384,512)"},
    {0, 256, 1, 0, 0, 256, 1, "SS N=256 D=[0,256) + ld/st pack [256,384)"},
    {1, 128, 0, 0, 256, 384, 0, "TS N=128 D=[0,128) A=[256,320) alone"},
    {1, 128, 0, 128, 0, 384, 0, "TS N=128 D=[128,256) A=[0,64) alone (same 256-col region, as in the kernel)"},
    {1, 128, 0, 384, 0, 384, 0, "TS N=128 D=[384,512) A=[0,64) alone (other region)"},
    {1, 128, 1, 128, 0, 384, 0, "TS N=128 D=[128,256) A=[0,64) + ld [384,512)"},
    {1, 128, 1, 128, 0, 384, 1, "TS N=128 D=[128,256) A=[0,64) + ld/st pack [384,512)"},
    {1, 128, 1, 128, 0, 256, 1, "TS N=128 D=[128,256) A=[0,64) + ld/st pack [256,384)"},
    {1, 16, 0, 192, 128, 384, 0, "TS N=16 D=[192,208) A=[128,192) alone"},
    {1, 256, 0, 256, 0, 384, 0, "TS N=256 D=[256,512) A=[0,64) alone"},
    {1, 128, 0, 128, 0, 384, 0, "TS N=128 D=[128,256) + commit every 8 MMAs", 2},
    {1, 128, 0, 128, 0, 384, 0, "TS N=128 D=[128,256) + commit every 4 MMAs", 1},
    {0, 256, 0, 0, 0, 384, 0, "SS N=256 + commit every 4 MMAs", 1},
    {1, 128, 1, 128, 0, 384, 1, "TS N=128 D=[128,256) + commit every 8 MMAs + ld/st pack", 2},
    {1, 128, 0, 128, 0, 384, 20, "TS N=128, issue thread idles 20 cycles after every 8 MMAs", 2, 1},
    {1, 128, 0, 128, 0, 384, 40, "TS N=128, issue thread idles 40 cycles after every 8 MMAs", 2, 1},
    {1, 128, 0, 128, 0, 384, 60, "TS N=128, issue thread idles 60 cycles after every 8 MMAs", 2, 1},
    {1, 128, 0, 128, 0, 384, 80, "TS N=128, issue thread idles 80 cycles after every 8 MMAs", 2, 1},
    {0, 256, 0, 0, 0, 384, 20, "SS N=256, issue thread idles 20 cycles after every 4 MMAs", 1, 1},
    {0, 256, 0, 0, 0, 384, 50, "SS N=256, issue thread idles 50 cycles after every 4 MMAs", 1, 1},
    {0, 256, 0, 0, 0, 384, 100, "SS N=256, issue thread idles 100 cycles after every 4 MMAs", 1, 1},
    {0, 256, 0, 0, 0, 384, 200, "SS N=256, issue thread idles 200 cycles after every 4 MMAs", 1, 1},
    {0, 256, 0, 0, 0, 384, 400, "SS N=256, issue thread idles 400 cycles after every 4 MMAs", 1, 1},
    {1, 256, 0, 256, 0, 384, 100, "TS N=256, issue thread idles 100 cycles after every 8 MMAs", 2, 1},
    {1, 128, 0, 128, 0, 384, 100, "TS N=128, issue thread idles 100 cycles after every 8 MMAs", 2, 1},
    {1, 128, 0, 128, 0, 384, 200, "TS N=128, issue thread idles 200 cycles after every 8 MMAs", 2, 1},
    {1, 128, 0, 128, 0, 384, 400, "TS N=128, issue thread idles 400 cycles after every 8 MMAs", 2, 1},
    {1, 128, 0, 128, 0, 384, 800, "TS N=128, issue thread idles 800 cycles after every 8 MMAs", 2, 1},
    {2, 128, 0, 128, 0, 384, 0, "TS N=128, kernel layer-2 pattern (D alternates every 32 MMAs, A over 4 regions)", 2, 1},
    {1, 128, 0, 128, 0, 384, 2, "TS N=128, random operands, 1 cluster", 2, 1},
    {1, 128, 0, 128, 0, 384, 2, "TS N=128, random operands, 74 clusters", 2, 74},
    {1, 128, 0, 128, 0, 384, 0, "TS N=128, zero/garbage operands, 74 clusters", 2, 74},
    {0, 256, 0, 0, 0, 384, 2, "SS N=256, random operands, 1 cluster", 1, 1},
    {0, 256, 0, 0, 0, 384, 2, "SS N=256, random operands, 74 clusters", 1, 74},
  };
  for (const Case& cs : cases) {
    const int iters = cs.clusters > 1 ? 20000 : 2000;
    const double ideal = 128.0 * cs.N / 256;
    const int ld_iters = cs.ld ? 4000000 : 0;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * (cs.clusters > 0 ? cs.clusters : 1)); cfg.blockDim = dim3(192); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaMemset(d, 0, 32);
    cudaLaunchKernelEx(&cfg, rate2_kernel, cs.mode, cs.N, iters, ld_iters, d, cs.d_col, cs.a_col, cs.ep_col, cs.ep_mode, cs.commit_every);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    long long c[4]; cudaMemcpy(c, d, 32, cudaMemcpyDeviceToHost);
    const double per = (double)c[0] / (iters * 4);
    printf("%-80s: %.1f cycles/MMA (ideal %.0f) -> %.1f%%", cs.what, per, ideal, 100.0 * ideal / per);
    if (cs.ld) printf("; epilogue: %.1f cycles per 4 KB warp-load (%.1f B/clk/SM)", (double)c[1] / (double)c[3], 4.0 * 4096 / ((double)c[1] / (double)c[3]));
    printf("\n");
  }
  return 0;
}
