// Microbenchmark: how many bytes per clock can ONE SM take from L2 through the TMA engine (cp.async.bulk), and does
// cluster multicast raise the delivered rate?  Decides whether the fused MLP kernel (35.7 B/clk/SM of L2->SM traffic,
// profiles/r1_ncu_fused_pair_summary.txt) is bound by the L2 side (then weight multicast over two CTA pairs helps) or by
// the per-SM ingest path (then only fewer bytes per unit help).
//
//   mode 0  every CTA streams its own bytes                      (cluster size CM, no multicast)
//   mode 1  per stage: half own bytes + half shared by the CM CTAs of the cluster (each CTA issues 1/CM of the shared
//           half with .multicast::cluster to all of them)        = the fused kernel with weight multicast
//   mode 2  per stage: everything shared by the cluster (each CTA issues 1/CM, multicast to all)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tools/tma_rate tools/tma_rate.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t cta) {
  asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\tmbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(bar), "r"(cta) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ int g_spin = 0;   // 1: poll with test_wait (no hardware suspend) instead of try_wait
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  if (g_spin) { while (!mbar_test_wait(bar, parity)) if (++spins > (1u << 28)) { printf("timeout block %d\n", blockIdx.x); __trap(); } return; }
  while (!mbar_try_wait(bar, parity)) if (++spins > (1u << 22)) { printf("timeout block %d\n", blockIdx.x); __trap(); }
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_load_mc(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }

extern __shared__ __align__(1024) uint8_t smem_raw[];
constexpr int kMaxStages = 16;

__global__ void __launch_bounds__(128) rate_kernel(const uint8_t* __restrict__ buf, size_t buf_bytes, int mode, int CM, int NST, int stage_bytes,
                                                  int piece, int iters, long long* out, int NP) {
  __shared__ __align__(8) uint64_t full[kMaxStages], empty[kMaxStages];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t s0 = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x / CM;
  const bool shared_any = mode != 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < NST; ++i) { mbar_init(smem_u32(&full[i]), 1); mbar_init(smem_u32(&empty[i]), shared_any ? CM : 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  cluster_sync();
  const uint16_t mask = (uint16_t)((1u << CM) - 1);
  const int own_bytes = mode == 0 ? stage_bytes : (mode == 1 ? stage_bytes / 2 : 0);
  const int sh_bytes = stage_bytes - own_bytes;          // shared by the cluster
  const int sh_mine = sh_bytes / CM;                     // the part this CTA issues
  long long t0 = 0, t1 = 0;
  if (warp < NP && lane == 0) {
    const size_t limit = (buf_bytes - 2 * (size_t)stage_bytes) & ~(size_t)1023;
    size_t off = ((((size_t)blockIdx.x * 7919u) % 1024u) * (buf_bytes / 1024u)) % limit; off &= ~(size_t)1023;       // no division inside the loop
    size_t so = ((((size_t)cluster_id * 104729u) % 1024u) * (buf_bytes / 1024u) + buf_bytes / 2) % limit; so &= ~(size_t)1023;
    int st = 0, ph = 0;
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (it % NP != warp) { off += stage_bytes; if (off >= limit) off -= limit; so += stage_bytes; if (so >= limit) so -= limit; if (++st == NST) { st = 0; ph ^= 1; } continue; }
      mbar_wait(smem_u32(&empty[st]), ph ^ 1);
      const uint32_t fb = smem_u32(&full[st]);
      mbar_expect_tx(fb, (uint32_t)stage_bytes);
      const uint32_t dst = s0 + st * stage_bytes;
      for (int b = 0; b < own_bytes; b += piece) bulk_load(dst + b, buf + off + b, piece, fb);
      if (sh_bytes) {
        const int base = own_bytes + (int)rank * sh_mine;
        for (int b = 0; b < sh_mine; b += piece) {
          const int pc = min(piece, sh_mine - b);
          bulk_load_mc(dst + base + b, buf + so + (int)rank * sh_mine + b, pc, fb, mask);
        }
      }
      off += stage_bytes; if (off >= limit) off -= limit;
      so += stage_bytes; if (so >= limit) so -= limit;
      if (++st == NST) { st = 0; ph ^= 1; }
    }
    t1 = clock64();
  } else if (warp >= NP && warp < 2 * NP && lane == 0) {
    int st = 0, ph = 0;
    for (int it = 0; it < iters; ++it) {
      if (it % NP != warp - NP) { if (++st == NST) { st = 0; ph ^= 1; } continue; }
      mbar_wait(smem_u32(&full[st]), ph);
      if (shared_any) { for (int r = 0; r < CM; ++r) mbar_arrive_cluster(smem_u32(&empty[st]), r); }
      else mbar_arrive_cluster(smem_u32(&empty[st]), rank);
      if (++st == NST) { st = 0; ph ^= 1; }
    }
  }
  __syncthreads();
  cluster_sync();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}

int main(int argc, char** argv) {
  int dev = 0; cudaSetDevice(dev);
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, dev);
  const int sms = pr.multiProcessorCount;
  const size_t buf_bytes = (size_t)48 << 20;     // fits the 126 MB L2
  uint8_t* buf; cudaMalloc(&buf, buf_bytes); cudaMemset(buf, 1, buf_bytes);
  long long* out; cudaMalloc(&out, sizeof(long long) * 1024);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  printf("SMs %d\n", sms);
  printf("%-5s %-3s %-4s %-6s %-6s %-9s %-10s %-10s %-9s\n", "mode", "CM", "NST", "stageK", "piece", "clusters", "B/clk/SM", "GB/s(dlv)", "us");
  struct Cfg { int mode, CM, NST, stage_kb, piece, spin, np; };
  const Cfg cfgs[] = {
    {0, 1, 6, 32, 16384, 0, 1}, {0, 1, 6, 32, 16384, 0, 2}, {0, 1, 6, 32, 8192, 0, 1}, {0, 1, 6, 32, 8192, 0, 2}, {0, 1, 3, 64, 16384, 0, 1}, {0, 1, 3, 64, 16384, 0, 2},
    {0, 1, 3, 64, 32768, 0, 1}, {0, 1, 3, 64, 65536, 0, 1}, {0, 1, 6, 32, 32768, 0, 1}, {0, 1, 6, 32, 32768, 0, 2}, {0, 1, 12, 16, 16384, 0, 1}, {0, 1, 12, 16, 16384, 0, 2},
    {0, 1, 4, 48, 16384, 0, 1}, {0, 1, 4, 48, 49152, 0, 1},
  };
  for (const Cfg& c : cfgs) {
    cudaMemcpyToSymbol(g_spin, &c.spin, sizeof(int));
    const int iters = 4000;
    int clusters = sms / c.CM;
    // how many clusters of this size fit?
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = (size_t)c.NST * c.stage_kb * 1024 + 2048;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = c.CM; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cfg.gridDim = dim3(clusters * c.CM);
    int maxc = 0;
    cudaOccupancyMaxActiveClusters(&maxc, rate_kernel, &cfg);
    if (maxc > 0 && maxc < clusters) clusters = maxc;
    cfg.gridDim = dim3(clusters * c.CM);
    float best = 1e30f; double bclk = 0;
    for (int rep = 0; rep < 3; ++rep) {
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0);
      cudaError_t e = cudaLaunchKernelEx(&cfg, rate_kernel, (const uint8_t*)buf, buf_bytes, c.mode, c.CM, c.NST, c.stage_kb * 1024, c.piece, iters, out, c.np);
      cudaEventRecord(e1);
      cudaError_t e2 = cudaDeviceSynchronize();
      if (e != cudaSuccess || e2 != cudaSuccess) { printf("launch failed: %s %s\n", cudaGetErrorString(e), cudaGetErrorString(e2)); return 1; }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best) {
        best = ms;
        long long h[1024]; cudaMemcpy(h, out, sizeof(long long) * clusters * c.CM, cudaMemcpyDeviceToHost);
        double s = 0; for (int i = 0; i < clusters * c.CM; ++i) s += (double)h[i];
        bclk = (double)iters * c.stage_kb * 1024 / (s / (clusters * c.CM));
      }
    }
    const double gbs = (double)clusters * c.CM * iters * c.stage_kb * 1024 / (best * 1e-3) / 1e9;
    printf("%-5d %-3d %-4d %-6d %-6d %-9d %-10.2f %-10.1f %-9.1f spin=%d producers=%d\n", c.mode, c.CM, c.NST, c.stage_kb, c.piece, clusters, bclk, gbs, best * 1e3, c.spin, c.np);
  }
  return 0;
}
