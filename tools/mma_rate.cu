// Microbenchmark: cycles per tcgen05.mma for the shapes the fused kernel issues (single CTA,
// operands resident in smem/TMEM, garbage data).  Reports cycles/MMA and the implied fraction
// of the 8192 FLOP/clk/SM tensor peak.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tools/mma_rate tools/mma_rate.cu
#include <cstdio>
#include <cstdlib>
#include "../deepbayes_b200/csrc/tc05.cuh"
using namespace tc05;

extern __shared__ __align__(1024) uint8_t smem_raw[];

// mode: 0 = SS (A smem K-major, B smem MN-major), 1 = TS (A tmem, B smem MN-major), 2 = SS with B K-major
__global__ void __launch_bounds__(128) rate_kernel(int mode, int N, int iters, int nrot, long long* out) {
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t s0 = smem_u32(smem) & 0x3FFFF;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc<512>(smem_u32(&slot));
  fence_before_sync(); __syncthreads(); fence_after_sync();
  const uint32_t tmem = slot;
  if (warp == 1) {
    const uint32_t idesc = make_idesc_bf16(128, N, 0, mode == 2 ? 0 : 1);
    constexpr uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    long long t0 = 0, t1 = 0;
    if (elect_one()) {
      t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        // rotate over nrot different smem stages (48 KB apart) like the real pipeline does
        const uint32_t base = s0 + (uint32_t)(i % nrot) * 49152;
        const uint32_t a_lo = (base >> 4) | (1u << 16);
        const uint32_t b_lo = mode == 2 ? (((base + 16384) >> 4) | (1u << 16)) : (((base + 16384) >> 4) | ((8192u >> 4) << 16));
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          if (mode == 1) mma_ts_lh(tmem, tmem + 256 + kk * 8, b_lo + kk * 128, hi, idesc, 1);
          else if (mode == 0) mma_ss_lh(tmem, a_lo + kk * 2, hi, b_lo + kk * 128, hi, idesc, 1);
          else mma_ss_lh(tmem, a_lo + kk * 2, hi, b_lo + kk * 2, hi, idesc, 1);
        }
      }
      mma_commit(smem_u32(&bar));
    }
    __syncwarp();
    mbar_wait(smem_u32(&bar), 0);
    if (elect_one()) { t1 = clock64(); out[0] = t1 - t0; }
  }
  fence_before_sync(); __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  const int smem = 4 * 49152 + 2048;
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const char* names[3] = {"SS A:K-major B:MN-major", "TS A:TMEM    B:MN-major", "SS A:K-major B:K-major "};
  for (int mode = 0; mode < 3; ++mode)
    for (int N : {64, 128, 256})
      for (int nrot : {1, 4}) {
        const int iters = 2000;
        rate_kernel<<<1, 128, smem>>>(mode, N, iters, nrot, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        long long c; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
        const double per = (double)c / (iters * 4);
        printf("%s N=%3d stages=%d: %.1f cycles/MMA (ideal %.0f) -> %.1f%% of tensor peak\n", names[mode], N, nrot, per,
               128.0 * N / 256, 100.0 * (128.0 * N / 256) / per);
      }
  // all 148 SMs at once (power / clock effects)
  for (int mode = 0; mode < 2; ++mode) {
    const int N = mode == 0 ? 256 : 128, iters = 20000;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    rate_kernel<<<148, 128, smem>>>(mode, N, 100, 4, d);
    cudaEventRecord(a);
    rate_kernel<<<148, 128, smem>>>(mode, N, iters, 4, d);
    cudaEventRecord(b);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double flops = 148.0 * iters * 4 * 2.0 * 128 * N * 16;
    printf("148 CTAs %s N=%d: %.3f ms -> %.0f TFLOP/s\n", names[mode], N, ms, flops / ms / 1e9);
  }
  return 0;
}
