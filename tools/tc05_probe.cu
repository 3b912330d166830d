// Hardware probe for the tcgen05 building blocks used by deepbayes_b200/csrc/fused_mlp.cu.
// Single CTA, D[128 x N] = A[128 x K] * B, checked against a CPU reference.
//   variant 0: SS  A K-major (TMA SW128), B K-major  [N x K] (TMA SW128)          -- sanity baseline
//   variant 1: SS  A K-major,             B MN-major [K x N] (TMA SW128)          -- Keras weight layout
//   variant 2: TS  A in TMEM (tcgen05.st packed bf16x2),  B MN-major
//   variant 3: SS  A K-major, B MN-major with N=16 in the no-swizzle core-matrix layout (last layer)
//   variant 4: TS chain: D1 = A*B (fp32, TMEM) -> relu -> packed bf16 IN PLACE -> D2 = relu(D1)*B2
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tools/tc05_probe tools/tc05_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_bf16.h>
#include "../deepbayes_b200/csrc/tc05.cuh"

using namespace tc05;

struct Params {
  int variant, N, K, N2;
  const __nv_bfloat16* A;   // [128,K] row-major (for TS variants: read directly)
  float* D;                 // [128,N] (or [128,N2] for variant 4)
};

extern __shared__ __align__(1024) uint8_t smem_raw[];

__global__ void __launch_bounds__(128) probe_kernel(const __grid_constant__ CUtensorMap tmA,
                                                    const __grid_constant__ CUtensorMap tmB,
                                                    const __grid_constant__ CUtensorMap tmB2, Params p) {
  __shared__ __align__(8) uint64_t bar_full, bar_mma, bar_mma2;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int N = p.N, K = p.K;
  const uint32_t sA = smem_u32(smem);
  const uint32_t sB = sA + 128 * K * 2;
  const uint32_t sB2 = sB + K * N * 2;
  if (tid == 0) {
    mbar_init(smem_u32(&bar_full), 1);
    mbar_init(smem_u32(&bar_mma), 1);
    mbar_init(smem_u32(&bar_mma2), 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<512>(smem_u32(&tmem_slot));
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_slot;

  if (tid == 0) {
    uint32_t bytes = 0;
    const uint32_t bf = smem_u32(&bar_full);
    if (p.variant != 2 && p.variant != 4) bytes += 128 * K * 2;
    else bytes += 0;
    bytes += K * N * 2;
    if (p.variant == 4) bytes += N * p.N2 * 2;
    mbar_expect_tx(bf, bytes);
    if (p.variant != 2 && p.variant != 4)
      for (int kb = 0; kb < K / 64; ++kb) tma_load_2d(sA + kb * 16384, &tmA, bf, kb * 64, 0);
    if (p.variant == 0) {
      for (int kb = 0; kb < K / 64; ++kb) tma_load_2d(sB + kb * (N * 128), &tmB, bf, kb * 64, 0);
    } else if (p.variant == 3) {
      // no-swizzle: boxes of [8 n x K rows]; core matrix = 8 k-rows x 16 B
      for (int nb = 0; nb < N / 8; ++nb) tma_load_2d(sB + nb * (K * 16), &tmB, bf, nb * 8, 0);
    } else {
      for (int nb = 0; nb < N / 64; ++nb) tma_load_2d(sB + nb * (K * 128), &tmB, bf, nb * 64, 0);
    }
    if (p.variant == 4)
      for (int nb = 0; nb < p.N2 / 64; ++nb) tma_load_2d(sB2 + nb * (N * 128), &tmB2, bf, nb * 64, 0);
  }
  const uint32_t a_col0 = 256;  // TMEM columns holding packed A for the TS variants
  if (p.variant == 2 || p.variant == 4) {
    // thread = row: pack A[row][k..k+31] into 16 columns per store
    const __nv_bfloat16* arow = p.A + (size_t)tid * K;
    for (int k0 = 0; k0 < K; k0 += 32) {
      uint32_t v[16];
      for (int i = 0; i < 16; ++i) {
        const uint32_t lo = __bfloat16_as_ushort(arow[k0 + 2 * i]), hi = __bfloat16_as_ushort(arow[k0 + 2 * i + 1]);
        v[i] = lo | (hi << 16);
      }
      tmem_st16(tmem + ((uint32_t)(warp * 32) << 16) + a_col0 + k0 / 2, v);
    }
    wait_st();
    fence_before_sync();
  }
  __syncthreads();
  fence_after_sync();

  if (tid == 0) {
    mbar_wait(smem_u32(&bar_full), 0);
    fence_after_sync();
    const uint32_t idesc = make_idesc_bf16(128, N, 0, p.variant == 0 ? 0 : 1);
    for (int k = 0; k < K / 16; ++k) {
      const int kb = k / 4, kk = k % 4;
      uint64_t bdesc;
      if (p.variant == 0) bdesc = make_smem_desc(sB + kb * (N * 128) + kk * 32, 16, 1024, SW_128B);
      else if (p.variant == 3) bdesc = make_smem_desc(sB + k * 16 * 16, /*LBO: k-groups*/ 128, /*SBO: n-blocks*/ K * 16, SW_NONE);
      else bdesc = make_smem_desc(sB + k * 16 * 128, K * 128, 1024, SW_128B);
      if (p.variant == 2 || p.variant == 4) {
        mma_ts(tmem, tmem + a_col0 + k * 8, bdesc, idesc, k > 0);
      } else {
        const uint64_t adesc = make_smem_desc(sA + kb * 16384 + kk * 32, 16, 1024, SW_128B);
        mma_ss(tmem, adesc, bdesc, idesc, k > 0);
      }
    }
    mma_commit(smem_u32(&bar_mma));
  }
  mbar_wait(smem_u32(&bar_mma), 0);
  fence_after_sync();
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  if (p.variant != 4) {
    for (int c0 = 0; c0 < N; c0 += 16) {
      uint32_t v[16];
      tmem_ld16(tmem + lane_base + c0, v);
      wait_ld();
      for (int i = 0; i < 16; ++i) p.D[(size_t)tid * N + c0 + i] = __uint_as_float(v[i]);
    }
  } else {
    // relu + pack in place: fp32 cols [0,N) -> bf16x2 cols [0,N/2)
    for (int c0 = 0; c0 < N; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem + lane_base + c0, v);
      wait_ld();
      uint32_t w[16];
      for (int i = 0; i < 16; ++i)
        w[i] = pack_bf16x2(fmaxf(__uint_as_float(v[2 * i]), 0.f), fmaxf(__uint_as_float(v[2 * i + 1]), 0.f));
      tmem_st16(tmem + lane_base + c0 / 2, w);
    }
    wait_st();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const int N2 = p.N2;
    const uint32_t d2 = 256;  // second accumulator
    if (tid == 0) {
      const uint32_t idesc2 = make_idesc_bf16(128, N2, 0, 1);
      for (int k = 0; k < N / 16; ++k) {
        const uint64_t bdesc = make_smem_desc(sB2 + k * 16 * 128, N * 128, 1024, SW_128B);
        mma_ts(tmem + d2, tmem + k * 8, bdesc, idesc2, k > 0);
      }
      mma_commit(smem_u32(&bar_mma2));
    }
    mbar_wait(smem_u32(&bar_mma2), 0);
    fence_after_sync();
    for (int c0 = 0; c0 < N2; c0 += 16) {
      uint32_t v[16];
      tmem_ld16(tmem + lane_base + d2 + c0, v);
      wait_ld();
      for (int i = 0; i < 16; ++i) p.D[(size_t)tid * N2 + c0 + i] = __uint_as_float(v[i]);
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

static int run_variant(int variant, int N, int K, int N2) {
  std::vector<float> A(128 * K), Bkn(K * N), B2(N * N2);
  srand(variant * 7 + 1);
  for (auto& v : A) v = bf((rand() % 17 - 8) / 8.0f);
  for (auto& v : Bkn) v = bf((rand() % 13 - 6) / 16.0f);
  for (auto& v : B2) v = bf((rand() % 11 - 5) / 16.0f);
  std::vector<__nv_bfloat16> hA(128 * K), hB(K * N), hB2(N * N2);
  for (int i = 0; i < 128 * K; ++i) hA[i] = __float2bfloat16(A[i]);
  if (variant == 0) for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) hB[n * K + k] = __float2bfloat16(Bkn[k * N + n]);
  else for (int i = 0; i < K * N; ++i) hB[i] = __float2bfloat16(Bkn[i]);
  for (int i = 0; i < N * N2; ++i) hB2[i] = __float2bfloat16(B2[i]);
  __nv_bfloat16 *dA, *dB, *dB2; float* dD;
  const int NO = variant == 4 ? N2 : N;
  CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dB2, hB2.size() * 2));
  CK(cudaMalloc(&dD, 128 * NO * 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB2, hB2.data(), hB2.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xff, 128 * NO * 4));
  CUtensorMap tmA, tmB, tmB2;
  {
    uint64_t d[2] = {(uint64_t)K, 128}, s[1] = {(uint64_t)K * 2}; uint32_t b[2] = {64, 128};
    if (!make_tmap_bf16(&tmA, dA, 2, d, s, b, CU_TENSOR_MAP_SWIZZLE_128B)) { printf("tmap A failed\n"); return 1; }
  }
  if (variant == 0) {
    uint64_t d[2] = {(uint64_t)K, (uint64_t)N}, s[1] = {(uint64_t)K * 2}; uint32_t b[2] = {64, (uint32_t)N};
    if (!make_tmap_bf16(&tmB, dB, 2, d, s, b, CU_TENSOR_MAP_SWIZZLE_128B)) { printf("tmap B failed\n"); return 1; }
  } else if (variant == 3) {
    uint64_t d[2] = {(uint64_t)N, (uint64_t)K}, s[1] = {(uint64_t)N * 2}; uint32_t b[2] = {8, (uint32_t)K};
    if (!make_tmap_bf16(&tmB, dB, 2, d, s, b, CU_TENSOR_MAP_SWIZZLE_NONE)) { printf("tmap B failed\n"); return 1; }
  } else {
    uint64_t d[2] = {(uint64_t)N, (uint64_t)K}, s[1] = {(uint64_t)N * 2}; uint32_t b[2] = {64, (uint32_t)K};
    if (!make_tmap_bf16(&tmB, dB, 2, d, s, b, CU_TENSOR_MAP_SWIZZLE_128B)) { printf("tmap B failed\n"); return 1; }
  }
  {
    uint64_t d[2] = {(uint64_t)N2, (uint64_t)N}, s[1] = {(uint64_t)N2 * 2}; uint32_t b[2] = {64, (uint32_t)N};
    if (!make_tmap_bf16(&tmB2, dB2, 2, d, s, b, CU_TENSOR_MAP_SWIZZLE_128B)) { printf("tmap B2 failed\n"); return 1; }
  }
  Params p{variant, N, K, N2, dA, dD};
  const int smem = 128 * K * 2 + K * N * 2 + N * N2 * 2 + 2048;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  probe_kernel<<<1, 128, smem>>>(tmA, tmB, tmB2, p);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<float> D(128 * NO);
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  std::vector<float> ref(128 * N);
  for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) {
    float acc = 0; for (int k = 0; k < K; ++k) acc += A[m * K + k] * Bkn[k * N + n]; ref[m * N + n] = acc;
  }
  double maxerr = 0; int bad = 0;
  if (variant != 4) {
    for (int i = 0; i < 128 * N; ++i) { double e = fabs(D[i] - ref[i]); if (!(e <= 1e-3)) ++bad; if (e > maxerr || e != e) maxerr = e; }
  } else {
    for (int m = 0; m < 128; ++m) for (int n = 0; n < N2; ++n) {
      float acc = 0; for (int k = 0; k < N; ++k) acc += bf(fmaxf(ref[m * N + k], 0.f)) * B2[k * N2 + n];
      double e = fabs(D[m * N2 + n] - acc); if (!(e <= 2e-2)) ++bad; if (e > maxerr || e != e) maxerr = e;
    }
  }
  printf("variant %d N=%d K=%d N2=%d: max|err|=%.4g bad=%d -> %s\n", variant, N, K, N2, maxerr, bad, bad ? "FAIL" : "ok");
  if (bad) { printf("  D[0,0:4]= %g %g %g %g   ref= %g %g %g %g\n", D[0], D[1], D[2], D[3], ref[0], ref[1], ref[2], ref[3]); }
  cudaFree(dA); cudaFree(dB); cudaFree(dB2); cudaFree(dD);
  return bad ? 2 : 0;
}

int main(int argc, char** argv) {
  int rc = 0;
  if (argc > 1) return run_variant(atoi(argv[1]), argc > 2 ? atoi(argv[2]) : 128, argc > 3 ? atoi(argv[3]) : 128, argc > 4 ? atoi(argv[4]) : 128);
  rc |= run_variant(0, 128, 128, 64);
  rc |= run_variant(0, 256, 256, 64);
  rc |= run_variant(1, 128, 128, 64);
  rc |= run_variant(1, 256, 256, 64);
  rc |= run_variant(1, 192, 128, 64);
  rc |= run_variant(2, 128, 128, 64);
  rc |= run_variant(2, 256, 256, 64);
  rc |= run_variant(3, 16, 128, 64);
  rc |= run_variant(3, 16, 256, 64);
  rc |= run_variant(4, 128, 128, 128);
  rc |= run_variant(4, 256, 128, 128);
  return rc;
}
