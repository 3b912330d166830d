// Shared device/host helpers for the deepbayes_b200 engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <atomic>
#include <string>

namespace dbx {

extern thread_local std::string g_err;
extern std::atomic<long long> g_launches;

int set_err(const char* fmt, ...);

// Process-wide A/B switches.  Read ONCE from the DBX_* environment (first use), afterwards only changed through
// dbx_set_option / dbx_reset_options (tests, tools/): nothing on a call path touches getenv().  None of them selects a
// CPU or library fallback -- every value still runs this library's own kernels.
struct Options {
  long long no_fused = 0;           // skip fused_mlp*_kernel, use the generic executor
  long long no_stream = 0;          // skip stream_mlp_kernel
  long long no_tc = 0;              // generic executor / IBP on the CUDA-core kernels only
  long long no_pair_gemm = 0;       // big per-sample GEMMs on the single-CTA kernels instead of dense_tc_pair_kernel
  long long no_tma_store = 0;       // dense_tc_pair_kernel stores its bf16 output directly
  long long no_persist = 0;         // one CTA per tile (dense_tc_kernel)
  long long persist_min_tiles = 296;
  long long no_f32w = 0;            // stored posteriors: convert the slab chunk to bf16 first
  long long no_f32w_wide = 0;       // dense_tc_f32w_kernel: 64-column tiles, two CTAs per SM (the round-1 shape) instead of 128-column tiles
  long long no_shared_a = 0;        // a small layer on the call's shared input: per-sample GEMMs instead of dense_tc_shareda_kernel (8 samples per tile)
  long long no_direct_conv = 0;     // Conv2D as im2col + GEMM
  long long no_pool_fuse = 0;       // separate max-pool launch
  long long no_ibp_tc = 0;          // interval forward: two GEMMs + combine per hidden layer
  long long ibp_overlap = 1, ibp_prio = 1, generic_overlap = 1, fused_overlap = 1, fused_prio = 1;
  long long fused_kernel = 0;       // 0 = by shape; 1 = single-CTA kernel; 2 = CTA-pair kernel
  long long fused_order = -1;       // -1 = by shape; 0 = [q][s] contiguous ranges; 1 = [s][q] round-robin
  long long fused_ns = 0;           // samples per fused launch (0 = by shape)
  long long fused_l3at = 16;        // CTA-pair kernel: layer-2 k-step at which the previous chunk's output-layer MMAs are slipped in (0 = after the chunk)
  long long fused_stream = 1;       // fused path: sampler and fused kernel of a chunk run concurrently, gated by ready flags (0 = one event per chunk)
  long long fused_np = 1;           // CTA-pair kernel: TMA producer warps (2 = even / odd stages from two threads)
  long long fused_kd = 128;         // CTA-pair kernel: k-depth of a layer-1 TMA stage (128 = 3 x 64 KB stages, 64 = 6 x 32 KB)
  long long fused_trace = 0, fused_trace_only = 0;   // in-kernel clock probes of one cluster (tools/trace_fused.py)
  long long ws_mb = 2048;           // workspace budget of the generic executor / IBP
  long long train_graph = 0;        // dbx_bbb_train_epoch (bf16 mode): 1 = replay one captured step as a CUDA graph instead of launching every
                                    // step's kernels (bit-identical; measured SLOWER on this driver -- 0.33 vs 0.17 ms per step of
                                    // 784-512-512-10 at mini-batch 128 over 400 steps -- so it is off by default)
};
Options& opt();

#define DBX_CUDA(expr)                                                                      \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess)                                                                  \
      return ::dbx::set_err("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
  } while (0)

#define DBX_CHECK(cond, ...)                         \
  do {                                               \
    if (!(cond)) return ::dbx::set_err(__VA_ARGS__); \
  } while (0)

// every kernel launch goes through this so `gpu_launches` is a count, not a guess
#define DBX_LAUNCH(kernel, grid, block, smem, stream, ...)                 \
  do {                                                                     \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);            \
    ::dbx::g_launches.fetch_add(1, std::memory_order_relaxed);             \
    DBX_CUDA(cudaGetLastError());                                          \
  } while (0)

// SM count of the current device (queried once; the grid-stride kernels size their grids as a multiple of it)
int sm_count();

static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t round_up(int64_t a, int64_t b) { return cdiv(a, b) * b; }

// ---------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. SC'11).  counter = (blk lo, blk hi, sample lo, sample hi),
// key = (seed lo, seed hi); block blk yields the noise of flat elements 4*blk .. 4*blk+3.
// ---------------------------------------------------------------------------------------
struct Philox4 { uint32_t x, y, z, w; };

__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                 uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    // written as mulhi + mul: nvcc still fuses each pair into one IMAD.WIDE, but (unlike the uint64 form) without the
    // 2-3 register-shuffling adds per round that cost 12 % of the sampler's instructions
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return Philox4{c0, c1, c2, c3};
}

// Box-Muller on two 32-bit words: u1 = (a+1)*2^-32 in (0,1], u2 = b*2^-32 in [0,1).
// rad = sqrt(-2 ln u1) straight on the special-function unit: lg2.approx and sqrt.approx are one MUFU each (u1 >= 2^-32
// is far from the denormals, so the IEEE slow paths sqrtf()/logf() carry are dead weight here: 12 of the 46
// instructions per normal).  Relative error ~2^-22, inside the 5e-6 absolute bound the parity tests put on eps.
__device__ __forceinline__ float dbx_lg2_approx(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float dbx_sqrt_approx(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& z0, float& z1) {
  // (a + 1) * 2^-32 computed so that it never rounds to 0; it can round up to exactly 1 (rad = 0), never above
  const float u1 = fmaf((float)a, 2.3283064365386963e-10f, 2.3283064365386963e-10f);
  const float u2 = (float)b * 2.3283064365386963e-10f;
  const float rad = dbx_sqrt_approx(-1.3862943611198906f * dbx_lg2_approx(u1));   // -2 ln 2 * log2(u1)
  float sn, cs;
  __sincosf(6.283185307179586f * (u2 - 0.5f), &sn, &cs);  // angle in [-pi, pi)
  z0 = rad * cs;
  z1 = rad * sn;
}

__device__ __forceinline__ void philox_normal4(uint64_t seed, uint64_t sample, uint64_t blk, float z[4]) {
  Philox4 r = philox4x32_10((uint32_t)blk, (uint32_t)(blk >> 32), (uint32_t)sample, (uint32_t)(sample >> 32),
                            (uint32_t)seed, (uint32_t)(seed >> 32));
  box_muller(r.x, r.y, z[0], z[1]);
  box_muller(r.z, r.w, z[2], z[3]);
}

// tf.math.softplus with TensorFlow's CPU thresholds.
__device__ __forceinline__ float softplus_tf(float x) {
  const float thr = -13.942385f;  // ln(FLT_EPSILON) + 2
  if (x > -thr) return x;
  const float ex = expf(x);
  if (x < thr) return ex;
  return log1pf(ex);
}

__device__ __forceinline__ float apply_act(int act, float z) {
  switch (act) {
    case 1: return fmaxf(z, 0.f);
    case 3: return tanhf(z);
    case 4: return 1.f / (1.f + expf(-z));
    default: return z;  // linear; softmax is applied by the accumulation kernels
  }
}

__device__ __forceinline__ uint32_t tc_pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

}  // namespace dbx
