// Batched per-sample GEMM on tcgen05 for the GENERIC bf16 path: one Dense layer (or an im2col'd
// Conv2D) of every posterior sample of a chunk,
//     Y[s] = act( A[s] (M x K, bf16, K-major)  *  W_s (K x N, bf16, Keras (fan_in, fan_out) layout)  + b_s )
// for models the fused MLP kernel does not cover (hidden width > 512, more than two hidden layers,
// convolutions).  Classic Blackwell GEMM structure: TMA producer warp, one MMA-issuing warp
// (tcgen05.mma, accumulator in TMEM), four epilogue warps (tcgen05.ld -> bias + activation ->
// bf16 / fp32 stores).  One CTA per (128-row tile, <=256-column tile, sample).
#include <algorithm>
#include <cstdlib>
#include "engine.h"
#include "tc05.cuh"

namespace dbx {

using namespace tc05;

constexpr int kGStages = 4;
constexpr int kGThreads = 192;

struct GemmParams {
  int M, N, K, NT;          // NT = columns handled per CTA (multiple of 64, <= 256)
  int a_shared;             // A is the same for every sample (layer 0)
  const float* bias; long long b_sstride; long long b_off;
  const int32_t* rows; long long row0;      // sample -> weight-row indirection of the bias array (chunk-local rows here)
  int act;
  __nv_bfloat16* out16; long long o16_sstride; int ld16;    // bf16 output [s][M][ld16] (next layer's A) or null
  float* out32; long long o32_sstride;                      // fp32 output [s][M][N] (logits of the last layer) or null
};

struct __align__(8) GBarriers {
  uint64_t full[kGStages], empty[kGStages];
  uint64_t acc_full;
  uint32_t tmem_slot;
};

extern __shared__ __align__(1024) uint8_t gemm_smem_raw[];

// The epilogues apply the activation to 32-128 values per thread and tile; with the activation as a run-time switch that
// was a branch per element (the conv kernel's hottest instructions).  kAct: 1 = relu, 0 = linear, -1 = look at `act`.
template <int kAct>
__device__ __forceinline__ float act_fixed(int act, float z) {
  if (kAct == 1) return fmaxf(z, 0.f);
  if (kAct == 0) return z;
  return apply_act(act, z);
}
static inline int act_class(int act) { return act == DBX_ACT_RELU ? 1 : ((act == DBX_ACT_LINEAR || act == DBX_ACT_SOFTMAX) ? 0 : -1); }

// kNTMax = widest column tile of the launch (64 / 128 / 256): it fixes the stage size (A block 16 KB + kNTMax/64 weight
// blocks of 8 KB) and the TMEM allocation, so narrow layers (conv kernels with 32 / 64 output channels) run 3-4 CTAs
// per SM instead of one -- their tiles have 1-5 k-blocks, and a lone CTA per SM spent most of its life in the prologue.
template <int kNTMax, int kAct>
__global__ void __launch_bounds__(kGThreads, kNTMax == 256 ? 1 : (kNTMax == 128 ? 2 : 3))
dense_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const GemmParams p) {
  constexpr int kStageB = 16384 + (kNTMax / 64) * 8192;
  constexpr int kSt = kNTMax == 64 ? 3 : kGStages;            // 3 x 24 KB: three CTAs of a narrow layer fit one SM
  uint8_t* smem = (uint8_t*)(((uintptr_t)gemm_smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t s_stage = smem_u32(smem);
  GBarriers* bars = (GBarriers*)(smem + kSt * kStageB);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n0 = blockIdx.x * p.NT, m0 = blockIdx.y * 128, s = blockIdx.z;
  const int nt = min(p.NT, ((p.N - n0 + 63) / 64) * 64);     // columns of this tile, multiple of 64
  const int nblk = nt / 64;
  const int kb_n = (p.K + 63) / 64;

  if (tid == 0) {
    for (int i = 0; i < kSt; ++i) { mbar_init(smem_u32(&bars->full[i]), 1); mbar_init(smem_u32(&bars->empty[i]), 1); }
    mbar_init(smem_u32(&bars->acc_full), 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<kNTMax>(smem_u32(&bars->tmem_slot));
  if (warp == 0 && lane == 0) { prefetch_tmap(&tmA); prefetch_tmap(&tmW); }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = bars->tmem_slot;

  if (warp == 0) {
    for (int kb = 0; kb < kb_n; ++kb) {
      const uint32_t st = kb % kSt, ph = (kb / kSt) & 1;
      mbar_wait(smem_u32(&bars->empty[st]), ph ^ 1);
      if (elect_one()) {
        const uint32_t fb = smem_u32(&bars->full[st]);
        const uint32_t base = s_stage + st * kStageB;
        mbar_expect_tx(fb, 16384 + nblk * 8192);
        if (p.a_shared) tma_load_2d(base, &tmA, fb, kb * 64, m0);
        else tma_load_3d(base, &tmA, fb, kb * 64, m0, s);
        for (int nb = 0; nb < nblk; ++nb) tma_load_3d(base + 16384 + nb * 8192, &tmW, fb, n0 + nb * 64, kb * 64, s);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    constexpr uint32_t kDescHi128 = (1024u >> 4) | (1u << 14) | (2u << 29);
    const uint32_t idesc = make_idesc_bf16(128, nt, 0, 1);
    for (int kb = 0; kb < kb_n; ++kb) {
      const uint32_t st = kb % kSt, ph = (kb / kSt) & 1;
      mbar_wait(smem_u32(&bars->full[st]), ph);
      fence_after_sync();
      if (elect_one()) {
        const uint32_t a16 = ((s_stage + st * kStageB) & 0x3FFFF) >> 4;
        const uint32_t a_lo = a16 | (1u << 16);
        const uint32_t b_lo = (a16 + (16384 >> 4)) | ((8192u >> 4) << 16);
        const int ksteps = min(4, (p.K - kb * 64 + 15) / 16);
        for (int kk = 0; kk < ksteps; ++kk)
          mma_ss_lh(tmem, a_lo + kk * 2, kDescHi128, b_lo + kk * 128, kDescHi128, idesc, (kb | kk) > 0);
        mma_commit(smem_u32(&bars->empty[st]));
        if (kb == kb_n - 1) mma_commit(smem_u32(&bars->acc_full));
      }
      __syncwarp();
    }
  } else {
    const int g = warp & 3;
    const int row = g * 32 + lane;
    const int gm = m0 + row;
    const uint32_t lane_base = (uint32_t)(g * 32) << 16;
    const long long brow = p.rows ? (long long)p.rows[s] : (p.row0 + s);
    const float* bias = p.bias + brow * p.b_sstride + p.b_off;
    mbar_wait(smem_u32(&bars->acc_full), 0);
    fence_after_sync();
    for (int c0 = 0; c0 < nt; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem + lane_base + c0, v);
      wait_ld();
      if (gm < p.M) {
        if (p.out16) {
          __nv_bfloat16* dst = p.out16 + (long long)s * p.o16_sstride + (long long)gm * p.ld16 + n0 + c0;
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {   // 8 bf16 = 16 bytes per store
            const int c = n0 + c0 + q4 * 8;
            if (c + 8 <= p.N) {
              uint32_t w[4];
#pragma unroll
              for (int i = 0; i < 4; ++i)
                w[i] = pack_bf16x2(act_fixed<kAct>(p.act, __uint_as_float(v[q4 * 8 + 2 * i]) + bias[c + 2 * i]),
                                   act_fixed<kAct>(p.act, __uint_as_float(v[q4 * 8 + 2 * i + 1]) + bias[c + 2 * i + 1]));
              *reinterpret_cast<uint4*>(dst + q4 * 8) = make_uint4(w[0], w[1], w[2], w[3]);
            } else {
              for (int i = 0; i < 8; ++i)
                if (c + i < p.N) dst[q4 * 8 + i] = __float2bfloat16_rn(act_fixed<kAct>(p.act, __uint_as_float(v[q4 * 8 + i]) + bias[c + i]));
            }
          }
        } else {
          float* dst = p.out32 + (long long)s * p.o32_sstride + (long long)gm * p.N;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int c = n0 + c0 + i;
            if (c < p.N) dst[c] = act_fixed<kAct>(p.act, __uint_as_float(v[i]) + bias[c]);
          }
        }
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc<kNTMax>(tmem);
}

// =======================================================================================
// Backward-data GEMM of a Dense layer for every sample of a chunk (expected input gradient, attacks.py:49-75; the FGSM step
// of massart_bound_check, verifiers.py:622-625):
//     dA[s] (M x Kin) = ( dZ[s] (M x N, bf16, N contiguous)  *  W_s^T ) .* act'(Aprev[s])
// The reduction runs over the layer's OUTPUT features, so the Keras (fan_in, fan_out) kernel as it lies in the chunk's bf16
// layout IS the K-major B operand (row = fan_in index = output column of this GEMM, fan_out contiguous): one TMA box
// [64 fan_out x NT fan_in rows], SWIZZLE_128B, the same canonical layout as the A block -- no transposed copy of W_s exists.
// Same warp roles as dense_tc_kernel; the epilogue multiplies by the derivative of the previous layer's activation taken
// from its stored (post-activation, bf16) output and writes bf16 (the next GEMM's dZ) or fp32 (the gradient at the input).
// =======================================================================================
struct __align__(8) PBarriers {
  uint64_t full[4], empty[4];
  uint64_t acc_full[2], acc_empty[2];
  uint32_t tmem_slot;
};

struct BwdParams {
  int M, N, K, NT;          // N = fan_in (output columns), K = fan_out (reduction); NT = output columns per CTA
  const __nv_bfloat16* aprev; long long ap_sstride; int ld_ap; int act;     // previous layer's output [s][M][ld_ap] or null
  __nv_bfloat16* out16; long long o16_sstride; int ld16;
  float* out32; long long o32_sstride;
  // FGSM epilogue (the gradient at the input is only needed for its sign, attacks.py:94-101): with fgsm_x = the call's input
  // [M][N] fp32, out16 receives clip(clip(x + eps sign(g), x - eps, x + eps), lower, upper) as bf16 -- the A operand of the
  // forward pass at every sample's own adversarial input; the fp32 gradient is never written
  const float* fgsm_x; float fgsm_eps, fgsm_lo, fgsm_hi;
};

__device__ __forceinline__ float act_deriv(int act, float a) {      // derivative written in terms of the activation's OUTPUT
  switch (act) {
    case DBX_ACT_RELU: return a > 0.f ? 1.f : 0.f;
    case DBX_ACT_TANH: return 1.f - a * a;
    case DBX_ACT_SIGMOID: return a * (1.f - a);
    default: return 1.f;
  }
}

// What the epilogue reads from global memory for one thread's 32 columns [c0g, c0g + 32) of row gm: the previous layer's
// output (4 x 16 bytes of bf16) or, for the FGSM epilogue, the call's input (8 x 16 bytes of fp32).  The loads are issued
// BEFORE the accumulator is waited for and before the previous group's stores (which the compiler cannot move loads across:
// it does not know the buffers are distinct) -- issued one at a time between the stores, each cost a full memory latency and the
// epilogue ran at a 32nd of the L2's rate (ncu: long-scoreboard stalls, 9 % warps active).
struct BwdPre { uint4 q[8]; };
__device__ __forceinline__ bool bwd_vec_ok(const BwdParams& p) {
  return p.fgsm_x ? ((p.N & 3) == 0) : (p.aprev == nullptr || (p.ld_ap & 7) == 0);
}
__device__ __forceinline__ void bwd_prefetch(const BwdParams& p, int s, int gm, int c0g, BwdPre& pre) {
  if (!bwd_vec_ok(p)) return;
  if (p.fgsm_x) {
    const float* xr = p.fgsm_x + (long long)gm * p.N + c0g;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      pre.q[i] = (c0g + i * 4 + 4 <= p.N) ? __ldg(reinterpret_cast<const uint4*>(xr + i * 4)) : make_uint4(0u, 0u, 0u, 0u);
  } else if (p.aprev) {
    const __nv_bfloat16* ap = p.aprev + (long long)s * p.ap_sstride + (long long)gm * p.ld_ap + c0g;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      pre.q[i] = (c0g + i * 8 + 8 <= p.N) ? __ldg(reinterpret_cast<const uint4*>(ap + i * 8)) : make_uint4(0u, 0u, 0u, 0u);
  }
}

// one thread's 32 accumulator columns [c0g, c0g + 32) of row gm, sample s
__device__ __forceinline__ void bwd_epilogue32(const BwdParams& p, int s, int gm, int c0g, const uint32_t (&v)[32], const BwdPre& pre) {
  const bool vec = bwd_vec_ok(p);
  // The epilogue is ONE warp per scheduler walking its rows: its time is instructions x issue latency (ncu: 18 cycles per
  // instruction, 360 instructions per 32 columns with per-element guards).  Whole groups inside the matrix with 16-byte
  // pitches take this guard-free form.
  if (vec && c0g + 32 <= p.N && p.out16 && (p.ld16 & 7) == 0 && !p.fgsm_x && (p.act == DBX_ACT_RELU || p.aprev == nullptr)) {
    uint4* dst = reinterpret_cast<uint4*>(p.out16 + (long long)s * p.o16_sstride + (long long)gm * p.ld16 + c0g);
    const bool mask = p.aprev != nullptr;
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4) {
      const uint32_t aw[4] = {pre.q[q4].x, pre.q[q4].y, pre.q[q4].z, pre.q[q4].w};
      uint32_t w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float lo = __uint_as_float(v[q4 * 8 + 2 * i]), hi = __uint_as_float(v[q4 * 8 + 2 * i + 1]);
        if (mask) {      // relu'(a) = [a > 0] on the stored bf16 output
          lo = __uint_as_float(aw[i] << 16) > 0.f ? lo : 0.f;
          hi = __uint_as_float(aw[i] & 0xFFFF0000u) > 0.f ? hi : 0.f;
        }
        w[i] = pack_bf16x2(lo, hi);
      }
      dst[q4] = make_uint4(w[0], w[1], w[2], w[3]);
    }
    return;
  }
  if (vec && c0g + 32 <= p.N && p.fgsm_x && p.out16 && (p.ld16 & 7) == 0 && p.aprev == nullptr) {
    uint4* dst = reinterpret_cast<uint4*>(p.out16 + (long long)s * p.o16_sstride + (long long)gm * p.ld16 + c0g);
    const float e = p.fgsm_eps, lo_b = p.fgsm_lo, hi_b = p.fgsm_hi;
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4) {
      const uint4 xl = pre.q[2 * q4], xh = pre.q[2 * q4 + 1];
      const float xe[8] = {__uint_as_float(xl.x), __uint_as_float(xl.y), __uint_as_float(xl.z), __uint_as_float(xl.w),
                           __uint_as_float(xh.x), __uint_as_float(xh.y), __uint_as_float(xh.z), __uint_as_float(xh.w)};
      float a[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float g = __uint_as_float(v[q4 * 8 + i]);
        // x + eps sign(g) lies in [x - eps, x + eps] already (attacks.py:94-99); only the model's input range clips
        const float step = g > 0.f ? e : (g < 0.f ? -e : 0.f);
        a[i] = fminf(fmaxf(fminf(fmaxf(xe[i] + step, xe[i] - e), xe[i] + e), lo_b), hi_b);
      }
      dst[q4] = make_uint4(pack_bf16x2(a[0], a[1]), pack_bf16x2(a[2], a[3]), pack_bf16x2(a[4], a[5]), pack_bf16x2(a[6], a[7]));
    }
    return;
  }
  const __nv_bfloat16* ap = p.aprev ? p.aprev + (long long)s * p.ap_sstride + (long long)gm * p.ld_ap : nullptr;
#pragma unroll
  for (int q4 = 0; q4 < 4; ++q4) {
    const int c = c0g + q4 * 8;
    if (c < p.N) {
      float d[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) d[i] = __uint_as_float(v[q4 * 8 + i]);
      const bool whole = c + 8 <= p.N;
      if (ap) {
        if (whole && vec) {
          const uint32_t aw[4] = {pre.q[q4].x, pre.q[q4].y, pre.q[q4].z, pre.q[q4].w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            d[2 * i] *= act_deriv(p.act, __uint_as_float(aw[i] << 16));
            d[2 * i + 1] *= act_deriv(p.act, __uint_as_float(aw[i] & 0xFFFF0000u));
          }
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (c + i < p.N) d[i] *= act_deriv(p.act, __bfloat162float(ap[c + i]));
        }
      }
      if (p.fgsm_x) {
        float xe[8];
        if (vec) {      // N % 4 == 0: a float4 is inside the row or outside (then zero, and its columns are not stored)
          const uint4 lo = pre.q[2 * q4], hi = pre.q[2 * q4 + 1];
          xe[0] = __uint_as_float(lo.x); xe[1] = __uint_as_float(lo.y); xe[2] = __uint_as_float(lo.z); xe[3] = __uint_as_float(lo.w);
          xe[4] = __uint_as_float(hi.x); xe[5] = __uint_as_float(hi.y); xe[6] = __uint_as_float(hi.z); xe[7] = __uint_as_float(hi.w);
        } else {
          const float* xr = p.fgsm_x + (long long)gm * p.N + c;
#pragma unroll
          for (int i = 0; i < 8; ++i) xe[i] = (c + i < p.N) ? xr[i] : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float sgn = d[i] > 0.f ? 1.f : (d[i] < 0.f ? -1.f : 0.f);
          float a = xe[i] + p.fgsm_eps * sgn;
          a = fminf(fmaxf(a, xe[i] - p.fgsm_eps), xe[i] + p.fgsm_eps);
          d[i] = fminf(fmaxf(a, p.fgsm_lo), p.fgsm_hi);
        }
      }
      if (p.out16) {
        __nv_bfloat16* dst = p.out16 + (long long)s * p.o16_sstride + (long long)gm * p.ld16 + c;
        if (whole && (p.ld16 & 7) == 0) {
          *reinterpret_cast<uint4*>(dst) = make_uint4(pack_bf16x2(d[0], d[1]), pack_bf16x2(d[2], d[3]), pack_bf16x2(d[4], d[5]),
                                                      pack_bf16x2(d[6], d[7]));
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) if (c + i < p.N) dst[i] = __float2bfloat16_rn(d[i]);
        }
      } else {
        float* dst = p.out32 + (long long)s * p.o32_sstride + (long long)gm * p.N + c;
        if (whole && (p.N & 3) == 0) {
          *reinterpret_cast<float4*>(dst) = make_float4(d[0], d[1], d[2], d[3]);
          *reinterpret_cast<float4*>(dst + 4) = make_float4(d[4], d[5], d[6], d[7]);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) if (c + i < p.N) dst[i] = d[i];
        }
      }
    }
  }
}

template <int kNTMax>
__global__ void __launch_bounds__(kGThreads, kNTMax == 256 ? 1 : 2)
dense_tc_bwd_kernel(const __grid_constant__ CUtensorMap tmZ, const __grid_constant__ CUtensorMap tmW, const BwdParams p) {
  constexpr int kStageB = 16384 + kNTMax * 128;
  constexpr int kSt = kNTMax == 64 ? 3 : kGStages;
  uint8_t* smem = (uint8_t*)(((uintptr_t)gemm_smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t s_stage = smem_u32(smem);
  GBarriers* bars = (GBarriers*)(smem + kSt * kStageB);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n0 = blockIdx.x * p.NT, m0 = blockIdx.y * 128, s = blockIdx.z;
  const int nt = min(p.NT, ((p.N - n0 + 63) / 64) * 64);
  const int kb_n = (p.K + 63) / 64;

  if (tid == 0) {
    for (int i = 0; i < kSt; ++i) { mbar_init(smem_u32(&bars->full[i]), 1); mbar_init(smem_u32(&bars->empty[i]), 1); }
    mbar_init(smem_u32(&bars->acc_full), 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<kNTMax>(smem_u32(&bars->tmem_slot));
  if (warp == 0 && lane == 0) { prefetch_tmap(&tmZ); prefetch_tmap(&tmW); }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = bars->tmem_slot;

  if (warp == 0) {
    for (int kb = 0; kb < kb_n; ++kb) {
      const uint32_t st = kb % kSt, ph = (kb / kSt) & 1;
      mbar_wait(smem_u32(&bars->empty[st]), ph ^ 1);
      if (elect_one()) {
        const uint32_t fb = smem_u32(&bars->full[st]);
        const uint32_t base = s_stage + st * kStageB;
        mbar_expect_tx(fb, 16384 + p.NT * 128);       // a box is written whole; what lies outside the tensor arrives as zeros
        tma_load_3d(base, &tmZ, fb, kb * 64, m0, s);
        tma_load_3d(base + 16384, &tmW, fb, kb * 64, n0, s);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    constexpr uint32_t kDescHi128 = (1024u >> 4) | (1u << 14) | (2u << 29);
    const uint32_t idesc = make_idesc_bf16(128, nt, 0, 0);
    for (int kb = 0; kb < kb_n; ++kb) {
      const uint32_t st = kb % kSt, ph = (kb / kSt) & 1;
      mbar_wait(smem_u32(&bars->full[st]), ph);
      fence_after_sync();
      if (elect_one()) {
        const uint32_t a16 = ((s_stage + st * kStageB) & 0x3FFFF) >> 4;
        const uint32_t a_lo = a16 | (1u << 16);
        const uint32_t b_lo = (a16 + (16384 >> 4)) | (1u << 16);
        const int ksteps = min(4, (p.K - kb * 64 + 15) / 16);
        for (int kk = 0; kk < ksteps; ++kk)
          mma_ss_lh(tmem, a_lo + kk * 2, kDescHi128, b_lo + kk * 2, kDescHi128, idesc, (kb | kk) > 0);
        mma_commit(smem_u32(&bars->empty[st]));
        if (kb == kb_n - 1) mma_commit(smem_u32(&bars->acc_full));
      }
      __syncwarp();
    }
  } else {
    const int g = warp & 3;
    const int gm = m0 + g * 32 + lane;
    const uint32_t lane_base = (uint32_t)(g * 32) << 16;
    const bool live = gm < p.M;
    BwdPre preA, preB;      // nt is a multiple of 64: column groups go in pairs, each group's loads in flight during the one before
    if (live) bwd_prefetch(p, s, gm, n0, preA);
    mbar_wait(smem_u32(&bars->acc_full), 0);
    fence_after_sync();
    for (int c0 = 0; c0 < nt; c0 += 64) {
      uint32_t v[32];
      tmem_ld32(tmem + lane_base + c0, v);
      if (live) bwd_prefetch(p, s, gm, n0 + c0 + 32, preB);
      wait_ld();
      if (live) bwd_epilogue32(p, s, gm, n0 + c0, v, preA);
      tmem_ld32(tmem + lane_base + c0 + 32, v);
      if (live && c0 + 64 < nt) bwd_prefetch(p, s, gm, n0 + c0 + 64, preA);
      wait_ld();
      if (live) bwd_epilogue32(p, s, gm, n0 + c0 + 32, v, preB);
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc<kNTMax>(tmem);
}

// Persistent form (the launches of the gradient loops have hundreds of (sample, column tile) pairs): CTAs walk a contiguous
// range of the tile list; TMA ring, two TMEM accumulators and the epilogue of tile i overlap the loads and MMAs of tile i + 1
// (the structure of dense_tc_persist_kernel).
template <int kNT>
__global__ void __launch_bounds__(kGThreads, kNT == 256 ? 1 : 2)
dense_tc_bwd_persist_kernel(const __grid_constant__ CUtensorMap tmZ, const __grid_constant__ CUtensorMap tmW, const BwdParams p, int ns,
                            int tiles_m, int tiles_n) {
  constexpr int kStB = 16384 + kNT * 128;
  constexpr int kSt = kNT == 128 ? 3 : 4;
  uint8_t* smem = (uint8_t*)(((uintptr_t)gemm_smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t s_stage = smem_u32(smem);
  PBarriers* bars = (PBarriers*)(smem + kSt * kStB);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int kb_n = (p.K + 63) / 64;
  const long long total = (long long)ns * tiles_m * tiles_n;
  const long long t_begin = (total * blockIdx.x) / gridDim.x, t_end = (total * (blockIdx.x + 1)) / gridDim.x;
  // (column tile, row tile, sample) of a tile: decoded once per role, then advanced -- four 64-bit divisions per tile and warp
  // were as many instructions as a short tile's whole epilogue
  struct TileIt { int nt_i, mt_i, s; };
  auto decode0 = [&](long long t) {
    TileIt ti;
    ti.nt_i = (int)(t % tiles_n);
    const long long r = t / tiles_n;
    ti.mt_i = (int)(r % tiles_m); ti.s = (int)(r / tiles_m);
    return ti;
  };
  auto advance = [&](TileIt& ti) {
    if (++ti.nt_i == tiles_n) { ti.nt_i = 0; if (++ti.mt_i == tiles_m) { ti.mt_i = 0; ++ti.s; } }
  };

  if (tid == 0) {
    for (int i = 0; i < kSt; ++i) { mbar_init(smem_u32(&bars->full[i]), 1); mbar_init(smem_u32(&bars->empty[i]), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&bars->acc_full[i]), 1); mbar_init(smem_u32(&bars->acc_empty[i]), 4); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<2 * kNT>(smem_u32(&bars->tmem_slot));
  if (warp == 0 && lane == 0) { prefetch_tmap(&tmZ); prefetch_tmap(&tmW); }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = bars->tmem_slot;

  if (warp == 0) {
    uint32_t g = 0;
    TileIt ti = decode0(t_begin);
    for (long long t = t_begin; t < t_end; ++t, advance(ti)) {
      const int n0 = ti.nt_i * kNT, m0 = ti.mt_i * 128, s = ti.s;
      for (int kb = 0; kb < kb_n; ++kb, ++g) {
        const uint32_t st = g % kSt, ph = (g / kSt) & 1;
        mbar_wait(smem_u32(&bars->empty[st]), ph ^ 1);
        if (elect_one()) {
          const uint32_t fb = smem_u32(&bars->full[st]);
          const uint32_t base = s_stage + st * kStB;
          mbar_expect_tx(fb, 16384 + kNT * 128);
          tma_load_3d(base, &tmZ, fb, kb * 64, m0, s);
          tma_load_3d(base + 16384, &tmW, fb, kb * 64, n0, s);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t kDescHi128 = (1024u >> 4) | (1u << 14) | (2u << 29);
    uint32_t g = 0, it = 0;
    TileIt ti = decode0(t_begin);
    for (long long t = t_begin; t < t_end; ++t, ++it, advance(ti)) {
      const int n0 = ti.nt_i * kNT;
      const int nt = min(kNT, ((p.N - n0 + 63) / 64) * 64);
      const uint32_t idesc = make_idesc_bf16(128, nt, 0, 0);
      const uint32_t ab = it & 1;
      mbar_wait(smem_u32(&bars->acc_empty[ab]), ((it >> 1) & 1) ^ 1);
      fence_after_sync();
      for (int kb = 0; kb < kb_n; ++kb, ++g) {
        const uint32_t st = g % kSt, ph = (g / kSt) & 1;
        mbar_wait(smem_u32(&bars->full[st]), ph);
        if (elect_one()) {
          const uint32_t a16 = ((s_stage + st * kStB) & 0x3FFFF) >> 4;
          const uint32_t a_lo = a16 | (1u << 16);
          const uint32_t b_lo = (a16 + (16384 >> 4)) | (1u << 16);
          const int ksteps = min(4, (p.K - kb * 64 + 15) / 16);
          for (int kk = 0; kk < ksteps; ++kk)
            mma_ss_lh(tmem + ab * kNT, a_lo + kk * 2, kDescHi128, b_lo + kk * 2, kDescHi128, idesc, (kb | kk) > 0);
          mma_commit(smem_u32(&bars->empty[st]));
          if (kb == kb_n - 1) mma_commit(smem_u32(&bars->acc_full[ab]));
        }
        __syncwarp();
      }
    }
  } else {
    const int gq = warp & 3;
    const uint32_t lane_base = (uint32_t)(gq * 32) << 16;
    uint32_t it = 0;
    TileIt ti = decode0(t_begin);
    for (long long t = t_begin; t < t_end; ++t, ++it, advance(ti)) {
      const int n0 = ti.nt_i * kNT, m0 = ti.mt_i * 128, s = ti.s;
      const int nt = min(kNT, ((p.N - n0 + 63) / 64) * 64);
      const int gm = m0 + gq * 32 + lane;
      const uint32_t ab = it & 1;
      const bool live = gm < p.M;
      BwdPre preA, preB;      // nt is a multiple of 64: column groups go in pairs, each group's loads in flight during the one before
      if (live) bwd_prefetch(p, s, gm, n0, preA);
      mbar_wait(smem_u32(&bars->acc_full[ab]), (it >> 1) & 1);
      fence_after_sync();
      for (int c0 = 0; c0 < nt; c0 += 64) {
        uint32_t v[32];
        tmem_ld32(tmem + lane_base + ab * kNT + c0, v);
        if (live) bwd_prefetch(p, s, gm, n0 + c0 + 32, preB);
        wait_ld();
        if (live) bwd_epilogue32(p, s, gm, n0 + c0, v, preA);
        tmem_ld32(tmem + lane_base + ab * kNT + c0 + 32, v);
        if (live && c0 + 64 < nt) bwd_prefetch(p, s, gm, n0 + c0 + 64, preA);
        wait_ld();
        if (c0 + 64 >= nt) {   // the whole accumulator is in registers: the MMA warp may reuse it
          fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&bars->acc_empty[ab]));
        }
        if (live) bwd_epilogue32(p, s, gm, n0 + c0 + 32, v, preB);
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc<2 * kNT>(tmem);
}

// =======================================================================================
// Weight-gradient GEMM of a Dense layer (Bayes-by-Backprop step, bayesbybackprop.py:138-170: dL/dW of the sampled weights):
//     dW (fan_in x fan_out, fp32, Keras layout)  (+)=  A^T (fan_in x B)  *  dZ (B x fan_out)
// The reduction runs over the BATCH rows, which is the slow axis of both operands as they lie in memory (A = the layer's
// bf16 input [B][fan_in], dZ = the bf16 upstream gradient [B][fan_out]): both are MN-major operands -- TMA boxes
// [64 columns x 64 batch rows], SWIZZLE_128B, 64-column blocks 8 KB apart (LBO), 8-row groups 1 KB apart (SBO), the layout the
// forward kernels use for the Keras kernel.  No transposed copy of the activations exists.  One CTA per 128 x <=256 tile of dW.
// =======================================================================================
struct WgradParams {
  int Kin, N, B, NT;        // dW is Kin x N; B = reduction length
  float* out; int accumulate;
};

template <int kNTMax>
__global__ void __launch_bounds__(kGThreads, kNTMax == 256 ? 1 : 2)
dense_tc_wgrad_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmZ, const WgradParams p) {
  constexpr int kStageB = 16384 + (kNTMax / 64) * 8192;
  constexpr int kSt = kNTMax == 64 ? 3 : kGStages;
  uint8_t* smem = (uint8_t*)(((uintptr_t)gemm_smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t s_stage = smem_u32(smem);
  GBarriers* bars = (GBarriers*)(smem + kSt * kStageB);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n0 = blockIdx.x * p.NT, m0 = blockIdx.y * 128;
  const int nt = min(p.NT, ((p.N - n0 + 63) / 64) * 64);
  const int nblk = nt / 64;
  const int kb_n = (p.B + 63) / 64;

  if (tid == 0) {
    for (int i = 0; i < kSt; ++i) { mbar_init(smem_u32(&bars->full[i]), 1); mbar_init(smem_u32(&bars->empty[i]), 1); }
    mbar_init(smem_u32(&bars->acc_full), 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<kNTMax>(smem_u32(&bars->tmem_slot));
  if (warp == 0 && lane == 0) { prefetch_tmap(&tmA); prefetch_tmap(&tmZ); }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = bars->tmem_slot;

  if (warp == 0) {
    for (int kb = 0; kb < kb_n; ++kb) {
      const uint32_t st = kb % kSt, ph = (kb / kSt) & 1;
      mbar_wait(smem_u32(&bars->empty[st]), ph ^ 1);
      if (elect_one()) {
        const uint32_t fb = smem_u32(&bars->full[st]);
        const uint32_t base = s_stage + st * kStageB;
        // (a tile's upper 64 rows may lie wholly outside dW: that block is not fetched, its accumulator rows are never stored)
        const bool up = m0 + 64 < p.Kin;
        mbar_expect_tx(fb, (up ? 16384 : 8192) + nblk * 8192);
        tma_load_2d(base, &tmA, fb, m0, kb * 64);                  // fan_in columns [m0, m0 + 64) of batch rows [kb*64, +64)
        if (up) tma_load_2d(base + 8192, &tmA, fb, m0 + 64, kb * 64);
        for (int nb = 0; nb < nblk; ++nb) tma_load_2d(base + 16384 + nb * 8192, &tmZ, fb, n0 + nb * 64, kb * 64);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    constexpr uint32_t kDescHi128 = (1024u >> 4) | (1u << 14) | (2u << 29);
    const uint32_t idesc = make_idesc_bf16(128, nt, 1, 1);
    for (int kb = 0; kb < kb_n; ++kb) {
      const uint32_t st = kb % kSt, ph = (kb / kSt) & 1;
      mbar_wait(smem_u32(&bars->full[st]), ph);
      fence_after_sync();
      if (elect_one()) {
        const uint32_t a16 = ((s_stage + st * kStageB) & 0x3FFFF) >> 4;
        const uint32_t a_lo = a16 | ((8192u >> 4) << 16);
        const uint32_t b_lo = (a16 + (16384 >> 4)) | ((8192u >> 4) << 16);
        const int ksteps = min(4, (p.B - kb * 64 + 15) / 16);
        for (int kk = 0; kk < ksteps; ++kk)
          mma_ss_lh(tmem, a_lo + kk * 128, kDescHi128, b_lo + kk * 128, kDescHi128, idesc, (kb | kk) > 0);
        mma_commit(smem_u32(&bars->empty[st]));
        if (kb == kb_n - 1) mma_commit(smem_u32(&bars->acc_full));
      }
      __syncwarp();
    }
  } else {
    const int g = warp & 3;
    const int gm = m0 + g * 32 + lane;
    const uint32_t lane_base = (uint32_t)(g * 32) << 16;
    mbar_wait(smem_u32(&bars->acc_full), 0);
    fence_after_sync();
    for (int c0 = 0; c0 < nt; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem + lane_base + c0, v);
      wait_ld();
      if (gm < p.Kin) {
        float* dst = p.out + (long long)gm * p.N + n0 + c0;
        if (n0 + c0 + 32 <= p.N && (p.N & 3) == 0) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            float4 o = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
            if (p.accumulate) { const float4 old = *reinterpret_cast<const float4*>(dst + 4 * q); o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w; }
            *reinterpret_cast<float4*>(dst + 4 * q) = o;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (n0 + c0 + i < p.N) dst[i] = p.accumulate ? dst[i] + __uint_as_float(v[i]) : __uint_as_float(v[i]);
        }
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc<kNTMax>(tmem);
}

// =======================================================================================
// Persistent variant of dense_tc_kernel.  A CTA per tile spends most of a short tile's life in the serial chain
// TMEM alloc -> barrier init -> TMA -> MMA -> epilogue (config 4's first conv: 6 us per 128 x 32 tile with one
// k-block; the IBP GEMMs at B = 128: 50 us per CTA at 1.1 TB/s).  Here the CTAs (1 or 2 per SM) walk a contiguous range
// of the (sample, m tile, n tile) list; the TMA ring, the two TMEM accumulators and the epilogue of tile i overlap the
// loads and MMAs of tile i+1.  kNT = column tile (64 / 128 / 256).
// =======================================================================================

template <int kNT, int kAct>
__global__ void __launch_bounds__(kGThreads, kNT == 256 ? 1 : 2)
dense_tc_persist_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const GemmParams p, int ns,
                        int tiles_m, int tiles_n) {
  constexpr int kStB = 16384 + (kNT / 64) * 8192;
  constexpr int kSt = kNT == 128 ? 3 : 4;
  __shared__ float s_bias[kNT];
  uint8_t* smem = (uint8_t*)(((uintptr_t)gemm_smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t s_stage = smem_u32(smem);
  PBarriers* bars = (PBarriers*)(smem + kSt * kStB);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int kb_n = (p.K + 63) / 64;
  const long long total = (long long)ns * tiles_m * tiles_n;
  const long long t_begin = (total * blockIdx.x) / gridDim.x, t_end = (total * (blockIdx.x + 1)) / gridDim.x;
  // (column tile, row tile, sample) of a tile: decoded once per role, then advanced -- four 64-bit divisions per tile and warp
  // were as many instructions as a short tile's whole epilogue
  struct TileIt { int nt_i, mt_i, s; };
  auto decode0 = [&](long long t) {
    TileIt ti;
    ti.nt_i = (int)(t % tiles_n);
    const long long r = t / tiles_n;
    ti.mt_i = (int)(r % tiles_m); ti.s = (int)(r / tiles_m);
    return ti;
  };
  auto advance = [&](TileIt& ti) {
    if (++ti.nt_i == tiles_n) { ti.nt_i = 0; if (++ti.mt_i == tiles_m) { ti.mt_i = 0; ++ti.s; } }
  };

  if (tid == 0) {
    for (int i = 0; i < kSt; ++i) { mbar_init(smem_u32(&bars->full[i]), 1); mbar_init(smem_u32(&bars->empty[i]), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&bars->acc_full[i]), 1); mbar_init(smem_u32(&bars->acc_empty[i]), 4); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<2 * kNT>(smem_u32(&bars->tmem_slot));
  if (warp == 0 && lane == 0) { prefetch_tmap(&tmA); prefetch_tmap(&tmW); }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = bars->tmem_slot;

  if (warp == 0) {
    uint32_t g = 0;
    TileIt ti = decode0(t_begin);
    for (long long t = t_begin; t < t_end; ++t, advance(ti)) {
      const int n0 = ti.nt_i * kNT, m0 = ti.mt_i * 128, s = ti.s;
      const int nblk = min(kNT, ((p.N - n0 + 63) / 64) * 64) / 64;
      for (int kb = 0; kb < kb_n; ++kb, ++g) {
        const uint32_t st = g % kSt, ph = (g / kSt) & 1;
        mbar_wait(smem_u32(&bars->empty[st]), ph ^ 1);
        if (elect_one()) {
          const uint32_t fb = smem_u32(&bars->full[st]);
          const uint32_t base = s_stage + st * kStB;
          mbar_expect_tx(fb, 16384 + nblk * 8192);
          if (p.a_shared) tma_load_2d(base, &tmA, fb, kb * 64, m0);
          else tma_load_3d(base, &tmA, fb, kb * 64, m0, s);
          for (int nb = 0; nb < nblk; ++nb) tma_load_3d(base + 16384 + nb * 8192, &tmW, fb, n0 + nb * 64, kb * 64, s);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t kDescHi128 = (1024u >> 4) | (1u << 14) | (2u << 29);
    uint32_t g = 0, it = 0;
    TileIt ti = decode0(t_begin);
    for (long long t = t_begin; t < t_end; ++t, ++it, advance(ti)) {
      const int n0 = ti.nt_i * kNT;
      const int nt = min(kNT, ((p.N - n0 + 63) / 64) * 64);
      const uint32_t idesc = make_idesc_bf16(128, nt, 0, 1);
      const uint32_t ab = it & 1;
      mbar_wait(smem_u32(&bars->acc_empty[ab]), ((it >> 1) & 1) ^ 1);
      fence_after_sync();
      for (int kb = 0; kb < kb_n; ++kb, ++g) {
        const uint32_t st = g % kSt, ph = (g / kSt) & 1;
        mbar_wait(smem_u32(&bars->full[st]), ph);
        if (elect_one()) {
          const uint32_t a16 = ((s_stage + st * kStB) & 0x3FFFF) >> 4;
          const uint32_t a_lo = a16 | (1u << 16);
          const uint32_t b_lo = (a16 + (16384 >> 4)) | ((8192u >> 4) << 16);
          const int ksteps = min(4, (p.K - kb * 64 + 15) / 16);
          for (int kk = 0; kk < ksteps; ++kk)
            mma_ss_lh(tmem + ab * kNT, a_lo + kk * 2, kDescHi128, b_lo + kk * 128, kDescHi128, idesc, (kb | kk) > 0);
          mma_commit(smem_u32(&bars->empty[st]));
          if (kb == kb_n - 1) mma_commit(smem_u32(&bars->acc_full[ab]));
        }
        __syncwarp();
      }
    }
  } else {
    const int gq = warp & 3;
    const int row = gq * 32 + lane;
    const uint32_t lane_base = (uint32_t)(gq * 32) << 16;
    uint32_t it = 0;
    int cur_s = -1, cur_n0 = -1;
    TileIt ti = decode0(t_begin);
    for (long long t = t_begin; t < t_end; ++t, ++it, advance(ti)) {
      const int n0 = ti.nt_i * kNT, m0 = ti.mt_i * 128, s = ti.s;
      const int nt = min(kNT, ((p.N - n0 + 63) / 64) * 64);
      const int gm = m0 + row;
      const uint32_t ab = it & 1;
      if (s != cur_s || n0 != cur_n0) {      // this (sample, column tile)'s bias -> smem (per-column global loads stall the epilogue)
        asm volatile("bar.sync 1, 128;" ::: "memory");
        for (int c = tid - 64; c < kNT; c += 128) s_bias[c] = (n0 + c < p.N) ? p.bias[(long long)s * p.b_sstride + p.b_off + n0 + c] : 0.f;
        asm volatile("bar.sync 1, 128;" ::: "memory");
        cur_s = s; cur_n0 = n0;
      }
      mbar_wait(smem_u32(&bars->acc_full[ab]), (it >> 1) & 1);
      fence_after_sync();
      for (int c0 = 0; c0 < nt; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem + lane_base + ab * kNT + c0, v);
        wait_ld();
        if (c0 + 32 >= nt) {   // the whole accumulator is in registers / stored: the MMA warp may reuse it
          fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&bars->acc_empty[ab]));
        }
        if (gm < p.M) {
          if (p.out16) {
            __nv_bfloat16* dst = p.out16 + (long long)s * p.o16_sstride + (long long)gm * p.ld16 + n0 + c0;
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              const int c = n0 + c0 + q4 * 8;
              if (c + 8 <= p.N) {
                uint32_t w[4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
                  w[i] = pack_bf16x2(act_fixed<kAct>(p.act, __uint_as_float(v[q4 * 8 + 2 * i]) + s_bias[c0 + q4 * 8 + 2 * i]),
                                     act_fixed<kAct>(p.act, __uint_as_float(v[q4 * 8 + 2 * i + 1]) + s_bias[c0 + q4 * 8 + 2 * i + 1]));
                *reinterpret_cast<uint4*>(dst + q4 * 8) = make_uint4(w[0], w[1], w[2], w[3]);
              } else {
                for (int i = 0; i < 8; ++i)
                  if (c + i < p.N) dst[q4 * 8 + i] = __float2bfloat16_rn(act_fixed<kAct>(p.act, __uint_as_float(v[q4 * 8 + i]) + s_bias[c0 + q4 * 8 + i]));
              }
            }
          } else {
            float* dst = p.out32 + (long long)s * p.o32_sstride + (long long)gm * p.N;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int c = n0 + c0 + i;
              if (c < p.N) dst[c] = act_fixed<kAct>(p.act, __uint_as_float(v[i]) + s_bias[c0 + i]);
            }
          }
        }
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc<2 * kNT>(tmem);
}

// =======================================================================================
// CTA-pair GEMM (cta_group::2) for big per-sample GEMMs: tile = 256 rows x 256 columns on the two SMs of a cluster.  CTA r of
// the pair owns rows [r*128, r*128+128) of the tile (its A block, its TMEM lanes, its epilogue) and stages columns
// [r*128, r*128+128) of the weight block; one tcgen05.mma issued by the leader reads both halves.  Per k-block a CTA now
// pulls 16 KB (A) + 16 KB (half of W) through its L2 port for 2 x the MMA work of the single-CTA 128 x 256 tile, which pulled
// 48 KB: that kernel was bound by the ~42 B/clk/SM L2 -> SM rate (0.5 PFLOP/s at B >= 512), not by the tensor pipe.
// Persistent clusters, 6 x 32 KB TMA ring, two TMEM accumulators, epilogue overlapped with the next tile.
// =======================================================================================
constexpr int kPrStages = 6;
constexpr int kPrStageB = 32768;
struct __align__(8) PairBarriers {
  uint64_t full[kPrStages], empty[kPrStages];
  uint64_t acc_full[2], acc_empty[2];
  uint32_t tmem_slot;
};

template <int kAct>
__global__ void __launch_bounds__(kGThreads, 1)
dense_tc_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmO,
                     const GemmParams p, int ns, int tiles_m, int tiles_n, int tma_out) {
  __shared__ float s_bias[256];
  uint8_t* smem = (uint8_t*)(((uintptr_t)gemm_smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t s_stage = smem_u32(smem);
  uint8_t* stage_out = smem + kPrStages * kPrStageB;            // per epilogue warp: a [32 rows x 64 columns] bf16 box for the TMA store
  PairBarriers* bars = (PairBarriers*)(stage_out + 4 * 4096);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  constexpr uint16_t kBoth = 0x3;
  const int kb_n = (p.K + 63) / 64;
  const long long total = (long long)ns * tiles_m * tiles_n;
  const long long ncl = gridDim.x / 2, cl = blockIdx.x / 2;
  const long long t_begin = (total * cl) / ncl, t_end = (total * (cl + 1)) / ncl;
  auto decode = [&](long long t, int& n0, int& m0, int& s) {
    const int nt_i = (int)(t % tiles_n);
    const long long r = t / tiles_n;
    n0 = nt_i * 256; m0 = (int)(r % tiles_m) * 256 + (int)rank * 128; s = (int)(r / tiles_m);
  };

  if (tid == 0) {
    for (int i = 0; i < kPrStages; ++i) { mbar_init(smem_u32(&bars->full[i]), 1); mbar_init(smem_u32(&bars->empty[i]), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&bars->acc_full[i]), 1); mbar_init(smem_u32(&bars->acc_empty[i]), 8); }   // 4 epilogue warps of each CTA
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc2<512>(smem_u32(&bars->tmem_slot));
  if (warp == 0 && lane == 0) { prefetch_tmap(&tmA); prefetch_tmap(&tmW); }
  fence_before_sync();
  __syncthreads();
  cluster_sync();
  fence_after_sync();
  const uint32_t tmem = bars->tmem_slot;

  if (warp == 0) {
    // ---- producer (both CTAs): own A rows, own half of the W block; complete_tx goes to the LEADER's barrier
    uint32_t g = 0;
    for (long long t = t_begin; t < t_end; ++t) {
      int n0, m0, s;
      decode(t, n0, m0, s);
      for (int kb = 0; kb < kb_n; ++kb, ++g) {
        const uint32_t st = g % kPrStages, ph = (g / kPrStages) & 1;
        mbar_wait(smem_u32(&bars->empty[st]), ph ^ 1);
        if (elect_one()) {
          const uint32_t fb = smem_u32(&bars->full[st]);
          const uint32_t base = s_stage + st * kPrStageB;
          if (leader) mbar_expect_tx(fb, 2u * kPrStageB);
          if (p.a_shared) tma2_load_2d(base, &tmA, fb, kb * 64, m0);
          else tma2_load_3d(base, &tmA, fb, kb * 64, m0, s);
          for (int nb = 0; nb < 2; ++nb) tma2_load_3d(base + 16384 + nb * 8192, &tmW, fb, n0 + (int)rank * 128 + nb * 64, kb * 64, s);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    if (leader) {
      constexpr uint32_t kDescHi128 = (1024u >> 4) | (1u << 14) | (2u << 29);
      const uint32_t idesc = make_idesc_bf16(256, 256, 0, 1);
      uint32_t g = 0, it = 0;
      for (long long t = t_begin; t < t_end; ++t, ++it) {
        const uint32_t ab = it & 1;
        mbar_wait_cluster(smem_u32(&bars->acc_empty[ab]), ((it >> 1) & 1) ^ 1);
        fence_after_sync();
        for (int kb = 0; kb < kb_n; ++kb, ++g) {
          const uint32_t st = g % kPrStages, ph = (g / kPrStages) & 1;
          mbar_wait(smem_u32(&bars->full[st]), ph);
          fence_after_sync();
          if (elect_one()) {
            const uint32_t a16 = ((s_stage + st * kPrStageB) & 0x3FFFF) >> 4;
            const uint32_t a_lo = a16 | (1u << 16);
            const uint32_t b_lo = (a16 + (16384 >> 4)) | ((8192u >> 4) << 16);
            const int ksteps = min(4, (p.K - kb * 64 + 15) / 16);
            for (int kk = 0; kk < ksteps; ++kk)
              mma2_ss_lh(tmem + ab * 256, a_lo + kk * 2, kDescHi128, b_lo + kk * 128, kDescHi128, idesc, (kb | kk) > 0);
            mma2_commit_mc(smem_u32(&bars->empty[st]), kBoth);
            if (kb == kb_n - 1) mma2_commit_mc(smem_u32(&bars->acc_full[ab]), kBoth);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ---- epilogue (both CTAs): this CTA's 128 rows x 256 columns
    const int gq = warp & 3;
    const int row = gq * 32 + lane;
    const uint32_t lane_base = (uint32_t)(gq * 32) << 16;
    uint32_t it = 0;
    int cur_s = -1, cur_n0 = -1;
    for (long long t = t_begin; t < t_end; ++t, ++it) {
      int n0, m0, s;
      decode(t, n0, m0, s);
      const int gm = m0 + row;
      const uint32_t ab = it & 1;
      if (s != cur_s || n0 != cur_n0) {
        asm volatile("bar.sync 1, 128;" ::: "memory");
        for (int c = tid - 64; c < 256; c += 128) s_bias[c] = (n0 + c < p.N) ? p.bias[(long long)s * p.b_sstride + p.b_off + n0 + c] : 0.f;
        asm volatile("bar.sync 1, 128;" ::: "memory");
        cur_s = s; cur_n0 = n0;
      }
      mbar_wait(smem_u32(&bars->acc_full[ab]), (it >> 1) & 1);
      fence_after_sync();
      const int nt = min(256, ((p.N - n0 + 31) / 32) * 32);
      for (int c0 = 0; c0 < nt; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem + lane_base + ab * 256 + c0, v);
        wait_ld();
        if (c0 + 32 >= nt) {   // the accumulator is in registers: tell the leader's MMA warp
          fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster_release(smem_u32(&bars->acc_empty[ab]), 0);
        }
        if (tma_out) {
          // bf16 output through a swizzled staging box and a TMA store per 64 columns: full-line writes, M / N tails clipped by
          // the tensor map (one 16-byte store per lane put 32 half-used sectors of 32 different rows into every instruction)
          uint8_t* stg = stage_out + (warp - 2) * 4096;
          if ((c0 & 63) == 0) {
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncwarp();
          }
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            uint32_t w[4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
              w[i] = pack_bf16x2(act_fixed<kAct>(p.act, __uint_as_float(v[q4 * 8 + 2 * i]) + s_bias[c0 + q4 * 8 + 2 * i]),
                                 act_fixed<kAct>(p.act, __uint_as_float(v[q4 * 8 + 2 * i + 1]) + s_bias[c0 + q4 * 8 + 2 * i + 1]));
            const int chunk = (((c0 & 63) >> 3) + q4) ^ (lane & 7);
            *reinterpret_cast<uint4*>(stg + lane * 128 + chunk * 16) = make_uint4(w[0], w[1], w[2], w[3]);
          }
          if ((c0 & 63) == 32 || c0 + 32 >= nt) {
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0 && m0 + gq * 32 < p.M) {
              tma_store_3d(&tmO, smem_u32(stg), n0 + (c0 & ~63), m0 + gq * 32, s);
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
          }
        } else if (gm < p.M) {
          if (p.out16) {
            __nv_bfloat16* dst = p.out16 + (long long)s * p.o16_sstride + (long long)gm * p.ld16 + n0 + c0;
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              const int c = n0 + c0 + q4 * 8;
              if (c + 8 <= p.N) {
                uint32_t w[4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
                  w[i] = pack_bf16x2(act_fixed<kAct>(p.act, __uint_as_float(v[q4 * 8 + 2 * i]) + s_bias[c0 + q4 * 8 + 2 * i]),
                                     act_fixed<kAct>(p.act, __uint_as_float(v[q4 * 8 + 2 * i + 1]) + s_bias[c0 + q4 * 8 + 2 * i + 1]));
                *reinterpret_cast<uint4*>(dst + q4 * 8) = make_uint4(w[0], w[1], w[2], w[3]);
              } else {
                for (int i = 0; i < 8; ++i)
                  if (c + i < p.N) dst[q4 * 8 + i] = __float2bfloat16_rn(act_fixed<kAct>(p.act, __uint_as_float(v[q4 * 8 + i]) + s_bias[c0 + q4 * 8 + i]));
              }
            }
          } else {
            float* dst = p.out32 + (long long)s * p.o32_sstride + (long long)gm * p.N;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int c = n0 + c0 + i;
              if (c < p.N) dst[c] = act_fixed<kAct>(p.act, __uint_as_float(v[i]) + s_bias[c0 + i]);
            }
          }
        }
      }
    }
  }
  if (warp >= 2 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  fence_before_sync();
  __syncthreads();
  cluster_sync();          // the peer's smem / barriers are in use until the leader's last MMA and commit have completed
  if (warp == 1) tmem_dealloc2<512>(tmem);
}

// =======================================================================================
// Stored posteriors (BASELINE config 3) at batch sizes above the streaming kernel's 16 rows: the SAME per-sample
// GEMM, but the weights are read by TMA straight from the fp32 `[S,P]` slab -- once, 4 B per parameter, which is
// the algorithmic traffic of that path -- and rounded to bf16 INSIDE the SM: the four epilogue warps spend the
// main loop turning each landed fp32 block into the swizzled bf16 B operand of the next MMA.  (Before: a separate
// pass wrote a bf16 copy of the slab chunk to HBM and the GEMM read it back, 2x the algorithmic bytes.)
//   warp 0     TMA producer: W block [64 k x 64 n] fp32 -> landing ring (3 x 16 KB); A block [128 x 64] bf16 -> operand ring
//   warps 2-5  converter: landing block -> bf16, MN-major SWIZZLE_128B layout (row k: 128 B, 16-byte chunk c at c ^ (k & 7))
//              -> operand ring (2 x 24 KB)
//   warp 1     MMA issuer (M = 128, N = 64, four K = 16 steps per block), two 64-column accumulators in TMEM
//   warps 6-9  epilogue of tile i (bias, activation, stores) while the main loop of tile i+1 runs
// 97 KB of shared memory per CTA -> two persistent CTAs per SM.
// =======================================================================================
// What bounds the kernel is SHARED-MEMORY bandwidth, not HBM: per fp32 byte of W the SM moved 1 B (TMA landing) + 1 B
// (converter read) + 0.5 B (converter write) + 0.5 B (MMA B read) + 1 B (TMA write of the [128 x 64] A block, re-fetched
// for every 64-column tile) + 1 B (MMA A read) = 5 B; at 5.6 TB/s and the 1.48 GHz the power cap leaves that is the
// 128 B/clk/SM limit (bench: 0.78 of the HBM peak; ncu at full clock saw 0.89).  Hence NT = 128: 128-column tiles
// (512-byte slab rows per TMA box, A traffic per W byte halved) and A boxes of 64 rows when the batch has at most 64
// (the other 64 rows of the tile are never written; their accumulator rows are never read): 3.75 B per W byte.
template <int NT> struct FCfg {
  static constexpr int LandBytes = 64 * NT * 4;      // [64 k][NT n] fp32, as TMA delivers it (no swizzle, NT*4-byte rows)
  static constexpr int WopBytes = NT * 128;          // converted [64 k x NT n] bf16: NT/64 MN-major SW128 blocks of 8 KB
  static constexpr int OpBytes = 16384 + WopBytes;   // A block 16 KB + converted W
  static constexpr int Land = NT == 128 ? 4 : 3;
  static constexpr int Op = 2;
  static constexpr int CtasPerSm = NT == 128 ? 1 : 2;
  static constexpr int Smem = Land * LandBytes + Op * OpBytes;
};
constexpr int kFLandMax = 4, kFOp = 2;
constexpr int kFThreads = 320;       // producer, issuer, 4 converter warps, 4 epilogue warps

struct __align__(8) FBarriers {
  uint64_t land_full[kFLandMax], land_empty[kFLandMax];
  uint64_t op_full[kFOp], op_empty[kFOp];
  uint64_t acc_full[2], acc_empty[2];
  uint32_t tmem_slot;
};

struct FGemmParams {
  GemmParams g;
  const int32_t* w_rows; long long w_row0;   // sample -> slab row (third TMA coordinate)
  int ns, tiles_n, tiles_m;
  int a_rep;                                 // > 0: A is the same for every sample and exists a_rep times (copy s % a_rep is read)
  int a_box_bytes;                           // bytes one A box delivers (64 or 128 rows of 128 B)
};

// PERSISTENT: every CTA walks tiles t = blockIdx.x, blockIdx.x + gridDim.x, ... of the flattened (sample, m tile, n tile)
// list (n fastest: CTAs running side by side read neighbouring segments of the same slab rows), and all three
// rings run across tile boundaries, so the HBM stream never drains between tiles.
template <int kAct, int NT>
__global__ void __launch_bounds__(kFThreads, FCfg<NT>::CtasPerSm)
dense_tc_f32w_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const FGemmParams fp) {
  using C = FCfg<NT>;
  constexpr int kFLand = C::Land;
  constexpr int kFLandBytes = C::LandBytes;
  constexpr int kFOpBytes = C::OpBytes;
  constexpr int kFNT = NT;
  const GemmParams& p = fp.g;
  uint8_t* smem = (uint8_t*)(((uintptr_t)gemm_smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t s_land = smem_u32(smem);
  uint8_t* op_smem = smem + kFLand * kFLandBytes;
  const uint32_t s_op = s_land + kFLand * kFLandBytes;
  FBarriers* bars = (FBarriers*)(op_smem + kFOp * kFOpBytes);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int kb_n = (p.K + 63) / 64;
  const long long total = (long long)fp.ns * fp.tiles_m * fp.tiles_n;
  auto decode = [&](long long t, int& n0, int& m0, int& s) {
    const int nt = (int)(t % fp.tiles_n);
    const long long r = t / fp.tiles_n;
    n0 = nt * kFNT; m0 = (int)(r % fp.tiles_m) * 128; s = (int)(r / fp.tiles_m);
  };

  if (tid == 0) {
    for (int i = 0; i < kFLand; ++i) { mbar_init(smem_u32(&bars->land_full[i]), 1); mbar_init(smem_u32(&bars->land_empty[i]), 4); }
    for (int i = 0; i < kFOp; ++i) { mbar_init(smem_u32(&bars->op_full[i]), 1 + 4); mbar_init(smem_u32(&bars->op_empty[i]), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&bars->acc_full[i]), 1); mbar_init(smem_u32(&bars->acc_empty[i]), 4); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<2 * NT>(smem_u32(&bars->tmem_slot));
  if (warp == 0 && lane == 0) { prefetch_tmap(&tmA); prefetch_tmap(&tmW); }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = bars->tmem_slot;

  if (warp == 0) {
    // ---- TMA producer.  The fp32 weight blocks (the bulk of the bytes) are requested kFLand blocks ahead of their
    // conversion, across tile boundaries; the small A blocks follow the operand ring.
    long long wt = blockIdx.x; int wkb = 0; uint32_t wg = 0;       // cursor of the next weight block to request
    int wn0 = 0, wm0 = 0, ws = 0, wrow = 0;
    auto w_setup = [&]() {
      if (wt < total) { decode(wt, wn0, wm0, ws); wrow = (int)(fp.w_rows ? (long long)fp.w_rows[ws] : (fp.w_row0 + ws)); }
    };
    auto issue_w = [&]() {
      const uint32_t l = wg % kFLand, lph = (wg / kFLand) & 1;
      mbar_wait(smem_u32(&bars->land_empty[l]), lph ^ 1);
      if (elect_one()) {
        const uint32_t fb = smem_u32(&bars->land_full[l]);
        mbar_expect_tx(fb, kFLandBytes);
        tma_load_3d(s_land + l * kFLandBytes, &tmW, fb, wn0, wkb * 64, wrow);
      }
      __syncwarp();
      ++wg;
      if (++wkb == kb_n) { wkb = 0; wt += gridDim.x; w_setup(); }
    };
    w_setup();
    for (int i = 0; i < kFLand && wt < total; ++i) issue_w();
    uint32_t g = 0;
    for (long long t = blockIdx.x; t < total; t += gridDim.x) {
      int n0, m0, s;
      decode(t, n0, m0, s);
      for (int kb = 0; kb < kb_n; ++kb, ++g) {
        const uint32_t o = g % kFOp, oph = (g / kFOp) & 1;
        mbar_wait(smem_u32(&bars->op_empty[o]), oph ^ 1);
        if (elect_one()) {
          const uint32_t fb = smem_u32(&bars->op_full[o]);
          mbar_expect_tx(fb, (uint32_t)fp.a_box_bytes);
          tma_load_3d(s_op + o * kFOpBytes, &tmA, fb, kb * 64, m0, fp.a_rep ? s % fp.a_rep : s);
        }
        __syncwarp();
        if (wt < total) issue_w();
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer
    constexpr uint32_t kDescHi128 = (1024u >> 4) | (1u << 14) | (2u << 29);
    constexpr uint32_t idesc = make_idesc_bf16(128, kFNT, 0, 1);
    uint32_t g = 0, it = 0;
    for (long long t = blockIdx.x; t < total; t += gridDim.x, ++it) {
      const uint32_t ab = it & 1;
      mbar_wait(smem_u32(&bars->acc_empty[ab]), ((it >> 1) & 1) ^ 1);
      fence_after_sync();
      for (int kb = 0; kb < kb_n; ++kb, ++g) {
        const uint32_t o = g % kFOp, oph = (g / kFOp) & 1;
        mbar_wait(smem_u32(&bars->op_full[o]), oph);
        fence_after_sync();
        if (elect_one()) {
          const uint32_t a16 = ((s_op + o * kFOpBytes) & 0x3FFFF) >> 4;
          const uint32_t a_lo = a16 | (1u << 16);
          const uint32_t b_lo = (a16 + (16384 >> 4)) | ((8192u >> 4) << 16);
          const int ksteps = min(4, (p.K - kb * 64 + 15) / 16);
          for (int kk = 0; kk < ksteps; ++kk)
            mma_ss_lh(tmem + ab * kFNT, a_lo + kk * 2, kDescHi128, b_lo + kk * 128, kDescHi128, idesc, (kb | kk) > 0);
          mma_commit(smem_u32(&bars->op_empty[o]));
          if (kb == kb_n - 1) mma_commit(smem_u32(&bars->acc_full[ab]));
        }
        __syncwarp();
      }
    }
  } else if (warp < 6) {
    // ---- converter: landed fp32 block -> bf16 B operand: 64-column blocks of 8 KB, row k = 128 B, 16-byte chunk c at c ^ (k & 7).
    // Sixteen threads per k row, each taking 4 consecutive columns of every 64-column block: a quarter-warp reads 128
    // contiguous bytes and a half-warp writes one 128-byte operand row -- no shared-memory bank conflicts (the former
    // 32-bytes-per-thread mapping had 40 % conflict wavefronts, ncu, on a kernel whose bound is shared-memory bandwidth).
    const int tc = tid - 64;                      // 0..127
    uint32_t g = 0;
    for (long long t = blockIdx.x; t < total; t += gridDim.x) {
      for (int kb = 0; kb < kb_n; ++kb, ++g) {
        const uint32_t l = g % kFLand, lph = (g / kFLand) & 1, o = g % kFOp, oph = (g / kFOp) & 1;
        mbar_wait(smem_u32(&bars->land_full[l]), lph);
        mbar_wait(smem_u32(&bars->op_empty[o]), oph ^ 1);
        const uint8_t* src = smem + l * kFLandBytes;
        uint8_t* dst = op_smem + o * kFOpBytes + 16384;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int u = i * 128 + tc;
          const int k = u >> 4, l16 = u & 15;     // row k, columns 4*l16 .. 4*l16+3 of every 64-column block
          uint8_t* drow = dst + (k >> 3) * 1024 + (k & 7) * 128 + ((((l16 >> 1) ^ (k & 7)) << 4) | ((l16 & 1) << 3));
#pragma unroll
          for (int nb = 0; nb < NT / 64; ++nb) {
            const float4 a = *reinterpret_cast<const float4*>(src + k * (NT * 4) + nb * 256 + l16 * 16);
            *reinterpret_cast<uint2*>(drow + nb * 8192) = make_uint2(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w));
          }
        }
        fence_proxy_async_smem();                 // generic-proxy stores -> visible to the tensor core
        __syncwarp();
        if (lane == 0) { mbar_arrive(smem_u32(&bars->op_full[o])); mbar_arrive(smem_u32(&bars->land_empty[l])); }
      }
    }
  } else {
    // ---- epilogue (warps 6..9 own TMEM lane quadrants 2, 3, 0, 1)
    const int gq = warp & 3;
    const int row = gq * 32 + lane;
    const uint32_t lane_base = (uint32_t)(gq * 32) << 16;
    uint32_t it = 0;
    for (long long t = blockIdx.x; t < total; t += gridDim.x, ++it) {
      int n0, m0, s;
      decode(t, n0, m0, s);
      const int gm = m0 + row;
      const uint32_t ab = it & 1;
      const long long brow = p.rows ? (long long)p.rows[s] : (p.row0 + s);
      const float* bias = p.bias + brow * p.b_sstride + p.b_off;
      mbar_wait(smem_u32(&bars->acc_full[ab]), (it >> 1) & 1);
      fence_after_sync();
#pragma unroll
      for (int hb = 0; hb < NT / 64; ++hb) {
        uint32_t v[64];
        tmem_ld32(tmem + lane_base + ab * kFNT + hb * 64, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
        tmem_ld32(tmem + lane_base + ab * kFNT + hb * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
        wait_ld();
        if (hb == NT / 64 - 1) {
          fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&bars->acc_empty[ab]));   // the accumulator is in registers: the next tile may start
        }
        const int nh = n0 + hb * 64;
        if (gm < p.M) {
          if (p.out16) {
            __nv_bfloat16* dst = p.out16 + (long long)s * p.o16_sstride + (long long)gm * p.ld16 + nh;
#pragma unroll
            for (int q8 = 0; q8 < 8; ++q8) {
              const int c = nh + q8 * 8;
              if (c + 8 <= p.N) {
                uint32_t w[4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
                  w[i] = pack_bf16x2(act_fixed<kAct>(p.act, __uint_as_float(v[q8 * 8 + 2 * i]) + bias[c + 2 * i]),
                                     act_fixed<kAct>(p.act, __uint_as_float(v[q8 * 8 + 2 * i + 1]) + bias[c + 2 * i + 1]));
                *reinterpret_cast<uint4*>(dst + q8 * 8) = make_uint4(w[0], w[1], w[2], w[3]);
              } else {
                for (int i = 0; i < 8; ++i)
                  if (c + i < p.N) dst[q8 * 8 + i] = __float2bfloat16_rn(act_fixed<kAct>(p.act, __uint_as_float(v[q8 * 8 + i]) + bias[c + i]));
              }
            }
          } else {
            float* dst = p.out32 + (long long)s * p.o32_sstride + (long long)gm * p.N;
#pragma unroll
            for (int i = 0; i < 64; ++i) {
              const int c = nh + i;
              if (c < p.N) dst[c] = act_fixed<kAct>(p.act, __uint_as_float(v[i]) + bias[c]);
            }
          }
        }
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc<2 * NT>(tmem);
}


// =======================================================================================
// Direct Conv2D (stride 1, Cin in {16, 32, 64}) on tcgen05 -- no im2col matrix.  A tile = R output rows x Wo output
// columns of one image (R * Wo <= 128 accumulator rows); for filter tap (i, j) the A operand is the box
// [Cin, Wo, R] of the NHWC input at (j - pad_l, ho0 + i - pad_t), which TMA delivers as R*Wo rows of Cin channels
// (out-of-range coordinates -- SAME padding -- arrive as zeros), and the B operand is rows [tap*Cin, (tap+1)*Cin) of
// the sample's HWIO kernel.  K loop = kh*kw taps x Cin/16 MMAs.  Persistent CTAs (tile list = sample x image x row
// group), two TMEM accumulators, epilogue (bias, activation, NHWC bf16 stores) overlapped with the next tile.
// Before: the conv input was expanded to an im2col matrix in HBM (115 MB per sample for config 4's second conv) and
// read back by the GEMM.
// =======================================================================================
#ifndef DBX_CSTAGES
#define DBX_CSTAGES 6
#endif
constexpr int kCStages = DBX_CSTAGES;
#ifndef DBX_CEPI
#define DBX_CEPI 4
#endif
constexpr int kCEpiWarps = DBX_CEPI;      // 4 or 8: with 8, two warps share a TMEM lane quarter and split the tile's columns.  Measured
                                          // SLOWER (config 4: 6.45 -> 6.77 ms): what bounds the epilogue is the number of instructions
                                          // issued per tile, and every extra warp repeats the per-tile bookkeeping
constexpr int kCThreads = 64 + 32 * kCEpiWarps;

struct ConvParams {
  int NB, Ho, Wo, Cin, Cout, kh, kw, pad_t, pad_l, R, tiles_per_img, ns, a_shared, act;
  int patch;   // one input patch [R+kh-1][Wo+kw-1][Cin] per tile serves all taps (see conv_tc_kernel); implies w_res
  int PW, base_off;
  int w_res;   // the whole kernel tensor of a sample stays in shared memory while the CTA works on that sample
  int stages;  // A-ring depth (<= kCStages; fewer when the pooled epilogue needs its staging tile)
  int pool;    // 1: the layer is followed by MaxPool 2x2 / stride 2 and the epilogue writes the POOLED map [ns][NB][Ho/2][Wo/2][Cout]
  const float* bias; long long b_sstride, b_off;
  __nv_bfloat16* out; long long o_sstride;
};

struct __align__(8) CBarriers {
  uint64_t full[kCStages], empty[kCStages];
  uint64_t acc_full[2], acc_empty[2];
  uint64_t w_full;
  uint32_t tmem_slot;
};

template <int kAct>
__global__ void __launch_bounds__(kCThreads, 2)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const ConvParams p) {
  __shared__ __align__(16) float s_bias[128];         // the current sample's bias (epilogue warps refill it per sample)
  uint8_t* smem = (uint8_t*)(((uintptr_t)gemm_smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t s_stage = smem_u32(smem);
  const int cin_b = p.Cin * 2;                        // bytes per A row = swizzle span (32 / 64 / 128)
  // patch mode: the slot holds (kh-1)*PW + kw-1 + 128 pixel rows (the last ones are only read for garbage accumulator rows)
  const int a_bytes = p.patch ? (((p.kh - 1) * p.PW + p.kw - 1 + 128) * cin_b + 1023) / 1024 * 1024 : 128 * cin_b;
  const int nblk = (p.Cout + 63) / 64;                // 64-column weight blocks, each Cin rows of 128 B
  const int w_tap_b = nblk * p.Cin * 128;             // one tap's weight rows
  const int taps = p.kh * p.kw;
  // w_res: [all taps' weights][A stages]; otherwise every stage carries its tap's weights behind the A block.  (With
  // per-tap weight loads every CTA of the GPU re-read the same few KB of the current sample's kernel for every tile.)
  const int w_res_b = p.w_res ? taps * w_tap_b : 0;
  const int stage_b = a_bytes + (p.w_res ? 0 : w_tap_b);
  const uint32_t s_wres = s_stage;
  const uint32_t s_ring = s_stage + w_res_b;
  CBarriers* bars = (CBarriers*)(smem + w_res_b + p.stages * stage_b);
  uint8_t* pool_stage = (uint8_t*)bars + ((sizeof(CBarriers) + 15) / 16) * 16;      // pool mode: [128 accumulator rows][Cout] bf16
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long total = (long long)p.ns * p.NB * p.tiles_per_img;
  const int NT = nblk * 64;
  // a CONTIGUOUS range of the (sample, image, row group) list per CTA: it changes sample (= reloads the resident kernel
  // tensor, draining its pipeline) once or twice per launch, and walks down its images row group by row group, so the
  // input rows its nine taps share stay in L2
  const long long t_begin = (total * blockIdx.x) / gridDim.x, t_end = (total * (blockIdx.x + 1)) / gridDim.x;

  if (tid == 0) {
    for (int i = 0; i < kCStages; ++i) { mbar_init(smem_u32(&bars->full[i]), 1); mbar_init(smem_u32(&bars->empty[i]), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&bars->acc_full[i]), 1); mbar_init(smem_u32(&bars->acc_empty[i]), kCEpiWarps); }
    mbar_init(smem_u32(&bars->w_full), 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<256>(smem_u32(&bars->tmem_slot));
  if (warp == 0 && lane == 0) { prefetch_tmap(&tmA); prefetch_tmap(&tmW); }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = bars->tmem_slot;

  // (row group, image, sample) of a tile: decoded ONCE per role and then advanced (three runtime divisions per tile and thread
  // were a tenth of the epilogue's instructions, and the epilogue's instruction count is what bounds this kernel)
  struct TileIt { int rg, nb, s; };
  auto decode0 = [&](long long t) {
    TileIt ti;
    ti.rg = (int)(t % p.tiles_per_img);
    const long long r = t / p.tiles_per_img;
    ti.nb = (int)(r % p.NB); ti.s = (int)(r / p.NB);
    return ti;
  };
  auto advance = [&](TileIt& ti) {
    if (++ti.rg == p.tiles_per_img) { ti.rg = 0; if (++ti.nb == p.NB) { ti.nb = 0; ++ti.s; } }
  };

  if (warp == 0) {
    uint32_t g = 0;
    const uint32_t tx = (uint32_t)(p.R * p.Wo * cin_b + (p.w_res ? 0 : w_tap_b));
    int cur_s = -1;
    TileIt ti = decode0(t_begin);
    for (long long t = t_begin; t < t_end; ++t, advance(ti)) {
      const int ho0 = ti.rg * p.R, nb = ti.nb, s = ti.s;
      if (p.w_res && s != cur_s) {
        // new sample: every MMA that read the old weights has retired once the last tap's commit has arrived
        if (g > 0) mbar_wait(smem_u32(&bars->empty[(g - 1) % p.stages]), ((g - 1) / p.stages) & 1);
        if (elect_one()) {
          const uint32_t wf = smem_u32(&bars->w_full);
          mbar_expect_tx(wf, (uint32_t)w_res_b);
          for (int tap = 0; tap < taps; ++tap)
            for (int b = 0; b < nblk; ++b) tma_load_3d(s_wres + tap * w_tap_b + b * p.Cin * 128, &tmW, wf, b * 64, tap * p.Cin, s);
        }
        __syncwarp();
        cur_s = s;
      }
      if (p.patch) {
        // ONE box per tile: input rows ho0-pad_t .. +R+kh-2, columns -pad_l .. +PW-1 (zeros outside the image); the taps
        // are address offsets into it.  (Per-tap boxes cost 9 x 112 TMA rows of 64 B per tile, and the TMA row rate, not
        // bandwidth or latency, bounded the kernel.)
        const uint32_t st = g % p.stages, ph = (g / p.stages) & 1;
        mbar_wait(smem_u32(&bars->empty[st]), ph ^ 1);
        if (elect_one()) {
          const uint32_t fb = smem_u32(&bars->full[st]);
          mbar_expect_tx(fb, (uint32_t)((p.R + p.kh - 1) * p.PW * cin_b));
          tma_load_5d(s_ring + st * stage_b, &tmA, fb, 0, -p.pad_l, ho0 - p.pad_t, nb, p.a_shared ? 0 : s);
        }
        __syncwarp();
        ++g;
        continue;
      }
      for (int tap = 0; tap < taps; ++tap, ++g) {
        const uint32_t st = g % p.stages, ph = (g / p.stages) & 1;
        mbar_wait(smem_u32(&bars->empty[st]), ph ^ 1);
        if (elect_one()) {
          const uint32_t fb = smem_u32(&bars->full[st]);
          const uint32_t base = s_ring + st * stage_b;
          const int i = tap / p.kw, j = tap - i * p.kw;
          mbar_expect_tx(fb, tx);
          tma_load_5d(base, &tmA, fb, 0, j - p.pad_l, ho0 + i - p.pad_t, nb, p.a_shared ? 0 : s);
          if (!p.w_res)
            for (int b = 0; b < nblk; ++b) tma_load_3d(base + a_bytes + b * p.Cin * 128, &tmW, fb, b * 64, tap * p.Cin, s);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // A: K-major rows of cin_b bytes, 8-row groups cin_b*8 apart, swizzle span = cin_b.  B: MN-major SWIZZLE_128B.
    const uint32_t swz = cin_b == 128 ? 2u : (cin_b == 64 ? 4u : 6u);
    const uint32_t a_hi = ((uint32_t)(cin_b * 8) >> 4) | (1u << 14) | (swz << 29);
    constexpr uint32_t b_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    const uint32_t idesc = make_idesc_bf16(128, NT, 0, 1);
    const int ksteps = p.Cin / 16;
    uint32_t g = 0, it = 0, w_n = 0;
    int cur_s = -1;
    TileIt ti = decode0(t_begin);
    for (long long t = t_begin; t < t_end; ++t, ++it, advance(ti)) {
      const uint32_t ab = it & 1;
      if (p.w_res) {
        const int s = ti.s;
        if (s != cur_s) { mbar_wait(smem_u32(&bars->w_full), w_n & 1); ++w_n; cur_s = s; }
      }
      mbar_wait(smem_u32(&bars->acc_empty[ab]), ((it >> 1) & 1) ^ 1);
      fence_after_sync();
      if (p.patch) {
        const uint32_t st = g % p.stages, ph = (g / p.stages) & 1;
        mbar_wait(smem_u32(&bars->full[st]), ph);
        if (elect_one()) {
          const uint32_t base = (s_ring + st * stage_b) & 0x3FFFF;
          for (int tap = 0; tap < taps; ++tap) {
            const int i = tap / p.kw, j = tap - i * p.kw;
            const uint32_t addr = base + (uint32_t)((i * p.PW + j) * cin_b);     // the patch shifted by (i, j) pixels
            // the swizzle XOR is a function of the absolute shared-memory address (what TMA wrote with), so a start address
            // shifted by whole pixel rows needs NO base-offset field -- measured: setting it gives wrong results
            const uint32_t hi = a_hi | (p.base_off ? (((addr >> 7) & 7u) << 17) : 0u);
            const uint32_t a_lo = (addr >> 4) | (1u << 16);
            const uint32_t b_lo = (((s_wres + tap * w_tap_b) & 0x3FFFF) >> 4) | ((uint32_t)((p.Cin * 128) >> 4) << 16);
            for (int kk = 0; kk < ksteps; ++kk)
              mma_ss_lh(tmem + ab * 128, a_lo + kk * 2, hi, b_lo + kk * 128, b_hi, idesc, (tap | kk) > 0);
          }
          mma_commit(smem_u32(&bars->empty[st]));
          mma_commit(smem_u32(&bars->acc_full[ab]));
        }
        __syncwarp();
        ++g;
        continue;
      }
      for (int tap = 0; tap < taps; ++tap, ++g) {
        const uint32_t st = g % p.stages, ph = (g / p.stages) & 1;
        mbar_wait(smem_u32(&bars->full[st]), ph);
        if (elect_one()) {
          const uint32_t a16 = ((s_ring + st * stage_b) & 0x3FFFF) >> 4;
          const uint32_t a_lo = a16 | (1u << 16);
          const uint32_t w16 = p.w_res ? (((s_wres + tap * w_tap_b) & 0x3FFFF) >> 4) : (a16 + (uint32_t)(a_bytes >> 4));
          const uint32_t b_lo = w16 | ((uint32_t)((p.Cin * 128) >> 4) << 16);
          for (int kk = 0; kk < ksteps; ++kk)
            mma_ss_lh(tmem + ab * 128, a_lo + kk * 2, a_hi, b_lo + kk * 128, b_hi, idesc, (tap | kk) > 0);
          mma_commit(smem_u32(&bars->empty[st]));
          if (tap == taps - 1) mma_commit(smem_u32(&bars->acc_full[ab]));
        }
        __syncwarp();
      }
    }
  } else {
    const int gq = warp & 3;
    const int ch = (warp - 2) >> 2;                  // column half of this warp (always 0 with four epilogue warps)
    const int row = gq * 32 + lane;                  // accumulator row = output pixel r * Wo + wo of the tile
    const uint32_t lane_base = (uint32_t)(gq * 32) << 16;
    const int rw = p.patch ? p.PW : p.Wo;          // accumulator rows per output row (patch mode keeps kw-1 garbage columns)
    const int r = row / rw, wo = row - r * rw;
    uint32_t it = 0;
    int cur_s = -1;
    const float* bias = s_bias;
    // pool mode: what a thread does in the pooling pass is the same for every tile -- its (up to two) tasks' four staged
    // addresses and output offsets are computed here, once (per tile they cost ~150 instructions of divisions and address
    // arithmetic per task); tasks beyond 256 per tile take the general loop
    const int nchunk_p = p.Cout >> 3, pw_n = p.Wo >> 1, PHo = p.Ho >> 1;
    const int tasks = p.pool ? (p.R >> 1) * pw_n * nchunk_p : 0;
    int pt_off[2][4], pt_out[2], pt_pr[2];
    bool pt_ok[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int j = tid - 64 + 32 * kCEpiWarps * k;
      pt_ok[k] = j < tasks;
      const int q = pt_ok[k] ? j % nchunk_p : 0, pp = pt_ok[k] ? j / nchunk_p : 0;
      const int pr = pp / max(pw_n, 1), pc = pp - pr * pw_n;
      const int a0 = (2 * pr) * rw + 2 * pc;
      const int rws[4] = {a0, a0 + 1, a0 + rw, a0 + rw + 1};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int sz = p.pool ? ((rws[e] / (8 / nchunk_p)) & (nchunk_p - 1)) : 0;
        pt_off[k][e] = rws[e] * (p.Cout * 2) + ((q ^ sz) << 4);
      }
      pt_out[k] = (pr * pw_n + pc) * p.Cout + q * 8;
      pt_pr[k] = pr;
    }
    TileIt ti = decode0(t_begin);
    for (long long t = t_begin; t < t_end; ++t, ++it, advance(ti)) {
      const int ho0 = ti.rg * p.R, nb = ti.nb, s = ti.s;
      const uint32_t ab = it & 1;
      if (s != cur_s) {      // per-column global loads in the epilogue were its longest stall (ncu: long scoreboard)
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kCEpiWarps) : "memory");
        const int te = tid - 64;
        if (te < p.Cout) s_bias[te] = p.bias[(long long)s * p.b_sstride + p.b_off + te];
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kCEpiWarps) : "memory");
        cur_s = s;
      }
      const bool valid = r < p.R && wo < p.Wo && ho0 + r < p.Ho;
      __nv_bfloat16* dst = p.out + (long long)s * p.o_sstride + ((((long long)nb * p.Ho + ho0 + r) * p.Wo + wo) * p.Cout);
      // pool mode: the activated bf16 row goes to the staging tile instead (16-byte chunks XOR-swizzled so that the 32 row-wise
      // writes of a warp spread over all banks: rows of Cout * 2 <= 128 bytes)
      const int nchunk = p.Cout >> 3;                                   // 16-byte chunks per row (1, 2, 4 or 8 in pool mode)
      const int swz = p.pool ? ((row / (8 / nchunk)) & (nchunk - 1)) : 0;
      uint8_t* srow = pool_stage + row * (p.Cout * 2);
      mbar_wait(smem_u32(&bars->acc_full[ab]), (it >> 1) & 1);
      fence_after_sync();
      const int c_per = NT / (kCEpiWarps / 4), c_lo = ch * c_per, c_hi = c_lo + c_per;      // NT is a multiple of 64
      for (int c0 = c_lo; c0 < c_hi; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem + lane_base + ab * 128 + c0, v);
        wait_ld();
        if (c0 + 32 >= c_hi) {   // everything is in registers: hand the accumulator back before the stores
          fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&bars->acc_empty[ab]));
        }
        if (valid || p.pool) {
#pragma unroll
          for (int q8 = 0; q8 < 4; ++q8) {
            const int c = c0 + q8 * 8;
            if (c + 8 <= p.Cout) {
              uint32_t w[4];
              const float4 bl = *reinterpret_cast<const float4*>(bias + c), bh = *reinterpret_cast<const float4*>(bias + c + 4);
              const float bb[8] = {bl.x, bl.y, bl.z, bl.w, bh.x, bh.y, bh.z, bh.w};
#pragma unroll
              for (int i = 0; i < 4; ++i)
                w[i] = pack_bf16x2(act_fixed<kAct>(p.act, __uint_as_float(v[q8 * 8 + 2 * i]) + bb[2 * i]),
                                   act_fixed<kAct>(p.act, __uint_as_float(v[q8 * 8 + 2 * i + 1]) + bb[2 * i + 1]));
              if (p.pool) *reinterpret_cast<uint4*>(srow + (((c >> 3) ^ swz) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
              else *reinterpret_cast<uint4*>(dst + c) = make_uint4(w[0], w[1], w[2], w[3]);
            } else if (!p.pool) {
              for (int i = 0; i < 8; ++i)
                if (c + i < p.Cout) dst[c + i] = __float2bfloat16_rn(act_fixed<kAct>(p.act, __uint_as_float(v[q8 * 8 + i]) + bias[c + i]));
            }
          }
        }
      }
      if (p.pool) {
        // 2x2 / stride-2 max over the staged tile: R (even) output rows x Wo columns -> R/2 x Wo/2 pooled pixels, one 16-byte
        // channel chunk per task, written as full NHWC rows of the pooled map (packed bf16x2 max, like maxpool_bf16_vec8_kernel)
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kCEpiWarps) : "memory");
        const int pitch = p.Cout * 2;
        __nv_bfloat16* obase = p.out + (long long)s * p.o_sstride + (((long long)nb * PHo + (ho0 >> 1)) * pw_n) * p.Cout;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          if (!pt_ok[k] || (ho0 >> 1) + pt_pr[k] >= PHo) continue;
          __nv_bfloat162 acc[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const uint4 t4 = *reinterpret_cast<const uint4*>(pool_stage + pt_off[k][e]);
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t4);
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[i] = e == 0 ? h[i] : __hmax2(acc[i], h[i]);
          }
          *reinterpret_cast<uint4*>(obase + pt_out[k]) = *reinterpret_cast<const uint4*>(acc);
        }
        for (int j = tid - 64 + 2 * 32 * kCEpiWarps; j < tasks; j += 32 * kCEpiWarps) {
          const int q = j % nchunk, pp = j / nchunk;
          const int pr = pp / pw_n, pc = pp - pr * pw_n;
          const int prow = (ho0 >> 1) + pr;
          if (prow >= PHo) continue;
          const int a0 = (2 * pr) * rw + 2 * pc;
          uint4 m4;
          {
            const int rws[4] = {a0, a0 + 1, a0 + rw, a0 + rw + 1};
            __nv_bfloat162 acc[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int rr = rws[e];
              const int sz = (rr / (8 / nchunk)) & (nchunk - 1);
              const uint4 t4 = *reinterpret_cast<const uint4*>(pool_stage + rr * pitch + ((q ^ sz) << 4));
              const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t4);
#pragma unroll
              for (int i = 0; i < 4; ++i) acc[i] = e == 0 ? h[i] : __hmax2(acc[i], h[i]);
            }
            m4 = *reinterpret_cast<const uint4*>(acc);
          }
          __nv_bfloat16* po = p.out + (long long)s * p.o_sstride + ((((long long)nb * PHo + prow) * pw_n + pc) * p.Cout) + q * 8;
          *reinterpret_cast<uint4*>(po) = m4;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kCEpiWarps) : "memory");       // the staging tile is free for the next tile
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc<256>(tmem);
}


// =======================================================================================
// One hidden Dense layer of the interval forward (ibp.cuh) in ONE kernel: centre MU = A_mu W + b and radius
// R = A_r |W| accumulate side by side in TMEM (2 x 128 columns per buffer, two buffers), |W| is formed INSIDE the SM
// from the W block TMA delivered (sign bits cleared, same swizzled layout) by four converter warps, and the epilogue
// combines them (u = act(MU + R), l = act(MU - R) -> next centre (u + l)/2 and radius (u - l)/2, bf16).  Before: a
// pass that wrote |W16| for the whole chunk, two GEMM launches writing fp32 MU / R, and a combine pass.
//   warp 0 producer (A_mu, A_r, W blocks), warp 1 MMA issuer (8 MMAs per k-block), warps 2-5 converter, warps 6.. epilogue
// Persistent, one CTA per SM (TMA ring 4 x 48 KB + |W| ring 2 x 16 KB), tiles = (sample, m tile, 128-column tile).
// =======================================================================================
constexpr int kIStages = 3;          // TMA ring (4 measured no faster; the space is the epilogue's staging)
constexpr int kIStageB = 49152;      // A_mu 16 KB | A_r 16 KB | W [64 k x 128 n] 16 KB
constexpr int kIAbs = 2;             // |W| ring (16 KB slots): written by the converter warps, separate so that the TMA ring is 4 deep
constexpr int kIAbsB = 16384;
constexpr int kIOutB = 4 * 8192;     // per epilogue warp: [32 rows x 64 columns] bf16 boxes of the next centre and radius, for the TMA stores
constexpr int kINT = 128;
constexpr int kIEpiWarps = 4;      // one per TMEM lane quarter
constexpr int kIThreads = 192 + 32 * kIEpiWarps;

struct __align__(8) IBarriers {
  uint64_t full[kIStages], empty[kIStages], conv[kIAbs], abs_empty[kIAbs];
  uint64_t acc_full[2], acc_empty[2];
  uint32_t tmem_slot;
};

struct IbpGemmParams {
  int M, N, K, ns, tiles_m, tiles_n, a_shared, act;
  const float* bias; long long b_sstride, b_off;
  __nv_bfloat16* n_mu; __nv_bfloat16* n_r; long long n_sstride;
};

template <int kAct>
__global__ void __launch_bounds__(kIThreads, 1)
ibp_tc_kernel(const __grid_constant__ CUtensorMap tmAmu, const __grid_constant__ CUtensorMap tmAr, const __grid_constant__ CUtensorMap tmW,
              const __grid_constant__ CUtensorMap tmOmu, const __grid_constant__ CUtensorMap tmOr, const IbpGemmParams p) {
  uint8_t* smem = (uint8_t*)(((uintptr_t)gemm_smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t s_stage = smem_u32(smem);
  uint8_t* abs_ring = smem + kIStages * kIStageB;
  uint8_t* stage_out = abs_ring + kIAbs * kIAbsB;
  float* s_bias = (float*)(stage_out + kIOutB);
  IBarriers* bars = (IBarriers*)(s_bias + kINT);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int kb_n = (p.K + 63) / 64;
  const long long total = (long long)p.ns * p.tiles_m * p.tiles_n;
  const long long t_begin = (total * blockIdx.x) / gridDim.x, t_end = (total * (blockIdx.x + 1)) / gridDim.x;
  auto decode = [&](long long t, int& n0, int& m0, int& s) {
    const int nt_i = (int)(t % p.tiles_n);
    const long long r = t / p.tiles_n;
    n0 = nt_i * kINT; m0 = (int)(r % p.tiles_m) * 128; s = (int)(r / p.tiles_m);
  };

  if (tid == 0) {
    for (int i = 0; i < kIStages; ++i) { mbar_init(smem_u32(&bars->full[i]), 1); mbar_init(smem_u32(&bars->empty[i]), 1); }
    for (int i = 0; i < kIAbs; ++i) { mbar_init(smem_u32(&bars->conv[i]), 4); mbar_init(smem_u32(&bars->abs_empty[i]), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&bars->acc_full[i]), 1); mbar_init(smem_u32(&bars->acc_empty[i]), kIEpiWarps); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(smem_u32(&bars->tmem_slot));
  if (warp == 0 && lane == 0) { prefetch_tmap(&tmAmu); prefetch_tmap(&tmAr); prefetch_tmap(&tmW); prefetch_tmap(&tmOmu); prefetch_tmap(&tmOr); }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = bars->tmem_slot;

  if (warp == 0) {
    uint32_t g = 0;
    for (long long t = t_begin; t < t_end; ++t) {
      int n0, m0, s;
      decode(t, n0, m0, s);
      for (int kb = 0; kb < kb_n; ++kb, ++g) {
        const uint32_t st = g % kIStages, ph = (g / kIStages) & 1;
        mbar_wait(smem_u32(&bars->empty[st]), ph ^ 1);
        if (elect_one()) {
          const uint32_t fb = smem_u32(&bars->full[st]);
          const uint32_t base = s_stage + st * kIStageB;
          mbar_expect_tx(fb, 3 * 16384);
          if (p.a_shared) { tma_load_2d(base, &tmAmu, fb, kb * 64, m0); tma_load_2d(base + 16384, &tmAr, fb, kb * 64, m0); }
          else { tma_load_3d(base, &tmAmu, fb, kb * 64, m0, s); tma_load_3d(base + 16384, &tmAr, fb, kb * 64, m0, s); }
          for (int nb = 0; nb < 2; ++nb) tma_load_3d(base + 32768 + nb * 8192, &tmW, fb, n0 + nb * 64, kb * 64, s);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t kDescHi128 = (1024u >> 4) | (1u << 14) | (2u << 29);
    constexpr uint32_t idesc = make_idesc_bf16(128, kINT, 0, 1);
    uint32_t g = 0, it = 0;
    for (long long t = t_begin; t < t_end; ++t, ++it) {
      const uint32_t ab = it & 1;
      mbar_wait(smem_u32(&bars->acc_empty[ab]), ((it >> 1) & 1) ^ 1);
      fence_after_sync();
      for (int kb = 0; kb < kb_n; ++kb, ++g) {
        const uint32_t st = g % kIStages, ph = (g / kIStages) & 1;
        const uint32_t aj = g % kIAbs, aph = (g / kIAbs) & 1;
        mbar_wait(smem_u32(&bars->full[st]), ph);
        fence_after_sync();
        const uint32_t a16 = ((s_stage + st * kIStageB) & 0x3FFFF) >> 4;
        const uint32_t amu = a16 | (1u << 16), ar = (a16 + (16384 >> 4)) | (1u << 16);
        const uint32_t bw = (a16 + (32768 >> 4)) | ((8192u >> 4) << 16);
        const uint32_t babs = (((s_stage + kIStages * kIStageB + aj * kIAbsB) & 0x3FFFF) >> 4) | ((8192u >> 4) << 16);
        const int ksteps = min(4, (p.K - kb * 64 + 15) / 16);
        // the centre product needs only what TMA delivered: issue it first, the |W| conversion runs meanwhile
        if (elect_one())
          for (int kk = 0; kk < ksteps; ++kk)
            mma_ss_lh(tmem + ab * 256, amu + kk * 2, kDescHi128, bw + kk * 128, kDescHi128, idesc, (kb | kk) > 0);
        __syncwarp();
        mbar_wait(smem_u32(&bars->conv[aj]), aph);
        fence_after_sync();
        if (elect_one()) {
          for (int kk = 0; kk < ksteps; ++kk)
            mma_ss_lh(tmem + ab * 256 + kINT, ar + kk * 2, kDescHi128, babs + kk * 128, kDescHi128, idesc, (kb | kk) > 0);
          mma_commit(smem_u32(&bars->empty[st]));
          mma_commit(smem_u32(&bars->abs_empty[aj]));
          if (kb == kb_n - 1) mma_commit(smem_u32(&bars->acc_full[ab]));
        }
        __syncwarp();
      }
    }
  } else if (warp < 6) {
    // ---- converter: |W| block = W block with the sign bits cleared (identical layout)
    const int tc = tid - 64;
    uint32_t g = 0;
    for (long long t = t_begin; t < t_end; ++t) {
      for (int kb = 0; kb < kb_n; ++kb, ++g) {
        const uint32_t st = g % kIStages, ph = (g / kIStages) & 1;
        const uint32_t aj = g % kIAbs, aph = (g / kIAbs) & 1;
        mbar_wait(smem_u32(&bars->abs_empty[aj]), aph ^ 1);   // the radius MMAs that read this |W| slot have completed
        mbar_wait(smem_u32(&bars->full[st]), ph);
        const uint4* src = reinterpret_cast<const uint4*>(smem + st * kIStageB + 32768);
        uint4* dst = reinterpret_cast<uint4*>(abs_ring + aj * kIAbsB);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          uint4 v = src[i * 128 + tc];
          v.x &= 0x7FFF7FFFu; v.y &= 0x7FFF7FFFu; v.z &= 0x7FFF7FFFu; v.w &= 0x7FFF7FFFu;
          dst[i * 128 + tc] = v;
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bars->conv[aj]));
      }
    }
  } else {
    // ---- epilogue
    const int gq = warp & 3;      // TMEM lane quarter
    const uint32_t lane_base = (uint32_t)(gq * 32) << 16;
    uint32_t it = 0;
    int cur_s = -1, cur_n0 = -1;
    for (long long t = t_begin; t < t_end; ++t, ++it) {
      int n0, m0, s;
      decode(t, n0, m0, s);
      const uint32_t ab = it & 1;
      if (s != cur_s || n0 != cur_n0) {
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kIEpiWarps) : "memory");
        for (int te = tid - 192; te < kINT; te += 32 * kIEpiWarps) s_bias[te] = (n0 + te < p.N) ? p.bias[(long long)s * p.b_sstride + p.b_off + n0 + te] : 0.f;
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kIEpiWarps) : "memory");
        cur_s = s; cur_n0 = n0;
      }
      mbar_wait(smem_u32(&bars->acc_full[ab]), (it >> 1) & 1);
      fence_after_sync();
      // 64 columns at a time: the warp's 32 rows x 64 columns of next centre / radius go to its staging boxes in shared memory
      // (128-byte rows, 16-byte chunks XOR-swizzled by row so that the 32 row-wise writes of a warp spread over all banks),
      // and one lane hands both boxes to TMA: full-line writes, M / N tails clipped by the tensor map.  (One 16-byte store per
      // lane straight to global put 32 half-sectors of 32 different rows into every store instruction: 27 % of the kernel.)
      uint8_t* stg_mu = stage_out + (warp - 6) * 8192;
      uint8_t* stg_r = stg_mu + 4096;
      for (int h0 = 0; h0 < kINT; h0 += 64) {
        for (int c0 = h0; c0 < h0 + 64; c0 += 32) {
          uint32_t vm[32], vr[32];
          tmem_ld32(tmem + lane_base + ab * 256 + c0, vm);
          tmem_ld32(tmem + lane_base + ab * 256 + kINT + c0, vr);
          wait_ld();
          if (c0 + 32 >= kINT) {
            fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bars->acc_empty[ab]));
          }
          if (c0 == h0) {      // the previous boxes have been read by TMA
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncwarp();
          }
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            uint32_t wm[4], wr[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float cm[2], cr[2];
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const float mu = __uint_as_float(vm[q4 * 8 + 2 * i + h]) + s_bias[c0 + q4 * 8 + 2 * i + h];
                const float rr = __uint_as_float(vr[q4 * 8 + 2 * i + h]);
                const float au = act_fixed<kAct>(p.act, mu + rr), al = act_fixed<kAct>(p.act, mu - rr);
                cm[h] = (au + al) / 2.f; cr[h] = (au - al) / 2.f;
              }
              wm[i] = pack_bf16x2(cm[0], cm[1]); wr[i] = pack_bf16x2(cr[0], cr[1]);
            }
            const int chunk = (((c0 - h0) >> 3) + q4) ^ (lane & 7);
            *reinterpret_cast<uint4*>(stg_mu + lane * 128 + chunk * 16) = make_uint4(wm[0], wm[1], wm[2], wm[3]);
            *reinterpret_cast<uint4*>(stg_r + lane * 128 + chunk * 16) = make_uint4(wr[0], wr[1], wr[2], wr[3]);
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0 && n0 + h0 < p.N && m0 + gq * 32 < p.M) {
          tma_store_3d(&tmOmu, smem_u32(stg_mu), n0 + h0, m0 + gq * 32, s);
          tma_store_3d(&tmOr, smem_u32(stg_r), n0 + h0, m0 + gq * 32, s);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

// NHWC bf16 activations -> im2col rows [n*Ho*Wo][kh*kw*C (padded to Kp)] so that Conv2D becomes the GEMM above
// (same (kh, kw, cin) ordering as the HWIO kernel flattened to [kh*kw*cin, cout]).
__global__ void im2col_bf16_kernel(const __nv_bfloat16* __restrict__ in, long long in_sstride, __nv_bfloat16* __restrict__ out,
                                   long long out_sstride, int NB, int H, int Wd, int C, int Ho, int Wo, int kh, int kw,
                                   int sh, int sw, int pt, int pl, int Kp) {
  const int s = blockIdx.y;
  const __nv_bfloat16* ip = in + (long long)s * in_sstride;
  __nv_bfloat16* op = out + (long long)s * out_sstride;
  const long long total = (long long)NB * Ho * Wo * Kp;
  const int K = kh * kw * C;
  for (long long o = blockIdx.x * (long long)blockDim.x + threadIdx.x; o < total; o += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(o % Kp);
    long long r = o / Kp;
    __nv_bfloat16 v = __float2bfloat16_rn(0.f);
    if (k < K) {
      const int c = k % C; const int t = k / C; const int j = t % kw, i = t / kw;
      const int wo = (int)(r % Wo); long long r2 = r / Wo;
      const int ho = (int)(r2 % Ho); const int nb = (int)(r2 / Ho);
      const int hi = ho * sh + i - pt, wi = wo * sw + j - pl;
      if (hi >= 0 && hi < H && wi >= 0 && wi < Wd) v = ip[(((long long)nb * H + hi) * Wd + wi) * C + c];
    }
    op[o] = v;
  }
}

// Same gather for C % 8 == 0 (and Kp == kh*kw*C): one thread moves 8 channels = 16 bytes.
__global__ void im2col_bf16_vec8_kernel(const __nv_bfloat16* __restrict__ in, long long in_sstride,
                                        __nv_bfloat16* __restrict__ out, long long out_sstride, int NB, int H, int Wd, int C,
                                        int Ho, int Wo, int kh, int kw, int sh, int sw, int pt, int pl) {
  const int s = blockIdx.y;
  const uint4* ip = reinterpret_cast<const uint4*>(in + (long long)s * in_sstride);
  uint4* op = reinterpret_cast<uint4*>(out + (long long)s * out_sstride);
  const int C8 = C >> 3, taps = kh * kw;
  const int K8 = taps * C8;                               // 16-byte groups per im2col row
  const long long total = (long long)NB * Ho * Wo * K8;
  for (long long o = blockIdx.x * (long long)blockDim.x + threadIdx.x; o < total; o += (long long)gridDim.x * blockDim.x) {
    const int k8 = (int)(o % K8);
    const long long r = o / K8;
    const int c8 = k8 % C8, t = k8 / C8;
    const int j = t % kw, i = t / kw;
    const int wo = (int)(r % Wo); const long long r2 = r / Wo;
    const int ho = (int)(r2 % Ho); const int nb = (int)(r2 / Ho);
    const int hi = ho * sh + i - pt, wi = wo * sw + j - pl;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (hi >= 0 && hi < H && wi >= 0 && wi < Wd) v = __ldg(ip + (((long long)nb * H + hi) * Wd + wi) * C8 + c8);
    op[o] = v;
  }
}

static bool g_tc_ok = true;

// Launch one layer.  A: bf16 [ns or 1][M][lda] (lda % 8 == 0).  W16 chunk layout as produced by sample_bf16_kernel.
int dense_tc_launch(const __nv_bfloat16* A, long long a_sstride, int lda, int a_shared, const __nv_bfloat16* Wbase,
                    long long w_sstride_elems, int ldw, int ns, int M, int N, int K, const float* bias, long long b_sstride,
                    long long b_off, int act, __nv_bfloat16* out16, long long o16_sstride, int ld16, float* out32,
                    long long o32_sstride, cudaStream_t st, int nt_cap) {
  if (!g_tc_ok || !get_encode_fn()) return 1;
  CUtensorMap tmA, tmW;
  if (a_shared) {
    uint64_t d[2] = {(uint64_t)K, (uint64_t)M}, sb[1] = {(uint64_t)lda * 2};
    uint32_t b[2] = {64, 128};
    if (!make_tmap_bf16(&tmA, (void*)A, 2, d, sb, b, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  } else {
    uint64_t d[3] = {(uint64_t)K, (uint64_t)M, (uint64_t)ns}, sb[2] = {(uint64_t)lda * 2, (uint64_t)a_sstride * 2};
    uint32_t b[3] = {64, 128, 1};
    if (!make_tmap_bf16(&tmA, (void*)A, 3, d, sb, b, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  }
  {
    uint64_t d[3] = {(uint64_t)N, (uint64_t)K, (uint64_t)ns}, sb[2] = {(uint64_t)ldw * 2, (uint64_t)w_sstride_elems * 2};
    uint32_t b[3] = {64, 64, 1};
    if (!make_tmap_bf16(&tmW, (void*)Wbase, 3, d, sb, b, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  }
  GemmParams p = {};
  p.M = M; p.N = N; p.K = K; p.NT = (int)std::min<int64_t>(nt_cap >= 128 ? std::min(nt_cap, 256) : 256, round_up(N, 64));
  p.a_shared = a_shared; p.bias = bias; p.b_sstride = b_sstride; p.b_off = b_off; p.rows = nullptr; p.row0 = 0; p.act = act;
  p.out16 = out16; p.o16_sstride = o16_sstride; p.ld16 = ld16; p.out32 = out32; p.o32_sstride = o32_sstride;
  const int ka = act_class(act);
  // enough tiles to keep persistent CTAs busy: dense_tc_persist_kernel (DBX_NO_PERSIST=1 disables; DBX_PERSIST_MIN_TILES
  // lowers the threshold so that tests reach it with small shapes)
  {
    // big GEMMs: CTA pairs on 256 x 256 tiles (dense_tc_pair_kernel; DBX_NO_PAIR_GEMM=1 disables)
    if (M >= 256 && N >= 256 && K >= 256 && !opt().no_pair_gemm) {
      static int num_sms2 = 0;
      if (!num_sms2) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&num_sms2, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 1;
      }
      const int ptm = (int)cdiv(M, 256), ptn = (int)cdiv(N, 256);
      const long long ptiles = (long long)ns * ptm * ptn;
      if (ptiles >= num_sms2 / 2) {
        const int smem = kPrStages * kPrStageB + 4 * 4096 + (int)sizeof(PairBarriers) + 1024;
        // bf16 output: TMA stores from a staging box (needs 16-byte pitches; otherwise the kernel stores directly)
        CUtensorMap tmO = tmW;
        int tma_out = 0;
        if (out16 && !opt().no_tma_store && (ld16 & 7) == 0 && (o16_sstride & 7) == 0 &&
            ((uintptr_t)out16 & 15) == 0) {
          uint64_t d[3] = {(uint64_t)N, (uint64_t)M, (uint64_t)ns}, sb[2] = {(uint64_t)ld16 * 2, (uint64_t)o16_sstride * 2};
          uint32_t b[3] = {64, 32, 1};
          tma_out = make_tmap_bf16(&tmO, (void*)out16, 3, d, sb, b, CU_TENSOR_MAP_SWIZZLE_128B) ? 1 : 0;
        }
        auto gpair = [&](auto kern) -> int {
          if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return 1;
          cudaLaunchConfig_t cfg = {};
          cfg.gridDim = dim3(2u * (unsigned)std::min<long long>(ptiles, num_sms2 / 2)); cfg.blockDim = dim3(kGThreads);
          cfg.dynamicSmemBytes = smem; cfg.stream = st;
          cudaLaunchAttribute attr[1];
          attr[0].id = cudaLaunchAttributeClusterDimension;
          attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
          cfg.attrs = attr; cfg.numAttrs = 1;
          return cudaLaunchKernelEx(&cfg, kern, tmA, tmW, tmO, p, ns, ptm, ptn, tma_out) == cudaSuccess ? 0 : 1;
        };
        const int prc = ka == 1 ? gpair(dense_tc_pair_kernel<1>) : ka == 0 ? gpair(dense_tc_pair_kernel<0>) : gpair(dense_tc_pair_kernel<-1>);
        if (prc) return 1;
        g_launches.fetch_add(1, std::memory_order_relaxed);
        return cudaGetLastError() == cudaSuccess ? 0 : 1;
      }
    }
    const int tiles_m = (int)cdiv(M, 128), tiles_n = (int)cdiv(N, p.NT);
    const long long tiles = (long long)ns * tiles_m * tiles_n;
    if (tiles >= opt().persist_min_tiles && !opt().no_persist) {
      static int num_sms = 0;
      if (!num_sms) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 1;
      }
      const int ntc = p.NT <= 64 ? 64 : (p.NT <= 128 ? 128 : 256);
      const int smem = (ntc == 128 ? 3 : 4) * (16384 + (ntc / 64) * 8192) + (int)sizeof(PBarriers) + 1024;
      dim3 pgrid((unsigned)std::min<long long>(tiles, (ntc == 256 ? 1LL : 2LL) * num_sms));
      auto gop = [&](auto kern) -> int {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return 1;
        kern<<<pgrid, kGThreads, smem, st>>>(tmA, tmW, p, ns, tiles_m, tiles_n);
        return 0;
      };
#define DBX_GOP(NTM) (ka == 1 ? gop(dense_tc_persist_kernel<NTM, 1>) : ka == 0 ? gop(dense_tc_persist_kernel<NTM, 0>) : gop(dense_tc_persist_kernel<NTM, -1>))
      const int prc = ntc == 64 ? DBX_GOP(64) : (ntc == 128 ? DBX_GOP(128) : DBX_GOP(256));
#undef DBX_GOP
      if (prc) return 1;
      g_launches.fetch_add(1, std::memory_order_relaxed);
      return cudaGetLastError() == cudaSuccess ? 0 : 1;
    }
  }
  dim3 grid((unsigned)cdiv(N, p.NT), (unsigned)cdiv(M, 128), (unsigned)ns);
  auto go = [&](auto kern, int ntmax) -> int {
    const int smem = (ntmax == 64 ? 3 : kGStages) * (16384 + (ntmax / 64) * 8192) + (int)sizeof(GBarriers) + 1024;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) { g_tc_ok = false; return 1; }
    kern<<<grid, kGThreads, smem, st>>>(tmA, tmW, p);
    return 0;
  };
  int rc;
#define DBX_GO(NTM) (ka == 1 ? go(dense_tc_kernel<NTM, 1>, NTM) : ka == 0 ? go(dense_tc_kernel<NTM, 0>, NTM) : go(dense_tc_kernel<NTM, -1>, NTM))
  if (p.NT <= 64) rc = DBX_GO(64);
  else if (p.NT <= 128) rc = DBX_GO(128);
  else rc = DBX_GO(256);
#undef DBX_GO
  if (rc) return 1;
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

// Backward-data pass of one Dense layer (dense_tc_bwd_kernel).  dZ: bf16 [ns][M][ldz] (N_out = layer units = reduction);
// W16 chunk layout as for dense_tc_launch (kernel [fan_in][ldw]); result [ns][M][fan_in] as bf16 (ld16) or fp32 (dense rows).
// Needs ldz, ldw multiples of 8 and 16-byte aligned bases (TMA); returns 1 when not applicable.
int dense_tc_bwd_launch(const __nv_bfloat16* dZ, long long dz_sstride, int ldz, const __nv_bfloat16* Wbase, long long w_sstride_elems,
                        int ldw, int ns, int M, int fan_in, int fan_out, const __nv_bfloat16* aprev, long long ap_sstride, int ld_ap,
                        int act, __nv_bfloat16* out16, long long o16_sstride, int ld16, float* out32, long long o32_sstride,
                        cudaStream_t st, const float* fgsm_x, float fgsm_eps, float fgsm_lo, float fgsm_hi) {
  if (!g_tc_ok || !get_encode_fn()) return 1;
  if ((ldz & 7) || (ldw & 7) || (dz_sstride & 7) || (w_sstride_elems & 7) || ((uintptr_t)dZ & 15) || ((uintptr_t)Wbase & 15)) return 1;
  BwdParams p = {};
  p.M = M; p.N = fan_in; p.K = fan_out;
  p.aprev = aprev; p.ap_sstride = ap_sstride; p.ld_ap = ld_ap; p.act = act;
  p.out16 = out16; p.o16_sstride = o16_sstride; p.ld16 = ld16; p.out32 = out32; p.o32_sstride = o32_sstride;
  p.fgsm_x = fgsm_x; p.fgsm_eps = fgsm_eps; p.fgsm_lo = fgsm_lo; p.fgsm_hi = fgsm_hi;
  // column tile = template width = rows of the weight box.  128 columns even for wide layers: two CTAs per SM = two epilogue
  // warps per scheduler (these GEMMs have few rows per sample and are bound by their epilogue's instruction latency)
  const int ntc = fan_in <= 64 ? 64 : 128;
  p.NT = ntc;
  CUtensorMap tmZ, tmW;
  {
    uint64_t d[3] = {(uint64_t)fan_out, (uint64_t)M, (uint64_t)ns}, sb[2] = {(uint64_t)ldz * 2, (uint64_t)dz_sstride * 2};
    uint32_t b[3] = {64, 128, 1};
    if (!make_tmap_bf16(&tmZ, (void*)dZ, 3, d, sb, b, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  }
  {
    uint64_t d[3] = {(uint64_t)fan_out, (uint64_t)fan_in, (uint64_t)ns}, sb[2] = {(uint64_t)ldw * 2, (uint64_t)w_sstride_elems * 2};
    uint32_t b[3] = {64, (uint32_t)ntc, 1};
    if (!make_tmap_bf16(&tmW, (void*)Wbase, 3, d, sb, b, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  }
  const int tiles_m = (int)cdiv(M, 128), tiles_n = (int)cdiv(fan_in, ntc);
  const long long tiles = (long long)ns * tiles_m * tiles_n;
  if (tiles >= opt().persist_min_tiles && !opt().no_persist) {
    const int smem = (ntc == 128 ? 3 : 4) * (16384 + ntc * 128) + (int)sizeof(PBarriers) + 1024;
    dim3 pgrid((unsigned)std::min<long long>(tiles, (ntc == 256 ? 1LL : 2LL) * sm_count()));
    auto gop = [&](auto kern) -> int {
      if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return 1;
      kern<<<pgrid, kGThreads, smem, st>>>(tmZ, tmW, p, ns, tiles_m, tiles_n);
      return 0;
    };
    const int prc = ntc == 64 ? gop(dense_tc_bwd_persist_kernel<64>) : (ntc == 128 ? gop(dense_tc_bwd_persist_kernel<128>) : gop(dense_tc_bwd_persist_kernel<256>));
    if (prc) return 1;
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
  }
  dim3 grid((unsigned)tiles_n, (unsigned)tiles_m, (unsigned)ns);
  auto go = [&](auto kern, int ntmax) -> int {
    const int smem = (ntmax == 64 ? 3 : kGStages) * (16384 + ntmax * 128) + (int)sizeof(GBarriers) + 1024;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return 1;
    kern<<<grid, kGThreads, smem, st>>>(tmZ, tmW, p);
    return 0;
  };
  const int rc = ntc == 64 ? go(dense_tc_bwd_kernel<64>, 64) : (ntc == 128 ? go(dense_tc_bwd_kernel<128>, 128) : go(dense_tc_bwd_kernel<256>, 256));
  if (rc) return 1;
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

// dW (+)= A^T dZ (dense_tc_wgrad_kernel).  A: bf16 [B][lda] (the layer's input), dZ: bf16 [B][ldz]; out fp32 [fan_in][fan_out].
// lda, ldz multiples of 8, 16-byte aligned bases; returns 1 when not applicable.
int dense_tc_wgrad_launch(const __nv_bfloat16* A, int lda, const __nv_bfloat16* dZ, int ldz, int B, int fan_in, int fan_out, float* out,
                          int accumulate, cudaStream_t st) {
  if (!g_tc_ok || !get_encode_fn()) return 1;
  if ((lda & 7) || (ldz & 7) || ((uintptr_t)A & 15) || ((uintptr_t)dZ & 15) || ((uintptr_t)out & 15)) return 1;
  WgradParams p = {};
  p.Kin = fan_in; p.N = fan_out; p.B = B; p.out = out; p.accumulate = accumulate;
  // 128-column tiles even for wide layers: dW has few tiles (784 x 512: 7 x 4) and every CTA walks the whole batch
  const int ntc = fan_out <= 64 ? 64 : 128;
  p.NT = ntc;
  CUtensorMap tmA, tmZ;
  {
    uint64_t d[2] = {(uint64_t)fan_in, (uint64_t)B}, sb[1] = {(uint64_t)lda * 2};
    uint32_t b[2] = {64, 64};
    if (!make_tmap_bf16(&tmA, (void*)A, 2, d, sb, b, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  }
  {
    uint64_t d[2] = {(uint64_t)fan_out, (uint64_t)B}, sb[1] = {(uint64_t)ldz * 2};
    uint32_t b[2] = {64, 64};
    if (!make_tmap_bf16(&tmZ, (void*)dZ, 2, d, sb, b, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  }
  dim3 grid((unsigned)cdiv(fan_out, ntc), (unsigned)cdiv(fan_in, 128), 1);
  auto go = [&](auto kern, int ntmax) -> int {
    const int smem = (ntmax == 64 ? 3 : kGStages) * (16384 + (ntmax / 64) * 8192) + (int)sizeof(GBarriers) + 1024;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return 1;
    kern<<<grid, kGThreads, smem, st>>>(tmA, tmZ, p);
    return 0;
  };
  const int rc = ntc == 64 ? go(dense_tc_wgrad_kernel<64>, 64) : (ntc == 128 ? go(dense_tc_wgrad_kernel<128>, 128) : go(dense_tc_wgrad_kernel<256>, 256));
  if (rc) return 1;
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

// Launch one Dense layer whose weights are rows of the fp32 stored-sample slab: W element (k, n) of sample s is
// slab[(w_rows ? w_rows[s] : w_row0 + s) * slab_stride + w_off + k * N + n].  Needs N % 4 == 0, w_off % 4 == 0,
// slab_stride % 4 == 0 (16-byte TMA strides); returns 1 when not applicable so the caller can fall back.
int dense_tc_f32w_launch(const __nv_bfloat16* A, long long a_sstride, int lda, int a_rep, const float* slab, long long slab_stride,
                         long long slab_rows, long long w_off, const int32_t* w_rows, long long w_row0, int ns, int M, int N, int K,
                         const float* bias, long long b_sstride, long long b_off, int act, __nv_bfloat16* out16,
                         long long o16_sstride, int ld16, float* out32, long long o32_sstride, cudaStream_t st) {
  if (!g_tc_ok || !get_encode_fn()) return 1;
  if ((N & 3) || (w_off & 3) || (slab_stride & 3) || (lda & 7) || ((uintptr_t)slab & 15)) return 1;
  const bool wide = (N % 128 == 0) && !opt().no_f32w_wide;   // 128-column tiles (see FCfg)
  const int NT = wide ? 128 : 64;
  const int a_rows = (M <= 64) ? 64 : 128;                    // rows one A box delivers
  CUtensorMap tmA, tmW;
  {
    // a_rep > 0: the layer input is shared by all samples and replicated a_rep times a_sstride apart, so that the CTAs
    // of different samples do not all pull the same L2 lines; a_rep == 0: one input per sample
    uint64_t d[3] = {(uint64_t)K, (uint64_t)M, (uint64_t)(a_rep ? a_rep : ns)}, sb[2] = {(uint64_t)lda * 2, (uint64_t)a_sstride * 2};
    uint32_t b[3] = {64, (uint32_t)a_rows, 1};
    if (!make_tmap_bf16(&tmA, (void*)A, 3, d, sb, b, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  }
  {
    uint64_t d[3] = {(uint64_t)N, (uint64_t)K, (uint64_t)slab_rows}, sb[2] = {(uint64_t)N * 4, (uint64_t)slab_stride * 4};
    uint32_t b[3] = {(uint32_t)NT, 64, 1};
    if (!make_tmap_f32(&tmW, (void*)(slab + w_off), 3, d, sb, b, CU_TENSOR_MAP_SWIZZLE_NONE)) return 1;
  }
  FGemmParams fp = {};
  GemmParams& p = fp.g;
  p.M = M; p.N = N; p.K = K; p.NT = NT;
  p.a_shared = a_rep ? 1 : 0; p.bias = bias; p.b_sstride = b_sstride; p.b_off = b_off; p.rows = nullptr; p.row0 = 0; p.act = act;
  p.out16 = out16; p.o16_sstride = o16_sstride; p.ld16 = ld16; p.out32 = out32; p.o32_sstride = o32_sstride;
  fp.a_rep = a_rep;
  fp.a_box_bytes = a_rows * 128;
  fp.w_rows = w_rows; fp.w_row0 = w_row0;
  fp.ns = ns; fp.tiles_n = (int)cdiv(N, NT); fp.tiles_m = (int)cdiv(M, 128);
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 1;
  }
  const long long total = (long long)ns * fp.tiles_m * fp.tiles_n;
  auto go = [&](auto kern, int smem_b, int ctas_per_sm) -> int {
    const int smem = smem_b + (int)sizeof(FBarriers) + 1024;
    dim3 grid((unsigned)std::min<long long>(total, (long long)ctas_per_sm * num_sms));   // persistent: the resident CTAs
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return 1;
    kern<<<grid, kFThreads, smem, st>>>(tmA, tmW, fp);
    return 0;
  };
  const int ka = act_class(act);
  int rc;
  if (wide) rc = ka == 1 ? go(dense_tc_f32w_kernel<1, 128>, FCfg<128>::Smem, 1) : ka == 0 ? go(dense_tc_f32w_kernel<0, 128>, FCfg<128>::Smem, 1)
                                                                                         : go(dense_tc_f32w_kernel<-1, 128>, FCfg<128>::Smem, 1);
  else rc = ka == 1 ? go(dense_tc_f32w_kernel<1, 64>, FCfg<64>::Smem, 2) : ka == 0 ? go(dense_tc_f32w_kernel<0, 64>, FCfg<64>::Smem, 2)
                                                                                   : go(dense_tc_f32w_kernel<-1, 64>, FCfg<64>::Smem, 2);
  if (rc) return 1;
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

// =======================================================================================
// A layer whose INPUT is shared by all posterior samples and whose GEMM is tiny in K and N (config 4's first convolution:
// im2col of the call's input, K = 27, N = 32 output channels): run per sample it is one 128 x 32 x 32 tile per 1.5 us of
// barrier round trips and reads the im2col rows once per SAMPLE (7.7 us per sample, 27 % of config 4).  Here EIGHT samples
// share a tile: their kernels sit side by side as the N dimension, B = [K x (8 x 32)], so a 128-row block of A is read once
// for eight samples and feeds one M = 128, N = 256 MMA per 16 k; what remains is the write of the outputs
// (M x 32 bf16 per sample), eight 64-byte row segments per thread and tile.
//   warp 0  producer: W box [32 n x 32 k x 8 samples] per sample group (SWIZZLE_64B, lands as eight MN-major blocks of
//           2 KB = the N = 256 operand with LBO 2048), A boxes [32 k x 128 rows] (SWIZZLE_64B K-major) through a 4-deep ring
//   warp 1  MMA issuer: two K = 16 steps per tile, two 256-column TMEM accumulators
//   warps 2-5 epilogue: bias + activation + bf16 pack, 4 x 16-byte stores per (row, sample)
// Persistent CTAs over contiguous ranges of the [group][m tile] list (a group's kernels are loaded once per CTA).
// =======================================================================================
constexpr int kSAStages = 4, kSAThreads = 192, kSAGroup = 8;
struct SAParams {
  int M, K, ns, tiles_m, groups, act;
  const float* bias; long long b_sstride, b_off;
  __nv_bfloat16* out; long long o_sstride;   // [s][M][32]
};
struct __align__(8) SABarriers {
  uint64_t full[kSAStages], empty[kSAStages];
  uint64_t b_full[2], b_empty[2];
  uint64_t acc_full[2], acc_empty[2];
  uint32_t tmem_slot;
};

template <int kAct>
__global__ void __launch_bounds__(kSAThreads, 1)
dense_tc_shareda_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const SAParams p) {
  __shared__ float s_bias[2][kSAGroup * 32];
  uint8_t* smem = (uint8_t*)(((uintptr_t)gemm_smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t s_a = smem_u32(smem);                       // kSAStages x 8 KB
  const uint32_t s_b = s_a + kSAStages * 8192;               // 2 x 16 KB
  SABarriers* bars = (SABarriers*)(smem + kSAStages * 8192 + 2 * 16384);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long total = (long long)p.groups * p.tiles_m;
  const long long t_begin = (total * blockIdx.x) / gridDim.x, t_end = (total * (blockIdx.x + 1)) / gridDim.x;

  if (tid == 0) {
    for (int i = 0; i < kSAStages; ++i) { mbar_init(smem_u32(&bars->full[i]), 1); mbar_init(smem_u32(&bars->empty[i]), 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&bars->b_full[i]), 1); mbar_init(smem_u32(&bars->b_empty[i]), 1);
      mbar_init(smem_u32(&bars->acc_full[i]), 1); mbar_init(smem_u32(&bars->acc_empty[i]), 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(smem_u32(&bars->tmem_slot));
  if (warp == 0 && lane == 0) { prefetch_tmap(&tmA); prefetch_tmap(&tmW); }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = bars->tmem_slot;

  if (warp == 0) {
    uint32_t it = 0, gn = 0;
    int cur_g = -1;
    for (long long t = t_begin; t < t_end; ++t, ++it) {
      const int g = (int)(t / p.tiles_m), mt = (int)(t - (long long)g * p.tiles_m);
      if (g != cur_g) {
        const uint32_t gb = gn & 1;
        mbar_wait(smem_u32(&bars->b_empty[gb]), ((gn >> 1) & 1) ^ 1);
        if (elect_one()) {
          const uint32_t fb = smem_u32(&bars->b_full[gb]);
          mbar_expect_tx(fb, 16384);                          // samples past ns / rows past K arrive as zeros, counted
          tma_load_3d(s_b + gb * 16384, &tmW, fb, 0, 0, g * kSAGroup);
        }
        __syncwarp();
        cur_g = g; ++gn;
      }
      const uint32_t st = it % kSAStages, ph = (it / kSAStages) & 1;
      mbar_wait(smem_u32(&bars->empty[st]), ph ^ 1);
      if (elect_one()) {
        const uint32_t fb = smem_u32(&bars->full[st]);
        mbar_expect_tx(fb, 8192);
        tma_load_2d(s_a + st * 8192, &tmA, fb, 0, mt * 128);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    constexpr uint32_t kHi64 = (512u >> 4) | (1u << 14) | (4u << 29);      // 8-row / 8-k groups 512 B apart, SWIZZLE_64B
    constexpr uint32_t idesc = make_idesc_bf16(128, 256, 0, 1);
    uint32_t it = 0, gn = 0;
    int cur_g = -1;
    uint32_t gb = 0;
    for (long long t = t_begin; t < t_end; ++t, ++it) {
      const int g = (int)(t / p.tiles_m);
      if (g != cur_g) {
        gb = gn & 1;
        mbar_wait(smem_u32(&bars->b_full[gb]), (gn >> 1) & 1);
        cur_g = g; ++gn;
      }
      const bool last_of_group = (t + 1 == t_end) || ((int)((t + 1) / p.tiles_m) != g);
      const uint32_t ab = it & 1, st = it % kSAStages, ph = (it / kSAStages) & 1;
      mbar_wait(smem_u32(&bars->acc_empty[ab]), ((it >> 1) & 1) ^ 1);
      mbar_wait(smem_u32(&bars->full[st]), ph);
      fence_after_sync();
      if (elect_one()) {
        const uint32_t a_lo = (((s_a + st * 8192) & 0x3FFFF) >> 4) | (1u << 16);
        const uint32_t b_lo = (((s_b + gb * 16384) & 0x3FFFF) >> 4) | ((2048u >> 4) << 16);   // n-blocks (samples) 2 KB apart
        mma_ss_lh(tmem + ab * 256, a_lo, kHi64, b_lo, kHi64, idesc, 0);
        if (p.K > 16) mma_ss_lh(tmem + ab * 256, a_lo + 2, kHi64, b_lo + 64, kHi64, idesc, 1);
        mma_commit(smem_u32(&bars->empty[st]));
        mma_commit(smem_u32(&bars->acc_full[ab]));
        if (last_of_group) mma_commit(smem_u32(&bars->b_empty[gb]));
      }
      __syncwarp();
    }
  } else {
    const int gq = warp & 3;
    const int row = gq * 32 + lane;
    const int et = tid - 64;                                  // 0..127 among the epilogue threads
    const uint32_t lane_base = (uint32_t)(gq * 32) << 16;
    uint32_t it = 0, gn = 0;
    int cur_g = -1;
    uint32_t gb = 0;
    for (long long t = t_begin; t < t_end; ++t, ++it) {
      const int g = (int)(t / p.tiles_m), mt = (int)(t - (long long)g * p.tiles_m);
      if (g != cur_g) {
        gb = gn & 1;
        // this group's biases (the previous user of s_bias[gb] was two groups ago: every warp has passed it)
        for (int i = et; i < kSAGroup * 32; i += 128) {
          const int sI = g * kSAGroup + (i >> 5);
          s_bias[gb][i] = sI < p.ns ? p.bias[(long long)sI * p.b_sstride + p.b_off + (i & 31)] : 0.f;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        cur_g = g; ++gn;
      }
      const uint32_t ab = it & 1;
      const int gm = mt * 128 + row;
      mbar_wait(smem_u32(&bars->acc_full[ab]), (it >> 1) & 1);
      fence_after_sync();
#pragma unroll 2
      for (int q = 0; q < kSAGroup; ++q) {
        uint32_t v[32];
        tmem_ld32(tmem + lane_base + ab * 256 + q * 32, v);
        wait_ld();
        if (q == kSAGroup - 1) {
          fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&bars->acc_empty[ab]));
        }
        const int sI = g * kSAGroup + q;
        if (sI < p.ns && gm < p.M) {
          const float* bq = &s_bias[gb][q * 32];
          uint4* dst = reinterpret_cast<uint4*>(p.out + (long long)sI * p.o_sstride + (long long)gm * 32);
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) {
            uint32_t w[4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
              w[i] = pack_bf16x2(act_fixed<kAct>(p.act, __uint_as_float(v[c4 * 8 + 2 * i]) + bq[c4 * 8 + 2 * i]),
                                 act_fixed<kAct>(p.act, __uint_as_float(v[c4 * 8 + 2 * i + 1]) + bq[c4 * 8 + 2 * i + 1]));
            dst[c4] = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

// returns 0 when the layer was launched, 1 when the shape does not fit (the caller uses the per-sample GEMM)
int dense_tc_shareda_launch(const __nv_bfloat16* A, int lda, const __nv_bfloat16* Wbase, long long w_sstride_elems, int ldw, int ns, int M,
                            int N, int K, const float* bias, long long b_sstride, long long b_off, int act, __nv_bfloat16* out16,
                            long long o16_sstride, int ld16, cudaStream_t st) {
  if (!g_tc_ok || !get_encode_fn() || opt().no_shared_a) return 1;
  if (N != 32 || K > 32 || K < 1 || ldw != 32 || ld16 != 32 || ns < 2 || (lda & 7) || (o16_sstride & 7) || ((uintptr_t)out16 & 15) ||
      (w_sstride_elems & 7) || ((uintptr_t)Wbase & 15) || M < 128)
    return 1;
  CUtensorMap tmA, tmW;
  {
    uint64_t d[2] = {(uint64_t)K, (uint64_t)M}, sb[1] = {(uint64_t)lda * 2};
    uint32_t b[2] = {32, 128};
    if (!make_tmap_bf16(&tmA, (void*)A, 2, d, sb, b, CU_TENSOR_MAP_SWIZZLE_64B)) return 1;
  }
  {
    uint64_t d[3] = {32, (uint64_t)K, (uint64_t)ns}, sb[2] = {(uint64_t)ldw * 2, (uint64_t)w_sstride_elems * 2};
    uint32_t b[3] = {32, 32, (uint32_t)kSAGroup};
    if (!make_tmap_bf16(&tmW, (void*)Wbase, 3, d, sb, b, CU_TENSOR_MAP_SWIZZLE_64B)) return 1;
  }
  SAParams p = {};
  p.M = M; p.K = K; p.ns = ns; p.tiles_m = (int)cdiv(M, 128); p.groups = (int)cdiv(ns, kSAGroup); p.act = act;
  p.bias = bias; p.b_sstride = b_sstride; p.b_off = b_off; p.out = out16; p.o_sstride = o16_sstride;
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 1;
  }
  // (asks for more than half an SM's shared memory: the kernel allocates all 512 TMEM columns, so two CTAs must never share an SM)
  const int smem = std::max(kSAStages * 8192 + 2 * 16384 + (int)sizeof(SABarriers) + 1024, 118 * 1024);
  const long long total = (long long)p.groups * p.tiles_m;
  dim3 grid((unsigned)std::min<long long>(total, num_sms));
  auto go = [&](auto kern) -> int {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return 1;
    kern<<<grid, kSAThreads, smem, st>>>(tmA, tmW, p);
    return 0;
  };
  const int ka = act_class(act);
  if (ka == 1 ? go(dense_tc_shareda_kernel<1>) : ka == 0 ? go(dense_tc_shareda_kernel<0>) : go(dense_tc_shareda_kernel<-1>)) return 1;
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

// dynamic shared memory of conv_tc_kernel.  The kernel tensor stays resident when it is at most 40 KB; then, if
// R = 128 / (Wo + kw - 1) >= 1 output rows fit the 128 accumulator rows, one input patch per tile serves all taps.
static int conv_tc_smem(const Layer& L, int* w_res, int* patch, int* R_out, int pool = 0, int* stages_out = nullptr) {
  const int Cin = L.in_c, taps = L.kh * L.kw, nblk = (L.out_c + 63) / 64;
  const int w_tap_b = nblk * Cin * 128;
  const int res = taps * w_tap_b <= 40 * 1024 ? 1 : 0;
  const int PW = L.out_w + L.kw - 1;
  const int pm = (res && PW <= 128) ? 1 : 0;
  int R = std::min(pm ? 128 / PW : 128 / L.out_w, L.out_h);
  if (pool) R &= ~1;                                   // pooling windows must not straddle tiles
  const int a_bytes = pm ? (((L.kh - 1) * PW + L.kw - 1 + 128) * Cin * 2 + 1023) / 1024 * 1024 : 128 * Cin * 2;
  if (w_res) *w_res = res;
  if (patch) *patch = pm;
  if (R_out) *R_out = R;
  const int fixed = (res ? taps * w_tap_b : 0) + ((int)sizeof(CBarriers) + 15) / 16 * 16 + 1024 + (pool ? 128 * L.out_c * 2 : 0);
  const int per_stage = res ? a_bytes : a_bytes + w_tap_b;
  int stages = kCStages;
  if (pool) while (stages > 2 && fixed + stages * per_stage > 110 * 1024) --stages;      // two CTAs per SM
  if (stages_out) *stages_out = stages;
  return fixed + stages * per_stage;
}

// layer-level part of conv_tc_launch's scope test (the executor sizes its im2col buffer with it)
bool conv_tc_eligible(const Layer& L) {
  const int Cin = L.in_c, Cout = L.out_c, Wo = L.out_w;
  if (L.kind != DBX_CONV2D || L.sh != 1 || L.sw != 1 || !(Cin == 16 || Cin == 32 || Cin == 64) || Cout > 128 || (Cout & 7) || Wo > 128 || Wo < 1)
    return false;
  return conv_tc_smem(L, nullptr, nullptr, nullptr) <= 110 * 1024;
}

// Direct convolution launch.  in: bf16 NHWC [ns or 1][NB][H][W][Cin]; W16: the sample's HWIO kernel as [kh*kw*Cin][ldw] bf16
// rows; out: bf16 NHWC [ns][NB][Ho][Wo][Cout].  Returns 1 when the layer is outside the kernel's scope (the caller then
// uses im2col + dense_tc_kernel).
// Can the 2x2 / stride-2 MaxPool behind this convolution be done in its epilogue?
bool conv_tc_pool_eligible(const Layer& L, const Layer& P) {
  if (!conv_tc_eligible(L) || P.kind != DBX_MAXPOOL2D || P.kh != 2 || P.kw != 2 || P.sh != 2 || P.sw != 2) return false;
  if (opt().no_pool_fuse) return false;
  const int Cout = L.out_c;
  if (!(Cout == 8 || Cout == 16 || Cout == 32 || Cout == 64) || L.out_h < 2 || L.out_w < 2) return false;
  int R = 0, stages = 0;
  const int smem = conv_tc_smem(L, nullptr, nullptr, &R, 1, &stages);
  return R >= 2 && stages >= 3 && smem <= 110 * 1024;
}

int conv_tc_launch(const __nv_bfloat16* in, long long in_sstride, int in_shared, const __nv_bfloat16* W16, long long w_sstride_elems,
                   int ldw, int ns, int NB, const Layer& L, const float* bias, long long b_sstride, long long b_off, int act,
                   __nv_bfloat16* out, long long o_sstride, cudaStream_t st, int pool) {
  if (!g_tc_ok || !get_encode_fn()) return 1;
  const int Cin = L.in_c, Cout = L.out_c, Wo = L.out_w, Ho = L.out_h;
  if (!conv_tc_eligible(L)) return 1;
  if ((in_sstride & 7) || (o_sstride & 7) || (ldw & 7) || ((uintptr_t)in & 15) || ((uintptr_t)out & 15)) return 1;
  int w_res = 0, patch = 0, R = 1, stages = kCStages;
  const int smem = conv_tc_smem(L, &w_res, &patch, &R, pool, &stages);
  const int PW = Wo + L.kw - 1;
  CUtensorMap tmA, tmW;
  {
    uint64_t d[5] = {(uint64_t)Cin, (uint64_t)L.in_w, (uint64_t)L.in_h, (uint64_t)NB, (uint64_t)(in_shared ? 1 : ns)};
    uint64_t sb[4] = {(uint64_t)Cin * 2, (uint64_t)L.in_w * Cin * 2, (uint64_t)L.in_h * L.in_w * Cin * 2,
                      (uint64_t)(in_shared ? (long long)NB * L.in_h * L.in_w * Cin : in_sstride) * 2};
    uint32_t b[5] = {(uint32_t)Cin, (uint32_t)(patch ? PW : Wo), (uint32_t)(patch ? R + L.kh - 1 : R), 1, 1};
    const CUtensorMapSwizzle swz = Cin == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (Cin == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    if (!make_tmap_bf16(&tmA, (void*)in, 5, d, sb, b, swz)) return 1;
  }
  {
    uint64_t d[3] = {(uint64_t)ldw, (uint64_t)L.w_rows, (uint64_t)ns}, sb[2] = {(uint64_t)ldw * 2, (uint64_t)w_sstride_elems * 2};
    uint32_t b[3] = {64, (uint32_t)Cin, 1};
    if (!make_tmap_bf16(&tmW, (void*)W16, 3, d, sb, b, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  }
  ConvParams p = {};
  p.NB = NB; p.Ho = Ho; p.Wo = Wo; p.Cin = Cin; p.Cout = Cout; p.kh = L.kh; p.kw = L.kw; p.pad_t = L.pad_t; p.pad_l = L.pad_l;
  p.R = R; p.tiles_per_img = (int)cdiv(Ho, R); p.ns = ns; p.a_shared = in_shared; p.act = act; p.w_res = w_res;
  p.patch = patch; p.PW = PW; p.base_off = 0; p.stages = stages; p.pool = pool;
  p.bias = bias; p.b_sstride = b_sstride; p.b_off = b_off; p.out = out; p.o_sstride = o_sstride;
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 1;
  }
  const long long total = (long long)ns * NB * p.tiles_per_img;
  dim3 grid((unsigned)std::min<long long>(total, 2LL * num_sms));
  auto go = [&](auto kern) -> int {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return 1;
    kern<<<grid, kCThreads, smem, st>>>(tmA, tmW, p);
    return 0;
  };
  const int ka = act_class(act);
  if (ka == 1 ? go(conv_tc_kernel<1>) : ka == 0 ? go(conv_tc_kernel<0>) : go(conv_tc_kernel<-1>)) return 1;
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

bool ibp_tc_eligible(int N, int K) {
  if (!g_tc_ok || !get_encode_fn()) return false;
  if ((K & 7) || (N & 7)) return false;
  if (opt().no_ibp_tc) return false;
  return true;
}

// One hidden layer of the interval forward on tensor cores (see ibp_tc_kernel).  A_mu / A_r: bf16 [ns or 1][M][K] (K % 8 == 0);
// outputs bf16 [ns][M][N] (N % 8 == 0) at n_sstride.  Returns 1 when not applicable.
int ibp_tc_launch(const __nv_bfloat16* Amu, const __nv_bfloat16* Ar, long long a_sstride, int a_shared, const __nv_bfloat16* Wbase,
                  long long w_sstride_elems, int ldw, int ns, int M, int N, int K, const float* bias, long long b_sstride, long long b_off,
                  int act, __nv_bfloat16* n_mu, __nv_bfloat16* n_r, long long n_sstride, cudaStream_t st) {
  if (!ibp_tc_eligible(N, K) || (a_sstride & 7) || (n_sstride & 7)) return 1;
  CUtensorMap tmAmu, tmAr, tmW;
  for (int which = 0; which < 2; ++which) {
    const __nv_bfloat16* A = which ? Ar : Amu;
    CUtensorMap* tm = which ? &tmAr : &tmAmu;
    if (a_shared) {
      uint64_t d[2] = {(uint64_t)K, (uint64_t)M}, sb[1] = {(uint64_t)K * 2};
      uint32_t b[2] = {64, 128};
      if (!make_tmap_bf16(tm, (void*)A, 2, d, sb, b, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
    } else {
      uint64_t d[3] = {(uint64_t)K, (uint64_t)M, (uint64_t)ns}, sb[2] = {(uint64_t)K * 2, (uint64_t)a_sstride * 2};
      uint32_t b[3] = {64, 128, 1};
      if (!make_tmap_bf16(tm, (void*)A, 3, d, sb, b, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
    }
  }
  {
    uint64_t d[3] = {(uint64_t)N, (uint64_t)K, (uint64_t)ns}, sb[2] = {(uint64_t)ldw * 2, (uint64_t)w_sstride_elems * 2};
    uint32_t b[3] = {64, 64, 1};
    if (!make_tmap_bf16(&tmW, (void*)Wbase, 3, d, sb, b, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  }
  CUtensorMap tmOmu, tmOr;
  for (int which = 0; which < 2; ++which) {
    uint64_t d[3] = {(uint64_t)N, (uint64_t)M, (uint64_t)ns}, sb[2] = {(uint64_t)N * 2, (uint64_t)n_sstride * 2};
    uint32_t b[3] = {64, 32, 1};
    if (!make_tmap_bf16(which ? &tmOr : &tmOmu, (void*)(which ? n_r : n_mu), 3, d, sb, b, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  }
  IbpGemmParams p = {};
  p.M = M; p.N = N; p.K = K; p.ns = ns; p.tiles_m = (int)cdiv(M, 128); p.tiles_n = (int)cdiv(N, kINT); p.a_shared = a_shared; p.act = act;
  p.bias = bias; p.b_sstride = b_sstride; p.b_off = b_off; p.n_mu = n_mu; p.n_r = n_r; p.n_sstride = n_sstride;
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 1;
  }
  const int smem = kIStages * kIStageB + kIAbs * kIAbsB + kIOutB + kINT * 4 + (int)sizeof(IBarriers) + 1024;
  const int ka = act_class(act);
  auto kern = ka == 1 ? ibp_tc_kernel<1> : (ka == 0 ? ibp_tc_kernel<0> : ibp_tc_kernel<-1>);
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return 1;
  const long long total = (long long)ns * p.tiles_m * p.tiles_n;
  dim3 grid((unsigned)std::min<long long>(total, (long long)num_sms));
  kern<<<grid, kIThreads, smem, st>>>(tmAmu, tmAr, tmW, tmOmu, tmOr, p);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int im2col_launch(const __nv_bfloat16* in, long long in_sstride, __nv_bfloat16* out, long long out_sstride, int ns, int NB,
                  const Layer& L, int Kp, cudaStream_t st) {
  if (L.in_c % 8 == 0 && Kp == L.kh * L.kw * L.in_c && (in_sstride % 8) == 0 && (out_sstride % 8) == 0) {
    const long long total8 = (long long)NB * L.out_h * L.out_w * (Kp / 8);
    dim3 grid8((unsigned)std::min<long long>(std::max<long long>(cdiv(total8, 256), 1), sm_count() * 32), (unsigned)ns);
    im2col_bf16_vec8_kernel<<<grid8, 256, 0, st>>>(in, in_sstride, out, out_sstride, NB, L.in_h, L.in_w, L.in_c, L.out_h,
                                                   L.out_w, L.kh, L.kw, L.sh, L.sw, L.pad_t, L.pad_l);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
  }
  const long long total = (long long)NB * L.out_h * L.out_w * Kp;
  dim3 grid((unsigned)std::min<long long>(std::max<long long>(cdiv(total, 256), 1), sm_count() * 32), (unsigned)ns);
  im2col_bf16_kernel<<<grid, 256, 0, st>>>(in, in_sstride, out, out_sstride, NB, L.in_h, L.in_w, L.in_c, L.out_h, L.out_w,
                                           L.kh, L.kw, L.sh, L.sw, L.pad_t, L.pad_l, Kp);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace dbx
