// deepbayes_b200 engine: model plan, generic sm_100a kernels (every layer kind, fp32 and
// bf16 storage) and the C ABI of include/deepbayes_b200.h.  The tcgen05 fast path for MLPs
// lives in fused_mlp.cu and is tried first by dbx_predict / dbx_count_predicate.
//
// Reference behaviour restated here (paths relative to the reference checkout):
//   sampling      deepbayes/posteriormodel.py:60-79, optimizers/optimizer.py:197-202,
//                 optimizers/bayesbybackprop.py:50-56,176-181
//   MC predictive deepbayes/posteriormodel.py:81-101 (probs), :114-139 (logits)
//   forward       Keras Sequential Dense/Conv2D/MaxPool/Flatten (third-party)
//   predicate     deepbayes/analyzers/verifiers.py:537-557, 563-653
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include "engine.h"

namespace dbx {

thread_local std::string g_err;
std::atomic<long long> g_launches{0};
static int g_last_path = 0;
static bool g_used_tc = false;

int set_err(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return 1;
}

// ---- process-wide options (see common.cuh) ------------------------------------------------------
struct OptEntry { const char* name; long long Options::*field; };
static const OptEntry kOptTable[] = {
  {"NO_FUSED", &Options::no_fused}, {"NO_STREAM", &Options::no_stream}, {"NO_TC", &Options::no_tc},
  {"NO_PAIR_GEMM", &Options::no_pair_gemm}, {"NO_TMA_STORE", &Options::no_tma_store}, {"NO_PERSIST", &Options::no_persist},
  {"PERSIST_MIN_TILES", &Options::persist_min_tiles}, {"NO_F32W", &Options::no_f32w},
  {"NO_DIRECT_CONV", &Options::no_direct_conv}, {"NO_POOL_FUSE", &Options::no_pool_fuse}, {"NO_IBP_TC", &Options::no_ibp_tc},
  {"IBP_OVERLAP", &Options::ibp_overlap}, {"IBP_PRIO", &Options::ibp_prio}, {"GENERIC_OVERLAP", &Options::generic_overlap},
  {"FUSED_OVERLAP", &Options::fused_overlap}, {"FUSED_PRIO", &Options::fused_prio}, {"FUSED_KERNEL", &Options::fused_kernel},
  {"FUSED_ORDER", &Options::fused_order}, {"FUSED_NS", &Options::fused_ns}, {"FUSED_KD", &Options::fused_kd}, {"NO_SHARED_A", &Options::no_shared_a}, {"NO_F32W_WIDE", &Options::no_f32w_wide}, {"FUSED_STREAM", &Options::fused_stream}, {"FUSED_NP", &Options::fused_np}, {"FUSED_L3AT", &Options::fused_l3at}, {"FUSED_TRACE", &Options::fused_trace}, {"FUSED_TRACE_ONLY", &Options::fused_trace_only}, {"WS_MB", &Options::ws_mb},
  {"TRAIN_GRAPH", &Options::train_graph},
};
static void options_from_env(Options& o) {
  o = Options();
  for (const OptEntry& e : kOptTable) {
    const std::string var = std::string("DBX_") + e.name;
    if (const char* v = getenv(var.c_str())) o.*(e.field) = atoll(v);
  }
}
int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

Options& opt() {
  static Options o;
  static bool init = false;
  if (!init) { options_from_env(o); init = true; }
  return o;
}

struct ProfRec { cudaEvent_t a, b; double flops; };
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
static cudaEvent_t g_prof_open = nullptr;

void prof_begin(cudaStream_t st) {
  if (!g_prof_on) return;
  cudaEventCreate(&g_prof_open);
  cudaEventRecord(g_prof_open, st);
}
void prof_end(cudaStream_t st, double flops) {
  if (!g_prof_on || !g_prof_open) return;
  ProfRec r; r.a = g_prof_open; r.flops = flops; g_prof_open = nullptr;
  cudaEventCreate(&r.b);
  cudaEventRecord(r.b, st);
  g_prof.push_back(r);
}

int ensure_ws(Workspace& w, size_t bytes) {
  if (w.bytes >= bytes) return 0;
  if (w.ptr) DBX_CUDA(cudaFree(w.ptr));
  w.ptr = nullptr; w.bytes = 0;
  DBX_CUDA(cudaMalloc(&w.ptr, bytes));
  w.bytes = bytes;
  return 0;
}

// ---------------------------------------------------------------------------------------
// Kernels
// ---------------------------------------------------------------------------------------

// W_s = mu + (inflate*sigma)*eps_s, written as fp32 in flat Keras order.
// One thread = one Philox block = 4 consecutive elements of one sample.
__global__ void sample_f32_kernel(const float* __restrict__ mu, const float* __restrict__ sigma,
                                  const float* __restrict__ eps, float inflate, uint64_t seed, int64_t s0,
                                  int64_t n, int64_t P, int unfused, float* __restrict__ W,
                                  float* __restrict__ eps_out) {
  const int64_t nblk = (P + 3) >> 2;
  const int64_t total = n * nblk;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s = t / nblk, blk = t - s * nblk;
    const int64_t j0 = blk << 2;
    float z[4];
    if (eps == nullptr) philox_normal4(seed, (uint64_t)(s0 + s), (uint64_t)blk, z);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t j = j0 + i;
      if (j < P) {
        const float e = eps ? eps[s * P + j] : z[i];
        const float sc = inflate * sigma[j];
        const float w = unfused ? __fadd_rn(mu[j], __fmul_rn(sc, e)) : fmaf(sc, e, mu[j]);
        W[s * P + j] = w;
        if (eps_out) eps_out[s * P + j] = e;
      }
    }
  }
}

// Same draw, scattered into the bf16 device layout (kernels -> padded bf16 rows, biases ->
// fp32 side array).  src != nullptr converts stored samples instead (rows idx[s] / j0+s).
// grid = (blocks over Philox blocks, samples); one thread = 4 consecutive flat elements.
__global__ void __launch_bounds__(256)
sample_bf16_kernel(const float* __restrict__ mu, const float* __restrict__ sigma, const float* __restrict__ eps,
                   float inflate, uint64_t seed, int64_t s0, int P, const float* __restrict__ src,
                   const int32_t* __restrict__ idx, int64_t j0row, int64_t src_stride, const TensorTable tab, const BlockRanges rg,
                   __nv_bfloat16* __restrict__ W16, int64_t P16, float* __restrict__ B32, int64_t PB) {
  const int s = blockIdx.y;
  const int nblk = (P + 3) >> 2;
  const uint64_t sample = (uint64_t)(s0 + s);
  const float* srow = src ? src + (idx ? (int64_t)idx[s] : (j0row + s)) * src_stride : nullptr;
  const float* erow = eps ? eps + (int64_t)s * P : nullptr;
  __nv_bfloat16* wrow = W16 + (int64_t)s * P16;
  float* brow = B32 + (int64_t)s * PB;
  (void)nblk;
  for (int ri = 0; ri < rg.n; ++ri) {
  int tfirst = 0;
  for (int blk = rg.lo[ri] + blockIdx.x * blockDim.x + threadIdx.x; blk - (int)threadIdx.x < rg.hi[ri]; blk += gridDim.x * blockDim.x) {
    if (blk >= rg.hi[ri]) continue;
    const int j0 = blk << 2;
    const bool full4 = (j0 + 4 <= P);
    float w[4];
    if (srow) {
      if (full4 && (src_stride & 3) == 0) { const float4 v = *reinterpret_cast<const float4*>(srow + j0); w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w; }
      else { for (int i = 0; i < 4; ++i) w[i] = (j0 + i < P) ? srow[j0 + i] : 0.f; }
    } else {
      float z[4];
      if (erow == nullptr) philox_normal4(seed, sample, (uint64_t)blk, z);
      else { for (int i = 0; i < 4; ++i) z[i] = (j0 + i < P) ? erow[j0 + i] : 0.f; }
      if (full4) {
        const float4 m4 = *reinterpret_cast<const float4*>(mu + j0);
        const float4 s4 = *reinterpret_cast<const float4*>(sigma + j0);
        w[0] = fmaf(inflate * s4.x, z[0], m4.x); w[1] = fmaf(inflate * s4.y, z[1], m4.y);
        w[2] = fmaf(inflate * s4.z, z[2], m4.z); w[3] = fmaf(inflate * s4.w, z[3], m4.w);
      } else {
        for (int i = 0; i < 4; ++i) w[i] = (j0 + i < P) ? fmaf(inflate * sigma[j0 + i], z[i], mu[j0 + i]) : 0.f;
      }
    }
    // which tensor holds element j0: the 1024 elements of this block iteration usually sit in
    // one tensor (block-uniform test), otherwise count the tensor ends below j0
    const int jb = (blk - (int)threadIdx.x) << 2;
    int ti = tfirst;
    while (ti + 1 < tab.n && jb >= tab.src_end[ti]) ++ti;   // uniform: advances rarely
    tfirst = ti;
    if (jb + 1024 > tab.src_end[ti])
      for (int i = ti; i + 1 < tab.n; ++i) ti += (j0 >= tab.src_end[i]) ? 1 : 0;
    const int rel = j0 - tab.src_off[ti];
    if (!tab.is_bias[ti] && !tab.blocked8[ti] && tab.cols_pad[ti] == tab.cols[ti] && rel + 4 <= tab.size[ti] && ((tab.dst_off[ti] + rel) & 3) == 0) {
      // fast path: 4 elements of an unpadded kernel tensor -> one 8-byte store (the tensor's flat offset need not be a
      // multiple of 4 -- bias vectors of 5 or 6 entries before it -- in which case the store would be misaligned)
      uint2 v;
      v.x = tc_pack_bf16x2(w[0], w[1]);
      v.y = tc_pack_bf16x2(w[2], w[3]);
      *reinterpret_cast<uint2*>(wrow + tab.dst_off[ti] + rel) = v;
    } else {
      for (int i = 0; i < 4; ++i) {
        const int j = j0 + i;
        if (j >= P) break;
        while (ti + 1 < tab.n && j >= tab.src_end[ti]) ++ti;
        const int r = j - tab.src_off[ti];
        if (tab.is_bias[ti] == 1) {
          brow[tab.dst_off[ti] + r] = w[i];
        } else if (tab.is_bias[ti] == 2) {
          // output-layer kernel kept as fp32 rows of 12 (bf16-rounded values) next to the biases: the
          // CTA-pair kernel applies it with FFMA in the epilogue
          const int row = r / tab.cols[ti], col = r - row * tab.cols[ti];
          brow[tab.dst_off[ti] + row * 12 + col] = __bfloat162float(__float2bfloat16_rn(w[i]));
        } else {
          const int row = r / tab.cols[ti], col = r - row * tab.cols[ti];
          const int d = tab.blocked8[ti] ? (((col >> 3) * tab.rows[ti] + row) * 8 + (col & 7)) : (row * tab.cols_pad[ti] + col);
          wrow[tab.dst_off[ti] + d] = __float2bfloat16_rn(w[i]);
        }
      }
    }
  }
  }
}

// Fast path for one large, unpadded kernel tensor whose flat range is 4-aligned: Philox block ->
// 4 normals -> W = mu + sigma*eps -> one 8-byte bf16 store.  No table, no division.  A thread keeps its 8 (mu, sigma)
// pairs in registers for kSamplesPerThread consecutive samples (blockIdx.y covers that many), so the fp32 posterior is
// read from L2 once per group instead of once per sample (8 B/parameter/sample was 4x the bf16 bytes written).
constexpr int kSamplesPerThread = 4;
#ifndef DBX_SAMPLER_MIN_BLOCKS
#define DBX_SAMPLER_MIN_BLOCKS 1   // (6 = a 40-register cap so that three blocks fit beside a fused CTA: it spills 88 bytes and config 5 went 86 -> 130 ms)
#endif
__global__ void __launch_bounds__(256, DBX_SAMPLER_MIN_BLOCKS)
sample_bf16_fast_kernel(const float* __restrict__ mu, const float* __restrict__ sigma, float inflate, uint64_t seed,
                        int64_t s0, int nc, int blk_lo, int blk_hi, int dst_delta, __nv_bfloat16* __restrict__ W16,
                        int64_t P16) {
  const int sy = blockIdx.y * kSamplesPerThread;
  const int nt = min(kSamplesPerThread, nc - sy);
  // two Philox blocks (8 consecutive weights) per thread and iteration: two independent Philox /
  // Box-Muller chains in flight per thread
  const int npair = (blk_hi - blk_lo + 1) >> 1;
  for (int pi = blockIdx.x * blockDim.x + threadIdx.x; pi < npair; pi += gridDim.x * blockDim.x) {
    const int blk = blk_lo + 2 * pi;
    const bool two = (blk + 1 < blk_hi);
    const int j0 = blk << 2;
    const float4 m0 = __ldg(reinterpret_cast<const float4*>(mu + j0));
    float4 g0 = __ldg(reinterpret_cast<const float4*>(sigma + j0));
    g0.x *= inflate; g0.y *= inflate; g0.z *= inflate; g0.w *= inflate;
    float4 m1 = make_float4(0.f, 0.f, 0.f, 0.f), g1 = m1;
    if (two) {
      m1 = __ldg(reinterpret_cast<const float4*>(mu + j0 + 4));
      g1 = __ldg(reinterpret_cast<const float4*>(sigma + j0 + 4));
      g1.x *= inflate; g1.y *= inflate; g1.z *= inflate; g1.w *= inflate;
    }
    for (int t = 0; t < nt; ++t) {
      const uint64_t sample = (uint64_t)(s0 + sy + t);
      __nv_bfloat16* wrow = W16 + (int64_t)(sy + t) * P16 + dst_delta;   // dst index = flat index + dst_delta
      float z[8];
      philox_normal4(seed, sample, (uint64_t)blk, z);
      uint2 v0;
      v0.x = tc_pack_bf16x2(fmaf(g0.x, z[0], m0.x), fmaf(g0.y, z[1], m0.y));
      v0.y = tc_pack_bf16x2(fmaf(g0.z, z[2], m0.z), fmaf(g0.w, z[3], m0.w));
      if (two) {
        philox_normal4(seed, sample, (uint64_t)(blk + 1), z + 4);
        uint2 v1;
        v1.x = tc_pack_bf16x2(fmaf(g1.x, z[4], m1.x), fmaf(g1.y, z[5], m1.y));
        v1.y = tc_pack_bf16x2(fmaf(g1.z, z[6], m1.z), fmaf(g1.w, z[7], m1.w));
        if ((((uintptr_t)(wrow + j0)) & 15) == 0) *reinterpret_cast<uint4*>(wrow + j0) = make_uint4(v0.x, v0.y, v1.x, v1.y);
        else { *reinterpret_cast<uint2*>(wrow + j0) = v0; *reinterpret_cast<uint2*>(wrow + j0 + 4) = v1; }
      } else {
        *reinterpret_cast<uint2*>(wrow + j0) = v0;
      }
    }
  }
}

__global__ void softplus_kernel(const float* __restrict__ rho, float* __restrict__ sigma, int64_t P) {
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < P; j += (int64_t)gridDim.x * blockDim.x)
    sigma[j] = softplus_tf(rho[j]);
}

__global__ void philox_raw_kernel(uint64_t seed, uint64_t s, int64_t P, uint32_t* __restrict__ out) {
  const int64_t nblk = (P + 3) >> 2;
  for (int64_t blk = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; blk < nblk; blk += (int64_t)gridDim.x * blockDim.x) {
    Philox4 r = philox4x32_10((uint32_t)blk, (uint32_t)(blk >> 32), (uint32_t)s, (uint32_t)(s >> 32),
                              (uint32_t)seed, (uint32_t)(seed >> 32));
    const uint32_t v[4] = {r.x, r.y, r.z, r.w};
    for (int i = 0; i < 4; ++i)
      if (blk * 4 + i < P) out[blk * 4 + i] = v[i];
  }
}

// fp32 -> T into `reps` copies `rep_stride` elements apart (the layer-0 input of the stored-posterior path is read by every
// CTA of every sample: replicas keep them from all pulling the same L2 lines)
template <typename T>
__global__ void cast_rep_kernel(const float* __restrict__ in, T* __restrict__ out, int64_t nelem, int reps, int64_t rep_stride) {
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < nelem; j += (int64_t)gridDim.x * blockDim.x) {
    const T v = from_f32<T>(in[j]);
    for (int r = 0; r < reps; ++r) out[r * rep_stride + j] = v;
  }
}

template <typename T>
__global__ void cast_kernel(const float* __restrict__ in, T* __restrict__ out, int64_t nelem) {
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < nelem; j += (int64_t)gridDim.x * blockDim.x)
    out[j] = from_f32<T>(in[j]);
}

// Batched Dense: out[s] = act(A[s] (M x K) * W[s] (K x N) + b[s]).  CUDA-core FFMA, fp32
// accumulation in ascending-k order.  64x64x16 tiles, 4x4 outputs per thread.
template <typename T, typename TO>
__global__ void __launch_bounds__(256)
dense_kernel(const T* __restrict__ A, int64_t a_sstride, const T* __restrict__ Wbase, int64_t w_sstride,
             const int32_t* __restrict__ rows, int64_t row0, int64_t w_off, int ldw,
             const float* __restrict__ Bbase, int64_t b_sstride, int64_t b_off,
             TO* __restrict__ out, int64_t o_sstride, int M, int N, int K, int act) {
  constexpr int BM = 64, BN = 64, BK = 16;
  __shared__ float As[BK][BM + 4];
  __shared__ float Ws[BK][BN + 4];
  const int s = blockIdx.z;
  const int64_t wrow = rows ? (int64_t)rows[s] : (row0 + s);
  const T* Ap = A + (int64_t)s * a_sstride;
  const T* Wp = Wbase + wrow * w_sstride + w_off;
  const float* bp = Bbase + wrow * b_sstride + b_off;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int t = threadIdx.x, ty = t >> 4, tx = t & 15;
  float acc[4][4] = {};
  const int ar = t >> 2, ak = (t & 3) << 2;   // A tile: row ar, k ak..ak+3
  const int wk = t >> 4, wn = (t & 15) << 2;  // W tile: k wk, n wn..wn+3
  for (int k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int gm = m0 + ar, gk = k0 + ak + i;
      As[ak + i][ar] = (gm < M && gk < K) ? to_f32<T>(Ap[(int64_t)gm * K + gk]) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int gk = k0 + wk, gn = n0 + wn + i;
      Ws[wk][wn + i] = (gk < K && gn < N) ? to_f32<T>(Wp[(int64_t)gk * ldw + gn]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[4], w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) w[j] = Ws[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }
  TO* op = out + (int64_t)s * o_sstride;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn < N) op[(int64_t)gm * N + gn] = from_f32<TO>(apply_act(act, acc[i][j] + bp[gn]));
    }
  }
}

// Batched direct Conv2D, NHWC x HWIO -> NHWC (one thread per output element, co fastest).
template <typename T, typename TO>
__global__ void conv2d_kernel(const T* __restrict__ A, int64_t a_sstride, const T* __restrict__ Wbase,
                              int64_t w_sstride, const int32_t* __restrict__ rows, int64_t row0, int64_t w_off,
                              int ldw, const float* __restrict__ Bbase, int64_t b_sstride, int64_t b_off,
                              TO* __restrict__ out, int64_t o_sstride, int NB, int H, int Wd, int Cin, int Ho,
                              int Wo, int Cout, int kh, int kw, int sh, int sw, int pt, int pl, int act) {
  const int s = blockIdx.y;
  const int64_t wrow = rows ? (int64_t)rows[s] : (row0 + s);
  const T* Ap = A + (int64_t)s * a_sstride;
  const T* Wp = Wbase + wrow * w_sstride + w_off;
  const float* bp = Bbase + wrow * b_sstride + b_off;
  TO* op = out + (int64_t)s * o_sstride;
  const int64_t total = (int64_t)NB * Ho * Wo * Cout;
  for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
    const int co = (int)(o % Cout);
    int64_t r = o / Cout;
    const int wo = (int)(r % Wo); r /= Wo;
    const int ho = (int)(r % Ho);
    const int nb = (int)(r / Ho);
    float acc = 0.f;
    for (int i = 0; i < kh; ++i) {
      const int hi = ho * sh + i - pt;
      if (hi < 0 || hi >= H) continue;
      for (int j = 0; j < kw; ++j) {
        const int wi = wo * sw + j - pl;
        if (wi < 0 || wi >= Wd) continue;
        const T* ap = Ap + (((int64_t)nb * H + hi) * Wd + wi) * Cin;
        const T* wp = Wp + ((int64_t)(i * kw + j) * Cin) * ldw + co;
        for (int c = 0; c < Cin; ++c) acc = fmaf(to_f32<T>(ap[c]), to_f32<T>(wp[(int64_t)c * ldw]), acc);
      }
    }
    op[o] = from_f32<TO>(apply_act(act, acc + bp[co]));
  }
}

template <typename T>
__global__ void maxpool_kernel(const T* __restrict__ A, int64_t a_sstride, T* __restrict__ out, int64_t o_sstride,
                               int NB, int H, int Wd, int C, int Ho, int Wo, int ph, int pw, int sh, int sw) {
  const int s = blockIdx.y;
  const T* Ap = A + (int64_t)s * a_sstride;
  T* op = out + (int64_t)s * o_sstride;
  const int64_t total = (int64_t)NB * Ho * Wo * C;
  for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(o % C);
    int64_t r = o / C;
    const int wo = (int)(r % Wo); r /= Wo;
    const int ho = (int)(r % Ho);
    const int nb = (int)(r / Ho);
    float m = -INFINITY;
    for (int i = 0; i < ph; ++i)
      for (int j = 0; j < pw; ++j)
        m = fmaxf(m, to_f32<T>(Ap[(((int64_t)nb * H + ho * sh + i) * Wd + wo * sw + j) * C + c]));
    op[o] = from_f32<T>(m);
  }
}

// bf16, C % 8 == 0: one thread = 8 channels (16-byte loads / stores, packed bf16x2 max)
__global__ void maxpool_bf16_vec8_kernel(const __nv_bfloat16* __restrict__ A, int64_t a_sstride, __nv_bfloat16* __restrict__ out,
                                         int64_t o_sstride, int NB, int H, int Wd, int C, int Ho, int Wo, int ph, int pw, int sh,
                                         int sw) {
  const int s = blockIdx.y;
  const uint4* Ap = reinterpret_cast<const uint4*>(A + (int64_t)s * a_sstride);
  uint4* op = reinterpret_cast<uint4*>(out + (int64_t)s * o_sstride);
  const int C8 = C >> 3;
  const int64_t total = (int64_t)NB * Ho * Wo * C8;
  for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(o % C8);
    int64_t r = o / C8;
    const int wo = (int)(r % Wo); r /= Wo;
    const int ho = (int)(r % Ho);
    const int nb = (int)(r / Ho);
    __nv_bfloat162 m[4];
    bool first = true;
    for (int i = 0; i < ph; ++i)
      for (int j = 0; j < pw; ++j) {
        const uint4 v = __ldg(Ap + (((int64_t)nb * H + ho * sh + i) * Wd + wo * sw + j) * C8 + c8);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
        for (int q = 0; q < 4; ++q) m[q] = first ? h[q] : __hmax2(m[q], h[q]);
        first = false;
      }
    op[o] = *reinterpret_cast<const uint4*>(m);
  }
}

// Sample-ordered accumulation of the model output over the chunk's samples:
//   sum1[b,c] (+)= sum_s w_s * f_s[b,c],  sum2 likewise with f^2  (posteriormodel.py:95-100).
// logits [ns,B,C] fp32 are the last layer's pre-activation values.
__device__ __forceinline__ float accum_value(const float* __restrict__ z, int c, int C, int act, int out_kind) {
  if (out_kind == DBX_OUT_LOGITS) return z[c];
  if (act == DBX_ACT_SOFTMAX) {
    float m = z[0];
    for (int j = 1; j < C; ++j) m = fmaxf(m, z[j]);
    float den = 0.f;
    for (int j = 0; j < C; ++j) den += expf(z[j] - m);
    return expf(z[c] - m) / den;
  }
  return apply_act(act, z[c]);
}

__global__ void accum_kernel(const float* __restrict__ logits, int ns, int64_t B, int C, int act, int out_kind,
                             const float* __restrict__ wts, int overwrite, float* __restrict__ sum1,
                             float* __restrict__ sum2) {
  const int64_t total = B * C;
  for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = o / C;
    const int c = (int)(o - b * C);
    float a1 = overwrite ? 0.f : sum1[o];
    float a2 = (sum2 && !overwrite) ? sum2[o] : 0.f;
    for (int s = 0; s < ns; ++s) {
      const float p = accum_value(logits + ((int64_t)s * B + b) * C, c, C, act, out_kind);
      const float w = wts ? wts[s] : 1.f;
      if (wts) { a1 = fmaf(w, p, a1); a2 = fmaf(w, p * p, a2); }
      else { a1 += p; a2 = fmaf(p, p, a2); }
    }
    sum1[o] = a1;
    if (sum2) sum2[o] = a2;
  }
}

// One BLOCK per input row b (C <= 32): thread t takes the contiguous sample range [t*ns/256, (t+1)*ns/256), evaluates the
// last activation ONCE per (sample, row) for all C classes and keeps C running sums in registers; the 256 partial sums
// are combined by a fixed shuffle tree and a fixed-order pass over the 8 warps -- deterministic.  (The element-per-thread
// kernel above recomputes the softmax C times and, at B = 64, left a 256-sample chunk to 640 threads: 224 us.)
constexpr int kAccumMaxC = 32;
__global__ void __launch_bounds__(256)
accum_block_kernel(const float* __restrict__ logits, int ns, int64_t B, int C, int act, int out_kind,
                   const float* __restrict__ wts, int overwrite, float* __restrict__ sum1, float* __restrict__ sum2) {
  __shared__ float part[8][2][kAccumMaxC];
  const int64_t b = blockIdx.x;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int s_lo = (int)(((int64_t)t * ns) / 256), s_hi = (int)(((int64_t)(t + 1) * ns) / 256);
  float a1[kAccumMaxC], a2[kAccumMaxC];
#pragma unroll
  for (int c = 0; c < kAccumMaxC; ++c) { a1[c] = 0.f; a2[c] = 0.f; }
  for (int s = s_lo; s < s_hi; ++s) {
    const float* z = logits + ((int64_t)s * B + b) * C;
    float v[kAccumMaxC];
#pragma unroll
    for (int c = 0; c < kAccumMaxC; ++c) v[c] = (c < C) ? z[c] : -INFINITY;
    if (out_kind != DBX_OUT_LOGITS) {
      if (act == DBX_ACT_SOFTMAX) {
        float m = v[0];
#pragma unroll
        for (int c = 1; c < kAccumMaxC; ++c) m = fmaxf(m, v[c]);
        float den = 0.f;
#pragma unroll
        for (int c = 0; c < kAccumMaxC; ++c) { v[c] = (c < C) ? expf(v[c] - m) : 0.f; den += v[c]; }
#pragma unroll
        for (int c = 0; c < kAccumMaxC; ++c) v[c] = v[c] / den;
      } else {
#pragma unroll
        for (int c = 0; c < kAccumMaxC; ++c) v[c] = (c < C) ? apply_act(act, v[c]) : 0.f;
      }
    }
    const float w = wts ? wts[s] : 1.f;
#pragma unroll
    for (int c = 0; c < kAccumMaxC; ++c)
      if (c < C) { a1[c] = fmaf(w, v[c], a1[c]); a2[c] = fmaf(w * v[c], v[c], a2[c]); }
  }
#pragma unroll
  for (int c = 0; c < kAccumMaxC; ++c) {
    if (c < C) {
      float x1 = a1[c], x2 = a2[c];
#pragma unroll
      for (int d = 16; d >= 1; d >>= 1) { x1 += __shfl_down_sync(0xffffffffu, x1, d); x2 += __shfl_down_sync(0xffffffffu, x2, d); }
      if (lane == 0) { part[warp][0][c] = x1; part[warp][1][c] = x2; }
    }
  }
  __syncthreads();
  if (t < C) {
    const int64_t o = b * C + t;
    float t1 = overwrite ? 0.f : sum1[o];
    float t2 = (sum2 && !overwrite) ? sum2[o] : 0.f;
    for (int w8 = 0; w8 < 8; ++w8) { t1 += part[w8][0][t]; t2 += part[w8][1][t]; }
    sum1[o] = t1;
    if (sum2) sum2[o] = t2;
  }
}

static inline void launch_accum(const float* logits, int ns, int64_t B, int C, int act, int out_kind, const float* wts, int overwrite,
                                float* sum1, float* sum2, cudaStream_t st) {
  if (C <= kAccumMaxC && ns >= 32 && B <= 65535 * 16) {
    accum_block_kernel<<<(unsigned)B, 256, 0, st>>>(logits, ns, B, C, act, out_kind, wts, overwrite, sum1, sum2);
  } else {
    accum_kernel<<<(unsigned)std::min<int64_t>(std::max<int64_t>(cdiv(B * C, 256), 1), sm_count() * 16), 256, 0, st>>>(
        logits, ns, B, C, act, out_kind, wts, overwrite, sum1, sum2);
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
}

// flags[s,k] = (argmax_c logits[s,k,:] == labels[k]) (first maximum, like np.argmax).
__global__ void count_kernel(const float* __restrict__ logits, int ns, int64_t K, int C,
                             const int32_t* __restrict__ labels, uint8_t* __restrict__ flags,
                             unsigned long long* __restrict__ counts) {
  const int64_t total = (int64_t)ns * K;
  for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
    const int64_t k = o % K;
    const float* z = logits + o * C;
    int best = 0;
    float bv = z[0];
    for (int j = 1; j < C; ++j)
      if (z[j] > bv) { bv = z[j]; best = j; }
    const int ok = (best == labels[k]);
    if (flags) flags[o] = (uint8_t)ok;
    if (ok) atomicAdd(counts + k, 1ULL);
  }
}

__global__ void act_rows_kernel(int act, float* __restrict__ x, int64_t rows, int64_t cols) {
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
    float* z = x + r * cols;
    if (act == DBX_ACT_SOFTMAX) {
      float m = z[0];
      for (int64_t j = 1; j < cols; ++j) m = fmaxf(m, z[j]);
      float den = 0.f;
      for (int64_t j = 0; j < cols; ++j) den += expf(z[j] - m);
      for (int64_t j = 0; j < cols; ++j) z[j] = expf(z[j] - m) / den;
    } else {
      for (int64_t j = 0; j < cols; ++j) z[j] = apply_act(act, z[j]);
    }
  }
}

__global__ void finalize_kernel(const float* __restrict__ s1, const float* __restrict__ s2, float inv_n,
                                int64_t total, float* __restrict__ mean, float* __restrict__ var) {
  for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
    const float m = s1[o] * inv_n;
    mean[o] = m;
    if (var) var[o] = s2[o] * inv_n - m * m;
  }
}

__global__ void zero_u64_kernel(unsigned long long* p, int64_t n) {
  for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < n; o += (int64_t)gridDim.x * blockDim.x) p[o] = 0ULL;
}

static inline int grid_for(int64_t total, int block = 256) {
  int64_t g = cdiv(total, block);
  return (int)std::min<int64_t>(std::max<int64_t>(g, 1), sm_count() * 16);
}

// ---------------------------------------------------------------------------------------
// Generic executor (all layer kinds; T = storage type of weights/activations)
// ---------------------------------------------------------------------------------------

int launch_sample_bf16(dbx_model* m, const TensorTable& tab, const Source& src, int64_t c0, int nc, __nv_bfloat16* W16,
                       float* B32, cudaStream_t st, int64_t pb_stride, const bool* skip, int64_t w16_stride) {
  if (pb_stride <= 0) pb_stride = m->PB;
  if (w16_stride <= 0) w16_stride = m->P16;
  const int nblk = (int)((m->P + 3) / 4);
  BlockRanges rg; rg.n = 0;
  const bool philox = !src.stored && !src.eps;
  int cursor = 0;  // first Philox block not yet covered
  if (philox) {
    for (int i = 0; i < tab.n; ++i) {
      const bool fast = !tab.is_bias[i] && !tab.blocked8[i] && tab.cols_pad[i] == tab.cols[i] && tab.size[i] >= 65536 &&
                        ((tab.dst_off[i] - tab.src_off[i]) & 3) == 0;
      if (!fast) continue;
      const int lo = (tab.src_off[i] + 3) >> 2, hi = tab.src_end[i] >> 2;   // whole Philox blocks inside the tensor
      if (hi <= lo || rg.n + 2 > DBX_MAX_RANGES) continue;
      if (lo > cursor) { rg.lo[rg.n] = cursor; rg.hi[rg.n] = lo; ++rg.n; }
      dim3 grid((unsigned)std::min<int64_t>(cdiv((hi - lo + 1) / 2, 256), sm_count() * 4), (unsigned)cdiv(nc, kSamplesPerThread));
      DBX_LAUNCH(sample_bf16_fast_kernel, grid, 256, 0, st, m->mu, m->sigma, src.inflate, src.seed, src.s0 + c0, nc, lo, hi,
                 tab.dst_off[i] - tab.src_off[i], W16, w16_stride);
      cursor = hi;
    }
  }
  if (skip) {
    // tensors somebody else consumes in their original form (stored fp32 kernels read by dense_tc_f32w_kernel): leave
    // their interior Philox blocks out
    for (int i = 0; i < tab.n; ++i) {
      if (!skip[i]) continue;
      const int lo = (tab.src_off[i] + 3) >> 2, hi = tab.src_end[i] >> 2;
      if (hi <= lo || lo < cursor || rg.n + 2 > DBX_MAX_RANGES) continue;
      if (lo > cursor) { rg.lo[rg.n] = cursor; rg.hi[rg.n] = lo; ++rg.n; }
      cursor = hi;
    }
  }
  if (cursor < nblk) { rg.lo[rg.n] = cursor; rg.hi[rg.n] = nblk; ++rg.n; }
  if (rg.n > 0) {
    int64_t most = 0;
    for (int i = 0; i < rg.n; ++i) most = std::max<int64_t>(most, rg.hi[i] - rg.lo[i]);
    dim3 grid((unsigned)std::max<int64_t>(1, std::min<int64_t>(cdiv(most, 256), sm_count() * 4)), (unsigned)nc);
    DBX_LAUNCH(sample_bf16_kernel, grid, 256, 0, st, m->mu, m->sigma,
               (!src.stored && src.eps) ? src.eps + c0 * m->P : nullptr, src.inflate, src.seed, src.s0 + c0, (int)m->P,
               src.stored ? m->slab : nullptr, (src.stored && src.idx) ? src.idx + c0 : nullptr, src.j0 + c0, m->slab_stride, tab, rg, W16,
               w16_stride, B32, pb_stride);
  }
  return 0;
}

static size_t ws_budget_bytes() { return (size_t)std::max<long long>(1, opt().ws_mb) << 20; }

// Can dense_tc_f32w_kernel take this layer's kernel straight from the fp32 slab?  (16-byte TMA strides and offsets; small
// kernels are not worth a launch of their own.)  DBX_NO_F32W=1 restores the two-pass path.
static bool f32w_layer_ok(const dbx_model* m, int li) {
  if (opt().no_f32w || opt().no_tc) return false;
  const Layer& L = m->layers[li];
  if (L.kind != DBX_DENSE || !L.has_w || !m->slab) return false;
  if ((L.in_feat & 7) || (L.w_cols & 3) || (L.w_off & 3) || (m->slab_stride & 3) || ((uintptr_t)m->slab & 15)) return false;
  return (int64_t)L.w_rows * L.w_cols >= 65536 && (li == m->last_weighted || L.units % 8 == 0);
}
// ... and do those layers hold most of the model (then the tensor-core path streams the slab at B > 8 faster than the
// FFMA streaming kernel, which is compute-bound beyond 8 rows)
static bool f32w_model_ok(const dbx_model* m) {
  int64_t covered = 0;
  for (int li = 0; li < (int)m->layers.size() && li < 64; ++li)
    if (f32w_layer_ok(m, li)) covered += (int64_t)m->layers[li].w_rows * m->layers[li].w_cols;
  return covered * 10 >= m->P * 9;
}

// side stream + events of the "weights of the next chunk are drawn ahead" pipelines (generic executor, gradient loops)
static int ensure_side_streams(dbx_model* m) {
  if (!m->ibp_side) {
    DBX_CUDA(cudaStreamCreateWithFlags(&m->ibp_side, cudaStreamNonBlocking));
    for (cudaEvent_t& e : m->ibp_ev) DBX_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  return 0;
}

// high-priority stream of the GEMM chains: their CTAs are placed ahead of the sampler's blocks, which fill what is left
static int ensure_main_stream(dbx_model* m) {
  if (!m->ibp_main) {
    int lo = 0, hi = 0;
    DBX_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    DBX_CUDA(cudaStreamCreateWithPriority(&m->ibp_main, cudaStreamNonBlocking, hi));
  }
  return 0;
}

template <typename T>
static int run_generic(dbx_model* m, const float* x_dev, int64_t B, int64_t n, const Source& src, const Sink& sink,
                       cudaStream_t st, cudaStream_t sst = nullptr) {
  // sst != nullptr (bf16 mode): the weights of chunk c + 1 are drawn / converted on that side stream while chunk c is in the
  // layers (two weight buffers; events gen_ev[1..4] = sampled[2], consumed[2]); the caller has ordered both streams after
  // everything that was enqueued before the call and joins them afterwards (see run())
  constexpr bool kBf16 = !std::is_same<T, float>::value;
  const int64_t C = m->out_feat;
  const bool overlap = kBf16 && sst != nullptr;
  const int nwb = overlap ? 2 : 1;
  // Stored posteriors in bf16 mode: large Dense kernels are NOT copied to bf16 first -- dense_tc_f32w_kernel reads them
  // from the fp32 slab by TMA and rounds them inside the SM (gemm_tc.cu).  DBX_NO_F32W=1 restores the two-pass path.
  // (Batches above 384 rows: every 128-row tile converts the same fp32 block again, and one conversion pass to bf16 + the
  // CTA-pair GEMM wins -- 7.6 vs 11.2 ms at B = 1024, 4.9 vs 5.9 ms at 512, equal at 384, 3.6 vs 3.1 ms at 256 for 1000
  // samples of config 3.)
  bool skip_tensor[DBX_MAX_TENSORS] = {};
  bool layer_f32w[64] = {};
  bool any_f32w = false;
  if (kBf16 && !opt().no_tc && src.stored && m->layers.size() <= 64 && B <= 384) {
    for (int li = 0; li < (int)m->layers.size(); ++li) {
      if (!f32w_layer_ok(m, li)) continue;
      const Layer& L = m->layers[li];
      for (int ti = 0; ti < m->table.n; ++ti)
        if (!m->table.is_bias[ti] && m->table.src_off[ti] == (int32_t)L.w_off) { skip_tensor[ti] = true; layer_f32w[li] = true; any_f32w = true; }
    }
  }
  // ... and then the bf16 copy of a sample holds only the remaining (small) kernels: a compact layout, so that a chunk
  // can span thousands of samples (launch boundaries drain the HBM pipeline: 40 chunks x 3 launches cost config 3 ~8 %)
  TensorTable tabc = m->table;
  int64_t p16_eff = m->P16;
  int64_t w16_off_c[64];
  for (int li = 0; li < (int)m->layers.size() && li < 64; ++li) w16_off_c[li] = m->layers[li].w16_off;
  if (any_f32w) {
    int64_t off = 0;
    for (int ti = 0; ti < tabc.n; ++ti) {
      if (tabc.is_bias[ti]) continue;
      if (skip_tensor[ti]) continue;
      for (int li = 0; li < (int)m->layers.size() && li < 64; ++li)
        if (m->layers[li].has_w && m->table.src_off[ti] == (int32_t)m->layers[li].w_off) w16_off_c[li] = off;
      tabc.dst_off[ti] = (int32_t)off;
      off += round_up((int64_t)tabc.rows[ti] * tabc.cols_pad[ti], 64);
    }
    p16_eff = std::max<int64_t>(round_up(off, 64), 64);
  }
  // per-sample workspace bytes
  const size_t act_b = (size_t)B * m->max_feat * sizeof(T);
  const size_t w1_b = kBf16 ? ((size_t)p16_eff * 2 + (size_t)m->PB * 4) : (src.stored ? 0 : (size_t)m->P * 4);
  const size_t w_b = w1_b * nwb;
  // bf16 mode: Conv2D layers run as im2col + tcgen05 GEMM and need a column buffer
  size_t col_b = 0;
  bool use_tc = kBf16;
  if (opt().no_tc) use_tc = false;
  // (not the layers conv_tc_kernel takes directly; a conv on the call's input, which every sample shares, needs ONE
  // column matrix per call, built before the chunk loop)
  const bool direct_conv = !opt().no_direct_conv;
  size_t col0_b = 0;
  int first_w = -1;
  for (int li = 0; li < (int)m->layers.size(); ++li) if (m->layers[li].has_w) { first_w = li; break; }
  if (use_tc)
    for (int li = 0; li < (int)m->layers.size(); ++li) {
      const Layer& L = m->layers[li];
      if (L.kind != DBX_CONV2D || (direct_conv && li != m->last_weighted && conv_tc_eligible(L))) continue;
      const size_t b = round_up((size_t)B * L.out_h * L.out_w * round_up(L.w_rows, 8) * 2, 256);
      if (li == first_w) col0_b = b; else col_b = std::max(col_b, b);
    }
  const size_t per = 2 * act_b + w_b + (size_t)B * C * 4 + col_b + 256;
  // the bf16 copy of x exists kXRep times when stored posteriors go through dense_tc_f32w_kernel (see its launcher)
  constexpr int kXRep = 32;
  const size_t xin1_b = kBf16 ? round_up((size_t)B * m->in_feat * 2, 256) : 0;
  const size_t xin_b = xin1_b * ((kBf16 && src.stored) ? kXRep : 1) + col0_b;
  int64_t ns = (int64_t)std::max<size_t>(1, (ws_budget_bytes() - std::min(ws_budget_bytes(), xin_b)) / per);
  ns = std::min<int64_t>(std::min<int64_t>(ns, n), any_f32w ? 2048 : 256);
  if (ensure_ws(m->ws, xin_b + (size_t)ns * per + 16384)) return 1;
  char* p = (char*)m->ws.ptr;
  T* xin = (T*)p; p += xin_b - col0_b;
  __nv_bfloat16* col0 = (__nv_bfloat16*)p; p += col0_b;
  bool col0_done = false;
  T* act0 = (T*)p; p += round_up((size_t)ns * act_b, 256);
  T* act1 = (T*)p; p += round_up((size_t)ns * act_b, 256);
  float* logits = (float*)p; p += round_up((size_t)ns * B * C * 4, 256);
  T* Wc = (T*)p;
  float* B32 = kBf16 ? (float*)(p + round_up((size_t)ns * p16_eff * 2, 256)) : nullptr;
  T* Wc2[2] = {Wc, Wc}; float* B32b[2] = {B32, B32};
  if (overlap) {
    const size_t one = round_up((size_t)ns * p16_eff * 2, 256) + round_up((size_t)ns * m->PB * 4, 256);
    Wc2[1] = (T*)(p + one); B32b[1] = (float*)(p + one + round_up((size_t)ns * p16_eff * 2, 256));
  }
  __nv_bfloat16* colbuf = (__nv_bfloat16*)(p + round_up((size_t)ns * w_b, 256) + 1024);

  const T* x_in;
  if (kBf16) {
    if (src.stored)   // one launch writes all kXRep replicas (was 31 device-to-device copies per call)
      DBX_LAUNCH(cast_rep_kernel<T>, grid_for(B * m->in_feat), 256, 0, st, x_dev, xin, B * m->in_feat, kXRep, (int64_t)(xin1_b / sizeof(T)));
    else
      DBX_LAUNCH(cast_kernel<T>, grid_for(B * m->in_feat), 256, 0, st, x_dev, xin, B * m->in_feat);
    x_in = xin;
  } else x_in = (const T*)x_dev;

  if (sink.kind == 1 && !sink.accumulate)
    DBX_LAUNCH(zero_u64_kernel, grid_for(B), 256, 0, st, sink.counts, B);

  for (int64_t c0 = 0; c0 < n; c0 += ns) {
    const int nc = (int)std::min<int64_t>(ns, n - c0);
    // ---- weights of this chunk
    const T* Wbase; int64_t w_sstride; const int32_t* rows = nullptr; int64_t row0 = 0;
    const float* Bbase; int64_t b_sstride;
    const int64_t ci = c0 / ns;
    const int wb = (int)(ci & 1) % nwb;
    if (kBf16) {
      auto draw_chunk = [&](int64_t cj) -> int {
        const int b = (int)(cj & 1) % nwb;
        const int64_t d0 = cj * ns;
        const int dn = (int)std::min<int64_t>(ns, n - d0);
        cudaStream_t ds = overlap ? sst : st;
        if (overlap && cj >= 2) DBX_CUDA(cudaStreamWaitEvent(ds, m->ibp_ev[3 + b], 0));
        if (launch_sample_bf16(m, tabc, src, d0, dn, (__nv_bfloat16*)Wc2[b], B32b[b], ds, 0, any_f32w ? skip_tensor : nullptr, p16_eff)) return 1;
        if (overlap) DBX_CUDA(cudaEventRecord(m->ibp_ev[1 + b], ds));
        return 0;
      };
      if (ci == 0 || !overlap) { if (draw_chunk(ci)) return 1; }
      if (overlap) {
        if (c0 + ns < n && draw_chunk(ci + 1)) return 1;       // next chunk's weights, concurrently
        DBX_CUDA(cudaStreamWaitEvent(st, m->ibp_ev[1 + wb], 0));
      }
      Wc = Wc2[wb]; B32 = B32b[wb];
      Wbase = Wc; w_sstride = p16_eff; Bbase = B32; b_sstride = m->PB;
    } else if (src.stored) {
      Wbase = (const T*)m->slab; w_sstride = m->slab_stride; rows = src.idx ? src.idx + c0 : nullptr; row0 = src.j0 + c0;
      Bbase = m->slab; b_sstride = m->slab_stride;
    } else {
      const int64_t total = (int64_t)nc * ((m->P + 3) / 4);
      DBX_LAUNCH(sample_f32_kernel, grid_for(total), 256, 0, st, m->mu, m->sigma,
                 src.eps ? src.eps + c0 * m->P : nullptr, src.inflate, src.seed, src.s0 + c0, (int64_t)nc, m->P,
                 src.unfused ? 1 : 0, (float*)Wc, (float*)nullptr);
      Wbase = Wc; w_sstride = m->P; Bbase = (const float*)Wc; b_sstride = m->P;
    }
    // ---- layers
    const T* cur = x_in; int64_t cur_ss = 0;  // layer-0 input is shared by all samples
    T* bufs[2] = {act0, act1}; int bi = 0;
    bool pooled_next = false;
    for (int li = 0; li < (int)m->layers.size(); ++li) {
      const Layer& L = m->layers[li];
      if (L.kind == DBX_FLATTEN) continue;
      if (pooled_next) { pooled_next = false; continue; }      // this MaxPool ran in the previous convolution's epilogue
      const bool last = (li == m->last_weighted);
      const int64_t woff = kBf16 ? w16_off_c[std::min(li, 63)] : L.w_off;
      const int64_t boff = kBf16 ? L.b32_off : L.b_off;
      const int ldw = (int)(kBf16 ? L.w_cols_pad : L.w_cols);
      const int act = last ? DBX_ACT_LINEAR : L.act;  // last activation applied by the sink
      if (L.has_w) prof_begin(st);
      bool done_tc = false;
      if (kBf16 && use_tc && L.has_w && (last || L.units % 8 == 0)) {
        // tensor-core path: per-sample GEMM on tcgen05 (Dense directly, Conv2D through im2col)
        const __nv_bfloat16* A16 = (const __nv_bfloat16*)cur;
        const int shared = (cur_ss == 0) ? 1 : 0;
        if (L.kind == DBX_DENSE && layer_f32w[li]) {
          if (dense_tc_f32w_launch(A16, shared ? (long long)(xin1_b / 2) : cur_ss, (int)L.in_feat, shared ? kXRep : 0, m->slab, m->slab_stride, m->S, L.w_off,
                                   src.idx ? src.idx + c0 : nullptr, src.j0 + c0, nc, (int)B, L.units, (int)L.in_feat, Bbase, b_sstride,
                                   boff, act, last ? nullptr : (__nv_bfloat16*)bufs[bi], B * L.out_feat, L.units,
                                   last ? logits : nullptr, B * C, st)) {
            set_err("dense_tc_f32w launch failed"); return 1;
          }
          done_tc = true;
        } else if (L.kind == DBX_DENSE && L.in_feat % 8 == 0) {
          done_tc = dense_tc_launch(A16, cur_ss, (int)L.in_feat, shared, (const __nv_bfloat16*)Wbase + woff, w_sstride, ldw, nc,
                                    (int)B, L.units, (int)L.in_feat, Bbase, b_sstride, boff, act,
                                    last ? nullptr : (__nv_bfloat16*)bufs[bi], B * L.out_feat, L.units, last ? logits : nullptr,
                                    B * C, st) == 0;
        } else if (L.kind == DBX_CONV2D && !last && !opt().no_direct_conv &&
                   li + 1 < (int)m->layers.size() && conv_tc_pool_eligible(L, m->layers[li + 1]) && (B * m->layers[li + 1].out_feat) % 8 == 0 &&
                   conv_tc_launch(A16, cur_ss, shared, (const __nv_bfloat16*)Wbase + woff, w_sstride, ldw, nc, (int)B, L, Bbase, b_sstride,
                                  boff, act, (__nv_bfloat16*)bufs[bi], B * m->layers[li + 1].out_feat, st, 1) == 0) {
          done_tc = true;      // direct convolution with the following 2x2 max-pool done in its epilogue: the full-resolution map
          pooled_next = true;  // (4x the bytes) is never written
        } else if (L.kind == DBX_CONV2D && !last && !opt().no_direct_conv &&
                   conv_tc_launch(A16, cur_ss, shared, (const __nv_bfloat16*)Wbase + woff, w_sstride, ldw, nc, (int)B, L, Bbase, b_sstride,
                                  boff, act, (__nv_bfloat16*)bufs[bi], B * L.out_feat, st) == 0) {
          done_tc = true;      // direct convolution: per-tap TMA boxes of the NHWC input, no im2col matrix
        } else if (L.kind == DBX_CONV2D) {
          const int Kp = (int)round_up(L.w_rows, 8);
          const long long Mrows = (long long)B * L.out_h * L.out_w;
          const long long col_ss = shared ? 0 : Mrows * Kp;
          __nv_bfloat16* cb = (shared && li == first_w && col0_b) ? col0 : colbuf;
          bool col_ok = true;
          if (!(cb == col0 && col0_done)) col_ok = im2col_launch(A16, cur_ss, cb, col_ss, shared ? 1 : nc, (int)B, L, Kp, st) == 0;
          if (cb == col0 && col_ok) col0_done = true;      // the input is the same for every chunk of this call
          // the call's own input under eight samples' kernels at a time (config 4's first convolution)
          if (col_ok && shared && !last &&
              dense_tc_shareda_launch(cb, Kp, (const __nv_bfloat16*)Wbase + woff, w_sstride, ldw, nc, (int)Mrows, L.units, (int)L.w_rows,
                                      Bbase, b_sstride, boff, act, (__nv_bfloat16*)bufs[bi], B * L.out_feat, L.units, st) == 0)
            done_tc = true;
          else if (col_ok)
            done_tc = dense_tc_launch(cb, col_ss, Kp, shared, (const __nv_bfloat16*)Wbase + woff, w_sstride, ldw, nc,
                                      (int)Mrows, L.units, (int)L.w_rows, Bbase, b_sstride, boff, act,
                                      last ? nullptr : (__nv_bfloat16*)bufs[bi], B * L.out_feat, L.units,
                                      last ? logits : nullptr, B * C, st) == 0;
        }
        if (done_tc) g_used_tc = true;
      }
      if (done_tc) {
      } else if (L.kind == DBX_DENSE) {
        dim3 grid((unsigned)cdiv(L.units, 64), (unsigned)cdiv(B, 64), (unsigned)nc);
        if (last) {
          DBX_LAUNCH((dense_kernel<T, float>), grid, 256, 0, st, cur, cur_ss, Wbase, w_sstride, rows, row0, woff, ldw,
                     Bbase, b_sstride, boff, logits, B * C, (int)B, L.units, (int)L.in_feat, act);
        } else {
          DBX_LAUNCH((dense_kernel<T, T>), grid, 256, 0, st, cur, cur_ss, Wbase, w_sstride, rows, row0, woff, ldw,
                     Bbase, b_sstride, boff, bufs[bi], B * L.out_feat, (int)B, L.units, (int)L.in_feat, act);
        }
      } else if (L.kind == DBX_CONV2D) {
        dim3 grid((unsigned)grid_for(B * L.out_feat), (unsigned)nc);
        if (last) {
          DBX_LAUNCH((conv2d_kernel<T, float>), grid, 256, 0, st, cur, cur_ss, Wbase, w_sstride, rows, row0, woff, ldw,
                     Bbase, b_sstride, boff, logits, B * C, (int)B, L.in_h, L.in_w, L.in_c, L.out_h, L.out_w, L.out_c,
                     L.kh, L.kw, L.sh, L.sw, L.pad_t, L.pad_l, act);
        } else {
          DBX_LAUNCH((conv2d_kernel<T, T>), grid, 256, 0, st, cur, cur_ss, Wbase, w_sstride, rows, row0, woff, ldw,
                     Bbase, b_sstride, boff, bufs[bi], B * L.out_feat, (int)B, L.in_h, L.in_w, L.in_c, L.out_h, L.out_w,
                     L.out_c, L.kh, L.kw, L.sh, L.sw, L.pad_t, L.pad_l, act);
        }
      } else if (L.kind == DBX_MAXPOOL2D) {
        if (kBf16 && L.in_c % 8 == 0 && cur_ss % 8 == 0 && (B * L.out_feat) % 8 == 0) {
          dim3 grid((unsigned)grid_for(B * L.out_feat / 8), (unsigned)nc);
          DBX_LAUNCH(maxpool_bf16_vec8_kernel, grid, 256, 0, st, (const __nv_bfloat16*)cur, cur_ss, (__nv_bfloat16*)bufs[bi],
                     B * L.out_feat, (int)B, L.in_h, L.in_w, L.in_c, L.out_h, L.out_w, L.kh, L.kw, L.sh, L.sw);
        } else {
          dim3 grid((unsigned)grid_for(B * L.out_feat), (unsigned)nc);
          DBX_LAUNCH(maxpool_kernel<T>, grid, 256, 0, st, cur, cur_ss, bufs[bi], B * L.out_feat, (int)B, L.in_h, L.in_w,
                     L.in_c, L.out_h, L.out_w, L.kh, L.kw, L.sh, L.sw);
        }
      }
      if (L.has_w)
        prof_end(st, 2.0 * (double)L.w_rows * L.w_cols * (L.kind == DBX_CONV2D ? (double)L.out_h * L.out_w : 1.0) * (double)B * nc);
      if (last) break;
      cur = bufs[bi]; cur_ss = B * (pooled_next ? m->layers[li + 1].out_feat : L.out_feat); bi ^= 1;
    }
    if (overlap) DBX_CUDA(cudaEventRecord(m->ibp_ev[3 + wb], st));      // this chunk's weights are free again
    // ---- sink
    const int act_last = m->layers[m->last_weighted].act;
    if (sink.kind == 0) {
      const int overwrite = (c0 == 0 && !sink.accumulate) ? 1 : 0;
      launch_accum(logits, nc, B, (int)C, act_last, sink.out_kind, sink.wts ? sink.wts + c0 : nullptr, overwrite, sink.sum1, sink.sum2, st);
      DBX_CUDA(cudaGetLastError());
    } else if (sink.kind == 2) {   // every sample's own output (what the reference's loops look at one sample at a time)
      float* dst = sink.sum1 + c0 * B * C;
      DBX_CUDA(cudaMemcpyAsync(dst, logits, (size_t)nc * B * C * 4, cudaMemcpyDeviceToDevice, st));
      if (sink.out_kind == DBX_OUT_PROBS && act_last != DBX_ACT_LINEAR)
        DBX_LAUNCH(act_rows_kernel, grid_for((int64_t)nc * B), 256, 0, st, act_last, dst, (int64_t)nc * B, C);
    } else {
      DBX_LAUNCH(count_kernel, grid_for((int64_t)nc * B), 256, 0, st, logits, nc, B, (int)C, sink.labels,
                 sink.flags ? sink.flags + c0 * B : nullptr, sink.counts);
    }
  }
  return 0;
}

}  // namespace dbx
#include "ibp.cuh"
#include "train.cuh"
#include "attack.cuh"
namespace dbx {

static int run(dbx_model* m, const float* x_dev, int64_t B, int64_t n, const Source& src, const Sink& sink, int mode,
               cudaStream_t st) {
  DBX_CHECK(B > 0 && n > 0, "B and n must be positive (B=%lld n=%lld)", (long long)B, (long long)n);
  DBX_CHECK(src.stored ? (m->slab != nullptr) : m->have_gaussian,
            src.stored ? "no stored-sample slab set (dbx_set_samples)" : "no Gaussian posterior set (dbx_set_gaussian)");
  DBX_CUDA(cudaSetDevice(m->device));
  g_last_path = 0; g_used_tc = false;
  DBX_CHECK(mode == DBX_MODE_BF16 || mode == DBX_MODE_FP32, "unknown mode %d", mode);
  if (sink.kind == 2) {
    if (mode == DBX_MODE_FP32) return run_generic<float>(m, x_dev, B, n, src, sink, st);
    const int rg2 = run_generic<__nv_bfloat16>(m, x_dev, B, n, src, sink, st);
    if (g_used_tc) g_last_path = 3;
    return rg2;
  }
  if (!(mode == DBX_MODE_BF16 && src.stored && B > 8 && f32w_model_ok(m))) {
    const int r = stream_mlp_run(m, x_dev, B, n, src, sink, st);
    if (r < 0) return 1;
    if (r == 1) { g_last_path = 2; return 0; }
  }
  if (mode == DBX_MODE_BF16) {
    if (!opt().no_fused) {
      const int r = fused_mlp_run(m, x_dev, B, n, src, sink, st);
      if (r < 0) return 1;
      if (r == 1) { g_last_path = 1; return 0; }
    }
    // generic executor: the layers on a high-priority stream of the engine's own, the sampling / conversion of the next chunk
    // on a default-priority side stream, both joined with the caller's stream here (DBX_GENERIC_OVERLAP=0: all on `st`)
    const bool overlap = opt().generic_overlap != 0;
    if (!overlap) {
      const int rg0 = run_generic<__nv_bfloat16>(m, x_dev, B, n, src, sink, st);
      if (g_used_tc) g_last_path = 3;
      return rg0;
    }
    if (ensure_side_streams(m)) return 1;
    if (ensure_main_stream(m)) return 1;
    DBX_CUDA(cudaEventRecord(m->ibp_ev[0], st));
    DBX_CUDA(cudaStreamWaitEvent(m->ibp_side, m->ibp_ev[0], 0));
    DBX_CUDA(cudaStreamWaitEvent(m->ibp_main, m->ibp_ev[0], 0));
    const int rg = run_generic<__nv_bfloat16>(m, x_dev, B, n, src, sink, m->ibp_main, m->ibp_side);
    // join: the side stream's last work was consumed by the main stream already, but an error path may have left some
    DBX_CUDA(cudaEventRecord(m->ibp_ev[5], m->ibp_main));
    DBX_CUDA(cudaStreamWaitEvent(st, m->ibp_ev[5], 0));
    DBX_CUDA(cudaEventRecord(m->ibp_ev[0], m->ibp_side));
    DBX_CUDA(cudaStreamWaitEvent(st, m->ibp_ev[0], 0));
    if (g_used_tc) g_last_path = 3;
    return rg;
  }
  DBX_CHECK(mode == DBX_MODE_FP32, "unknown mode %d", mode);
  return run_generic<float>(m, x_dev, B, n, src, sink, st);
}

}  // namespace dbx

// ---------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------
using namespace dbx;

extern "C" {

const char* dbx_last_error(void) { return g_err.c_str(); }
int dbx_version(void) { return 100; }
int64_t dbx_launch_count(void) { return (int64_t)g_launches.load(); }
int dbx_last_path(void) { return g_last_path; }

int dbx_set_option(const char* name, int64_t value) {
  DBX_CHECK(name != nullptr, "dbx_set_option: null name");
  for (const OptEntry& e : kOptTable)
    if (strcmp(e.name, name) == 0) { opt().*(e.field) = (long long)value; return 0; }
  return set_err("dbx_set_option: unknown option '%s'", name);
}
int dbx_get_option(const char* name, int64_t* value) {
  DBX_CHECK(name != nullptr && value != nullptr, "dbx_get_option: null argument");
  for (const OptEntry& e : kOptTable)
    if (strcmp(e.name, name) == 0) { *value = (int64_t)(opt().*(e.field)); return 0; }
  return set_err("dbx_get_option: unknown option '%s'", name);
}
int dbx_reset_options(void) { options_from_env(opt()); return 0; }

int dbx_profile(int enable) {
  g_prof_on = enable != 0;
  if (!g_prof_on) {
    for (auto& r : g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    g_prof.clear();
  }
  return 0;
}

int dbx_profile_read(double* ms_total, double* flops_total, int64_t* launches) {
  double ms = 0, fl = 0;
  for (auto& r : g_prof) {
    DBX_CUDA(cudaEventSynchronize(r.b));
    float t = 0;
    DBX_CUDA(cudaEventElapsedTime(&t, r.a, r.b));
    ms += t; fl += r.flops;
    cudaEventDestroy(r.a); cudaEventDestroy(r.b);
  }
  if (ms_total) *ms_total = ms;
  if (flops_total) *flops_total = fl;
  if (launches) *launches = (int64_t)g_prof.size();
  g_prof.clear();
  return 0;
}

int dbx_create(const dbx_layer_desc* layers, int32_t n_layers, const int32_t* input_shape, int32_t input_rank,
               int32_t device, dbx_handle* out) {
  DBX_CHECK(layers && n_layers > 0 && input_shape && out, "dbx_create: null argument");
  DBX_CHECK(input_rank == 1 || input_rank == 3, "input_rank must be 1 ({K}) or 3 ({H,W,C}), got %d", input_rank);
  int ndev = 0;
  DBX_CUDA(cudaGetDeviceCount(&ndev));
  DBX_CHECK(device >= 0 && device < ndev, "device %d not available (%d CUDA devices)", device, ndev);
  DBX_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  DBX_CUDA(cudaGetDeviceProperties(&prop, device));
  DBX_CHECK(prop.major == 10, "deepbayes_b200 is built for sm_100a only; device %d is sm_%d%d", device, prop.major, prop.minor);

  dbx_model* m = new dbx_model();
  m->device = device;
  int h = 0, w = 0, c = 0;
  int64_t feat = 0;
  bool spatial = (input_rank == 3);
  if (spatial) { h = input_shape[0]; w = input_shape[1]; c = input_shape[2]; feat = (int64_t)h * w * c; }
  else feat = input_shape[0];
  m->in_feat = feat; m->max_feat = feat;
  int64_t off = 0, off16 = 0, offb = 0;
  memset(&m->table, 0, sizeof m->table);
  m->table.n = 0; m->is_mlp = true;
  for (int i = 0; i < n_layers; ++i) {
    const dbx_layer_desc& d = layers[i];
    Layer L;
    L.kind = d.kind; L.act = d.activation; L.units = d.units; L.kh = d.kh; L.kw = d.kw;
    L.sh = d.sh > 0 ? d.sh : 1; L.sw = d.sw > 0 ? d.sw : 1; L.pad = d.padding;
    L.in_feat = feat; L.in_h = h; L.in_w = w; L.in_c = c;
    if (d.kind == DBX_DENSE) {
      if (spatial) { delete m; return set_err("layer %d: Dense after a spatial layer needs a Flatten first", i); }
      L.has_w = true; L.w_rows = feat; L.w_cols = d.units; L.out_feat = d.units;
    } else if (d.kind == DBX_CONV2D) {
      if (!spatial) { delete m; return set_err("layer %d: Conv2D needs an {H,W,C} input", i); }
      m->is_mlp = false;
      L.has_w = true; L.w_rows = (int64_t)d.kh * d.kw * c; L.w_cols = d.units;
      if (d.padding == DBX_PAD_SAME) {
        L.out_h = (h + L.sh - 1) / L.sh; L.out_w = (w + L.sw - 1) / L.sw;
        const int ph = std::max((L.out_h - 1) * L.sh + d.kh - h, 0), pw = std::max((L.out_w - 1) * L.sw + d.kw - w, 0);
        L.pad_t = ph / 2; L.pad_l = pw / 2;
      } else { L.out_h = (h - d.kh) / L.sh + 1; L.out_w = (w - d.kw) / L.sw + 1; }
      L.out_c = d.units; L.out_feat = (int64_t)L.out_h * L.out_w * L.out_c;
      h = L.out_h; w = L.out_w; c = L.out_c;
    } else if (d.kind == DBX_MAXPOOL2D) {
      if (!spatial) { delete m; return set_err("layer %d: MaxPool2D needs an {H,W,C} input", i); }
      m->is_mlp = false;
      L.out_h = (h - d.kh) / L.sh + 1; L.out_w = (w - d.kw) / L.sw + 1; L.out_c = c;
      L.out_feat = (int64_t)L.out_h * L.out_w * c;
      h = L.out_h; w = L.out_w;
    } else if (d.kind == DBX_FLATTEN) {
      L.out_feat = feat; spatial = false; h = w = c = 0;
    } else { delete m; return set_err("layer %d: unknown kind %d", i, d.kind); }
    if (L.out_feat <= 0) { delete m; return set_err("layer %d: empty output", i); }
    if (L.has_w) {
      if (m->table.n + 2 > DBX_MAX_TENSORS) { delete m; return set_err("too many weight tensors"); }
      L.w_off = off; off += L.w_rows * L.w_cols;
      L.b_off = off; off += L.w_cols;
      L.w_cols_pad = round_up(L.w_cols, 8);
      L.w16_off = off16; off16 += round_up(L.w_rows * L.w_cols_pad, 64);
      L.b32_off = offb; offb += round_up(L.w_cols, 4);
      TensorTable& T = m->table;
      T.src_off[T.n] = (int32_t)L.w_off; T.size[T.n] = (int32_t)(L.w_rows * L.w_cols); T.dst_off[T.n] = (int32_t)L.w16_off;
      T.src_end[T.n] = T.src_off[T.n] + T.size[T.n];
      T.cols[T.n] = (int)L.w_cols; T.cols_pad[T.n] = (int)L.w_cols_pad; T.rows[T.n] = (int)L.w_rows;
      T.is_bias[T.n] = 0; T.blocked8[T.n] = 0; ++T.n;
      T.src_off[T.n] = (int32_t)L.b_off; T.size[T.n] = (int32_t)L.w_cols; T.dst_off[T.n] = (int32_t)L.b32_off;
      T.src_end[T.n] = T.src_off[T.n] + T.size[T.n];
      T.cols[T.n] = (int)L.w_cols; T.cols_pad[T.n] = (int)L.w_cols; T.rows[T.n] = 1;
      T.is_bias[T.n] = 1; T.blocked8[T.n] = 0; ++T.n;
      m->flops += 2 * L.w_rows * L.w_cols * (L.kind == DBX_CONV2D ? (int64_t)L.out_h * L.out_w : 1);
      m->last_weighted = i;
    }
    feat = L.out_feat;
    m->max_feat = std::max(m->max_feat, feat);
    m->layers.push_back(L);
  }
  if (m->last_weighted < 0) { delete m; return set_err("model has no weighted layer"); }
  for (int i = m->last_weighted + 1; i < n_layers; ++i)
    if (m->layers[i].kind != DBX_FLATTEN) { delete m; return set_err("layers after the last weighted layer are not supported"); }
  if (off >= (1ll << 31) - 8 || off16 >= (1ll << 31) - 8) { delete m; return set_err("model too large (> 2^31 parameters)"); }
  m->P = off; m->P16 = off16; m->PB = offb; m->out_feat = m->layers[m->last_weighted].out_feat;
  if (cudaMalloc(&m->mu, (size_t)m->P * 4) != cudaSuccess || cudaMalloc(&m->sigma, (size_t)m->P * 4) != cudaSuccess) {
    delete m; return set_err("cudaMalloc of posterior buffers failed");
  }
  *out = m;
  return 0;
}

int dbx_destroy(dbx_handle h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  fused_mlp_free(h);
  cudaFree(h->mu); cudaFree(h->sigma); cudaFree(h->ws.ptr); cudaFree(h->fused_ws.ptr);
  if (h->h_pin) cudaFreeHost(h->h_pin);
  for (int i = 0; i < 2; ++i) { if (h->up_stage[i]) cudaFreeHost(h->up_stage[i]); if (h->up_ev[i]) cudaEventDestroy(h->up_ev[i]); }
  cudaFree(h->d_x); cudaFree(h->d_s1); cudaFree(h->d_s2);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  if (h->ev_posterior) cudaEventDestroy(h->ev_posterior);
  if (h->ibp_side) cudaStreamDestroy(h->ibp_side);
  if (h->ibp_main) cudaStreamDestroy(h->ibp_main);
  for (cudaEvent_t e : h->ibp_ev) if (e) cudaEventDestroy(e);
  delete h;
  return 0;
}

int dbx_param_count(dbx_handle h, int64_t* P) { DBX_CHECK(h && P, "null argument"); *P = h->P; return 0; }
int dbx_output_dim(dbx_handle h, int64_t* C) { DBX_CHECK(h && C, "null argument"); *C = h->out_feat; return 0; }
int dbx_input_dim(dbx_handle h, int64_t* K) { DBX_CHECK(h && K, "null argument"); *K = h->in_feat; return 0; }
int dbx_flops_per_forward(dbx_handle h, int64_t* f) { DBX_CHECK(h && f, "null argument"); *f = h->flops; return 0; }

int dbx_set_gaussian(dbx_handle h, const float* mu, const float* sigma, int64_t P, int32_t on_device, void* stream) {
  DBX_CHECK(h && mu && sigma, "null argument");
  DBX_CHECK(P == h->P, "parameter count mismatch: model has %lld, got %lld", (long long)h->P, (long long)P);
  DBX_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  const cudaMemcpyKind k = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  DBX_CUDA(cudaMemcpyAsync(h->mu, mu, (size_t)P * 4, k, st));
  DBX_CUDA(cudaMemcpyAsync(h->sigma, sigma, (size_t)P * 4, k, st));
  if (!on_device) DBX_CUDA(cudaStreamSynchronize(st));
  if (!h->ev_posterior) DBX_CUDA(cudaEventCreateWithFlags(&h->ev_posterior, cudaEventDisableTiming));
  DBX_CUDA(cudaEventRecord(h->ev_posterior, st));
  h->have_gaussian = true;
  return 0;
}

int dbx_set_gaussian_rho(dbx_handle h, const float* mu, const float* rho, int64_t P, int32_t on_device, void* stream) {
  if (dbx_set_gaussian(h, mu, rho, P, on_device, stream)) return 1;
  cudaStream_t st = (cudaStream_t)stream;
  DBX_LAUNCH(softplus_kernel, grid_for(P), 256, 0, st, h->sigma, h->sigma, P);
  DBX_CUDA(cudaEventRecord(h->ev_posterior, st));
  return 0;
}

int dbx_get_gaussian(dbx_handle h, float* mu_host, float* sigma_host, int64_t P, void* stream) {
  DBX_CHECK(h && h->have_gaussian, "no Gaussian posterior set");
  DBX_CHECK(P == h->P, "parameter count mismatch");
  DBX_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  if (mu_host) DBX_CUDA(cudaMemcpyAsync(mu_host, h->mu, (size_t)P * 4, cudaMemcpyDeviceToHost, st));
  if (sigma_host) DBX_CUDA(cudaMemcpyAsync(sigma_host, h->sigma, (size_t)P * 4, cudaMemcpyDeviceToHost, st));
  DBX_CUDA(cudaStreamSynchronize(st));
  return 0;
}

int dbx_set_samples(dbx_handle h, const float* slab_dev, int64_t S, int64_t P) {
  DBX_CHECK(h && slab_dev && S > 0, "null/empty slab");
  DBX_CHECK(P == h->P, "parameter count mismatch: model has %lld, slab rows have %lld", (long long)h->P, (long long)P);
  h->slab = slab_dev; h->S = S; h->slab_stride = P;
  return 0;
}

int dbx_set_samples_strided(dbx_handle h, const float* slab_dev, int64_t S, int64_t P, int64_t row_stride) {
  if (dbx_set_samples(h, slab_dev, S, P)) return 1;
  DBX_CHECK(row_stride >= P, "row_stride %lld < P %lld", (long long)row_stride, (long long)P);
  h->slab_stride = row_stride;
  return 0;
}

int dbx_sample(dbx_handle h, uint64_t seed, int64_t s0, int64_t n, const float* eps_dev, float inflate,
               float* W_out_dev, float* eps_out_dev, void* stream) {
  DBX_CHECK(h && W_out_dev && n > 0, "null argument");
  DBX_CHECK(h->have_gaussian, "no Gaussian posterior set (dbx_set_gaussian)");
  DBX_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  DBX_LAUNCH(sample_f32_kernel, grid_for(n * ((h->P + 3) / 4)), 256, 0, st, h->mu, h->sigma, eps_dev, inflate, seed, s0,
             n, h->P, 0, W_out_dev, eps_out_dev);
  return 0;
}

int dbx_predict(dbx_handle h, const float* x_dev, int64_t B, int64_t n, uint64_t seed, int64_t s0, const float* eps_dev,
                float inflate, int32_t mode, int32_t out_kind, int32_t accumulate, float* sum1_dev, float* sum2_dev,
                void* stream) {
  DBX_CHECK(h && x_dev && sum1_dev, "null argument");
  Source src; src.eps = eps_dev; src.inflate = inflate; src.seed = seed; src.s0 = s0;
  Sink sink; sink.kind = 0; sink.out_kind = out_kind; sink.accumulate = accumulate; sink.sum1 = sum1_dev; sink.sum2 = sum2_dev;
  return run(h, x_dev, B, n, src, sink, mode, (cudaStream_t)stream);
}

int dbx_predict_stored(dbx_handle h, const float* x_dev, int64_t B, int64_t j0, int64_t n, const int32_t* idx_dev,
                       const float* wts_dev, int32_t mode, int32_t out_kind, int32_t accumulate, float* sum1_dev,
                       float* sum2_dev, void* stream) {
  DBX_CHECK(h && x_dev && sum1_dev, "null argument");
  DBX_CHECK(idx_dev || (j0 >= 0 && j0 + n <= h->S), "stored-sample range [%lld,%lld) outside [0,%lld)", (long long)j0,
            (long long)(j0 + n), (long long)h->S);
  Source src; src.stored = true; src.j0 = j0; src.idx = idx_dev;
  Sink sink; sink.kind = 0; sink.out_kind = out_kind; sink.accumulate = accumulate; sink.sum1 = sum1_dev; sink.sum2 = sum2_dev;
  sink.wts = wts_dev;
  return run(h, x_dev, B, n, src, sink, mode, (cudaStream_t)stream);
}

int dbx_forward_one(dbx_handle h, const float* x_dev, int64_t B, const float* W_dev, int32_t mode, int32_t out_kind,
                    float* out_dev, void* stream) {
  DBX_CHECK(h && x_dev && W_dev && out_dev, "null argument");
  // a single explicit weight vector is a stored-sample posterior with S = 1
  const float* keep = h->slab; const int64_t keepS = h->S, keepStride = h->slab_stride;
  h->slab = W_dev; h->S = 1; h->slab_stride = h->P;
  Source src; src.stored = true; src.j0 = 0;
  Sink sink; sink.kind = 0; sink.out_kind = out_kind; sink.sum1 = out_dev;
  const int r = run(h, x_dev, B, 1, src, sink, mode, (cudaStream_t)stream);
  h->slab = keep; h->S = keepS; h->slab_stride = keepStride;
  return r;
}

int dbx_bbb_forward(dbx_handle h, const float* x_dev, int64_t B, uint64_t seed, int64_t step, const float* eps_dev,
                    int32_t mode, float* out_dev, float* W_out_dev, float* eps_out_dev, void* stream) {
  DBX_CHECK(h && x_dev && out_dev, "null argument");
  DBX_CHECK(h->have_gaussian, "no posterior set (dbx_set_gaussian_rho)");
  DBX_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  if (W_out_dev || eps_out_dev) {
    // the caller wants the sampled weights / noise_used: materialise once, forward with them
    float* Wtmp = W_out_dev;
    Workspace tmp;
    if (!Wtmp) { if (ensure_ws(tmp, (size_t)h->P * 4)) return 1; Wtmp = (float*)tmp.ptr; }
    DBX_LAUNCH(sample_f32_kernel, grid_for((h->P + 3) / 4), 256, 0, st, h->mu, h->sigma, eps_dev, 1.0f, seed, step,
               (int64_t)1, h->P, 1, Wtmp, eps_out_dev);
    const int r = dbx_forward_one(h, x_dev, B, Wtmp, mode, DBX_OUT_PROBS, out_dev, stream);
    if (tmp.ptr) { cudaStreamSynchronize(st); cudaFree(tmp.ptr); }
    return r;
  }
  Source src; src.eps = eps_dev; src.inflate = 1.f; src.seed = seed; src.s0 = step; src.unfused = true;
  Sink sink; sink.kind = 0; sink.out_kind = DBX_OUT_PROBS; sink.sum1 = out_dev;
  return run(h, x_dev, B, 1, src, sink, mode, st);
}

int dbx_bbb_step(dbx_handle h, const float* x_dev, int64_t B, const int32_t* labels_dev, float* mean_dev, float* rho_dev,
                 const float* prior_mean_dev, const float* prior_var_dev, const float* noise_dev, uint64_t seed, int64_t step,
                 float lrate, float kl_weight, int32_t robust_mode, float epsilon, float robust_lambda, float in_lower, float in_upper,
                 int32_t mode, float* loss_out_dev, float* probs_out_dev, float* W_out_dev, void* stream) {
  DBX_CHECK(h && x_dev && labels_dev && mean_dev && rho_dev && prior_mean_dev && prior_var_dev, "null argument");
  DBX_CHECK(B > 0, "B must be positive");
  DBX_CHECK(mode == DBX_MODE_BF16 || mode == DBX_MODE_FP32, "unknown mode %d", mode);
  DBX_CHECK(robust_mode >= 0 && robust_mode <= 2, "robust_mode %d: 0, 1, 2 here; the epsilon-draw mode 3 is dbx_bbb_step_ibp_mc", robust_mode);
  DBX_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  if (run_bbb_step(h, x_dev, B, labels_dev, mean_dev, rho_dev, prior_mean_dev, prior_var_dev, noise_dev, seed, step, lrate, kl_weight,
                   loss_out_dev, probs_out_dev, robust_mode, epsilon, robust_lambda, in_lower, in_upper, st, 0, nullptr, mode))
    return 1;
  if (W_out_dev) DBX_CUDA(cudaMemcpyAsync(W_out_dev, h->ws.ptr, (size_t)h->P * 4, cudaMemcpyDeviceToDevice, st));   // W is the first workspace block
  return 0;
}

int dbx_bbb_train_epoch(dbx_handle h, const float* x_dev, const int32_t* labels_dev, int64_t N, int64_t batch, float* mean_dev, float* rho_dev,
                        const float* prior_mean_dev, const float* prior_var_dev, uint64_t seed, int64_t step0, float lrate, float kl_weight,
                        int32_t robust_mode, float epsilon, float robust_lambda, float in_lower, float in_upper, int32_t mode,
                        float* loss_out_dev, float* probs_out_dev, float* W_out_dev, void* stream) {
  DBX_CHECK(h && x_dev && labels_dev && mean_dev && rho_dev && prior_mean_dev && prior_var_dev, "null argument");
  DBX_CHECK(N > 0 && batch > 0, "N and batch must be positive");
  DBX_CHECK(mode == DBX_MODE_BF16 || mode == DBX_MODE_FP32, "unknown mode %d", mode);
  DBX_CHECK(robust_mode >= 0 && robust_mode <= 2, "robust_mode %d: 0, 1, 2 here (the epsilon-draw mode 3 needs host draws per step: dbx_bbb_step_ibp_mc)", robust_mode);
  DBX_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  int64_t i = 0;
  // Option TRAIN_GRAPH=1 (bf16 mode, standard step): the mini-batches after the first are ONE captured step replayed as a CUDA graph --
  // what differs between them (Philox step id, first row, output slot) lives in a device-side StepCtx that the graph's last node
  // advances.  Same kernels, same order, same arguments as the step-by-step loop (bit-identical parameters, tested).  It is NOT the
  // default: measured with tools/time_train_graph.py the replay was slower than the plain launches (784-512-512-10, mini-batch 128:
  // 0.18 vs 0.165 ms per step over 40 steps, 0.33 vs 0.17 over 400; mini-batch 1024: 0.22 vs 0.25 over 200, 0.59 vs 0.25 over 40),
  // i.e. the step is bound by its ~20 dependent kernels' own latencies, not by launch gaps.
  const int64_t nfull = N / batch;
  if (mode == DBX_MODE_BF16 && robust_mode == 0 && h->is_mlp && grad_tc_ok(h) && nfull >= 4 && opt().train_graph) {
    if (!h->own_stream) DBX_CUDA(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
    cudaStream_t gs = h->own_stream;      // (stream capture is not possible on the legacy default stream a caller may pass)
    cudaEvent_t ev_in = nullptr, ev_out = nullptr;
    DBX_CUDA(cudaEventCreateWithFlags(&ev_in, cudaEventDisableTiming));
    DBX_CUDA(cudaEventCreateWithFlags(&ev_out, cudaEventDisableTiming));
    DBX_CUDA(cudaEventRecord(ev_in, st));
    DBX_CUDA(cudaStreamWaitEvent(gs, ev_in, 0));
    // step 0 directly: sizes the workspace and sets the kernels' attributes outside the capture
    int rc = run_bbb_step(h, x_dev, batch, labels_dev, mean_dev, rho_dev, prior_mean_dev, prior_var_dev, nullptr, seed, step0, lrate, kl_weight,
                          loss_out_dev, probs_out_dev, 0, epsilon, robust_lambda, in_lower, in_upper, gs, 0, nullptr, mode);
    StepCtx* ctx = nullptr;
    cudaGraph_t graph = nullptr; cudaGraphExec_t exec = nullptr;
    if (!rc && cudaMalloc((void**)&ctx, sizeof(StepCtx)) != cudaSuccess) rc = set_err("cudaMalloc(StepCtx) failed");
    int64_t done = 1;
    if (!rc) {
      const StepCtx init = {(long long)(step0 + 1), (long long)batch, 1};
      if (cudaMemcpyAsync(ctx, &init, sizeof init, cudaMemcpyHostToDevice, gs) != cudaSuccess || cudaStreamSynchronize(gs) != cudaSuccess)
        rc = set_err("StepCtx upload failed");
    }
    bool captured = false;
    if (!rc && cudaStreamBeginCapture(gs, cudaStreamCaptureModeRelaxed) == cudaSuccess) {
      int crc = run_bbb_step_tc(h, x_dev, batch, labels_dev, mean_dev, rho_dev, prior_mean_dev, prior_var_dev, nullptr, seed, 0, lrate, kl_weight,
                                loss_out_dev, probs_out_dev, 0, epsilon, robust_lambda, in_lower, in_upper, gs, ctx);
      if (!crc) { step_ctx_advance_kernel<<<1, 32, 0, gs>>>(ctx, (long long)batch); g_launches.fetch_add(1, std::memory_order_relaxed); }
      const cudaError_t ce = cudaStreamEndCapture(gs, &graph);
      if (!crc && ce == cudaSuccess && graph && cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess) captured = true;
      else cudaGetLastError();      // capture not possible here: the remaining steps run one by one below
    }
    if (captured) {
      const long long per_step = g_launches.load(std::memory_order_relaxed);
      for (; done < nfull && !rc; ++done)
        if (cudaGraphLaunch(exec, gs) != cudaSuccess) rc = set_err("cudaGraphLaunch failed: %s", cudaGetErrorString(cudaGetLastError()));
      (void)per_step;
    }
    for (int64_t b0 = done * batch; b0 < N && !rc; b0 += batch, ++done) {      // what the graph did not cover (a ragged last mini-batch)
      const int64_t Bb = std::min(batch, N - b0);
      rc = run_bbb_step(h, x_dev + b0 * h->in_feat, Bb, labels_dev + b0, mean_dev, rho_dev, prior_mean_dev, prior_var_dev, nullptr, seed, step0 + done,
                        lrate, kl_weight, loss_out_dev ? loss_out_dev + 2 * done : nullptr, probs_out_dev ? probs_out_dev + b0 * h->out_feat : nullptr,
                        0, epsilon, robust_lambda, in_lower, in_upper, gs, 0, nullptr, mode);
    }
    if (!rc && W_out_dev && cudaMemcpyAsync(W_out_dev, h->ws.ptr, (size_t)h->P * 4, cudaMemcpyDeviceToDevice, gs) != cudaSuccess)
      rc = set_err("copy of the sampled weights failed");
    cudaEventRecord(ev_out, gs);
    cudaStreamWaitEvent(st, ev_out, 0);
    if (exec || ctx) cudaStreamSynchronize(gs);      // the graph and its context are freed below
    if (exec) cudaGraphExecDestroy(exec);
    if (graph) cudaGraphDestroy(graph);
    if (ctx) cudaFree(ctx);
    cudaEventDestroy(ev_in); cudaEventDestroy(ev_out);
    return rc;
  }
  for (int64_t b0 = 0; b0 < N; b0 += batch, ++i) {
    const int64_t Bb = std::min(batch, N - b0);
    if (run_bbb_step(h, x_dev + b0 * h->in_feat, Bb, labels_dev + b0, mean_dev, rho_dev, prior_mean_dev, prior_var_dev, nullptr, seed, step0 + i,
                     lrate, kl_weight, loss_out_dev ? loss_out_dev + 2 * i : nullptr, probs_out_dev ? probs_out_dev + b0 * h->out_feat : nullptr,
                     robust_mode, epsilon, robust_lambda, in_lower, in_upper, st, 0, nullptr, mode))
      return 1;
  }
  if (W_out_dev) DBX_CUDA(cudaMemcpyAsync(W_out_dev, h->ws.ptr, (size_t)h->P * 4, cudaMemcpyDeviceToDevice, st));   // the last step's sample
  return 0;
}

int dbx_bbb_step_ibp_mc(dbx_handle h, const float* x_dev, int64_t B, const int32_t* labels_dev, float* mean_dev, float* rho_dev,
                        const float* prior_mean_dev, const float* prior_var_dev, const float* noise_dev, uint64_t seed, int64_t step,
                        float lrate, float kl_weight, int32_t n_eps, const float* eps_host, float* loss_out_dev, float* probs_out_dev,
                        float* W_out_dev, void* stream) {
  DBX_CHECK(h && x_dev && labels_dev && mean_dev && rho_dev && prior_mean_dev && prior_var_dev && eps_host, "null argument");
  DBX_CHECK(B > 0 && n_eps >= 1 && n_eps <= 64, "B must be positive and 1 <= n_eps <= 64");
  for (int k = 0; k < n_eps; ++k) DBX_CHECK(eps_host[k] >= 0.f, "epsilon draw %d is negative", k);
  DBX_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  if (run_bbb_step(h, x_dev, B, labels_dev, mean_dev, rho_dev, prior_mean_dev, prior_var_dev, noise_dev, seed, step, lrate, kl_weight,
                   loss_out_dev, probs_out_dev, 3, 0.f, 0.f, 0.f, 0.f, st, n_eps, eps_host))
    return 1;
  if (W_out_dev) DBX_CUDA(cudaMemcpyAsync(W_out_dev, h->ws.ptr, (size_t)h->P * 4, cudaMemcpyDeviceToDevice, st));
  return 0;
}

int dbx_input_gradient(dbx_handle h, const float* x_dev, int64_t B, const int32_t* labels_dev, int64_t n_outer, int64_t n_inner, uint64_t seed,
                       int64_t s0, const float* noise_dev, float inflate, int32_t stored, int64_t j0, const int32_t* idx_dev, int32_t mode,
                       float* grad_dev, void* stream) {
  DBX_CHECK(h && x_dev && labels_dev && grad_dev, "null argument");
  DBX_CHECK(mode == DBX_MODE_BF16 || mode == DBX_MODE_FP32, "unknown mode %d", mode);
  DBX_CHECK(B > 0 && n_outer > 0 && n_inner > 0 && n_inner <= 4096, "B, n_outer, n_inner must be positive (n_inner <= 4096)");
  Source src; src.stored = stored != 0; src.eps = noise_dev; src.inflate = inflate; src.seed = seed; src.s0 = s0; src.j0 = j0;
  src.idx = src.stored ? idx_dev : nullptr;      // rows in [0, S): checked by the caller (Engine.input_gradient)
  DBX_CHECK(src.stored ? (h->slab != nullptr && (src.idx || (j0 >= 0 && j0 + n_outer * n_inner <= h->S))) : h->have_gaussian,
            src.stored ? "stored rows out of range / no slab set" : "no Gaussian posterior set (dbx_set_gaussian)");
  DBX_CUDA(cudaSetDevice(h->device));
  g_last_path = 0;
  if (mode == DBX_MODE_BF16 && h->is_mlp && grad_tc_ok(h)) {
    g_last_path = 3;
    return run_input_gradient_tc(h, x_dev, B, labels_dev, n_outer, n_inner, src, grad_dev, (cudaStream_t)stream);
  }
  return run_input_gradient(h, x_dev, B, labels_dev, n_outer, n_inner, src, grad_dev, (cudaStream_t)stream);
}

int dbx_input_gradient_one(dbx_handle h, const float* x_dev, int64_t B, const int32_t* labels_dev, const float* W_dev, float* grad_dev,
                           void* stream) {
  DBX_CHECK(h && W_dev, "null argument");
  const float* keep = h->slab; const int64_t keepS = h->S, keepStride = h->slab_stride;
  h->slab = W_dev; h->S = 1; h->slab_stride = h->P;
  const int rc = dbx_input_gradient(h, x_dev, B, labels_dev, 1, 1, 0, 0, nullptr, 1.f, 1, 0, nullptr, DBX_MODE_FP32, grad_dev, stream);
  h->slab = keep; h->S = keepS; h->slab_stride = keepStride;
  return rc;
}

int dbx_count_predicate_fgsm(dbx_handle h, const float* x_dev, int64_t B, const int32_t* attack_labels_dev, const int32_t* labels_dev,
                             float eps, float lower, float upper, int64_t n, uint64_t seed, int64_t s0, const float* noise_dev, float inflate,
                             int32_t mode, int32_t accumulate, uint8_t* flags_dev, int64_t* counts_dev, void* stream) {
  DBX_CHECK(h && x_dev && attack_labels_dev && labels_dev && counts_dev, "null argument");
  DBX_CHECK(B > 0 && n > 0 && eps >= 0.f, "B, n must be positive and eps >= 0");
  DBX_CHECK(mode == DBX_MODE_BF16 || mode == DBX_MODE_FP32, "unknown mode %d", mode);
  DBX_CHECK(h->have_gaussian, "no Gaussian posterior set (dbx_set_gaussian)");
  DBX_CUDA(cudaSetDevice(h->device));
  Source src; src.eps = noise_dev; src.inflate = inflate; src.seed = seed; src.s0 = s0;
  g_last_path = 0;
  if (mode == DBX_MODE_BF16 && h->is_mlp && grad_tc_ok(h)) {
    g_last_path = 3;
    return run_count_fgsm_tc(h, x_dev, B, attack_labels_dev, labels_dev, eps, lower, upper, n, src, accumulate, flags_dev,
                             (unsigned long long*)counts_dev, (cudaStream_t)stream);
  }
  return run_count_fgsm(h, x_dev, B, attack_labels_dev, labels_dev, eps, lower, upper, n, src, accumulate, flags_dev,
                        (unsigned long long*)counts_dev, (cudaStream_t)stream);
}

int dbx_count_predicate(dbx_handle h, const float* x_dev, int64_t K, const int32_t* labels_dev, int64_t n, uint64_t seed,
                        int64_t s0, const float* eps_dev, float inflate, int32_t mode, int32_t accumulate,
                        uint8_t* flags_dev, int64_t* counts_dev, void* stream) {
  DBX_CHECK(h && x_dev && labels_dev && counts_dev, "null argument");
  Source src; src.eps = eps_dev; src.inflate = inflate; src.seed = seed; src.s0 = s0;
  Sink sink; sink.kind = 1; sink.accumulate = accumulate; sink.labels = labels_dev; sink.flags = flags_dev;
  sink.counts = (unsigned long long*)counts_dev;
  return run(h, x_dev, K, n, src, sink, mode, (cudaStream_t)stream);
}

int dbx_ibp_predict(dbx_handle h, const float* x_dev, int64_t B, float eps, int32_t clip01, const int32_t* labels_dev, int64_t n,
                    uint64_t seed, int64_t s0, const float* noise_dev, float inflate, int32_t stored, int64_t j0, int32_t mode,
                    int32_t accumulate, float* sum_worst_dev, float* sum_lower_dev, float* sum_upper_dev, uint8_t* flags_dev,
                    int64_t* counts_dev, void* stream) {
  DBX_CHECK(h && x_dev, "null argument");
  DBX_CHECK(B > 0 && n > 0, "B and n must be positive (B=%lld n=%lld)", (long long)B, (long long)n);
  DBX_CHECK(eps >= 0.f, "eps must be >= 0");
  DBX_CHECK(labels_dev || (!sum_worst_dev && !counts_dev), "worst-case outputs need the class labels");
  DBX_CHECK(sum_worst_dev || sum_lower_dev || sum_upper_dev || counts_dev, "no output requested");
  DBX_CHECK(!flags_dev || counts_dev, "flags need counts");
  Source src; src.stored = stored != 0; src.eps = noise_dev; src.inflate = inflate; src.seed = seed; src.s0 = s0; src.j0 = j0;
  DBX_CHECK(src.stored ? (h->slab != nullptr && j0 >= 0 && j0 + n <= h->S) : h->have_gaussian,
            src.stored ? "stored rows out of range / no slab set" : "no Gaussian posterior set (dbx_set_gaussian)");
  DBX_CUDA(cudaSetDevice(h->device));
  IbpSink sink; sink.labels = labels_dev; sink.sum_worst = sum_worst_dev; sink.sum_lower = sum_lower_dev; sink.sum_upper = sum_upper_dev;
  sink.counts = (unsigned long long*)counts_dev; sink.flags = flags_dev; sink.accumulate = accumulate;
  g_last_path = 0; g_used_tc = false;
  int rc;
  if (mode == DBX_MODE_BF16) { rc = run_ibp<__nv_bfloat16>(h, x_dev, B, eps, clip01, n, src, sink, (cudaStream_t)stream); if (g_used_tc) g_last_path = 3; }
  else { DBX_CHECK(mode == DBX_MODE_FP32, "unknown mode %d", mode); rc = run_ibp<float>(h, x_dev, B, eps, clip01, n, src, sink, (cudaStream_t)stream); }
  return rc;
}

int dbx_predict_samples(dbx_handle h, const float* x_dev, int64_t B, int64_t n, uint64_t seed, int64_t s0, const float* eps_dev,
                        float inflate, int32_t mode, int32_t out_kind, float* out_dev, void* stream) {
  DBX_CHECK(h && x_dev && out_dev, "null argument");
  Source src; src.eps = eps_dev; src.inflate = inflate; src.seed = seed; src.s0 = s0;
  Sink sink; sink.kind = 2; sink.out_kind = out_kind; sink.sum1 = out_dev;
  return run(h, x_dev, B, n, src, sink, mode, (cudaStream_t)stream);
}

int dbx_fgsm_samples(dbx_handle h, const float* x_dev, int64_t B, const int32_t* attack_labels_dev, float eps, float lower, float upper,
                     int64_t n, uint64_t seed, int64_t s0, const float* noise_dev, float inflate, int32_t mode, float* probs_dev, void* stream) {
  DBX_CHECK(h && x_dev && attack_labels_dev && probs_dev, "null argument");
  DBX_CHECK(B > 0 && n > 0 && eps >= 0.f, "B, n must be positive and eps >= 0");
  DBX_CHECK(mode == DBX_MODE_BF16 || mode == DBX_MODE_FP32, "unknown mode %d", mode);
  DBX_CHECK(h->have_gaussian, "no Gaussian posterior set (dbx_set_gaussian)");
  DBX_CUDA(cudaSetDevice(h->device));
  Source src; src.eps = noise_dev; src.inflate = inflate; src.seed = seed; src.s0 = s0;
  g_last_path = 0;
  if (mode == DBX_MODE_BF16 && h->is_mlp && grad_tc_ok(h)) {
    g_last_path = 3;
    return run_count_fgsm_tc(h, x_dev, B, attack_labels_dev, nullptr, eps, lower, upper, n, src, 1, nullptr, nullptr, (cudaStream_t)stream,
                             probs_dev);
  }
  return run_count_fgsm(h, x_dev, B, attack_labels_dev, nullptr, eps, lower, upper, n, src, 1, nullptr, nullptr, (cudaStream_t)stream,
                        probs_dev);
}

int dbx_ibp_samples(dbx_handle h, const float* x_dev, const float* box_hi_dev, int64_t B, float eps, int32_t clip01, int64_t n,
                    uint64_t seed, int64_t s0, const float* noise_dev, float inflate, int32_t mode, float* lower_dev, float* upper_dev,
                    void* stream) {
  DBX_CHECK(h && x_dev && lower_dev && upper_dev, "null argument");
  DBX_CHECK(B > 0 && n > 0 && eps >= 0.f, "B, n must be positive and eps >= 0");
  DBX_CHECK(h->have_gaussian, "no Gaussian posterior set (dbx_set_gaussian)");
  DBX_CUDA(cudaSetDevice(h->device));
  Source src; src.eps = noise_dev; src.inflate = inflate; src.seed = seed; src.s0 = s0;
  IbpSink sink; sink.per_lower = lower_dev; sink.per_upper = upper_dev;
  g_last_path = 0; g_used_tc = false;
  int rc;
  if (mode == DBX_MODE_BF16) { rc = run_ibp<__nv_bfloat16>(h, x_dev, B, eps, clip01, n, src, sink, (cudaStream_t)stream, box_hi_dev); if (g_used_tc) g_last_path = 3; }
  else { DBX_CHECK(mode == DBX_MODE_FP32, "unknown mode %d", mode); rc = run_ibp<float>(h, x_dev, B, eps, clip01, n, src, sink, (cudaStream_t)stream, box_hi_dev); }
  return rc;
}

int dbx_ibp_one(dbx_handle h, const float* x_dev, int64_t B, float eps, int32_t clip01, const float* W_dev, int32_t mode,
                float* lower_dev, float* upper_dev, void* stream) {
  DBX_CHECK(h && x_dev && W_dev && lower_dev && upper_dev, "null argument");
  const float* keep = h->slab; const int64_t keepS = h->S, keepStride = h->slab_stride;
  h->slab = W_dev; h->S = 1; h->slab_stride = h->P;
  const int rc = dbx_ibp_predict(h, x_dev, B, eps, clip01, nullptr, 1, 0, 0, nullptr, 1.f, 1, 0, mode, 0, nullptr, lower_dev, upper_dev,
                                 nullptr, nullptr, stream);
  h->slab = keep; h->S = keepS; h->slab_stride = keepStride;
  return rc;
}

int dbx_ibp_state(dbx_handle h, const float* lo_dev, const float* hi_dev, int64_t B, const float* W_dev, float weight_margin,
                  float* lower_dev, float* upper_dev, void* stream) {
  DBX_CHECK(h && lo_dev && hi_dev && W_dev && lower_dev && upper_dev, "null argument");
  DBX_CHECK(B > 0 && weight_margin >= 0.f, "B must be positive and weight_margin >= 0");
  DBX_CUDA(cudaSetDevice(h->device));
  const float* keep = h->slab; const int64_t keepS = h->S, keepStride = h->slab_stride;
  h->slab = W_dev; h->S = 1; h->slab_stride = h->P;
  Source src; src.stored = true; src.j0 = 0;
  IbpSink sink; sink.sum_lower = lower_dev; sink.sum_upper = upper_dev;
  g_last_path = 0; g_used_tc = false;
  const int rc = run_ibp<float>(h, lo_dev, B, 0.f, 0, 1, src, sink, (cudaStream_t)stream, hi_dev, weight_margin);
  h->slab = keep; h->S = keepS; h->slab_stride = keepStride;
  return rc;
}

static bool host_ptr_is_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}

constexpr size_t kUpPiece = 4u << 20;

int dbx_upload_f32(dbx_handle h, const float* src_host, float* dst_dev, int64_t count, void* stream) {
  DBX_CHECK(h && src_host && dst_dev && count >= 0, "dbx_upload_f32: bad argument");
  DBX_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t bytes = (size_t)count * 4;
  if (bytes == 0) return 0;
  if (host_ptr_is_pinned(src_host)) {   // page-locked by the caller: one DMA, nothing staged; the caller keeps it alive until the stream has passed
    DBX_CUDA(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, st));
    return 0;
  }
  for (int i = 0; i < 2; ++i)
    if (!h->up_stage[i]) {
      DBX_CUDA(cudaMallocHost(&h->up_stage[i], kUpPiece));
      DBX_CUDA(cudaEventCreateWithFlags(&h->up_ev[i], cudaEventDisableTiming));
      DBX_CUDA(cudaEventRecord(h->up_ev[i], st));
    }
  // pageable memory: pieces of 4 MB through two pinned buffers, the host memcpy of piece i+1 under the DMA of piece i; a
  // buffer is rewritten only after the event of its previous copy has completed
  for (size_t off = 0; off < bytes; off += kUpPiece) {
    const int b = h->up_next; h->up_next ^= 1;
    const size_t nb = std::min(kUpPiece, bytes - off);
    DBX_CUDA(cudaEventSynchronize(h->up_ev[b]));
    memcpy(h->up_stage[b], (const char*)src_host + off, nb);
    DBX_CUDA(cudaMemcpyAsync((char*)dst_dev + off, h->up_stage[b], nb, cudaMemcpyHostToDevice, st));
    DBX_CUDA(cudaEventRecord(h->up_ev[b], st));
  }
  return 0;
}

int dbx_predict_host(dbx_handle h, const float* x_host, int64_t B, int64_t n, uint64_t seed, int64_t s0, float inflate,
                     int32_t mode, int32_t out_kind, float* mean_host, float* var_host) {
  DBX_CHECK(h && x_host && mean_host, "null argument");
  DBX_CUDA(cudaSetDevice(h->device));
  if (!h->own_stream) DBX_CUDA(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
  cudaStream_t st = h->own_stream;
  const size_t xb = (size_t)B * h->in_feat * 4, ob = (size_t)B * h->out_feat * 4;
  if (h->h_pin_bytes < 2 * ob) {
    if (h->h_pin) DBX_CUDA(cudaFreeHost(h->h_pin));
    h->h_pin = nullptr; h->h_pin_bytes = 0;
    DBX_CUDA(cudaMallocHost(&h->h_pin, 2 * ob));
    h->h_pin_bytes = 2 * ob;
  }
  if (h->d_x_bytes < xb) {
    cudaFree(h->d_x); h->d_x = nullptr; h->d_x_bytes = 0;
    DBX_CUDA(cudaMalloc(&h->d_x, xb)); h->d_x_bytes = xb;
  }
  if (h->d_s_bytes < ob) {
    cudaFree(h->d_s1); cudaFree(h->d_s2); h->d_s1 = h->d_s2 = nullptr; h->d_s_bytes = 0;
    DBX_CUDA(cudaMalloc(&h->d_s1, ob)); DBX_CUDA(cudaMalloc(&h->d_s2, ob)); h->d_s_bytes = ob;
  }
  if (dbx_upload_f32(h, x_host, h->d_x, B * h->in_feat, st)) return 1;
  if (dbx_predict(h, h->d_x, B, n, seed, s0, nullptr, inflate, mode, out_kind, 0, h->d_s1, var_host ? h->d_s2 : nullptr, st))
    return 1;
  DBX_LAUNCH(finalize_kernel, grid_for(B * h->out_feat), 256, 0, st, h->d_s1, var_host ? h->d_s2 : nullptr,
             1.0f / (float)n, B * h->out_feat, h->d_s1, var_host ? h->d_s2 : nullptr);
  const bool direct = host_ptr_is_pinned(mean_host) && (!var_host || host_ptr_is_pinned(var_host));
  float* om = direct ? mean_host : (float*)h->h_pin;
  float* ov = direct ? var_host : (float*)((char*)h->h_pin + ob);
  DBX_CUDA(cudaMemcpyAsync(om, h->d_s1, ob, cudaMemcpyDeviceToHost, st));
  if (var_host) DBX_CUDA(cudaMemcpyAsync(ov, h->d_s2, ob, cudaMemcpyDeviceToHost, st));
  DBX_CUDA(cudaStreamSynchronize(st));
  if (!direct) {
    memcpy(mean_host, om, ob);
    if (var_host) memcpy(var_host, ov, ob);
  }
  return 0;
}

int dbx_activation(int32_t act, float* x_dev, int64_t rows, int64_t cols, void* stream) {
  DBX_CHECK(x_dev && rows > 0 && cols > 0, "bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  DBX_LAUNCH(act_rows_kernel, grid_for(rows), 256, 0, st, act, x_dev, rows, cols);
  return 0;
}

int dbx_philox_raw(uint64_t seed, int64_t s, int64_t P, uint32_t* out_dev, void* stream) {
  DBX_CHECK(out_dev && P > 0, "bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  DBX_LAUNCH(philox_raw_kernel, grid_for((P + 3) / 4), 256, 0, st, seed, (uint64_t)s, P, out_dev);
  return 0;
}

}  // extern "C"
