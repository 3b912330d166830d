// Streaming MLP kernel for SMALL input batches (B <= 16): the HBM-bound regime of the path.
//
// Stored posteriors (hmc.py:308-322 -> posteriormodel.py:77-78): every stored weight sample is
// read from the HBM-resident [S,P] fp32 slab exactly once, with coalesced 128-bit loads, and used
// for all B rows while it is in registers -- 4*P algorithmic bytes per sample, no intermediate
// copy (BASELINE config 3).  Gaussian posteriors at small B (the reference's tutorials call
// predict / massart_bound_check with ONE input): W = mu + sigma*eps is formed in registers from
// the same Philox stream dbx_sample uses, so W_s never exists in memory at all.
//
// One CTA walks samples blockIdx.x, blockIdx.x + gridDim.x, ...; activations live in shared
// memory as [feature][BT]; all arithmetic is fp32 FFMA (this path serves both DBX_MODE_FP32 and
// DBX_MODE_BF16 -- it is more precise than the bf16 tensor path and the work is bandwidth / RNG
// bound, not FLOP bound).  Per-CTA running sums are written once and added in CTA order by
// stream_finish_kernel (deterministic).
#include <algorithm>
#include <cstdlib>
#include "engine.h"

namespace dbx {

constexpr int kStreamThreads = 256;
constexpr int kStreamMaxLayers = 8;

struct StreamLayer { int K, N, act; int w_off, b_off; int vec_ok; };

struct StreamParams {
  int nlayers, B, C, in_feat, maxw;
  StreamLayer L[kStreamMaxLayers];
  long long n;                 // samples in this call
  long long P; long long slab_stride;
  // source
  int stored;
  const float* slab; const int32_t* idx; long long j0;          // stored
  const float* mu; const float* sigma; const float* eps; float inflate; unsigned long long seed; long long s0;
  int unfused;
  // input
  const float* x;              // [B, in_feat]
  // sink
  int sink_kind, out_kind, act_last;
  const float* wts; const int32_t* labels; uint8_t* flags; unsigned long long* counts;
  float* partial;              // [gridDim][2][BT*C]
};

// 4 consecutive weights starting at flat index j (j % 4 == 0, aligned source rows)
template <int SRC>
__device__ __forceinline__ float4 load_w4(const StreamParams& p, const float* __restrict__ srow, const float* __restrict__ erow,
                                          unsigned long long sample, int j) {
  if (SRC == 0) return __ldcs(reinterpret_cast<const float4*>(srow + j));   // streaming: read once
  const float4 m4 = __ldg(reinterpret_cast<const float4*>(p.mu + j));
  const float4 s4 = __ldg(reinterpret_cast<const float4*>(p.sigma + j));
  float z[4];
  if (SRC == 2) { const float4 e4 = __ldcs(reinterpret_cast<const float4*>(erow + j)); z[0] = e4.x; z[1] = e4.y; z[2] = e4.z; z[3] = e4.w; }
  else philox_normal4(p.seed, sample, (unsigned long long)(j >> 2), z);
  float4 w;
  if (p.unfused) {
    w.x = __fadd_rn(m4.x, __fmul_rn(p.inflate * s4.x, z[0])); w.y = __fadd_rn(m4.y, __fmul_rn(p.inflate * s4.y, z[1]));
    w.z = __fadd_rn(m4.z, __fmul_rn(p.inflate * s4.z, z[2])); w.w = __fadd_rn(m4.w, __fmul_rn(p.inflate * s4.w, z[3]));
  } else {
    w.x = fmaf(p.inflate * s4.x, z[0], m4.x); w.y = fmaf(p.inflate * s4.y, z[1], m4.y);
    w.z = fmaf(p.inflate * s4.z, z[2], m4.z); w.w = fmaf(p.inflate * s4.w, z[3], m4.w);
  }
  return w;
}

// one weight at flat index j (any alignment)
template <int SRC>
__device__ __forceinline__ float load_w1(const StreamParams& p, const float* __restrict__ srow, const float* __restrict__ erow,
                                         unsigned long long sample, int j) {
  if (SRC == 0) return __ldcs(srow + j);
  float e;
  if (SRC == 2) e = erow[j];
  else { float z[4]; philox_normal4(p.seed, sample, (unsigned long long)(j >> 2), z); e = z[j & 3]; }
  const float sc = p.inflate * p.sigma[j];
  return p.unfused ? __fadd_rn(p.mu[j], __fmul_rn(sc, e)) : fmaf(sc, e, p.mu[j]);
}

extern __shared__ __align__(16) float stream_smem[];

template <int BT, int SRC>
__global__ void __launch_bounds__(kStreamThreads, (BT <= 2 ? 3 : 2))
stream_mlp_kernel(const StreamParams p) {
  float* xs = stream_smem;                       // [in_feat][BT]   (input, shared by all samples)
  float* hA = xs + (size_t)p.in_feat * BT;       // [maxw][BT]
  float* hB = hA + (size_t)p.maxw * BT;          // [maxw][BT]
  float* acc1 = hB + (size_t)p.maxw * BT;        // [BT*C] running sum of outputs over this CTA's samples
  float* acc2 = acc1 + BT * p.C;                 // [BT*C] running sum of squares
  const int t = threadIdx.x;
  for (int i = t; i < p.in_feat * BT; i += kStreamThreads) {
    const int k = i / BT, b = i - k * BT;
    xs[i] = (b < p.B) ? p.x[(size_t)b * p.in_feat + k] : 0.f;
  }
  for (int i = t; i < 2 * BT * p.C; i += kStreamThreads) acc1[i] = 0.f;
  __syncthreads();

  for (long long si = blockIdx.x; si < p.n; si += gridDim.x) {
    const unsigned long long sample = (unsigned long long)(p.s0 + si);
    const float* srow = p.stored ? p.slab + (p.idx ? (long long)p.idx[si] : (p.j0 + si)) * p.slab_stride : nullptr;
    const float* erow = (!p.stored && p.eps) ? p.eps + si * p.P : nullptr;
    const float* hin = xs;
    float* hout = hA;
    for (int l = 0; l < p.nlayers; ++l) {
      const StreamLayer L = p.L[l];
      const bool last = (l == p.nlayers - 1);
      const int act = last ? DBX_ACT_LINEAR : L.act;   // the sink applies the last activation
      if (L.N > 16) {
        if (L.vec_ok) {
          // ---- column mapping, 128-bit loads: thread owns 4 adjacent output features
          for (int g = t; g < (L.N >> 2); g += kStreamThreads) {
            float a[BT][4];
#pragma unroll
            for (int b = 0; b < BT; ++b) { a[b][0] = a[b][1] = a[b][2] = a[b][3] = 0.f; }
            const int col = g << 2;
            int k = 0;
            constexpr int U = (SRC == 0 && BT <= 2) ? 12 : 8;   // independent 128-bit loads in flight per thread
            for (; k + U <= L.K; k += U) {
              float4 w[U];
#pragma unroll
              for (int u = 0; u < U; ++u) w[u] = load_w4<SRC>(p, srow, erow, sample, L.w_off + (k + u) * L.N + col);
#pragma unroll
              for (int u = 0; u < U; ++u) {
                const float* hk = hin + (size_t)(k + u) * BT;
                float hv[BT];
                if (BT >= 4) {   // activations are [feature][BT]: one 128-bit broadcast load per 4 rows
#pragma unroll
                  for (int b4 = 0; b4 < BT / 4; ++b4) {
                    const float4 h4 = *reinterpret_cast<const float4*>(hk + 4 * b4);
                    hv[4 * b4] = h4.x; hv[4 * b4 + 1] = h4.y; hv[4 * b4 + 2] = h4.z; hv[4 * b4 + 3] = h4.w;
                  }
                } else {
#pragma unroll
                  for (int b = 0; b < BT; ++b) hv[b] = hk[b];
                }
#pragma unroll
                for (int b = 0; b < BT; ++b) {
                  const float h = hv[b];
                  a[b][0] = fmaf(h, w[u].x, a[b][0]); a[b][1] = fmaf(h, w[u].y, a[b][1]);
                  a[b][2] = fmaf(h, w[u].z, a[b][2]); a[b][3] = fmaf(h, w[u].w, a[b][3]);
                }
              }
            }
            for (; k < L.K; ++k) {
              const float4 w = load_w4<SRC>(p, srow, erow, sample, L.w_off + k * L.N + col);
              const float* hk = hin + (size_t)k * BT;
#pragma unroll
              for (int b = 0; b < BT; ++b) {
                const float h = hk[b];
                a[b][0] = fmaf(h, w.x, a[b][0]); a[b][1] = fmaf(h, w.y, a[b][1]);
                a[b][2] = fmaf(h, w.z, a[b][2]); a[b][3] = fmaf(h, w.w, a[b][3]);
              }
            }
            const float4 bias = load_w4<SRC>(p, srow, erow, sample, L.b_off + col);
            const float bb[4] = {bias.x, bias.y, bias.z, bias.w};
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
              for (int b = 0; b < BT; ++b) hout[(size_t)(col + j) * BT + b] = apply_act(act, a[b][j] + bb[j]);
          }
        } else {
          // ---- column mapping, scalar loads (shapes that are not 4-aligned)
          for (int col = t; col < L.N; col += kStreamThreads) {
            float a[BT];
#pragma unroll
            for (int b = 0; b < BT; ++b) a[b] = 0.f;
            for (int k = 0; k < L.K; ++k) {
              const float w = load_w1<SRC>(p, srow, erow, sample, L.w_off + k * L.N + col);
              const float* hk = hin + (size_t)k * BT;
#pragma unroll
              for (int b = 0; b < BT; ++b) a[b] = fmaf(hk[b], w, a[b]);
            }
            const float bias = load_w1<SRC>(p, srow, erow, sample, L.b_off + col);
#pragma unroll
            for (int b = 0; b < BT; ++b) hout[(size_t)col * BT + b] = apply_act(act, a[b] + bias);
          }
        }
      } else {
        // ---- narrow layer (N <= 16, typically the output layer): threads split K, every thread
        // produces partial sums for all N outputs, reduced through shared memory
        // hout doubles as scratch: [N][BT] result; partials go through warp shuffles first
        for (int b0 = 0; b0 < BT; b0 += 4) {
          float a[4][16];
#pragma unroll
          for (int bb = 0; bb < 4; ++bb)
#pragma unroll
            for (int c = 0; c < 16; ++c) a[bb][c] = 0.f;
          for (int k = t; k < L.K; k += kStreamThreads) {
            const float* hk = hin + (size_t)k * BT + b0;
            float h[4];
#pragma unroll
            for (int bb = 0; bb < 4; ++bb) h[bb] = (b0 + bb < BT) ? hk[bb] : 0.f;
#pragma unroll
            for (int c = 0; c < 16; ++c) {
              if (c < L.N) {   // consecutive threads read consecutive rows: the warp's loads are contiguous
                const float w = load_w1<SRC>(p, srow, erow, sample, L.w_off + k * L.N + c);
#pragma unroll
                for (int bb = 0; bb < 4; ++bb) a[bb][c] = fmaf(h[bb], w, a[bb][c]);
              }
            }
          }
          // block reduction: warp shuffle, then one smem slot per warp
          __shared__ float red[kStreamThreads / 32][4][16];
#pragma unroll
          for (int bb = 0; bb < 4; ++bb)
#pragma unroll
            for (int c = 0; c < 16; ++c) {
              if (c < L.N) {
                float v = a[bb][c];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if ((t & 31) == 0) red[t >> 5][bb][c] = v;
              }
            }
          __syncthreads();
          if (t < 4 * L.N) {
            const int bb = t / L.N, c = t - bb * L.N;
            if (b0 + bb < BT) {
              float v = 0.f;
#pragma unroll
              for (int w = 0; w < kStreamThreads / 32; ++w) v += red[w][bb][c];
              const float bias = load_w1<SRC>(p, srow, erow, sample, L.b_off + c);
              hout[(size_t)c * BT + b0 + bb] = apply_act(act, v + bias);
            }
          }
          __syncthreads();
        }
      }
      __syncthreads();
      hin = hout;
      hout = (hout == hA) ? hB : hA;
    }
    // ---- sink: hin holds the logits [C][BT]
    if (t < p.B) {
      const int b = t;
      if (p.sink_kind == 0) {
        const float w = p.wts ? p.wts[si] : 1.f;
        if (p.out_kind == DBX_OUT_PROBS && p.act_last == DBX_ACT_SOFTMAX) {
          float m = hin[b];
          for (int c = 1; c < p.C; ++c) m = fmaxf(m, hin[(size_t)c * BT + b]);
          float den = 0.f;
          for (int c = 0; c < p.C; ++c) den += expf(hin[(size_t)c * BT + b] - m);
          for (int c = 0; c < p.C; ++c) {
            const float pr = expf(hin[(size_t)c * BT + b] - m) / den;
            acc1[b * p.C + c] = fmaf(w, pr, acc1[b * p.C + c]);
            acc2[b * p.C + c] = fmaf(w * pr, pr, acc2[b * p.C + c]);
          }
        } else {
          for (int c = 0; c < p.C; ++c) {
            const float z = hin[(size_t)c * BT + b];
            const float pr = (p.out_kind == DBX_OUT_LOGITS) ? z : apply_act(p.act_last, z);
            acc1[b * p.C + c] = fmaf(w, pr, acc1[b * p.C + c]);
            acc2[b * p.C + c] = fmaf(w * pr, pr, acc2[b * p.C + c]);
          }
        }
      } else {
        int best = 0; float bv = hin[b];
        for (int c = 1; c < p.C; ++c) { const float z = hin[(size_t)c * BT + b]; if (z > bv) { bv = z; best = c; } }
        const int ok = (best == p.labels[b]);
        if (p.flags) p.flags[si * p.B + b] = (uint8_t)ok;
        if (ok) atomicAdd(p.counts + b, 1ULL);
      }
    }
    __syncthreads();
  }
  if (p.sink_kind == 0) {
    float* dst = p.partial + (size_t)blockIdx.x * 2 * BT * p.C;
    for (int i = t; i < 2 * BT * p.C; i += kStreamThreads) dst[i] = acc1[i];
  }
}

__global__ void stream_finish_kernel(const float* __restrict__ partial, int nblocks, int BT, int B, int C, int overwrite,
                                     float* __restrict__ sum1, float* __restrict__ sum2) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= B * C) return;
  float a1 = overwrite ? 0.f : sum1[o];
  float a2 = (sum2 && !overwrite) ? sum2[o] : 0.f;
  for (int k = 0; k < nblocks; ++k) {
    const float* src = partial + (size_t)k * 2 * BT * C;
    a1 += src[o];
    a2 += src[BT * C + o];
  }
  sum1[o] = a1;
  if (sum2) sum2[o] = a2;
}

__global__ void stream_zero_counts(unsigned long long* p, int n) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o < n) p[o] = 0ULL;
}

template <int BT, int SRC>
static int launch_stream2(StreamParams& p, int num_sms, size_t smem, cudaStream_t st, int* grid_out) {
  DBX_CUDA(cudaFuncSetAttribute(stream_mlp_kernel<BT, SRC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 1;
  DBX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, stream_mlp_kernel<BT, SRC>, kStreamThreads, smem));
  per_sm = std::max(per_sm, 1);
  // one resident wave; if there are fewer samples than that, one CTA per sample
  const int grid = (int)std::min<long long>(p.n, (long long)num_sms * per_sm);
  *grid_out = grid;
  stream_mlp_kernel<BT, SRC><<<grid, kStreamThreads, smem, st>>>(p);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  DBX_CUDA(cudaGetLastError());
  return 0;
}
template <int BT>
static int launch_stream(StreamParams& p, int num_sms, size_t smem, cudaStream_t st, int* grid_out) {
  if (p.stored) return launch_stream2<BT, 0>(p, num_sms, smem, st, grid_out);
  if (p.eps) return launch_stream2<BT, 2>(p, num_sms, smem, st, grid_out);
  return launch_stream2<BT, 1>(p, num_sms, smem, st, grid_out);
}

// returns 1 if handled, 0 if the shape is not for this path, <0 on error
int stream_mlp_run(dbx_model* m, const float* x_dev, int64_t B, int64_t n, const Source& src, const Sink& sink,
                   cudaStream_t st) {
  if (!m->is_mlp || B > 16 || B < 1) return 0;
  if (opt().no_stream) return 0;
  StreamParams p = {};
  int nl = 0;
  int maxw = 1;
  for (auto& L : m->layers) {
    if (L.kind != DBX_DENSE) continue;
    if (nl >= kStreamMaxLayers) return 0;
    if (L.w_rows > 8192 || L.w_cols > 8192) return 0;
    StreamLayer& S = p.L[nl++];
    S.K = (int)L.w_rows; S.N = (int)L.w_cols; S.act = L.act; S.w_off = (int)L.w_off; S.b_off = (int)L.b_off;
    // 128-bit loads need 4-aligned rows / offsets (and, for slab rows, P % 4 == 0)
    S.vec_ok = (S.N % 4 == 0) && (S.w_off % 4 == 0) && (S.b_off % 4 == 0) &&
               (src.stored ? (m->slab_stride % 4 == 0) : (!src.eps || m->P % 4 == 0));   // slab / eps row starts
    maxw = std::max(maxw, S.N);
  }
  if (nl == 0) return 0;
  int BT = 1;
  while (BT < B) BT <<= 1;
  const int C = (int)m->out_feat;
  const size_t smem = ((size_t)m->in_feat * BT + 2 * (size_t)maxw * BT + 2 * (size_t)BT * C) * sizeof(float);
  if (smem > 200 * 1024) return 0;
  static int num_sms = 0;
  if (!num_sms && cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, m->device) != cudaSuccess) return -1;
  // workspace for per-CTA partials (upper bound: 8 CTAs per SM)
  const size_t part_b = (size_t)num_sms * 8 * 2 * BT * C * sizeof(float);
  if (ensure_ws(m->ws, std::max<size_t>(part_b, 256))) return -1;
  p.nlayers = nl; p.B = (int)B; p.C = C; p.in_feat = (int)m->in_feat; p.maxw = maxw; p.n = n; p.P = m->P;
  p.stored = src.stored ? 1 : 0; p.slab = m->slab; p.slab_stride = m->slab_stride; p.idx = src.idx; p.j0 = src.j0;
  p.mu = m->mu; p.sigma = m->sigma; p.eps = src.eps; p.inflate = src.inflate; p.seed = src.seed; p.s0 = src.s0;
  p.unfused = src.unfused ? 1 : 0;
  p.x = x_dev;
  p.sink_kind = sink.kind; p.out_kind = sink.out_kind; p.act_last = m->layers[m->last_weighted].act;
  p.wts = sink.wts; p.labels = sink.labels; p.flags = sink.flags; p.counts = sink.counts;
  p.partial = (float*)m->ws.ptr;
  if (sink.kind == 1 && !sink.accumulate) {
    stream_zero_counts<<<(int)cdiv(B, 64), 64, 0, st>>>(sink.counts, (int)B);
    g_launches.fetch_add(1);
  }
  prof_begin(st);
  int rc = 0;
  int grid = 1;
  switch (BT) {
    case 1: rc = launch_stream<1>(p, num_sms, smem, st, &grid); break;
    case 2: rc = launch_stream<2>(p, num_sms, smem, st, &grid); break;
    case 4: rc = launch_stream<4>(p, num_sms, smem, st, &grid); break;
    case 8: rc = launch_stream<8>(p, num_sms, smem, st, &grid); break;
    default: rc = launch_stream<16>(p, num_sms, smem, st, &grid); break;
  }
  if (rc) return -1;
  prof_end(st, (double)m->flops * (double)B * (double)n);
  if (sink.kind == 0) {
    stream_finish_kernel<<<(int)cdiv(B * C, 128), 128, 0, st>>>((const float*)m->ws.ptr, grid, BT, (int)B, C,
                                                                 sink.accumulate ? 0 : 1, sink.sum1, sink.sum2);
    g_launches.fetch_add(1);
    if (cudaGetLastError() != cudaSuccess) { set_err("stream finish launch failed"); return -1; }
  }
  return 1;
}

}  // namespace dbx
