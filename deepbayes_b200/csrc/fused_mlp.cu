// Fused Bayesian-MLP posterior-predictive kernel for sm_100a (tcgen05 / TMEM / TMA).
//
// One launch evaluates a chunk of `ns` posterior samples for all B input rows of an MLP
//     x[K0] -> Dense(H, relu) [-> Dense(H, relu)] -> Dense(C <= 16, softmax | linear)
// (BASELINE configs 1, 2, 5: 784-512-10 and 784-512-512-10).  The per-sample weights W_s were
// materialised just before by sample_bf16_kernel (Philox + mu/sigma transform) in Keras layout,
// bf16, and are streamed with TMA; activations never leave the SM:
//
//   unit = (posterior sample s, 128-row tile m).  Per unit, in one CTA:
//     L1  : X tile (smem, K-major SW128)  x  W1_s (smem, MN-major SW128)  -> fp32 acc in TMEM
//           two N=256 chunks -> TMEM regions R0=[0,256), R1=[256,512)
//     E1  : epilogue warps: acc + b1, relu, pack bf16x2 IN PLACE -> R[0:128) (next A operand)
//     L2  : A = packed H1 straight from TMEM (tcgen05.mma .ts), W2_s from smem, four N=128
//           chunks into the freed halves R0[128:256) / R1[128:256)
//     E2  : acc + b2, relu, pack in place -> buf[0:64)
//     L3  : per chunk, A = packed H2 chunk (TMEM), W3_s (smem, no-swizzle, N=16) -> 16 cols,
//           drained to registers and summed over chunks
//     fin : + b3, softmax (fp32), running sum of p and p^2 over this CTA's samples in registers
//   (posteriormodel.py:94-101: per-sample logits / probabilities never reach HBM.)
//
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM owner), warps 2-5 = epilogue
// (one thread per accumulator row).  CTAs of a cluster work on the same sample and adjacent row
// tiles, so every weight box is fetched from L2 once per cluster and multicast.
//
// Work split: units are flattened [tile-group q][sample s] and cut into contiguous ranges, one
// per cluster; a CTA keeps its running sums in registers while q stays the same and writes one
// partial block per (cluster, segment); finish_kernel adds the partials in sample order, so
// results are deterministic (no float atomics).
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include "engine.h"
#include "tc05.cuh"

namespace dbx {

using namespace tc05;

constexpr int kStages = 4;
constexpr int kStageBytes = 49152;           // X block 16 KB + weight blocks 32 KB
constexpr int kXBytes = 16384;
constexpr int kBlkBytes = 8192;              // one [64 n x 64 k] MN-major weight box
constexpr int kW3Bytes = 16384;              // [H<=512 x 16] bf16, no-swizzle core-matrix layout
constexpr int kBiasFloats = 1056;            // b1 | b2 | b3 (each padded to 4 floats)
constexpr int kThreads = 192;
constexpr int kMaxC = 16;

struct FusedParams {
  int K0, H, NH, C, B, ns, CM, mc;
  long long U;                 // units in this launch = Q * ns
  int Q;                       // tile groups (CM row tiles each)
  int NC;                      // clusters
  int maxseg;
  int sink_kind, out_kind, act_last, want_sum2;
  const float* bias;           // [ns, PB] fp32
  long long PB;
  int b1_off, b2_off, b3_off;  // offsets inside one sample's bias row
  const float* wts;            // [ns] or null
  const int32_t* labels;       // [B] (count sink)
  uint8_t* flags;              // [ns, B] or null
  unsigned long long* counts;  // [B]
  float* partial;              // [NC][maxseg][CM*128][2][16]
  int w3f_off;                 // v3: floats of (b1|b2|b3) per sample = offset of the fp32 [H][12] output-layer kernel
  const __nv_bfloat16* w3;     // last-layer kernels of the chunk, blocked8 layout: sample s at w3 + s*P16
  long long P16; int w3_bytes;
  int chunk;                   // index of this launch within the call (order 1 keeps its running sums across launches)
  int order;                   // v4: 0 = [q][s] contiguous ranges, 1 = [s][q] round-robin with running sums in `partial`
  unsigned long long* visited; // v4 order 1: per cluster, bit q set when the cluster wrote its (cluster, q) slice
  const unsigned int* ready;   // stream mode: ready[k] != 0 once sub-chunk k of the launch's samples has been drawn (null: all drawn before the launch)
  int sub0, subc;              // sub-chunk 0 holds samples [0, sub0), sub-chunk k >= 1 samples [sub0 + (k-1)*subc, sub0 + k*subc)
  int l3_at;                   // pair kernel: k-step of a layer-2 chunk at which the previous chunk's output-layer MMAs may be slipped in (0 = never)
  unsigned trace_only;            // 0 = all probes, else up to four probe ids, one per byte (a probe costs ~180 cycles on the warp that takes it)
  long long* trace;            // debug timeline of block 0 (null in production): [0]=count, then (id, clock) pairs
};

struct __align__(8) Barriers {
  uint64_t full[kStages], empty[kStages];
  uint64_t w3_full, w3_empty;
  uint64_t bias_full[2], bias_empty[2];
  uint64_t acc_full[2], act1_ready[2];
  uint64_t acc2_full[2], act2_ready[2];
  uint64_t l3_full[2], l3_drained[2];
  uint32_t tmem_slot;
};

extern __shared__ __align__(1024) uint8_t fused_smem_raw[];

// After the wait on a TMA stage's `full` barrier the MMA warp needs no tcgen05 fence: the mbarrier completion
// already orders the async-proxy writes before the MMA's operand reads, and tcgen05.fence::after_thread_sync is only
// for ordering against OTHER threads' tcgen05 work (the epilogue barriers keep theirs).  -DDBX_STAGE_FENCE_ON=1
// restores the per-stage fence for A/B measurements.
#if defined(DBX_STAGE_FENCE_ON) && DBX_STAGE_FENCE_ON
#define DBX_STAGE_FENCE() fence_after_sync()
#else
#define DBX_STAGE_FENCE() do {} while (0)
#endif

// Stream mode: the sampler of a chunk runs CONCURRENTLY with the fused kernel that consumes it (side stream, launched
// first); after each sub-chunk of samples a one-thread kernel publishes ready[k] (the kernel boundary before it makes the
// sampler's writes visible).  A producer warp acquires the flag of a unit's sample before its first copy of that sample's
// weights; the copies read global memory through the async proxy, hence the proxy fence after the acquire.
__device__ __forceinline__ void wait_sample_ready(const FusedParams& p, int s, int& ready_upto, int lane) {
  if (p.ready == nullptr) return;
  const int sub = s < p.sub0 ? 0 : 1 + (s - p.sub0) / p.subc;
  if (sub <= ready_upto) return;
  const unsigned int* f = p.ready + sub;
  unsigned int v = 0, spins = 0;
  do {
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
    if (v == 0) {
      __nanosleep(200);
      if (++spins > (1u << 24)) { if (lane == 0) printf("fused kernel: sample %d never became ready (block %d)\n", s, blockIdx.x); __trap(); }
    }
  } while (v == 0);
  asm volatile("fence.proxy.async.global;" ::: "memory");
  __syncwarp();
  ready_upto = sub;
}
__global__ void flag_set_kernel(unsigned int* flag) {
  if (threadIdx.x == 0) { __threadfence(); asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flag), "r"(1u) : "memory"); }
}

// debug timeline (block 0 only): each traced warp appends (id, clock) to its own lane of the
// buffer with a private counter -- no atomics, so the probe costs a clock read and two stores.
#define DBX_TRACE(id)                                                                         \
  do {                                                                                        \
    if (p.trace != nullptr && (p.trace_only == 0 || (p.trace_only & 0xFF) == (id) || ((p.trace_only >> 8) & 0xFF) == (id) || ((p.trace_only >> 16) & 0xFF) == (id) || ((p.trace_only >> 24) & 0xFF) == (id)) && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && trace_n < 2600) { \
      long long* _t = p.trace + 1 + (threadIdx.x >> 5) * 5200 + 2 * trace_n;                  \
      _t[0] = (id); _t[1] = clock64(); ++trace_n;                                             \
    }                                                                                         \
  } while (0)

// (a register cap of 128 -- no spills -- would let a third sampler block run beside this CTA: measured SLOWER, config 5
// 85.1 -> 88.5 ms; the extra sampler warps take issue slots from the epilogue)
template <int CM>
__global__ void __launch_bounds__(kThreads, 1)
fused_mlp_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
                 const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmW3,
                 const FusedParams p) {
  uint8_t* smem = (uint8_t*)(((uintptr_t)fused_smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t s_stage = smem_u32(smem);
  const uint32_t s_w3 = s_stage + kStages * kStageBytes;
  float* bias_smem = (float*)(smem + kStages * kStageBytes + kW3Bytes);
  const uint32_t s_bias = smem_u32(bias_smem);
  Barriers* bars = (Barriers*)(smem + kStages * kStageBytes + kW3Bytes + 2 * kBiasFloats * 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int trace_n = 0;
  const uint32_t rank = (CM > 1) ? cluster_ctarank() : 0;
  const int cluster_id = blockIdx.x / CM;
  constexpr uint16_t kMask = (uint16_t)((1u << CM) - 1);

  const int H = p.H, NH = p.NH, K0 = p.K0;
  const int n_l1 = (H + 255) / 256;             // L1 chunks (N = 256, last may be 128)
  const int n_l2 = (NH == 2) ? H / 128 : 0;     // L2 chunks (N = 128)
  const int kb1 = (K0 + 63) / 64;               // k-blocks of layer 1
  const int l2_stages = H / 128;                // 2 k-blocks per stage

  if (tid == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(smem_u32(&bars->full[i]), 1); mbar_init(smem_u32(&bars->empty[i]), p.mc ? CM : 1); }
    mbar_init(smem_u32(&bars->w3_full), 1); mbar_init(smem_u32(&bars->w3_empty), 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&bars->bias_full[i]), 1); mbar_init(smem_u32(&bars->bias_empty[i]), 128);
      mbar_init(smem_u32(&bars->acc_full[i]), 1); mbar_init(smem_u32(&bars->act1_ready[i]), 128);
      mbar_init(smem_u32(&bars->acc2_full[i]), 1); mbar_init(smem_u32(&bars->act2_ready[i]), 128);
      mbar_init(smem_u32(&bars->l3_full[i]), 1); mbar_init(smem_u32(&bars->l3_drained[i]), 128);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(smem_u32(&bars->tmem_slot));
  if (warp == 0 && lane == 0) { prefetch_tmap(&tmX); prefetch_tmap(&tmW1); prefetch_tmap(&tmW2); prefetch_tmap(&tmW3); }
  fence_before_sync();
  __syncthreads();
  if (CM > 1) cluster_sync();
  fence_after_sync();
  const uint32_t tmem = bars->tmem_slot;

  const long long u0 = ((long long)cluster_id * p.U) / p.NC;
  const long long u1 = ((long long)(cluster_id + 1) * p.U) / p.NC;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    // The whole warp runs the loop with warp-uniform values (so addresses / descriptors live
    // in uniform registers); one elected lane issues the copies.
    uint32_t it = 0;
    uint32_t w3_n = 0, unit_n = 0;
    int ready_upto = -1;
    for (long long u = u0; u < u1; ++u, ++unit_n) {
      const int q = (int)(u / p.ns), s = (int)(u - (long long)q * p.ns);
      const int m0 = (q * CM + (int)rank) * 128;
      wait_sample_ready(p, s, ready_upto, lane);
      {
        const uint32_t b = unit_n & 1;
        mbar_wait(smem_u32(&bars->bias_empty[b]), ((unit_n >> 1) & 1) ^ 1);
        if (elect_one()) {
          mbar_expect_tx(smem_u32(&bars->bias_full[b]), (uint32_t)(p.PB * 4));
          bulk_load(s_bias + b * kBiasFloats * 4, p.bias + (long long)s * p.PB, (uint32_t)(p.PB * 4), smem_u32(&bars->bias_full[b]));
        }
        __syncwarp();
      }
      for (int c = 0; c < n_l1; ++c) {
        const int ncols = min(256, H - c * 256);
        const int nblk = ncols / 64;
        for (int kb = 0; kb < kb1; ++kb, ++it) {
          const uint32_t st = it % kStages, ph = (it / kStages) & 1;
          mbar_wait(smem_u32(&bars->empty[st]), ph ^ 1);
          if (elect_one()) {
            const uint32_t fb = smem_u32(&bars->full[st]);
            const uint32_t base = s_stage + st * kStageBytes;
            if (CM == 1 || !p.mc) {
              // ONE copy for the chunk's (up to) four 64-column weight blocks through the 4-D (64 n, k, n-block, sample) view
              // of the kernel: the issuing thread pays ~95 cycles per copy (tools/tma_rate.cu), and five copies per stage
              // made it, not the MMAs, the pace of layer 1.  The box always spans four n-blocks (zeros / the next chunk's
              // columns beyond this one's; all counted).
              mbar_expect_tx(fb, kXBytes + 4 * kBlkBytes);
              tma_load_2d(base, &tmX, fb, kb * 64, m0);
              tma_load_4d(base + kXBytes, &tmW1, fb, 0, kb * 64, c * 4, s);
            } else {
              mbar_expect_tx(fb, kXBytes + nblk * kBlkBytes);
              tma_load_2d(base, &tmX, fb, kb * 64, m0);
#pragma unroll
              for (int nb = 0; nb < 4; ++nb)
                if (nb < nblk && (nb % CM) == (int)rank)
                  tma_load_3d_mc(base + kXBytes + nb * kBlkBytes, &tmW1, fb, c * 256 + nb * 64, kb * 64, s, kMask);
            }
          }
          __syncwarp();
        }
        if (c == 0) {
          // W3 of this sample (single buffer): the MMA warp releases it after the previous
          // unit's last L3 partial, which it force-issues right after L1 chunk 0.
          mbar_wait(smem_u32(&bars->w3_empty), (w3_n & 1) ^ 1);
          if (elect_one()) {
            const uint32_t wf = smem_u32(&bars->w3_full);
            mbar_expect_tx(wf, (uint32_t)p.w3_bytes);
            bulk_load(s_w3, p.w3 + (long long)s * p.P16, (uint32_t)p.w3_bytes, wf);   // already in core-matrix order
          }
          __syncwarp();
          ++w3_n;
        }
      }
      for (int j = 0; j < n_l2; ++j) {
        for (int sg = 0; sg < l2_stages; ++sg, ++it) {
          const uint32_t st = it % kStages, ph = (it / kStages) & 1;
          mbar_wait(smem_u32(&bars->empty[st]), ph ^ 1);
          if (elect_one()) {
            const uint32_t fb = smem_u32(&bars->full[st]);
            const uint32_t base = s_stage + st * kStageBytes + kXBytes;
            mbar_expect_tx(fb, 4 * kBlkBytes);
            if (CM == 1 || !p.mc) {
              // one copy: [2 n-blocks][128 k][64 n] (n-block stride 16 KB)
              tma_load_4d(base, &tmW2, fb, 0, sg * 128, j * 2, s);
            } else {
#pragma unroll
              for (int bx = 0; bx < 4; ++bx) {  // (k-block kblk, n-block nb)
                const int kblk = bx >> 1, nb = bx & 1;
                const int kabs = sg * 2 + kblk;
                if ((bx % CM) == (int)rank)
                  tma_load_3d_mc(base + bx * kBlkBytes, &tmW2, fb, j * 128 + nb * 64, kabs * 64, s, kMask);
              }
            }
          }
          __syncwarp();
        }
      }
    }
    // producer tail: peers' MMA warps still arrive (multicast commit) on this CTA's `empty`
    // barriers for the last ring pass; do not let the CTA exit before those have landed.
    if (CM > 1 && p.mc) {
      for (int k = 0; k < kStages; ++k, ++it) {
        const uint32_t st = it % kStages, ph = (it / kStages) & 1;
        mbar_wait(smem_u32(&bars->empty[st]), ph ^ 1);
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    // Warp-uniform control flow; one elected lane issues tcgen05.mma / commit.
    // running pipeline-stage state (no div/mod, no descriptor rebuilds in the hot loop)
    uint32_t st_i = 0, st_ph = 0;
    const uint32_t full0 = smem_u32(&bars->full[0]), empty0 = smem_u32(&bars->empty[0]);
    // NB: inside a cluster the shared-window address carries the CTA rank in its upper bits
    // (rank 1: 0x01000400); matrix descriptors take the CTA-local 18-bit offset only.
    const uint32_t a16_0 = (s_stage & 0x3FFFF) >> 4;
    uint32_t st_full = full0, st_empty = empty0, st_a16 = a16_0;   // stage base address >> 4
    auto advance = [&]() {
      if (++st_i == kStages) { st_i = 0; st_ph ^= 1; st_full = full0; st_empty = empty0; st_a16 = a16_0; }
      else { st_full += 8; st_empty += 8; st_a16 += kStageBytes >> 4; }
    };
    constexpr uint32_t kDescHi128 = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO 1024, version 1, SWIZZLE_128B
    uint32_t ph_act1_0 = 0, ph_act1_1 = 0, ph_act2_0 = 0, ph_act2_1 = 0;
    uint32_t l3_issued0 = 0, l3_issued1 = 0, l3_waited0 = 0, l3_waited1 = 0;
    uint32_t w3_n = 0;
    // pending L3 partials: FIFO of at most 2 entries, kept in scalars
    // entry = (buf, a_tmem, d_tmem, w3 byte offset, ksteps, last-of-unit, kind 1|2)
    int pn = 0;
    int pb0 = 0, pk0 = 0, pl0 = 0, pkind0 = 0; uint32_t pa0 = 0, pd0 = 0, pw0 = 0;
    int pb1 = 0, pk1 = 0, pl1 = 0, pkind1 = 0; uint32_t pa1 = 0, pd1 = 0, pw1 = 0;
    const uint32_t idesc_l3 = make_idesc_bf16(128, 16, 0, 1);
    const uint32_t idesc_l2 = make_idesc_bf16(128, 128, 0, 1);
    const uint64_t w3desc0 = make_smem_desc(s_w3, 128, (uint32_t)(H * 16), SW_NONE);

    // try (or force) to issue the oldest pending L3 partial; returns true if one was issued
    auto pump = [&](bool force) -> bool {
      if (pn == 0) return false;
      const uint32_t bar = smem_u32(pkind0 == 1 ? &bars->act1_ready[pb0] : &bars->act2_ready[pb0]);
      uint32_t ph;
      if (pkind0 == 1) ph = pb0 ? ph_act1_1 : ph_act1_0; else ph = pb0 ? ph_act2_1 : ph_act2_0;
      if (force) mbar_wait(bar, ph & 1);
      else if (!mbar_test_wait(bar, ph & 1)) return false;   // must not block: the next stage's MMAs are waiting to be issued
      if (pkind0 == 1) { if (pb0) ++ph_act1_1; else ++ph_act1_0; } else { if (pb0) ++ph_act2_1; else ++ph_act2_0; }
      fence_after_sync();
      DBX_TRACE(140);
      if (elect_one()) {
        const uint64_t bd = w3desc0 + (uint64_t)(pw0 >> 4);
        for (int kk = 0; kk < pk0; ++kk)   // four independent accumulator chains (see drain_l3)
          mma_ts(pd0 + (uint32_t)(kk & 3) * 16, pa0 + kk * 8, bd + (uint64_t)(kk * 16), idesc_l3, kk > 3);
        mma_commit(smem_u32(&bars->l3_full[pb0]));
        if (pl0) mma_commit(smem_u32(&bars->w3_empty));
      }
      __syncwarp();
      if (pb0) ++l3_issued1; else ++l3_issued0;
      pb0 = pb1; pk0 = pk1; pl0 = pl1; pkind0 = pkind1; pa0 = pa1; pd0 = pd1; pw0 = pw1;
      --pn;
      return true;
    };
    auto push = [&](int buf, uint32_t a, uint32_t d, uint32_t woff, int ksteps, int last, int kind) {
      if (pn == 0) { pb0 = buf; pa0 = a; pd0 = d; pw0 = woff; pk0 = ksteps; pl0 = last; pkind0 = kind; }
      else { pb1 = buf; pa1 = a; pd1 = d; pw1 = woff; pk1 = ksteps; pl1 = last; pkind1 = kind; }
      ++pn;
    };
    auto flush_all = [&]() { while (pump(true)) {} };
    // issue every pending partial up to the last one that touches buffer b
    auto flush_buf = [&](int b) {
      if (pn == 2 && pb1 == b) { pump(true); pump(true); }
      else if (pn >= 1 && pb0 == b) pump(true);
    };
    auto wait_drained = [&](int b) {
      if (b == 0) { while (l3_waited0 < l3_issued0) { mbar_wait(smem_u32(&bars->l3_drained[0]), l3_waited0 & 1); ++l3_waited0; } }
      else { while (l3_waited1 < l3_issued1) { mbar_wait(smem_u32(&bars->l3_drained[1]), l3_waited1 & 1); ++l3_waited1; } }
      fence_after_sync();
    };

    for (long long u = u0; u < u1; ++u) {
      // ---------------- layer 1 (SS): X tile x W1 chunk -> region c
      for (int c = 0; c < n_l1; ++c) {
        const int ncols = min(256, H - c * 256);
        const uint32_t idesc = make_idesc_bf16(128, ncols, 0, 1);
        const uint32_t dreg = tmem + c * 256;
        if (c == 1) flush_all();        // everything that still reads region 1 / W3 of the previous unit
        else flush_buf(0);
        wait_drained(c);                // region c is free once its last L3 partial was drained
        DBX_TRACE(100 + c);
        for (int kb = 0; kb < kb1; ++kb) {
          mbar_wait(st_full, st_ph);
          DBX_STAGE_FENCE();
          if (elect_one()) {
            const uint32_t a_lo = st_a16 | (1u << 16);                               // A: X block, K-major
            const uint32_t b_lo = (st_a16 + (kXBytes >> 4)) | ((kBlkBytes >> 4) << 16);  // B: LBO = n-block stride
            if (kb + 1 < kb1 || (K0 & 63) == 0) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                mma_ss_lh(dreg, a_lo + kk * 2, kDescHi128, b_lo + kk * 128, kDescHi128, idesc, (kb | kk) > 0);
            } else {
              const int ksteps = ((K0 & 63) + 15) >> 4;
              for (int kk = 0; kk < ksteps; ++kk)
                mma_ss_lh(dreg, a_lo + kk * 2, kDescHi128, b_lo + kk * 128, kDescHi128, idesc, (kb | kk) > 0);
            }
            if (CM == 1 || !p.mc) mma_commit(st_empty); else mma_commit_mc(st_empty, kMask);
          }
          __syncwarp();
          advance();
          pump(false);
        }
        if (elect_one()) mma_commit(smem_u32(&bars->acc_full[c]));
        __syncwarp();
        DBX_TRACE(110 + c);
        if (c == 0) {
          flush_all();  // releases W3 of the previous unit (the producer waits for it here)
          mbar_wait(smem_u32(&bars->w3_full), w3_n & 1);
          ++w3_n;
        }
        if (NH == 1)  // output layer straight from H1 chunk c
          push(c, tmem + c * 256, tmem + c * 256 + 128, (uint32_t)(c * 256 * 16), ncols / 16, c == n_l1 - 1, 1);
      }
      // ---------------- layer 2 (TS): packed H1 (TMEM) x W2 chunk -> buffer j&1
      for (int j = 0; j < n_l2; ++j) {
        const int b = j & 1;
        const uint32_t dbuf = tmem + b * 256 + 128;
        flush_buf(b);      // buffer b must be free: force its pending L3 partial out ...
        wait_drained(b);   // ... and wait until the epilogue drained it
        DBX_TRACE(120 + j);
        for (int sg = 0; sg < l2_stages; ++sg) {
          // the A operand of k-blocks [2sg, 2sg+1] lives in region (sg*128)/256
          if (j == 0 && (sg & 1) == 0) {  // first touch of that region in this unit
            if ((sg >> 1) == 0) { mbar_wait(smem_u32(&bars->act1_ready[0]), ph_act1_0 & 1); ++ph_act1_0; }
            else { mbar_wait(smem_u32(&bars->act1_ready[1]), ph_act1_1 & 1); ++ph_act1_1; }
          }
          mbar_wait(st_full, st_ph);
          DBX_STAGE_FENCE();
          if (elect_one()) {
            // 128 consecutive k of packed H1: region sg>>1, packed column (sg&1)*64
            const uint32_t a0 = tmem + (uint32_t)(sg >> 1) * 256 + (uint32_t)(sg & 1) * 64;
            if (CM == 1 || !p.mc) {
              const uint32_t b_lo = (st_a16 + (kXBytes >> 4)) | ((2 * kBlkBytes >> 4) << 16);   // [n-block][128 k][64 n]
#pragma unroll
              for (int i = 0; i < 8; ++i)   // consecutive 16-wide k-steps: 8 TMEM columns / 2 KB of the weight block apart
                mma_ts_lh(dbuf, a0 + i * 8, b_lo + i * 128, kDescHi128, idesc_l2, (sg | i) > 0);
            } else {
              const uint32_t b_lo = (st_a16 + (kXBytes >> 4)) | ((kBlkBytes >> 4) << 16);
#pragma unroll
              for (int i = 0; i < 8; ++i)   // i = kblk*4 + kk
                mma_ts_lh(dbuf, a0 + i * 8, b_lo + (i >> 2) * (2 * kBlkBytes / 16) + (i & 3) * 128, kDescHi128, idesc_l2,
                          (sg | i) > 0);
            }
            if (CM == 1 || !p.mc) mma_commit(st_empty); else mma_commit_mc(st_empty, kMask);
          }
          __syncwarp();
          advance();
          pump(false);
        }
        if (elect_one()) mma_commit(smem_u32(&bars->acc2_full[b]));
        __syncwarp();
        DBX_TRACE(130 + j);
        push(b, dbuf, dbuf + 64, (uint32_t)(j * 128 * 16), 8, j == n_l2 - 1, 2);
      }
    }
    flush_all();
  } else {
    // =========================== epilogue (warps 2..5) ===========================
    const int g = warp & 3;                       // TMEM lane group this warp may access
    const int row = g * 32 + lane;
    const uint32_t lane_base = (uint32_t)(g * 32) << 16;
    uint32_t ph_acc[2] = {0, 0}, ph_acc2[2] = {0, 0}, ph_l3[2] = {0, 0};
    uint32_t unit_n = 0;
    float acc1[kMaxC], acc2[kMaxC];
#pragma unroll
    for (int i = 0; i < kMaxC; ++i) { acc1[i] = 0.f; acc2[i] = 0.f; }
    unsigned int cnt = 0;
    int cur_q = (u0 < u1) ? (int)(u0 / p.ns) : -1;
    const int q_first = cur_q;

    auto flush_segment = [&](int q) {
      const int seg = q - q_first;
      const int grow = (q * CM + (int)rank) * 128 + row;
      if (p.sink_kind == 0) {
        float* dst = p.partial + ((((long long)cluster_id * p.maxseg + seg) * CM + rank) * 128 + row) * 32;
#pragma unroll
        for (int i = 0; i < kMaxC; i += 4) {
          *reinterpret_cast<float4*>(dst + i) = make_float4(acc1[i], acc1[i + 1], acc1[i + 2], acc1[i + 3]);
          *reinterpret_cast<float4*>(dst + 16 + i) = make_float4(acc2[i], acc2[i + 1], acc2[i + 2], acc2[i + 3]);
        }
      } else if (grow < p.B && cnt) {
        atomicAdd(p.counts + grow, (unsigned long long)cnt);
      }
#pragma unroll
      for (int i = 0; i < kMaxC; ++i) { acc1[i] = 0.f; acc2[i] = 0.f; }
      cnt = 0;
    };

    // pack one accumulator chunk: fp32 cols [0,ncols) at `src` (+bias, relu) -> bf16x2 cols [0,ncols/2)
    auto pack_chunk = [&](uint32_t src, int ncols, const float* bias) {
      for (int c0 = 0; c0 < ncols; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(src + c0, v);
        wait_ld();
        uint32_t w[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float2 bb = *reinterpret_cast<const float2*>(bias + c0 + 2 * i);
          const float lo = fmaxf(__uint_as_float(v[2 * i]) + bb.x, 0.f);
          const float hi = fmaxf(__uint_as_float(v[2 * i + 1]) + bb.y, 0.f);
          w[i] = pack_bf16x2(lo, hi);
        }
        tmem_st16(src + c0 / 2, w);
      }
      wait_st();
      fence_before_sync();
    };

    for (long long u = u0; u < u1; ++u, ++unit_n) {
      const int q = (int)(u / p.ns), s = (int)(u - (long long)q * p.ns);
      if (q != cur_q) { flush_segment(cur_q); cur_q = q; }
      const uint32_t bb = unit_n & 1;
      const float* bias = bias_smem + bb * kBiasFloats;
      mbar_wait(smem_u32(&bars->bias_full[bb]), (unit_n >> 1) & 1);
      float logit[kMaxC];
#pragma unroll
      for (int i = 0; i < kMaxC; ++i) logit[i] = 0.f;

      auto drain_l3 = [&](int b, uint32_t daddr) {
        mbar_wait(smem_u32(&bars->l3_full[b]), ph_l3[b] & 1);
        ++ph_l3[b];
        fence_after_sync();
        // four partial accumulators of 16 columns each (the output layer's k-steps go round-robin into them: a single
        // dependent chain of N = 16 MMAs runs at the pipe's accumulate latency, 53 cycles each)
        uint32_t v[32], v2[32];
        tmem_ld32(daddr, v);
        tmem_ld32(daddr + 32, v2);
        wait_ld();
        fence_before_sync();
        mbar_arrive(smem_u32(&bars->l3_drained[b]));
#pragma unroll
        for (int i = 0; i < kMaxC; ++i)
          logit[i] += (__uint_as_float(v[i]) + __uint_as_float(v[16 + i])) + (__uint_as_float(v2[i]) + __uint_as_float(v2[16 + i]));
      };

      for (int c = 0; c < n_l1; ++c) {
        const int ncols = min(256, H - c * 256);
        mbar_wait(smem_u32(&bars->acc_full[c]), ph_acc[c] & 1);
        ++ph_acc[c];
        fence_after_sync();
        if (warp == 2) DBX_TRACE(200 + c);
        pack_chunk(tmem + lane_base + c * 256, ncols, bias + p.b1_off + c * 256);
        mbar_arrive(smem_u32(&bars->act1_ready[c]));
        if (warp == 2) DBX_TRACE(210 + c);
        if (NH == 1) drain_l3(c, tmem + lane_base + c * 256 + 128);
      }
      for (int j = 0; j < n_l2; ++j) {
        const int b = j & 1;
        mbar_wait(smem_u32(&bars->acc2_full[b]), ph_acc2[b] & 1);
        ++ph_acc2[b];
        fence_after_sync();
        if (warp == 2) DBX_TRACE(220 + j);
        pack_chunk(tmem + lane_base + b * 256 + 128, 128, bias + p.b2_off + j * 128);
        mbar_arrive(smem_u32(&bars->act2_ready[b]));
        if (warp == 2) DBX_TRACE(230 + j);
        drain_l3(b, tmem + lane_base + b * 256 + 128 + 64);
        if (warp == 2) DBX_TRACE(240 + j);
      }
      // ---------------- output: + b3, softmax / linear, running sums over this CTA's samples
      const float* b3 = bias + p.b3_off;
      float z[kMaxC];
#pragma unroll
      for (int i = 0; i < kMaxC; ++i) z[i] = (i < p.C) ? logit[i] + b3[i] : -INFINITY;
      mbar_arrive(smem_u32(&bars->bias_empty[bb]));
      const int grow = (q * CM + (int)rank) * 128 + row;
      if (p.sink_kind == 0) {
        const float w = p.wts ? p.wts[s] : 1.f;
        if (p.out_kind == DBX_OUT_LOGITS) {
#pragma unroll
          for (int i = 0; i < kMaxC; ++i) if (i < p.C) { acc1[i] = fmaf(w, z[i], acc1[i]); acc2[i] = fmaf(w * z[i], z[i], acc2[i]); }
        } else if (p.act_last == DBX_ACT_SOFTMAX) {
          float mx = z[0];
#pragma unroll
          for (int i = 1; i < kMaxC; ++i) mx = fmaxf(mx, z[i]);
          float e[kMaxC], den = 0.f;
#pragma unroll
          for (int i = 0; i < kMaxC; ++i) { e[i] = (i < p.C) ? expf(z[i] - mx) : 0.f; den += e[i]; }
          const float inv = 1.f / den;
#pragma unroll
          for (int i = 0; i < kMaxC; ++i) { const float pr = e[i] * inv; acc1[i] = fmaf(w, pr, acc1[i]); acc2[i] = fmaf(w * pr, pr, acc2[i]); }
        } else {
#pragma unroll
          for (int i = 0; i < kMaxC; ++i) if (i < p.C) { const float pr = apply_act(p.act_last, z[i]); acc1[i] = fmaf(w, pr, acc1[i]); acc2[i] = fmaf(w * pr, pr, acc2[i]); }
        }
      } else {
        int best = 0; float bv = z[0];
#pragma unroll
        for (int i = 1; i < kMaxC; ++i) if (z[i] > bv) { bv = z[i]; best = i; }
        if (grow < p.B) {
          const int ok = (best == p.labels[grow]);
          if (p.flags) p.flags[(long long)s * p.B + grow] = (uint8_t)ok;
          cnt += ok;
        }
      }
    }
    if (cur_q >= 0) flush_segment(cur_q);
  }

  // teardown: nobody may exit while a peer can still multicast into / arrive on this CTA
  fence_before_sync();
  __syncthreads();
  if (CM > 1) cluster_sync();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

// =======================================================================================
// CTA-pair variant (tcgen05 cta_group::2): the two CTAs of a cluster own rows [0,128) and
// [128,256) of a 256-row unit; ONE tcgen05.mma issued by the leader drives both SMs' tensor
// cores, and each CTA's shared memory holds only HALF of every weight tile (N/2 columns).  That
// halves the per-SM shared-memory operand traffic (8 KB instead of 12 KB per layer-1 MMA), which
// is what limits the single-CTA SS MMA to ~77 % of the tensor peak (tools/mma_rate.cu).
// Same TMEM plan, same epilogue and same work split as fused_mlp_kernel<2>.
// =======================================================================================
// Stage geometry of the CTA-pair kernel.  KD = k-depth of one layer-1 stage.  What bounds the TMA ring is not bytes in
// flight but the PRODUCER THREAD: every stage costs it ~400 cycles of barrier handshake plus ~95 cycles per bulk-tensor
// copy it issues (tools/tma_rate.cu, profiles/r2_tma_rate.log: one thread sustains 54 B/clk/SM with 2 x 16 KB copies per
// 32 KB stage, 41 with 4 x 8 KB, 82-87 with 64 KB stages -- the same for 2, 3 or 6 stages in flight).  With 32 KB stages
// and three copies per stage the layer-1 loop needed a stage per 512 MMA cycles from a thread that could not deliver one
// in under ~700.  So: KD = 128 -> 64 KB stages (3 of them), ONE copy for the X block (box 64k x 128 rows x 2 k-blocks through
// a 3-D view of X) and ONE for the CTA's weight half (box 64n x 128k x 2 n-blocks through a 4-D view of the Keras kernel);
// a layer-2 chunk's whole [64n x 512k] weight half is one stage (two copies).  KD = 64 keeps 6 x 32 KB stages with the
// same merged copies (option FUSED_KD=64, for A/B runs).
template <int KD> struct K2 {
  static constexpr int XB = KD * 256;                  // X block: 128 rows x KD x bf16
  static constexpr int WB = KD * 128;                  // one [64 n x KD k] MN-major weight block
  static constexpr int STAGE = XB + 2 * WB;            // 32 KB / 64 KB
  static constexpr int NST = 196608 / STAGE;           // 6 / 3
  static constexpr int L2K = KD == 128 ? 512 : 128;    // k per layer-2 stage
  static constexpr int L2BOXK = KD == 128 ? 256 : 128; // k rows of one layer-2 copy (box dims are <= 256)
};
constexpr int k2MaxStages = 6;

struct __align__(8) Barriers2 {
  uint64_t full[k2MaxStages], empty[k2MaxStages];
  uint64_t w3_full, w3_empty;
  uint64_t bias_full[2], bias_empty[2];
  uint64_t acc_full[2], act1_ready[2];
  uint64_t acc2_full[2], act2_ready[2];
  uint64_t l3_full[2], l3_drained[2];
  uint32_t tmem_slot;
};

constexpr int k2Threads = 224;   // warp 0 producer, 1 MMA issuer, 2-5 epilogue, 6 second producer (NP = 2)
template <int KD, int NP>
__global__ void __launch_bounds__(k2Threads, 1)
fused_mlp2_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
                  const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmW3,
                  const FusedParams p) {
  using G = K2<KD>;
  constexpr int k2StageBytes = G::STAGE;
  constexpr int CM = 2;
  uint8_t* smem = (uint8_t*)(((uintptr_t)fused_smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t s_stage = smem_u32(smem);
  constexpr int NST = G::NST;
  const uint32_t s_w3 = s_stage + NST * k2StageBytes;
  float* bias_smem = (float*)(smem + NST * k2StageBytes + kW3Bytes / 2);
  const uint32_t s_bias = smem_u32(bias_smem);
  Barriers2* bars = (Barriers2*)((uint8_t*)bias_smem + 2 * kBiasFloats * 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int trace_n = 0;
  const uint32_t rank = cluster_ctarank();
  const bool leader = (rank == 0);
  const int cluster_id = blockIdx.x / CM;
  constexpr uint16_t kBoth = 0x3;

  const int H = p.H, NH = p.NH, K0 = p.K0;
  const int n_l1 = (H + 255) / 256;
  const int n_l2 = (NH == 2) ? H / 128 : 0;
  const int kb1 = (K0 + KD - 1) / KD;
  const int l2_stages = (H + G::L2K - 1) / G::L2K;

  if (tid == 0) {
    for (int i = 0; i < NST; ++i) { mbar_init(smem_u32(&bars->full[i]), 1); mbar_init(smem_u32(&bars->empty[i]), 1); }
    mbar_init(smem_u32(&bars->w3_full), 1); mbar_init(smem_u32(&bars->w3_empty), 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&bars->bias_full[i]), 1); mbar_init(smem_u32(&bars->bias_empty[i]), 128);
      mbar_init(smem_u32(&bars->acc_full[i]), 1); mbar_init(smem_u32(&bars->act1_ready[i]), 8);   // one arrival per epilogue warp of both CTAs
      mbar_init(smem_u32(&bars->acc2_full[i]), 1); mbar_init(smem_u32(&bars->act2_ready[i]), 8);
      mbar_init(smem_u32(&bars->l3_full[i]), 1); mbar_init(smem_u32(&bars->l3_drained[i]), 8);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc2<512>(smem_u32(&bars->tmem_slot));
  if (warp == 0 && lane == 0) { prefetch_tmap(&tmX); prefetch_tmap(&tmW1); prefetch_tmap(&tmW2); prefetch_tmap(&tmW3); }
  fence_before_sync();
  __syncthreads();
  cluster_sync();
  fence_after_sync();
  const uint32_t tmem = bars->tmem_slot;

  // this cluster's units: order 0 = a contiguous range of the [q][s] flattening,
  //                       order 1 = cluster_id, cluster_id + NC, ... of the [s][q] flattening (see fused_mlp4.cuh)
  const long long u0 = p.order ? (long long)cluster_id : ((long long)cluster_id * p.U) / p.NC;
  const long long u1 = p.order ? p.U : ((long long)(cluster_id + 1) * p.U) / p.NC;
  const long long u_step = p.order ? (long long)p.NC : 1;
  auto decode = [&](long long u, int& q, int& s) {
    if (p.order) { s = (int)(u / p.Q); q = (int)(u - (long long)s * p.Q); }
    else { q = (int)(u / p.ns); s = (int)(u - (long long)q * p.ns); }
  };

  if (warp == 0 || (NP == 2 && warp == 6)) {
    // =========================== TMA producer(s) (both CTAs) ===========================
    // Every copy lands in the issuing CTA's smem and signals the LEADER's `full` barrier; only
    // the leader arms it (expect_tx of the pair's total bytes).  NP = 2: warp 0 fills the even stages of the sequence,
    // warp 6 the odd ones (a stage costs the issuing thread ~600 cycles, see K2; two threads keep a ring of 32 KB stages fed).
    const uint32_t pi = warp == 0 ? 0u : 1u;
    uint32_t it = 0, w3_n = 0, unit_n = 0;
    int ready_upto = -1;
    for (long long u = u0; u < u1; u += u_step, ++unit_n) {
      int q, s;
      decode(u, q, s);
      const int m0 = (q * CM + (int)rank) * 128;
      wait_sample_ready(p, s, ready_upto, lane);
      if (pi == 0) {
        const uint32_t b = unit_n & 1;
        mbar_wait(smem_u32(&bars->bias_empty[b]), ((unit_n >> 1) & 1) ^ 1);
        if (elect_one()) {
          mbar_expect_tx(smem_u32(&bars->bias_full[b]), (uint32_t)(p.PB * 4));
          bulk_load(s_bias + b * kBiasFloats * 4, p.bias + (long long)s * p.PB, (uint32_t)(p.PB * 4), smem_u32(&bars->bias_full[b]));
        }
        __syncwarp();
      }
      for (int c = 0; c < n_l1; ++c) {
        const int ncols = min(256, H - c * 256);
        const int nhalf = ncols / 128;            // n-blocks (64 cols) per CTA
        for (int kb = 0; kb < kb1; ++kb, ++it) {
          if (NP == 2 && (it & 1u) != pi) continue;
          const uint32_t st = it % NST, ph = (it / NST) & 1;
          mbar_wait(smem_u32(&bars->empty[st]), ph ^ 1);
          if (elect_one()) {
            const uint32_t fb = smem_u32(&bars->full[st]);
            const uint32_t base = s_stage + st * k2StageBytes;
            // full boxes always: what lies beyond K0 (k) or beyond the kernel's columns is zero-filled and counted
            if (leader) mbar_expect_tx(fb, 2u * (uint32_t)G::STAGE);
            tma2_load_3d(base, &tmX, fb, 0, m0, kb * (KD / 64));
            tma2_load_4d(base + G::XB, &tmW1, fb, 0, kb * KD, c * 4 + (int)rank * nhalf, s);
          }
          __syncwarp();
        }
        if (c == 0 && pi == 0) {
          mbar_wait(smem_u32(&bars->w3_empty), (w3_n & 1) ^ 1);
          if (elect_one()) {
            const uint32_t wf = smem_u32(&bars->w3_full);
            if (leader) mbar_expect_tx(wf, 2u * (uint32_t)(H * 16));
            // this CTA's 8 output columns = core-matrix n-block `rank` of the blocked8 layout
            tma2_load_3d(s_w3, &tmW3, wf, 0, (int)rank * (H / 8), s);
          }
          __syncwarp();
          ++w3_n;
        }
      }
      for (int j = 0; j < n_l2; ++j) {
        for (int sg = 0; sg < l2_stages; ++sg, ++it) {
          if (NP == 2 && (it & 1u) != pi) continue;
          const uint32_t st = it % NST, ph = (it / NST) & 1;
          mbar_wait(smem_u32(&bars->empty[st]), ph ^ 1);
          if (elect_one()) {
            const uint32_t fb = smem_u32(&bars->full[st]);
            const uint32_t base = s_stage + st * k2StageBytes;
            // this CTA's 64 of the chunk's 128 columns, k rows [sg * L2K, ...): one or two [64 n x L2BOXK k] boxes
            const int krem = H - sg * G::L2K;
            const int nops = krem > G::L2BOXK && G::L2K > G::L2BOXK ? 2 : 1;
            if (leader) mbar_expect_tx(fb, 2u * (uint32_t)(nops * G::L2BOXK * 128));
            for (int o = 0; o < nops; ++o)
              tma2_load_4d(base + o * (G::L2BOXK * 128), &tmW2, fb, 0, sg * G::L2K + o * G::L2BOXK, j * 2 + (int)rank, s);
          }
          __syncwarp();
        }
      }
    }
    if (pi == 0)
      for (int k = 0; k < NST; ++k, ++it) {   // tail: the leader's last commits must have landed here
        const uint32_t st = it % NST, ph = (it / NST) & 1;
        mbar_wait(smem_u32(&bars->empty[st]), ph ^ 1);
      }
  } else if (warp == 1) {
    // =========================== MMA issuer (leader CTA only) ===========================
    if (leader) {
      uint32_t st_i = 0, st_ph = 0;
      const uint32_t full0 = smem_u32(&bars->full[0]), empty0 = smem_u32(&bars->empty[0]);
      const uint32_t a16_0 = (s_stage & 0x3FFFF) >> 4;
      uint32_t st_full = full0, st_empty = empty0, st_a16 = a16_0;
      auto advance = [&]() {
        if (++st_i == NST) { st_i = 0; st_ph ^= 1; st_full = full0; st_empty = empty0; st_a16 = a16_0; }
        else { st_full += 8; st_empty += 8; st_a16 += k2StageBytes >> 4; }
      };
      constexpr uint32_t kDescHi128 = (1024u >> 4) | (1u << 14) | (2u << 29);
      uint32_t ph_act1_0 = 0, ph_act1_1 = 0, ph_act2_0 = 0, ph_act2_1 = 0;
      uint32_t l3_issued0 = 0, l3_issued1 = 0, l3_waited0 = 0, l3_waited1 = 0;
      uint32_t w3_n = 0;
      int pn = 0;
      int pb0 = 0, pk0 = 0, pl0 = 0, pkind0 = 0; uint32_t pa0 = 0, pd0 = 0, pw0 = 0;
      int pb1 = 0, pk1 = 0, pl1 = 0, pkind1 = 0; uint32_t pa1 = 0, pd1 = 0, pw1 = 0;
      const uint32_t idesc_l3 = make_idesc_bf16(256, 16, 0, 1);
      const uint32_t idesc_l2 = make_idesc_bf16(256, 128, 0, 1);
      // W3: one 8-column core-matrix block per CTA, k-groups 128 B apart
      const uint64_t w3desc0 = make_smem_desc(s_w3, 128, (uint32_t)(H * 16), SW_NONE);

      auto pump = [&](bool force) -> bool {
        if (pn == 0) return false;
        const uint32_t bar = smem_u32(pkind0 == 1 ? &bars->act1_ready[pb0] : &bars->act2_ready[pb0]);
        uint32_t ph;
        if (pkind0 == 1) ph = pb0 ? ph_act1_1 : ph_act1_0; else ph = pb0 ? ph_act2_1 : ph_act2_0;
        if (force) mbar_wait(bar, ph & 1);
        else if (!mbar_test_wait(bar, ph & 1)) return false;   // must not block: the next stage's MMAs are waiting to be issued
        if (pkind0 == 1) { if (pb0) ++ph_act1_1; else ++ph_act1_0; } else { if (pb0) ++ph_act2_1; else ++ph_act2_0; }
        fence_after_sync();
        DBX_TRACE(140);
        if (elect_one()) {
          const uint64_t bd = w3desc0 + (uint64_t)(pw0 >> 4);
          for (int kk = 0; kk < pk0; ++kk) {   // four independent accumulator chains (see drain_l3)
            const uint64_t d = bd + (uint64_t)(kk * 16);
            mma2_ts_lh(pd0 + (uint32_t)(kk & 3) * 16, pa0 + kk * 8, (uint32_t)d, (uint32_t)(d >> 32), idesc_l3, kk > 3);
          }
          mma2_commit_mc(smem_u32(&bars->l3_full[pb0]), kBoth);
          if (pl0) mma2_commit_mc(smem_u32(&bars->w3_empty), kBoth);
        }
        __syncwarp();
        if (pb0) ++l3_issued1; else ++l3_issued0;
        pb0 = pb1; pk0 = pk1; pl0 = pl1; pkind0 = pkind1; pa0 = pa1; pd0 = pd1; pw0 = pw1;
        --pn;
        return true;
      };
      auto push = [&](int buf, uint32_t a, uint32_t d, uint32_t woff, int ksteps, int last, int kind) {
        if (pn == 0) { pb0 = buf; pa0 = a; pd0 = d; pw0 = woff; pk0 = ksteps; pl0 = last; pkind0 = kind; }
        else { pb1 = buf; pa1 = a; pd1 = d; pw1 = woff; pk1 = ksteps; pl1 = last; pkind1 = kind; }
        ++pn;
      };
      auto flush_all = [&]() { while (pump(true)) {} };
      auto flush_buf = [&](int b) {
        if (pn == 2 && pb1 == b) { pump(true); pump(true); }
        else if (pn >= 1 && pb0 == b) pump(true);
      };
      auto wait_drained = [&](int b) {
        if (b == 0) { while (l3_waited0 < l3_issued0) { mbar_wait(smem_u32(&bars->l3_drained[0]), l3_waited0 & 1); ++l3_waited0; } }
        else { while (l3_waited1 < l3_issued1) { mbar_wait(smem_u32(&bars->l3_drained[1]), l3_waited1 & 1); ++l3_waited1; } }
        fence_after_sync();
      };

      for (long long u = u0; u < u1; u += u_step) {
        for (int c = 0; c < n_l1; ++c) {
          const int ncols = min(256, H - c * 256);
          const uint32_t idesc = make_idesc_bf16(256, ncols, 0, 1);
          const uint32_t dreg = tmem + c * 256;
          if (c == 1) flush_all(); else flush_buf(0);
          wait_drained(c);
          DBX_TRACE(100 + c);
          for (int kb = 0; kb < kb1; ++kb) {
            mbar_wait(st_full, st_ph);
            DBX_STAGE_FENCE();
            if (elect_one()) {
              // A: [64 k x 128 rows] K-major SW128 tiles, 16 KB apart; B: [KD k x 64 n] MN-major SW128 blocks, WB apart
              const uint32_t a_lo = st_a16 | (1u << 16);
              const uint32_t b_lo = (st_a16 + (G::XB >> 4)) | ((G::WB >> 4) << 16);
              const int kleft = K0 - kb * KD;
              if (kleft >= KD) {
#pragma unroll
                for (int kk = 0; kk < KD / 16; ++kk)
                  mma2_ss_lh(dreg, a_lo + (kk >> 2) * (16384 >> 4) + (kk & 3) * 2, kDescHi128, b_lo + kk * 128, kDescHi128, idesc, (kb | kk) > 0);
              } else {
                const int ksteps = (kleft + 15) >> 4;
                for (int kk = 0; kk < ksteps; ++kk)
                  mma2_ss_lh(dreg, a_lo + (kk >> 2) * (16384 >> 4) + (kk & 3) * 2, kDescHi128, b_lo + kk * 128, kDescHi128, idesc, (kb | kk) > 0);
              }
              mma2_commit_mc(st_empty, kBoth);
            }
            __syncwarp();
            advance();
            pump(false);
          }
          if (elect_one()) mma2_commit_mc(smem_u32(&bars->acc_full[c]), kBoth);
          __syncwarp();
          DBX_TRACE(110 + c);
          if (c == 0) {
            flush_all();
            mbar_wait(smem_u32(&bars->w3_full), w3_n & 1);
            ++w3_n;
          }
          if (NH == 1)
            push(c, tmem + c * 256, tmem + c * 256 + 128, (uint32_t)(c * 256 * 16), ncols / 16, c == n_l1 - 1, 1);
        }
        // Layer 2.  tools/mma_rate2.cu: a .ts N=128 pair MMA takes 73.5 cycles when the MMAs are issued back to back,
        // but ANY pause of the issuing thread between groups of 8 (20 cycles are enough) costs ~330 extra cycles per
        // group -- the tensor pipe restarts -- which is what the per-stage bookkeeping (barrier wait, commit, pump)
        // used to do four times per chunk.  So a whole chunk (chunk 0: each half, because its second half needs the
        // epilogue of layer-1 chunk 1) is issued as ONE block: wait for all its TMA stages first, then 32 MMAs and
        // their commits without anything in between.
        for (int j = 0; j < n_l2; ++j) {
          const int b = j & 1;
          const uint32_t dbuf = tmem + b * 256 + 128;
          flush_buf(b);
          wait_drained(b);
          DBX_TRACE(120 + j);
          // k-steps [0, KS) of this chunk; chunk 0 is cut at k-step 16 (the features of layer-1 chunk 1 follow there,
          // and their epilogue may still be running); each piece is issued as ONE block after all its stages have landed
          const int KS = H >> 4;
          constexpr int SPS = G::L2K / 16;   // k-steps per stage
          const int l3_at = (p.l3_at > 0 && p.l3_at < KS) ? p.l3_at : KS + 1;   // k-step at which the previous chunk's output-layer MMAs may be slipped in
          int ks = 0, waited = 0;
          while (ks < KS) {
            int ke = KS;
            if (j == 0) {
              if (ks == 0) {
                mbar_wait(smem_u32(&bars->act1_ready[0]), ph_act1_0 & 1); ++ph_act1_0;
                ke = min(16, KS);
              } else {
                mbar_wait(smem_u32(&bars->act1_ready[1]), ph_act1_1 & 1); ++ph_act1_1;
              }
              fence_after_sync();
            }
            {   // every stage of the piece has landed (normally long ago: the producer runs ahead)
              const int hi = (ke - 1) / SPS;
              for (; waited <= hi; ++waited) {
                uint32_t i2 = st_i + (uint32_t)waited, ph = st_ph;
                while (i2 >= (uint32_t)NST) { i2 -= NST; ph ^= 1; }
                mbar_wait(full0 + 8 * i2, ph);
              }
            }
            // The piece is issued as ONE block by the elected thread: a gap of even ~20 cycles between dependent MMAs costs
            // ~330 cycles (the accumulator chain restarts; splitting the 32 MMAs into 4 groups with a barrier probe between
            // them measured 2350 -> 4000 cycles per chunk).  The pending output-layer MMAs of the PREVIOUS chunk are slipped
            // in ONCE, at k-step `l3_at`, if its packed H2 is there by then (one in-thread non-blocking probe): issued
            // only after this chunk, they, the drain of their 16 columns and with it the start of the chunk after next
            // came ~1300 cycles late.
            bool slipped = false;
            if (elect_one()) {
              int g = ks;
              while (g < ke) {
                const int t = g / SPS, kk0 = g - t * SPS;
                int kk1 = min(SPS, ke - t * SPS);
                uint32_t i2 = st_i + (uint32_t)t;
                while (i2 >= (uint32_t)NST) i2 -= NST;
                const uint32_t b_base = (a16_0 + i2 * (k2StageBytes >> 4)) | ((G::WB >> 4) << 16);
                if (j > 0 && !slipped && pn > 0 && g < l3_at && t * SPS + kk1 > l3_at) kk1 = l3_at - t * SPS;   // stop at the probe point
#pragma unroll 8
                for (int kk = kk0; kk < kk1; ++kk) {
                  const int gg = t * SPS + kk;
                  mma2_ts_lh(dbuf, tmem + (uint32_t)(gg >> 4) * 256 + (uint32_t)(gg & 15) * 8, b_base + kk * 128, kDescHi128, idesc_l2, gg > 0);
                }
                g = t * SPS + kk1;
                if (kk1 == SPS || g == KS) mma2_commit_mc(empty0 + 8 * i2, kBoth);
                if (j > 0 && !slipped && pn > 0 && g == l3_at) {
                  const uint32_t bar = smem_u32(&bars->act2_ready[pb0]);
                  const uint32_t ph = pb0 ? ph_act2_1 : ph_act2_0;
                  if (pkind0 == 2 && mbar_test_wait(bar, ph & 1)) {
                    fence_after_sync();
                    const uint64_t bd = w3desc0 + (uint64_t)(pw0 >> 4);
                    for (int kk = 0; kk < pk0; ++kk) {
                      const uint64_t d = bd + (uint64_t)(kk * 16);
                      mma2_ts_lh(pd0 + (uint32_t)(kk & 3) * 16, pa0 + kk * 8, (uint32_t)d, (uint32_t)(d >> 32), idesc_l3, kk > 3);
                    }
                    mma2_commit_mc(smem_u32(&bars->l3_full[pb0]), kBoth);
                    if (pl0) mma2_commit_mc(smem_u32(&bars->w3_empty), kBoth);
                    slipped = true;
                  }
                }
              }
              if (ke == KS) mma2_commit_mc(smem_u32(&bars->acc2_full[b]), kBoth);
            }
            slipped = __any_sync(0xffffffffu, slipped);
            if (slipped) {   // book-keeping of pump(), warp-uniform
              if (pb0) { ++ph_act2_1; ++l3_issued1; } else { ++ph_act2_0; ++l3_issued0; }
              pb0 = pb1; pk0 = pk1; pl0 = pl1; pkind0 = pkind1; pa0 = pa1; pd0 = pd1; pw0 = pw1;
              --pn;
            }
            __syncwarp();
            ks = ke;
            pump(false);
          }
          for (int t = 0; t < l2_stages; ++t) advance();
          DBX_TRACE(130 + j);
          push(b, dbuf, dbuf + 64, (uint32_t)(j * 128 * 16), 8, j == n_l2 - 1, 2);
        }
      }
      flush_all();
    }
  } else if (warp >= 2 && warp < 6) {
    // =========================== epilogue (warps 2..5, both CTAs) ===========================
    const int g = warp & 3;
    const int row = g * 32 + lane;
    const uint32_t lane_base = (uint32_t)(g * 32) << 16;
    uint32_t ph_acc[2] = {0, 0}, ph_acc2[2] = {0, 0}, ph_l3[2] = {0, 0};
    uint32_t unit_n = 0;
    float acc1[kMaxC], acc2[kMaxC];
#pragma unroll
    for (int i = 0; i < kMaxC; ++i) { acc1[i] = 0.f; acc2[i] = 0.f; }
    unsigned int cnt = 0;
    int cur_q = -1;
    if (u0 < u1) { int s_first; decode(u0, cur_q, s_first); }
    const int q_first = cur_q;
    // order 1: the (cluster, q) slices persist over the chunk launches of one call; finish4_kernel runs once at the end
    unsigned long long visited = (p.order && p.chunk > 0 && p.visited) ? p.visited[cluster_id] : 0ull;

    auto flush_segment = [&](int q) {
      const int seg = q - q_first;
      const int grow = (q * CM + (int)rank) * 128 + row;
      if (p.sink_kind == 0) {
        float* dst = p.partial + ((((long long)cluster_id * p.maxseg + seg) * CM + rank) * 128 + row) * 32;
#pragma unroll
        for (int i = 0; i < kMaxC; i += 4) {
          *reinterpret_cast<float4*>(dst + i) = make_float4(acc1[i], acc1[i + 1], acc1[i + 2], acc1[i + 3]);
          *reinterpret_cast<float4*>(dst + 16 + i) = make_float4(acc2[i], acc2[i + 1], acc2[i + 2], acc2[i + 3]);
        }
      } else if (grow < p.B && cnt) {
        atomicAdd(p.counts + grow, (unsigned long long)cnt);
      }
#pragma unroll
      for (int i = 0; i < kMaxC; ++i) { acc1[i] = 0.f; acc2[i] = 0.f; }
      cnt = 0;
    };
    auto pack_chunk = [&](uint32_t src, int ncols, const float* bias) {
      for (int c0 = 0; c0 < ncols; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(src + c0, v);
        wait_ld();
        uint32_t w[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float2 bb = *reinterpret_cast<const float2*>(bias + c0 + 2 * i);
          const float lo = fmaxf(__uint_as_float(v[2 * i]) + bb.x, 0.f);
          const float hi = fmaxf(__uint_as_float(v[2 * i + 1]) + bb.y, 0.f);
          w[i] = pack_bf16x2(lo, hi);
        }
        tmem_st16(src + c0 / 2, w);
      }
      wait_st();
      fence_before_sync();
    };

    for (long long u = u0; u < u1; u += u_step, ++unit_n) {
      int q, s;
      decode(u, q, s);
      if (!p.order && q != cur_q) { flush_segment(cur_q); cur_q = q; }
      const uint32_t bb = unit_n & 1;
      const float* bias = bias_smem + bb * kBiasFloats;
      mbar_wait(smem_u32(&bars->bias_full[bb]), (unit_n >> 1) & 1);
      float logit[kMaxC];
#pragma unroll
      for (int i = 0; i < kMaxC; ++i) logit[i] = 0.f;

      auto drain_l3 = [&](int b, uint32_t daddr) {
        mbar_wait(smem_u32(&bars->l3_full[b]), ph_l3[b] & 1);
        ++ph_l3[b];
        fence_after_sync();
        // The output layer's k-steps accumulate round-robin into FOUR 16-column accumulators: an N = 16 MMA lasts a few
        // cycles, so a single dependent chain of 8 (16) of them ran at the pipe's accumulate latency (~100 cycles each,
        // ~900 cycles of an in-order tensor pipe per chunk); the four partial sums are added here.
        uint32_t v[32], v2[32];
        tmem_ld32(daddr, v);
        tmem_ld32(daddr + 32, v2);
        wait_ld();
        fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(smem_u32(&bars->l3_drained[b]), 0);   // the leader's MMA warp waits for both CTAs
#pragma unroll
        for (int i = 0; i < kMaxC; ++i)
          logit[i] += (__uint_as_float(v[i]) + __uint_as_float(v[16 + i])) + (__uint_as_float(v2[i]) + __uint_as_float(v2[16 + i]));
      };

      for (int c = 0; c < n_l1; ++c) {
        const int ncols = min(256, H - c * 256);
        mbar_wait(smem_u32(&bars->acc_full[c]), ph_acc[c] & 1);
        ++ph_acc[c];
        fence_after_sync();
        if (warp == 2) DBX_TRACE(200 + c);
        pack_chunk(tmem + lane_base + c * 256, ncols, bias + p.b1_off + c * 256);
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(smem_u32(&bars->act1_ready[c]), 0);
        if (warp == 2) DBX_TRACE(210 + c);
        if (NH == 1) drain_l3(c, tmem + lane_base + c * 256 + 128);
      }
      for (int j = 0; j < n_l2; ++j) {
        const int b = j & 1;
        mbar_wait(smem_u32(&bars->acc2_full[b]), ph_acc2[b] & 1);
        ++ph_acc2[b];
        fence_after_sync();
        if (warp == 2) DBX_TRACE(220 + j);
        pack_chunk(tmem + lane_base + b * 256 + 128, 128, bias + p.b2_off + j * 128);
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(smem_u32(&bars->act2_ready[b]), 0);
        if (warp == 2) DBX_TRACE(230 + j);
        drain_l3(b, tmem + lane_base + b * 256 + 128 + 64);
        if (warp == 2) DBX_TRACE(240 + j);
      }
      const float* b3 = bias + p.b3_off;
      float z[kMaxC];
#pragma unroll
      for (int i = 0; i < kMaxC; ++i) z[i] = (i < p.C) ? logit[i] + b3[i] : -INFINITY;
      mbar_arrive(smem_u32(&bars->bias_empty[bb]));
      const int grow = (q * CM + (int)rank) * 128 + row;
      if (p.sink_kind == 0) {
        const float w = p.wts ? p.wts[s] : 1.f;
        float pr[kMaxC];
        if (p.out_kind == DBX_OUT_LOGITS) {
#pragma unroll
          for (int i = 0; i < kMaxC; ++i) pr[i] = (i < p.C) ? z[i] : 0.f;
        } else if (p.act_last == DBX_ACT_SOFTMAX) {
          float mx = z[0];
#pragma unroll
          for (int i = 1; i < kMaxC; ++i) mx = fmaxf(mx, z[i]);
          float den = 0.f;
#pragma unroll
          for (int i = 0; i < kMaxC; ++i) { pr[i] = (i < p.C) ? expf(z[i] - mx) : 0.f; den += pr[i]; }
          const float inv = 1.f / den;
#pragma unroll
          for (int i = 0; i < kMaxC; ++i) pr[i] *= inv;
        } else {
#pragma unroll
          for (int i = 0; i < kMaxC; ++i) pr[i] = (i < p.C) ? apply_act(p.act_last, z[i]) : 0.f;
        }
        if (!p.order) {
#pragma unroll
          for (int i = 0; i < kMaxC; ++i) { acc1[i] = fmaf(w, pr[i], acc1[i]); acc2[i] = fmaf(w * pr[i], pr[i], acc2[i]); }
        } else {
          // running sums of (cluster, q) live in global memory (L2 resident); only this thread touches this row
          float* dst = p.partial + ((((long long)cluster_id * p.Q + q) * CM + rank) * 128 + row) * 32;
          const bool first = !((visited >> q) & 1ull);
          visited |= 1ull << q;
#pragma unroll
          for (int i = 0; i < kMaxC; i += 4) {
            if (i < p.C) {
              float4 o = first ? make_float4(0.f, 0.f, 0.f, 0.f) : *reinterpret_cast<const float4*>(dst + i);
              o.x = fmaf(w, pr[i], o.x); o.y = fmaf(w, pr[i + 1], o.y); o.z = fmaf(w, pr[i + 2], o.z); o.w = fmaf(w, pr[i + 3], o.w);
              *reinterpret_cast<float4*>(dst + i) = o;
              if (p.want_sum2) {
                float4 o2 = first ? make_float4(0.f, 0.f, 0.f, 0.f) : *reinterpret_cast<const float4*>(dst + 16 + i);
                o2.x = fmaf(w * pr[i], pr[i], o2.x); o2.y = fmaf(w * pr[i + 1], pr[i + 1], o2.y);
                o2.z = fmaf(w * pr[i + 2], pr[i + 2], o2.z); o2.w = fmaf(w * pr[i + 3], pr[i + 3], o2.w);
                *reinterpret_cast<float4*>(dst + 16 + i) = o2;
              }
            }
          }
        }
      } else {
        int best = 0; float bv = z[0];
#pragma unroll
        for (int i = 1; i < kMaxC; ++i) if (z[i] > bv) { bv = z[i]; best = i; }
        if (grow < p.B) {
          const int ok = (best == p.labels[grow]);
          if (p.flags) p.flags[(long long)s * p.B + grow] = (uint8_t)ok;
          if (!p.order) cnt += ok;
          else if (ok) atomicAdd(p.counts + grow, 1ull);
        }
      }
    }
    if (!p.order && cur_q >= 0) flush_segment(cur_q);
    if (p.order && p.visited && tid == 64 && rank == 0) p.visited[cluster_id] = visited;
  }

  fence_before_sync();
  __syncthreads();
  cluster_sync();
  if (warp == 1) tmem_dealloc2<512>(tmem);
}


// order 1: adds the per-(cluster, q) running sums in cluster order.  One thread per (input row, group of 4 floats of the
// row's 32-float slice: sum p in [0,16), sum p^2 in [16,32)); the row's 8 threads read one 128-byte line per cluster, eight
// independent 128-bit loads in flight per thread.  (The element-per-thread form took 27 us for B = 4096 -- 4 % of the
// per-rank step of the 8-GPU strong split.)
__global__ void __launch_bounds__(256)
finish4_kernel(const float* __restrict__ partial, const unsigned long long* __restrict__ visited, int NC, int Q,
               int CM, int B, int C, int overwrite, float* __restrict__ sum1, float* __restrict__ sum2) {
  const long long total = (long long)B * 8;
  for (long long o = blockIdx.x * (long long)blockDim.x + threadIdx.x; o < total; o += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(o >> 3), j = (int)(o & 7);
    const int c0 = (j & 3) * 4;
    const bool second = j >= 4;
    if (c0 >= C || (second && !sum2)) continue;
    const int tile = b / 128, row = b % 128, q = tile / CM, rank = tile % CM;
    const float* base = partial + ((((long long)q) * CM + rank) * 128 + row) * 32 + j * 4;
    const long long kstride = (long long)Q * CM * 128 * 32;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int k = 0;
    for (; k + 8 <= NC; k += 8) {
      float4 v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
        v[i] = ((visited[k + i] >> q) & 1ull) ? __ldcg(reinterpret_cast<const float4*>(base + (k + i) * kstride)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int i = 0; i < 8; ++i) { acc.x += v[i].x; acc.y += v[i].y; acc.z += v[i].z; acc.w += v[i].w; }   // cluster order
    }
    for (; k < NC; ++k) {
      if (!((visited[k] >> q) & 1ull)) continue;
      const float4 v = __ldcg(reinterpret_cast<const float4*>(base + k * kstride));
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    float* dst = (second ? sum2 : sum1) + (long long)b * C + c0;
    const float a[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (c0 + i < C) dst[i] = (overwrite ? 0.f : dst[i]) + a[i];
  }
}

// Adds the per-(cluster, segment) partial sums in sample order.
__global__ void finish_kernel(const float* __restrict__ partial, int NC, int maxseg, int CM, long long U, int ns, int B,
                              int C, int overwrite, float* __restrict__ sum1, float* __restrict__ sum2) {
  const long long total = (long long)B * C;
  for (long long o = blockIdx.x * (long long)blockDim.x + threadIdx.x; o < total; o += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(o / C), c = (int)(o - (long long)b * C);
    const int tile = b / 128, row = b % 128, q = tile / CM, rank = tile % CM;
    float a1 = overwrite ? 0.f : sum1[o];
    float a2 = (sum2 && !overwrite) ? sum2[o] : 0.f;
    const long long lo = (long long)q * ns, hi = lo + ns;
    // clusters whose unit range [k*U/NC, (k+1)*U/NC) meets [lo, hi): a handful, found directly
    int k0 = (int)((lo * NC) / U);
    while (k0 > 0 && ((long long)k0 * U) / NC > lo) --k0;
    for (int k = k0; k < NC; ++k) {
      const long long u0 = ((long long)k * U) / NC, u1 = ((long long)(k + 1) * U) / NC;
      if (u0 >= hi) break;
      if (u1 <= lo || u0 >= u1) continue;
      const int seg = q - (int)(u0 / ns);
      const float* src = partial + ((((long long)k * maxseg + seg) * CM + rank) * 128 + row) * 32;
      a1 += src[c];
      a2 += src[16 + c];
    }
    sum1[o] = a1;
    if (sum2) sum2[o] = a2;
  }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
__global__ void cast_rows_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int64_t rows, int K,
                                      int Kp) {
  const int64_t total = rows * Kp;
  for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = o / Kp; const int k = (int)(o - r * Kp);
    out[o] = __float2bfloat16_rn(k < K ? in[r * K + k] : 0.f);
  }
}
// K % 8 == 0 and 16-byte aligned rows: 8 elements per thread (two 128-bit loads, one 128-bit store); Kp % 8 == 0 always
__global__ void __launch_bounds__(256)
cast_rows_bf16_vec8_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int64_t rows, int K, int Kp) {
  const int g = Kp >> 3;
  const int64_t total = rows * g;
  for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = o / g; const int k = (int)(o - r * g) << 3;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (k < K) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(in + r * K + k));
      const float4 b = __ldg(reinterpret_cast<const float4*>(in + r * K + k + 4));
      v = make_uint4(tc_pack_bf16x2(a.x, a.y), tc_pack_bf16x2(a.z, a.w), tc_pack_bf16x2(b.x, b.y), tc_pack_bf16x2(b.z, b.w));
    }
    *reinterpret_cast<uint4*>(out + r * Kp + k) = v;
  }
}
__global__ void zero_counts_kernel(unsigned long long* p, int64_t n) {
  for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < n; o += (int64_t)gridDim.x * blockDim.x) p[o] = 0ULL;
}

long long* g_trace_buf = nullptr;

// Tensor maps of one (workspace, batch, chunk size, kernel) combination: encoded once, reused by every later call
// with the same key (a repeated predict call encodes nothing).
struct FusedPlan {
  const void* ws = nullptr; size_t ws_bytes = 0;
  int64_t B = -1, ns = -1; int variant = 0, kd = 0;
  CUtensorMap tmX, tmW1[2], tmW2[2], tmW3[2];
};

struct FusedState {
  int num_sms = 0;
  // weight sampling of chunk c+1 runs on a side stream while chunk c is in the fused kernel
  cudaStream_t side = nullptr;
  cudaStream_t main = nullptr;   // high priority: the fused kernels (see fused_mlp_run)
  cudaEvent_t ev_join = nullptr, ev_done = nullptr;
  bool used = false;   // ev_entry holds the end of the previous call's device work
  bool consumed_rec[2] = {false, false};          // ev_consumed[b] was recorded by a previous call ...
  size_t lay_x = 0, lay_w = 0, lay_b = 0;         // ... whose workspace layout was this
  bool lay_stream = false;
  int buf_base = 0;                               // weight buffer of a call's first chunk: alternates from call to call
  cudaEvent_t ev_entry = nullptr, ev_sampled[2] = {nullptr, nullptr}, ev_consumed[2] = {nullptr, nullptr};
  cudaEvent_t ev_in = nullptr;
  FusedPlan plan;
  bool attr_set[2] = {false, false};
};

static bool shape_supported(const dbx_model* m) {
  if (!m->is_mlp) return false;
  std::vector<const Layer*> dl;
  for (auto& L : m->layers) if (L.kind == DBX_DENSE) dl.push_back(&L);
  if (dl.size() != 2 && dl.size() != 3) return false;
  const int H = dl[0]->units;
  if (H % 128 != 0 || H > 512 || H < 128) return false;
  if (dl.size() == 3 && dl[1]->units != H) return false;
  for (size_t i = 0; i + 1 < dl.size(); ++i) if (dl[i]->act != DBX_ACT_RELU) return false;
  if (dl.back()->units > kMaxC) return false;
  if (m->in_feat < 16) return false;
  return true;
}

constexpr int kSmem1 = kStages * kStageBytes + kW3Bytes + 2 * kBiasFloats * 4 + (int)sizeof(Barriers) + 1024;
constexpr int kSmem2 = 196608 + kW3Bytes / 2 + 2 * kBiasFloats * 4 + (int)sizeof(Barriers2) + 1024;   // 3 x 64 KB or 6 x 32 KB of stages

// single-CTA kernel (a lone 128-row tile: a CTA pair would idle its second half)
static int launch_fused1(FusedState* fs, const CUtensorMap& tmX, const CUtensorMap& tmW1, const CUtensorMap& tmW2,
                         const FusedParams& p, cudaStream_t st) {
  if (!fs->attr_set[0]) {
    DBX_CUDA(cudaFuncSetAttribute(fused_mlp_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem1));
    fs->attr_set[0] = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)p.NC);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = (size_t)kSmem1;
  cfg.stream = st;
  DBX_CUDA(cudaLaunchKernelEx(&cfg, fused_mlp_kernel<1>, tmX, tmW1, tmW2, tmW2, p));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

// CTA-pair kernel (cta_group::2): the default
static int launch_fused2(FusedState* fs, int kd, const CUtensorMap& tmX, const CUtensorMap& tmW1, const CUtensorMap& tmW2,
                         const CUtensorMap& tmW3, const FusedParams& p, cudaStream_t st) {
  // The shipped library holds ONE instantiation: 64 KB stages, one producer warp.  The measured-slower geometries (32 KB
  // stages: 39 k cycles per unit against 30.8 k; a second producer warp: no gain) are compiled only with
  // -DDBX_FUSED_VARIANTS (tools/, A/B runs), where options FUSED_KD / FUSED_NP select them.
#ifdef DBX_FUSED_VARIANTS
  const int np = opt().fused_np == 2 ? 2 : 1;
#else
  const int np = 1;
  (void)kd;
#endif
  if (!fs->attr_set[1]) {
    DBX_CUDA(cudaFuncSetAttribute(fused_mlp2_kernel<128, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem2));
#ifdef DBX_FUSED_VARIANTS
    DBX_CUDA(cudaFuncSetAttribute(fused_mlp2_kernel<128, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem2));
    DBX_CUDA(cudaFuncSetAttribute(fused_mlp2_kernel<64, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem2));
    DBX_CUDA(cudaFuncSetAttribute(fused_mlp2_kernel<64, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem2));
#endif
    fs->attr_set[1] = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(p.NC * 2));
  cfg.blockDim = dim3(k2Threads);
  cfg.dynamicSmemBytes = (size_t)kSmem2;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
#ifdef DBX_FUSED_VARIANTS
  if (kd == 128 && np == 1) DBX_CUDA(cudaLaunchKernelEx(&cfg, fused_mlp2_kernel<128, 1>, tmX, tmW1, tmW2, tmW3, p));
  else if (kd == 128) DBX_CUDA(cudaLaunchKernelEx(&cfg, fused_mlp2_kernel<128, 2>, tmX, tmW1, tmW2, tmW3, p));
  else if (np == 1) DBX_CUDA(cudaLaunchKernelEx(&cfg, fused_mlp2_kernel<64, 1>, tmX, tmW1, tmW2, tmW3, p));
  else DBX_CUDA(cudaLaunchKernelEx(&cfg, fused_mlp2_kernel<64, 2>, tmX, tmW1, tmW2, tmW3, p));
#else
  DBX_CUDA(cudaLaunchKernelEx(&cfg, fused_mlp2_kernel<128, 1>, tmX, tmW1, tmW2, tmW3, p));
#endif
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

static FusedState* fused_state_of(dbx_model* m) {
  FusedState* fs = (FusedState*)m->fused_state;
  if (!fs) {
    fs = new FusedState();
    if (cudaDeviceGetAttribute(&fs->num_sms, cudaDevAttrMultiProcessorCount, m->device) != cudaSuccess) { delete fs; return nullptr; }
    if (cudaStreamCreateWithFlags(&fs->side, cudaStreamNonBlocking) != cudaSuccess) { delete fs; return nullptr; }
    int lo = 0, hi = 0;
    if (cudaDeviceGetStreamPriorityRange(&lo, &hi) != cudaSuccess ||
        cudaStreamCreateWithPriority(&fs->main, cudaStreamNonBlocking, hi) != cudaSuccess) { delete fs; return nullptr; }
    cudaEventCreateWithFlags(&fs->ev_entry, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&fs->ev_join, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&fs->ev_done, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&fs->ev_in, cudaEventDisableTiming);
    for (int i = 0; i < 2; ++i) {
      cudaEventCreateWithFlags(&fs->ev_sampled[i], cudaEventDisableTiming);
      cudaEventCreateWithFlags(&fs->ev_consumed[i], cudaEventDisableTiming);
    }
    m->fused_state = fs;
  }
  return fs;
}

static inline int grid_for2(const FusedState* fs, int64_t total, int block = 256) {
  return (int)std::min<int64_t>(std::max<int64_t>(cdiv(total, block), 1), (int64_t)fs->num_sms * 16);
}

static int fused_mlp_run_on(dbx_model* m, FusedState* fs, const float* x_dev, int64_t B, int64_t n, const Source& src,
                            const Sink& sink, cudaStream_t st);

// Stream mode (see fused_mlp_run_on): when the fused kernel is the bulk of the work, i.e. >= 6 row-tile pairs per sample.
static bool stream_mode_for(int64_t B) {
  if (!opt().fused_overlap || !opt().fused_stream || opt().fused_ns != 0) return false;
  int variant = (B <= 128) ? 1 : 2;
  if (opt().fused_kernel == 1 || opt().fused_kernel == 2) variant = (int)opt().fused_kernel;
  const int CM = variant == 2 ? 2 : 1;
  const int64_t Q = cdiv(cdiv(B, 128), CM);
  return Q >= 6 || opt().fused_stream == 2;
}

// The fused kernels run on a HIGH-priority stream of the engine's own, joined with the caller's stream at both ends: the
// sampler of the next chunk shares the SMs with them from a default-priority side stream, and at every chunk boundary the
// next fused launch's CTA pairs then get SM slots ahead of the sampler's queued blocks (option FUSED_PRIO=0: caller's stream).
int fused_mlp_run(dbx_model* m, const float* x_dev, int64_t B, int64_t n, const Source& src, const Sink& sink,
                  cudaStream_t st_user) {
  if (!shape_supported(m)) return 0;
  if (!get_encode_fn()) return 0;
  FusedState* fs = fused_state_of(m);
  if (!fs) return -1;
  // stream mode holds ONE long fused launch per chunk: nothing is queued behind the sampler's blocks at chunk boundaries,
  // so the priority stream buys nothing there and its two cross-stream joins cost ~10 us per call (1.5 % of the per-rank
  // step of the 8-GPU strong split)
  if (!opt().fused_prio || stream_mode_for(B)) return fused_mlp_run_on(m, fs, x_dev, B, n, src, sink, st_user);
  if (cudaEventRecord(fs->ev_join, st_user) != cudaSuccess || cudaStreamWaitEvent(fs->main, fs->ev_join, 0) != cudaSuccess) return -1;
  const int rc = fused_mlp_run_on(m, fs, x_dev, B, n, src, sink, fs->main);
  if (cudaEventRecord(fs->ev_done, fs->main) != cudaSuccess || cudaStreamWaitEvent(st_user, fs->ev_done, 0) != cudaSuccess) return -1;
  return rc;
}

// Samples per fused launch.  256 x Q row-tile pairs keep every cluster busy for large batches.  With one or two row tiles
// per sample (config 5: 128 inputs x 100k samples) a launch of 256 samples is over before its pipeline has filled, so
// long runs use up to 1024 (still >= 4 chunks, so that sampling keeps overlapping).  SHORT calls (n <= 256: the per-rank
// share of a sample-sharded predict) are cut into 2-4 launches so that only the first quarter of the
// sampling is exposed before the first fused launch; each launch keeps >= 4 units per cluster.
static int64_t samples_per_launch(const FusedState* fs, int64_t n, int Q, int CM) {
  if (opt().fused_ns > 0) return std::min<int64_t>(n, opt().fused_ns);
  int64_t ns_max = 256;
  if (Q <= 4 && n >= 2048) ns_max = std::min<int64_t>(1024, (n / 4) / 256 * 256);
  if (n > ns_max) return ns_max;
  const int64_t NC = std::max(1, fs->num_sms / CM);
  const int64_t min_ns = cdiv(4 * NC, Q);                  // >= 4 units per cluster and launch
  int64_t parts = std::min<int64_t>(4, std::max<int64_t>(1, n / std::max<int64_t>(min_ns, 1)));
  return cdiv(n, parts);
}

static int fused_mlp_run_on(dbx_model* m, FusedState* fs, const float* x_dev, int64_t B, int64_t n, const Source& src,
                            const Sink& sink, cudaStream_t st) {
  std::vector<const Layer*> dl;
  for (auto& L : m->layers) if (L.kind == DBX_DENSE) dl.push_back(&L);
  const int NH = (int)dl.size() - 1, H = dl[0]->units, K0 = (int)m->in_feat, C = (int)m->out_feat;
  const Layer* Lout = dl.back();

  // kernel: 2 = CTA pair (cta_group::2), the default; 1 = single-CTA kernel for a lone 128-row tile (option FUSED_KERNEL forces one)
  int variant = (B <= 128) ? 1 : 2;
  if (opt().fused_kernel == 1 || opt().fused_kernel == 2) variant = (int)opt().fused_kernel;
  const bool two_cta = variant == 2;
  const int CM = two_cta ? 2 : 1;
  const int T = (int)cdiv(B, 128);
  const int Q = (int)cdiv(T, CM);
  // stream mode (default): ONE fused launch per chunk of up to 1024 samples runs concurrently with the sampler that feeds
  // it, gated by per-sub-chunk flags (wait_sample_ready) -- a short call (the per-rank share of a sample-sharded predict,
  // SURVEY 8(e)) exposes only its first 16 samples of sampling and has no launch boundaries inside
  const bool overlap = opt().fused_overlap != 0;
  // ... when the fused kernel is the bulk of the work (>= 6 row-tile groups per sample).  With one or two tiles per sample
  // (config 5: 128 inputs) the SAMPLER is the bottleneck, and a fused kernel that sits on every SM for the whole chunk,
  // waiting for samples, leaves it two resident blocks per SM instead of eight: 119 ms against 86 ms for 100k samples.
  const bool stream = stream_mode_for(B);
  const int64_t per_sample_b = (int64_t)m->P16 * 2 + (int64_t)m->PB * 4;
  const int64_t ns = stream ? std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(n, 1024), ((int64_t)3 << 29) / per_sample_b))
                            : samples_per_launch(fs, n, Q, CM);
  constexpr int kMaxSub = 72;
  const int sub0 = (int)std::min<int64_t>(ns, 16);
  const int subc = (int)std::min<int64_t>(128, std::max<int64_t>(32, round_up(cdiv(ns, 8), 4)));

  // ---- workspace: X bf16 | W16 [ns,P16] | B32 [ns,PB] | partials
  const int K0p = (int)round_up(K0, 64);   // whole 64-wide k-blocks: the pad columns hold zeros (the pair kernel reads X through a 3-D view)
  const size_t x_b = round_up((size_t)B * K0p * 2, 1024);
  const size_t w_b = round_up((size_t)ns * m->P16 * 2, 1024);
  const size_t b_b = round_up((size_t)ns * m->PB * 4, 1024);
  const int NCmax = std::max(1, fs->num_sms / CM);
  const int maxseg = (int)cdiv(Q, NCmax) + 2;
  // work order: 1 = [s][q] round-robin (L2-friendly; CTA-pair kernel, needs Q <= 64 for the visited mask), 0 = [q][s] ranges
  const bool order_ok = two_cta && Q <= 64;
  int order = order_ok ? 1 : 0;
  if (opt().fused_order >= 0) order = (order_ok && opt().fused_order != 0) ? 1 : 0;
  const size_t part_b = round_up(std::max((size_t)NCmax * maxseg * CM * 128 * 32 * 4,
                                          order ? (size_t)NCmax * Q * CM * 128 * 32 * 4 + (size_t)NCmax * 8 : (size_t)0), 1024);
  const size_t flag_b = 1024;   // 2 x kMaxSub flags, keeps the 1024-byte alignment of what follows
  const size_t need = x_b + 2 * (w_b + b_b) + part_b + flag_b;
  if (m->fused_ws.bytes < need) {
    if (fs->used) cudaDeviceSynchronize();   // the old workspace may still be in use by a side stream
    if (ensure_ws(m->fused_ws, need)) return -1;
    if (cudaMemsetAsync(m->fused_ws.ptr, 0, m->fused_ws.bytes, st) != cudaSuccess) return -1;  // pad columns and ready flags start at zero
    if (cudaEventRecord(fs->ev_in, st) != cudaSuccess || cudaStreamWaitEvent(fs->side, fs->ev_in, 0) != cudaSuccess) return -1;
    fs->plan.ws = nullptr;
    fs->used = false; fs->consumed_rec[0] = fs->consumed_rec[1] = false;
  }
  // ready flags FIRST: their place must not depend on the call's sizes (they are zero whenever no chunk is in flight)
  char* wp = (char*)m->fused_ws.ptr;
  unsigned int* flags[2] = {(unsigned int*)wp, (unsigned int*)wp + kMaxSub};
  wp += flag_b;
  __nv_bfloat16* X16 = (__nv_bfloat16*)wp; wp += x_b;
  __nv_bfloat16* W16buf[2]; float* B32buf[2];
  for (int i = 0; i < 2; ++i) { W16buf[i] = (__nv_bfloat16*)wp; wp += w_b; B32buf[i] = (float*)wp; wp += b_b; }
  float* partial = (float*)wp;

  if ((K0 & 7) == 0 && ((uintptr_t)x_dev & 15) == 0)
    cast_rows_bf16_vec8_kernel<<<grid_for2(fs, B * (K0p / 8)), 256, 0, st>>>(x_dev, X16, B, K0, K0p);
  else
    cast_rows_bf16_kernel<<<grid_for2(fs, B * K0p), 256, 0, st>>>(x_dev, X16, B, K0, K0p);
  g_launches.fetch_add(1);
  if (cudaGetLastError() != cudaSuccess) { set_err("cast launch failed"); return -1; }

  // ---- tensor maps (cached: the workspace layout is a function of B and ns)
  FusedPlan& pl = fs->plan;
#ifdef DBX_FUSED_VARIANTS
  const int kd = (two_cta && opt().fused_kd == 64) ? 64 : 128;
#else
  const int kd = 128;
#endif
  if (pl.ws != m->fused_ws.ptr || pl.ws_bytes != m->fused_ws.bytes || pl.B != B || pl.ns != ns || pl.variant != variant || pl.kd != kd) {
    if (two_cta) {
      // X [B, K0p] seen as (64 k, rows, k-block): one box = KD/64 K-major SW128 tiles of 128 rows
      uint64_t d[3] = {64, (uint64_t)B, (uint64_t)(K0p / 64)}, s[2] = {(uint64_t)K0p * 2, 128};
      uint32_t b[3] = {64u, 128u, (uint32_t)(kd / 64)};
      if (!make_tmap_bf16(&pl.tmX, X16, 3, d, s, b, CU_TENSOR_MAP_SWIZZLE_128B)) { set_err("tensor map X (3-D view) failed"); return -1; }
    } else {
      uint64_t d[2] = {(uint64_t)K0, (uint64_t)B}, s[1] = {(uint64_t)K0p * 2};
      uint32_t b[2] = {64u, 128};
      if (!make_tmap_bf16(&pl.tmX, X16, 2, d, s, b, CU_TENSOR_MAP_SWIZZLE_128B)) { set_err("tensor map X failed"); return -1; }
    }
    for (int i = 0; i < 2; ++i) {
      // the Keras (fan_in, fan_out) kernel seen as (64 n, k, n-block, sample); a box spans `kbox` k rows of `nblk` n-blocks
      auto make_w4 = [&](CUtensorMap* tm, const Layer* L, uint32_t kbox, uint32_t nblk) {
        uint64_t d[4] = {64, (uint64_t)L->w_rows, (uint64_t)cdiv(L->w_cols_pad, 64), (uint64_t)ns};
        uint64_t sb[3] = {(uint64_t)L->w_cols_pad * 2, 128, (uint64_t)m->P16 * 2};
        uint32_t b[4] = {64, kbox, nblk, 1};
        return make_tmap_bf16(tm, W16buf[i] + L->w16_off, 4, d, sb, b, CU_TENSOR_MAP_SWIZZLE_128B);
      };
      if (two_cta) {
        if (!make_w4(&pl.tmW1[i], dl[0], (uint32_t)kd, 2)) { set_err("tensor map W1 (4-D view) failed"); return -1; }
        if (!make_w4(&pl.tmW2[i], NH == 2 ? dl[1] : dl[0], kd == 128 ? 256u : 128u, 1)) { set_err("tensor map W2 (4-D view) failed"); return -1; }
      } else {
        // single-CTA kernel: all four 64-column blocks of a layer-1 chunk / both blocks x 128 k of a layer-2 stage per copy
        if (!make_w4(&pl.tmW1[i], dl[0], 64, 4)) { set_err("tensor map W1 (4-D view) failed"); return -1; }
        if (!make_w4(&pl.tmW2[i], NH == 2 ? dl[1] : dl[0], 128, 2)) { set_err("tensor map W2 (4-D view) failed"); return -1; }
      }
      // last layer's kernel in blocked8 order viewed as rows of 64 bf16: CTA r of a pair fetches rows [r*H/8, (r+1)*H/8)
      uint64_t d[3] = {64, (uint64_t)(Lout->w_rows * Lout->w_cols_pad / 64), (uint64_t)ns};
      uint64_t sb[2] = {128, (uint64_t)m->P16 * 2};
      uint32_t b[3] = {64, (uint32_t)(H / 8), 1};
      if (!make_tmap_bf16(&pl.tmW3[i], W16buf[i] + Lout->w16_off, 3, d, sb, b, CU_TENSOR_MAP_SWIZZLE_NONE)) { set_err("tensor map W3 failed"); return -1; }
    }
    pl.ws = m->fused_ws.ptr; pl.ws_bytes = m->fused_ws.bytes; pl.B = B; pl.ns = ns; pl.variant = variant; pl.kd = kd;
  }
  // the last layer's kernel goes out in tcgen05 core-matrix order (one bulk copy per sample)
  TensorTable tab = m->table;
  for (int i = 0; i < tab.n; ++i)
    if (!tab.is_bias[i] && tab.src_off[i] == (int32_t)Lout->w_off) tab.blocked8[i] = 1;

  if (sink.kind == 1 && !sink.accumulate) {
    zero_counts_kernel<<<grid_for2(fs, B), 256, 0, st>>>(sink.counts, B);
    g_launches.fetch_add(1);
  }

  // overlap: sampling runs one chunk ahead on the side stream (option FUSED_OVERLAP=0 disables)
  cudaStream_t sst = overlap ? fs->side : st;
  if (overlap) {
    // Sampling does not depend on this call's input, only on the previous call having finished with the weight
    // buffers: wait for the event recorded at the END of the previous call, so the first chunk is drawn while the
    // caller's H2D copy of x (enqueued on `st` just before this call) is still in flight.
    // Same workspace layout as the previous call: each weight buffer only has to wait for ITS last consumer, so the
    // first chunk of this call is drawn while the previous call's last chunk is still in the fused kernel
    // (back-to-back predict calls pipeline; a call that ends with a device->host read does not see this).
    const bool same_layout = fs->used && fs->lay_x == x_b && fs->lay_w == w_b && fs->lay_b == b_b && fs->lay_stream == stream;
    // (each chunk waits for its own buffer's last consumer in sample_chunk)
    if (!same_layout && fs->used && cudaStreamWaitEvent(sst, fs->ev_entry, 0) != cudaSuccess) return -1;
    if (m->ev_posterior && cudaStreamWaitEvent(sst, m->ev_posterior, 0) != cudaSuccess) return -1;
    if (src.stored || src.eps) {   // slab / injected eps may have been produced on `st` just before this call
      if (cudaEventRecord(fs->ev_in, st) != cudaSuccess || cudaStreamWaitEvent(sst, fs->ev_in, 0) != cudaSuccess) return -1;
    }
  }
  const int64_t nchunks = cdiv(n, ns);
  int nc_first = 1;
  // A call's first chunk takes the buffer the PREVIOUS call's first chunk did not: back-to-back single-chunk calls (the
  // per-rank share of a sharded predict) then draw call i+1's weights while call i's fused kernel is still reading its own.
  const int bb = overlap ? fs->buf_base : 0;
  auto sample_chunk = [&](int64_t ci) -> int {
    const int b = (int)((ci + bb) & 1);
    const int64_t c0 = ci * ns;
    const int nc = (int)std::min<int64_t>(ns, n - c0);
    if (overlap && fs->consumed_rec[b] && cudaStreamWaitEvent(sst, fs->ev_consumed[b], 0) != cudaSuccess) return -1;
    if (!stream) {
      if (launch_sample_bf16(m, tab, src, c0, nc, W16buf[b], B32buf[b], sst, m->PB)) return -1;
    } else {
      // sub-chunks in sample order, each followed by its ready flag (the fused kernel of this chunk is launched AFTER these,
      // so even if the two streams shared one hardware queue the flags would be set before it could spin on them)
      int k = 0;
      for (int o = 0; o < nc; ++k) {
        const int cnt = std::min(nc - o, k == 0 ? sub0 : subc);
        if (launch_sample_bf16(m, tab, src, c0 + o, cnt, W16buf[b] + (int64_t)o * m->P16, B32buf[b] + (int64_t)o * m->PB, sst, m->PB)) return -1;
        flag_set_kernel<<<1, 32, 0, sst>>>(flags[b] + k);
        g_launches.fetch_add(1);
        o += cnt;
      }
      if (cudaGetLastError() != cudaSuccess) { set_err("flag launch failed"); return -1; }
    }
    if (overlap && cudaEventRecord(fs->ev_sampled[b], sst) != cudaSuccess) return -1;
    return 0;
  };
  if (sample_chunk(0)) return -1;
  for (int64_t ci = 0; ci < nchunks; ++ci) {
    const int64_t c0 = ci * ns;
    const int bsel = (int)((ci + bb) & 1);
    const int nc = (int)std::min<int64_t>(ns, n - c0);
    if (ci + 1 < nchunks && sample_chunk(ci + 1)) return -1;   // next chunk's weights, concurrently
    if (overlap && !stream && cudaStreamWaitEvent(st, fs->ev_sampled[bsel], 0) != cudaSuccess) return -1;
    __nv_bfloat16* W16 = W16buf[bsel];

    FusedParams p = {};
    p.K0 = K0; p.H = H; p.NH = NH; p.C = C; p.B = (int)B; p.ns = nc; p.CM = CM;
    p.mc = 0;
    p.Q = Q; p.U = (long long)Q * nc;
    p.NC = (int)std::min<long long>(NCmax, std::max<long long>(1, p.U));
    p.maxseg = maxseg;
    p.sink_kind = sink.kind; p.out_kind = sink.out_kind; p.act_last = Lout->act; p.want_sum2 = sink.sum2 != nullptr;
    p.bias = B32buf[bsel]; p.PB = m->PB; p.w3f_off = 0;
    p.b1_off = (int)dl[0]->b32_off; p.b2_off = (int)(NH == 2 ? dl[1]->b32_off : 0); p.b3_off = (int)Lout->b32_off;
    p.wts = sink.wts ? sink.wts + c0 : nullptr;
    p.labels = sink.labels; p.flags = sink.flags ? sink.flags + c0 * B : nullptr; p.counts = sink.counts;
    p.partial = partial;
    p.w3 = W16 + Lout->w16_off; p.P16 = m->P16; p.w3_bytes = (int)(Lout->w_rows * Lout->w_cols_pad * 2);
    p.chunk = (int)ci; p.order = order; p.visited = (unsigned long long*)((char*)partial + (size_t)NCmax * Q * CM * 128 * 32 * 4);
    p.trace = nullptr; p.trace_only = (unsigned)opt().fused_trace_only;
    p.l3_at = (int)opt().fused_l3at;
    p.ready = stream ? flags[bsel] : nullptr; p.sub0 = sub0; p.subc = subc;
    if (opt().fused_trace) {
      static long long* tb = nullptr;
      if (!tb) cudaMalloc(&tb, 8 * 31208);
      cudaMemsetAsync(tb, 0, 8 * 31208, st);
      p.trace = tb; g_trace_buf = tb;
    }
    prof_begin(st);
    const int rc = two_cta ? launch_fused2(fs, kd, pl.tmX, pl.tmW1[bsel], pl.tmW2[bsel], pl.tmW3[bsel], p, st)
                           : launch_fused1(fs, pl.tmX, pl.tmW1[bsel], pl.tmW2[bsel], p, st);
    if (rc) return -1;
    prof_end(st, (double)m->flops * (double)B * nc);
    if (sink.kind == 0) {
      const int overwrite = (c0 == 0 && !sink.accumulate) ? 1 : 0;
      if (order) {
        if (ci == 0) nc_first = p.NC;
        if (ci + 1 == nchunks) {   // the slices carried the sums of every chunk
          finish4_kernel<<<grid_for2(fs, B * 8), 256, 0, st>>>(partial, p.visited, nc_first, Q, CM, (int)B, C, sink.accumulate ? 0 : 1,
                                                               sink.sum1, sink.sum2);
          g_launches.fetch_add(1);
        }
      } else {
        finish_kernel<<<grid_for2(fs, B * C), 256, 0, st>>>(partial, p.NC, maxseg, CM, p.U, nc, (int)B, C, overwrite, sink.sum1,
                                                             sink.sum2);
        g_launches.fetch_add(1);
      }
      if (cudaGetLastError() != cudaSuccess) { set_err("finish launch failed"); return -1; }
    }
    if (overlap) {
      if (stream && cudaMemsetAsync(flags[bsel], 0, kMaxSub * sizeof(unsigned int), st) != cudaSuccess) return -1;   // for the buffer's next user
      if (cudaEventRecord(fs->ev_consumed[bsel], st) != cudaSuccess) return -1;
      fs->consumed_rec[bsel] = true;
    }
  }
  if (cudaEventRecord(fs->ev_entry, st) != cudaSuccess) return -1;
  fs->used = true;
  fs->lay_x = x_b; fs->lay_w = w_b; fs->lay_b = b_b; fs->lay_stream = stream;
  if (overlap) fs->buf_base = (int)((bb + nchunks) & 1);   // the buffer after this call's last one
  if (!overlap) fs->consumed_rec[0] = fs->consumed_rec[1] = false;
  return 1;
}

// debug: copy the timeline of the last traced launch to the host (count, then (id, clock) pairs)
extern "C" __attribute__((visibility("default"))) int dbx_debug_trace(long long* out, int max_pairs) {
  if (!g_trace_buf) return 0;
  cudaDeviceSynchronize();
  static long long host[31208];
  cudaMemcpy(host, g_trace_buf, sizeof host, cudaMemcpyDeviceToHost);
  int cnt = 0;
  for (int w = 0; w < 6; ++w)
    for (int i = 0; i < 2600 && cnt < max_pairs; ++i) {
      const long long* e = host + 1 + w * 5200 + 2 * i;
      if (e[1] == 0) break;
      out[2 * cnt] = e[0]; out[2 * cnt + 1] = e[1]; ++cnt;
    }
  return cnt;
}

void fused_mlp_free(dbx_model* m) {
  if (m->fused_state) {
    FusedState* fs = (FusedState*)m->fused_state;
    if (fs->side) cudaStreamDestroy(fs->side);
    if (fs->ev_entry) cudaEventDestroy(fs->ev_entry);
    if (fs->main) cudaStreamDestroy(fs->main);
    if (fs->ev_join) cudaEventDestroy(fs->ev_join);
    if (fs->ev_done) cudaEventDestroy(fs->ev_done);
    if (fs->ev_in) cudaEventDestroy(fs->ev_in);
    for (int i = 0; i < 2; ++i) { if (fs->ev_sampled[i]) cudaEventDestroy(fs->ev_sampled[i]); if (fs->ev_consumed[i]) cudaEventDestroy(fs->ev_consumed[i]); }
    delete fs; m->fused_state = nullptr;
  }
}

}  // namespace dbx
