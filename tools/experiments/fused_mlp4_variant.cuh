// NOT BUILT.  Round-1 experiment kept for the record (measured negative result): CTA-pair fused
// MLP kernel with all of H1 in SHARED memory and N=256 SS layer-2 MMAs -- 17 % slower than fused_mlp2_kernel.  It was
// #included by deepbayes_b200/csrc/fused_mlp.cu until round 2; finish4_kernel now lives in that file.
// CTA-pair kernel, version 4 (included by fused_mlp.cu): 784-H-H-C MLPs with H = 512.
//
// What limited version 2 (profiles/r1_timeline_pair_kernel.log): the hidden activations H1 were
// packed to bf16 IN TMEM and layer 2 read them back as the A operand (tcgen05.mma .ts).  TMEM has
// ONE 64 B/clk read port per SM, shared by the A-operand reads and the epilogue's tcgen05.ld, and
// TMEM capacity (512 columns = H1 packed + two 128-column accumulators) forced N = 128 layer-2
// MMAs, whose A reads alone saturate that port: layer 2 ran at half the tensor rate.
//
// Here H1 goes to SHARED memory instead (bf16, K-major SWIZZLE_128B, 128 KB per CTA), so
//   * layer 2 is an SS MMA with N = 256 like layer 1 (8 KB of smem operands per 128-cycle MMA),
//   * TMEM holds only accumulators: R0 = [0,256), R1 = [256,512), each used by a layer-1 chunk,
//     then by a layer-2 chunk, then (packed in place) as the A operand of the tiny output layer,
//   * every epilogue overlaps an MMA phase:
//        L1a->R0 | L1b->R1 || E1(R0)->smem | L2a->R0 (k<256 first) || E1(R1)->smem |
//        L2b->R1 || E2(R0) + L3(R0) | next unit's L1a->R0 || E2(R1) + L3(R1) | ...
// The price is shared memory: 128 KB for H1 leaves 80 KB of TMA stages, so a stage is a 32-wide
// k-block (X [128 x 32] SWIZZLE_64B + this CTA's [32 x 128] half of the weight block = 16 KB).
//
// Work order (p.order):
//   0  units flattened [tile-group q][sample s], one contiguous range per cluster, running sums
//      in registers (as versions 1-3);
//   1  units flattened [sample s][tile-group q], dealt round-robin to the clusters, so that at any
//      moment the whole GPU works on ~NC/Q consecutive samples and every weight tile is fetched
//      from HBM once and then served by L2 to the other tile groups.  The running sums live in
//      a per-(cluster, q) slice of `partial` (read-modify-write by the owning thread, L2
//      resident); finish4_kernel adds the slices in cluster order -- still no float atomics.
#if 0

constexpr int k4Stages = 5;
constexpr int k4StageBytes = 16384;   // X block 8 KB + this CTA's half of the weight block 8 KB
constexpr int k4XBytes = 8192;
constexpr int k4BlkBytes = 4096;      // one [64 n x 32 k] MN-major weight box
constexpr int k4H = 512;
constexpr int k4H1Bytes = 128 * k4H * 2;

struct __align__(8) Barriers4 {
  uint64_t full[k4Stages], empty[k4Stages];
  uint64_t w3_full, w3_empty;
  uint64_t bias_full[2], bias_empty[2];
  uint64_t acc_full[2], act1_ready[2];
  uint64_t acc2_full[2], act2_ready[2];
  uint64_t l3_full[2], l3_drained[2];
  uint32_t tmem_slot;
};

constexpr int k4SmemBytes = k4Stages * k4StageBytes + k4H1Bytes + kW3Bytes / 2 + 2 * kBiasFloats * 4 + (int)sizeof(Barriers4) + 1024;

__global__ void __launch_bounds__(kThreads, 1)
fused_mlp4_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
                  const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmW3,
                  const FusedParams p) {
  constexpr int CM = 2;
  constexpr int H = k4H;
  uint8_t* smem = (uint8_t*)(((uintptr_t)fused_smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t s_stage = smem_u32(smem);
  uint8_t* h1_smem = smem + k4Stages * k4StageBytes;
  const uint32_t s_h1 = smem_u32(h1_smem);
  const uint32_t s_w3 = s_h1 + k4H1Bytes;
  float* bias_smem = (float*)(h1_smem + k4H1Bytes + kW3Bytes / 2);
  const uint32_t s_bias = smem_u32(bias_smem);
  Barriers4* bars = (Barriers4*)((uint8_t*)bias_smem + 2 * kBiasFloats * 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int trace_n = 0;
  const uint32_t rank = cluster_ctarank();
  const bool leader = (rank == 0);
  const int cluster_id = blockIdx.x / CM;
  constexpr uint16_t kBoth = 0x3;

  const int K0 = p.K0;
  const int kb1 = (K0 + 31) / 32;
  constexpr int kL2Stages = H / 32;

  if (tid == 0) {
    for (int i = 0; i < k4Stages; ++i) { mbar_init(smem_u32(&bars->full[i]), 1); mbar_init(smem_u32(&bars->empty[i]), 1); }
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

  // this cluster's units: order 0 = [u_begin, u_end) step 1 of the [q][s] flattening,
  //                       order 1 = cluster_id, cluster_id + NC, ... of the [s][q] flattening
  const long long u_begin = p.order ? (long long)cluster_id : ((long long)cluster_id * p.U) / p.NC;
  const long long u_end = p.order ? p.U : ((long long)(cluster_id + 1) * p.U) / p.NC;
  const long long u_step = p.order ? (long long)p.NC : 1;
  auto decode = [&](long long u, int& q, int& s) {
    if (p.order) { s = (int)(u / p.Q); q = (int)(u - (long long)s * p.Q); }
    else { q = (int)(u / p.ns); s = (int)(u - (long long)q * p.ns); }
  };

  if (warp == 0) {
    // =========================== TMA producer (both CTAs) ===========================
    uint32_t it = 0, w3_n = 0, unit_n = 0;
    for (long long u = u_begin; u < u_end; u += u_step, ++unit_n) {
      int q, s;
      decode(u, q, s);
      const int m0 = (q * CM + (int)rank) * 128;
      {
        const uint32_t b = unit_n & 1;
        mbar_wait(smem_u32(&bars->bias_empty[b]), ((unit_n >> 1) & 1) ^ 1);
        if (elect_one()) {
          mbar_expect_tx(smem_u32(&bars->bias_full[b]), (uint32_t)(p.PB * 4));
          bulk_load(s_bias + b * kBiasFloats * 4, p.bias + (long long)s * p.PB, (uint32_t)(p.PB * 4), smem_u32(&bars->bias_full[b]));
        }
        __syncwarp();
      }
      for (int c = 0; c < 2; ++c) {
        for (int kb = 0; kb < kb1; ++kb, ++it) {
          const uint32_t st = it % k4Stages, ph = (it / k4Stages) & 1;
          mbar_wait(smem_u32(&bars->empty[st]), ph ^ 1);
          if (elect_one()) {
            const uint32_t fb = smem_u32(&bars->full[st]);
            const uint32_t base = s_stage + st * k4StageBytes;
            if (leader) mbar_expect_tx(fb, 2u * (uint32_t)k4StageBytes);
            tma2_load_2d(base, &tmX, fb, kb * 32, m0);
#pragma unroll
            for (int nb = 0; nb < 2; ++nb)
              tma2_load_3d(base + k4XBytes + nb * k4BlkBytes, &tmW1, fb, c * 256 + ((int)rank * 2 + nb) * 64, kb * 32, s);
          }
          __syncwarp();
        }
        if (c == 0) {
          mbar_wait(smem_u32(&bars->w3_empty), (w3_n & 1) ^ 1);
          if (elect_one()) {
            const uint32_t wf = smem_u32(&bars->w3_full);
            if (leader) mbar_expect_tx(wf, 2u * (uint32_t)(H * 16));
            tma2_load_3d(s_w3, &tmW3, wf, 0, (int)rank * (H / 8), s);
          }
          __syncwarp();
          ++w3_n;
        }
      }
      for (int j = 0; j < 2; ++j) {
        for (int kb = 0; kb < kL2Stages; ++kb, ++it) {
          const uint32_t st = it % k4Stages, ph = (it / k4Stages) & 1;
          mbar_wait(smem_u32(&bars->empty[st]), ph ^ 1);
          if (elect_one()) {
            const uint32_t fb = smem_u32(&bars->full[st]);
            const uint32_t base = s_stage + st * k4StageBytes + k4XBytes;
            if (leader) mbar_expect_tx(fb, 2u * 2u * k4BlkBytes);
#pragma unroll
            for (int nb = 0; nb < 2; ++nb)
              tma2_load_3d(base + nb * k4BlkBytes, &tmW2, fb, j * 256 + ((int)rank * 2 + nb) * 64, kb * 32, s);
          }
          __syncwarp();
        }
      }
    }
    for (int k = 0; k < k4Stages; ++k, ++it) {   // tail: the leader's last commits must have landed here
      const uint32_t st = it % k4Stages, ph = (it / k4Stages) & 1;
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
        if (++st_i == k4Stages) { st_i = 0; st_ph ^= 1; st_full = full0; st_empty = empty0; st_a16 = a16_0; }
        else { st_full += 8; st_empty += 8; st_a16 += k4StageBytes >> 4; }
      };
      constexpr uint32_t kDescHi128 = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO 1024, SWIZZLE_128B
      constexpr uint32_t kDescHi64 = (512u >> 4) | (1u << 14) | (4u << 29);     // SBO 512, SWIZZLE_64B
      const uint32_t h1_16 = (s_h1 & 0x3FFFF) >> 4;
      uint32_t ph_act1_0 = 0, ph_act1_1 = 0, ph_act2_0 = 0, ph_act2_1 = 0;
      uint32_t l3_issued0 = 0, l3_issued1 = 0, l3_waited0 = 0, l3_waited1 = 0;
      uint32_t w3_n = 0;
      int pn = 0;              // pending output-layer jobs: buffer ids, oldest first
      int pb0 = 0, pb1 = 0;
      constexpr uint32_t idesc = make_idesc_bf16(256, 256, 0, 1);
      constexpr uint32_t idesc_l3 = make_idesc_bf16(256, 16, 0, 1);
      const uint64_t w3desc0 = make_smem_desc(s_w3, 128, (uint32_t)(H * 16), SW_NONE);

      // output layer of buffer b: A = relu(H2) chunk packed in place (128 columns = 256 k), D = 16 columns behind it
      auto pump = [&](bool force) -> bool {
        if (pn == 0) return false;
        const int b = pb0;
        const uint32_t bar = smem_u32(&bars->act2_ready[b]);
        const uint32_t ph = b ? ph_act2_1 : ph_act2_0;
        if (force) mbar_wait(bar, ph & 1);
        else if (!mbar_test_wait(bar, ph & 1)) return false;   // must not block: the next stage's MMAs are waiting to be issued
        if (b) ++ph_act2_1; else ++ph_act2_0;
        fence_after_sync();
        DBX_TRACE(140);
        if (elect_one()) {
          const uint64_t bd = w3desc0 + (uint64_t)((b * 256 * 16) >> 4);
          const uint32_t a = tmem + b * 256, d = a + 128;
#pragma unroll
          for (int kk = 0; kk < 16; ++kk) {
            const uint64_t dd = bd + (uint64_t)(kk * 16);
            mma2_ts_lh(d, a + kk * 8, (uint32_t)dd, (uint32_t)(dd >> 32), idesc_l3, kk > 0);
          }
          mma2_commit_mc(smem_u32(&bars->l3_full[b]), kBoth);
          if (b == 1) mma2_commit_mc(smem_u32(&bars->w3_empty), kBoth);
        }
        __syncwarp();
        if (b) ++l3_issued1; else ++l3_issued0;
        pb0 = pb1;
        --pn;
        return true;
      };
      auto push = [&](int b) { if (pn == 0) pb0 = b; else pb1 = b; ++pn; };
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

      for (long long u = u_begin; u < u_end; u += u_step) {
        // ---- layer 1: two N=256 chunks, A = X block (SW64), B = W1 block
        for (int c = 0; c < 2; ++c) {
          const uint32_t dreg = tmem + c * 256;
          if (c == 1) flush_all(); else flush_buf(0);
          wait_drained(c);
          DBX_TRACE(100 + c);
          for (int kb = 0; kb < kb1; ++kb) {
            mbar_wait(st_full, st_ph);
            DBX_STAGE_FENCE();
            if (elect_one()) {
              const uint32_t a_lo = st_a16 | (1u << 16);
              const uint32_t b_lo = (st_a16 + (k4XBytes >> 4)) | ((k4BlkBytes >> 4) << 16);
              mma2_ss_lh(dreg, a_lo, kDescHi64, b_lo, kDescHi128, idesc, kb > 0);
              if (kb + 1 < kb1 || (K0 & 31) == 0 || (K0 & 31) > 16)
                mma2_ss_lh(dreg, a_lo + 2, kDescHi64, b_lo + 128, kDescHi128, idesc, 1);
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
        }
        // ---- layer 2: two N=256 chunks, A = H1 from shared memory (K-major SW128), B = W2 block
        for (int j = 0; j < 2; ++j) {
          const uint32_t dreg = tmem + j * 256;
          DBX_TRACE(120 + j);
          for (int kb = 0; kb < kL2Stages; ++kb) {
            if (j == 0 && kb == 0) { mbar_wait_cluster(smem_u32(&bars->act1_ready[0]), ph_act1_0 & 1); ++ph_act1_0; }
            if (j == 0 && kb == kL2Stages / 2) { mbar_wait_cluster(smem_u32(&bars->act1_ready[1]), ph_act1_1 & 1); ++ph_act1_1; }
            mbar_wait(st_full, st_ph);
            DBX_STAGE_FENCE();
            if (elect_one()) {
              // 32-wide k-block kb = half (kb & 1) of the 64-wide H1 block kb >> 1
              const uint32_t a_lo = (h1_16 + (uint32_t)(kb >> 1) * (16384u >> 4) + (uint32_t)(kb & 1) * 4u) | (1u << 16);
              const uint32_t b_lo = (st_a16 + (k4XBytes >> 4)) | ((k4BlkBytes >> 4) << 16);
              mma2_ss_lh(dreg, a_lo, kDescHi128, b_lo, kDescHi128, idesc, kb > 0);
              mma2_ss_lh(dreg, a_lo + 2, kDescHi128, b_lo + 128, kDescHi128, idesc, 1);
              mma2_commit_mc(st_empty, kBoth);
            }
            __syncwarp();
            advance();
            pump(false);
          }
          if (elect_one()) mma2_commit_mc(smem_u32(&bars->acc2_full[j]), kBoth);
          __syncwarp();
          DBX_TRACE(130 + j);
          push(j);
        }
      }
      flush_all();
    }
  } else {
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
    int cur_q = -1, q_first = -1;
    unsigned long long visited = (p.order && p.chunk > 0 && p.visited) ? p.visited[cluster_id] : 0ull;   // slices persist over a call's launches
    if (u_begin < u_end) { int s0; decode(u_begin, cur_q, s0); q_first = cur_q; }
    // this thread's row inside an H1 k-block: 8-row groups of 1024 B, 128 B per row, 16-byte chunks XOR-swizzled by (row & 7)
    uint8_t* h1_row = h1_smem + (row >> 3) * 1024 + (row & 7) * 128;
    const int sw = row & 7;

    auto flush_segment = [&](int q) {   // order 0 only
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
    // relu(acc + b) -> bf16 -> this row of the H1 operand in shared memory
    auto store_h1 = [&](uint32_t src, int kcol0, const float* bias) {
      for (int c0 = 0; c0 < 256; c0 += 32) {
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
        const int k = kcol0 + c0;
        uint8_t* blk = h1_row + (k >> 6) * 16384;
        const int j0 = (k & 63) >> 3;            // first 16-byte chunk (0 or 4)
#pragma unroll
        for (int i = 0; i < 4; ++i)
          *reinterpret_cast<uint4*>(blk + (((j0 + i) ^ sw) << 4)) = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
      }
      fence_proxy_async_smem();   // generic-proxy stores -> visible to the tensor core's async-proxy reads
      fence_before_sync();
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

    for (long long u = u_begin; u < u_end; u += u_step, ++unit_n) {
      int q, s;
      decode(u, q, s);
      if (!p.order && q != cur_q) { flush_segment(cur_q); cur_q = q; }
      const uint32_t bb = unit_n & 1;
      const float* bias = bias_smem + bb * kBiasFloats;
      mbar_wait(smem_u32(&bars->bias_full[bb]), (unit_n >> 1) & 1);
      float logit[kMaxC];
#pragma unroll
      for (int i = 0; i < kMaxC; ++i) logit[i] = 0.f;

      for (int c = 0; c < 2; ++c) {
        mbar_wait(smem_u32(&bars->acc_full[c]), ph_acc[c] & 1);
        ++ph_acc[c];
        fence_after_sync();
        if (warp == 2) DBX_TRACE(200 + c);
        store_h1(tmem + lane_base + c * 256, c * 256, bias + p.b1_off + c * 256);
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster_release(smem_u32(&bars->act1_ready[c]), 0);
        if (warp == 2) DBX_TRACE(210 + c);
      }
      for (int j = 0; j < 2; ++j) {
        mbar_wait(smem_u32(&bars->acc2_full[j]), ph_acc2[j] & 1);
        ++ph_acc2[j];
        fence_after_sync();
        if (warp == 2) DBX_TRACE(220 + j);
        pack_chunk(tmem + lane_base + j * 256, 256, bias + p.b2_off + j * 256);
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(smem_u32(&bars->act2_ready[j]), 0);
        if (warp == 2) DBX_TRACE(230 + j);
        mbar_wait(smem_u32(&bars->l3_full[j]), ph_l3[j] & 1);
        ++ph_l3[j];
        fence_after_sync();
        uint32_t v[16];
        tmem_ld16(tmem + lane_base + j * 256 + 128, v);
        wait_ld();
        fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(smem_u32(&bars->l3_drained[j]), 0);
#pragma unroll
        for (int i = 0; i < kMaxC; ++i) logit[i] += __uint_as_float(v[i]);
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
          // running sums of (cluster, q) live in global memory; only this thread touches this row
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

// order 1: adds the per-(cluster, q) running sums in cluster order
__global__ void finish4_kernel(const float* __restrict__ partial, const unsigned long long* __restrict__ visited, int NC, int Q,
                               int CM, int B, int C, int overwrite, float* __restrict__ sum1, float* __restrict__ sum2) {
  const long long total = (long long)B * C;
  for (long long o = blockIdx.x * (long long)blockDim.x + threadIdx.x; o < total; o += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(o / C), c = (int)(o - (long long)b * C);
    const int tile = b / 128, row = b % 128, q = tile / CM, rank = tile % CM;
    float a1 = overwrite ? 0.f : sum1[o];
    float a2 = (sum2 && !overwrite) ? sum2[o] : 0.f;
    for (int k = 0; k < NC; ++k) {
      if (!((visited[k] >> q) & 1ull)) continue;
      const float* src = partial + ((((long long)k * Q + q) * CM + rank) * 128 + row) * 32;
      a1 += src[c];
      if (sum2) a2 += src[16 + c];
    }
    sum1[o] = a1;
    if (sum2) sum2[o] = a2;
  }
}

#endif
