// NOT BUILT.  Round-1 experiment kept for the record (measured negative result):
// CTA-pair fused MLP kernel with the output layer on FFMA in two epilogue warpgroups -- 25 % slower than
// fused_mlp2_kernel (broadcasting W3 to every row-thread is shared-memory-bandwidth bound).  It was part of
// deepbayes_b200/csrc/fused_mlp.cu until round 2 and compiled against that file's helpers.
#if 0
// =======================================================================================
// CTA-pair kernel, version 3: as fused_mlp2_kernel, but the tiny output layer (C <= 12) leaves
// the tensor pipe.  Measured (tools/mma_rate.cu, timeline traces): a tcgen05.mma costs the
// issuing thread >= ~63 cycles whatever its N, so the 32 N=16 output-layer MMAs per unit took
// as long as a quarter of layer 2 and sat on the critical path together with their drain
// round-trips.  Here the epilogue applies W3 with FFMA while it already holds relu(H2) in
// registers: H2 is never packed or written back, and the MMA warp only ever issues full-rate
// layer-1/2 MMAs.  Two epilogue warpgroups (8 warps) split the chunks (WG0: L1 chunk 0 + L2
// chunks 0,2; WG1: L1 chunk 1 + L2 chunks 1,3) and hand the partial logits over in smem.
// W3_s arrives as fp32 [H][12] (bf16-rounded values) right behind the biases in B32.
// =======================================================================================
constexpr int k3Stages = 5;
constexpr int k3StageBytes = 32768;
constexpr int k3Threads = 320;       // producer, issuer, 2 x 4 epilogue warps
constexpr int kW3Pitch = 12;

struct __align__(8) Barriers3 {
  uint64_t full[k3Stages], empty[k3Stages];
  uint64_t w3_full, w3_empty;
  uint64_t bias_full[2], bias_empty[2];
  uint64_t acc_full[2], act1_ready[2];
  uint64_t acc2_full[2], buf_free[2];
  uint64_t logit_full[2], logit_empty[2];
  uint32_t tmem_slot;
};

__global__ void __launch_bounds__(k3Threads, 1)
fused_mlp3_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
                  const __grid_constant__ CUtensorMap tmW2, const FusedParams p) {
  constexpr int CM = 2;
  uint8_t* smem = (uint8_t*)(((uintptr_t)fused_smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t s_stage = smem_u32(smem);
  float* w3f = (float*)(smem + k3Stages * k3StageBytes);                       // [H][12] fp32
  float* bias_smem = w3f + 512 * kW3Pitch;                                     // 2 x kBiasFloats
  float* lslot = bias_smem + 2 * kBiasFloats;                                  // 2 x [128][12] partial logits (WG1 -> WG0)
  Barriers3* bars = (Barriers3*)(lslot + 2 * 128 * kW3Pitch);
  const uint32_t s_w3f = smem_u32(w3f), s_bias = smem_u32(bias_smem);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int trace_n = 0;
  const uint32_t rank = cluster_ctarank();
  const bool leader = (rank == 0);
  const int cluster_id = blockIdx.x / CM;
  constexpr uint16_t kBoth = 0x3;

  const int H = p.H, NH = p.NH, K0 = p.K0;
  const int n_l1 = (H + 255) / 256;
  const int n_l2 = (NH == 2) ? H / 128 : 0;
  const int kb1 = (K0 + 63) / 64;
  const int l2_stages = H / 128;

  if (tid == 0) {
    for (int i = 0; i < k3Stages; ++i) { mbar_init(smem_u32(&bars->full[i]), 1); mbar_init(smem_u32(&bars->empty[i]), 1); }
    mbar_init(smem_u32(&bars->w3_full), 1); mbar_init(smem_u32(&bars->w3_empty), 8);
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&bars->bias_full[i]), 1); mbar_init(smem_u32(&bars->bias_empty[i]), 8);
      mbar_init(smem_u32(&bars->acc_full[i]), 1); mbar_init(smem_u32(&bars->act1_ready[i]), 8);   // 4 warps x 2 CTAs
      mbar_init(smem_u32(&bars->acc2_full[i]), 1); mbar_init(smem_u32(&bars->buf_free[i]), 8);
      mbar_init(smem_u32(&bars->logit_full[i]), 4); mbar_init(smem_u32(&bars->logit_empty[i]), 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc2<512>(smem_u32(&bars->tmem_slot));
  if (warp == 0 && lane == 0) { prefetch_tmap(&tmX); prefetch_tmap(&tmW1); prefetch_tmap(&tmW2); }
  fence_before_sync();
  __syncthreads();
  cluster_sync();
  fence_after_sync();
  const uint32_t tmem = bars->tmem_slot;

  const long long u0 = ((long long)cluster_id * p.U) / p.NC;
  const long long u1 = ((long long)(cluster_id + 1) * p.U) / p.NC;

  if (warp == 0) {
    // =========================== TMA producer (both CTAs) ===========================
    uint32_t it = 0, w3_n = 0, unit_n = 0;
    for (long long u = u0; u < u1; ++u, ++unit_n) {
      const int q = (int)(u / p.ns), s = (int)(u - (long long)q * p.ns);
      const int m0 = (q * CM + (int)rank) * 128;
      {
        const uint32_t b = unit_n & 1;
        mbar_wait(smem_u32(&bars->bias_empty[b]), ((unit_n >> 1) & 1) ^ 1);
        if (elect_one()) {
          mbar_expect_tx(smem_u32(&bars->bias_full[b]), (uint32_t)(p.w3f_off * 4));
          bulk_load(s_bias + b * kBiasFloats * 4, p.bias + (long long)s * p.PB, (uint32_t)(p.w3f_off * 4), smem_u32(&bars->bias_full[b]));
        }
        __syncwarp();
      }
      for (int c = 0; c < n_l1; ++c) {
        const int ncols = min(256, H - c * 256);
        const int nhalf = ncols / 128;
        for (int kb = 0; kb < kb1; ++kb, ++it) {
          const uint32_t st = it % k3Stages, ph = (it / k3Stages) & 1;
          mbar_wait(smem_u32(&bars->empty[st]), ph ^ 1);
          if (elect_one()) {
            const uint32_t fb = smem_u32(&bars->full[st]);
            const uint32_t base = s_stage + st * k3StageBytes;
            if (leader) mbar_expect_tx(fb, 2u * (uint32_t)(kXBytes + nhalf * kBlkBytes));
            tma2_load_2d(base, &tmX, fb, kb * 64, m0);
#pragma unroll
            for (int nb = 0; nb < 2; ++nb)
              if (nb < nhalf)
                tma2_load_3d(base + kXBytes + nb * kBlkBytes, &tmW1, fb, c * 256 + ((int)rank * nhalf + nb) * 64, kb * 64, s);
          }
          __syncwarp();
        }
        if (c == 0) {
          // output-layer kernel of this sample (fp32 [H][12], single buffer): free once both
          // warpgroups finished the previous unit's last chunk
          mbar_wait(smem_u32(&bars->w3_empty), (w3_n & 1) ^ 1);
          if (elect_one()) {
            const uint32_t wf = smem_u32(&bars->w3_full);
            mbar_expect_tx(wf, (uint32_t)(H * kW3Pitch * 4));
            bulk_load(s_w3f, p.bias + (long long)s * p.PB + p.w3f_off, (uint32_t)(H * kW3Pitch * 4), wf);
          }
          __syncwarp();
          ++w3_n;
        }
      }
      for (int j = 0; j < n_l2; ++j) {
        for (int sg = 0; sg < l2_stages; ++sg, ++it) {
          const uint32_t st = it % k3Stages, ph = (it / k3Stages) & 1;
          mbar_wait(smem_u32(&bars->empty[st]), ph ^ 1);
          if (elect_one()) {
            const uint32_t fb = smem_u32(&bars->full[st]);
            const uint32_t base = s_stage + st * k3StageBytes + kXBytes;
            if (leader) mbar_expect_tx(fb, 2u * 2u * kBlkBytes);
#pragma unroll
            for (int kblk = 0; kblk < 2; ++kblk)
              tma2_load_3d(base + kblk * kBlkBytes, &tmW2, fb, j * 128 + (int)rank * 64, (sg * 2 + kblk) * 64, s);
          }
          __syncwarp();
        }
      }
    }
    for (int k = 0; k < k3Stages; ++k, ++it) {
      const uint32_t st = it % k3Stages, ph = (it / k3Stages) & 1;
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
        if (++st_i == k3Stages) { st_i = 0; st_ph ^= 1; st_full = full0; st_empty = empty0; st_a16 = a16_0; }
        else { st_full += 8; st_empty += 8; st_a16 += k3StageBytes >> 4; }
      };
      constexpr uint32_t kDescHi128 = (1024u >> 4) | (1u << 14) | (2u << 29);
      uint32_t ph_act1_0 = 0, ph_act1_1 = 0;
      uint32_t issued0 = 0, issued1 = 0, waited0 = 0, waited1 = 0;   // accumulator hand-offs per buffer
      const uint32_t idesc_l2 = make_idesc_bf16(256, 128, 0, 1);
      // buffer / region b may be overwritten once every accumulator handed to its warpgroup was read
      auto wait_free = [&](int b) {
        if (b == 0) { while (waited0 < issued0) { mbar_wait(smem_u32(&bars->buf_free[0]), waited0 & 1); ++waited0; } }
        else { while (waited1 < issued1) { mbar_wait(smem_u32(&bars->buf_free[1]), waited1 & 1); ++waited1; } }
        fence_after_sync();
      };

      for (long long u = u0; u < u1; ++u) {
        for (int c = 0; c < n_l1; ++c) {
          const int ncols = min(256, H - c * 256);
          const uint32_t idesc = make_idesc_bf16(256, ncols, 0, 1);
          const uint32_t dreg = tmem + c * 256;
          wait_free(c);
          if (n_l1 == 1) wait_free(1);     // H <= 256: buffer 1 lives in the otherwise unused region 1
          DBX_TRACE(100 + c);
          for (int kb = 0; kb < kb1; ++kb) {
            mbar_wait(st_full, st_ph);
            DBX_STAGE_FENCE();
            if (elect_one()) {
              const uint32_t a_lo = st_a16 | (1u << 16);
              const uint32_t b_lo = (st_a16 + (kXBytes >> 4)) | ((kBlkBytes >> 4) << 16);
              if (kb + 1 < kb1 || (K0 & 63) == 0) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                  mma2_ss_lh(dreg, a_lo + kk * 2, kDescHi128, b_lo + kk * 128, kDescHi128, idesc, (kb | kk) > 0);
              } else {
                const int ksteps = ((K0 & 63) + 15) >> 4;
                for (int kk = 0; kk < ksteps; ++kk)
                  mma2_ss_lh(dreg, a_lo + kk * 2, kDescHi128, b_lo + kk * 128, kDescHi128, idesc, (kb | kk) > 0);
              }
              mma2_commit_mc(st_empty, kBoth);
            }
            __syncwarp();
            advance();
          }
          if (elect_one()) mma2_commit_mc(smem_u32(&bars->acc_full[c]), kBoth);
          __syncwarp();
          if (NH == 1) { if (c == 0) ++issued0; else ++issued1; }
          DBX_TRACE(110 + c);
        }
        for (int j = 0; j < n_l2; ++j) {
          const int b = j & 1;
          const uint32_t dbuf = tmem + b * 256 + 128;
          wait_free(b);
          DBX_TRACE(120 + j);
          for (int sg = 0; sg < l2_stages; ++sg) {
            if (j == 0 && (sg & 1) == 0) {   // first touch of packed H1 region sg>>1 (also frees its upper half)
              if ((sg >> 1) == 0) { mbar_wait(smem_u32(&bars->act1_ready[0]), ph_act1_0 & 1); ++ph_act1_0; }
              else { mbar_wait(smem_u32(&bars->act1_ready[1]), ph_act1_1 & 1); ++ph_act1_1; }
            }
            mbar_wait(st_full, st_ph);
            DBX_STAGE_FENCE();
            if (elect_one()) {
              const uint32_t b_lo = (st_a16 + (kXBytes >> 4)) | ((kBlkBytes >> 4) << 16);
              const uint32_t a0 = tmem + (uint32_t)(sg >> 1) * 256 + (uint32_t)(sg & 1) * 64;
#pragma unroll
              for (int i = 0; i < 8; ++i)
                mma2_ts_lh(dbuf, a0 + i * 8, b_lo + (i >> 2) * (kBlkBytes / 16) + (i & 3) * 128, kDescHi128, idesc_l2, (sg | i) > 0);
              mma2_commit_mc(st_empty, kBoth);
            }
            __syncwarp();
            advance();
          }
          if (elect_one()) mma2_commit_mc(smem_u32(&bars->acc2_full[b]), kBoth);
          __syncwarp();
          if (b == 0) ++issued0; else ++issued1;
          DBX_TRACE(130 + j);
        }
        // H <= 256 with two hidden layers: region 1 is never packed, but chunk 1 of layer 2 used its upper
        // half as buffer 1 -- nothing else to do, wait_free(1) above covers it.
      }
    }
  } else {
    // =========================== epilogue: warpgroup 0 = warps 2-5, warpgroup 1 = warps 6-9 ===========================
    const int wg = (warp - 2) >> 2;
    const int g = warp & 3;                       // TMEM lane group of this warp
    const int row = g * 32 + lane;
    const uint32_t lane_base = (uint32_t)(g * 32) << 16;
    uint32_t ph_acc = 0, ph_acc2 = 0, unit_n = 0;
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
    // relu(acc + bias) as packed bf16 in place (next layer's A operand)
    auto pack_chunk = [&](uint32_t src, int ncols, const float* bias) {
      for (int c0 = 0; c0 < ncols; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(src + c0, v);
        wait_ld();
        uint32_t w[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float2 bb = *reinterpret_cast<const float2*>(bias + c0 + 2 * i);
          w[i] = pack_bf16x2(fmaxf(__uint_as_float(v[2 * i]) + bb.x, 0.f), fmaxf(__uint_as_float(v[2 * i + 1]) + bb.y, 0.f));
        }
        tmem_st16(src + c0 / 2, w);
      }
      wait_st();
      fence_before_sync();
    };

    for (long long u = u0; u < u1; ++u, ++unit_n) {
      const int q = (int)(u / p.ns), s = (int)(u - (long long)q * p.ns);
      if (wg == 0 && q != cur_q) { flush_segment(cur_q); cur_q = q; }
      const uint32_t bb = unit_n & 1;
      const float* bias = bias_smem + bb * kBiasFloats;
      mbar_wait(smem_u32(&bars->bias_full[bb]), (unit_n >> 1) & 1);
      float logit[kW3Pitch];
#pragma unroll
      for (int i = 0; i < kW3Pitch; ++i) logit[i] = 0.f;

      // logit += bf16(relu(acc + bias)) * W3[k0 .. k0+ncols), accumulators read from TMEM at `src`
      auto out_layer_chunk = [&](uint32_t src, int ncols, const float* bvec, int k0) {
        for (int c0 = 0; c0 < ncols; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(src + c0, v);
          wait_ld();
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float2 b2 = *reinterpret_cast<const float2*>(bvec + c0 + i);
            // round to bf16 exactly as the packed tensor-core operand would be
            const uint32_t pk = pack_bf16x2(fmaxf(__uint_as_float(v[i]) + b2.x, 0.f), fmaxf(__uint_as_float(v[i + 1]) + b2.y, 0.f));
            const float h0 = __uint_as_float(pk << 16), h1 = __uint_as_float(pk & 0xFFFF0000u);
            const float4* w0 = reinterpret_cast<const float4*>(w3f + (size_t)(k0 + c0 + i) * kW3Pitch);
            const float4 a0 = w0[0], a1 = w0[1], a2 = w0[2], c0v = w0[3], c1v = w0[4], c2v = w0[5];
            logit[0] = fmaf(h0, a0.x, logit[0]); logit[1] = fmaf(h0, a0.y, logit[1]); logit[2] = fmaf(h0, a0.z, logit[2]); logit[3] = fmaf(h0, a0.w, logit[3]);
            logit[4] = fmaf(h0, a1.x, logit[4]); logit[5] = fmaf(h0, a1.y, logit[5]); logit[6] = fmaf(h0, a1.z, logit[6]); logit[7] = fmaf(h0, a1.w, logit[7]);
            logit[8] = fmaf(h0, a2.x, logit[8]); logit[9] = fmaf(h0, a2.y, logit[9]); logit[10] = fmaf(h0, a2.z, logit[10]); logit[11] = fmaf(h0, a2.w, logit[11]);
            logit[0] = fmaf(h1, c0v.x, logit[0]); logit[1] = fmaf(h1, c0v.y, logit[1]); logit[2] = fmaf(h1, c0v.z, logit[2]); logit[3] = fmaf(h1, c0v.w, logit[3]);
            logit[4] = fmaf(h1, c1v.x, logit[4]); logit[5] = fmaf(h1, c1v.y, logit[5]); logit[6] = fmaf(h1, c1v.z, logit[6]); logit[7] = fmaf(h1, c1v.w, logit[7]);
            logit[8] = fmaf(h1, c2v.x, logit[8]); logit[9] = fmaf(h1, c2v.y, logit[9]); logit[10] = fmaf(h1, c2v.z, logit[10]); logit[11] = fmaf(h1, c2v.w, logit[11]);
          }
        }
        fence_before_sync();
      };

      bool w3_waited = false;
      auto wait_w3 = [&]() { if (!w3_waited) { mbar_wait(smem_u32(&bars->w3_full), unit_n & 1); w3_waited = true; } };

      // ---- layer-1 chunk of this warpgroup
      if (wg < n_l1) {
        const int c = wg;
        const int ncols = min(256, H - c * 256);
        mbar_wait(smem_u32(&bars->acc_full[c]), ph_acc & 1);
        ++ph_acc;
        fence_after_sync();
        if (warp == 2) DBX_TRACE(200 + c);
        if (NH == 2) {
          pack_chunk(tmem + lane_base + c * 256, ncols, bias + p.b1_off + c * 256);
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(smem_u32(&bars->act1_ready[c]), 0);
        } else {
          wait_w3();
          out_layer_chunk(tmem + lane_base + c * 256, ncols, bias + p.b1_off + c * 256, c * 256);
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(smem_u32(&bars->buf_free[c]), 0);
        }
        if (warp == 2) DBX_TRACE(210 + c);
      }
      // ---- layer-2 chunks of this warpgroup (buffer == warpgroup)
      for (int j = wg; j < n_l2; j += 2) {
        mbar_wait(smem_u32(&bars->acc2_full[wg]), ph_acc2 & 1);
        ++ph_acc2;
        fence_after_sync();
        if (warp == 2) DBX_TRACE(220 + j);
        wait_w3();
        out_layer_chunk(tmem + lane_base + wg * 256 + 128, 128, bias + p.b2_off + j * 128, j * 128);
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(smem_u32(&bars->buf_free[wg]), 0);
        if (warp == 2) DBX_TRACE(230 + j);
      }
      // this warpgroup is done with W3 (and, for WG1, with the biases) of this unit
      wait_w3();   // keeps the w3 phase in step even if this warpgroup had no chunk
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bars->w3_empty));
      if (wg == 1) {
        // hand the partial logits to warpgroup 0
        mbar_wait(smem_u32(&bars->logit_empty[bb]), ((unit_n >> 1) & 1) ^ 1);
        float4* dst = reinterpret_cast<float4*>(lslot + ((size_t)bb * 128 + row) * kW3Pitch);
        dst[0] = make_float4(logit[0], logit[1], logit[2], logit[3]);
        dst[1] = make_float4(logit[4], logit[5], logit[6], logit[7]);
        dst[2] = make_float4(logit[8], logit[9], logit[10], logit[11]);
        __syncwarp();
        if (lane == 0) { mbar_arrive(smem_u32(&bars->logit_full[bb])); mbar_arrive(smem_u32(&bars->bias_empty[bb])); }
        continue;
      }
      // ---- warpgroup 0: total logits, output activation, running sums
      mbar_wait(smem_u32(&bars->logit_full[bb]), (unit_n >> 1) & 1);
      {
        const float4* src = reinterpret_cast<const float4*>(lslot + ((size_t)bb * 128 + row) * kW3Pitch);
        const float4 x0 = src[0], x1 = src[1], x2 = src[2];
        logit[0] += x0.x; logit[1] += x0.y; logit[2] += x0.z; logit[3] += x0.w;
        logit[4] += x1.x; logit[5] += x1.y; logit[6] += x1.z; logit[7] += x1.w;
        logit[8] += x2.x; logit[9] += x2.y; logit[10] += x2.z; logit[11] += x2.w;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bars->logit_empty[bb]));
      const float* b3 = bias + p.b3_off;
      float z[kW3Pitch];
#pragma unroll
      for (int i = 0; i < kW3Pitch; ++i) z[i] = (i < p.C) ? logit[i] + b3[i] : -INFINITY;
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bars->bias_empty[bb]));
      const int grow = (q * CM + (int)rank) * 128 + row;
      if (p.sink_kind == 0) {
        const float w = p.wts ? p.wts[s] : 1.f;
        if (p.out_kind == DBX_OUT_LOGITS) {
#pragma unroll
          for (int i = 0; i < kW3Pitch; ++i) if (i < p.C) { acc1[i] = fmaf(w, z[i], acc1[i]); acc2[i] = fmaf(w * z[i], z[i], acc2[i]); }
        } else if (p.act_last == DBX_ACT_SOFTMAX) {
          float mx = z[0];
#pragma unroll
          for (int i = 1; i < kW3Pitch; ++i) mx = fmaxf(mx, z[i]);
          float e[kW3Pitch], den = 0.f;
#pragma unroll
          for (int i = 0; i < kW3Pitch; ++i) { e[i] = (i < p.C) ? expf(z[i] - mx) : 0.f; den += e[i]; }
          const float inv = 1.f / den;
#pragma unroll
          for (int i = 0; i < kW3Pitch; ++i) { const float pr = e[i] * inv; acc1[i] = fmaf(w, pr, acc1[i]); acc2[i] = fmaf(w * pr, pr, acc2[i]); }
        } else {
#pragma unroll
          for (int i = 0; i < kW3Pitch; ++i) if (i < p.C) { const float pr = apply_act(p.act_last, z[i]); acc1[i] = fmaf(w, pr, acc1[i]); acc2[i] = fmaf(w * pr, pr, acc2[i]); }
        }
      } else {
        int best = 0; float bv = z[0];
#pragma unroll
        for (int i = 1; i < kW3Pitch; ++i) if (z[i] > bv) { bv = z[i]; best = i; }
        if (grow < p.B) {
          const int ok = (best == p.labels[grow]);
          if (p.flags) p.flags[(long long)s * p.B + grow] = (uint8_t)ok;
          cnt += ok;
        }
      }
    }
    if (wg == 0 && cur_q >= 0) flush_segment(cur_q);
  }

  fence_before_sync();
  __syncthreads();
  cluster_sync();
  if (warp == 1) tmem_dealloc2<512>(tmem);
}

#endif
