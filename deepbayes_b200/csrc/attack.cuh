// Expected input gradient over posterior samples (included by engine.cu).
//
// Reference: deepbayes/analyzers/attacks.py gradient_expectation :49-75 -- for each of num_models iterations,
// d loss_fn(direction, model.predict(inp)) / d inp by tf.GradientTape, summed.  model.predict of a PosteriorModel is
// itself the mean of n forward passes under fresh weight samples (posteriormodel.py:81-101), of an optimizer a single
// forward.  Here, for one iteration with n_inner samples s:
//     pbar = (1/n) sum_s softmax(f_s(x)),   loss = -(1/B) sum_b log clip(pbar_b[y_b])      (sparse categorical CE)
//     dLoss/dz_s[b, c] = g_b * p_s[b, y_b] * ([c == y_b] - p_s[b, c]),   g_b = -1 / (B n pbar_b[y_b])
//     dLoss/dx = sum_s backprop_s(dLoss/dz_s)
// evaluated with the chunk's samples side by side (fp32, CUDA cores, the sample is the grid's z dimension).
#pragma once

namespace dbx {

// in place: logits[s][b][:] -> p_s[b][:];  pbar[b][:] = mean over s
__global__ void softmax_mean_kernel(float* __restrict__ z, int ns, int B, int C, float* __restrict__ pbar) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  for (int c = 0; c < C; ++c) pbar[(int64_t)b * C + c] = 0.f;
  for (int s = 0; s < ns; ++s) {
    float* row = z + ((int64_t)s * B + b) * C;
    float m = row[0];
    for (int c = 1; c < C; ++c) m = fmaxf(m, row[c]);
    float den = 0.f;
    for (int c = 0; c < C; ++c) den += expf(row[c] - m);
    for (int c = 0; c < C; ++c) { const float p = expf(row[c] - m) / den; row[c] = p; pbar[(int64_t)b * C + c] += p; }
  }
  for (int c = 0; c < C; ++c) pbar[(int64_t)b * C + c] /= (float)ns;
}

// dz[s][b][c] from p_s and pbar (see the header); zero where Keras' clip of pbar[y] to [1e-7, 1-1e-7] is active
__global__ void ce_mean_upstream_kernel(const float* __restrict__ p, const float* __restrict__ pbar, const int32_t* __restrict__ labels,
                                        int ns, int B, int C, float* __restrict__ dz) {
  const int64_t total = (int64_t)ns * B;
  for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(o % B);
    const int y = labels[b];
    const float pb = pbar[(int64_t)b * C + y];
    const float g = (pb > 1e-7f && pb < 1.f - 1e-7f) ? -1.f / ((float)B * (float)ns * pb) : 0.f;
    const float* ps = p + o * C;
    const float gy = g * ps[y];
    for (int c = 0; c < C; ++c) dz[o * C + c] = gy * ((c == y ? 1.f : 0.f) - ps[c]);
  }
}

// batched over samples (blockIdx.z): dZprev[s][b, k] = (sum_n dZ[s][b, n] * W_s[k, n]) * act'(Aprev[s][b, k]); aprev == nullptr: no act'
__global__ void __launch_bounds__(256)
grad_a_batched_kernel(const float* __restrict__ dZ, int64_t dz_ss, const float* __restrict__ Wbase, int64_t w_ss, int64_t w_off,
                      const int32_t* __restrict__ rows, int64_t row0, const float* __restrict__ Aprev, int64_t a_ss, int B, int K, int N,
                      int act, float* __restrict__ out, int64_t o_ss) {
  __shared__ float Zs[32][33];
  __shared__ float Ws[32][33];
  const int s = blockIdx.z;
  const int64_t wrow = rows ? (int64_t)rows[s] : (row0 + s);
  const float* dZp = dZ + (int64_t)s * dz_ss;
  const float* W = Wbase + wrow * w_ss + w_off;
  const int b0 = blockIdx.y * 32, k0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int n0 = 0; n0 < N; n0 += 32) {
    for (int i = ty; i < 32; i += 8) {
      Zs[i][tx] = (b0 + i < B && n0 + tx < N) ? dZp[(int64_t)(b0 + i) * N + n0 + tx] : 0.f;
      Ws[i][tx] = (k0 + i < K && n0 + tx < N) ? W[(int64_t)(k0 + i) * N + n0 + tx] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int n = 0; n < 32; ++n) {
      const float w = Ws[tx][n];
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[r] = fmaf(Zs[ty + 8 * r][n], w, acc[r]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int b = b0 + ty + 8 * r, k = k0 + tx;
    if (b < B && k < K) {
      float d = 1.f;
      if (Aprev) {
        const float a = Aprev[(int64_t)s * a_ss + (int64_t)b * K + k];
        switch (act) {
          case DBX_ACT_RELU: d = a > 0.f ? 1.f : 0.f; break;
          case DBX_ACT_TANH: d = 1.f - a * a; break;
          case DBX_ACT_SIGMOID: d = a * (1.f - a); break;
          default: d = 1.f;
        }
      }
      out[(int64_t)s * o_ss + (int64_t)b * K + k] = acc[r] * d;
    }
  }
}

// grad[e] (+)= sum_s dx[s][e], in sample order
__global__ void sum_samples_kernel(const float* __restrict__ dx, int ns, int64_t n, int accumulate, float* __restrict__ grad) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    float a = accumulate ? grad[e] : 0.f;
    for (int s = 0; s < ns; ++s) a += dx[(int64_t)s * n + e];
    grad[e] = a;
  }
}

static int run_input_gradient(dbx_model* m, const float* x_dev, int64_t B, const int32_t* labels_dev, int64_t n_outer, int64_t n_inner,
                              const Source& src0, float* grad_dev, cudaStream_t st) {
  DBX_CHECK(m->is_mlp, "the device input gradient covers Dense MLPs (attacks.py:49-75)");
  DBX_CHECK(m->layers[m->last_weighted].act == DBX_ACT_SOFTMAX, "the device input gradient needs a softmax output (sparse categorical cross-entropy)");
  std::vector<const Layer*> dl;
  for (auto& L : m->layers) if (L.kind == DBX_DENSE) dl.push_back(&L);
  const int nl = (int)dl.size();
  const int64_t C = m->out_feat, K0 = m->in_feat;
  const int nc = (int)n_inner;
  size_t act_total = 0;
  for (auto* L : dl) act_total += (size_t)B * L->units * 4;      // per-sample blocks are dense: [s][B][units]
  const size_t dzb = round_up((size_t)B * std::max<int64_t>(m->max_feat, K0) * 4, 256);
  const size_t wb = src0.stored ? 0 : (size_t)m->P * 4;
  const size_t per = act_total + 2 * dzb + wb;
  const size_t need = (size_t)nc * per + round_up((size_t)B * C * 4, 256) + 8192;
  DBX_CHECK(need <= ws_budget_bytes() * 2, "n_inner = %d samples need %zu MB of workspace", nc, need >> 20);
  if (ensure_ws(m->ws, need)) return 1;
  char* p = (char*)m->ws.ptr;
  std::vector<float*> acts(nl);
  std::vector<int64_t> act_ss(nl);
  for (int i = 0; i < nl; ++i) { acts[i] = (float*)p; act_ss[i] = B * dl[i]->units; p += (size_t)nc * act_ss[i] * 4; }
  p = (char*)(((uintptr_t)p + 255) & ~(uintptr_t)255);
  float* dz[2];
  dz[0] = (float*)p; p += (size_t)nc * dzb; dz[1] = (float*)p; p += (size_t)nc * dzb;
  float* pbar = (float*)p; p += round_up((size_t)B * C * 4, 256);
  float* Wc = (float*)p;

  for (int64_t it = 0; it < n_outer; ++it) {
    Source src = src0;
    src.s0 = src0.s0 + it * n_inner; src.j0 = src0.j0 + (src0.idx ? 0 : it * n_inner);
    const float* Wbase; int64_t w_ss; const int32_t* rows = nullptr; int64_t row0 = 0;
    if (src.stored) { Wbase = m->slab; w_ss = m->slab_stride; rows = src.idx ? src.idx + it * n_inner : nullptr; row0 = src.j0; }
    else {
      const int64_t total = (int64_t)nc * ((m->P + 3) / 4);
      DBX_LAUNCH(sample_f32_kernel, grid_for(total), 256, 0, st, m->mu, m->sigma, src.eps ? src.eps + it * n_inner * m->P : nullptr,
                 src.inflate, src.seed, src.s0, (int64_t)nc, m->P, 0, Wc, (float*)nullptr);
      Wbase = Wc; w_ss = m->P;      // sample_f32_kernel writes rows m->P apart
    }
    // forward, all samples side by side
    const float* cur = x_dev; int64_t cur_ss = 0;
    for (int i = 0; i < nl; ++i) {
      const Layer& L = *dl[i];
      dim3 grid((unsigned)cdiv(L.units, 64), (unsigned)cdiv(B, 64), (unsigned)nc);
      DBX_LAUNCH((dense_kernel<float, float>), grid, 256, 0, st, cur, cur_ss, Wbase, w_ss, rows, row0, L.w_off, (int)L.w_cols, Wbase, w_ss,
                 L.b_off, acts[i], act_ss[i], (int)B, L.units, (int)L.in_feat, i == nl - 1 ? DBX_ACT_LINEAR : L.act);
      cur = acts[i]; cur_ss = act_ss[i];
    }
    float* logits = acts[nl - 1];
    DBX_LAUNCH(softmax_mean_kernel, (int)cdiv(B, 128), 128, 0, st, logits, nc, (int)B, (int)C, pbar);
    DBX_LAUNCH(ce_mean_upstream_kernel, grid_for((int64_t)nc * B), 256, 0, st, (const float*)logits, (const float*)pbar, labels_dev, nc, (int)B,
               (int)C, dz[0]);
    // backward to the input
    int zi = 0; int64_t zs = B * C;
    for (int i = nl - 1; i >= 0; --i) {
      const Layer& L = *dl[i];
      const int K = (int)L.in_feat, N = L.units;
      dim3 ga((unsigned)cdiv(K, 32), (unsigned)cdiv(B, 32), (unsigned)nc);
      const float* aprev = i > 0 ? acts[i - 1] : nullptr;
      DBX_LAUNCH(grad_a_batched_kernel, ga, 256, 0, st, (const float*)dz[zi], zs, Wbase, w_ss, L.w_off, rows, row0, aprev, i > 0 ? act_ss[i - 1] : 0,
                 (int)B, K, N, i > 0 ? dl[i - 1]->act : DBX_ACT_LINEAR, dz[zi ^ 1], (int64_t)B * K);
      zi ^= 1; zs = (int64_t)B * K;
    }
    DBX_LAUNCH(sum_samples_kernel, grid_for(B * K0), 256, 0, st, (const float*)dz[zi], nc, B * K0, it > 0 ? 1 : 0, grad_dev);
  }
  return 0;
}

// per-sample upstream gradient (n_inner = 1): dz[s][b][c] = (p_s[b,c] - [c == y_b]) / B, zero where the clip is active; logits -> p in place
__global__ void softmax_ce_per_sample_kernel(float* __restrict__ z, const int32_t* __restrict__ labels, int ns, int B, int C,
                                             float* __restrict__ dz) {
  const int64_t total = (int64_t)ns * B;
  for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(o % B);
    float* row = z + o * C;
    float m = row[0];
    for (int c = 1; c < C; ++c) m = fmaxf(m, row[c]);
    float den = 0.f;
    for (int c = 0; c < C; ++c) den += expf(row[c] - m);
    const int y = labels[b];
    const float py = expf(row[y] - m) / den;
    const float active = (py > 1e-7f && py < 1.f - 1e-7f) ? 1.f / (float)B : 0.f;
    for (int c = 0; c < C; ++c) {
      const float p = expf(row[c] - m) / den;
      row[c] = p;
      dz[o * C + c] = (p - (c == y ? 1.f : 0.f)) * active;
    }
  }
}

// adv[s][e] = clip(clip(x[e] + eps * sign(g[s][e]), x[e] - eps, x[e] + eps), lower, upper)   (attacks.py:94-101), in place over g
__global__ void fgsm_adv_kernel(const float* __restrict__ x, float* __restrict__ g, int ns, int64_t n, float eps, float lower, float upper) {
  const int64_t total = (int64_t)ns * n;
  for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
    const float xe = x[o % n], ge = g[o];
    const float sgn = ge > 0.f ? 1.f : (ge < 0.f ? -1.f : 0.f);
    float a = xe + eps * sgn;
    a = fminf(fmaxf(a, xe - eps), xe + eps);
    g[o] = fminf(fmaxf(a, lower), upper);
  }
}

// massart_bound_check with verify=False (verifiers.py:622-634): for every posterior sample, an FGSM step against THAT
// sample's weights, then the predicate argmax == label on the sample's output at the adversarial input.
static int run_count_fgsm(dbx_model* m, const float* x_dev, int64_t B, const int32_t* attack_labels_dev, const int32_t* labels_dev,
                          float eps, float lower, float upper, int64_t n, const Source& src0, int accumulate, uint8_t* flags_dev,
                          unsigned long long* counts_dev, cudaStream_t st, float* probs_out = nullptr) {
  DBX_CHECK(m->is_mlp, "the device FGSM loop covers Dense MLPs");
  DBX_CHECK(m->layers[m->last_weighted].act == DBX_ACT_SOFTMAX, "the device FGSM loop needs a softmax output");
  std::vector<const Layer*> dl;
  for (auto& L : m->layers) if (L.kind == DBX_DENSE) dl.push_back(&L);
  const int nl = (int)dl.size();
  const int64_t C = m->out_feat, K0 = m->in_feat;
  size_t act_total = 0;
  for (auto* L : dl) act_total += (size_t)B * L->units * 4;
  const size_t dzb = round_up((size_t)B * std::max<int64_t>(m->max_feat, K0) * 4, 256);
  const size_t wb = src0.stored ? 0 : (size_t)m->P * 4;
  const size_t per = act_total + 2 * dzb + wb + 64;
  int64_t ns = (int64_t)std::max<size_t>(1, ws_budget_bytes() / per);
  ns = std::min<int64_t>(std::min<int64_t>(ns, n), 256);
  if (ensure_ws(m->ws, (size_t)ns * per + 8192)) return 1;
  char* p = (char*)m->ws.ptr;
  std::vector<float*> acts(nl);
  std::vector<int64_t> act_ss(nl);
  for (int i = 0; i < nl; ++i) { acts[i] = (float*)p; act_ss[i] = B * dl[i]->units; p += (size_t)ns * act_ss[i] * 4; }
  p = (char*)(((uintptr_t)p + 255) & ~(uintptr_t)255);
  float* dz[2];
  dz[0] = (float*)p; p += (size_t)ns * dzb; dz[1] = (float*)p; p += (size_t)ns * dzb;
  float* Wc = (float*)p;
  if (!accumulate && counts_dev) DBX_LAUNCH(zero_u64_kernel, grid_for(B), 256, 0, st, counts_dev, B);

  for (int64_t c0 = 0; c0 < n; c0 += ns) {
    const int nc = (int)std::min<int64_t>(ns, n - c0);
    const float* Wbase; int64_t w_ss; const int32_t* rows = nullptr; int64_t row0 = 0;
    if (src0.stored) { Wbase = m->slab; w_ss = m->slab_stride; rows = src0.idx ? src0.idx + c0 : nullptr; row0 = src0.j0 + c0; }
    else {
      const int64_t total = (int64_t)nc * ((m->P + 3) / 4);
      DBX_LAUNCH(sample_f32_kernel, grid_for(total), 256, 0, st, m->mu, m->sigma, src0.eps ? src0.eps + c0 * m->P : nullptr, src0.inflate,
                 src0.seed, src0.s0 + c0, (int64_t)nc, m->P, 0, Wc, (float*)nullptr);
      Wbase = Wc; w_ss = m->P;
    }
    auto forward = [&](const float* in, int64_t in_ss) -> int {
      const float* cur = in; int64_t cur_ss = in_ss;
      for (int i = 0; i < nl; ++i) {
        const Layer& L = *dl[i];
        dim3 grid((unsigned)cdiv(L.units, 64), (unsigned)cdiv(B, 64), (unsigned)nc);
        DBX_LAUNCH((dense_kernel<float, float>), grid, 256, 0, st, cur, cur_ss, Wbase, w_ss, rows, row0, L.w_off, (int)L.w_cols, Wbase, w_ss,
                   L.b_off, acts[i], act_ss[i], (int)B, L.units, (int)L.in_feat, i == nl - 1 ? DBX_ACT_LINEAR : L.act);
        cur = acts[i]; cur_ss = act_ss[i];
      }
      return 0;
    };
    if (forward(x_dev, 0)) return 1;
    DBX_LAUNCH(softmax_ce_per_sample_kernel, grid_for((int64_t)nc * B), 256, 0, st, acts[nl - 1], attack_labels_dev, nc, (int)B, (int)C, dz[0]);
    int zi = 0; int64_t zs = B * C;
    for (int i = nl - 1; i >= 0; --i) {
      const Layer& L = *dl[i];
      const int K = (int)L.in_feat, N = L.units;
      dim3 ga((unsigned)cdiv(K, 32), (unsigned)cdiv(B, 32), (unsigned)nc);
      DBX_LAUNCH(grad_a_batched_kernel, ga, 256, 0, st, (const float*)dz[zi], zs, Wbase, w_ss, L.w_off, rows, row0,
                 i > 0 ? (const float*)acts[i - 1] : (const float*)nullptr, i > 0 ? act_ss[i - 1] : 0, (int)B, K, N,
                 i > 0 ? dl[i - 1]->act : DBX_ACT_LINEAR, dz[zi ^ 1], (int64_t)B * K);
      zi ^= 1; zs = (int64_t)B * K;
    }
    DBX_LAUNCH(fgsm_adv_kernel, grid_for((int64_t)nc * B * K0), 256, 0, st, x_dev, dz[zi], nc, B * K0, eps, lower, upper);
    if (forward(dz[zi], B * K0)) return 1;
    if (counts_dev)
      DBX_LAUNCH(count_kernel, grid_for((int64_t)nc * B), 256, 0, st, (const float*)acts[nl - 1], nc, B, (int)C, labels_dev,
                 flags_dev ? flags_dev + c0 * B : nullptr, counts_dev);
    if (probs_out) {   // every sample's own output at its adversarial input (model._predict(adv), verifiers.py:624-625)
      float* dst = probs_out + c0 * B * C;
      DBX_CUDA(cudaMemcpyAsync(dst, acts[nl - 1], (size_t)nc * B * C * 4, cudaMemcpyDeviceToDevice, st));
      DBX_LAUNCH(act_rows_kernel, grid_for((int64_t)nc * B), 256, 0, st, (int)DBX_ACT_SOFTMAX, dst, (int64_t)nc * B, C);
    }
  }
  return 0;
}

// ---------------------------------------------------------------------------------------
// bf16 mode: the same two loops on the tensor cores.  Forward = the generic executor's per-sample GEMMs (dense_tc_kernel /
// dense_tc_persist_kernel / dense_tc_pair_kernel, bf16 activations KEPT per layer), backward = dense_tc_bwd_kernel per layer
// (dZ * W_s^T with the activation derivative in its epilogue); softmax / cross-entropy upstream and the sums over samples stay
// fp32.  Operands are rounded to bf16 (weights, activations, upstream gradients), accumulation is fp32.
// ---------------------------------------------------------------------------------------
static bool grad_tc_ok(const dbx_model* m) {
  if (opt().no_tc || !m->is_mlp || m->layers[m->last_weighted].act != DBX_ACT_SOFTMAX) return false;
  for (int li = 0; li < (int)m->layers.size(); ++li) {
    const Layer& L = m->layers[li];
    if (L.kind != DBX_DENSE) continue;
    if ((L.in_feat & 7) || (L.w_cols_pad & 7) || (L.w16_off & 7)) return false;
    if (li != m->last_weighted && (L.units & 7)) return false;
  }
  return (m->P16 & 7) == 0;
}

// dz16[s][b][0..Cp) from p_s and pbar (ce_mean_upstream_kernel), columns >= C zero
__global__ void ce_mean_upstream_bf16_kernel(const float* __restrict__ p, const float* __restrict__ pbar, const int32_t* __restrict__ labels,
                                             int ns, int B, int C, int Cp, __nv_bfloat16* __restrict__ dz) {
  const int64_t total = (int64_t)ns * B;
  for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(o % B);
    const int y = labels[b];
    const float pb = pbar[(int64_t)b * C + y];
    const float g = (pb > 1e-7f && pb < 1.f - 1e-7f) ? -1.f / ((float)B * (float)ns * pb) : 0.f;
    const float* ps = p + o * C;
    const float gy = g * ps[y];
    for (int c = 0; c < Cp; ++c) dz[o * Cp + c] = __float2bfloat16_rn(c < C ? gy * ((c == y ? 1.f : 0.f) - ps[c]) : 0.f);
  }
}

// per-sample upstream gradient (softmax_ce_per_sample_kernel) as bf16 [s][b][Cp]; logits -> p in place
__global__ void softmax_ce_per_sample_bf16_kernel(float* __restrict__ z, const int32_t* __restrict__ labels, int ns, int B, int C, int Cp,
                                                  __nv_bfloat16* __restrict__ dz) {
  const int64_t total = (int64_t)ns * B;
  for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(o % B);
    float* row = z + o * C;
    float m = row[0];
    for (int c = 1; c < C; ++c) m = fmaxf(m, row[c]);
    float den = 0.f;
    for (int c = 0; c < C; ++c) den += expf(row[c] - m);
    const int y = labels[b];
    const float py = expf(row[y] - m) / den;
    const float active = (py > 1e-7f && py < 1.f - 1e-7f) ? 1.f / (float)B : 0.f;
    for (int c = 0; c < Cp; ++c) {
      float d = 0.f;
      if (c < C) { const float pc = expf(row[c] - m) / den; row[c] = pc; d = (pc - (c == y ? 1.f : 0.f)) * active; }
      dz[o * Cp + c] = __float2bfloat16_rn(d);
    }
  }
}

// Workspace of one chunk of the tensor-core gradient loops.
struct GradTcBufs {
  std::vector<const Layer*> dl;
  std::vector<__nv_bfloat16*> acts;      // hidden layers' outputs [nc][B][units]
  __nv_bfloat16* xin = nullptr;          // [B][K0]
  __nv_bfloat16* adv = nullptr;          // [nc][B][K0] (FGSM loop)
  __nv_bfloat16* dz16 = nullptr;         // [nc][B][Cp]
  __nv_bfloat16* dbuf[2] = {nullptr, nullptr};   // [nc][B][max hidden]
  float* logits = nullptr; float* dx = nullptr; float* pbar = nullptr;
  __nv_bfloat16* W16b[2] = {nullptr, nullptr}; float* B32b[2] = {nullptr, nullptr};   // two chunks of weights (drawn ahead on a side stream)
  __nv_bfloat16* W16 = nullptr; float* B32 = nullptr;                                 // the chunk in use
  int Cp = 0; int64_t hmax = 0;
};

static size_t grad_tc_per_sample(const dbx_model* m, int64_t B, bool with_adv, GradTcBufs& g, int nwb = 1) {
  g.dl.clear();
  for (auto& L : m->layers) if (L.kind == DBX_DENSE) g.dl.push_back(&L);
  const int nl = (int)g.dl.size();
  g.Cp = (int)round_up(m->out_feat, 8);
  g.hmax = 8;
  size_t per = 0;
  for (int i = 0; i + 1 < nl; ++i) { per += round_up((size_t)B * g.dl[i]->units * 2, 256); g.hmax = std::max<int64_t>(g.hmax, g.dl[i]->units); }
  per += round_up((size_t)B * m->out_feat * 4, 256) + round_up((size_t)B * g.Cp * 2, 256) + 2 * round_up((size_t)B * g.hmax * 2, 256) +
         nwb * (round_up((size_t)m->P16 * 2, 256) + round_up((size_t)m->PB * 4, 256));
  // the FGSM loop writes each sample's adversarial input (bf16) straight from the last backward GEMM; the expected-gradient
  // loop keeps every sample's fp32 gradient for the sum in sample order
  per += with_adv ? round_up((size_t)B * m->in_feat * 2, 256) : round_up((size_t)B * m->in_feat * 4, 256);
  return per;
}

static void grad_tc_carve(const dbx_model* m, int64_t B, int ns, bool with_adv, char* p, GradTcBufs& g, int nwb = 1) {
  const int nl = (int)g.dl.size();
  auto take = [&](size_t bytes) { char* r = p; p += round_up(bytes, 256); return r; };
  g.xin = (__nv_bfloat16*)take((size_t)B * m->in_feat * 2);
  g.pbar = (float*)take((size_t)B * m->out_feat * 4);
  g.acts.resize(std::max(0, nl - 1));
  for (int i = 0; i + 1 < nl; ++i) g.acts[i] = (__nv_bfloat16*)take((size_t)ns * round_up((size_t)B * g.dl[i]->units * 2, 256));
  g.logits = (float*)take((size_t)ns * B * m->out_feat * 4);
  g.dz16 = (__nv_bfloat16*)take((size_t)ns * B * g.Cp * 2);
  g.dbuf[0] = (__nv_bfloat16*)take((size_t)ns * B * g.hmax * 2);
  g.dbuf[1] = (__nv_bfloat16*)take((size_t)ns * B * g.hmax * 2);
  for (int b = 0; b < nwb; ++b) {
    g.W16b[b] = (__nv_bfloat16*)take((size_t)ns * m->P16 * 2);
    g.B32b[b] = (float*)take((size_t)ns * m->PB * 4);
  }
  g.W16 = g.W16b[0]; g.B32 = g.B32b[0];
  if (with_adv) g.adv = (__nv_bfloat16*)take((size_t)ns * B * m->in_feat * 2);
  else g.dx = (float*)take((size_t)ns * B * m->in_feat * 4);
}

// forward of nc samples, hidden outputs kept in g.acts, logits [nc][B][C] fp32
static int grad_tc_forward(dbx_model* m, const GradTcBufs& g, const __nv_bfloat16* in, int64_t in_ss, int64_t B, int nc, cudaStream_t st) {
  const int nl = (int)g.dl.size();
  const __nv_bfloat16* cur = in; int64_t cur_ss = in_ss;
  for (int i = 0; i < nl; ++i) {
    const Layer& L = *g.dl[i];
    const bool last = i == nl - 1;
    prof_begin(st);
    if (dense_tc_launch(cur, cur_ss, (int)L.in_feat, cur_ss == 0 ? 1 : 0, g.W16 + L.w16_off, m->P16, (int)L.w_cols_pad, nc, (int)B, L.units,
                        (int)L.in_feat, g.B32, m->PB, L.b32_off, last ? DBX_ACT_LINEAR : L.act, last ? nullptr : g.acts[i], B * L.units, L.units,
                        last ? g.logits : nullptr, B * m->out_feat, st, B <= 128 ? 128 : 256)) {
      set_err("dense_tc launch failed (gradient forward)"); return 1;
    }
    prof_end(st, 2.0 * (double)L.w_rows * L.w_cols * (double)B * nc);
    if (!last) { cur = g.acts[i]; cur_ss = B * L.units; }
  }
  return 0;
}

// backward of nc samples from g.dz16 down to g.dx [nc][B][K0] fp32 -- or, with fgsm_x, to every sample's adversarial input
// g.adv [nc][B][K0] bf16 (the FGSM step in the epilogue of the last GEMM)
static int grad_tc_backward(dbx_model* m, const GradTcBufs& g, int64_t B, int nc, cudaStream_t st, const float* fgsm_x = nullptr,
                            float eps = 0.f, float lower = 0.f, float upper = 0.f) {
  const int nl = (int)g.dl.size();
  const __nv_bfloat16* dz = g.dz16; int64_t dz_ss = B * g.Cp; int ldz = g.Cp; int zi = 0;
  for (int i = nl - 1; i >= 0; --i) {
    const Layer& L = *g.dl[i];
    const bool first = i == 0;
    prof_begin(st);
    if (dense_tc_bwd_launch(dz, dz_ss, ldz, g.W16 + L.w16_off, m->P16, (int)L.w_cols_pad, nc, (int)B, (int)L.in_feat, L.units,
                            first ? nullptr : g.acts[i - 1], first ? 0 : B * g.dl[i - 1]->units, first ? 0 : g.dl[i - 1]->units,
                            first ? DBX_ACT_LINEAR : g.dl[i - 1]->act, first ? (fgsm_x ? g.adv : nullptr) : g.dbuf[zi], B * L.in_feat,
                            (int)L.in_feat, (first && !fgsm_x) ? g.dx : nullptr, B * L.in_feat, st, first ? fgsm_x : nullptr, eps, lower, upper)) {
      set_err("dense_tc_bwd launch failed"); return 1;
    }
    prof_end(st, 2.0 * (double)L.w_rows * L.w_cols * (double)B * nc);
    if (!first) { dz = g.dbuf[zi]; dz_ss = B * L.in_feat; ldz = (int)L.in_feat; zi ^= 1; }
  }
  return 0;
}

static int run_input_gradient_tc(dbx_model* m, const float* x_dev, int64_t B, const int32_t* labels_dev, int64_t n_outer, int64_t n_inner,
                                 const Source& src0, float* grad_dev, cudaStream_t st) {
  GradTcBufs g;
  const int nc = (int)n_inner;
  const int64_t C = m->out_feat, K0 = m->in_feat;
  // several iterations: the next iteration's weights are drawn on the side stream while this one's GEMM chain runs on the
  // engine's high-priority stream (the same pipeline as the FGSM loop below)
  const bool overlap = opt().generic_overlap != 0 && n_outer > 1;
  const int nwb = overlap ? 2 : 1;
  const size_t per = grad_tc_per_sample(m, B, false, g, nwb);
  const size_t need = (size_t)nc * per + round_up((size_t)B * K0 * 2, 256) + round_up((size_t)B * C * 4, 256) + 8192;
  DBX_CHECK(need <= ws_budget_bytes() * 2, "n_inner = %d samples need %zu MB of workspace", nc, need >> 20);
  if (ensure_ws(m->ws, need)) return 1;
  grad_tc_carve(m, B, nc, false, (char*)m->ws.ptr, g, nwb);
  cudaStream_t sst = st, caller = st;
  if (overlap) {
    if (ensure_side_streams(m) || ensure_main_stream(m)) return 1;
    sst = m->ibp_side;
    DBX_CUDA(cudaEventRecord(m->ibp_ev[0], caller));
    DBX_CUDA(cudaStreamWaitEvent(sst, m->ibp_ev[0], 0));
    DBX_CUDA(cudaStreamWaitEvent(m->ibp_main, m->ibp_ev[0], 0));
    st = m->ibp_main;
  }
  auto draw = [&](int64_t it) -> int {
    const int b = (int)(it & 1) % nwb;
    if (overlap && it >= 2) DBX_CUDA(cudaStreamWaitEvent(sst, m->ibp_ev[3 + b], 0));
    if (launch_sample_bf16(m, m->table, src0, it * n_inner, nc, g.W16b[b], g.B32b[b], sst)) return 1;
    if (overlap) DBX_CUDA(cudaEventRecord(m->ibp_ev[1 + b], sst));
    return 0;
  };
  DBX_LAUNCH(cast_kernel<__nv_bfloat16>, grid_for(B * K0), 256, 0, st, x_dev, g.xin, B * K0);
  for (int64_t it = 0; it < n_outer; ++it) {
    const int wb = (int)(it & 1) % nwb;
    if (it == 0 || !overlap) { if (draw(it)) return 1; }
    if (overlap) {
      if (it + 1 < n_outer && draw(it + 1)) return 1;
      DBX_CUDA(cudaStreamWaitEvent(st, m->ibp_ev[1 + wb], 0));
    }
    g.W16 = g.W16b[wb]; g.B32 = g.B32b[wb];
    if (grad_tc_forward(m, g, g.xin, 0, B, nc, st)) return 1;
    DBX_LAUNCH(softmax_mean_kernel, (int)cdiv(B, 128), 128, 0, st, g.logits, nc, (int)B, (int)C, g.pbar);
    DBX_LAUNCH(ce_mean_upstream_bf16_kernel, grid_for((int64_t)nc * B), 256, 0, st, (const float*)g.logits, (const float*)g.pbar, labels_dev,
               nc, (int)B, (int)C, g.Cp, g.dz16);
    if (grad_tc_backward(m, g, B, nc, st)) return 1;
    if (overlap) DBX_CUDA(cudaEventRecord(m->ibp_ev[3 + wb], st));
    DBX_LAUNCH(sum_samples_kernel, grid_for(B * K0), 256, 0, st, (const float*)g.dx, nc, B * K0, it > 0 ? 1 : 0, grad_dev);
  }
  if (overlap) {      // join
    DBX_CUDA(cudaEventRecord(m->ibp_ev[5], st));
    DBX_CUDA(cudaStreamWaitEvent(caller, m->ibp_ev[5], 0));
    DBX_CUDA(cudaEventRecord(m->ibp_ev[0], sst));
    DBX_CUDA(cudaStreamWaitEvent(caller, m->ibp_ev[0], 0));
  }
  return 0;
}

static int run_count_fgsm_tc(dbx_model* m, const float* x_dev, int64_t B, const int32_t* attack_labels_dev, const int32_t* labels_dev,
                             float eps, float lower, float upper, int64_t n, const Source& src0, int accumulate, uint8_t* flags_dev,
                             unsigned long long* counts_dev, cudaStream_t st, float* probs_out = nullptr) {
  GradTcBufs g;
  const int64_t C = m->out_feat, K0 = m->in_feat;
  // the weights of chunk c + 1 are drawn on a side stream while chunk c is in the GEMMs (two weight buffers; events
  // ibp_ev[1..4] = sampled[2], consumed[2], as in the generic executor)
  const bool overlap = opt().generic_overlap != 0;
  const int nwb = overlap ? 2 : 1;
  const size_t per = grad_tc_per_sample(m, B, true, g, nwb);
  const size_t fixed = round_up((size_t)B * K0 * 2, 256) + round_up((size_t)B * C * 4, 256) + 8192;
  int64_t ns = (int64_t)std::max<size_t>(1, (ws_budget_bytes() - std::min(ws_budget_bytes(), fixed)) / per);
  ns = std::min<int64_t>(std::min<int64_t>(ns, n), 256);
  if (ensure_ws(m->ws, (size_t)ns * per + fixed)) return 1;
  grad_tc_carve(m, B, (int)ns, true, (char*)m->ws.ptr, g, nwb);
  cudaStream_t sst = st, caller = st;
  if (overlap) {      // GEMM chain on the engine's high-priority stream, sampling on the side stream, both joined with the caller's
    if (ensure_side_streams(m) || ensure_main_stream(m)) return 1;
    sst = m->ibp_side;
    DBX_CUDA(cudaEventRecord(m->ibp_ev[0], caller));
    DBX_CUDA(cudaStreamWaitEvent(sst, m->ibp_ev[0], 0));
    DBX_CUDA(cudaStreamWaitEvent(m->ibp_main, m->ibp_ev[0], 0));
    st = m->ibp_main;
  }
  auto draw_chunk = [&](int64_t cj) -> int {
    const int b = (int)(cj & 1) % nwb;
    const int64_t d0 = cj * ns;
    const int dn = (int)std::min<int64_t>(ns, n - d0);
    if (overlap && cj >= 2) DBX_CUDA(cudaStreamWaitEvent(sst, m->ibp_ev[3 + b], 0));
    if (launch_sample_bf16(m, m->table, src0, d0, dn, g.W16b[b], g.B32b[b], sst)) return 1;
    if (overlap) DBX_CUDA(cudaEventRecord(m->ibp_ev[1 + b], sst));
    return 0;
  };
  DBX_LAUNCH(cast_kernel<__nv_bfloat16>, grid_for(B * K0), 256, 0, st, x_dev, g.xin, B * K0);
  if (!accumulate && counts_dev) DBX_LAUNCH(zero_u64_kernel, grid_for(B), 256, 0, st, counts_dev, B);
  for (int64_t c0 = 0; c0 < n; c0 += ns) {
    const int nc = (int)std::min<int64_t>(ns, n - c0);
    const int64_t ci = c0 / ns;
    const int wb = (int)(ci & 1) % nwb;
    if (ci == 0 || !overlap) { if (draw_chunk(ci)) return 1; }
    if (overlap) {
      if (c0 + ns < n && draw_chunk(ci + 1)) return 1;
      DBX_CUDA(cudaStreamWaitEvent(st, m->ibp_ev[1 + wb], 0));
    }
    g.W16 = g.W16b[wb]; g.B32 = g.B32b[wb];
    if (grad_tc_forward(m, g, g.xin, 0, B, nc, st)) return 1;
    DBX_LAUNCH(softmax_ce_per_sample_bf16_kernel, grid_for((int64_t)nc * B), 256, 0, st, g.logits, attack_labels_dev, nc, (int)B, (int)C, g.Cp,
               g.dz16);
    if (grad_tc_backward(m, g, B, nc, st, x_dev, eps, lower, upper)) return 1;      // ... ends in every sample's adversarial input
    if (grad_tc_forward(m, g, g.adv, B * K0, B, nc, st)) return 1;
    if (overlap) DBX_CUDA(cudaEventRecord(m->ibp_ev[3 + wb], st));
    if (counts_dev)
      DBX_LAUNCH(count_kernel, grid_for((int64_t)nc * B), 256, 0, st, (const float*)g.logits, nc, B, (int)C, labels_dev,
                 flags_dev ? flags_dev + c0 * B : nullptr, counts_dev);
    if (probs_out) {
      float* dst = probs_out + c0 * B * C;
      DBX_CUDA(cudaMemcpyAsync(dst, g.logits, (size_t)nc * B * C * 4, cudaMemcpyDeviceToDevice, st));
      DBX_LAUNCH(act_rows_kernel, grid_for((int64_t)nc * B), 256, 0, st, (int)DBX_ACT_SOFTMAX, dst, (int64_t)nc * B, C);
    }
  }
  if (overlap) {      // join
    DBX_CUDA(cudaEventRecord(m->ibp_ev[5], st));
    DBX_CUDA(cudaStreamWaitEvent(caller, m->ibp_ev[5], 0));
    DBX_CUDA(cudaEventRecord(m->ibp_ev[0], sst));
    DBX_CUDA(cudaStreamWaitEvent(caller, m->ibp_ev[0], 0));
  }
  return 0;
}

}  // namespace dbx
