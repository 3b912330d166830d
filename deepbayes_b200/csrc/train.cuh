// Bayes-by-Backprop training step on the device (included by engine.cu).
//
// Reference: deepbayes/optimizers/bayesbybackprop.py:46-170 (robust_train == 0) + losses.py:10-45.  One mini-batch:
//   noise ~ N(0,1) per weight (Philox, or injected), W = mean + softplus(rho) * noise            (:50-59)
//   predictions = model_W(features)                                                               (:65)
//   loss = loss_fn(labels, predictions) + kl_weight * (log_q(W; mean, rho) - log_q(W; prior))    (losses.py:40-45)
//   weight_gradient = dLoss/dW, mean_gradient = dLoss/dmean, var_gradient = dLoss/drho            (:138-140, tf.GradientTape)
//   mean -= lrate * (weight_gradient + mean_gradient)                                             (:146, :158-161)
//   rho  -= lrate * (noise * sigmoid(rho) * weight_gradient + var_gradient)                       (:148-150, :157-160)
// Here the tape is replaced by explicit kernels: fp32 forward keeping the layer outputs, softmax + sparse categorical
// cross-entropy, backward GEMMs (dW = A^T dZ, dA = dZ W^T) on CUDA cores with ascending-k fp32 accumulation, and one
// element-wise kernel that adds the closed-form log_q terms and applies the update in place.  Dense MLPs with a
// softmax output; log_q takes the MEAN over a tensor's elements (losses.py:15), hence the 1/n_i factors, and the
// prior's scale is softplus(prior_var) (the prior goes through the same log_q, losses.py:42).
#pragma once

namespace dbx {

// W = mean + softplus(rho) * noise, noise from Philox (counter = flat index / 4, sample id = step) or injected
__global__ void bbb_sample_kernel(const float* __restrict__ mean, const float* __restrict__ rho, const float* __restrict__ noise_in,
                                  uint64_t seed, int64_t step, int64_t P, float* __restrict__ W, float* __restrict__ noise_out) {
  const int64_t nblk = (P + 3) >> 2;
  for (int64_t blk = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; blk < nblk; blk += (int64_t)gridDim.x * blockDim.x) {
    float z[4];
    if (noise_in == nullptr) philox_normal4(seed, (uint64_t)step, (uint64_t)blk, z);
    for (int i = 0; i < 4; ++i) {
      const int64_t j = blk * 4 + i;
      if (j >= P) break;
      const float e = noise_in ? noise_in[j] : z[i];
      W[j] = __fadd_rn(mean[j], __fmul_rn(softplus_tf(rho[j]), e));     // tf.add(mean, tf.multiply(softplus(rho), noise)), :54-56
      noise_out[j] = e;
    }
  }
}

// dlogits = (softmax(z) - onehot) / B (gradient of the batch-mean cross-entropy; zero where Keras' clip of p to
// [1e-7, 1-1e-7] is active), probabilities out, per-row -log(clip(p[label])) for the loss value
__global__ void softmax_ce_kernel(const float* __restrict__ logits, const int32_t* __restrict__ labels, int B, int C,
                                  float* __restrict__ probs, float* __restrict__ dz, float* __restrict__ row_loss) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* z = logits + (int64_t)b * C;
  float m = z[0];
  for (int c = 1; c < C; ++c) m = fmaxf(m, z[c]);
  float den = 0.f;
  for (int c = 0; c < C; ++c) den += expf(z[c] - m);
  const int lab = labels[b];
  const float pl = expf(z[lab] - m) / den;
  const float active = (pl > 1e-7f && pl < 1.f - 1e-7f) ? 1.f : 0.f;
  row_loss[b] = -logf(fminf(fmaxf(pl, 1e-7f), 1.f - 1e-7f));
  for (int c = 0; c < C; ++c) {
    const float p = expf(z[c] - m) / den;
    probs[(int64_t)b * C + c] = p;
    dz[(int64_t)b * C + c] = (p - (c == lab ? 1.f : 0.f)) * active / (float)B;
  }
}

// dW[k, n] = sum_b A[b, k] * dZ[b, n]   (A: [B, K], dZ: [B, N]); 32x32 tiles, ascending b
__global__ void __launch_bounds__(256)
grad_w_kernel(const float* __restrict__ A, const float* __restrict__ dZ, int B, int K, int N, float* __restrict__ dW) {
  __shared__ float As[32][33];
  __shared__ float Zs[32][33];
  const int k0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int b0 = 0; b0 < B; b0 += 32) {
    for (int i = ty; i < 32; i += 8) {
      const int b = b0 + i;
      As[i][tx] = (b < B && k0 + tx < K) ? A[(int64_t)b * K + k0 + tx] : 0.f;
      Zs[i][tx] = (b < B && n0 + tx < N) ? dZ[(int64_t)b * N + n0 + tx] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int i = 0; i < 32; ++i) {
      const float z = Zs[i][tx];
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[r] = fmaf(As[i][ty + 8 * r], z, acc[r]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int k = k0 + ty + 8 * r, n = n0 + tx;
    if (k < K && n < N) dW[(int64_t)k * N + n] = acc[r];
  }
}

// db[n] = sum_b dZ[b, n]
__global__ void grad_b_kernel(const float* __restrict__ dZ, int B, int N, float* __restrict__ db) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float a = 0.f;
  for (int b = 0; b < B; ++b) a += dZ[(int64_t)b * N + n];
  db[n] = a;
}

// dZprev[b, k] = (sum_n dZ[b, n] * W[k, n]) * act'(Aprev[b, k])   (W: [K, N] row-major)
__global__ void __launch_bounds__(256)
grad_a_kernel(const float* __restrict__ dZ, const float* __restrict__ W, const float* __restrict__ Aprev, int B, int K, int N,
              int act, float* __restrict__ dZprev) {
  __shared__ float Zs[32][33];
  __shared__ float Ws[32][33];
  const int b0 = blockIdx.y * 32, k0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int n0 = 0; n0 < N; n0 += 32) {
    for (int i = ty; i < 32; i += 8) {
      Zs[i][tx] = (b0 + i < B && n0 + tx < N) ? dZ[(int64_t)(b0 + i) * N + n0 + tx] : 0.f;
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
      const float a = Aprev[(int64_t)b * K + k];
      float d;
      switch (act) {
        case DBX_ACT_RELU: d = a > 0.f ? 1.f : 0.f; break;
        case DBX_ACT_TANH: d = 1.f - a * a; break;
        case DBX_ACT_SIGMOID: d = a * (1.f - a); break;
        default: d = 1.f;
      }
      dZprev[(int64_t)b * K + k] = acc[r] * d;
    }
  }
}

struct TrainTensors {
  int n;
  int32_t off[DBX_MAX_TENSORS], end[DBX_MAX_TENSORS];
};

// Element-wise: add the log_q terms to the data gradient and update mean / rho in place; accumulate the two log_q sums
// (per-block partials, summed in block order by the finish kernel: deterministic).
__global__ void __launch_bounds__(256)
bbb_update_kernel(float* __restrict__ mean, float* __restrict__ rho, const float* __restrict__ W, const float* __restrict__ noise,
                  const float* __restrict__ gdata, const float* __restrict__ prior_mean, const float* __restrict__ prior_var,
                  int64_t P, TrainTensors tt, float lrate, float kl_weight, float* __restrict__ partial /*[grid][2]*/) {
  __shared__ float red[2][256];
  float lq_post = 0.f, lq_prior = 0.f;
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < P; j += (int64_t)gridDim.x * blockDim.x) {
    int ti = 0;
    while (ti + 1 < tt.n && j >= tt.end[ti]) ++ti;
    const float inv_n = 1.f / (float)(tt.end[ti] - tt.off[ti]);
    const float r = rho[j], m = mean[j], w = W[j], e = noise[j];
    const float sg = softplus_tf(r);
    const float ps = softplus_tf(prior_var[j]);                      // losses.py:12 applied to the prior (:42)
    const float pm = prior_mean[j];
    const float sig = 1.f / (1.f + expf(-r));
    const float dq = (w - m) / (sg * sg);                            // -d log_q / dW = +d log_q / dmean   (times n_i)
    const float wg = gdata[j] + kl_weight * (-dq + (w - pm) / (ps * ps)) * inv_n;
    const float mg = kl_weight * dq * inv_n;
    const float vg = kl_weight * ((w - m) * (w - m) / (sg * sg * sg) - 1.f / sg) * sig * inv_n;
    mean[j] = m - lrate * (wg + mg);
    rho[j] = r - lrate * (e * sig * wg + vg);
    const float zq = (w - m) / sg, zp = (w - pm) / ps;
    lq_post += (-0.5f * zq * zq - logf(sg) - 0.9189385332046727f) * inv_n;
    lq_prior += (-0.5f * zp * zp - logf(ps) - 0.9189385332046727f) * inv_n;
  }
  red[0][threadIdx.x] = lq_post; red[1][threadIdx.x] = lq_prior;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) { red[0][threadIdx.x] += red[0][threadIdx.x + s]; red[1][threadIdx.x] += red[1][threadIdx.x + s]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { partial[2 * blockIdx.x] = red[0][0]; partial[2 * blockIdx.x + 1] = red[1][0]; }
}

// out[0] = loss = mean(row_loss) + kl_weight * (lq_post - lq_prior); out[1] = kl component
__global__ void bbb_loss_kernel(const float* __restrict__ row_loss, int B, const float* __restrict__ partial, int nblocks,
                                float kl_weight, float* __restrict__ out) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  float data = 0.f;
  for (int b = 0; b < B; ++b) data += row_loss[b];
  data /= (float)B;
  float qp = 0.f, pr = 0.f;
  for (int i = 0; i < nblocks; ++i) { qp += partial[2 * i]; pr += partial[2 * i + 1]; }
  out[0] = data + kl_weight * (qp - pr);
  out[1] = qp - pr;
}

static int run_bbb_step(dbx_model* m, const float* x_dev, int64_t B, const int32_t* labels_dev, float* mean_dev, float* rho_dev,
                        const float* prior_mean_dev, const float* prior_var_dev, const float* noise_dev, uint64_t seed, int64_t step,
                        float lrate, float kl_weight, float* loss_out_dev, float* probs_out_dev, cudaStream_t st) {
  DBX_CHECK(m->is_mlp, "the device training step covers Dense MLPs (bayesbybackprop.py:46-170)");
  const Layer& Lout = m->layers[m->last_weighted];
  DBX_CHECK(Lout.act == DBX_ACT_SOFTMAX, "the device training step needs a softmax output (sparse categorical cross-entropy)");
  const int64_t P = m->P, C = m->out_feat;
  // workspace: W | noise | gdata | activations of every layer (outputs) | dZ ping-pong | logits | row_loss | partials
  std::vector<const Layer*> dl;
  for (auto& L : m->layers) if (L.kind == DBX_DENSE) dl.push_back(&L);
  const int nl = (int)dl.size();
  size_t act_total = 0;
  for (auto* L : dl) act_total += round_up((size_t)B * L->units * 4, 256);
  const size_t pb = round_up((size_t)P * 4, 256), dzb = round_up((size_t)B * m->max_feat * 4, 256);
  const int ublocks = 148 * 4;
  const size_t need = 3 * pb + act_total + 2 * dzb + round_up((size_t)B * 4, 256) + round_up((size_t)ublocks * 8, 256) + 4096;
  if (ensure_ws(m->ws, need)) return 1;
  char* p = (char*)m->ws.ptr;
  float* W = (float*)p; p += pb;
  float* noise = (float*)p; p += pb;
  float* gdata = (float*)p; p += pb;
  std::vector<float*> acts(nl);
  for (int i = 0; i < nl; ++i) { acts[i] = (float*)p; p += round_up((size_t)B * dl[i]->units * 4, 256); }
  float* dz[2]; dz[0] = (float*)p; p += dzb; dz[1] = (float*)p; p += dzb;
  float* row_loss = (float*)p; p += round_up((size_t)B * 4, 256);
  float* partial = (float*)p;

  DBX_LAUNCH(bbb_sample_kernel, grid_for((P + 3) / 4), 256, 0, st, mean_dev, rho_dev, noise_dev, seed, step, P, W, noise);
  // ---- forward (fp32, layer outputs kept)
  const float* cur = x_dev;
  for (int i = 0; i < nl; ++i) {
    const Layer& L = *dl[i];
    const bool last = (i == nl - 1);
    dim3 grid((unsigned)cdiv(L.units, 64), (unsigned)cdiv(B, 64), 1);
    DBX_LAUNCH((dense_kernel<float, float>), grid, 256, 0, st, cur, (int64_t)0, (const float*)W, (int64_t)0, (const int32_t*)nullptr,
               (int64_t)0, L.w_off, (int)L.w_cols, (const float*)W, (int64_t)0, L.b_off, acts[i], (int64_t)0, (int)B, L.units,
               (int)L.in_feat, last ? DBX_ACT_LINEAR : L.act);
    cur = acts[i];
  }
  // ---- loss gradient at the logits; acts[nl-1] becomes the probabilities
  float* probs = probs_out_dev ? probs_out_dev : dz[1];
  DBX_LAUNCH(softmax_ce_kernel, (int)cdiv(B, 128), 128, 0, st, acts[nl - 1], labels_dev, (int)B, (int)C, probs, dz[0], row_loss);
  // ---- backward
  int zi = 0;
  for (int i = nl - 1; i >= 0; --i) {
    const Layer& L = *dl[i];
    const float* a_prev = (i == 0) ? x_dev : acts[i - 1];
    const int K = (int)L.in_feat, N = L.units;
    dim3 gw((unsigned)cdiv(N, 32), (unsigned)cdiv(K, 32));
    DBX_LAUNCH(grad_w_kernel, gw, 256, 0, st, a_prev, dz[zi], (int)B, K, N, gdata + L.w_off);
    DBX_LAUNCH(grad_b_kernel, (int)cdiv(N, 128), 128, 0, st, dz[zi], (int)B, N, gdata + L.b_off);
    if (i > 0) {
      dim3 ga((unsigned)cdiv(K, 32), (unsigned)cdiv(B, 32));
      // dz[1] may hold the probabilities of this call only when the caller did not ask for them; they are dead by now
      DBX_LAUNCH(grad_a_kernel, ga, 256, 0, st, dz[zi], (const float*)W + L.w_off, a_prev, (int)B, K, N, dl[i - 1]->act, dz[zi ^ 1]);
      zi ^= 1;
    }
  }
  // ---- log_q terms + update
  TrainTensors tt; tt.n = m->table.n;
  for (int i = 0; i < tt.n; ++i) { tt.off[i] = m->table.src_off[i]; tt.end[i] = m->table.src_end[i]; }
  DBX_LAUNCH(bbb_update_kernel, ublocks, 256, 0, st, mean_dev, rho_dev, (const float*)W, (const float*)noise, (const float*)gdata,
             prior_mean_dev, prior_var_dev, P, tt, lrate, kl_weight, partial);
  if (loss_out_dev) DBX_LAUNCH(bbb_loss_kernel, 1, 32, 0, st, (const float*)row_loss, (int)B, (const float*)partial, ublocks, kl_weight, loss_out_dev);
  return 0;
}

}  // namespace dbx
