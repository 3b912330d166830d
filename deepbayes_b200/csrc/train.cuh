// Bayes-by-Backprop training step on the device (included by engine.cu).
//
// Reference: deepbayes/optimizers/bayesbybackprop.py:46-170 (robust_train 0, 1, 2) + losses.py:10-45.  One mini-batch:
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

// What changes from one mini-batch of an epoch to the next, kept in DEVICE memory so that one captured step can be replayed as a CUDA
// graph (dbx_bbb_train_epoch): the Philox step id, the first row of the mini-batch in the epoch's arrays, the step's index in the
// call.  Kernels that get a StepCtx take these from it and treat their pointers as the epoch's base pointers.
struct StepCtx { long long step, row0, idx; };
__global__ void step_ctx_advance_kernel(StepCtx* c, long long rows) {
  if (blockIdx.x == 0 && threadIdx.x == 0) { c->step += 1; c->row0 += rows; c->idx += 1; }
}
// x[row0 .. row0 + B) of the epoch's [N][K] fp32 array as the bf16 A operand of the first layer
__global__ void cast_rows_ctx_kernel(const float* __restrict__ x_base, const StepCtx* __restrict__ ctx, int64_t K, __nv_bfloat16* __restrict__ out,
                                     int64_t nelem) {
  const float* in = x_base + ctx->row0 * K;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nelem; i += (int64_t)gridDim.x * blockDim.x) out[i] = __float2bfloat16_rn(in[i]);
}

// W = mean + softplus(rho) * noise, noise from Philox (counter = flat index / 4, sample id = step) or injected
__global__ void bbb_sample_kernel(const float* __restrict__ mean, const float* __restrict__ rho, const float* __restrict__ noise_in,
                                  uint64_t seed, int64_t step, int64_t P, float* __restrict__ W, float* __restrict__ noise_out,
                                  const StepCtx* __restrict__ ctx = nullptr) {
  if (ctx) step = ctx->step;
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
                                  float* __restrict__ probs, float* __restrict__ dz, float* __restrict__ row_loss,
                                  const StepCtx* __restrict__ ctx = nullptr, int probs_epoch = 0) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  if (ctx) { labels += ctx->row0; if (probs_epoch) probs += ctx->row0 * C; }
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
grad_w_kernel(const float* __restrict__ A, const float* __restrict__ dZ, int B, int K, int N, float* __restrict__ dW, int accumulate) {
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
    if (k < K && n < N) dW[(int64_t)k * N + n] = accumulate ? dW[(int64_t)k * N + n] + acc[r] : acc[r];
  }
}

// db[n] = sum_b dZ[b, n].  One block per 32 columns, 32 row groups: thread (r, c) adds rows r, r + 32, ... of column c, the 32
// partial sums are added in row-group order (deterministic).  (One thread per column walking all B rows took 40-70 us at
// mini-batch 1024 -- more than the step's tensor-core GEMMs together; profiles/r2_train_step_launches.csv.)
constexpr int kGradBThreads = 1024;
template <typename T>
__device__ __forceinline__ void grad_b_body(const T* __restrict__ dZ, int ldz, int B, int N, float* __restrict__ db, int accumulate) {
  __shared__ float part[32][33];
  const int cl = threadIdx.x & 31, r = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  float a = 0.f;
  if (c < N)
    for (int b = r; b < B; b += 32) a += (float)dZ[(int64_t)b * ldz + c];
  part[r][cl] = a;
  __syncthreads();
  if (r == 0 && c < N) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += part[i][cl];
    db[c] = accumulate ? db[c] + s : s;
  }
}
__global__ void __launch_bounds__(kGradBThreads) grad_b_kernel(const float* __restrict__ dZ, int B, int N, float* __restrict__ db, int accumulate) {
  grad_b_body<float>(dZ, N, B, N, db, accumulate);
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

// robust_train == 2 (bayesbybackprop.py:91-101): FGSM against the step's own weights, direction = the predicted class
// (attacks.py:86-91 with direction = -1): dz_att[b][c] = (p[b][c] - [c == argmax p[b]]) / B
__global__ void attack_upstream_kernel(const float* __restrict__ logits, int B, int C, float* __restrict__ dz) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* z = logits + (int64_t)b * C;
  int best = 0; float m = z[0];
  for (int c = 1; c < C; ++c) if (z[c] > m) { m = z[c]; best = c; }
  float den = 0.f;
  for (int c = 0; c < C; ++c) den += expf(z[c] - m);
  const float pb = 1.f / den;
  const float active = (pb > 1e-7f && pb < 1.f - 1e-7f) ? 1.f / (float)B : 0.f;
  for (int c = 0; c < C; ++c) dz[(int64_t)b * C + c] = (expf(z[c] - m) / den - (c == best ? 1.f : 0.f)) * active;
}

__global__ void fgsm_step_kernel(const float* __restrict__ x, const float* __restrict__ g, int64_t n, float eps, float lower, float upper,
                                 float* __restrict__ adv) {
  for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < n; o += (int64_t)gridDim.x * blockDim.x) {
    const float xe = x[o], ge = g[o];
    float a = xe + eps * (ge > 0.f ? 1.f : (ge < 0.f ? -1.f : 0.f));
    a = fminf(fmaxf(a, xe - eps), xe + eps);
    adv[o] = fminf(fmaxf(a, lower), upper);
  }
}

// output = lambda * p + (1 - lambda) * p' (:94, :101); loss row, and the logit gradients of BOTH passes
__global__ void mix_ce_kernel(const float* __restrict__ logits, const float* __restrict__ logits_adv, const int32_t* __restrict__ labels, int B,
                              int C, float lambda, float* __restrict__ probs, float* __restrict__ dz, float* __restrict__ dz_adv,
                              float* __restrict__ row_loss) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* z = logits + (int64_t)b * C;
  const float* za = logits_adv + (int64_t)b * C;
  float m = z[0], ma = za[0];
  for (int c = 1; c < C; ++c) { m = fmaxf(m, z[c]); ma = fmaxf(ma, za[c]); }
  float den = 0.f, dena = 0.f;
  for (int c = 0; c < C; ++c) { den += expf(z[c] - m); dena += expf(za[c] - ma); }
  const int y = labels[b];
  const float py = expf(z[y] - m) / den, pay = expf(za[y] - ma) / dena;
  const float oy = lambda * py + (1.f - lambda) * pay;
  const float g = (oy > 1e-7f && oy < 1.f - 1e-7f) ? -1.f / ((float)B * oy) : 0.f;
  row_loss[b] = -logf(fminf(fmaxf(oy, 1e-7f), 1.f - 1e-7f));
  for (int c = 0; c < C; ++c) {
    const float p = expf(z[c] - m) / den, pa = expf(za[c] - ma) / dena;
    probs[(int64_t)b * C + c] = p;                                   // train_metric sees `predictions`, :166
    dz[(int64_t)b * C + c] = lambda * g * py * ((c == y ? 1.f : 0.f) - p);
    dz_adv[(int64_t)b * C + c] = (1.f - lambda) * g * pay * ((c == y ? 1.f : 0.f) - pa);
  }
}

// ---------------------------------------------------------------------------------------
// robust_train == 1 (bayesbybackprop.py:73-90): the interval forward of verifiers.IBP (:68-95) through the step's own
// weights, kept in (upper, lower) form per layer, and its backward.  With mu = (u + l)/2, r = (u - l)/2 of a layer's
// input: MU = mu W + b, R = r |W| (propagate_interval, :23-41), z_u = MU + R, z_l = MU - R.
// ---------------------------------------------------------------------------------------
__global__ void ibp_train_box_kernel(const float* __restrict__ x, int64_t n, float eps, float* __restrict__ U, float* __restrict__ L) {
  for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < n; o += (int64_t)gridDim.x * blockDim.x) {
    const float v = x[o];
    U[o] = fminf(fmaxf(v + eps, 0.f), 1.f);          // tf.clip_by_value(inp + eps, 0, 1), verifiers.py:70-71
    L[o] = fminf(fmaxf(v - eps, 0.f), 1.f);
  }
}

// U / L [B, N] = act(MU +- R); 32 x 32 output tiles, ascending k
__global__ void __launch_bounds__(256)
ibp_train_fwd_kernel(const float* __restrict__ Up, const float* __restrict__ Lp, const float* __restrict__ W, const float* __restrict__ bias,
                     int B, int K, int N, int act, float* __restrict__ U, float* __restrict__ L) {
  __shared__ float Ms[32][33];
  __shared__ float Rs[32][33];
  __shared__ float Ws[32][33];
  const int b0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  float am[4] = {0.f, 0.f, 0.f, 0.f}, ar[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k0 = 0; k0 < K; k0 += 32) {
    for (int i = ty; i < 32; i += 8) {
      const bool in = b0 + i < B && k0 + tx < K;
      const float u = in ? Up[(int64_t)(b0 + i) * K + k0 + tx] : 0.f, l = in ? Lp[(int64_t)(b0 + i) * K + k0 + tx] : 0.f;
      Ms[i][tx] = (u + l) / 2.f; Rs[i][tx] = (u - l) / 2.f;
      Ws[i][tx] = (k0 + i < K && n0 + tx < N) ? W[(int64_t)(k0 + i) * N + n0 + tx] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < 32; ++k) {
      const float w = Ws[k][tx], wa = fabsf(w);
#pragma unroll
      for (int r = 0; r < 4; ++r) { am[r] = fmaf(Ms[ty + 8 * r][k], w, am[r]); ar[r] = fmaf(Rs[ty + 8 * r][k], wa, ar[r]); }
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int b = b0 + ty + 8 * r, n = n0 + tx;
    if (b < B && n < N) {
      const float mu = am[r] + bias[n];
      U[(int64_t)b * N + n] = apply_act(act, mu + ar[r]);
      L[(int64_t)b * N + n] = apply_act(act, mu - ar[r]);
    }
  }
}

// worst[b][c] = lower on the label, upper elsewhere (:78-82)
__global__ void ibp_worst_kernel(const float* __restrict__ U, const float* __restrict__ L, const int32_t* __restrict__ labels, int B, int C,
                                 float* __restrict__ worst) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= B * C) return;
  worst[o] = (labels[o / C] == o % C) ? L[o] : U[o];
}
// gradient at the worst-case logits -> gradients at the upper / lower logits
__global__ void ibp_split_kernel(const float* __restrict__ dzw, const int32_t* __restrict__ labels, int B, int C, float* __restrict__ dzu,
                                 float* __restrict__ dzl) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= B * C) return;
  const bool lab = labels[o / C] == o % C;
  dzu[o] = lab ? 0.f : dzw[o];
  dzl[o] = lab ? dzw[o] : 0.f;
}

// dW[k, n] (+)= sum_b mu[b, k] dMU[b, n] + sign(W[k, n]) sum_b r[b, k] dR[b, n],  dMU = dzu + dzl, dR = dzu - dzl
__global__ void __launch_bounds__(256)
ibp_grad_w_kernel(const float* __restrict__ Up, const float* __restrict__ Lp, const float* __restrict__ dzu, const float* __restrict__ dzl,
                  const float* __restrict__ W, int B, int K, int N, float* __restrict__ dW, int accumulate) {
  __shared__ float Ms[32][33];
  __shared__ float Rs[32][33];
  __shared__ float Zm[32][33];
  __shared__ float Zr[32][33];
  const int k0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  float am[4] = {0.f, 0.f, 0.f, 0.f}, ar[4] = {0.f, 0.f, 0.f, 0.f};
  for (int b0 = 0; b0 < B; b0 += 32) {
    for (int i = ty; i < 32; i += 8) {
      const int b = b0 + i;
      const bool ia = b < B && k0 + tx < K, iz = b < B && n0 + tx < N;
      const float u = ia ? Up[(int64_t)b * K + k0 + tx] : 0.f, l = ia ? Lp[(int64_t)b * K + k0 + tx] : 0.f;
      const float zu = iz ? dzu[(int64_t)b * N + n0 + tx] : 0.f, zl = iz ? dzl[(int64_t)b * N + n0 + tx] : 0.f;
      Ms[i][tx] = (u + l) / 2.f; Rs[i][tx] = (u - l) / 2.f;
      Zm[i][tx] = zu + zl; Zr[i][tx] = zu - zl;
    }
    __syncthreads();
#pragma unroll 8
    for (int i = 0; i < 32; ++i) {
      const float zm = Zm[i][tx], zr = Zr[i][tx];
#pragma unroll
      for (int r = 0; r < 4; ++r) { am[r] = fmaf(Ms[i][ty + 8 * r], zm, am[r]); ar[r] = fmaf(Rs[i][ty + 8 * r], zr, ar[r]); }
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int k = k0 + ty + 8 * r, n = n0 + tx;
    if (k < K && n < N) {
      const float w = W[(int64_t)k * N + n];
      const float g = am[r] + (w > 0.f ? ar[r] : (w < 0.f ? -ar[r] : 0.f));      // d|w|/dw = sign(w)
      dW[(int64_t)k * N + n] = accumulate ? dW[(int64_t)k * N + n] + g : g;
    }
  }
}

// gradients at the previous layer's pre-activation bounds: dmu = dMU W^T, dr = dR |W|^T, du = (dmu + dr)/2, dl = (dmu - dr)/2,
// each times act' evaluated from the previous layer's OUTPUT bound (Up / Lp)
__global__ void __launch_bounds__(256)
ibp_grad_a_kernel(const float* __restrict__ dzu, const float* __restrict__ dzl, const float* __restrict__ W, const float* __restrict__ Up,
                  const float* __restrict__ Lp, int B, int K, int N, int act, float* __restrict__ dzu_prev, float* __restrict__ dzl_prev) {
  __shared__ float Zm[32][33];
  __shared__ float Zr[32][33];
  __shared__ float Ws[32][33];
  const int b0 = blockIdx.y * 32, k0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  float am[4] = {0.f, 0.f, 0.f, 0.f}, ar[4] = {0.f, 0.f, 0.f, 0.f};
  for (int n0 = 0; n0 < N; n0 += 32) {
    for (int i = ty; i < 32; i += 8) {
      const bool iz = b0 + i < B && n0 + tx < N;
      const float zu = iz ? dzu[(int64_t)(b0 + i) * N + n0 + tx] : 0.f, zl = iz ? dzl[(int64_t)(b0 + i) * N + n0 + tx] : 0.f;
      Zm[i][tx] = zu + zl; Zr[i][tx] = zu - zl;
      Ws[i][tx] = (k0 + i < K && n0 + tx < N) ? W[(int64_t)(k0 + i) * N + n0 + tx] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int n = 0; n < 32; ++n) {
      const float w = Ws[tx][n], wa = fabsf(w);
#pragma unroll
      for (int r = 0; r < 4; ++r) { am[r] = fmaf(Zm[ty + 8 * r][n], w, am[r]); ar[r] = fmaf(Zr[ty + 8 * r][n], wa, ar[r]); }
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int b = b0 + ty + 8 * r, k = k0 + tx;
    if (b < B && k < K) {
      const float u = Up[(int64_t)b * K + k], l = Lp[(int64_t)b * K + k];
      float du, dl;
      switch (act) {
        case DBX_ACT_RELU: du = u > 0.f ? 1.f : 0.f; dl = l > 0.f ? 1.f : 0.f; break;
        case DBX_ACT_TANH: du = 1.f - u * u; dl = 1.f - l * l; break;
        case DBX_ACT_SIGMOID: du = u * (1.f - u); dl = l * (1.f - l); break;
        default: du = 1.f; dl = 1.f;
      }
      dzu_prev[(int64_t)b * K + k] = (am[r] + ar[r]) / 2.f * du;
      dzl_prev[(int64_t)b * K + k] = (am[r] - ar[r]) / 2.f * dl;
    }
  }
}

// robust_train == 3 (bayesbybackprop.py:102-121): output = sum_k (1/K) softmax(worst_k), worst_k the worst-case IBP logits at
// the k-th draw of epsilon; loss row, the clean probabilities (train_metric, :166) and the gradient at every worst_k
__global__ void mc_ce_kernel(const float* __restrict__ logits, const float* __restrict__ worstK, int K, const int32_t* __restrict__ labels, int B,
                             int C, float* __restrict__ probs, float* __restrict__ dzwK, float* __restrict__ row_loss) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int y = labels[b];
  const float invK = 1.f / (float)K;
  float oy = 0.f;
  for (int k = 0; k < K; ++k) {
    const float* z = worstK + ((int64_t)k * B + b) * C;
    float m = z[0];
    for (int c = 1; c < C; ++c) m = fmaxf(m, z[c]);
    float den = 0.f;
    for (int c = 0; c < C; ++c) den += expf(z[c] - m);
    oy += invK * (expf(z[y] - m) / den);
  }
  const float g = (oy > 1e-7f && oy < 1.f - 1e-7f) ? -1.f / ((float)B * oy) : 0.f;
  row_loss[b] = -logf(fminf(fmaxf(oy, 1e-7f), 1.f - 1e-7f));
  for (int k = 0; k < K; ++k) {
    const float* z = worstK + ((int64_t)k * B + b) * C;
    float m = z[0];
    for (int c = 1; c < C; ++c) m = fmaxf(m, z[c]);
    float den = 0.f;
    for (int c = 0; c < C; ++c) den += expf(z[c] - m);
    const float py = expf(z[y] - m) / den;
    for (int c = 0; c < C; ++c)
      dzwK[((int64_t)k * B + b) * C + c] = invK * g * py * ((c == y ? 1.f : 0.f) - expf(z[c] - m) / den);
  }
  const float* z = logits + (int64_t)b * C;
  float m = z[0];
  for (int c = 1; c < C; ++c) m = fmaxf(m, z[c]);
  float den = 0.f;
  for (int c = 0; c < C; ++c) den += expf(z[c] - m);
  for (int c = 0; c < C; ++c) probs[(int64_t)b * C + c] = expf(z[c] - m) / den;
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
                                float kl_weight, float* __restrict__ out, const StepCtx* __restrict__ ctx = nullptr) {
  // one warp (launched with 32 threads): lane l adds elements l, l + 32, ..., then a fixed butterfly -- deterministic.  (One thread
  // walking B row losses and the update kernel's ~600 partial pairs took 50 us at mini-batch 1024.)
  if (blockIdx.x != 0 || threadIdx.x >= 32) return;
  const int lane = threadIdx.x;
  float data = 0.f, qp = 0.f, pr = 0.f;
  for (int b = lane; b < B; b += 32) data += row_loss[b];
  for (int i = lane; i < nblocks; i += 32) { qp += partial[2 * i]; pr += partial[2 * i + 1]; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    data += __shfl_xor_sync(0xffffffffu, data, o);
    qp += __shfl_xor_sync(0xffffffffu, qp, o);
    pr += __shfl_xor_sync(0xffffffffu, pr, o);
  }
  if (lane != 0) return;
  if (ctx) out += 2 * ctx->idx;
  data /= (float)B;
  out[0] = data + kl_weight * (qp - pr);
  out[1] = qp - pr;
}

static bool grad_tc_ok(const dbx_model* m);      // attack.cuh
static int run_bbb_step_tc_fwd(dbx_model* m, const float* x_dev, int64_t B, const int32_t* labels_dev, float* mean_dev, float* rho_dev,
                               const float* prior_mean_dev, const float* prior_var_dev, const float* noise_dev, uint64_t seed, int64_t step,
                               float lrate, float kl_weight, float* loss_out_dev, float* probs_out_dev, int robust_mode, float epsilon,
                               float robust_lambda, float in_lower, float in_upper, cudaStream_t st);
static int run_bbb_step(dbx_model* m, const float* x_dev, int64_t B, const int32_t* labels_dev, float* mean_dev, float* rho_dev,
                        const float* prior_mean_dev, const float* prior_var_dev, const float* noise_dev, uint64_t seed, int64_t step,
                        float lrate, float kl_weight, float* loss_out_dev, float* probs_out_dev, int robust_mode, float epsilon,
                        float robust_lambda, float in_lower, float in_upper, cudaStream_t st, int n_eps = 0, const float* eps_list = nullptr,
                        int mode = DBX_MODE_FP32) {
  DBX_CHECK(m->is_mlp, "the device training step covers Dense MLPs (bayesbybackprop.py:46-170)");
  DBX_CHECK(m->layers[m->last_weighted].act == DBX_ACT_SOFTMAX, "the device training step needs a softmax output (sparse categorical cross-entropy)");
  // bf16 mode: tensor-core GEMMs for the standard and the FGSM step when every Dense width suits TMA; the interval modes and
  // everything else stay on the fp32 kernels
  g_last_path = 0;
  if (mode == DBX_MODE_BF16 && (robust_mode == 0 || robust_mode == 2) && grad_tc_ok(m)) {
    g_last_path = 3;
    return run_bbb_step_tc_fwd(m, x_dev, B, labels_dev, mean_dev, rho_dev, prior_mean_dev, prior_var_dev, noise_dev, seed, step, lrate, kl_weight,
                               loss_out_dev, probs_out_dev, robust_mode, epsilon, robust_lambda, in_lower, in_upper, st);
  }
  DBX_CHECK(robust_mode >= 0 && robust_mode <= 3, "robust_train %d is not built (0 = standard, 1 = IBP, 2 = FGSM adversarial training, 3 = IBP over draws of epsilon)", robust_mode);
  DBX_CHECK(robust_mode != 3 || (n_eps >= 1 && n_eps <= 64 && eps_list), "robust_train 3 needs 1..64 epsilon draws");
  const Layer& Lout = m->layers[m->last_weighted];
  DBX_CHECK(Lout.act == DBX_ACT_SOFTMAX, "the device training step needs a softmax output (sparse categorical cross-entropy)");
  const int64_t P = m->P, C = m->out_feat, K0 = m->in_feat;
  std::vector<const Layer*> dl;
  for (auto& L : m->layers) if (L.kind == DBX_DENSE) dl.push_back(&L);
  const int nl = (int)dl.size();
  // workspace: W | noise | gdata | layer outputs (clean pass, adversarial pass) | x_adv | dZ ping-pong | probs | dz_adv | row_loss | partials
  size_t act_total = 0;
  for (auto* L : dl) act_total += round_up((size_t)B * L->units * 4, 256);
  const size_t pb = round_up((size_t)P * 4, 256), dzb = round_up((size_t)B * std::max<int64_t>(m->max_feat, K0) * 4, 256);
  const size_t xb = round_up((size_t)B * K0 * 4, 256), ob = round_up((size_t)B * C * 4, 256);
  const int ublocks = sm_count() * 4;
  const size_t need = 3 * pb + 3 * act_total + 2 * xb + 4 * dzb + (3 + 2 * (size_t)std::max(n_eps, 0)) * ob + round_up((size_t)B * 4, 256) + round_up((size_t)ublocks * 8, 256) + 4096;
  if (ensure_ws(m->ws, need)) return 1;
  char* p = (char*)m->ws.ptr;
  float* W = (float*)p; p += pb;            // (dbx_bbb_step copies the sampled weights out of this first block)
  float* noise = (float*)p; p += pb;
  float* gdata = (float*)p; p += pb;
  std::vector<float*> acts(nl), acts_adv(nl);
  for (int i = 0; i < nl; ++i) { acts[i] = (float*)p; p += round_up((size_t)B * dl[i]->units * 4, 256); }
  for (int i = 0; i < nl; ++i) { acts_adv[i] = (float*)p; p += round_up((size_t)B * dl[i]->units * 4, 256); }
  std::vector<float*> acts_low(nl);     // robust_train == 1: acts_adv holds the upper bounds, acts_low the lower ones
  for (int i = 0; i < nl; ++i) { acts_low[i] = (float*)p; p += round_up((size_t)B * dl[i]->units * 4, 256); }
  float* x_adv = (float*)p; p += xb;      // robust_train == 1: upper input bound
  float* x_low = (float*)p; p += xb;
  float* dz[2]; dz[0] = (float*)p; p += dzb; dz[1] = (float*)p; p += dzb;
  float* dzl[2]; dzl[0] = (float*)p; p += dzb; dzl[1] = (float*)p; p += dzb;
  float* worst = (float*)p; p += ob;
  float* worstK = (float*)p; p += (size_t)std::max(n_eps, 0) * ob;       // robust_train == 3: [K][B][C] (ob-pitched)
  float* dzwK = (float*)p; p += (size_t)std::max(n_eps, 0) * ob;
  float* probs_scratch = (float*)p; p += ob;
  float* dz_adv = (float*)p; p += ob;
  float* row_loss = (float*)p; p += round_up((size_t)B * 4, 256);
  float* partial = (float*)p;
  float* probs = probs_out_dev ? probs_out_dev : probs_scratch;

  DBX_LAUNCH(bbb_sample_kernel, grid_for((P + 3) / 4), 256, 0, st, mean_dev, rho_dev, noise_dev, seed, step, P, W, noise);
  // forward with the sampled weights, fp32, layer outputs kept
  auto forward = [&](const float* in, std::vector<float*>& out) -> int {
    const float* cur = in;
    for (int i = 0; i < nl; ++i) {
      const Layer& L = *dl[i];
      dim3 grid((unsigned)cdiv(L.units, 64), (unsigned)cdiv(B, 64), 1);
      DBX_LAUNCH((dense_kernel<float, float>), grid, 256, 0, st, cur, (int64_t)0, (const float*)W, (int64_t)0, (const int32_t*)nullptr,
                 (int64_t)0, L.w_off, (int)L.w_cols, (const float*)W, (int64_t)0, L.b_off, out[i], (int64_t)0, (int)B, L.units,
                 (int)L.in_feat, i == nl - 1 ? DBX_ACT_LINEAR : L.act);
      cur = out[i];
    }
    return 0;
  };
  // backward from dz[0] (the gradient at the logits).  to_input: propagate down to dX and return the index of the dz
  // buffer holding it; otherwise produce the weight / bias gradients into gdata (accumulate: add to what is there)
  auto backward = [&](const float* in, std::vector<float*>& a, bool to_input, int accumulate) -> int {
    int zi = 0;
    for (int i = nl - 1; i >= 0; --i) {
      const Layer& L = *dl[i];
      const float* a_prev = (i == 0) ? in : a[i - 1];
      const int K = (int)L.in_feat, N = L.units;
      if (!to_input) {
        dim3 gw((unsigned)cdiv(N, 32), (unsigned)cdiv(K, 32));
        DBX_LAUNCH(grad_w_kernel, gw, 256, 0, st, a_prev, (const float*)dz[zi], (int)B, K, N, gdata + L.w_off, accumulate);
        DBX_LAUNCH(grad_b_kernel, (int)cdiv(N, 32), kGradBThreads, 0, st, (const float*)dz[zi], (int)B, N, gdata + L.b_off, accumulate);
      }
      if (i > 0 || to_input) {
        dim3 ga((unsigned)cdiv(K, 32), (unsigned)cdiv(B, 32));
        DBX_LAUNCH(grad_a_kernel, ga, 256, 0, st, (const float*)dz[zi], (const float*)W + L.w_off, a_prev, (int)B, K, N,
                   i > 0 ? dl[i - 1]->act : DBX_ACT_LINEAR, dz[zi ^ 1]);
        zi ^= 1;
      }
    }
    return zi;
  };
  if (forward(x_dev, acts)) return 1;
  if (robust_mode == 0) {
    DBX_LAUNCH(softmax_ce_kernel, (int)cdiv(B, 128), 128, 0, st, (const float*)acts[nl - 1], labels_dev, (int)B, (int)C, probs, dz[0], row_loss);
    backward(x_dev, acts, false, 0);
  } else if (robust_mode == 1 || robust_mode == 3) {
    // interval forward through the step's own weights at one epsilon: bounds per layer in acts_adv (upper) / acts_low (lower),
    // worst-case logits into `wdst`
    auto ibp_forward = [&](float eps_k, float* wdst) -> int {
      DBX_LAUNCH(ibp_train_box_kernel, grid_for(B * K0), 256, 0, st, x_dev, B * K0, eps_k, x_adv, x_low);
      for (int i = 0; i < nl; ++i) {
        const Layer& L = *dl[i];
        dim3 grid((unsigned)cdiv(L.units, 32), (unsigned)cdiv(B, 32));
        DBX_LAUNCH(ibp_train_fwd_kernel, grid, 256, 0, st, (const float*)(i ? acts_adv[i - 1] : x_adv), (const float*)(i ? acts_low[i - 1] : x_low),
                   (const float*)W + L.w_off, (const float*)W + L.b_off, (int)B, (int)L.in_feat, L.units, i == nl - 1 ? DBX_ACT_LINEAR : L.act,
                   acts_adv[i], acts_low[i]);
      }
      DBX_LAUNCH(ibp_worst_kernel, (int)cdiv(B * C, 256), 256, 0, st, (const float*)acts_adv[nl - 1], (const float*)acts_low[nl - 1], labels_dev,
                 (int)B, (int)C, wdst);
      return 0;
    };
    // backward through the interval forward from the gradient at the worst-case logits
    auto ibp_backward = [&](const float* dzw, int accumulate) -> int {
    DBX_LAUNCH(ibp_split_kernel, (int)cdiv(B * C, 256), 256, 0, st, dzw, labels_dev, (int)B, (int)C, dz[0], dzl[0]);
    int zi = 0;
    for (int i = nl - 1; i >= 0; --i) {
      const Layer& L = *dl[i];
      const float* up = i ? acts_adv[i - 1] : x_adv;
      const float* lp = i ? acts_low[i - 1] : x_low;
      const int K = (int)L.in_feat, N = L.units;
      dim3 gw((unsigned)cdiv(N, 32), (unsigned)cdiv(K, 32));
      DBX_LAUNCH(ibp_grad_w_kernel, gw, 256, 0, st, up, lp, (const float*)dz[zi], (const float*)dzl[zi], (const float*)W + L.w_off, (int)B, K, N,
                 gdata + L.w_off, accumulate);
      DBX_LAUNCH(grad_b_kernel, (int)cdiv(N, 32), kGradBThreads, 0, st, (const float*)dz[zi], (int)B, N, gdata + L.b_off, accumulate);
      DBX_LAUNCH(grad_b_kernel, (int)cdiv(N, 32), kGradBThreads, 0, st, (const float*)dzl[zi], (int)B, N, gdata + L.b_off, 1);
      if (i > 0) {
        dim3 ga((unsigned)cdiv(K, 32), (unsigned)cdiv(B, 32));
        DBX_LAUNCH(ibp_grad_a_kernel, ga, 256, 0, st, (const float*)dz[zi], (const float*)dzl[zi], (const float*)W + L.w_off, up, lp, (int)B, K, N,
                   dl[i - 1]->act, dz[zi ^ 1], dzl[zi ^ 1]);
        zi ^= 1;
      }
    }
    return 0;
    };
    if (robust_mode == 1) {
      // the worst-case logits take the adversarial pass's place in the mixed loss
      if (ibp_forward(epsilon, worst)) return 1;
      DBX_LAUNCH(mix_ce_kernel, (int)cdiv(B, 128), 128, 0, st, (const float*)acts[nl - 1], (const float*)worst, labels_dev, (int)B, (int)C,
                 robust_lambda, probs, dz[0], dz_adv, row_loss);
      backward(x_dev, acts, false, 0);
      if (ibp_backward(dz_adv, 1)) return 1;
    } else {
      // robust_train == 3: the mean over the epsilon draws of softmax(worst_k); the clean prediction is not part of the loss.
      // Pass 1 keeps only the worst-case logits of every draw; pass 2 repeats each interval forward right before its backward.
      for (int k = 0; k < n_eps; ++k) if (ibp_forward(eps_list[k], worstK + (int64_t)k * B * C)) return 1;
      DBX_LAUNCH(mc_ce_kernel, (int)cdiv(B, 128), 128, 0, st, (const float*)acts[nl - 1], (const float*)worstK, n_eps, labels_dev, (int)B, (int)C,
                 probs, dzwK, row_loss);
      for (int k = 0; k < n_eps; ++k) {
        if (n_eps > 1 || k > 0) { if (ibp_forward(eps_list[k], worst)) return 1; }
        if (ibp_backward(dzwK + (int64_t)k * B * C, k > 0 ? 1 : 0)) return 1;
      }
    }
  } else {
    // robust_train == 2: FGSM against the step's own weights (attacks.py:81-102 with num_models = 1 -> current weights, :57-58)
    DBX_LAUNCH(attack_upstream_kernel, (int)cdiv(B, 128), 128, 0, st, (const float*)acts[nl - 1], (int)B, (int)C, dz[0]);
    const int zi = backward(x_dev, acts, true, 0);
    DBX_LAUNCH(fgsm_step_kernel, grid_for(B * K0), 256, 0, st, x_dev, (const float*)dz[zi], B * K0, epsilon, in_lower, in_upper, x_adv);
    if (forward(x_adv, acts_adv)) return 1;
    // loss on lambda * p + (1 - lambda) * p_adv; logit gradients of both passes
    DBX_LAUNCH(mix_ce_kernel, (int)cdiv(B, 128), 128, 0, st, (const float*)acts[nl - 1], (const float*)acts_adv[nl - 1], labels_dev, (int)B, (int)C,
               robust_lambda, probs, dz[0], dz_adv, row_loss);
    backward(x_dev, acts, false, 0);
    DBX_CUDA(cudaMemcpyAsync(dz[0], dz_adv, (size_t)B * C * 4, cudaMemcpyDeviceToDevice, st));
    backward(x_adv, acts_adv, false, 1);
  }
  DBX_CUDA(cudaGetLastError());
  // ---- log_q terms + update
  TrainTensors tt; tt.n = m->table.n;
  for (int i = 0; i < tt.n; ++i) { tt.off[i] = m->table.src_off[i]; tt.end[i] = m->table.src_end[i]; }
  DBX_LAUNCH(bbb_update_kernel, ublocks, 256, 0, st, mean_dev, rho_dev, (const float*)W, (const float*)noise, (const float*)gdata,
             prior_mean_dev, prior_var_dev, P, tt, lrate, kl_weight, partial);
  if (loss_out_dev) DBX_LAUNCH(bbb_loss_kernel, 1, 32, 0, st, (const float*)row_loss, (int)B, (const float*)partial, ublocks, kl_weight, loss_out_dev);
  return 0;
}

// ---------------------------------------------------------------------------------------
// bf16 mode of the training step (robust_train 0 and 2): the three GEMMs of every layer on the tensor cores --
//   forward   Z = A W + b            dense_tc_kernel (the inference GEMM, one "sample" = the step's own weights)
//   backward  dA = (dZ W^T) .* act'  dense_tc_bwd_kernel (K-major view of the Keras kernel; FGSM step in its epilogue for mode 2)
//   weights   dW = A^T dZ            dense_tc_wgrad_kernel (both operands MN-major as they lie in memory)
// with the sampled weights, the activations and the upstream gradients rounded to bf16 as GEMM operands and everything else
// (the draw W = mean + softplus(rho) * noise, biases, softmax / cross-entropy, the KL terms, the update of mean and rho) in fp32
// exactly as in the fp32 step.  The parameters themselves never leave fp32.
// ---------------------------------------------------------------------------------------
__global__ void dz_to_bf16_kernel(const float* __restrict__ dz, int B, int C, int Cp, __nv_bfloat16* __restrict__ out) {
  const int64_t total = (int64_t)B * Cp;
  for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(o % Cp);
    out[o] = __float2bfloat16_rn(c < C ? dz[(o / Cp) * C + c] : 0.f);
  }
}

// db[n] = sum_b dZ[b, n], dZ as the GEMMs see it (bf16), summed in fp32 in row order
__global__ void __launch_bounds__(kGradBThreads) grad_b_bf16_kernel(const __nv_bfloat16* __restrict__ dZ, int ldz, int B, int N,
                                                                    float* __restrict__ db, int accumulate) {
  grad_b_body<__nv_bfloat16>(dZ, ldz, B, N, db, accumulate);
}

static int run_bbb_step_tc(dbx_model* m, const float* x_dev, int64_t B, const int32_t* labels_dev, float* mean_dev, float* rho_dev,
                           const float* prior_mean_dev, const float* prior_var_dev, const float* noise_dev, uint64_t seed, int64_t step,
                           float lrate, float kl_weight, float* loss_out_dev, float* probs_out_dev, int robust_mode, float epsilon,
                           float robust_lambda, float in_lower, float in_upper, cudaStream_t st, const StepCtx* ctx = nullptr) {
  // ctx (device memory, robust_mode 0 only): this is the step a CUDA graph replays -- x_dev / labels_dev / probs_out_dev /
  // loss_out_dev are then the EPOCH's base pointers and the step id, first row and output slot come from *ctx
  const int64_t P = m->P, C = m->out_feat, K0 = m->in_feat;
  std::vector<const Layer*> dl;
  for (auto& L : m->layers) if (L.kind == DBX_DENSE) dl.push_back(&L);
  const int nl = (int)dl.size();
  const int Cp = (int)round_up(C, 8);
  int64_t hmax = 8;
  size_t act16_total = 0;
  for (int i = 0; i + 1 < nl; ++i) { act16_total += round_up((size_t)B * dl[i]->units * 2, 256); hmax = std::max<int64_t>(hmax, dl[i]->units); }
  const size_t pb = round_up((size_t)P * 4, 256), ob = round_up((size_t)B * C * 4, 256), x16b = round_up((size_t)B * K0 * 2, 256);
  const size_t dbb = round_up((size_t)B * hmax * 2, 256);
  const int ublocks = sm_count() * 4;
  const size_t need = 3 * pb + round_up((size_t)m->P16 * 2, 256) + round_up((size_t)m->PB * 4, 256) + 2 * x16b + 2 * act16_total + 5 * ob +
                      round_up((size_t)B * Cp * 2, 256) + 2 * dbb + round_up((size_t)B * 4, 256) + round_up((size_t)ublocks * 8, 256) + 4096;
  if (ensure_ws(m->ws, need)) return 1;
  char* p = (char*)m->ws.ptr;
  auto take = [&](size_t bytes) { char* r = p; p += round_up(bytes, 256); return r; };
  float* W = (float*)take(pb);              // (dbx_bbb_step copies the sampled weights out of this first block)
  float* noise = (float*)take(pb);
  float* gdata = (float*)take(pb);
  __nv_bfloat16* W16 = (__nv_bfloat16*)take((size_t)m->P16 * 2);
  float* B32 = (float*)take((size_t)m->PB * 4);
  __nv_bfloat16* xin = (__nv_bfloat16*)take(x16b);
  __nv_bfloat16* xadv = (__nv_bfloat16*)take(x16b);
  std::vector<__nv_bfloat16*> acts(std::max(0, nl - 1)), acts_adv(std::max(0, nl - 1));
  for (int i = 0; i + 1 < nl; ++i) acts[i] = (__nv_bfloat16*)take((size_t)B * dl[i]->units * 2);
  for (int i = 0; i + 1 < nl; ++i) acts_adv[i] = (__nv_bfloat16*)take((size_t)B * dl[i]->units * 2);
  float* logits = (float*)take(ob);
  float* logits_adv = (float*)take(ob);
  float* dz32 = (float*)take(ob);
  float* dz_adv32 = (float*)take(ob);
  float* probs_scratch = (float*)take(ob);
  __nv_bfloat16* dz16 = (__nv_bfloat16*)take((size_t)B * Cp * 2);
  __nv_bfloat16* dbuf[2] = {(__nv_bfloat16*)take(dbb), (__nv_bfloat16*)take(dbb)};
  float* row_loss = (float*)take((size_t)B * 4);
  float* partial = (float*)p;
  float* probs = probs_out_dev ? probs_out_dev : probs_scratch;

  DBX_LAUNCH(bbb_sample_kernel, grid_for((P + 3) / 4), 256, 0, st, mean_dev, rho_dev, noise_dev, seed, step, P, W, noise, ctx);
  {   // the step's weights in the GEMMs' bf16 layout (kernels padded / aligned for TMA, biases fp32): W read as ONE stored sample
    const float* keep = m->slab; const int64_t keepS = m->S, keepStride = m->slab_stride;
    m->slab = W; m->S = 1; m->slab_stride = P;
    Source src; src.stored = true; src.j0 = 0;
    const int rc = launch_sample_bf16(m, m->table, src, 0, 1, W16, B32, st);
    m->slab = keep; m->S = keepS; m->slab_stride = keepStride;
    if (rc) return 1;
  }
  if (ctx) DBX_LAUNCH(cast_rows_ctx_kernel, grid_for(B * K0), 256, 0, st, x_dev, ctx, K0, xin, B * K0);
  else DBX_LAUNCH(cast_kernel<__nv_bfloat16>, grid_for(B * K0), 256, 0, st, x_dev, xin, B * K0);

  auto forward = [&](const __nv_bfloat16* in, std::vector<__nv_bfloat16*>& out, float* z) -> int {
    const __nv_bfloat16* cur = in;
    for (int i = 0; i < nl; ++i) {
      const Layer& L = *dl[i];
      const bool last = i == nl - 1;
      if (dense_tc_launch(cur, 0, (int)L.in_feat, 1, W16 + L.w16_off, m->P16, (int)L.w_cols_pad, 1, (int)B, L.units, (int)L.in_feat, B32, m->PB,
                          L.b32_off, last ? DBX_ACT_LINEAR : L.act, last ? nullptr : out[i], B * L.units, L.units, last ? z : nullptr, B * C, st)) {
        set_err("dense_tc launch failed (training forward)"); return 1;
      }
      if (!last) cur = out[i];
    }
    return 0;
  };
  // from dz16 (bf16 [B][Cp], the gradient at the logits): weight / bias gradients into gdata, or (to_input) the FGSM input
  auto backward = [&](const __nv_bfloat16* in, std::vector<__nv_bfloat16*>& a, bool to_input, int accumulate) -> int {
    const __nv_bfloat16* dz = dz16; int ldz = Cp; int zi = 0;
    for (int i = nl - 1; i >= 0; --i) {
      const Layer& L = *dl[i];
      const __nv_bfloat16* a_prev = (i == 0) ? in : a[i - 1];
      const int K = (int)L.in_feat, N = L.units;
      if (!to_input) {
        if (dense_tc_wgrad_launch(a_prev, K, dz, ldz, (int)B, K, N, gdata + L.w_off, accumulate, st)) { set_err("dense_tc_wgrad launch failed"); return 1; }
        DBX_LAUNCH(grad_b_bf16_kernel, (int)cdiv(N, 32), kGradBThreads, 0, st, dz, ldz, (int)B, N, gdata + L.b_off, accumulate);
      }
      if (i > 0) {
        if (dense_tc_bwd_launch(dz, 0, ldz, W16 + L.w16_off, m->P16, (int)L.w_cols_pad, 1, (int)B, K, N, a[i - 1], 0, dl[i - 1]->units, dl[i - 1]->act,
                                dbuf[zi], 0, K, nullptr, 0, st)) { set_err("dense_tc_bwd launch failed"); return 1; }
        dz = dbuf[zi]; ldz = K; zi ^= 1;
      } else if (to_input) {      // FGSM step against the step's own weights in the epilogue: the adversarial batch as the next A operand
        if (dense_tc_bwd_launch(dz, 0, ldz, W16 + L.w16_off, m->P16, (int)L.w_cols_pad, 1, (int)B, K, N, nullptr, 0, 0, DBX_ACT_LINEAR, xadv, 0, K,
                                nullptr, 0, st, x_dev, epsilon, in_lower, in_upper)) { set_err("dense_tc_bwd launch failed"); return 1; }
      }
    }
    return 0;
  };
  if (forward(xin, acts, logits)) return 1;
  if (robust_mode == 0) {
    // (with ctx, a null probs_out_dev must not be offset: the scratch buffer holds one mini-batch)
    DBX_LAUNCH(softmax_ce_kernel, (int)cdiv(B, 128), 128, 0, st, (const float*)logits, labels_dev, (int)B, (int)C, probs, dz32, row_loss, ctx,
               probs_out_dev ? 1 : 0);
    DBX_LAUNCH(dz_to_bf16_kernel, grid_for(B * Cp), 256, 0, st, (const float*)dz32, (int)B, (int)C, Cp, dz16);
    if (backward(xin, acts, false, 0)) return 1;
  } else {
    DBX_LAUNCH(attack_upstream_kernel, (int)cdiv(B, 128), 128, 0, st, (const float*)logits, (int)B, (int)C, dz32);
    DBX_LAUNCH(dz_to_bf16_kernel, grid_for(B * Cp), 256, 0, st, (const float*)dz32, (int)B, (int)C, Cp, dz16);
    if (backward(xin, acts, true, 0)) return 1;
    if (forward(xadv, acts_adv, logits_adv)) return 1;
    DBX_LAUNCH(mix_ce_kernel, (int)cdiv(B, 128), 128, 0, st, (const float*)logits, (const float*)logits_adv, labels_dev, (int)B, (int)C,
               robust_lambda, probs, dz32, dz_adv32, row_loss);
    DBX_LAUNCH(dz_to_bf16_kernel, grid_for(B * Cp), 256, 0, st, (const float*)dz32, (int)B, (int)C, Cp, dz16);
    if (backward(xin, acts, false, 0)) return 1;
    DBX_LAUNCH(dz_to_bf16_kernel, grid_for(B * Cp), 256, 0, st, (const float*)dz_adv32, (int)B, (int)C, Cp, dz16);
    if (backward(xadv, acts_adv, false, 1)) return 1;
  }
  DBX_CUDA(cudaGetLastError());
  TrainTensors tt; tt.n = m->table.n;
  for (int i = 0; i < tt.n; ++i) { tt.off[i] = m->table.src_off[i]; tt.end[i] = m->table.src_end[i]; }
  DBX_LAUNCH(bbb_update_kernel, ublocks, 256, 0, st, mean_dev, rho_dev, (const float*)W, (const float*)noise, (const float*)gdata,
             prior_mean_dev, prior_var_dev, P, tt, lrate, kl_weight, partial);
  if (loss_out_dev) DBX_LAUNCH(bbb_loss_kernel, 1, 32, 0, st, (const float*)row_loss, (int)B, (const float*)partial, ublocks, kl_weight, loss_out_dev, ctx);
  return 0;
}

static int run_bbb_step_tc_fwd(dbx_model* m, const float* x_dev, int64_t B, const int32_t* labels_dev, float* mean_dev, float* rho_dev,
                               const float* prior_mean_dev, const float* prior_var_dev, const float* noise_dev, uint64_t seed, int64_t step,
                               float lrate, float kl_weight, float* loss_out_dev, float* probs_out_dev, int robust_mode, float epsilon,
                               float robust_lambda, float in_lower, float in_upper, cudaStream_t st) {
  return run_bbb_step_tc(m, x_dev, B, labels_dev, mean_dev, rho_dev, prior_mean_dev, prior_var_dev, noise_dev, seed, step, lrate, kl_weight,
                         loss_out_dev, probs_out_dev, robust_mode, epsilon, robust_lambda, in_lower, in_upper, st);
}

}  // namespace dbx
