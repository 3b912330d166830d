// Interval bound propagation over posterior samples (included by engine.cu).
//
// Reference: deepbayes/analyzers/verifiers.py -- propagate_interval :23-41 (centre/radius form:
// h_mu = x_mu W, x_rad = x_r |W|, h_u = h_mu + x_rad + b, h_l = h_mu - x_rad + b), IBP :68-95 (input box
// clip(x -+ eps, 0, 1), activation on both bounds, none after the last layer), and the per-sample use in
// chernoff_bound_verification :537-557 / massart_bound_check :588-634 (worst case = lower bound of the true
// class, upper bound of every other class; softmax of that; sum / predicate count over samples).
// The reference evaluates the layer in float64; here fp32 mode accumulates in fp32 (ascending k) and bf16 mode
// runs both products on the tensor cores (operands rounded to bf16, fp32 accumulation).
//
// Per sample and Dense layer:  MU = A_mu W + b,  R = A_r |W|   (A_r >= 0, so R >= 0)
//   hidden layer: u = act(MU + R), l = act(MU - R)  (act monotone), next A_mu = (u + l)/2, A_r = (u - l)/2
//   last layer  : lower = MU - R, upper = MU + R, worst[c] = c == label ? lower[c] : upper[c]
#pragma once

namespace dbx {

template <typename T>
__global__ void ibp_input_kernel(const float* __restrict__ x, float eps, int clip01, T* __restrict__ a_mu,
                                 T* __restrict__ a_r, int64_t total) {
  for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
    float u = x[o] + eps, l = x[o] - eps;                    // verifiers.py:70-74, in the input's dtype (fp32)
    if (clip01) { u = fminf(fmaxf(u, 0.f), 1.f); l = fminf(fmaxf(l, 0.f), 1.f); }
    a_mu[o] = from_f32<T>((u + l) / 2.f);                    // :26-27
    a_r[o] = from_f32<T>((u - l) / 2.f);
  }
}

// explicit input box [lo, hi] (IBP_state, verifiers.py:97-99)
template <typename T>
__global__ void ibp_box_kernel(const float* __restrict__ lo, const float* __restrict__ hi, T* __restrict__ a_mu, T* __restrict__ a_r,
                               int64_t total) {
  for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
    a_mu[o] = from_f32<T>((hi[o] + lo[o]) / 2.f);
    a_r[o] = from_f32<T>((hi[o] - lo[o]) / 2.f);
  }
}

// |W16|: the radius product needs the element-wise absolute value of the sampled kernels (sign bits cleared)
__global__ void abs_bf16_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int64_t n16) {
  for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < n16; o += (int64_t)gridDim.x * blockDim.x) {
    uint4 v = in[o];
    v.x &= 0x7FFF7FFFu; v.y &= 0x7FFF7FFFu; v.z &= 0x7FFF7FFFu; v.w &= 0x7FFF7FFFu;
    out[o] = v;
  }
}

// epilogue shared by both GEMM back ends
struct IbpOut {
  int last;                       // 1: write lower / upper / worst (fp32 [s][M][N]); 0: write next centre / radius
  int act;
  const int32_t* labels;          // [M] (last layer)
  float* lower; float* upper; float* worst; int64_t o_sstride;
};

template <typename T>
__device__ __forceinline__ void ibp_store(const IbpOut& o, T* __restrict__ n_mu, T* __restrict__ n_r, int64_t n_sstride,
                                          int s, int gm, int gn, int N, float mu, float r) {
  const float hi = mu + r, lo = mu - r;
  if (o.last) {
    const int64_t at = (int64_t)s * o.o_sstride + (int64_t)gm * N + gn;
    if (o.lower) o.lower[at] = lo;
    if (o.upper) o.upper[at] = hi;
    o.worst[at] = (o.labels && o.labels[gm] == gn) ? lo : hi;
  } else {
    const float au = apply_act(o.act, hi), al = apply_act(o.act, lo);
    const int64_t at = (int64_t)s * n_sstride + (int64_t)gm * N + gn;
    n_mu[at] = from_f32<T>((au + al) / 2.f);
    n_r[at] = from_f32<T>((au - al) / 2.f);
  }
}

// CUDA-core back end: ONE pass over W_s gives both products (64x64x16 tiles, 4x4 outputs per thread, fp32
// accumulation in ascending-k order), epilogue fused.
template <typename T>
__global__ void __launch_bounds__(256)
ibp_dense_kernel(const T* __restrict__ Amu, const T* __restrict__ Ar, int64_t a_sstride, const T* __restrict__ Wbase,
                 int64_t w_sstride, const int32_t* __restrict__ rows, int64_t row0, int64_t w_off, int ldw,
                 const float* __restrict__ Bbase, int64_t b_sstride, int64_t b_off, T* __restrict__ n_mu,
                 T* __restrict__ n_r, int64_t n_sstride, int M, int N, int K, IbpOut o,
                 const float* __restrict__ sigma, float margin, int64_t sig_w_off, int64_t sig_b_off) {
  // margin > 0 (IBP_state, verifiers.py:107-112): weight interval W -+ margin * sigma, bias interval b -+ margin * sigma_b:
  //   radius += |x_mu| W_r + x_r W_r   (propagate_interval :37-38),  upper += b_marg, lower -= b_marg
  constexpr int BM = 64, BN = 64, BK = 16;
  __shared__ float As[BK][BM + 4];
  __shared__ float Rs[BK][BM + 4];
  __shared__ float Ws[BK][BN + 4];
  __shared__ float Ss[BK][BN + 4];
  const bool marg = margin > 0.f && sigma != nullptr;
  const int s = blockIdx.z;
  const int64_t wrow = rows ? (int64_t)rows[s] : (row0 + s);
  const T* Ap = Amu + (int64_t)s * a_sstride;
  const T* Rp = Ar + (int64_t)s * a_sstride;
  const T* Wp = Wbase + wrow * w_sstride + w_off;
  const float* bp = Bbase + wrow * b_sstride + b_off;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int t = threadIdx.x, ty = t >> 4, tx = t & 15;
  float am[4][4] = {}, ar[4][4] = {};
  const int arow = t >> 2, ak = (t & 3) << 2;
  const int wk = t >> 4, wn = (t & 15) << 2;
  for (int k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int gm = m0 + arow, gk = k0 + ak + i;
      const bool ok = gm < M && gk < K;
      As[ak + i][arow] = ok ? to_f32<T>(Ap[(int64_t)gm * K + gk]) : 0.f;
      Rs[ak + i][arow] = ok ? to_f32<T>(Rp[(int64_t)gm * K + gk]) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int gk = k0 + wk, gn = n0 + wn + i;
      Ws[wk][wn + i] = (gk < K && gn < N) ? to_f32<T>(Wp[(int64_t)gk * ldw + gn]) : 0.f;
      if (marg) Ss[wk][wn + i] = (gk < K && gn < N) ? margin * sigma[sig_w_off + (int64_t)gk * N + gn] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[4], r[4], w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[k][ty * 4 + i]; r[i] = Rs[k][ty * 4 + i]; }
#pragma unroll
      for (int j = 0; j < 4; ++j) w[j] = Ws[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { am[i][j] = fmaf(a[i], w[j], am[i][j]); ar[i][j] = fmaf(r[i], fabsf(w[j]), ar[i][j]); }
      if (marg) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) ar[i][j] = fmaf(fabsf(a[i]) + r[i], Ss[k][tx * 4 + j], ar[i][j]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn < N) ibp_store<T>(o, n_mu, n_r, n_sstride, s, gm, gn, N, am[i][j] + bp[gn], ar[i][j] + (marg ? margin * sigma[sig_b_off + gn] : 0.f));
    }
  }
}

// propagate_conv2d (verifiers.py:11-21).  Two quirks of the reference are kept: tf.nn.convolution is called with its defaults
// (VALID padding, stride 1 -- the layer's own settings are not passed), and the nominal convolution is ADDED to both interval
// sums: h_l = conv(mid, W) + [conv(x_l, W+) + conv(x_u, W-)] + b = 2 conv(mid, W) - conv(rad, |W|) + b, h_u likewise with +.
// One thread per output element (NHWC x HWIO), centre and radius sums side by side.
template <typename T>
__global__ void ibp_conv_kernel(const T* __restrict__ Amu, const T* __restrict__ Ar, int64_t a_sstride, const T* __restrict__ Wbase,
                                int64_t w_sstride, const int32_t* __restrict__ rows, int64_t row0, int64_t w_off, int ldw,
                                const float* __restrict__ Bbase, int64_t b_sstride, int64_t b_off, T* __restrict__ n_mu,
                                T* __restrict__ n_r, int64_t n_sstride, int NB, int H, int Wd, int Cin, int Ho, int Wo, int Cout,
                                int kh, int kw, IbpOut o) {
  const int s = blockIdx.y;
  const int64_t wrow = rows ? (int64_t)rows[s] : (row0 + s);
  const T* Mp = Amu + (int64_t)s * a_sstride;
  const T* Rp = Ar + (int64_t)s * a_sstride;
  const T* Wp = Wbase + wrow * w_sstride + w_off;
  const float* bp = Bbase + wrow * b_sstride + b_off;
  const int64_t total = (int64_t)NB * Ho * Wo * Cout;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int co = (int)(e % Cout);
    int64_t r = e / Cout;
    const int wo = (int)(r % Wo); r /= Wo;
    const int ho = (int)(r % Ho);
    const int nb = (int)(r / Ho);
    float am = 0.f, ar = 0.f;
    for (int i = 0; i < kh; ++i)
      for (int j = 0; j < kw; ++j) {
        const int64_t at = (((int64_t)nb * H + ho + i) * Wd + wo + j) * Cin;
        const T* wp = Wp + ((int64_t)(i * kw + j) * Cin) * ldw + co;
        for (int c = 0; c < Cin; ++c) {
          const float w = to_f32<T>(wp[(int64_t)c * ldw]);
          am = fmaf(to_f32<T>(Mp[at + c]), w, am);
          ar = fmaf(to_f32<T>(Rp[at + c]), fabsf(w), ar);
        }
      }
    ibp_store<T>(o, n_mu, n_r, n_sstride, s, (int)(e / Cout), co, Cout, 2.f * am + bp[co], ar);
  }
}

// a weightless pooling layer is applied to each bound (verifiers.py:78-81): max-pool the upper and the lower bound
template <typename T>
__global__ void ibp_pool_kernel(const T* __restrict__ Amu, const T* __restrict__ Ar, int64_t a_sstride, T* __restrict__ n_mu,
                                T* __restrict__ n_r, int64_t n_sstride, int NB, int H, int Wd, int C, int Ho, int Wo, int ph, int pw,
                                int sh, int sw) {
  const int s = blockIdx.y;
  const T* Mp = Amu + (int64_t)s * a_sstride;
  const T* Rp = Ar + (int64_t)s * a_sstride;
  const int64_t total = (int64_t)NB * Ho * Wo * C;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(e % C);
    int64_t r = e / C;
    const int wo = (int)(r % Wo); r /= Wo;
    const int ho = (int)(r % Ho);
    const int nb = (int)(r / Ho);
    float hi = -INFINITY, lo = -INFINITY;
    for (int i = 0; i < ph; ++i)
      for (int j = 0; j < pw; ++j) {
        const int64_t at = (((int64_t)nb * H + ho * sh + i) * Wd + wo * sw + j) * C + c;
        const float m = to_f32<T>(Mp[at]), rr = to_f32<T>(Rp[at]);
        hi = fmaxf(hi, m + rr); lo = fmaxf(lo, m - rr);
      }
    n_mu[(int64_t)s * n_sstride + e] = from_f32<T>((hi + lo) / 2.f);
    n_r[(int64_t)s * n_sstride + e] = from_f32<T>((hi - lo) / 2.f);
  }
}

// |W16| of one parameter range [off, off + len) of every sample (units of 8 bf16)
__global__ void abs_bf16_range_kernel(const uint4* __restrict__ W, uint4* __restrict__ Wabs, int64_t sstride8, int64_t off8, int64_t len8) {
  const int64_t base = (int64_t)blockIdx.y * sstride8 + off8;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < len8; e += (int64_t)gridDim.x * blockDim.x) {
    uint4 v = W[base + e];
    v.x &= 0x7FFF7FFFu; v.y &= 0x7FFF7FFFu; v.z &= 0x7FFF7FFFu; v.w &= 0x7FFF7FFFu;
    Wabs[base + e] = v;
  }
}

// tensor-core back end: MU (bias added) and R come from two dense_tc_kernel launches as fp32 [s][M][N]
__global__ void ibp_combine_kernel(const float* __restrict__ MU, const float* __restrict__ R, int64_t total_per_s, int M, int N,
                                   __nv_bfloat16* __restrict__ n_mu, __nv_bfloat16* __restrict__ n_r, int64_t n_sstride, IbpOut o) {
  const int s = blockIdx.y;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total_per_s; e += (int64_t)gridDim.x * blockDim.x) {
    const int gm = (int)(e / N), gn = (int)(e - (int64_t)gm * N);
    const int64_t at = (int64_t)s * total_per_s + e;
    ibp_store<__nv_bfloat16>(o, n_mu, n_r, n_sstride, s, gm, gn, N, MU[at], R[at]);
  }
}

struct IbpSink {
  const int32_t* labels = nullptr;
  float* sum_worst = nullptr; float* sum_lower = nullptr; float* sum_upper = nullptr;   // [B,C] each, optional
  unsigned long long* counts = nullptr; uint8_t* flags = nullptr;                        // optional
  float* per_lower = nullptr; float* per_upper = nullptr;                                // [n,B,C] each: every sample's own logit bounds, optional
  int accumulate = 0;
};

template <typename T>
static int run_ibp(dbx_model* m, const float* x_dev, int64_t B, float eps, int clip01, int64_t n, const Source& src,
                   const IbpSink& sink, cudaStream_t st_user, const float* box_hi_dev = nullptr, float weight_margin = 0.f) {
  cudaStream_t st = st_user;
  // box_hi_dev != nullptr: x_dev / box_hi_dev are the lower / upper input bounds themselves (IBP_state); weight_margin > 0
  // widens every weight by margin * sigma (needs the Gaussian posterior's sigma; CUDA-core back end)
  constexpr bool kBf16 = !std::is_same<T, float>::value;
  for (size_t li = 0; li < m->layers.size(); ++li) {
    const Layer& L = m->layers[li];
    if (L.kind != DBX_CONV2D) continue;
    // tf.nn.convolution(x, w) in propagate_conv2d runs VALID / stride 1 whatever the layer says: any other layer would leave the
    // reference with shapes that do not match its own next layer
    DBX_CHECK(L.pad == DBX_PAD_VALID && L.sh == 1 && L.sw == 1, "IBP through Conv2D needs padding='valid' and strides 1 (propagate_conv2d, verifiers.py:11-21, calls tf.nn.convolution with its defaults)");
    DBX_CHECK((int)li != m->last_weighted, "IBP: the output layer must be Dense");
    DBX_CHECK(!(weight_margin > 0.f), "weight margins through Conv2D are not built");
  }
  for (size_t li = 0; li + 1 < m->layers.size(); ++li) {
    const Layer& L = m->layers[li];
    DBX_CHECK(!L.has_w || (int)li == m->last_weighted || L.act != DBX_ACT_SOFTMAX,
              "IBP needs monotone element-wise hidden activations (layer %d is softmax)", (int)li);
  }
  const int64_t C = m->out_feat;
  bool use_tc = kBf16 && !(weight_margin > 0.f);
  if (opt().no_tc) use_tc = false;
  DBX_CHECK(!(weight_margin > 0.f) || (m->have_gaussian && !kBf16), "weight margins need a Gaussian posterior's sigma and the fp32 mode");
  // ---- workspace per sample: 4 activation buffers (centre/radius, ping-pong), fp32 MU/R (tensor-core back end),
  //      lower/upper/worst, weights (+ |W16|)
  const size_t act_b = round_up((size_t)B * m->max_feat * sizeof(T), 256);
  const size_t f32_b = use_tc ? round_up((size_t)B * m->max_feat * 4, 256) : 0;
  const size_t out_b = (size_t)B * C * 4;                     // lower / upper / worst stay dense [s][B][C] for the sink kernels
  const size_t w16_b = (size_t)m->P16 * 2;
  DBX_CHECK(!kBf16 || m->P16 % 8 == 0, "bf16 parameter pitch %lld is not a multiple of 8", (long long)m->P16);
  // bf16: the weights of chunk c + 1 are drawn on a side stream while chunk c is in the GEMMs (two W16 / bias buffers;
  // the sampler is ALU work, the GEMMs tensor-pipe work); DBX_IBP_OVERLAP=0 serialises them
  const bool overlap = kBf16 && opt().ibp_overlap != 0;
  const int nwb = overlap ? 2 : 1;
  const size_t pb_b = round_up((size_t)m->PB * 4, 16);
  const size_t w_b = kBf16 ? (w16_b * (nwb + (use_tc ? 1 : 0)) + nwb * pb_b) : (src.stored ? 0 : (size_t)m->P * 4);
  const size_t per = 4 * act_b + 2 * f32_b + 3 * out_b + w_b + 64;
  const size_t zb_b = round_up(std::max<size_t>(4096, (size_t)m->max_feat * 4), 256);
  const size_t in_b = 2 * round_up((size_t)B * m->in_feat * sizeof(T), 256) + zb_b + 4096;
  int64_t ns = (int64_t)std::max<size_t>(1, (ws_budget_bytes() - std::min(ws_budget_bytes(), in_b)) / per);
  ns = std::min<int64_t>(std::min<int64_t>(ns, n), 256);
  if (ensure_ws(m->ws, in_b + (size_t)ns * per + 16384)) return 1;
  char* p = (char*)m->ws.ptr;
  T* in_mu = (T*)p; p += round_up((size_t)B * m->in_feat * sizeof(T), 256);
  T* in_r = (T*)p; p += round_up((size_t)B * m->in_feat * sizeof(T), 256);
  float* zero_bias = (float*)p; p += zb_b;
  T* abuf[4];
  for (int i = 0; i < 4; ++i) { abuf[i] = (T*)p; p += (size_t)ns * act_b; }
  float* MU = (float*)p; p += (size_t)ns * f32_b;
  float* R = (float*)p; p += (size_t)ns * f32_b;
  float* lower = (float*)p; p += round_up((size_t)ns * out_b, 256);
  float* upper = (float*)p; p += round_up((size_t)ns * out_b, 256);
  float* worst = (float*)p; p += round_up((size_t)ns * out_b, 256);
  T* Wc = (T*)p;
  T* Wc2[2] = {Wc, Wc}; float* B32b[2] = {nullptr, nullptr}; __nv_bfloat16* Wabs = nullptr;
  if (kBf16) {
    const size_t wslab = round_up((size_t)ns * w16_b, 256);
    char* q = p;
    for (int i = 0; i < nwb; ++i) { Wc2[i] = (T*)q; q += wslab; }
    if (nwb == 1) Wc2[1] = Wc2[0];
    Wabs = (__nv_bfloat16*)q; if (use_tc) q += wslab;
    for (int i = 0; i < nwb; ++i) { B32b[i] = (float*)q; q += round_up((size_t)ns * pb_b, 256); }
    if (nwb == 1) B32b[1] = B32b[0];
  }
  cudaStream_t sst = st;
  if (overlap) {
    if (!m->ibp_side) {
      DBX_CUDA(cudaStreamCreateWithFlags(&m->ibp_side, cudaStreamNonBlocking));
      for (cudaEvent_t& e : m->ibp_ev) DBX_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    sst = m->ibp_side;
    // everything enqueued on the caller's stream before this call (previous users of the workspace, the posterior upload, a
    // slab or injected noise written just before) is ordered before the first draw
    DBX_CUDA(cudaEventRecord(m->ibp_ev[0], st_user));
    DBX_CUDA(cudaStreamWaitEvent(sst, m->ibp_ev[0], 0));
    // the GEMM chain runs on a HIGH-priority stream of our own (joined with the caller's stream at both ends): its short
    // persistent kernels then get SM slots ahead of the sampler's queued blocks instead of waiting for the sampler grid to
    // drain (DBX_IBP_PRIO=0: run it on the caller's stream)
    const bool prio = opt().ibp_prio != 0;
    if (prio) {
      if (!m->ibp_main) {
        int lo = 0, hi = 0;
        DBX_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        DBX_CUDA(cudaStreamCreateWithPriority(&m->ibp_main, cudaStreamNonBlocking, hi));
      }
      st = m->ibp_main;
      DBX_CUDA(cudaStreamWaitEvent(st, m->ibp_ev[0], 0));
    }
  }
  auto draw_chunk = [&](int64_t ci) -> int {     // bf16 only
    const int b = (int)(ci & 1) % nwb;
    const int64_t c0 = ci * ns;
    const int nc = (int)std::min<int64_t>(ns, n - c0);
    if (overlap && ci >= 2) DBX_CUDA(cudaStreamWaitEvent(sst, m->ibp_ev[3 + b], 0));
    if (launch_sample_bf16(m, m->table, src, c0, nc, (__nv_bfloat16*)Wc2[b], B32b[b], sst)) return 1;
    if (overlap) DBX_CUDA(cudaEventRecord(m->ibp_ev[1 + b], sst));
    return 0;
  };
  const int64_t act_ss = (int64_t)(act_b / sizeof(T)), out_ss = B * C;

  if (box_hi_dev) DBX_LAUNCH(ibp_box_kernel<T>, grid_for(B * m->in_feat), 256, 0, st, x_dev, box_hi_dev, in_mu, in_r, B * m->in_feat);
  else DBX_LAUNCH(ibp_input_kernel<T>, grid_for(B * m->in_feat), 256, 0, st, x_dev, eps, clip01, in_mu, in_r, B * m->in_feat);
  DBX_CUDA(cudaMemsetAsync(zero_bias, 0, zb_b, st));
  if (sink.counts && !sink.accumulate) DBX_LAUNCH(zero_u64_kernel, grid_for(B), 256, 0, st, sink.counts, B);

  // hidden layers that the fused interval kernel takes need no |W16| copy
  int64_t abs_off = 0, abs_len = kBf16 ? m->P16 : 0;
  std::vector<char> fused_layer(m->layers.size(), 0);
  if (use_tc) {
    bool all_hidden = true;
    for (int li = 0; li < (int)m->layers.size(); ++li) {
      const Layer& L = m->layers[li];
      if (!L.has_w || li == m->last_weighted || L.kind != DBX_DENSE) continue;
      fused_layer[li] = ibp_tc_eligible(L.units, (int)L.in_feat) && (act_b / sizeof(T)) % 8 == 0;
      all_hidden = all_hidden && fused_layer[li];
    }
    const Layer& LL = m->layers[m->last_weighted];
    if (all_hidden && LL.w16_off % 8 == 0 && (LL.w_rows * LL.w_cols_pad) % 8 == 0) { abs_off = LL.w16_off; abs_len = LL.w_rows * LL.w_cols_pad; }
  }
  for (int64_t c0 = 0; c0 < n; c0 += ns) {
    const int nc = (int)std::min<int64_t>(ns, n - c0);
    const T* Wbase; int64_t w_sstride; const int32_t* rows = nullptr; int64_t row0 = 0;
    const float* Bbase; int64_t b_sstride;
    const int64_t ci = c0 / ns;
    const int wb = (int)(ci & 1) % nwb;
    if (kBf16) {
      if (ci == 0 || !overlap) { if (draw_chunk(ci)) return 1; }
      if (overlap) {
        if (c0 + ns < n && draw_chunk(ci + 1)) return 1;       // next chunk's weights, concurrently
        DBX_CUDA(cudaStreamWaitEvent(st, m->ibp_ev[1 + wb], 0));
      }
      Wc = Wc2[wb];
      Wbase = Wc; w_sstride = m->P16; Bbase = B32b[wb]; b_sstride = m->PB;
      if (use_tc) {  // |W16| for the layers that run as two GEMMs: all of them, or (hidden layers fused) the output layer alone
        if (abs_len == m->P16)
          DBX_LAUNCH(abs_bf16_kernel, grid_for((int64_t)nc * m->P16 / 8), 256, 0, st, (const uint4*)Wc, (uint4*)Wabs, (int64_t)nc * m->P16 / 8);
        else if (abs_len > 0)
          DBX_LAUNCH(abs_bf16_range_kernel, dim3((unsigned)std::min<int64_t>(cdiv(abs_len / 8, 256), 1024), (unsigned)nc), 256, 0, st,
                     (const uint4*)Wc, (uint4*)Wabs, m->P16 / 8, abs_off / 8, abs_len / 8);
      }
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
    const T* cur_mu = in_mu; const T* cur_r = in_r; int64_t cur_ss = 0;
    int bi = 0;
    for (int li = 0; li < (int)m->layers.size(); ++li) {
      const Layer& L = m->layers[li];
      if (L.kind == DBX_MAXPOOL2D) {
        T* p_mu = abuf[2 * bi]; T* p_r = abuf[2 * bi + 1];
        dim3 grid((unsigned)grid_for(B * L.out_feat), (unsigned)nc);
        DBX_LAUNCH(ibp_pool_kernel<T>, grid, 256, 0, st, cur_mu, cur_r, cur_ss, p_mu, p_r, act_ss, (int)B, L.in_h, L.in_w, L.in_c, L.out_h, L.out_w,
                   L.kh, L.kw, L.sh, L.sw);
        cur_mu = p_mu; cur_r = p_r; cur_ss = act_ss; bi ^= 1;
        continue;
      }
      if (!L.has_w) continue;    // Flatten (NHWC order is the flat order already)
      const bool last = (li == m->last_weighted);
      const int64_t woff = kBf16 ? L.w16_off : L.w_off;
      const int64_t boff = kBf16 ? L.b32_off : L.b_off;
      const int ldw = (int)(kBf16 ? L.w_cols_pad : L.w_cols);
      IbpOut o; o.last = last ? 1 : 0; o.act = L.act; o.labels = sink.labels;
      o.lower = lower; o.upper = upper; o.worst = worst; o.o_sstride = out_ss;
      T* n_mu = abuf[2 * bi]; T* n_r = abuf[2 * bi + 1];
      prof_begin(st);
      bool done = false;
      if (L.kind == DBX_CONV2D) {
        dim3 grid((unsigned)grid_for(B * L.out_feat), (unsigned)nc);
        DBX_LAUNCH(ibp_conv_kernel<T>, grid, 256, 0, st, cur_mu, cur_r, cur_ss, Wbase, w_sstride, rows, row0, woff, ldw, Bbase, b_sstride, boff,
                   n_mu, n_r, act_ss, (int)B, L.in_h, L.in_w, L.in_c, L.out_h, L.out_w, L.out_c, L.kh, L.kw, o);
        done = true;
      } else if (kBf16 && use_tc && !last && fused_layer[li]) {
        // hidden layer: centre + radius + combine in one kernel, |W| formed in the SM (no fallback: |W16| was not written for it)
        DBX_CHECK(ibp_tc_launch((const __nv_bfloat16*)cur_mu, (const __nv_bfloat16*)cur_r, cur_ss, cur_ss == 0 ? 1 : 0,
                                (const __nv_bfloat16*)Wbase + woff, w_sstride, ldw, nc, (int)B, L.units, (int)L.in_feat, Bbase, b_sstride, boff,
                                L.act, (__nv_bfloat16*)n_mu, (__nv_bfloat16*)n_r, act_ss, st) == 0,
                  "fused interval layer %d failed to launch", li);
        done = true; g_used_tc = true;
      } else if (kBf16 && use_tc && L.in_feat % 8 == 0 && (last || L.units % 8 == 0) && (size_t)L.units * 4 <= zb_b) {
        const int shared = cur_ss == 0 ? 1 : 0;
        const int rc1 = dense_tc_launch((const __nv_bfloat16*)cur_mu, cur_ss, (int)L.in_feat, shared, (const __nv_bfloat16*)Wbase + woff,
                                        w_sstride, ldw, nc, (int)B, L.units, (int)L.in_feat, Bbase, b_sstride, boff, DBX_ACT_LINEAR,
                                        nullptr, 0, 0, MU, B * L.units, st);
        const int rc2 = rc1 ? 1 : dense_tc_launch((const __nv_bfloat16*)cur_r, cur_ss, (int)L.in_feat, shared, Wabs + woff, w_sstride, ldw, nc,
                                                  (int)B, L.units, (int)L.in_feat, zero_bias, 0, 0, DBX_ACT_LINEAR, nullptr, 0, 0, R,
                                                  B * L.units, st);
        if (rc1 == 0 && rc2 == 0) {
          dim3 grid((unsigned)grid_for(B * L.units), (unsigned)nc);
          DBX_LAUNCH(ibp_combine_kernel, grid, 256, 0, st, MU, R, B * L.units, (int)B, L.units, (__nv_bfloat16*)n_mu,
                     (__nv_bfloat16*)n_r, act_ss, o);
          done = true; g_used_tc = true;
        }
      }
      if (!done) {
        dim3 grid((unsigned)cdiv(L.units, 64), (unsigned)cdiv(B, 64), (unsigned)nc);
        DBX_LAUNCH(ibp_dense_kernel<T>, grid, 256, 0, st, cur_mu, cur_r, cur_ss, Wbase, w_sstride, rows, row0, woff, ldw, Bbase,
                   b_sstride, boff, n_mu, n_r, act_ss, (int)B, L.units, (int)L.in_feat, o, (const float*)m->sigma, weight_margin, L.w_off, L.b_off);
      }
      prof_end(st, 4.0 * (double)L.w_rows * L.w_cols * (double)B * nc);
      if (last) break;
      cur_mu = n_mu; cur_r = n_r; cur_ss = act_ss; bi ^= 1;
    }
    if (overlap) DBX_CUDA(cudaEventRecord(m->ibp_ev[3 + wb], st));     // this chunk's weights are free again
    // ---- sinks
    const int act_last = m->layers[m->last_weighted].act;
    const int overwrite = (c0 == 0 && !sink.accumulate) ? 1 : 0;
    if (sink.sum_worst)
      launch_accum(worst, nc, B, (int)C, act_last, DBX_OUT_PROBS, nullptr, overwrite, sink.sum_worst, nullptr, st);
    if (sink.sum_lower)
      launch_accum(lower, nc, B, (int)C, act_last, DBX_OUT_LOGITS, nullptr, overwrite, sink.sum_lower, nullptr, st);
    if (sink.sum_upper)
      launch_accum(upper, nc, B, (int)C, act_last, DBX_OUT_LOGITS, nullptr, overwrite, sink.sum_upper, nullptr, st);
    if (sink.per_lower) DBX_CUDA(cudaMemcpyAsync(sink.per_lower + c0 * out_ss, lower, (size_t)nc * out_ss * 4, cudaMemcpyDeviceToDevice, st));
    if (sink.per_upper) DBX_CUDA(cudaMemcpyAsync(sink.per_upper + c0 * out_ss, upper, (size_t)nc * out_ss * 4, cudaMemcpyDeviceToDevice, st));
    if (sink.counts)
      DBX_LAUNCH(count_kernel, grid_for((int64_t)nc * B), 256, 0, st, worst, nc, B, (int)C, sink.labels,
                 sink.flags ? sink.flags + c0 * B : nullptr, sink.counts);
  }
  if (st != st_user) {       // hand the results back to the caller's stream
    DBX_CUDA(cudaEventRecord(m->ibp_ev[5], st));
    DBX_CUDA(cudaStreamWaitEvent(st_user, m->ibp_ev[5], 0));
  }
  return 0;
}

}  // namespace dbx
