// Internal model / plan structures shared by engine.cu (generic kernels + C ABI) and
// fused_mlp.cu (tcgen05 fast path).  Not part of the public ABI.
#pragma once
#include <vector>
#include "common.cuh"
#include "../../include/deepbayes_b200.h"

namespace dbx {

struct Layer {
  int kind = 0, act = 0, units = 0, kh = 0, kw = 0, sh = 1, sw = 1, pad = 0;
  int in_h = 0, in_w = 0, in_c = 0, out_h = 0, out_w = 0, out_c = 0, pad_t = 0, pad_l = 0;
  int64_t in_feat = 0, out_feat = 0;
  bool has_w = false;
  // GEMM view of the kernel tensor: [w_rows (=fan_in or kh*kw*cin), w_cols (=units)] row-major
  int64_t w_rows = 0, w_cols = 0;
  int64_t w_off = 0, b_off = 0;      // offsets in the flat fp32 Keras-order vector
  int64_t w16_off = 0, w_cols_pad = 0;  // offsets / row pitch in the padded bf16 weight layout
  int64_t b32_off = 0;               // offset in the fp32 bias side-array used by bf16 mode
};

// One entry per Keras tensor, used by the sampling / conversion kernels to scatter the flat
// Keras-order element index into the bf16 device layout.
#define DBX_MAX_TENSORS 16
struct TensorTable {
  int n;
  int32_t src_off[DBX_MAX_TENSORS];   // start in flat Keras order
  int32_t src_end[DBX_MAX_TENSORS];   // src_off + size
  int32_t size[DBX_MAX_TENSORS];
  int32_t dst_off[DBX_MAX_TENSORS];   // start in bf16 weight layout (kernels) or bias array (biases)
  int32_t cols[DBX_MAX_TENSORS];
  int32_t cols_pad[DBX_MAX_TENSORS];
  int32_t rows[DBX_MAX_TENSORS];
  int8_t is_bias[DBX_MAX_TENSORS];
  // blocked8: element (k, n) of a kernel tensor goes to dst_off + ((n>>3)*rows + k)*8 + (n&7): the
  // tcgen05 no-swizzle core-matrix order, so the fused kernel fetches it with ONE bulk copy
  int8_t blocked8[DBX_MAX_TENSORS];
};

// Philox-block ranges [lo, hi) handled by the general sampling kernel
#define DBX_MAX_RANGES 12
struct BlockRanges {
  int n;
  int32_t lo[DBX_MAX_RANGES], hi[DBX_MAX_RANGES];
};

struct Workspace {
  void* ptr = nullptr;
  size_t bytes = 0;
};

}  // namespace dbx

struct dbx_model {
  int device = 0;
  std::vector<dbx::Layer> layers;
  int last_weighted = -1;
  int64_t P = 0, in_feat = 0, out_feat = 0, flops = 0;
  int64_t P16 = 0;   // elements of the padded bf16 weight layout (kernels only)
  int64_t PB = 0;    // elements of the fp32 bias side-array
  int64_t max_feat = 0;
  bool is_mlp = false;  // only Dense layers (flatten tolerated)
  dbx::TensorTable table;
  float* mu = nullptr;
  float* sigma = nullptr;
  bool have_gaussian = false;
  cudaEvent_t ev_posterior = nullptr;   // recorded after the last upload of mu / sigma (side streams wait on it)
  const float* slab = nullptr;
  int64_t S = 0;
  int64_t slab_stride = 0;   // elements between stored samples (>= P; a multiple of 4 enables 128-bit loads)
  dbx::Workspace ws;        // generic-path workspace
  dbx::Workspace fused_ws;  // fused-path workspace
  // host-call staging (dbx_predict_host)
  float* h_pin = nullptr; size_t h_pin_bytes = 0;
  float* d_x = nullptr; size_t d_x_bytes = 0;
  float* d_s1 = nullptr; float* d_s2 = nullptr; size_t d_s_bytes = 0;
  cudaStream_t own_stream = nullptr;
  // pageable host -> device uploads: two pinned staging pieces, each guarded by the event of its last copy (dbx_upload_f32)
  void* up_stage[2] = {nullptr, nullptr}; cudaEvent_t up_ev[2] = {nullptr, nullptr}; int up_next = 0;
  void* fused_state = nullptr;  // owned by fused_mlp.cu
  cudaStream_t ibp_side = nullptr;  // interval forward: weights of chunk c+1 are drawn here while chunk c is in the GEMMs
  cudaStream_t ibp_main = nullptr;  // ... and the GEMM chain here (high priority)
  cudaEvent_t ibp_ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // entry, sampled[2], consumed[2], done
};

namespace dbx {

// Where the per-sample weights come from.
struct Source {
  bool stored = false;
  // gaussian
  const float* eps = nullptr; float inflate = 1.f; uint64_t seed = 0; int64_t s0 = 0; bool unfused = false;
  // stored
  int64_t j0 = 0; const int32_t* idx = nullptr;
};

// What happens to each sample's output.
struct Sink {
  int kind = 0;  // 0 accumulate moments, 1 count predicate, 2 per-sample outputs [n,B,C] into `sum1` (generic executor only)
  int out_kind = 0; int accumulate = 0;
  float* sum1 = nullptr; float* sum2 = nullptr; const float* wts = nullptr;
  const int32_t* labels = nullptr; uint8_t* flags = nullptr; unsigned long long* counts = nullptr;
};

int ensure_ws(Workspace& w, size_t bytes);

// Optional timing of the dominant (GEMM-carrying) kernel with CUDA events on the launching
// stream -- read by bench.py for the roofline line (dbx_profile / dbx_profile_read).
void prof_begin(cudaStream_t st);
void prof_end(cudaStream_t st, double flops);

// Materialise the chunk's per-sample weights in the bf16 device layout (kernels -> W16 [nc,P16]
// padded bf16, biases -> B32 [nc,PB] fp32): Gaussian draw (injected eps or Philox) or conversion
// of stored samples.  Defined in engine.cu.
int launch_sample_bf16(dbx_model* m, const TensorTable& tab, const Source& src, int64_t c0, int nc, __nv_bfloat16* W16,
                       float* B32, cudaStream_t st, int64_t pb_stride = 0, const bool* skip = nullptr, int64_t w16_stride = 0);

// fused_mlp.cu: returns 1 if the tcgen05 fused path handled the request, 0 if the shape is
// not supported by it (caller uses the generic kernels), <0 on error.
int fused_mlp_run(dbx_model* m, const float* x_dev, int64_t B, int64_t n, const Source& src, const Sink& sink,
                  cudaStream_t st);
void fused_mlp_free(dbx_model* m);

// gemm_tc.cu: one layer of every sample of a chunk on tcgen05 (generic bf16 path); return 0 on success
int dense_tc_launch(const __nv_bfloat16* A, long long a_sstride, int lda, int a_shared, const __nv_bfloat16* Wbase,
                    long long w_sstride_elems, int ldw, int ns, int M, int N, int K, const float* bias, long long b_sstride,
                    long long b_off, int act, __nv_bfloat16* out16, long long o16_sstride, int ld16, float* out32,
                    long long o32_sstride, cudaStream_t st, int nt_cap = 256);      // nt_cap 128: narrower column tiles, two CTAs per SM
int dense_tc_bwd_launch(const __nv_bfloat16* dZ, long long dz_sstride, int ldz, const __nv_bfloat16* Wbase, long long w_sstride_elems,
                        int ldw, int ns, int M, int fan_in, int fan_out, const __nv_bfloat16* aprev, long long ap_sstride, int ld_ap,
                        int act, __nv_bfloat16* out16, long long o16_sstride, int ld16, float* out32, long long o32_sstride,
                        cudaStream_t st, const float* fgsm_x = nullptr, float fgsm_eps = 0.f, float fgsm_lo = 0.f, float fgsm_hi = 0.f);
int dense_tc_wgrad_launch(const __nv_bfloat16* A, int lda, const __nv_bfloat16* dZ, int ldz, int B, int fan_in, int fan_out, float* out,
                          int accumulate, cudaStream_t st);
int dense_tc_f32w_launch(const __nv_bfloat16* A, long long a_sstride, int lda, int a_rep, const float* slab, long long slab_stride,
                         long long slab_rows, long long w_off, const int32_t* w_rows, long long w_row0, int ns, int M, int N, int K,
                         const float* bias, long long b_sstride, long long b_off, int act, __nv_bfloat16* out16,
                         long long o16_sstride, int ld16, float* out32, long long o32_sstride, cudaStream_t st);
int dense_tc_shareda_launch(const __nv_bfloat16* A, int lda, const __nv_bfloat16* Wbase, long long w_sstride_elems, int ldw, int ns, int M,
                            int N, int K, const float* bias, long long b_sstride, long long b_off, int act, __nv_bfloat16* out16,
                            long long o16_sstride, int ld16, cudaStream_t st);
bool conv_tc_eligible(const Layer& L);
bool conv_tc_pool_eligible(const Layer& L, const Layer& P);
int conv_tc_launch(const __nv_bfloat16* in, long long in_sstride, int in_shared, const __nv_bfloat16* W16, long long w_sstride_elems,
                   int ldw, int ns, int NB, const Layer& L, const float* bias, long long b_sstride, long long b_off, int act,
                   __nv_bfloat16* out, long long o_sstride, cudaStream_t st, int pool = 0);
bool ibp_tc_eligible(int N, int K);
int ibp_tc_launch(const __nv_bfloat16* Amu, const __nv_bfloat16* Ar, long long a_sstride, int a_shared, const __nv_bfloat16* Wbase,
                  long long w_sstride_elems, int ldw, int ns, int M, int N, int K, const float* bias, long long b_sstride, long long b_off,
                  int act, __nv_bfloat16* n_mu, __nv_bfloat16* n_r, long long n_sstride, cudaStream_t st);
int im2col_launch(const __nv_bfloat16* in, long long in_sstride, __nv_bfloat16* out, long long out_sstride, int ns, int NB,
                  const Layer& L, int Kp, cudaStream_t st);

// stream_mlp.cu: small-batch (B <= 16) streaming path, fp32 FFMA, weights read / generated once.
int stream_mlp_run(dbx_model* m, const float* x_dev, int64_t B, int64_t n, const Source& src, const Sink& sink,
                   cudaStream_t st);

}  // namespace dbx
