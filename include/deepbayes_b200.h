/*
 * deepbayes_b200 -- C ABI of the B200 (sm_100a) posterior-predictive engine.
 *
 * The reference (matthewwicker/deepbayes) has no FFI layer: its boundary is the Python object
 * protocol the analyzers rely on (deepbayes/analyzers/attacks.py:2-7).  This header declares
 * the entry points a binding for that protocol needs; each cites the reference code it
 * replaces (paths relative to the reference checkout).
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; dbx_last_error() gives text.
 *   - pointers named *_dev are CUDA device pointers on the handle's device, *_host are host
 *     pointers; `stream` is a cudaStream_t passed as void* (NULL = default stream).
 *   - parameters are ONE flat fp32 vector of P elements in Keras tensor order
 *     [W0,b0,W1,b1,...], each tensor C-contiguous (Dense kernel (fan_in,fan_out); Conv2D
 *     kernel HWIO); activations are row-major [rows, features] / NHWC.
 *   - posterior sample ids are global (s0 + local index): results do not depend on how the
 *     samples are split over calls or over GPUs.
 *   - no CPU fallback: every compute entry point launches sm_100a kernels.
 */
#ifndef DEEPBAYES_B200_H
#define DEEPBAYES_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

typedef struct dbx_model* dbx_handle;

enum { DBX_DENSE = 0, DBX_CONV2D = 1, DBX_MAXPOOL2D = 2, DBX_FLATTEN = 3 };
enum { DBX_ACT_LINEAR = 0, DBX_ACT_RELU = 1, DBX_ACT_SOFTMAX = 2, DBX_ACT_TANH = 3, DBX_ACT_SIGMOID = 4 };
enum { DBX_PAD_VALID = 0, DBX_PAD_SAME = 1 };
/* arithmetic mode: FP32 = fp32 FFMA everywhere (parity <=1e-5 rel); BF16 = bf16 GEMM
 * operands on tcgen05 tensor cores, fp32 accumulate / bias / softmax / sample sums. */
enum { DBX_MODE_FP32 = 0, DBX_MODE_BF16 = 1 };
/* what is accumulated over samples: the model output (softmax probabilities when the last
 * layer is softmax; posteriormodel.py:81-101) or the pre-activation logits (:114-139). */
enum { DBX_OUT_PROBS = 0, DBX_OUT_LOGITS = 1 };

/* One Keras layer (the subset the reference's forward touches). */
typedef struct {
  int32_t kind;       /* DBX_DENSE ... */
  int32_t activation; /* DBX_ACT_* */
  int32_t units;      /* Dense: fan_out; Conv2D: filters */
  int32_t kh, kw;     /* Conv2D kernel, MaxPool window */
  int32_t sh, sw;     /* strides */
  int32_t padding;    /* DBX_PAD_* */
} dbx_layer_desc;

const char* dbx_last_error(void);
int dbx_version(void);

/* Build the forward plan for a Keras Sequential (replaces tf.keras.models.load_model at
 * posteriormodel.py:37,50 / the `keras_model` argument of Optimizer.compile,
 * optimizer.py:33-40).  input_shape = per-row feature shape: {K} for Dense-first models,
 * {H,W,C} for Conv2D-first models. */
int dbx_create(const dbx_layer_desc* layers, int32_t n_layers, const int32_t* input_shape,
               int32_t input_rank, int32_t device, dbx_handle* out);
int dbx_destroy(dbx_handle h);
int dbx_param_count(dbx_handle h, int64_t* P);
int dbx_output_dim(dbx_handle h, int64_t* C);
int dbx_input_dim(dbx_handle h, int64_t* K);
/* algorithmic FLOPs of one (posterior sample x input row) forward: 2*sum(fan_in*fan_out). */
int dbx_flops_per_forward(dbx_handle h, int64_t* flops);

/* Gaussian posterior (posteriormodel.py:40-41; optimizer.py:51-52).  `sigma` is the array
 * the reference passes as `scale=` to np.random.normal, i.e. var.npy used AS A STD
 * (posteriormodel.py:74-75).  Copies P floats each; on_device selects the pointer kind. */
int dbx_set_gaussian(dbx_handle h, const float* mu, const float* sigma, int64_t P, int32_t on_device, void* stream);
/* Bayes-by-Backprop parameterisation: sigma = softplus(rho) computed on the device with
 * TensorFlow's thresholds (bayesbybackprop.py:22-23,54; optimizer.py:286). */
int dbx_set_gaussian_rho(dbx_handle h, const float* mu, const float* rho, int64_t P, int32_t on_device, void* stream);
/* Read back the device-resident mu / sigma (e.g. to write var.npy, bayesbybackprop.py:187-191). */
int dbx_get_gaussian(dbx_handle h, float* mu_host, float* sigma_host, int64_t P, void* stream);

/* Stored-sample posterior (hmc.py:308-322, posteriormodel.py:46-54,77-78): slab_dev is an
 * HBM-resident [S,P] fp32 matrix, one stored weight sample per row (borrowed, not copied). */
int dbx_set_samples(dbx_handle h, const float* slab_dev, int64_t S, int64_t P);
/* Same with an explicit row pitch (elements, >= P): a pitch that is a multiple of 4 keeps every stored sample
 * 16-byte aligned, which lets the streaming kernel use 128-bit loads. */
int dbx_set_samples_strided(dbx_handle h, const float* slab_dev, int64_t S, int64_t P, int64_t row_stride);

/* W_s = mu + (inflate*sigma) * eps_s for s in [s0, s0+n)  (posteriormodel.py:60-79,
 * optimizer.py:197-202, bayesbybackprop.py:176-181).  eps_dev: injected noise [n,P] fp32 or
 * NULL -> in-register Philox4x32-10 + Box-Muller keyed by (seed, sample id, element).
 * W_out_dev: [n,P] fp32.  eps_out_dev (optional, [n,P]) receives the noise that was used
 * (`noise_used`, bayesbybackprop.py:57). */
int dbx_sample(dbx_handle h, uint64_t seed, int64_t s0, int64_t n, const float* eps_dev, float inflate,
               float* W_out_dev, float* eps_out_dev, void* stream);

/* Monte-Carlo posterior predictive, Gaussian posterior (posteriormodel.py:81-101,114-139):
 *   sum1[b,c]  = sum_{s in [s0,s0+n)} f(x_b ; W_s)      (fp32, accumulated in sample order)
 *   sum2[b,c]  = sum_s f(...)^2                           (optional, uncertainty.py:33-35)
 * x_dev [B,K] fp32.  The caller divides by float(n) (after the all-reduce when sharded).
 * accumulate != 0 adds into sum1/sum2 instead of overwriting. */
int dbx_predict(dbx_handle h, const float* x_dev, int64_t B, int64_t n, uint64_t seed, int64_t s0,
                const float* eps_dev, float inflate, int32_t mode, int32_t out_kind, int32_t accumulate,
                float* sum1_dev, float* sum2_dev, void* stream);

/* Same loop over the stored-sample posterior: visit rows idx[j] of the slab with weight
 * wts[j] (idx_dev NULL -> rows j0..j0+n-1; wts_dev NULL -> weight 1).  The MC draw
 * with replacement ~ freq (posteriormodel.py:77) is idx = draws, wts = NULL; the deterministic
 * expectation sum_i freq_i f(W_i) (probverification.py:619-650) is idx = NULL, wts = freq. */
int dbx_predict_stored(dbx_handle h, const float* x_dev, int64_t B, int64_t j0, int64_t n,
                       const int32_t* idx_dev, const float* wts_dev, int32_t mode, int32_t out_kind,
                       int32_t accumulate, float* sum1_dev, float* sum2_dev, void* stream);

/* One forward with explicit weights W_dev [P] (PosteriorModel._predict, posteriormodel.py:
 * 112-113; det=True predict :92-93; Optimizer.predict optimizer.py:293-298).  out [B,C]. */
int dbx_forward_one(dbx_handle h, const float* x_dev, int64_t B, const float* W_dev, int32_t mode,
                    int32_t out_kind, float* out_dev, void* stream);

/* Bayes-by-Backprop sampling + forward of one step (bayesbybackprop.py:46-65): draws (or takes)
 * eps, forms W = mu + softplus(rho)*eps, forwards the mini-batch.  W_out_dev / eps_out_dev
 * (optional, [P]) return the sampled weights and `noise_used`. */
int dbx_bbb_forward(dbx_handle h, const float* x_dev, int64_t B, uint64_t seed, int64_t step,
                    const float* eps_dev, int32_t mode, float* out_dev, float* W_out_dev,
                    float* eps_out_dev, void* stream);

/* Statistical-verification outer loop (verifiers.py:537-557, 563-653 with the tutorial
 * predicate argmax == label): for each posterior sample s and each of the K fixed inputs,
 * flag = (argmax_c logits_s(x_k) == labels[k]).  flags_dev [n,K] uint8 (optional);
 * counts_dev [K] int64 (+= number of successes; zeroed first unless accumulate). */
int dbx_count_predicate(dbx_handle h, const float* x_dev, int64_t K, const int32_t* labels_dev, int64_t n,
                        uint64_t seed, int64_t s0, const float* eps_dev, float inflate, int32_t mode,
                        int32_t accumulate, uint8_t* flags_dev, int64_t* counts_dev, void* stream);

/* One Bayes-by-Backprop training step on a mini-batch (deepbayes/optimizers/bayesbybackprop.py:46-170 with
 * robust_train == 0, losses.py:10-45): noise ~ N(0,1) (Philox keyed by (seed, step), or noise_dev [P]),
 * W = mean + softplus(rho) * noise, forward, loss = sparse categorical cross-entropy (batch mean, Keras' clip of p to
 * [1e-7, 1-1e-7]) + kl_weight * (log_q(W; mean, rho) - log_q(W; prior_mean, prior_var)), the three gradients the
 * reference takes from tf.GradientTape, and the in-place update mean -= lrate * (gW + gmean),
 * rho -= lrate * (noise * sigmoid(rho) * gW + grho).  All parameter vectors are flat fp32 [P] in Keras order, on the
 * device, owned by the caller.  Optional outputs: loss_out_dev[2] = (loss, kl component), probs_out_dev [B,C] = the
 * predictions of this step, W_out_dev [P] = the sampled weights (the reference leaves them in the model, :59).
 * robust_mode = 1 (bayesbybackprop.py:73-90): (logit_l, logit_u) = IBP(features, W, eps = epsilon) (verifiers.py:68-95, input box
 * clip(features -+ epsilon, 0, 1)), worst = softmax(logit_l on the label, logit_u elsewhere), the data term is evaluated on
 * robust_lambda * predictions + (1 - robust_lambda) * worst and differentiated through the interval forward as well.
 * robust_mode = 2 (bayesbybackprop.py:91-101): features_adv = FGSM(features, eps = epsilon) against the step's own
 * weights towards the predicted class, clipped to [in_lower, in_upper], and the data term is evaluated on
 * robust_lambda * predictions + (1 - robust_lambda) * model(features_adv); robust_mode = 0 ignores those arguments.
 * Dense MLPs with a softmax output.  mode = DBX_MODE_FP32: fp32 arithmetic throughout (what the reference computes);
 * DBX_MODE_BF16 (robust_mode 0 and 2, Dense widths multiples of 8; the fp32 kernels otherwise): the forward, backward-data and
 * weight-gradient GEMMs on the tensor cores with bf16 operands and fp32 accumulation -- the draw of W, biases, losses, KL terms
 * and the update of mean / rho stay fp32, the parameters never leave fp32. */
int dbx_bbb_step(dbx_handle h, const float* x_dev, int64_t B, const int32_t* labels_dev, float* mean_dev, float* rho_dev,
                 const float* prior_mean_dev, const float* prior_var_dev, const float* noise_dev, uint64_t seed, int64_t step,
                 float lrate, float kl_weight, int32_t robust_mode, float epsilon, float robust_lambda, float in_lower, float in_upper,
                 int32_t mode, float* loss_out_dev, float* probs_out_dev, float* W_out_dev, void* stream);

/* robust_train == 3 of the same step (bayesbybackprop.py:102-121): the data term is evaluated on
 * output = sum_k (1 / n_eps) * softmax(worst-case IBP logits at eps_host[k]) -- the clean prediction is NOT part of the loss
 * (it is still returned in probs_out_dev for the train metric, :166).  eps_host[n_eps] is HOST memory: the reference draws
 * every eps from Exponential(rate = 1 / epsilon) (:106-109); the caller draws them.  Everything else as dbx_bbb_step.
 * (robust_train == 4, :123-136, repeats the SAME FGSM pass loss_mc times -- the drawn eps is not used, :131 -- and is
 * dbx_bbb_step with robust_mode = 2 and robust_lambda = 0.) */
/* One EPOCH of such steps in a single call (optimizer.py:100-133, the mini-batch loop `for (features, labels) in train_ds:
 * self.step(...)`, :112-116): x_dev [N,K] / labels_dev [N] hold the epoch's examples in visiting order (the caller applies
 * its shuffle), mini-batch i = rows [i*batch, min(N, (i+1)*batch)) takes Philox step id step0 + i; lrate / epsilon are the
 * epoch's (the reference changes them per epoch, :102-108, :131-132).  loss_out_dev [ceil(N/batch), 2] and probs_out_dev
 * [N,C] (optional) receive every step's (loss, KL term) and predictions; W_out_dev (optional) the last step's sampled weights
 * (what the reference leaves in the model, :59).  The steps are the same kernels in the same order as N/batch calls of
 * dbx_bbb_step: bit-identical parameters, without a host round trip per mini-batch. */
int dbx_bbb_train_epoch(dbx_handle h, const float* x_dev, const int32_t* labels_dev, int64_t N, int64_t batch, float* mean_dev, float* rho_dev,
                        const float* prior_mean_dev, const float* prior_var_dev, uint64_t seed, int64_t step0, float lrate, float kl_weight,
                        int32_t robust_mode, float epsilon, float robust_lambda, float in_lower, float in_upper, int32_t mode,
                        float* loss_out_dev, float* probs_out_dev, float* W_out_dev, void* stream);
int dbx_bbb_step_ibp_mc(dbx_handle h, const float* x_dev, int64_t B, const int32_t* labels_dev, float* mean_dev, float* rho_dev,
                        const float* prior_mean_dev, const float* prior_var_dev, const float* noise_dev, uint64_t seed, int64_t step,
                        float lrate, float kl_weight, int32_t n_eps, const float* eps_host, float* loss_out_dev, float* probs_out_dev,
                        float* W_out_dev, void* stream);

/* `IBP_state(model, s0, s1, weights, weight_margin)` (verifiers.py:97-123; probverification.py IBP_prob :84-110): the
 * input box [lo, hi] itself is given, the weights are one explicit flat vector, and every weight / bias is widened by
 * weight_margin * sigma (sigma = the Gaussian posterior's scale set with dbx_set_gaussian): the |x_mu| W_r and
 * |x_r| |W_r| terms of propagate_interval (:37-38).  Logit bounds [B,C]; fp32. */
int dbx_ibp_state(dbx_handle h, const float* lo_dev, const float* hi_dev, int64_t B, const float* W_dev, float weight_margin,
                  float* lower_dev, float* upper_dev, void* stream);

/* Expected input gradient (deepbayes/analyzers/attacks.py gradient_expectation :49-75): the sum over n_outer iterations
 * of d CE(labels, mean over n_inner posterior samples of softmax(f_W(x))) / d x, CE = batch-mean sparse categorical
 * cross-entropy.  n_inner = 35 reproduces a PosteriorModel (its predict() averages 35 samples, posteriormodel.py:81),
 * n_inner = 1 an optimizer object.  Sample ids s0 + it * n_inner + s (stored != 0: slab rows j0 + ..., or, when idx_dev is
 * given, the rows idx_dev[it * n_inner + s] -- the draws with replacement of posteriormodel.py:77).  grad_dev [B,K]
 * fp32.  Dense MLPs with a softmax output.  mode = DBX_MODE_FP32: forward and backward on fp32 FFMA kernels (what the
 * reference computes); DBX_MODE_BF16: both passes on the tensor cores (per-sample tcgen05 GEMMs, operands rounded to bf16,
 * fp32 accumulation, softmax / cross-entropy in fp32) when every Dense width is a multiple of 8, the fp32 kernels otherwise. */
int dbx_input_gradient(dbx_handle h, const float* x_dev, int64_t B, const int32_t* labels_dev, int64_t n_outer, int64_t n_inner, uint64_t seed,
                       int64_t s0, const float* noise_dev, float inflate, int32_t stored, int64_t j0, const int32_t* idx_dev, int32_t mode,
                       float* grad_dev, void* stream);
/* same with one explicit flat weight vector (the det / current-weights branch, attacks.py:57-58) */
int dbx_input_gradient_one(dbx_handle h, const float* x_dev, int64_t B, const int32_t* labels_dev, const float* W_dev, float* grad_dev,
                           void* stream);

/* massart_bound_check / chernoff loop with verify=False (verifiers.py:588-634, :622-625): for each of the n posterior
 * samples, an FGSM step (attacks.py:81-102, first order, sparse categorical cross-entropy towards attack_labels) against
 * THAT sample's weights, then the predicate argmax == labels[b] on the sample's output at the adversarial input.
 * counts_dev [B] (+ per-sample flags [n,B]).  lower / upper = the model's input_lower / input_upper.  mode as for
 * dbx_input_gradient (bf16: forward, backward and the second forward at each sample's own adversarial input on tcgen05). */
int dbx_count_predicate_fgsm(dbx_handle h, const float* x_dev, int64_t B, const int32_t* attack_labels_dev, const int32_t* labels_dev,
                             float eps, float lower, float upper, int64_t n, uint64_t seed, int64_t s0, const float* noise_dev, float inflate,
                             int32_t mode, int32_t accumulate, uint8_t* flags_dev, int64_t* counts_dev, void* stream);

/* Interval bound propagation over posterior samples (deepbayes/analyzers/verifiers.py: propagate_interval :23-41,
 * IBP :68-95, and its per-sample use in chernoff_bound_verification :537-557 / massart_bound_check :588-634).
 * For each of the n samples (Gaussian: global ids s0.., optional injected noise [n,P]; stored != 0: slab rows j0..)
 * the input box clip(x -+ eps, 0, 1) (clip01 = 0: x -+ eps) goes through the Dense layers in centre/radius form;
 * outputs, all optional: sum over samples of softmax(worst case) [B,C] (worst case = lower logit of labels[b], upper
 * logit of every other class), of the lower / upper logits [B,C], and per input the number of samples (and the
 * per-sample flags [n,B]) whose worst case still has argmax == labels[b].  Dense / Conv2D (VALID, stride 1) / MaxPool / Flatten models. */
int dbx_ibp_predict(dbx_handle h, const float* x_dev, int64_t B, float eps, int32_t clip01, const int32_t* labels_dev, int64_t n,
                    uint64_t seed, int64_t s0, const float* noise_dev, float inflate, int32_t stored, int64_t j0, int32_t mode,
                    int32_t accumulate, float* sum_worst_dev, float* sum_lower_dev, float* sum_upper_dev, uint8_t* flags_dev,
                    int64_t* counts_dev, void* stream);
/* The same loops with EVERY SAMPLE'S OWN OUTPUT returned instead of sums / flags, for callers that apply their own predicate
 * to each iteration's `worst_case` like massart_bound_check does (verifiers.py:563-653: `predicate(worst_case)`, :628):
 *   dbx_predict_samples  out [n,B,C] = model output (or logits) of posterior sample s0+i at x          (adv given, :625)
 *   dbx_fgsm_samples     probs [n,B,C] = output of sample i at ITS OWN FGSM input (attack label per row)  (:622-625)
 *   dbx_ibp_samples      lower / upper [n,B,C] = logit bounds of sample i over the eps-ball (clip01) or, when box_hi_dev is
 *                        given, over the explicit box [x_dev, box_hi_dev] (IBP(..., predict=True), :605-606)  (:593, :606)
 * also what analyzers/uncertainty.py:15-36 stacks (per-sample predictions). */
int dbx_predict_samples(dbx_handle h, const float* x_dev, int64_t B, int64_t n, uint64_t seed, int64_t s0, const float* eps_dev,
                        float inflate, int32_t mode, int32_t out_kind, float* out_dev, void* stream);
int dbx_fgsm_samples(dbx_handle h, const float* x_dev, int64_t B, const int32_t* attack_labels_dev, float eps, float lower, float upper,
                     int64_t n, uint64_t seed, int64_t s0, const float* noise_dev, float inflate, int32_t mode, float* probs_dev,
                     void* stream);
int dbx_ibp_samples(dbx_handle h, const float* x_dev, const float* box_hi_dev, int64_t B, float eps, int32_t clip01, int64_t n,
                    uint64_t seed, int64_t s0, const float* noise_dev, float inflate, int32_t mode, float* lower_dev, float* upper_dev,
                    void* stream);
/* `IBP(model, inp, weights, eps)` with one explicit flat weight vector (verifiers.py:68-95): logit bounds [B,C]. */
int dbx_ibp_one(dbx_handle h, const float* x_dev, int64_t B, float eps, int32_t clip01, const float* W_dev, int32_t mode,
                float* lower_dev, float* upper_dev, void* stream);

/* Host -> device upload of `count` fp32 values on `stream` (the input of predict: posteriormodel.py:81 takes any host array).
 * Page-locked sources are copied with one DMA (the caller keeps them alive until the stream has passed the copy); pageable
 * sources go through two pinned staging pieces owned by the handle, each reused only after the event of its previous copy,
 * so the call returns with `src_host` free to be modified. */
int dbx_upload_f32(dbx_handle h, const float* src_host, float* dst_dev, int64_t count, void* stream);

/* Host-buffer form of dbx_predict, the call PosteriorModel.predict(x, n) binds to: copies x
 * (host, [B,K] fp32) to the device, runs the loop, writes mean = sum1/n ([B,C]) and, if
 * var_host != NULL, the per-class variance sum2/n - mean^2 back to host memory. */
int dbx_predict_host(dbx_handle h, const float* x_host, int64_t B, int64_t n, uint64_t seed, int64_t s0,
                     float inflate, int32_t mode, int32_t out_kind, float* mean_host, float* var_host);

/* ---- multi-GPU: the one exchange of the path -------------------------------------------------------
 * Posterior samples shard over ranks (posteriormodel.py:94-101: the iterations of the loop are independent); the only
 * collective is the SUM over ranks of the [B,C] sample-sums.  A communicator owns a window in its GPU's HBM that every
 * peer rank of the node maps through CUDA IPC; dbx_comm_allreduce is ONE kernel on `stream`: local sums -> own window,
 * release-flag to every peer over NVLink, wait for all peers' flags, read all windows in rank order and write the sum in
 * place (bit-identical on every rank, independent of arrival order).  One process per GPU:
 *   dbx_comm_create   allocate the window (max_floats = largest na + nb of a later call)
 *   dbx_comm_handle   64-byte IPC handle of the own window; the host side all-gathers these (torch.distributed / MPI)
 *   dbx_comm_connect  map the `world` windows (handles in rank order, world x 64 bytes)
 *   dbx_comm_allreduce  a_dev[na] (and b_dev[nb], optional) += the same buffers of every other rank; calls on one
 *                     communicator must be issued in the same order on every rank; na, nb multiples of 4, 16-byte aligned
 *   dbx_comm_status   timed_out != 0 after a peer failed to arrive within 10 s (the communicator is then unusable) */
typedef struct dbx_comm_s* dbx_comm;
int dbx_comm_create(int32_t rank, int32_t world, int32_t device, int64_t max_floats, dbx_comm* out);
int dbx_comm_handle(dbx_comm c, void* handle64);
int dbx_comm_connect(dbx_comm c, const void* handles);
int dbx_comm_allreduce(dbx_comm c, float* a_dev, int64_t na, float* b_dev, int64_t nb, void* stream);
int dbx_comm_status(dbx_comm c, int32_t* timed_out, int64_t* calls);
int dbx_comm_destroy(dbx_comm c);

/* Elementwise activation on a device buffer [rows, cols] (model.layers[-1].activation(...),
 * verifiers.py:552-555). */
int dbx_activation(int32_t act, float* x_dev, int64_t rows, int64_t cols, void* stream);

/* Raw Philox4x32-10 words for sample id s, elements [0,P) (bit-exactness test hook). */
int dbx_philox_raw(uint64_t seed, int64_t s, int64_t P, uint32_t* out_dev, void* stream);

/* Number of kernel launches issued by this library since process start (bench `gpu_launches`). */
int64_t dbx_launch_count(void);
/* Time the dominant (GEMM-carrying) kernel of every call with CUDA events on the launching
 * stream while enabled; dbx_profile_read synchronises, returns the summed device time (ms),
 * the algorithmic FLOPs those launches carried and their count, and clears the record. */
int dbx_profile(int enable);
int dbx_profile_read(double* ms_total, double* flops_total, int64_t* launches);
/* which kernels served the last compute call: 0 generic (CUDA cores), 1 fused tcgen05 MLP,
 * 2 small-batch streaming, 3 generic with per-layer tcgen05 GEMMs (bf16 mode). */
int dbx_last_path(void);
/* Process-wide A/B switches of the kernel dispatch (DESIGN.md section 5 lists them; none selects a CPU or library
 * fallback).  Their defaults are read ONCE from the DBX_<NAME> environment variables; afterwards only these calls change
 * them (the parity tests compare one kernel against another this way).  `name` without the DBX_ prefix, e.g. "NO_FUSED".
 * dbx_reset_options re-reads the environment defaults. */
int dbx_set_option(const char* name, int64_t value);
int dbx_get_option(const char* name, int64_t* value);
int dbx_reset_options(void);
/* Development hook (tools/trace_fused.py): with option FUSED_TRACE=1 the fused kernel records (probe id, clock) pairs of one
 * cluster; this copies up to max_pairs of them to host memory and returns the count. */
int dbx_debug_trace(long long* out_host, int max_pairs);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif

#ifdef __cplusplus
}
#endif
#endif /* DEEPBAYES_B200_H */
