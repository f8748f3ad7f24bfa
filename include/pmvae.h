/* pmvae.h -- C ABI of libpmvae.so: the B200 (sm_100a) PM-VAE hot path.
 *
 * The reference (lupalab/posterior-matching) is pure Python/JAX and exposes a Haiku
 * module class, not an FFI.  These entry points are what a
 * jax.ffi / XLA custom-call (or ctypes) binding for that path would bind; each one
 * cites the reference code it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error; the message is
 *     available from pmvae_last_error() (thread-local);
 *   - all data pointers are DEVICE pointers owned by the caller unless the name says
 *     `host`; outputs are pre-allocated by the caller; nothing is retained;
 *   - every device function only ENQUEUES work on `stream` (no sync, no allocation);
 *     scratch comes from the caller (`ws`, sized by pmvae_workspace_bytes);
 *   - all tensors are float32 row-major; masks are float32 0/1 (masking.py:11,17);
 *   - keys are the two uint32 words of a JAX threefry PRNGKey, passed by value from
 *     the host.
 */
#ifndef PMVAE_H_
#define PMVAE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* pmvae_stream_t; /* cudaStream_t */

/* What PosteriorMatchingVAE.from_config (posterior_matching/models/vae.py:61-118)
 * resolves a `config.model` mapping to, for the ResidualMLP/TriLGaussian/
 * IdentityGaussian family of configs/pm_vae_{gas,power,hepmass,bsds}.py. */
typedef struct pmvae_config {
  int32_t D;        /* features (decoder_dist_config.event_size)            */
  int32_t d;        /* latent_dim                                            */
  int32_t H;        /* hidden_units (256 in every config)                    */
  int32_t R_enc;    /* encoder_net_config.residual_blocks                    */
  int32_t R_dec;    /* decoder_net_config.residual_blocks                    */
  int32_t R_part;   /* partial_encoder_net_config (defaults to the encoder's)*/
  int32_t ln_enc;   /* layer_norm flags (networks.py:117-118,123-124,128-129)*/
  int32_t ln_dec;
  int32_t ln_part;
  int32_t stop_grad;/* matching_ll_stop_gradients (vae.py:136-137)           */
  int32_t precision;/* PMVAE_PREC_*: arithmetic of the dense contractions    */
  int32_t reserved[5];
} pmvae_config;

enum { PMVAE_PREC_F32 = 0,  /* fp32 FMA tiles (exact-parity path)                    */
       PMVAE_PREC_BF16 = 1  /* bf16 operands, fp32 accumulate, tcgen05/TMEM tiles    */ };

/* One Haiku parameter leaf pair (SURVEY §3.3 order): w[rows, cols] and b[cols] inside
 * the flat float32 parameter arena.  The scalar decoder_dist/log_scale is reported as
 * rows = cols = 0 with w_off = its offset. */
typedef struct pmvae_leaf {
  char name[64];
  int32_t rows, cols;
  uint64_t w_off, b_off; /* offsets in floats */
} pmvae_leaf;

const char* pmvae_last_error(void);
int pmvae_version(void);
/* Number of kernels this library has launched in this process (bench.py's gpu_launches). */
uint64_t pmvae_launch_count(void);

/* ---- parameter arena ------------------------------------------------------------ */
/* Number of floats in the (padded) flat arena; gradients and both Adam moments use
 * the same layout. */
uint64_t pmvae_param_count(const pmvae_config* cfg);
/* Fills up to `cap` leaves, returns how many the config has (or <0 on error). */
int pmvae_layout(const pmvae_config* cfg, pmvae_leaf* out, int cap);

/* ---- PRNG (jax.random threefry2x32 stream; reference draws via hk.next_rng_key():
 *      vae.py:124,162,192-195) ------------------------------- */
/* host: jax.random.split / fold_in / haiku PRNGSequence.next (key <- row 0, sub = row 1) */
int pmvae_key_split_host(const uint32_t key[2], int n, uint32_t* out_keys /* [n,2] */);
int pmvae_key_fold_in_host(const uint32_t key[2], uint32_t data, uint32_t out_key[2]);
/* device: elements [start, start+count) of the n_total-element draw */
int pmvae_random_bits(const uint32_t key[2], uint64_t n_total, uint64_t start, uint64_t count,
                      uint32_t* out, pmvae_stream_t stream);
int pmvae_uniform(const uint32_t key[2], uint64_t n_total, uint64_t start, uint64_t count,
                  float* out, pmvae_stream_t stream);
int pmvae_normal(const uint32_t key[2], uint64_t n_total, uint64_t start, uint64_t count,
                 float* out, pmvae_stream_t stream);

/* ---- masks (masking.py:84-91 BernoulliMaskGenerator; :235-249 MNISTMaskGenerator),
 *      redefined on the JAX stream: b = uniform(key,[B_total,D]) < p -------------- */
int pmvae_mask_bernoulli(const uint32_t key[2], float p, uint64_t B_total, uint64_t row_start,
                         uint64_t rows, int32_t D, float* out, pmvae_stream_t stream);
int pmvae_mask_mnist(const uint32_t key[2], uint64_t B_total, uint64_t row_start, uint64_t rows,
                     float* out /* [rows,28,28,1] */, pmvae_stream_t stream);

/* ---- one hk.Linear (networks.py:116,122,127; distributions.py:44,104) ------------------
 * y[B,N] = relu?(x[B,K]) @ w[K,N] + bias[N] in the given PMVAE_PREC_* arithmetic: the
 * dominant kernel of the path, exposed for parity tests and bench.py's roofline leg.
 * `ws` is scratch for the bf16 operand images (may be NULL for PMVAE_PREC_F32). */
int pmvae_linear(int32_t precision, const float* x, const float* w, const float* bias, int64_t B,
                 int32_t K, int32_t N, int32_t relu_in, float* y, void* ws, uint64_t ws_bytes,
                 pmvae_stream_t stream);

/* The two tcgen05 kernels behind PMVAE_PREC_BF16, on raw bf16 operands (uint16 storage):
 *   nt: y[M,N] (fp32) = A[M,K] . Bt[N,K]^T + bias[N]      (forward / input-gradient shape)
 *   tn: y[M,N] (fp32) += A[rows,M]^T . B[rows,N]          (weight-gradient shape; y must be
 *       zeroed by the caller, the contraction is split over CTAs and summed with atomics)
 * Row pitches lda/ldb are in elements and must be multiples of 8; N % 8 == 0. */
int pmvae_tc_gemm_nt(const void* A, int64_t lda, const void* Bt, int64_t ldb, const float* bias,
                     int64_t M, int32_t N, int32_t K, float* y, pmvae_stream_t stream);
int pmvae_tc_gemm_tn(const void* A, int64_t lda, const void* B, int64_t ldb, int32_t M, int32_t N,
                     int64_t rows, float* y, pmvae_stream_t stream);

/* ---- model ------------------------------------------------------------------------ */
/* Bytes of scratch for a batch of B rows (training forward+backward keeps its saved
 * activations here) and, for the evaluators, K importance samples per row. */
uint64_t pmvae_workspace_bytes(const pmvae_config* cfg, int64_t B, int64_t K);

/* Must be called after the float32 parameters change and before the next
 * forward/eval when precision == PMVAE_PREC_BF16: refreshes the bf16 operand images
 * kept at the head of `ws` (no-op for PMVAE_PREC_F32). */
int pmvae_prepare_params(const pmvae_config* cfg, const float* params, void* ws, uint64_t ws_bytes,
                         pmvae_stream_t stream);

/* PosteriorMatchingVAE.__call__ (vae.py:120-144).  eps [B,d] is the N(0,I) draw behind
 * posterior.sample (z = mu + L eps).  Writes the three per-row terms and leaves what
 * pmvae_backward needs in `ws`. */
int pmvae_forward(const pmvae_config* cfg, const float* params, const float* x, const float* b,
                  const float* eps, int64_t B, float* out_rec, float* out_kl, float* out_match,
                  void* ws, uint64_t ws_bytes, pmvae_stream_t stream);

/* Vector-Jacobian product of pmvae_forward: cotangents g_* [B] for the three outputs ->
 * gradient arena (overwritten).  (Replaces jax.value_and_grad through the module in
 * bax.Trainer, train_pm_vae.py:85,96.) */
int pmvae_backward(const pmvae_config* cfg, const float* params, const float* x, const float* b,
                   const float* eps, int64_t B, const float* g_rec, const float* g_kl,
                   const float* g_match, float* grads, void* ws, uint64_t ws_bytes,
                   pmvae_stream_t stream);

/* loss_fn (train_pm_vae.py:58-72): cotangents of loss = -mean(rec - beta*kl) - coef*mean(match)
 * over B_global rows, and the batch sums of the three terms (out_sums[3], accumulated
 * with atomics after being zeroed here). */
int pmvae_loss_cotangents(int64_t B, int64_t B_global, float beta, float coef, const float* rec,
                          const float* kl, const float* match, float* g_rec, float* g_kl,
                          float* g_match, float* out_sums, pmvae_stream_t stream);

/* optax chain of train_pm_vae.py:74-83: Adam (b1,b2,eps) -> + wd*p on leaves with
 * ndim != 1 -> * lr -> * -1.  `count` = updates already applied. */
int pmvae_adamw(const pmvae_config* cfg, float* params, const float* grads, float* m, float* v,
                int64_t count, float lr, float wd, float b1, float b2, float eps,
                pmvae_stream_t stream);

/* One network + its distribution head on its own: the module attributes .encoder, .decoder and
 * .partial_encoder (vae.py:47-53; used by lookahead.py:77,126,133,219 and the MNIST notebook).
 *   which = 0: encoder(x[B,D])             -> out[B,P]  raw TriL parameters (loc | FillScaleTriL input)
 *   which = 1: decoder(z[B,d])             -> out[B,D]  IdentityGaussian loc
 *   which = 2: partial_encoder([x*b, b])   -> out[B,P]  (in = x, msk = b, both [B,D])
 * P = d + d(d+1)/2. */
enum { PMVAE_NET_ENCODER = 0, PMVAE_NET_DECODER = 1, PMVAE_NET_PARTIAL_ENCODER = 2,
       /* OR-ed into `which`: run the training-mode forward of that net (what pmvae_forward launches for it: the
        * saved activations / relu bits of pmvae_backward are written into `ws` as well) -- bench.py's roofline leg */
       PMVAE_NET_SAVE = 0x100 };
int pmvae_net_apply(const pmvae_config* cfg, const float* params, int32_t which, const float* in,
                    const float* msk, int64_t B, float* out, void* ws, uint64_t ws_bytes,
                    pmvae_stream_t stream);

/* PosteriorMatchingVAE.is_log_prob (vae.py:171-226) with K samples per row; eps drawn
 * on device from key_z / key_zxo as normal(key, [K, B_total, d]) restricted to rows
 * [row_start, row_start+B).  Either output may be NULL. */
int pmvae_is_log_prob(const pmvae_config* cfg, const float* params, const float* x, const float* b,
                      int64_t B, int64_t K, const uint32_t key_z[2], const uint32_t key_zxo[2],
                      int64_t B_total, int64_t row_start, float* out_log_p_x,
                      float* out_log_p_xu_given_xo, void* ws, uint64_t ws_bytes,
                      pmvae_stream_t stream);

/* mean over K of PosteriorMatchingVAE.impute (vae.py:146-169; the mean is what
 * eval_pm_vae_uci.py:88-89 keeps). */
int pmvae_impute_mean(const pmvae_config* cfg, const float* params, const float* x, const float* b,
                      int64_t B, int64_t K, const uint32_t key[2], int64_t B_total,
                      int64_t row_start, float* out /* [B,D] */, void* ws, uint64_t ws_bytes,
                      pmvae_stream_t stream);

/* PosteriorMatchingVAE.impute (vae.py:146-169): K imputations per row, out_samples[K, B, D] =
 * where(b == 1, x_o, decoder mean of z_k), z_k ~ q(z | x_o) drawn as in pmvae_impute_mean (same key, same stream);
 * out_mean[B, D] (optional) is their mean over K.  Either output may be NULL, not both. */
int pmvae_impute(const pmvae_config* cfg, const float* params, const float* x, const float* b,
                 int64_t B, int64_t K, const uint32_t key[2], int64_t B_total, int64_t row_start,
                 float* out_samples /* [K,B,D] */, float* out_mean /* [B,D] or NULL */, void* ws,
                 uint64_t ws_bytes, pmvae_stream_t stream);

/* ---- distribution objects behind .encoder / .partial_encoder / .decoder / .prior (vae.py:47-57) ---------------
 * What lookahead.py:126-133,219-222 and the evaluation scripts call on the returned tfd objects, as row kernels over
 * the raw head output `par[B, d + d(d+1)/2]` of pmvae_net_apply (TriLGaussian, distributions.py:101-113):
 *   pmvae_tril_log_prob   MultivariateNormalTriL.log_prob(z)            out[B]
 *   pmvae_tril_entropy    MultivariateNormalTriL.entropy()              out[B]
 *   pmvae_tril_sample     .sample(seed=key, sample_shape=K): z[K,B,d] = mu + L eps, eps = rows
 *                         [row_start, row_start+B) of normal(key, [K, B_total, d]); log_ratio[K,B] =
 *                         log N(z; 0, I) - log q(z) (the importance-weight term of vae.py:203-212)
 * (one sample with caller-provided eps + the KL to the prior: pmvae_tril_sample_kl below)
 *   pmvae_normal_log_prob IdentityGaussian (distributions.py:41-55): Normal(loc, exp(log_scale)).log_prob(x)
 *                         elementwise, out[rows, D]; x has rows_x rows and is repeated rows / rows_x times
 *   pmvae_std_normal_log_prob  the prior MultivariateNormalDiag(0, 1).log_prob(z) (vae.py:55-57), out[B] */
int pmvae_tril_log_prob(const float* par, const float* z, int64_t B, int32_t d, float* out, pmvae_stream_t stream);
int pmvae_tril_entropy(const float* par, int64_t B, int32_t d, float* out, pmvae_stream_t stream);
/* VJP of pmvae_tril_log_prob for the cotangent g[B]: dpar[B, d + d(d+1)/2] and (optionally) dz[B, d] -- what a
 * TriLGaussian partial posterior contributes to the backward pass of vae.py:136-138 when it sits on host-composed
 * networks (configs/pm_vae_mnist16.py).  ws: (B P + B d + B) floats of scratch. */
int pmvae_tril_log_prob_backward(const float* par, const float* z, const float* g, int64_t B, int32_t d, float* dpar,
                                 float* dz, void* ws, uint64_t ws_bytes, pmvae_stream_t stream);
int pmvae_tril_sample(const float* par, const uint32_t key[2], int64_t B, int64_t K, int64_t B_total,
                      int64_t row_start, int32_t d, float* z, float* log_ratio, pmvae_stream_t stream);
int pmvae_normal_log_prob(const float* x, const float* loc, const float* log_scale /* device scalar */,
                          int64_t rows, int64_t rows_x, int32_t D, float* out, pmvae_stream_t stream);
int pmvae_std_normal_log_prob(const float* z, int64_t B, int32_t d, float* out, pmvae_stream_t stream);
/* DiagonalGaussian (distributions.py:58-84), the posterior of the VaDE models (vade.py:61-63): par[B, 2d] = raw head
 * output [loc | raw scale], scale = softplus(raw) + 1e-5.
 *   pmvae_diag_sample    z = loc + scale * eps            (posterior.sample, vade.py:259, for the PM-VaDE matching term)
 *   pmvae_diag_log_prob  MultivariateNormalDiag.log_prob(z) and / or .entropy(), each [B] (either output may be NULL) */
int pmvae_diag_sample(const float* par, const float* eps, int64_t B, int32_t d, float* z, pmvae_stream_t stream);
int pmvae_diag_log_prob(const float* par, const float* z, int64_t B, int32_t d, float* out_log_prob,
                        float* out_entropy, pmvae_stream_t stream);
/* LookaheadPosterior objective (lookahead.py:183-199): par[B,S,2d] = the selected LookaheadBlock
 * outputs [loc | raw scale] (lookahead.py:14-39), z[K,B,S,d] one-step latent samples, valid[B,S] in {0,1}:
 *   ll[b] = sum_s valid[b,s] mean_k MultivariateNormalDiag(loc, softplus(raw) + 1e-5).log_prob(z[k,b,s]) / #valid[b]
 * (0 where no s is valid), and its VJP dpar[B,S,2d] for the cotangent g[B]. */
int pmvae_lookahead_ll(const float* par, const float* z, const float* valid, int64_t K, int64_t B, int32_t S, int32_t d,
                       float* out_ll, pmvae_stream_t stream);
int pmvae_lookahead_ll_backward(const float* par, const float* z, const float* valid, const float* g, int64_t K,
                                int64_t B, int32_t S, int32_t d, float* dpar, pmvae_stream_t stream);

/* ---- the whole training step as one launch sequence ---------------------------------------------
 * train_pm_vae.py's step (mask draw, eps draw, loss_fn forward, value_and_grad, optax update) enqueued by ONE
 * call with nothing step-dependent passed from the host: per-step PRNG keys, beta (train_pm_vae.py:28-43,
 * utils.py:124-136), the learning rate and Adam's bias corrections (train_pm_vae.py:74-83) are derived on the
 * device from a small state block, so the sequence can be captured once in a CUDA graph and replayed.
 * Same arithmetic and key chain as pmvae_mask_bernoulli / pmvae_normal / pmvae_forward / pmvae_loss_cotangents /
 * pmvae_backward / pmvae_adamw driven from the host (Trainer.train_step). */
enum { PMVAE_BETA_CONST = 0, PMVAE_BETA_CYCLIC = 1, PMVAE_BETA_MONOTONIC = 2 };
typedef struct pmvae_train_config {
  int32_t beta_schedule;                       /* PMVAE_BETA_*                                         */
  float beta_low, beta_high;                   /* low_value / high_value                               */
  int64_t beta_period, beta_delay;             /* cyclic                                               */
  int64_t beta_transition_steps, beta_transition_begin; /* monotonic                                  */
  float matching_coef;                         /* config.get("matching_coef", 1.0)                     */
  float lr_init, lr_decay_rate;                /* optax.exponential_decay                              */
  int64_t lr_transition_steps;
  float weight_decay, adam_b1, adam_b2, adam_eps;
  float mask_p;                                /* BernoulliMaskGenerator p                             */
  int32_t reserved[3];
} pmvae_train_config;
typedef struct pmvae_train_state_host {        /* host copy of the device state, for checkpoints/tests */
  uint32_t seq_key[2], mask_key[2], eps_key[2];
  uint32_t mask_calls, pad;
  int64_t step;
  float beta, lr, bc1, bc2;
} pmvae_train_state_host;
uint64_t pmvae_train_state_bytes(void);
uint64_t pmvae_train_scratch_floats(const pmvae_config* cfg, int64_t B);
/* seq_key: key of the Haiku PRNGSequence handing out the per-step rng; mask_key: the mask generator's key
 * (call c draws with fold_in(mask_key, c)); step: optimizer updates already applied.  Synchronises `stream`. */
int pmvae_train_state_init(void* state, const uint32_t seq_key[2], const uint32_t mask_key[2], int64_t step,
                           uint32_t mask_calls, pmvae_stream_t stream);
int pmvae_train_state_read(const void* state, pmvae_train_state_host* out, pmvae_stream_t stream);
/* phase bit 0: advance state, draw mask + eps, forward, loss cotangents (+ batch sums into out_sums[3]), backward
 * (grads overwritten);  phase bit 1: AdamW, refresh of the operand images, step counter.  With several ranks call
 * phase 1, all-reduce `grads` and `out_sums`, then phase 2.  `scratch`: pmvae_train_scratch_floats floats.
 * Bucketed exchange: PMVAE_STEP_FWD_BWD | PMVAE_STEP_SPLIT_BWD stops the backward after the decoder + latent stage
 * (the decoder's gradient range [decoder_net/linear .. decoder_dist] is final), PMVAE_STEP_BWD_ENC then
 * PMVAE_STEP_BWD_PART finish the encoder's and the partial encoder's ranges; each range can be all-reduced on a side
 * stream while the next stage runs (Trainer.train_step_fused does, inside one captured CUDA graph). */
enum { PMVAE_STEP_FWD_BWD = 1, PMVAE_STEP_UPDATE = 2, PMVAE_STEP_BWD_ENC = 4, PMVAE_STEP_BWD_PART = 8,
       PMVAE_STEP_SPLIT_BWD = 16 };
int pmvae_train_step(const pmvae_config* cfg, const pmvae_train_config* tc, float* params, float* m, float* v,
                     float* grads, void* state, const float* x, int64_t B, int64_t B_global, int64_t row_start,
                     float* scratch, float* out_sums, void* ws, uint64_t ws_bytes, int32_t phase,
                     pmvae_stream_t stream);

/* ---- distribution heads of configs/pm_vae_mnist.py (float32 arithmetic) -------------------------
 * Bernoulli decoder (distributions.py:20-25; summed over the event at vae.py:127-128):
 *   out[r] = sum_j w[r,j] * Bernoulli(logits[r,j]).log_prob(x[r,j]),  x a float in [0,1], w optional (NULL = 1)
 * and its VJP dlogits[r,j] = g[r] * w[r,j] * (x - sigmoid(logits)). */
int pmvae_bernoulli_ll(const float* logits, const float* x, const float* w, int64_t B, int32_t D,
                       float* out, pmvae_stream_t stream);
int pmvae_bernoulli_ll_backward(const float* logits, const float* x, const float* w, const float* g,
                                int64_t B, int32_t D, float* dlogits, pmvae_stream_t stream);

/* AutoregressiveGMM (distributions.py:192-223) = ResidualMLP(R, H) + OneDimensionalGMM(d, n_comp)
 * (:116-134) over [z * (arange(d) < i), (arange(d) < i), context]; log_prob follows
 * _AutoregressiveDistribution.log_prob (:152-166) with the d steps batched as d*B rows.
 * Parameters: one flat float32 arena, leaves `partial_posterior_dist/residual_mlp/linear{,_1..}` and
 * `partial_posterior_dist/one_dimensional_gmm/linear` (pmvae_argmm_layout). */
typedef struct pmvae_argmm_config {
  int32_t d;       /* event_size (latent_dim)            */
  int32_t n_comp;  /* num_components                      */
  int32_t R;       /* residual_blocks                     */
  int32_t H;       /* hidden_units                        */
  int32_t C;       /* flattened context features          */
  int32_t reserved[3]; /* [0] = 1: hidden 256 x 256 Linears on the tcgen05 GEMMs with bf16 operands and fp32
                        * accumulation (first Linear, mixture head and density algebra stay float32); else 0 */
} pmvae_argmm_config;
uint64_t pmvae_argmm_param_count(const pmvae_argmm_config* cfg);
int pmvae_argmm_layout(const pmvae_argmm_config* cfg, pmvae_leaf* out, int cap);
uint64_t pmvae_argmm_workspace_bytes(const pmvae_argmm_config* cfg, int64_t B);
/* out[b] = log q(z[b] | context[b]) */
int pmvae_argmm_log_prob(const pmvae_argmm_config* cfg, const float* params, const float* z,
                         const float* context, int64_t B, float* out, void* ws, uint64_t ws_bytes,
                         pmvae_stream_t stream);
/* VJP of the preceding pmvae_argmm_log_prob on the same `ws`: g[B] -> grads (arena, overwritten),
 * dz[B,d] and dcontext[B,C] (either may be NULL). */
int pmvae_argmm_backward(const pmvae_argmm_config* cfg, const float* params, const float* z,
                         const float* context, int64_t B, const float* g, float* grads, float* dz,
                         float* dcontext, void* ws, uint64_t ws_bytes, pmvae_stream_t stream);

/* _AutoregressiveDistribution._sample_n (distributions.py:168-189): out[n, B, d], n samples per context row, drawn
 * dimension by dimension (one ResidualMLP pass over the n * B rows per dimension).  Noise contract [R: TFP's seed
 * plumbing is recollection, so parity with the reference is distributional]: (k_comp, k_cat) = split(key);
 * eps = normal(k_comp, [n, d, K]), u = uniform(k_cat, [n, d, K]); at step i component c = argmax_k(logits_k -
 * log(-log(u[s, i, k]))) and x_i = mean_c + scale_c * eps[s, i, c] -- the same draws for every batch row and every
 * step, as the reference's jax.vmap over the batch with one key does. */
uint64_t pmvae_argmm_sample_workspace_bytes(const pmvae_argmm_config* cfg, int64_t B, int64_t n);
int pmvae_argmm_sample(const pmvae_argmm_config* cfg, const float* params, const float* context, int64_t B,
                       int64_t n, const uint32_t key[2], float* out /* [n,B,d] */, void* ws, uint64_t ws_bytes,
                       pmvae_stream_t stream);
/* reduce_logmeanexp over the leading axis (vae.py:222-223): out[r] = logsumexp_k a[k*B + r] - log K
 * [ - (logsumexp_k c[k*B + r] - log K) when c is given: log p(x) - log p(x_o), vae.py:224 ]. */
int pmvae_logmeanexp_rows(const float* a, const float* c /* or NULL */, float* out, int64_t B, int64_t K,
                          pmvae_stream_t stream);

/* ---- convolutions of the MNIST config's networks (networks.py:9-72), NHWC float32 ----------------
 * One general operator covers hk.Conv2D and hk.Conv2DTranspose (what lax.conv_general_dilated reduces both to):
 *   y[b,oy,ox,co] = act(bias[co] + sum_{ky,kx,ci} Xd(b, oy*stride+ky-pad_top, ox*stride+kx-pad_left, ci)
 *                                                 * w[(ky*KW+kx)*Cin*Cout + ci*w_ci + co*w_co])
 *   Xd = x dilated by `dil` (zeros between samples and outside), act = leaky_relu(slope) (slope 1 = identity).
 *   hk.Conv2D(k, s, SAME|VALID), weights [kh,kw,in,out]:          stride = s, dil = 1, w_ci = Cout, w_co = 1
 *   hk.Conv2DTranspose(k, s, SAME|VALID), weights [kh,kw,out,in]: stride = 1, dil = s, w_ci = 1, w_co = Cin
 * (padding per lax.padtype_to_pads / lax.conv_transpose; posterior_matching_b200/conv.py builds the descriptors).
 * Float32 (exact-parity) or bf16-operand tensor-core GEMMs, selected per descriptor (`reserved` = precision). */
typedef struct pmvae_conv_desc {
  int32_t H, W, Cin;          /* input  [B, H, W, Cin]    */
  int32_t OH, OW, Cout;       /* output [B, OH, OW, Cout] */
  int32_t KH, KW, stride, dil, pad_top, pad_left;
  int32_t w_ci, w_co;         /* element strides of ci / co inside one tap of w */
  float slope;                /* leaky_relu negative slope */
  int32_t reserved;           /* precision: 0 = float32 GEMMs, 1 = bf16 operands on the tcgen05 GEMMs (fp32 accumulate) */
} pmvae_conv_desc;
/* With a workspace (pmvae_conv2d_workspace_bytes, 256-byte aligned) the operator runs as im2col + float32 GEMM;
 * with ws = NULL as direct one-thread-per-element kernels (slow; kept as an independent cross-check). */
uint64_t pmvae_conv2d_workspace_bytes(const pmvae_conv_desc* desc, int64_t B);
int pmvae_conv2d_forward(const pmvae_conv_desc* desc, const float* x, const float* w, const float* bias,
                         int64_t B, float* y, void* ws, uint64_t ws_bytes, pmvae_stream_t stream);
/* VJP: dy (cotangent of y) is overwritten with the pre-activation cotangent; dx may be NULL; dw and dbias are
 * accumulated into (zero them first). */
int pmvae_conv2d_backward(const pmvae_conv_desc* desc, const float* x, const float* w, const float* y,
                          float* dy, int64_t B, float* dx, float* dw, float* dbias, void* ws, uint64_t ws_bytes,
                          pmvae_stream_t stream);

/* ---- building blocks for models composed on the host (the MNIST config: posterior_matching_b200/conv_vae.py) --
 * VJP of pmvae_linear (float32 arithmetic): dw[K,N] += relu?(x)^T dy, db[N] += colsum(dy), dx[B,K] = dy w^T
 * (dx only for relu_in = 0); any of dx / dw / db may be NULL. */
int pmvae_linear_backward(const float* x, const float* w, const float* dy, int64_t B, int32_t K, int32_t N,
                          int32_t relu_in, float* dx, float* dw, float* db, pmvae_stream_t stream);
/* TriLGaussian head algebra on its own (distributions.py:101-113, vae.py:124,130): par[B, d + d(d+1)/2] raw head
 * output, eps[B,d] -> z = mu + L eps, kl = KL(q || N(0,I)); and the VJP for cotangents dz[B,d], g_kl[B]. */
int pmvae_tril_sample_kl(const float* par, const float* eps, int64_t B, int32_t d, float* z, float* kl,
                         pmvae_stream_t stream);
int pmvae_tril_sample_kl_backward(const float* par, const float* eps, const float* dz, const float* g_kl,
                                  int64_t B, int32_t d, float* dpar, pmvae_stream_t stream);
/* optax chain of train_pm_vae.py:74-83 over a flat arena where weight decay applies to every element
 * (or wd = 0, the MNIST config). */
int pmvae_adamw_flat(float* params, const float* grads, float* m, float* v, uint64_t n, int64_t count, float lr,
                     float wd, float b1, float b2, float eps, pmvae_stream_t stream);

/* ---- XLA custom-call targets (jax.ffi / xla_client registration; CustomCallApiVersion
 *      API_VERSION_STATUS_RETURNING = 2, i.e. `custom_call_api_version=2`) --------------
 * The reference is driven by jax.jit / jax.value_and_grad (bax.Trainer, train_pm_vae.py:85,96;
 * eval_pm_vae_uci.py:96), so the binding a maintainer adds is an XLA custom call per entry
 * point.  Each target has the status-returning legacy signature
 *     void target(cudaStream_t stream, void** buffers, const char* opaque, size_t opaque_len,
 *                 XlaCustomCallStatus* status)
 * with `buffers` = the operands followed by the results (all device pointers owned by XLA) and
 * `opaque` = one pmvae_xla_opaque.  Scratch (`ws`) is a result buffer of ws_bytes + PMVAE_XLA_WS_SLACK bytes that
 * XLA allocates (XLA aligns to 256 bytes at most, the tensor path needs 1024: the targets round the pointer up); the forward's
 * `ws` is a residual of the custom_vjp and is passed to the backward as an operand aliased to a
 * result.  posterior_matching_b200/jax_ffi.py registers them; INTEGRATION.md shows the wiring.
 *
 *   pmvae_xla_forward      operands [params, x, b, eps]                        results [rec, kl, match, ws]
 *   pmvae_xla_backward     operands [params, x, b, eps, g_rec, g_kl, g_match, ws]  results [grads, ws (aliased)]
 *   pmvae_xla_is_log_prob  operands [params, x, b]                             results [log_p_x, log_p_xu_given_xo, ws]
 *   pmvae_xla_impute_mean  operands [params, x_o, b]                           results [mean, ws]
 *   pmvae_xla_mask_bernoulli  operands []                                      results [mask]
 */
#define PMVAE_XLA_WS_SLACK 1024
typedef struct pmvae_xla_opaque {
  pmvae_config cfg;
  int64_t B, K, B_total, row_start;
  uint64_t ws_bytes;
  uint32_t key0[2], key1[2];
  float p;           /* Bernoulli rate (mask target)                                   */
  int32_t prepare;   /* != 0: refresh the bf16 operand images first (params changed)   */
  int32_t D;         /* mask width                                                     */
  int32_t reserved;
} pmvae_xla_opaque;
uint64_t pmvae_xla_opaque_size(void);
void pmvae_xla_forward(pmvae_stream_t stream, void** buffers, const char* opaque, size_t opaque_len, void* status);
void pmvae_xla_backward(pmvae_stream_t stream, void** buffers, const char* opaque, size_t opaque_len, void* status);
void pmvae_xla_is_log_prob(pmvae_stream_t stream, void** buffers, const char* opaque, size_t opaque_len, void* status);
void pmvae_xla_impute_mean(pmvae_stream_t stream, void** buffers, const char* opaque, size_t opaque_len, void* status);
void pmvae_xla_mask_bernoulli(pmvae_stream_t stream, void** buffers, const char* opaque, size_t opaque_len, void* status);

#ifdef __cplusplus
}
#endif
#endif /* PMVAE_H_ */
