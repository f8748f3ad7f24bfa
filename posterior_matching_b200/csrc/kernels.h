// Internal launchers shared between the translation units of libpmvae.
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace pmvae {

// Device-resident state of the fused training step (train_step.cu): everything a step needs from the host-side
// loop of train_pm_vae.py (per-step PRNG keys, beta, learning rate, Adam bias corrections) is derived on the
// device from this block, so that a whole step is a fixed launch sequence (CUDA-graph replayable).
struct StepState {
  uint32_t seq_key[2];    // Haiku PRNGSequence key behind the per-step rng of the transformed loss_fn
  uint32_t mask_base[2];  // mask generator key; call c draws with fold_in(mask_base, c)
  uint32_t mask_key[2];   // derived for the current step
  uint32_t eps_key[2];    // derived for the current step (key of z ~ q(z|x), vae.py:124)
  uint32_t mask_calls, pad0;
  int64_t step;           // optimizer updates applied before the current step
  float beta, lr, bc1, bc2;
};

// ---- gemm_f32.cu
struct GemmF32Args {
  int64_t M, N, K;
  const float* A; int64_t lda;
  const float* B; int64_t ldb;
  float* C; int64_t ldc;
  const float* bias = nullptr;
  const float* mask = nullptr; int64_t ldmask = 0;   // out *= (mask > 0)
  const float* resid = nullptr; int64_t ldresid = 0; // out += resid
  int relu_a = 0;
  int atomic = 0;   // atomicAdd the raw product into C (split-K over gridDim.z)
  int split_k = 1;
  // `batch` independent problems of the same shape in one launch: member i uses A + i sA, B + i sB, C + i sC, ... (elements)
  int batch = 1;
  int64_t sA = 0, sB = 0, sC = 0, sBias = 0, sMask = 0, sResid = 0;
};
int gemm_f32(const GemmF32Args& a, bool ta, bool tb, cudaStream_t stream);

// ---- elementwise.cu
int concat_masked(const float* x, const float* b, float* out, int64_t B, int D, cudaStream_t s);
// y <- LN(y) in place (hk.LayerNorm(-1,False,False), eps 1e-5), rstd[B] saved;
// if resid: out_sum = resid + LN(y)
int ln_fwd(float* y, float* rstd, const float* resid, float* out_sum, int64_t B, int N, cudaStream_t s);
// dx = rstd * (dy - mean(dy) - xhat * mean(dy * xhat)); dx may alias dy
int ln_bwd(const float* dy, const float* xhat, const float* rstd, float* dx, int64_t B, int N, cudaStream_t s);
// out[n] += sum_m dY[m, n]   (atomics)
int colsum_add(const float* dY, int64_t ld, float* out, int64_t B, int N, cudaStream_t s);
// rec[r] = sum_j w[r,j] * logN(x[r,j]; loc[r,j], exp(ls))   (w == nullptr -> 1)
int rec_ll(const float* x, const float* loc, int64_t ld_loc, const float* log_scale, const float* w, float* out,
           int64_t B, int D, cudaStream_t s);
// dloc[r,j] = g[r] * (x - loc) * exp(-2 ls);  *dls += sum_r g[r] * sum_j ((x-loc)^2 exp(-2 ls) - 1)
// (either dloc output may be null; the bf16 one is zero-padded to its pitch)
// db (optional): += column sums of the bf16 dloc (the decoder head's bias gradient)
int rec_ll_bwd(const float* x, const float* loc, int64_t ld_loc, const float* log_scale, const float* g, float* dloc,
               __nv_bfloat16* dloc_bf16, int64_t ld_dloc, float* dls, int64_t B, int D, cudaStream_t s,
               float* db = nullptr);
// st (optional): beta is read from the device step state instead
int loss_cotangents(int64_t B, int64_t B_global, float beta, float coef, const float* rec, const float* kl,
                    const float* match, float* g_rec, float* g_kl, float* g_match, float* out_sums, cudaStream_t s,
                    const StepState* st = nullptr);
// no-decay (bias) ranges: one per Linear of the three nets (<= 2 * kMaxBlocks + 1 = 17 each) + three heads
constexpr int kAdamSegCap = 3 * (2 * 8 + 1) + 3;
struct AdamSegs { int n; uint32_t beg[kAdamSegCap]; uint32_t end[kAdamSegCap]; };
// st (optional): lr and the bias corrections are read from the device step state instead
int adamw(float* p, const float* g, float* m, float* v, uint64_t n, const AdamSegs& nodecay, float lr, float wd,
          float b1, float b2, float eps, float bc1, float bc2, cudaStream_t s, const StepState* st = nullptr);
// rng.cu: draws whose key lives in the device step state
int mask_bernoulli_dev(const StepState* st, float p, uint64_t B_total, uint64_t row_start, uint64_t rows, int D, float* out,
                       cudaStream_t s);
int normal_dev(const StepState* st, uint64_t n_total, uint64_t start, uint64_t count, float* out, cudaStream_t s);
// evaluators
// ll[k*B + r] = sum_j w * logN(x[r]; loc[k*B + r]) + base[k*B + r]
int eval_rows_ll(const float* x, const float* w, const float* loc, int64_t ld_loc, const float* log_scale,
                 const float* base, float* out, int64_t B, int64_t K, int D, cudaStream_t s);
// out[r] = logsumexp_k(a[k*B + r]) - log K  [ - (logsumexp_k(c[k*B+r]) - log K) if c ]
int logmeanexp_rows(const float* a, const float* c, float* out, int64_t B, int64_t K, cudaStream_t s);
int impute_mean(const float* x, const float* b, const float* loc, int64_t ld_loc, float* out, int64_t B, int64_t K,
                int D, cudaStream_t s);
// out[(k * B_all + r) * D + j] = b ? x * b : loc[(k * nb + r) * ld_loc + j]   (one chunk of nb data rows, K samples)
int impute_samples(const float* x, const float* b, const float* loc, int64_t ld_loc, float* out, int64_t nb, int64_t B_all,
                   int64_t K, int D, cudaStream_t s);

// ---- latent.cu  (par = raw TriL head output [B, P], P = d + d(d+1)/2)
int latent_fwd(const float* par, const float* eps, float* z, float* kl, int64_t B, int d, cudaStream_t s);
// `saved` (training forward, d = 64 only): three [*, d] vectors at saved, saved + saved_stride, saved + 2 saved_stride that
// latent_bwd takes back through the same two arguments (match_fwd_saves(d) tells whether the pair is used).
int match_fwd(const float* par_p, const float* z, float* match, int64_t B, int d, cudaStream_t s, float* saved = nullptr,
              int64_t saved_stride = 0);
bool match_fwd_saves(int d);
int latent_bwd(const float* par_e, const float* par_p, const float* eps, const float* z, const float* dz_dec,
               const float* g_kl, const float* g_match, int stop_grad, float* dpar_e, float* dpar_p,
               __nv_bfloat16* dpar_e_b, __nv_bfloat16* dpar_p_b, int64_t B, int d, cudaStream_t s,
               float* db_e = nullptr, float* db_p = nullptr, bool* db_done = nullptr, const float* saved = nullptr,
               int64_t saved_stride = 0);
bool latent_bwd_bias_fused(int d);
int tril_sample_bwd(const float* par, const float* eps, const float* dz, const float* g_kl, float* dpar, int64_t B, int d,
                    cudaStream_t s);
// latent16.cu: thread-per-row versions for d = 16 (bf16 gradient outputs only)
int latent_fwd16(const float* par, const float* eps, float* z, float* kl, int64_t B, cudaStream_t s);
int match_fwd16(const float* par_p, const float* z, float* match, int64_t B, cudaStream_t s);
int sample_latents16(const float* par, Key2 key, int64_t B, int64_t K, int64_t B_total, int64_t row_start, float* z,
                     float* base, cudaStream_t s);
int latent_bwd16(const float* par_e, const float* par_p, const float* eps, const float* z, const float* dz_dec,
                 const float* g_kl, const float* g_match, int stop_grad, __nv_bfloat16* dpar_e_b,
                 __nv_bfloat16* dpar_p_b, float* db_e, float* db_p, int64_t B, cudaStream_t s);
// tensor.cu: one hidden 256 x 256 Linear of a float32 net on the tcgen05 GEMMs (bf16 operand copies, fp32 results)
uint64_t hidden_images_bytes(int n_leaves);
int hidden_images_pack(const float* params, const struct Leaf* leaves, int n, void* img, const __nv_bfloat16** wn,
                       const __nv_bfloat16** wt, cudaStream_t s);
int hidden_fwd_tc(const float* x, const __nv_bfloat16* wt, const float* bias, const float* resid, int64_t M, float* y,
                  __nv_bfloat16* xb, cudaStream_t s);
int hidden_bwd_tc(const float* x, const float* dy, const __nv_bfloat16* wn, float* gW, float* dx, const float* resid,
                  int64_t M, __nv_bfloat16* xb, __nv_bfloat16* dyb, cudaStream_t s);
// latent64.cu: warp-per-row versions for d = 64 (packed factor image in shared memory; bf16 gradient outputs only)
int latent_fwd64(const float* par, const float* eps, float* z, float* kl, int64_t B, cudaStream_t s);
// match_fwd64 with save_* != NULL (training forward) also writes r = L_p^-1 (z - mu_p), g = L_p^-T r and the diagonal
// terms qd, [B, 64] each; latent_bwd64 consumes them instead of re-reading par_p.
int match_fwd64(const float* par_p, const float* z, float* match, int64_t B, cudaStream_t s, float* save_r = nullptr,
                float* save_g = nullptr, float* save_qd = nullptr);
int latent_bwd64(const float* par_e, const float* eps, const float* dz_dec, const float* g_kl, const float* g_match,
                 int stop_grad, __nv_bfloat16* dpar_e_b, __nv_bfloat16* dpar_p_b, float* db_e, float* db_p,
                 const float* vec_r, const float* vec_g, const float* vec_qd, int64_t B, cudaStream_t s);
// z[k,r,:] = mu_r + L_r eps[k,r,:], eps = normal(key, [K, B_total, d]) rows row_start..;
// base[k,r] = log N(z;0,I) - log q(z) = -0.5|z|^2 + 0.5|eps|^2 + sum log L_ii
int sample_latents(const float* par, Key2 key, int64_t B, int64_t K, int64_t B_total, int64_t row_start, int d,
                   float* z, float* base, cudaStream_t s);

// ---- dists_mnist.cu (MNIST-config heads: Bernoulli decoder, autoregressive GMM partial posterior)
int bernoulli_ll(const float* logits, const float* x, const float* w, int64_t B, int D, float* out, cudaStream_t s);
int bernoulli_ll_bwd(const float* logits, const float* x, const float* w, const float* g, int64_t B, int D,
                     float* dlogits, cudaStream_t s);
int argmm_input(const float* z, const float* ctx, int64_t B, int d, int C, float* X, cudaStream_t s);
int argmm_lp(const float* head_out, const float* z, int64_t B, int d, int K, float* out, cudaStream_t s);
int argmm_lp_bwd(const float* head_out, const float* z, const float* g, int64_t B, int d, int K, float* d_head,
                 float* dz_direct, cudaStream_t s);
int argmm_sample_input(const float* x, const float* ctx, int64_t M, int64_t B, int d, int C, int step, float* X, cudaStream_t s);
int argmm_sample_step(const float* head_out, const float* eps, const float* u, int64_t M, int64_t B, int d, int K, int step,
                      float* x, cudaStream_t s);
int argmm_reduce_dx(const float* dX, const float* dz_direct, int64_t B, int d, int C, float* dz, float* dctx,
                    cudaStream_t s);

}  // namespace pmvae
