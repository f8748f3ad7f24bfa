// The training step of train_pm_vae.py as ONE fixed launch sequence on a stream (CUDA-graph replayable):
// per-step keys / beta / learning rate are derived on the device from a small state block instead of coming
// from the host loop, so replaying the captured graph advances the run.
//
// Reference: train_pm_vae.py:58-72 (loss_fn), :28-43 + utils.py:124-136 (beta schedules), :74-83 (optax chain),
// bax.Trainer's per-step rng, networks.py:126 (the dropout keys drawn before z's key),
// masking.py:84-91 (Bernoulli masks; on the JAX stream).
#include <math.h>

#include "kernels.h"
#include "model.h"

namespace pmvae {

__device__ __forceinline__ void dev_split2(uint32_t k0, uint32_t k1, uint32_t (&a)[2], uint32_t (&b)[2]) {
  // jax.random.split(key, 2) = random_bits(key, 4).reshape(2, 2): counters (0,2) and (1,3); out = [x0, y0, x1, y1]
  uint32_t x0 = 0u, x1 = 2u, y0 = 1u, y1 = 3u;
  threefry2x32(k0, k1, x0, x1);
  threefry2x32(k0, k1, y0, y1);
  a[0] = x0; a[1] = y0; b[0] = x1; b[1] = y1;
}

__global__ void advance_state_kernel(StepState* __restrict__ st, pmvae_train_config tc, int R_enc) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  // per-step rng of the transformed loss_fn: hk.PRNGSequence.next() -> (keep, hand out)
  uint32_t keep[2], sub[2];
  dev_split2(st->seq_key[0], st->seq_key[1], keep, sub);
  st->seq_key[0] = keep[0]; st->seq_key[1] = keep[1];
  // inside __call__: the encoder ResidualMLP draws one dropout key per block before z's key (networks.py:126)
  uint32_t k[2] = {sub[0], sub[1]}, nk[2], out[2];
  for (int i = 0; i < R_enc; ++i) { dev_split2(k[0], k[1], nk, out); k[0] = nk[0]; k[1] = nk[1]; }
  dev_split2(k[0], k[1], nk, out);
  st->eps_key[0] = out[0]; st->eps_key[1] = out[1];
  // mask generator: fold_in(base, calls)
  uint32_t m0 = 0u, m1 = st->mask_calls;
  threefry2x32(st->mask_base[0], st->mask_base[1], m0, m1);
  st->mask_key[0] = m0; st->mask_key[1] = m1;
  st->mask_calls += 1u;
  // schedules (double arithmetic like the host loop), train_pm_vae.py:28-43, utils.py:124-136
  const int64_t step = st->step;
  double beta = 1.0;
  if (tc.beta_schedule == PMVAE_BETA_CYCLIC) {
    int64_t c = step - tc.beta_delay;
    const int64_t per = tc.beta_period > 0 ? tc.beta_period : 1;
    int64_t cm = c % per; if (cm < 0) cm += per;               // Python's modulo
    const int64_t half = per / 2;
    if (cm > half) cm = half;
    const double frac = 1.0 - (double)cm / (double)(half > 0 ? half : 1);
    const double xv = ((double)tc.beta_low - (double)tc.beta_high) * frac + (double)tc.beta_high;
    beta = (step >= tc.beta_delay) ? xv : 0.0;
  } else if (tc.beta_schedule == PMVAE_BETA_MONOTONIC) {
    double f = (double)(step - tc.beta_transition_begin) / (double)(tc.beta_transition_steps > 0 ? tc.beta_transition_steps : 1);
    f = f < 0.0 ? 0.0 : (f > 1.0 ? 1.0 : f);
    beta = (double)tc.beta_low + ((double)tc.beta_high - (double)tc.beta_low) * f;
  }
  st->beta = (float)beta;
  st->lr = (float)((double)tc.lr_init * pow((double)tc.lr_decay_rate, (double)step / (double)(tc.lr_transition_steps > 0 ? tc.lr_transition_steps : 1)));
  const double t = (double)step + 1.0;
  st->bc1 = (float)(1.0 - pow((double)tc.adam_b1, t));
  st->bc2 = (float)(1.0 - pow((double)tc.adam_b2, t));
}

__global__ void bump_step_kernel(StepState* __restrict__ st) {
  if (threadIdx.x == 0 && blockIdx.x == 0) st->step += 1;
}

int adamw_step_dev(const pmvae_config* c, float* params, const float* grads, float* m, float* v, float wd, float b1,
                   float b2, float eps, const StepState* st, cudaStream_t s);   // model.cu

}  // namespace pmvae

using namespace pmvae;

extern "C" {

uint64_t pmvae_train_state_bytes(void) { return sizeof(StepState); }
uint64_t pmvae_train_scratch_floats(const pmvae_config* cfg, int64_t B) {
  if (!cfg || B < 0) return 0;
  return (uint64_t)B * (uint64_t)(cfg->D + cfg->d + 6) + 64;
}

int pmvae_train_state_init(void* state, const uint32_t seq_key[2], const uint32_t mask_key[2], int64_t step,
                           uint32_t mask_calls, pmvae_stream_t stream) {
  PMVAE_CHECK(state && seq_key && mask_key && step >= 0, "bad arguments");
  StepState h{};
  h.seq_key[0] = seq_key[0]; h.seq_key[1] = seq_key[1];
  h.mask_base[0] = mask_key[0]; h.mask_base[1] = mask_key[1];
  h.mask_calls = mask_calls; h.step = step; h.beta = 1.f;
  PMVAE_CUDA(cudaMemcpyAsync(state, &h, sizeof(h), cudaMemcpyHostToDevice, as_stream(stream)));
  PMVAE_CUDA(cudaStreamSynchronize(as_stream(stream)));      // `h` lives on this stack frame
  return 0;
}

int pmvae_train_state_read(const void* state, pmvae_train_state_host* out, pmvae_stream_t stream) {
  PMVAE_CHECK(state && out, "bad arguments");
  StepState h{};
  PMVAE_CUDA(cudaMemcpyAsync(&h, state, sizeof(h), cudaMemcpyDeviceToHost, as_stream(stream)));
  PMVAE_CUDA(cudaStreamSynchronize(as_stream(stream)));
  out->seq_key[0] = h.seq_key[0]; out->seq_key[1] = h.seq_key[1];
  out->mask_key[0] = h.mask_key[0]; out->mask_key[1] = h.mask_key[1];
  out->eps_key[0] = h.eps_key[0]; out->eps_key[1] = h.eps_key[1];
  out->mask_calls = h.mask_calls; out->step = h.step;
  out->beta = h.beta; out->lr = h.lr; out->bc1 = h.bc1; out->bc2 = h.bc2;
  return 0;
}

// phase bit 0 (PMVAE_STEP_FWD_BWD): state advance, mask + eps draw, forward, loss cotangents (+ batch sums), backward
// phase bit 1 (PMVAE_STEP_UPDATE):  AdamW + refresh of the bf16 operand images + step counter
// With PMVAE_STEP_SPLIT_BWD OR-ed into bit 0 the backward stops after the decoder + latent stage; PMVAE_STEP_BWD_ENC and
// PMVAE_STEP_BWD_PART run the remaining two stages, so that the caller can exchange each finished gradient range
// (decoder / encoder / partial encoder are contiguous in the arena) while the next stage computes.
int pmvae_train_step(const pmvae_config* cfg, const pmvae_train_config* tc, float* params, float* m, float* v,
                     float* grads, void* state, const float* x, int64_t B, int64_t B_global, int64_t row_start,
                     float* scratch, float* out_sums, void* ws, uint64_t ws_bytes, int32_t phase,
                     pmvae_stream_t stream) {
  PMVAE_CHECK(cfg && tc && params && m && v && grads && state && out_sums && ws && (x || B == 0), "null pointer");
  PMVAE_CHECK(B >= 0 && B_global >= B && row_start >= 0 && row_start + B <= B_global, "bad row range");
  cudaStream_t s = as_stream(stream);
  StepState* st = reinterpret_cast<StepState*>(state);
  const int D = cfg->D, d = cfg->d;
  float* b = scratch;
  float* eps = b + (uint64_t)B * D;
  float* terms = eps + (uint64_t)B * d;          // rec | kl | match
  float* cot = terms + 3ull * B;                 // g_rec | g_kl | g_match
  if (phase & 1) {
    PMVAE_CHECK(scratch != nullptr || B == 0, "null scratch");
    advance_state_kernel<<<1, 32, 0, s>>>(st, *tc, cfg->R_enc);
    PMVAE_LAUNCH_CHECK();
    PMVAE_TRY(mask_bernoulli_dev(st, tc->mask_p, (uint64_t)B_global, (uint64_t)row_start, (uint64_t)B, D, b, s));
    PMVAE_TRY(normal_dev(st, (uint64_t)B_global * d, (uint64_t)row_start * d, (uint64_t)B * d, eps, s));
    PMVAE_TRY(pmvae_forward(cfg, params, x, b, eps, B, terms, terms + B, terms + 2 * B, ws, ws_bytes, stream));
    PMVAE_TRY(loss_cotangents(B, B_global, 0.f, tc->matching_coef, terms, terms + B, terms + 2 * B, cot, cot + B,
                              cot + 2 * B, out_sums, s, st));
    PMVAE_TRY(backward_staged(cfg, params, x, b, eps, B, cot, cot + B, cot + 2 * B, grads,
                              (phase & PMVAE_STEP_SPLIT_BWD) ? 1 : 7, ws, ws_bytes, s));
  }
  if (phase & (PMVAE_STEP_BWD_ENC | PMVAE_STEP_BWD_PART)) {
    PMVAE_CHECK(!(phase & 1) || (phase & PMVAE_STEP_SPLIT_BWD), "BWD_ENC / BWD_PART follow a SPLIT_BWD call");
    const int st_bits = ((phase & PMVAE_STEP_BWD_ENC) ? 2 : 0) | ((phase & PMVAE_STEP_BWD_PART) ? 4 : 0);
    PMVAE_TRY(backward_staged(cfg, params, x, b, eps, B, cot, cot + B, cot + 2 * B, grads, st_bits, ws, ws_bytes, s));
  }
  if (phase & 2) {
    PMVAE_TRY(adamw_step_dev(cfg, params, grads, m, v, tc->weight_decay, tc->adam_b1, tc->adam_b2, tc->adam_eps, st, s));
    PMVAE_TRY(pmvae_prepare_params(cfg, params, ws, ws_bytes, stream));
    bump_step_kernel<<<1, 32, 0, s>>>(st);
    PMVAE_LAUNCH_CHECK();
  }
  return 0;
}

}  // extern "C"
