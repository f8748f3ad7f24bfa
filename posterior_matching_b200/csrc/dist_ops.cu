// Stand-alone distribution algebra behind the module attributes of PosteriorMatchingVAE (vae.py:47-57): the
// objects `.encoder(x)`, `.partial_encoder(x_o_b)` (tfd.MultivariateNormalTriL, distributions.py:101-113),
// `.decoder(z)` (tfd.Normal with one shared scale, distributions.py:41-55) and `.prior`
// (tfd.MultivariateNormalDiag(0, 1), vae.py:55-57) expose .mean() / .sample() / .log_prob() / .entropy();
// lookahead.py:126-133,219-222 and the evaluation scripts call them.  Small HBM-bound row kernels; the
// K-sample draw and the TriL log-prob reuse the kernels of the training / evaluation path (latent.cu).
#include "kernels.h"

namespace pmvae {

static int grid_rows(int64_t work, int block) {
  int64_t g = ceil_div(work, block);
  if (g > 148 * 8) g = 148 * 8;
  return (int)(g < 1 ? 1 : g);
}

// entropy of N(mu, L L^T): d/2 (1 + log 2 pi) + sum_i log L_ii, L_ii = softplus(raw_ii) + 1e-5 (FillScaleTriL);
// the diagonal (i, i) of tfp's fill_triangular sits at c[i*d + i], c = concat(v[d:], reverse(v))
__global__ void __launch_bounds__(256) tril_entropy_kernel(const float* __restrict__ par, int64_t B, int d,
                                                           float* __restrict__ out) {
  const int m = d * (d + 1) / 2, P = d + m;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < B; r += (int64_t)gridDim.x * blockDim.x) {
    const float* v = par + r * P + d;
    float acc = 0.f;
    for (int i = 0; i < d; ++i) {
      const int k = i * d + i;
      const float raw = v[(k < m - d) ? (d + k) : (m - 1 - (k - (m - d)))];
      acc += logf(softplus_f(raw) + 1e-5f);
    }
    out[r] = 0.5f * d * (1.0f + kLog2Pi) + acc;
  }
}

// tfd.Normal(loc, exp(log_scale)).log_prob(x), elementwise over [rows, D] (loc has row pitch ld_loc); x is
// broadcast over `reps` leading repetitions of its `rows_x` rows (the [K, B, D] case of vae.py:197-199)
__global__ void __launch_bounds__(256) normal_log_prob_kernel(const float* __restrict__ x, const float* __restrict__ loc,
                                                              int64_t ld_loc, const float* __restrict__ log_scale,
                                                              int64_t rows, int64_t rows_x, int D,
                                                              float* __restrict__ out) {
  const float ls = *log_scale, inv = __expf(-ls);
  const int64_t n = rows * D;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / D;
    const int j = (int)(i - r * D);
    const float t = (x[(r % rows_x) * D + j] - loc[r * ld_loc + j]) * inv;
    out[i] = -0.5f * t * t - ls - 0.5f * kLog2Pi;
  }
}

// tfd.MultivariateNormalDiag(0, 1).log_prob(z): -|z|^2 / 2 - d/2 log 2 pi
__global__ void __launch_bounds__(256) std_normal_log_prob_kernel(const float* __restrict__ z, int64_t B, int d,
                                                                  float* __restrict__ out) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < B; r += (int64_t)gridDim.x * blockDim.x) {
    float q = 0.f;
    for (int j = 0; j < d; ++j) { const float v = z[r * d + j]; q = fmaf(v, v, q); }
    out[r] = -0.5f * q - 0.5f * d * kLog2Pi;
  }
}

// DiagonalGaussian (distributions.py:58-84): par[B, 2d] = [loc | raw], scale = softplus(raw) + 1e-5.
// z = loc + scale * eps (the reparameterised sample behind posterior.sample, vade.py:259);
// log_prob(z) = sum_j -0.5 ((z - loc) / scale)^2 - log scale - 0.5 log 2 pi;  entropy = sum_j 0.5 (1 + log 2 pi) + log scale
__global__ void __launch_bounds__(256) diag_sample_kernel(const float* __restrict__ par, const float* __restrict__ eps,
                                                          int64_t B, int d, float* __restrict__ z) {
  const int64_t n = B * d;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / d;
    const int j = (int)(i - r * d);
    z[i] = par[r * 2 * d + j] + (softplus_f(par[r * 2 * d + d + j]) + 1e-5f) * eps[i];
  }
}
__global__ void __launch_bounds__(256) diag_log_prob_kernel(const float* __restrict__ par, const float* __restrict__ z,
                                                            int64_t B, int d, float* __restrict__ out_lp,
                                                            float* __restrict__ out_ent) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < B; r += (int64_t)gridDim.x * blockDim.x) {
    float lp = 0.f, ent = 0.f;
    for (int j = 0; j < d; ++j) {
      const float sc = softplus_f(par[r * 2 * d + d + j]) + 1e-5f;
      const float ls = logf(sc);
      if (z) { const float t = (z[r * d + j] - par[r * 2 * d + j]) / sc; lp += -0.5f * t * t - ls - 0.5f * kLog2Pi; }
      ent += 0.5f * (1.0f + kLog2Pi) + ls;
    }
    if (out_lp) out_lp[r] = lp;
    if (out_ent) out_ent[r] = ent;
  }
}


// ---------------------------------------------------------------- lookahead posteriors (lookahead.py:14-39,183-199)
// par [B, S, 2d] = the selected LookaheadBlock outputs (loc | raw scale, scale = softplus(raw) + 1e-5),
// z [K, B, S, d] = one-step latent samples of the model, valid [B, S] in {0, 1}:
//     ll[b] = sum_s valid[b,s] * mean_k log N(z[k,b,s] ; loc[b,s], diag(scale[b,s])^2) / #valid[b]     (0 if none valid)
// One block per row b; warp = one (b, s) pair at a time, lanes over the latent dimension.
__global__ void __launch_bounds__(256) lookahead_ll_kernel(const float* __restrict__ par, const float* __restrict__ z,
                                                           const float* __restrict__ valid, int64_t K, int64_t B, int S,
                                                           int d, float* __restrict__ out) {
  __shared__ float s_sum[8], s_cnt[8];
  const int64_t b = blockIdx.x;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float sum = 0.f, cnt = 0.f;
  for (int s = w; s < S; s += 8) {
    if (valid[b * S + s] == 0.f) continue;
    const float* pr = par + (b * S + s) * 2 * d;
    float acc = 0.f;
    for (int j = lane; j < d; j += 32) {
      const float loc = pr[j], sc = softplus_f(pr[d + j]) + 1e-5f, inv = 1.0f / sc;
      float q = 0.f;
      for (int64_t k = 0; k < K; ++k) {
        const float t = (z[((k * B + b) * S + s) * d + j] - loc) * inv;
        q = fmaf(t, t, q);
      }
      acc += -0.5f * q / (float)K - logf(sc) - 0.5f * kLog2Pi;
    }
    sum += warp_sum(acc);
    cnt += 1.f;
  }
  if (lane == 0) { s_sum[w] = sum; s_cnt[w] = cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f, c = 0.f;
    for (int i = 0; i < 8; ++i) { t += s_sum[i]; c += s_cnt[i]; }
    out[b] = c > 0.f ? t / c : 0.f;
  }
}

// d (sum_b g[b] ll[b]) / d par: thread = one (b, s, j).
__global__ void __launch_bounds__(256) lookahead_ll_bwd_kernel(const float* __restrict__ par, const float* __restrict__ z,
                                                               const float* __restrict__ valid,
                                                               const float* __restrict__ g, int64_t K, int64_t B, int S,
                                                               int d, float* __restrict__ dpar) {
  const int64_t n = B * S * d;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(t % d);
    const int64_t bs = t / d, b = bs / S;
    float cnt = 0.f;
    for (int s = 0; s < S; ++s) cnt += valid[b * S + s];
    float dloc = 0.f, draw = 0.f;
    if (valid[bs] != 0.f && cnt > 0.f) {
      const float loc = par[bs * 2 * d + j], raw = par[bs * 2 * d + d + j];
      const float sc = softplus_f(raw) + 1e-5f, inv = 1.0f / sc;
      float a1 = 0.f, a2 = 0.f;
      for (int64_t k = 0; k < K; ++k) {
        const float u = z[(k * B * S + bs) * d + j] - loc;
        a1 += u; a2 = fmaf(u, u, a2);
      }
      const float coef = g[b] / (cnt * (float)K);
      dloc = coef * a1 * inv * inv;
      draw = coef * (a2 * inv * inv * inv - (float)K * inv) * sigmoid_f(raw);
    }
    dpar[bs * 2 * d + j] = dloc;
    dpar[bs * 2 * d + d + j] = draw;
  }
}

}  // namespace pmvae

using namespace pmvae;

extern "C" {

int pmvae_tril_log_prob(const float* par, const float* z, int64_t B, int32_t d, float* out, pmvae_stream_t stream) {
  PMVAE_CHECK(B >= 0 && (B == 0 || (par && z && out)), "bad arguments");
  return match_fwd(par, z, out, B, d, as_stream(stream));
}

int pmvae_tril_entropy(const float* par, int64_t B, int32_t d, float* out, pmvae_stream_t stream) {
  PMVAE_CHECK(B >= 0 && d >= 1 && d <= 64 && (B == 0 || (par && out)), "bad arguments");
  if (B == 0) return 0;
  tril_entropy_kernel<<<grid_rows(B, 256), 256, 0, as_stream(stream)>>>(par, B, d, out);
  PMVAE_LAUNCH_CHECK();
  return 0;
}

int pmvae_tril_sample(const float* par, const uint32_t key[2], int64_t B, int64_t K, int64_t B_total, int64_t row_start,
                      int32_t d, float* z, float* log_ratio, pmvae_stream_t stream) {
  PMVAE_CHECK(B >= 0 && K >= 1 && row_start >= 0 && row_start + B <= B_total, "bad K / row range");
  PMVAE_CHECK(B == 0 || (par && key && z && log_ratio), "null pointer");
  return sample_latents(par, Key2{key[0], key[1]}, B, K, B_total, row_start, d, z, log_ratio, as_stream(stream));
}

int pmvae_normal_log_prob(const float* x, const float* loc, const float* log_scale, int64_t rows, int64_t rows_x,
                          int32_t D, float* out, pmvae_stream_t stream) {
  PMVAE_CHECK(rows >= 0 && D >= 1 && rows_x >= 1 && (rows == 0 || (x && loc && log_scale && out)), "bad arguments");
  PMVAE_CHECK(rows % rows_x == 0, "rows must be a multiple of the rows of x");
  if (rows == 0) return 0;
  normal_log_prob_kernel<<<grid_rows(rows * D, 256), 256, 0, as_stream(stream)>>>(x, loc, D, log_scale, rows, rows_x, D, out);
  PMVAE_LAUNCH_CHECK();
  return 0;
}

int pmvae_std_normal_log_prob(const float* z, int64_t B, int32_t d, float* out, pmvae_stream_t stream) {
  PMVAE_CHECK(B >= 0 && d >= 1 && (B == 0 || (z && out)), "bad arguments");
  if (B == 0) return 0;
  std_normal_log_prob_kernel<<<grid_rows(B, 256), 256, 0, as_stream(stream)>>>(z, B, d, out);
  PMVAE_LAUNCH_CHECK();
  return 0;
}

int pmvae_diag_sample(const float* par, const float* eps, int64_t B, int32_t d, float* z, pmvae_stream_t stream) {
  PMVAE_CHECK(B >= 0 && d >= 1 && (B == 0 || (par && eps && z)), "bad arguments");
  if (B == 0) return 0;
  diag_sample_kernel<<<grid_rows(B * d, 256), 256, 0, as_stream(stream)>>>(par, eps, B, d, z);
  PMVAE_LAUNCH_CHECK();
  return 0;
}

int pmvae_diag_log_prob(const float* par, const float* z, int64_t B, int32_t d, float* out_log_prob, float* out_entropy,
                        pmvae_stream_t stream) {
  PMVAE_CHECK(B >= 0 && d >= 1 && (B == 0 || (par && (out_log_prob || out_entropy))), "bad arguments");
  PMVAE_CHECK(out_log_prob == nullptr || z != nullptr, "log_prob needs z");
  if (B == 0) return 0;
  diag_log_prob_kernel<<<grid_rows(B, 256), 256, 0, as_stream(stream)>>>(par, z, B, d, out_log_prob, out_entropy);
  PMVAE_LAUNCH_CHECK();
  return 0;
}

int pmvae_lookahead_ll(const float* par, const float* z, const float* valid, int64_t K, int64_t B, int32_t S, int32_t d,
                       float* out_ll, pmvae_stream_t stream) {
  PMVAE_CHECK(K >= 1 && B >= 0 && S >= 1 && d >= 1 && (B == 0 || (par && z && valid && out_ll)), "bad arguments");
  if (B == 0) return 0;
  lookahead_ll_kernel<<<(unsigned)B, 256, 0, as_stream(stream)>>>(par, z, valid, K, B, S, d, out_ll);
  PMVAE_LAUNCH_CHECK();
  return 0;
}

int pmvae_lookahead_ll_backward(const float* par, const float* z, const float* valid, const float* g, int64_t K,
                                int64_t B, int32_t S, int32_t d, float* dpar, pmvae_stream_t stream) {
  PMVAE_CHECK(K >= 1 && B >= 0 && S >= 1 && d >= 1 && (B == 0 || (par && z && valid && g && dpar)), "bad arguments");
  if (B == 0) return 0;
  lookahead_ll_bwd_kernel<<<grid_rows(B * S * d, 256), 256, 0, as_stream(stream)>>>(par, z, valid, g, K, B, S, d, dpar);
  PMVAE_LAUNCH_CHECK();
  return 0;
}

int pmvae_tril_log_prob_backward(const float* par, const float* z, const float* g, int64_t B, int32_t d, float* dpar,
                                 float* dz, void* ws, uint64_t ws_bytes, pmvae_stream_t stream) {
  PMVAE_CHECK(B >= 0 && d >= 1 && d <= 64 && (B == 0 || (par && z && g && dpar && ws)), "bad arguments");
  if (B == 0) return 0;
  const int64_t P = d + (int64_t)d * (d + 1) / 2;
  const uint64_t need = (uint64_t)(B * P + B * d + B) * sizeof(float);
  PMVAE_CHECK(ws_bytes >= need, "workspace too small: (B P + B d + B) floats");
  cudaStream_t s = as_stream(stream);
  // The two-head backward kernel with an inert "encoder" head: eps = 0 and g_kl = 0 leave d / d(loc of that head) =
  // the total gradient into z, which without a decoder term and without stop_gradient is exactly -g * L^-T L^-1 (z - mu).
  float* tmp = reinterpret_cast<float*>(ws);            // [B, P] gradient of the inert head
  float* zeros = tmp + B * P;                           // eps [B, d] and g_kl [B]
  PMVAE_CUDA(cudaMemsetAsync(zeros, 0, (size_t)(B * d + B) * sizeof(float), s));
  PMVAE_TRY(latent_bwd(par, par, zeros, z, nullptr, zeros + B * d, g, 0, tmp, dpar, nullptr, nullptr, B, d, s));
  if (dz) PMVAE_CUDA(cudaMemcpy2DAsync(dz, (size_t)d * sizeof(float), tmp, (size_t)P * sizeof(float), (size_t)d * sizeof(float),
                                       (size_t)B, cudaMemcpyDeviceToDevice, s));
  return 0;
}

}  // extern "C"
