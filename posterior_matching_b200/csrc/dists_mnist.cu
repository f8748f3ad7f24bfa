// Distribution heads only the MNIST config uses (configs/pm_vae_mnist.py:19-21), fp32:
//   Bernoulli decoder log-prob            distributions.py:20-25, vae.py:127-128
//   AutoregressiveGMM partial posterior   distributions.py:116-134 (OneDimensionalGMM),
//                                         :152-166 (_AutoregressiveDistribution.log_prob)
// The d autoregressive steps are independent given z, so they are batched as d*B rows (row i*B + b is
// step i of sample b) through one ResidualMLP pass; these kernels build that batch, evaluate the mixture
// log-density of dimension i from the 3K head columns of step i, and run the matching backward.
#include "kernels.h"

namespace pmvae {

static int grid1d(int64_t work, int block, int per_sm = 8) {
  int64_t g = ceil_div(work, block);
  const int64_t cap = 148ll * per_sm;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// ---------------------------------------------------------------- Bernoulli(logits).log_prob(x), row sums
// ll[r] = sum_j w[r,j] * (-x softplus(-l) - (1 - x) softplus(l))   (x is a float in [0, 1]; w optional)
__global__ void __launch_bounds__(256) bernoulli_ll_kernel(const float* __restrict__ logits, const float* __restrict__ x,
                                                           const float* __restrict__ w, int64_t B, int D,
                                                           float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < B; r += nwarps) {
    float acc = 0.f;
    for (int j = lane; j < D; j += 32) {
      const float l = logits[r * D + j], xv = x[r * D + j];
      float t = -xv * softplus_f(-l) - (1.0f - xv) * softplus_f(l);
      if (w) t *= w[r * D + j];
      acc += t;
    }
    acc = warp_sum(acc);
    if (lane == 0) out[r] = acc;
  }
}
// dlogits[r,j] = g[r] * w[r,j] * (x - sigmoid(l))
__global__ void __launch_bounds__(256) bernoulli_ll_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ x,
                                                               const float* __restrict__ w, const float* __restrict__ g,
                                                               int64_t B, int D, float* __restrict__ dlogits) {
  const int64_t n = B * D;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / D;
    float t = g[r] * (x[i] - sigmoid_f(logits[i]));
    if (w) t *= w[i];
    dlogits[i] = t;
  }
}
int bernoulli_ll(const float* logits, const float* x, const float* w, int64_t B, int D, float* out, cudaStream_t s) {
  if (B == 0) return 0;
  bernoulli_ll_kernel<<<grid1d(B * 32, 256), 256, 0, s>>>(logits, x, w, B, D, out);
  PMVAE_LAUNCH_CHECK();
  return 0;
}
int bernoulli_ll_bwd(const float* logits, const float* x, const float* w, const float* g, int64_t B, int D,
                     float* dlogits, cudaStream_t s) {
  if (B == 0) return 0;
  bernoulli_ll_bwd_kernel<<<grid1d(B * D, 256), 256, 0, s>>>(logits, x, w, g, B, D, dlogits);
  PMVAE_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------- AR-GMM: batched step inputs
// X[i*B + b] = [ z[b] * (arange(d) < i), (arange(d) < i), context[b] ]      (distributions.py:153-161)
__global__ void __launch_bounds__(256) argmm_input_kernel(const float* __restrict__ z, const float* __restrict__ ctx,
                                                          int64_t B, int d, int C, float* __restrict__ X) {
  const int F = 2 * d + C;
  const int64_t n = (int64_t)d * B * F;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = t / F;
    const int j = (int)(t - row * F);
    const int64_t i = row / B, b = row - i * B;
    float v;
    if (j < d) v = (j < i) ? z[b * d + j] : 0.f;
    else if (j < 2 * d) v = (j - d < i) ? 1.f : 0.f;
    else v = ctx[b * C + (j - 2 * d)];
    X[t] = v;
  }
}

constexpr int kMaxComp = 32;

// mixture terms of one (step, sample): a_k = log_softmax(logits)_k + log N(zv; mean_k, scale_k); returns logsumexp
__device__ __forceinline__ float gmm_terms(const float* __restrict__ p, int K, float zv, float* a, float* scale,
                                           float* lsm) {
  float mx = -INFINITY;
  for (int k = 0; k < K; ++k) mx = fmaxf(mx, p[k]);
  float se = 0.f;
  for (int k = 0; k < K; ++k) se += expf(p[k] - mx);
  const float lse = mx + logf(se);
  float amx = -INFINITY;
  for (int k = 0; k < K; ++k) {
    const float sc = softplus_f(p[2 * K + k]) + 1e-5f;
    const float u = (zv - p[K + k]) / sc;
    scale[k] = sc;
    lsm[k] = p[k] - lse;
    a[k] = lsm[k] - 0.5f * u * u - logf(sc) - 0.5f * kLog2Pi;
    amx = fmaxf(amx, a[k]);
  }
  float sa = 0.f;
  for (int k = 0; k < K; ++k) sa += expf(a[k] - amx);
  return amx + logf(sa);
}

// out[b] = sum_i log p(z_i | z_<i, context): thread per sample, steps in order (deterministic sum)
__global__ void __launch_bounds__(128) argmm_lp_kernel(const float* __restrict__ head_out, const float* __restrict__ z,
                                                       int64_t B, int d, int K, float* __restrict__ out) {
  const int ld = 3 * K * d;
  for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < B; b += (int64_t)gridDim.x * blockDim.x) {
    float total = 0.f;
    float a[kMaxComp], sc[kMaxComp], lsm[kMaxComp];
    for (int i = 0; i < d; ++i) {
      const float* p = head_out + ((int64_t)i * B + b) * ld + (int64_t)i * 3 * K;
      total += gmm_terms(p, K, z[b * d + i], a, sc, lsm);
    }
    out[b] = total;
  }
}

// d(out[b] * g[b]) / d head_out (only the 3K columns of step i in row i*B + b are non-zero; the caller zeroes
// the buffer) and the direct term d / d z[b,i]
__global__ void __launch_bounds__(128) argmm_lp_bwd_kernel(const float* __restrict__ head_out, const float* __restrict__ z,
                                                           const float* __restrict__ g, int64_t B, int d, int K,
                                                           float* __restrict__ d_head, float* __restrict__ dz_direct) {
  const int ld = 3 * K * d;
  const int64_t n = (int64_t)d * B;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = t / B, b = t - i * B;
    const int64_t off = t * ld + i * 3 * K;
    const float* p = head_out + off;
    float* dp = d_head + off;
    float a[kMaxComp], sc[kMaxComp], lsm[kMaxComp];
    const float zv = z[b * d + i];
    const float lp = gmm_terms(p, K, zv, a, sc, lsm);
    const float gb = g[b];
    float dzv = 0.f;
    for (int k = 0; k < K; ++k) {
      const float resp = expf(a[k] - lp);                 // posterior responsibility of component k
      const float u = (zv - p[K + k]) / sc[k];
      dp[k] = gb * (resp - expf(lsm[k]));
      dp[K + k] = gb * resp * u / sc[k];
      dp[2 * K + k] = gb * resp * (u * u - 1.0f) / sc[k] * sigmoid_f(p[2 * K + k]);
      dzv -= resp * u / sc[k];
    }
    dz_direct[b * d + i] = gb * dzv;
  }
}

// dz[b,j] = dz_direct[b,j] + sum_{i > j} dX[i*B + b, j];   dctx[b,c] = sum_i dX[i*B + b, 2d + c]
__global__ void __launch_bounds__(256) argmm_reduce_dx_kernel(const float* __restrict__ dX, const float* __restrict__ dz_direct,
                                                              int64_t B, int d, int C, float* __restrict__ dz,
                                                              float* __restrict__ dctx) {
  const int F = 2 * d + C;
  const int W = d + C;
  const int64_t n = B * W;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = t / W;
    const int j = (int)(t - b * W);
    if (j < d) {
      if (dz) {
        float acc = dz_direct[b * d + j];
        for (int i = j + 1; i < d; ++i) acc += dX[((int64_t)i * B + b) * F + j];
        dz[b * d + j] = acc;
      }
    } else if (dctx) {
      const int c = j - d;
      float acc = 0.f;
      for (int i = 0; i < d; ++i) acc += dX[((int64_t)i * B + b) * F + 2 * d + c];
      dctx[b * C + c] = acc;
    }
  }
}

int argmm_input(const float* z, const float* ctx, int64_t B, int d, int C, float* X, cudaStream_t s) {
  argmm_input_kernel<<<grid1d((int64_t)d * B * (2 * d + C), 256), 256, 0, s>>>(z, ctx, B, d, C, X);
  PMVAE_LAUNCH_CHECK();
  return 0;
}
int argmm_lp(const float* head_out, const float* z, int64_t B, int d, int K, float* out, cudaStream_t s) {
  PMVAE_CHECK(K >= 1 && K <= kMaxComp, "num_components must be in [1, 32]");
  argmm_lp_kernel<<<grid1d(B, 128), 128, 0, s>>>(head_out, z, B, d, K, out);
  PMVAE_LAUNCH_CHECK();
  return 0;
}
int argmm_lp_bwd(const float* head_out, const float* z, const float* g, int64_t B, int d, int K, float* d_head,
                 float* dz_direct, cudaStream_t s) {
  PMVAE_CHECK(K >= 1 && K <= kMaxComp, "num_components must be in [1, 32]");
  argmm_lp_bwd_kernel<<<grid1d((int64_t)d * B, 128), 128, 0, s>>>(head_out, z, g, B, d, K, d_head, dz_direct);
  PMVAE_LAUNCH_CHECK();
  return 0;
}
// ---- sampling (distributions.py:168-189 `_sample_n`) ------------------------------------------------------------
// input rows of sampling step i for M = n * B rows (row m = s * B + b): [x * (arange(d) < i), (arange(d) < i), context[b]]
__global__ void __launch_bounds__(256) argmm_sample_input_kernel(const float* __restrict__ x, const float* __restrict__ ctx,
                                                                 int64_t M, int64_t B, int d, int C, int step,
                                                                 float* __restrict__ X) {
  const int F = 2 * d + C;
  const int64_t n = M * F;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = t / F;
    const int j = (int)(t - m * F);
    float v;
    if (j < d) v = (j < step) ? x[m * d + j] : 0.f;
    else if (j < 2 * d) v = (j - d < step) ? 1.f : 0.f;
    else v = ctx[(m % B) * C + (j - 2 * d)];
    X[t] = v;
  }
}

// x[m, step] = mean_c + scale_c * eps[s, step, c], c = argmax_k(logits_k + gumbel[s, step, k]) with
// gumbel = -log(-log(u)), u = uniform(minval = tiny, maxval = 1) (jax.random.categorical / gumbel); eps and u are
// [n, d, K] draws shared by every batch row and every step (the reference passes the same key to every step under
// jax.vmap over the batch: distributions.py:168-189).
__global__ void __launch_bounds__(128) argmm_sample_step_kernel(const float* __restrict__ head_out, const float* __restrict__ eps,
                                                                const float* __restrict__ u, int64_t M, int64_t B, int d, int K,
                                                                int step, float* __restrict__ x) {
  const int ld = 3 * K * d;
  for (int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; m < M; m += (int64_t)gridDim.x * blockDim.x) {
    const float* p = head_out + m * ld + (int64_t)step * 3 * K;
    const int64_t s = m / B;
    const float* us = u + (s * d + step) * K;
    int best = 0;
    float bv = -INFINITY;
    for (int k = 0; k < K; ++k) {
      float uk = us[k];
      uk = uk > 0.f ? uk : 1.17549435e-38f;
      const float v = p[k] - logf(-logf(uk));
      if (v > bv) { bv = v; best = k; }
    }
    const float sc = softplus_f(p[2 * K + best]) + 1e-5f;
    x[m * d + step] = p[K + best] + sc * eps[(s * d + step) * K + best];
  }
}

int argmm_sample_input(const float* x, const float* ctx, int64_t M, int64_t B, int d, int C, int step, float* X, cudaStream_t s) {
  if (M == 0) return 0;
  argmm_sample_input_kernel<<<grid1d(M * (2 * d + C), 256), 256, 0, s>>>(x, ctx, M, B, d, C, step, X);
  PMVAE_LAUNCH_CHECK();
  return 0;
}
int argmm_sample_step(const float* head_out, const float* eps, const float* u, int64_t M, int64_t B, int d, int K, int step,
                      float* x, cudaStream_t s) {
  PMVAE_CHECK(K >= 1 && K <= kMaxComp, "num_components must be in [1, 32]");
  if (M == 0) return 0;
  argmm_sample_step_kernel<<<grid1d(M, 128), 128, 0, s>>>(head_out, eps, u, M, B, d, K, step, x);
  PMVAE_LAUNCH_CHECK();
  return 0;
}

int argmm_reduce_dx(const float* dX, const float* dz_direct, int64_t B, int d, int C, float* dz, float* dctx,
                    cudaStream_t s) {
  argmm_reduce_dx_kernel<<<grid1d(B * (d + C), 256), 256, 0, s>>>(dX, dz_direct, B, d, C, dz, dctx);
  PMVAE_LAUNCH_CHECK();
  return 0;
}

}  // namespace pmvae
