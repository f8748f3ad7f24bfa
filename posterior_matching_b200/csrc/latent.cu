// Per-row TriL-Gaussian algebra of the PM-VAE latent: one warp per row, the raw head
// output (P = d + d(d+1)/2 floats) staged in shared memory, all fp32.
//
// Reference sites:
//   distributions.py:101-113  TriLGaussian: loc = p[:d], L = FillScaleTriL(p[d:])
//   vae.py:124                z = mu + L eps                      (latent_fwd)
//   vae.py:130                KL(q(z|x) || N(0,I))                (latent_fwd)
//   vae.py:136-138            log q(stop_grad(z) | x_o)           (match_fwd)
//   vae.py:192-195,203-212    K samples + prior/posterior log-probs (sample_latents)
// Closed forms and backward formulas: SURVEY.md §8a-E/F and Appendix A.3/A.6.
#include "kernels.h"

namespace pmvae {

constexpr int kWarpsPerBlock = 4;

// index into the raw head vector of L[i][j] (j <= i): tfp fill_triangular (lower) after
// the d loc entries: c = concat(v[d:], reverse(v)), L[i][j] = c[i*d + j].
__device__ __forceinline__ int tril_src(int i, int j, int d, int m) {
  const int k = i * d + j;
  return d + ((k < m - d) ? (d + k) : (m - 1 - (k - (m - d))));
}
// inverse: position s in v -> (i, j)
__device__ __forceinline__ void tril_ij(int s, int d, int m, int& i, int& j) {
  if (s >= d) {
    const int k = s - d;
    i = k / d; j = k - i * d;
    if (j <= i && k < m - d) return;
  }
  const int k = (m - d) + (m - 1 - s);
  i = k / d; j = k - i * d;
}
__device__ __forceinline__ float tril_diag(const float* sp, int i, int d, int m) {
  return softplus_f(sp[tril_src(i, i, d, m)]) + 1e-5f;
}

__device__ __forceinline__ void stage_row(float* sp, const float* __restrict__ src, int P, int lane) {
  for (int q = lane; q < P; q += 32) sp[q] = src[q];
  __syncwarp();
}

// ---------------------------------------------------------------- z, KL
__global__ void __launch_bounds__(32 * kWarpsPerBlock) latent_fwd_kernel(const float* __restrict__ par,
                                                                         const float* __restrict__ eps,
                                                                         float* __restrict__ z, float* __restrict__ kl,
                                                                         int64_t B, int d) {
  extern __shared__ float smem[];
  const int P = d + d * (d + 1) / 2, m = d * (d + 1) / 2;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float* sp = smem + (size_t)w * (P + d);
  float* se = sp + P;
  for (int64_t r = (int64_t)blockIdx.x * kWarpsPerBlock + w; r < B; r += (int64_t)gridDim.x * kWarpsPerBlock) {
    __syncwarp();
    stage_row(sp, par + r * P, P, lane);
    for (int q = lane; q < d; q += 32) se[q] = eps[r * d + q];
    __syncwarp();
    float fro = 0.f, logd = 0.f, mu2 = 0.f;
    for (int i = lane; i < d; i += 32) {
      const float mu = sp[i];
      float acc = mu;
      for (int j = 0; j < i; ++j) {
        const float l = sp[tril_src(i, j, d, m)];
        acc = fmaf(l, se[j], acc);
        fro = fmaf(l, l, fro);
      }
      const float dg = tril_diag(sp, i, d, m);
      acc = fmaf(dg, se[i], acc);
      fro = fmaf(dg, dg, fro);
      logd += logf(dg);
      mu2 = fmaf(mu, mu, mu2);
      z[r * d + i] = acc;
    }
    fro = warp_sum(fro); logd = warp_sum(logd); mu2 = warp_sum(mu2);
    if (lane == 0) kl[r] = -logd + 0.5f * (-(float)d + fro + mu2);
  }
}

int latent_fwd(const float* par, const float* eps, float* z, float* kl, int64_t B, int d, cudaStream_t s) {
  PMVAE_CHECK(d >= 1 && d <= 64, "latent_dim must be in [1, 64]");
  if (B == 0) return 0;
  const int P = d + d * (d + 1) / 2;
  const size_t smem = (size_t)kWarpsPerBlock * (P + d) * sizeof(float);
  int64_t g = ceil_div(B, kWarpsPerBlock);
  if (g > 148 * 16) g = 148 * 16;
  latent_fwd_kernel<<<(int)g, 32 * kWarpsPerBlock, smem, s>>>(par, eps, z, kl, B, d);
  PMVAE_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------- forward substitution (column oriented)
// lanes hold s[k] for k = lane (s0) and k = lane + 32 (s1); returns r in sr[] (smem),
// and sum r^2 / sum log L_ii in all lanes.
__device__ __forceinline__ void solve_lower(const float* sp, float* sr, const float* sz, int d, int m, int lane,
                                            float& sumsq, float& logd) {
  float s0 = (lane < d) ? sz[lane] - sp[lane] : 0.f;
  float s1 = (lane + 32 < d) ? sz[lane + 32] - sp[lane + 32] : 0.f;
  sumsq = 0.f; logd = 0.f;
  for (int i = 0; i < d; ++i) {
    const float lii = tril_diag(sp, i, d, m);
    const float si = __shfl_sync(0xffffffffu, (i < 32) ? s0 : s1, i & 31);
    const float ri = si / lii;
    sumsq = fmaf(ri, ri, sumsq);
    logd += logf(lii);
    if (lane == 0) sr[i] = ri;
    if (lane > i && lane < d) s0 = fmaf(-sp[tril_src(lane, i, d, m)], ri, s0);
    if (lane + 32 > i && lane + 32 < d) s1 = fmaf(-sp[tril_src(lane + 32, i, d, m)], ri, s1);
  }
  __syncwarp();
}

__global__ void __launch_bounds__(32 * kWarpsPerBlock) match_fwd_kernel(const float* __restrict__ par_p,
                                                                        const float* __restrict__ z,
                                                                        float* __restrict__ match, int64_t B, int d) {
  extern __shared__ float smem[];
  const int P = d + d * (d + 1) / 2, m = d * (d + 1) / 2;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float* sp = smem + (size_t)w * (P + 2 * d);
  float* sz = sp + P;
  float* sr = sz + d;
  for (int64_t r = (int64_t)blockIdx.x * kWarpsPerBlock + w; r < B; r += (int64_t)gridDim.x * kWarpsPerBlock) {
    __syncwarp();
    stage_row(sp, par_p + r * P, P, lane);
    for (int q = lane; q < d; q += 32) sz[q] = z[r * d + q];
    __syncwarp();
    float sumsq, logd;
    solve_lower(sp, sr, sz, d, m, lane, sumsq, logd);
    if (lane == 0) match[r] = -0.5f * sumsq - logd - 0.5f * (float)d * kLog2Pi;
  }
}

int match_fwd(const float* par_p, const float* z, float* match, int64_t B, int d, cudaStream_t s) {
  PMVAE_CHECK(d >= 1 && d <= 64, "latent_dim must be in [1, 64]");
  if (B == 0) return 0;
  const int P = d + d * (d + 1) / 2;
  const size_t smem = (size_t)kWarpsPerBlock * (P + 2 * d) * sizeof(float);
  int64_t g = ceil_div(B, kWarpsPerBlock);
  if (g > 148 * 16) g = 148 * 16;
  match_fwd_kernel<<<(int)g, 32 * kWarpsPerBlock, smem, s>>>(par_p, z, match, B, d);
  PMVAE_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------- backward of both heads
__global__ void __launch_bounds__(32 * kWarpsPerBlock) latent_bwd_kernel(
    const float* __restrict__ par_e, const float* __restrict__ par_p, const float* __restrict__ eps,
    const float* __restrict__ z, const float* __restrict__ dz_dec, const float* __restrict__ g_kl,
    const float* __restrict__ g_match, int stop_grad, float* __restrict__ dpar_e, float* __restrict__ dpar_p, int64_t B,
    int d) {
  extern __shared__ float smem[];
  const int P = d + d * (d + 1) / 2, m = d * (d + 1) / 2;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float* sp = smem + (size_t)w * (P + 4 * d);
  float* sz = sp + P;      // z, later dz_total
  float* sr = sz + d;      // r = L_p^-1 (z - mu_p)
  float* sg = sr + d;      // g = L_p^-T r
  float* se = sg + d;      // eps
  for (int64_t r = (int64_t)blockIdx.x * kWarpsPerBlock + w; r < B; r += (int64_t)gridDim.x * kWarpsPerBlock) {
    __syncwarp();
    // ---- partial posterior: d match / d par_p
    stage_row(sp, par_p + r * P, P, lane);
    for (int q = lane; q < d; q += 32) { sz[q] = z[r * d + q]; se[q] = eps[r * d + q]; }
    __syncwarp();
    float sumsq, logd;
    solve_lower(sp, sr, sz, d, m, lane, sumsq, logd);
    // backward substitution g = L^-T r: t_j = r_j; for i = d-1..0: g_i = t_i / L_ii; t_j -= L_ij g_i (j < i)
    {
      float t0 = (lane < d) ? sr[lane] : 0.f;
      float t1 = (lane + 32 < d) ? sr[lane + 32] : 0.f;
      for (int i = d - 1; i >= 0; --i) {
        const float lii = tril_diag(sp, i, d, m);
        const float ti = __shfl_sync(0xffffffffu, (i < 32) ? t0 : t1, i & 31);
        const float gi = ti / lii;
        if (lane == 0) sg[i] = gi;
        if (lane < i) t0 = fmaf(-sp[tril_src(i, lane, d, m)], gi, t0);
        if (lane + 32 < i) t1 = fmaf(-sp[tril_src(i, lane + 32, d, m)], gi, t1);
      }
      __syncwarp();
    }
    const float mw = g_match[r];
    for (int q = lane; q < P; q += 32) {
      float val;
      if (q < d) {
        val = mw * sg[q];
      } else {
        int i, j;
        tril_ij(q - d, d, m, i, j);
        val = sg[i] * sr[j];
        if (i == j) {
          const float raw = sp[q];
          val -= 1.0f / (softplus_f(raw) + 1e-5f);
          val *= sigmoid_f(raw);
        }
        val *= mw;
      }
      dpar_p[r * P + q] = val;
    }
    __syncwarp();
    // ---- dz_total = dz_dec - (stop_grad ? 0 : mw * g)
    for (int q = lane; q < d; q += 32) {
      float v = dz_dec ? dz_dec[r * d + q] : 0.f;
      if (!stop_grad) v -= mw * sg[q];
      sz[q] = v;
    }
    __syncwarp();
    // ---- posterior: z = mu + L eps and kw * KL
    stage_row(sp, par_e + r * P, P, lane);
    const float kw = g_kl[r];
    for (int q = lane; q < P; q += 32) {
      float val;
      if (q < d) {
        val = sz[q] + kw * sp[q];
      } else {
        int i, j;
        tril_ij(q - d, d, m, i, j);
        const float raw = sp[q];
        if (i == j) {
          const float dg = softplus_f(raw) + 1e-5f;
          val = sz[i] * se[j] + kw * (dg - 1.0f / dg);
          val *= sigmoid_f(raw);
        } else {
          val = sz[i] * se[j] + kw * raw;
        }
      }
      dpar_e[r * P + q] = val;
    }
  }
}

int latent_bwd(const float* par_e, const float* par_p, const float* eps, const float* z, const float* dz_dec,
               const float* g_kl, const float* g_match, int stop_grad, float* dpar_e, float* dpar_p, int64_t B, int d,
               cudaStream_t s) {
  PMVAE_CHECK(d >= 1 && d <= 64, "latent_dim must be in [1, 64]");
  if (B == 0) return 0;
  const int P = d + d * (d + 1) / 2;
  const size_t smem = (size_t)kWarpsPerBlock * (P + 4 * d) * sizeof(float);
  int64_t g = ceil_div(B, kWarpsPerBlock);
  if (g > 148 * 16) g = 148 * 16;
  latent_bwd_kernel<<<(int)g, 32 * kWarpsPerBlock, smem, s>>>(par_e, par_p, eps, z, dz_dec, g_kl, g_match, stop_grad,
                                                             dpar_e, dpar_p, B, d);
  PMVAE_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------- K importance samples per row
__global__ void __launch_bounds__(32 * kWarpsPerBlock) sample_latents_kernel(const float* __restrict__ par, Key2 key,
                                                                             int64_t B, int64_t K, int64_t B_total,
                                                                             int64_t row_start, int d,
                                                                             float* __restrict__ z,
                                                                             float* __restrict__ base) {
  extern __shared__ float smem[];
  const int P = d + d * (d + 1) / 2, m = d * (d + 1) / 2;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float* sp = smem + (size_t)w * (P + d);
  float* se = sp + P;
  const uint64_t n_total = (uint64_t)K * (uint64_t)B_total * (uint64_t)d;
  // work item = (row r, sample chunk): chunks of 32 samples keep small batches parallel
  const int64_t chunks = ceil_div(K, 32);
  for (int64_t item = (int64_t)blockIdx.x * kWarpsPerBlock + w; item < B * chunks;
       item += (int64_t)gridDim.x * kWarpsPerBlock) {
    const int64_t r = item / chunks, ck = item - r * chunks;
    __syncwarp();
    stage_row(sp, par + r * P, P, lane);
    float logd = 0.f;
    for (int i = lane; i < d; i += 32) logd += logf(tril_diag(sp, i, d, m));
    logd = warp_sum(logd);
    const int64_t kend = min(K, (ck + 1) * 32);
    for (int64_t k = ck * 32; k < kend; ++k) {
      __syncwarp();
      float e2 = 0.f;
      for (int q = lane; q < d; q += 32) {
        const uint64_t idx = ((uint64_t)k * B_total + (uint64_t)(row_start + r)) * d + q;
        const float e = bits_to_normal(jax_random_word(key, n_total, idx));
        se[q] = e;
        e2 = fmaf(e, e, e2);
      }
      __syncwarp();
      float z2 = 0.f;
      for (int i = lane; i < d; i += 32) {
        float acc = sp[i];
        for (int j = 0; j < i; ++j) acc = fmaf(sp[tril_src(i, j, d, m)], se[j], acc);
        acc = fmaf(tril_diag(sp, i, d, m), se[i], acc);
        z[(k * B + r) * d + i] = acc;
        z2 = fmaf(acc, acc, z2);
      }
      e2 = warp_sum(e2); z2 = warp_sum(z2);
      if (lane == 0) base[k * B + r] = -0.5f * z2 + 0.5f * e2 + logd;
    }
  }
}

int sample_latents(const float* par, Key2 key, int64_t B, int64_t K, int64_t B_total, int64_t row_start, int d,
                   float* z, float* base, cudaStream_t s) {
  PMVAE_CHECK(d >= 1 && d <= 64, "latent_dim must be in [1, 64]");
  PMVAE_CHECK((uint64_t)K * (uint64_t)B_total * (uint64_t)d <= 0xFFFFFFFFull,
              "K*B_total*d exceeds one 2^32-1 element draw");
  if (B * K == 0) return 0;
  const int P = d + d * (d + 1) / 2;
  const size_t smem = (size_t)kWarpsPerBlock * (P + d) * sizeof(float);
  int64_t g = ceil_div(B * ceil_div(K, 32), kWarpsPerBlock);
  if (g > 148 * 16) g = 148 * 16;
  sample_latents_kernel<<<(int)g, 32 * kWarpsPerBlock, smem, s>>>(par, key, B, K, B_total, row_start, d, z, base);
  PMVAE_LAUNCH_CHECK();
  return 0;
}

}  // namespace pmvae
