// Per-row TriL-Gaussian algebra of the PM-VAE latent, all fp32.
//
// A row is handled by a group of G lanes (G = 16 for d <= 16, so a warp works on two rows;
// G = 32 otherwise, with two elements per lane for d in (32, 64]).  The raw head output
// (P = d + d(d+1)/2 floats) is staged in shared memory with coalesced loads; diagonals
// (softplus + 1e-5), their reciprocals and logs are computed once per row.
//
// Reference sites:
//   distributions.py:101-113  TriLGaussian: loc = p[:d], L = FillScaleTriL(p[d:])
//   vae.py:124                z = mu + L eps                      (latent_fwd)
//   vae.py:130                KL(q(z|x) || N(0,I))                (latent_fwd)
//   vae.py:136-138            log q(stop_grad(z) | x_o)           (match_fwd)
//   vae.py:192-195,203-212    K samples + prior/posterior log-probs (sample_latents)
// Closed forms and backward formulas: SURVEY.md §8a-E/F and Appendix A.3/A.6.
#include <stdlib.h>

#include "kernels.h"

namespace pmvae {

// d = 16 takes the thread-per-row kernels of latent16.cu (PMVAE_LATENT16=0 keeps the group-per-row ones)
static bool fast16() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("PMVAE_LATENT16"); v = e ? atoi(e) : 1; }
  return v != 0;
}
// d = 64: the K-sample draw of the evaluators has its own kernel (PMVAE_LATENT64=0 keeps the general one)
static bool fast64() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("PMVAE_LATENT64"); v = e ? atoi(e) : 1; }
  return v != 0;
}

constexpr int kThreadsL = 128;

// index into the raw head vector of L[i][j] (j <= i): tfp fill_triangular (lower) after the d
// loc entries: c = concat(v[d:], reverse(v)), L[i][j] = c[i*d + j].
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

template <int G>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Work distribution shared by the kernels: every group of a warp runs the same number of
// iterations (shuffles use the full mask); rows past the end are clamped and not stored.
template <int G>
struct RowIter {
  int gl, grp;
  int64_t rows_per_iter, first, iters;
  __device__ RowIter(int64_t B) {
    gl = threadIdx.x % G;
    grp = threadIdx.x / G;
    const int gpb = blockDim.x / G;
    rows_per_iter = (int64_t)gridDim.x * gpb;
    first = (int64_t)blockIdx.x * gpb + grp;
    iters = (B + rows_per_iter - 1) / rows_per_iter;
  }
};

// diagonal terms of one row: sd = softplus(raw)+1e-5, si = 1/sd; returns this lane's sum of log(sd)
template <int G, int E>
__device__ __forceinline__ float stage_diag(const float* sp, float* sd, float* si, int d, int m, int gl) {
  float logd = 0.f;
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int i = gl + G * e;
    if (i < d) {
      const float dg = softplus_f(sp[tril_src(i, i, d, m)]) + 1e-5f;
      sd[i] = dg;
      si[i] = 1.0f / dg;
      logd += logf(dg);
    }
  }
  return logd;
}

// ---------------------------------------------------------------- z, KL
template <int G, int E, int DCT>
__global__ void __launch_bounds__(kThreadsL) latent_fwd_kernel(const float* __restrict__ par,
                                                               const float* __restrict__ eps, float* __restrict__ z,
                                                               float* __restrict__ kl, int64_t B, int d_rt) {
  extern __shared__ float smem[];
  const int d = DCT ? DCT : d_rt;
  const int m = d * (d + 1) / 2, P = d + m;
  RowIter<G> it(B);
  float* sp = smem + (size_t)it.grp * (P + 3 * d);
  float* se = sp + P;
  float* sd = se + d;
  float* si = sd + d;
  const int gl = it.gl;
  for (int64_t t = 0; t < it.iters; ++t) {
    const int64_t r_raw = it.first + t * it.rows_per_iter;
    const bool valid = r_raw < B;
    const int64_t r = valid ? r_raw : B - 1;
    __syncwarp();
    for (int q = gl; q < P; q += G) sp[q] = par[r * P + q];
    for (int q = gl; q < d; q += G) se[q] = eps[r * d + q];
    __syncwarp();
    float logd = stage_diag<G, E>(sp, sd, si, d, m, gl);
    __syncwarp();
    float fro = 0.f, mu2 = 0.f;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int i = gl + G * e;
      if (i < d) {
        const float mu = sp[i];
        float acc = mu;
        for (int j = 0; j < i; ++j) {
          const float l = sp[tril_src(i, j, d, m)];
          acc = fmaf(l, se[j], acc);
          fro = fmaf(l, l, fro);
        }
        const float dg = sd[i];
        acc = fmaf(dg, se[i], acc);
        fro = fmaf(dg, dg, fro);
        mu2 = fmaf(mu, mu, mu2);
        if (valid) z[r * d + i] = acc;
      }
    }
    fro = group_sum<G>(fro); logd = group_sum<G>(logd); mu2 = group_sum<G>(mu2);
    if (valid && gl == 0) kl[r] = -logd + 0.5f * (-(float)d + fro + mu2);
  }
}

// ---------------------------------------------------------------- forward substitution (column oriented)
// On entry sz = z (or z - mu is formed here), si = 1/diag.  Lanes hold s[k] for k = gl + G*e.
// Writes r = L^-1 (z - mu) to sr[] and returns sum r^2 (same value in every lane of the group).
template <int G, int E>
__device__ __forceinline__ float solve_lower(const float* sp, const float* sz, const float* si, float* sr, int d, int m,
                                             int gl) {
  float s[E];
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int k = gl + G * e;
    s[e] = (k < d) ? sz[k] - sp[k] : 0.f;
  }
  float sumsq = 0.f;
  for (int i = 0; i < d; ++i) {
    float src = s[0];
#pragma unroll
    for (int e = 1; e < E; ++e) if (i >= G * e) src = s[e];
    const float ri = __shfl_sync(0xffffffffu, src, i % G, G) * si[i];
    sumsq = fmaf(ri, ri, sumsq);
    if (gl == 0) sr[i] = ri;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int k = gl + G * e;
      if (k > i && k < d) s[e] = fmaf(-sp[tril_src(k, i, d, m)], ri, s[e]);
    }
  }
  __syncwarp();
  return sumsq;
}

template <int G, int E, int DCT>
__global__ void __launch_bounds__(kThreadsL) match_fwd_kernel(const float* __restrict__ par_p,
                                                              const float* __restrict__ z, float* __restrict__ match,
                                                              int64_t B, int d_rt) {
  extern __shared__ float smem[];
  const int d = DCT ? DCT : d_rt;
  const int m = d * (d + 1) / 2, P = d + m;
  RowIter<G> it(B);
  float* sp = smem + (size_t)it.grp * (P + 4 * d);
  float* sz = sp + P;
  float* sr = sz + d;
  float* sd = sr + d;
  float* si = sd + d;
  const int gl = it.gl;
  for (int64_t t = 0; t < it.iters; ++t) {
    const int64_t r_raw = it.first + t * it.rows_per_iter;
    const bool valid = r_raw < B;
    const int64_t r = valid ? r_raw : B - 1;
    __syncwarp();
    for (int q = gl; q < P; q += G) sp[q] = par_p[r * P + q];
    for (int q = gl; q < d; q += G) sz[q] = z[r * d + q];
    __syncwarp();
    float logd = stage_diag<G, E>(sp, sd, si, d, m, gl);
    __syncwarp();
    const float sumsq = solve_lower<G, E>(sp, sz, si, sr, d, m, gl);
    logd = group_sum<G>(logd);
    if (valid && gl == 0) match[r] = -0.5f * sumsq - logd - 0.5f * (float)d * kLog2Pi;
  }
}

// ---------------------------------------------------------------- backward of both heads
template <int G, int E, int DCT>
__global__ void __launch_bounds__(kThreadsL) latent_bwd_kernel(
    const float* __restrict__ par_e, const float* __restrict__ par_p, const float* __restrict__ eps,
    const float* __restrict__ z, const float* __restrict__ dz_dec, const float* __restrict__ g_kl,
    const float* __restrict__ g_match, int stop_grad, float* __restrict__ dpar_e, float* __restrict__ dpar_p,
    __nv_bfloat16* __restrict__ dpar_e_b, __nv_bfloat16* __restrict__ dpar_p_b, int64_t B, int d_rt) {
  extern __shared__ float smem[];
  const int d = DCT ? DCT : d_rt;
  const int m = d * (d + 1) / 2, P = d + m;
  RowIter<G> it(B);
  float* sp = smem + (size_t)it.grp * (P + 6 * d);
  float* sz = sp + P;      // z, later dz_total
  float* sr = sz + d;      // r = L_p^-1 (z - mu_p)
  float* sg = sr + d;      // g = L_p^-T r
  float* se = sg + d;      // eps
  float* sd = se + d;      // diag
  float* si = sd + d;      // 1 / diag
  const int gl = it.gl;
  for (int64_t t = 0; t < it.iters; ++t) {
    const int64_t r_raw = it.first + t * it.rows_per_iter;
    const bool valid = r_raw < B;
    const int64_t r = valid ? r_raw : B - 1;
    __syncwarp();
    // ---- partial posterior: d match / d par_p
    for (int q = gl; q < P; q += G) sp[q] = par_p[r * P + q];
    for (int q = gl; q < d; q += G) { sz[q] = z[r * d + q]; se[q] = eps[r * d + q]; }
    __syncwarp();
    stage_diag<G, E>(sp, sd, si, d, m, gl);
    __syncwarp();
    solve_lower<G, E>(sp, sz, si, sr, d, m, gl);
    // backward substitution g = L^-T r: t = r; for i = d-1..0: g_i = t_i / L_ii; t_j -= L_ij g_i (j < i)
    {
      float tt[E];
#pragma unroll
      for (int e = 0; e < E; ++e) { const int k = gl + G * e; tt[e] = (k < d) ? sr[k] : 0.f; }
      for (int i = d - 1; i >= 0; --i) {
        float src = tt[0];
#pragma unroll
        for (int e = 1; e < E; ++e) if (i >= G * e) src = tt[e];
        const float gi = __shfl_sync(0xffffffffu, src, i % G, G) * si[i];
        if (gl == 0) sg[i] = gi;
#pragma unroll
        for (int e = 0; e < E; ++e) {
          const int k = gl + G * e;
          if (k < i) tt[e] = fmaf(-sp[tril_src(i, k, d, m)], gi, tt[e]);
        }
      }
      __syncwarp();
    }
    const float mw = g_match[r];
    for (int q = gl; q < P; q += G) {
      float val;
      if (q < d) {
        val = mw * sg[q];
      } else {
        int i, j;
        tril_ij(q - d, d, m, i, j);
        val = sg[i] * sr[j];
        if (i == j) { val -= si[i]; val *= sigmoid_f(sp[q]); }
        val *= mw;
      }
      if (valid) {
        if (dpar_p) dpar_p[r * P + q] = val;
        if (dpar_p_b) dpar_p_b[r * P + q] = __float2bfloat16(val);
      }
    }
    __syncwarp();
    // ---- dz_total = dz_dec - (stop_grad ? 0 : mw * g)
    for (int q = gl; q < d; q += G) {
      float v = dz_dec ? dz_dec[r * d + q] : 0.f;
      if (!stop_grad) v -= mw * sg[q];
      sz[q] = v;
    }
    __syncwarp();
    // ---- posterior: z = mu + L eps and kw * KL
    for (int q = gl; q < P; q += G) sp[q] = par_e[r * P + q];
    __syncwarp();
    const float kw = g_kl[r];
    for (int q = gl; q < P; q += G) {
      float val;
      if (q < d) {
        val = sz[q] + kw * sp[q];
      } else {
        int i, j;
        tril_ij(q - d, d, m, i, j);
        const float raw = sp[q];
        if (i == j) {
          const float dg = softplus_f(raw) + 1e-5f;
          val = (sz[i] * se[j] + kw * (dg - 1.0f / dg)) * sigmoid_f(raw);
        } else {
          val = sz[i] * se[j] + kw * raw;
        }
      }
      if (valid) {
        if (dpar_e) dpar_e[r * P + q] = val;
        if (dpar_e_b) dpar_e_b[r * P + q] = __float2bfloat16(val);
      }
    }
  }
}

// ---------------------------------------------------------------- K importance samples per row
template <int G, int E, int DCT>
__global__ void __launch_bounds__(kThreadsL) sample_latents_kernel(const float* __restrict__ par, Key2 key, int64_t B,
                                                                   int64_t K, int64_t B_total, int64_t row_start,
                                                                   int d_rt, float* __restrict__ z,
                                                                   float* __restrict__ base) {
  extern __shared__ float smem[];
  const int d = DCT ? DCT : d_rt;
  const int m = d * (d + 1) / 2, P = d + m;
  constexpr int kChunk = 16;                       // samples per work item
  const int64_t chunks = (K + kChunk - 1) / kChunk;
  RowIter<G> it(B * chunks);
  float* sp = smem + (size_t)it.grp * (P + 3 * d);
  float* se = sp + P;
  float* sd = se + d;
  float* si = sd + d;
  const int gl = it.gl;
  const uint64_t n_total = (uint64_t)K * (uint64_t)B_total * (uint64_t)d;
  for (int64_t t = 0; t < it.iters; ++t) {
    const int64_t item_raw = it.first + t * it.rows_per_iter;
    const bool valid = item_raw < B * chunks;
    const int64_t item = valid ? item_raw : B * chunks - 1;
    const int64_t r = item / chunks, ck = item - r * chunks;
    __syncwarp();
    for (int q = gl; q < P; q += G) sp[q] = par[r * P + q];
    __syncwarp();
    float logd = stage_diag<G, E>(sp, sd, si, d, m, gl);
    logd = group_sum<G>(logd);
    const int64_t kend = min(K, (ck + 1) * kChunk);
    for (int64_t k = ck * kChunk; k < kend; ++k) {
      __syncwarp();
      float e2 = 0.f;
      for (int q = gl; q < d; q += G) {
        const uint64_t idx = ((uint64_t)k * B_total + (uint64_t)(row_start + r)) * d + q;
        const float ev = bits_to_normal(jax_random_word(key, n_total, idx));
        se[q] = ev;
        e2 = fmaf(ev, ev, e2);
      }
      __syncwarp();
      float z2 = 0.f;
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const int i = gl + G * e;
        if (i < d) {
          float acc = sp[i];
          for (int j = 0; j < i; ++j) acc = fmaf(sp[tril_src(i, j, d, m)], se[j], acc);
          acc = fmaf(sd[i], se[i], acc);
          if (valid) z[(k * B + r) * d + i] = acc;
          z2 = fmaf(acc, acc, z2);
        }
      }
      e2 = group_sum<G>(e2); z2 = group_sum<G>(z2);
      if (valid && gl == 0) base[k * B + r] = -0.5f * z2 + 0.5f * e2 + logd;
    }
  }
}

// d = 64 (bsds): one warp per (row, 64 samples).  The triangular factor is expanded ONCE per item into a dense,
// zero-padded, transposed matrix Lt[j][i] in shared memory (the fill_triangular index map is paid 4096 times per item
// instead of 2080 times per sample), then the samples go through in batches of 16 as a small matrix product:
// z[s][i] = mu[i] + sum_j Lt[j][i] e[j][s], lane = rows i and i + 32, 32 accumulators per lane, e[j][0..15] read as four
// broadcast 16-byte loads.  Same words of the same threefry stream as the general kernel (index (k, row, q)).
constexpr int kS64Batch = 16, kS64Chunk = 64;
constexpr int kS64SmemFloats = 64 * 64 + 64 * kS64Batch + 64;      // Lt | e | mu per warp
__global__ void __launch_bounds__(kThreadsL) sample_latents64_kernel(const float* __restrict__ par, Key2 key, int64_t B,
                                                                     int64_t K, int64_t B_total, int64_t row_start,
                                                                     float* __restrict__ z, float* __restrict__ base) {
  extern __shared__ __align__(16) float smem[];
  constexpr int d = 64, m = d * (d + 1) / 2, P = d + m;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  float* Lt = smem + (size_t)wib * kS64SmemFloats;
  float* se = Lt + 64 * 64;                         // [j][16 samples]
  float* smu = se + 64 * kS64Batch;
  const int64_t chunks = (K + kS64Chunk - 1) / kS64Chunk;
  const int64_t items = B * chunks;
  const uint64_t n_total = (uint64_t)K * (uint64_t)B_total * (uint64_t)d;
  for (int64_t item = (int64_t)blockIdx.x * wpb + wib; item < items; item += (int64_t)gridDim.x * wpb) {
    const int64_t r = item / chunks, ck = item - r * chunks;
    const float* pr = par + r * P;
    __syncwarp();
    // dense transposed factor: lane owns columns i = lane and lane + 32 of every row j
    float logd = 0.f;
    for (int j = 0; j < 64; ++j) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int i = lane + 32 * h;
        float v = 0.f;
        if (i >= j) {
          v = __ldg(pr + tril_src(i, j, d, m));
          if (i == j) { v = softplus_f(v) + 1e-5f; logd += logf(v); }
        }
        Lt[j * 64 + i] = v;
      }
    }
    smu[lane] = __ldg(pr + lane); smu[lane + 32] = __ldg(pr + lane + 32);
    logd = group_sum<32>(logd);
    __syncwarp();
    const float mu0 = smu[lane], mu1 = smu[lane + 32];
    const int64_t kend = min(K, (ck + 1) * kS64Chunk);
    for (int64_t k0 = ck * kS64Chunk; k0 < kend; k0 += kS64Batch) {
      const int ns = (int)min((int64_t)kS64Batch, kend - k0);
      __syncwarp();
      // eps: lane draws elements q = lane and lane + 32 of every sample of the batch
      float e2[kS64Batch];
#pragma unroll
      for (int sI = 0; sI < kS64Batch; ++sI) {
        float a = 0.f, b = 0.f;
        if (sI < ns) {
          const uint64_t idx = ((uint64_t)(k0 + sI) * B_total + (uint64_t)(row_start + r)) * d;
          a = bits_to_normal(jax_random_word(key, n_total, idx + lane));
          b = bits_to_normal(jax_random_word(key, n_total, idx + lane + 32));
        }
        se[lane * kS64Batch + sI] = a;
        se[(lane + 32) * kS64Batch + sI] = b;
        e2[sI] = fmaf(a, a, b * b);
      }
      __syncwarp();
      float acc0[kS64Batch], acc1[kS64Batch];
#pragma unroll
      for (int sI = 0; sI < kS64Batch; ++sI) { acc0[sI] = mu0; acc1[sI] = mu1; }
      // rows i < 32 only see columns j < 32 (the rest of the factor is zero)
      for (int j = 0; j < 64; ++j) {
        const float l1 = Lt[j * 64 + lane + 32];
        const float4* ev = reinterpret_cast<const float4*>(se + j * kS64Batch);
        const float4 e0 = ev[0], e1 = ev[1], e2v = ev[2], e3 = ev[3];
        const float ee[kS64Batch] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w, e2v.x, e2v.y, e2v.z, e2v.w, e3.x, e3.y, e3.z, e3.w};
#pragma unroll
        for (int sI = 0; sI < kS64Batch; ++sI) acc1[sI] = fmaf(l1, ee[sI], acc1[sI]);
        if (j < 32) {
          const float l0 = Lt[j * 64 + lane];
#pragma unroll
          for (int sI = 0; sI < kS64Batch; ++sI) acc0[sI] = fmaf(l0, ee[sI], acc0[sI]);
        }
      }
#pragma unroll
      for (int sI = 0; sI < kS64Batch; ++sI) {
        if (sI < ns) {                                   // warp-uniform
          const int64_t k = k0 + sI;
          z[(k * B + r) * d + lane] = acc0[sI];
          z[(k * B + r) * d + lane + 32] = acc1[sI];
          const float z2 = group_sum<32>(fmaf(acc0[sI], acc0[sI], acc1[sI] * acc1[sI]));
          const float ee2 = group_sum<32>(e2[sI]);
          if (lane == 0) base[k * B + r] = -0.5f * z2 + 0.5f * ee2 + logd;
        }
      }
    }
  }
}

// ---------------------------------------------------------------- stand-alone VJP of (z, kl) = f(par, eps)
// d/dpar of  sum_r [ dz[r] . z[r] + g_kl[r] * kl[r] ]  with z = mu + L eps, kl = KL(N(mu, LL^T) || N(0, I)):
//   loc:  dz + g_kl * mu;   off-diagonal L_ij: dz_i eps_j + g_kl * L_ij;
//   diagonal raw_ii: (dz_i eps_i + g_kl * (D_ii - 1 / D_ii)) * sigmoid(raw_ii),  D_ii = softplus(raw_ii) + 1e-5
__global__ void __launch_bounds__(256) tril_sample_bwd_kernel(const float* __restrict__ par, const float* __restrict__ eps,
                                                              const float* __restrict__ dz, const float* __restrict__ g_kl,
                                                              float* __restrict__ dpar, int64_t B, int d) {
  const int m = d * (d + 1) / 2, P = d + m;
  const int64_t n = B * P;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = t / P;
    const int q = (int)(t - r * P);
    const float kw = g_kl[r];
    const float raw = par[t];
    float val;
    if (q < d) {
      val = dz[r * d + q] + kw * raw;
    } else {
      int i, j;
      tril_ij(q - d, d, m, i, j);
      const float base = dz[r * d + i] * eps[r * d + j];
      if (i == j) {
        const float dg = softplus_f(raw) + 1e-5f;
        val = (base + kw * (dg - 1.0f / dg)) * sigmoid_f(raw);
      } else {
        val = base + kw * raw;
      }
    }
    dpar[t] = val;
  }
}

int tril_sample_bwd(const float* par, const float* eps, const float* dz, const float* g_kl, float* dpar, int64_t B, int d,
                    cudaStream_t s) {
  PMVAE_CHECK(d >= 1 && d <= 64, "latent_dim must be in [1, 64]");
  if (B == 0) return 0;
  const int P = d + d * (d + 1) / 2;
  int64_t g = (B * P + 255) / 256;
  if (g > 148 * 8) g = 148 * 8;
  tril_sample_bwd_kernel<<<(int)g, 256, 0, s>>>(par, eps, dz, g_kl, dpar, B, d);
  PMVAE_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------- launchers
static int grid_for_rows(int64_t rows, int G) {
  const int gpb = kThreadsL / G;
  int64_t g = (rows + gpb - 1) / gpb;
  if (g > 148 * 8) g = 148 * 8;
  if (g < 1) g = 1;
  return (int)g;
}

#define DISPATCH_D(KERNEL, d, rows, smem_per_group, stream, ...)                                         \
  do {                                                                                                   \
    if ((d) == 16) {                                                                                     \
      KERNEL<16, 1, 16><<<grid_for_rows(rows, 16), kThreadsL, (kThreadsL / 16) * (smem_per_group), stream>>>(__VA_ARGS__); \
    } else if ((d) <= 16) {                                                                              \
      KERNEL<16, 1, 0><<<grid_for_rows(rows, 16), kThreadsL, (kThreadsL / 16) * (smem_per_group), stream>>>(__VA_ARGS__); \
    } else if ((d) <= 32) {                                                                              \
      KERNEL<32, 1, 0><<<grid_for_rows(rows, 32), kThreadsL, (kThreadsL / 32) * (smem_per_group), stream>>>(__VA_ARGS__); \
    } else if ((d) == 64) {                                                                              \
      KERNEL<32, 2, 64><<<grid_for_rows(rows, 32), kThreadsL, (kThreadsL / 32) * (smem_per_group), stream>>>(__VA_ARGS__); \
    } else {                                                                                             \
      KERNEL<32, 2, 0><<<grid_for_rows(rows, 32), kThreadsL, (kThreadsL / 32) * (smem_per_group), stream>>>(__VA_ARGS__); \
    }                                                                                                    \
  } while (0)

int latent_fwd(const float* par, const float* eps, float* z, float* kl, int64_t B, int d, cudaStream_t s) {
  PMVAE_CHECK(d >= 1 && d <= 64, "latent_dim must be in [1, 64]");
  if (B == 0) return 0;
  if (d == 16 && fast16()) return latent_fwd16(par, eps, z, kl, B, s);
  if (d == 64 && fast64()) return latent_fwd64(par, eps, z, kl, B, s);
  const int P = d + d * (d + 1) / 2;
  const size_t spg = (size_t)(P + 3 * d) * sizeof(float);
  DISPATCH_D(latent_fwd_kernel, d, B, spg, s, par, eps, z, kl, B, d);
  PMVAE_LAUNCH_CHECK();
  return 0;
}

bool match_fwd_saves(int d) { return d == 64 && fast64(); }

int match_fwd(const float* par_p, const float* z, float* match, int64_t B, int d, cudaStream_t s, float* saved,
              int64_t saved_stride) {
  PMVAE_CHECK(d >= 1 && d <= 64, "latent_dim must be in [1, 64]");
  if (B == 0) return 0;
  if (d == 16 && fast16()) return match_fwd16(par_p, z, match, B, s);
  if (d == 64 && fast64())
    return saved ? match_fwd64(par_p, z, match, B, s, saved, saved + saved_stride, saved + 2 * saved_stride)
                 : match_fwd64(par_p, z, match, B, s);
  const int P = d + d * (d + 1) / 2;
  const size_t spg = (size_t)(P + 4 * d) * sizeof(float);
  DISPATCH_D(match_fwd_kernel, d, B, spg, s, par_p, z, match, B, d);
  PMVAE_LAUNCH_CHECK();
  return 0;
}

// whether latent_bwd with bf16 outputs and db_e / db_p also accumulates the two head bias gradients (d = 16 kernels)
bool latent_bwd_bias_fused(int d) { return (d == 16 && fast16()) || (d == 64 && fast64()); }

int latent_bwd(const float* par_e, const float* par_p, const float* eps, const float* z, const float* dz_dec,
               const float* g_kl, const float* g_match, int stop_grad, float* dpar_e, float* dpar_p,
               __nv_bfloat16* dpar_e_b, __nv_bfloat16* dpar_p_b, int64_t B, int d, cudaStream_t s, float* db_e,
               float* db_p, bool* db_done, const float* saved, int64_t saved_stride) {
  if (db_done) *db_done = false;
  PMVAE_CHECK(d >= 1 && d <= 64, "latent_dim must be in [1, 64]");
  if (B == 0) return 0;
  if (d == 16 && fast16() && dpar_e == nullptr && dpar_p == nullptr && dpar_e_b && dpar_p_b) {
    if (db_done) *db_done = (db_e != nullptr && db_p != nullptr);     // head bias gradients are taken here as well
    return latent_bwd16(par_e, par_p, eps, z, dz_dec, g_kl, g_match, stop_grad, dpar_e_b, dpar_p_b, db_e, db_p, B, s);
  }
  if (d == 64 && fast64() && dpar_e == nullptr && dpar_p == nullptr && dpar_e_b && dpar_p_b && saved) {
    if (db_done) *db_done = (db_e != nullptr && db_p != nullptr);
    return latent_bwd64(par_e, eps, dz_dec, g_kl, g_match, stop_grad, dpar_e_b, dpar_p_b, db_e, db_p, saved,
                        saved + saved_stride, saved + 2 * saved_stride, B, s);
  }
  const int P = d + d * (d + 1) / 2;
  const size_t spg = (size_t)(P + 6 * d) * sizeof(float);
  DISPATCH_D(latent_bwd_kernel, d, B, spg, s, par_e, par_p, eps, z, dz_dec, g_kl, g_match, stop_grad, dpar_e, dpar_p,
             dpar_e_b, dpar_p_b, B, d);
  PMVAE_LAUNCH_CHECK();
  return 0;
}

int sample_latents(const float* par, Key2 key, int64_t B, int64_t K, int64_t B_total, int64_t row_start, int d,
                   float* z, float* base, cudaStream_t s) {
  PMVAE_CHECK(d >= 1 && d <= 64, "latent_dim must be in [1, 64]");
  PMVAE_CHECK((uint64_t)K * (uint64_t)B_total * (uint64_t)d <= 0xFFFFFFFFull,
              "K*B_total*d exceeds one 2^32-1 element draw");
  if (B * K == 0) return 0;
  if (d == 16 && fast16()) return sample_latents16(par, key, B, K, B_total, row_start, z, base, s);
  if (d == 64 && fast64()) {
    const size_t smem = (size_t)(kThreadsL / 32) * kS64SmemFloats * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
      PMVAE_CUDA(cudaFuncSetAttribute(sample_latents64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_set = true;
    }
    const int64_t items = B * ((K + kS64Chunk - 1) / kS64Chunk);
    int64_t g = (items + kThreadsL / 32 - 1) / (kThreadsL / 32);
    if (g > 148 * 2) g = 148 * 2;
    if (g < 1) g = 1;
    sample_latents64_kernel<<<(int)g, kThreadsL, smem, s>>>(par, key, B, K, B_total, row_start, z, base);
    PMVAE_LAUNCH_CHECK();
    return 0;
  }
  const int P = d + d * (d + 1) / 2;
  const size_t spg = (size_t)(P + 3 * d) * sizeof(float);
  const int64_t items = B * ((K + 15) / 16);
  DISPATCH_D(sample_latents_kernel, d, items, spg, s, par, key, B, K, B_total, row_start, d, z, base);
  PMVAE_LAUNCH_CHECK();
  return 0;
}

}  // namespace pmvae
