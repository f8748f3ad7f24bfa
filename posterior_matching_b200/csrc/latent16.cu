// Thread-per-row TriL-Gaussian algebra for latent_dim = 16 (gas / power / hepmass): the whole raw head
// vector of a row (P = 152 floats: loc | FillScaleTriL input) lives in registers, every index of the
// fill_triangular permutation is a compile-time constant, and rows move between HBM and registers through
// a per-warp shared-memory tile so that global traffic is fully coalesced (a warp's 32 rows are one
// contiguous 19 KB span).  Same formulas as latent.cu.
//
// Reference sites: distributions.py:101-113 (TriLGaussian), vae.py:124 (z = mu + L eps), vae.py:130 (KL),
// vae.py:136-138 (log q(z | x_o)).
#include <cuda_bf16.h>

#include "kernels.h"

namespace pmvae {
namespace l16 {

constexpr int D = 16;
constexpr int M = D * (D + 1) / 2;     // 136
constexpr int P = D + M;               // 152
constexpr int kPitch = 156;            // floats per staged row: LDS.128 by row is bank-conflict free
constexpr int kWarps = 4;
constexpr int kThreads = kWarps * 32;
constexpr int kTileFloats = 32 * kPitch;

// position s (0..M-1) of the tril part of the raw vector -> (i, j) with j <= i (tfp fill_triangular, lower;
// same map as latent.cu::tril_ij)
struct IJ { int i, j; };
__host__ __device__ constexpr IJ tril_ij(int s) {
  if (s >= D) {
    const int k = s - D;
    const int i = k / D, j = k - (k / D) * D;
    if (j <= i && k < M - D) return IJ{i, j};
  }
  const int k = (M - D) + (M - 1 - s);
  return IJ{k / D, k - (k / D) * D};
}

__device__ __forceinline__ float softplus16(float x) { return fmaxf(x, 0.0f) + log1pf(expf(-fabsf(x))); }
__device__ __forceinline__ float sigmoid16(float x) { return 1.0f / (1.0f + expf(-x)); }

// ---- staging: global [rows, P] fp32 <-> registers of the owning lane, through the warp's tile
__device__ __forceinline__ void load_rows(const float* __restrict__ src, int64_t row0, int64_t B, float* tile, int lane,
                                          float (&v)[P]) {
  // rows row0 .. row0+31 are contiguous in global memory (pitch P)
  const int64_t nrows = (B - row0 < 32) ? (B - row0) : 32;
  const int nvec = (int)nrows * (P / 4);
  const float4* s4 = reinterpret_cast<const float4*>(src + row0 * P);
  __syncwarp();
  if (nrows == 32) {
    // full tile: all 38 loads of a lane are issued before the first one is consumed
    float4 buf[P / 4];
#pragma unroll
    for (int it = 0; it < P / 4; ++it) buf[it] = __ldg(s4 + it * 32 + lane);
#pragma unroll
    for (int it = 0; it < P / 4; ++it) {
      const int f = it * 32 + lane;
      const int r = f / (P / 4), q = f - r * (P / 4);
      *reinterpret_cast<float4*>(tile + r * kPitch + q * 4) = buf[it];
    }
  } else {
    for (int f = lane; f < nvec; f += 32) {
      const int r = f / (P / 4), q = f - r * (P / 4);
      *reinterpret_cast<float4*>(tile + r * kPitch + q * 4) = __ldg(s4 + f);
    }
  }
  __syncwarp();
  const float4* t4 = reinterpret_cast<const float4*>(tile + lane * kPitch);
#pragma unroll
  for (int q = 0; q < P / 4; ++q) {
    const float4 x = t4[q];
    v[4 * q] = x.x; v[4 * q + 1] = x.y; v[4 * q + 2] = x.z; v[4 * q + 3] = x.w;
  }
}

// registers (bf16-rounded) -> global [rows, P] bf16, through the same tile (pitch kPitch/2 words per row)
// csum (optional): per-lane partial column sums of the staged bf16 values, words lane, lane + 32, lane + 64
__device__ __forceinline__ void store_rows_bf16(const float (&v)[P], __nv_bfloat16* __restrict__ dst, int64_t row0, int64_t B,
                                                float* tile, int lane, float (*csum)[2] = nullptr) {
  uint32_t* tw = reinterpret_cast<uint32_t*>(tile);
  constexpr int kPitchW = 84;            // words per staged bf16 row (76 used): 84 = 20 mod 32 keeps uint4 stores by row conflict free
  __syncwarp();
#pragma unroll
  for (int q = 0; q < P / 8; ++q) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(v[8 * q + 2 * i], v[8 * q + 2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(tw + lane * kPitchW + 4 * q) = make_uint4(w[0], w[1], w[2], w[3]);
  }
  __syncwarp();
  const int64_t nrows = (B - row0 < 32) ? (B - row0) : 32;
  if (csum) {
    // bias gradient of the head Linear = column sums of these rows (the values the weight-gradient GEMM reads)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int w = lane + 32 * c;
      if (w < P / 2) {
        float a0 = 0.f, a1 = 0.f;
        for (int r = 0; r < (int)nrows; ++r) {
          const uint32_t x = tw[r * kPitchW + w];
          a0 += __uint_as_float(x << 16); a1 += __uint_as_float(x & 0xFFFF0000u);
        }
        csum[c][0] += a0; csum[c][1] += a1;
      }
    }
  }
  const int nvec = (int)nrows * (P / 8);                 // uint4 = 8 bf16; 19 per row
  uint4* d4 = reinterpret_cast<uint4*>(dst + row0 * P);
#pragma unroll 2
  for (int f = lane; f < nvec; f += 32) {
    const int r = f / (P / 8), q = f - r * (P / 8);
    d4[f] = *reinterpret_cast<const uint4*>(tw + r * kPitchW + 4 * q);
  }
}

__device__ __forceinline__ void load16(const float* __restrict__ src, int64_t row, bool ok, float (&e)[D]) {
  const float4* s4 = reinterpret_cast<const float4*>(src + row * D);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 x = ok ? __ldg(s4 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
    e[4 * q] = x.x; e[4 * q + 1] = x.y; e[4 * q + 2] = x.z; e[4 * q + 3] = x.w;
  }
}
__device__ __forceinline__ void store16(float* __restrict__ dst, int64_t row, const float (&e)[D]) {
  float4* d4 = reinterpret_cast<float4*>(dst + row * D);
#pragma unroll
  for (int q = 0; q < 4; ++q) d4[q] = make_float4(e[4 * q], e[4 * q + 1], e[4 * q + 2], e[4 * q + 3]);
}

// diag[i] = softplus(raw_ii) + 1e-5 for the row in v; returns sum log diag
__device__ __forceinline__ float diagonals(const float (&v)[P], float (&dg)[D]) {
  float logd = 0.f;
#pragma unroll
  for (int t = 0; t < M; ++t) {
    const IJ ij = tril_ij(t);
    if (ij.i == ij.j) {
      const float x = softplus16(v[D + t]) + 1e-5f;
      dg[ij.i] = x;
      logd += logf(x);
    }
  }
  return logd;
}

// r = L^-1 (z - mu) by forward substitution, L given by (v, dg); returns sum r^2
__device__ __forceinline__ float solve_lower(const float (&v)[P], const float (&dg)[D], const float (&z)[D], float (&r)[D]) {
  float s[D];
#pragma unroll
  for (int i = 0; i < D; ++i) s[i] = z[i] - v[i];
  // s_i -= L_ij r_j needs r_j first: walk the rows in order, gathering the (compile-time) positions of row i
  float sumsq = 0.f;
#pragma unroll
  for (int i = 0; i < D; ++i) {
#pragma unroll
    for (int t = 0; t < M; ++t) {
      const IJ ij = tril_ij(t);
      if (ij.i == i && ij.j < i) s[i] = fmaf(-v[D + t], r[ij.j], s[i]);
    }
    r[i] = s[i] / dg[i];
    sumsq = fmaf(r[i], r[i], sumsq);
  }
  return sumsq;
}

// ---------------------------------------------------------------- z = mu + L eps, KL(q || N(0, I))
__global__ void __launch_bounds__(kThreads) latent_fwd16_kernel(const float* __restrict__ par, const float* __restrict__ eps,
                                                                float* __restrict__ z, float* __restrict__ kl, int64_t B) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* tile = smem + warp * kTileFloats;
  const int64_t nwt = (B + 31) / 32;
  for (int64_t wt = (int64_t)blockIdx.x * kWarps + warp; wt < nwt; wt += (int64_t)gridDim.x * kWarps) {
    const int64_t row0 = wt * 32, row = row0 + lane;
    const bool ok = row < B;
    float v[P], e[D], zz[D];
    load_rows(par, row0, B, tile, lane, v);
    load16(eps, row, ok, e);
    float fro = 0.f, mu2 = 0.f, logd = 0.f;
#pragma unroll
    for (int i = 0; i < D; ++i) { zz[i] = v[i]; mu2 = fmaf(v[i], v[i], mu2); }
#pragma unroll
    for (int t = 0; t < M; ++t) {
      const IJ ij = tril_ij(t);
      float l = v[D + t];
      if (ij.i == ij.j) { l = softplus16(l) + 1e-5f; logd += logf(l); }
      zz[ij.i] = fmaf(l, e[ij.j], zz[ij.i]);
      fro = fmaf(l, l, fro);
    }
    if (ok) {
      store16(z, row, zz);
      kl[row] = -logd + 0.5f * (-(float)D + fro + mu2);
    }
  }
}

// ---------------------------------------------------------------- log q(z | x_o)
__global__ void __launch_bounds__(kThreads) match_fwd16_kernel(const float* __restrict__ par_p, const float* __restrict__ z,
                                                               float* __restrict__ match, int64_t B) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* tile = smem + warp * kTileFloats;
  const int64_t nwt = (B + 31) / 32;
  for (int64_t wt = (int64_t)blockIdx.x * kWarps + warp; wt < nwt; wt += (int64_t)gridDim.x * kWarps) {
    const int64_t row0 = wt * 32, row = row0 + lane;
    const bool ok = row < B;
    float v[P], zz[D], dg[D], r[D];
    load_rows(par_p, row0, B, tile, lane, v);
    load16(z, row, ok, zz);
    const float logd = diagonals(v, dg);
    const float sumsq = solve_lower(v, dg, zz, r);
    if (ok) match[row] = -0.5f * sumsq - logd - 0.5f * (float)D * kLog2Pi;
  }
}

// ---------------------------------------------------------------- backward of both heads
__global__ void __launch_bounds__(kThreads) latent_bwd16_kernel(
    const float* __restrict__ par_e, const float* __restrict__ par_p, const float* __restrict__ eps,
    const float* __restrict__ z, const float* __restrict__ dz_dec, const float* __restrict__ g_kl,
    const float* __restrict__ g_match, int stop_grad, __nv_bfloat16* __restrict__ dpar_e_b,
    __nv_bfloat16* __restrict__ dpar_p_b, float* __restrict__ db_e, float* __restrict__ db_p, int64_t B) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* tile = smem + warp * kTileFloats;
  const int64_t nwt = (B + 31) / 32;
  float cse[3][2] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}}, csp[3][2] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
  for (int64_t wt = (int64_t)blockIdx.x * kWarps + warp; wt < nwt; wt += (int64_t)gridDim.x * kWarps) {
    const int64_t row0 = wt * 32, row = row0 + lane;
    const bool ok = row < B;
    float v[P], zz[D], dz[D];
    load16(z, row, ok, zz);
    const float mw = ok ? g_match[row] : 0.f;
    // ---- partial posterior: d match / d par_p;  r = L_p^-1 (z - mu_p), g = L_p^-T r
    load_rows(par_p, row0, B, tile, lane, v);
    {
      float dg[D], r[D], g[D], tt[D];
      diagonals(v, dg);
      solve_lower(v, dg, zz, r);
#pragma unroll
      for (int i = 0; i < D; ++i) tt[i] = r[i];
#pragma unroll
      for (int i = D - 1; i >= 0; --i) {
        g[i] = tt[i] / dg[i];
#pragma unroll
        for (int t = 0; t < M; ++t) {
          const IJ ij = tril_ij(t);
          if (ij.i == i && ij.j < i) tt[ij.j] = fmaf(-v[D + t], g[i], tt[ij.j]);
        }
      }
      // gradient with respect to the raw vector, in place
#pragma unroll
      for (int t = 0; t < M; ++t) {
        const IJ ij = tril_ij(t);
        float val = g[ij.i] * r[ij.j];
        if (ij.i == ij.j) val = (val - 1.0f / dg[ij.i]) * sigmoid16(v[D + t]);
        v[D + t] = mw * val;
      }
#pragma unroll
      for (int i = 0; i < D; ++i) v[i] = mw * g[i];
      load16(dz_dec, row, ok && dz_dec != nullptr, dz);
      if (!stop_grad) {
#pragma unroll
        for (int i = 0; i < D; ++i) dz[i] -= mw * g[i];
      }
    }
    store_rows_bf16(v, dpar_p_b, row0, B, tile, lane, db_p ? csp : nullptr);
    // ---- posterior: z = mu + L eps (cotangent dz) and kw * KL
    load_rows(par_e, row0, B, tile, lane, v);
    {
      float e[D];
      load16(eps, row, ok, e);
      const float kw = ok ? g_kl[row] : 0.f;
#pragma unroll
      for (int t = 0; t < M; ++t) {
        const IJ ij = tril_ij(t);
        const float raw = v[D + t];
        float val;
        if (ij.i == ij.j) {
          const float dgv = softplus16(raw) + 1e-5f;
          val = (dz[ij.i] * e[ij.j] + kw * (dgv - 1.0f / dgv)) * sigmoid16(raw);
        } else {
          val = dz[ij.i] * e[ij.j] + kw * raw;
        }
        v[D + t] = val;
      }
#pragma unroll
      for (int i = 0; i < D; ++i) v[i] = dz[i] + kw * v[i];
    }
    store_rows_bf16(v, dpar_e_b, row0, B, tile, lane, db_e ? cse : nullptr);
  }
  // head bias gradients: one atomic per column per warp
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int w = lane + 32 * c;
    if (w < P / 2) {
      if (db_e) { atomicAdd(db_e + 2 * w, cse[c][0]); atomicAdd(db_e + 2 * w + 1, cse[c][1]); }
      if (db_p) { atomicAdd(db_p + 2 * w, csp[c][0]); atomicAdd(db_p + 2 * w + 1, csp[c][1]); }
    }
  }
}

// ---------------------------------------------------------------- K importance samples per row (evaluators)
// z[k, r, :] = mu_r + L_r eps[k, r, :] with eps = jax.random.normal(key, [K, B_total, 16]) restricted to rows
// row_start + r (vae.py:192-195), base[k, r] = log N(z; 0, I) - log q(z) = -|z|^2/2 + |eps|^2/2 + sum log L_ii
// (vae.py:203-212 at the posterior's own samples).  A warp takes 32 rows x kS samples: the rows' head vectors sit in
// registers, consecutive lanes write consecutive rows of z (coalesced).
constexpr int kS = 4;
__global__ void __launch_bounds__(kThreads) sample_latents16_kernel(const float* __restrict__ par, Key2 key, int64_t B,
                                                                    int64_t K, int64_t B_total, int64_t row_start,
                                                                    float* __restrict__ z, float* __restrict__ base) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* tile = smem + warp * kTileFloats;
  const int64_t nwt = (B + 31) / 32, nks = (K + kS - 1) / kS;
  const uint64_t n_total = (uint64_t)K * (uint64_t)B_total * (uint64_t)D;
  for (int64_t item = (int64_t)blockIdx.x * kWarps + warp; item < nwt * nks; item += (int64_t)gridDim.x * kWarps) {
    const int64_t wt = item % nwt, ks = item / nwt;
    const int64_t row0 = wt * 32, row = row0 + lane;
    const bool ok = row < B;
    float v[P], dg[D];
    load_rows(par, row0, B, tile, lane, v);
    const float logd = diagonals(v, dg);
    for (int kk = 0; kk < kS; ++kk) {
      const int64_t k = ks * kS + kk;
      if (k >= K) break;
      const uint64_t idx0 = ((uint64_t)k * (uint64_t)B_total + (uint64_t)(row_start + (ok ? row : 0))) * D;
      float e[D], zz[D];
      float e2 = 0.f, z2 = 0.f;
#pragma unroll
      for (int i = 0; i < D; ++i) {
        e[i] = bits_to_normal(jax_random_word(key, n_total, idx0 + i));
        e2 = fmaf(e[i], e[i], e2);
        zz[i] = v[i];
      }
#pragma unroll
      for (int t = 0; t < M; ++t) {
        const IJ ij = tril_ij(t);
        const float l = (ij.i == ij.j) ? dg[ij.i] : v[D + t];
        zz[ij.i] = fmaf(l, e[ij.j], zz[ij.i]);
      }
#pragma unroll
      for (int i = 0; i < D; ++i) z2 = fmaf(zz[i], zz[i], z2);
      if (ok) {
        store16(z, k * B + row, zz);
        base[k * B + row] = -0.5f * z2 + 0.5f * e2 + logd;
      }
    }
  }
}

static int grid_for(int64_t B) {
  int64_t g = ((B + 31) / 32 + kWarps - 1) / kWarps;
  if (g > 148 * 2) g = 148 * 2;
  if (g < 1) g = 1;
  return (int)g;
}
constexpr size_t kSmem = (size_t)kWarps * kTileFloats * sizeof(float);   // 79872 B

template <typename K>
static int set_smem(K kernel) {
  PMVAE_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem));
  return 0;
}

}  // namespace l16

int latent_fwd16(const float* par, const float* eps, float* z, float* kl, int64_t B, cudaStream_t s) {
  static bool once = false;
  if (!once) { PMVAE_TRY(l16::set_smem(l16::latent_fwd16_kernel)); once = true; }
  l16::latent_fwd16_kernel<<<l16::grid_for(B), l16::kThreads, l16::kSmem, s>>>(par, eps, z, kl, B);
  PMVAE_LAUNCH_CHECK();
  return 0;
}
int match_fwd16(const float* par_p, const float* z, float* match, int64_t B, cudaStream_t s) {
  static bool once = false;
  if (!once) { PMVAE_TRY(l16::set_smem(l16::match_fwd16_kernel)); once = true; }
  l16::match_fwd16_kernel<<<l16::grid_for(B), l16::kThreads, l16::kSmem, s>>>(par_p, z, match, B);
  PMVAE_LAUNCH_CHECK();
  return 0;
}
int sample_latents16(const float* par, Key2 key, int64_t B, int64_t K, int64_t B_total, int64_t row_start, float* z,
                     float* base, cudaStream_t s) {
  static bool once = false;
  if (!once) { PMVAE_TRY(l16::set_smem(l16::sample_latents16_kernel)); once = true; }
  const int64_t items = ((B + 31) / 32) * ((K + l16::kS - 1) / l16::kS);
  int64_t g = (items + l16::kWarps - 1) / l16::kWarps;
  if (g > 148 * 4) g = 148 * 4;
  if (g < 1) g = 1;
  l16::sample_latents16_kernel<<<(int)g, l16::kThreads, l16::kSmem, s>>>(par, key, B, K, B_total, row_start, z, base);
  PMVAE_LAUNCH_CHECK();
  return 0;
}
int latent_bwd16(const float* par_e, const float* par_p, const float* eps, const float* z, const float* dz_dec,
                 const float* g_kl, const float* g_match, int stop_grad, __nv_bfloat16* dpar_e_b,
                 __nv_bfloat16* dpar_p_b, float* db_e, float* db_p, int64_t B, cudaStream_t s) {
  static bool once = false;
  if (!once) { PMVAE_TRY(l16::set_smem(l16::latent_bwd16_kernel)); once = true; }
  l16::latent_bwd16_kernel<<<l16::grid_for(B), l16::kThreads, l16::kSmem, s>>>(par_e, par_p, eps, z, dz_dec, g_kl, g_match,
                                                                              stop_grad, dpar_e_b, dpar_p_b, db_e, db_p, B);
  PMVAE_LAUNCH_CHECK();
  return 0;
}

}  // namespace pmvae
