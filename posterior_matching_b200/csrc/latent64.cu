// TriL-Gaussian algebra for latent_dim = 64 (the bsds config), one warp per row, all fp32.
//
// The raw head vector of a row is P = 64 + 2080 floats: loc, then tfp's fill_triangular input v.  For d = 64 the
// fill_triangular map puts every row of L into ONE contiguous run of v:
//     rows i <= 31 :  L[i][j] = v[64 + 64 i + j]          (forward run)
//     rows i >= 32 :  L[i][j] = v[4095 - 64 i - j]        (reversed run; v[0..63] is row 63)
// so the factor is staged straight from global memory (coalesced, consecutive lanes = consecutive elements) into a
// dense [64][65] tile in shared memory.  The odd pitch makes both access patterns of the triangular algebra
// bank-conflict free: fixed row / varying column (dot products, back substitution) and fixed column / varying row
// (forward substitution) -- the general-d kernels of latent.cu address the packed vector with a stride of 64 floats
// between lanes (32-way conflicts) and pay the index map per element, which made them 47 % of the bsds train step.
//
// Reference sites as in latent.cu: distributions.py:101-113 (TriLGaussian / FillScaleTriL), vae.py:124,130,136-138.
#include "kernels.h"

namespace pmvae {
namespace l64 {

constexpr int D = 64, M = D * (D + 1) / 2, P = D + M;      // 2080, 2144
constexpr int kPitch = D + 1;
constexpr int kWarps = 4;                                  // per block
constexpr int kThreads = kWarps * 32;
constexpr int kLFloats = D * kPitch;                       // dense factor
constexpr int kVecs = 7;                                   // per-warp vectors of D floats
constexpr int kWarpFloats = kLFloats + kVecs * D;
constexpr size_t kSmemBytes = (size_t)kWarps * kWarpFloats * sizeof(float);

// position q of v (0 .. 2079) -> (i, j), j <= i
__device__ __forceinline__ void v_to_ij(int q, int& i, int& j) {
  if (q >= D) {
    const int k = q - D;
    i = k >> 6; j = k & 63;
    if (j <= i) return;
  }
  const int k2 = 4095 - q;
  i = k2 >> 6; j = k2 & 63;
}

// Stages L (diagonal = softplus(raw) + 1e-5) into Lp[64][65]; sraw[i] = raw diagonal, sinv[i] = 1 / L_ii.
// Returns this lane's share of sum_i log L_ii.
__device__ __forceinline__ float stage_factor(const float* __restrict__ pr, float* Lp, float* sraw, float* sinv, int lane) {
  float logd = 0.f;
  const float* v = pr + D;
#pragma unroll 5
  for (int it = 0; it < M / 32; ++it) {
    const int q = lane + 32 * it;
    float val = __ldg(v + q);
    int i, j;
    v_to_ij(q, i, j);
    if (i == j) {
      sraw[i] = val;
      val = softplus_f(val) + 1e-5f;
      sinv[i] = 1.0f / val;
      logd += logf(val);
    }
    Lp[i * kPitch + j] = val;
  }
  return logd;
}

// r = L^-1 s (forward substitution, column oriented): lanes own s_k for k = lane, lane + 32; r -> sr[], returns |r|^2.
__device__ __forceinline__ float solve_lower(const float* Lp, const float* sinv, float s0, float s1, float* sr, int lane) {
  float sumsq = 0.f;
#pragma unroll 4
  for (int i = 0; i < D; ++i) {
    const float src = (i < 32) ? s0 : s1;
    const float ri = __shfl_sync(0xffffffffu, src, i & 31) * sinv[i];
    sumsq = fmaf(ri, ri, sumsq);
    if (lane == 0) sr[i] = ri;
    if (lane > i) s0 = fmaf(-Lp[lane * kPitch + i], ri, s0);
    if (lane + 32 > i) s1 = fmaf(-Lp[(lane + 32) * kPitch + i], ri, s1);
  }
  __syncwarp();
  return sumsq;
}

// g = L^-T t (back substitution): lanes own t_k; g -> sg[]
__device__ __forceinline__ void solve_upper_t(const float* Lp, const float* sinv, float t0, float t1, float* sg, int lane) {
#pragma unroll 4
  for (int i = D - 1; i >= 0; --i) {
    const float src = (i < 32) ? t0 : t1;
    const float gi = __shfl_sync(0xffffffffu, src, i & 31) * sinv[i];
    if (lane == 0) sg[i] = gi;
    if (lane < i) t0 = fmaf(-Lp[i * kPitch + lane], gi, t0);
    if (lane + 32 < i) t1 = fmaf(-Lp[i * kPitch + lane + 32], gi, t1);
  }
  __syncwarp();
}

// ---------------------------------------------------------------- z = mu + L eps, KL(q || N(0, I))
__global__ void __launch_bounds__(kThreads) latent_fwd64_kernel(const float* __restrict__ par, const float* __restrict__ eps,
                                                                float* __restrict__ z, float* __restrict__ kl, int64_t B) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  float* Lp = smem + (size_t)wib * kWarpFloats;
  float* se = Lp + kLFloats;
  float* sraw = se + D;
  float* sinv = sraw + D;
  for (int64_t r = (int64_t)blockIdx.x * kWarps + wib; r < B; r += (int64_t)gridDim.x * kWarps) {
    const float* pr = par + r * P;
    __syncwarp();
    se[lane] = eps[r * D + lane]; se[lane + 32] = eps[r * D + lane + 32];
    float logd = stage_factor(pr, Lp, sraw, sinv, lane);
    __syncwarp();
    const float mu0 = __ldg(pr + lane), mu1 = __ldg(pr + lane + 32);
    float a0 = mu0, a1 = mu1, fro = 0.f;
#pragma unroll 8
    for (int j = 0; j < D; ++j) {
      const float e = se[j];
      const float l0 = (j <= lane) ? Lp[lane * kPitch + j] : 0.f;
      const float l1 = (j <= lane + 32) ? Lp[(lane + 32) * kPitch + j] : 0.f;
      a0 = fmaf(l0, e, a0); a1 = fmaf(l1, e, a1);
      fro = fmaf(l0, l0, fro); fro = fmaf(l1, l1, fro);
    }
    z[r * D + lane] = a0; z[r * D + lane + 32] = a1;
    const float mu2 = warp_sum(mu0 * mu0 + mu1 * mu1);
    fro = warp_sum(fro); logd = warp_sum(logd);
    if (lane == 0) kl[r] = -logd + 0.5f * (-(float)D + fro + mu2);
  }
}

// ---------------------------------------------------------------- log q(z | x_o) (vae.py:136-138)
__global__ void __launch_bounds__(kThreads) match_fwd64_kernel(const float* __restrict__ par_p, const float* __restrict__ z,
                                                               float* __restrict__ match, int64_t B) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  float* Lp = smem + (size_t)wib * kWarpFloats;
  float* sr = Lp + kLFloats;
  float* sraw = sr + D;
  float* sinv = sraw + D;
  for (int64_t r = (int64_t)blockIdx.x * kWarps + wib; r < B; r += (int64_t)gridDim.x * kWarps) {
    const float* pr = par_p + r * P;
    __syncwarp();
    float logd = stage_factor(pr, Lp, sraw, sinv, lane);
    __syncwarp();
    const float s0 = z[r * D + lane] - __ldg(pr + lane), s1 = z[r * D + lane + 32] - __ldg(pr + lane + 32);
    const float sumsq = solve_lower(Lp, sinv, s0, s1, sr, lane);
    logd = warp_sum(logd);
    if (lane == 0) match[r] = -0.5f * sumsq - logd - 0.5f * (float)D * kLog2Pi;
  }
}

// ---------------------------------------------------------------- backward of both heads
// Cotangents g_kl[r], g_match[r] and dz_dec (decoder) -> d / d par_e, d / d par_p (bf16, the A operands of the two head
// Linears' backward), formulas as latent.cu::latent_bwd_kernel.  Optionally also the two head bias gradients (column
// sums of the bf16 values), accumulated per block in shared memory and flushed with one atomicAdd per column.
__global__ void __launch_bounds__(kThreads) latent_bwd64_kernel(
    const float* __restrict__ par_e, const float* __restrict__ par_p, const float* __restrict__ eps,
    const float* __restrict__ z, const float* __restrict__ dz_dec, const float* __restrict__ g_kl,
    const float* __restrict__ g_match, int stop_grad, __nv_bfloat16* __restrict__ dpar_e_b,
    __nv_bfloat16* __restrict__ dpar_p_b, float* __restrict__ db_e, float* __restrict__ db_p, int64_t B) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  float* Lp = smem + (size_t)wib * kWarpFloats;
  float* sr = Lp + kLFloats;     // r = L_p^-1 (z - mu_p)
  float* sg = sr + D;            // g = L_p^-T r
  float* se = sg + D;            // eps
  float* sdz = se + D;           // dz_total
  float* sraw = sdz + D;
  float* sinv = sraw + D;
  float* cs = smem + (size_t)kWarps * kWarpFloats;      // [2][P] column sums of this block (only with db_e / db_p)
  const bool do_cs = db_e != nullptr && db_p != nullptr;
  if (do_cs) {
    for (int q = threadIdx.x; q < 2 * P; q += kThreads) cs[q] = 0.f;
    __syncthreads();
  }
  for (int64_t r = (int64_t)blockIdx.x * kWarps + wib; r < B; r += (int64_t)gridDim.x * kWarps) {
    const float* pp = par_p + r * P;
    const float* pe = par_e + r * P;
    __syncwarp();
    se[lane] = eps[r * D + lane]; se[lane + 32] = eps[r * D + lane + 32];
    stage_factor(pp, Lp, sraw, sinv, lane);
    __syncwarp();
    // ---- partial posterior: r, g
    const float s0 = z[r * D + lane] - __ldg(pp + lane), s1 = z[r * D + lane + 32] - __ldg(pp + lane + 32);
    solve_lower(Lp, sinv, s0, s1, sr, lane);
    solve_upper_t(Lp, sinv, sr[lane], sr[lane + 32], sg, lane);
    const float mw = g_match[r];
    // d match / d par_p: loc -> mw g; L_ij -> mw (g_i r_j - [i == j] / L_ii) (diagonal through softplus)
    __nv_bfloat16* op = dpar_p_b + r * P;
#pragma unroll 1
    for (int q0 = 0; q0 < P; q0 += 32) {
      const int q = q0 + lane;
      float val;
      if (q < D) {
        val = mw * sg[q];
      } else {
        int i, j;
        v_to_ij(q - D, i, j);
        val = sg[i] * sr[j];
        if (i == j) val = (val - sinv[i]) * sigmoid_f(sraw[i]);
        val *= mw;
      }
      const __nv_bfloat16 hv = __float2bfloat16(val);
      op[q] = hv;
      if (do_cs) atomicAdd(cs + P + q, __bfloat162float(hv));
    }
    // ---- dz_total = dz_dec - (stop_grad ? 0 : mw g)
    {
      float v0 = dz_dec ? dz_dec[r * D + lane] : 0.f, v1 = dz_dec ? dz_dec[r * D + lane + 32] : 0.f;
      if (!stop_grad) { v0 -= mw * sg[lane]; v1 -= mw * sg[lane + 32]; }
      sdz[lane] = v0; sdz[lane + 32] = v1;
    }
    __syncwarp();
    // ---- posterior: z = mu + L eps and kw * KL
    const float kw = g_kl[r];
    __nv_bfloat16* oe = dpar_e_b + r * P;
#pragma unroll 1
    for (int q0 = 0; q0 < P; q0 += 32) {
      const int q = q0 + lane;
      const float raw = __ldg(pe + q);
      float val;
      if (q < D) {
        val = sdz[q] + kw * raw;
      } else {
        int i, j;
        v_to_ij(q - D, i, j);
        if (i == j) {
          const float dg = softplus_f(raw) + 1e-5f;
          val = (sdz[i] * se[j] + kw * (dg - 1.0f / dg)) * sigmoid_f(raw);
        } else {
          val = sdz[i] * se[j] + kw * raw;
        }
      }
      const __nv_bfloat16 hv = __float2bfloat16(val);
      oe[q] = hv;
      if (do_cs) atomicAdd(cs + q, __bfloat162float(hv));
    }
  }
  if (do_cs) {
    __syncthreads();
    for (int q = threadIdx.x; q < P; q += kThreads) {
      atomicAdd(db_e + q, cs[q]);
      atomicAdd(db_p + q, cs[P + q]);
    }
  }
}

static int grid_rows(int64_t B, int blocks_per_sm) {
  int64_t g = (B + kWarps - 1) / kWarps;
  const int64_t cap = 148ll * blocks_per_sm;
  if (g > cap) g = cap;
  return (int)(g < 1 ? 1 : g);
}

}  // namespace l64

int latent_fwd64(const float* par, const float* eps, float* z, float* kl, int64_t B, cudaStream_t s) {
  using namespace l64;
  static bool attr = false;
  if (!attr) { PMVAE_CUDA(cudaFuncSetAttribute(latent_fwd64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes)); attr = true; }
  latent_fwd64_kernel<<<grid_rows(B, 3), kThreads, kSmemBytes, s>>>(par, eps, z, kl, B);
  PMVAE_LAUNCH_CHECK();
  return 0;
}

int match_fwd64(const float* par_p, const float* z, float* match, int64_t B, cudaStream_t s) {
  using namespace l64;
  static bool attr = false;
  if (!attr) { PMVAE_CUDA(cudaFuncSetAttribute(match_fwd64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes)); attr = true; }
  match_fwd64_kernel<<<grid_rows(B, 3), kThreads, kSmemBytes, s>>>(par_p, z, match, B);
  PMVAE_LAUNCH_CHECK();
  return 0;
}

int latent_bwd64(const float* par_e, const float* par_p, const float* eps, const float* z, const float* dz_dec,
                 const float* g_kl, const float* g_match, int stop_grad, __nv_bfloat16* dpar_e_b, __nv_bfloat16* dpar_p_b,
                 float* db_e, float* db_p, int64_t B, cudaStream_t s) {
  using namespace l64;
  const bool do_cs = db_e != nullptr && db_p != nullptr;
  const size_t smem = kSmemBytes + (do_cs ? (size_t)2 * P * sizeof(float) : 0);
  static size_t attr = 0;
  if (attr < smem) { PMVAE_CUDA(cudaFuncSetAttribute(latent_bwd64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = smem; }
  latent_bwd64_kernel<<<grid_rows(B, do_cs ? 2 : 3), kThreads, smem, s>>>(par_e, par_p, eps, z, dz_dec, g_kl, g_match, stop_grad,
                                                                         dpar_e_b, dpar_p_b, db_e, db_p, B);
  PMVAE_LAUNCH_CHECK();
  return 0;
}

}  // namespace pmvae
