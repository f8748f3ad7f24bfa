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
constexpr int kVecs = 4;                                   // per-warp vectors of D floats
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

// index into v of the diagonal element (i, i)
__device__ __forceinline__ int diag_q(int i) { return i < 32 ? D + 65 * i : 4095 - 65 * i; }

// Stages L (diagonal = softplus(raw) + 1e-5) into Lp[64][65]; sraw[i] = raw diagonal, sinv[i] = 1 / L_ii.
// Returns this lane's share of sum_i log L_ii.  The 65 loads of a lane are issued in batches of 13 before anything is
// stored (stores to shared memory would otherwise pin every load behind the previous iteration: one HBM round trip
// per element), and the transcendental work of the 64 diagonal elements is done once, two per lane, after the copy.
__device__ __forceinline__ float stage_factor(const float* __restrict__ pr, float* Lp, float* sraw, float* sinv, int lane) {
  const float* v = pr + D;
#pragma unroll 1
  for (int it0 = 0; it0 < M / 32; it0 += 13) {
    float buf[13];
#pragma unroll
    for (int u = 0; u < 13; ++u) buf[u] = __ldg(v + lane + 32 * (it0 + u));
#pragma unroll
    for (int u = 0; u < 13; ++u) {
      int i, j;
      v_to_ij(lane + 32 * (it0 + u), i, j);
      Lp[i * kPitch + j] = buf[u];
    }
  }
  __syncwarp();
  float logd = 0.f;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int i = lane + 32 * h;
    const float raw = Lp[i * kPitch + i];
    const float dg = softplus_f(raw) + 1e-5f;
    sraw[i] = raw;
    sinv[i] = 1.0f / dg;
    Lp[i * kPitch + i] = dg;
    logd += logf(dg);
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
template <bool CS>
__global__ void __launch_bounds__(kThreads) latent_bwd64_kernel(
    const float* __restrict__ par_e, const float* __restrict__ par_p, const float* __restrict__ eps,
    const float* __restrict__ z, const float* __restrict__ dz_dec, const float* __restrict__ g_kl,
    const float* __restrict__ g_match, int stop_grad, __nv_bfloat16* __restrict__ dpar_e_b,
    __nv_bfloat16* __restrict__ dpar_p_b, float* __restrict__ db_e, float* __restrict__ db_p, float* __restrict__ dz_total,
    int64_t B) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  float* Lp = smem + (size_t)wib * kWarpFloats;
  float* sr = Lp + kLFloats;     // r = L_p^-1 (z - mu_p)
  float* sg = sr + D;            // g = L_p^-T r
  float* sraw = sg + D;
  float* sinv = sraw + D;
  // CS: column sums of the bf16 outputs (the two head bias gradients) in registers: lane keeps column 32 it + lane of
  // both outputs over all rows of its warp (shared-memory atomics cost 64 cycles per warp and column: 4 ms per step)
  constexpr int kIt = P / 32;                 // 67
  float acc_p[CS ? kIt : 1], dacc_p[2] = {0.f, 0.f};
  if (CS) {
#pragma unroll
    for (int it = 0; it < kIt; ++it) acc_p[it] = 0.f;
  }
  for (int64_t r = (int64_t)blockIdx.x * kWarps + wib; r < B; r += (int64_t)gridDim.x * kWarps) {
    const float* pp = par_p + r * P;
    __syncwarp();
    stage_factor(pp, Lp, sraw, sinv, lane);
    __syncwarp();
    // ---- partial posterior: r, g
    const float s0 = z[r * D + lane] - __ldg(pp + lane), s1 = z[r * D + lane + 32] - __ldg(pp + lane + 32);
    solve_lower(Lp, sinv, s0, s1, sr, lane);
    solve_upper_t(Lp, sinv, sr[lane], sr[lane + 32], sg, lane);
    const float mw = g_match[r];
    // d match / d par_p: loc -> mw g; L_ij -> mw (g_i r_j - [i == j] / L_ii) (diagonal through softplus: second pass)
    __nv_bfloat16* op = dpar_p_b + r * P;
#pragma unroll
    for (int it = 0; it < kIt; ++it) {
      const int q = 32 * it + lane;
      float val;
      bool diag = false;
      if (it < D / 32) {
        val = mw * sg[q];
      } else {
        int i, j;
        v_to_ij(q - D, i, j);
        val = mw * sg[i] * sr[j];
        diag = i == j;
      }
      if (!diag) {
        const __nv_bfloat16 hv = __float2bfloat16(val);
        op[q] = hv;
        if (CS) acc_p[it] += __bfloat162float(hv);
      }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int i = lane + 32 * h;
      const float val = mw * (sg[i] * sr[i] - sinv[i]) * sigmoid_f(sraw[i]);
      const __nv_bfloat16 hv = __float2bfloat16(val);
      op[D + diag_q(i)] = hv;
      if (CS) dacc_p[h] += __bfloat162float(hv);
    }
    // ---- dz_total = dz_dec - (stop_grad ? 0 : mw g): consumed by the posterior part (post_bwd64_kernel)
    {
      float v0 = dz_dec ? dz_dec[r * D + lane] : 0.f, v1 = dz_dec ? dz_dec[r * D + lane + 32] : 0.f;
      if (!stop_grad) { v0 -= mw * sg[lane]; v1 -= mw * sg[lane + 32]; }
      dz_total[r * D + lane] = v0; dz_total[r * D + lane + 32] = v1;
    }
  }
  if (CS) {
    // block reduction in the (now idle) factor tiles: warp w writes its 2 x P sums, then every column is summed over
    // the four warps and added to the gradient arena with one atomic per column and block
    __syncthreads();
    float* mine = smem + (size_t)wib * kWarpFloats;
#pragma unroll
    for (int it = 0; it < kIt; ++it) mine[32 * it + lane] = acc_p[it];
    __syncwarp();
#pragma unroll
    for (int h = 0; h < 2; ++h) mine[D + diag_q(lane + 32 * h)] += dacc_p[h];
    __syncthreads();
    for (int q = threadIdx.x; q < P; q += kThreads) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) t += smem[(size_t)w * kWarpFloats + q];
      atomicAdd(db_p + q, t);
    }
  }
}

// ---------------------------------------------------------------- posterior part: d / d par_e (streaming)
// d / d par_e of z = mu + L eps (cotangent dz_total) and kw KL: loc -> dz_i + kw mu_i; L_ij -> dz_i eps_j + kw L_ij;
// diagonal (through softplus) -> (dz_i eps_i + kw (D_ii - 1 / D_ii)) sigmoid(raw).  No triangular algebra: thread =
// column q, block = kPostRows consecutive rows, so reads and writes are coalesced and the head bias gradient (column
// sums of the bf16 values) is one register per thread and one atomicAdd per column and block.
constexpr int kPostThreads = 256, kPostRows = 64;
__global__ void __launch_bounds__(kPostThreads) post_bwd64_kernel(const float* __restrict__ par_e, const float* __restrict__ eps,
                                                                  const float* __restrict__ dz_total,
                                                                  const float* __restrict__ g_kl,
                                                                  __nv_bfloat16* __restrict__ dpar_e_b,
                                                                  float* __restrict__ db_e, int64_t B) {
  __shared__ float sdz[kPostRows][D], se[kPostRows][D], skw[kPostRows];
  const int64_t r0 = (int64_t)blockIdx.y * kPostRows;
  const int nr = (int)((B - r0 < kPostRows) ? (B - r0) : kPostRows);
  const int q = blockIdx.x * kPostThreads + threadIdx.x;
  for (int t = threadIdx.x; t < nr * D; t += kPostThreads) {
    sdz[t / D][t % D] = dz_total[r0 * D + t];
    se[t / D][t % D] = eps[r0 * D + t];
  }
  if (threadIdx.x < nr) skw[threadIdx.x] = g_kl[r0 + threadIdx.x];
  __syncthreads();
  if (q >= P) return;
  int i = 0, j = 0;
  if (q >= D) v_to_ij(q - D, i, j);
  const bool is_loc = q < D, diag = !is_loc && i == j;
  float acc = 0.f;
#pragma unroll 8
  for (int rr = 0; rr < nr; ++rr) {
    const float raw = __ldg(par_e + (r0 + rr) * P + q);
    const float kw = skw[rr];
    float val;
    if (is_loc) {
      val = sdz[rr][q] + kw * raw;
    } else if (diag) {
      const float dg = softplus_f(raw) + 1e-5f;
      val = (sdz[rr][i] * se[rr][i] + kw * (dg - 1.0f / dg)) * sigmoid_f(raw);
    } else {
      val = sdz[rr][i] * se[rr][j] + kw * raw;
    }
    const __nv_bfloat16 hv = __float2bfloat16(val);
    dpar_e_b[(r0 + rr) * P + q] = hv;
    acc += __bfloat162float(hv);
  }
  if (db_e) atomicAdd(db_e + q, acc);
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
                 float* db_e, float* db_p, float* dz_total, int64_t B, cudaStream_t s) {
  using namespace l64;
  const bool do_cs = db_e != nullptr && db_p != nullptr;
  static bool attr = false;
  if (!attr) {
    PMVAE_CUDA(cudaFuncSetAttribute(latent_bwd64_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    PMVAE_CUDA(cudaFuncSetAttribute(latent_bwd64_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    attr = true;
  }
  // partial posterior (two triangular solves per row) -> d / d par_p, dz_total
  if (do_cs)
    latent_bwd64_kernel<true><<<grid_rows(B, 3), kThreads, kSmemBytes, s>>>(par_e, par_p, eps, z, dz_dec, g_kl, g_match, stop_grad,
                                                                             dpar_e_b, dpar_p_b, db_e, db_p, dz_total, B);
  else
    latent_bwd64_kernel<false><<<grid_rows(B, 3), kThreads, kSmemBytes, s>>>(par_e, par_p, eps, z, dz_dec, g_kl, g_match, stop_grad,
                                                                              dpar_e_b, dpar_p_b, db_e, db_p, dz_total, B);
  PMVAE_LAUNCH_CHECK();
  // posterior: streaming elementwise kernel
  const dim3 grid((P + kPostThreads - 1) / kPostThreads, (unsigned)((B + kPostRows - 1) / kPostRows));
  post_bwd64_kernel<<<grid, kPostThreads, 0, s>>>(par_e, eps, dz_total, g_kl, dpar_e_b, do_cs ? db_e : nullptr, B);
  PMVAE_LAUNCH_CHECK();
  return 0;
}

}  // namespace pmvae
