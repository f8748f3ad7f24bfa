// TriL-Gaussian algebra for latent_dim = 64 (the bsds config), one warp per row, all fp32.
//
// The raw head vector of a row is P = 64 + 2080 floats: loc, then tfp's fill_triangular input v.  For d = 64 the
// fill_triangular map puts every row of L into ONE contiguous run of v:
//     rows i <= 31 :  L[i][j] = v[64 + 64 i + j]          (forward run)
//     rows i >= 32 :  L[i][j] = v[4095 - 64 i - j]        (reversed run; v[0..63] is row 63)
// i.e. v, read as a matrix X of 32.5 rows of 64, holds row a - 1 of L in the left part of its row a and row 63 - a,
// reversed, in the right part.  The kernels keep exactly that packed image in shared memory (33 rows, pitch 65 floats,
// 8.6 KB per warp: 20 warps per SM where a dense [64][65] tile allowed 12), so staging is a straight coalesced copy with
// compile-time offsets, and
//     L[i][j] = X[i + 1][j]            (i <= 31)          L[i][j] = X[63 - i][63 - j]      (i >= 32).
// The odd pitch makes both access patterns of the triangular algebra bank-conflict free: fixed row / varying column
// (dot products, back substitution; contiguous, ascending or descending) and fixed column / varying row (forward
// substitution; 65 floats apart) -- the general-d kernels of latent.cu address the packed vector with a stride of 64
// floats between lanes (32-way conflicts) and pay the index map per element, which made them 47 % of the bsds step.
//
// Reference sites as in latent.cu: distributions.py:101-113 (TriLGaussian / FillScaleTriL), vae.py:124,130,136-138.
#include <type_traits>

#include "kernels.h"

namespace pmvae {
namespace l64 {

constexpr int D = 64, M = D * (D + 1) / 2, P = D + M;      // 2080, 2144
constexpr int kPitch = D + 1;
constexpr int kWarps = 4;                                  // per block
constexpr int kThreads = kWarps * 32;
constexpr int kLFloats = 33 * kPitch;                      // packed factor (see above)
constexpr int kVecs = 4;                                   // per-warp vectors of D floats
constexpr int kWarpFloats = kLFloats + kVecs * D;
constexpr size_t kSmemBytes = (size_t)kWarps * kWarpFloats * sizeof(float);
constexpr int kBlocksPerSm = 5;                            // 38.4 KB and <= 96 registers x 128 threads per block

// offset of L[i][j] (j <= i) in the packed image
__device__ __forceinline__ int xoff(int i, int j) { return i < 32 ? (i + 1) * kPitch + j : (63 - i) * kPitch + 63 - j; }

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

// The factor of a row is 65 floats per lane.  `load_row` issues all of them (one HBM round trip instead of one per
// element or per batch: stores to shared memory in between would pin every load behind the previous one), and the
// kernels call it for the NEXT row before they start the triangular algebra of the current one, so that the latency
// is hidden behind ~4 K cycles of dependent solve steps.
constexpr int kRowVals = M / 32;       // 65
__device__ __forceinline__ void load_row(const float* __restrict__ pr, float (&buf)[kRowVals], int lane) {
  const float* v = pr + D;
#pragma unroll
  for (int u = 0; u < kRowVals; ++u) buf[u] = __ldg(v + lane + 32 * u);
}

// Stages L (diagonal = softplus(raw) + 1e-5) from the registers filled by load_row into the packed image Xs; sraw[i] =
// raw diagonal, sinv[i] = 1 / L_ii.  Returns this lane's share of sum_i log L_ii.  The transcendental work of the 64
// diagonal elements is done once, two per lane, after the copy.
__device__ __forceinline__ float stage_factor(const float (&buf)[kRowVals], float* Xs, float* sraw, float* sinv, int lane) {
#pragma unroll
  for (int u = 0; u < kRowVals; ++u) Xs[(u >> 1) * kPitch + 32 * (u & 1) + lane] = buf[u];
  __syncwarp();
  float logd = 0.f;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int i = lane + 32 * h, o = xoff(i, i);
    const float raw = Xs[o];
    const float dg = softplus_f(raw) + 1e-5f;
    sraw[i] = raw;
    sinv[i] = 1.0f / dg;
    Xs[o] = dg;
    logd += logf(dg);
  }
  return logd;
}

// r = L^-1 s (forward substitution, column oriented): lanes own s_k for k = lane, lane + 32; r -> sr[], returns |r|^2.
__device__ __forceinline__ float solve_lower(const float* Xs, const float* sinv, float s0, float s1, float* sr, int lane) {
  const float* c0 = Xs + (lane + 1) * kPitch;           // L[lane][i]      = c0[i]
  const float* c1 = Xs + (31 - lane) * kPitch + 63;     // L[lane + 32][i] = c1[-i]
  float sumsq = 0.f;
#pragma unroll 4
  for (int i = 0; i < D; ++i) {
    const float src = (i < 32) ? s0 : s1;
    const float ri = __shfl_sync(0xffffffffu, src, i & 31) * sinv[i];
    sumsq = fmaf(ri, ri, sumsq);
    if (lane == 0) sr[i] = ri;
    if (lane > i) s0 = fmaf(-c0[i], ri, s0);
    if (lane + 32 > i) s1 = fmaf(-c1[-i], ri, s1);
  }
  __syncwarp();
  return sumsq;
}

// g = L^-T t (back substitution): lanes own t_k; g -> sg[]
__device__ __forceinline__ void solve_upper_t(const float* Xs, const float* sinv, float t0, float t1, float* sg, int lane) {
#pragma unroll 4
  for (int i = D - 1; i >= 32; --i) {                   // row i >= 32: L[i][j] = row[-j]
    const float* row = Xs + (63 - i) * kPitch + 63;
    const float gi = __shfl_sync(0xffffffffu, t1, i & 31) * sinv[i];
    if (lane == 0) sg[i] = gi;
    t0 = fmaf(-row[-lane], gi, t0);
    if (lane + 32 < i) t1 = fmaf(-row[-lane - 32], gi, t1);
  }
#pragma unroll 4
  for (int i = 31; i >= 0; --i) {                       // row i <= 31: L[i][j] = row[j]
    const float* row = Xs + (i + 1) * kPitch;
    const float gi = __shfl_sync(0xffffffffu, t0, i) * sinv[i];
    if (lane == 0) sg[i] = gi;
    if (lane < i) t0 = fmaf(-row[lane], gi, t0);
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
  const int64_t stride = (int64_t)gridDim.x * kWarps;
  float buf[kRowVals];
  int64_t r = (int64_t)blockIdx.x * kWarps + wib;
  if (r < B) load_row(par + r * P, buf, lane);
  for (; r < B; r += stride) {
    const float* pr = par + r * P;
    __syncwarp();
    se[lane] = eps[r * D + lane]; se[lane + 32] = eps[r * D + lane + 32];
    float logd = stage_factor(buf, Lp, sraw, sinv, lane);
    if (r + stride < B) load_row(par + (r + stride) * P, buf, lane);
    __syncwarp();
    const float mu0 = __ldg(pr + lane), mu1 = __ldg(pr + lane + 32);
    float a0 = mu0, a1 = mu1, fro = 0.f;
    const float* c0 = Lp + (lane + 1) * kPitch;
    const float* c1 = Lp + (31 - lane) * kPitch + 63;
#pragma unroll 8
    for (int j = 0; j < D; ++j) {
      const float e = se[j];
      const float l0 = (j <= lane) ? c0[j] : 0.f;
      const float l1 = (j <= lane + 32) ? c1[-j] : 0.f;
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
// Training forward (SAVE): the factor is staged here anyway, so the backward's triangular algebra is done on it now and
// three 64-vectors per row are kept instead of re-reading the 8.6 KB of par_p in the backward:
//     r = L_p^-1 (z - mu_p),   g = L_p^-T r,   qd_i = (g_i r_i - 1 / D_ii) sigmoid(raw_ii)   (diagonal slots of d / d par_p
// before the per-row cotangent).
template <bool SAVE>
__global__ void __launch_bounds__(kThreads) match_fwd64_kernel(const float* __restrict__ par_p, const float* __restrict__ z,
                                                               float* __restrict__ match, float* __restrict__ out_r,
                                                               float* __restrict__ out_g, float* __restrict__ out_qd,
                                                               int64_t B) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  float* Lp = smem + (size_t)wib * kWarpFloats;
  float* sr = Lp + kLFloats;
  float* sraw = sr + D;
  float* sinv = sraw + D;
  float* sg = sinv + D;
  const int64_t stride = (int64_t)gridDim.x * kWarps;
  float buf[kRowVals];
  int64_t r = (int64_t)blockIdx.x * kWarps + wib;
  if (r < B) load_row(par_p + r * P, buf, lane);
  for (; r < B; r += stride) {
    const float* pr = par_p + r * P;
    __syncwarp();
    float logd = stage_factor(buf, Lp, sraw, sinv, lane);
    if (r + stride < B) load_row(par_p + (r + stride) * P, buf, lane);
    __syncwarp();
    const float s0 = z[r * D + lane] - __ldg(pr + lane), s1 = z[r * D + lane + 32] - __ldg(pr + lane + 32);
    const float sumsq = solve_lower(Lp, sinv, s0, s1, sr, lane);
    logd = warp_sum(logd);
    if (lane == 0) match[r] = -0.5f * sumsq - logd - 0.5f * (float)D * kLog2Pi;
    if (SAVE) {
      solve_upper_t(Lp, sinv, sr[lane], sr[lane + 32], sg, lane);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int i = lane + 32 * h;
        out_r[r * D + i] = sr[i];
        out_g[r * D + i] = sg[i];
        out_qd[r * D + i] = (sg[i] * sr[i] - sinv[i]) * sigmoid_f(sraw[i]);
      }
    }
  }
}

// ---------------------------------------------------------------- backward of both heads
// The triangular algebra of the partial posterior (r, g and the diagonal terms qd) was done by match_fwd64_kernel<true>
// in the forward; what is left is elementwise in the P = 2144 head columns, so it streams: thread = a pair of columns,
// block = kPostRows consecutive rows (coalesced reads and bf16 writes, the head bias gradients are one register per
// thread and output):
//     d / d par_p: loc -> mw g;  L_ij -> mw g_i r_j;  diagonal -> mw (g_i r_i - 1 / D_ii) sigmoid(raw_ii)
//     d / d par_e: loc -> dz_i + kw mu_i;  L_ij -> dz_i eps_j + kw L_ij;  diagonal -> (dz_i eps_i + kw (D_ii - 1 / D_ii)) sigmoid(raw_ii)
// (formulas as latent.cu::latent_bwd_kernel; D_ii = softplus(raw_ii) + 1e-5).
// Thread = a PAIR of adjacent head columns (P and D are even, so a pair never straddles loc | tril), four rows per
// inner step.  The per-row vectors are staged TRANSPOSED ([element][row], pitch 36 floats) so that the four rows of a
// step are one LDS.128 per operand; with 8-byte loads and 4-byte bf16x2 stores that is a third of the memory
// instructions of the one-column / one-row form, which was bound by the load/store issue rate.  One kernel per head
// (ROLE 0: d / d par_e, reads the head output; ROLE 1: d / d par_p, a pure write stream): each needs two of the four
// vectors, i.e. half the operand registers and half the shared memory of a combined kernel, and runs at its own pace.
constexpr int kPostThreads = 256, kPostRows = 32, kPostPitch = kPostRows + 4, kPairs = P / 2;
constexpr int kPostBlocksX = (kPairs + D + kPostThreads - 1) / kPostThreads;
template <int ROLE>
__global__ void __launch_bounds__(kPostThreads, 4) heads_bwd64_kernel(
    const float* __restrict__ par_e, const float* __restrict__ eps, const float* __restrict__ dz_dec, int stop_grad,
    const float* __restrict__ vec_r, const float* __restrict__ vec_g, const float* __restrict__ vec_qd,
    const float* __restrict__ g_kl, const float* __restrict__ g_match, __nv_bfloat16* __restrict__ dpar_b,
    float* __restrict__ db, int64_t B) {
  // ROLE 0: sa = dz_total, sb = eps, sw = g_kl;   ROLE 1: sa = g, sb = r, sw = g_match
  // One block = one tile of 32 rows x 512 columns; the five column blocks of a row range are adjacent in launch order,
  // so they run together and share the row vectors and DRAM pages through L2.  Measured and dropped: 64-row tiles
  // (0.66 vs 0.49 ms), 2 / 4 consecutive tiles per block to divide the bias-gradient atomics (0.43 / 0.51 vs 0.40 ms),
  // blocks striding over the whole batch (2.4x slower, 1.7x the DRAM reads).
  __shared__ __align__(16) float sa[D][kPostPitch], sb[D][kPostPitch];
  __shared__ float sd[kPostRows][D];                 // the diagonal threads' inputs (last column block only)
  __shared__ __align__(16) float sw[kPostRows];
  const int64_t r0 = (int64_t)blockIdx.y * kPostRows;
  const int nr = (int)((B - r0 < kPostRows) ? (B - r0) : kPostRows);
  const int t = blockIdx.x * kPostThreads + threadIdx.x;
  const int q0 = 2 * t;
  // The streaming loop below keeps at most eight 8-byte loads per thread in flight, which at DRAM latency is ~3 TB/s
  // for the whole chip (ncu: long-scoreboard stalls).  Ask L2 for the block's whole tile of head outputs up front -
  // prefetches hold no register and no scoreboard - so that the demand loads find it there.
  if (ROLE == 0 && t < kPairs && (threadIdx.x & 15) == 0) {
#pragma unroll 8
    for (int rr = 0; rr < nr; ++rr)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(par_e + (r0 + rr) * P + q0));
  }
  // All loads of the staging phase are issued before the first shared-memory store: a load / store loop made one L2
  // round trip per iteration (the stores may alias the loads as far as the compiler knows), ~5 us per block.
  constexpr int kFill = kPostRows * D / kPostThreads;     // 8 elements per thread
  {
    float va[kFill], vb[kFill], vd[kFill];
    if (threadIdx.x < kPostRows) {
      const float* w = ROLE == 0 ? g_kl : g_match;
      sw[threadIdx.x] = threadIdx.x < nr ? w[r0 + threadIdx.x] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < kFill; ++u) {
      const int e = threadIdx.x + u * kPostThreads;
      const bool ok = e < nr * D;
      if (ROLE == 0) {
        // total gradient into z: the decoder's, minus the matching term's unless it carries a stop_gradient (vae.py:136-138)
        float dzv = (ok && dz_dec) ? dz_dec[r0 * D + e] : 0.f;
        if (!stop_grad && ok) dzv -= g_match[r0 + e / D] * vec_g[r0 * D + e];
        va[u] = dzv;
        vb[u] = ok ? eps[r0 * D + e] : 0.f;
      } else {
        va[u] = ok ? vec_g[r0 * D + e] : 0.f;
        vb[u] = ok ? vec_r[r0 * D + e] : 0.f;
      }
    }
    const bool last = blockIdx.x == kPostBlocksX - 1;
    if (last) {
      // what the diagonal threads below read, fetched by the whole block with every load in flight at once
#pragma unroll
      for (int u = 0; u < kFill; ++u) {
        const int e = threadIdx.x + u * kPostThreads;
        const int row = e / D, i = e % D;
        vd[u] = 0.f;
        if (row < nr) vd[u] = ROLE == 0 ? __ldg(par_e + (r0 + row) * P + D + diag_q(i)) : vec_qd[(r0 + row) * D + i];
      }
    }
#pragma unroll
    for (int u = 0; u < kFill; ++u) {
      const int e = threadIdx.x + u * kPostThreads;
      sa[e % D][e / D] = va[u];
      sb[e % D][e / D] = vb[u];
      if (last) sd[e / D][e % D] = vd[u];
    }
  }
  __syncthreads();
  // The 64 diagonal elements go through softplus / sigmoid.  They sit 65 columns apart, i.e. one in every other warp:
  // handled in line, those warps would pay the transcendental path (one active lane) for every row.  The spare threads
  // past the last pair take them instead, thread = diagonal index.
  if (t >= kPairs) {
    const int i = t - kPairs;
    if (i >= D) return;
    const int qd = D + diag_q(i);
    float acc = 0.f;
#pragma unroll 4
    for (int rr = 0; rr < nr; ++rr) {
      float v;
      if (ROLE == 0) {
        const float raw_e = sd[rr][i];
        const float dg = softplus_f(raw_e) + 1e-5f;
        v = (sa[i][rr] * sb[i][rr] + sw[rr] * (dg - 1.0f / dg)) * sigmoid_f(raw_e);
      } else {
        v = sw[rr] * sd[rr][i];
      }
      const __nv_bfloat16 hv = __float2bfloat16(v);
      dpar_b[(r0 + rr) * P + qd] = hv;
      acc += __bfloat162float(hv);
    }
    if (db) atomicAdd(db + qd, acc);
    return;
  }
  const bool is_loc = q0 < D;
  int i0 = q0, j0 = 0, i1 = q0 + 1, j1 = 0;
  if (!is_loc) { v_to_ij(q0 - D, i0, j0); v_to_ij(q0 + 1 - D, i1, j1); }
  const bool d0 = !is_loc && i0 == j0, d1 = !is_loc && i1 == j1;     // a diagonal slot: left to the spare threads
  float a0 = 0.f, a1 = 0.f;                         // bias-gradient partial sums
  auto el = [](const float4& v, int u) { return u == 0 ? v.x : u == 1 ? v.y : u == 2 ? v.z : v.w; };
  // ncu: this loop is bound by instruction issue and the L1 data stage (31 instructions per thread and row in the
  // first version), so the full-tile case runs without row guards, with one row pointer and immediate offsets
  __nv_bfloat16* const orow = dpar_b + r0 * P + q0;
  const float2* const prow = reinterpret_cast<const float2*>(par_e + r0 * P + q0);
  auto tile = [&](auto full_tag) {
    constexpr bool kFull = decltype(full_tag)::value;
#pragma unroll 2
    for (int rb = 0; rb < kPostRows; rb += 4) {
      float2 raw[4];
      if (ROLE == 0) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
          raw[u] = (kFull || rb + u < nr) ? __ldg(prow + (rb + u) * (P / 2)) : make_float2(0.f, 0.f);
      }
      const float4 w = *reinterpret_cast<const float4*>(&sw[rb]);
      const float4 aA = *reinterpret_cast<const float4*>(&sa[i0][rb]), aB = *reinterpret_cast<const float4*>(&sa[i1][rb]);
      float4 bA = *reinterpret_cast<const float4*>(&sb[j0][rb]), bB = *reinterpret_cast<const float4*>(&sb[j1][rb]);
      if (is_loc) { bA = bB = make_float4(1.f, 1.f, 1.f, 1.f); }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (kFull || rb + u < nr) {
          float v0, v1;
          if (ROLE == 0) {
            v0 = fmaf(el(aA, u), el(bA, u), el(w, u) * raw[u].x);
            v1 = fmaf(el(aB, u), el(bB, u), el(w, u) * raw[u].y);
          } else {
            v0 = el(w, u) * el(aA, u) * el(bA, u);
            v1 = el(w, u) * el(aB, u) * el(bB, u);
          }
          const __nv_bfloat162 hv = __floats2bfloat162_rn(v0, v1);
          __nv_bfloat16* o = orow + (rb + u) * P;
          if (!(d0 | d1)) {
            *reinterpret_cast<__nv_bfloat162*>(o) = hv;
          } else {                       // (both can be diagonal: a forward run of v ends where a reversed one starts)
            if (!d0) o[0] = hv.x;
            if (!d1) o[1] = hv.y;
          }
          a0 += v0; a1 += v1;
        }
      }
    }
  };
  if (nr == kPostRows) tile(std::true_type{}); else tile(std::false_type{});
  if (db) { if (!d0) atomicAdd(db + q0, a0); if (!d1) atomicAdd(db + q0 + 1, a1); }
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
  latent_fwd64_kernel<<<grid_rows(B, kBlocksPerSm), kThreads, kSmemBytes, s>>>(par, eps, z, kl, B);
  PMVAE_LAUNCH_CHECK();
  return 0;
}

int match_fwd64(const float* par_p, const float* z, float* match, int64_t B, cudaStream_t s, float* save_r, float* save_g,
                float* save_qd) {
  using namespace l64;
  static bool attr = false;
  if (!attr) {
    PMVAE_CUDA(cudaFuncSetAttribute(match_fwd64_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    PMVAE_CUDA(cudaFuncSetAttribute(match_fwd64_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    attr = true;
  }
  if (save_r)
    match_fwd64_kernel<true><<<grid_rows(B, kBlocksPerSm), kThreads, kSmemBytes, s>>>(par_p, z, match, save_r, save_g, save_qd, B);
  else
    match_fwd64_kernel<false><<<grid_rows(B, kBlocksPerSm), kThreads, kSmemBytes, s>>>(par_p, z, match, nullptr, nullptr, nullptr, B);
  PMVAE_LAUNCH_CHECK();
  return 0;
}

// r / g / qd: the three [B, 64] vectors match_fwd64 saved for these rows in the forward.
int latent_bwd64(const float* par_e, const float* eps, const float* dz_dec, const float* g_kl, const float* g_match,
                 int stop_grad, __nv_bfloat16* dpar_e_b, __nv_bfloat16* dpar_p_b, float* db_e, float* db_p,
                 const float* vec_r, const float* vec_g, const float* vec_qd, int64_t B, cudaStream_t s) {
  using namespace l64;
  const dim3 grid(kPostBlocksX, (unsigned)((B + kPostRows - 1) / kPostRows));
  heads_bwd64_kernel<0><<<grid, kPostThreads, 0, s>>>(par_e, eps, dz_dec, stop_grad, vec_r, vec_g, vec_qd, g_kl, g_match,
                                                      dpar_e_b, db_e, B);
  PMVAE_LAUNCH_CHECK();
  heads_bwd64_kernel<1><<<grid, kPostThreads, 0, s>>>(par_e, eps, dz_dec, stop_grad, vec_r, vec_g, vec_qd, g_kl, g_match,
                                                      dpar_p_b, db_p, B);
  PMVAE_LAUNCH_CHECK();
  return 0;
}

}  // namespace pmvae
