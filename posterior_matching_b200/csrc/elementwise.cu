// Row-wise and elementwise kernels of the PM-VAE path (all fp32, coalesced).
//
// Reference sites: networks.py:117-118 (LayerNorm), vae.py:127-128 (Normal log-prob
// row-sum), vae.py:132-133 ([x*b, b]), vae.py:222-223 (reduce_logmeanexp),
// vae.py:165-167 (impute where), train_pm_vae.py:58-83 (loss weights, optax chain).
#include <cuda_bf16.h>

#include "kernels.h"

namespace pmvae {

static int grid1d(int64_t work, int block, int per_sm = 8) {
  int64_t g = ceil_div(work, block);
  const int64_t cap = 148ll * per_sm;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// ---------------------------------------------------------------- [x*b, b]
__global__ void __launch_bounds__(256) concat_masked_kernel(const float* __restrict__ x, const float* __restrict__ b,
                                                            float* __restrict__ out, int64_t B, int D) {
  const int64_t n = B * D;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / D;
    const int c = (int)(i - r * D);
    const float bv = b[i];
    out[r * 2 * D + c] = x[i] * bv;
    out[r * 2 * D + D + c] = bv;
  }
}
int concat_masked(const float* x, const float* b, float* out, int64_t B, int D, cudaStream_t s) {
  if (B == 0) return 0;
  concat_masked_kernel<<<grid1d(B * D, 256), 256, 0, s>>>(x, b, out, B, D);
  PMVAE_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------- LayerNorm (warp per row)
__global__ void __launch_bounds__(256) ln_fwd_kernel(float* __restrict__ y, float* __restrict__ rstd,
                                                     const float* __restrict__ resid, float* __restrict__ out_sum,
                                                     int64_t B, int N) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < B; r += nwarps) {
    float* row = y + r * N;
    float s = 0.f;
    for (int c = lane; c < N; c += 32) s += row[c];
    const float mean = warp_sum(s) / N;
    float q = 0.f;
    for (int c = lane; c < N; c += 32) { const float t = row[c] - mean; q += t * t; }
    const float rs = rsqrtf(warp_sum(q) / N + 1e-5f);
    for (int c = lane; c < N; c += 32) {
      const float xh = (row[c] - mean) * rs;
      row[c] = xh;
      if (out_sum) out_sum[r * N + c] = resid[r * N + c] + xh;
    }
    if (lane == 0) rstd[r] = rs;
  }
}
int ln_fwd(float* y, float* rstd, const float* resid, float* out_sum, int64_t B, int N, cudaStream_t s) {
  if (B == 0) return 0;
  ln_fwd_kernel<<<grid1d(B * 32, 256), 256, 0, s>>>(y, rstd, resid, out_sum, B, N);
  PMVAE_LAUNCH_CHECK();
  return 0;
}

__global__ void __launch_bounds__(256) ln_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ xhat,
                                                     const float* __restrict__ rstd, float* __restrict__ dx, int64_t B,
                                                     int N) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < B; r += nwarps) {
    float s1 = 0.f, s2 = 0.f;
    for (int c = lane; c < N; c += 32) {
      const float g = dy[r * N + c];
      s1 += g;
      s2 += g * xhat[r * N + c];
    }
    s1 = warp_sum(s1) / N;
    s2 = warp_sum(s2) / N;
    const float rs = rstd[r];
    for (int c = lane; c < N; c += 32) dx[r * N + c] = rs * (dy[r * N + c] - s1 - xhat[r * N + c] * s2);
  }
}
int ln_bwd(const float* dy, const float* xhat, const float* rstd, float* dx, int64_t B, int N, cudaStream_t s) {
  if (B == 0) return 0;
  ln_bwd_kernel<<<grid1d(B * 32, 256), 256, 0, s>>>(dy, xhat, rstd, dx, B, N);
  PMVAE_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------- column sums (bias grads)
// block (32 x 8): 32 columns, 8 row lanes; grid.x over column groups, grid.y over row chunks
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ dY, int64_t ld, float* __restrict__ out,
                                                     int64_t B, int N) {
  __shared__ float sm[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float acc = 0.f;
  if (c < N)
    for (int64_t r = (int64_t)blockIdx.y * 8 + threadIdx.y; r < B; r += (int64_t)gridDim.y * 8) acc += dY[r * ld + c];
  sm[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && c < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += sm[i][threadIdx.x];
    atomicAdd(out + c, t);
  }
}
int colsum_add(const float* dY, int64_t ld, float* out, int64_t B, int N, cudaStream_t s) {
  if (B == 0) return 0;
  dim3 block(32, 8);
  int64_t gy = ceil_div(B, 8 * 16);
  const int64_t gx = ceil_div(N, 32);
  const int64_t cap = ceil_div(148 * 8, gx);
  if (gy > cap) gy = cap;
  if (gy < 1) gy = 1;
  colsum_kernel<<<dim3((unsigned)gx, (unsigned)gy), block, 0, s>>>(dY, ld, out, B, N);
  PMVAE_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------- Normal log-prob row sums
__global__ void __launch_bounds__(256) rec_ll_kernel(const float* __restrict__ x, const float* __restrict__ loc,
                                                     int64_t ld_loc, const float* __restrict__ log_scale,
                                                     const float* __restrict__ w, float* __restrict__ out, int64_t B,
                                                     int D) {
  const float ls = *log_scale;
  const float inv = expf(-ls);
  const float cst = -ls - 0.5f * kLog2Pi;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < B; r += (int64_t)gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int j = 0; j < D; ++j) {
      const float t = (x[r * D + j] - loc[r * ld_loc + j]) * inv;
      const float ll = -0.5f * t * t + cst;
      acc += w ? ll * w[r * D + j] : ll;
    }
    out[r] = acc;
  }
}
int rec_ll(const float* x, const float* loc, int64_t ld_loc, const float* log_scale, const float* w, float* out,
           int64_t B, int D, cudaStream_t s) {
  if (B == 0) return 0;
  rec_ll_kernel<<<grid1d(B, 256), 256, 0, s>>>(x, loc, ld_loc, log_scale, w, out, B, D);
  PMVAE_LAUNCH_CHECK();
  return 0;
}

// Same VJP without the fused bias gradient, one thread per output element (the padded pitch included): contiguous
// reads and writes per warp for any D (thread-per-row left every access of a warp on its own line: 133 us at D = 63).
__global__ void __launch_bounds__(256) rec_ll_bwd_flat_kernel(const float* __restrict__ x, const float* __restrict__ loc,
                                                              int64_t ld_loc, const float* __restrict__ log_scale,
                                                              const float* __restrict__ g, float* __restrict__ dloc,
                                                              __nv_bfloat16* __restrict__ dloc_bf16, int64_t ld_dloc,
                                                              float* __restrict__ dls, int64_t B, int D) {
  const float inv2 = expf(-2.0f * *log_scale);
  float part = 0.f;
  const int64_t n = B * ld_dloc;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = t / ld_dloc;
    const int j = (int)(t - r * ld_dloc);
    float dv = 0.f;
    if (j < D) {
      const float gr = g[r];
      const float df = x[r * D + j] - loc[r * ld_loc + j];
      dv = gr * df * inv2;
      part += gr * (df * df * inv2 - 1.0f);
      if (dloc) dloc[t] = dv;
    }
    if (dloc_bf16) dloc_bf16[t] = __float2bfloat16(dv);
  }
  part = warp_sum(part);
  __shared__ float sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tsum = 0.f;
    for (int i = 0; i < 8; ++i) tsum += sm[i];
    atomicAdd(dls, tsum);
  }
}

__global__ void __launch_bounds__(256) rec_ll_bwd_kernel(const float* __restrict__ x, const float* __restrict__ loc,
                                                         int64_t ld_loc, const float* __restrict__ log_scale,
                                                         const float* __restrict__ g, float* __restrict__ dloc,
                                                         __nv_bfloat16* __restrict__ dloc_bf16, int64_t ld_dloc,
                                                         float* __restrict__ dls, int64_t B, int D, float* __restrict__ db) {
  __shared__ float dbs[64];
  if (db) { if (threadIdx.x < 64) dbs[threadIdx.x] = 0.f; __syncthreads(); }
  const bool do_db = db != nullptr && D <= 64;
  const float ls = *log_scale;
  const float inv2 = expf(-2.0f * ls);
  float part = 0.f;
  // every warp runs the same number of iterations so that the column sums can be reduced with full-warp shuffles
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t iters = (B + stride - 1) / stride;
  for (int64_t it = 0; it < iters; ++it) {
    const int64_t r = it * stride + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool ok = r < B;
    const float gr = ok ? g[r] : 0.f;
    float acc = 0.f;
    for (int j = 0; j < D; ++j) {
      float hbv = 0.f;
      if (ok) {
        const float df = x[r * D + j] - loc[r * ld_loc + j];
        const float dv = gr * df * inv2;
        if (dloc) dloc[r * ld_dloc + j] = dv;
        if (dloc_bf16) {
          const __nv_bfloat16 hb = __float2bfloat16(dv);
          dloc_bf16[r * ld_dloc + j] = hb;
          hbv = __bfloat162float(hb);
        }
        acc += df * df * inv2 - 1.0f;
      }
      if (do_db) {
        hbv = warp_sum(hbv);
        if ((threadIdx.x & 31) == 0) atomicAdd(&dbs[j], hbv);     // one add per warp and column; flushed once per block
      }
    }
    if (ok && dloc_bf16)
      for (int j = D; j < ld_dloc; ++j) dloc_bf16[r * ld_dloc + j] = __float2bfloat16(0.f);
    part += gr * acc;
  }
  part = warp_sum(part);
  __shared__ float sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += sm[i];
    atomicAdd(dls, t);
  }
  if (do_db && (int)threadIdx.x < D) atomicAdd(db + threadIdx.x, dbs[threadIdx.x]);   // (after the __syncthreads above)
}
int rec_ll_bwd(const float* x, const float* loc, int64_t ld_loc, const float* log_scale, const float* g, float* dloc,
               __nv_bfloat16* dloc_bf16, int64_t ld_dloc, float* dls, int64_t B, int D, cudaStream_t s, float* db) {
  if (B == 0) return 0;
  PMVAE_CHECK(db == nullptr || (D <= 64 && dloc_bf16 != nullptr), "fused decoder-head bias gradient needs D <= 64");
  if (db == nullptr && D > 16)
    rec_ll_bwd_flat_kernel<<<grid1d(B * ld_dloc, 256, 8), 256, 0, s>>>(x, loc, ld_loc, log_scale, g, dloc, dloc_bf16, ld_dloc,
                                                                       dls, B, D);
  else
    rec_ll_bwd_kernel<<<grid1d(B, 256, 4), 256, 0, s>>>(x, loc, ld_loc, log_scale, g, dloc, dloc_bf16, ld_dloc, dls, B, D, db);
  PMVAE_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------- loss weights + batch sums
__global__ void __launch_bounds__(256) loss_cot_kernel(int64_t B, float a, float k, float m,
                                                       const float* __restrict__ rec, const float* __restrict__ kl,
                                                       const float* __restrict__ match, float* __restrict__ g_rec,
                                                       float* __restrict__ g_kl, float* __restrict__ g_match,
                                                       float* __restrict__ sums, const StepState* __restrict__ st) {
  if (st) k = st->beta * (-a);        // a = -1/B_global
  float s0 = 0.f, s1 = 0.f, s2 = 0.f;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < B; r += (int64_t)gridDim.x * blockDim.x) {
    s0 += rec[r]; s1 += kl[r]; s2 += match[r];
    g_rec[r] = a; g_kl[r] = k; g_match[r] = m;
  }
  s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2);
  __shared__ float sm[3][8];
  if ((threadIdx.x & 31) == 0) { sm[0][threadIdx.x >> 5] = s0; sm[1][threadIdx.x >> 5] = s1; sm[2][threadIdx.x >> 5] = s2; }
  __syncthreads();
  if (threadIdx.x < 3) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += sm[threadIdx.x][i];
    atomicAdd(sums + threadIdx.x, t);
  }
}
int loss_cotangents(int64_t B, int64_t B_global, float beta, float coef, const float* rec, const float* kl,
                    const float* match, float* g_rec, float* g_kl, float* g_match, float* out_sums, cudaStream_t s,
                    const StepState* st) {
  PMVAE_CHECK(B_global > 0, "B_global must be positive");
  PMVAE_CUDA(cudaMemsetAsync(out_sums, 0, 3 * sizeof(float), s));
  if (B == 0) return 0;
  const float inv = 1.0f / (float)B_global;
  loss_cot_kernel<<<grid1d(B, 256, 2), 256, 0, s>>>(B, -inv, beta * inv, -coef * inv, rec, kl, match, g_rec, g_kl,
                                                    g_match, out_sums, st);
  PMVAE_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------- AdamW over the flat arena
// Four consecutive elements per thread and iteration (16-byte loads / stores, all four tensors' loads in flight
// together): with one element per iteration the kernel ran at DRAM latency, 0.6 TB/s.  `vec` = the four pointers are
// 16-byte aligned; the tail (n % 4) and unaligned arenas take the scalar path.
__device__ __forceinline__ float adamw_one(float pi, float gi, float& mi, float& vi, bool decay, float lr, float wd,
                                           float b1, float b2, float eps, float bc1, float bc2) {
  mi = b1 * mi + (1.0f - b1) * gi;
  vi = b2 * vi + (1.0f - b2) * gi * gi;
  float u = (mi / bc1) / (sqrtf(vi / bc2) + eps);
  if (decay) u += wd * pi;
  return pi - lr * u;
}
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                    float* __restrict__ m, float* __restrict__ v, uint64_t n,
                                                    AdamSegs nd, float lr, float wd, float b1, float b2, float eps,
                                                    float bc1, float bc2, const StepState* __restrict__ st, int vec) {
  if (st) { lr = st->lr; bc1 = st->bc1; bc2 = st->bc2; }
  // no-decay ranges (the bias leaves) are ascending, disjoint and 64-float aligned at both ends (model.cu::bias_ranges,
  // checked in adamw()): binary search, and one answer holds for a whole aligned group of four elements.  The linear
  // scan per element this replaces was 40 ranges x 2 compares = most of the kernel's instructions.
  auto decays = [&](uint64_t i) {
    int lo = 0, hi = nd.n;                       // first range with beg > i
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (nd.beg[mid] <= i) lo = mid + 1; else hi = mid;
    }
    return !(lo > 0 && i < nd.end[lo - 1]);
  };
  const uint64_t n4 = vec ? n / 4 : 0;
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t q = tid; q < n4; q += nth) {
    const float4 g4 = reinterpret_cast<const float4*>(g)[q];
    float4 m4 = reinterpret_cast<float4*>(m)[q], v4 = reinterpret_cast<float4*>(v)[q], p4 = reinterpret_cast<float4*>(p)[q];
    const bool dk = decays(4 * q);
    p4.x = adamw_one(p4.x, g4.x, m4.x, v4.x, dk, lr, wd, b1, b2, eps, bc1, bc2);
    p4.y = adamw_one(p4.y, g4.y, m4.y, v4.y, dk, lr, wd, b1, b2, eps, bc1, bc2);
    p4.z = adamw_one(p4.z, g4.z, m4.z, v4.z, dk, lr, wd, b1, b2, eps, bc1, bc2);
    p4.w = adamw_one(p4.w, g4.w, m4.w, v4.w, dk, lr, wd, b1, b2, eps, bc1, bc2);
    reinterpret_cast<float4*>(m)[q] = m4;
    reinterpret_cast<float4*>(v)[q] = v4;
    reinterpret_cast<float4*>(p)[q] = p4;
  }
  for (uint64_t i = 4 * n4 + tid; i < n; i += nth) {
    float mi = m[i], vi = v[i];
    p[i] = adamw_one(p[i], g[i], mi, vi, decays(i), lr, wd, b1, b2, eps, bc1, bc2);
    m[i] = mi; v[i] = vi;
  }
}
int adamw(float* p, const float* g, float* m, float* v, uint64_t n, const AdamSegs& nodecay, float lr, float wd,
          float b1, float b2, float eps, float bc1, float bc2, cudaStream_t s, const StepState* st) {
  for (int i = 0; i < nodecay.n; ++i)
    PMVAE_CHECK(nodecay.beg[i] % 4 == 0 && nodecay.end[i] % 4 == 0 && nodecay.beg[i] < nodecay.end[i] &&
                    (i == 0 || nodecay.end[i - 1] <= nodecay.beg[i]),
                "no-decay ranges must be ascending, disjoint and aligned to four elements");
  const int vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                    reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  adamw_kernel<<<grid1d((int64_t)(n / 4 + 1), 256), 256, 0, s>>>(p, g, m, v, n, nodecay, lr, wd, b1, b2, eps, bc1, bc2, st, vec);
  PMVAE_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------- evaluators
__global__ void __launch_bounds__(256) eval_rows_ll_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                           const float* __restrict__ loc, int64_t ld_loc,
                                                           const float* __restrict__ log_scale,
                                                           const float* __restrict__ base, float* __restrict__ out,
                                                           int64_t B, int64_t K, int D) {
  const float ls = *log_scale;
  const float inv = expf(-ls);
  const float cst = -ls - 0.5f * kLog2Pi;
  const int64_t n = B * K;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i % B;
    float acc = 0.f;
    for (int j = 0; j < D; ++j) {
      const float t = (x[r * D + j] - loc[i * ld_loc + j]) * inv;
      const float ll = -0.5f * t * t + cst;
      acc += w ? ll * w[r * D + j] : ll;
    }
    out[i] = acc + base[i];
  }
}
int eval_rows_ll(const float* x, const float* w, const float* loc, int64_t ld_loc, const float* log_scale,
                 const float* base, float* out, int64_t B, int64_t K, int D, cudaStream_t s) {
  if (B * K == 0) return 0;
  eval_rows_ll_kernel<<<grid1d(B * K, 256), 256, 0, s>>>(x, w, loc, ld_loc, log_scale, base, out, B, K, D);
  PMVAE_LAUNCH_CHECK();
  return 0;
}

// block = 32 rows x 32 slices of K: thread (row, slice) covers k = slice, slice + 32, ...
__global__ void __launch_bounds__(1024) logmeanexp_kernel(const float* __restrict__ a, const float* __restrict__ c,
                                                          float* __restrict__ out, int64_t B, int64_t K) {
  __shared__ float red[32][33];
  const float logK = logf((float)K);
  const int rl = threadIdx.x & 31, sl = threadIdx.x >> 5;
  for (int64_t r0 = (int64_t)blockIdx.x * 32; r0 < B; r0 += (int64_t)gridDim.x * 32) {
    const int64_t r = r0 + rl;
    const bool ok = r < B;
    float res[2] = {0.f, 0.f};
    for (int t = 0; t < (c ? 2 : 1); ++t) {
      const float* src = t ? c : a;
      float mx = -INFINITY;
      if (ok) for (int64_t k = sl; k < K; k += 32) mx = fmaxf(mx, src[k * B + r]);
      __syncthreads();
      red[sl][rl] = mx;
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 32; ++i) mx = fmaxf(mx, red[i][rl]);
      const bool fin = mx > -INFINITY && mx < INFINITY;
      float sum = 0.f;
      if (ok && fin) for (int64_t k = sl; k < K; k += 32) sum += expf(src[k * B + r] - mx);
      __syncthreads();
      red[sl][rl] = sum;
      __syncthreads();
      sum = 0.f;
#pragma unroll
      for (int i = 0; i < 32; ++i) sum += red[i][rl];
      res[t] = fin ? (mx + logf(sum) - logK) : mx;
    }
    if (ok && sl == 0) out[r] = res[0] - res[1];
  }
}
int logmeanexp_rows(const float* a, const float* c, float* out, int64_t B, int64_t K, cudaStream_t s) {
  if (B == 0) return 0;
  int64_t g = ceil_div(B, 32);
  if (g > 148 * 2) g = 148 * 2;
  logmeanexp_kernel<<<(int)g, 1024, 0, s>>>(a, c, out, B, K);
  PMVAE_LAUNCH_CHECK();
  return 0;
}

__global__ void __launch_bounds__(256) impute_mean_kernel(const float* __restrict__ x, const float* __restrict__ b,
                                                          const float* __restrict__ loc, int64_t ld_loc,
                                                          float* __restrict__ out, int64_t B, int64_t K, int D) {
  const int64_t n = B * D;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float bv = b[i];
    float acc = 0.f;
    if (bv != 0.f) {
      acc = x[i] * bv;  // where(b, x_o, x_hat): every sample equals x_o
    } else {
      const int64_t r = i / D;
      const int j = (int)(i - r * D);
      for (int64_t k = 0; k < K; ++k) acc += loc[(k * B + r) * ld_loc + j];
      acc /= (float)K;
    }
    out[i] = acc;
  }
}
int impute_mean(const float* x, const float* b, const float* loc, int64_t ld_loc, float* out, int64_t B, int64_t K,
                int D, cudaStream_t s) {
  if (B == 0) return 0;
  impute_mean_kernel<<<grid1d(B * D, 256), 256, 0, s>>>(x, b, loc, ld_loc, out, B, K, D);
  PMVAE_LAUNCH_CHECK();
  return 0;
}

// vae.py:164-165: imputations = where(b == 1, x_o, decoder mean) for every sample k of a chunk of nb data rows:
// out[(k * B_all + r) * D + j] with r counted inside the whole call (out already points at the chunk's first row)
__global__ void __launch_bounds__(256) impute_samples_kernel(const float* __restrict__ x, const float* __restrict__ b,
                                                             const float* __restrict__ loc, int64_t ld_loc,
                                                             float* __restrict__ out, int64_t nb, int64_t B_all, int64_t K,
                                                             int D) {
  const int64_t n = K * nb * D;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t kr = i / D;
    const int j = (int)(i - kr * D);
    const int64_t k = kr / nb, r = kr - k * nb;
    const float bv = b[r * D + j];
    out[(k * B_all + r) * D + j] = bv != 0.f ? x[r * D + j] * bv : loc[kr * ld_loc + j];
  }
}
int impute_samples(const float* x, const float* b, const float* loc, int64_t ld_loc, float* out, int64_t nb, int64_t B_all,
                   int64_t K, int D, cudaStream_t s) {
  if (nb == 0) return 0;
  impute_samples_kernel<<<grid1d(K * nb * D, 256), 256, 0, s>>>(x, b, loc, ld_loc, out, nb, B_all, K, D);
  PMVAE_LAUNCH_CHECK();
  return 0;
}

}  // namespace pmvae
