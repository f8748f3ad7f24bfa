// fp32 FMA-tile GEMM with fused prologue/epilogue: the exact-parity arithmetic of
// PMVAE_PREC_F32 and the on-device check for the tcgen05 path.
//
//   C[M,N] = epi( A'[M,K] @ B'[K,N] )
//   A'(m,k) = TA ? A[k*lda + m] : A[m*lda + k]      (optionally relu'd on load)
//   B'(k,n) = TB ? B[n*ldb + k] : B[k*ldb + n]
//   epi(acc) = ((acc + bias[n]) [* (mask[m,n] > 0)]) + resid[m,n]     or atomicAdd into C
//
// Used for hk.Linear forward (networks.py:116,122,127; distributions.py:44,104), its
// input gradient (dY @ W^T) and its weight gradient (relu(S)^T @ dY, split over the
// batch with fp32 atomics).
#include "kernels.h"

namespace pmvae {

constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4;

template <bool TA, bool TB, bool VEC>
__global__ void __launch_bounds__(256) gemm_f32_kernel(GemmF32Args p) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int tx = tid % 16, ty = tid / 16;  // thread tile: rows ty*4.., cols tx*4..
  // blockIdx.z = batch index * splits + split-K index; a batch member is the same problem at strided pointers
  const int nsplit = (int)gridDim.z / p.batch;
  const int bi = (int)blockIdx.z / nsplit, zi = (int)blockIdx.z - bi * nsplit;
  if (p.batch > 1) {
    p.A += bi * p.sA; p.B += bi * p.sB; p.C += bi * p.sC;
    if (p.bias) p.bias += bi * p.sBias;
    if (p.mask) p.mask += bi * p.sMask;
    if (p.resid) p.resid += bi * p.sResid;
  }
  // split-K range
  const int64_t kchunk = ((p.K + nsplit - 1) / nsplit + BK - 1) / BK * BK;
  const int64_t kbeg = (int64_t)zi * kchunk;
  const int64_t kend = min((int64_t)p.K, kbeg + kchunk);

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;

  for (int64_t k0 = kbeg; k0 < kend; k0 += BK) {
    // ---- A tile -> As[k][m]
    if (!TA) {  // k contiguous: thread -> (row = tid/4, 4 consecutive k)
      const int r = tid / 4, kq = (tid % 4) * 4;
      const int64_t m = m0 + r;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (m < p.M) {
        const float* src = p.A + m * p.lda + k0 + kq;
        if (VEC && k0 + kq + 3 < kend) {
          const float4 t = *reinterpret_cast<const float4*>(src);
          v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) if (k0 + kq + i < kend) v[i] = src[i];
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) As[kq + i][r] = p.relu_a ? fmaxf(v[i], 0.f) : v[i];
    } else {  // m contiguous: thread -> (k = tid/16, 4 consecutive m)
      const int kk = tid / 16, mq = (tid % 16) * 4;
      const int64_t k = k0 + kk;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (k < kend) {
        const float* src = p.A + k * p.lda + m0 + mq;
        if (VEC && m0 + mq + 3 < p.M) {
          const float4 t = *reinterpret_cast<const float4*>(src);
          v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) if (m0 + mq + i < p.M) v[i] = src[i];
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) As[kk][mq + i] = p.relu_a ? fmaxf(v[i], 0.f) : v[i];
    }
    // ---- B tile -> Bs[k][n]
    if (!TB) {  // n contiguous
      const int kk = tid / 16, nq = (tid % 16) * 4;
      const int64_t k = k0 + kk;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (k < kend) {
        const float* src = p.B + k * p.ldb + n0 + nq;
        if (VEC && n0 + nq + 3 < p.N) {
          const float4 t = *reinterpret_cast<const float4*>(src);
          v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) if (n0 + nq + i < p.N) v[i] = src[i];
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) Bs[kk][nq + i] = v[i];
    } else {  // k contiguous
      const int r = tid / 4, kq = (tid % 4) * 4;
      const int64_t n = n0 + r;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (n < p.N) {
        const float* src = p.B + n * p.ldb + k0 + kq;
        if (VEC && k0 + kq + 3 < kend) {
          const float4 t = *reinterpret_cast<const float4*>(src);
          v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) if (k0 + kq + i < kend) v[i] = src[i];
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) Bs[kq + i][r] = v[i];
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int64_t n = n0 + tx * 4 + j;
      if (n >= p.N) continue;
      float v = acc[i][j];
      if (p.atomic) {
        atomicAdd(p.C + m * p.ldc + n, v);
        continue;
      }
      if (p.bias) v += p.bias[n];
      if (p.mask) v = (p.mask[m * p.ldmask + n] > 0.f) ? v : 0.f;
      if (p.resid) v += p.resid[m * p.ldresid + n];
      p.C[m * p.ldc + n] = v;
    }
  }
}

int gemm_f32(const GemmF32Args& a, bool ta, bool tb, cudaStream_t stream) {
  PMVAE_CHECK(a.M >= 0 && a.N > 0 && a.K >= 0, "bad gemm shape");
  if (a.M == 0) return 0;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  PMVAE_CHECK(a.batch >= 1, "bad batch count");
  const bool vec = al16(a.A) && al16(a.B) && a.lda % 4 == 0 && a.ldb % 4 == 0 &&
                   (a.batch == 1 || (a.sA % 4 == 0 && a.sB % 4 == 0));
  int split = a.atomic ? a.split_k : 1;
  if (split < 1) split = 1;
  PMVAE_CHECK((int64_t)split * a.batch <= 65535, "batch x split-K exceeds the grid's z extent");
  dim3 grid((unsigned)ceil_div(a.N, BN), (unsigned)ceil_div(a.M, BM), (unsigned)(split * a.batch));
  PMVAE_CHECK(grid.y <= 65535u * 1u || true, "");
  // blockIdx.y is limited to 65535: M up to 4.1M rows per call
  PMVAE_CHECK(ceil_div(a.M, BM) <= 65535, "M too large for one gemm_f32 launch");
#define LAUNCH(TA_, TB_)                                                                  \
  do {                                                                                    \
    if (vec) gemm_f32_kernel<TA_, TB_, true><<<grid, 256, 0, stream>>>(a);                \
    else gemm_f32_kernel<TA_, TB_, false><<<grid, 256, 0, stream>>>(a);                   \
  } while (0)
  if (!ta && !tb) LAUNCH(false, false);
  else if (!ta && tb) LAUNCH(false, true);
  else if (ta && !tb) LAUNCH(true, false);
  else LAUNCH(true, true);
#undef LAUNCH
  PMVAE_LAUNCH_CHECK();
  return 0;
}

}  // namespace pmvae
