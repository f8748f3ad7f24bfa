// Launch interface of the tcgen05 GEMM (tc_gemm.cu).
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace pmvae {
namespace tc {

// Epilogue of the NT kernel, per element (row m, column n):
//   v = acc + bias[n];  if mask: v = mask[m,n] > 0 ? v : 0;  v += resid_f32[m,n] + resid_bf16[m,n];
//   out_f32[m,n] = v;  out_bf16[m,n] = bf16(relu_out ? max(v, 0) : v)
// (every pointer optional).  The TN kernel only uses out_f32 / atomic.
struct TcGemmArgs {
  // filled by the launcher
  int M = 0, N = 0, K = 0, n_tile = 0, num_m_tiles = 0, num_n_tiles = 0, num_k_blocks = 0, split_k = 1,
      kb_per_split = 0, atomic = 0;
  // epilogue
  const float* bias = nullptr;
  const __nv_bfloat16* mask_bf16 = nullptr; int64_t ld_mask = 0;
  const float* resid_f32 = nullptr; int64_t ld_resid_f32 = 0;
  const __nv_bfloat16* resid_bf16 = nullptr; int64_t ld_resid_bf16 = 0;
  float* out_f32 = nullptr; int64_t ld_out_f32 = 0;
  __nv_bfloat16* out_bf16 = nullptr; int64_t ld_out_bf16 = 0;
  int relu_out = 0;
  float* colsum_out = nullptr;   // += column sums of the bf16 output (needs a single N tile)
  int debug = 0;                 // PMVAE_TC_DEBUG bits (profiling only): 1 = no epilogue global I/O, 2 = no epilogue work
};

int gemm_nt(const __nv_bfloat16* A, int64_t lda, const __nv_bfloat16* Bt, int64_t ldb, int64_t M, int N, int K,
            TcGemmArgs ep, cudaStream_t s);
int gemm_tn(const __nv_bfloat16* A, int64_t lda, const __nv_bfloat16* B, int64_t ldb, int M, int N, int64_t rows,
            float* out, int64_t ld_out, int atomic, int max_split, int* split_out, cudaStream_t s);

// One launch for several weight gradients: out_i[M_i, N_i] (fp32, pitch ld_out, atomically accumulated) +=
// A_i[rows, M_i]^T . B_i[rows, N_i]; up to 8 problems, N_i <= 256.
struct TnDesc {
  const __nv_bfloat16* A; int64_t lda;
  const __nv_bfloat16* B; int64_t ldb;
  int M, N; int64_t rows;
  float* out; int64_t ld_out;
};
int gemm_tn_grouped(const TnDesc* d, int count, cudaStream_t s);

}  // namespace tc
}  // namespace pmvae
