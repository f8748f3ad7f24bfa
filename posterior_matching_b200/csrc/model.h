// Parameter-arena layout and the entry points of the two arithmetic paths.
#pragma once
#include "common.cuh"

namespace pmvae {

constexpr int kMaxBlocks = 8;

struct Leaf { int rows, cols; uint64_t w, b; };            // offsets in floats
struct Net { int in_dim, R, ln; Leaf lin[2 * kMaxBlocks + 1]; };
struct Layout {
  Net enc, dec, part;
  Leaf post, ddist, ppost;
  uint64_t log_scale, total;
  int P;
};

int build_layout(const pmvae_config* c, Layout* L);
const char* last_error();

// bf16-operand tcgen05 path (tensor.cu)
int linear_bf16(const float* x, const float* w, const float* bias, int64_t B, int K, int N, int relu_in, float* y,
                void* ws, uint64_t ws_bytes, cudaStream_t s);
uint64_t workspace_bytes_bf16(const pmvae_config* c, int64_t B, int64_t K);
int prepare_params_bf16(const pmvae_config* c, const float* params, void* ws, uint64_t ws_bytes, cudaStream_t s);
int forward_bf16(const pmvae_config* c, const Layout& L, const float* params, const float* x, const float* b,
                 const float* eps, int64_t B, float* out_rec, float* out_kl, float* out_match, void* ws,
                 uint64_t ws_bytes, cudaStream_t s);
int backward_bf16(const pmvae_config* c, const Layout& L, const float* params, const float* x, const float* b,
                  const float* eps, int64_t B, const float* g_rec, const float* g_kl, const float* g_match,
                  float* grads, int stages, void* ws, uint64_t ws_bytes, cudaStream_t s);
// backward of the whole model in stages (bit 0: decoder + latent algebra, bit 1: encoder, bit 2: partial encoder)
int backward_staged(const pmvae_config* c, const float* params, const float* x, const float* b, const float* eps, int64_t B,
                    const float* g_rec, const float* g_kl, const float* g_match, float* grads, int stages, void* ws,
                    uint64_t ws_bytes, cudaStream_t s);
int net_apply_bf16(const pmvae_config* c, const Layout& L, const float* params, int which, const float* in,
                   const float* msk, int64_t B, float* out, bool save, void* ws, uint64_t ws_bytes, cudaStream_t s);
int is_log_prob_bf16(const pmvae_config* c, const Layout& L, const float* params, const float* x, const float* b,
                     int64_t B, int64_t K, const uint32_t key_z[2], const uint32_t key_zxo[2], int64_t B_total,
                     int64_t row_start, float* out_log_p_x, float* out_cond, void* ws, uint64_t ws_bytes,
                     cudaStream_t s);
int impute_mean_bf16(const pmvae_config* c, const Layout& L, const float* params, const float* x, const float* b,
                     int64_t B, int64_t K, const uint32_t key[2], int64_t B_total, int64_t row_start, float* out,
                     float* out_samples, void* ws, uint64_t ws_bytes, cudaStream_t s);

}  // namespace pmvae
