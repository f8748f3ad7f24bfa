// bf16-operand tcgen05/TMEM path (placeholder until the tensor kernels land).
#include "kernels.h"
#include "model.h"

namespace pmvae {
#define NOT_BUILT() do { set_error("PMVAE_PREC_BF16 path is not built yet"); return 3; } while (0)
int linear_bf16(const float*, const float*, const float*, int64_t, int, int, int, float*, void*, uint64_t,
                cudaStream_t) { NOT_BUILT(); }
uint64_t workspace_bytes_bf16(const pmvae_config*, int64_t, int64_t) { return 0; }
int prepare_params_bf16(const pmvae_config*, const float*, void*, uint64_t, cudaStream_t) { NOT_BUILT(); }
int forward_bf16(const pmvae_config*, const Layout&, const float*, const float*, const float*, const float*, int64_t,
                 float*, float*, float*, void*, uint64_t, cudaStream_t) { NOT_BUILT(); }
int backward_bf16(const pmvae_config*, const Layout&, const float*, const float*, const float*, const float*, int64_t,
                  const float*, const float*, const float*, float*, void*, uint64_t, cudaStream_t) { NOT_BUILT(); }
int is_log_prob_bf16(const pmvae_config*, const Layout&, const float*, const float*, const float*, int64_t, int64_t,
                     const uint32_t*, const uint32_t*, int64_t, int64_t, float*, float*, void*, uint64_t,
                     cudaStream_t) { NOT_BUILT(); }
int impute_mean_bf16(const pmvae_config*, const Layout&, const float*, const float*, const float*, int64_t, int64_t,
                     const uint32_t*, int64_t, int64_t, float*, void*, uint64_t, cudaStream_t) { NOT_BUILT(); }
}  // namespace pmvae
