// Launch interface of the fused ResidualMLP kernels (fused_mlp.cu).
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"
#include "model.h"

namespace pmvae {
namespace fused {

// bf16 operand images of one ResidualMLP + its head Linear, laid out for the fused kernels:
//   stack_t : [(1 + 2R) * 256, 256]  slab 0 = first Linear (transposed, input columns expanded to the
//             hi/lo operand layout below), slab l = W_l^T of hidden Linear l
//   head_t  : [head_tiles * head_NT, 256]   W_head^T, rows >= N zero
//   stack_n : [2R * 256, 256]        W_l (natural [in, out]) of hidden Linear l = 1..2R   (input gradients)
//   head_n  : [256, head_Kp]         W_head (natural), columns >= N zero
// First-layer operand layout (columns of the A operand, K_ext <= 256):
//   in_kind 0: [hi(v) (D) | lo(v) (D)]              v = x or z,   hi = bf16(v), lo = bf16(v - hi)
//   in_kind 1: [hi(x*b) (D) | lo(x*b) (D) | b (D)]  the masked input [x*b, b] of vae.py:132-133
struct NetImages {
  const __nv_bfloat16* stack_t; const __nv_bfloat16* head_t;
  const __nv_bfloat16* stack_n; const __nv_bfloat16* head_n;
  const __nv_bfloat16* w0_n;            // [din_N, 256] first Linear (natural), in_kind 0 only: dL/d(input)
  bool has_w0_n; int din_N;
  int R, D_in, in_kind, in_lo, k16_0;   // k16_0 = ceil(K_ext / 16); in_lo: the [lo(v)] columns are present
  int in_wide;                          // masked input with 3 D > 64: first Linear as two K-blocks, [x*b] @ W[:D] + b @ W[D:]
  int head_N, head_NT, head_tiles, head_Kp;
  uint64_t elems;                       // bf16 elements used by the four images
};

// forward_supported: the fused forward covers this net (H = 256, expanded fan-in <= 64; LayerNorm nets forward only);
// supported: forward with saved activations + backward (no LayerNorm)
bool forward_supported(const Net& n, int H, int in_kind);
bool supported(const Net& n, int H, int in_kind);
bool ln_train_supported(const Net& n, int H, int in_kind);

// plans the images at `base` (may be null to size only)
NetImages plan_images(const Net& n, const Leaf& head, int in_kind, __nv_bfloat16* base);
// one launch: refreshes all four images from the float32 parameter arena
int pack_images(const float* params, const Net& n, const Leaf& head, const NetImages& im, cudaStream_t s);

// Forward of one ResidualMLP + head over B rows (networks.py:111-135 + the hk.Linear of the
// distribution head).  Activations stay on chip; when `saved` is given, relu(h_0), relu(linear1_r),
// relu(h_{r+1}) (bf16, slab l of [(2R+1), Bpad, 256]) are also streamed to HBM for the backward, with
// one relu bit per element in `masks` ([(2R+1), Bpad, 8] words; word 4*half + j of a row covers columns
// [64 j + 32 half, +32)).
//   in/msk : [B, D_in] float32 (msk only for in_kind 1)
//   out    : [B, ld_out] float32 head output (columns < head_N written)
// LayerNorm nets in training mode additionally keep xhat of every LayerNorm (`xhat`, bf16, same slab layout as
// `saved`) and 1 / sigma (`rstd`, [(2R+1), Bpad] float32).
int net_forward(const float* params, const Net& n, const Leaf& head, const NetImages& im, const float* in,
                const float* msk, int64_t B, __nv_bfloat16* saved, uint32_t* masks, int64_t Bpad, float* out,
                int64_t ld_out, cudaStream_t s, __nv_bfloat16* xhat = nullptr, float* rstd = nullptr);

// Input-gradient chain of the same net (the activations' VJP): reads dHead [B, ld_dhead] (bf16, columns >=
// head_N zero) and the relu bits of the forward, writes dY_l (bf16, slab l of [(2R+1), Bpad, 256]) = the
// gradient with respect to the output of Linear l (operand of its weight-gradient GEMM), adds the bias
// gradients of Linear 0..2R into `grads`, and optionally dL/d(input) [B, in_dim] (fp32).
bool backward_supported(const Net& n, int H, int in_kind);
int net_backward(const Net& n, const Leaf& head, const NetImages& im, const __nv_bfloat16* dHead, int64_t ld_dhead,
                 int64_t B, const uint32_t* masks, int64_t Bpad, __nv_bfloat16* dY, float* grads, float* dIn,
                 cudaStream_t s);

// The same chain for LayerNorm nets (VJP of the training-mode LayerNorm forward): additionally reads xhat / 1/sigma of
// every LayerNorm; dY_l is the gradient with respect to the (pre-LayerNorm) output of Linear l.  Any head width.
bool backward_ln_supported(const Net& n, int H, int in_kind);
int net_backward_ln(const Net& n, const Leaf& head, const NetImages& im, const __nv_bfloat16* dHead, int64_t ld_dhead,
                    int64_t B, const uint32_t* masks, const __nv_bfloat16* xhat, const float* rstd, int64_t Bpad,
                    __nv_bfloat16* dY, float* grads, float* dIn, cudaStream_t s);

}  // namespace fused
}  // namespace pmvae
