// Shared device/host helpers for libpmvae (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/pmvae.h"

namespace pmvae {

// ---------------------------------------------------------------- errors
void set_error(const std::string& msg);
#define PMVAE_CHECK(cond, msg)                                                        \
  do {                                                                                \
    if (!(cond)) {                                                                    \
      ::pmvae::set_error(std::string(__FILE__) + ":" + std::to_string(__LINE__) + ": " + (msg)); \
      return 1;                                                                       \
    }                                                                                 \
  } while (0)
#define PMVAE_CUDA(call)                                                              \
  do {                                                                                \
    cudaError_t e__ = (call);                                                         \
    if (e__ != cudaSuccess) {                                                         \
      ::pmvae::set_error(std::string(__FILE__) + ":" + std::to_string(__LINE__) + ": " + \
                         cudaGetErrorString(e__));                                    \
      return 2;                                                                       \
    }                                                                                 \
  } while (0)
void count_launch();
#define PMVAE_LAUNCH_CHECK()            \
  do {                                  \
    ::pmvae::count_launch();            \
    PMVAE_CUDA(cudaPeekAtLastError());  \
  } while (0)
#define PMVAE_TRY(call)        \
  do {                         \
    int r__ = (call);          \
    if (r__ != 0) return r__;  \
  } while (0)

static inline cudaStream_t as_stream(pmvae_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
__host__ __device__ static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline uint64_t align_up(uint64_t a, uint64_t b) { return (a + b - 1) / b * b; }

constexpr float kLog2Pi = 1.8378770664093453f;

// ---------------------------------------------------------------- threefry2x32 (jax.random)
struct Key2 { uint32_t k0, k1; };

__host__ __device__ __forceinline__ uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

__host__ __device__ __forceinline__ void threefry2x32(uint32_t k0, uint32_t k1, uint32_t& x0, uint32_t& x1) {
  const uint32_t ks0 = k0, ks1 = k1, ks2 = k0 ^ k1 ^ 0x1BD11BDAu;
  x0 += ks0; x1 += ks1;
#define TF_R(r) { x0 += x1; x1 = rotl32(x1, r); x1 ^= x0; }
  TF_R(13) TF_R(15) TF_R(26) TF_R(6)   x0 += ks1; x1 += ks2 + 1u;
  TF_R(17) TF_R(29) TF_R(16) TF_R(24)  x0 += ks2; x1 += ks0 + 2u;
  TF_R(13) TF_R(15) TF_R(26) TF_R(6)   x0 += ks0; x1 += ks1 + 3u;
  TF_R(17) TF_R(29) TF_R(16) TF_R(24)  x0 += ks1; x1 += ks2 + 4u;
  TF_R(13) TF_R(15) TF_R(26) TF_R(6)   x0 += ks2; x1 += ks0 + 5u;
#undef TF_R
}

// Element i of jax's random_bits(key, n) (counters iota(n) padded to even, split in
// halves; one threefry call per pair; out = concat(first words, second words)).
__host__ __device__ __forceinline__ uint32_t jax_random_word(Key2 key, uint64_t n, uint64_t i) {
  const uint64_t h = (n + (n & 1ull)) >> 1;
  uint32_t x0, x1;
  if (i < h) {
    uint64_t c1 = h + i;
    x0 = (uint32_t)i; x1 = (c1 < n) ? (uint32_t)c1 : 0u;
    threefry2x32(key.k0, key.k1, x0, x1);
    return x0;
  }
  x0 = (uint32_t)(i - h); x1 = (uint32_t)i;
  threefry2x32(key.k0, key.k1, x0, x1);
  return x1;
}

__host__ __device__ __forceinline__ float bits_to_unit_float(uint32_t bits) {
  // jax.random.uniform: bitcast((bits >> 9) | 0x3F800000) - 1  in [0, 1)
#ifdef __CUDA_ARCH__
  return __uint_as_float((bits >> 9) | 0x3F800000u) - 1.0f;
#else
  union { uint32_t u; float f; } c; c.u = (bits >> 9) | 0x3F800000u; return c.f - 1.0f;
#endif
}

#ifdef __CUDACC__
// XLA's float32 ErfInv (Giles' polynomial); w = -log1p(-x*x).
__device__ __forceinline__ float erfinv_giles(float x) {
  float w = -log1pf(-x * x);
  float p;
  if (w < 5.0f) {
    w = w - 2.5f;
    p = 2.81022636e-08f;
    p = fmaf(p, w, 3.43273939e-07f);
    p = fmaf(p, w, -3.5233877e-06f);
    p = fmaf(p, w, -4.39150654e-06f);
    p = fmaf(p, w, 0.00021858087f);
    p = fmaf(p, w, -0.00125372503f);
    p = fmaf(p, w, -0.00417768164f);
    p = fmaf(p, w, 0.246640727f);
    p = fmaf(p, w, 1.50140941f);
  } else {
    w = sqrtf(w) - 3.0f;
    p = -0.000200214257f;
    p = fmaf(p, w, 0.000100950558f);
    p = fmaf(p, w, 0.00134934322f);
    p = fmaf(p, w, -0.00367342844f);
    p = fmaf(p, w, 0.00573950773f);
    p = fmaf(p, w, -0.0076224613f);
    p = fmaf(p, w, 0.00943887047f);
    p = fmaf(p, w, 1.00167406f);
    p = fmaf(p, w, 2.83297682f);
  }
  return p * x;
}

// jax.random.normal float32: sqrt(2) * erfinv(uniform(lo = nextafter(-1, 0), hi = 1))
__device__ __forceinline__ float bits_to_normal(uint32_t bits) {
  const float lo = -0.99999994f;  // nextafter(-1f, 0f)
  float u = bits_to_unit_float(bits);
  u = __fadd_rn(__fmul_rn(u, __fsub_rn(1.0f, lo)), lo);
  u = fmaxf(lo, u);
  return 1.41421356237309515f * erfinv_giles(u);
}

__device__ __forceinline__ float softplus_f(float x) {
  // logaddexp(x, 0), stable on both sides
  return fmaxf(x, 0.0f) + log1pf(expf(-fabsf(x)));
}
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
#endif  // __CUDACC__

}  // namespace pmvae
