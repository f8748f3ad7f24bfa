// jax.random-compatible threefry2x32 streams and the mask generators, on device.
//
// Reference sites: posterior_matching/models/vae.py:124,162,192-195 (sample keys);
// posterior_matching/masking.py:84-91 (BernoulliMaskGenerator), :24-47,94-174,235-249
// (MNISTMaskGenerator = mixture of ImageBernoulli / 4 FixedRectangle / Square /
// Rectangle).  The reference draws masks from unseeded host MT19937 streams
// (SURVEY F3); the device contract re-defines them on the JAX threefry stream and is
// specified by oracle/prng.py + oracle/masks.py, which these kernels match bit for bit.
#include "common.cuh"
#include "kernels.h"

namespace pmvae {

// ---------------------------------------------------------------- host key utilities
static void host_random_bits(Key2 key, uint64_t n, uint32_t* out) {
  for (uint64_t i = 0; i < n; ++i) out[i] = jax_random_word(key, n, i);
}

// ---------------------------------------------------------------- bulk kernels
// kind: 0 raw bits, 1 uniform [0,1), 2 normal.  Each thread produces elements
// start + t (one threefry call each: the slice interface does not let a thread own
// both words of a pair in general; the full-range fast path below does).
template <int KIND>
__global__ void __launch_bounds__(256) rng_slice_kernel(Key2 key, uint64_t n_total, uint64_t start,
                                                        uint64_t count, void* out_) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < count; t += stride) {
    const uint32_t w = jax_random_word(key, n_total, start + t);
    if (KIND == 0) reinterpret_cast<uint32_t*>(out_)[t] = w;
    else if (KIND == 1) reinterpret_cast<float*>(out_)[t] = bits_to_unit_float(w);
    else reinterpret_cast<float*>(out_)[t] = bits_to_normal(w);
  }
}

// Full draw: thread j owns the pair (j, h + j) and writes both output words, so each
// threefry call yields two elements (2x fewer integer ops than the slice kernel).
template <int KIND>
__global__ void __launch_bounds__(256) rng_full_kernel(Key2 key, uint64_t n, void* out_) {
  const uint64_t h = (n + (n & 1ull)) >> 1;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < h; j += stride) {
    uint32_t x0 = (uint32_t)j, x1 = (h + j < n) ? (uint32_t)(h + j) : 0u;
    threefry2x32(key.k0, key.k1, x0, x1);
    if (KIND == 0) {
      reinterpret_cast<uint32_t*>(out_)[j] = x0;
      if (h + j < n) reinterpret_cast<uint32_t*>(out_)[h + j] = x1;
    } else if (KIND == 1) {
      reinterpret_cast<float*>(out_)[j] = bits_to_unit_float(x0);
      if (h + j < n) reinterpret_cast<float*>(out_)[h + j] = bits_to_unit_float(x1);
    } else {
      reinterpret_cast<float*>(out_)[j] = bits_to_normal(x0);
      if (h + j < n) reinterpret_cast<float*>(out_)[h + j] = bits_to_normal(x1);
    }
  }
}

static int grid_for(uint64_t work, int block) {
  uint64_t g = (work + block - 1) / block;
  const uint64_t cap = 148ull * 16ull;  // 16 resident 256-thread CTAs per SM
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

template <int KIND>
static int launch_rng(const uint32_t key[2], uint64_t n_total, uint64_t start, uint64_t count, void* out,
                      pmvae_stream_t stream) {
  PMVAE_CHECK(key != nullptr && (out != nullptr || count == 0), "null pointer");
  PMVAE_CHECK(n_total <= 0xFFFFFFFFull, "a single draw is limited to 2^32-1 elements (uint32 counters)");
  PMVAE_CHECK(start + count <= n_total, "slice out of range");
  if (count == 0) return 0;
  Key2 k{key[0], key[1]};
  if (start == 0 && count == n_total) {
    const uint64_t h = (n_total + 1) / 2;
    rng_full_kernel<KIND><<<grid_for(h, 256), 256, 0, as_stream(stream)>>>(k, n_total, out);
  } else {
    rng_slice_kernel<KIND><<<grid_for(count, 256), 256, 0, as_stream(stream)>>>(k, n_total, start, count, out);
  }
  PMVAE_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------- Bernoulli mask
// b[r, c] = uniform(key, [B_total, D])[r, c] < p, 1 = observed.
__global__ void __launch_bounds__(256) mask_bernoulli_kernel(Key2 key, float p, uint64_t n_total, uint64_t start,
                                                             uint64_t count, float* out) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < count; t += stride) {
    const float u = bits_to_unit_float(jax_random_word(key, n_total, start + t));
    out[t] = (u < p) ? 1.0f : 0.0f;
  }
}

// same draws with the key read from the device step state (fused training step)
__global__ void __launch_bounds__(256) mask_bernoulli_dev_kernel(const StepState* __restrict__ st, float p, uint64_t n_total,
                                                                 uint64_t start, uint64_t count, float* out) {
  const Key2 key{st->mask_key[0], st->mask_key[1]};
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < count; t += stride) {
    const float u = bits_to_unit_float(jax_random_word(key, n_total, start + t));
    out[t] = (u < p) ? 1.0f : 0.0f;
  }
}
__global__ void __launch_bounds__(256) normal_dev_kernel(const StepState* __restrict__ st, uint64_t n_total, uint64_t start,
                                                         uint64_t count, float* out) {
  const Key2 key{st->eps_key[0], st->eps_key[1]};
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < count; t += stride)
    out[t] = bits_to_normal(jax_random_word(key, n_total, start + t));
}
int mask_bernoulli_dev(const StepState* st, float p, uint64_t B_total, uint64_t row_start, uint64_t rows, int D, float* out,
                       cudaStream_t s) {
  const uint64_t n_total = B_total * (uint64_t)D;
  PMVAE_CHECK(n_total <= 0xFFFFFFFFull && row_start + rows <= B_total, "bad mask draw");
  if (rows == 0) return 0;
  const uint64_t count = rows * (uint64_t)D;
  mask_bernoulli_dev_kernel<<<grid_for(count, 256), 256, 0, s>>>(st, p, n_total, row_start * (uint64_t)D, count, out);
  PMVAE_LAUNCH_CHECK();
  return 0;
}
int normal_dev(const StepState* st, uint64_t n_total, uint64_t start, uint64_t count, float* out, cudaStream_t s) {
  PMVAE_CHECK(n_total <= 0xFFFFFFFFull && start + count <= n_total, "bad normal draw");
  if (count == 0) return 0;
  normal_dev_kernel<<<grid_for(count, 256), 256, 0, s>>>(st, n_total, start, count, out);
  PMVAE_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------- MNIST mixture mask
// Contract (oracle/masks.py::mnist_mask): k_cat, k_bern, k_sq, k_rect = split(key, 4)
//   cat[r]  = choice(k_cat, 7, [B], p = [2,1,1,1,1,2,2]/10)
//   cat 0   : bernoulli(k_bern, 0.5, [B,28,28,1])[r]
//   cat 1-4 : zero [y1:y2, x1:x2] for (0,0,28,14) (0,0,14,28) (0,14,28,28) (14,0,28,28)
//   cat 5   : (x, y) = randint(k_sq, [B,2], 0, 14)[r]; zero [y:y+14, x:x+14]
//   cat 6   : attempt t = 0,1,..: (x1,x2,y1,y2) = randint(fold_in(k_rect,t), [B,4], 0, 28)[r],
//             sorted pairs; accept first t with 0.3*784 <= (x2-x1+1)(y2-y1+1) <= 784;
//             zero [y1:y2+1, x1:x2+1]
struct MnistMaskKeys {
  Key2 cat, bern, sq_hi, sq_lo, rect;
  float cum[7];
};

__device__ __forceinline__ uint32_t jax_randint_word(Key2 khi, Key2 klo, uint64_t n, uint64_t i, uint32_t span) {
  const uint32_t hi = jax_random_word(khi, n, i), lo = jax_random_word(klo, n, i);
  uint32_t mult = 65536u % span;
  mult = (mult * mult) % span;
  return ((hi % span) * mult + (lo % span)) % span;
}

constexpr int kMnistMaxAttempts = 4096;

__global__ void __launch_bounds__(256) mask_mnist_kernel(MnistMaskKeys K, uint64_t B_total, uint64_t row_start,
                                                         uint64_t rows, float* out) {
  __shared__ int s_cat, s_x1, s_x2, s_y1, s_y2;
  for (uint64_t rr = blockIdx.x; rr < rows; rr += gridDim.x) {
    const uint64_t r = row_start + rr;
    __syncthreads();
    if (threadIdx.x == 0) {
      const float u = bits_to_unit_float(jax_random_word(K.cat, B_total, r));
      const float q = __fmul_rn(K.cum[6], __fsub_rn(1.0f, u));
      int c = 0;
      while (c < 7 && K.cum[c] < q) ++c;  // searchsorted(cum, q, side='left')
      int x1 = 0, x2 = -1, y1 = 0, y2 = -1;  // zeroed block, inclusive corners
      if (c == 1) { y1 = 0; x1 = 0; y2 = 27; x2 = 13; }
      else if (c == 2) { y1 = 0; x1 = 0; y2 = 13; x2 = 27; }
      else if (c == 3) { y1 = 0; x1 = 14; y2 = 27; x2 = 27; }
      else if (c == 4) { y1 = 14; x1 = 0; y2 = 27; x2 = 27; }
      else if (c == 5) {
        const int x = (int)jax_randint_word(K.sq_hi, K.sq_lo, B_total * 2, r * 2 + 0, 14u);
        const int y = (int)jax_randint_word(K.sq_hi, K.sq_lo, B_total * 2, r * 2 + 1, 14u);
        x1 = x; x2 = x + 13; y1 = y; y2 = y + 13;
      } else if (c == 6) {
        x1 = 0; x2 = 27; y1 = 0; y2 = 27;  // fallback if no attempt is accepted
        for (int t = 0; t < kMnistMaxAttempts; ++t) {
          uint32_t f0 = 0u, f1 = (uint32_t)t;
          threefry2x32(K.rect.k0, K.rect.k1, f0, f1);  // fold_in(k_rect, t)
          const Key2 kt{f0, f1};
          // split(kt) -> (k_hi, k_lo): random_bits(kt, 4) reshaped [2,2]
          uint32_t a0 = 0u, a1 = 2u, b0 = 1u, b1 = 3u;
          threefry2x32(kt.k0, kt.k1, a0, a1);
          threefry2x32(kt.k0, kt.k1, b0, b1);
          const Key2 khi{a0, b0}, klo{a1, b1};
          const int c0 = (int)jax_randint_word(khi, klo, B_total * 4, r * 4 + 0, 28u);
          const int c1 = (int)jax_randint_word(khi, klo, B_total * 4, r * 4 + 1, 28u);
          const int c2 = (int)jax_randint_word(khi, klo, B_total * 4, r * 4 + 2, 28u);
          const int c3 = (int)jax_randint_word(khi, klo, B_total * 4, r * 4 + 3, 28u);
          const int xa = min(c0, c1), xb = max(c0, c1), ya = min(c2, c3), yb = max(c2, c3);
          const int area = (xb - xa + 1) * (yb - ya + 1);
          // 0.3 * 784 = 235.2 -> area >= 236 ; area <= 784 always
          if (area * 10 >= 2352) { x1 = xa; x2 = xb; y1 = ya; y2 = yb; break; }
        }
      }
      s_cat = c; s_x1 = x1; s_x2 = x2; s_y1 = y1; s_y2 = y2;
    }
    __syncthreads();
    const int c = s_cat, x1 = s_x1, x2 = s_x2, y1 = s_y1, y2 = s_y2;
    float* o = out + rr * 784;
    for (int px = threadIdx.x; px < 784; px += blockDim.x) {
      float v;
      if (c == 0) {
        const uint32_t w = jax_random_word(K.bern, B_total * 784, r * 784 + px);
        v = (bits_to_unit_float(w) < 0.5f) ? 1.0f : 0.0f;
      } else {
        const int y = px / 28, x = px % 28;
        v = (y >= y1 && y <= y2 && x >= x1 && x <= x2) ? 0.0f : 1.0f;
      }
      o[px] = v;
    }
  }
}

}  // namespace pmvae

using namespace pmvae;

extern "C" {

int pmvae_key_split_host(const uint32_t key[2], int n, uint32_t* out_keys) {
  PMVAE_CHECK(key && out_keys && n >= 0, "bad arguments");
  host_random_bits(Key2{key[0], key[1]}, 2ull * (uint64_t)n, out_keys);
  return 0;
}

int pmvae_key_fold_in_host(const uint32_t key[2], uint32_t data, uint32_t out_key[2]) {
  PMVAE_CHECK(key && out_key, "bad arguments");
  uint32_t x0 = 0u, x1 = data;
  threefry2x32(key[0], key[1], x0, x1);
  out_key[0] = x0; out_key[1] = x1;
  return 0;
}

int pmvae_random_bits(const uint32_t key[2], uint64_t n_total, uint64_t start, uint64_t count, uint32_t* out,
                      pmvae_stream_t stream) {
  return launch_rng<0>(key, n_total, start, count, out, stream);
}
int pmvae_uniform(const uint32_t key[2], uint64_t n_total, uint64_t start, uint64_t count, float* out,
                  pmvae_stream_t stream) {
  return launch_rng<1>(key, n_total, start, count, out, stream);
}
int pmvae_normal(const uint32_t key[2], uint64_t n_total, uint64_t start, uint64_t count, float* out,
                 pmvae_stream_t stream) {
  return launch_rng<2>(key, n_total, start, count, out, stream);
}

int pmvae_mask_bernoulli(const uint32_t key[2], float p, uint64_t B_total, uint64_t row_start, uint64_t rows,
                         int32_t D, float* out, pmvae_stream_t stream) {
  PMVAE_CHECK(key && (out || rows == 0) && D > 0, "bad arguments");
  const uint64_t n_total = B_total * (uint64_t)D;
  PMVAE_CHECK(n_total <= 0xFFFFFFFFull, "a single draw is limited to 2^32-1 elements");
  PMVAE_CHECK(row_start + rows <= B_total, "row slice out of range");
  if (rows == 0) return 0;
  const uint64_t count = rows * (uint64_t)D;
  mask_bernoulli_kernel<<<grid_for(count, 256), 256, 0, as_stream(stream)>>>(Key2{key[0], key[1]}, p, n_total,
                                                                            row_start * (uint64_t)D, count, out);
  PMVAE_LAUNCH_CHECK();
  return 0;
}

int pmvae_mask_mnist(const uint32_t key[2], uint64_t B_total, uint64_t row_start, uint64_t rows, float* out,
                     pmvae_stream_t stream) {
  PMVAE_CHECK(key && (out || rows == 0), "bad arguments");
  PMVAE_CHECK(B_total * 784ull <= 0xFFFFFFFFull, "a single draw is limited to 2^32-1 elements");
  PMVAE_CHECK(row_start + rows <= B_total, "row slice out of range");
  if (rows == 0) return 0;
  uint32_t ks[8];
  host_random_bits(Key2{key[0], key[1]}, 8, ks);  // split(key, 4)
  MnistMaskKeys K;
  K.cat = Key2{ks[0], ks[1]};
  K.bern = Key2{ks[2], ks[3]};
  uint32_t sq[4];
  host_random_bits(Key2{ks[4], ks[5]}, 4, sq);  // randint splits its key once more
  K.sq_hi = Key2{sq[0], sq[1]};
  K.sq_lo = Key2{sq[2], sq[3]};
  K.rect = Key2{ks[6], ks[7]};
  const float w[7] = {2, 1, 1, 1, 1, 2, 2};
  float acc = 0.0f;
  for (int i = 0; i < 7; ++i) {
    volatile float pi = w[i] / 10.0f;  // float32 division, then sequential float32 cumsum
    acc = acc + pi;
    K.cum[i] = acc;
  }
  uint64_t g = rows < 148ull * 8ull ? rows : 148ull * 8ull;
  mask_mnist_kernel<<<(int)g, 256, 0, as_stream(stream)>>>(K, B_total, row_start, rows, out);
  PMVAE_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
