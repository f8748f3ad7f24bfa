// 2-D convolution / transposed convolution of the MNIST config's ConvEncoder / ConvDecoder (networks.py:9-72),
// NHWC float32, as ONE general operator (the form lax.conv_general_dilated reduces both to):
//
//   y[b, oy, ox, co] = act( bias[co] + sum_{ky, kx, ci} Xd(b, oy*stride + ky - pad_top, ox*stride + kx - pad_left, ci)
//                                                       * W[(ky*KW + kx) * Cin*Cout + ci*w_ci + co*w_co] )
//   Xd(b, v, u, ci) = X[b, v/dil, u/dil, ci] when v, u >= 0, both divisible by dil and inside the image, else 0
//   act(t) = t > 0 ? t : slope * t                      (jax.nn.leaky_relu, slope 0.01; slope 1 = identity)
//
// hk.Conv2D (weights HWIO):            stride = s, dil = 1, w_ci = Cout, w_co = 1
// hk.Conv2DTranspose (weights HWOI):   stride = 1, dil = s, w_ci = 1,    w_co = Cin   (lax.conv_transpose, kernel not flipped)
// Correctness-first direct kernels (row N1 of SURVEY §8f): one thread per output / input / weight element.
#include "kernels.h"
#include "tc_gemm.h"

namespace pmvae {

static int grid1d_c(int64_t work, int block) {
  int64_t g = ceil_div(work, block);
  if (g > 148ll * 16) g = 148ll * 16;
  if (g < 1) g = 1;
  return (int)g;
}

__global__ void __launch_bounds__(256) conv_fwd_kernel(const float* __restrict__ X, const float* __restrict__ W,
                                                       const float* __restrict__ bias, float* __restrict__ Y, int64_t B,
                                                       pmvae_conv_desc d) {
  const int64_t n = B * d.OH * d.OW * d.Cout;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const int co = (int)(t % d.Cout);
    int64_t r = t / d.Cout;
    const int ox = (int)(r % d.OW); r /= d.OW;
    const int oy = (int)(r % d.OH);
    const int64_t b = r / d.OH;
    float acc = bias ? bias[co] : 0.f;
    for (int ky = 0; ky < d.KH; ++ky) {
      const int v = oy * d.stride + ky - d.pad_top;
      if (v < 0 || v % d.dil != 0) continue;
      const int iy = v / d.dil;
      if (iy >= d.H) continue;
      for (int kx = 0; kx < d.KW; ++kx) {
        const int u = ox * d.stride + kx - d.pad_left;
        if (u < 0 || u % d.dil != 0) continue;
        const int ix = u / d.dil;
        if (ix >= d.W) continue;
        const float* xp = X + ((b * d.H + iy) * d.W + ix) * d.Cin;
        const float* wp = W + (int64_t)(ky * d.KW + kx) * d.Cin * d.Cout + (int64_t)co * d.w_co;
        for (int ci = 0; ci < d.Cin; ++ci) acc = fmaf(xp[ci], wp[(int64_t)ci * d.w_ci], acc);
      }
    }
    Y[t] = acc > 0.f ? acc : d.slope * acc;
  }
}

// dpre = dY * act'(y) from the stored activation (sign(y) = sign(pre-activation) for slope > 0)
__global__ void __launch_bounds__(256) conv_dpre_kernel(const float* __restrict__ dY, const float* __restrict__ Y,
                                                        float* __restrict__ dpre, int64_t n, float slope) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x)
    dpre[t] = Y[t] > 0.f ? dY[t] : slope * dY[t];
}

// dX[b, iy, ix, ci] = sum over the taps that reach it
__global__ void __launch_bounds__(256) conv_bwd_data_kernel(const float* __restrict__ dpre, const float* __restrict__ W,
                                                            float* __restrict__ dX, int64_t B, pmvae_conv_desc d) {
  const int64_t n = B * d.H * d.W * d.Cin;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const int ci = (int)(t % d.Cin);
    int64_t r = t / d.Cin;
    const int ix = (int)(r % d.W); r /= d.W;
    const int iy = (int)(r % d.H);
    const int64_t b = r / d.H;
    float acc = 0.f;
    for (int ky = 0; ky < d.KH; ++ky) {
      const int vy = iy * d.dil + d.pad_top - ky;          // = oy * stride
      if (vy < 0 || vy % d.stride != 0) continue;
      const int oy = vy / d.stride;
      if (oy >= d.OH) continue;
      for (int kx = 0; kx < d.KW; ++kx) {
        const int vx = ix * d.dil + d.pad_left - kx;
        if (vx < 0 || vx % d.stride != 0) continue;
        const int ox = vx / d.stride;
        if (ox >= d.OW) continue;
        const float* gp = dpre + ((b * d.OH + oy) * d.OW + ox) * d.Cout;
        const float* wp = W + (int64_t)(ky * d.KW + kx) * d.Cin * d.Cout + (int64_t)ci * d.w_ci;
        for (int co = 0; co < d.Cout; ++co) acc = fmaf(gp[co], wp[(int64_t)co * d.w_co], acc);
      }
    }
    dX[t] = acc;
  }
}

// dW[tap, ci, co] += sum over (b, oy, ox) in this block's slice of the batch; db handled by colsum_add
__global__ void __launch_bounds__(256) conv_bwd_weight_kernel(const float* __restrict__ X, const float* __restrict__ dpre,
                                                              float* __restrict__ dW, int64_t B, pmvae_conv_desc d) {
  const int64_t nw = (int64_t)d.KH * d.KW * d.Cin * d.Cout;
  const int64_t b_per = ceil_div(B, gridDim.y);
  const int64_t b0 = (int64_t)blockIdx.y * b_per, b1 = (b0 + b_per < B) ? b0 + b_per : B;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < nw; t += (int64_t)gridDim.x * blockDim.x) {
    // t enumerates (tap, ci, co) with co fastest so that dpre reads coalesce; the element lives at the layout's offset
    const int co = (int)(t % d.Cout);
    int64_t r = t / d.Cout;
    const int ci = (int)(r % d.Cin);
    const int tap = (int)(r / d.Cin);
    const int ky = tap / d.KW, kx = tap % d.KW;
    float acc = 0.f;
    for (int64_t b = b0; b < b1; ++b)
      for (int oy = 0; oy < d.OH; ++oy) {
        const int v = oy * d.stride + ky - d.pad_top;
        if (v < 0 || v % d.dil != 0) continue;
        const int iy = v / d.dil;
        if (iy >= d.H) continue;
        for (int ox = 0; ox < d.OW; ++ox) {
          const int u = ox * d.stride + kx - d.pad_left;
          if (u < 0 || u % d.dil != 0) continue;
          const int ix = u / d.dil;
          if (ix >= d.W) continue;
          acc = fmaf(X[((b * d.H + iy) * d.W + ix) * d.Cin + ci], dpre[((b * d.OH + oy) * d.OW + ox) * d.Cout + co], acc);
        }
      }
    atomicAdd(dW + (int64_t)tap * d.Cin * d.Cout + (int64_t)ci * d.w_ci + (int64_t)co * d.w_co, acc);
  }
}

// ---------------------------------------------------------------- im2col + GEMM path (used when a workspace is given)
// col[(b, oy, ox), (ky, kx, ci)] = Xd(b, oy*stride + ky - pad_top, ox*stride + kx - pad_left, ci)
__global__ void __launch_bounds__(256) im2col_kernel(const float* __restrict__ X, float* __restrict__ col, int64_t B,
                                                     pmvae_conv_desc d) {
  const int Kc = d.KH * d.KW * d.Cin;
  const int64_t n = B * d.OH * d.OW * Kc;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(t % Kc);
    int64_t r = t / Kc;
    const int ci = k % d.Cin, tap = k / d.Cin;
    const int ky = tap / d.KW, kx = tap % d.KW;
    const int ox = (int)(r % d.OW); r /= d.OW;
    const int oy = (int)(r % d.OH);
    const int64_t b = r / d.OH;
    const int v = oy * d.stride + ky - d.pad_top, u = ox * d.stride + kx - d.pad_left;
    float val = 0.f;
    if (v >= 0 && u >= 0 && v % d.dil == 0 && u % d.dil == 0) {
      const int iy = v / d.dil, ix = u / d.dil;
      if (iy < d.H && ix < d.W) val = X[((b * d.H + iy) * d.W + ix) * d.Cin + ci];
    }
    col[t] = val;
  }
}
// dX[b, iy, ix, ci] = sum over the taps that reach it of dcol[(b, oy, ox), (ky, kx, ci)]
__global__ void __launch_bounds__(256) col2im_kernel(const float* __restrict__ dcol, float* __restrict__ dX, int64_t B,
                                                     pmvae_conv_desc d) {
  const int Kc = d.KH * d.KW * d.Cin;
  const int64_t n = B * d.H * d.W * d.Cin;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const int ci = (int)(t % d.Cin);
    int64_t r = t / d.Cin;
    const int ix = (int)(r % d.W); r /= d.W;
    const int iy = (int)(r % d.H);
    const int64_t b = r / d.H;
    float acc = 0.f;
    for (int ky = 0; ky < d.KH; ++ky) {
      const int vy = iy * d.dil + d.pad_top - ky;
      if (vy < 0 || vy % d.stride != 0) continue;
      const int oy = vy / d.stride;
      if (oy >= d.OH) continue;
      for (int kx = 0; kx < d.KW; ++kx) {
        const int vx = ix * d.dil + d.pad_left - kx;
        if (vx < 0 || vx % d.stride != 0) continue;
        const int ox = vx / d.stride;
        if (ox >= d.OW) continue;
        acc += dcol[((b * d.OH + oy) * d.OW + ox) * (int64_t)Kc + (ky * d.KW + kx) * d.Cin + ci];
      }
    }
    dX[t] = acc;
  }
}
// weight tensor (either layout) <-> GEMM matrix Wm[(tap, ci), co]
__global__ void __launch_bounds__(256) wmat_pack_kernel(const float* __restrict__ w, float* __restrict__ wm, pmvae_conv_desc d) {
  const int64_t n = (int64_t)d.KH * d.KW * d.Cin * d.Cout;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const int co = (int)(t % d.Cout);
    const int64_t r = t / d.Cout;
    const int ci = (int)(r % d.Cin), tap = (int)(r / d.Cin);
    wm[t] = w[(int64_t)tap * d.Cin * d.Cout + (int64_t)ci * d.w_ci + (int64_t)co * d.w_co];
  }
}
__global__ void __launch_bounds__(256) wmat_unpack_add_kernel(const float* __restrict__ wm, float* __restrict__ dw, pmvae_conv_desc d) {
  const int64_t n = (int64_t)d.KH * d.KW * d.Cin * d.Cout;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const int co = (int)(t % d.Cout);
    const int64_t r = t / d.Cout;
    const int ci = (int)(r % d.Cin), tap = (int)(r / d.Cin);
    dw[(int64_t)tap * d.Cin * d.Cout + (int64_t)ci * d.w_ci + (int64_t)co * d.w_co] += wm[t];
  }
}
__global__ void __launch_bounds__(256) leaky_kernel(float* __restrict__ y, int64_t n, float slope) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const float v = y[t];
    y[t] = v > 0.f ? v : slope * v;
  }
}

struct ConvWs { float *col, *dcol, *wm; uint64_t bytes; };
static ConvWs plan_conv_ws(const pmvae_conv_desc* d, int64_t B, void* ws) {
  ConvWs p{};
  const uint64_t rows = (uint64_t)B * d->OH * d->OW, Kc = (uint64_t)d->KH * d->KW * d->Cin;
  uint64_t off = 0;
  auto take = [&](uint64_t n) { float* q = ws ? reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + off) : nullptr; off += align_up(n * 4, 256); return q; };
  p.col = take(rows * Kc);
  p.dcol = take(rows * Kc);
  p.wm = take(Kc * d->Cout);
  p.bytes = off;
  return p;
}
static int split_rows(int64_t out_rows, int64_t out_cols, int64_t K) {
  const int64_t tiles = ceil_div(out_rows, 64) * ceil_div(out_cols, 64);
  int64_t sp = ceil_div(148 * 4, tiles);
  const int64_t maxsp = ceil_div(K, 256);
  if (sp > maxsp) sp = maxsp;
  if (sp < 1) sp = 1;
  if (sp > 65535) sp = 65535;
  return (int)sp;
}

// ---------------------------------------------------------------- bf16 tensor-core path (desc.precision = 1)
// The same im2col formulation with bf16 GEMM operands and fp32 accumulation on the tcgen05 GEMMs of tc_gemm.cu:
//   forward   y    = leaky(col @ Wm + b)          gemm_nt(col [rows, Kc], Wm^T [Cout, Kc])
//   weights   dWm += col^T @ dpre                  gemm_tn(col, dpre)                (fp32 atomics)
//   data      dcol = dpre @ Wm^T                   gemm_nt(dpre [rows, Cout], Wm [Kc, Cout]) -> bf16, then col2im
// Row pitches are padded to 8 elements (TMA needs 16-byte pitches; the tensor maps' extents stay exact).
typedef __nv_bfloat16 bf16c;
__device__ __forceinline__ float conv_tap_value(const float* __restrict__ X, const pmvae_conv_desc& d, int64_t b, int oy, int ox,
                                                int k) {
  const int ci = k % d.Cin, tap = k / d.Cin;
  const int ky = tap / d.KW, kx = tap - ky * d.KW;
  const int v = oy * d.stride + ky - d.pad_top, u = ox * d.stride + kx - d.pad_left;
  if (v < 0 || u < 0) return 0.f;
  int iy = v, ix = u;
  if (d.dil != 1) {
    if (v % d.dil != 0 || u % d.dil != 0) return 0.f;
    iy = v / d.dil; ix = u / d.dil;
  }
  if (iy >= d.H || ix >= d.W) return 0.f;
  return X[((b * d.H + iy) * d.W + ix) * d.Cin + ci];
}
// source pixel of output position (oy, ox) under tap (ky, kx), or false when it falls into padding / a dilation hole
__device__ __forceinline__ bool conv_src(const pmvae_conv_desc& d, int oy, int ox, int ky, int kx, int& iy, int& ix) {
  const int v = oy * d.stride + ky - d.pad_top, u = ox * d.stride + kx - d.pad_left;
  if (v < 0 || u < 0) return false;
  iy = v; ix = u;
  if (d.dil != 1) {
    if (v % d.dil != 0 || u % d.dil != 0) return false;
    iy = v / d.dil; ix = u / d.dil;
  }
  return iy < d.H && ix < d.W;
}
// Cin % 8 == 0: one thread per (row, group of 8 channels), looping over the taps: for a tap the 8 channels are
// contiguous in the image (NHWC) and in the column (k = tap * Cin + ci): two 16-byte loads, one 16-byte store, and
// the only divisions are the per-thread row decomposition.
__global__ void __launch_bounds__(256) im2col_bf16_c8_kernel(const float* __restrict__ X, bf16c* __restrict__ col, int64_t B,
                                                             pmvae_conv_desc d, int Kp) {
  const int cgs = d.Cin >> 3;
  const int64_t n = B * d.OH * d.OW * cgs;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const int cg = (int)(t % cgs);
    const int64_t r = t / cgs;
    const int ox = (int)(r % d.OW);
    const int64_t r2 = r / d.OW;
    const int oy = (int)(r2 % d.OH);
    const int64_t b = r2 / d.OH;
    bf16c* dst = col + r * Kp + cg * 8;
    for (int ky = 0; ky < d.KH; ++ky)
      for (int kx = 0; kx < d.KW; ++kx) {
        int iy, ix;
        uint4 o = make_uint4(0u, 0u, 0u, 0u);
        if (conv_src(d, oy, ox, ky, kx, iy, ix)) {
          const float4* src = reinterpret_cast<const float4*>(X + ((b * d.H + iy) * d.W + ix) * d.Cin + cg * 8);
          const float4 v0 = __ldg(src), v1 = __ldg(src + 1);
          __nv_bfloat162 h0 = __floats2bfloat162_rn(v0.x, v0.y), h1 = __floats2bfloat162_rn(v0.z, v0.w);
          __nv_bfloat162 h2 = __floats2bfloat162_rn(v1.x, v1.y), h3 = __floats2bfloat162_rn(v1.z, v1.w);
          o = make_uint4(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1), *reinterpret_cast<uint32_t*>(&h2),
                         *reinterpret_cast<uint32_t*>(&h3));
        }
        *reinterpret_cast<uint4*>(dst + (ky * d.KW + kx) * d.Cin) = o;
      }
  }
}
// any Cin: one thread per (row, group of 8 k): a 16-byte store; columns [Kc, Kp) are zero
__global__ void __launch_bounds__(256) im2col_bf16_kernel(const float* __restrict__ X, bf16c* __restrict__ col, int64_t B,
                                                          pmvae_conv_desc d, int Kc, int Kp) {
  const int groups = Kp >> 3;
  const int64_t n = B * d.OH * d.OW * groups;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const int gk = (int)(t % groups);
    int64_t r = t / groups;
    const int ox = (int)(r % d.OW);
    const int64_t r2 = r / d.OW;
    const int oy = (int)(r2 % d.OH);
    const int64_t b = r2 / d.OH;
    uint32_t pk[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int k0 = gk * 8 + 2 * i;
      const float v0 = k0 < Kc ? conv_tap_value(X, d, b, oy, ox, k0) : 0.f;
      const float v1 = k0 + 1 < Kc ? conv_tap_value(X, d, b, oy, ox, k0 + 1) : 0.f;
      __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
      pk[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(col + r * Kp + gk * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
}
static int im2col_bf16(const float* x, bf16c* col, int64_t B, const pmvae_conv_desc& d, int Kc, int Kp, cudaStream_t s) {
  const int64_t rows = B * d.OH * d.OW;
  if (d.Cin % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15u) == 0)
    im2col_bf16_c8_kernel<<<grid1d_c(rows * (d.Cin >> 3), 256), 256, 0, s>>>(x, col, B, d, Kp);
  else
    im2col_bf16_kernel<<<grid1d_c(rows * (Kp >> 3), 256), 256, 0, s>>>(x, col, B, d, Kc, Kp);
  PMVAE_LAUNCH_CHECK();
  return 0;
}
// dX[b, iy, ix, ci..ci+V) = sum over the taps that reach the pixel of dcol[(b, oy, ox), tap * Cin + ci..]; V = 8 when
// Cin % 8 == 0 (one 16-byte load per tap), else 1
template <int V>
__global__ void __launch_bounds__(256) col2im_bf16_kernel(const bf16c* __restrict__ dcol, float* __restrict__ dX, int64_t B,
                                                          pmvae_conv_desc d, int Kp) {
  const int cgs = d.Cin / V;
  const int64_t n = B * d.H * d.W * cgs;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const int cg = (int)(t % cgs);
    int64_t r = t / cgs;
    const int ix = (int)(r % d.W); r /= d.W;
    const int iy = (int)(r % d.H);
    const int64_t b = r / d.H;
    float acc[V];
#pragma unroll
    for (int i = 0; i < V; ++i) acc[i] = 0.f;
    for (int ky = 0; ky < d.KH; ++ky) {
      const int vy = iy * d.dil + d.pad_top - ky;
      if (vy < 0 || vy % d.stride != 0) continue;
      const int oy = vy / d.stride;
      if (oy >= d.OH) continue;
      for (int kx = 0; kx < d.KW; ++kx) {
        const int vx = ix * d.dil + d.pad_left - kx;
        if (vx < 0 || vx % d.stride != 0) continue;
        const int ox = vx / d.stride;
        if (ox >= d.OW) continue;
        const bf16c* src = dcol + ((b * d.OH + oy) * d.OW + ox) * (int64_t)Kp + (ky * d.KW + kx) * d.Cin + cg * V;
        if (V == 8) {
          const uint4 q = __ldg(reinterpret_cast<const uint4*>(src));
          const uint32_t w4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            acc[2 * i] += __uint_as_float(w4[i] << 16);
            acc[2 * i + 1] += __uint_as_float(w4[i] & 0xFFFF0000u);
          }
        } else {
          acc[0] += __bfloat162float(src[0]);
        }
      }
    }
    float* dst = dX + (((b * d.H + iy) * d.W + ix) * d.Cin + cg * V);
#pragma unroll
    for (int i = 0; i < V; ++i) dst[i] = acc[i];
  }
}
// bf16 weight images with the output channels padded to Cp: nk[co][k] (pitch Kp) for the forward, kn[k][co] (pitch Cp)
// for the data gradient; padded rows / columns are zero
// flip = 1 reverses the tap order (the adjoint convolution of the data gradient); kn may be NULL
__global__ void __launch_bounds__(256) wpack_bf16_kernel(const float* __restrict__ w, bf16c* __restrict__ nk, bf16c* __restrict__ kn,
                                                         pmvae_conv_desc d, int Kc, int Kp, int Cp, int flip) {
  const int64_t n = (int64_t)Kp * Cp;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const int co = (int)(t % Cp);
    const int k = (int)(t / Cp);
    float v = 0.f;
    if (k < Kc && co < d.Cout) {
      const int ci = k % d.Cin;
      int tap = k / d.Cin;
      if (flip) tap = d.KH * d.KW - 1 - tap;
      v = w[(int64_t)tap * d.Cin * d.Cout + (int64_t)ci * d.w_ci + (int64_t)co * d.w_co];
    }
    const bf16c h = __float2bfloat16(v);
    nk[(int64_t)co * Kp + k] = h;
    if (kn && k < Kc) kn[(int64_t)k * Cp + co] = h;
  }
}
// dpre = dy * act'(y) in place (float32, for the bias gradient) and as bf16 with the channel pitch padded to Cp
__global__ void __launch_bounds__(256) conv_dpre_bf16_kernel(float* __restrict__ dY, const float* __restrict__ Y,
                                                             bf16c* __restrict__ out, int64_t rows, int Cout, int Cp, float slope) {
  const int64_t n = rows * Cp;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(t % Cp);
    const int64_t r = t / Cp;
    float g = 0.f;
    if (c < Cout) {
      const int64_t i = r * Cout + c;
      g = dY[i] * (Y[i] > 0.f ? 1.f : slope);
      dY[i] = g;
    }
    out[t] = __float2bfloat16(g);
  }
}
// y[r, c] = act(tmp[r, c]) from the GEMM's padded output pitch
__global__ void __launch_bounds__(256) leaky_gather_kernel(const float* __restrict__ tmp, float* __restrict__ y, int64_t rows,
                                                           int Cout, int Cp, float slope) {
  const int64_t n = rows * Cout;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const float v = tmp[(t / Cout) * Cp + (t % Cout)];
    y[t] = v > 0.f ? v : slope * v;
  }
}
// dw (either layout) += wm[(tap, ci), co] with column pitch Cp
__global__ void __launch_bounds__(256) wmat_unpack_add_p_kernel(const float* __restrict__ wm, float* __restrict__ dw, pmvae_conv_desc d,
                                                                int Cp) {
  const int64_t n = (int64_t)d.KH * d.KW * d.Cin * d.Cout;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const int co = (int)(t % d.Cout);
    const int64_t r = t / d.Cout;
    const int ci = (int)(r % d.Cin), tap = (int)(r / d.Cin);
    dw[(int64_t)tap * d.Cin * d.Cout + (int64_t)ci * d.w_ci + (int64_t)co * d.w_co] += wm[r * Cp + co];
  }
}
// The data gradient as a convolution of its own: dX = conv(dpre) with stride and dilation swapped, pads K-1-pad, the
// taps reversed and the channel roles exchanged (same operator, so the same im2col + NT GEMM, K = KH*KW*Cout, N = Cin).
static pmvae_conv_desc adjoint_desc(const pmvae_conv_desc& d) {
  pmvae_conv_desc a = d;
  a.H = d.OH; a.W = d.OW; a.Cin = d.Cout;
  a.OH = d.H; a.OW = d.W; a.Cout = d.Cin;
  a.stride = d.dil; a.dil = d.stride;
  a.pad_top = d.KH - 1 - d.pad_top; a.pad_left = d.KW - 1 - d.pad_left;
  a.w_ci = d.w_co; a.w_co = d.w_ci;
  a.slope = 1.0f;
  return a;
}
// column-matrix elements of the two ways to get dX: rows(out) x K x Cin (dcol + col2im) against rows(in) x K x Cout
static bool use_adjoint(const pmvae_conv_desc& d) {
  if (d.pad_top > d.KH - 1 || d.pad_left > d.KW - 1) return false;
  const double direct = (double)d.OH * d.OW * d.Cin, adj = (double)d.H * d.W * d.Cout;
  return adj <= direct;
}
struct ConvWsB { bf16c *col, *dcol, *nk, *kn, *dpre; float *wm, *bias, *ytmp; uint64_t bytes; };
static ConvWsB plan_conv_ws_b(const pmvae_conv_desc* d, int64_t B, void* ws) {
  ConvWsB p{};
  const uint64_t rows = (uint64_t)B * d->OH * d->OW, Kc = (uint64_t)d->KH * d->KW * d->Cin, Kp = (Kc + 7) & ~7ull;
  const uint64_t Cp = ((uint64_t)d->Cout + 7) & ~7ull;
  // the adjoint convolution of the data gradient reuses col / nk / ytmp with its own shapes
  const uint64_t rows_a = (uint64_t)B * d->H * d->W, Kc_a = (uint64_t)d->KH * d->KW * d->Cout, Kp_a = (Kc_a + 7) & ~7ull;
  const uint64_t Cp_a = ((uint64_t)d->Cin + 7) & ~7ull;
  const bool adj = use_adjoint(*d);
  auto mx = [](uint64_t a, uint64_t b) { return a > b ? a : b; };
  uint64_t off = 0;
  auto take = [&](uint64_t bytes) { char* q = ws ? reinterpret_cast<char*>(ws) + off : nullptr; off += align_up(bytes, 256); return q; };
  p.col = reinterpret_cast<bf16c*>(take(mx(rows * Kp, adj ? rows_a * Kp_a : 0) * 2));
  p.dcol = reinterpret_cast<bf16c*>(take(adj ? 0 : rows * Kp * 2));
  p.nk = reinterpret_cast<bf16c*>(take(mx(Cp * Kp, adj ? Cp_a * Kp_a : 0) * 2));
  p.kn = reinterpret_cast<bf16c*>(take(Kc * Cp * 2));
  p.dpre = reinterpret_cast<bf16c*>(take(rows * Cp * 2));
  p.wm = reinterpret_cast<float*>(take(Kc * Cp * 4));
  p.bias = reinterpret_cast<float*>(take(Cp * 4));
  p.ytmp = reinterpret_cast<float*>(take(mx(Cp != (uint64_t)d->Cout ? rows * Cp * 4 : 0, (adj && Cp_a != (uint64_t)d->Cin) ? rows_a * Cp_a * 4 : 0)));
  p.bytes = off;
  return p;
}
static bool conv_bf16_ok(const pmvae_conv_desc* d) { return d->reserved == 1; }

static int check_desc(const pmvae_conv_desc* d) {
  PMVAE_CHECK(d != nullptr, "null conv descriptor");
  PMVAE_CHECK(d->H > 0 && d->W > 0 && d->Cin > 0 && d->OH > 0 && d->OW > 0 && d->Cout > 0 && d->KH > 0 && d->KW > 0 &&
                  d->stride > 0 && d->dil > 0 && d->pad_top >= 0 && d->pad_left >= 0 && d->w_ci > 0 && d->w_co > 0 &&
                  d->slope > 0.f,
              "bad conv descriptor");
  return 0;
}

}  // namespace pmvae

using namespace pmvae;

extern "C" {

uint64_t pmvae_conv2d_workspace_bytes(const pmvae_conv_desc* desc, int64_t B) {
  if (check_desc(desc) != 0 || B < 0) return 0;
  const uint64_t f32 = plan_conv_ws(desc, B < 1 ? 1 : B, nullptr).bytes, b16 = plan_conv_ws_b(desc, B < 1 ? 1 : B, nullptr).bytes;
  return (f32 > b16 ? f32 : b16) + 256;
}

int pmvae_conv2d_forward(const pmvae_conv_desc* desc, const float* x, const float* w, const float* bias, int64_t B,
                         float* y, void* ws, uint64_t ws_bytes, pmvae_stream_t stream) {
  PMVAE_TRY(check_desc(desc));
  if (B == 0) return 0;
  PMVAE_CHECK(x && w && y && B > 0, "null pointer");
  if (ws && conv_bf16_ok(desc)) {
    ConvWsB p = plan_conv_ws_b(desc, B, ws);
    PMVAE_CHECK(p.bytes <= ws_bytes && (reinterpret_cast<uintptr_t>(ws) & 255u) == 0, "conv workspace too small or misaligned");
    cudaStream_t s = as_stream(stream);
    const int64_t rows = B * desc->OH * desc->OW;
    const int Kc = desc->KH * desc->KW * desc->Cin, Kp = (Kc + 7) & ~7, Cp = (desc->Cout + 7) & ~7;
    PMVAE_TRY(im2col_bf16(x, p.col, B, *desc, Kc, Kp, s));
    wpack_bf16_kernel<<<grid1d_c((int64_t)Kp * Cp, 256), 256, 0, s>>>(w, p.nk, p.kn, *desc, Kc, Kp, Cp, 0);
    PMVAE_LAUNCH_CHECK();
    tc::TcGemmArgs e{};
    if (bias) {
      // parameter leaves are views into a flat arena (4-byte aligned): the epilogue reads the bias in 16-byte vectors
      PMVAE_CUDA(cudaMemsetAsync(p.bias, 0, (size_t)Cp * sizeof(float), s));
      PMVAE_CUDA(cudaMemcpyAsync(p.bias, bias, (size_t)desc->Cout * sizeof(float), cudaMemcpyDeviceToDevice, s));
      e.bias = p.bias;
    }
    const bool padded = Cp != desc->Cout;
    PMVAE_CHECK(padded || (reinterpret_cast<uintptr_t>(y) & 15u) == 0, "bf16 convolution path: y must be 16-byte aligned");
    e.out_f32 = padded ? p.ytmp : y; e.ld_out_f32 = Cp;
    PMVAE_TRY(tc::gemm_nt(p.col, Kp, p.nk, Kp, rows, Cp, Kc, e, s));
    if (padded) {
      leaky_gather_kernel<<<grid1d_c(rows * desc->Cout, 256), 256, 0, s>>>(p.ytmp, y, rows, desc->Cout, Cp, desc->slope);
      PMVAE_LAUNCH_CHECK();
    } else if (desc->slope != 1.0f) {
      leaky_kernel<<<grid1d_c(rows * desc->Cout, 256), 256, 0, s>>>(y, rows * desc->Cout, desc->slope);
      PMVAE_LAUNCH_CHECK();
    }
    return 0;
  }
  if (ws) {
    // im2col + fp32 GEMM: y = leaky(col @ Wm + bias)
    ConvWs p = plan_conv_ws(desc, B, ws);
    PMVAE_CHECK(p.bytes <= ws_bytes && (reinterpret_cast<uintptr_t>(ws) & 255u) == 0, "conv workspace too small or misaligned");
    cudaStream_t s = as_stream(stream);
    const int64_t rows = B * desc->OH * desc->OW;
    const int Kc = desc->KH * desc->KW * desc->Cin;
    im2col_kernel<<<grid1d_c(rows * Kc, 256), 256, 0, s>>>(x, p.col, B, *desc);
    PMVAE_LAUNCH_CHECK();
    const float* wm = w;
    if (!(desc->w_co == 1 && desc->w_ci == desc->Cout)) {
      wmat_pack_kernel<<<grid1d_c((int64_t)Kc * desc->Cout, 256), 256, 0, s>>>(w, p.wm, *desc);
      PMVAE_LAUNCH_CHECK();
      wm = p.wm;
    }
    GemmF32Args a{};
    a.M = rows; a.N = desc->Cout; a.K = Kc;
    a.A = p.col; a.lda = Kc; a.B = wm; a.ldb = desc->Cout; a.C = y; a.ldc = desc->Cout; a.bias = bias;
    PMVAE_TRY(gemm_f32(a, false, false, s));
    if (desc->slope != 1.0f) {
      leaky_kernel<<<grid1d_c(rows * desc->Cout, 256), 256, 0, s>>>(y, rows * desc->Cout, desc->slope);
      PMVAE_LAUNCH_CHECK();
    }
    return 0;
  }
  conv_fwd_kernel<<<grid1d_c(B * desc->OH * desc->OW * desc->Cout, 256), 256, 0, as_stream(stream)>>>(x, w, bias, y, B, *desc);
  PMVAE_LAUNCH_CHECK();
  return 0;
}

// dy: cotangent of y (overwritten with the pre-activation cotangent); dx may be NULL; dw / dbias are ACCUMULATED into
int pmvae_conv2d_backward(const pmvae_conv_desc* desc, const float* x, const float* w, const float* y, float* dy,
                          int64_t B, float* dx, float* dw, float* dbias, void* ws, uint64_t ws_bytes,
                          pmvae_stream_t stream) {
  PMVAE_TRY(check_desc(desc));
  if (B == 0) return 0;
  PMVAE_CHECK(x && w && y && dy && B > 0, "null pointer");
  cudaStream_t s = as_stream(stream);
  const int64_t ny = B * desc->OH * desc->OW * desc->Cout;
  const int Kc_all = desc->KH * desc->KW * desc->Cin;
  if (ws && conv_bf16_ok(desc) && (dx == nullptr || Kc_all % 8 == 0 || use_adjoint(*desc))) {
    ConvWsB p = plan_conv_ws_b(desc, B, ws);
    PMVAE_CHECK(p.bytes <= ws_bytes && (reinterpret_cast<uintptr_t>(ws) & 255u) == 0, "conv workspace too small or misaligned");
    const int64_t rows = B * desc->OH * desc->OW;
    const int Kc = Kc_all, Kp = (Kc + 7) & ~7, Cp = (desc->Cout + 7) & ~7;
    // the weight-gradient GEMM accumulates with vector atomics: straight into dw only when it is laid out as the
    // GEMM matrix and 16-byte aligned (leaves of a flat arena need not be)
    const bool direct_w = desc->w_co == 1 && desc->w_ci == desc->Cout && Cp == desc->Cout &&
                          (reinterpret_cast<uintptr_t>(dw) & 15u) == 0;
    conv_dpre_bf16_kernel<<<grid1d_c(rows * Cp, 256), 256, 0, s>>>(dy, y, p.dpre, rows, desc->Cout, Cp, desc->slope);
    PMVAE_LAUNCH_CHECK();
    if (dw) {
      PMVAE_TRY(im2col_bf16(x, p.col, B, *desc, Kc, Kp, s));
      float* target = dw;
      if (!direct_w) {
        PMVAE_CUDA(cudaMemsetAsync(p.wm, 0, (size_t)Kc * Cp * sizeof(float), s));
        target = p.wm;
      }
      PMVAE_TRY(tc::gemm_tn(p.col, Kp, p.dpre, Cp, Kc, Cp, rows, target, Cp, 1, 0, nullptr, s));
      if (!direct_w) {
        wmat_unpack_add_p_kernel<<<grid1d_c((int64_t)Kc * desc->Cout, 256), 256, 0, s>>>(p.wm, dw, *desc, Cp);
        PMVAE_LAUNCH_CHECK();
      }
    }
    if (dx && use_adjoint(*desc)) {
      const pmvae_conv_desc a = adjoint_desc(*desc);
      const int64_t rows_a = B * a.OH * a.OW;
      const int Kc_a = a.KH * a.KW * a.Cin, Kp_a = (Kc_a + 7) & ~7, Cp_a = (a.Cout + 7) & ~7;
      PMVAE_TRY(im2col_bf16(dy, p.col, B, a, Kc_a, Kp_a, s));          // dy holds dpre (float32) by now
      wpack_bf16_kernel<<<grid1d_c((int64_t)Kp_a * Cp_a, 256), 256, 0, s>>>(w, p.nk, nullptr, a, Kc_a, Kp_a, Cp_a, 1);
      PMVAE_LAUNCH_CHECK();
      tc::TcGemmArgs e{};
      const bool padded = Cp_a != a.Cout;
      PMVAE_CHECK(padded || (reinterpret_cast<uintptr_t>(dx) & 15u) == 0, "bf16 convolution path: dx must be 16-byte aligned");
      e.out_f32 = padded ? p.ytmp : dx; e.ld_out_f32 = Cp_a;
      PMVAE_TRY(tc::gemm_nt(p.col, Kp_a, p.nk, Kp_a, rows_a, Cp_a, Kc_a, e, s));
      if (padded) {
        leaky_gather_kernel<<<grid1d_c(rows_a * a.Cout, 256), 256, 0, s>>>(p.ytmp, dx, rows_a, a.Cout, Cp_a, 1.0f);
        PMVAE_LAUNCH_CHECK();
      }
    } else if (dx) {
      wpack_bf16_kernel<<<grid1d_c((int64_t)Kp * Cp, 256), 256, 0, s>>>(w, p.nk, p.kn, *desc, Kc, Kp, Cp, 0);
      PMVAE_LAUNCH_CHECK();
      tc::TcGemmArgs e{};
      e.out_bf16 = p.dcol; e.ld_out_bf16 = Kp;
      PMVAE_TRY(tc::gemm_nt(p.dpre, Cp, p.kn, Cp, rows, Kc, Cp, e, s));
      const int64_t npix = B * desc->H * desc->W;
      if (desc->Cin % 8 == 0)
        col2im_bf16_kernel<8><<<grid1d_c(npix * (desc->Cin >> 3), 256), 256, 0, s>>>(p.dcol, dx, B, *desc, Kp);
      else
        col2im_bf16_kernel<1><<<grid1d_c(npix * desc->Cin, 256), 256, 0, s>>>(p.dcol, dx, B, *desc, Kp);
      PMVAE_LAUNCH_CHECK();
    }
    if (dbias) PMVAE_TRY(colsum_add(dy, desc->Cout, dbias, rows, desc->Cout, s));
    return 0;
  }
  conv_dpre_kernel<<<grid1d_c(ny, 256), 256, 0, s>>>(dy, y, dy, ny, desc->slope);
  PMVAE_LAUNCH_CHECK();
  if (ws) {
    // dWm += col^T @ dpre;  dcol = dpre @ Wm^T -> col2im
    ConvWs p = plan_conv_ws(desc, B, ws);
    PMVAE_CHECK(p.bytes <= ws_bytes && (reinterpret_cast<uintptr_t>(ws) & 255u) == 0, "conv workspace too small or misaligned");
    const int64_t rows = B * desc->OH * desc->OW;
    const int Kc = desc->KH * desc->KW * desc->Cin;
    const bool direct_w = desc->w_co == 1 && desc->w_ci == desc->Cout;
    if (dw) {
      im2col_kernel<<<grid1d_c(rows * Kc, 256), 256, 0, s>>>(x, p.col, B, *desc);
      PMVAE_LAUNCH_CHECK();
      float* target = dw;
      if (!direct_w) {
        PMVAE_CUDA(cudaMemsetAsync(p.wm, 0, (size_t)Kc * desc->Cout * sizeof(float), s));
        target = p.wm;
      }
      GemmF32Args a{};
      a.M = Kc; a.N = desc->Cout; a.K = rows;
      a.A = p.col; a.lda = Kc; a.B = dy; a.ldb = desc->Cout; a.C = target; a.ldc = desc->Cout;
      a.atomic = 1; a.split_k = split_rows(Kc, desc->Cout, rows);
      PMVAE_TRY(gemm_f32(a, true, false, s));
      if (!direct_w) {
        wmat_unpack_add_kernel<<<grid1d_c((int64_t)Kc * desc->Cout, 256), 256, 0, s>>>(p.wm, dw, *desc);
        PMVAE_LAUNCH_CHECK();
      }
    }
    if (dx) {
      const float* wm = w;
      if (!direct_w) {
        wmat_pack_kernel<<<grid1d_c((int64_t)Kc * desc->Cout, 256), 256, 0, s>>>(w, p.wm, *desc);
        PMVAE_LAUNCH_CHECK();
        wm = p.wm;
      }
      GemmF32Args a{};
      a.M = rows; a.N = Kc; a.K = desc->Cout;
      a.A = dy; a.lda = desc->Cout; a.B = wm; a.ldb = desc->Cout; a.C = p.dcol; a.ldc = Kc;
      PMVAE_TRY(gemm_f32(a, false, true, s));
      col2im_kernel<<<grid1d_c(B * desc->H * desc->W * desc->Cin, 256), 256, 0, s>>>(p.dcol, dx, B, *desc);
      PMVAE_LAUNCH_CHECK();
    }
    if (dbias) PMVAE_TRY(colsum_add(dy, desc->Cout, dbias, rows, desc->Cout, s));
    return 0;
  }
  if (dx) {
    conv_bwd_data_kernel<<<grid1d_c(B * desc->H * desc->W * desc->Cin, 256), 256, 0, s>>>(dy, w, dx, B, *desc);
    PMVAE_LAUNCH_CHECK();
  }
  if (dw) {
    const int64_t nw = (int64_t)desc->KH * desc->KW * desc->Cin * desc->Cout;
    int gx = (int)ceil_div(nw, 256);
    int gy = (int)((148 * 8 + gx - 1) / gx);
    if (gy > B) gy = (int)B;
    if (gy < 1) gy = 1;
    conv_bwd_weight_kernel<<<dim3((unsigned)gx, (unsigned)gy), 256, 0, s>>>(x, dy, dw, B, *desc);
    PMVAE_LAUNCH_CHECK();
  }
  if (dbias) PMVAE_TRY(colsum_add(dy, desc->Cout, dbias, B * desc->OH * desc->OW, desc->Cout, s));
  return 0;
}

}  // extern "C"
