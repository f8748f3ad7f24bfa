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
  return plan_conv_ws(desc, B < 1 ? 1 : B, nullptr).bytes + 256;
}

int pmvae_conv2d_forward(const pmvae_conv_desc* desc, const float* x, const float* w, const float* bias, int64_t B,
                         float* y, void* ws, uint64_t ws_bytes, pmvae_stream_t stream) {
  PMVAE_TRY(check_desc(desc));
  if (B == 0) return 0;
  PMVAE_CHECK(x && w && y && B > 0, "null pointer");
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
