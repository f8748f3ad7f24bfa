"""CPU restatement (PyTorch) of hk.Conv2D / hk.Conv2DTranspose + leaky_relu as used by ConvEncoder / ConvDecoder
(reference posterior_matching/models/networks.py:9-72).  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Semantics restated from jax.lax (0.2.26) and dm-haiku 0.0.5, all [R] (recollection, unverifiable here):
  * Conv2D: NHWC input, weights [kh, kw, in, out]; SAME pads so that out = ceil(in / stride) with the extra
    element on the high side (lax.padtype_to_pads); VALID pads nothing.
  * Conv2DTranspose: weights [kh, kw, out, in]; lax.conv_transpose(transpose_kernel=False) = a stride-1
    correlation (kernel NOT flipped) over the input dilated by the stride, padded by
    lax._conv_transpose_padding: SAME -> pad_len = k + s - 2, pad_lo = k - 1 if s > k - 1 else ceil(pad_len / 2);
    VALID -> pad_len = k + s - 2 + max(k - s, 0), pad_lo = k - 1.
The dilation and padding are materialised explicitly here (zeros inserted / appended), so this does not share
the index arithmetic of the CUDA kernels it checks.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def _nchw(x):
    return x.permute(0, 3, 1, 2)


def _nhwc(x):
    return x.permute(0, 2, 3, 1)


def conv2d(x, w, b, stride: int, padding: str, slope: float = 0.01):
    """x [B,H,W,Cin], w [kh,kw,Cin,Cout] -> leaky_relu(conv + b) [B,OH,OW,Cout]."""
    k = w.shape[0]
    H, W = x.shape[1], x.shape[2]
    xp = _nchw(x)
    if padding == "SAME":
        pads = []
        for size in (W, H):               # F.pad order: last dim first
            out = -(-size // stride)
            total = max((out - 1) * stride + k - size, 0)
            pads += [total // 2, total - total // 2]
        xp = F.pad(xp, pads)
    y = F.conv2d(xp, w.permute(3, 2, 0, 1).contiguous(), b, stride=stride)
    return F.leaky_relu(_nhwc(y), slope)


def conv2d_transpose(x, w, b, stride: int, padding: str, slope: float = 0.01):
    """x [B,H,W,Cin], w [kh,kw,Cout,Cin] -> leaky_relu(conv_transpose + b) [B,OH,OW,Cout]."""
    k = w.shape[0]
    B, H, W, Cin = x.shape
    xd = torch.zeros(B, (H - 1) * stride + 1, (W - 1) * stride + 1, Cin, dtype=x.dtype)
    xd[:, ::stride, ::stride, :] = x
    if padding == "SAME":
        pad_len = k + stride - 2
        pad_a = k - 1 if stride > k - 1 else int(math.ceil(pad_len / 2))
    else:
        pad_len = k + stride - 2 + max(k - stride, 0)
        pad_a = k - 1
    pad_b = pad_len - pad_a
    xp = F.pad(_nchw(xd), [pad_a, pad_b, pad_a, pad_b])
    y = F.conv2d(xp, w.permute(2, 3, 0, 1).contiguous(), b, stride=1)      # [kh,kw,O,I] -> [O,I,kh,kw], no flip
    return F.leaky_relu(_nhwc(y), slope)


MNIST_ENCODER = [(32, 5, 1), (32, 5, 2), (64, 5, 1), (64, 5, 2), (128, 7, 1)]       # configs/pm_vae_mnist.py:24-30
MNIST_DECODER = [(64, 7, 1), (64, 5, 2), (32, 5, 1), (32, 5, 2), (32, 5, 1), (1, 5, 1)]  # :32-40


def conv_encoder(params, x, layers=MNIST_ENCODER):
    """ConvEncoder.__call__ (networks.py:24-38): SAME except the last layer (VALID); leaky_relu after every layer."""
    h = x
    for i, (_, _, s) in enumerate(layers):
        w, b = params[i]
        h = conv2d(h, w, b, s, "VALID" if i == len(layers) - 1 else "SAME")
    return h


def conv_decoder(params, z, layers=MNIST_DECODER):
    """ConvDecoder.__call__ (networks.py:55-72): z -> [B,1,1,d]; first layer VALID, the rest SAME; leaky_relu after
    every layer, including the last (the Bernoulli logits)."""
    h = z.reshape(z.shape[0], 1, 1, z.shape[1])
    for i, (_, _, s) in enumerate(layers):
        w, b = params[i]
        h = conv2d_transpose(h, w, b, s, "VALID" if i == 0 else "SAME")
    return h


# ---- the general operator of csrc/conv.cu, restated (test infrastructure; pins the descriptor algebra on the CPU) ----
def general_conv(x, w_taps, d):
    """y[b,oy,ox,co] = sum_{ky,kx,ci} Xd(b, oy*stride+ky-pad_top, ox*stride+kx-pad_left, ci) * w_taps[ky,kx,ci,co], where Xd is
    x with dil-1 zeros inserted between pixels (csrc/conv.cu header).  `d`: dict with H, W, Cin, OH, OW, Cout, KH, KW,
    stride, dil, pad_top, pad_left.  x [B,H,W,Cin], w_taps [KH,KW,Cin,Cout] (already in tap-major HWIO order)."""
    B = x.shape[0]
    Hd, Wd = (d["H"] - 1) * d["dil"] + 1, (d["W"] - 1) * d["dil"] + 1
    xd = x.new_zeros((B, Hd, Wd, d["Cin"]))
    xd[:, ::d["dil"], ::d["dil"], :] = x
    need_h, need_w = (d["OH"] - 1) * d["stride"] + d["KH"], (d["OW"] - 1) * d["stride"] + d["KW"]
    pb, pr = need_h - Hd - d["pad_top"], need_w - Wd - d["pad_left"]
    xp = F.pad(_nchw(xd), (d["pad_left"], max(pr, 0), d["pad_top"], max(pb, 0)))
    xp = xp[:, :, :need_h, :need_w]
    return _nhwc(F.conv2d(xp, w_taps.permute(3, 2, 0, 1).contiguous(), stride=d["stride"]))


def adjoint_desc(d):
    """Descriptor of the data gradient as a convolution of its own (csrc/conv.cu::adjoint_desc): stride and dilation swap,
    pads become K-1-pad, channels swap; the weights are used with reversed taps and transposed channels."""
    return dict(H=d["OH"], W=d["OW"], Cin=d["Cout"], OH=d["H"], OW=d["W"], Cout=d["Cin"], KH=d["KH"], KW=d["KW"],
                stride=d["dil"], dil=d["stride"], pad_top=d["KH"] - 1 - d["pad_top"], pad_left=d["KW"] - 1 - d["pad_left"])
