"""hk.Conv2D / hk.Conv2DTranspose + leaky_relu of the MNIST config's ConvEncoder / ConvDecoder
(reference posterior_matching/models/networks.py:9-72) on the device: descriptor builders for the general
convolution operator of libpmvae (`pmvae_conv2d_forward / backward`, NHWC float32) and thin functional wrappers.

Padding follows `lax.padtype_to_pads` (Conv2D) and `lax.conv_transpose` with `transpose_kernel=False`
(Conv2DTranspose) as recollected -- JAX is not installable here, so those semantics are [R] (unpinned); the
oracle (`oracle/conv.py`) states the same rules independently with explicit zero-insertion and padding.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional, Tuple

import torch

from . import _lib


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def conv_pads(size: int, k: int, s: int, padding: str) -> Tuple[int, int, int]:
    """(out, pad_lo, pad_hi) of hk.Conv2D along one axis."""
    if padding == "VALID":
        return (size - k) // s + 1, 0, 0
    out = -(-size // s)
    total = max((out - 1) * s + k - size, 0)
    return out, total // 2, total - total // 2


def conv_transpose_pads(size: int, k: int, s: int, padding: str) -> Tuple[int, int, int]:
    """(out, pad_lo, pad_hi) of hk.Conv2DTranspose along one axis (lax._conv_transpose_padding)."""
    if padding == "SAME":
        pad_len = k + s - 2
        pad_a = k - 1 if s > k - 1 else int(math.ceil(pad_len / 2))
    else:
        pad_len = k + s - 2 + max(k - s, 0)
        pad_a = k - 1
    pad_b = pad_len - pad_a
    return (size - 1) * s + 1 + pad_a + pad_b - k + 1, pad_a, pad_b


def conv_desc(H: int, W: int, Cin: int, Cout: int, k: int, s: int, padding: str, *, transpose: bool = False,
              slope: float = 0.01, precision: str = "fp32") -> _lib.ConvDesc:
    """`precision="bf16"`: bf16 GEMM operands on the tcgen05 GEMMs with fp32 accumulation (layers whose channel counts
    do not give 16-byte operand pitches, e.g. the decoder's 1-channel output layer, stay on the float32 GEMM)."""
    if precision not in ("fp32", "bf16"):
        raise ValueError("precision must be 'fp32' or 'bf16'")
    d = _lib.ConvDesc()
    d.reserved = 1 if precision == "bf16" else 0
    pads = conv_transpose_pads if transpose else conv_pads
    OH, pt, _ = pads(H, k, s, padding)
    OW, pl, _ = pads(W, k, s, padding)
    d.H, d.W, d.Cin, d.OH, d.OW, d.Cout, d.KH, d.KW = H, W, Cin, OH, OW, Cout, k, k
    d.pad_top, d.pad_left = pt, pl
    if transpose:        # weights [kh, kw, out, in]
        d.stride, d.dil, d.w_ci, d.w_co = 1, s, 1, Cin
    else:                # weights [kh, kw, in, out]
        d.stride, d.dil, d.w_ci, d.w_co = s, 1, Cout, 1
    d.slope = float(slope)
    return d


_WS = {}


def conv_workspace(d: _lib.ConvDesc, B: int, device) -> torch.Tensor:
    """One shared, growing scratch buffer per device for the im2col + GEMM form of the operator."""
    need = int(_lib.lib.pmvae_conv2d_workspace_bytes(C.byref(d), B))
    ws = _WS.get(device)
    if ws is None or ws.numel() < need:
        ws = torch.empty(need + 256, dtype=torch.uint8, device=device)
        _WS[device] = ws
    off = (-ws.data_ptr()) % 256
    return ws[off:]


def conv2d_forward(d: _lib.ConvDesc, x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor], *,
                   direct: bool = False) -> torch.Tensor:
    """`direct=True` runs the one-thread-per-element kernels instead of im2col + GEMM (cross-check)."""
    B = x.shape[0]
    y = torch.empty((B, d.OH, d.OW, d.Cout), dtype=torch.float32, device=x.device)
    ws = None if direct else conv_workspace(d, B, x.device)
    _lib.check(_lib.lib.pmvae_conv2d_forward(C.byref(d), x.data_ptr(), w.data_ptr(), b.data_ptr() if b is not None else None,
                                             B, y.data_ptr(), ws.data_ptr() if ws is not None else None,
                                             ws.numel() if ws is not None else 0, _stream()), "pmvae_conv2d_forward")
    return y


def conv2d_backward(d: _lib.ConvDesc, x: torch.Tensor, w: torch.Tensor, y: torch.Tensor, dy: torch.Tensor,
                    dw: torch.Tensor, db: Optional[torch.Tensor], need_dx: bool = True, *,
                    direct: bool = False) -> Optional[torch.Tensor]:
    """Accumulates into dw / db, overwrites dy with the pre-activation cotangent, returns dx (or None)."""
    B = x.shape[0]
    dx = torch.empty_like(x) if need_dx else None
    ws = None if direct else conv_workspace(d, B, x.device)
    _lib.check(_lib.lib.pmvae_conv2d_backward(C.byref(d), x.data_ptr(), w.data_ptr(), y.data_ptr(), dy.data_ptr(), B,
                                              dx.data_ptr() if need_dx else None, dw.data_ptr(),
                                              db.data_ptr() if db is not None else None,
                                              ws.data_ptr() if ws is not None else None, ws.numel() if ws is not None else 0,
                                              _stream()), "pmvae_conv2d_backward")
    return dx
