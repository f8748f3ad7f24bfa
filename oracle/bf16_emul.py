"""The oracle with ONLY its dense-contraction operands rounded to bfloat16 (test infrastructure, see
oracle/__init__.py): every hk.Linear computes round(x) @ round(w) + b in float64, its backward
round(dy) @ round(w)^T and round(x)^T @ round(dy) -- the arithmetic contract of PMVAE_PREC_BF16 (bf16 operands,
fp32-or-better accumulation, dY carried in bf16 between Linears; DESIGN.md §5 Numerics) with everything else exact.
tests/test_oracle_bf16_emulation.py uses it to show which part of the CUDA path's gradient tolerance is intrinsic
to operand rounding."""
from __future__ import annotations

import contextlib

import torch

from . import model as M


def bf16_round(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.float32).to(torch.bfloat16).to(t.dtype)


class _BF16Linear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b):
        xr, wr = bf16_round(x), bf16_round(w)
        ctx.save_for_backward(xr, wr)
        return xr @ wr + b

    @staticmethod
    def backward(ctx, dy):
        xr, wr = ctx.saved_tensors
        dyr = bf16_round(dy)
        return dyr @ wr.t(), xr.t() @ dyr, dy.sum(0)


@contextlib.contextmanager
def bf16_operands():
    """Inside the context `oracle.model.linear` rounds its operands (forward and backward)."""
    orig = M.linear
    M.linear = lambda p, name, x: _BF16Linear.apply(x, p[name]["w"], p[name]["b"])
    try:
        yield
    finally:
        M.linear = orig
