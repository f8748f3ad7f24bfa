"""The XLA custom-call targets, driven through ctypes with XLA's calling convention
(stream, buffers = operands then results, opaque bytes, status): same results as the
direct C ABI.  (JAX itself is not installed here; posterior_matching_b200/jax_ffi.py is
the registration a maintainer adds.)"""
import ctypes as C

import numpy as np
import pytest
import torch

from tests.util import conditioned_params, make_inputs, spec_of

pytestmark = pytest.mark.gpu


def _aligned_ws(nbytes):
    """A workspace the way XLA hands it over: ws_bytes + PMVAE_XLA_WS_SLACK bytes, aligned to 256 only (the targets
    round the pointer up to the 1024 bytes the tensor path needs)."""
    raw = torch.empty(nbytes + 2048, dtype=torch.uint8, device="cuda")
    off = (256 - raw.data_ptr()) % 1024
    ws = raw[off:off + nbytes + 1024]
    assert ws.data_ptr() % 1024 == 256
    return ws


def _buffers(*tensors):
    arr = (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
    return arr


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_forward_backward_targets_match_direct_calls(precision):
    from posterior_matching_b200 import PosteriorMatchingVAE, pm_vae_config, _lib
    from posterior_matching_b200.jax_ffi import opaque
    spec = spec_of("gas")
    p = conditioned_params(spec)
    B = 200
    x, b, eps = (t.float().cuda() for t in make_inputs(spec, B, seed=5))
    m = PosteriorMatchingVAE.from_config(pm_vae_config("gas").model, precision=precision)
    m.load_params(p)
    want = {k: v.clone() for k, v in m(x, b, eps=eps).items()}
    g = [torch.full((B,), 1.0 / B, device="cuda"), torch.full((B,), -0.3 / B, device="cuda"),
         torch.full((B,), 1.0 / B, device="cuda")]
    want_grads = m.backward(*g)
    want_arena = m.grad_arena.clone()

    S = torch.cuda.current_stream().cuda_stream
    ws_bytes = int(_lib.lib.pmvae_workspace_bytes(C.byref(m.cfg), B, 0))
    ws = _aligned_ws(ws_bytes)
    out = torch.empty(3, B, device="cuda")
    op = opaque(m.cfg, B=B, ws_bytes=ws_bytes, prepare=True)
    _lib.lib.pmvae_xla_forward(S, _buffers(m.arena, x, b, eps, out[0], out[1], out[2], ws), op, len(op), None)
    grads = torch.full_like(m.arena, 7.0)
    op2 = opaque(m.cfg, B=B, ws_bytes=ws_bytes, prepare=False)
    _lib.lib.pmvae_xla_backward(S, _buffers(m.arena, x, b, eps, g[0], g[1], g[2], ws, grads, ws), op2, len(op2), None)
    torch.cuda.synchronize()
    for i, k in enumerate(("reconstruction_ll", "kl", "matching_ll")):
        assert torch.equal(out[i], want[k]), k
    if precision == "fp32":
        # same kernels, same order; only the atomics of the weight-gradient reduction may reorder
        assert float((grads - want_arena).abs().max()) <= 1e-5 * float(want_arena.abs().max())
    else:
        assert float((grads - want_arena).norm() / want_arena.norm()) < 1e-3


def test_eval_and_mask_targets_match_direct_calls():
    from posterior_matching_b200 import PosteriorMatchingVAE, pm_vae_config, _lib
    from posterior_matching_b200.jax_ffi import opaque
    from posterior_matching_b200.masking import get_mask_generator
    spec = spec_of("power")
    p = conditioned_params(spec)
    B, K = 24, 32
    x, b, _ = (t.float().cuda() for t in make_inputs(spec, B, seed=6))
    m = PosteriorMatchingVAE.from_config(pm_vae_config("power").model, precision="bf16")
    m.load_params(p)
    keys = ((11, 12), (13, 14))
    lpx, cond = (t.clone() for t in m.is_log_prob(x, b, K, keys=keys))
    imp = m.impute_mean(x, b, K, key=(5, 6)).clone()
    S = torch.cuda.current_stream().cuda_stream
    ws_bytes = int(_lib.lib.pmvae_workspace_bytes(C.byref(m.cfg), B, K))
    ws = _aligned_ws(ws_bytes)
    o1, o2 = torch.empty(B, device="cuda"), torch.empty(B, device="cuda")
    op = opaque(m.cfg, B=B, K=K, ws_bytes=ws_bytes, key0=keys[0], key1=keys[1], prepare=True)
    _lib.lib.pmvae_xla_is_log_prob(S, _buffers(m.arena, x, b, o1, o2, ws), op, len(op), None)
    o3 = torch.empty(B, spec.D, device="cuda")
    op = opaque(m.cfg, B=B, K=K, ws_bytes=ws_bytes, key0=(5, 6), prepare=False)
    _lib.lib.pmvae_xla_impute_mean(S, _buffers(m.arena, x, b, o3, ws), op, len(op), None)
    mk = torch.empty(B, spec.D, device="cuda")
    op = opaque(m.cfg, B=B, B_total=B, key0=(3, 9), p=0.5, D=spec.D)
    _lib.lib.pmvae_xla_mask_bernoulli(S, _buffers(mk), op, len(op), None)
    torch.cuda.synchronize()
    assert torch.equal(o1, lpx) and torch.equal(o2, cond) and torch.equal(o3, imp)
    from oracle import prng as oprng
    want = oprng.bernoulli(np.array([3, 9], dtype=np.uint32), 0.5, (B, spec.D)).astype(np.float32)
    assert np.array_equal(mk.cpu().numpy(), want)


def test_bad_opaque_is_reported_not_executed(capfd):
    from posterior_matching_b200 import _lib
    out = torch.zeros(4, device="cuda")
    _lib.lib.pmvae_xla_mask_bernoulli(torch.cuda.current_stream().cuda_stream, _buffers(out), b"xx", 2, None)
    torch.cuda.synchronize()
    assert float(out.abs().max()) == 0.0
    assert b"opaque" in _lib.lib.pmvae_last_error()
