"""Registration of libpmvae.so's XLA custom-call targets with JAX, and the `jax.custom_vjp`
wrapper through which the reference's `loss_fn` (train_pm_vae.py:58-72) and `eval_fn`
(eval_pm_vae_uci.py:82-94) would call the CUDA path from inside `jax.jit`.

EXPERIMENTAL: JAX is NOT installed in the image this repository is built and tested in,
so this module is import-guarded and UNEXERCISED here: the targets themselves are tested
through ctypes with the exact calling convention XLA uses (tests/test_gpu_xla_shim.py);
the JAX side below is the thin part a maintainer checks once against their jaxlib.
INTEGRATION.md walks through it.
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Mapping

from . import _lib

WS_SLACK = 1024      # PMVAE_XLA_WS_SLACK (include/pmvae.h)
TARGETS = ("pmvae_xla_forward", "pmvae_xla_backward", "pmvae_xla_is_log_prob", "pmvae_xla_impute_mean",
           "pmvae_xla_mask_bernoulli")


def _capsule(name: str):
    """PyCapsule around the function pointer, named as XLA expects."""
    fn = getattr(_lib.lib, name)
    C.pythonapi.PyCapsule_New.restype = C.py_object
    C.pythonapi.PyCapsule_New.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
    return C.pythonapi.PyCapsule_New(C.cast(fn, C.c_void_p), b"xla._CUSTOM_CALL_TARGET", None)


def opaque(cfg: _lib.Config, *, B: int, K: int = 0, B_total: int = 0, row_start: int = 0, ws_bytes: int = 0,
           key0=(0, 0), key1=(0, 0), p: float = 0.5, prepare: bool = True, D: int = 0) -> bytes:
    """Serialises one pmvae_xla_opaque (include/pmvae.h)."""
    o = _lib.XlaOpaque()
    o.cfg = cfg
    o.B, o.K, o.B_total, o.row_start, o.ws_bytes = int(B), int(K), int(B_total or B), int(row_start), int(ws_bytes)
    o.key0[0], o.key0[1] = int(key0[0]) & 0xFFFFFFFF, int(key0[1]) & 0xFFFFFFFF
    o.key1[0], o.key1[1] = int(key1[0]) & 0xFFFFFFFF, int(key1[1]) & 0xFFFFFFFF
    o.p, o.prepare, o.D = float(p), int(bool(prepare)), int(D)
    return bytes(o)


def register() -> None:
    """Registers every target for platform "CUDA".  Raises ImportError without JAX.  The targets have the legacy
    (untyped) custom-call signature with a trailing XlaCustomCallStatus*: `api_version=0` here selects the legacy
    registration path of jax.ffi, and the call sites pass `custom_call_api_version=2` (API_VERSION_STATUS_RETURNING)."""
    import jax  # noqa: F401  (ImportError here = this environment has no JAX)
    try:  # jax >= 0.4.31
        from jax import ffi as jffi
        for name in TARGETS:
            jffi.register_ffi_target(name, _capsule(name), platform="CUDA", api_version=0)
    except (ImportError, AttributeError):  # jax 0.2.x - 0.4.30 (the reference pins 0.2.26)
        from jax.lib import xla_client
        for name in TARGETS:
            xla_client.register_custom_call_target(name.encode(), _capsule(name), platform="CUDA")


def make_pmvae_call(config: Mapping[str, Any], B: int, precision: str = "bf16"):
    """Returns `f(params_arena, x, b, eps) -> (rec, kl, match)` with a custom VJP whose
    backward is the second custom call; `params_arena` is the flat float32 arena
    (pmvae_layout order).  Needs jax >= 0.4.31 (`jax.ffi.ffi_call`)."""
    import jax
    import jax.numpy as jnp
    from jax import ffi as jffi

    from .vae import PosteriorMatchingVAE
    shell = PosteriorMatchingVAE.from_config(config, precision=precision)   # config -> pmvae_config, arena size
    cfg = shell.cfg
    ws_bytes = int(_lib.lib.pmvae_workspace_bytes(C.byref(cfg), B, 0))
    n_arena = shell.n_arena
    f32 = jnp.float32
    fwd_opaque = opaque(cfg, B=B, ws_bytes=ws_bytes, prepare=True)
    bwd_opaque = opaque(cfg, B=B, ws_bytes=ws_bytes, prepare=False)
    row = jax.ShapeDtypeStruct((B,), f32)
    # + PMVAE_XLA_WS_SLACK: XLA aligns buffers to <= 256 bytes, the targets round the pointer up to 1024
    ws_t = jax.ShapeDtypeStruct((ws_bytes + WS_SLACK,), jnp.uint8)

    def _fwd_call(arena, x, b, eps):
        return jffi.ffi_call("pmvae_xla_forward", (row, row, row, ws_t), custom_call_api_version=2,
                             legacy_backend_config=fwd_opaque)(arena, x, b, eps)

    @jax.custom_vjp
    def pmvae_call(arena, x, b, eps):
        rec, kl, match, _ = _fwd_call(arena, x, b, eps)
        return rec, kl, match

    def fwd(arena, x, b, eps):
        rec, kl, match, ws = _fwd_call(arena, x, b, eps)
        return (rec, kl, match), (arena, x, b, eps, ws)

    def bwd(res, cot):
        arena, x, b, eps, ws = res
        g_rec, g_kl, g_match = cot
        grads, _ = jffi.ffi_call("pmvae_xla_backward", (jax.ShapeDtypeStruct((n_arena,), f32), ws_t),
                                 custom_call_api_version=2, legacy_backend_config=bwd_opaque,
                                 input_output_aliases={7: 1})(arena, x, b, eps, g_rec, g_kl, g_match, ws)
        return grads, None, None, None     # x, b, eps are data: no cotangent is produced for them

    pmvae_call.defvjp(fwd, bwd)
    return pmvae_call
