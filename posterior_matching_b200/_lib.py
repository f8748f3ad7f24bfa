"""ctypes binding of libpmvae.so (the C ABI in include/pmvae.h).

There is no CPU fallback: importing this module without the built library, or
calling a device entry point without a CUDA device, raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libpmvae.so")

PREC_F32, PREC_BF16 = 0, 1
NET_SAVE = 0x100      # PMVAE_NET_SAVE


class Config(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("D", "d", "H", "R_enc", "R_dec", "R_part", "ln_enc", "ln_dec", "ln_part", "stop_grad", "precision")]
    _fields_.append(("reserved", C.c_int32 * 5))


class Leaf(C.Structure):
    _fields_ = [("name", C.c_char * 64), ("rows", C.c_int32), ("cols", C.c_int32),
                ("w_off", C.c_uint64), ("b_off", C.c_uint64)]


class ArgmmConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("d", "n_comp", "R", "H", "C")] + [("reserved", C.c_int32 * 3)]


class TrainConfig(C.Structure):
    """pmvae_train_config (include/pmvae.h)."""
    _fields_ = [("beta_schedule", C.c_int32), ("beta_low", C.c_float), ("beta_high", C.c_float),
                ("beta_period", C.c_int64), ("beta_delay", C.c_int64), ("beta_transition_steps", C.c_int64),
                ("beta_transition_begin", C.c_int64), ("matching_coef", C.c_float), ("lr_init", C.c_float),
                ("lr_decay_rate", C.c_float), ("lr_transition_steps", C.c_int64), ("weight_decay", C.c_float),
                ("adam_b1", C.c_float), ("adam_b2", C.c_float), ("adam_eps", C.c_float), ("mask_p", C.c_float),
                ("reserved", C.c_int32 * 3)]


class TrainStateHost(C.Structure):
    _fields_ = [("seq_key", C.c_uint32 * 2), ("mask_key", C.c_uint32 * 2), ("eps_key", C.c_uint32 * 2),
                ("mask_calls", C.c_uint32), ("pad", C.c_uint32), ("step", C.c_int64), ("beta", C.c_float),
                ("lr", C.c_float), ("bc1", C.c_float), ("bc2", C.c_float)]


BETA_CONST, BETA_CYCLIC, BETA_MONOTONIC = 0, 1, 2


class ConvDesc(C.Structure):
    """pmvae_conv_desc (include/pmvae.h)."""
    _fields_ = [(n, C.c_int32) for n in ("H", "W", "Cin", "OH", "OW", "Cout", "KH", "KW", "stride", "dil", "pad_top",
                                         "pad_left", "w_ci", "w_co")] + [("slope", C.c_float), ("reserved", C.c_int32)]


class XlaOpaque(C.Structure):
    """pmvae_xla_opaque (include/pmvae.h): the `opaque` descriptor of the XLA custom-call targets."""
    _fields_ = [("cfg", Config), ("B", C.c_int64), ("K", C.c_int64), ("B_total", C.c_int64), ("row_start", C.c_int64),
                ("ws_bytes", C.c_uint64), ("key0", C.c_uint32 * 2), ("key1", C.c_uint32 * 2), ("p", C.c_float),
                ("prepare", C.c_int32), ("D", C.c_int32), ("reserved", C.c_int32)]


class PmvaeError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python posterior_matching_b200/csrc/build.py` "
            "(nvcc, sm_100a).  The PM-VAE hot path has no CPU fallback.")
    return C.CDLL(LIB_PATH)


lib = _load()

_u32p = C.POINTER(C.c_uint32)
_vp = C.c_void_p
_cfgp = C.POINTER(Config)
_i64, _u64, _f32, _i32 = C.c_int64, C.c_uint64, C.c_float, C.c_int32

_SIGS = {
    "pmvae_last_error": (C.c_char_p, []),
    "pmvae_version": (_i32, []),
    "pmvae_launch_count": (_u64, []),
    "pmvae_linear": (_i32, [_i32, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp, _u64, _vp]),
    "pmvae_tc_gemm_nt": (_i32, [_vp, _i64, _vp, _i64, _vp, _i64, _i32, _i32, _vp, _vp]),
    "pmvae_tc_gemm_tn": (_i32, [_vp, _i64, _vp, _i64, _i32, _i32, _i64, _vp, _vp]),
    "pmvae_param_count": (_u64, [_cfgp]),
    "pmvae_layout": (_i32, [_cfgp, C.POINTER(Leaf), _i32]),
    "pmvae_key_split_host": (_i32, [_u32p, _i32, _u32p]),
    "pmvae_key_fold_in_host": (_i32, [_u32p, C.c_uint32, _u32p]),
    "pmvae_random_bits": (_i32, [_u32p, _u64, _u64, _u64, _vp, _vp]),
    "pmvae_uniform": (_i32, [_u32p, _u64, _u64, _u64, _vp, _vp]),
    "pmvae_normal": (_i32, [_u32p, _u64, _u64, _u64, _vp, _vp]),
    "pmvae_mask_bernoulli": (_i32, [_u32p, _f32, _u64, _u64, _u64, _i32, _vp, _vp]),
    "pmvae_mask_mnist": (_i32, [_u32p, _u64, _u64, _u64, _vp, _vp]),
    "pmvae_workspace_bytes": (_u64, [_cfgp, _i64, _i64]),
    "pmvae_prepare_params": (_i32, [_cfgp, _vp, _vp, _u64, _vp]),
    "pmvae_forward": (_i32, [_cfgp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _u64, _vp]),
    "pmvae_backward": (_i32, [_cfgp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _u64, _vp]),
    "pmvae_loss_cotangents": (_i32, [_i64, _i64, _f32, _f32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pmvae_adamw": (_i32, [_cfgp, _vp, _vp, _vp, _vp, _i64, _f32, _f32, _f32, _f32, _f32, _vp]),
    "pmvae_is_log_prob": (_i32, [_cfgp, _vp, _vp, _vp, _i64, _i64, _u32p, _u32p, _i64, _i64, _vp, _vp, _vp, _u64, _vp]),
    "pmvae_net_apply": (_i32, [_cfgp, _vp, _i32, _vp, _vp, _i64, _vp, _vp, _u64, _vp]),
    "pmvae_train_state_bytes": (_u64, []),
    "pmvae_train_scratch_floats": (_u64, [_cfgp, _i64]),
    "pmvae_train_state_init": (_i32, [_vp, _u32p, _u32p, _i64, C.c_uint32, _vp]),
    "pmvae_train_state_read": (_i32, [_vp, C.POINTER(TrainStateHost), _vp]),
    "pmvae_train_step": (_i32, [_cfgp, C.POINTER(TrainConfig), _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp,
                                _vp, _u64, _i32, _vp]),
    "pmvae_bernoulli_ll": (_i32, [_vp, _vp, _vp, _i64, _i32, _vp, _vp]),
    "pmvae_bernoulli_ll_backward": (_i32, [_vp, _vp, _vp, _vp, _i64, _i32, _vp, _vp]),
    "pmvae_argmm_param_count": (_u64, [C.POINTER(ArgmmConfig)]),
    "pmvae_argmm_layout": (_i32, [C.POINTER(ArgmmConfig), C.POINTER(Leaf), _i32]),
    "pmvae_argmm_workspace_bytes": (_u64, [C.POINTER(ArgmmConfig), _i64]),
    "pmvae_argmm_log_prob": (_i32, [C.POINTER(ArgmmConfig), _vp, _vp, _vp, _i64, _vp, _vp, _u64, _vp]),
    "pmvae_argmm_backward": (_i32, [C.POINTER(ArgmmConfig), _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _u64, _vp]),
    "pmvae_argmm_sample_workspace_bytes": (_u64, [C.POINTER(ArgmmConfig), _i64, _i64]),
    "pmvae_argmm_sample": (_i32, [C.POINTER(ArgmmConfig), _vp, _vp, _i64, _i64, _u32p, _vp, _vp, _u64, _vp]),
    "pmvae_logmeanexp_rows": (_i32, [_vp, _vp, _vp, _i64, _i64, _vp]),
    "pmvae_linear_backward": (_i32, [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp, _vp, _vp]),
    "pmvae_tril_sample_kl": (_i32, [_vp, _vp, _i64, _i32, _vp, _vp, _vp]),
    "pmvae_tril_sample_kl_backward": (_i32, [_vp, _vp, _vp, _vp, _i64, _i32, _vp, _vp]),
    "pmvae_adamw_flat": (_i32, [_vp, _vp, _vp, _vp, _u64, _i64, _f32, _f32, _f32, _f32, _f32, _vp]),
    "pmvae_conv2d_workspace_bytes": (_u64, [C.POINTER(ConvDesc), _i64]),
    "pmvae_conv2d_forward": (_i32, [C.POINTER(ConvDesc), _vp, _vp, _vp, _i64, _vp, _vp, _u64, _vp]),
    "pmvae_conv2d_backward": (_i32, [C.POINTER(ConvDesc), _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _u64, _vp]),
    "pmvae_xla_opaque_size": (_u64, []),
    "pmvae_xla_forward": (None, [_vp, C.POINTER(_vp), C.c_char_p, C.c_size_t, _vp]),
    "pmvae_xla_backward": (None, [_vp, C.POINTER(_vp), C.c_char_p, C.c_size_t, _vp]),
    "pmvae_xla_is_log_prob": (None, [_vp, C.POINTER(_vp), C.c_char_p, C.c_size_t, _vp]),
    "pmvae_xla_impute_mean": (None, [_vp, C.POINTER(_vp), C.c_char_p, C.c_size_t, _vp]),
    "pmvae_xla_mask_bernoulli": (None, [_vp, C.POINTER(_vp), C.c_char_p, C.c_size_t, _vp]),
    "pmvae_impute_mean": (_i32, [_cfgp, _vp, _vp, _vp, _i64, _i64, _u32p, _i64, _i64, _vp, _vp, _u64, _vp]),
    "pmvae_tril_log_prob": (_i32, [_vp, _vp, _i64, _i32, _vp, _vp]),
    "pmvae_tril_entropy": (_i32, [_vp, _i64, _i32, _vp, _vp]),
    "pmvae_tril_sample": (_i32, [_vp, _u32p, _i64, _i64, _i64, _i64, _i32, _vp, _vp, _vp]),
    "pmvae_normal_log_prob": (_i32, [_vp, _vp, _vp, _i64, _i64, _i32, _vp, _vp]),
    "pmvae_std_normal_log_prob": (_i32, [_vp, _i64, _i32, _vp, _vp]),
    "pmvae_diag_sample": (_i32, [_vp, _vp, _i64, _i32, _vp, _vp]),
    "pmvae_diag_log_prob": (_i32, [_vp, _vp, _i64, _i32, _vp, _vp, _vp]),
    "pmvae_impute": (_i32, [_cfgp, _vp, _vp, _vp, _i64, _i64, _u32p, _i64, _i64, _vp, _vp, _vp, _u64, _vp]),
    "pmvae_tril_log_prob_backward": (_i32, [_vp, _vp, _vp, _i64, _i32, _vp, _vp, _vp, _u64, _vp]),
    "pmvae_lookahead_ll": (_i32, [_vp, _vp, _vp, _i64, _i64, _i32, _i32, _vp, _vp]),
    "pmvae_lookahead_ll_backward": (_i32, [_vp, _vp, _vp, _vp, _i64, _i64, _i32, _i32, _vp, _vp]),
}
EXPORTS = tuple(_SIGS)

for _name, (_res, _args) in _SIGS.items():
    _fn = getattr(lib, _name)  # AttributeError here = the library does not export the header's symbol
    _fn.restype = _res
    _fn.argtypes = _args


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib.pmvae_last_error()
        raise PmvaeError(f"{what} failed (rc={rc}): {msg.decode() if msg else '?'}")


def key_arg(key):
    """(k0, k1) -> ctypes uint32[2]."""
    arr = (C.c_uint32 * 2)(int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF)
    return arr


def make_config(D, d, H, R_enc, R_dec, R_part, ln_enc, ln_dec, ln_part, stop_grad, precision) -> Config:
    c = Config()
    c.D, c.d, c.H = int(D), int(d), int(H)
    c.R_enc, c.R_dec, c.R_part = int(R_enc), int(R_dec), int(R_part)
    c.ln_enc, c.ln_dec, c.ln_part = int(bool(ln_enc)), int(bool(ln_dec)), int(bool(ln_part))
    c.stop_grad = int(bool(stop_grad))
    c.precision = int(precision)
    return c


def layout(cfg: Config):
    n = lib.pmvae_layout(C.byref(cfg), None, 0)
    if n < 0:
        check(1, "pmvae_layout")
    arr = (Leaf * n)()
    lib.pmvae_layout(C.byref(cfg), arr, n)
    return [(l.name.decode(), l.rows, l.cols, int(l.w_off), int(l.b_off)) for l in arr]
