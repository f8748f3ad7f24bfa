"""CPU-side checks of the boundary: the library loads, exports every symbol the header
declares, and its host logic (layout, key utilities, schedules, config surface) agrees
with the oracle.  No device calls."""
import ctypes as C
import importlib.util
import os
import re

import numpy as np
import pytest

from oracle import model as M, prng as oprng
from posterior_matching_b200 import _lib, prng, train
from posterior_matching_b200.config import pm_vae_config

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_header_symbol():
    hdr = open(os.path.join(ROOT, "include", "pmvae.h")).read()
    declared = set(re.findall(r"\b(pmvae_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(_lib.lib, name), f"libpmvae.so does not export {name}"
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    assert _lib.lib.pmvae_version() >= 1


@pytest.mark.parametrize("name", ["gas", "power", "hepmass", "bsds"])
def test_layout_matches_oracle_leaves(name):
    spec = M.spec_from_config(pm_vae_config(name).model.to_dict())
    cfg = _lib.make_config(spec.D, spec.d, spec.H, spec.R_enc, spec.R_dec, spec.R_part, spec.ln_enc, spec.ln_dec,
                           spec.ln_part, spec.stop_grad, _lib.PREC_F32)
    leaves = _lib.layout(cfg)
    want = M.leaf_shapes(spec)
    got = [(n, r, c) for n, r, c, _, _ in leaves if (r, c) != (0, 0)]
    assert got == want
    assert [n for n, r, c, _, _ in leaves if (r, c) == (0, 0)] == ["decoder_dist"]
    # leaves do not overlap and fit in the arena
    total = _lib.lib.pmvae_param_count(C.byref(cfg))
    spans = []
    for n, r, c, w, b in leaves:
        if (r, c) == (0, 0):
            spans.append((w, w + 1))
        else:
            spans += [(w, w + r * c), (b, b + c)]
    spans.sort()
    assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:])) and spans[-1][1] <= total
    assert sum(e - s for s, e in spans) == M.n_params(spec)


def test_bad_config_reports_error():
    cfg = _lib.make_config(8, 65, 256, 2, 2, 2, 0, 0, 0, 1, 0)
    assert _lib.lib.pmvae_param_count(C.byref(cfg)) == 0
    assert b"latent_dim" in _lib.lib.pmvae_last_error()
    with pytest.raises(_lib.PmvaeError):
        _lib.check(_lib.lib.pmvae_key_split_host(None, 2, None), "split")


def test_host_keys_match_oracle():
    for seed in (0, 42, 2 ** 40 + 17):
        k = prng.PRNGKey(seed)
        assert list(k) == oprng.PRNGKey(seed).tolist()
        assert [list(s) for s in prng.split(k, 5)] == oprng.split(oprng.PRNGKey(seed), 5).tolist()
        assert list(prng.fold_in(k, 9)) == oprng.fold_in(oprng.PRNGKey(seed), 9).tolist()
    a, b = prng.PRNGSequence(91), oprng.PRNGSequence(91)
    for _ in range(4):
        assert list(a.next()) == b.next().tolist()


def test_schedules_match_oracle():
    for name in ("gas", "bsds"):
        cfg = pm_vae_config(name)
        mine, ref = train.get_beta_schedule(cfg.beta), M.beta_schedule(cfg.beta.to_dict())
        for step in (0, 999, 1000, 1001, 13500, 26000, 30000, 51000, 63500, 130000, 250000):
            assert abs(mine(step) - ref(step)) < 1e-12
    assert train.get_beta_schedule({})(3) == 1.0
    lr, ref = train.exponential_decay(1e-3, 5000, 0.9), M.lr_schedule(1e-3, 0.9, 5000)
    assert all(abs(lr(t) - ref(t)) < 1e-18 for t in (0, 1, 5000, 123456))


@pytest.mark.parametrize("name", ["gas", "power", "hepmass", "bsds", "mnist"])
def test_config_files_keep_the_reference_surface(name):
    path = os.path.join(ROOT, "configs", f"pm_vae_{name}.py")
    spec = importlib.util.spec_from_file_location(f"cfg_{name}", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    cfg = mod.get_config()
    assert cfg.data.dataset == name and "model" in cfg and cfg.lr_schedule.init_value == 0.001
    cfg.lock()
    d = cfg.model.to_dict()
    assert d["latent_dim"] == {"gas": 16, "power": 16, "hepmass": 16, "bsds": 64, "mnist": 32}[name]
    if name != "mnist":
        assert cfg.model.decoder_dist_config.event_size == {"gas": 8, "power": 6, "hepmass": 21, "bsds": 63}[name]
        assert cfg.model.matching_ll_stop_gradients is True and cfg.weight_decay == 1e-5
        assert cfg.data.train_batch_size == 512 and cfg.steps == 200000
    else:
        assert cfg.model.partial_posterior_dist == "AutoregressiveGMM" and cfg.data.train_batch_size == 256


def test_mnist16_and_lookahead_config_files():
    """configs/pm_vae_mnist16.py and configs/lookahead_mnist16.py carry the reference's keys and values."""
    mods = {}
    for fname in ("pm_vae_mnist16", "lookahead_mnist16"):
        spec = importlib.util.spec_from_file_location(f"cfg_{fname}", os.path.join(ROOT, "configs", f"{fname}.py"))
        mods[fname] = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mods[fname])
    cfg = mods["pm_vae_mnist16"].get_config()
    assert cfg.data.dataset == "mnist16" and cfg.data.mask_generator == "UniformMaskGenerator"
    assert tuple(cfg.data.mask_generator_kwargs.bounds) == (0.0, 0.2) and cfg.data.train_batch_size == 128
    assert cfg.model.latent_dim == 10 and "partial_posterior_dist" not in cfg.model      # falls back to posterior_dist
    assert [tuple(l) for l in cfg.model.encoder_net_config.conv_layers] == [(32, 3, 1), (32, 3, 2), (64, 3, 2), (64, 1, 1)]
    assert [tuple(l) for l in cfg.model.decoder_net_config.conv_layers][0] == (64, 8, 1) and cfg.steps == 200000
    look = mods["lookahead_mnist16"].get_config()
    assert look.model.lookahead_subsample == 16 and look.model.model_samples == 64 and look.steps == 40000
    assert look.data.train_batch_size == 32 and "pm_vae_dir" in look and look.lr_schedule.decay_rate == 0.9


def test_device_entry_points_fail_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from posterior_matching_b200 import PosteriorMatchingVAE
    with pytest.raises(RuntimeError):
        PosteriorMatchingVAE.from_config(pm_vae_config("gas").model)


def test_xla_opaque_layout_and_guarded_jax_module():
    """The ctypes mirror of pmvae_xla_opaque matches the C struct; the JAX registration module imports
    without JAX and only fails (ImportError) when asked to register."""
    assert _lib.lib.pmvae_xla_opaque_size() == C.sizeof(_lib.XlaOpaque)
    from posterior_matching_b200 import jax_ffi
    cfg = _lib.make_config(8, 16, 256, 2, 2, 2, 0, 0, 0, 1, _lib.PREC_BF16)
    blob = jax_ffi.opaque(cfg, B=5, K=7, key0=(1, 2), ws_bytes=99)
    assert len(blob) == C.sizeof(_lib.XlaOpaque)
    back = _lib.XlaOpaque.from_buffer_copy(blob)
    assert (back.B, back.K, back.B_total, back.ws_bytes, back.key0[1], back.cfg.d) == (5, 7, 5, 99, 2, 16)
    try:
        import jax  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError):
            jax_ffi.register()
