"""Run-directory formats: a pickle shaped like the reference's train_state.pkl -- written here with
stand-in `bax` / `haiku` / `optax` modules, then read back with those modules gone, as on a machine without the JAX stack
-- model_config.json and the uci_results .npy files.  CPU only (no device work: parameters stay numpy arrays)."""
import json
import os
import pickle
import sys
import types
from collections import namedtuple

import numpy as np
import pytest

from posterior_matching_b200 import checkpoint as ck


class _FlatMapping(dict):
    """what haiku's FlatMapping looks like to pickle [R]: a Mapping rebuilt from a plain dict of dicts"""

    def __reduce__(self):
        return (type(self), (dict(self),))


def _fake_modules():
    bax_trainer = types.ModuleType("bax.trainer")
    bax = types.ModuleType("bax")
    TrainState = namedtuple("TrainState", ["step", "rng", "params", "state", "opt_state"])
    TrainState.__module__ = "bax.trainer"
    bax_trainer.TrainState = TrainState
    bax.trainer = bax_trainer
    optax_t = types.ModuleType("optax._src.transform")
    optax = types.ModuleType("optax")
    optax_src = types.ModuleType("optax._src")
    Adam = namedtuple("ScaleByAdamState", ["count", "mu", "nu"])
    Adam.__module__ = "optax._src.transform"
    optax_t.ScaleByAdamState = Adam
    hk_ds = types.ModuleType("haiku._src.data_structures")
    hk = types.ModuleType("haiku")
    hk_src = types.ModuleType("haiku._src")
    _FlatMapping.__module__ = "haiku._src.data_structures"
    _FlatMapping.__qualname__ = "FlatMapping"
    hk_ds.FlatMapping = _FlatMapping
    return {"bax": bax, "bax.trainer": bax_trainer, "optax": optax, "optax._src": optax_src,
            "optax._src.transform": optax_t, "haiku": hk, "haiku._src": hk_src, "haiku._src.data_structures": hk_ds}, \
        TrainState, Adam


def _params(rng):
    return {"encoder_net/linear": {"w": rng.standard_normal((8, 256)).astype(np.float32), "b": np.zeros(256, np.float32)},
            "decoder_dist": {"log_scale": np.float32(0.25)}}


def test_reference_shaped_pickle_is_read_without_the_jax_stack(tmp_path):
    mods, TrainState, Adam = _fake_modules()
    rng = np.random.default_rng(0)
    params = _params(rng)
    mu = {m: {k: np.ones_like(v) for k, v in leaves.items()} for m, leaves in params.items()}
    state = TrainState(step=np.int32(1234), rng=np.array([7, 9], np.uint32), params=_FlatMapping({m: _FlatMapping(l) for m, l in params.items()}),
                       state=_FlatMapping(), opt_state=(Adam(np.int32(1234), mu, mu), (), ()))
    path = tmp_path / "train_state.pkl"
    sys.modules.update(mods)
    try:
        with open(path, "wb") as fp:
            pickle.dump(state, fp)
    finally:
        for k in mods:
            sys.modules.pop(k, None)
    with pytest.raises((ImportError, ModuleNotFoundError, AttributeError)):
        with open(path, "rb") as fp:
            pickle.load(fp)                     # the plain unpickler needs bax / haiku / optax
    ts = ck.load_train_state(str(path))
    assert isinstance(ts, ck.TrainState) and int(ts.step) == 1234
    got = ck.haiku_params(ts.params)
    assert set(got) == set(params)
    for m in params:
        for k in params[m]:
            assert np.array_equal(got[m][k], params[m][k])
    adam = ts.opt_state[0]
    assert isinstance(adam, ck.ScaleByAdamState) and int(adam.count) == 1234
    assert np.array_equal(ck.haiku_params(adam.mu)["encoder_net/linear"]["w"], mu["encoder_net/linear"]["w"])


def test_model_config_and_results_files(tmp_path):
    from posterior_matching_b200.config import pm_vae_config
    cfg = pm_vae_config("gas").model
    ck.save_model_config(str(tmp_path), cfg)
    back = ck.load_model_config(str(tmp_path))
    assert back == json.loads(json.dumps(cfg.to_dict())) and back["latent_dim"] == 16
    d = ck.save_uci_results(str(tmp_path), np.array([0.5, 0.6]), np.array([-1.0, -1.1]))
    assert np.load(os.path.join(d, "nrmse.npy")).shape == (2,) and np.load(os.path.join(d, "ac_lls.npy"))[1] == -1.1
