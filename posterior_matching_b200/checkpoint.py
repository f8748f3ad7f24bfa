"""Run-directory formats of the reference scripts, host side only:

  train_state.pkl    pickle of bax's TrainState written by CheckpointCallback / train_pm_vae.py:104-106 and read by
                     eval_pm_vae_uci.py:79-80 (`model_state.params`, `model_state.state`)
  model_config.json  `config.model.to_dict()` (train_pm_vae.py:108-109, eval_pm_vae_uci.py:76-77)
  uci_results/nrmse.npy, uci_results/ac_lls.npy   (eval_pm_vae_uci.py:127-135)

bax, jax, haiku and optax are not installed here, so a reference checkpoint is read with an unpickler that maps every
class of those packages to a plain record (their instances only carry data: NamedTuples / FlatMappings of arrays; jax
DeviceArrays pickle as numpy arrays).  [R] the exact field list of bax.TrainState (step, rng, params, state, opt_state)
and haiku's FlatMapping pickling are recollection: the reader therefore looks parameters up structurally (the first
mapping of `{module_name: {leaf_name: array}}` found under `.params`), not by position.
"""
from __future__ import annotations

import io
import json
import os
import pickle
from typing import Any, Dict, Mapping, NamedTuple, Optional

import numpy as np

_FOREIGN = ("bax", "jax", "jaxlib", "haiku", "optax", "flax", "chex", "tensorflow_probability", "ml_collections")


class TrainState(NamedTuple):
    """Field names of bax's TrainState [R]; every leaf a numpy array."""
    step: Any
    rng: Any
    params: Any
    state: Any
    opt_state: Any


class ScaleByAdamState(NamedTuple):
    """optax.ScaleByAdamState [R]: count (int32 scalar), mu, nu (trees shaped like params)."""
    count: Any
    mu: Any
    nu: Any


class ForeignRecord:
    """Stand-in for an instance of a class from a package that is not installed: keeps what pickle hands over
    (constructor arguments via NEWOBJ / REDUCE, state via BUILD, items via SETITEMS)."""

    def __new__(cls, *args, **kwargs):
        self = object.__new__(cls)
        object.__setattr__(self, "_args", args)
        object.__setattr__(self, "_kwargs", kwargs)
        object.__setattr__(self, "_state", None)
        object.__setattr__(self, "_items", {})
        return self

    def __init__(self, *args, **kwargs):
        pass

    def __setstate__(self, state):
        object.__setattr__(self, "_state", state)

    def __setitem__(self, k, v):
        self._items[k] = v

    def _fields_dict(self) -> Dict[str, Any]:
        out: Dict[str, Any] = {}
        st = self._state
        if isinstance(st, tuple) and len(st) == 2 and isinstance(st[1], dict):     # (dict_state, slots_state)
            st = {**(st[0] or {}), **st[1]}
        if isinstance(st, dict):
            out.update(st)
        out.update(self._kwargs)
        return out

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        d = self._fields_dict()
        if name in d:
            return d[name]
        raise AttributeError(name)


class _ForeignTrainState(ForeignRecord):
    """bax.TrainState is a NamedTuple [R]: rebuilt positionally as this module's TrainState."""

    def __new__(cls, *a, **k):
        vals = list(a) + [None] * max(0, 5 - len(a))
        ts = TrainState(*vals[:5])
        return ts._replace(**{kk: v for kk, v in k.items() if kk in TrainState._fields}) if k else ts


class _ForeignAdamState(ForeignRecord):
    def __new__(cls, *a, **k):
        return ScaleByAdamState(*a, **k)


def _make_foreign(module: str, name: str):
    if name == "TrainState":
        return _ForeignTrainState
    if name == "ScaleByAdamState":
        return _ForeignAdamState
    return type(name, (ForeignRecord,), {"_foreign_origin": f"{module}.{name}"})


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        try:
            return super().find_class(module, name)
        except (ImportError, AttributeError):
            if module.split(".")[0] in _FOREIGN:
                return _make_foreign(module, name)
            raise


def _as_mapping(obj) -> Optional[Dict[str, Any]]:
    """dict view of a Mapping / ForeignRecord holding a mapping (haiku FlatMapping [R])."""
    if isinstance(obj, Mapping):
        return dict(obj)
    if isinstance(obj, ForeignRecord):
        if obj._items:
            return dict(obj._items)
        for cand in list(obj._args) + list(obj._fields_dict().values()):
            m = _as_mapping(cand)
            if m is not None:
                return m
    return None


def haiku_params(tree) -> Dict[str, Dict[str, np.ndarray]]:
    """`{module: {leaf: ndarray}}` out of whatever holds the Haiku parameters (dict, FlatMapping stand-in, ...)."""
    top = _as_mapping(tree)
    if top is None:
        raise ValueError("no parameter mapping found in the checkpoint")
    out: Dict[str, Dict[str, np.ndarray]] = {}
    for mod, leaves in top.items():
        lm = _as_mapping(leaves)
        if lm is None:
            raise ValueError(f"parameters of module {mod!r} are not a mapping")
        out[str(mod)] = {str(k): np.asarray(v) for k, v in lm.items()}
    return out


def load_train_state(path: str) -> TrainState:
    """Reads a reference `train_state.pkl` (or one written by `save_train_state`)."""
    with open(path, "rb") as fp:
        obj = _Unpickler(fp).load()
    if isinstance(obj, TrainState):
        return obj
    if isinstance(obj, Mapping):
        return TrainState(obj.get("step"), obj.get("rng"), obj.get("params"), obj.get("state"), obj.get("opt_state"))
    if hasattr(obj, "params"):
        g = lambda n: getattr(obj, n, None)                      # noqa: E731
        return TrainState(g("step"), g("rng"), g("params"), g("state"), g("opt_state"))
    raise ValueError(f"{path}: not a TrainState")


def load_params_into(model, path: str) -> TrainState:
    """eval_pm_vae_uci.py:76-80,113: the parameters of a run directory's train_state.pkl go into `model`'s arena."""
    ts = load_train_state(path)
    model.load_params(haiku_params(ts.params))
    return ts


def _tree_numpy(views) -> Dict[str, Dict[str, np.ndarray]]:
    return {mod: {k: np.array(v.detach().cpu().numpy() if hasattr(v, "detach") else v) for k, v in leaves.items()}
            for mod, leaves in views.items()}


def save_train_state(path: str, trainer) -> TrainState:
    """Writes the trainer's state with TrainState's field names (train_pm_vae.py:104-106): params, the optax chain's
    state as (ScaleByAdamState(count, mu, nu), ...) [R: the remaining links of the chain are stateless or hold only the
    schedule count], step, rng = the key of the per-step PRNGSequence."""
    mdl = trainer.model
    mu = _tree_numpy(mdl._views(trainer.m))
    nu = _tree_numpy(mdl._views(trainer.v))
    count = np.asarray(trainer.step, dtype=np.int32)
    ts = TrainState(step=np.asarray(trainer.step, dtype=np.int64), rng=np.asarray(trainer._rng.key, dtype=np.uint32),
                    params=_tree_numpy(mdl.params), state={},
                    opt_state=(ScaleByAdamState(count, mu, nu), (), {"count": count}, ()))
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "wb") as fp:
        pickle.dump(ts, fp)
    return ts


def restore_trainer(trainer, path: str) -> TrainState:
    """Resumes: parameters, Adam moments, step counter and the rng sequence key of a saved TrainState."""
    import torch
    ts = load_train_state(path)
    mdl = trainer.model
    mdl.load_params(haiku_params(ts.params))
    adam = None
    for link in (ts.opt_state if isinstance(ts.opt_state, (tuple, list)) else [ts.opt_state]):
        if isinstance(link, ScaleByAdamState) or (hasattr(link, "mu") and hasattr(link, "nu")):
            adam = link
            break
    if adam is not None:
        for arena, tree in ((trainer.m, adam.mu), (trainer.v, adam.nu)):
            views = mdl._views(arena)
            src = haiku_params(tree)
            for mod, leaves in views.items():
                for k, dst in leaves.items():
                    dst.copy_(torch.as_tensor(src[mod][k]).to(dst.device, torch.float32).reshape(dst.shape))
    if ts.step is not None:
        trainer.step = int(np.asarray(ts.step))
    if ts.rng is not None:
        trainer._rng.key = tuple(int(v) for v in np.asarray(ts.rng).reshape(-1)[:2])
    if getattr(trainer, "_fused", None) is not None:
        trainer._fused["stale"] = True
    return ts


def save_model_config(run_dir: str, model_config: Mapping[str, Any]) -> str:
    """train_pm_vae.py:108-109."""
    os.makedirs(run_dir, exist_ok=True)
    path = os.path.join(run_dir, "model_config.json")
    cfg = model_config.to_dict() if hasattr(model_config, "to_dict") else dict(model_config)
    with open(path, "w") as fp:
        json.dump(cfg, fp)
    return path


def load_model_config(run_dir: str) -> Dict[str, Any]:
    """eval_pm_vae_uci.py:76-77."""
    with open(os.path.join(run_dir, "model_config.json")) as fp:
        return json.load(fp)


def save_uci_results(run_dir: str, nrmse: np.ndarray, lls: np.ndarray) -> str:
    """eval_pm_vae_uci.py:127-135: uci_results/nrmse.npy ([num_trials]) and uci_results/ac_lls.npy ([num_trials])."""
    results_dir = os.path.join(run_dir, "uci_results")
    os.makedirs(results_dir, exist_ok=True)
    np.save(os.path.join(results_dir, "nrmse.npy"), np.asarray(nrmse))
    np.save(os.path.join(results_dir, "ac_lls.npy"), np.asarray(lls))
    return results_dir
