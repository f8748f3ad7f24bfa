"""The `configs/pm_vae_*.py` surface (reference: configs/pm_vae_{gas,power,hepmass,
bsds,mnist}.py) rebuilt from one table.

`configs/<name>.py:get_config()` returns the same tree of keys and values the
reference's files build with `ml_collections.ConfigDict`; ml_collections is not
installed in this image, so a minimal attribute-dict stands in when it is
missing (it supports what train_pm_vae.py:47-52,58-83 and vae.py:74-117 use:
attribute and item access, `in`, `.get`, `.lock()`, `.to_dict()`).
"""
from __future__ import annotations

from typing import Any, Dict

try:  # pragma: no cover - not installed in this image
    from ml_collections import ConfigDict  # type: ignore
except Exception:  # noqa: BLE001
    class ConfigDict(dict):
        """Small stand-in for ml_collections.ConfigDict."""

        def __init__(self, initial: Dict[str, Any] | None = None):
            super().__init__()
            object.__setattr__(self, "_locked", False)
            for k, v in (initial or {}).items():
                self[k] = v

        def __setitem__(self, k, v):
            if self._locked and k not in self:
                raise KeyError(f"config is locked; cannot add {k!r}")
            if isinstance(v, dict) and not isinstance(v, ConfigDict):
                v = ConfigDict(v)
            super().__setitem__(k, v)

        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError as e:
                raise AttributeError(k) from e

        def __setattr__(self, k, v):
            self[k] = v

        def lock(self):
            object.__setattr__(self, "_locked", True)
            for v in self.values():
                if isinstance(v, ConfigDict):
                    v.lock()
            return self

        def to_dict(self):
            return {k: (v.to_dict() if isinstance(v, ConfigDict) else v) for k, v in self.items()}


# name -> (features D, latent d, residual blocks, layer_norm, beta schedule)
_UCI = {
    "gas": (8, 16, 2, False, "cyclic"),
    "power": (6, 16, 2, False, "cyclic"),
    "hepmass": (21, 16, 2, False, "cyclic"),
    "bsds": (63, 64, 5, True, "monotonic"),
}
DATASET_FEATURES = {k: v[0] for k, v in _UCI.items()}


def uci_config(name: str) -> ConfigDict:
    D, d, R, ln, sched = _UCI[name]
    net = {"residual_blocks": R, "hidden_units": 256, "layer_norm": ln}
    if sched == "cyclic":
        beta = {"schedule": "cyclic", "low_value": 0.0, "high_value": 1.0, "period": 50000, "delay": 1000}
    else:
        beta = {"schedule": "monotonic", "low_value": 0.0, "high_value": 1.0,
                "transition_steps": 200000, "transition_begin": 30000}
    return ConfigDict({
        "data": {"dataset": name, "train_split": "train", "validation_split": "val",
                 "train_batch_size": 512, "val_batch_size": 512, "training_noise": 0.001,
                 "mask_generator": "BernoulliMaskGenerator"},
        "model": {"latent_dim": d, "encoder_net": "ResidualMLP", "decoder_net": "ResidualMLP",
                  "decoder_dist": "IdentityGaussian", "posterior_dist": "TriLGaussian",
                  "decoder_dist_config": {"event_size": D},
                  # never read by from_config; kept because the reference carries them
                  "masked_posterior_dist": "AutoregressiveGMM",
                  "masked_posterior_config": {"hidden_units": 256, "residual_blocks": 3},
                  "encoder_net_config": dict(net), "decoder_net_config": dict(net),
                  "matching_ll_stop_gradients": True},
        "beta": beta,
        "steps": 200000, "validation_freq": 1000, "save_final_state": True,
        "weight_decay": 0.00001,
        "lr_schedule": {"init_value": 0.001, "decay_rate": 0.9, "transition_steps": 5000},
    })


def mnist_config() -> ConfigDict:
    return ConfigDict({
        "data": {"dataset": "mnist", "train_split": "train", "validation_split": "test",
                 "train_batch_size": 256, "val_batch_size": 256, "mask_generator": "MNISTMaskGenerator"},
        "model": {"latent_dim": 32, "encoder_net": "ConvEncoder", "decoder_net": "ConvDecoder",
                  "posterior_dist": "TriLGaussian", "partial_posterior_dist": "AutoregressiveGMM",
                  "decoder_dist": "Bernoulli",
                  "encoder_net_config": {"conv_layers": [(32, 5, 1), (32, 5, 2), (64, 5, 1), (64, 5, 2), (128, 7, 1)]},
                  "decoder_net_config": {"conv_layers": [(64, 7, 1), (64, 5, 2), (32, 5, 1), (32, 5, 2),
                                                         (32, 5, 1), (1, 5, 1)]}},
        "steps": 80000, "validation_freq": 1000,
        "lr_schedule": {"init_value": 0.001, "decay_rate": 0.9, "transition_steps": 5000},
    })


def mnist16_config() -> ConfigDict:
    """configs/pm_vae_mnist16.py: the model `LookaheadPosterior` is trained over (TriLGaussian posterior and, by the
    fallback of vae.py:97-105, partial posterior)."""
    return ConfigDict({
        "data": {"dataset": "mnist16", "train_split": "train", "validation_split": "test",
                 "train_batch_size": 128, "val_batch_size": 128, "mask_generator": "UniformMaskGenerator",
                 "mask_generator_kwargs": {"bounds": (0.0, 0.2)}},
        "model": {"latent_dim": 10, "encoder_net": "ConvEncoder", "decoder_net": "ConvDecoder",
                  "posterior_dist": "TriLGaussian", "decoder_dist": "Bernoulli",
                  "encoder_net_config": {"conv_layers": [(32, 3, 1), (32, 3, 2), (64, 3, 2), (64, 1, 1)]},
                  "decoder_net_config": {"conv_layers": [(64, 8, 1), (64, 5, 2), (32, 5, 1), (32, 5, 1), (1, 3, 1)]}},
        "steps": 200000, "validation_freq": 10000,
        "lr_schedule": {"init_value": 0.001, "decay_rate": 0.9, "transition_steps": 5000},
    })


def lookahead_mnist16_config() -> ConfigDict:
    """configs/lookahead_mnist16.py (train_lookahead_posterior.py reads `pm_vae_dir` for the frozen model)."""
    return ConfigDict({
        "data": {"dataset": "mnist16", "train_split": "train", "validation_split": "test",
                 "train_batch_size": 32, "val_batch_size": 32, "mask_generator": "UniformMaskGenerator",
                 "mask_generator_kwargs": {"bounds": (0.0, 0.20)}},
        "pm_vae_dir": "runs/pm-vae-mnist16-20220302-160842",
        "model": {"lookahead_subsample": 16, "model_samples": 64},
        "steps": 40000, "validation_freq": 5000,
        "lr_schedule": {"init_value": 0.001, "decay_rate": 0.9, "transition_steps": 5000},
    })


def pm_vae_config(name: str) -> ConfigDict:
    if name == "mnist":
        return mnist_config()
    if name == "mnist16":
        return mnist16_config()
    return uci_config(name)
