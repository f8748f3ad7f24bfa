"""eval_pm_vae_uci.py's eval_fn (:82-94) and NRMSE (:60-66) over the CUDA model."""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch

from .vae import PosteriorMatchingVAE


def eval_fn(model: PosteriorMatchingVAE, rng, x: torch.Tensor, b: torch.Tensor, num_samples: int = 512, *,
            row_start: int = 0, total_rows=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """(mean-over-K imputation [B, D], log p(x_u | x_o) [B]) for one batch; `rng` is the
    key `prng.next()` yields at eval_pm_vae_uci.py:113."""
    k_imp, k_z, k_zxo = model.eval_keys(rng)
    imputed = model.impute_mean(x, b, num_samples, key=k_imp, row_start=row_start, total_rows=total_rows)
    _, ll = model.is_log_prob(x, b, num_samples, keys=(k_z, k_zxo), row_start=row_start, total_rows=total_rows)
    return imputed, ll


def nrmse_score(imputations: np.ndarray, true_data: np.ndarray, observed_mask: np.ndarray) -> np.ndarray:
    """eval_pm_vae_uci.py:60-66."""
    error = (imputations - true_data) ** 2
    mse = np.sum(error, axis=-2) / np.count_nonzero(1.0 - observed_mask, axis=-2)
    nrmse = np.sqrt(mse) / np.std(true_data, axis=-2)
    return np.mean(nrmse, axis=-1)
