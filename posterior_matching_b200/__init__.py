"""B200-native (sm_100a) Posterior-Matching VAE hot path.

Drop-in for the reference's `posterior_matching.models.vae.PosteriorMatchingVAE`
training / conditional-likelihood path; everything numeric runs in libpmvae.so
(hand-written CUDA behind the C ABI of include/pmvae.h).  No CPU fallback.
"""
from . import _lib  # noqa: F401  (raises if the CUDA library is missing)
from .config import pm_vae_config  # noqa: F401
from .vae import PosteriorMatchingVAE, get_distribution, get_network  # noqa: F401
from .masking import BernoulliMaskGenerator, MNISTMaskGenerator, get_mask_generator  # noqa: F401
from .train import HostFeeder, Trainer, get_beta_schedule, cyclical_annealing_schedule  # noqa: F401
from .evaluate import eval_fn, nrmse_score  # noqa: F401
from .distributions import AutoregressiveGMM, Bernoulli  # noqa: F401
from .conv_vae import ConvPosteriorMatchingVAE  # noqa: F401
from .lookahead import LookaheadBlock, LookaheadPosterior  # noqa: F401
