"""Same keys/values as the reference's configs/pm_vae_mnist16.py (built from the shared table)."""
from posterior_matching_b200.config import pm_vae_config


def get_config():
    return pm_vae_config("mnist16")
