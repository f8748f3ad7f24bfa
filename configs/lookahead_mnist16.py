"""Same keys/values as the reference's configs/lookahead_mnist16.py (built from the shared table)."""
from posterior_matching_b200.config import lookahead_mnist16_config


def get_config():
    return lookahead_mnist16_config()
