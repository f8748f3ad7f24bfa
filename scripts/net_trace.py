"""Prints the per-phase clock stamps of block 0 of one fused net launch (PMVAE_FUSED_TRACE=1)."""
import os, sys
os.environ["PMVAE_FUSED_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from posterior_matching_b200 import PosteriorMatchingVAE, pm_vae_config
m = PosteriorMatchingVAE.from_config(pm_vae_config(os.environ.get("CFG", "power")).model, precision="bf16"); m.init(0)
M = 131072
which = int(os.environ.get("NET", 0))      # 0 encoder, 1 decoder
x = torch.randn(M, m.latent_dim if which == 1 else m.num_features, device="cuda")
m.net_apply(which, x)
torch.cuda.synchronize()
