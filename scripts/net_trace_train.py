"""Per-phase clock stamps of block 0 of the training-mode (activation-saving) fused forward launches."""
import os, sys
os.environ["PMVAE_FUSED_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from posterior_matching_b200 import PosteriorMatchingVAE, pm_vae_config
m = PosteriorMatchingVAE.from_config(pm_vae_config("power").model, precision="bf16"); m.init(0)
M = 131072
x = torch.randn(M, m.num_features, device="cuda")
b = (torch.rand(M, m.num_features, device="cuda") < 0.5).float()
m(x, b, rng=(0, 1))
torch.cuda.synchronize()
