"""is_log_prob (K = 512) of one config, a few calls: for an ncu launch list of the evaluator alone."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from posterior_matching_b200 import PosteriorMatchingVAE, pm_vae_config
name = os.environ.get("CFG", "bsds")
m = PosteriorMatchingVAE.from_config(pm_vae_config(name).model, precision="bf16"); m.init(0)
for hn in ("posterior_dist/linear", "partial_posterior_dist/linear"):
    m.params[hn]["w"].mul_(0.1)
m.mark_params_changed()
B = int(os.environ.get("ROWS", 2048))
x = torch.randn(B, m.num_features, device="cuda"); b = (torch.rand(B, m.num_features, device="cuda") < 0.5).float()
for i in range(3):
    out = m.is_log_prob(x, b, num_samples=512, rng=(0, i))
torch.cuda.synchronize()
