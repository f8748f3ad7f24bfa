"""Times the bsds forward (training mode) and backward through the host mirror, CUDA events, 131072 rows."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from posterior_matching_b200 import PosteriorMatchingVAE, pm_vae_config
B = 131072
m = PosteriorMatchingVAE.from_config(pm_vae_config("bsds").model, precision="bf16"); m.init(3)
for hn in ("posterior_dist/linear", "partial_posterior_dist/linear"): m.params[hn]["w"].mul_(0.1)
m.mark_params_changed()
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(B, 63, device="cuda", generator=g); b = (torch.rand(B, 63, device="cuda", generator=g) < 0.5).float()
eps = torch.randn(B, 64, device="cuda", generator=g)
c = torch.full((B,), 1.0 / B, device="cuda")
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print("PMVAE_FUSED_DEBUG", os.environ.get("PMVAE_FUSED_DEBUG", "0"), "fwd ms", round(t(lambda: m(x, b, eps=eps)), 3),
      "bwd ms", round(t(lambda: m.backward(-c, 0.3 * c, -c)), 3), flush=True)
