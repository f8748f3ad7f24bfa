"""Per-Linear steady-state cost of the fused forward: encoder nets with R = 2 and R = 8 blocks, same rows;
(t8 - t2) / 12 is the time one more hidden Linear adds."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from posterior_matching_b200 import PosteriorMatchingVAE, pm_vae_config
M = int(os.environ.get("ROWS", 131072))
ts = {}
for R in (2, 8):
    mc = pm_vae_config("power").model.to_dict()
    mc["encoder_net_config"] = dict(mc["encoder_net_config"], residual_blocks=R)
    m = PosteriorMatchingVAE.from_config(mc, precision="bf16"); m.init(0)
    x = torch.randn(M, m.num_features, device="cuda")
    out = torch.empty(M, 152, device="cuda")
    f = lambda: m.net_apply(0, x, None, out)
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): f()
    e1.record(); torch.cuda.synchronize()
    ts[R] = e0.elapsed_time(e1) / 20 * 1e3
per = (ts[8] - ts[2]) / 12
tiles = (M + 127) // 128
per_tile_cycles = per * 1e-6 * 1.965e9 / ((tiles + 147) // 148)
print(f"CTA2={os.environ.get('PMVAE_FUSED_CTA2','0')} DEBUG={os.environ.get('PMVAE_FUSED_DEBUG','0')}: R2 {ts[2]:.1f} us, R8 {ts[8]:.1f} us, "
      f"per Linear {per:.2f} us = {per_tile_cycles:.0f} cycles per tile-Linear (MMA floor 2048), "
      f"{2*M*65536/per/1e6:.0f} TFLOP/s marginal", flush=True)
