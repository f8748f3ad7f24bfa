"""Times one net + head (pmvae_net_apply) alone: decoder / encoder / partial encoder of a config."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
from posterior_matching_b200 import _lib, PosteriorMatchingVAE, pm_vae_config
name = os.environ.get("CFG", "power")
M = int(os.environ.get("ROWS", 131072))
m = PosteriorMatchingVAE.from_config(pm_vae_config(name).model, precision="bf16")
m.init(0)
S = torch.cuda.current_stream().cuda_stream
D, d = m.num_features, m.latent_dim
P = d + d * (d + 1) // 2
x = torch.randn(M, D, device="cuda"); b = (torch.rand(M, D, device="cuda") < 0.5).float(); z = torch.randn(M, d, device="cuda")
ws = m._workspace(M, 0) if hasattr(m, "_workspace") else None
for which, (inp, msk, cols, macs) in {0: (x, None, P, D * 256 + 4 * 65536 + 256 * P), 1: (z, None, D, d * 256 + 4 * 65536 + 256 * D),
                                      2: (x, b, P, 2 * D * 256 + 4 * 65536 + 256 * P)}.items():
    out = torch.empty(M, cols, device="cuda")
    f = lambda: m.net_apply(which, inp, msk, out)
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"DEBUG={os.environ.get('PMVAE_FUSED_DEBUG','0')} net {which} rows {M}: {ms*1e3:.1f} us  {2*macs*M/ms/1e9:.1f} TFLOP/s", flush=True)
