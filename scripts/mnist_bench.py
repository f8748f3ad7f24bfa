"""Throughput of the MNIST-config train step (float32 first cut, batch 256 as in configs/pm_vae_mnist.py)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from posterior_matching_b200 import PosteriorMatchingVAE, MNISTMaskGenerator, pm_vae_config
B = int(os.environ.get("ROWS", 256))
PREC = os.environ.get("PREC", "bf16")
m = PosteriorMatchingVAE.from_config(pm_vae_config("mnist").model.to_dict(), precision=PREC); m.init(0)
m.params["posterior_dist/linear"]["w"].mul_(0.1)
x = (torch.rand(B, 28, 28, 1, device="cuda") < 0.13).float()
gen = MNISTMaskGenerator(seed=1)
for i in range(3):
    m.train_step(x, gen((B, 28, 28, 1)), rng=(0, i))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 10
e0.record()
for i in range(n):
    out = m.train_step(x, gen((B, 28, 28, 1)), rng=(1, i))
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(f"mnist ({PREC} conv GEMMs) train step: {ms:.2f} ms for {B} rows = {B / ms * 1e3:.0f} samples/s, {609.5e6 * B / ms / 1e9:.1f} TFLOP/s algorithmic; loss {out['loss']:.2f}")
