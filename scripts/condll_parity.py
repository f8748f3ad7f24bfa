"""Prints the cond-LL / log p(x) / imputation errors of the CUDA evaluators against the float64 oracle at benchmark scale
(what tests/test_gpu_condll_scale.py asserts), plus raw-Haiku-init fp32 parity for d = 16 (no conditioning of the TriL
heads).  Run on the GPU box: python scripts/condll_parity.py > gpurun_out/condll_parity.txt"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import model as M, prng as oprng  # noqa: E402
from tests.util import conditioned_params, make_inputs, oracle_eval_chunked, rel_err, spec_of  # noqa: E402
from posterior_matching_b200 import PosteriorMatchingVAE, eval_fn, pm_vae_config  # noqa: E402

cases = [("gas", 2048, 512), ("power", 2048, 512), ("hepmass", 2048, 512), ("bsds", 2048, 512), ("bsds", 256, 4096),
         ("gas", 32, 64)]
if len(sys.argv) > 1:
    cases = [c for c in cases if c[0] in sys.argv[1:]]
for name, B, K in cases:
    spec = spec_of(name)
    p = conditioned_params(spec)
    x, b, _ = make_inputs(spec, B, seed=31)
    rng = oprng.PRNGKey(91)
    keys = M.eval_keys(rng, spec)
    t0 = time.time()
    want_imp, want_ll, want_lpx = oracle_eval_chunked(p, spec, x, b, keys, K)
    t_or = time.time() - t0
    for precision in ("bf16", "fp32"):
        if precision == "fp32" and B * K > 2048 * 512:
            continue
        m = PosteriorMatchingVAE.from_config(pm_vae_config(name).model, precision=precision)
        m.load_params(p)
        xc, bc = x.float().cuda(), b.float().cuda()
        imp, ll = eval_fn(m, tuple(int(v) for v in rng), xc, bc, K)
        lpx, _ = m.is_log_prob(xc, bc, K, keys=tuple(tuple(int(v) for v in k) for k in keys[1:]))
        torch.cuda.synchronize()
        ll, lpx, imp = ll.cpu().double(), lpx.cpu().double(), imp.cpu().double()
        print(f"{name:8s} B={B:5d} K={K:5d} {precision}: cond-LL mean {float(ll.mean()):+.5f} (oracle {float(want_ll.mean()):+.5f}) "
              f"rel {abs(float(ll.mean() - want_ll.mean())) / abs(float(want_ll.mean())):.2e}  row max abs {float((ll - want_ll).abs().max()):.2e} "
              f"row rms {float((ll - want_ll).pow(2).mean().sqrt()):.2e} bias {float((ll - want_ll).mean()):+.2e} | log p(x) rel "
              f"{abs(float(lpx.mean() - want_lpx.mean())) / abs(float(want_lpx.mean())):.2e} row max {float((lpx - want_lpx).abs().max()):.2e} | "
              f"impute rel {rel_err(imp[b == 0].numpy(), want_imp[b == 0].numpy()):.2e}  (oracle {t_or:.1f}s)", flush=True)

# raw Haiku init (no x0.1 on the TriL heads), d = 16, fp32 path against the float64 oracle
for name in ("gas", "power", "hepmass"):
    spec = spec_of(name)
    p = M.init_params(spec, 3)
    B = 512
    x, b, eps = make_inputs(spec, B, seed=5)
    want = M.forward(p, spec, x, b, eps)
    for precision in ("fp32", "bf16"):
        m = PosteriorMatchingVAE.from_config(pm_vae_config(name).model, precision=precision)
        m.load_params(p)
        got = m(x.float().cuda(), b.float().cuda(), eps=eps.float().cuda())
        torch.cuda.synchronize()
        line = f"raw-init {name} {precision}:"
        for k in ("reconstruction_ll", "kl", "matching_ll"):
            g, w = got[k].cpu().double(), want[k].detach()
            line += f" {k} mean {float(g.mean()):+.4e} vs {float(w.mean()):+.4e} (rel {abs(float(g.mean() - w.mean())) / abs(float(w.mean())):.2e}, row rel max {float(((g - w).abs() / w.abs().clamp_min(1.0)).max()):.2e});"
        print(line, flush=True)
