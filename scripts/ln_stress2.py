"""Isolates the flaky chain: each net of the bsds model alone (pmvae_net_apply), evaluation and training mode, many
repetitions on the same input; prints the repetitions / row blocks whose output differs from the majority."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.util import conditioned_params, make_inputs, spec_of
from posterior_matching_b200 import PosteriorMatchingVAE, pm_vae_config, _lib
name = sys.argv[1] if len(sys.argv) > 1 else "bsds"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 100
print("env:", {k: v for k, v in os.environ.items() if k.startswith("PMVAE")}, flush=True)
spec = spec_of(name)
p = conditioned_params(spec)
for B in (2048, 18944 * 2):
    x, b, eps = (t.float().cuda() for t in make_inputs(spec, B, seed=4))
    z = eps[:, :spec.d].contiguous()
    m = PosteriorMatchingVAE.from_config(pm_vae_config(name).model, precision="bf16"); m.load_params(p)
    for which, inp, msk in ((0, x, None), (2, x, b)):
        for save in (_lib.NET_SAVE,):
            outs = []
            for it in range(iters):
                outs.append(m.net_apply(which | save, inp, msk).clone())
            torch.cuda.synchronize()
            # majority = the most common output per 128-row block
            bad = []
            for blk in range((B + 127) // 128):
                rows = slice(blk * 128, min(B, (blk + 1) * 128))
                ref = outs[-1][rows]
                n_ref = sum(torch.equal(o[rows], ref) for o in outs)
                if n_ref < iters:
                    diffs = []
                    for i, o in enumerate(outs):
                        if not torch.equal(o[rows], ref):
                            d = (o[rows] - ref).abs()
                            cols = (d.amax(0) > 0).nonzero().flatten()
                            rws = (d.amax(1) > 0).nonzero().flatten()
                            diffs.append((i, round(float(d.max()), 2), f"cols {sorted(set((cols // 32 * 32).tolist()))} ({cols.numel()})",
                                          f"rows {int(rws.min())}..{int(rws.max())} ({rws.numel()})"))
                    bad.append((blk, diffs[:4], len(diffs)))
            print(f"{name} B={B} net {which} {'train' if save else 'eval '}: " + ("ok" if not bad else f"FLAKY {bad[:4]}"), flush=True)
