"""Determinism stress of the bsds (LayerNorm) chains: the forward (training mode), the three nets in evaluation mode and
the evaluator are deterministic, so repeated calls must agree bit for bit; prints which output / rows ever differ."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.util import conditioned_params, make_inputs, spec_of
from posterior_matching_b200 import PosteriorMatchingVAE, pm_vae_config, eval_fn
name = sys.argv[1] if len(sys.argv) > 1 else "bsds"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 30
spec = spec_of(name)
p = conditioned_params(spec)
for B in (130, 300, 256, 2048):
    x, b, eps = (t.float().cuda() for t in make_inputs(spec, B, seed=4))
    m = PosteriorMatchingVAE.from_config(pm_vae_config(name).model, precision="bf16"); m.load_params(p)
    ref = None
    bad = {}
    for it in range(iters):
        out = {k: v.clone() for k, v in m(x, b, eps=eps).items()}
        out["enc"] = m.encoder(x).parameters.clone()
        out["dec"] = m.decoder(eps[:, :spec.d].contiguous()).mean().clone()
        out["part"] = m.partial_encoder(torch.cat([x * b, b], -1)).parameters.clone()
        if B <= 300:
            imp, ll = eval_fn(m, (0, 91), x, b, 64)
            out["eval_ll"] = ll.clone(); out["eval_imp"] = imp.clone()
        torch.cuda.synchronize()
        if ref is None:
            ref = out
            continue
        for k in out:
            if not torch.equal(out[k], ref[k]):
                d = (out[k] - ref[k]).abs()
                rows = (d.reshape(d.shape[0], -1).amax(1) > 0).nonzero().flatten().tolist()
                bad.setdefault(k, []).append((it, float(d.max()), rows[:6], len(rows)))
    print(f"{name} B={B}: " + ("deterministic over %d iterations" % iters if not bad else "MISMATCH " + str({k: v[:3] for k, v in bad.items()})), flush=True)
