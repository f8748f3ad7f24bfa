"""bsds (LayerNorm) diagnostics: per-term forward errors vs the float64 oracle at several batch sizes, fused vs unfused
(PMVAE_FUSED read once per process -> run twice), gradient errors."""
import dataclasses, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import model as M
from tests.util import conditioned_params, make_inputs, spec_of, rel_l2, rel_err
from posterior_matching_b200 import PosteriorMatchingVAE, pm_vae_config
print("PMVAE_FUSED =", os.environ.get("PMVAE_FUSED", "1"), flush=True)
name = "bsds"
spec = dataclasses.replace(spec_of(name), stop_grad=True)
p = conditioned_params(spec)
for B in (64, 130, 256, 300, 1100):
    x, b, eps = make_inputs(spec, B, seed=4)
    want = M.forward(p, spec, x, b, eps)
    m = PosteriorMatchingVAE.from_config(pm_vae_config(name).model, precision="bf16"); m.load_params(p)
    got = m(x.float().cuda(), b.float().cuda(), eps=eps.float().cuda())
    torch.cuda.synchronize()
    line = f"B={B:5d} fwd:"
    for k in ("reconstruction_ll", "kl", "matching_ll"):
        g, w = got[k].cpu().double(), want[k].detach()
        bad = int(((g - w).abs() > 2e-2 * w.abs().max()).sum())
        line += f" {k[:5]} mean rel {abs(float(g.mean()-w.mean()))/abs(float(w.mean())):.2e} rowmax {float((g-w).abs().max()/w.abs().max()):.2e} bad {bad} first_bad {int(((g - w).abs() > 2e-2 * w.abs().max()).nonzero()[0]) if bad else -1};"
    print(line, flush=True)
    # net outputs of the three nets (eval mode)
    pe = m.partial_encoder(torch.cat([x * b, b], -1).float().cuda()).parameters.cpu().double()
    we = M.net_head(p, spec, "partial_encoder_net", "partial_posterior_dist/linear", torch.cat([x * b, b], -1))
    rowerr = (pe - we).abs().amax(1) / we.abs().max()
    print(f"        partial_encoder eval: rel err {float(rowerr.max()):.2e}, rows > 3e-2: {int((rowerr > 3e-2).sum())} first {int((rowerr > 3e-2).nonzero()[0]) if (rowerr > 3e-2).any() else -1}", flush=True)
    if B in (64, 300):
        loss, aux, grads = M.loss_and_grads(p, spec, x, b, eps, 0.37)
        g = torch.full((B,), 1.0 / B, device="cuda")
        m.backward(-g, 0.37 * g, -g)
        torch.cuda.synchronize()
        errs = {}
        for n in grads:
            for k in grads[n]:
                w = grads[n][k].numpy()
                if np.linalg.norm(w) > 0:
                    errs[(n, k)] = rel_l2(m.grads[n][k].cpu().numpy(), w)
        worst = max(errs, key=errs.get)
        print(f"        grads: worst {errs[worst]:.3f} at {worst}; per net max: " + ", ".join(
            f"{pre}: {max(v for (n, k), v in errs.items() if n.startswith(pre)):.3f}" for pre in ("encoder", "posterior", "decoder_net", "decoder_dist", "partial_enc", "partial_post")), flush=True)
