"""pmvae_net_apply (the module's .encoder / .decoder / .partial_encoder) against the oracle's
networks, fused and unfused tensor paths and the fp32 path; and fused == unfused on the
quantities the backward consumes."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import model as M
from tests.util import conditioned_params, make_inputs, rel_err, spec_of

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("name,B", [("gas", 300), ("power", 129), ("hepmass", 7), ("bsds", 40)])
def test_net_apply_matches_oracle_networks(name, B, precision):
    from posterior_matching_b200 import PosteriorMatchingVAE, pm_vae_config
    spec = spec_of(name)
    p = conditioned_params(spec)
    x, b, eps = make_inputs(spec, B, seed=3)
    z = torch.tensor(np.random.default_rng(5).standard_normal((B, spec.d)))
    m = PosteriorMatchingVAE.from_config(pm_vae_config(name).model, precision=precision)
    m.load_params(p)
    want_enc = M.net_head(p, spec, "encoder_net", "posterior_dist/linear", x)
    want_dec = M.net_head(p, spec, "decoder_net", "decoder_dist/linear", z)
    want_part = M.net_head(p, spec, "partial_encoder_net", "partial_posterior_dist/linear", torch.cat([x * b, b], -1))
    got_enc = m.encoder(x.float().cuda()).parameters
    got_dec = m.decoder(z.float().cuda()).mean()
    got_part = m.partial_encoder(torch.cat([x * b, b], -1).float().cuda()).parameters
    torch.cuda.synchronize()
    tol = 1e-4 if precision == "fp32" else 3e-2
    for g, w in ((got_enc, want_enc), (got_dec, want_dec), (got_part, want_part)):
        assert g.shape == w.shape and torch.isfinite(g).all()
        assert rel_err(g.cpu().numpy(), w.detach().numpy()) < tol


def test_fused_and_unfused_tensor_paths_agree():
    """PMVAE_FUSED is read once per process: run the unfused path in a child process and
    compare losses and gradients with the fused path on the same inputs."""
    code = r'''
import sys, torch, numpy as np
sys.path.insert(0, %r)
from tests.util import conditioned_params, make_inputs, spec_of
from posterior_matching_b200 import PosteriorMatchingVAE, pm_vae_config
spec = spec_of("hepmass"); p = conditioned_params(spec); B = 700
x, b, eps = (t.float().cuda() for t in make_inputs(spec, B, seed=8))
m = PosteriorMatchingVAE.from_config(pm_vae_config("hepmass").model, precision="bf16"); m.load_params(p)
out = m(x, b, eps=eps)
g = torch.full((B,), 1.0 / B, device="cuda")
m.backward(g, -0.5 * g, g)
torch.cuda.synchronize()
np.savez(sys.argv[1], rec=out["reconstruction_ll"].cpu().numpy(), kl=out["kl"].cpu().numpy(),
         match=out["matching_ll"].cpu().numpy(), grads=m.grad_arena.cpu().numpy())
''' % ROOT
    import tempfile
    res = {}
    with tempfile.TemporaryDirectory() as td:
        for fused in ("1", "0"):
            path = os.path.join(td, f"o{fused}.npz")
            env = dict(os.environ, PMVAE_FUSED=fused)
            r = subprocess.run([sys.executable, "-c", code, path], env=env, capture_output=True, text=True, timeout=300)
            assert r.returncode == 0, r.stderr[-2000:]
            res[fused] = dict(np.load(path))
    for k in ("rec", "kl", "match"):
        assert rel_err(res["1"][k], res["0"][k]) < 2e-2, k
        assert abs(res["1"][k].mean() - res["0"][k].mean()) < 1e-3 * abs(res["0"][k].mean())
    g1, g0 = res["1"]["grads"], res["0"]["grads"]
    assert np.linalg.norm(g1 - g0) / np.linalg.norm(g0) < 5e-2
