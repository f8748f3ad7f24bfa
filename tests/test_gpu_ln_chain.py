"""The fused LayerNorm training chains (bsds: hk.LayerNorm after every Linear, networks.py:117-129): training-mode forward
with saved xhat / 1/sigma and the LayerNorm-aware backward chain, including the wide masked first Linear of the partial
encoder (2 D = 126 > 64 columns), with several 128-row tiles per CTA (PMVAE_FUSED_MAXGRID caps the persistent grids, read
once per process: the checks run in a child process) against the float64 oracle and against the unfused per-Linear path."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CODE = r'''
import sys, dataclasses, numpy as np, torch
sys.path.insert(0, %r)
from oracle import model as M
from tests.util import conditioned_params, make_inputs, spec_of, rel_l2
from posterior_matching_b200 import PosteriorMatchingVAE, pm_vae_config
name, B = "bsds", 1100                      # 9 tiles: 4-5 per CTA under PMVAE_FUSED_MAXGRID=2, ragged last tile
spec = dataclasses.replace(spec_of(name), stop_grad=True)
p = conditioned_params(spec)
x, b, eps = make_inputs(spec, B, seed=6)
m = PosteriorMatchingVAE.from_config(pm_vae_config(name).model, precision="bf16"); m.load_params(p)
out = m(x.float().cuda(), b.float().cuda(), eps=eps.float().cuda())
g = torch.full((B,), 1.0 / B, device="cuda")
m.backward(-g, 0.37 * g, -g)
torch.cuda.synchronize()
res = dict(rec=out["reconstruction_ll"].cpu().numpy(), kl=out["kl"].cpu().numpy(), match=out["matching_ll"].cpu().numpy(),
           grads=m.grad_arena.cpu().numpy())
if sys.argv[2] == "oracle":
    loss, aux, grads = M.loss_and_grads(p, spec, x, b, eps, 0.37)
    want = M.forward(p, spec, x, b, eps)
    worst = 0.0
    for n in grads:
        for k in grads[n]:
            w = grads[n][k].numpy()
            if np.linalg.norm(w) > 0:
                worst = max(worst, rel_l2(m.grads[n][k].cpu().numpy(), w))
    res["worst_grad"] = np.array(worst)
    for k, kk in (("rec", "reconstruction_ll"), ("kl", "kl"), ("match", "matching_ll")):
        res["want_" + k] = want[kk].detach().numpy()
np.savez(sys.argv[1], **res)
''' % ROOT


def _run(env_extra, mode):
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "o.npz")
        env = dict(os.environ, **env_extra)
        r = subprocess.run([sys.executable, "-c", CODE, path, mode], env=env, capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stderr[-3000:]
        return dict(np.load(path))


def test_fused_ln_chains_match_oracle_with_several_tiles_per_cta():
    res = _run({"PMVAE_FUSED_MAXGRID": "2"}, "oracle")
    for k in ("rec", "kl", "match"):
        got, want = res[k], res["want_" + k]
        assert np.isfinite(got).all()
        assert abs(got.mean() - want.mean()) <= 1e-3 * abs(want.mean()), (k, got.mean(), want.mean())
    assert float(res["worst_grad"]) < 2e-1, float(res["worst_grad"])      # GRAD_TOL["bf16-bsds"] of tests/test_gpu_model.py


def test_fused_ln_chains_agree_with_the_unfused_path():
    fused = _run({"PMVAE_FUSED_MAXGRID": "3"}, "plain")
    plain = _run({"PMVAE_FUSED": "0"}, "plain")
    for k in ("rec", "kl", "match"):
        assert abs(fused[k].mean() - plain[k].mean()) <= 1e-3 * abs(plain[k].mean()), k
    g1, g0 = fused["grads"], plain["grads"]
    assert np.isfinite(g1).all()
    assert np.linalg.norm(g1 - g0) / np.linalg.norm(g0) < 1e-1
