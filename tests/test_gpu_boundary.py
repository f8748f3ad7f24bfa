"""The module-interface pieces beyond __call__ / is_log_prob: `impute` -> [K, B, D] (vae.py:146-169), `.prior`, and the
distribution objects behind `.encoder` / `.decoder` / `.partial_encoder` (vae.py:47-57; lookahead.py:126-133,219-222),
each against the float64 oracle / torch.distributions; and eval_pm_vae_uci.py's eval_fn body (:82-94) run line for line
against the mirror."""
import math

import numpy as np
import pytest
import torch

from oracle import model as M, prng as oprng
from tests.util import conditioned_params, make_inputs, rel_err, spec_of

pytestmark = pytest.mark.gpu


def _model(name, precision, params):
    from posterior_matching_b200 import PosteriorMatchingVAE, pm_vae_config
    m = PosteriorMatchingVAE.from_config(pm_vae_config(name).model, precision=precision)
    m.load_params(params)
    return m


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("name,B,K", [("gas", 37, 9), ("bsds", 6, 20)])
def test_impute_samples_match_oracle(name, B, K, precision):
    spec = spec_of(name)
    p = conditioned_params(spec)
    x, b, _ = make_inputs(spec, B, seed=12)
    key = oprng.PRNGKey(17)
    eps = torch.tensor(oprng.normal(key, (K, B, spec.d)).astype(np.float64))
    want = M.impute(p, spec, x, b, eps)
    m = _model(name, precision, p)
    got = m.impute(x.float().cuda(), b.float().cuda(), K, key=tuple(int(v) for v in key))
    mean = m.impute_mean(x.float().cuda(), b.float().cuda(), K, key=tuple(int(v) for v in key))
    torch.cuda.synchronize()
    assert got.shape == (K, B, spec.D)
    tol = 1e-4 if precision == "fp32" else 2e-2
    assert rel_err(got.cpu().numpy(), want.numpy()) < tol
    # observed entries are x_o bit for bit; the mean over K is what pmvae_impute_mean returns
    bb = b.bool().unsqueeze(0).expand(K, -1, -1)
    assert torch.equal(got.cpu()[bb], (x * b).float().unsqueeze(0).expand(K, -1, -1)[bb])
    assert rel_err(got.mean(0).cpu().numpy(), mean.cpu().numpy()) < 1e-5


@pytest.mark.parametrize("name,B", [("gas", 33), ("bsds", 5)])
def test_distribution_objects_match_torch_distributions(name, B):
    spec = spec_of(name)
    p = conditioned_params(spec)
    x, b, _ = make_inputs(spec, B, seed=2)
    m = _model(name, "fp32", p)
    xc, bc = x.float().cuda(), b.float().cuda()
    post = m.encoder(xc)
    part = m.partial_encoder(torch.cat([xc * bc, bc], -1))
    torch.cuda.synchronize()
    for dist, net, head, inp in ((post, "encoder_net", "posterior_dist/linear", x),
                                 (part, "partial_encoder_net", "partial_posterior_dist/linear", torch.cat([x * b, b], -1))):
        raw = M.net_head(p, spec, net, head, inp)
        mu, L = raw[:, :spec.d], M.fill_scale_tril(raw[:, spec.d:], spec.d)
        ref = torch.distributions.MultivariateNormal(mu, scale_tril=L)
        assert rel_err(dist.mean().cpu().numpy(), mu.numpy()) < 1e-4
        assert rel_err(dist.entropy().cpu().numpy(), ref.entropy().numpy()) < 1e-4
        z = torch.tensor(np.random.default_rng(3).standard_normal((B, spec.d)))
        assert rel_err(dist.log_prob(z.float().cuda()).cpu().numpy(), ref.log_prob(z).numpy()) < 2e-4
        zk = torch.tensor(np.random.default_rng(4).standard_normal((3, B, spec.d)))
        assert rel_err(dist.log_prob(zk.float().cuda()).cpu().numpy(), ref.log_prob(zk).numpy()) < 2e-4
        std = torch.distributions.MultivariateNormal(torch.zeros(spec.d, dtype=torch.float64),
                                                     scale_tril=torch.eye(spec.d, dtype=torch.float64))
        want_kl = torch.distributions.kl_divergence(ref, std)
        assert rel_err(dist.kl_divergence(m.prior).cpu().numpy(), want_kl.numpy()) < 1e-4
        # .sample(seed=key, sample_shape=K) = mu + L normal(key, [K, B, d])
        key = oprng.PRNGKey(5)
        eps = torch.tensor(oprng.normal(key, (4, B, spec.d)).astype(np.float64))
        want_z = mu.unsqueeze(0) + torch.einsum("bij,kbj->kbi", L, eps)
        got_z = dist.sample(seed=tuple(int(v) for v in key), sample_shape=4)
        assert got_z.shape == (4, B, spec.d)
        assert rel_err(got_z.cpu().numpy(), want_z.numpy()) < 1e-4
        one = dist.sample(seed=tuple(int(v) for v in key))
        eps1 = torch.tensor(oprng.normal(key, (1, B, spec.d)).astype(np.float64))
        assert one.shape == (B, spec.d)
        assert rel_err(one.cpu().numpy(), (mu + torch.einsum("bij,bj->bi", L, eps1[0])).numpy()) < 1e-4
    # decoder object: Normal(loc, exp(log_scale)), elementwise log_prob
    z = torch.tensor(np.random.default_rng(6).standard_normal((B, spec.d)))
    dec = m.decoder(z.float().cuda())
    loc, ls = M.decoder(p, spec, z)
    ref = torch.distributions.Normal(loc, torch.exp(ls))
    assert rel_err(dec.mean().cpu().numpy(), loc.numpy()) < 1e-4
    assert rel_err(dec.log_prob(xc).cpu().numpy(), ref.log_prob(x).numpy()) < 2e-4
    # prior
    zz = torch.tensor(np.random.default_rng(7).standard_normal((2, B, spec.d)))
    want = -0.5 * (zz ** 2).sum(-1) - 0.5 * spec.d * math.log(2 * math.pi)
    assert rel_err(m.prior.log_prob(zz.float().cuda()).cpu().numpy(), want.numpy()) < 1e-5
    assert m.prior.sample(seed=(0, 5), sample_shape=7).shape == (7, spec.d)


def test_reference_eval_fn_body_runs_against_the_mirror():
    """eval_pm_vae_uci.py:82-94, line for line: `model.impute(x, b, num_samples=K)`, mean over axis 0,
    `model.is_log_prob(x, b, num_samples=K)`; the rng of hk.transform's apply is threaded explicitly."""
    name, B, K = "power", 48, 32
    spec = spec_of(name)
    p = conditioned_params(spec)
    x, b, _ = make_inputs(spec, B, seed=21)
    rng = oprng.PRNGKey(91)
    k_imp, k_z, k_zxo = M.eval_keys(rng, spec)
    e = [torch.tensor(oprng.normal(k, (K, B, spec.d)).astype(np.float64)) for k in (k_imp, k_z, k_zxo)]
    want_imp, want_ll = M.eval_fn(p, spec, x, b, *e)
    model = _model(name, "fp32", p)
    batch = {"features": x.float().cuda(), "mask": b.float().cuda()}
    keys = model.eval_keys(tuple(int(v) for v in rng))

    def eval_fn(batch):
        x = batch["features"]
        b = batch["mask"]
        imputed = model.impute(x, b, num_samples=K, key=keys[0])
        imputed = torch.mean(imputed, axis=0)
        _, log_p_xu_given_xo = model.is_log_prob(x, b, num_samples=K, keys=keys[1:])
        return imputed, log_p_xu_given_xo

    im, ll = eval_fn(batch)
    torch.cuda.synchronize()
    assert rel_err(im.cpu().numpy(), want_imp.numpy()) < 1e-4
    assert np.abs(ll.cpu().numpy() - want_ll.numpy()).max() < 1e-3
