"""`LookaheadPosterior` (lookahead.py:122-227; train_lookahead_posterior.py:47-70) against the float64
oracle: the subsampled features bit for bit, the model's one-step latent samples, the objective, its gradients with
respect to the lookahead modules, one optimiser step, and `expected_info_gains`."""
import numpy as np
import pytest
import torch

from oracle import model as M, model_lookahead as OL, prng as oprng
from tests.util import conditioned_params, make_inputs, rel_err, rel_l2, spec_of

pytestmark = pytest.mark.gpu


def _build(name, K, S, precision="fp32"):
    from posterior_matching_b200 import pm_vae_config
    from posterior_matching_b200.lookahead import LookaheadPosterior
    spec = spec_of(name)
    p = conditioned_params(spec)
    cfg = pm_vae_config(name).model
    look = LookaheadPosterior.from_config({"num_features": spec.D, "lookahead_subsample": S, "model_samples": K}, cfg,
                                          precision=precision)
    look.pm_vae.load_params(p)
    lp = OL.init_params(spec, look.net.residual_blocks, look.net.hidden_units)
    look.load_params(lp)
    return spec, p, lp, look


@pytest.mark.parametrize("name,B,K,S", [("gas", 9, 5, 4), ("power", 6, 7, 6), ("bsds", 3, 3, 16)])
def test_lookahead_objective_and_gradients_match_oracle(name, B, K, S):
    spec, p, lp, look = _build(name, K, S)
    R, ln = look.net.residual_blocks, look.net.layer_norm
    x, b, _ = make_inputs(spec, B, seed=21)
    b[0] = 1.0                                      # a row with nothing left to acquire: no valid lookahead, ll = 0
    key = oprng.PRNGKey(123)
    loss_o, ll_o, g_o, (inds_o, valid_o, z1_o) = OL.loss_and_grads(p, lp, spec, R, ln, x, b, key, K, S)
    rng = tuple(int(v) for v in key)
    xc, bc = x.float().cuda(), b.float().cuda()
    inds, valid, z1 = look.model_one_step_samples(xc, bc, rng)
    torch.cuda.synchronize()
    assert inds.cpu().tolist() == list(inds_o)       # jax.random.choice(replace=False): integer work, exact
    assert torch.equal(valid.cpu().double(), valid_o)
    assert z1.shape == (K, B, S, spec.d)
    assert rel_l2(z1.cpu().numpy(), z1_o.numpy()) < 2e-4
    ll = look(xc, bc, rng=rng)
    assert ll.shape == (1, B) and float(ll[0, 0]) == 0.0
    assert rel_err(ll[0].detach().cpu().numpy(), ll_o.numpy()) < 1e-3
    loss, grads = look.loss_and_grads(xc, bc, rng=rng)
    torch.cuda.synchronize()
    assert abs(float(loss) - float(loss_o)) < 1e-3 * max(1.0, abs(float(loss_o)))
    for n in g_o:
        for k in ("w", "b"):
            assert rel_l2(grads[n][k].cpu().numpy(), g_o[n][k].numpy()) < 2e-3, (n, k)
    # only the selected features' LookaheadBlock columns receive gradient
    d = spec.d
    hb = grads[OL.HEAD]["b"].view(spec.D, 2 * d).abs().sum(-1).cpu()
    assert set(torch.nonzero(hb).flatten().tolist()) <= set(inds_o.tolist())


def test_lookahead_train_step_and_info_gains():
    spec, p, lp, look = _build("gas", 6, 4)
    R, ln = look.net.residual_blocks, look.net.layer_norm
    x, b, _ = make_inputs(spec, 16, seed=5)
    xc, bc = x.float().cuda(), b.float().cuda()
    key = oprng.PRNGKey(9)
    rng = tuple(int(v) for v in key)
    # one update = scale_by_adam -> schedule -> -1 on the lookahead leaves (train_lookahead_posterior.py:53-61)
    _, _, g_o, _ = OL.loss_and_grads(p, lp, spec, R, ln, x, b, key, 6, 4)
    want = {n: {k: t.clone() for k, t in leaf.items()} for n, leaf in lp.items()}
    M.adamw_update(want, g_o, M.zeros_like_params(lp), M.zeros_like_params(lp), 0, 1e-3, 0.0)
    before = look.arena.clone()
    pm_before = look.pm_vae.arena.clone()
    out = look.train_step(xc, bc, rng=rng)
    torch.cuda.synchronize()
    assert np.isfinite(out["loss"]) and look.step == 1
    assert torch.equal(pm_before, look.pm_vae.arena)         # the PM-VAE is frozen (trainable_predicate)
    assert not torch.equal(before, look.arena)
    for n in want:
        for k in ("w", "b"):
            upd_o = (want[n][k] - lp[n][k]).numpy()
            upd = (look.params[n][k].cpu().double() - lp[n][k]).numpy()
            # Adam's first step is lr * sign(g) wherever |g| >> eps: compare where the oracle gradient is not ~0
            sel = np.abs(g_o[n][k].numpy()) > 1e-6
            assert np.allclose(upd[sel], upd_o[sel], atol=2e-5), (n, k)
    # losses go down over a few steps on a fixed batch
    l0 = out["loss"]
    for _ in range(25):
        out = look.train_step(xc, bc, rng=rng)
    assert out["loss"] < l0
    # expected_info_gains: one instance, -inf on observed features
    look.load_params(lp)
    xi, bi = x[3], b[3].clone()
    bi[:2] = 1.0; bi[2:] = 0.0
    want_g = OL.expected_info_gains(p, lp, spec, R, ln, xi, bi)
    got_g = look.expected_info_gains(xi.float().cuda(), bi.float().cuda()).cpu().double()
    assert got_g.shape == (spec.D,)
    assert torch.isinf(got_g[:2]).all() and (got_g[:2] < 0).all()
    assert torch.allclose(got_g[2:], want_g[2:], rtol=1e-4, atol=1e-4)


MNIST16_ENC = [(32, 3, 1), (32, 3, 2), (64, 3, 2), (64, 1, 1)]                 # configs/pm_vae_mnist16.py:24-29
MNIST16_DEC = [(64, 8, 1), (64, 5, 2), (32, 5, 1), (32, 5, 1), (1, 3, 1)]      # :31-38


def test_lookahead_over_the_convolutional_mnist16_model():
    """configs/lookahead_mnist16.py: a frozen ConvEncoder / ConvDecoder PM-VAE with TriLGaussian posteriors (latent 10,
    16 x 16 x 1 images) and a ConvEncoder lookahead net (the PM-VAE encoder's config by default, lookahead.py:107-113)."""
    from posterior_matching_b200.lookahead import LookaheadPosterior
    from posterior_matching_b200.conv_vae import ConvPosteriorMatchingVAE
    K, S, B = 3, 6, 4
    spec = OL.ConvLookSpec(16, 1, 10, MNIST16_ENC, MNIST16_DEC, MNIST16_ENC)
    p, lp = OL.conv_init(spec)
    cfg = {"latent_dim": 10, "encoder_net": "ConvEncoder", "decoder_net": "ConvDecoder", "posterior_dist": "TriLGaussian",
           "decoder_dist": "Bernoulli", "encoder_net_config": {"conv_layers": MNIST16_ENC},
           "decoder_net_config": {"conv_layers": MNIST16_DEC}}
    look = LookaheadPosterior.from_config({"num_features": 256, "lookahead_subsample": S, "model_samples": K}, cfg,
                                          image_size=16)
    assert isinstance(look.pm_vae, ConvPosteriorMatchingVAE) and look.pm_vae.argmm is None
    look.pm_vae.load_params(p)
    look.load_params(lp)
    g = torch.Generator().manual_seed(3)
    x = (torch.rand(B, 16, 16, 1, generator=g, dtype=torch.float64) < 0.3).double()
    b = (torch.rand(B, 16, 16, 1, generator=g, dtype=torch.float64) < 0.15).double()
    b[0] = 1.0
    key = oprng.PRNGKey(31)
    rng = tuple(int(v) for v in key)
    loss_o, ll_o, g_o, (inds_o, valid_o, z1_o) = OL.conv_loss_and_grads(p, lp, spec, x, b, key, K, S)
    xc, bc = x.float().cuda(), b.float().cuda()
    inds, valid, z1 = look.model_one_step_samples(xc, bc, rng)
    torch.cuda.synchronize()
    assert inds.cpu().tolist() == list(inds_o)
    assert torch.equal(valid.cpu().double(), valid_o)
    assert rel_l2(z1.cpu().numpy(), z1_o.numpy()) < 1e-3
    ll = look(xc, bc, rng=rng)
    assert ll.shape == (1, B) and float(ll[0, 0]) == 0.0
    assert rel_err(ll[0].detach().cpu().numpy(), ll_o.numpy()) < 2e-3
    loss, grads = look.loss_and_grads(xc, bc, rng=rng)
    torch.cuda.synchronize()
    assert abs(float(loss) - float(loss_o)) < 2e-3 * max(1.0, abs(float(loss_o)))
    for n in g_o:
        for k in ("w", "b"):
            assert rel_l2(grads[n][k].cpu().numpy(), g_o[n][k].numpy()) < 1e-2, (n, k)
    out = look.train_step(xc, bc, rng=rng)
    assert np.isfinite(out["loss"]) and look.step == 1
    look.load_params(lp)
    bi = b[1].clone()
    want_g = OL.conv_expected_info_gains(p, lp, spec, x[1], bi)
    got_g = look.expected_info_gains(x[1].float().cuda(), bi.float().cuda()).cpu().double()
    obs = bi.reshape(-1) == 1
    assert got_g.shape == (256,) and torch.isinf(got_g[obs]).all()
    assert torch.allclose(got_g[~obs], want_g[~obs], rtol=1e-3, atol=1e-3)
