"""configs/pm_vae_mnist16.py (convolutional networks, TriLGaussian posterior AND partial posterior, Bernoulli decoder; the
model `LookaheadPosterior` is trained over in configs/lookahead_mnist16.py): per-row terms, every gradient leaf, one train
step, `impute`, `is_log_prob` and the VJP operator `pmvae_tril_log_prob_backward` against the float64 oracle."""
import numpy as np
import pytest
import torch

from oracle import model as M, model_lookahead as OL, model_mnist16 as O16, prng as oprng
from tests.util import rel_err, rel_l2

pytestmark = pytest.mark.gpu

ENC = [(32, 3, 1), (32, 3, 2), (64, 3, 2), (64, 1, 1)]                 # configs/pm_vae_mnist16.py:24-29
DEC = [(64, 8, 1), (64, 5, 2), (32, 5, 1), (32, 5, 1), (1, 3, 1)]      # :31-38
CFG = {"latent_dim": 10, "encoder_net": "ConvEncoder", "decoder_net": "ConvDecoder", "posterior_dist": "TriLGaussian",
       "decoder_dist": "Bernoulli", "encoder_net_config": {"conv_layers": ENC}, "decoder_net_config": {"conv_layers": DEC}}


def _setup(B, seed=2):
    from posterior_matching_b200 import PosteriorMatchingVAE
    spec = OL.ConvLookSpec(16, 1, 10, ENC, DEC, ENC)
    p, _ = OL.conv_init(spec)
    m = PosteriorMatchingVAE.from_config(CFG, image_size=16)
    m.load_params(p)
    g = torch.Generator().manual_seed(seed)
    x = (torch.rand(B, 16, 16, 1, generator=g, dtype=torch.float64) < 0.3).double()
    b = (torch.rand(B, 16, 16, 1, generator=g, dtype=torch.float64) < 0.2).double()
    return spec, p, m, x, b


def test_tril_log_prob_backward_matches_autograd():
    from posterior_matching_b200 import _lib
    for d, B in ((10, 7), (16, 5), (64, 3)):
        P = d + d * (d + 1) // 2
        g = torch.Generator().manual_seed(d)
        par = (0.3 * torch.randn(B, P, generator=g, dtype=torch.float64)).requires_grad_(True)
        z = torch.randn(B, d, generator=g, dtype=torch.float64, requires_grad=True)
        cot = torch.randn(B, generator=g, dtype=torch.float64)
        lp = M.tril_log_prob(z, par[:, :d], M.fill_scale_tril(par[:, d:], d))
        (lp * cot).sum().backward()
        pc, zc, gc = par.detach().float().cuda(), z.detach().float().cuda(), cot.float().cuda()
        dpar, dz = torch.empty_like(pc), torch.empty_like(zc)
        ws = torch.empty(B * (P + d + 1), dtype=torch.float32, device="cuda")
        S = torch.cuda.current_stream().cuda_stream
        _lib.check(_lib.lib.pmvae_tril_log_prob_backward(pc.data_ptr(), zc.data_ptr(), gc.data_ptr(), B, d, dpar.data_ptr(),
                                                         dz.data_ptr(), ws.data_ptr(), ws.numel() * 4, S))
        torch.cuda.synchronize()
        assert rel_l2(dpar.cpu().numpy(), par.grad.numpy()) < 2e-4, d
        assert rel_l2(dz.cpu().numpy(), z.grad.numpy()) < 2e-4, d
        with pytest.raises(_lib.PmvaeError):
            _lib.check(_lib.lib.pmvae_tril_log_prob_backward(pc.data_ptr(), zc.data_ptr(), gc.data_ptr(), B, d, dpar.data_ptr(),
                                                             dz.data_ptr(), ws.data_ptr(), 16, S))


def test_mnist16_terms_gradients_and_train_step_match_oracle():
    B = 6
    spec, p, m, x, b = _setup(B)
    eps = torch.randn(B, 10, generator=torch.Generator().manual_seed(9), dtype=torch.float64)
    loss_o, out_o, g_o = O16.loss_and_grads(p, spec, x, b, eps)
    out = m(x.float().cuda(), b.float().cuda(), eps=eps.float().cuda())
    ones = torch.full((B,), 1.0 / B, device="cuda")
    grads = m.backward(-ones, ones, -ones)
    torch.cuda.synchronize()
    for k in ("reconstruction_ll", "kl", "matching_ll"):
        assert rel_err(out[k].cpu().numpy(), out_o[k].numpy()) < 2e-4, k
    for n in g_o:
        for k in ("w", "b"):
            assert rel_l2(grads[n][k].cpu().numpy(), g_o[n][k].numpy()) < 3e-3, (n, k)
    before = m.arena.clone()
    met = m.train_step(x.float().cuda(), b.float().cuda(), eps=eps.float().cuda())
    assert abs(met["loss"] - float(loss_o)) < 1e-3 * max(1.0, abs(float(loss_o))) and m.step == 1
    assert not torch.equal(before, m.arena)


def test_mnist16_impute_and_is_log_prob_match_oracle():
    B, K = 4, 5
    spec, p, m, x, b = _setup(B, seed=4)
    k_imp, k_z, k_zxo = oprng.PRNGKey(1), oprng.PRNGKey(2), oprng.PRNGKey(3)
    as64 = lambda key: torch.tensor(oprng.normal(key, (K, B, 10)).astype(np.float64))
    want_imp = O16.impute(p, spec, x, b, as64(k_imp))
    want_lp, want_cond = O16.is_log_prob(p, spec, x, b, as64(k_z), as64(k_zxo))
    xc, bc = x.float().cuda(), b.float().cuda()
    tk = lambda key: tuple(int(v) for v in key)
    imp = m.impute(xc, bc, K, key=tk(k_imp))
    lp, cond = m.is_log_prob(xc, bc, K, keys=(tk(k_z), tk(k_zxo)))
    torch.cuda.synchronize()
    assert imp.shape == (K, B, 16, 16, 1)
    assert rel_err(imp.cpu().numpy(), want_imp.numpy()) < 5e-4
    assert rel_err(lp.cpu().numpy(), want_lp.numpy()) < 5e-4
    assert rel_err(cond.cpu().numpy(), want_cond.numpy()) < 5e-4
