"""the Posterior Matching term of PosteriorMatchingVADE (vade.py:246-265) and its training step
(train_pm_vade.py:38-61) on the device against the float64 oracle (oracle/model_vade.py); DiagonalGaussian objects."""
import numpy as np
import pytest
import torch

from oracle import model_vade as MV
from tests.util import rel_err, rel_l2

pytestmark = pytest.mark.gpu

CONFIG = {"encoder_net": "ConvEncoder", "decoder_net": "ConvDecoder", "decoder_dist": "Bernoulli", "latent_dim": 10,
          "num_components": 10, "partial_posterior_dist": "AutoregressiveGMM",
          "partial_posterior_dist_config": {"num_components": 10, "residual_blocks": 2, "hidden_units": 256},
          "encoder_net_config": {"conv_layers": [(32, 5, 1), (32, 5, 2), (64, 5, 1), (64, 5, 2), (128, 7, 1)]},
          "decoder_net_config": {"conv_layers": [(64, 7, 1), (64, 5, 2), (32, 5, 1), (32, 5, 2), (32, 5, 1), (1, 5, 1)]}}


def _inputs(B, seed=0):
    rng = np.random.default_rng(seed)
    x = torch.tensor((rng.random((B, 28, 28, 1)) < 0.13).astype(np.float64))
    b = torch.tensor((rng.random((B, 28, 28, 1)) < 0.5).astype(np.float64))
    eps = torch.tensor(rng.standard_normal((B, MV.LATENT)))
    return x, b, eps


def test_posterior_matching_ll_and_partial_gradients():
    from posterior_matching_b200.pm_vade import PosteriorMatchingVADE
    m = PosteriorMatchingVADE.from_config(CONFIG)
    assert [(n, tuple(s), nb) for n, s, nb in m.frozen_leaves + m.train_leaves] == \
        [(n, tuple(s), nb) for n, s, nb in MV.leaf_shapes()[:len(m.frozen_leaves) + len(m.train_leaves)]]
    p = MV.init_params()
    m.load_params(p)
    B = 5
    x, b, eps = _inputs(B)
    loss, want_ll, grads = MV.loss_and_grads(p, x, b, eps)
    ll = m.posterior_matching_ll(x.float().cuda(), b.float().cuda(), eps=eps.float().cuda())
    g = m.backward(torch.full((B,), -1.0 / B, device="cuda"))
    torch.cuda.synchronize()
    assert rel_err(ll.cpu().numpy(), want_ll.numpy()) < 2e-4
    assert set(g) == set(grads)                      # exactly the `partial_*` modules are trainable
    for n in grads:
        for k in grads[n]:
            w = grads[n][k].numpy()
            assert rel_l2(g[n][k].cpu().numpy().reshape(w.shape), w) < 2e-3, (n, k)
    # one optimizer step moves only the partial modules and lowers the loss on the same batch
    frozen = m.frozen_arena.clone()
    l0 = m.train_step(x.float().cuda(), b.float().cuda(), eps=eps.float().cuda())["loss"]
    for _ in range(5):
        l1 = m.train_step(x.float().cuda(), b.float().cuda(), eps=eps.float().cuda())["loss"]
    assert torch.equal(frozen, m.frozen_arena) and l1 < l0
    assert abs(l0 - float(loss)) < 2e-4 * abs(float(loss))


def test_diagonal_gaussian_kernels_match_torch():
    from posterior_matching_b200 import _lib
    B, d = 37, 10
    torch.manual_seed(0)
    par = torch.randn(B, 2 * d, dtype=torch.float64)
    eps = torch.randn(B, d, dtype=torch.float64)
    scale = torch.nn.functional.softplus(par[:, d:]) + 1e-5
    ref = torch.distributions.Independent(torch.distributions.Normal(par[:, :d], scale), 1)
    z_want = par[:, :d] + scale * eps
    S = torch.cuda.current_stream().cuda_stream
    pc, ec = par.float().cuda(), eps.float().cuda()
    z = torch.empty(B, d, device="cuda"); lp = torch.empty(B, device="cuda"); ent = torch.empty(B, device="cuda")
    _lib.check(_lib.lib.pmvae_diag_sample(pc.data_ptr(), ec.data_ptr(), B, d, z.data_ptr(), S), "diag_sample")
    _lib.check(_lib.lib.pmvae_diag_log_prob(pc.data_ptr(), z.data_ptr(), B, d, lp.data_ptr(), ent.data_ptr(), S), "diag_lp")
    torch.cuda.synchronize()
    assert rel_err(z.cpu().numpy(), z_want.numpy()) < 1e-5
    assert rel_err(lp.cpu().numpy(), ref.log_prob(z_want).numpy()) < 1e-4
    assert rel_err(ent.cpu().numpy(), ref.entropy().numpy()) < 1e-5
