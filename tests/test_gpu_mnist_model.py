"""The MNIST-config PM-VAE composed from libpmvae operators (posterior_matching_b200/conv_vae.py) against the float64
oracle (oracle/model_mnist.py): per-row terms, loss and every gradient leaf; masks from the MNIST mask generator."""
import numpy as np
import pytest
import torch

from oracle import model_mnist as MM
from tests.util import rel_err, rel_l2

pytestmark = pytest.mark.gpu


def _inputs(B, seed=0):
    from posterior_matching_b200 import MNISTMaskGenerator
    rng = np.random.default_rng(seed)
    x = torch.tensor((rng.random((B, 28, 28, 1)) < 0.13).astype(np.float64))
    b = MNISTMaskGenerator(seed=seed + 1)((B, 28, 28, 1)).cpu().double()
    eps = torch.tensor(rng.standard_normal((B, MM.LATENT)))
    return x, b, eps


def test_mnist_model_forward_loss_and_gradients():
    from posterior_matching_b200 import pm_vae_config
    from posterior_matching_b200.conv_vae import ConvPosteriorMatchingVAE
    cfg = pm_vae_config("mnist")
    m = ConvPosteriorMatchingVAE.from_config(cfg.model.to_dict())
    assert [(n, tuple(s), nb) for n, s, nb in m.leaves] == [(n, tuple(s), nb) for n, s, nb in MM.leaf_shapes()[:len(m.leaves)]]
    p = MM.init_params()
    m.load_params(p)
    B = 6
    x, b, eps = _inputs(B)
    loss, out, grads = MM.loss_and_grads(p, x, b, eps)
    got = m(x.float().cuda(), b.float().cuda(), eps=eps.float().cuda())
    for k in ("reconstruction_ll", "kl", "matching_ll"):
        assert rel_err(got[k].cpu().numpy(), out[k].numpy()) < 2e-4, k
    ones = torch.full((B,), 1.0 / B, device="cuda")
    g = m.backward(-ones, ones, -ones)
    torch.cuda.synchronize()
    got_loss = float(-(got["reconstruction_ll"] - got["kl"]).mean() - got["matching_ll"].mean())
    assert abs(got_loss - float(loss)) < 2e-5 * abs(float(loss))
    for n in grads:
        for k in grads[n]:
            w = grads[n][k].numpy()
            gg = g[n][k].cpu().numpy().reshape(w.shape)
            e = rel_l2(gg, w) if np.linalg.norm(w) > 0 else float(np.abs(gg).max())
            assert e < 2e-3, (n, k, e)


def test_mnist_model_bf16_convolutions_within_tolerance():
    """precision="bf16": convolution GEMM operands rounded to bf16 (fp32 accumulation on the tensor cores), everything
    else float32.  north_star tolerance on the per-batch loss: 1e-3 relative; gradient leaves are held to the figure
    operand rounding plus leaky-relu sign flips of near-zero pre-activations give (as for the UCI bf16 path)."""
    from posterior_matching_b200 import pm_vae_config
    from posterior_matching_b200.conv_vae import ConvPosteriorMatchingVAE
    m = ConvPosteriorMatchingVAE.from_config(pm_vae_config("mnist").model.to_dict(), precision="bf16")
    p = MM.init_params()
    m.load_params(p)
    B = 6
    x, b, eps = _inputs(B, seed=5)
    loss, out, grads = MM.loss_and_grads(p, x, b, eps)
    got = m(x.float().cuda(), b.float().cuda(), eps=eps.float().cuda())
    ones = torch.full((B,), 1.0 / B, device="cuda")
    g = m.backward(-ones, ones, -ones)
    torch.cuda.synchronize()
    got_loss = float(-(got["reconstruction_ll"] - got["kl"]).mean() - got["matching_ll"].mean())
    assert abs(got_loss - float(loss)) < 1e-3 * abs(float(loss)), (got_loss, float(loss))
    for k in ("reconstruction_ll", "kl", "matching_ll"):
        assert rel_l2(got[k].cpu().numpy(), out[k].numpy()) < 2e-3, k
    worst = 0.0
    for n in grads:
        for k in grads[n]:
            w = grads[n][k].numpy()
            if np.linalg.norm(w) == 0:
                continue
            worst = max(worst, rel_l2(g[n][k].cpu().numpy().reshape(w.shape), w))
    assert worst < 0.1, worst


def test_mnist_train_step_reduces_the_loss():
    from posterior_matching_b200 import pm_vae_config
    from posterior_matching_b200.conv_vae import ConvPosteriorMatchingVAE
    m = ConvPosteriorMatchingVAE.from_config(pm_vae_config("mnist").model.to_dict())
    m.load_params(MM.init_params())
    x, b, eps = (t.float().cuda() for t in _inputs(16, seed=3))
    first = m.train_step(x, b, eps=eps)
    for _ in range(12):
        last = m.train_step(x, b, eps=eps)
    assert np.isfinite(last["loss"]) and last["loss"] < first["loss"]


def test_mnist_train_step_data_parallel_hooks_match_the_full_batch():
    """Two ranks' worth of rows through `global_rows` + `grad_sync` (the other shard's gradients are added by the hook,
    as an all-reduce would) give the same parameter update as one full-batch step."""
    from posterior_matching_b200 import pm_vae_config
    from posterior_matching_b200.conv_vae import ConvPosteriorMatchingVAE
    cfg = pm_vae_config("mnist").model.to_dict()
    p = MM.init_params()
    full, rank0, rank1 = (ConvPosteriorMatchingVAE.from_config(cfg) for _ in range(3))
    for m in (full, rank0, rank1):
        m.load_params(p)
    B = 8
    x, b, eps = (t.float().cuda() for t in _inputs(B, seed=11))
    full.train_step(x, b, eps=eps)
    # rank 1's shard: gradients only (cotangents scaled by 1 / global rows), no update
    h = B // 2
    rank1(x[h:], b[h:], eps=eps[h:])
    ones = torch.full((h,), 1.0 / B, device="cuda")
    rank1.backward(-ones, ones, -ones)
    other = [rank1.grad_arena.clone(), rank1.argmm.grad_arena.clone()]

    def sync(arenas):
        for a, o in zip(arenas, other):
            a.add_(o)

    rank0.train_step(x[:h], b[:h], eps=eps[:h], grad_sync=sync, global_rows=B)
    torch.cuda.synchronize()
    for a, c in ((full.arena, rank0.arena), (full.argmm.arena, rank0.argmm.arena)):
        # Adam's first step moves every weight by ~lr * sign(g): compare the updates, not the weights
        assert rel_l2((c - a).cpu().numpy() + 1.0, np.ones(a.numel())) < 1e-5
        assert float((a - c).abs().max()) < 2e-4


def test_mnist_model_impute_and_is_log_prob():
    """`impute` / `is_log_prob` of the MNIST config (vae.py:146-226) -- AR-GMM samples from the device
    sampler (its noise contract is tested in test_gpu_mnist_dists.py), everything downstream against the float64 oracle
    fed the same samples."""
    from oracle import prng as oprng
    from posterior_matching_b200 import pm_vae_config
    from posterior_matching_b200.conv_vae import ConvPosteriorMatchingVAE
    m = ConvPosteriorMatchingVAE.from_config(pm_vae_config("mnist").model.to_dict())
    p = MM.init_params()
    m.load_params(p)
    B, K = 3, 6
    x, b, _ = _inputs(B, seed=8)
    xc, bc = x.float().cuda(), b.float().cuda()
    k_z, k_zxo = (5, 6), (7, 8)
    ctx = m._context(xc * bc, bc)
    z_xo = m.argmm.sample(ctx, K, key=k_zxo).cpu().double()
    # impute: observed pixels kept, unobserved = Bernoulli mean of the decoded samples
    imp = m.impute(xc, bc, K, key=k_zxo)
    want_imp = MM.impute(p, x, b, z_xo)
    torch.cuda.synchronize()
    assert imp.shape == (K, B, 28, 28, 1)
    assert rel_err(imp.cpu().numpy(), want_imp.numpy()) < 2e-4
    # is_log_prob
    eps_z = torch.tensor(oprng.normal(np.array(k_z, dtype=np.uint32), (K, B, MM.LATENT)).astype(np.float64))
    want_lpx, want_cond = MM.is_log_prob(p, x, b, eps_z, z_xo)
    lpx, cond = m.is_log_prob(xc, bc, K, keys=(k_z, k_zxo))
    torch.cuda.synchronize()
    assert rel_err(lpx.cpu().numpy(), want_lpx.numpy()) < 2e-4
    assert np.abs(cond.cpu().numpy() - want_cond.numpy()).max() < 2e-3 * max(1.0, np.abs(want_cond.numpy()).max())
