"""Pins oracle/model.py: TFP closed forms vs torch.distributions (float64), the
fill_triangular doc example, schedules, the reference's own masking.py outputs, and
the committed golden fixture."""
import math
import os

import numpy as np
import torch

from oracle import model as M
from posterior_matching_b200.config import pm_vae_config


def _spec(name):
    return M.spec_from_config(pm_vae_config(name).model.to_dict())


def _cond_params(spec, seed=3):
    p = M.init_params(spec, seed)
    for hn in ("posterior_dist/linear", "partial_posterior_dist/linear"):
        p[hn]["w"] *= 0.1
    return p


def test_fill_triangular_doc_example():
    v = torch.arange(1.0, 7.0, dtype=torch.float64)
    idx = M.fill_triangular_index(3)
    L = torch.zeros(3, 3, dtype=torch.float64)
    for i in range(3):
        for j in range(i + 1):
            L[i, j] = v[idx[i, j]]
    assert L.tolist() == [[4, 0, 0], [6, 5, 0], [3, 2, 1]]


def test_tril_closed_forms_vs_torch_distributions():
    torch.manual_seed(0)
    for d in (16, 64):
        B = 5
        m = d * (d + 1) // 2
        v = torch.randn(B, m, dtype=torch.float64) * 0.3
        mu = torch.randn(B, d, dtype=torch.float64)
        L = M.fill_scale_tril(v, d)
        assert torch.equal(L, torch.tril(L)) and (torch.diagonal(L, dim1=-2, dim2=-1) > 0).all()
        dist = torch.distributions.MultivariateNormal(mu, scale_tril=L)
        prior = torch.distributions.MultivariateNormal(torch.zeros(d, dtype=torch.float64),
                                                       torch.eye(d, dtype=torch.float64))
        z = torch.randn(B, d, dtype=torch.float64)
        assert torch.allclose(M.tril_log_prob(z, mu, L), dist.log_prob(z), atol=1e-9)
        assert torch.allclose(M.tril_kl_std_normal(mu, L), torch.distributions.kl_divergence(dist, prior), atol=1e-9)
        assert torch.allclose(M.std_normal_log_prob(z), prior.log_prob(z), atol=1e-10)
        # own-sample identity used by the eval kernel (SURVEY a-S)
        eps = torch.randn(B, d, dtype=torch.float64)
        zz = mu + (L @ eps.unsqueeze(-1)).squeeze(-1)
        ident = -0.5 * (eps ** 2).sum(-1) - torch.log(torch.diagonal(L, dim1=-2, dim2=-1)).sum(-1) - 0.5 * d * M.LOG2PI
        assert torch.allclose(M.tril_log_prob(zz, mu, L), ident, atol=1e-8)


def test_param_counts_match_survey_table():
    want = {"gas": 880697, "power": 878647, "hepmass": 894022, "bsds": 3157504}
    for k, n in want.items():
        assert M.n_params(_spec(k)) == n


def test_layer_norm_and_residual_mlp_shapes():
    x = torch.randn(7, 256, dtype=torch.float64)
    y = M.layer_norm(x)
    assert torch.allclose(y, torch.nn.functional.layer_norm(x, (256,), eps=1e-5), atol=1e-10)
    spec = _spec("bsds")
    p = _cond_params(spec)
    h = M.residual_mlp(p, "encoder_net", torch.randn(4, spec.D, dtype=torch.float64), spec.R_enc, True)
    assert h.shape == (4, 256) and (h >= 0).all()


def test_beta_and_lr_schedules():
    cyc = M.beta_schedule(pm_vae_config("gas").beta.to_dict())
    assert cyc(0) == 0 and cyc(999) == 0 and cyc(1000) == 0.0
    assert abs(cyc(1000 + 12500) - 0.5) < 1e-12 and cyc(1000 + 25000) == 1.0 and cyc(1000 + 40000) == 1.0
    assert abs(cyc(1000 + 50000 + 12500) - 0.5) < 1e-12  # the cycle restarts
    mono = M.beta_schedule(pm_vae_config("bsds").beta.to_dict())
    assert mono(0) == 0 and mono(30000) == 0 and abs(mono(130000) - 0.5) < 1e-12 and mono(10 ** 6) == 1.0
    assert M.beta_schedule({})(5) == 1.0
    lr = M.lr_schedule(1e-3, 0.9, 5000)
    assert lr(0) == 1e-3 and abs(lr(5000) - 9e-4) < 1e-15


def test_handwritten_backward_formulas_vs_autograd():
    """the formulas the CUDA latent-backward kernel implements."""
    torch.manual_seed(1)
    d, B = 16, 6
    m = d * (d + 1) // 2
    raw = (torch.randn(B, d + m, dtype=torch.float64) * 0.3).requires_grad_(True)
    rawp = (torch.randn(B, d + m, dtype=torch.float64) * 0.3).requires_grad_(True)
    eps = torch.randn(B, d, dtype=torch.float64)
    mu, L = raw[:, :d], M.fill_scale_tril(raw[:, d:], d)
    mup, Lp = rawp[:, :d], M.fill_scale_tril(rawp[:, d:], d)
    z = mu + (L @ eps.unsqueeze(-1)).squeeze(-1)
    kl = M.tril_kl_std_normal(mu, L)
    match = M.tril_log_prob(z.detach(), mup, Lp)
    gz_up = torch.randn(B, d, dtype=torch.float64)      # stands for the decoder's dX
    kw, mw = 0.7, -1.3
    total = (gz_up * z).sum() + kw * kl.sum() + mw * match.sum()
    g_raw, g_rawp = torch.autograd.grad(total, [raw, rawp])
    # hand formulas
    Ld, Lpd = L.detach(), Lp.detach()
    dmu = gz_up + kw * mu.detach()
    dL = torch.tril(gz_up.unsqueeze(-1) * eps.unsqueeze(-2)) + kw * (Ld - torch.diag_embed(1 / torch.diagonal(Ld, dim1=-2, dim2=-1)))
    r = torch.linalg.solve_triangular(Lpd, (z.detach() - mup.detach()).unsqueeze(-1), upper=False).squeeze(-1)
    g = torch.linalg.solve_triangular(Lpd.transpose(-1, -2), r.unsqueeze(-1), upper=True).squeeze(-1)
    dmup = mw * g
    dLp = mw * (torch.tril(g.unsqueeze(-1) * r.unsqueeze(-2)) - torch.diag_embed(1 / torch.diagonal(Lpd, dim1=-2, dim2=-1)))
    idx = M.fill_triangular_index(d)

    def scatter(dLm, rawv):
        out = torch.zeros(B, m, dtype=torch.float64)
        for i in range(d):
            for j in range(i + 1):
                gij = dLm[:, i, j]
                if i == j:
                    gij = gij * torch.sigmoid(rawv[:, d + idx[i, j]].detach())
                out[:, idx[i, j]] += gij
        return out
    assert torch.allclose(torch.cat([dmu, scatter(dL, raw)], 1), g_raw, atol=1e-10)
    assert torch.allclose(torch.cat([dmup, scatter(dLp, rawp)], 1), g_rawp, atol=1e-10)


def test_data_parallel_shards_sum_to_full_batch_grads():
    """grads of G equal shards, each pre-scaled by 1/B_global, sum to the
    full-batch gradient."""
    spec = _spec("power")
    p = _cond_params(spec)
    rng = np.random.default_rng(0)
    B, G = 32, 4
    x = torch.tensor(rng.standard_normal((B, spec.D)))
    b = torch.tensor((rng.random((B, spec.D)) < 0.5).astype(np.float64))
    eps = torch.tensor(rng.standard_normal((B, spec.d)))
    _, _, full = M.loss_and_grads(p, spec, x, b, eps, 0.3)
    acc = M.zeros_like_params(p)
    for g in range(G):
        s = slice(g * B // G, (g + 1) * B // G)
        _, _, part = M.loss_and_grads(p, spec, x[s], b[s], eps[s], 0.3)
        for n in acc:
            for k in acc[n]:
                acc[n][k] += part[n][k] / G
    for n in acc:
        for k in acc[n]:
            assert torch.allclose(acc[n][k], full[n][k], atol=1e-12), (n, k)


def test_adamw_decay_mask_and_first_step():
    spec = _spec("gas")
    p = _cond_params(spec)
    p0 = {n: {k: t.clone() for k, t in d.items()} for n, d in p.items()}
    g = {n: {k: torch.ones_like(t) for k, t in d.items()} for n, d in p.items()}
    m, v = M.zeros_like_params(p), M.zeros_like_params(p)
    M.adamw_update(p, g, m, v, count=0, lr=1e-3, wd=1e-5)
    # first Adam step with g=1: u = 1/(1+1e-8); biases (ndim 1) get no decay, weights and log_scale do
    u = 1 / (1 + 1e-8)
    n = "encoder_net/linear"
    assert torch.allclose(p[n]["b"], p0[n]["b"] - 1e-3 * u)
    assert torch.allclose(p[n]["w"], p0[n]["w"] - 1e-3 * (u + 1e-5 * p0[n]["w"]))
    assert torch.allclose(p["decoder_dist"]["log_scale"], p0["decoder_dist"]["log_scale"] - 1e-3 * u)


def test_reference_bernoulli_mask_is_random_sample_threshold(golden_dir):
    """masking.py:84-91 [verified against the live reference, fixture made by
    tests/golden/make_golden.py]: binomial(1, p) == (random_sample > 1-p)... for p=.5
    the reference equals (RandomState(s).random_sample(shape) > 0.5)."""
    g = np.load(os.path.join(golden_dir, "masks_reference.npz"))
    for seed in (0, 1, 7):
        want = g[f"bernoulli_seed{seed}"]
        assert want.dtype == np.float32 and set(np.unique(want)) <= {0.0, 1.0}
        got = (np.random.RandomState(seed).random_sample((64, 8)) > 0.5).astype(np.float32)
        assert np.array_equal(got, want)
    assert abs(g["bernoulli_p03_seed5"].mean() - 0.3) < 0.02


def test_model_golden_fixture(golden_dir):
    from oracle import prng
    g = np.load(os.path.join(golden_dir, "model_golden.npz"))
    for name, B, K in (("gas", 16, 8), ("bsds", 8, 4)):
        spec = _spec(name)
        p = _cond_params(spec)
        x = torch.tensor(g[f"{name}_x"])
        b = torch.tensor(g[f"{name}_b"])
        eps = torch.tensor(prng.normal(prng.PRNGKey(2), (B, spec.d)).astype(np.float64))
        o = M.forward(p, spec, x, b, eps)
        for k in ("reconstruction_ll", "kl", "matching_ll"):
            assert np.allclose(o[k].numpy(), g[f"{name}_{k}"], rtol=1e-9, atol=1e-9)
        ek = [torch.tensor(prng.normal(prng.PRNGKey(10 + i), (K, B, spec.d)).astype(np.float64)) for i in range(3)]
        imp, ll = M.eval_fn(p, spec, x, b, *ek)
        assert np.allclose(imp.numpy(), g[f"{name}_impute_mean"], rtol=1e-9, atol=1e-9)
        assert np.allclose(ll.numpy(), g[f"{name}_log_p_xu_given_xo"], rtol=1e-9, atol=1e-9)


def test_is_log_prob_converges_to_exact_for_linear_gaussian_sanity():
    """IS estimate of log p(x) is finite, and log p(x_u|x_o) <= ~0-ish sanity:
    with all features observed log p(x_u|x_o) == 0 exactly when z == z_xo draws agree."""
    spec = _spec("gas")
    p = _cond_params(spec)
    rng = np.random.default_rng(1)
    B, K = 4, 64
    x = torch.tensor(rng.standard_normal((B, spec.D)))
    b = torch.ones(B, spec.D, dtype=torch.float64)
    e1 = torch.tensor(rng.standard_normal((K, B, spec.d)))
    lpx, cond = M.is_log_prob(p, spec, x, b, e1, e1)
    assert torch.isfinite(lpx).all() and torch.isfinite(cond).all()
    assert math.isfinite(float(cond.mean()))


def test_shifted_one_pass_layernorm_statistics_are_exact():
    """The row statistics of the fused LayerNorm epilogue (csrc/fused_mlp.cu::epi_ln): two halves of 128 columns each sum
    S = sum(v - c), Q = sum((v - c)^2) around their own first value c and are combined as
    mean = (S0 + S1 + 128 (c0 + c1)) / 256, sum (v - mean)^2 = sum_h Q_h - 2 (mean - c_h) S_h + 128 (mean - c_h)^2.
    In float32 this must agree with the two-pass variance also when |mean| >> std."""
    import numpy as np
    rng = np.random.default_rng(0)
    for mean, std in ((0.0, 1.0), (300.0, 0.5), (-4000.0, 2.0), (1e-3, 1e-4)):
        v = (mean + std * rng.standard_normal((64, 256))).astype(np.float32)
        halves = [v[:, :128], v[:, 128:]]
        c = [h[:, :1] for h in halves]
        S = [np.sum(h - ch, axis=1, dtype=np.float32) for h, ch in zip(halves, c)]
        Q = [np.sum((h - ch) ** 2, axis=1, dtype=np.float32) for h, ch in zip(halves, c)]
        mu = (S[0] + S[1] + np.float32(128) * (c[0][:, 0] + c[1][:, 0])) * np.float32(1 / 256)
        ss = np.zeros_like(mu)
        for h in range(2):
            dm = mu - c[h][:, 0]
            ss += Q[h] - 2 * dm * S[h] + np.float32(128) * dm * dm
        var = np.maximum(ss, 0) * np.float32(1 / 256)
        v64 = v.astype(np.float64)
        assert np.allclose(mu, v64.mean(axis=1), rtol=1e-6, atol=1e-6 * max(1.0, abs(mean)))
        assert np.allclose(var, v64.var(axis=1), rtol=2e-4), (mean, std)
