"""Pins oracle/dists_mnist.py against torch.distributions (float64)."""
import torch

from oracle import dists_mnist as DM


def test_gmm_log_prob_vs_torch_mixture_same_family():
    torch.manual_seed(0)
    B, d, K = 6, 5, 10
    params = torch.randn(B, d, 3 * K, dtype=torch.float64)
    value = torch.randn(B, d, dtype=torch.float64)
    mix = torch.distributions.Categorical(logits=params[..., :K])
    comp = torch.distributions.Normal(params[..., K:2 * K], torch.nn.functional.softplus(params[..., 2 * K:]) + 1e-5)
    want = torch.distributions.MixtureSameFamily(mix, comp).log_prob(value)
    assert torch.allclose(DM.gmm_log_prob(params, value, K), want, atol=1e-10)


def test_bernoulli_log_prob_vs_torch():
    torch.manual_seed(1)
    logits = torch.randn(7, 11, dtype=torch.float64) * 3
    x = (torch.rand(7, 11, dtype=torch.float64) < 0.3).double()
    want = torch.distributions.Bernoulli(logits=logits).log_prob(x)
    assert torch.allclose(DM.bernoulli_log_prob(logits, x), want, atol=1e-12)
    # float events interpolate linearly between the two outcomes
    xf = torch.rand(7, 11, dtype=torch.float64)
    lo = torch.distributions.Bernoulli(logits=logits).log_prob(torch.zeros_like(xf))
    hi = torch.distributions.Bernoulli(logits=logits).log_prob(torch.ones_like(xf))
    assert torch.allclose(DM.bernoulli_log_prob(logits, xf), xf * hi + (1 - xf) * lo, atol=1e-12)


def test_argmm_is_a_normalised_autoregressive_density():
    """Step i only sees dimensions < i (changing later dimensions cannot change earlier terms), and in one
    dimension the density integrates to one."""
    spec = DM.ArgmmSpec(d=4, n_comp=3, R=1, H=32, C=5)
    p = DM.argmm_init(spec)
    torch.manual_seed(2)
    ctx = torch.randn(3, spec.C, dtype=torch.float64)
    v = torch.randn(3, spec.d, dtype=torch.float64)
    base = DM.argmm_log_prob(p, spec, v, ctx)
    v2 = v.clone(); v2[:, -1] += 1.0        # only the last term may change
    spec1 = DM.ArgmmSpec(d=1, n_comp=3, R=1, H=32, C=5)
    p1 = DM.argmm_init(spec1)
    grid = torch.linspace(-12, 12, 4801, dtype=torch.float64).unsqueeze(-1)
    dens = DM.argmm_log_prob(p1, spec1, grid, ctx[:1].expand(grid.shape[0], -1)).exp()
    assert abs(float(torch.trapezoid(dens, grid.squeeze(-1))) - 1.0) < 1e-6
    # prefix property: recompute with the last dimension dropped from the sum
    def terms(val):
        B, d = val.shape
        out = []
        ar = torch.arange(d, dtype=val.dtype)
        for i in range(d):
            mask = (ar < i).to(val.dtype).expand(B, d)
            h = DM.residual_mlp(p, DM.NET, torch.cat([val * mask, mask, ctx], -1), spec.R, False)
            params = DM.linear(p, DM.HEAD, h).reshape(B, d, 3 * spec.n_comp)
            out.append(DM.gmm_log_prob(params, val, spec.n_comp)[:, i])
        return torch.stack(out, 0)
    t1, t2 = terms(v), terms(v2)
    assert torch.allclose(t1[:-1], t2[:-1], atol=1e-12) and not torch.allclose(t1[-1], t2[-1])
    assert torch.allclose(t1.sum(0), base, atol=1e-12)


def test_argmm_sample_follows_the_mixture_of_the_first_dimension():
    """Dimension 0 of the autoregressive sampler sees only the context: its samples follow the first 1-D mixture
    (analytic mean / variance), and the samples' own log-density is what argmm_log_prob computes (finite, and higher
    on average than that of shuffled samples)."""
    from oracle import prng
    spec = DM.ArgmmSpec(d=3, n_comp=4, R=1, H=32, C=5)
    p = DM.argmm_init(spec)
    torch.manual_seed(3)
    ctx = torch.randn(2, spec.C, dtype=torch.float64)
    n = 4000
    x = DM.argmm_sample(p, spec, ctx, n, prng.PRNGKey(11))
    assert x.shape == (n, 2, 3) and torch.isfinite(x).all()
    from oracle.model import linear, residual_mlp
    inp = torch.cat([torch.zeros(2, 3, dtype=torch.float64), torch.zeros(2, 3, dtype=torch.float64), ctx], -1)
    par = linear(p, DM.HEAD, residual_mlp(p, DM.NET, inp, spec.R, False)).reshape(2, 3, 12)[:, 0]
    w = torch.softmax(par[:, :4], -1)
    mu, sc = par[:, 4:8], torch.nn.functional.softplus(par[:, 8:]) + 1e-5
    mean = (w * mu).sum(-1)
    var = (w * (sc ** 2 + mu ** 2)).sum(-1) - mean ** 2
    se = (var / n).sqrt()
    assert torch.all((x[:, :, 0].mean(0) - mean).abs() < 5 * se)
    assert torch.all((x[:, :, 0].var(0) / var - 1).abs() < 0.15)
    lp = DM.argmm_log_prob(p, spec, x.reshape(n * 2, 3), ctx.repeat(n, 1))
    perm = torch.randperm(n * 2)
    lp_shuffled = DM.argmm_log_prob(p, spec, x.reshape(n * 2, 3)[perm], ctx.repeat(n, 1))
    assert torch.isfinite(lp).all() and lp.mean() >= lp_shuffled.mean() - 1e-9
