"""Regenerates tests/golden/*.  Run in the BUILD container (needs /root/reference).

  python tests/golden/make_golden.py

* masks_reference.npz -- outputs of the LIVE reference `posterior_matching/masking.py`
  (imported from /root/reference with a stub `tensorflow` module: only
  `get_add_mask_fn` touches TF).  Seeded BernoulliMaskGenerator draws pin what
  `masking.py:84-91` produces; a large unseeded-equivalent MNISTMaskGenerator sample
  pins the mixture's category frequencies / geometry (masking.py:235-249) that the
  device contract must reproduce in distribution.
* model_golden.npz -- float64 oracle outputs on small fixed inputs (regression
  fixture for the oracle itself and a GPU-side parity fixture).
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def reference_masks():
    sys.modules.setdefault("tensorflow", types.ModuleType("tensorflow"))
    sys.path.insert(0, "/root/reference")
    from posterior_matching import masking  # the reference itself

    out = {}
    for seed in (0, 1, 7):
        g = masking.BernoulliMaskGenerator(seed=seed)
        out[f"bernoulli_seed{seed}"] = g((64, 8))
    g = masking.BernoulliMaskGenerator(p=0.3, seed=5)
    out["bernoulli_p03_seed5"] = g((256, 21))

    # MNIST mixture: the sub-generators are unseeded in the reference (masking.py:238-246),
    # so seed the global state they fall back to for a reproducible sample.
    np.random.seed(1234)
    gen = masking.MNISTMaskGenerator(seed=11)
    n = 20000
    m = gen((n, 28, 28, 1))[..., 0]
    assert m.shape == (n, 28, 28)
    # classify each row into the 7 categories by geometry
    cats = np.full(n, -1)
    frac = m.mean((1, 2))
    half = {1: (slice(0, 28), slice(0, 14)), 2: (slice(0, 14), slice(0, 28)),
            3: (slice(0, 28), slice(14, 28)), 4: (slice(14, 28), slice(0, 28))}
    rect_area = np.zeros(n)
    for i in range(n):
        z = m[i] == 0
        ys, xs = np.where(z)
        if len(ys) == 0:
            cats[i] = 0
            continue
        y0, y1, x0, x1 = ys.min(), ys.max(), xs.min(), xs.max()
        solid = z[y0:y1 + 1, x0:x1 + 1].all() and z.sum() == (y1 - y0 + 1) * (x1 - x0 + 1)
        if not solid:
            cats[i] = 0
            continue
        h, w = y1 - y0 + 1, x1 - x0 + 1
        rect_area[i] = h * w
        hit = None
        # FixedRectangle(y1, x1, y2, x2) zeroes mask[y1:y2, x1:x2]
        if (y0, x0, h, w) == (0, 0, 28, 14):
            hit = 1
        elif (y0, x0, h, w) == (0, 0, 14, 28):
            hit = 2
        elif (y0, x0, h, w) == (0, 14, 28, 14):
            hit = 3
        elif (y0, x0, h, w) == (14, 0, 14, 28):
            hit = 4
        elif h == 14 and w == 14:
            hit = 5
        else:
            hit = 6
        cats[i] = hit
    out["mnist_cat_freq"] = np.bincount(cats, minlength=7) / n
    out["mnist_rect_area_min"] = np.array(rect_area[cats == 6].min())
    out["mnist_rect_area_max"] = np.array(rect_area[cats == 6].max())
    out["mnist_bern_mean"] = np.array(frac[cats == 0].mean())
    sq = np.where(cats == 5)[0]
    tl = np.array([[np.where(m[i] == 0)[0].min(), np.where(m[i] == 0)[1].min()] for i in sq])
    out["mnist_square_tl_max"] = tl.max(0)
    out["mnist_square_tl_min"] = tl.min(0)
    out["mnist_n"] = np.array(n)
    np.savez_compressed(os.path.join(HERE, "masks_reference.npz"), **out)
    print({k: (v if v.size < 10 else v.shape) for k, v in out.items()})


def model_golden():
    import torch
    from oracle import model as M, prng
    from posterior_matching_b200.config import pm_vae_config

    out = {}
    for name, B, K in (("gas", 16, 8), ("bsds", 8, 4)):
        spec = M.spec_from_config(pm_vae_config(name).model.to_dict())
        p = M.init_params(spec, seed=3)
        for hn in ("posterior_dist/linear", "partial_posterior_dist/linear"):
            p[hn]["w"] *= 0.1
        rng = np.random.default_rng(0)
        x = torch.tensor(rng.standard_normal((B, spec.D)).astype(np.float32), dtype=torch.float64)
        b = torch.tensor(prng.bernoulli(prng.PRNGKey(1), 0.5, (B, spec.D)).astype(np.float64))
        eps = torch.tensor(prng.normal(prng.PRNGKey(2), (B, spec.d)).astype(np.float64))
        o = M.forward(p, spec, x, b, eps)
        loss, aux, _ = M.loss_and_grads(p, spec, x, b, eps, 0.5)
        ek = [torch.tensor(prng.normal(prng.PRNGKey(10 + i), (K, B, spec.d)).astype(np.float64)) for i in range(3)]
        imp, ll = M.eval_fn(p, spec, x, b, *ek)
        out[f"{name}_x"] = x.numpy()
        out[f"{name}_b"] = b.numpy()
        for k in ("reconstruction_ll", "kl", "matching_ll"):
            out[f"{name}_{k}"] = o[k].detach().numpy()
        out[f"{name}_loss_beta0.5"] = loss.numpy()
        out[f"{name}_impute_mean"] = imp.detach().numpy()
        out[f"{name}_log_p_xu_given_xo"] = ll.detach().numpy()
    np.savez_compressed(os.path.join(HERE, "model_golden.npz"), **out)
    print("model_golden:", sorted(out))


if __name__ == "__main__":
    reference_masks()
    model_golden()
