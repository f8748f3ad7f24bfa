"""CPU restatement (PyTorch float64) of the PM-VAE of configs/pm_vae_mnist16.py: ConvEncoder / ConvDecoder (networks.py:9-72),
TriLGaussian posterior AND partial posterior (the config names no partial_posterior_dist, so vae.py:97-105 falls back to
posterior_dist), Bernoulli decoder; `PosteriorMatchingVAE.__call__` (vae.py:120-144), loss_fn (train_pm_vae.py:58-72 with
beta = 1, no stop_gradient), `impute` (vae.py:146-169) and `is_log_prob` (vae.py:171-226) given the normal draws.
TEST INFRASTRUCTURE ONLY.  Parameters and layer lists: oracle/model_lookahead.py (`ConvLookSpec`, `conv_init`)."""
from __future__ import annotations

import math

import torch

from . import conv as OC
from . import dists_mnist as DM
from . import model as OM
from .model_lookahead import ConvLookSpec, _convs, _tril_of, conv_partial


def posterior(p, spec: ConvLookSpec, x):
    h = OC.conv_encoder(_convs(p, "encoder_net", "conv2_d", len(spec.enc_layers)), x, spec.enc_layers)
    return _tril_of(p, "posterior_dist/linear", h.reshape(x.shape[0], -1), spec.d)


def decode(p, spec: ConvLookSpec, z):
    return OC.conv_decoder(_convs(p, "decoder_net", "conv2_d_transpose", len(spec.dec_layers)), z, spec.dec_layers)


def forward(p, spec: ConvLookSpec, x, b, eps):
    B = x.shape[0]
    mu, L = posterior(p, spec, x)
    z = mu + (L @ eps.unsqueeze(-1)).squeeze(-1)
    rec = DM.bernoulli_log_prob(decode(p, spec, z), x).reshape(B, -1).sum(-1)
    kl = OM.tril_kl_std_normal(mu, L)
    mu_p, L_p = conv_partial(p, spec, torch.cat([x * b, b], -1))
    match = OM.tril_log_prob(z, mu_p, L_p)                       # no stop_gradient in this config
    return {"reconstruction_ll": rec, "kl": kl, "matching_ll": match, "z": z}


def loss_and_grads(p, spec: ConvLookSpec, x, b, eps, coef: float = 1.0):
    q = {n: {k: t.detach().clone().requires_grad_(True) for k, t in leaf.items()} for n, leaf in p.items()}
    out = forward(q, spec, x, b, eps)
    loss = -(out["reconstruction_ll"] - out["kl"]).mean() + coef * (-out["matching_ll"].mean())
    loss.backward()
    return loss.detach(), {k: v.detach() for k, v in out.items()}, {n: {k: t.grad for k, t in leaf.items()} for n, leaf in q.items()}


def impute(p, spec: ConvLookSpec, x_o, b, eps):
    """eps [K,B,d] drives z ~ q(z | x_o) -> [K, B, H, W, C]."""
    K, B, d = eps.shape
    x_o = x_o * b
    mu_p, L_p = conv_partial(p, spec, torch.cat([x_o, b], -1))
    z = mu_p.unsqueeze(0) + torch.einsum("bij,kbj->kbi", L_p, eps)
    mean = torch.sigmoid(decode(p, spec, z.reshape(K * B, d))).reshape(K, *x_o.shape)
    return torch.where(b.unsqueeze(0) != 0, x_o.unsqueeze(0), mean)


def is_log_prob(p, spec: ConvLookSpec, x, b, eps_z, eps_zxo):
    K, B, d = eps_z.shape
    mu, L = posterior(p, spec, x)
    mu_p, L_p = conv_partial(p, spec, torch.cat([x * b, b], -1))
    z = mu.unsqueeze(0) + torch.einsum("bij,kbj->kbi", L, eps_z)
    z_xo = mu_p.unsqueeze(0) + torch.einsum("bij,kbj->kbi", L_p, eps_zxo)

    def dec_ll(zz, weight):
        ll = DM.bernoulli_log_prob(decode(p, spec, zz.reshape(K * B, d)).reshape(K, *x.shape), x.unsqueeze(0))
        if weight is not None:
            ll = ll * weight.unsqueeze(0)
        return ll.reshape(K, B, -1).sum(-1)

    lk = math.log(K)
    log_p_x = torch.logsumexp(dec_ll(z, None) + OM.std_normal_log_prob(z) - OM.tril_log_prob(z, mu.unsqueeze(0), L.unsqueeze(0)), 0) - lk
    log_p_xo = torch.logsumexp(dec_ll(z_xo, b) + OM.std_normal_log_prob(z_xo)
                               - OM.tril_log_prob(z_xo, mu_p.unsqueeze(0), L_p.unsqueeze(0)), 0) - lk
    return log_p_x, log_p_x - log_p_xo
