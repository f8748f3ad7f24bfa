"""world_size-2 gloo test of the data-parallel contract on CPU: row-sharded
RNG slices reproduce the global stream, and the all-reduced sum of 1/B_global-scaled
shard gradients equals the full-batch gradient."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    from oracle import model as M, prng
    from tests.util import spec_of, conditioned_params
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    spec = spec_of("power")
    p = conditioned_params(spec)
    Bg = 16
    B = Bg // world
    rng = np.random.default_rng(0)
    x = torch.tensor(rng.standard_normal((Bg, spec.D)))
    # every rank draws only ITS rows of the global mask / eps draw
    bits = prng.random_bits(prng.PRNGKey(1), Bg * spec.D).reshape(Bg, spec.D)
    b_full = torch.tensor(((bits >> 31) == 0).astype(np.float64))
    eps_full = torch.tensor(prng.normal(prng.PRNGKey(2), (Bg, spec.d)).astype(np.float64))
    sl = slice(rank * B, (rank + 1) * B)
    _, aux, g = M.loss_and_grads(p, spec, x[sl], b_full[sl], eps_full[sl], 0.4)
    flat = torch.cat([g[n][k].reshape(-1) for n in g for k in g[n]]) / world   # mean over B -> 1/B_global scaling
    dist.all_reduce(flat)
    _, _, gf = M.loss_and_grads(p, spec, x, b_full, eps_full, 0.4)
    full = torch.cat([gf[n][k].reshape(-1) for n in gf for k in gf[n]])
    q.put((rank, float((flat - full).abs().max()), float(full.abs().max())))
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_equals_full_batch():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, err, scale in res:
        assert err < 1e-12 * max(scale, 1.0), (rank, err)
