"""Worker of tests/test_gpu_dist.py (launched by torch.distributed.run, one rank per GPU): runs a few fused training steps
on this rank's rows of a global batch and lets rank 0 save the resulting parameters and metrics."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.util import conditioned_params, spec_of  # noqa: E402
from posterior_matching_b200 import PosteriorMatchingVAE, Trainer, pm_vae_config  # noqa: E402


def main():
    out_path, name, steps, graph = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4] == "graph"
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    cfg = pm_vae_config(name)
    spec = spec_of(name)
    m = PosteriorMatchingVAE.from_config(cfg.model, precision=os.environ.get("PMVAE_TEST_PRECISION", "fp32"))
    m.load_params(conditioned_params(spec))
    init = m.arena.clone()
    tr = Trainer(cfg, seed=4, precision=m.precision, model=m)
    tr.step = 24000
    Bg = 512                                  # global batch; every rank takes a contiguous slice
    Bl = Bg // world
    metrics = []
    for i in range(steps):
        xg = torch.randn(Bg, spec.D, device="cuda", generator=torch.Generator(device="cuda").manual_seed(300 + i))
        x = xg[rank * Bl:(rank + 1) * Bl].contiguous()
        tr.train_step_fused(x, graph=graph)
        metrics.append([tr.metrics()[k] for k in ("reconstruction_ll", "kl", "matching_ll", "loss")])
    torch.cuda.synchronize()
    if rank == 0:
        np.savez(out_path, params=m.arena.cpu().numpy(), init=init.cpu().numpy(), metrics=np.array(metrics))
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        os._exit(0)      # see bench.py Dist.close: communicator teardown blocks while captured NCCL graphs are alive


if __name__ == "__main__":
    main()
