"""Data parallelism on real GPUs (skipped with fewer than two): the 2-rank Trainer -- rows sharded by rank, masks and eps
drawn from the global counter stream, bucketed gradient all-reduce (NCCL) overlapped with the remaining backward inside
one captured CUDA graph -- walks the same trajectory as (a) the same two ranks with one serial all-reduce after the
backward and (b) a single rank on the concatenated batch (SURVEY §8e's invariant)."""
import os
import socket
import subprocess
import sys
import tempfile

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "dist_worker.py")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(nproc, out, name, steps, mode, env_extra=None):
    env = dict(os.environ, **(env_extra or {}))
    env.pop("NCCL_DEBUG", None)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), WORKER, out, name, str(steps), mode]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=240, cwd=ROOT)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-3000:])
    return dict(np.load(out))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_two_rank_overlapped_step_equals_serial_and_single_rank(precision):
    steps = 4
    with tempfile.TemporaryDirectory() as td:
        env = {"PMVAE_TEST_PRECISION": precision}
        over = _run(2, os.path.join(td, "over.npz"), "gas", steps, "graph", env)
        eager = _run(2, os.path.join(td, "eager.npz"), "gas", steps, "eager", env)
        serial = _run(2, os.path.join(td, "serial.npz"), "gas", steps, "graph", dict(env, PMVAE_DP_OVERLAP="0"))
        single = _run(1, os.path.join(td, "single.npz"), "gas", steps, "graph", env)
    tol = 2e-5 if precision == "fp32" else 1e-3
    for other in (eager, serial, single):
        for i in range(steps):
            a, c = over["metrics"][i], other["metrics"][i]
            assert np.all(np.abs(a - c) <= tol * (1 + i) * np.maximum(1.0, np.abs(c))), (i, a, c)
    # parameters after the last step: same updates up to the reordering of the weight-gradient atomics, which Adam
    # with fresh moments amplifies (cf. tests/test_gpu_model.py::test_fused_graph_train_step_equals_host_driven_step)
    ptol = 5e-3 if precision == "fp32" else 2.5e-1
    move = np.linalg.norm(single["params"] - single["init"])
    assert move > 0
    for other in (eager, serial, single):
        assert np.linalg.norm(over["params"] - other["params"]) <= ptol * move
