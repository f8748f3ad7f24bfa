#!/usr/bin/env python
"""PM-VAE hot-path benchmark (contract: one JSON line on rank 0).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                  [--config auto|gas|power|hepmass|bsds|mnist] [--batch ROWS_PER_GPU] [--precision auto|fp32|bf16]

A "step" is one training step of train_pm_vae.py on one batch of synthetic rows of the config's shape: device mask
draw (threefry), eps draw, forward, loss, backward, gradient all-reduce (N > 1), AdamW.  `value` = rows of all ranks /
max-over-ranks device time with the inputs resident in HBM; `e2e` = the same step fed from pinned host memory with the
H2D copy of x and a D2H read of the step's metrics inside the timed region.

Workload (`--config auto`): N = 1 -> configs/pm_vae_power.py (BASELINE.json configs[1], "large batch on 1 B200");
N > 1 -> configs/pm_vae_hepmass.py (configs[2], "data-parallel at 1/2/4/8 B200").  Every line also carries
`all_configs`: a short train + cond-LL measurement of every config (gas, power, hepmass, bsds incl. K = 4096, mnist),
so the 1-GPU value of the N > 1 workload is `all_configs.hepmass.train.value` of the N = 1 line.

`--impl reference` times the reference's CPU implementation of the same step.  The reference is pure JAX/Haiku/TFP and
cannot be installed here, so this arm runs the oracle port (PyTorch CPU float32, all host threads) on
the same config and the same rows per step; it imports nothing from the product package.
"""
from __future__ import annotations

import argparse
import importlib.util
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "PM-VAE train samples/s"
UNIT = "samples/s"
TRAIN_MFLOP = {"gas": 5.259, "power": 5.247, "hepmass": 5.339, "bsds": 18.868,    # SURVEY §8 (6 x sum in*out)
               "mnist": 609.5}                                                        # conv stacks + AR-GMM, per image
CONDLL_GFLOP = {"gas": 0.5507, "power": 0.5496, "hepmass": 0.5575, "bsds": 1.4137}  # is_log_prob, K = 512
EVALFN_GFLOP = {"gas": 0.8260, "power": 0.8244, "hepmass": 0.8363, "bsds": 2.1205}  # impute + is_log_prob, K = 512
UCI = ("gas", "power", "hepmass", "bsds")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="auto", choices=["auto"] + sorted(TRAIN_MFLOP))
    ap.add_argument("--batch", type=int, default=0, help="rows per GPU per step (0 = default for the precision)")
    ap.add_argument("--precision", default="auto", choices=["auto", "fp32", "bf16"])
    ap.add_argument("--eval-rows", type=int, default=2048, help="rows per cond-LL eval call (K = 512)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eval", action="store_true")
    ap.add_argument("--no-all-configs", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="drive the step from the host (Trainer.train_step) instead of "
                    "replaying the CUDA graph of pmvae_train_step")
    return ap.parse_args()


def workload_name(args) -> str:
    if args.config != "auto":
        return args.config
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    return "power" if max(world, args.gpus) <= 1 else "hepmass"


def config_module():
    """posterior_matching_b200/config.py loaded by path: the configs table without importing the package (whose
    __init__ maps libpmvae.so) -- the reference arm must not touch the product."""
    spec = importlib.util.spec_from_file_location("_pmvae_config_table", os.path.join(ROOT, "posterior_matching_b200", "config.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = mod
    spec.loader.exec_module(mod)
    return mod


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained"),
                "hbm_gbs": d["hbm_gbs"], "source": "measured"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def synthetic_x(D, B, seed):
    import numpy as np
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((B, D)).astype(np.float32)
    x += (1e-3 * rng.standard_normal((B, D))).astype(np.float32)   # training_noise folded once (utils.py:108-116)
    return x


def conditioned_init(model, seed):
    """Haiku-default init with the two TriL head matrices scaled by 0.1 (well-conditioned
    triangular solves; same weights the parity tests use)."""
    model.init(seed)
    for hn in ("posterior_dist/linear", "partial_posterior_dist/linear"):
        model.params[hn]["w"].mul_(0.1)
    model.mark_params_changed()


def default_rows(name, precision):
    if name == "mnist":
        return 2048
    if precision != "bf16":
        return 65536
    return 131072


# ----------------------------------------------------------------------------- CPU arm (oracle port; no product import)
def _oracle_setup(name, rows):
    import numpy as np
    import torch
    from oracle import model as M
    cfg = config_module().pm_vae_config(name)
    spec = M.spec_from_config(cfg.model.to_dict())
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    p = M.cast_params(M.init_params(spec, 3), torch.float32)
    for hn in ("posterior_dist/linear", "partial_posterior_dist/linear"):
        p[hn]["w"] *= 0.1
    x = torch.tensor(synthetic_x(spec.D, rows, 0))
    rng = np.random.default_rng(1)
    b = torch.tensor((rng.random(x.shape) < 0.5).astype(np.float32))
    return M, spec, p, x, b, rng, cores


def cpu_train_baseline(name, rows, min_seconds=8.0, max_iters=20):
    """The oracle restatement (PyTorch CPU float32, all host threads) doing the same train
    step (forward + backward + AdamW) on a bounded sample."""
    import torch
    M, spec, p, x, b, rng, cores = _oracle_setup(name, rows)
    m, v = M.zeros_like_params(p), M.zeros_like_params(p)
    eps = torch.tensor(rng.standard_normal((rows, spec.d)).astype("float32"))

    def step(i):
        _, _, g = M.loss_and_grads(p, spec, x, b, eps, 0.5)
        M.adamw_update(p, g, m, v, count=i, lr=1e-3, wd=1e-5)

    step(0)
    times = []
    t_all = time.perf_counter()
    i = 1
    while (time.perf_counter() - t_all < min_seconds or len(times) < 3) and len(times) < max_iters:
        t0 = time.perf_counter()
        step(i)
        times.append(time.perf_counter() - t0)
        i += 1
    times.sort()
    med = times[len(times) // 2]
    return {"value": rows / med, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{len(times)} train steps of {rows} rows ({name}; the GPU arm's step has more rows), PyTorch CPU "
                      "float32 oracle, median", "ms_per_step": med * 1e3}


def cpu_eval_baseline(name, rows, K, max_iters=3):
    """oracle eval_fn (impute + is_log_prob, eval_pm_vae_uci.py:82-94) in PyTorch CPU float32 on a bounded sample."""
    import torch
    M, spec, p, x, b, rng, cores = _oracle_setup(name, rows)
    e = [torch.tensor(rng.standard_normal((K, rows, spec.d)).astype("float32")) for _ in range(3)]
    times = []
    with torch.no_grad():
        for _ in range(max_iters):
            t0 = time.perf_counter()
            M.eval_fn(p, spec, x, b, *e)
            times.append(time.perf_counter() - t0)
    times.sort()
    med = times[len(times) // 2]
    return {"value": rows / med, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{len(times)} eval_fn calls of {rows} rows, K = {K} ({name}), PyTorch CPU float32 oracle, median",
            "ms_per_call": med * 1e3}


def run_reference(args):
    """--impl reference: the reference's own implementation of the path is JAX and cannot be installed here, so this
    arm times the oracle port on the host cores (rank 0 only), same config and rows per step as the GPU arm."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    name = workload_name(args)
    if name == "mnist":
        run_reference_mnist(args)
        return
    rows = args.batch or default_rows(name, "bf16")
    steps = max(1, min(args.steps, 20))
    warm = max(1, min(args.warmup, 3))
    import torch
    M, spec, p, x, b, rng, cores = _oracle_setup(name, rows)
    m, v = M.zeros_like_params(p), M.zeros_like_params(p)
    eps = torch.tensor(rng.standard_normal((rows, spec.d)).astype("float32"))
    for i in range(warm):
        _, _, g = M.loss_and_grads(p, spec, x, b, eps, 0.5)
        M.adamw_update(p, g, m, v, count=i, lr=1e-3, wd=1e-5)
    t0 = time.perf_counter()
    for i in range(steps):
        _, _, g = M.loss_and_grads(p, spec, x, b, eps, 0.5)
        M.adamw_update(p, g, m, v, count=warm + i, lr=1e-3, wd=1e-5)
    dt = time.perf_counter() - t0
    val = rows * steps / dt
    sample = (f"{steps} train steps of {rows} rows ({name}), the GPU arm's rows per GPU per step; oracle port (PyTorch CPU "
              "float32, all host threads): the JAX reference is not installable in this image")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"configs/pm_vae_{name}.py train step (mask+eps given, fwd, bwd, AdamW)",
                   "rows_per_gpu_per_step": rows, "features": spec.D,
                   "note": "host CPU only; one process regardless of --gpus (rank 0)"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ----------------------------------------------------------------------------- MNIST config
def mnist_cpu_baseline(rows=32, min_seconds=6.0, max_iters=6):
    """oracle/model_mnist.py (PyTorch CPU float64 autograd, all host threads): loss + every gradient of one batch."""
    import numpy as np
    import torch
    from oracle import model_mnist as MM
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    p = MM.init_params()
    rng = np.random.default_rng(0)
    x = torch.tensor((rng.random((rows, 28, 28, 1)) < 0.13).astype(np.float64))
    b = torch.tensor((rng.random((rows, 28, 28, 1)) < 0.5).astype(np.float64))
    eps = torch.tensor(rng.standard_normal((rows, MM.LATENT)))
    MM.loss_and_grads(p, x, b, eps)
    times, t_all = [], time.perf_counter()
    while (time.perf_counter() - t_all < min_seconds or len(times) < 2) and len(times) < max_iters:
        t0 = time.perf_counter()
        MM.loss_and_grads(p, x, b, eps)
        times.append(time.perf_counter() - t0)
    times.sort()
    med = times[len(times) // 2]
    return {"value": rows / med, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{len(times)} loss+gradient passes of {rows} images (mnist), PyTorch CPU float64 oracle, median",
            "ms_per_step": med * 1e3}


def run_reference_mnist(args):
    cb = mnist_cpu_baseline(rows=32, min_seconds=3.0 * max(1, min(args.steps, 5)), max_iters=max(2, min(args.steps, 8)))
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
                      "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                      "config": {"workload": "pm_vae_mnist loss + gradients", "rows_per_step": 32,
                                 "note": "bounded sample: the float64 autograd oracle needs ~0.1 s per image"},
                      "cpu_baseline": cb,
                      "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


class Dist:
    """Rank plumbing shared by the GPU legs."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() in ("VERSION", "WARN"):
            os.environ.pop("NCCL_DEBUG", None)     # keep stdout to the one JSON line (both print a banner there)
            os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        if self.world == 1:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, fn, reps):
        """Device time of `reps` calls of fn (ms per call), bracketed by barrier + synchronize, max over ranks."""
        torch = self.torch
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        self.barrier()
        return self.max_over_ranks(e0.elapsed_time(e1)) / reps

    def close(self):
        """Multi-rank exit: synchronise, then leave without tearing the NCCL communicator down.  With CUDA graphs alive
        that captured NCCL kernels (the overlapped gradient exchange), destroy_process_group / interpreter exit blocked
        for minutes on this driver + NCCL build (measured in round 2); every result has been printed by now."""
        if self.world > 1:
            self.torch.cuda.synchronize()
            try:
                self.dist.barrier()
                self.torch.cuda.synchronize()
            except Exception:  # noqa: BLE001
                pass
            sys.stdout.flush()
            sys.stderr.flush()
            os._exit(0)


def mnist_leg(dd: Dist, precision, B, K, W, with_e2e=True, load_seconds=0.6):
    """configs/pm_vae_mnist.py train step: ConvEncoder / ConvDecoder, Bernoulli decoder, AutoregressiveGMM partial
    posterior, MNISTMaskGenerator masks (conv_vae.py)."""
    import numpy as np
    import torch
    from posterior_matching_b200 import MNISTMaskGenerator, PosteriorMatchingVAE, _lib, pm_vae_config
    dist, world, rank = dd.dist, dd.world, dd.rank
    m = PosteriorMatchingVAE.from_config(pm_vae_config("mnist").model.to_dict(), precision=precision)
    m.init(3)
    m.params["posterior_dist/linear"]["w"].mul_(0.1)
    rng = np.random.default_rng(100 + rank)
    x_host = torch.from_numpy((rng.random((B, 28, 28, 1)) < 0.13).astype(np.float32)).pin_memory()
    x_dev = x_host.cuda()
    gen = MNISTMaskGenerator(seed=1 + rank)
    sync = None
    if world > 1:
        def sync(ts):
            for t in ts:
                dist.all_reduce(t)

    def step(xd, i, read):
        return m.train_step(xd, gen((B, 28, 28, 1)), rng=(7, i), grad_sync=sync, global_rows=B * world, row_start=rank * B,
                            sync_metrics=read)

    for i in range(W):
        step(x_dev, i, False)
    dd.barrier()
    t_load = time.perf_counter()
    while time.perf_counter() - t_load < load_seconds:
        step(x_dev, 0, False)
        torch.cuda.synchronize()
    dd.barrier()
    l0 = int(_lib.lib.pmvae_launch_count())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        step(x_dev, W + i, False)
    e1.record()
    dd.barrier()
    ms = dd.max_over_ranks(e0.elapsed_time(e1))
    launches = int(_lib.lib.pmvae_launch_count()) - l0
    out = {"value": world * B * K / (ms * 1e-3), "ms_per_step": ms / K, "launches": launches, "rows_per_gpu_per_step": B}
    if with_e2e:
        # end to end: the batch comes from pinned host memory every step, the batch means are read back every step
        x_buf = torch.empty_like(x_dev)
        step(x_dev, 0, True)
        dd.barrier()
        t0 = time.perf_counter()
        last = None
        for i in range(K):
            x_buf.copy_(x_host, non_blocking=True)
            last = step(x_buf, W + K + i, True)
        dd.barrier()
        e2e_s = dd.max_over_ranks(time.perf_counter() - t0)
        out["e2e"] = {"value": world * B * K / e2e_s, "unit": UNIT, "h2d_bytes_per_step": B * 784 * 4, "d2h_bytes_per_step": 12,
                      "ms_per_step": e2e_s / K * 1e3, "feed": "pinned host images copied in every step; masks drawn on the device"}
        out["loss"] = last.get("loss") if isinstance(last, dict) else None
    del m
    torch.cuda.empty_cache()
    return out


def main_mnist(args):
    import torch
    from posterior_matching_b200 import conv as PC
    dd = Dist()
    world, rank = dd.world, dd.rank
    precision = "fp32" if args.precision == "fp32" else "bf16"
    B = args.batch or 2048
    K, W = args.steps, max(args.warmup, 3)
    sampler = ClockSampler(dd.local)
    if rank == 0:
        sampler.start()
    leg = mnist_leg(dd, precision, B, K, W)
    clocks = sampler.stop() if rank == 0 else None
    ms = leg["ms_per_step"]
    # dominant operator alone: the largest convolution of the encoders (14x14, 32 -> 64 channels, 5x5) forward
    peaks = measured_peaks()
    d = PC.conv_desc(14, 14, 32, 64, 5, 1, "SAME", precision=precision)
    xin = torch.randn(B, 14, 14, 32, device="cuda")
    wc = torch.randn(5, 5, 32, 64, device="cuda") / 28.0
    bc = torch.zeros(64, device="cuda")
    for _ in range(3):
        PC.conv2d_forward(d, xin, wc, bc)
    torch.cuda.synchronize()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record()
    for _ in range(10):
        PC.conv2d_forward(d, xin, wc, bc)
    k1.record()
    torch.cuda.synchronize()
    k_ms = k0.elapsed_time(k1) / 10
    k_flop = 2.0 * B * 14 * 14 * 800 * 64
    step_tflops = TRAIN_MFLOP["mnist"] * 1e6 * world * B / (ms * 1e-3) / 1e12
    roofline = {"bound": "tensor", "achieved": k_flop / (k_ms * 1e-3) / 1e12, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                "frac": k_flop / (k_ms * 1e-3) / 1e12 / peaks["bf16_tflops"], "traffic": None,
                "kernel": f"pmvae_conv2d_forward 14x14x32 -> 64, 5x5 ({B} images), timed as one operator (10 calls)",
                "kernel_ms": k_ms, "peak_source": peaks["source"] + " (bf16 burst)", "algorithmic_flop_per_launch": k_flop,
                "step_tflops_per_gpu": step_tflops / world,
                "step_frac_of_sustained_peak": step_tflops / world / (peaks["bf16_tflops_sustained"] or peaks["bf16_tflops"])}
    if rank == 0:
        cb = None if (args.no_cpu_baseline or world > 1) else mnist_cpu_baseline()      # rank 0 at N = 1 only
        print(json.dumps({
            "metric": METRIC, "value": leg["value"], "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": precision if precision != "fp32" else "f32",
            "data": "synthetic",
            "config": {"workload": "configs/pm_vae_mnist.py train step (MNIST masks, conv encoders / decoder, Bernoulli + "
                                   "AR-GMM terms, fwd, bwd, grad all-reduce, Adam)", "rows_per_gpu_per_step": B,
                       "global_batch": B * world, "parallelism": f"dp{world}", "l2": "activations + im2col buffers >> L2",
                       "launch": "host-composed libpmvae operator calls (conv_vae.py)"},
            "clocks": clocks, "e2e": leg["e2e"], "gpu_launches": leg["launches"], "roofline": roofline, "cpu_baseline": cb,
            "loss": leg.get("loss")}))
    dd.close()


# ----------------------------------------------------------------------------- UCI legs
def uci_setup(name, precision, B, rank, seed=0):
    import torch
    from posterior_matching_b200 import PosteriorMatchingVAE, Trainer, pm_vae_config
    cfg = pm_vae_config(name)
    model = PosteriorMatchingVAE.from_config(cfg.model, precision=precision)
    conditioned_init(model, 3)
    tr = Trainer(cfg, seed=seed, precision=precision, model=model)
    x_host = torch.from_numpy(synthetic_x(model.num_features, B, 100 + rank)).pin_memory()
    return cfg, model, tr, x_host


def eval_leg(dd: Dist, model, name, x_dev, Be, Ke, n_eval=3, with_e2e=True):
    """eval_pm_vae_uci.py's eval_fn (impute, mean over K, is_log_prob) on Be rows per rank, K = Ke samples."""
    import torch
    from posterior_matching_b200 import eval_fn
    world, rank = dd.world, dd.rank
    D = model.num_features
    xe = x_dev[:Be].contiguous()
    be = (torch.rand(Be, D, device="cuda", generator=torch.Generator(device="cuda").manual_seed(5 + rank)) < 0.5).float()

    def call(xd, bd):
        return eval_fn(model, (0, 91), xd, bd, Ke, row_start=rank * Be, total_rows=world * Be)

    for _ in range(2):
        call(xe, be)
    cms = dd.timed(lambda: call(xe, be), n_eval)
    cval = world * Be / (cms * 1e-3)
    gf = EVALFN_GFLOP[name] * Ke / 512.0
    peaks = measured_peaks()
    out = {"metric": "PM-VAE cond-LL eval samples/s (eval_fn = impute + is_log_prob)", "value": cval, "unit": UNIT, "K": Ke,
           "rows_per_gpu_per_call": Be, "ms_per_call": cms, "tflops_per_gpu": gf * 1e9 * cval / world / 1e12,
           "frac_of_tensor_peak": gf * 1e9 * cval / world / 1e12 / peaks["bf16_tflops"]}
    if with_e2e:
        # host buffers: x and b come from pinned host memory, imputations [Be, D] and cond-LL [Be] go back to the host
        xh, bh = xe.cpu().pin_memory(), be.cpu().pin_memory()
        imp_h = torch.empty((Be, D), dtype=torch.float32).pin_memory()
        ll_h = torch.empty(Be, dtype=torch.float32).pin_memory()
        xb, bb = torch.empty_like(xe), torch.empty_like(be)

        def call_host():
            xb.copy_(xh, non_blocking=True)
            bb.copy_(bh, non_blocking=True)
            imp, ll = call(xb, bb)
            imp_h.copy_(imp, non_blocking=True)
            ll_h.copy_(ll, non_blocking=True)
            torch.cuda.synchronize()

        call_host()
        dd.barrier()
        t0 = time.perf_counter()
        for _ in range(n_eval):
            call_host()
        dd.barrier()
        es = dd.max_over_ranks(time.perf_counter() - t0) / n_eval
        out["e2e"] = {"value": world * Be / es, "unit": UNIT, "h2d_bytes_per_call": 2 * Be * D * 4, "d2h_bytes_per_call": Be * (D + 1) * 4,
                      "ms_per_call": es * 1e3}
    return out


def train_leg(dd: Dist, tr, x_dev, K, W, graph=True):
    step = tr.train_step_fused if graph else tr.train_step
    for _ in range(W):
        step(x_dev)
    ms = dd.timed(lambda: step(x_dev), K)
    return ms


def all_configs_block(dd: Dist, precision, eval_rows=2048):
    """Every config of BASELINE.json in one short measurement each (3 warm-up + 8 timed train steps; 2 + 2 eval_fn calls)."""
    import torch
    peaks = measured_peaks()
    sus = peaks["bf16_tflops_sustained"] or peaks["bf16_tflops"]
    keep = ("value", "unit", "K", "rows_per_gpu_per_call", "ms_per_call", "tflops_per_gpu", "frac_of_tensor_peak")
    out = {}
    for name in UCI:
        B = default_rows(name, precision)
        cfg, model, tr, x_host = uci_setup(name, precision, B, dd.rank)
        x_dev = x_host.cuda()
        entry = {}
        ms = train_leg(dd, tr, x_dev, 8, 3)
        val = dd.world * B / (ms * 1e-3)
        tf = TRAIN_MFLOP[name] * 1e6 * val / dd.world / 1e12
        entry["train"] = {"value": val, "unit": UNIT, "ms_per_step": ms, "rows_per_gpu_per_step": B,
                          "tflops_per_gpu": tf, "frac_of_sustained_peak": tf / sus}
        ev = eval_leg(dd, model, name, x_dev, eval_rows, 512, n_eval=2, with_e2e=False)
        entry["cond_ll"] = {k: ev[k] for k in keep}
        if name == "bsds":      # BASELINE.json configs[3]: "importance-sampled cond-LL with large K"
            ev = eval_leg(dd, model, name, x_dev, 256, 4096, n_eval=2, with_e2e=False)
            entry["cond_ll_K4096"] = {k: ev[k] for k in keep}
        out[name] = entry
        del model, tr, x_dev
        torch.cuda.empty_cache()
    leg = mnist_leg(dd, "bf16" if precision == "bf16" else "fp32", 2048, 5, 3, with_e2e=False, load_seconds=0.0)
    tf = TRAIN_MFLOP["mnist"] * 1e6 * leg["value"] / dd.world / 1e12
    out["mnist"] = {"train": {"value": leg["value"], "unit": UNIT, "ms_per_step": leg["ms_per_step"], "rows_per_gpu_per_step": 2048,
                              "tflops_per_gpu": tf, "frac_of_sustained_peak": tf / sus}}
    return out


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return
    name = workload_name(args)
    if name == "mnist":
        main_mnist(args)
        return
    import ctypes as C
    import torch
    from posterior_matching_b200 import HostFeeder, _lib

    dd = Dist()
    world, rank = dd.world, dd.rank
    precision = args.precision
    if precision == "auto":
        probe = _lib.make_config(8, 16, 256, 2, 2, 2, 0, 0, 0, 1, _lib.PREC_BF16)
        precision = "bf16" if _lib.lib.pmvae_workspace_bytes(C.byref(probe), 128, 0) > 0 else "fp32"
    B = args.batch or default_rows(name, precision)
    K, W = args.steps, max(args.warmup, 3)
    cfg, model, tr, x_host = uci_setup(name, precision, B, rank)
    D = model.num_features
    x_dev = x_host.cuda()
    step_fn = tr.train_step if args.no_graph else tr.train_step_fused   # one pmvae_train_step per step, CUDA graph

    # ---- device-resident timing
    sampler = ClockSampler(dd.local)
    if rank == 0:
        sampler.start()
    for _ in range(W):
        step_fn(x_dev)
    dd.barrier()
    # nvidia-smi samples every 100 ms and a step is a few ms: keep the same load running (untimed) long enough
    # for the clock record to describe the state the timed steps run in
    t_load = time.perf_counter()
    while time.perf_counter() - t_load < 0.6:
        for _ in range(10):
            step_fn(x_dev)
        torch.cuda.synchronize()
    l0 = int(_lib.lib.pmvae_launch_count())
    ms = dd.timed(lambda: step_fn(x_dev), K) * K
    launches = int(_lib.lib.pmvae_launch_count()) - l0
    if not args.no_graph and tr.graph_launches_per_step:
        launches = K * tr.graph_launches_per_step     # replayed graph nodes are not seen by the enqueue-time counter
    clocks = sampler.stop() if rank == 0 else None
    metrics = tr.metrics()
    value = world * B * K / (ms * 1e-3)

    # ---- end-to-end: host buffers; every step's x comes from pinned host memory (H2D inside the timed region,
    #      double-buffered on a copy stream like the reference's prefetching input pipeline) and every step's
    #      metrics are read back to the host (D2H, synchronises the step)
    feeder = HostFeeder((B, D))
    slot = feeder.put(x_host)
    for _ in range(2):
        xb = feeder.get(slot); slot = feeder.put(x_host); step_fn(xb); tr.metrics()
    dd.barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        xb = feeder.get(slot)           # this step's batch (its copy was started during the previous step)
        slot = feeder.put(x_host)       # start the next batch's H2D copy
        step_fn(xb)
        tr.metrics()                    # D2H of the three batch sums
    dd.barrier()
    e2e_s = dd.max_over_ranks(time.perf_counter() - t0)
    e2e = {"value": world * B * K / e2e_s, "unit": UNIT, "h2d_bytes_per_step": B * D * 4, "d2h_bytes_per_step": 12,
           "ms_per_step": e2e_s / K * 1e3, "feed": "pinned host batches, double-buffered H2D on a copy stream"}

    # ---- dominant kernel alone: the TRAINING-mode fused ResidualMLP forward (net_fwd_kernel<SAVE=1>: encoder net + TriL
    #      head, every operand tile + relu bits stored for the backward), the variant the timed step launches three times
    peaks = measured_peaks()
    H = 256
    stream = torch.cuda.current_stream().cuda_stream
    d_lat = model.latent_dim
    P = d_lat + d_lat * (d_lat + 1) // 2
    R_enc = int(model.cfg.R_enc)
    enc_macs = D * H + 2 * R_enc * H * H + H * P            # Linear MACs per row (SURVEY §8 convention)
    extra = []

    def time_alone(fn, reps=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record()
        for _ in range(reps):
            fn()
        k1.record()
        torch.cuda.synchronize()
        return k0.elapsed_time(k1) / reps

    reps = 20
    if precision == "bf16":
        out_par = torch.empty(B, P, device="cuda")
        kname = (f"fused::net_fwd_kernel<SAVE> training-mode forward (tcgen05 chain: {1 + 2 * R_enc} hidden Linears + TriL head; "
                 f"operand tiles + relu bits stored by TMA) {B} rows via pmvae_net_apply(encoder | PMVAE_NET_SAVE)")
        k_flop = 2.0 * enc_macs * B
        k_ms = time_alone(lambda: model.net_apply(0 | _lib.NET_SAVE, x_dev, None, out_par), reps)
        t_eval = time_alone(lambda: model.net_apply(0, x_dev, None, out_par), reps)
        extra.append({"kernel": "fused::net_fwd_kernel<SAVE=0> (evaluation-mode forward of the same net: no activation stores)",
                      "ms": t_eval, "tflops": k_flop / (t_eval * 1e-3) / 1e12,
                      "frac_of_burst_peak": k_flop / (t_eval * 1e-3) / 1e12 / peaks["bf16_tflops"]})
        # algorithmic HBM bytes of the training-mode forward: x in, head out, (2R+1) bf16 operand tiles + relu bits
        k_bytes = B * (4.0 * D + 4.0 * P + (2 * R_enc + 1) * (512 + 32))
    else:
        xin = torch.randn(B, H, device="cuda")
        wt = torch.randn(H, H, device="cuda") / 16
        bias = torch.zeros(H, device="cuda")
        y = torch.empty(B, H, device="cuda")
        kname = f"gemm_f32_kernel {B}x{H}x{H} via pmvae_linear"
        k_flop = 2.0 * B * H * H
        k_bytes = B * H * 8.0
        k_ms = time_alone(lambda: _lib.check(_lib.lib.pmvae_linear(_lib.PREC_F32, xin.data_ptr(), wt.data_ptr(), bias.data_ptr(),
                                                                   B, H, H, 1, y.data_ptr(), None, 0, stream), "pmvae_linear"), reps)
    k_tflops = k_flop / (k_ms * 1e-3) / 1e12
    if precision == "bf16":
        # the two stand-alone tcgen05 GEMM shapes of the step, for reference
        xin = torch.randn(B, H, device="cuda").to(torch.bfloat16)
        wt = (torch.randn(H, H, device="cuda") / 16).to(torch.bfloat16)
        gy = torch.randn(B, H, device="cuda").to(torch.bfloat16)
        bias = torch.zeros(H, device="cuda")
        y = torch.empty(B, H, device="cuda")
        gw = torch.zeros(H, H, device="cuda")
        t_nt = time_alone(lambda: _lib.check(_lib.lib.pmvae_tc_gemm_nt(xin.data_ptr(), H, wt.data_ptr(), H, bias.data_ptr(),
                                                                       B, H, H, y.data_ptr(), stream), "nt"))
        t_tn = time_alone(lambda: _lib.check(_lib.lib.pmvae_tc_gemm_tn(xin.data_ptr(), H, gy.data_ptr(), H, H, H, B,
                                                                       gw.data_ptr(), stream), "tn"))
        for nm, t, byts in (("tc_gemm_kernel<NT> one hidden Linear, fp32 out", t_nt, B * H * 6.0),
                            ("tc_gemm_kernel<TN> one weight gradient (act^T @ dY, split over rows)", t_tn, B * H * 4.0)):
            extra.append({"kernel": nm, "ms": t, "tflops": 2.0 * B * H * H / (t * 1e-3) / 1e12,
                          "hbm_gbs": byts / (t * 1e-3) / 1e9, "hbm_frac": byts / (t * 1e-3) / 1e9 / peaks["hbm_gbs"]})
        del xin, y, gy
    traffic, step_dram, traffic_src = None, None, None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        try:
            with open(tpath) as f:
                tall = json.load(f)
                tj = tall.get(f"{precision}:{name}", tall.get(precision, {}))     # per-config entry, else the headline's
            if tj.get("config", name) == name and tj.get("rows", B) == B:
                traffic = tj.get("dram_bytes_per_launch_train")
                step_dram = tj.get("step_dram_bytes")
                traffic_src = tj.get("source")
        except Exception:  # noqa: BLE001
            traffic = None
    step_s = ms / K * 1e-3
    step_tflops = TRAIN_MFLOP[name] * 1e6 * world * B / step_s / 1e12
    roofline = {"bound": "tensor", "achieved": k_tflops, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                "frac": k_tflops / peaks["bf16_tflops"], "traffic": traffic, "traffic_source": traffic_src,
                "kernel": kname + f", timed alone ({reps} launches)",
                "kernel_ms": k_ms, "peak_source": peaks["source"] + " (bf16 burst)",
                "algorithmic_flop_per_launch": k_flop, "algorithmic_bytes_per_launch": k_bytes,
                "kernel_hbm_gbs_algorithmic": k_bytes / (k_ms * 1e-3) / 1e9,
                "step_tflops_per_gpu": step_tflops / world,
                "step_frac_of_sustained_peak": step_tflops / world / (peaks["bf16_tflops_sustained"] or peaks["bf16_tflops"]),
                "other_kernels": extra}
    if step_dram:
        # whole step against the HBM roofline: DRAM bytes of one step (sum over its kernels, one ncu --set full capture of
        # this command: profiles/roofline_traffic.json) / measured step time
        roofline["step_hbm"] = {"dram_bytes_per_step": step_dram, "gbs": step_dram / step_s / 1e9,
                                "frac_of_hbm_peak": step_dram / step_s / 1e9 / peaks["hbm_gbs"],
                                "algorithmic_bytes_per_step": B * (4.0 * D) + 16.0 * model.n_arena}

    # ---- cond-LL evaluation throughput: eval_pm_vae_uci.py's eval_fn (impute + is_log_prob), K = 512
    cond = None
    if not args.no_eval:
        cond = eval_leg(dd, model, name, x_dev, args.eval_rows, 512)
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            cond["cpu_baseline"] = cpu_eval_baseline(name, 64, 512)

    # ---- latency point: the reference's own batch size (configs/pm_vae_*.py train_batch_size = 512)
    ref_batch = None
    if not args.no_graph:
        Br = int(cfg.data.train_batch_size)
        xr = x_dev[:Br].contiguous()
        for _ in range(5):
            tr.train_step_fused(xr)
        rms = dd.timed(lambda: tr.train_step_fused(xr), 100)
        ref_batch = {"rows_per_gpu_per_step": Br, "ms_per_step": rms, "value": world * Br / (rms * 1e-3), "unit": UNIT,
                     "note": "train_batch_size of the reference config; launch-latency bound"}

    del tr, model, x_dev
    torch.cuda.empty_cache()
    allc = None
    if not args.no_all_configs:
        allc = all_configs_block(dd, precision, eval_rows=args.eval_rows)

    # ---- CPU baseline (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_train_baseline(name, 8192)

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if precision == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": f"configs/pm_vae_{name}.py train step (mask+eps draw, fwd, bwd, "
                                   f"{'NCCL grad all-reduce, ' if world > 1 else ''}AdamW)",
                       "rows_per_gpu_per_step": B, "global_batch": world * B, "features": D,
                       "parallelism": f"dp{world}", "accumulate": "fp32",
                       "launch": "host-driven C-ABI calls" if args.no_graph else "pmvae_train_step replayed as a CUDA graph",
                       "l2": "saved activations of one step exceed the 126 MB L2 (no explicit flush)",
                       "weights": "Haiku-default init, TriL heads x0.1",
                       "workload_choice": "N = 1: power (BASELINE.json configs[1]); N > 1: hepmass (configs[2]); the 1-GPU "
                                          "value of the N > 1 workload is all_configs.hepmass.train of the N = 1 line"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
            "train_metrics": metrics,
        }
        if cond is not None:
            out["cond_ll_eval"] = cond
        if ref_batch is not None:
            out["reference_batch"] = ref_batch
        if allc is not None:
            out["all_configs"] = allc
        if cpu is not None:
            out["cpu_baseline"] = cpu
        print(json.dumps(out))
    dd.close()


if __name__ == "__main__":
    main()
