"""Training step of train_pm_vae.py on the device: loss_fn (:58-72), the beta schedules
(:28-43, utils.py:124-136), the optax chain (:74-83) and the step/pmean semantics of
bax.Trainer, with the batch sharded by rows across ranks and one
all-reduce (NCCL through torch.distributed) of the flat gradient arena.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Any, Callable, Dict, Mapping, Optional

import torch

from . import _lib, prng
from .masking import get_mask_generator
from .vae import PosteriorMatchingVAE, _stream


def cyclical_annealing_schedule(low_value: float, high_value: float, period: int, delay: int = 0) -> Callable:
    """utils.py:124-136."""
    def schedule(count):
        true_count = count
        count = count - delay
        count = min(max(count % period, 0), period // 2)
        frac = 1 - count / (period // 2)
        x = (low_value - high_value) * frac + high_value
        return x * (1.0 if true_count >= delay else 0.0)
    return schedule


def linear_schedule(init_value, end_value, transition_steps, transition_begin=0) -> Callable:
    """optax.linear_schedule."""
    def schedule(count):
        frac = min(max((count - transition_begin) / transition_steps, 0.0), 1.0)
        return init_value + (end_value - init_value) * frac
    return schedule


def exponential_decay(init_value, transition_steps, decay_rate) -> Callable:
    """optax.exponential_decay (continuous)."""
    return lambda count: init_value * decay_rate ** (count / transition_steps)


def get_beta_schedule(config: Mapping[str, Any]) -> Callable:
    """train_pm_vae.py:28-43."""
    if "schedule" not in config:
        return lambda x: 1.0
    if config["schedule"] == "monotonic":
        return linear_schedule(config["low_value"], config["high_value"], config["transition_steps"],
                               config["transition_begin"])
    if config["schedule"] == "cyclic":
        return cyclical_annealing_schedule(config["low_value"], config["high_value"], config["period"],
                                           config["delay"])
    raise ValueError(config["schedule"])


class HostFeeder:
    """Double-buffered host -> device feed (what the reference's tf.data prefetch + device_put does for
    train_pm_vae.py's loop): batch i+1 is copied from pinned host memory on a side stream while step i runs."""

    def __init__(self, shape, device=None):
        self.device = torch.device("cuda" if device is None else device)
        self.bufs = [torch.empty(shape, dtype=torch.float32, device=self.device) for _ in range(2)]
        self.events = [torch.cuda.Event(), torch.cuda.Event()]
        self.stream = torch.cuda.Stream(device=self.device)
        self.i = 0

    def put(self, x_host: torch.Tensor) -> int:
        """Starts the copy of `x_host` (pinned) into the next buffer; returns its slot."""
        slot = self.i & 1
        self.i += 1
        self.stream.wait_stream(torch.cuda.current_stream())     # the step that last read this buffer has been issued
        with torch.cuda.stream(self.stream):
            self.bufs[slot].copy_(x_host, non_blocking=True)
            self.events[slot].record(self.stream)
        return slot

    def get(self, slot: int) -> torch.Tensor:
        """The device batch of `slot`, ordered after its copy on the current stream."""
        torch.cuda.current_stream().wait_event(self.events[slot])
        return self.bufs[slot]


class Trainer:
    """One object per rank.  `train_step(x)` = mask draw, eps draw, forward, loss
    cotangents, backward, gradient all-reduce, AdamW -- every launch on the current CUDA
    stream, no host sync; metrics stay on the device until `metrics()` is called."""

    def __init__(self, config: Mapping[str, Any], *, seed: int = 0, precision: str = "bf16", device=None,
                 process_group=None, model: Optional[PosteriorMatchingVAE] = None):
        self.config = config
        self.model = model or PosteriorMatchingVAE.from_config(config["model"], precision=precision, device=device)
        if not isinstance(self.model, PosteriorMatchingVAE):
            raise NotImplementedError(
                "Trainer drives the ResidualMLP / TriLGaussian model (the four UCI configs); "
                "ConvPosteriorMatchingVAE (configs/pm_vae_mnist.py) has its own train_step")
        if model is None:
            self.model.init(seed)
        self.device = self.model.device
        self.beta_schedule = get_beta_schedule(config.get("beta", {}) or {})
        self.matching_coef = float(config.get("matching_coef", 1.0))
        ls = config["lr_schedule"]
        self.lr_schedule = exponential_decay(ls["init_value"], ls["transition_steps"], ls["decay_rate"])
        self.weight_decay = float(config.get("weight_decay", 0.0))
        adam = dict(config.get("adam", {}) or {})
        self.b1, self.b2, self.adam_eps = adam.get("b1", 0.9), adam.get("b2", 0.999), adam.get("eps", 1e-8)
        self.m = torch.zeros_like(self.model.arena)
        self.v = torch.zeros_like(self.model.arena)
        self.step = 0
        self.pg = process_group
        self.world = 1
        self.rank = 0
        if process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.world = torch.distributed.get_world_size(process_group)
            self.rank = torch.distributed.get_rank(process_group)
        data = config.get("data", {}) or {}
        self.mask_generator = get_mask_generator(data.get("mask_generator", "BernoulliMaskGenerator"),
                                                 seed=seed + 1, device=self.device)
        self._rng = prng.PRNGSequence(prng.PRNGKey(seed + 2))
        self._cot = None
        self._sums = self.model.grad_tail[:3]       # batch sums ride along with the gradient all-reduce
        self.last_beta = 1.0
        self._seed = seed
        self._fused = None                          # state of train_step_fused (device step state, graphs)
        self._comm_stream = None
        # gradient ranges of the three nets (contiguous in the arena, in backward order: decoder, encoder, partial
        # encoder); the last one carries the 64-float tail with the batch sums
        starts = {}
        for name, rows, cols, w_off, b_off in self.model.leaves:
            grp = "enc" if name.startswith(("encoder_net", "posterior_dist")) else (
                "dec" if name.startswith(("decoder_net", "decoder_dist")) else "part")
            starts[grp] = min(starts.get(grp, w_off), w_off)
        n_store = self.model._grad_store.numel()
        assert starts["enc"] == 0 and starts["enc"] < starts["dec"] < starts["part"]
        self._buckets = {"enc": (0, starts["dec"]), "dec": (starts["dec"], starts["part"]), "part": (starts["part"], n_store)}

    # ---- the step as one launch sequence / CUDA graph --------------------------------------------
    def _train_config(self) -> "_lib.TrainConfig":
        tc = _lib.TrainConfig()
        beta = dict(self.config.get("beta", {}) or {})
        kind = beta.get("schedule")
        tc.beta_schedule = {None: _lib.BETA_CONST, "cyclic": _lib.BETA_CYCLIC, "monotonic": _lib.BETA_MONOTONIC}[kind]
        tc.beta_low, tc.beta_high = float(beta.get("low_value", 1.0)), float(beta.get("high_value", 1.0))
        tc.beta_period, tc.beta_delay = int(beta.get("period", 1)), int(beta.get("delay", 0))
        tc.beta_transition_steps = int(beta.get("transition_steps", 1))
        tc.beta_transition_begin = int(beta.get("transition_begin", 0))
        tc.matching_coef = self.matching_coef
        ls = self.config["lr_schedule"]
        tc.lr_init, tc.lr_decay_rate = float(ls["init_value"]), float(ls["decay_rate"])
        tc.lr_transition_steps = int(ls["transition_steps"])
        tc.weight_decay, tc.adam_b1, tc.adam_b2, tc.adam_eps = self.weight_decay, self.b1, self.b2, self.adam_eps
        tc.mask_p = float(getattr(self.mask_generator, "p", 0.5))
        return tc

    def _fused_setup(self, B: int):
        """Moves the host loop's state (rng sequence, mask-call counter, step) into the device state block."""
        mdl = self.model
        f = {"B": B, "tc": self._train_config(), "graphs": None}
        f["state"] = torch.zeros(int(_lib.lib.pmvae_train_state_bytes()), dtype=torch.uint8, device=self.device)
        f["scratch"] = torch.empty(int(_lib.lib.pmvae_train_scratch_floats(mdl._cfgp, B)), dtype=torch.float32,
                                   device=self.device)
        f["x"] = torch.empty((B, mdl.num_features), dtype=torch.float32, device=self.device)
        _lib.check(_lib.lib.pmvae_train_state_init(f["state"].data_ptr(), _lib.key_arg(self._rng.key),
                                                   _lib.key_arg(self.mask_generator._key), self.step,
                                                   self.mask_generator._calls, _stream()), "pmvae_train_state_init")
        self._fused = f
        return f

    def _fused_call(self, phase: int):
        f, mdl = self._fused, self.model
        ws = mdl.workspace(f["B"])
        _lib.check(_lib.lib.pmvae_train_step(
            mdl._cfgp, C.byref(f["tc"]), mdl.arena.data_ptr(), self.m.data_ptr(), self.v.data_ptr(),
            mdl.grad_arena.data_ptr(), f["state"].data_ptr(), f["x"].data_ptr(), f["B"], f["B"] * self.world,
            self.rank * f["B"], f["scratch"].data_ptr(), self._sums.data_ptr(), ws.data_ptr(), ws.numel(), phase,
            _stream()), "pmvae_train_step")

    def _step_sequence(self):
        """The launch sequence of one step on the current stream.  One rank: a single pmvae_train_step.  Several ranks:
        the backward runs in three stages (decoder, encoder, partial encoder) and each finished gradient range is
        all-reduced (NCCL) on a communication stream while the next stage computes; only the last range's exchange is
        exposed before AdamW.  The whole sequence (both streams) is what `train_step_fused` captures in ONE CUDA graph."""
        mdl = self.model
        if self.world == 1:
            self._fused_call(3)
            return
        if os.environ.get("PMVAE_DP_OVERLAP", "1") == "0":      # serial exchange (tests compare the two)
            self._fused_call(1)
            torch.distributed.all_reduce(mdl._grad_store, group=self.pg)
            self._fused_call(2)
            return
        if self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream(device=self.device)
        main, comm = torch.cuda.current_stream(), self._comm_stream
        store = mdl._grad_store

        def exchange(bucket):
            lo, hi = self._buckets[bucket]
            comm.wait_stream(main)
            with torch.cuda.stream(comm):
                torch.distributed.all_reduce(store[lo:hi], group=self.pg)

        self._fused_call(1 | 16)          # forward, loss, decoder + latent backward
        exchange("dec")
        self._fused_call(4)               # encoder backward
        exchange("enc")
        self._fused_call(8)               # partial-encoder backward
        exchange("part")                  # + the three batch sums in the arena's tail
        main.wait_stream(comm)
        self._fused_call(2)               # AdamW on the summed gradients, operand-image refresh

    def train_step_fused(self, x: torch.Tensor, graph: bool = True):
        """Same step as `train_step(x)` (same keys, schedules and arithmetic) issued through pmvae_train_step:
        one C call per step, every step-dependent scalar derived on the device.  With `graph=True` the launch
        sequence is captured once per batch size in a CUDA graph and replayed (the second call onwards)."""
        mdl = self.model
        B = x.shape[0]
        if not isinstance(self.mask_generator, __import__("posterior_matching_b200.masking", fromlist=["x"]).BernoulliMaskGenerator):
            raise NotImplementedError("the fused step draws Bernoulli masks (the four UCI configs)")
        f = self._fused if (self._fused is not None and self._fused["B"] == B) else self._fused_setup(B)
        if f.get("stale"):
            # host-driven steps (train_step) ran since the last fused one: the device state block (keys, mask-call
            # counter, step) is re-seeded from the host mirrors; same buffer, so captured graphs stay valid
            _lib.check(_lib.lib.pmvae_train_state_init(f["state"].data_ptr(), _lib.key_arg(self._rng.key),
                                                       _lib.key_arg(self.mask_generator._key), self.step,
                                                       self.mask_generator._calls, _stream()), "pmvae_train_state_init")
            f["stale"] = False
        f["x"].copy_(x, non_blocking=True)
        ws = mdl.workspace(B)
        if f.get("ws_ptr") != ws.data_ptr():          # the workspace moved (e.g. an evaluator grew it): re-capture
            f["ws_ptr"], f["graphs"] = ws.data_ptr(), None
        mdl._prepare(ws)
        if not graph:
            self._step_sequence()
        else:
            if f["graphs"] is None:
                if f.get("warm", 0) < 1:             # one eager step first (lazy kernel attributes, workspace, NCCL channels)
                    f["warm"] = 1
                    return self.train_step_fused(x, graph=False)
                torch.cuda.synchronize()
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                n0 = int(_lib.lib.pmvae_launch_count())
                with torch.cuda.stream(side):
                    g = torch.cuda.CUDAGraph()
                    # thread_local: the NCCL watchdog thread may query events while this thread captures
                    with torch.cuda.graph(g, stream=side, capture_error_mode="thread_local"):
                        self._step_sequence()
                torch.cuda.current_stream().wait_stream(side)
                f["launches"] = int(_lib.lib.pmvae_launch_count()) - n0    # kernels one replayed step launches
                # capturing does not execute: the state block is untouched, so the first replay is this step
                f["graphs"] = [g]
            f["graphs"][0].replay()
        mdl._params_dirty = False                    # phase 2 refreshed the operand images
        # keep the host-side mirrors in step (so train_step / metrics keep working)
        self._rng.next()
        self.mask_generator._calls += 1
        self.last_beta = float(self.beta_schedule(self.step))
        self.step += 1
        self._last_Bg = B * self.world
        return self._sums

    def release_graphs(self):
        """Drops the captured step graph(s) and the device step state (re-created by the next fused step).  Call it
        before torch.distributed.destroy_process_group(): graphs that captured NCCL kernels keep the communicator busy."""
        if self._fused is not None:
            torch.cuda.synchronize()
            self._fused = None

    @property
    def graph_launches_per_step(self) -> int:
        """Kernels of this library inside one replayed step (counted while the graph was captured)."""
        return int(self._fused.get("launches", 0)) if self._fused else 0

    def fused_state(self) -> "_lib.TrainStateHost":
        out = _lib.TrainStateHost()
        _lib.check(_lib.lib.pmvae_train_state_read(self._fused["state"].data_ptr(), C.byref(out), _stream()),
                   "pmvae_train_state_read")
        return out

    def _buffers(self, B):
        if self._cot is None or self._cot.shape[1] < B:
            self._cot = torch.empty((3, B), dtype=torch.float32, device=self.device)
        return self._cot

    def train_step(self, x: torch.Tensor, b: Optional[torch.Tensor] = None, eps: Optional[torch.Tensor] = None):
        """x: this rank's rows [B_local, D] of a global batch of world*B_local rows."""
        mdl = self.model
        B = x.shape[0]
        Bg = B * self.world
        row0 = self.rank * B
        step_key = self._rng.next()     # the per-step key bax hands to the transformed loss_fn
        if b is None:
            b = self.mask_generator((B, mdl.num_features), row_start=row0, total_rows=Bg)
        out = mdl(x, b, is_training=True, rng=step_key if eps is None else None, eps=eps,
                  row_start=row0, total_rows=Bg)
        beta = float(self.beta_schedule(self.step))
        self.last_beta = beta
        cot = self._buffers(B)
        _lib.check(_lib.lib.pmvae_loss_cotangents(
            B, Bg, beta, self.matching_coef, out["reconstruction_ll"].data_ptr(), out["kl"].data_ptr(),
            out["matching_ll"].data_ptr(), cot[0].data_ptr(), cot[1].data_ptr(), cot[2].data_ptr(),
            self._sums.data_ptr(), _stream()), "pmvae_loss_cotangents")
        mdl.backward(cot[0, :B], cot[1, :B], cot[2, :B])
        if self.world > 1:
            # mean of per-rank mean-gradients == sum of the 1/B_global-scaled shard gradients
            torch.distributed.all_reduce(mdl._grad_store, group=self.pg)    # gradients + the three batch sums
        lr = float(self.lr_schedule(self.step))
        _lib.check(_lib.lib.pmvae_adamw(mdl._cfgp, mdl.arena.data_ptr(), mdl.grad_arena.data_ptr(),
                                        self.m.data_ptr(), self.v.data_ptr(), self.step, lr, self.weight_decay,
                                        self.b1, self.b2, self.adam_eps, _stream()), "pmvae_adamw")
        mdl.mark_params_changed()
        self.step += 1
        self._last_Bg = Bg
        if self._fused is not None:
            self._fused["stale"] = True      # the device step state is now one step behind the host mirrors
        return self._sums

    def metrics(self) -> Dict[str, float]:
        """Batch means of the last step (one D2H read): the aux dict of train_pm_vae.py:72."""
        s = (self._sums / float(self._last_Bg)).tolist()
        rec, kl, match = s
        return {"reconstruction_ll": rec, "kl": kl, "matching_ll": match, "beta": self.last_beta,
                "loss": -(rec - self.last_beta * kl) + self.matching_coef * (-match)}
