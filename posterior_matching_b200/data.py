"""Host input pipeline of the reference's scripts (posterior_matching/utils.py:36-121 `load_datasets`; eval_pm_vae_uci.py
:46-57 `load_dataset`) for arrays already in host memory: shuffle buffer, fixed-size batches with drop_remainder, image
rescaling, training noise, and the device-side mask draw that replaces `get_add_mask_fn`'s `tf.py_function`
(masking.py:338-350).  TFDS itself is outside the path: the arrays come from the caller
(`np.load`, a TFDS export, synthetic data).

Batches are dicts with the reference's keys -- "features" or "image", plus "mask" -- whose values are CUDA tensors; the
H2D copy of batch i + 1 overlaps step i (`HostFeeder`, pinned staging buffers).
"""
from __future__ import annotations

from typing import Any, Dict, Iterator, Mapping, Optional

import numpy as np
import torch

from .masking import get_mask_generator
from .train import HostFeeder


class ArrayDataset:
    """`tf.data` pipeline of utils.py:36-121 over one host array.

    config keys honoured (same names as `config.data`): train_batch_size / val_batch_size (via `batch_size`),
    buffer_size (shuffle buffer, default 40000 as utils.py:44), training_noise (utils.py:108-116, training split only),
    mask_generator (+ mask_generator_kwargs).  Images (uint8 [N, H, W, C]) are cast to float32 and divided by 255
    (utils.py:49-57)."""

    def __init__(self, array: np.ndarray, batch_size: int, *, training: bool, config: Optional[Mapping[str, Any]] = None,
                 seed: int = 0, device=None, drop_remainder: bool = True):
        config = dict(config or {})
        self.is_image = array.ndim == 4
        self.key = "image" if self.is_image else "features"
        arr = np.asarray(array)
        if self.is_image and arr.dtype == np.uint8:
            arr = arr.astype(np.float32) / 255.0
        self.array = np.ascontiguousarray(arr, dtype=np.float32)
        self.batch_size, self.training, self.drop_remainder = int(batch_size), bool(training), drop_remainder
        self.buffer_size = int(config.get("buffer_size", 40000))
        self.noise = float(config.get("training_noise", 0.0)) if training else 0.0
        self.rng = np.random.default_rng(seed)
        self.device = torch.device("cuda" if device is None else device)
        self.mask_generator = None
        if "mask_generator" in config:
            self.mask_generator = get_mask_generator(config["mask_generator"], seed=seed + 1, device=self.device,
                                                     **dict(config.get("mask_generator_kwargs", {}) or {}))
        shape = (self.batch_size,) + self.array.shape[1:]
        self._staging = [torch.empty(shape, dtype=torch.float32).pin_memory() for _ in range(2)]
        self._feeder = HostFeeder(shape, device=self.device)

    def __len__(self) -> int:
        n = self.array.shape[0]
        return n // self.batch_size if self.drop_remainder else -(-n // self.batch_size)

    def _order(self) -> np.ndarray:
        """tf.data's shuffle(buffer_size): a sliding buffer, not a full permutation (utils.py:44)."""
        n = self.array.shape[0]
        if not self.training:
            return np.arange(n)
        buf = list(range(min(self.buffer_size, n)))
        nxt = len(buf)
        out = np.empty(n, dtype=np.int64)
        for i in range(n):
            j = int(self.rng.integers(len(buf)))
            out[i] = buf[j]
            if nxt < n:
                buf[j] = nxt
                nxt += 1
            else:
                buf[j] = buf[-1]
                buf.pop()
        return out

    def _host_batch(self, idx: np.ndarray, slot: int) -> torch.Tensor:
        st = self._staging[slot]
        np.take(self.array, idx, axis=0, out=st.numpy())
        if self.noise > 0.0:
            st.numpy()[...] += (self.noise * self.rng.standard_normal(st.shape)).astype(np.float32)
        return st

    def __iter__(self) -> Iterator[Dict[str, torch.Tensor]]:
        order = self._order()
        nb = len(self)
        if nb == 0:
            return
        B = self.batch_size
        slot = self._feeder.put(self._host_batch(order[:B], 0))
        for i in range(nb):
            x = self._feeder.get(slot)
            if i + 1 < nb:
                # the staging buffer of the batch after next is free once its copy has been issued on the feeder stream
                self._feeder.stream.synchronize()
                slot = self._feeder.put(self._host_batch(order[(i + 1) * B:(i + 2) * B], (i + 1) & 1))
            batch = {self.key: x}
            if self.mask_generator is not None:
                batch["mask"] = self.mask_generator(tuple(x.shape))
            yield batch


def load_datasets(arrays: Mapping[str, np.ndarray], config: Mapping[str, Any], *, seed: int = 0, device=None):
    """utils.py:36-121 for host arrays: `arrays` maps split names to arrays; returns (train, val) iterables."""
    train = ArrayDataset(arrays[config.get("train_split", "train")], config["train_batch_size"], training=True,
                         config=config, seed=seed, device=device)
    val = ArrayDataset(arrays[config.get("validation_split", "validation")], config["val_batch_size"], training=False,
                       config=config, seed=seed + 7, device=device)
    return train, val
