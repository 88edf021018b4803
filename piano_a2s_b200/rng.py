"""Randomness used by the hot path, in the reference's consumption order.

The reference draws one python `random.random()` coin per executed note step and one per bar (models.py:404, :289)
and uses `F.dropout` on token embeddings (p=0.1, models.py:239,:391) and conv features (p=0.2, models.py:541).
The CUDA decoder runs all steps of a (bar, staff) inside one call, so coins and masks are drawn *before* the call,
in the same order, and handed to the kernels.  `SOURCE` can be swapped (tests replay recorded coins/masks).
"""
import contextlib
import random

import torch


class DeviceRandom:
    def coin(self) -> float:
        return random.random()

    def coins(self, n: int):
        return [random.random() for _ in range(n)]

    def dropout_mask(self, shape, p: float, device, kind: str) -> torch.Tensor:
        """{0, 1/(1-p)} mask; `kind` in {"conv", "bar_token", "note_steps"} tells a replaying source what is asked for."""
        return (torch.rand(shape, device=device) >= p).to(torch.float32).div_(1.0 - p)


SOURCE = DeviceRandom()


@contextlib.contextmanager
def use_source(src):
    global SOURCE
    old = SOURCE
    SOURCE = src
    try:
        yield src
    finally:
        SOURCE = old


def source():
    return SOURCE
