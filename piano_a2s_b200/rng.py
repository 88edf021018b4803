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
    # masks are i.i.d. Bernoulli draws from the device generator: a caller may ask for them in any order and in bulk (one launch
    # for all steps of all bars) instead of the reference's call-by-call order, which only a replaying source has to honour
    iid = True

    def coin(self) -> float:
        return random.random()

    def coins(self, n: int):
        return [random.random() for _ in range(n)]

    def dropout_mask(self, shape, p: float, device, kind: str) -> torch.Tensor:
        """{0, 1/(1-p)} mask; `kind` in {"conv", "bar_token", "note_steps"} tells a replaying source what is asked for."""
        return torch.empty(shape, device=device, dtype=torch.float32).bernoulli_(1.0 - p).mul_(1.0 / (1.0 - p))


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
