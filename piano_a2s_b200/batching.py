"""Batch collation in front of the hot path (SURVEY 8 a2 + the target half of N2): what the reference does per item in
`Dataset.__getitem__` -- `pad_spectrogram`, `pad_score`, `key_to_int` (datasets/syn.py:38-74,88-121; asap.py:330-366), one
`.to(device)` per tensor per item -- done once per BATCH:

* spectrograms: the clips' (n_b, F) arrays are packed back to back into one pinned staging buffer, copied with ONE
  asynchronous H2D transfer and zero-padded / truncated to (B, 1, max_frame_num, F) by one libpa2s kernel
  (`pa2s_pad_spectrograms`); a list of device tensors (e.g. VQT outputs of clips of different length) skips the copy;
* targets: the ragged token lists become the six int64 arrays of models.py:26-31 with numpy on the host (integer work on
  a few thousand elements), in pinned memory, copied asynchronously, and carry the decoder step counts with them
  (train.targets_to_device) so that forward() needs no device read.

There is no CPU fallback for the device half: a missing libpa2s raises.
"""
from __future__ import annotations

import numpy as np
import torch

from ._lib import lib, ptr, stream
from .models import EOS, PAD


def pad_spectrograms(spectrograms, max_frame_num: int, device, truncate: bool = False) -> torch.Tensor:
    """list of (n_b, F) float arrays / tensors (host or device) -> (B, 1, max_frame_num, F) float32 on `device`.
    A clip with more than max_frame_num frames raises like the reference does (syn.py:56-57 assigns the whole spectrogram to the
    first max rows) unless `truncate=True` (keep the first max_frame_num frames)."""
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("piano_a2s_b200.batching pads on the GPU only; there is no CPU fallback")
    B = len(spectrograms)
    if B == 0:
        raise ValueError("empty batch")
    F = int(spectrograms[0].shape[-1])
    rows = [int(s.shape[0]) for s in spectrograms]
    for s in spectrograms:
        if s.ndim != 2 or int(s.shape[-1]) != F:
            raise ValueError("every spectrogram must be (frames, %d)" % F)
    if not truncate and max(rows) > max_frame_num:
        raise RuntimeError("spectrogram has %d frames, more than max_frame_num = %d" % (max(rows), max_frame_num))
    off = np.zeros(B + 1, dtype=np.int64)
    np.cumsum(rows, out=off[1:])
    total = int(off[-1])
    if all(torch.is_tensor(s) and s.is_cuda for s in spectrograms):
        packed = torch.cat([s.to(torch.float32) for s in spectrograms]) if total else torch.empty(0, F, device=device)
    else:
        stage = torch.empty((max(total, 1), F), dtype=torch.float32, pin_memory=True)
        for s, a, b in zip(spectrograms, off[:-1], off[1:]):
            if b > a:
                stage[a:b].copy_(s if torch.is_tensor(s) else torch.from_numpy(np.ascontiguousarray(s, dtype=np.float32)))
        packed = stage.to(device, non_blocking=True)
    row_off = torch.from_numpy(off).pin_memory().to(device, non_blocking=True)
    out = torch.empty((B, 1, max_frame_num, F), device=device, dtype=torch.float32)
    lib.pa2s_pad_spectrograms(stream(), ptr(packed.contiguous()), ptr(row_off), B, int(max_frame_num), F, ptr(out))
    return out


def pad_scores(scores, max_length: int, pad: int = PAD, eos: int = EOS):
    """scores[b][bar] = token list -> ((B, bars, max_length) int64, (B, bars) int64 lengths) pinned host tensors:
    tokens truncated to max_length, <eos> after them when there is room, <pad> elsewhere (syn.py:60-74)."""
    B = len(scores)
    bars = len(scores[0]) if B else 0
    tok = torch.empty((B, bars, max_length), dtype=torch.int64, pin_memory=torch.cuda.is_available())
    ln = torch.empty((B, bars), dtype=torch.int64, pin_memory=torch.cuda.is_available())
    t, l = tok.numpy(), ln.numpy()
    t[...] = pad
    for b, score in enumerate(scores):
        if len(score) != bars:
            raise ValueError("every clip must have the same number of bars")
        for k, measure in enumerate(score):
            n = min(len(measure), max_length)
            if n:
                t[b, k, :n] = np.asarray(measure[:n], dtype=np.int64)
            if n < max_length:
                t[b, k, n] = eos
            l[b, k] = n
    return tok, ln


def collate(items, max_frame_num: int, max_length, device):
    """items: [(spectrogram (n,F), time_sig ints (bars), key sharps (bars), upper bars, lower bars)] of ONE batch ->
    (spectrogram (B,1,max_frame_num,F), ground_truth list of models.py:26-31) on `device`."""
    from .train import targets_to_device
    spec = pad_spectrograms([it[0] for it in items], max_frame_num, device)
    ts = torch.from_numpy(np.asarray([it[1] for it in items], dtype=np.int64))
    key = torch.from_numpy(np.asarray([it[2] for it in items], dtype=np.int64) + 6)        # key_to_int: sharps + 6 (syn.py:38-40)
    up, ul = pad_scores([it[3] for it in items], max_length[0])
    lo, ll = pad_scores([it[4] for it in items], max_length[1])
    return spec, targets_to_device([ts, key, up, ul, lo, ll], device)
