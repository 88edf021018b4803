"""Audio ingest in front of the VQT: what `datasets/asap.py:80-103` does to a decoded wav before it is cut into
clips -- mono mix-down, peak normalisation, the 4..12 s duration filter -- on the GPU (`pa2s_mono_peak_normalize`), so that
decoded PCM -> VQT -> tokens never leaves the device.

Out of scope here: decoding (torchaudio.load / librosa.load) and RESAMPLING to 16 kHz (`librosa.load(sr=16000)` of
utilities.py:242 resamples with soxr 0.3.7 `soxr_hq`; neither soxr nor its filter design is available in this image, so a resampler
written here could not be shown to match it).  Feed 16 kHz PCM.
"""
from __future__ import annotations

import torch

from ._lib import lib, ptr, stream


def mono_peak_normalize(audio: torch.Tensor) -> torch.Tensor:
    """(channels, n) float32 on the GPU -> (1, n): `torch.mean(audio, 0, keepdim=True)` (only when channels > 1) followed by
    `audio / torch.max(torch.abs(audio))` (asap.py:83-86), bit-identical for mono and stereo input."""
    if not audio.is_cuda:
        raise RuntimeError("piano_a2s_b200.audio runs on CUDA tensors only; there is no CPU fallback")
    if audio.dtype != torch.float32 or audio.ndim != 2:
        raise TypeError("audio must be a (channels, samples) float32 tensor")
    a = audio.contiguous()
    C, n = a.shape
    out = torch.empty((1, n), device=a.device, dtype=torch.float32)
    scratch = torch.empty(1, device=a.device, dtype=torch.int32)
    lib.pa2s_mono_peak_normalize(stream(), ptr(a), C, n, ptr(out), ptr(scratch))
    return out


def keep_clip(n_samples: int, sample_rate: int, min_s: float = 4.0, max_s: float = 12.0) -> bool:
    """The duration filter of asap.py:100-102: a clip is dropped when it is longer than 12 s or shorter than 4 s."""
    return not (n_samples > max_s * sample_rate or n_samples < min_s * sample_rate)


def cut_clips(audio: torch.Tensor, sample_rate: int, bounds_s):
    """audio (1, n) -> [(start, stop) clips that pass the duration filter]: `audio[:, int(t0*sr): int(t1*sr)]` (asap.py:98-103)."""
    out = []
    for t0, t1 in bounds_s:
        clip = audio[:, int(t0 * sample_rate): int(t1 * sample_rate)]
        if keep_clip(clip.shape[1], sample_rate):
            out.append(clip)
    return out
