"""Audio ingest in front of the VQT: what `datasets/asap.py:80-103` does to a decoded wav before it is cut into
clips -- mono mix-down, peak normalisation, the 4..12 s duration filter -- on the GPU (`pa2s_mono_peak_normalize`), so that
decoded PCM -> VQT -> tokens never leaves the device.

`load(path, sr=16000)` is the `librosa.load(audio, sr=16000)` of utilities.py:241-242: RIFF/WAVE decoding on the host (PCM 8/16/24/32
bit and IEEE float; the only container the reference's pipeline produces: FluidSynth / ffmpeg write .wav, render.py:470-489), mono
mix-down and rational resampling to the target rate on the GPU (`pa2s_resample_poly`: polyphase Kaiser-windowed sinc, the filter of
`scipy.signal.resample_poly`, against which it is pinned).  librosa resamples with soxr 0.3.7 `soxr_hq`, which is neither vendored nor
installed here: parity with soxr is UNPINNED (both are linear-phase low-pass resamplers with > 100 dB stop-band, so the spectrograms
agree to the filter's transition band, but not bit for bit).
"""
from __future__ import annotations

import math
import struct

import numpy as np
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


def read_wav(path):
    """RIFF/WAVE file -> ((channels, n) float32 numpy in [-1, 1), sample_rate).  PCM 8 (unsigned) / 16 / 24 / 32 bit, IEEE float 32 / 64,
    incl. WAVE_FORMAT_EXTENSIBLE; scaling like soundfile / librosa (int / 2^(bits-1))."""
    with open(path, "rb") as f:
        data = f.read()
    if len(data) < 12 or data[:4] != b"RIFF" or data[8:12] != b"WAVE":
        raise ValueError(f"{path}: not a RIFF/WAVE file")
    pos, fmt, pcm = 12, None, None
    while pos + 8 <= len(data):
        cid, size = data[pos:pos + 4], struct.unpack("<I", data[pos + 4:pos + 8])[0]
        body = data[pos + 8:pos + 8 + size]
        if cid == b"fmt ":
            tag, ch, sr, _, _, bits = struct.unpack("<HHIIHH", body[:16])
            if tag == 0xFFFE and len(body) >= 26:                      # WAVE_FORMAT_EXTENSIBLE: the sub-format GUID starts with the tag
                tag = struct.unpack("<H", body[24:26])[0]
            fmt = (tag, ch, sr, bits)
        elif cid == b"data":
            pcm = body
        pos += 8 + size + (size & 1)
    if fmt is None or pcm is None:
        raise ValueError(f"{path}: missing fmt / data chunk")
    tag, ch, sr, bits = fmt
    if tag == 1 and bits == 8:
        x = (np.frombuffer(pcm, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
    elif tag == 1 and bits == 16:
        x = np.frombuffer(pcm, dtype="<i2").astype(np.float32) / 32768.0
    elif tag == 1 and bits == 24:
        b = np.frombuffer(pcm[:len(pcm) // 3 * 3], dtype=np.uint8).reshape(-1, 3).astype(np.int32)
        v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
        x = (v - ((v & 0x800000) << 1)).astype(np.float32) / 8388608.0
    elif tag == 1 and bits == 32:
        x = (np.frombuffer(pcm, dtype="<i4").astype(np.float64) / 2147483648.0).astype(np.float32)
    elif tag == 3 and bits in (32, 64):
        x = np.frombuffer(pcm, dtype="<f4" if bits == 32 else "<f8").astype(np.float32)
    else:
        raise ValueError(f"{path}: unsupported WAVE format tag {tag} with {bits} bits")
    n = len(x) // ch
    return np.ascontiguousarray(x[:n * ch].reshape(n, ch).T), int(sr)


_FILTERS = {}


def resample_filter(up: int, down: int, device):
    """The FIR of scipy.signal.resample_poly(x, up, down): firwin(2*10*max(up,down)+1, 1/max(up,down), window=('kaiser', 5.0)) * up,
    designed in float64 on the host (restated here: numpy only), cached per (up, down, device)."""
    key = (up, down, str(device))
    if key not in _FILTERS:
        mx = max(up, down)
        half = 10 * mx
        n = np.arange(-half, half + 1, dtype=np.float64)
        fc = 1.0 / mx
        h = fc * np.sinc(fc * n) * np.kaiser(2 * half + 1, 5.0)
        h /= h.sum()                                                  # firwin scales the pass band to unit gain at DC
        _FILTERS[key] = torch.from_numpy((h * up).astype(np.float32)).to(device)
    return _FILTERS[key]


def resample(audio: torch.Tensor, sr_in: int, sr_out: int) -> torch.Tensor:
    """(channels, n) float32 on the GPU at sr_in -> (channels, ceil(n * sr_out / sr_in)) at sr_out."""
    if not audio.is_cuda:
        raise RuntimeError("piano_a2s_b200.audio runs on CUDA tensors only; there is no CPU fallback")
    if sr_in == sr_out:
        return audio
    g = math.gcd(int(sr_in), int(sr_out))
    up, down = int(sr_out) // g, int(sr_in) // g
    a = audio.contiguous().float()
    C, n = a.shape
    n_out = -(-n * up // down)
    h = resample_filter(up, down, a.device)
    out = torch.empty((C, n_out), device=a.device, dtype=torch.float32)
    lib.pa2s_resample_poly(stream(), ptr(a), C, n, up, down, ptr(h), h.numel(), ptr(out), n_out)
    return out


def load(path, sr: int = 16000, mono: bool = True, device="cuda") -> torch.Tensor:
    """`librosa.load(path, sr=sr)` (utilities.py:242) -> (n,) float32 on the GPU (mono) or (channels, n): decode on the host, one H2D copy,
    mix-down (mean over channels, like librosa.to_mono) and resampling on the device."""
    x, sr_in = read_wav(path)
    a = torch.from_numpy(x).to(device)
    if mono and a.shape[0] > 1:
        a = a.mean(0, keepdim=True)
    a = resample(a, sr_in, sr)
    return a[0] if mono else a
