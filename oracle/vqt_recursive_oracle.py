"""Second, independent CPU evaluation of the VQT front end: librosa's OWN algorithm (octave-recursive, FFT-domain, sparsified).

TEST INFRASTRUCTURE ONLY.  `oracle/vqt_oracle.py` restates the transform's definition in direct (time-domain) form; the product
filter design (`piano_a2s_b200/vqt.py`) follows the same definition, so a comparison of those two is partly self-referential.  This
file restates how librosa 0.10.1 actually COMPUTES `librosa.vqt` [recalled from librosa/core/constantq.py: `vqt`,
`__vqt_filter_fft`, `__cqt_response`, `__trim_stack`; librosa/filters.py: `wavelet(pad_fft=True)`; util.sparsify_rows]:

  for each of the 8 octaves, top one first:  wavelets of the octave at the CURRENT sample rate, zero-padded to n_fft = next power of
  two, L1-normalised, scaled by N_k / n_fft, FFT'd (positive half), each row SPARSIFIED (the smallest-magnitude entries holding 1 %
  of the row's L1 mass are zeroed, sparsity=0.01), scaled by sqrt(sr / my_sr); response = basis @ STFT(y, n_fft, hop, window='ones',
  center=True, pad_mode='constant'); then, while the hop is even, y is decimated by 2 (librosa: soxr_hq resampler with scale=True,
  i.e. * sqrt(2)), hop //= 2, my_sr /= 2.  Finally the octave stacks are trimmed to the common frame count and V /= sqrt(N_k).

The one substitution: soxr (absent from this image) is replaced by scipy.signal.resample_poly for the decimation by 2.  What this
oracle is for: it shares NO code path with the direct form (multi-rate, frequency domain, sparsified), so the difference between the
two bounds how far librosa's approximations sit from the transform's definition -- the "~1e-3 after the /80 dB scaling" claim of
DESIGN.md -- and tests/test_vqt_oracles.py asserts that bound.  It does not pin either file to real librosa output (parity with
librosa itself stays unpinned: librosa cannot be imported here).
"""
from __future__ import annotations

import numpy as np
from scipy import signal

from . import vqt_oracle as VO


def _wavelets(freqs, lengths, sr):
    """filters.wavelet(..., pad_fft=True, norm=1, window='hann'): (n, n_fft) complex, centred, L1-normalised."""
    max_len = int(2.0 ** np.ceil(np.log2(lengths.max())))
    out = np.zeros((len(freqs), max_len), dtype=np.complex128)
    for i, (f, ilen) in enumerate(zip(freqs, lengths)):
        o = np.arange(-ilen // 2, ilen // 2, dtype=np.float64)
        n = len(o)
        sig = np.exp(2j * np.pi * f * o / sr) * (0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n))
        sig = sig / np.sum(np.abs(sig))
        lpad = (max_len - n) // 2                                             # util.pad_center
        out[i, lpad:lpad + n] = sig
    return out, max_len


def _sparsify_rows(x, quantile=0.01):
    """util.sparsify_rows: zero the smallest-magnitude entries of each row whose cumulative L1 mass is < quantile of the row's."""
    out = np.zeros_like(x)
    mags = np.abs(x)
    norms = mags.sum(axis=1, keepdims=True)
    srt = np.sort(mags, axis=1)
    cum = np.cumsum(srt / norms, axis=1)
    idx = np.argmin(cum < quantile, axis=1)
    for i, j in enumerate(idx):
        keep = mags[i] >= srt[i, j]
        out[i, keep] = x[i, keep]
    return out


def _stft_ones(y, n_fft, hop):
    """librosa.stft(y, n_fft, hop_length=hop, window='ones', center=True, pad_mode='constant') -> (1 + n_fft//2, frames)."""
    ypad = np.concatenate([np.zeros(n_fft // 2), y, np.zeros(n_fft // 2)])
    n_frames = 1 + (len(ypad) - n_fft) // hop
    idx = np.arange(n_frames)[None, :] * hop + np.arange(n_fft)[:, None]
    return np.fft.rfft(ypad[idx], axis=0)


EXTEND_FRAMES = 32        # 32 frames = 5 120 samples = a multiple of hop * 2^5: padding by it keeps every octave's frame grid aligned


def vqt_magnitude(y, params=VO.DEFAULT_PARAMS, sparsity=0.01, extend=True):
    """|V| (n_bins, frames).  extend=True (what the product is tested against): the clip is treated as an infinite signal that is zero
    outside [0, n) -- it is padded with EXTEND_FRAMES frames of zeros on both sides before the recursion and the frames are trimmed
    afterwards, so the decimated signals keep the filter tails that spill over the clip's ends.  extend=False: the literal recursion,
    where each decimation TRUNCATES its output to ceil(n/2) samples (what `audio.resample` returns); the two differ only in the first
    and last ~5 frames, which is also where librosa's real resampler (soxr) has an edge convention of its own."""
    if extend:
        hop = params["hop_length"]
        pad = EXTEND_FRAMES * hop
        n_frames = 1 + len(y) // hop
        yy = np.concatenate([np.zeros(pad), np.asarray(y, dtype=np.float64), np.zeros(pad)])
        return vqt_magnitude(yy, params, sparsity, extend=False)[:, EXTEND_FRAMES:EXTEND_FRAMES + n_frames]
    sr, hop, bpo, n_oct = params["sample_rate"], params["hop_length"], params["bins_per_octave"], params["n_octaves"]
    n_bins = bpo * n_oct
    freqs, lengths = VO.wavelet_lengths(params)                               # lengths at the ORIGINAL rate
    y = np.asarray(y, dtype=np.float64)
    my_y, my_sr, my_hop = y, float(sr), hop
    resp = []
    for i in range(n_oct):
        sl = slice(n_bins - bpo * (i + 1), n_bins - bpo * i)
        f_oct = freqs[sl]
        len_oct = lengths[sl] * (my_sr / sr)                                  # the same filters at the current rate
        basis, n_fft = _wavelets(f_oct, len_oct, my_sr)
        if n_fft < 2.0 ** (1 + np.ceil(np.log2(my_hop))):
            n_fft = int(2.0 ** (1 + np.ceil(np.log2(my_hop))))
            b2 = np.zeros((basis.shape[0], n_fft), dtype=np.complex128)
            lpad = (n_fft - basis.shape[1]) // 2
            b2[:, lpad:lpad + basis.shape[1]] = basis
            basis = b2
        basis = basis * (len_oct[:, None] / float(n_fft))
        fft_basis = np.fft.fft(basis, n=n_fft, axis=1)[:, : n_fft // 2 + 1]
        if sparsity > 0:
            fft_basis = _sparsify_rows(fft_basis, sparsity)
        fft_basis = fft_basis * np.sqrt(sr / my_sr)
        resp.append(fft_basis @ _stft_ones(my_y, n_fft, my_hop))
        if my_hop % 2 == 0:
            my_hop //= 2
            my_sr /= 2.0
            my_y = signal.resample_poly(my_y, 1, 2) * np.sqrt(2.0)             # audio.resample(..., scale=True): /= sqrt(ratio)
    n_frames = min(r.shape[1] for r in resp)
    V = np.concatenate([r[:, :n_frames] for r in resp[::-1]], axis=0)         # __trim_stack: lowest octave first
    V = V / np.sqrt(lengths[:, None])
    return np.abs(V)


def get_vqt(y, params=VO.DEFAULT_PARAMS, sparsity=0.01, extend=True):
    log_vqt = VO.amplitude_to_db(vqt_magnitude(y, params, sparsity, extend)) / 80.0 + 1.0
    return log_vqt.T.astype(np.float32)
