"""CPU oracle for the VQT front end (`utilities.get_VQT`, /root/reference/utilities.py:240-254).

TEST INFRASTRUCTURE ONLY (see oracle/a2s_oracle.py for the import rules).

*** PARITY UNPINNED ***  The arithmetic of the reference front end lives in a third-party
dependency that is absent from /root/reference and from this image: `librosa==0.10.1`
(environment.yaml:52; `librosa.vqt` + `librosa.amplitude_to_db`) with `soxr==0.3.7` (:104) for
its octave-by-octave decimation.  The reference ships no golden spectrogram, no test and no
fixture for this boundary, and librosa cannot be imported here, so nothing pins this file
to real librosa output.  What it restates is librosa's *published definition* of the
transform, in float64 and in direct (time-domain) form:

  librosa/core/constantq.py::vqt   (sr=16000, hop=160, fmin=A0=27.5 Hz, n_bins=480, bpo=60, gamma=20,
                                    filter_scale=1, norm=1, window='hann', scale=True, pad_mode='constant')
  librosa/filters.py::wavelet_lengths   alpha = (2^(2/bpo)-1)/(2^(2/bpo)+1);  Q = filter_scale/alpha;
                                        N_k = Q*sr/(f_k + gamma/alpha)
  librosa/filters.py::wavelet           taps o = arange(-N_k//2, N_k//2);  b_k[o] = exp(2i*pi*f_k*o/sr) * hann_periodic(len)
                                        L1-normalised (norm=1)
  __cqt_response + stft(window='ones', center=True) + fft_basis.dot(D) with basis*N_k/n_fft and the final
  V /= sqrt(N_k) amount to     V[k,t] = sqrt(N_k) * | sum_o b_k[o] * y[t*hop - o] |        (y zero outside the clip)
  librosa/core/spectrum.py::amplitude_to_db(|V|, ref=max, amin=1e-5, top_db=80), then /80 + 1, transposed.

librosa itself evaluates the same filters octave by octave on a decimated signal and keeps only 99 % of
each filter's spectral L1 mass (sparsity=0.01).  `oracle/vqt_recursive_oracle.py` restates THAT computation
(with scipy's decimator in place of soxr) and is what `get_vqt` returns by default; measured against it
(tests/test_vqt_oracles.py) this direct form agrees to 0.2-0.5 % at spectral peaks but reads up to 0.38 (of the
0..1 scale) higher in quiet low-octave bins: without the recursion's anti-aliasing low-passes the 787-tap
low-octave filters pick up Hann side-lobe leakage of strong components at -45..-55 dB, which librosa's
low-passes remove.  The product therefore implements the recursion (composed into one filter bank).
"""
from __future__ import annotations

import numpy as np

DEFAULT_PARAMS = dict(sample_rate=16000, hop_length=160, bins_per_octave=60, n_octaves=8, gamma=20)  # pretrain.yaml:30-35
FMIN_A0 = 27.5          # librosa.note_to_hz('A0'), utilities.py:249
WINDOW = 1024           # analysis window that holds the longest filter (787 taps), centred on t*hop


def wavelet_lengths(params=DEFAULT_PARAMS):
    sr, bpo = params["sample_rate"], params["bins_per_octave"]
    n_bins = bpo * params["n_octaves"]
    freqs = FMIN_A0 * 2.0 ** (np.arange(n_bins, dtype=np.float64) / bpo)
    r = 2.0 ** (2.0 / bpo)
    alpha = (r - 1.0) / (r + 1.0)
    q = 1.0 / alpha
    lengths = q * sr / (freqs + params["gamma"] / alpha)
    return freqs, lengths


def filter_bank(params=DEFAULT_PARAMS, window=WINDOW):
    """Dense (n_bins, window) complex128 matrix G with G[k, window/2 - o] = b_k[o] * sqrt(N_k), so that
    V[k,t] = | sum_j G[k,j] * ypad[t*hop + j] |  for ypad = y zero-padded by window/2 on both sides."""
    sr = params["sample_rate"]
    freqs, lengths = wavelet_lengths(params)
    G = np.zeros((len(freqs), window), dtype=np.complex128)
    for k, (f, ilen) in enumerate(zip(freqs, lengths)):
        o = np.arange(-ilen // 2, ilen // 2, dtype=np.float64)            # filters.py::wavelet
        n = len(o)
        sig = np.exp(2j * np.pi * f * o / sr)
        sig = sig * (0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n))     # scipy get_window('hann', n, fftbins=True)
        sig = sig / np.sum(np.abs(sig))                                     # util.normalize(norm=1)
        j = (window // 2 - o).astype(np.int64)
        assert j.min() >= 0 and j.max() < window
        G[k, j] = sig * np.sqrt(ilen)
    return G


def vqt_magnitude(y, params=DEFAULT_PARAMS):
    """|V| as (n_bins, frames) float64; frames = 1 + len(y)//hop (centred)."""
    hop = params["hop_length"]
    y = np.asarray(y, dtype=np.float64)
    n_frames = 1 + len(y) // hop
    G = filter_bank(params)
    half = WINDOW // 2
    ypad = np.concatenate([np.zeros(half), y, np.zeros(half + hop)])
    idx = np.arange(n_frames)[:, None] * hop + np.arange(WINDOW)[None, :]
    frames = ypad[idx]                                                       # (T, WINDOW)
    return np.abs(frames @ G.T).T


def amplitude_to_db(mag, amin=1e-5, top_db=80.0):
    """librosa.amplitude_to_db(S, ref=np.max) for one clip."""
    ref = np.max(mag)
    log_spec = 20.0 * np.log10(np.maximum(amin, mag)) - 20.0 * np.log10(np.maximum(amin, ref))
    return np.maximum(log_spec, log_spec.max() - top_db)


def get_vqt(y, params=DEFAULT_PARAMS, algorithm=None):
    """utilities.get_VQT for an in-memory mono clip: (frames, n_bins) float32 in [0, 1].
    algorithm "librosa" (default, like the product's VQT module; env PA2S_VQT_ALGO): the octave-recursive computation librosa actually
    performs (oracle/vqt_recursive_oracle.py); "direct": this file's time-domain definition (no decimation, no sparsification)."""
    import os
    algorithm = algorithm or os.environ.get("PA2S_VQT_ALGO", "librosa")
    if algorithm == "librosa":
        from . import vqt_recursive_oracle as VR
        return VR.get_vqt(y, params)
    log_vqt = amplitude_to_db(vqt_magnitude(y, params)) / 80.0 + 1.0
    return log_vqt.T.astype(np.float32)
