"""Batched on-GPU variable-Q transform front end: drop-in for `utilities.get_VQT` (utilities.py:240-254).

The reference calls `librosa.vqt(y, sr=16000, hop_length=160, fmin=A0, n_bins=480, bins_per_octave=60, gamma=20)`
on the CPU, one clip at a time, offline.  Here the transform is evaluated in its direct (time-domain) form as ONE
strided complex filterbank contraction over overlapping frames of the zero-padded clip
(row stride = hop, so the frame matrix is never materialised), followed by a fused |.| / per-clip max / dB / scale
epilogue.  Filter design (host, float64, once): librosa's wavelet definition -- see DESIGN.md "VQT".
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import ops
from ._lib import lib, ptr, stream

FMIN_A0 = 27.5     # librosa.note_to_hz('A0')
WINDOW = 1024      # analysis window holding the longest filter (787 taps), centred on t*hop


def design_filters(sample_rate=16000, bins_per_octave=60, n_octaves=8, gamma=20, window=WINDOW):
    """-> (2*n_bins, K) float32 rows [re_0, im_0, re_1, ...] restricted to the K-column support, and the support offset.

    b_k[o] = exp(2i*pi*f_k*o/sr) * hann_periodic(n)[o - o_min], o = arange(-N_k//2, N_k//2), L1-normalised, times
    sqrt(N_k); N_k = Q*sr/(f_k + gamma/alpha), alpha = (2^(2/bpo)-1)/(2^(2/bpo)+1), Q = 1/alpha
    (librosa filters.wavelet / wavelet_lengths with filter_scale=1, norm=1, window='hann'; vqt(scale=True))."""
    n_bins = bins_per_octave * n_octaves
    freqs = FMIN_A0 * 2.0 ** (np.arange(n_bins, dtype=np.float64) / bins_per_octave)
    r = 2.0 ** (2.0 / bins_per_octave)
    alpha = (r - 1.0) / (r + 1.0)
    lengths = (1.0 / alpha) * sample_rate / (freqs + gamma / alpha)
    G = np.zeros((n_bins, window), dtype=np.complex128)
    for k in range(n_bins):
        o = np.arange(-lengths[k] // 2, lengths[k] // 2, dtype=np.float64)
        n = len(o)
        sig = np.exp(2j * np.pi * freqs[k] * o / sample_rate) * (0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n))
        sig /= np.abs(sig).sum()
        j = (window // 2 - o).astype(np.int64)
        if j.min() < 0 or j.max() >= window:
            raise ValueError("filter longer than the analysis window")
        G[k, j] = sig * np.sqrt(lengths[k])
    nz = np.nonzero(np.abs(G).sum(0))[0]
    j0 = (nz.min() // 8) * 8                               # keep 16-byte alignment of the (fp32 and bf16) frame rows
    j1 = ((nz.max() + 1 + 15) // 16) * 16
    j1 = min(max(j1, j0 + 16), window)
    W = np.empty((2 * n_bins, j1 - j0), dtype=np.float32)
    W[0::2] = G.real[:, j0:j1]
    W[1::2] = G.imag[:, j0:j1]
    return W, int(j0)


def _resample2_filter():
    """The decimate-by-2 low-pass of the octave recursion: scipy.signal.resample_poly(x, 1, 2)'s FIR (firwin(41, 0.5, kaiser 5.0)),
    restated in numpy.  librosa decimates with soxr_hq (soxr is not available here): same class of filter, not the same taps."""
    half = 20
    n = np.arange(-half, half + 1, dtype=np.float64)
    h = 0.5 * np.sinc(0.5 * n) * np.kaiser(2 * half + 1, 5.0)
    return h / h.sum(), half


def design_filters_librosa(sample_rate=16000, hop_length=160, bins_per_octave=60, n_octaves=8, gamma=20, sparsity=0.01, tail_tol=1e-9):
    """The filter bank librosa.vqt EFFECTIVELY applies, as one full-rate direct-form bank -> (W (2*n_bins, K) float32 rows
    [re_0, im_0, ...], p_min) with  V[k,t] = | sum_j W[k,j] * y[t*hop + p_min + j] |  (y zero outside the clip).

    librosa 0.10.1 evaluates the transform octave by octave [recalled: core/constantq.py vqt / __vqt_filter_fft / __cqt_response]:
    the octave's wavelets at the current rate, zero-padded to n_fft, L1-normalised, * N_k / n_fft, FFT'd (positive half), every row
    sparsified (entries holding the smallest 1 % of the row's L1 mass dropped), * sqrt(sr / rate); response = basis . STFT(window =
    ones, centred); then y is decimated by 2 (* sqrt(2)) and the hop halved while it is even (5 times for hop 160), and finally
    V /= sqrt(N_k).  Every step is linear in y and the frame positions t*hop are multiples of the total decimation factor, so the
    response at a frame is a fixed FIR of y: the octave's time-domain kernel (inverse transform of the sparsified basis) upsampled by
    2^d and convolved with the cascade of the d decimation low-passes.  Composing them here, in float64, turns the whole recursion
    into ONE filterbank contraction (K = 2 304 taps for the lowest octaves) that the tensor-core GEMM evaluates in a single pass --
    including the anti-aliasing of the low octaves that the plain time-domain definition (design_filters) lacks."""
    n_bins = bins_per_octave * n_octaves
    freqs = FMIN_A0 * 2.0 ** (np.arange(n_bins, dtype=np.float64) / bins_per_octave)
    r = 2.0 ** (2.0 / bins_per_octave)
    alpha = (r - 1.0) / (r + 1.0)
    lengths = (1.0 / alpha) * sample_rate / (freqs + gamma / alpha)
    h2, half2 = _resample2_filter()
    kernels = []                                                            # per bin: (p_first, complex taps)
    H, c, d, hop = np.ones(1), 0, 0, hop_length                             # cascade low-pass, its centre, stages, current hop
    for i in range(n_octaves):
        rate = sample_rate / 2.0 ** d
        sl = slice(n_bins - bins_per_octave * (i + 1), n_bins - bins_per_octave * i)
        f_oct, len_oct = freqs[sl], lengths[sl] * (rate / sample_rate)
        n_fft = int(2.0 ** np.ceil(np.log2(len_oct.max())))
        n_fft = max(n_fft, int(2.0 ** (1 + np.ceil(np.log2(hop)))))
        basis = np.zeros((bins_per_octave, n_fft), dtype=np.complex128)
        for j, (f, ilen) in enumerate(zip(f_oct, len_oct)):
            o = np.arange(-ilen // 2, ilen // 2, dtype=np.float64)
            n = len(o)
            sig = np.exp(2j * np.pi * f * o / rate) * (0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n))
            sig /= np.abs(sig).sum()
            lpad = (n_fft - n) // 2
            basis[j, lpad:lpad + n] = sig * (ilen / float(n_fft))
        fb = np.fft.fft(basis, axis=1)[:, : n_fft // 2 + 1]
        if sparsity > 0:                                                    # util.sparsify_rows(quantile=sparsity)
            mags = np.abs(fb)
            srt = np.sort(mags, axis=1)
            cum = np.cumsum(srt / mags.sum(axis=1, keepdims=True), axis=1)
            thr = srt[np.arange(len(srt)), np.argmin(cum < sparsity, axis=1)]
            fb = np.where(mags >= thr[:, None], fb, 0.0)
        fb = fb * np.sqrt(sample_rate / rate)
        # time-domain kernel of `fb . rfft(frame)`: c[n] = sum_f fb[f] exp(-2 pi i f n / n_fft), frame sample n <-> y_d[t*hop_d + n - n_fft/2]
        nn = np.arange(n_fft)
        ck = fb @ np.exp(-2j * np.pi * np.outer(np.arange(n_fft // 2 + 1), nn) / n_fft)
        scale = 2.0 ** (d / 2.0)                                            # sqrt(2) per decimation (resample(scale=True))
        step = 2 ** d
        for j in range(bins_per_octave):
            up = np.zeros((n_fft - 1) * step + 1, dtype=np.complex128)
            up[::step] = ck[j]
            g = np.convolve(up, H) * scale / np.sqrt(lengths[sl][j])        # taps over p = step*(n - n_fft/2) + c - q, q index of H
            # np.convolve index m = step*n + q'  with q' = index into H (symmetric: H[q] = H[len-1-q]);  p = step*n - step*n_fft/2 + c - q
            # taking q = len(H)-1-q':  p = m - (len(H)-1) - step*n_fft/2 + c
            p_first = -(len(H) - 1) - step * n_fft // 2 + c
            kernels.append((sl.start + j, p_first, g))
        if hop % 2 == 0:
            hop //= 2
            H = np.convolve(H, np.repeat(h2, 1) if d == 0 else np.kron(h2, np.eye(1, 2 ** d)[0])[: (len(h2) - 1) * 2 ** d + 1])
            c += half2 * 2 ** d
            d += 1
    p_lo = min(p for _, p, _ in kernels)
    p_hi = max(p + len(g) for _, p, g in kernels)
    G = np.zeros((n_bins, p_hi - p_lo), dtype=np.complex128)
    for k, p, g in kernels:
        G[k, p - p_lo:p - p_lo + len(g)] = g
    # the cascade's far tails are ~1e-10 of a filter's L1 mass: drop leading / trailing taps that hold < tail_tol of EVERY row's mass
    mass = np.abs(G) / np.abs(G).sum(axis=1, keepdims=True)
    lead = np.cumsum(mass, axis=1).max(axis=0)
    trail = np.cumsum(mass[:, ::-1], axis=1).max(axis=0)[::-1]
    lo = int(np.argmax(lead >= tail_tol))
    hi = len(trail) - int(np.argmax(trail[::-1] >= tail_tol))
    p_min = -((-(p_lo + lo) + 7) // 8 * 8)                                  # 16-byte alignment of the bf16 / fp32 frame rows
    K = ((p_lo + hi - p_min + 63) // 64) * 64                               # whole k-blocks of the tensor-core GEMM
    W = np.zeros((2 * n_bins, K), dtype=np.float32)
    a, b = max(p_min, p_lo), min(p_min + K, p_hi)
    W[0::2, a - p_min:b - p_min] = G.real[:, a - p_lo:b - p_lo]
    W[1::2, a - p_min:b - p_min] = G.imag[:, a - p_lo:b - p_lo]
    return W, int(p_min)


class VQT(torch.nn.Module):
    """audio (B, n_samples) float32 on CUDA -> (B, 1 + n_samples//hop, n_bins) float32 in [0, 1].
    algorithm = "librosa" (default): the filter bank librosa.vqt effectively applies (octave recursion with decimation and basis
    sparsification composed into one direct-form bank, design_filters_librosa); "direct": the transform's plain time-domain definition
    (design_filters; no anti-aliasing of the low octaves, window 1024)."""

    def __init__(self, sample_rate=16000, hop_length=160, bins_per_octave=60, n_octaves=8, gamma=20, algorithm=None):
        super().__init__()
        self.hop = hop_length
        self.n_bins = bins_per_octave * n_octaves
        self.algorithm = algorithm or os.environ.get("PA2S_VQT_ALGO", "librosa")
        if self.algorithm == "librosa":
            W, p_min = design_filters_librosa(sample_rate, hop_length, bins_per_octave, n_octaves, gamma)
            self.left = -p_min                                             # zeros in front of the clip; frame t starts at t*hop
            self.j0 = 0
        else:
            W, j0 = design_filters(sample_rate, bins_per_octave, n_octaves, gamma)
            self.left, self.j0 = WINDOW // 2, j0
        self.register_buffer("filters", torch.from_numpy(W), persistent=False)
        # bins 80 dB below the clip maximum must keep ~1e-3 relative accuracy through the dB epilogue, which a bf16x3
        # product (error ~5e-6 of the *dominant* terms) cannot give: audio and filters are split into THREE bf16 pieces
        # and the six leading piece products are accumulated in TMEM ("bf16x6", ~2^-24: fp32-level accuracy on tcgen05).
        self.precision = os.environ.get("PA2S_VQT_PRECISION", "bf16x6")
        self._fop = None

    @torch.no_grad()
    def forward(self, audio, n_samples=None):
        """`n_samples` (B,) int tensor on the device, or None: true length of each zero-padded clip.  Frames a clip does not have
        (t >= 1 + n_b // hop) come out as zeros and stay out of the clip maximum, i.e. the result equals get_VQT of the un-padded clip
        followed by pad_spectrogram (datasets/asap.py:345-349, 383)."""
        if not audio.is_cuda:
            raise RuntimeError("VQT runs on CUDA only")
        audio = audio.float()
        B, n = audio.shape
        T = 1 + n // self.hop
        K = self.filters.shape[1]
        half = self.left
        plen = ((half + n + K + self.hop + 7) // 8) * 8
        ypad = torch.zeros(B, plen, device=audio.device, dtype=torch.float32)
        ypad[:, half:half + n] = audio
        valid = None
        if n_samples is not None:
            valid = (1 + torch.div(n_samples.to(audio.device), self.hop, rounding_mode="floor")).to(torch.int32).contiguous()
        out = torch.empty(B, T, self.n_bins, device=audio.device, dtype=torch.float32)
        cmax = torch.zeros(B, device=audio.device, dtype=torch.int32)
        # frames[t, j] = ypad[t*hop + j0 + j]: overlapping rows, lda = hop
        if self.precision == "fp32":
            C = torch.empty(B, T, 2 * self.n_bins, device=audio.device, dtype=torch.float32)
            with ops.ktime("vqt_filterbank"):
                ops.gemm(ypad, self.filters, C, T, 2 * self.n_bins, K, transB=True, lda=self.hop, ldb=K, ldc=2 * self.n_bins,
                         batch=B, strideA=plen, strideB=0, strideC=T * 2 * self.n_bins, a_off=self.j0, precision="fp32")
            lib.pa2s_vqt_post(stream(), ptr(C), ptr(out), ptr(cmax), B, T, self.n_bins, ptr(valid))
            return out
        # tensor-core path: the filter bank is constant (split into bf16 pieces once); |.| and the per-clip maximum are the EPILOGUE of
        # the contraction (the (B, T, 960) complex responses never reach HBM), the dB / scale pass then works in place on the magnitudes
        npc = ops.npieces_for(self.precision)
        if self._fop is None or self._fop[0] != (self.precision, self.filters.device):
            self._fop = ((self.precision, self.filters.device), ops.split_operand(self.filters, 2 * self.n_bins, K, K, npieces=npc))
        Bop = self._fop[1]
        none = dict(t_scale=None, t_shift=None, t_period=1, t_relu=False)
        with ops.ktime("vqt_filterbank"):
            Aop, a_mn, a_ld = ops._operand(ypad, False, T, K, self.hop, self.j0, B, plen, npc, none)
            lib.pa2s_gemm_bf16_tma_mag(stream(), T, 2 * self.n_bins, K,
                                       ptr(Aop.buf), a_ld or Aop.ld, Aop.piece_stride, Aop.batch_stride, Aop.npieces, int(a_mn),
                                       ptr(Bop.buf), Bop.ld, Bop.piece_stride, Bop.batch_stride, Bop.npieces, 0,
                                       ptr(out), T * self.n_bins, ptr(cmax), ptr(valid), B)
        lib.pa2s_vqt_logscale(stream(), ptr(out), ptr(cmax), B, T, self.n_bins, ptr(valid))
        return out


_CACHE = {}


def get_VQT(audio_or_path, hparams):
    """Same signature as utilities.get_VQT: a 16 kHz mono array or the path of a .wav file -> (frames, n_bins) float32 numpy."""
    key = (hparams["sample_rate"], hparams["hop_length"], hparams["bins_per_octave"], hparams["n_octaves"], hparams["gamma"])
    if key not in _CACHE:
        _CACHE[key] = VQT(*key).cuda()
    if isinstance(audio_or_path, (str, os.PathLike)):
        # `librosa.load(audio, sr=hparams['sample_rate'])` of utilities.py:241-242: WAVE decode + mono + resampling (audio.load)
        from .audio import load
        y = load(os.fspath(audio_or_path), sr=hparams["sample_rate"]).reshape(1, -1)
    else:
        y = torch.as_tensor(np.asarray(audio_or_path), dtype=torch.float32).cuda().reshape(1, -1)
    return _CACHE[key](y)[0].cpu().numpy()
