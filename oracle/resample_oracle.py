"""CPU oracle of the audio-ingest resampler.  TEST INFRASTRUCTURE ONLY.

`resample_poly` is `scipy.signal.resample_poly` itself (float64): the published polyphase algorithm the GPU kernel
`pa2s_resample_poly` implements.  The reference resamples inside `librosa.load(sr=16000)` (utilities.py:242) with soxr 0.3.7
`soxr_hq`; soxr is neither vendored under /root/reference nor installed in this image, so parity with soxr is UNPINNED -- the pin
is against scipy's resampler (same class of filter: linear-phase Kaiser-windowed sinc low-pass)."""
import numpy as np
from scipy import signal


def resample_poly(x, sr_in, sr_out):
    """(channels, n) or (n,) float array at sr_in -> float64 at sr_out."""
    g = np.gcd(int(sr_in), int(sr_out))
    return signal.resample_poly(np.asarray(x, dtype=np.float64), int(sr_out) // g, int(sr_in) // g, axis=-1)


def write_wav_pcm16(path, x, sr):
    """(channels, n) float in [-1, 1) -> 16-bit PCM RIFF/WAVE (stdlib `wave`), returns the quantised samples as float32."""
    import wave
    q = np.clip(np.round(np.asarray(x) * 32768.0), -32768, 32767).astype("<i2")
    with wave.open(path, "wb") as w:
        w.setnchannels(q.shape[0])
        w.setsampwidth(2)
        w.setframerate(sr)
        w.writeframes(np.ascontiguousarray(q.T).tobytes())
    return q.astype(np.float32) / 32768.0
