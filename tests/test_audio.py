"""Audio ingest: the two arithmetic lines of datasets/asap.py:83-86 are plain torch ops, so torch on the CPU IS the
reference for them; the device kernel has to match bit for bit (mono / stereo)."""
import numpy as np
import pytest
import torch


def _reference(audio):
    """asap.py:82-86, verbatim arithmetic."""
    if audio.shape[0] > 1:
        audio = torch.mean(audio, dim=0, keepdim=True)
    return audio / torch.max(torch.abs(audio))


def test_duration_filter():
    from piano_a2s_b200.audio import keep_clip
    sr = 16000
    assert keep_clip(4 * sr, sr) and keep_clip(12 * sr, sr) and keep_clip(8 * sr, sr)
    assert not keep_clip(4 * sr - 1, sr) and not keep_clip(12 * sr + 1, sr)


def test_audio_refuses_cpu():
    from piano_a2s_b200.audio import mono_peak_normalize
    with pytest.raises(RuntimeError):
        mono_peak_normalize(torch.zeros(2, 8))


@pytest.mark.gpu
@pytest.mark.parametrize("C,n", [(1, 192000), (2, 192000), (2, 1), (2, 529201), (1, 7)])
def test_mono_peak_normalize_bit_exact(cuda, C, n):
    from piano_a2s_b200.audio import mono_peak_normalize
    g = torch.Generator().manual_seed(C * 1000 + n % 977)
    a = torch.clamp(0.3 * torch.randn(C, n, generator=g), -1, 1)
    want = _reference(a)
    got = mono_peak_normalize(a.to(cuda))
    assert got.shape == (1, n) and torch.equal(got.cpu(), want)
    assert float(got.abs().max()) == 1.0


@pytest.mark.gpu
def test_mono_peak_normalize_edge_cases(cuda):
    from piano_a2s_b200.audio import cut_clips, mono_peak_normalize
    z = mono_peak_normalize(torch.zeros(2, 64, device=cuda))            # silence: 0/0 = NaN in the reference too
    assert torch.isnan(z).all() and torch.isnan(_reference(torch.zeros(2, 64))).all()
    a = torch.randn(4, 3001)                                            # > 2 channels: the mean's rounding may differ in the last bit
    got = mono_peak_normalize(a.to(cuda)).cpu()
    assert torch.allclose(got, _reference(a), rtol=3e-7, atol=0)
    x = mono_peak_normalize(torch.randn(1, 16000 * 30).to(cuda))
    clips = cut_clips(x, 16000, [(0.0, 3.0), (3.0, 9.5), (9.5, 22.0), (18.0, 30.0)])
    assert [c.shape[1] for c in clips] == [int(6.5 * 16000), 12 * 16000]


# ------------------------------------------------------------------------------------------------ decode + resample (N1)
def test_wav_decoder_formats(tmp_path):
    """read_wav: 16-bit PCM written by the stdlib `wave` module, and hand-built 8 / 24 / 32-bit PCM and float32 files."""
    import struct
    from oracle import resample_oracle as R
    from piano_a2s_b200 import audio
    rng = np.random.default_rng(3)
    x = rng.uniform(-0.9, 0.9, (2, 257))
    p = str(tmp_path / "pcm16.wav")
    q = R.write_wav_pcm16(p, x, 44100)
    y, sr = audio.read_wav(p)
    assert sr == 44100 and y.shape == (2, 257) and np.array_equal(y, q)

    def riff(tag, bits, payload, ch=2, sr=22050):
        fmt = struct.pack("<HHIIHH", tag, ch, sr, sr * ch * bits // 8, ch * bits // 8, bits)
        body = b"WAVE" + b"fmt " + struct.pack("<I", len(fmt)) + fmt + b"LIST" + struct.pack("<I", 4) + b"abcd" + b"data" + struct.pack("<I", len(payload)) + payload
        return b"RIFF" + struct.pack("<I", len(body)) + body
    inter = np.ascontiguousarray(x.T)                                             # (n, ch) interleaved
    cases = {
        "u8": (1, 8, (np.round(inter * 128) + 128).clip(0, 255).astype(np.uint8).tobytes(), lambda: (np.round(inter * 128).clip(-128, 127)) / 128),
        "i32": (1, 32, np.round(inter * 2 ** 31).astype("<i4").tobytes(), lambda: np.round(inter * 2 ** 31) / 2 ** 31),
        "f32": (3, 32, inter.astype("<f4").tobytes(), lambda: inter.astype(np.float32)),
    }
    i24 = np.round(inter * 2 ** 23).astype(np.int32)
    cases["i24"] = (1, 24, b"".join(int(v).to_bytes(3, "little", signed=True) for v in i24.reshape(-1)), lambda: i24 / 2 ** 23)
    for name, (tag, bits, payload, want) in cases.items():
        p = str(tmp_path / f"{name}.wav")
        open(p, "wb").write(riff(tag, bits, payload))
        y, sr = audio.read_wav(p)
        assert sr == 22050 and y.shape == (2, 257), name
        assert np.allclose(y, want().T.astype(np.float32), atol=1e-7), name


@pytest.mark.gpu
@pytest.mark.parametrize("sr_in,n", [(44100, 30011), (48000, 16001), (22050, 9999), (8000, 4000), (16000, 5000)])
def test_resampler_matches_scipy_resample_poly(cuda, sr_in, n):
    """pa2s_resample_poly against scipy.signal.resample_poly (float64) -- the published algorithm it implements; soxr parity is unpinned."""
    from oracle import resample_oracle as R
    from piano_a2s_b200 import audio
    x = np.random.default_rng(sr_in).uniform(-1, 1, (2, n)).astype(np.float32)
    y = audio.resample(torch.from_numpy(x).to(cuda), sr_in, 16000).cpu().numpy()
    ref = R.resample_poly(x, sr_in, 16000)
    assert y.shape == ref.shape
    assert np.abs(y - ref).max() < 2e-6


@pytest.mark.gpu
def test_load_and_get_vqt_from_wav_path(cuda, tmp_path):
    """`get_VQT(path, hparams)` == utilities.get_VQT's path branch: decode + mono + resample (audio.load) in front of the VQT."""
    from oracle import resample_oracle as R
    from oracle import vqt_oracle as VO
    from piano_a2s_b200 import audio
    from piano_a2s_b200.vqt import get_VQT
    rng = np.random.default_rng(11)
    t = np.arange(44100) / 44100.0
    x = np.stack([0.4 * np.sin(2 * np.pi * 440 * t) + 0.05 * rng.standard_normal(t.size), 0.3 * np.sin(2 * np.pi * 660 * t)])
    p = str(tmp_path / "clip.wav")
    q = R.write_wav_pcm16(p, x, 44100)
    y = audio.load(p, sr=16000).cpu().numpy()
    ref = R.resample_poly(q.astype(np.float64).mean(0), 44100, 16000)
    assert y.shape == ref.shape and np.abs(y - ref).max() < 2e-6
    hp = dict(sample_rate=16000, hop_length=160, bins_per_octave=60, n_octaves=8, gamma=20)
    got = get_VQT(p, hp)
    want = VO.get_vqt(ref.astype(np.float32))
    assert got.shape == want.shape and np.abs(got - want).max() < 2e-4


@pytest.mark.gpu
def test_vqt_of_zero_padded_ragged_clips_equals_per_clip_vqt(cuda):
    """BASELINE configs[3]: VQT(zero-padded batch, n_samples) == get_VQT of every un-padded clip followed by pad_spectrogram
    (frames the clip does not have are exact zeros and stay out of the per-clip maximum; asap.py:345-349, 383)."""
    from oracle import vqt_oracle as VO
    from piano_a2s_b200.vqt import VQT
    rng = np.random.default_rng(5)
    N = 8000
    lens = [8000, 5120, 3333]
    a = np.zeros((3, N), dtype=np.float32)
    for b, n in enumerate(lens):
        a[b, :n] = np.clip(0.25 * rng.standard_normal(n), -1, 1)
    out = VQT().to(cuda)(torch.from_numpy(a).to(cuda), torch.tensor(lens, device=cuda)).cpu().numpy()
    for b, n in enumerate(lens):
        ref = VO.get_vqt(a[b, :n])
        tb = ref.shape[0]
        assert np.abs(out[b, :tb] - ref).max() < 2e-4
        assert not out[b, tb:].any()
