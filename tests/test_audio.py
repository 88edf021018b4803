"""Audio ingest: the two arithmetic lines of datasets/asap.py:83-86 are plain torch ops, so torch on the CPU IS the
reference for them; the device kernel has to match bit for bit (mono / stereo)."""
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
