"""Batch collation (SURVEY 8 a2 + target half of N2): oracle vs the reference (golden file, and live when /root/reference is
mounted), host-side target padding vs the oracle, and the device padding kernel vs the oracle (bit-exact)."""
import os

import numpy as np
import pytest
import torch

from oracle import collate_oracle as CO
from refimport import have_reference_tree as have_reference

GOLD = os.path.join(os.path.dirname(__file__), "golden", "collate_golden.npz")
PAD_, EOS_ = 147, 146


def _unpack(mat):
    return [[int(v) for v in row if v >= 0] for row in mat]


def test_vocabulary_ids():
    import models
    assert (models.PAD if hasattr(models, "PAD") else models.labels.labels_map["<pad>"]) == PAD_ and models.EOS == EOS_


def test_oracle_matches_reference_golden():
    g = np.load(GOLD)
    n_spec = sum(1 for k in g.files if k.startswith("spec_in_"))
    raised = 0
    for i in range(n_spec):
        if f"spec_raises_{i}" in g.files:
            raised += 1
            with pytest.raises(RuntimeError):
                CO.pad_spectrogram(g[f"spec_in_{i}"], 30)
            assert CO.pad_spectrogram(g[f"spec_in_{i}"], 30, truncate=True).shape == (1, 30, 12)
        else:
            out = CO.pad_spectrogram(g[f"spec_in_{i}"], 30)
            assert out.dtype == np.float32 and np.array_equal(out, g[f"spec_out_{i}"])
    assert raised == 2                                                    # the 31- and 45-frame clips
    for i in range(3):
        rows, ln = CO.pad_score(_unpack(g[f"score_in_{i}"]), 10, PAD_, EOS_)
        assert np.array_equal(rows, g[f"score_out_{i}"]) and np.array_equal(ln, g[f"score_len_{i}"])
        assert np.array_equal(CO.key_to_int(g[f"key_in_{i}"]), g[f"key_out_{i}"])


@pytest.mark.skipif(not have_reference(), reason="/root/reference is only mounted in the build container")
def test_oracle_matches_live_reference():
    from refimport import reference_dataset_stub
    ds = reference_dataset_stub(max_frame_num=1201)
    rng = np.random.RandomState(11)
    for n in (0, 5, 1200, 1201):
        s = rng.rand(n, 480).astype(np.float32)
        assert np.array_equal(ds.pad_spectrogram(s).numpy(), CO.pad_spectrogram(s, 1201))
    with pytest.raises(RuntimeError):
        ds.pad_spectrogram(rng.rand(1202, 480).astype(np.float32))
    for L in (398, 189):
        score = [list(rng.randint(0, 144, size=n)) for n in (0, 1, L - 1, L, L + 7)]
        rows, ln = ds.pad_score(score, L)
        orows, oln = CO.pad_score(score, L, PAD_, EOS_)
        assert np.array_equal(rows.numpy(), orows) and np.array_equal(ln.numpy(), oln)


def test_host_target_padding_matches_oracle():
    from piano_a2s_b200.batching import pad_scores
    rng = np.random.RandomState(3)
    scores = [[list(rng.randint(0, 144, size=n)) for n in ns] for ns in ((0, 3, 9, 10), (10, 11, 4, 1), (1, 9, 25, 0))]
    tok, ln = pad_scores(scores, 10)
    assert tok.dtype == torch.int64 and tok.shape == (3, 4, 10) and ln.shape == (3, 4)
    for b, sc in enumerate(scores):
        rows, l = CO.pad_score(sc, 10, PAD_, EOS_)
        assert np.array_equal(tok[b].numpy(), rows) and np.array_equal(ln[b].numpy(), l)
    with pytest.raises(ValueError):
        pad_scores([[[1]], [[1], [2]]], 10)


def test_pad_spectrograms_refuses_cpu():
    from piano_a2s_b200.batching import pad_spectrograms
    with pytest.raises(RuntimeError):
        pad_spectrograms([np.zeros((3, 8), np.float32)], 5, "cpu")


@pytest.mark.gpu
@pytest.mark.parametrize("F,Tmax,rows", [(480, 1201, (1201, 0, 400, 1, 977)), (12, 30, (0, 1, 7, 30)), (7, 9, (9, 3, 0)), (480, 64, (64,))])
def test_device_padding_bit_exact(cuda, F, Tmax, rows):
    from piano_a2s_b200.batching import pad_spectrograms
    rng = np.random.RandomState(F + Tmax)
    specs = [rng.randn(n, F).astype(np.float32) for n in rows]
    want = np.stack([CO.pad_spectrogram(s, Tmax) for s in specs])
    got = pad_spectrograms(specs, Tmax, cuda)
    assert got.shape == (len(rows), 1, Tmax, F) and got.dtype == torch.float32
    assert np.array_equal(got.cpu().numpy(), want)
    # device-resident clips (e.g. VQT outputs of different lengths) take the same kernel without the staging copy
    got2 = pad_spectrograms([torch.from_numpy(s).to(cuda) for s in specs], Tmax, cuda)
    assert torch.equal(got, got2)
    # over-long clip: raises like the reference, truncates on request
    long = specs + [rng.randn(Tmax + 3, F).astype(np.float32)]
    with pytest.raises(RuntimeError):
        pad_spectrograms(long, Tmax, cuda)
    got3 = pad_spectrograms(long, Tmax, cuda, truncate=True)
    assert np.array_equal(got3[-1].cpu().numpy(), CO.pad_spectrogram(long[-1], Tmax, truncate=True))
    assert torch.equal(got3[:-1], got)


@pytest.mark.gpu
def test_collate_feeds_the_model(cuda):
    """collate() -> the six target tensors + step counts the decoder launches with, identical to the oracle's collation."""
    import models
    from piano_a2s_b200.batching import collate
    rng = np.random.RandomState(9)
    items = []
    for b in range(3):
        items.append((rng.rand(10 + 5 * b, 32).astype(np.float32), list(rng.randint(0, 7, 2)), list(rng.randint(-6, 8, 2)),
                      [list(rng.randint(0, 144, size=rng.randint(1, 14))) for _ in range(2)],
                      [list(rng.randint(0, 144, size=rng.randint(1, 9))) for _ in range(2)]))
    spec, gt = collate(items, 24, (14, 9), cuda)
    ospec, ogt = CO.collate(items, 24, (14, 9), PAD_, EOS_)
    assert np.array_equal(spec.cpu().numpy(), ospec)
    for a, b in zip(gt, ogt):
        assert a.dtype == torch.int64 and np.array_equal(a.cpu().numpy(), b)
    assert gt[2]._pa2s_steps == models.HierarchicalDecoder._steps_from_gt(gt[2].cpu()).tolist()
    assert gt[4]._pa2s_steps == models.HierarchicalDecoder._steps_from_gt(gt[4].cpu()).tolist()
