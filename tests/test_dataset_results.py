"""N2 (feature-cache reader + double-buffered batch loader) and N3 (evaluation records + result JSON files): a small synthetic feature
folder in the reference's layout (datasets/syn.py:28-36, 88-121) is read back and compared with the collate oracle; the result files
follow pretrain.py:189-214 field by field."""
import json
import os
import pickle

import numpy as np
import pytest
import torch

from oracle import a2s_oracle as O
from oracle import collate_oracle as CO

SPLIT = "valid"
TS = ["4/4", "3/4", "2/4", "6/8", "2/2", "12/8", "3/8"]


def _make_folder(root, n_songs=5, versions=(0, 1), bars=2, seed=0):
    rng = np.random.default_rng(seed)
    songs = {}
    for v in versions:
        for sub in ("spectrogram", "target", "info"):
            os.makedirs(os.path.join(root, SPLIT, str(v), sub), exist_ok=True)
        for i in range(n_songs):
            chunk = f"{'c' if i % 2 else 'P'}hunk{i}"
            score = [[int(rng.integers(-6, 8)), TS[int(rng.integers(7))], [int(t) for t in rng.integers(0, 144, rng.integers(1, 8))],
                      [int(t) for t in rng.integers(0, 144, rng.integers(1, 12))]] for _ in range(bars)]
            with open(os.path.join(root, SPLIT, str(v), "target", f"{chunk}.pkl"), "wb") as f:
                pickle.dump(score, f)
            with open(os.path.join(root, SPLIT, str(v), "info", f"{chunk}.json"), "w") as f:
                json.dump({"composer": f"composer{i}"}, f)
            spec = rng.random((int(rng.integers(5, 20)), 32), dtype=np.float32)
            np.save(os.path.join(root, SPLIT, str(v), "spectrogram", f"{chunk}~sf{v}.npy"), spec)
            songs[(v, f"{chunk}~sf{v}")] = (spec, score)
    return songs


def test_feature_folder_lists_and_reads_like_the_reference(tmp_path):
    from piano_a2s_b200.dataset import FeatureFolder
    songs = _make_folder(str(tmp_path))
    test = FeatureFolder(str(tmp_path), SPLIT, versions=(0, 1), train=False)
    assert len(test) == 10
    for i in range(len(test)):
        spec, ts, key, up, lo, name, v = test.item(i)
        ref_spec, score = songs[(v, name)]
        assert np.array_equal(spec, ref_spec)
        assert ts == [TS.index(b[1]) for b in score] and key == [b[0] for b in score]
        assert up == [b[3] for b in score] and lo == [b[2] for b in score]
    train = FeatureFolder(str(tmp_path), SPLIT, versions=(0, 1), train=True, seed=1)
    assert len(train) == 5
    assert {train.item(7)[5].split("~")[0]} == {sorted(s for (v, s) in songs if v == 0)[7 % 5].split("~")[0]}


@pytest.mark.gpu
def test_batch_loader_matches_collate_oracle(cuda, tmp_path):
    from piano_a2s_b200.dataset import BatchLoader, FeatureFolder
    songs = _make_folder(str(tmp_path))
    folder = FeatureFolder(str(tmp_path), SPLIT, versions=(0, 1), train=False)
    seen = 0
    for spec, gt, names, versions in BatchLoader(folder, 4, 24, (14, 9), cuda):
        B = len(names)
        for b in range(B):
            ref_spec, score = songs[(versions[b], names[b])]
            want = CO.pad_spectrogram(ref_spec, 24)
            assert torch.equal(spec[b].cpu(), torch.from_numpy(want))
            up, ul = CO.pad_score([bar[3] for bar in score], 14, O.PAD, O.EOS)
            lo, ll = CO.pad_score([bar[2] for bar in score], 9, O.PAD, O.EOS)
            assert torch.equal(gt[2][b].cpu(), torch.from_numpy(up)) and torch.equal(gt[3][b].cpu(), torch.from_numpy(ul))
            assert torch.equal(gt[4][b].cpu(), torch.from_numpy(lo)) and torch.equal(gt[5][b].cpu(), torch.from_numpy(ll))
            assert gt[0][b].cpu().tolist() == [TS.index(bar[1]) for bar in score]
            assert gt[1][b].cpu().tolist() == [bar[0] + 6 for bar in score]
        seen += B
    assert seen == 10


@pytest.mark.gpu
def test_result_files_follow_the_reference_layout(cuda, tmp_path):
    """ResultRecorder.add_batch + write == what compute_objectives records and on_stage_end saves (pretrain.py:95-117, 189-214)."""
    from piano_a2s_b200 import results
    from piano_a2s_b200.dataset import BatchLoader, FeatureFolder
    _make_folder(str(tmp_path))
    folder = FeatureFolder(str(tmp_path), SPLIT, versions=(0,), train=False)
    rec = results.ResultRecorder()
    g = torch.Generator().manual_seed(3)
    kept = {}
    for spec, gt, names, versions in BatchLoader(folder, 3, 24, (14, 9), cuda):
        B = len(names)
        outs = [torch.log_softmax(torch.randn(B, 2, 7, generator=g), -1), torch.log_softmax(torch.randn(B, 2, 14, generator=g), -1),
                torch.log_softmax(torch.randn(B, 2, 14, 173, generator=g), -1), torch.log_softmax(torch.randn(B, 2, 9, 173, generator=g), -1)]
        rec.add_batch([o.to(cuda) for o in outs], gt, names, versions)
        want = O.greedy_tokens(outs)
        for b in range(B):
            kept["~".join([str(versions[b]), names[b]])] = (want["key"][b], want["time_sig"][b], want["lower"][b], want["upper"][b])
    paths = rec.write(str(tmp_path / "out"), str(tmp_path), SPLIT)
    assert len(paths) == 5
    for p in paths:
        cid = os.path.basename(p)[:-5]
        res = json.load(open(p))
        assert set(res) == {"style", "soundfont", "composer", "target_path", "pred", "wer_upper", "wer_lower", "key_f1", "time_f1"}
        key, ts, lo, up = kept[cid]
        assert res["pred"] == [[key[i] - 6, TS[ts[i]], lo[i], up[i]] for i in range(2)]
        version, chunk, sf = cid.split("~")
        assert res["soundfont"] == sf and res["style"] == ("classical" if chunk[0].islower() else "pop")
        assert res["composer"] == f"composer{chunk[-1]}" and res["target_path"].endswith(os.path.join(SPLIT, version, "target", f"{chunk}.pkl"))
        assert 0.0 <= res["key_f1"] <= 1.0 and res["wer_upper"] >= 0.0
    st = rec.stage_stats()
    assert abs(st["WER"] - (st["wer_upper"] + st["wer_lower"]) / 2) < 1e-12
