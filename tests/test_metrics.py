"""Evaluation metrics after the hot path: the oracle's word-error restatement on known answers, macro-F1 pinned
against scikit-learn, and the device WER kernel bit-exact against the oracle (small cases) plus properties at full size."""
import numpy as np
import pytest
import torch

from oracle import metrics_oracle as MO

EOS_, PAD_, TAB_, NL_ = 146, 147, 142, 143


def _inv():
    import models
    return models.labels.labels_map_inv


def test_oracle_word_error_known_answers():
    inv = _inv()
    assert MO.jiwer_words("a \n = \n b  c \t d") == ["a", "=", "b", "c", "d"]
    assert MO.levenshtein(list("kitten"), list("sitting")) == 3
    assert MO.levenshtein([], [1, 2]) == 2 and MO.levenshtein([1, 2], []) == 2 and MO.levenshtein([1], [1]) == 0
    tgt = [[1, 2, 3, EOS_, PAD_], [4, TAB_, 5, EOS_, PAD_]]
    same = MO.clip_wer(tgt, tgt, inv, EOS_)
    assert same == (0.0, 0, 6, 6)                                       # 3 + "=" + 2 words: the "\t" label is whitespace to jiwer
    one_sub = MO.clip_wer([[1, 9, 3, EOS_, 0], [4, 5, NL_, 7, 7]], tgt, inv, EOS_)     # no <eos> in bar 2: all 5 tokens, "\n" dropped
    assert one_sub[1:] == (3, 6, 8)                                     # 1 substitution + 2 insertions
    with pytest.raises(ValueError):
        MO.clip_wer(tgt, [[EOS_, 0, 0, 0, 0]], inv, EOS_)


def test_macro_f1_matches_scikit_learn():
    sk = pytest.importorskip("sklearn.metrics")
    import warnings
    from piano_a2s_b200.metrics import macro_f1
    rng = np.random.RandomState(4)
    for n, k in ((5, 14), (5, 7), (5, 2), (40, 14), (1, 3)):
        for _ in range(20):
            t, p = rng.randint(0, k, n), rng.randint(0, k, n)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                want = sk.f1_score(t, p, average="macro")
            assert abs(macro_f1(t, p) - want) < 1e-12
            assert abs(MO.f1_macro(t.tolist(), p.tolist()) - want) < 1e-12


def test_metrics_refuse_cpu_tensors():
    from piano_a2s_b200.metrics import wer_counts
    with pytest.raises(RuntimeError):
        wer_counts(torch.zeros(1, 1, 4, dtype=torch.int64), torch.zeros(1, 1, 4, dtype=torch.int64))


def _random_rows(rng, B, bars, L, p_eos=0.8):
    tok = rng.randint(0, 148, size=(B, bars, L)).astype(np.int64)
    tok[tok == EOS_] = 3
    for b in range(B):
        for k in range(bars):
            if rng.rand() < p_eos:
                tok[b, k, rng.randint(0, L)] = EOS_
    return tok


@pytest.mark.gpu
@pytest.mark.parametrize("B,bars,Lh,Lr", [(7, 2, 14, 14), (3, 5, 40, 33), (4, 1, 9, 9), (2, 3, 1, 5)])
def test_device_wer_counts_bit_exact(cuda, B, bars, Lh, Lr):
    from piano_a2s_b200.metrics import calculate_wer, wer_counts
    inv = _inv()
    rng = np.random.RandomState(B * 100 + Lh)
    hyp, ref = _random_rows(rng, B, bars, Lh), _random_rows(rng, B, bars, Lr)
    hyp[0] = 0
    hyp[0, :, :min(Lh, Lr)] = ref[0, :, :min(Lh, Lr)]                   # a clip that mostly agrees
    ref[:, 0, 0] = 5                                                    # no empty reference
    ref[-1, -1, 0 if bars > 1 else 1] = EOS_                            # an empty last bar: "... =" with nothing after it (one word when it is the only bar)
    hyp[-1, 0, 0] = EOS_                                                # an empty first hypothesis bar
    d, nr, nh = (t.cpu().numpy() for t in wer_counts(torch.from_numpy(hyp).to(cuda), torch.from_numpy(ref).to(cuda)))
    for b in range(B):
        w, dist, n_ref, n_hyp = MO.clip_wer(hyp[b], ref[b], inv, EOS_)
        assert (d[b], nr[b], nh[b]) == (dist, n_ref, n_hyp), b
    mean, per = calculate_wer(torch.from_numpy(hyp).to(cuda), torch.from_numpy(ref).to(cuda))
    assert per == [MO.clip_wer(hyp[b], ref[b], inv, EOS_)[0] for b in range(B)] and abs(mean - np.mean(per)) < 1e-15
    empty = ref.copy()
    empty[1, :, 0] = EOS_                                               # every bar of clip 1 empty: only the bars - 1 "=" words are left
    if bars == 1:
        with pytest.raises(ValueError):                                 # jiwer: "one or more references are empty strings"
            calculate_wer(torch.from_numpy(hyp).to(cuda), torch.from_numpy(empty).to(cuda))
    else:
        assert wer_counts(torch.from_numpy(hyp).to(cuda), torch.from_numpy(empty).to(cuda))[1][1].item() == bars - 1
        assert MO.clip_wer(hyp[1], empty[1], inv, EOS_)[2] == bars - 1


@pytest.mark.gpu
def test_device_wer_full_size_properties(cuda):
    """pretrain.yaml sizes (5 bars x 398 tokens, 32 clips): identity, k substitutions, symmetry, a deleted bar."""
    from piano_a2s_b200.metrics import wer_counts
    rng = np.random.RandomState(1)
    B, bars, L = 32, 5, 398
    ref = rng.randint(0, 140, size=(B, bars, L)).astype(np.int64)       # no <eos>, no whitespace labels: 5*398 + 4 words per clip
    r = torch.from_numpy(ref).to(cuda)
    d, nr, nh = wer_counts(r, r)
    assert d.tolist() == [0] * B and nr.tolist() == [bars * L + bars - 1] * B and nh.tolist() == nr.tolist()
    hyp = ref.copy()
    ks = rng.randint(0, 300, size=B)
    for b in range(B):
        pos = rng.choice(bars * L, size=ks[b], replace=False)
        flat = hyp[b].reshape(-1)
        flat[pos] = 141                                                 # a label that never occurs in ref: exactly k substitutions
    h = torch.from_numpy(hyp).to(cuda)
    d1 = wer_counts(h, r)[0].tolist()
    assert d1 == ks.tolist()
    assert wer_counts(r, h)[0].tolist() == d1                           # symmetric
    cut = ref.copy()
    cut[:, 2, 0] = EOS_                                                 # bar 3 empty: its 398 words are deleted, the "=" stays
    assert wer_counts(torch.from_numpy(cut).to(cuda), r)[0].tolist() == [L] * B


@pytest.mark.gpu
def test_evaluate_batch_on_model_outputs(cuda):
    import models
    from helpers import make_ground_truth, synth_state_dict, lcg_uniform
    from piano_a2s_b200 import kern
    from piano_a2s_b200.metrics import evaluate_batch
    cfg = dict(freq_bins=32, max_bars=2, max_length=(14, 9))
    torch.manual_seed(1234)
    m = models.ScoreTranscription(**cfg)
    m.load_state_dict(synth_state_dict(m))
    m = m.to(cuda).eval()
    gt = make_ground_truth(4, 2, 14, 9, seed=2, lo_up=(3, 13), lo_lo=(2, 9))
    with torch.no_grad():
        outs = m(lcg_uniform((4, 1, 20, 32), seed=8).to(cuda), device=cuda)
    res = evaluate_batch(outs, [g.to(cuda) for g in gt])
    toks = kern.greedy_tokens(outs)
    inv = models.labels.labels_map_inv
    for name, gi in (("upper", 2), ("lower", 4)):
        for b in range(4):
            pred_rows = [row + [146] for row in toks[name][b]]          # unpadded lists -> rows that end at <eos>
            want = MO.clip_wer(pred_rows, gt[gi][b].tolist(), inv, 146)[0]
            assert res[f"wer_{name}_per_clip"][b] == want
    for name, gi in (("key", 1), ("time_sig", 0)):
        k = "key_f1_per_clip" if name == "key" else "time_f1_per_clip"
        for b in range(4):
            assert abs(res[k][b] - MO.f1_macro(gt[gi][b].tolist(), toks[name][b])) < 1e-12
