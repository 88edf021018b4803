"""CPU: the oracle (oracle/a2s_oracle.py) against the golden vectors produced from the unmodified reference
(tests/golden/make_golden.py), and -- when /root/reference is present (build container) -- against the reference itself."""
import os
import random

import numpy as np
import pytest
import torch

from helpers import lcg_uniform, make_ground_truth, synth_state_dict
from oracle import a2s_oracle as O
from refimport import have_reference, import_reference_models

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_golden.npz"))
SMALL = dict(freq_bins=32, max_bars=2, max_length=(14, 9))
FULL = dict(max_length=(398, 189))


def _small_sd():
    import models
    return synth_state_dict(models.ScoreTranscription(**SMALL))


def golden_masks():
    n = int(GOLD["small_train_nmasks"])
    p = GOLD["small_train_mask_p"]
    return [torch.from_numpy(GOLD[f"small_train_mask_{i}"]).float() / (1.0 - float(p[i])) for i in range(n)]


def grad_digest(g):
    g = g.detach().double().reshape(-1)
    idx = np.linspace(0, g.numel() - 1, 8).astype(np.int64)
    return np.concatenate([[g.sum().item(), g.abs().sum().item(), g.abs().max().item()], g[idx].numpy()])


def test_oracle_small_eval_matches_golden():
    sd = _small_sd()
    x = lcg_uniform((3, 1, 24, 32), seed=9)
    with torch.no_grad():
        outs = O.score_transcription(sd, x, SMALL)
    for n, t in zip(("ts", "key", "up", "lo"), outs):
        assert np.abs(t.numpy() - GOLD["small_eval_" + n]).max() < 2e-5, n


def test_oracle_small_training_matches_golden():
    sd = _small_sd()
    x = lcg_uniform((3, 1, 24, 32), seed=9)
    gt = make_ground_truth(3, 2, 14, 9, seed=4, lo_up=(3, 13), lo_lo=(2, 9))
    sdg = {k: v.clone().requires_grad_(v.dtype == torch.float32 and "running" not in k) for k, v in sd.items()}
    ns = {}
    outs = O.score_transcription(sdg, x, SMALL, False, gt, 0.6, True, O.ReplaySource(GOLD["small_train_coins"].tolist(), golden_masks()),
                                 new_stats=ns)
    loss = O.training_loss(outs, gt)
    loss.backward()
    for n, t in zip(("ts", "key", "up", "lo"), outs):
        assert np.abs(t.detach().numpy() - GOLD["small_train_" + n]).max() < 2e-5, n
    assert abs(loss.item() - float(GOLD["small_train_loss"])) < 1e-5
    for k, v in sdg.items():
        if v.requires_grad:
            d, g = grad_digest(v.grad), GOLD["small_grad_" + k]
            assert np.abs(d - g).max() <= 2e-4 * max(1e-6, np.abs(g).max()), k
    for k, v in ns.items():
        if "running" in k:
            assert np.abs(v.numpy() - GOLD["small_stat_" + k]).max() < 1e-6, k


def test_oracle_full_size_greedy_tokens_match_golden():
    import models
    sd = synth_state_dict(models.ScoreTranscription(**FULL))
    x = lcg_uniform((1, 1, 1201, 480), seed=1234)
    with torch.no_grad():
        outs = O.score_transcription(sd, x, FULL)
    assert np.array_equal(outs[2].argmax(-1).numpy(), GOLD["full_up_tokens"])
    assert np.array_equal(outs[3].argmax(-1).numpy(), GOLD["full_lo_tokens"])
    assert np.abs(outs[0].numpy() - GOLD["full_ts"]).max() < 1e-4
    assert np.abs(outs[2].max(-1).values.numpy() - GOLD["full_up_top"]).max() < 1e-3


@pytest.mark.skipif(not have_reference(), reason="/root/reference only exists in the build container")
def test_oracle_matches_live_reference_with_native_rng():
    """Same torch / python seeds -> the oracle consumes both RNG streams exactly like the reference (train mode)."""
    rm = import_reference_models()
    torch.manual_seed(0)
    m = rm.ScoreTranscription(**SMALL)
    for k, v in m.state_dict().items():
        if "bn" in k and v.dtype == torch.float32:
            v.copy_(torch.rand_like(v) + 0.5)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    x = torch.rand(2, 1, 19, 32)
    gt = make_ground_truth(2, 2, 14, 9, seed=2, lo_up=(3, 13), lo_lo=(2, 9))
    m.train()
    torch.manual_seed(5); random.seed(5)
    r = m(x, inference=False, ground_truth=gt, teacher_forcing_ratio=0.6, device="cpu")
    O.training_loss(r, gt).backward()
    sdg = {k: v.clone().requires_grad_(v.dtype == torch.float32 and "running" not in k) for k, v in sd.items()}
    torch.manual_seed(5); random.seed(5)
    o = O.score_transcription(sdg, x, SMALL, False, gt, 0.6, True)
    O.training_loss(o, gt).backward()
    for a, b in zip(r, o):
        assert (a - b).abs().max() < 2e-5
    for k, p in m.named_parameters():
        assert (p.grad - sdg[k].grad).abs().max() <= 1e-4 * (p.grad.abs().max() + 1e-12), k
