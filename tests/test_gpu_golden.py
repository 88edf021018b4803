"""GPU: the CUDA path (through the C ABI) against the golden vectors produced from the UNMODIFIED reference
(tests/golden/make_golden.py, run in the build container where /root/reference exists).  Nothing here touches the oracle."""
import os

import numpy as np
import pytest
import torch

from helpers import ReplayDeviceSource, lcg_uniform, make_ground_truth, synth_state_dict

pytestmark = pytest.mark.gpu

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_golden.npz"))
SMALL = dict(freq_bins=32, max_bars=2, max_length=(14, 9))
FULL = dict(max_length=(398, 189))


def _model(cuda, cfg):
    import models
    m = models.ScoreTranscription(**cfg)
    m.load_state_dict(synth_state_dict(m))
    return m.to(cuda)


def _masks():
    n = int(GOLD["small_train_nmasks"])
    p = GOLD["small_train_mask_p"]
    return [torch.from_numpy(GOLD[f"small_train_mask_{i}"]).float() / (1.0 - float(p[i])) for i in range(n)]


def _digest(g):
    g = g.detach().double().reshape(-1).cpu()
    idx = np.linspace(0, g.numel() - 1, 8).astype(np.int64)
    return np.concatenate([[g.sum().item(), g.abs().sum().item(), g.abs().max().item()], g[idx].numpy()])


def test_small_eval_matches_reference_golden(cuda):
    from piano_a2s_b200 import ops
    m = _model(cuda, SMALL).eval()
    x = lcg_uniform((3, 1, 24, 32), seed=9)
    with torch.no_grad():
        outs = m(x.to(cuda), device=cuda)
    for n, t in zip(("ts", "key", "up", "lo"), outs):
        g = GOLD["small_eval_" + n]
        assert np.abs(t.cpu().numpy() - g).max() < 1e-4 * max(1.0, np.abs(g).max()), n
        assert np.array_equal(t.argmax(-1).cpu().numpy(), g.argmax(-1)), n      # greedy tokens bit-identical
    ops.check_sync_flags()


def test_small_training_matches_reference_golden(cuda):
    from piano_a2s_b200 import ops, rng
    from piano_a2s_b200.train import compute_objectives
    m = _model(cuda, SMALL).train()
    x = lcg_uniform((3, 1, 24, 32), seed=9)
    gt = [g.to(cuda) for g in make_ground_truth(3, 2, 14, 9, seed=4, lo_up=(3, 13), lo_lo=(2, 9))]
    with rng.use_source(ReplayDeviceSource(GOLD["small_train_coins"].tolist(), _masks())):
        outs = m(x.to(cuda), inference=False, ground_truth=gt, teacher_forcing_ratio=0.6, device=cuda)
    loss, _ = compute_objectives(outs, gt)
    loss.backward()
    for n, t in zip(("ts", "key", "up", "lo"), outs):
        g = GOLD["small_train_" + n]
        assert np.abs(t.detach().cpu().numpy() - g).max() < 2e-4 * max(1.0, np.abs(g).max()), n
    assert abs(loss.item() - float(GOLD["small_train_loss"])) < 1e-4 * float(GOLD["small_train_loss"])
    for k, p in m.named_parameters():
        g = GOLD["small_grad_" + k]
        d = _digest(p.grad) if p.grad is not None else np.zeros_like(g)
        assert np.abs(d - g).max() <= 2e-3 * max(1e-6, np.abs(g).max()), (k, d, g)
    sd = m.state_dict()
    for k in GOLD.files:
        if k.startswith("small_stat_") and "running" in k:
            assert np.abs(sd[k[len("small_stat_"):]].cpu().numpy() - GOLD[k]).max() < 1e-5, k
    ops.check_sync_flags()


def test_full_size_greedy_tokens_bit_identical_to_reference(cuda):
    """pretrain.yaml model, one 12-s clip, 5 bars x (398 + 189) greedy steps in the fp32 path: every token equals the
    reference's (BASELINE north_star: bit-identical greedy **kern token sequences)."""
    from piano_a2s_b200 import ops
    m = _model(cuda, FULL).eval()
    x = lcg_uniform((1, 1, 1201, 480), seed=1234)
    with torch.no_grad():
        outs = m(x.to(cuda), device=cuda)
    up = outs[2].argmax(-1).cpu().numpy()
    lo = outs[3].argmax(-1).cpu().numpy()
    assert np.array_equal(up, GOLD["full_up_tokens"]), int((up != GOLD["full_up_tokens"]).sum())
    assert np.array_equal(lo, GOLD["full_lo_tokens"]), int((lo != GOLD["full_lo_tokens"]).sum())
    assert np.abs(outs[0].cpu().numpy() - GOLD["full_ts"]).max() < 1e-4
    assert np.abs(outs[1].cpu().numpy() - GOLD["full_key"]).max() < 1e-4
    assert np.abs(outs[2].max(-1).values.cpu().numpy() - GOLD["full_up_top"]).max() < 1e-3
    assert np.abs(outs[3].max(-1).values.cpu().numpy() - GOLD["full_lo_top"]).max() < 1e-3
    ops.check_sync_flags()
