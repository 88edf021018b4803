"""The alternative implementations kept as comparators behind switches must agree with the default path on a whole training
step (same weights, same replayed coins / dropout masks): every output and every gradient.

  * ops.BAR_CHAIN = False      the bar-level chain as individual autograd ops instead of one BarStepFn node per bar
  * pa2s_conv_tma_set_impl(0)  conv_tma_kernel (one instruction group per tap) instead of conv_tma3_kernel
  * pa2s_gru_seq_set_exchange  0: DSMEM stores + cluster barrier, 3: st.async with the column-owner reverse kernel (default 1)
"""
import random

import pytest
import torch

from helpers import ReplayDeviceSource, make_ground_truth, rel_err, synth_state_dict
from oracle import a2s_oracle as O

pytestmark = pytest.mark.gpu

CFG = dict(freq_bins=32, max_bars=3, max_length=(14, 9))


def _step(cuda, rec, gt, x, sd):
    import models
    from piano_a2s_b200 import rng
    from piano_a2s_b200.train import compute_objectives
    torch.manual_seed(1234)
    m = models.ScoreTranscription(**CFG)
    m.load_state_dict(sd)
    m = m.to(cuda).train()
    with rng.use_source(ReplayDeviceSource(rec.coins, rec.masks)):
        outs = m(x.to(cuda), inference=False, ground_truth=[g.to(cuda) for g in gt], teacher_forcing_ratio=0.6, device=cuda)
    loss, _ = compute_objectives(outs, [g.to(cuda) for g in gt])
    loss.backward()
    torch.cuda.synchronize()
    return [o.detach().clone() for o in outs], {k: p.grad.detach().clone() for k, p in m.named_parameters() if p.grad is not None}


@pytest.fixture(scope="module")
def case():
    import models
    torch.manual_seed(1234)
    sd = synth_state_dict(models.ScoreTranscription(**CFG))
    B, T = 5, 40
    x = torch.rand(B, 1, T, 32, generator=torch.Generator().manual_seed(9))
    gt = make_ground_truth(B, 3, 14, 9, seed=4, lo_up=(3, 13), lo_lo=(2, 9))
    rec = O.RecordingSource()
    torch.manual_seed(21)
    random.seed(21)
    sdg = {k: v.clone() for k, v in sd.items()}
    with torch.no_grad():
        O.score_transcription(sdg, x, CFG, False, gt, 0.6, True, rec)       # only to record coins / masks in the reference's order
    return rec, gt, x, sd


def _compare(base, other, tol_out=2e-5, tol_grad=2e-4):
    for a, b in zip(base[0], other[0]):
        assert rel_err(a, b) < tol_out
    assert base[1].keys() == other[1].keys()
    worst = max((rel_err(other[1][k], g), k) for k, g in base[1].items() if g.abs().max() > 0)
    print("worst gradient difference", worst)
    assert worst[0] < tol_grad, worst


def test_bar_chain_nodes_match_individual_ops(cuda, case):
    from piano_a2s_b200 import ops
    base = _step(cuda, *case)
    ops.BAR_CHAIN = False
    try:
        other = _step(cuda, *case)
    finally:
        ops.BAR_CHAIN = True
    _compare(base, other)


def test_conv_tap_kernel_matches_shared_window_kernel(cuda, case):
    from piano_a2s_b200._lib import lib
    base = _step(cuda, *case)
    lib.pa2s_conv_tma_set_impl(0)
    try:
        other = _step(cuda, *case)
    finally:
        lib.pa2s_conv_tma_set_impl(1)
    _compare(base, other)


@pytest.mark.parametrize("mode", [0, 3])
def test_encoder_exchange_modes_match(cuda, case, mode):
    from piano_a2s_b200._lib import lib
    base = _step(cuda, *case)
    lib.pa2s_gru_seq_set_exchange(mode)
    try:
        other = _step(cuda, *case)
    finally:
        lib.pa2s_gru_seq_set_exchange(1)
    _compare(base, other)
