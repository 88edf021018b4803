"""CPU: C-ABI library loads and exports every declared symbol; drop-in contract of the `models` module; host logic."""
import ctypes
import os

import numpy as np
import pytest
import torch

from oracle import vqt_oracle as VO
from refimport import have_reference, import_reference_models


def test_library_exports_every_declared_symbol():
    from piano_a2s_b200 import _lib
    protos = _lib.parse_header()
    assert len(protos) >= 29
    dll = ctypes.CDLL(_lib.LIB_PATH)
    for name in protos:
        assert hasattr(dll, name), name
    dll.pa2s_dec_args_size.restype = ctypes.c_int
    assert dll.pa2s_dec_args_size() == ctypes.sizeof(_lib.DecArgs)


def test_cpu_tensors_are_refused():
    import models
    m = models.ScoreTranscription(freq_bins=32, max_bars=1, max_length=(4, 3))
    with pytest.raises(RuntimeError):
        m(torch.rand(1, 1, 8, 32), device="cpu")


def test_vocabulary_and_state_dict_contract():
    import models
    assert (models.vocab_size, models.SOS, models.EOS, models.PAD) == (173, 145, 146, 147)
    assert models.labels.checksum() == "2b6fdcff65e5dd2eef3f4492419190653582d0d8ab9ea21f558abc9a0436c1ec"
    m = models.ScoreTranscription(max_length=(398, 189))
    sd = m.state_dict()
    assert len(sd) == 98 and sum(p.numel() for p in m.parameters()) == 16358675
    assert sd["convstack.out.weight"].shape == (256, 19200)
    assert sd["decoder.upper_decoder.gru.weight_ih_l0"].shape == (1536, 528)
    assert sd["decoder.gru.weight_ih_l0"].shape == (1536, 653)


@pytest.mark.skipif(not have_reference(), reason="/root/reference only exists in the build container")
def test_same_seed_gives_the_reference_weights_and_vocabulary():
    import models
    rm = import_reference_models()
    assert rm.labels.labels == models.labels.labels
    torch.manual_seed(1234); a = rm.ScoreTranscription(max_length=(398, 189)).state_dict()
    torch.manual_seed(1234); b = models.ScoreTranscription(max_length=(398, 189)).state_dict()
    assert list(a.keys()) == list(b.keys())
    assert all(torch.equal(a[k], b[k]) for k in a)


def test_filter_design_matches_oracle_definition():
    from piano_a2s_b200.vqt import design_filters
    W, j0 = design_filters()
    G = VO.filter_bank()
    assert np.abs(G[:, :j0]).max() == 0 and np.abs(G[:, j0 + W.shape[1]:]).sum() == 0
    assert np.abs(W[0::2] - G.real[:, j0:j0 + W.shape[1]]).max() < 1e-7
    assert np.abs(W[1::2] - G.imag[:, j0:j0 + W.shape[1]]).max() < 1e-7
    f, l = VO.wavelet_lengths()
    assert abs(l[0] - 787.4916) < 1e-3 and abs(l.sum() - 273041.47) < 0.1


def test_vqt_oracle_known_answers():
    """Pure tone lands in its own bin at full scale; silence maps to the floor; frames = 1 + n//hop."""
    t = np.arange(32000) / 16000.0
    f, _ = VO.wavelet_lengths()
    S = VO.get_vqt(np.sin(2 * np.pi * f[200] * t).astype(np.float32))
    assert S.shape == (201, 480) and S[100].argmax() == 200 and abs(S.max() - 1.0) < 1e-6 and S.min() >= 0.0
    Z = VO.get_vqt(np.zeros(1600, dtype=np.float32))
    assert Z.shape == (11, 480) and np.all(Z == 1.0)   # all-equal magnitudes: ref=max -> 0 dB everywhere (librosa semantics)


def test_steps_from_ground_truth():
    import models
    from helpers import make_ground_truth
    gt = make_ground_truth(4, 3, 14, 9, seed=3, lo_up=(3, 13), lo_lo=(2, 8))
    s = models.HierarchicalDecoder._steps_from_gt(gt[2])
    for bar in range(3):
        assert int(s[bar]) == int(gt[3][:, bar].max()) + 1          # max length + the <eos> step
    full = gt[2].clone()
    full[0, 1, :] = 5                                                 # a row without <eos>: loop never exits early
    assert int(models.HierarchicalDecoder._steps_from_gt(full)[1]) == 14
    # the loader counts the same numbers on the host and attaches them to the (device) targets
    from piano_a2s_b200.train import targets_to_device
    out = targets_to_device(gt, "cpu")
    assert len(out) == 6 and out[2]._pa2s_steps == s.tolist()
    assert out[4]._pa2s_steps == [int(gt[5][:, bar].max()) + 1 for bar in range(3)]


def test_attention_split_covers_all_frames():
    from piano_a2s_b200.ops import attn_split
    for B in (1, 2, 3, 16, 32, 200):
        for T in (1, 7, 24, 1201):
            ns, tile = attn_split(B, T)
            assert ns >= 1 and ns * tile >= T and (ns - 1) * tile < T


def test_flat_adadelta_gathers_gradients_with_one_copy():
    """train.FlatAdadelta: zero_grad() leaves p.grad = None (autograd then keeps the tensors it is handed, no add per parameter);
    gather() moves them into the flat buffer and re-points p.grad at the flat views; parameters without a gradient read as zeros."""
    import torch
    from piano_a2s_b200.train import FlatAdadelta
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Linear(7, 3), torch.nn.Linear(3, 2))
    opt = FlatAdadelta(net)
    params = list(net.parameters())
    assert all(p.grad is not None and p.grad.data_ptr() == v.data_ptr() for p, v in zip(params, opt.gviews))
    opt.zero_grad()
    assert all(p.grad is None for p in params) and float(opt.grad.abs().sum()) == 0.0
    x = torch.randn(4, 5)
    net[1](net[0](x)).square().sum().backward()               # the last layer gets no gradient
    want = [None if p.grad is None else p.grad.clone() for p in params]
    assert want[4] is None and want[5] is None and want[0] is not None
    assert all(p.grad is None or p.grad.data_ptr() != v.data_ptr() for p, v in zip(params, opt.gviews))
    opt.gather()
    for p, v, w, off in zip(params, opt.gviews, want, opt.offsets):
        assert p.grad.data_ptr() == v.data_ptr()
        flat = opt.grad[off:off + p.numel()].view_as(p)
        assert torch.equal(flat, w if w is not None else torch.zeros_like(p))
    opt.gather()                                               # idempotent
    assert torch.equal(opt.grad[opt.offsets[0]:opt.offsets[0] + params[0].numel()].view_as(params[0]), want[0])
    # a second backward without zero_grad accumulates into the flat views, like torch
    net[1](net[0](x)).square().sum().backward()
    assert torch.allclose(params[0].grad, 2 * want[0])
