"""Generates tests/golden/*.npz from the UNMODIFIED reference (/root/reference/models.py, music21 stubbed).

Run in the build container only (the reference does not travel to the GPU box):
    python tests/golden/make_golden.py
Weights and inputs are produced by `helpers.synth_state_dict` / `lcg_uniform` (pure integer arithmetic), so the CUDA
path and the oracle can rebuild them bit-exactly anywhere; only the reference's OUTPUTS are stored here.
"""
import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from helpers import make_ground_truth, synth_state_dict, lcg_uniform  # noqa: E402
from oracle import a2s_oracle as O  # noqa: E402
from refimport import import_reference_models  # noqa: E402

SMALL = dict(freq_bins=32, max_bars=2, max_length=(14, 9))
FULL = dict(max_length=(398, 189))


def grad_digest(g):
    g = g.detach().double().reshape(-1)
    idx = np.linspace(0, g.numel() - 1, 8).astype(np.int64)
    return np.concatenate([[g.sum().item(), g.abs().sum().item(), g.abs().max().item()], g[idx].numpy()])


def main():
    rm = import_reference_models()
    out = {}
    # ---- small config: eval greedy + teacher-forced training forward/backward --------------------------------
    m = rm.ScoreTranscription(**SMALL)
    m.load_state_dict(synth_state_dict(m))
    B, T = 3, 24
    x = lcg_uniform((B, 1, T, 32), seed=9)
    m.eval()
    with torch.no_grad():
        ev = m(x, device="cpu")
    for n, t in zip(("ts", "key", "up", "lo"), ev):
        out["small_eval_" + n] = t.numpy()
    gt = make_ground_truth(B, 2, 14, 9, seed=4, lo_up=(3, 13), lo_lo=(2, 9))
    m.train()
    rec = O.RecordingSource()
    # record the reference's own randomness by routing its F.dropout / random.random through the recorder
    import torch.nn.functional as F
    orig_F, orig_random = rm.F, rm.random      # the reference's module-level names `F` and `random`
    torch.manual_seed(21)
    random.seed(21)

    def rec_dropout(t, p=0.5, training=True, inplace=False):
        if not training:
            return t
        if t.dim() == 3 and t.stride(2) != 1:          # conv features: transposed view of (B,C,T) (models.py:539-541)
            mk = rec.dropout_mask((t.shape[0], t.shape[2], t.shape[1]), p).transpose(1, 2)
        else:
            mk = rec.dropout_mask(tuple(t.shape), p)
        return t * mk
    class _Shim:                               # proxy a module, overriding one attribute (no global monkeypatch)
        def __init__(self, mod, **over):
            self._mod, self._over = mod, over

        def __getattr__(self, name):
            return self._over[name] if name in self._over else getattr(self._mod, name)
    rm.F = _Shim(orig_F, dropout=rec_dropout)
    rm.random = _Shim(orig_random, random=rec.coin)
    try:
        tr = m(x, inference=False, ground_truth=gt, teacher_forcing_ratio=0.6, device="cpu")
    finally:
        rm.F, rm.random = orig_F, orig_random
    loss = O.training_loss(tr, gt)
    loss.backward()
    for n, t in zip(("ts", "key", "up", "lo"), tr):
        out["small_train_" + n] = t.detach().numpy()
    out["small_train_loss"] = np.array(loss.item())
    out["small_train_coins"] = np.array(rec.coins)
    out["small_train_nmasks"] = np.array(len(rec.masks))
    for i, mk in enumerate(rec.masks):
        out[f"small_train_mask_{i}"] = (mk != 0).numpy()
    out["small_train_mask_p"] = np.array([0.2] + [0.1] * (len(rec.masks) - 1))
    for k, p in m.named_parameters():
        out["small_grad_" + k] = grad_digest(p.grad)
    for k, v in m.state_dict().items():
        if "running" in k:
            out["small_stat_" + k] = v.numpy()
    # ---- full config (pretrain.yaml), one clip, greedy: token ids + log-prob digests -------------------------
    m = rm.ScoreTranscription(**FULL)
    m.load_state_dict(synth_state_dict(m))
    m.eval()
    x = lcg_uniform((1, 1, 1201, 480), seed=1234)
    with torch.no_grad():
        fv = m(x, device="cpu")
    up, lo = fv[2], fv[3]
    out["full_ts"] = fv[0].numpy()
    out["full_key"] = fv[1].numpy()
    out["full_up_tokens"] = up.argmax(-1).numpy().astype(np.int16)
    out["full_lo_tokens"] = lo.argmax(-1).numpy().astype(np.int16)
    for n, t in (("up", up), ("lo", lo)):
        top2 = t.topk(2, -1).values
        out[f"full_{n}_margin"] = (top2[..., 0] - top2[..., 1]).numpy().astype(np.float32)
        out[f"full_{n}_top"] = top2[..., 0].numpy().astype(np.float32)
    print("min margins", out["full_up_margin"].min(), out["full_lo_margin"].min())
    np.savez_compressed(os.path.join(HERE, "reference_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "reference_golden.npz"), os.path.getsize(os.path.join(HERE, "reference_golden.npz")))


if __name__ == "__main__":
    main()
