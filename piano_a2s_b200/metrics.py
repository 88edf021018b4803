"""Evaluation metrics right after the hot path: `calculate_wer` / `caculate_f1` of pretrain.py:216-243.

The reference scores every validation clip on the host: token ids -> label strings -> `jiwer.wer(target, pred)` (jiwer 3.0.3,
environment.yaml:47) and `sklearn.metrics.f1_score(average="macro")`.  Here the word error counts of all clips of a batch come
from ONE libpa2s launch on the token tensors the decoder left on the GPU (`pa2s_wer_counts`, csrc/metrics.cu: per clip, build both
word sequences, anti-diagonal Levenshtein), and one small device->host copy; macro-F1 over <= 14 classes x 5 bars stays numpy.
jiwer itself is not in this image: what its default transform does to these strings (whitespace runs collapse, so the "\\t" / "\\n"
labels vanish and the bar separator " \\n = \\n " becomes one "=" word) is restated from its published source -- parity unpinned,
see oracle/metrics_oracle.py.
"""
from __future__ import annotations

import numpy as np
import torch

from ._lib import lib, ptr, stream
from .models import EOS, labels, vocab_size

_SKIP = tuple(i for i, l in enumerate(labels.labels) if l.strip() == "")      # "\t", "\n": removed by jiwer's whitespace collapsing
_SEP = vocab_size                                                                # the "=" word between bars: not a label
assert len(_SKIP) == 2


def wer_counts(hyp_tokens: torch.Tensor, ref_tokens: torch.Tensor, eos: int = EOS):
    """(B, bars, Lh) / (B, bars, Lr) int64 token rows on the GPU -> (dist, nref, nhyp) int32 (B,) on the GPU."""
    if not (hyp_tokens.is_cuda and ref_tokens.is_cuda):
        raise RuntimeError("piano_a2s_b200.metrics runs on CUDA tensors only; there is no CPU fallback")
    if hyp_tokens.dtype != torch.int64 or ref_tokens.dtype != torch.int64:
        raise TypeError("token tensors must be int64")
    B, bars, Lh = hyp_tokens.shape
    if ref_tokens.shape[:2] != (B, bars):
        raise ValueError("hypothesis and reference must have the same (clips, bars)")
    Lr = ref_tokens.shape[2]
    hyp, ref = hyp_tokens.contiguous(), ref_tokens.contiguous()
    out = torch.empty((3, B), device=hyp.device, dtype=torch.int32)
    lib.pa2s_wer_counts(stream(), ptr(hyp), ptr(ref), B, bars, Lh, Lr, int(eos), _SKIP[0], _SKIP[1], _SEP, ptr(out[0]), ptr(out[1]), ptr(out[2]))
    return out[0], out[1], out[2]


def calculate_wer(pred_tokens: torch.Tensor, target_tokens: torch.Tensor):
    """pretrain.py:216-227 for one staff of a batch: -> (mean WER, [WER per clip]).  pred_tokens = argmax of the staff's
    log-probabilities (kern.greedy_staff_tokens), target_tokens = the padded targets.  A clip whose reference has no words raises
    ValueError, as jiwer does."""
    dist, nref, _ = wer_counts(pred_tokens, target_tokens)
    d, n = torch.stack([dist, nref]).cpu().numpy()               # one copy
    if (n == 0).any():
        raise ValueError("one or more references are empty strings")
    per_clip = (d.astype(np.float64) / n.astype(np.float64)).tolist()
    return float(np.mean(per_clip)), per_clip


def macro_f1(target, pred) -> float:
    """sklearn.metrics.f1_score(target, pred, average="macro") for two integer label lists: unweighted mean over the labels present
    in either list of 2TP / (2TP + FP + FN) (0 when a label is never predicted or never true)."""
    t, p = np.asarray(target).reshape(-1), np.asarray(pred).reshape(-1)
    scores = []
    for c in np.union1d(t, p):
        tp = np.sum((t == c) & (p == c))
        fp = np.sum((t != c) & (p == c))
        fn = np.sum((t == c) & (p != c))
        den = 2 * tp + fp + fn
        scores.append(2.0 * tp / den if den else 0.0)
    return float(np.mean(scores))


def calculate_f1(pred: torch.Tensor, target: torch.Tensor):
    """pretrain.py:236-243 (`caculate_f1`) for key or time signature: (B, bars) predicted / target classes -> (mean, [per clip])."""
    p, t = pred.cpu().numpy(), target.cpu().numpy()
    per_clip = [macro_f1(t[b], p[b]) for b in range(p.shape[0])]
    return float(np.mean(per_clip)), per_clip


def evaluate_batch(predictions, ground_truth):
    """The numbers `on_stage_end` logs for a validation batch (pretrain.py:150-214): WER of both staves, macro-F1 of key and time
    signature, from the model's four outputs and the six target tensors."""
    from .kern import greedy_staff_tokens
    ts, key, up, lo = predictions
    ts_gt, key_gt, up_gt, _, lo_gt, _ = ground_truth
    out = {}
    out["wer_upper"], out["wer_upper_per_clip"] = calculate_wer(greedy_staff_tokens(up)[0], up_gt)
    out["wer_lower"], out["wer_lower_per_clip"] = calculate_wer(greedy_staff_tokens(lo)[0], lo_gt)
    out["key_f1"], out["key_f1_per_clip"] = calculate_f1(greedy_staff_tokens(key)[0], key_gt)
    out["time_f1"], out["time_f1_per_clip"] = calculate_f1(greedy_staff_tokens(ts)[0], ts_gt)
    return out
