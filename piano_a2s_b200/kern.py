"""Token post-processing right after the hot path: greedy tokens -> `unpad` -> `**kern` token strings.

Mirrors the per-sample Python of the reference trainer -- `pred = outs[b].argmax(-1)`, `unpad(p).tolist()`
(pretrain.py:97-117, 245-249; finetune.py:86-108) and `idx2string` (pretrain.py:229-234) -- but runs the argmax and the
first-<eos> search for every (clip, bar) sequence of a staff in ONE libpa2s launch and reads the result back with one
device->host copy per staff instead of one `.nonzero()`/`.cpu()` sync per sequence.
"""
from __future__ import annotations

import torch

from ._lib import lib, ptr, stream
from .models import EOS, labels


def greedy_staff_tokens(logp: torch.Tensor, eos: int = EOS):
    """(..., L, V) log-probabilities on the GPU -> (tokens int64 (..., L), lengths int32 (...)) on the GPU.
    tokens = argmax over V (lowest index on ties), lengths = index of the first <eos> (L when there is none)."""
    if not logp.is_cuda:
        raise RuntimeError("piano_a2s_b200.kern runs on CUDA tensors only; there is no CPU fallback")
    if logp.dtype != torch.float32:
        raise TypeError("log-probabilities must be float32")
    x = logp.contiguous()
    lead, L, V = x.shape[:-2], x.shape[-2], x.shape[-1]
    nseq = 1
    for d in lead:
        nseq *= d
    tokens = torch.empty(lead + (L,), device=x.device, dtype=torch.int64)
    lengths = torch.empty(lead, device=x.device, dtype=torch.int32)
    if nseq:
        lib.pa2s_greedy_tokens(stream(), ptr(x), nseq, L, V, int(eos), ptr(tokens), ptr(lengths))
    return tokens, lengths


def unpad(full_seq, eos: int = EOS):
    """pretrain.py:245-249 for ONE token row (tensor or list): everything before the first <eos>, as a python list."""
    seq = full_seq.tolist() if hasattr(full_seq, "tolist") else list(full_seq)
    return seq[:seq.index(eos)] if eos in seq else seq


def idx2string(idx_seq):
    """pretrain.py:229-234."""
    return " ".join(labels.labels_map_inv[int(i)] for i in idx_seq)


def greedy_tokens(predictions):
    """The four model outputs -> what `compute_objectives` records for evaluation (pretrain.py:97-117):
    {"upper": [clip][bar] -> unpadded token list, "lower": ..., "key": [clip][bar] -> int, "time_sig": ...}."""
    ts, key, up, lo = predictions
    out = {}
    for name, lp in (("upper", up), ("lower", lo)):
        tok, ln = greedy_staff_tokens(lp)
        tok, ln = tok.cpu().tolist(), ln.cpu().tolist()          # one copy per staff
        out[name] = [[row[:n] for row, n in zip(tb, nb)] for tb, nb in zip(tok, ln)]
    for name, lp in (("key", key), ("time_sig", ts)):
        out[name] = greedy_staff_tokens(lp)[0].cpu().tolist()
    return out


def kern_strings(predictions, bar_separator=" \n = \n "):
    """Per clip and staff the string the reference scores with WER (`calculate_wer`, pretrain.py:216-227):
    bars joined by " \\n = \\n ", tokens by " "."""
    toks = greedy_tokens(predictions)
    return {s: [bar_separator.join(idx2string(b) for b in clip) for clip in toks[s]] for s in ("upper", "lower")}
