"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the evaluation metrics of pretrain.py:216-243.

* `calculate_wer` (pretrain.py:216-227): per clip `jiwer.wer(target, pred)` of " \\n = \\n ".join(idx2string(bar) for bar in bars).
  jiwer==3.0.3 (environment.yaml:47) is a third-party dependency that is NOT vendored under /root/reference and not installed in
  this image, and the reference holds no test or golden value for it: **parity unpinned**.  What is restated here is jiwer's
  published algorithm [recalled]: default transform `Compose([RemoveMultipleSpaces(), Strip(), ReduceToListOfListOfWords()])`
  -- `re.sub(r"\\s\\s+", " ", s)`, `s.strip()`, split on " " dropping empty words -- then
  wer = (S + D + I) / (H + S + D) = Levenshtein(reference words, hypothesis words) / len(reference words); an empty reference
  raises ValueError.  The strings are built exactly as the reference builds them (idx2string, pretrain.py:229-234), the transform
  is applied to the STRINGS (so the whitespace labels really go through the regex), and the distance is the textbook
  row-by-row dynamic programme.
* `caculate_f1` (pretrain.py:236-243): `sklearn.metrics.f1_score(target, pred, average="macro")`; scikit-learn IS installed
  here, so tests/test_metrics.py pins `f1_macro` against it directly.
Only tests/ may import this file.
"""
import re


def idx2string(idx_seq, labels_map_inv):
    return " ".join(labels_map_inv[int(i)] for i in idx_seq)


def unpad(seq, eos):
    seq = [int(t) for t in seq]
    return seq[:seq.index(eos)] if eos in seq else seq


def jiwer_words(s):
    s = re.sub(r"\s\s+", " ", s)
    s = s.strip()
    return [w for w in s.split(" ") if len(w) >= 1]


def levenshtein(ref, hyp):
    prev = list(range(len(hyp) + 1))
    for i in range(1, len(ref) + 1):
        cur = [i] + [0] * len(hyp)
        for j in range(1, len(hyp) + 1):
            cur[j] = min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ref[i - 1] != hyp[j - 1]))
        prev = cur
    return prev[len(hyp)]


def clip_wer(pred_bars, target_bars, labels_map_inv, eos):
    """(bars, L) predicted / target token rows of ONE clip -> (wer, distance, #reference words, #hypothesis words)."""
    pred = " \n = \n ".join(idx2string(unpad(b, eos), labels_map_inv) for b in pred_bars)
    target = " \n = \n ".join(idx2string(unpad(b, eos), labels_map_inv) for b in target_bars)
    rw, hw = jiwer_words(target), jiwer_words(pred)
    if not rw:
        raise ValueError("one or more references are empty strings")
    d = levenshtein(rw, hw)
    return d / len(rw), d, len(rw), len(hw)


def f1_macro(target, pred):
    labs = sorted(set(int(x) for x in target) | set(int(x) for x in pred))
    tot = 0.0
    for c in labs:
        tp = sum(1 for t, p in zip(target, pred) if t == c and p == c)
        fp = sum(1 for t, p in zip(target, pred) if t != c and p == c)
        fn = sum(1 for t, p in zip(target, pred) if t == c and p != c)
        tot += (2.0 * tp / (2 * tp + fp + fn)) if (2 * tp + fp + fn) else 0.0
    return tot / len(labs)
