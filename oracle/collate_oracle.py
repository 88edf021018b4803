"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy) of the reference's per-item padding / collation, the row a2 of
SURVEY section 8 and the target half of `__getitem__`:

* `pad_spectrogram`      datasets/syn.py:46-58   (identical copy in datasets/asap.py:338-350)
* `pad_single_measure`   datasets/syn.py:67-74   (asap.py:359-366)
* `pad_score`            datasets/syn.py:60-65   (asap.py:352-357)
* `key_to_int`           datasets/syn.py:38-40   (asap.py:330-332): sharps + 6
* batch stacking         torch DataLoader default collate over `__getitem__` (syn.py:88-121)

Pinned: live against the unmodified reference methods in the build container (tests/test_collate.py, reference imported
with its unused third-party imports stubbed) and against tests/golden/collate_golden.npz written from the reference by
tests/golden/make_collate_golden.py.  Only tests/ may import this file; the product path (piano_a2s_b200/batching.py) does not.
"""
import numpy as np


def pad_spectrogram(spectrogram, max_frame_num, truncate=False):
    """(T, F) -> (1, max_frame_num, F) float32: the T frames, zeros after (syn.py:53-58).  As written the reference assigns the
    WHOLE spectrogram to the first min(T, max) rows (syn.py:56-57), i.e. it raises for T > max_frame_num (the loaders filter
    clips to <= 12 s, asap.py:101, so it never sees one); `truncate=True` is the evident intent (keep the first max frames)."""
    spec = np.asarray(spectrogram).astype(np.float32)
    out = np.zeros((max_frame_num, spec.shape[-1]), dtype=np.float32)
    n = min(spec.shape[0], max_frame_num)
    if spec.shape[0] > max_frame_num and not truncate:
        raise RuntimeError("spectrogram has %d frames, more than max_frame_num = %d" % (spec.shape[0], max_frame_num))
    out[:n] = spec[:n]
    return out[None]


def pad_single_measure(measure, max_length, pad, eos):
    """token list -> (max_length,) int64: tokens (truncated), <eos> right after them if there is room, <pad> elsewhere (syn.py:67-74)."""
    row = np.full((max_length,), pad, dtype=np.int64)
    m = list(measure)[:max_length]
    row[:len(m)] = np.asarray(m, dtype=np.int64)
    if len(m) < max_length:
        row[len(m)] = eos
    return row


def pad_score(score, max_length, pad, eos):
    """list of bars -> ((bars, max_length) int64, (bars,) int64 lengths = min(len, max_length)) (syn.py:60-65)."""
    rows = np.stack([pad_single_measure(m, max_length, pad, eos) for m in score]) if len(score) else np.zeros((0, max_length), np.int64)
    lengths = np.asarray([min(len(m), max_length) for m in score], dtype=np.int64)
    return rows, lengths


def key_to_int(key_signatures):
    return np.asarray(key_signatures, dtype=np.int64) + 6


def collate(items, max_frame_num, max_length, pad, eos):
    """items: [(spectrogram (T,F), time_sig ints, key sharps, upper bars, lower bars)] -> the stacked batch the trainer sees:
    spectrogram (B,1,max_frame_num,F) and the six target arrays of models.py:26-31."""
    spec = np.stack([pad_spectrogram(it[0], max_frame_num) for it in items])
    ts = np.stack([np.asarray(it[1], dtype=np.int64) for it in items])
    key = np.stack([key_to_int(it[2]) for it in items])
    up = [pad_score(it[3], max_length[0], pad, eos) for it in items]
    lo = [pad_score(it[4], max_length[1], pad, eos) for it in items]
    return spec, [ts, key, np.stack([u[0] for u in up]), np.stack([u[1] for u in up]),
                  np.stack([l[0] for l in lo]), np.stack([l[1] for l in lo])]
