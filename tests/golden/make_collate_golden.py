"""Writes tests/golden/collate_golden.npz from the UNMODIFIED reference dataset methods (datasets/syn.py:38-74), imported in the
build container with the third-party modules the methods never touch stubbed out.  Run: python tests/golden/make_collate_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from refimport import reference_dataset_stub  # noqa: E402


def cases():
    rng = np.random.RandomState(5)
    specs = [rng.rand(n, 12).astype(np.float32) for n in (0, 1, 7, 30, 31, 45)]       # max_frame_num = 30: empty, short, exact, over
    scores = [[list(rng.randint(0, 144, size=n)) for n in ns] for ns in ((0, 3, 9), (10, 11, 4), (1, 9, 25))]   # max_length = 10
    keys = [[-6, 0, 7], [1, -1, 3], [0, 0, 0]]
    return specs, scores, keys


def main():
    ds = reference_dataset_stub(max_frame_num=30)
    specs, scores, keys = cases()
    out = {}
    for i, s in enumerate(specs):
        out[f"spec_in_{i}"] = s
        try:
            out[f"spec_out_{i}"] = ds.pad_spectrogram(s).numpy()
        except RuntimeError:                       # more frames than max_frame_num: the reference raises (syn.py:56-57)
            out[f"spec_raises_{i}"] = np.asarray(1)
    for i, sc in enumerate(scores):
        rows, ln = ds.pad_score(sc, 10)
        out[f"score_in_{i}"] = np.asarray([np.pad(np.asarray(m, dtype=np.int64), (0, 32 - len(m)), constant_values=-1) for m in sc])
        out[f"score_out_{i}"] = rows.numpy()
        out[f"score_len_{i}"] = ln.numpy()
        out[f"key_out_{i}"] = ds.key_to_int(keys[i]).numpy()
        out[f"key_in_{i}"] = np.asarray(keys[i], dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "collate_golden.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
