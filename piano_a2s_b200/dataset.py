"""Feature-cache reading in front of the hot path: `SyntheticDataset` / `TrainDataset` / `TestDataset` of
datasets/syn.py:10-170 and `utilities.load` for `.npy` / `.pkl` / `.json` (utilities.py:27-58).

The reference loads every item synchronously in `__getitem__`, builds each target tensor with its own `torch.tensor(...).to(device)`
and copies each padded spectrogram to the device on its own (syn.py:99-121).  Here the files of a whole BATCH are read by a worker
thread into one pinned staging buffer while the previous batch trains (double buffering), and reach the GPU with one asynchronous
H2D copy on a copy stream + one padding kernel (batching.collate); the training stream only waits for the batch's event.
"""
from __future__ import annotations

import json
import os
import pickle
import queue
import threading

import numpy as np
import torch

from . import batching
from .results import TIME_SIGNATURES


def load(path):
    """utilities.load for the three formats of the feature cache."""
    if path.endswith(".npy"):
        return np.load(path)
    if path.endswith(".json"):
        with open(path) as f:
            return json.load(f)
    if path.endswith(".pkl"):
        with open(path, "rb") as f:
            return pickle.load(f)
    raise ValueError(f"unsupported feature file: {path}")


class FeatureFolder:
    """`{feature_folder}/{split}/{version}/spectrogram/{chunk}~{soundfont}.npy` + `.../target/{chunk}.pkl` (syn.py:28-36, 88-121).
    train=True: `TrainDataset` semantics (length = the longest version's list, a random version per item, idx modulo its length);
    train=False: `TestDataset` semantics (every (song, version) pair once)."""

    def __init__(self, feature_folder, split, versions=(0,), train=True, time_sig_list=TIME_SIGNATURES, seed=None):
        self.root, self.split, self.versions, self.train = feature_folder, split, list(versions), train
        self.time_sig_dict = {t: i for i, t in enumerate(time_sig_list)}
        self.song_list = {}
        for v in self.versions:
            folder = os.path.join(feature_folder, str(split), str(v), "spectrogram")
            self.song_list[v] = sorted(s[:-4] for s in os.listdir(folder) if s.endswith(".npy"))
        self.pairs = [(s, v) for v in self.versions for s in self.song_list[v]]
        self.rng = np.random.default_rng(seed)

    def __len__(self):
        return max(len(s) for s in self.song_list.values()) if self.train else len(self.pairs)

    def item(self, idx):
        """-> (spectrogram (n, F) float32, time_sig classes, key sharps, upper bars, lower bars, spectrogram_name, version):
        the tuple batching.collate takes, + the two identifiers `compute_objectives` records (pretrain.py:98)."""
        if self.train:
            v = self.versions[int(self.rng.integers(len(self.versions)))]
            name = self.song_list[v][idx % len(self.song_list[v])]
        else:
            name, v = self.pairs[idx]
        folder = os.path.join(self.root, str(self.split), str(v))
        spec = load(os.path.join(folder, "spectrogram", f"{name}.npy"))
        score = load(os.path.join(folder, "target", f"{name.split('~')[0]}.pkl"))         # bars of [key, time_sig, lower, upper]
        return (spec, [self.time_sig_dict[b[1]] for b in score], [b[0] for b in score], [b[3] for b in score], [b[2] for b in score], name, v)


class BatchLoader:
    """Iterates (spectrogram (B,1,max_frame_num,F), ground_truth[6], names, versions) batches on `device`.  A worker thread reads and
    collates batch i+1 (files -> pinned buffers -> async H2D + pad kernel on its own CUDA stream) while batch i is being consumed."""

    def __init__(self, folder: FeatureFolder, batch_size, max_frame_num, max_length, device, shuffle=False, drop_last=False, depth=2, seed=0):
        self.folder, self.bs, self.max_frame_num, self.max_length = folder, int(batch_size), int(max_frame_num), tuple(max_length)
        self.device, self.shuffle, self.drop_last, self.depth = torch.device(device), shuffle, drop_last, depth
        self.rng = np.random.default_rng(seed)

    def __len__(self):
        n = len(self.folder)
        return n // self.bs if self.drop_last else -(-n // self.bs)

    def _batches(self):
        idx = np.arange(len(self.folder))
        if self.shuffle:
            self.rng.shuffle(idx)
        for i in range(0, len(idx), self.bs):
            chunk = idx[i:i + self.bs]
            if len(chunk) == self.bs or not self.drop_last:
                yield chunk

    def __iter__(self):
        q: queue.Queue = queue.Queue(maxsize=self.depth)
        copy_stream = torch.cuda.Stream(device=self.device)

        def work():
            try:
                with torch.cuda.device(self.device), torch.cuda.stream(copy_stream):
                    for chunk in self._batches():
                        items = [self.folder.item(int(i)) for i in chunk]
                        spec, gt = batching.collate([it[:5] for it in items], self.max_frame_num, self.max_length, self.device)
                        q.put((spec, gt, [it[5] for it in items], [it[6] for it in items], copy_stream.record_event()))
                q.put(None)
            except BaseException as e:                                 # surface reader errors in the consuming thread
                q.put(e)

        th = threading.Thread(target=work, daemon=True)
        th.start()
        while True:
            got = q.get()
            if got is None:
                break
            if isinstance(got, BaseException):
                raise got
            spec, gt, names, versions, ev = got
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(ev)
            for t in [spec] + list(gt):
                t.record_stream(cur)
            yield spec, gt, names, versions
        th.join()
