"""Evaluation records and result files right after the hot path: what `ASR.compute_objectives` keeps for every
validation / test clip (pretrain.py:95-117) and what `on_stage_end` writes to `results/{split}/{id}.json` (pretrain.py:189-214) --
the only input of the reference's evaluate.py.

The reference does this per sample with `.argmax()`, `.nonzero()`, `.item()` and `.cpu()` syncs; here a batch's token ids come from
one libpa2s launch per staff (kern.greedy_tokens), the word error counts from one launch per staff (metrics.wer_counts) and the
files are plain json written on the host.  Field names, nesting, value conventions (key = class - 6 sharps, time signature as the
string of `time_signature_list.json`, bars as [key, time_sig, lower, upper]) are the reference's.
"""
from __future__ import annotations

import json
import os

from . import kern, metrics

# data_processing/metadata/time_signature_list.json (7 entries -> num_time_sig: 7, pretrain.yaml:90)
TIME_SIGNATURES = ["4/4", "3/4", "2/4", "6/8", "2/2", "12/8", "3/8"]


class ResultRecorder:
    """Accumulates the records of an evaluation stage; `ids` follow the reference: '~'.join([str(version), song_name]) with
    song_name = '<chunk>~<soundfont>' (pretrain.py:98, 200)."""

    def __init__(self, time_sig_list=TIME_SIGNATURES):
        self.time_sig_list = list(time_sig_list)
        self.upper_pred, self.upper_target, self.lower_pred, self.lower_target = {}, {}, {}, {}
        self.key_pred, self.key_target, self.time_sig_pred, self.time_sig_target = {}, {}, {}, {}
        self.wer_upper, self.wer_lower, self.key_f1, self.time_f1 = {}, {}, {}, {}

    def add_batch(self, predictions, ground_truth, song_names, versions):
        """predictions: the model's four outputs; ground_truth: the six target tensors; one id per clip (pretrain.py:95-117)."""
        ts, key, up, lo = predictions
        ts_gt, key_gt, up_gt, _, lo_gt, _ = ground_truth
        toks = kern.greedy_tokens(predictions)
        ev = metrics.evaluate_batch(predictions, ground_truth)
        up_t, lo_t = up_gt.cpu().tolist(), lo_gt.cpu().tolist()
        key_t, ts_t = key_gt.cpu().tolist(), ts_gt.cpu().tolist()
        for b, (name, v) in enumerate(zip(song_names, versions)):
            cid = "~".join([str(int(v)), name])
            self.upper_pred[cid] = toks["upper"][b]
            self.lower_pred[cid] = toks["lower"][b]
            self.upper_target[cid] = [kern.unpad(r) for r in up_t[b]]
            self.lower_target[cid] = [kern.unpad(r) for r in lo_t[b]]
            self.key_pred[cid], self.key_target[cid] = toks["key"][b], key_t[b]
            self.time_sig_pred[cid], self.time_sig_target[cid] = toks["time_sig"][b], ts_t[b]
            self.wer_upper[cid], self.wer_lower[cid] = ev["wer_upper_per_clip"][b], ev["wer_lower_per_clip"][b]
            self.key_f1[cid], self.time_f1[cid] = ev["key_f1_per_clip"][b], ev["time_f1_per_clip"][b]

    def stage_stats(self):
        """WER (mean of the two staves' mean WER), key / time-signature macro-F1: the numbers `on_stage_end` logs (pretrain.py:160-178)."""
        n = max(len(self.upper_pred), 1)
        mean = lambda d: sum(d.values()) / n
        return {"wer_upper": mean(self.wer_upper), "wer_lower": mean(self.wer_lower), "WER": (mean(self.wer_upper) + mean(self.wer_lower)) / 2,
                "key_f1": mean(self.key_f1), "time_f1": mean(self.time_f1)}

    def result(self, cid, feature_folder, split, composer=None):
        """The dict pretrain.py:191-212 saves for one clip."""
        pred = [[self.key_pred[cid][i] - 6, self.time_sig_list[self.time_sig_pred[cid][i]], self.lower_pred[cid][i], self.upper_pred[cid][i]]
                for i in range(len(self.upper_pred[cid]))]
        version, chunk_name, soundfont = cid.split("~")
        if composer is None:
            info_path = os.path.join(feature_folder, split, version, "info", f"{chunk_name}.json")
            with open(info_path) as f:
                composer = json.load(f)["composer"]
        return {"style": "classical" if chunk_name[0].islower() else "pop", "soundfont": soundfont, "composer": composer,
                "target_path": os.path.join(feature_folder, split, version, "target", f"{chunk_name}.pkl"), "pred": pred,
                "wer_upper": self.wer_upper[cid], "wer_lower": self.wer_lower[cid], "key_f1": self.key_f1[cid], "time_f1": self.time_f1[cid]}

    def write(self, output_folder, feature_folder, split, composers=None):
        """results/{split}/{id}.json for every recorded clip (pretrain.py:189-214) -> list of paths."""
        out_dir = os.path.join(output_folder, "results", split)
        os.makedirs(out_dir, exist_ok=True)
        paths = []
        for cid in self.upper_pred:
            res = self.result(cid, feature_folder, split, None if composers is None else composers[cid])
            path = os.path.join(out_dir, f"{cid}.json")
            with open(path, "w") as f:
                json.dump(res, f)
            paths.append(path)
        return paths
