"""Generates tests/golden/score_golden.json: inputs (token measures of one staff) and the `.krn` text the REFERENCE's own code
(data_processing/humdrum.py:846-858: LabelsMultiple.decode, add_split_token, Kern, eliminate_duplicate_chords) produces for them,
plus kern_to_midi (humdrum.py:600-622) on a list of pitches.  Run in the build container (needs /root/reference):
    python tests/golden/make_score_golden.py
The reference removes duplicated chord notes through `set`, so the order of the notes of such a chord is hash order: the
fixture stores the text as produced and the test compares chords as sorted note lists."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from refimport import import_reference_models  # noqa: E402

CASES = {
    "one_voice": ["4c\n4d\n4e\n4f", "2g\n2cc", "1r"],
    "chords_and_dups": ["4c 4e 4g\n4c 4c 4e\n2G 2B 2d 2B", "4.f# 4.a\n8b-\n2cc# 2ee 2cc#"],
    "two_voices": ["4c\t4cc\n4d\t4dd\n2e\t2ee", "4f\n4g\n2a", "2c 2e\t4g\n.\t4a\n2r\t2b"],
    "ties": ["[2c\n2c_", "4c]\n4r\n[2e 2g", "2e] 2g]\n2r"],
    "split_then_merge_then_split": ["4c\t4e", "4d", "4e\t4g\n4f\t4a", "1c"],
    "empty_measure": ["4c\n4d", "", "2e"],
}
PITCHES = ["c", "cc", "ccc#", "C", "CC", "BBB#", "AAA", "b-", "f#", "ffff", "CCC", "d-", "E#", "gg-"]


def main():
    ref = import_reference_models()
    g = ref.LabelsMultiple.__init__.__globals__
    labels = ref.LabelsMultiple(extended=True)
    out = {"cases": {}, "pitches": {p: g["kern_to_midi"](p) for p in PITCHES}}
    for name, measures in CASES.items():
        toks = [labels.encode(m) if m else [] for m in measures]
        kern_data = ["**kern"]
        for t in toks:
            kern_data.append("".join(labels.decode(t)))
        kern_data = "\n=\n".join(kern_data) + "\n="
        kern_data = "\n".join(g["add_split_token"](kern_data.split("\n")))
        kern = g["Kern"](data=kern_data + "\n*-\n")
        kern = g["eliminate_duplicate_chords"](kern)
        out["cases"][name] = {"tokens": toks, "krn": kern.dump()}
    with open(os.path.join(HERE, "score_golden.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", len(out["cases"]), "cases")


if __name__ == "__main__":
    main()
