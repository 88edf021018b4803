"""predicted tokens -> `.krn` text (humdrum.py:846-858, before the external `tiefix`), kern pitch -> MIDI
number (humdrum.py:600-622), and the repo's own kern -> note events -> Standard MIDI File writer."""
import json
import os
import struct

import pytest

from piano_a2s_b200 import score
from refimport import have_reference, import_reference_models

GOLD = os.path.join(os.path.dirname(__file__), "golden", "score_golden.json")


def norm(krn):
    """chords as sorted note lists: the reference dedupes through `set` (hash order)"""
    return [[sorted(ch.split(" ")) for ch in line.split("\t")] for line in krn.split("\n")]


def test_staff_kern_and_pitches_match_golden():
    gold = json.load(open(GOLD))
    for name, case in gold["cases"].items():
        got = score.staff_kern(case["tokens"])
        assert norm(got) == norm(case["krn"]), name
    for p, n in gold["pitches"].items():
        assert score.kern_pitch_to_midi(p) == n, p


@pytest.mark.skipif(not have_reference(), reason="reference model not available")
def test_staff_kern_matches_live_reference_on_random_token_sequences():
    import random
    ref = import_reference_models()
    g = ref.LabelsMultiple.__init__.__globals__
    labels = ref.LabelsMultiple(extended=True)
    rnd = random.Random(7)
    vocab = len(labels.labels)
    for trial in range(40):
        measures = []
        for _ in range(5):
            n = rnd.randrange(0, 40)
            # what a (badly trained) model emits: arbitrary ids, including <sos>/<pad>, tabs, newlines, ids beyond the label list
            measures.append([rnd.randrange(0, vocab + 3) for _ in range(n)])
        kern_data = ["**kern"] + ["".join(labels.decode(m)) for m in measures]
        kern_data = "\n=\n".join(kern_data) + "\n="
        kern_data = "\n".join(g["add_split_token"](kern_data.split("\n")))
        try:
            kern = g["Kern"](data=kern_data + "\n*-\n")
            kern = g["eliminate_duplicate_chords"](kern)
            want = kern.dump()
        except Exception as e:                      # the reference may choke on garbage; then this implementation is free
            print("reference raised", type(e).__name__)
            continue
        assert norm(score.staff_kern(measures)) == norm(want), trial
    for p in ("c", "CC#", "bbb-", "F", "eeee", "AAA#"):
        assert score.kern_pitch_to_midi(p) == g["kern_to_midi"](p)
    for text in ("4c\t4cc\n4d\t4dd", "[2c 2e\n2c] 2e_\n.\n8.f#;", "1r", "16BB- 16D 16F\t2.ccc#"):
        assert score.encode_kern(text) == labels.encode(text)


def test_note_events_and_midi_file(tmp_path):
    enc = score.encode_kern
    upper = [enc("4c\t4cc\n4d\t4dd\n2e\t2ee"), enc("[2c [2e\n2c] 2e_"), enc("4e]\n4r\n2g")]
    lower = [enc("1C"), enc("2.D\n4r"), enc("1E")]
    pred = [(0, "4/4", lo, up) for lo, up in zip(lower, upper)]
    files = score.result_kern_files(pred)
    ev = score.kern_note_events(files["upper"])
    # voice 1: c d e(2), voice 2: cc dd ee(2); measure 2 starts at 4: tied c (2+2), e tied over the barline: 2 + 2 + 1
    assert (0, 1, 60) in [(float(o), float(d), p) for o, d, p in ev]
    assert (0, 1, 72) in [(float(o), float(d), p) for o, d, p in ev]
    assert (4.0, 4.0, 60) in [(float(o), float(d), p) for o, d, p in ev]
    assert (4.0, 5.0, 64) in [(float(o), float(d), p) for o, d, p in ev]
    assert (10.0, 2.0, 67) in [(float(o), float(d), p) for o, d, p in ev]
    lo = score.kern_note_events(files["lower"])
    assert [(float(o), float(d), p) for o, d, p in lo] == [(0.0, 4.0, 48), (4.0, 3.0, 50), (8.0, 4.0, 52)]
    path = tmp_path / "pred.mid"
    score.result_to_midi(pred, str(path))
    data = path.read_bytes()
    assert data[:4] == b"MThd" and struct.unpack(">IHHH", data[4:14]) == (6, 1, 3, 480)
    assert data.count(b"MTrk") == 3
    # every note-on has its note-off
    assert sum(1 for i in range(len(data) - 2) if data[i] == 0x91 and data[i + 2] == 80) == len(lo)


def test_musicxml_file(tmp_path):
    import xml.etree.ElementTree as ET
    enc = score.encode_kern
    upper = [enc("4c\t4cc\n4d\t4dd\n2e\t2ee"), enc("[2c [2e\n2c] 2e_"), enc("4e]\n4r\n12g\n12a\n12b\n8.cc#;\n16dd-")]
    lower = [enc("1C"), enc("2.D\n4r"), enc("1EE-")]
    pred = [(-1, "4/4", lower[0], upper[0]), (-1, "4/4", lower[1], upper[1]), (2, "3/4", lower[2], upper[2])]
    path = tmp_path / "pred.xml"
    score.write_musicxml(str(path), pred)
    root = ET.parse(str(path)).getroot()
    assert root.tag == "score-partwise" and [p.get("id") for p in root.findall("part")] == ["P1", "P2"]
    up, lo = root.findall("part")
    assert len(up.findall("measure")) == 3 and len(lo.findall("measure")) == 3
    D = score._DIVISIONS
    m1 = up.findall("measure")[0]
    assert m1.find("attributes/key/fifths").text == "-1" and m1.find("attributes/time/beats").text == "4"
    assert m1.find("attributes/clef/sign").text == "G" and lo.find("measure/attributes/clef/sign").text == "F"
    assert m1.find("backup/duration").text == str(4 * D)                       # second voice starts over
    assert [n.find("voice").text for n in m1.findall("note")] == ["1", "1", "1", "2", "2", "2"]
    assert [int(n.find("duration").text) for n in m1.findall("note")] == [D, D, 2 * D, D, D, 2 * D]
    m2 = up.findall("measure")[1]
    assert m2.find("attributes") is None                                       # nothing changed
    notes = m2.findall("note")
    assert notes[1].find("chord") is not None and notes[0].find("tie").get("type") == "start"
    assert [t.get("type") for t in notes[3].findall("tie")] == ["stop", "start"]
    m3 = up.findall("measure")[2]
    assert m3.find("attributes/key/fifths").text == "2" and m3.find("attributes/time/beat-type").text == "4"
    n3 = m3.findall("note")
    assert n3[1].find("rest") is not None
    trip = n3[2]
    assert trip.find("type").text == "eighth" and trip.find("time-modification/actual-notes").text == "3" and int(trip.find("duration").text) == D // 3
    assert n3[5].find("pitch/alter").text == "1" and n3[5].find("dot") is not None and n3[5].find("notations/fermata") is not None
    assert n3[6].find("pitch/step").text == "D" and n3[6].find("pitch/alter").text == "-1" and n3[6].find("pitch/octave").text == "5"
    low3 = lo.findall("measure")[2].find("note/pitch")
    assert (low3.find("step").text, low3.find("alter").text, low3.find("octave").text) == ("E", "-1", "2")
