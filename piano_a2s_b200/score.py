"""After the path: predicted tokens -> `**kern` files -> MIDI.

The reference's `get_xml_from_target` (data_processing/humdrum.py:841-891) builds, per staff, a one-spine `**kern` text from the
decoded measures, repairs the spine-split tokens, removes duplicated chord notes and then hands the file to EXTERNAL programs:
`tiefix` and `hum2xml` (humextra), music21 for MusicXML/MIDI export and the Java MV2H evaluator (evaluate.py:10-65).  None of
those exist in this image.  What is built here:

* `staff_kern` / `result_kern_files`: the `.krn` text exactly as the reference writes it to `temp/{staff}.krn` BEFORE `tiefix`
  (humdrum.py:846-858) -- pinned against the reference's own `add_split_token`, `Kern`, `eliminate_duplicate_chords`
  (tests/test_score.py, live in the build container + golden fixture);
* `kern_pitch_to_midi`: humdrum.py:600-622, pinned the same way;
* `kern_note_events` + `write_midi`: an OWN reader of that one-spine (optionally two-voice) kern text and a Standard MIDI File
  writer, so that predictions can be listened to / fed to an evaluator without the external tool chain.  Parity with
  music21's MIDI export is UNPINNED (music21, hum2xml absent); it is not on the measured path.
"""
from __future__ import annotations

import re
import struct
from fractions import Fraction

from .models import labels as _labels

_NOTE_RE = re.compile(r"(\[?)(\d+)(\.*)([a-gA-Gr]{1,4})([\-#]*)(;?)([\]_]?)")


def decode_measure(tokens, labels=_labels):
    """LabelsMultiple.decode (humdrum.py:130-132) joined: ids without a label vanish, `<b>` becomes the chord-note separator."""
    out = []
    for t in tokens:
        s = labels.labels_map_inv.get(int(t))
        if s:
            out.append(" " if s == "<b>" else s)
    return "".join(out)


_ENC_RE = re.compile(r"(\[?)(\d+\.*)([a-gA-Gr]{1,4}[\-#]*)(;?)([\]_]?)")


def encode_kern(text, labels=_labels):
    """LabelsMultiple.encode (humdrum.py:103-128): kern lines -> token ids.  A note splits into [tie start][duration][pitch][fermata]
    [tie end]; notes of a chord are separated by `<b>`, voices by the tab token, lines by the newline token."""
    lm = labels.labels_map
    tokens = []
    for line in text.splitlines():
        for chord in line.split("\t"):
            for note in chord.split(" "):
                if len(note) == 1:
                    tokens.append(lm[note])
                else:
                    m = _ENC_RE.fullmatch(note)
                    if not m:
                        raise ValueError(f"item {note!r} in {line!r} is not a kern note")
                    tokens.extend(lm[x] for x in m.groups() if x)
                tokens.append(lm["<b>"])
            if tokens[-1] == lm["<b>"]:
                tokens.pop()
            tokens.append(lm["\t"])
        tokens[-1] = lm["\n"]
    tokens.pop()
    return tokens


def add_split_tokens(lines):
    """humdrum.py:760-772: `*^` before the first two-voice line after a one-voice line, `*v\\t*v` where the voices merge again;
    comment lines are dropped."""
    out, prev = [], 1
    for line in lines:
        if line.startswith("!"):
            continue
        cur = len(line.split("\t"))
        if cur == 2 and prev == 1:
            out.append("*^")
        elif cur == 1 and prev == 2:
            out.append("*v\t*v")
        out.append(line)
        prev = cur
    return out


def dedupe_chords(lines):
    """humdrum.py:821-839: within a chord (space-separated notes of one voice) every distinct note once, empty items dropped.
    The reference goes through `set`, so ITS order of the surviving notes is hash order; here first occurrence wins."""
    out = []
    for line in lines:
        if line.startswith("=") or line.startswith("*"):
            out.append(line)
            continue
        voices = []
        for chord in line.split("\t"):
            notes = chord.split(" ")
            if len(notes) > 1:
                seen = []
                for n in notes:
                    if n and n not in seen:
                        seen.append(n)
                voices.append(" ".join(seen))
            else:
                voices.append(notes[0])
        out.append("\t".join(voices))
    return out


def staff_kern(measures, labels=_labels):
    """Token lists of the measures of ONE staff -> the `.krn` text of humdrum.py:846-858 (before `tiefix`)."""
    data = ["**kern"] + [decode_measure(m, labels) for m in measures]
    text = "\n=\n".join(data) + "\n="
    lines = add_split_tokens(text.split("\n"))
    # Kern(data=...): header = everything up to and including the `**kern` line, footer from the first `*-` line
    full = ("\n".join(lines) + "\n*-\n").splitlines()
    begin, end = 0, 0
    for i, line in enumerate(full):
        if line.startswith("**"):
            begin = i + 1
        if line.startswith("*-"):
            end = i
            break
    header, body, footer = full[:begin], full[begin:end], full[end:]
    return "\n".join(header + dedupe_chords(body) + footer)


def result_kern_files(pred):
    """One evaluation record's `pred` (results.ResultRecorder / pretrain.py:189-214: per measure (key, time signature, lower tokens,
    upper tokens)) -> {"lower": krn text, "upper": krn text, "keys": [...], "time_sigs": [...]}, the inputs of humdrum.py:841-891."""
    return {"lower": staff_kern([m[2] for m in pred]), "upper": staff_kern([m[3] for m in pred]),
            "keys": [m[0] for m in pred], "time_sigs": [m[1] for m in pred]}


def kern_pitch_to_midi(kern_note: str) -> int:
    """humdrum.py:600-622: `c` = 60, upper-case letters go down an octave per repetition, lower-case up; one trailing # / -."""
    base = {"c": 60, "d": 62, "e": 64, "f": 65, "g": 67, "a": 69, "b": 71, "C": 48, "D": 50, "E": 52, "F": 53, "G": 55, "A": 57, "B": 59}
    n = 0
    if kern_note[-1] == "#":
        n, kern_note = 1, kern_note[:-1]
    elif kern_note[-1] == "-":
        n, kern_note = -1, kern_note[:-1]
    n += base[kern_note[0]]
    octaves = 12 * (len(kern_note) - 1)
    return n - octaves if kern_note[0].isupper() else n + octaves


def _duration(recip: str, dots: str) -> Fraction:
    """kern reciprocal duration in quarter notes: `4` = 1, `8.` = 3/4, `0` (breve) = 8."""
    r = int(recip)
    d = Fraction(8) if r == 0 else Fraction(4, r)
    total, add = d, d
    for _ in dots:
        add /= 2
        total += add
    return total


def kern_note_events(krn_text: str):
    """One-spine kern text (with `*^` / `*v` two-voice sections) -> [(onset, duration, midi pitch)] in quarter notes, tied notes
    (`[` ... `_` ... `]`) merged.  Every voice keeps its own clock inside a measure; a barline advances all voices to the longest."""
    events, open_ties = [], {}
    clock = [Fraction(0)]
    for line in krn_text.splitlines():
        if not line or line.startswith(("!", "**", "*-")):
            continue
        if line.startswith("*^"):
            clock = [clock[0], clock[0]]
            continue
        if line.startswith("*v"):
            clock = [max(clock)]
            continue
        if line.startswith("*"):
            continue
        if line.startswith("="):
            clock = [max(clock)] * len(clock)
            continue
        for v, chord in enumerate(line.split("\t")):
            if v >= len(clock):
                clock.append(clock[-1])
            step = None
            for note in chord.split(" "):
                m = _NOTE_RE.fullmatch(note)
                if not m:
                    continue                                   # `.` place holders and anything the model garbled
                dur = _duration(m[2], m[3])
                step = dur if step is None else min(step, dur)
                if m[4] == "r":
                    continue
                pitch = kern_pitch_to_midi(m[4] + m[5][:1])
                if m[7] in ("_", "]") and pitch in open_ties:
                    i = open_ties[pitch]
                    events[i] = (events[i][0], events[i][1] + dur, pitch)
                    if m[7] == "]":
                        del open_ties[pitch]
                    continue
                events.append((clock[v], dur, pitch))
                if m[1] == "[":
                    open_ties[pitch] = len(events) - 1
            if step is not None:
                clock[v] += step
    return events


def _vlq(n: int) -> bytes:
    out = [n & 0x7F]
    n >>= 7
    while n:
        out.append((n & 0x7F) | 0x80)
        n >>= 7
    return bytes(reversed(out))


def write_midi(path, staves, ticks_per_quarter=480, tempo_bpm=120, velocity=80):
    """Standard MIDI File, format 1: one tempo track + one track per staff; `staves` = list of kern_note_events lists."""
    def track(body: bytes) -> bytes:
        body += b"\x00\xff\x2f\x00"
        return b"MTrk" + struct.pack(">I", len(body)) + body
    us = int(round(60e6 / tempo_bpm))
    chunks = [track(b"\x00\xff\x51\x03" + struct.pack(">I", us)[1:])]
    for ch, events in enumerate(staves):
        msgs = []
        for onset, dur, pitch in events:
            if not 0 <= pitch <= 127 or dur <= 0:
                continue
            t0, t1 = int(round(onset * ticks_per_quarter)), int(round((onset + dur) * ticks_per_quarter))
            msgs.append((t0, 1, bytes([0x90 | (ch & 15), pitch, velocity])))
            msgs.append((max(t1, t0 + 1), 0, bytes([0x80 | (ch & 15), pitch, 0])))
        msgs.sort(key=lambda m: (m[0], m[1]))
        body, now = b"", 0
        for t, _, data in msgs:
            body += _vlq(t - now) + data
            now = t
        chunks.append(track(body))
    data = b"MThd" + struct.pack(">IHHH", 6, 1, len(chunks), ticks_per_quarter) + b"".join(chunks)
    with open(path, "wb") as f:
        f.write(data)
    return len(data)


def result_to_midi(pred, path):
    """`pred` of one evaluation record -> a two-track MIDI file (upper staff, lower staff); returns the kern texts used."""
    files = result_kern_files(pred)
    write_midi(path, [kern_note_events(files["upper"]), kern_note_events(files["lower"])])
    return files
