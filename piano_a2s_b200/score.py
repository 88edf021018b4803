"""After the path: predicted tokens -> `**kern` files -> MIDI.

The reference's `get_xml_from_target` (data_processing/humdrum.py:841-891) builds, per staff, a one-spine `**kern` text from the
decoded measures, repairs the spine-split tokens, removes duplicated chord notes and then hands the file to EXTERNAL programs:
`tiefix` and `hum2xml` (humextra), music21 for MusicXML/MIDI export and the Java MV2H evaluator (evaluate.py:10-65).  None of
those exist in this image.  What is built here:

* `staff_kern` / `result_kern_files`: the `.krn` text exactly as the reference writes it to `temp/{staff}.krn` BEFORE `tiefix`
  (humdrum.py:846-858) -- pinned against the reference's own `add_split_token`, `Kern`, `eliminate_duplicate_chords`
  (tests/test_score.py, live in the build container + golden fixture);
* `kern_pitch_to_midi`: humdrum.py:600-622, pinned the same way;
* `kern_note_events` + `write_midi`, `write_musicxml`: an OWN reader of that one-spine (optionally two-voice) kern text, a Standard
  MIDI File writer and a MusicXML (score-partwise) writer with the predicted key / time signatures, so that predictions can be
  listened to, engraved or fed to an evaluator without the external tool chain.  Parity with hum2xml / music21's exports is
  UNPINNED (both absent); none of this is on the measured path.
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


# ------------------------------------------------------------------------------------------------------------------
# MusicXML (score-partwise 3.1) from the same kern text: an OWN writer, where the reference goes through hum2xml + music21
# (humdrum.py:859-891); unpinned like write_midi.
# ------------------------------------------------------------------------------------------------------------------
_DIVISIONS = 10080                                    # per quarter note: divisible by 2^5, 3^2, 5, 7 (all reciprocals of the vocabulary)
_TYPES = {1: "whole", 2: "half", 4: "quarter", 8: "eighth", 16: "16th", 32: "32nd", 64: "64th", 128: "128th"}


def _note_type(recip: int):
    """kern reciprocal -> (MusicXML type or None, (actual, normal) tuplet ratio or None)."""
    if recip in _TYPES:
        return _TYPES[recip], None
    for actual, normal in ((3, 2), (5, 4), (7, 4)):
        if recip % actual == 0 and (recip // actual * normal) in _TYPES:
            return _TYPES[recip // actual * normal], (actual, normal)
    return None, None


def _pitch_xml(name: str, acc: str):
    step = name[0].upper()
    octave = 4 + (len(name) - 1) if name[0].islower() else 3 - (len(name) - 1)
    alter = acc.count("#") - acc.count("-")
    s = f"<pitch><step>{step}</step>"
    if alter:
        s += f"<alter>{alter}</alter>"
    return s + f"<octave>{octave}</octave></pitch>"


def _measure_voices(lines):
    """kern lines of ONE measure -> list of voices, each a list of (chord note strings, duration in divisions)."""
    voices = [[], []]
    for line in lines:
        if not line or line.startswith(("*", "!")):
            continue
        for v, chord in enumerate(line.split("\t")[:2]):
            notes = [n for n in chord.split(" ") if _NOTE_RE.fullmatch(n)]
            if not notes:
                continue
            dur = min(_duration(*_NOTE_RE.fullmatch(n).group(2, 3)) for n in notes)
            voices[v].append((notes, int(dur * _DIVISIONS)))
    return [v for v in voices if v] or [[]]


def _part_xml(krn_text, part_id, clef, keys, time_sigs):
    sign, line = clef
    measures, cur = [], []
    for ln in krn_text.splitlines():
        if ln.startswith("**") or ln.startswith("*-"):
            continue
        if ln.startswith("="):
            measures.append(cur)
            cur = []
        else:
            cur.append(ln)
    measures = measures[1:] if measures and not measures[0] else measures      # the text starts with a barline
    out = [f'<part id="{part_id}">']
    key_now, ts_now = None, None
    for i, lines in enumerate(measures):
        out.append(f'<measure number="{i + 1}">')
        attrs = []
        if i == 0:
            attrs.append(f"<divisions>{_DIVISIONS}</divisions>")
        if i < len(keys) and keys[i] != key_now:
            key_now = keys[i]
            attrs.append(f"<key><fifths>{int(key_now)}</fifths></key>")
        if i < len(time_sigs) and time_sigs[i] != ts_now:
            ts_now = time_sigs[i]
            beats, _, beat_type = str(ts_now).partition("/")
            attrs.append(f"<time><beats>{beats}</beats><beat-type>{beat_type or 4}</beat-type></time>")
        if i == 0:
            attrs.append(f"<clef><sign>{sign}</sign><line>{line}</line></clef>")
        if attrs:
            out.append("<attributes>" + "".join(attrs) + "</attributes>")
        voices = _measure_voices(lines)
        for v, events in enumerate(voices):
            if v > 0:
                back = sum(d for _, d in voices[v - 1])
                if back:
                    out.append(f"<backup><duration>{back}</duration></backup>")
            for notes, dur in events:
                for j, n in enumerate(notes):
                    m = _NOTE_RE.fullmatch(n)
                    x = "<note>"
                    if j > 0:
                        x += "<chord/>"
                    x += "<rest/>" if m[4] == "r" else _pitch_xml(m[4], m[5])
                    x += f"<duration>{dur}</duration>"
                    if m[4] != "r":
                        if m[7] in ("]", "_"):
                            x += '<tie type="stop"/>'
                        if m[1] == "[" or m[7] == "_":
                            x += '<tie type="start"/>'
                    x += f"<voice>{v + 1}</voice>"
                    typ, tup = _note_type(int(m[2]))
                    if typ:
                        x += f"<type>{typ}</type>" + "<dot/>" * len(m[3])
                        if tup:
                            x += f"<time-modification><actual-notes>{tup[0]}</actual-notes><normal-notes>{tup[1]}</normal-notes></time-modification>"
                    nota = ""
                    if m[4] != "r" and m[7] in ("]", "_"):
                        nota += '<tied type="stop"/>'
                    if m[4] != "r" and (m[1] == "[" or m[7] == "_"):
                        nota += '<tied type="start"/>'
                    if m[6] == ";":
                        nota += "<fermata/>"
                    if nota:
                        x += f"<notations>{nota}</notations>"
                    out.append(x + "</note>")
        out.append("</measure>")
    out.append("</part>")
    return "\n".join(out)


def write_musicxml(path, pred):
    """`pred` of one evaluation record -> a two-part MusicXML file (upper staff, treble clef; lower staff, bass clef) with the
    predicted key and time signatures where they change (what humdrum.py:859-891 assembles with music21).  Returns the kern texts."""
    files = result_kern_files(pred)
    body = ['<?xml version="1.0" encoding="UTF-8"?>',
            '<score-partwise version="3.1">',
            '<part-list><score-part id="P1"><part-name>Piano</part-name></score-part>'
            '<score-part id="P2"><part-name>Piano</part-name></score-part></part-list>',
            _part_xml(files["upper"], "P1", ("G", 2), files["keys"], files["time_sigs"]),
            _part_xml(files["lower"], "P2", ("F", 4), files["keys"], files["time_sigs"]),
            "</score-partwise>"]
    with open(path, "w", encoding="utf-8") as f:
        f.write("\n".join(body) + "\n")
    return files
