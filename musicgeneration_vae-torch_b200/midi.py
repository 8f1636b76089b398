"""Standard-MIDI-file tail of the sampling script (reference: maker_bar.py:46-54), without pypianoroll.

The reference pads the generated [T,60] roll to 128 pitches with ``np.pad(note, [[0,0],[27,41]])`` (pitch 27 is the
lowest cell), multiplies by 127 and writes it through ``pypianoroll.Multitrack(tracks=[Track(note, name='piano')],
beat_resolution=24, name='test').write(path)``.  pypianoroll is a third-party dependency that is neither vendored nor
pinned by the reference (no requirements file; the ``beat_resolution`` keyword dates it to the 0.5.x series) and is
absent from this image, so its published conversion is restated here:

  * a note is a maximal run of consecutive non-zero time steps of one pitch (``Multitrack.to_pretty_midi`` binarises
    the roll, pads it with a zero step on both sides and takes the rising / falling edges of ``np.diff``);
  * every note gets the constant velocity 100 (``constant_velocity`` default -- the x127 scaling only matters for
    the binarisation), program 0, not a drum track; constant tempo 120 bpm (the Multitrack default);
  * one time step is 1/beat_resolution of a beat, so a 96-step bar is four beats.

This writer emits SMF format 1 with ``division = beat_resolution`` ticks per quarter note, so one time step is exactly
one tick (pretty_midi would use 220 ticks per beat and round; the notes -- pitch, onset and length in beats,
velocity -- are the same).  ``read_midi`` / ``midi_to_roll`` parse the file back (tests round-trip through them).
"""
from __future__ import annotations

import struct

import numpy as np

LOWEST_PITCH = 27          # maker_bar.py:49: np.pad(note, [[0, 0], [27, 41]])
N_PITCHES = 60
BEAT_RESOLUTION = 24       # maker_bar.py:53
VELOCITY = 100             # pypianoroll to_pretty_midi(constant_velocity=100)
TEMPO_BPM = 120.0


def roll_to_notes(roll, lowest_pitch: int = LOWEST_PITCH):
    """[T, P] roll (non-zero = sounding) -> sorted list of (midi_pitch, start_step, end_step), end exclusive."""
    r = np.asarray(roll)
    if r.ndim != 2:
        raise ValueError("roll_to_notes: expected a [time, pitch] array, got shape %r" % (r.shape,))
    on = np.zeros((r.shape[0] + 2, r.shape[1]), dtype=np.int8)
    on[1:-1] = r > 0
    d = np.diff(on, axis=0)                       # +1 at the first sounding step, -1 one past the last
    notes = []
    for p in range(r.shape[1]):
        starts = np.nonzero(d[:, p] > 0)[0]
        ends = np.nonzero(d[:, p] < 0)[0]
        notes.extend((p + lowest_pitch, int(s), int(e)) for s, e in zip(starts, ends))
    notes.sort(key=lambda n: (n[1], n[0]))
    return notes


def _vlq(v: int) -> bytes:
    out = [v & 0x7F]
    v >>= 7
    while v:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    return bytes(reversed(out))


def _track(events) -> bytes:
    """events: (tick, order, payload) -> MTrk chunk with delta times and an end-of-track meta event"""
    body, t = bytearray(), 0
    for tick, _, payload in sorted(events, key=lambda e: (e[0], e[1])):
        body += _vlq(tick - t) + payload
        t = tick
    body += b"\x00\xff\x2f\x00"
    return b"MTrk" + struct.pack(">I", len(body)) + bytes(body)


def write_midi(roll, path, beat_resolution: int = BEAT_RESOLUTION, lowest_pitch: int = LOWEST_PITCH,
               velocity: int = VELOCITY, tempo_bpm: float = TEMPO_BPM, track_name: str = "piano",
               song_name: str = "test") -> int:
    """Write the [T, 60] roll as a format-1 MIDI file; returns the number of notes written."""
    notes = roll_to_notes(roll, lowest_pitch)
    if any(not 0 <= p <= 127 for p, _, _ in notes):
        raise ValueError("write_midi: pitch outside 0..127 (lowest_pitch=%d, %d columns)" %
                         (lowest_pitch, np.asarray(roll).shape[1]))
    usec = int(round(60e6 / tempo_bpm))
    name = song_name.encode()
    conductor = [(0, 0, b"\xff\x03" + _vlq(len(name)) + name),
                 (0, 1, b"\xff\x51\x03" + struct.pack(">I", usec)[1:]),
                 (0, 2, b"\xff\x58\x04\x04\x02\x18\x08")]                      # 4/4
    tname = track_name.encode()
    ev = [(0, 0, b"\xff\x03" + _vlq(len(tname)) + tname), (0, 1, b"\xc0\x00")]   # program 0 (acoustic grand)
    for p, s, e in notes:
        ev.append((s, 3, bytes((0x90, p, velocity))))
        ev.append((e, 2, bytes((0x80, p, 0))))                                  # offs sort before ons at a tick
    data = b"MThd" + struct.pack(">IHHH", 6, 1, 2, beat_resolution) + _track(conductor) + _track(ev)
    with open(path, "wb") as f:
        f.write(data)
    return len(notes)


def read_midi(path):
    """Minimal SMF parser (formats 0/1, running status): returns (division, tempo_usec, [(pitch, start, end, vel)])."""
    data = open(path, "rb").read()
    if data[:4] != b"MThd":
        raise ValueError("read_midi: not a MIDI file")
    hlen, _fmt, ntracks, division = struct.unpack(">IHHH", data[4:14])
    pos, tempo, notes = 8 + hlen, 500000, []
    for _ in range(ntracks):
        if data[pos:pos + 4] != b"MTrk":
            raise ValueError("read_midi: missing track chunk")
        (tlen,) = struct.unpack(">I", data[pos + 4:pos + 8])
        i, end, tick, status, open_notes = pos + 8, pos + 8 + tlen, 0, 0, {}
        while i < end:
            delta = 0
            while True:
                b = data[i]
                i += 1
                delta = (delta << 7) | (b & 0x7F)
                if not b & 0x80:
                    break
            tick += delta
            if data[i] & 0x80:
                status = data[i]
                i += 1
            if status == 0xFF:
                mtype = data[i]
                i += 1
                ln = 0
                while True:
                    b = data[i]
                    i += 1
                    ln = (ln << 7) | (b & 0x7F)
                    if not b & 0x80:
                        break
                if mtype == 0x51:
                    tempo = int.from_bytes(data[i:i + ln], "big")
                i += ln
            elif status in (0xF0, 0xF7):
                ln = 0
                while True:
                    b = data[i]
                    i += 1
                    ln = (ln << 7) | (b & 0x7F)
                    if not b & 0x80:
                        break
                i += ln
            else:
                kind = status & 0xF0
                nargs = 1 if kind in (0xC0, 0xD0) else 2
                a = data[i:i + nargs]
                i += nargs
                if kind == 0x90 and a[1] > 0:
                    open_notes.setdefault(a[0], []).append((tick, a[1]))
                elif kind == 0x80 or (kind == 0x90 and a[1] == 0):
                    if open_notes.get(a[0]):
                        s, v = open_notes[a[0]].pop(0)
                        notes.append((a[0], s, tick, v))
        pos = end
    notes.sort(key=lambda n: (n[1], n[0]))
    return division, tempo, notes


def midi_to_roll(path, n_steps: int, lowest_pitch: int = LOWEST_PITCH, n_pitches: int = N_PITCHES,
                 beat_resolution: int = BEAT_RESOLUTION):
    """inverse of write_midi on the [T, 60] grid (float32 {0,1})"""
    division, _, notes = read_midi(path)
    roll = np.zeros((n_steps, n_pitches), dtype=np.float32)
    for p, s, e, _ in notes:
        s, e = s * beat_resolution // division, e * beat_resolution // division
        if lowest_pitch <= p < lowest_pitch + n_pitches:
            roll[s:e, p - lowest_pitch] = 1.0
    return roll
