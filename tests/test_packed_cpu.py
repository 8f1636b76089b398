"""CPU: the bit-packed piano-roll format (host packer, oracle, layout) and the MIDI tail."""
import os
import sys

import numpy as np
import pytest
import torch

from gpu_util import ROOT, pkg

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import bits_oracle as BO  # noqa: E402


def test_oracle_known_answers():
    assert BO.pack([1, 0, 0, 0, 0, 0, 0, 1]).tolist() == [0x81]
    assert BO.pack([0, 1, 1]).tolist() == [0x60]                     # ragged: low bits of the last byte are 0
    assert BO.pack([1] * 9).tolist() == [0xFF, 0x80]
    assert BO.pack([]).tolist() == []
    assert BO.unpack([0xA5, 0xC0], 10).tolist() == [1, 0, 1, 0, 0, 1, 0, 1, 1, 1]
    bits, hard = BO.threshold_pack([0.3, 0.30000004, 0.9, 0.0], 0.3)    # strict >, as torch.gt (maker_bar.py:39)
    assert hard.tolist() == [0, 1, 1, 0] and bits.tolist() == [0x60]


@pytest.mark.parametrize("n", [0, 1, 7, 8, 9, 63, 5760, 5760 * 3 + 5])
def test_oracle_equals_numpy_packbits_and_roundtrips(n):
    r = np.random.RandomState(n)
    cells = (r.rand(n) < 0.3).astype(np.float32)
    bits = BO.pack(cells)
    assert np.array_equal(bits, np.packbits(cells.astype(np.uint8)))
    assert np.array_equal(BO.unpack(bits, n), cells)


def test_host_packer_matches_oracle_and_rejects_non_binary():
    P = pkg("data.packed")
    r = np.random.RandomState(3)
    for density in (0.0, 0.05, 0.5, 1.0):
        x = (r.rand(5, 1, 96, 60) < density).astype(np.float32)
        assert np.array_equal(P.pack_cells(x), BO.pack(x))
        assert np.array_equal(P.pack_cells(torch.from_numpy(x)), BO.pack(x))
        assert np.array_equal(P.unpack_cells_host(P.pack_cells(x), x.shape), x)
    with pytest.raises(ValueError):
        P.pack_cells(np.array([0.0, 0.5, 1.0]))
    with pytest.raises(ValueError):
        P.pack_cells(np.array([0, 2], dtype=np.int64))


def test_packed_batch_layout_and_roundtrip():
    """reference-format batch (agent/barGen.py:134-141 shapes) -> bits -> the same arrays, bit-exact; 32x smaller"""
    P = pkg("data.packed")
    r = np.random.RandomState(7)
    B = 6
    note = (r.rand(B, 1, 96, 60) < 0.05).astype(np.float32)
    pre = (r.rand(B, 1, 96, 60) < 0.05).astype(np.float32)
    phrase = (r.rand(B, 1, 384, 60) < 0.05).astype(np.float32)
    pos = r.randint(0, 332, size=(B,)).astype(np.int64)
    pb = P.PackedBatch.from_arrays(note, pre, phrase, pos)
    assert pb.bits.dtype == torch.uint8 and pb.bits.numel() == B * 4320
    assert np.array_equal(pb.bits.numpy(), BO.batch_layout(note, pre, phrase))
    n2, p2, ph2, pos2 = pb.to_host_arrays()
    assert np.array_equal(n2, note) and np.array_equal(p2, pre) and np.array_equal(ph2, phrase)
    assert np.array_equal(pos2, pos)
    assert (note.nbytes + pre.nbytes + phrase.nbytes) == 32 * pb.bits.numel()
    with pytest.raises(ValueError):
        P.PackedBatch.from_arrays(note[:2], pre, phrase, pos)
    with pytest.raises(RuntimeError):                                 # device expansion has no CPU fallback
        P.unpack_bits(pb.bits, 8, None, torch.zeros(8), 8)
    with pytest.raises(RuntimeError):
        P.threshold_pack(torch.zeros(8), 0.3)


# ---------------------------------------------------------------------------------------------------------------
# MIDI tail (maker_bar.py:46-54)
# ---------------------------------------------------------------------------------------------------------------
def test_roll_to_notes_merges_runs_and_offsets_pitch():
    M = pkg("midi")
    roll = np.zeros((8, 60), dtype=np.float32)
    roll[1:4, 0] = 127          # one 3-step note on the lowest cell -> MIDI pitch 27 (np.pad [27, 41])
    roll[5, 0] = 127            # a second note on the same pitch after a gap
    roll[0:8, 59] = 1           # sounding through both ends: closed by the zero padding step
    assert M.roll_to_notes(roll) == [(86, 0, 8), (27, 1, 4), (27, 5, 6)]
    assert M.roll_to_notes(np.zeros((4, 60))) == []
    with pytest.raises(ValueError):
        M.roll_to_notes(np.zeros(5))


def test_midi_bytes_known_answer(tmp_path):
    """a one-note file, byte for byte (SMF format 1, 24 ticks per quarter, tempo 120, velocity 100)"""
    M = pkg("midi")
    roll = np.zeros((4, 60), dtype=np.float32)
    roll[1:3, 33] = 1                                            # pitch 60, steps 1..2
    path = tmp_path / "one.mid"
    assert M.write_midi(roll, path) == 1
    want = (b"MThd" + bytes([0, 0, 0, 6, 0, 1, 0, 2, 0, 24])
            + b"MTrk" + bytes([0, 0, 0, 27])
            + b"\x00\xff\x03\x04test" + b"\x00\xff\x51\x03\x07\xa1\x20" + b"\x00\xff\x58\x04\x04\x02\x18\x08"
            + b"\x00\xff\x2f\x00"
            + b"MTrk" + bytes([0, 0, 0, 24])
            + b"\x00\xff\x03\x05piano" + b"\x00\xc0\x00" + b"\x01\x90\x3c\x64" + b"\x02\x80\x3c\x00"
            + b"\x00\xff\x2f\x00")
    assert open(path, "rb").read() == want


@pytest.mark.parametrize("density", [0.0, 0.05, 0.5])
def test_midi_roundtrip(tmp_path, density):
    M = pkg("midi")
    r = np.random.RandomState(11)
    roll = (r.rand(2 * 4 * 96, 60) < density).astype(np.float32)      # music_length 2: 8 bars
    path = tmp_path / "t.mid"
    n = M.write_midi(roll * 127, path)
    division, tempo, notes = M.read_midi(path)
    assert division == 24 and tempo == 500000 and len(notes) == n
    assert all(v == 100 and 27 <= p <= 86 for p, _, _, v in notes)
    assert np.array_equal(M.midi_to_roll(path, roll.shape[0]), roll)
    # long delta times use multi-byte variable-length quantities
    sparse = np.zeros((40000, 60), dtype=np.float32)
    sparse[39990:39995, 5] = 1
    M.write_midi(sparse, path)
    assert np.array_equal(M.midi_to_roll(path, 40000), sparse)


def test_packed_dataset_on_disk_and_collate(tmp_path):
    """reference-format .npz items -> convert_dataset -> PackedNoteDataset -> collate_packed == packing the
    reference's own collate (np.concatenate along axis 0, agent/barGen.py:134-141); ragged items (different bar counts)"""
    P = pkg("data.packed")
    src, dst = tmp_path / "f32", tmp_path / "bits"
    src.mkdir()
    r = np.random.RandomState(5)
    items = []
    for i, n in enumerate((3, 1, 4)):
        it = {"note": (r.rand(n, 1, 96, 60) < 0.05).astype(np.float32),
              "pre_note": (r.rand(n, 1, 96, 60) < 0.05).astype(np.float32),
              "pre_phrase": (r.rand(n, 1, 384, 60) < 0.05).astype(np.float32),
              "position": r.randint(0, 332, size=(n,)).astype(np.int64)}
        np.savez(src / ("%03d.npz" % i), **it)
        items.append(it)
    assert P.convert_dataset(str(src), str(dst)) == 8

    class Cfg:
        data_path, packed_data_path, batch_size = "f32", "bits", 2
    ds = P.PackedNoteDataset(str(tmp_path), Cfg)
    assert len(ds) == 3 and ds.num_iterations == 2
    assert ds[0]["note_bits"].shape == (3, 720) and ds[2]["pre_phrase_bits"].shape == (4, 2880)
    back = P.unpack_item(ds[1])
    assert all(np.array_equal(back[k], items[1][k]) for k in items[1])
    pb = P.collate_packed([ds[i] for i in range(3)])
    cat = lambda k: np.concatenate([it[k] for it in items], axis=0)
    want = P.PackedBatch.from_arrays(cat("note"), cat("pre_note"), cat("pre_phrase"), cat("position"))
    assert pb.batch == 8 and torch.equal(pb.bits, want.bits) and torch.equal(pb.position, want.position)
    assert np.array_equal(pb.bits.numpy(), BO.batch_layout(cat("note"), cat("pre_note"), cat("pre_phrase")))
    bad = dict(ds[0])
    bad["note_bits"] = bad["note_bits"][:, :-1]
    with pytest.raises(ValueError):
        P.collate_packed([bad])


def test_memmap_dataset_batches(tmp_path):
    """convert_dataset_to_arrays + PackedMemmapDataset: contiguous and shuffled batches equal packing the reference-
    format bars directly; rank shards are disjoint and cover everything (agent/barGen_horovod.py:49-50 semantics)"""
    P = pkg("data.packed")
    src, dst = tmp_path / "f32", tmp_path / "arr"
    src.mkdir()
    r = np.random.RandomState(9)
    items = []
    for i, n in enumerate((3, 2, 4, 1)):
        it = {"note": (r.rand(n, 1, 96, 60) < 0.05).astype(np.float32),
              "pre_note": (r.rand(n, 1, 96, 60) < 0.05).astype(np.float32),
              "pre_phrase": (r.rand(n, 1, 384, 60) < 0.05).astype(np.float32),
              "position": r.randint(0, 332, size=(n,)).astype(np.int64)}
        np.savez(src / ("%03d.npz" % i), **it)
        items.append(it)
    assert P.convert_dataset_to_arrays(str(src), str(dst)) == 10
    ds = P.PackedMemmapDataset(str(dst))
    assert len(ds) == 10
    cat = lambda k: np.concatenate([it[k] for it in items], axis=0)

    def want(idx):
        return P.PackedBatch.from_arrays(cat("note")[idx], cat("pre_note")[idx], cat("pre_phrase")[idx], cat("position")[idx])

    for idx in ([2, 3, 4, 5], [7, 0, 9, 3, 3], [4]):
        got, w = ds.batch(idx), want(np.asarray(idx))
        assert got.batch == len(idx) and torch.equal(got.bits, w.bits) and torch.equal(got.position, w.position)
    one = P.collate_packed([ds[6], ds[1]])
    w = want(np.asarray([6, 1]))
    assert torch.equal(one.bits, w.bits) and torch.equal(one.position, w.position)
    seen = []
    for rank in range(3):
        for b in ds.batches(2, shuffle=True, seed=4, rank=rank, world=3):
            seen.extend(b.position.tolist())
            assert b.batch <= 2
    # 10 bars over 3 ranks: every rank gets 4 (two bars are visited twice -- DistributedSampler's wrap-around padding,
    # so that all ranks run the same number of steps), every bar is visited
    from collections import Counter
    have, need = Counter(seen), Counter(cat("position").tolist())
    assert len(seen) == 12 and all(have[k] >= v for k, v in need.items())
    assert sum(b.batch for b in ds.batches(4, drop_last=True)) == 8


def test_memmap_loader_is_reiterable_and_sharded(tmp_path):
    P = pkg("data.packed")
    Loader = pkg("agent.barGen")._MemmapLoader
    ds0 = pkg("data.bar_dataset").SyntheticBars(n_items=5, bars_per_item=2, batch_size=2, seed=2)
    src, dst = tmp_path / "f32", tmp_path / "arr"
    src.mkdir()
    for i in range(5):
        np.savez(src / ("%03d.npz" % i), **ds0[i])
    P.convert_dataset_to_arrays(str(src), str(dst))
    ds = P.PackedMemmapDataset(str(dst))
    for world in (1, 2):
        total = []
        for rank in range(world):
            ld = Loader(ds, 4, rank, world, False)
            first = [b.position.tolist() for b in ld]
            assert first == [b.position.tolist() for b in ld]              # a second epoch sees the same batches
            total += [p for b in first for p in b]
        assert total == np.load(dst / "position.npy").tolist()


def test_golden_reference_format_batch():
    """tests/golden/packed_v1.npz (oracle/gen_golden_bits.py: a batch loaded by the reference's own NoteDataset and
    collated as agent/barGen.py:134-141): oracle, host packer and collate all produce the committed bits, and the bits
    expand back to the reference-format arrays"""
    P = pkg("data.packed")
    g = np.load(os.path.join(ROOT, "tests", "golden", "packed_v1.npz"))
    note, pre, phrase = (g[k].astype(np.float32) for k in ("note", "pre_note", "pre_phrase"))
    assert note.shape == (6, 1, 96, 60) and phrase.shape == (6, 1, 384, 60)
    assert np.array_equal(BO.batch_layout(note, pre, phrase), g["bits"])
    pb = P.PackedBatch.from_arrays(note, pre, phrase, g["position"])
    assert np.array_equal(pb.bits.numpy(), g["bits"])
    items = [P.pack_item({"note": note[a:b], "pre_note": pre[a:b], "pre_phrase": phrase[a:b], "position": g["position"][a:b]})
             for a, b in ((0, 2), (2, 3), (3, 6))]
    assert np.array_equal(P.collate_packed(items).bits.numpy(), g["bits"])
    n2, p2, ph2, pos2 = pb.to_host_arrays()
    assert np.array_equal(n2, note) and np.array_equal(p2, pre) and np.array_equal(ph2, phrase)
    assert np.array_equal(BO.unpack(g["bits"][:6 * 720], 6 * 5760).reshape(note.shape), note)
