#!/usr/bin/env python
"""Convert a reference-format dataset directory (data/bar_dataset.py: one .npz per item with fp32 note / pre_note /
pre_phrase / position) into the bit-packed format of data/packed.py, and report loader throughput of both.

    python tools/pack_dataset.py SRC_DIR DST_DIR            # convert to packed .npz items
    python tools/pack_dataset.py --arrays SRC_DIR DST_DIR   # convert to flat memory-mappable arrays (PackedMemmapDataset)
    python tools/pack_dataset.py --selftest [--items 64]    # synthetic items in a temp dir: sizes + collate bars/s
"""
import argparse
import importlib
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
P = importlib.import_module("musicgeneration_vae-torch_b200.data.packed")
D = importlib.import_module("musicgeneration_vae-torch_b200.data.bar_dataset")


def dir_bytes(d):
    return sum(os.path.getsize(os.path.join(d, f)) for f in os.listdir(d))


def selftest(items, bars_per_item):
    ds = D.SyntheticBars(items, bars_per_item, 8, seed=1)
    with tempfile.TemporaryDirectory() as tmp:
        src, dst = os.path.join(tmp, "f32"), os.path.join(tmp, "bits")
        os.makedirs(src)
        for i in range(items):
            np.savez(os.path.join(src, "%05d.npz" % i), **ds[i])
        t0 = time.perf_counter()
        bars = P.convert_dataset(src, dst)
        t_conv = time.perf_counter() - t0

        class Cfg:
            data_path, packed_data_path, batch_size = "f32", "bits", 8
        f32 = D.NoteDataset(tmp, Cfg)
        pk = P.PackedNoteDataset(tmp, Cfg)
        cat = lambda samples, k: np.concatenate([s[k] for s in samples], axis=0)
        t0 = time.perf_counter()
        s = [f32[i] for i in range(items)]
        ref_batch = tuple(np.ascontiguousarray(cat(s, k), dtype=np.float32) for k in ("note", "pre_note", "pre_phrase"))
        t_f32 = time.perf_counter() - t0
        t0 = time.perf_counter()
        pb = P.collate_packed([pk[i] for i in range(items)])
        t_pk = time.perf_counter() - t0
        n, p, ph, _ = pb.to_host_arrays()
        assert all(np.array_equal(a, b) for a, b in zip((n, p, ph), ref_batch)), "packed round trip differs"
        arr = os.path.join(tmp, "arr")
        P.convert_dataset_to_arrays(dst, arr)
        mm = P.PackedMemmapDataset(arr)
        t0 = time.perf_counter()
        reps = 20
        for r in range(reps):
            for b in mm.batches(bars, shuffle=True, seed=r):
                pass
        t_mm = (time.perf_counter() - t0) / reps
        assert np.array_equal(mm.batch(np.arange(bars)).bits.numpy(), pb.bits.numpy())
        print("memory-mapped arrays (%.2f MB): shuffled %d-bar batches at %.0f bars/s (one process)" %
              (dir_bytes(arr) / 1e6, bars, bars / t_mm))
        print("bars %d | disk fp32 %.1f MB, packed %.2f MB | convert %.2f s | load+collate: fp32 %.0f bars/s, "
              "packed %.0f bars/s (one process)" % (bars, dir_bytes(src) / 1e6, dir_bytes(dst) / 1e6, t_conv,
                                                    bars / t_f32, bars / t_pk))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("src", nargs="?")
    ap.add_argument("dst", nargs="?")
    ap.add_argument("--selftest", action="store_true")
    ap.add_argument("--arrays", action="store_true")
    ap.add_argument("--items", type=int, default=64)
    ap.add_argument("--bars-per-item", type=int, default=8)
    a = ap.parse_args()
    if a.selftest:
        return selftest(a.items, a.bars_per_item)
    if not a.src or not a.dst:
        ap.error("SRC_DIR and DST_DIR are required")
    conv = P.convert_dataset_to_arrays if a.arrays else P.convert_dataset
    print("packed %d bars into %s" % (conv(a.src, a.dst), a.dst))


if __name__ == "__main__":
    main()
