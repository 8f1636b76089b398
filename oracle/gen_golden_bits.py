#!/usr/bin/env python
"""Generate tests/golden/packed_v1.npz: a batch in the REFERENCE's own data format, loaded by the reference's own
``data/bar_dataset.py:NoteDataset`` from ``.npz`` items and collated as ``agent/barGen.py:134-141`` does
(np.concatenate along axis 0), next to the bits the packed path must produce for it.  TEST INFRASTRUCTURE; build
container only (needs /root/reference):    python oracle/gen_golden_bits.py"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("BARVAE_REFERENCE", "/root/reference")
sys.path.insert(0, HERE)
sys.path.insert(0, REF)

import bits_oracle as BO  # noqa: E402


def main():
    from data.bar_dataset import NoteDataset          # the reference's loader, unmodified

    class Cfg:
        data_path, batch_size = "dataset", 2
    r = np.random.RandomState(2024)
    with tempfile.TemporaryDirectory() as root:
        os.makedirs(os.path.join(root, "dataset"))
        for i, n in enumerate((2, 1, 3)):              # items hold several bars each
            np.savez(os.path.join(root, "dataset", "%02d.npz" % i),
                     note=(r.rand(n, 1, 96, 60) < 0.05).astype(np.float32),
                     pre_note=(r.rand(n, 1, 96, 60) < 0.07).astype(np.float32),
                     pre_phrase=(r.rand(n, 1, 384, 60) < 0.04).astype(np.float32),
                     position=r.randint(0, 332, size=(n,)).astype(np.int64))
        ds = NoteDataset(root, Cfg)
        ds.file_list = sorted(ds.file_list)            # os.listdir order is not defined
        samples = [ds[i] for i in range(len(ds))]
    cat = {k: np.concatenate([s[k] for s in samples], axis=0) for k in ("note", "pre_note", "pre_phrase", "position")}
    out = os.path.join(os.path.dirname(HERE), "tests", "golden", "packed_v1.npz")
    np.savez_compressed(out, bits=BO.batch_layout(cat["note"], cat["pre_note"], cat["pre_phrase"]),
                        note=cat["note"].astype(np.uint8), pre_note=cat["pre_note"].astype(np.uint8),
                        pre_phrase=cat["pre_phrase"].astype(np.uint8), position=cat["position"])
    print("wrote", out, os.path.getsize(out), "bytes;", cat["note"].shape[0], "bars")


if __name__ == "__main__":
    main()
