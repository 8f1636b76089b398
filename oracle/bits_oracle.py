"""CPU oracle for the bit-packed piano-roll path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Integer / byte work, restated with explicit shifts in numpy (no numpy.packbits, so that the product's host packer --
which does use packbits -- and the CUDA kernels in csrc/bits.cu are checked against an independent statement of the
layout).  Only ``tests/`` and ``__graft_entry__.smoke()`` import this file.

What it restates: the reference has no packed format; its format is fp32 {0,1} cells in C order
(data/bar_dataset.py:22-25 -> ``np.concatenate`` along axis 0 at agent/barGen.py:134-141; generated bars
``torch.gt(pre_bar, 0.3)`` at maker_bar.py:39).  The packed layout is DEFINED as: cell i of the C-order flattening is
bit ``7 - i % 8`` of byte ``i // 8``; unused bits of the last byte are 0.  Parity bar: bit-exact.  Pinning: the
layout equals numpy.packbits(bitorder='big') by definition, checked in tests/test_packed_cpu.py together with
hand-written known-answer bytes; the round trip through it must reproduce the reference-format arrays exactly.
"""
import numpy as np


def pack(cells) -> np.ndarray:
    c = np.asarray(cells).reshape(-1)
    n = c.size
    out = np.zeros((n + 7) // 8, dtype=np.uint8)
    for j in range(8):                                   # bit 7-j of every byte <- cells j, j+8, j+16, ...
        col = (c[j::8] != 0).astype(np.uint8)
        out[:col.size] |= col << np.uint8(7 - j)
    return out


def unpack(bits, n: int) -> np.ndarray:
    b = np.asarray(bits, dtype=np.uint8)
    out = np.zeros(((n + 7) // 8) * 8, dtype=np.float32)
    for j in range(8):
        out[j::8] = (b[:(n + 7) // 8] >> np.uint8(7 - j)) & 1
    return out[:n]


def threshold_pack(probs, threshold: float):
    """maker_bar.py:39 (``torch.gt(pre_bar, 0.3)``, strict) followed by ``pack``; returns (bits, {0,1} float32)"""
    p = np.asarray(probs, dtype=np.float32).reshape(-1)
    hard = (p > np.float32(threshold)).astype(np.float32)
    return pack(hard), hard


def batch_layout(note, pre_note, pre_phrase) -> np.ndarray:
    """[note bits | pre_note bits | pre_phrase bits]: the PackedBatch buffer of data/packed.py"""
    return np.concatenate([pack(note), pack(pre_note), pack(pre_phrase)])
