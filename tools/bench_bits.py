#!/usr/bin/env python
"""Achieved HBM throughput of the bit-packing kernels (csrc/bits.cu) at a size well beyond L2.
    python tools/bench_bits.py [bars]       # default 16384 bars + phrases = 70.8 MB of bits -> 1.13 GB of bf16 cells"""
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
P = importlib.import_module("musicgeneration_vae-torch_b200.data.packed")
bars = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
nbytes = bars * 4320
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6547.8) \
    if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6547.8
bits = torch.randint(0, 256, (nbytes,), device="cuda", dtype=torch.uint8)
ob = torch.empty(nbytes * 8, device="cuda", dtype=torch.bfloat16)
of = torch.empty(nbytes * 8, device="cuda", dtype=torch.float32)
back = torch.empty(nbytes, device="cuda", dtype=torch.uint8)


def timed(fn, n=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e-3


cases = {
    "unpack -> bf16 (1 B read + 16 B written per byte of bits)": (lambda: P.unpack_bits(bits, nbytes * 8, ob, None, 0), 17),
    "unpack -> bf16 + fp32 (1 + 16 + 32 B)": (lambda: P.unpack_bits(bits, nbytes * 8, ob, of, nbytes * 8), 49),
    "threshold_pack (32 B read + 1 B written)": (lambda: P._lib.check(P._lib.lib().bvae_threshold_pack(
        of.data_ptr(), nbytes * 8, 0.5, back.data_ptr(), None, P._lib.stream_ptr())), 33),
}
for name, (fn, bpb) in cases.items():
    t = timed(fn)
    gbs = nbytes * bpb / t / 1e9
    print("%-62s %8.1f us  %7.1f GB/s  %.0f %% of the measured HBM peak (%.0f GB/s)" % (name, t * 1e6, gbs,
                                                                                       100 * gbs / peak, peak))
assert torch.equal(back, bits)
