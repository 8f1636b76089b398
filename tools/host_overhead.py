"""How long does the HOST need to enqueue one training step (no device sync inside), eager vs CUDA-graph replay, and what
does the device need per step at small batches?  Prints one JSON line per (batch, mode):
    python tools/host_overhead.py [B ...]          (default: 512 64)"""
import importlib
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "musicgeneration_vae-torch_b200"
bench = importlib.import_module("bench")
Model = importlib.import_module(PKG + ".graph.model").Model
Trainer = importlib.import_module(PKG + ".trainer").GeneratorTrainer


def measure(B, graph, steps=10):
    torch.manual_seed(0)
    model = Model().cuda().train()
    tr = Trainer(model, use_graph=graph)
    batch = bench.synthetic_batch(B, 1, "cuda")
    for _ in range(6):                       # 3 eager steps, the capture, 2 replays
        tr.step(*batch)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        tr.step(*batch)
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    dev_ms = e0.elapsed_time(e1) / steps
    line = {"tool": "host_overhead", "bars": B, "mode": "graph" if graph else "eager", "captured": len(tr._graphs),
            "host_enqueue_ms_per_step": (t1 - t0) / steps * 1e3, "device_ms_per_step": dev_ms,
            "bars_per_sec": B / (dev_ms * 1e-3)}
    print(json.dumps(line), flush=True)
    del tr, model
    torch.cuda.empty_cache()
    return line


if __name__ == "__main__":
    sizes = [int(a) for a in sys.argv[1:]] or [512, 64]
    for B in sizes:
        for graph in (False, True):
            measure(B, graph)
