"""Experiment: how much would running two half-batches concurrently (two stream sets) buy?  Two independent model replicas,
256 bars each, replayed as CUDA graphs on two streams, against one replica at 512 bars.  (Separate replicas: no shared
gradient buffers, so this only measures the achievable overlap.)"""
import importlib
import json
import os
import sys

import warnings

import torch

warnings.simplefilter("error")            # a failed capture must stop the experiment, with its message

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "musicgeneration_vae-torch_b200"
bench = importlib.import_module("bench")
Model = importlib.import_module(PKG + ".graph.model").Model
Trainer = importlib.import_module(PKG + ".trainer").GeneratorTrainer


def make(B, seed):
    torch.manual_seed(seed)
    tr = Trainer(Model().cuda().train(), use_graph=True)
    return tr, bench.synthetic_batch(B, seed, "cuda")


def run(pairs, steps):
    streams = [torch.cuda.Stream() for _ in pairs]
    for _ in range(6):
        for (tr, b), s in zip(pairs, streams):
            with torch.cuda.stream(s):
                tr.step(*b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in streams:
        s.wait_stream(torch.cuda.current_stream())
    for _ in range(steps):
        for (tr, b), s in zip(pairs, streams):
            with torch.cuda.stream(s):
                tr.step(*b)
    for s in streams:
        torch.cuda.current_stream().wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


one = run([make(512, 1)], 10)
print(json.dumps({"config": "1 x 512 bars", "ms_per_512_bars": one}), flush=True)
torch.cuda.empty_cache()
two = run([make(256, 1), make(256, 2)], 10)
print(json.dumps({"config": "2 x 256 bars concurrently", "ms_per_512_bars": two}), flush=True)
four = run([make(128, i) for i in range(4)], 10)
print(json.dumps({"config": "4 x 128 bars concurrently", "ms_per_512_bars": four}), flush=True)
