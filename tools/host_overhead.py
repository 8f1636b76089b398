"""How long does the HOST need to enqueue one training step (no device sync inside)?"""
import cProfile
import importlib
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "musicgeneration_vae-torch_b200"
sys.argv = [sys.argv[0]] + sys.argv[1:]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
bench = importlib.import_module("bench")
Model = importlib.import_module(PKG + ".graph.model").Model
Trainer = importlib.import_module(PKG + ".trainer").GeneratorTrainer
torch.manual_seed(0)
model = Model().cuda().train()
tr = Trainer(model)
batch = bench.synthetic_batch(B, 1, "cuda")
for _ in range(3):
    tr.step(*batch)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    tr.step(*batch)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("B=%d host enqueue %.1f ms/step, incl. drain %.1f ms/step" % (B, (t1 - t0) / 3 * 1e3, (t2 - t0) / 3 * 1e3))
pr = cProfile.Profile()
pr.enable()
tr.step(*batch)
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
