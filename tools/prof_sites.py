"""Per-site time of one 512-bar training step: every library call bracketed by CUDA events on its launching stream (stream
overlap off, eager launches), grouped by call tag + site detail.  Prints one JSON line per site, slowest first.
usage: prof_sites.py [bars] [steps]"""
import importlib
import json
import os
import sys

os.environ.setdefault("BVAE_GRAPH", "0")
os.environ.setdefault("BVAE_STREAMS", "0")
os.environ.setdefault("BVAE_WGRAD_STREAM", "0")
os.environ.setdefault("BVAE_DEC_STREAMS", "0")
import torch  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "musicgeneration_vae-torch_b200"
bench = importlib.import_module("bench")
eng = importlib.import_module(PKG + ".engine")
Model = importlib.import_module(PKG + ".graph.model").Model
Trainer = importlib.import_module(PKG + ".trainer").GeneratorTrainer
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
torch.manual_seed(0)
model = Model().cuda().train()
tr = Trainer(model, use_graph=False)
batch = bench.synthetic_batch(B, 1234, "cuda")
for _ in range(3):
    tr.step(*batch)
torch.cuda.synchronize()
eng.profile_begin()
for _ in range(steps):
    tr.step(*batch)
p = eng.profile_end()
rows = sorted(p["detail"].items(), key=lambda kv: -kv[1][0])
tot = {}
for k, (ms, n) in rows:
    tag = k.split(":")[0]
    tot[tag] = tot.get(tag, 0.0) + ms / steps
print(json.dumps({"per_step_ms_by_tag": {k: round(v, 3) for k, v in tot.items()}, "total_ms": round(p["total_ms"] / steps, 2)}))
for k, (ms, n) in rows:
    print(json.dumps({"site": k, "ms_per_step": round(ms / steps, 4), "calls_per_step": n / steps,
                      "us_per_call": round(ms / n * 1e3, 1)}))
