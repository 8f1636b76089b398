"""One 512-bar training step between cudaProfilerStart / cudaProfilerStop (for `ncu --profile-from-start off`): eager
launches (no graph replay) after 3 warm-up steps.  usage: ncu ... python tools/one_step.py [bars]"""
import importlib
import os
import sys

os.environ.setdefault("BVAE_GRAPH", "0")
import torch  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "musicgeneration_vae-torch_b200"
bench = importlib.import_module("bench")
Model = importlib.import_module(PKG + ".graph.model").Model
Trainer = importlib.import_module(PKG + ".trainer").GeneratorTrainer
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
torch.manual_seed(0)
model = Model().cuda().train()
tr = Trainer(model, use_graph=False)
batch = bench.synthetic_batch(B, 1234, "cuda")
for _ in range(3):
    tr.step(*batch)
torch.cuda.synchronize()
torch.cuda.profiler.start()
tr.step(*batch)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("one step done, launches counted by the library:", importlib.import_module(PKG).launch_count())
