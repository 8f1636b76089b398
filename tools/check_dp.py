"""Data-parallel consistency check (run under torchrun, >= 2 GPUs):
  * after two optimisation steps every rank must hold bit-identical parameters (same all-reduced gradients);
  * N ranks x B bars must match ONE process stepping on the N*B bars (the generator has no batch-coupled op), up to the
    fp32 atomics' summation order.
usage: python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_dp.py"""
import importlib
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "musicgeneration_vae-torch_b200"
par = importlib.import_module(PKG + ".parallel")
mdl = importlib.import_module(PKG + ".graph.model")
trn = importlib.import_module(PKG + ".trainer")
data = importlib.import_module(PKG + ".data.bar_dataset")

rank, world, local = par.init_from_env()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
B = 8


def make(seed):
    torch.manual_seed(seed)
    m = mdl.Model().to(dev)
    m.train()
    return m


def batch(n, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    note = (torch.rand(n, 1, 96, 60, generator=g) < 0.05).float()
    pre = (torch.rand(n, 1, 96, 60, generator=g) < 0.05).float()
    phr = (torch.rand(n, 1, 384, 60, generator=g) < 0.05).float()
    pos = torch.randint(0, 4, (n,), generator=g)
    return note, pre, phr, pos


full = batch(B * world, 123)
mine = [t[rank * B:(rank + 1) * B].to(dev) for t in full]
masks_full = None

model = make(7)
flat = model.flatten_parameters()
red = par.GradReducer.for_model(model, flat)
red.broadcast_parameters()
tr = trn.GeneratorTrainer(model, lr=0.002, reducer=red)
model.decoder.dropout.p = 0.0
for _ in range(2):
    tr.step(*mine)
torch.cuda.synchronize()
mineflat = flat.data.clone()
gathered = [torch.empty_like(mineflat) for _ in range(world)]
dist.all_gather(gathered, mineflat)
same = all(torch.equal(gathered[0], g) for g in gathered)
ok = same
msg = "ranks identical: %s" % same
if rank == 0:
    ref = make(7)
    rflat = ref.flatten_parameters()
    ref.decoder.dropout.p = 0.0
    rt = trn.GeneratorTrainer(ref, lr=0.002 , reducer=None)
    allb = [t.to(dev) for t in full]
    for _ in range(2):
        rt.step(*allb)
    torch.cuda.synchronize()
    # DP averages per-rank mean losses == the mean over the whole batch (equal shard sizes)
    diff = (rflat.data - mineflat).abs()
    moved = (rflat.data - make(7).flatten_parameters().data).abs().mean().item()
    rel = diff.mean().item() / max(moved, 1e-12)
    msg += "; vs single process on %d bars: mean|diff| / mean|update| = %.3e" % (B * world, rel)
    ok = ok and rel < 0.2
    print(("DP CHECK OK: " if ok else "DP CHECK FAILED: ") + msg, flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
