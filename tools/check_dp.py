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

def one_process(steps):
    ref = make(7)
    rflat = ref.flatten_parameters()
    ref.decoder.dropout.p = 0.0
    rt = trn.GeneratorTrainer(ref, lr=0.002, reducer=None)
    allb = [t.to(dev) for t in full]
    g1 = None
    for i in range(steps):
        rt.step(*allb)
        if i == 0:
            g1 = rflat.grad.clone()
    torch.cuda.synchronize()
    return rflat.data.clone(), g1


model = make(7)
flat = model.flatten_parameters()
start = flat.data.clone()
red = par.GradReducer.for_model(model, flat)
red.broadcast_parameters()
tr = trn.GeneratorTrainer(model, lr=0.002, reducer=red)
model.decoder.dropout.p = 0.0
grad1 = None
for i in range(2):
    tr.step(*mine)
    if i == 0:
        grad1 = flat.grad.clone() / world          # the all-reduced SUM; Adam applies the 1/world (grad_scale)
torch.cuda.synchronize()
mineflat = flat.data.clone()
gathered = [torch.empty_like(mineflat) for _ in range(world)]
dist.all_gather(gathered, mineflat)
same = all(torch.equal(gathered[0], g) for g in gathered)
ggrad = [torch.empty_like(grad1) for _ in range(world)]
dist.all_gather(ggrad, grad1)
same_grad = all(torch.equal(ggrad[0], g) for g in ggrad)
ok = same and same_grad
if rank == 0:
    import json
    p1, g1 = one_process(2)
    p2, g2 = one_process(2)                         # the single process against itself: fp32 atomics' order
    moved = (p1 - start).abs().mean().item()
    rel = (p1 - mineflat).abs().mean().item() / max(moved, 1e-12)
    floor = (p1 - p2).abs().mean().item() / max(moved, 1e-12)
    gn = g1.double().norm().item()
    grel = (grad1.double() - g1.double()).norm().item() / gn
    gfloor = (g2.double() - g1.double()).norm().item() / gn
    ok = ok and rel < 0.05 + 2 * floor
    line = {"tool": "check_dp", "world": world, "bars_per_rank": B, "steps": 2, "ranks_params_bit_identical": same,
            "ranks_allreduced_grad_bit_identical": same_grad,
            "params_vs_one_process_mean_abs_diff_over_mean_update": rel,
            "one_process_run_to_run_same_metric": floor,
            "first_step_grad_vs_one_process_rel_fro": grel, "one_process_grad_run_to_run_rel_fro": gfloor,
            "bound": "ranks bit-identical and params within 5 % of the mean update (+ 2 x the run-to-run floor)", "ok": ok}
    print(json.dumps(line), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "check_dp_n%d.json" % world), "w") as f:
        f.write(json.dumps(line) + "\n")
flag = torch.tensor([1 if ok else 0], device=dev)
dist.broadcast(flag, src=0)
ok = bool(flag.item())
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
