"""agent.barGen.BarGen under torchrun (>= 2 GPUs) on a dataset whose size is NOT a multiple of the world size (ADVICE
round 1: unequal shards made ranks run different numbers of steps and hang): two epochs must complete on every rank, with
the same number of steps, identical parameters, identical learning rate and identical epoch loss on all ranks.
usage: python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_bargen_dp.py"""
import importlib
import json
import os
import sys
import tempfile

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "musicgeneration_vae-torch_b200"
Config = importlib.import_module(PKG + ".config").Config
BarGen = importlib.import_module(PKG + ".agent.barGen").BarGen
SyntheticBars = importlib.import_module(PKG + ".data.bar_dataset").SyntheticBars


class Cfg(Config):
    root_path = tempfile.mkdtemp(prefix="bvae_dp_")
    batch_size = 2
    epoch = 2
    pretraining_step_size = 100


world = int(os.environ.get("WORLD_SIZE", "1"))
agent = BarGen(Cfg(), dataset=SyntheticBars(n_items=4 * world + 1, bars_per_item=2, batch_size=2, seed=3))
losses = []
for _ in range(2):
    agent.epoch += 1
    losses.append(agent.train_epoch())
torch.cuda.synchronize()
flat = agent.opt_gen1.flat.data
mine = torch.cat((flat.double().sum().view(1), flat.double().abs().sum().view(1),
                  torch.tensor([float(agent.iteration), agent.opt_gen1.lr, losses[-1], float(len(agent.indices))],
                               dtype=torch.float64, device=flat.device)))
allv = [torch.empty_like(mine) for _ in range(world)]
dist.all_gather(allv, mine)
full = [torch.empty_like(flat) for _ in range(world)]
dist.all_gather(full, flat)
same_params = all(torch.equal(full[0], f) for f in full)
same_host = all(torch.equal(allv[0][2:], v[2:]) for v in allv)
ok = same_params and same_host
if agent.rank == 0:
    print(json.dumps({"tool": "check_bargen_dp", "world": world, "items": 4 * world + 1, "items_per_rank": len(agent.indices),
                      "steps_per_rank": agent.iteration, "params_bit_identical": same_params,
                      "steps_lr_loss_identical": same_host, "epoch_losses": losses, "ok": ok}), flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
