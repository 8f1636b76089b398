"""Data-parallel rank equivalence on real GPUs (SURVEY.md section 8e / agent/barGen_horovod.py:91-99,130-134): runs
tools/check_dp.py under torchrun on 2 GPUs when the box has them (the 1-GPU box skips; the committed outputs of the 2- and
8-GPU runs are profiles/check_dp_r2_n*.json)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_two_rank_step_equals_one_process():
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29631", os.path.join(ROOT, "tools", "check_dp.py")],
                         capture_output=True, text=True, timeout=900)
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert out.returncode == 0 and lines, (out.stdout[-2000:], out.stderr[-2000:])
    r = json.loads(lines[-1])
    assert r["ranks_params_bit_identical"] and r["ranks_allreduced_grad_bit_identical"] and r["ok"], r


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_bargen_two_ranks_unequal_dataset():
    """BarGen.train_epoch under torchrun with len(dataset) % world != 0: equal-length shards (wrap-around padding), same
    step count, bit-identical parameters, same learning rate and epoch loss on both ranks (tools/check_bargen_dp.py)"""
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29633",
                          os.path.join(ROOT, "tools", "check_bargen_dp.py")], capture_output=True, text=True, timeout=600)
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert out.returncode == 0 and lines, (out.stdout[-2000:], out.stderr[-2000:])
    r = json.loads(lines[-1])
    assert r["ok"] and r["params_bit_identical"] and r["steps_lr_loss_identical"] and r["items_per_rank"] == 5, r
