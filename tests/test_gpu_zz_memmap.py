"""GPU: BarGen fed from flat memory-mapped bit arrays (config.packed_array_path) -- the loader path with no worker
process and no per-item work.  (Runs last: added after the round's last GPU minute; every piece it composes --
PackedMemmapDataset.batches on the CPU, HostPrefetcher + step_batch on PackedBatch on the GPU -- is tested separately.)"""
import os

import numpy as np
import pytest
import torch

from gpu_util import pkg, report

pytestmark = pytest.mark.gpu


def test_bargen_trains_from_memmapped_bits(tmp_path):
    Config = pkg("config").Config
    BarGen = pkg("agent.barGen").BarGen
    P = pkg("data.packed")
    ds0 = pkg("data.bar_dataset").SyntheticBars(n_items=4, bars_per_item=2, batch_size=2, seed=3)
    src = tmp_path / "data" / "dataset"
    os.makedirs(src)
    for i in range(4):
        np.savez(src / ("%03d.npz" % i), **ds0[i])
    assert P.convert_dataset_to_arrays(str(src), str(tmp_path / "bits")) == 8

    class Cfg(Config):
        root_path = str(tmp_path)
        batch_size = 4                     # bars per step on this path
        packed_array_path = "bits"

    agent = BarGen(Cfg())
    assert isinstance(agent.dataset, P.PackedMemmapDataset) and len(agent.dataset) == 8
    l1 = agent.train_epoch()
    agent.epoch += 1
    l2 = agent.train_epoch()
    report(test="bargen_memmap", loss_epoch1=l1, loss_epoch2=l2, iterations=agent.iteration)
    assert agent.iteration == 4 and l1 == l1 and l2 == l2 and l2 < l1 * 1.5


def test_unpack_kernel_against_committed_golden():
    """tests/golden/packed_v1.npz (a reference-format batch loaded by the reference's own NoteDataset,
    oracle/gen_golden_bits.py): the committed bits expand on the device to exactly the committed cells"""
    from gpu_util import ROOT
    P = pkg("data.packed")
    g = np.load(os.path.join(ROOT, "tests", "golden", "packed_v1.npz"))
    B = g["position"].shape[0]
    pb = P.PackedBatch(torch.from_numpy(g["bits"].copy()), torch.from_numpy(g["position"].copy()), B)
    note32, bars, phrase16, pos, _ = pb.to_device(torch.device("cuda", 0))
    torch.cuda.synchronize()
    assert np.array_equal(note32.cpu().numpy(), g["note"].astype(np.float32))
    assert np.array_equal(bars.float().cpu().numpy(), np.concatenate([g["note"], g["pre_note"]]).astype(np.float32))
    assert np.array_equal(phrase16.float().cpu().numpy(), g["pre_phrase"].astype(np.float32))
    back, _ = P.threshold_pack(torch.cat((bars.float().reshape(-1), phrase16.float().reshape(-1))), 0.5)
    assert np.array_equal(back.cpu().numpy(), g["bits"])
