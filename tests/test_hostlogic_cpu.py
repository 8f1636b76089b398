"""CPU: host logic of the training entry point that mirrors torch / reference behaviour."""
import random

import torch

from gpu_util import pkg


def test_plateau_scheduler_matches_torch():
    """agent/barGen.py:70-81,361-367: ReduceLROnPlateau(opt, mode='min', factor=0.8, cooldown=6) stepped once per epoch
    on the mean loss -- the learning-rate trajectory of our _Plateau equals torch's on plateaus, noise and improvements"""
    Plateau = pkg("agent.barGen")._Plateau
    rnd = random.Random(4)
    seqs = {
        "flat": [1.0] * 80,
        "improving": [1.0 / (1 + 0.05 * i) for i in range(80)],
        "noisy plateau": [1.0 + 0.01 * rnd.uniform(-1, 1) for _ in range(120)],
        "steps": [1.0] * 15 + [0.5] * 30 + [0.5001] * 40 + [0.2] * 20,
        "tiny improvements below the 1e-4 threshold": [1.0 - 1e-6 * i for i in range(60)],
    }
    for name, losses in seqs.items():
        p = torch.nn.Parameter(torch.zeros(1))
        opt = torch.optim.Adam([p], lr=0.002)
        ref = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, mode="min", factor=0.8, cooldown=6)
        ours, lr = Plateau(factor=0.8, cooldown=6), 0.002
        for i, v in enumerate(losses):
            ref.step(v)
            lr = ours.step(v, lr)
            assert abs(lr - opt.param_groups[0]["lr"]) < 1e-12, (name, i, lr, opt.param_groups[0]["lr"])
        if name in ("flat", "noisy plateau"):
            assert lr < 0.002                       # the sequence did trigger reductions


def test_make_batch_layouts():
    """agent/barGen.py:134-141: items concatenate along axis 0; with config.packed_input the same batch as bits"""
    import numpy as np
    B = pkg("agent.barGen").BarGen
    P = pkg("data.packed")
    ds = pkg("data.bar_dataset").SyntheticBars(n_items=3, bars_per_item=2, batch_size=2, seed=1)

    class Stub:
        pass
    s = Stub()
    s.config = Stub()
    s.config.packed_input = False
    note, pre, phrase, pos = B.make_batch(s, [ds[0], ds[2]])
    assert note.shape == (4, 1, 96, 60) and phrase.shape == (4, 1, 384, 60) and pos.dtype == torch.long
    assert torch.equal(note[2:], torch.from_numpy(ds[2]["note"]))
    s.config.packed_input = True
    pb = B.make_batch(s, [ds[0], ds[2]])
    assert isinstance(pb, P.PackedBatch) and pb.batch == 4
    n2, p2, ph2, pos2 = pb.to_host_arrays()
    assert np.array_equal(n2, note.numpy()) and np.array_equal(ph2, phrase.numpy()) and np.array_equal(pos2, pos.numpy())
    packed_items = [P.pack_item(ds[0]), P.pack_item(ds[2])]
    pb2 = B.make_batch(s, packed_items)                       # items already stored as bits
    assert torch.equal(pb2.bits, pb.bits) and torch.equal(pb2.position, pb.position)


def test_main_entry_overrides_and_missing_dataset(tmp_path, monkeypatch):
    """main.py (reference main.py:7-12): Config overrides from the command line; a missing dataset directory is an error
    unless synthetic data is asked for explicitly -- checked up to the point where a GPU would be needed"""
    import pytest
    main = pkg("main")
    with pytest.raises(SystemExit):
        main.main(["--set", "no_such_attribute=1"])
    BarGen = pkg("agent.barGen").BarGen
    Config = pkg("config").Config

    class Cfg(Config):
        root_path = str(tmp_path)

    import torch
    monkeypatch.setattr(torch.cuda, "set_device", lambda *a, **k: None)
    with pytest.raises(FileNotFoundError):
        BarGen(Cfg())
