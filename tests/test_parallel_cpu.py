"""CPU (gloo, world_size 2): host logic of the data-parallel path -- sharding, flat-bucket all-reduce, the
"N ranks x local batch == 1 rank x global batch" identity the generator's gradient exchange relies on."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

from gpu_util import pkg


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _net():
    torch.manual_seed(0)
    return nn.Sequential(nn.Linear(13, 29), nn.Tanh(), nn.Linear(29, 7))


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    par, eng = pkg("parallel"), pkg("engine")
    r, w, _ = par.init_from_env("gloo")
    assert (r, w) == (rank, world)
    net = _net()
    if rank == 1:                                  # rank 0's parameters must win
        with torch.no_grad():
            for p in net.parameters():
                p.add_(1.0)
    flat = eng.flatten(net)
    segs = [("a", 0, flat.offsets[2]), ("b", flat.offsets[2], flat.numel)]
    red = par.GradReducer(flat, segs, bucket_mb=1)
    red.broadcast_parameters(0)
    torch.manual_seed(1)
    x = torch.randn(8, 13)
    b, e = par.shard_range(8, rank, world)
    flat.attach_grads()
    net(x[b:e]).pow(2).mean().backward()
    red.segment_ready("b")                         # backward order: last layer first
    red.segment_ready("a")
    scale = red.finish()
    out[rank] = (flat.grad.clone() * scale, flat.data.clone())
    dist.destroy_process_group()


def test_two_rank_gradient_equals_single_rank():
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        g0, p0 = out[0]
        g1, p1 = out[1]
    assert torch.equal(p0, p1)                     # broadcast from rank 0
    assert torch.allclose(g0, g1)
    eng = pkg("engine")
    net = _net()
    flat = eng.flatten(net)
    torch.manual_seed(1)
    x = torch.randn(8, 13)
    flat.attach_grads()
    net(x).pow(2).mean().backward()
    assert torch.allclose(flat.grad, g0, rtol=1e-5, atol=1e-7)
    assert torch.equal(flat.data, p0)


def test_shard_range_partitions():
    par = pkg("parallel")
    for n in (0, 1, 7, 8, 4096):
        for w in (1, 2, 3, 8):
            spans = [par.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))


def test_segment_completed_twice_is_an_error():
    import pytest
    par, eng = pkg("parallel"), pkg("engine")
    net = _net()
    flat = eng.flatten(net)
    red = par.GradReducer(flat, [("a", 0, flat.numel)])
    red.world = 2                                  # pretend; no collective is reached before the error
    red._pending.append("a")
    with pytest.raises(RuntimeError):
        red.segment_ready("a")


def test_shard_indices_give_every_rank_the_same_count():
    """n % world != 0 (shards 3,3,3,1 / an empty last shard made ranks run different numbers of steps
    and hang in their last all-reduce): DistributedSampler-style wrap-around padding, every index still covered."""
    par = pkg("parallel")
    for n in (1, 7, 9, 10, 4097):
        for w in (1, 2, 3, 4, 8):
            shards = [par.shard_indices(n, r, w) for r in range(w)]
            assert len({len(s) for s in shards}) == 1 and len(shards[0]) == -(-n // w), (n, w)
            assert set(i for s in shards for i in s) == set(range(n))
            dropped = [par.shard_indices(n, r, w, drop_last=True) for r in range(w)]
            assert len({len(s) for s in dropped}) == 1 and len(dropped[0]) == n // w
    assert par.shard_indices(0, 0, 2) == []


def test_memmap_batches_equal_count_per_rank(tmp_path):
    import numpy as np
    P = pkg("data.packed")
    n = 10
    rng = np.random.RandomState(0)
    np.save(tmp_path / "note_bits.npy", rng.randint(0, 256, (n, P.BAR_BYTES), dtype=np.uint8))
    np.save(tmp_path / "pre_note_bits.npy", rng.randint(0, 256, (n, P.BAR_BYTES), dtype=np.uint8))
    np.save(tmp_path / "pre_phrase_bits.npy", rng.randint(0, 256, (n, P.PHRASE_BYTES), dtype=np.uint8))
    np.save(tmp_path / "position.npy", np.arange(n, dtype=np.int64))
    ds = P.PackedMemmapDataset(str(tmp_path))
    for world in (3, 4):
        per_rank = [[b.batch for b in ds.batches(2, rank=r, world=world)] for r in range(world)]
        assert len({tuple(x) for x in per_rank}) == 1, per_rank            # same number AND sizes of batches on every rank
        seen = set()
        for r in range(world):
            for b in ds.batches(2, rank=r, world=world):
                seen.update(int(i) for i in b.position)
        assert seen == set(range(n))


def _bcast_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    par, eng = pkg("parallel"), pkg("engine")
    par.init_from_env("gloo")
    net = _net()
    flat = eng.flatten(net)

    class T:                                        # what broadcast_parameters reads from / writes to a trainer
        step_count, lr = 0, 0.002

    t = T()
    if rank == 0:                                   # only the root "read a checkpoint"
        flat.exp_avg = torch.full_like(flat.data, 0.5)
        flat.exp_avg_sq = torch.full_like(flat.data, 0.25)
        t.step_count, t.lr = 17, 0.00128
    red = par.GradReducer(flat, [("a", 0, flat.numel)])
    epoch0 = eng._PARAM_EPOCH[0]
    red.broadcast_parameters(0, trainer=t)
    out[rank] = (flat.exp_avg.clone(), flat.exp_avg_sq.clone(), t.step_count, t.lr, eng._PARAM_EPOCH[0] > epoch0)
    dist.destroy_process_group()


def test_broadcast_when_only_the_root_has_optimizer_state():
    """the collective count must not depend on local state (it hung when only rank 0 had moments), and
    step count / lr travel with the moments; the packed-operand epoch is bumped"""
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_bcast_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        a, b = out[0], out[1]
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and float(b[0][0]) == 0.5
    assert (a[2], b[2]) == (17, 17) and abs(b[3] - 0.00128) < 1e-12 and a[4] and b[4]
