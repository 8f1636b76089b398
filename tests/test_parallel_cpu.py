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
