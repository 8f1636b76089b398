"""Data-parallel gradient exchange: one process per GPU, NCCL all-reduce of the flat gradient bucket.

Replaces the reference's two data-parallel mechanisms:
  * nn.DataParallel (agent/barGen.py:96-105): per-step parameter broadcast + gradient reduce to GPU 0;
  * Horovod (agent/barGen_horovod.py:35-36,49-50,91-99,130-134): per-parameter hooks -> fused-buffer
    allreduce-average, hvd.broadcast_parameters from rank 0.
Here every generator gradient already lives in ONE contiguous fp32 buffer (engine.FlatParams), split into the three
segments the backward pass completes in order (decoder first, then the two encoders).  As soon as a segment's
backward node finishes, its slice is all-reduced (SUM) on a side stream while the remaining backward kernels keep
the compute stream busy; the 1/world_size average is folded into the fused Adam step (grad_scale).  The generator
has no batch-coupled op (InstanceNorm is per sample, BCE-mean over equal shards averages exactly), so N ranks x
local batch B are mathematically one rank with batch N*B."""
from __future__ import annotations

import os
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment; initialises the process group if world > 1."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [begin, end) of n_items for this rank (DistributedSampler without shuffling/padding).
    Shards of unequal length make ranks run different numbers of optimiser steps; training loops use
    ``shard_indices`` instead."""
    per = (n_items + world - 1) // world
    b = min(n_items, rank * per)
    return b, min(n_items, b + per)


def shard_indices(n_items: int, rank: int, world: int, drop_last: bool = False) -> List[int]:
    """This rank's item indices with EVERY rank getting the same count, as torch's DistributedSampler does
    (agent/barGen_horovod.py:49-50): the index list is padded by wrapping around to ``world * ceil(n / world)`` entries
    (``drop_last``: truncated to ``world * floor(n / world)``) and rank r takes the contiguous block r.  Equal counts
    mean equal numbers of batches, hence equal numbers of collectives per epoch on every rank -- unequal shards pair
    all-reduces of different steps and hang the longer ranks at the end of the epoch."""
    if n_items <= 0:
        return []
    if drop_last:
        per = n_items // world
        order = list(range(per * world))
    else:
        per = (n_items + world - 1) // world
        order = [i % n_items for i in range(per * world)]
    return order[rank * per:(rank + 1) * per]


class GradReducer:
    """Bucketed all-reduce of a FlatParams gradient buffer, overlapped with the backward pass."""

    def __init__(self, flat, segments: List[Tuple[str, int, int]], bucket_mb: int = 64, group=None,
                 overlap: bool = True):
        self.flat, self.group, self.overlap = flat, group, overlap
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.segments = {name: (b, e) for name, b, e in segments}
        self.bucket_elems = max(1, bucket_mb) * (1 << 20) // 4
        self.comm_stream = torch.cuda.Stream() if flat.grad.is_cuda else None
        self._pending = []
        self._finishing = False
        self.hold = False          # gradient accumulation: True while more backward passes of this step are to come

    @staticmethod
    def for_model(model, flat, bucket_mb: int = 64, group=None, overlap: bool = True) -> "GradReducer":
        """segments = contiguous flat ranges of model.encoder / decoder / phrase_encoder parameters"""
        segs = []
        for name in ("encoder", "decoder", "phrase_encoder"):
            ps = list(getattr(model, name).parameters())
            b = min(p._bvae_off for p in ps)
            e = max(-(-(p._bvae_off + p.numel()) // flat.ALIGN) * flat.ALIGN for p in ps)
            segs.append((name, b, e))
        ref = getattr(model, "refiner", None)
        if ref is not None:                        # Model(refiner=True): no backward hook, reduced by finish()
            ps = list(ref.parameters())
            segs.append(("refiner", min(p._bvae_off for p in ps),
                         max(-(-(p._bvae_off + p.numel()) // flat.ALIGN) * flat.ALIGN for p in ps)))
        r = GradReducer(flat, segs, bucket_mb, group, overlap)
        for name in ("encoder", "decoder"):
            getattr(model, name)._bvae_on_bwd_done = (lambda n=name: r.segment_ready(n))
        model.phrase_encoder.phrase_encoder._bvae_on_bwd_done = lambda: r.segment_ready("phrase_encoder")
        return r

    def broadcast_parameters(self, root: int = 0, trainer=None):
        """hvd.broadcast_parameters(state_dict, root_rank=0) (agent/barGen_horovod.py:130-134) plus -- unlike the reference
        -- the optimiser state.  Every rank issues the SAME collectives whatever it holds locally (only the root may have
        read a checkpoint): a header [has_state, step_count, lr] goes first, ranks without moments allocate zeros.  The
        bf16 GEMM operands packed from the old parameter values are invalidated and repacked."""
        if self.world <= 1:
            return
        from . import engine
        flat = self.flat
        dev = flat.data.device
        has = flat.exp_avg is not None
        head = torch.tensor([1.0 if has else 0.0, float(trainer.step_count) if trainer is not None else 0.0,
                             float(trainer.lr) if trainer is not None else 0.0], dtype=torch.float64, device=dev)
        dist.broadcast(head, src=root, group=self.group)
        dist.broadcast(flat.data, src=root, group=self.group)
        if head[0].item() > 0:
            if flat.exp_avg is None:
                flat.exp_avg = torch.zeros_like(flat.data)
                flat.exp_avg_sq = torch.zeros_like(flat.data)
            dist.broadcast(flat.exp_avg, src=root, group=self.group)
            dist.broadcast(flat.exp_avg_sq, src=root, group=self.group)
        elif has:                                   # the root has no state: nobody keeps one
            flat.exp_avg = flat.exp_avg_sq = None
        if trainer is not None:
            trainer.step_count, trainer.lr = int(head[1].item()), float(head[2].item())
        engine.bump_param_epoch()                   # parameters are .data views: _version did not change
        if flat.data.is_cuda:
            engine.repack_weights(flat)

    def segment_ready(self, name: str):
        """Called right after a module's backward kernels were enqueued: all-reduce its gradient slice.
        With overlap=False (a module runs several backward passes per step, e.g. the GAN schedules) nothing happens
        here and finish() reduces everything."""
        if self.world == 1 or ((not self.overlap or self.hold) and not self._finishing):
            return
        if name in self._pending:
            raise RuntimeError("gradient segment %r was completed twice in one step; build the reducer with "
                               "overlap=False for schedules that run a module's backward more than once" % name)
        b, e = self.segments[name]
        g = self.flat.grad
        if self.comm_stream is not None:
            self.comm_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.comm_stream):
                for s in range(b, e, self.bucket_elems):
                    dist.all_reduce(g[s:min(e, s + self.bucket_elems)], op=dist.ReduceOp.SUM, group=self.group)
        else:
            for s in range(b, e, self.bucket_elems):
                dist.all_reduce(g[s:min(e, s + self.bucket_elems)], op=dist.ReduceOp.SUM, group=self.group)
        self._pending.append(name)

    def finish(self) -> float:
        """Make the compute stream wait for the reductions; returns the grad scale (1/world) for the optimiser.
        Segments whose backward did not run this step (frozen / unused) are reduced here so ranks stay in step."""
        if self.world == 1:
            return 1.0
        self._finishing = True
        for name in self.segments:
            if name not in self._pending:
                self.segment_ready(name)
        self._finishing = False
        if self.comm_stream is not None:
            torch.cuda.current_stream().wait_stream(self.comm_stream)
        self._pending = []
        return 1.0 / self.world
