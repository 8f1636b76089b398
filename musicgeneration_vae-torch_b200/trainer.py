"""The generator training step of the pre-training phase (reference: agent/barGen.py:249-252,302-335) as one
function: zero gradients (one memset of the flat bucket), forward, Loss, backward (gradient slices all-reduced over
NCCL as they complete), fused flat Adam."""
from __future__ import annotations

import os
import warnings

import torch

from . import _lib, engine
from .graph.loss.bar_loss import Loss


class DeviceBatch:
    """one batch resident on the device: what GeneratorTrainer.step_batch consumes"""
    __slots__ = ("note", "pre_note", "pre_phrase", "position", "target")

    def __init__(self, note, pre_note, pre_phrase, position, target=None):
        self.note, self.pre_note, self.pre_phrase, self.position, self.target = note, pre_note, pre_phrase, position, target

    def __iter__(self):                      # unpacks like the reference's batch tuple (agent/barGen.py:302)
        return iter((self.note, self.pre_note, self.pre_phrase, self.position))


class HostPrefetcher:
    """Iterate HOST batches (4-tuples of pinned CPU tensors as agent/barGen.py:134-141 collates them, or
    data.packed.PackedBatch) as DeviceBatch, with the NEXT batch's host->device copy already in flight on a copy
    stream when the current one is handed out -- so it overlaps the current step's kernels even in a loop that reads
    the loss back every step (agent/barGen.py:302-335 copies, steps and ``.item()``s strictly in sequence: 70.8 MB =
    1.3 ms of exposed PCIe time per 512-bar step).  Two device buffer sets are reused alternately (no allocator
    traffic, no record_stream); a set is overwritten only after the main stream has passed the point where the step
    that used it was fully enqueued (the step itself joins the phrase / weight-gradient streams before that)."""

    def __init__(self, batches, device):
        self.batches, self.device = batches, torch.device(device)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.slots = [dict(), dict()]

    def _buf(self, slot, name, shape, dtype):
        t = slot.get(name)
        if t is None or t.shape != torch.Size(shape) or t.dtype != dtype:
            t = slot[name] = torch.empty(shape, dtype=dtype, device=self.device)
        return t

    def _issue(self, hb, slot):
        """enqueue the H2D copies of host batch `hb` into `slot` on the copy stream; returns a finisher that, called on
        the consumer's stream after it waited for the copy, produces the DeviceBatch"""
        from .data.packed import BAR_BYTES, BAR_CELLS, PHRASE_CELLS, PackedBatch, unpack_bits
        cs = self.copy_stream
        free = slot.get("free")
        if free is not None:
            cs.wait_event(free)
        packed = isinstance(hb, PackedBatch)
        if packed:
            B = hb.batch
            dbits = self._buf(slot, "bits", hb.bits.shape, torch.uint8)
            pos = self._buf(slot, "pos", hb.position.shape, hb.position.dtype)
            bars = self._buf(slot, "bars", (2 * B, 1, 96, 60), torch.bfloat16)
            note32 = self._buf(slot, "note32", (B, 1, 96, 60), torch.float32)
            phrase = self._buf(slot, "phrase16", (B, 1, 384, 60), torch.bfloat16)
            with torch.cuda.stream(cs):
                dbits.copy_(hb.bits, non_blocking=True)
                pos.copy_(hb.position, non_blocking=True)

            def finish():
                unpack_bits(dbits[:2 * B * BAR_BYTES], 2 * B * BAR_CELLS, bars, note32, B * BAR_CELLS)
                unpack_bits(dbits[2 * B * BAR_BYTES:], B * PHRASE_CELLS, phrase, None, 0)
                return DeviceBatch(bars[:B], bars[B:], phrase, pos, target=note32)
        else:
            names = ("note", "pre_note", "phrase", "pos")
            dst = [self._buf(slot, n, t.shape, t.dtype) for n, t in zip(names, hb)]
            with torch.cuda.stream(cs):
                for d, t in zip(dst, hb):
                    d.copy_(t, non_blocking=True)

            def finish():
                return DeviceBatch(*dst)
        ready = torch.cuda.Event()
        ready.record(cs)
        return ready, finish

    def __iter__(self):
        it = iter(self.batches)
        try:
            pending = self._issue(next(it), self.slots[0])
        except StopIteration:
            return
        k = 0
        while pending is not None:
            ready, finish = pending
            try:
                pending = self._issue(next(it), self.slots[(k + 1) % 2])     # in flight while batch k is being used
            except StopIteration:
                pending = None
            main = torch.cuda.current_stream(self.device)
            main.wait_event(ready)
            yield finish()
            done = torch.cuda.Event()
            done.record(torch.cuda.current_stream(self.device))              # batch k's step is fully enqueued
            self.slots[k % 2]["free"] = done
            k += 1


class GeneratorTrainer:
    def __init__(self, model, lr: float = 0.002, betas=(0.9, 0.999), eps: float = 1e-8, reducer=None,
                 is_pretraining: bool = True, micro_bars: int = 0, use_graph=None):
        """``micro_bars`` > 0: a step over more bars than that is run as ceil(B / micro_bars) forward/backward passes
        whose gradients accumulate in the flat bucket before ONE all-reduce and ONE Adam step (exactly equivalent: the
        generator has no batch-coupled op and BCE-mean over the batch is the size-weighted mean of the chunk means).
        BASELINE config 3 (global batch 4096 on 2 / 4 GPUs = 2048 / 1024 bars per GPU) runs this way: the activations
        saved for backward are ~25 MB per bar."""
        self.micro_bars = int(micro_bars)
        # CUDA-graph replay of the whole step (zero-grad memset, forward, loss, backward on all streams, NCCL all-reduces,
        # Adam, operand repack: ~520 launches) -- see _graph_step.  BVAE_GRAPH=0 or use_graph=False keeps every step eager.
        # Default: on for single-process training; OFF under torch.distributed unless BVAE_GRAPH=1 / use_graph=True asks for
        # it: capturing the NCCL all-reduces works (measured: 2 GPUs, 23.8 k bars/s, same as eager), but a process that
        # still holds such a graph when the process group is destroyed hangs in NCCL teardown until the watchdog aborts it
        # (measured: 16 minutes) -- call release_graphs() before dist.destroy_process_group().
        env = os.environ.get("BVAE_GRAPH")
        multi = reducer is not None and getattr(reducer, "world", 1) > 1
        self.use_graph = ((env == "1") if (env is not None or multi) else True) if use_graph is None else bool(use_graph)
        if env == "0":
            self.use_graph = False
        self.graph_after = 3               # eager steps per input signature before capturing (kernel attributes, pack plan,
        self._graphs, self._graph_seen = {}, {}      # allocator warm-up all happen there)
        # every captured graph keeps its own memory pool (~25 MB of saved activations per bar): a loader whose batches vary
        # in size (the reference's .npz items hold different numbers of bars, agent/barGen.py:134-141) must not collect one
        # graph per size -- only the first `max_graphs` signatures that recur are captured, everything else stays eager
        self.max_graphs = 4
        self._hyper_dev = None
        self.model, self.lr, self.betas, self.eps = model, lr, betas, eps
        self.flat = model.flatten_parameters()
        self.reducer = reducer
        self.loss_fn = Loss()
        self.is_pretraining = is_pretraining
        self.step_count = 0

    def step(self, note, pre_note, pre_phrase, position, dropout_masks=None, target=None):
        """One optimisation step; returns the (device) loss tensor without synchronising.  ``target`` (default: ``note``)
        is the fp32 BCE target when ``note`` itself arrives as bf16 (bit-packed input path).  After ``graph_after`` eager
        steps with the same input signature the step is captured once as a CUDA graph and replayed from then on."""
        if self.use_graph and note.is_cuda and not engine.profiling():
            loss = self._graph_step(note, pre_note, pre_phrase, position, dropout_masks, target)
            if loss is not None:
                return loss
        return self._eager_step(note, pre_note, pre_phrase, position, dropout_masks, target)

    # ---- CUDA-graph replay --------------------------------------------------------------------------------------
    def _graph_step(self, note, pre_note, pre_phrase, position, dropout_masks, target):
        """The reference's loop (agent/barGen.py:302-335) enqueues ~2000 ATen kernels per step from Python; this path's eager
        step still issues ~520 launches through ctypes (~25 ms of host time per 42 ms step at 512 bars, the limiter at
        smaller batches).  Everything in the step is stream-ordered device work with fixed shapes, so it is captured ONCE
        per input signature -- forward and backward on the branch / weight-gradient streams (fork/join inside the
        capture), the NCCL all-reduces of the gradient segments, the fused Adam (step-dependent scalars read from device
        memory: bvae_adam_step_dev) and the operand repack -- and replayed with one cudaGraphLaunch.  Inputs are copied
        into the graph's static buffers (D2D, ~20 us for 70 MB); dropout masks are drawn inside the graph by torch's
        graph-safe Philox generator (a fresh mask every replay), injected masks are copied like inputs."""
        ins = [note, pre_note, pre_phrase, position] + list(dropout_masks or ()) + ([target] if target is not None else [])
        key = (tuple((tuple(t.shape), t.dtype) for t in ins), dropout_masks is not None, target is not None,
               self.is_pretraining, self.model.training, self.micro_bars)
        ent = self._graphs.get(key)
        if ent is None:
            seen = self._graph_seen.get(key, 0)
            if seen < self.graph_after or len(self._graphs) >= self.max_graphs:
                self._graph_seen[key] = seen + 1
                return None
            ent = self._capture(key, ins, dropout_masks is not None, target is not None)
            if ent is None:
                return None
        graph, static, loss, n_launch = ent
        for s, t in zip(static, ins):
            if s.data_ptr() != t.data_ptr():
                s.copy_(t, non_blocking=True)
        self.step_count += 1
        scale = 1.0 / self.reducer.world if self.reducer is not None and self.reducer.world > 1 else 1.0
        engine.adam_hyper_upload(self._hyper_dev, self.lr, self.step_count, self.betas, self.eps, scale)
        graph.replay()
        _lib.load().bvae_launch_count_add(n_launch)     # the replay launches what the capture recorded
        return loss.clone()          # the static loss buffer is overwritten by the next replay

    def release_graphs(self):
        """drop every captured graph (and its private memory pool); required before dist.destroy_process_group() when the
        graphs contain NCCL collectives"""
        if self._graphs:
            torch.cuda.synchronize()
            self._graphs.clear()
            self._graph_seen.clear()
            torch.cuda.synchronize()

    def _capture(self, key, ins, has_masks, has_target):
        dev = self.flat.data.device
        if self._hyper_dev is None:
            self._hyper_dev = torch.zeros(8, dtype=torch.float32, device=dev)
        static = [torch.empty_like(t).copy_(t) for t in ins]
        n_m = 2 if has_masks else 0
        masks = tuple(static[4:4 + n_m]) if has_masks else None
        target = static[4 + n_m] if has_target else None
        # the capture itself performs one real step: give it this step's scalars
        scale = 1.0 / self.reducer.world if self.reducer is not None and self.reducer.world > 1 else 1.0
        engine.adam_hyper_upload(self._hyper_dev, self.lr, self.step_count + 1, self.betas, self.eps, scale)
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count()
        try:
            with torch.cuda.graph(graph):
                loss = self._eager_step(static[0], static[1], static[2], static[3], masks, target, hyper_dev=self._hyper_dev)
        except Exception as exc:                  # loud, and the eager path keeps working
            self.use_graph = False
            warnings.warn("CUDA-graph capture of the training step failed (%s: %s); continuing with eager steps"
                          % (type(exc).__name__, exc))
            return None
        # capturing does not execute: the step this call stands for is the first replay (done by the caller)
        self.step_count -= 1
        n_launch = _lib.launch_count() - n0
        _lib.load().bvae_launch_count_add(-n_launch & 0xFFFFFFFFFFFFFFFF)      # recorded, not executed
        ent = self._graphs[key] = (graph, static, loss, n_launch)
        return ent

    def _eager_step(self, note, pre_note, pre_phrase, position, dropout_masks=None, target=None, hyper_dev=None):
        self.flat.attach_grads(zero=True)
        B = note.shape[0]
        if self.micro_bars and B > self.micro_bars:
            loss = self._accumulate(note, pre_note, pre_phrase, position, dropout_masks, target)
        else:
            gen, z, pre_z, pf = self.model(note, pre_note, pre_phrase, position, True, dropout_masks)
            loss = self.loss_fn(gen, note if target is None else target, self.is_pretraining)
            loss.backward()
        scale = self.reducer.finish() if self.reducer is not None else 1.0
        self.step_count += 1
        if hyper_dev is not None:
            engine.adam_step_dev(self.flat, hyper_dev)
        else:
            engine.adam_step(self.flat, self.lr, self.step_count, self.betas, self.eps, scale)
        return loss.detach()

    def _accumulate(self, note, pre_note, pre_phrase, position, dropout_masks, target):
        """gradient accumulation over chunks of ``micro_bars`` bars; the reducer's per-segment all-reduces are held back
        until the last chunk's backward pass"""
        B, mb = note.shape[0], self.micro_bars
        bce_tot = torch.zeros((), device=note.device)
        cnt_tot = torch.zeros((), device=note.device)
        starts = list(range(0, B, mb))
        for i, s in enumerate(starts):
            e = min(B, s + mb)
            if self.reducer is not None:
                self.reducer.hold = i + 1 < len(starts)
            masks = None if dropout_masks is None else tuple(m[s:e] for m in dropout_masks)
            gen = self.model(note[s:e], pre_note[s:e], pre_phrase[s:e], position[s:e], True, masks)[0]
            tgt = (note if target is None else target)[s:e]
            bce, cnt = self.loss_fn.parts(gen, tgt, self.is_pretraining)
            w = (e - s) / B
            (bce * w).backward()
            bce_tot += bce.detach() * w
            cnt_tot += cnt.detach()
        return bce_tot + 0.005 * cnt_tot

    def step_from_host(self, note, pre_note, pre_phrase, position, dropout_masks=None):
        """The same step from (pinned) HOST tensors, as the reference's loop feeds it (agent/barGen.py:302-311 moves
        every batch with .cuda()).  The phrase tensor -- two thirds of the bytes -- is copied on the phrase encoder's
        stream, so the bar encoder starts as soon as its own inputs have arrived."""
        dev = self.flat.data.device
        note_d = note.to(dev, non_blocking=True)
        pre_d = pre_note.to(dev, non_blocking=True)
        pos_d = position.to(dev, non_blocking=True)
        side = self.model.side_stream(dev) if hasattr(self.model, "side_stream") else None
        if side is None:
            phrase_d = pre_phrase.to(dev, non_blocking=True)
        else:
            with torch.cuda.stream(side):
                phrase_d = pre_phrase.to(dev, non_blocking=True)
            if self.use_graph:          # a graph replay reads its inputs on the main stream (copy into static buffers)
                torch.cuda.current_stream().wait_stream(side)
        return self.step(note_d, pre_d, phrase_d, pos_d, dropout_masks)

    def prefetch(self, host_batches):
        """``for batch in trainer.prefetch(loader): loss = trainer.step_batch(batch)`` -- see HostPrefetcher"""
        pf = getattr(self, "_prefetcher", None)
        if pf is None:
            pf = self._prefetcher = HostPrefetcher(host_batches, self.flat.data.device)
        pf.batches = host_batches           # one copy stream and one pair of device buffer sets per trainer
        return pf

    def step_batch(self, batch: DeviceBatch, dropout_masks=None):
        return self.step(batch.note, batch.pre_note, batch.pre_phrase, batch.position, dropout_masks,
                         target=batch.target)

    def step_from_packed(self, packed, dropout_masks=None):
        """The same step from a bit-packed host batch (data/packed.py: 4320 B per sample instead of 138 KB): ONE H2D
        copy of the bits, expanded on the device by bvae_unpack_bits into the bf16 encoder inputs (note and pre_note
        contiguous, so Model.forward's torch.cat is a view) and the fp32 BCE target; the phrase bits are expanded on the
        phrase encoder's stream."""
        dev = self.flat.data.device
        side = self.model.side_stream(dev) if hasattr(self.model, "side_stream") else None
        note_f32, bars, phrase, pos, dbits = packed.to_device(dev, side)
        if self.use_graph and side is not None:
            torch.cuda.current_stream().wait_stream(side)
        B = packed.batch
        loss = self.step(bars[:B], bars[B:], phrase, pos, dropout_masks, target=note_f32)
        del dbits        # held until here: the main stream has joined the phrase stream inside the step
        return loss

    # torch.optim-style state for checkpoints (agent/barGen.py:174-197)
    def state_dict(self):
        return {"step": self.step_count, "lr": self.lr,
                "exp_avg": None if self.flat.exp_avg is None else self.flat.exp_avg.clone(),
                "exp_avg_sq": None if self.flat.exp_avg_sq is None else self.flat.exp_avg_sq.clone()}

    def torch_state_dict(self):
        """The same state in ``torch.optim.Adam.state_dict()`` form -- what the reference stores under
        ``gen_optimizer1`` (agent/barGen.py:60,176): ``state[i] = {step, exp_avg, exp_avg_sq}`` for the i-th entry of
        ``generator.parameters()``, one param group.  Parameters that never received a gradient (the unused ``bn1``
        affine pairs) carry zero moments here; torch would have no entry for them."""
        flat, state = self.flat, {}
        if flat.exp_avg is not None:
            for i, (p, o) in enumerate(zip(flat.params, flat.offsets)):
                n = p.numel()
                state[i] = {"step": torch.tensor(float(self.step_count)),
                            "exp_avg": flat.exp_avg[o:o + n].view(p.shape).clone(),
                            "exp_avg_sq": flat.exp_avg_sq[o:o + n].view(p.shape).clone()}
        group = {"lr": self.lr, "betas": tuple(self.betas), "eps": self.eps, "weight_decay": 0, "amsgrad": False,
                 "params": list(range(len(flat.params)))}
        return {"state": state, "param_groups": [group]}

    def _load_torch_state_dict(self, sd):
        """``torch.optim.Adam.state_dict()`` of the reference's generator optimiser.  Entries are matched by position in
        ``generator.parameters()`` and must agree in shape; trailing entries this model does not have (the reference's
        Refiner, graph/model.py:18) are ignored.  The flat Adam keeps ONE step counter (bias correction), so per-parameter
        counters collapse to their maximum -- they are all equal in the reference's loop except for parameters that
        never get a gradient."""
        flat = self.flat
        group = sd["param_groups"][0]
        self.lr = group.get("lr", self.lr)
        self.betas = tuple(group.get("betas", self.betas))
        self.eps = group.get("eps", self.eps)
        ids = list(group["params"])
        dev = flat.data.device
        m, v, step = torch.zeros_like(flat.data), torch.zeros_like(flat.data), 0
        for i, (p, o) in enumerate(zip(flat.params, flat.offsets)):
            if i >= len(ids):
                break
            st = sd["state"].get(ids[i])
            if st is None:
                continue
            if tuple(st["exp_avg"].shape) != tuple(p.shape):
                raise ValueError("optimizer state %d has shape %s, parameter %d has %s" %
                                 (ids[i], tuple(st["exp_avg"].shape), i, tuple(p.shape)))
            n = p.numel()
            m[o:o + n].copy_(st["exp_avg"].reshape(-1).to(dev))
            v[o:o + n].copy_(st["exp_avg_sq"].reshape(-1).to(dev))
            step = max(step, int(st["step"]))
        flat.exp_avg, flat.exp_avg_sq, self.step_count = m, v, step

    def load_state_dict(self, sd):
        if "param_groups" in sd and "state" in sd:
            return self._load_torch_state_dict(sd)
        self.step_count, self.lr = sd["step"], sd["lr"]
        if sd["exp_avg"] is not None:
            self.flat.exp_avg = sd["exp_avg"].to(self.flat.data.device).clone()
            self.flat.exp_avg_sq = sd["exp_avg_sq"].to(self.flat.data.device).clone()
