"""The generator training step of the pre-training phase (reference: agent/barGen.py:249-252,302-335) as one
function: zero gradients (one memset of the flat bucket), forward, Loss, backward (gradient slices all-reduced over
NCCL as they complete), fused flat Adam."""
from __future__ import annotations

import torch

from . import engine
from .graph.loss.bar_loss import Loss


class GeneratorTrainer:
    def __init__(self, model, lr: float = 0.002, betas=(0.9, 0.999), eps: float = 1e-8, reducer=None,
                 is_pretraining: bool = True):
        self.model, self.lr, self.betas, self.eps = model, lr, betas, eps
        self.flat = model.flatten_parameters()
        self.reducer = reducer
        self.loss_fn = Loss()
        self.is_pretraining = is_pretraining
        self.step_count = 0

    def step(self, note, pre_note, pre_phrase, position, dropout_masks=None):
        """One optimisation step; returns the (device) loss tensor without synchronising."""
        self.flat.attach_grads(zero=True)
        gen, z, pre_z, pf = self.model(note, pre_note, pre_phrase, position, True, dropout_masks)
        loss = self.loss_fn(gen, note, self.is_pretraining)
        loss.backward()
        scale = self.reducer.finish() if self.reducer is not None else 1.0
        self.step_count += 1
        engine.adam_step(self.flat, self.lr, self.step_count, self.betas, self.eps, scale)
        return loss.detach()

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
        return self.step(note_d, pre_d, phrase_d, pos_d, dropout_masks)

    # torch.optim-style state for checkpoints (agent/barGen.py:174-197)
    def state_dict(self):
        return {"step": self.step_count, "lr": self.lr,
                "exp_avg": None if self.flat.exp_avg is None else self.flat.exp_avg.clone(),
                "exp_avg_sq": None if self.flat.exp_avg_sq is None else self.flat.exp_avg_sq.clone()}

    def load_state_dict(self, sd):
        self.step_count, self.lr = sd["step"], sd["lr"]
        if sd["exp_avg"] is not None:
            self.flat.exp_avg = sd["exp_avg"].to(self.flat.data.device).clone()
            self.flat.exp_avg_sq = sd["exp_avg_sq"].to(self.flat.data.device).clone()
