"""Reconstruction loss (reference: graph/loss/bar_loss.py:7-42).

``Loss()(probs, labels, is_pretraining)`` = BCE(mean, each log clamped at -100) on the post-sigmoid probabilities,
with the pitch-prior label smoothing when not pre-training, plus the constant 0.005 * #missed-notes term.  One
single-pass reduction kernel forward; backward is autograd's BCELoss gradient, (p - t') / max(p(1-p), 1e-12) / N,
which the decoder then folds through the sigmoid and fit2 in bvae_fit_sigmoid_bce_bwd.  Unlike the reference this
class has no hard-coded ``.cuda()`` in its constructor."""
import torch
import torch.nn as nn

from ... import _lib


class _BCEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, probs, labels, smoothing, with_count):
        p = probs.contiguous().float()
        t = labels.contiguous().float()
        if not p.is_cuda:
            raise RuntimeError("the B200 path needs CUDA tensors; there is no CPU fallback")
        rows = p.numel()
        acc = torch.zeros(2, device=p.device)
        _lib.check(_lib.lib().bvae_bce_fwd(p.data_ptr(), t.data_ptr(), rows, int(smoothing), acc.data_ptr(),
                                           _lib.stream_ptr()), "bce_fwd")
        ctx.save_for_backward(p, t)
        ctx.smoothing, ctx.shape = int(smoothing), probs.shape
        ctx.parts = with_count == 2
        if ctx.parts:                        # [BCE mean, missed-note count] (micro-batched steps weight them differently)
            return acc
        return acc[0] + acc[1] * 0.005 if with_count else acc[0]

    @staticmethod
    def backward(ctx, g):
        p, t = ctx.saved_tensors
        d = torch.empty_like(p)
        _lib.check(_lib.lib().bvae_bce_bwd(p.data_ptr(), t.data_ptr(), p.numel(), ctx.smoothing, 1.0 / p.numel(),
                                           d.data_ptr(), _lib.stream_ptr()), "bce_bwd")
        if ctx.parts:
            g = g[0]                         # the count is not differentiable (graph/loss/bar_loss.py:31-32)
        return (d * g).view(ctx.shape), None, None, None


class Loss(nn.Module):
    def forward(self, logits, labels, is_pretraining=False):
        return _BCEFn.apply(logits, labels, not is_pretraining, True)

    def parts(self, logits, labels, is_pretraining=False):
        """(BCE mean, missed-note count) of the same single-pass kernel: ``forward`` == parts[0] + 0.005 * parts[1]"""
        acc = _BCEFn.apply(logits, labels, not is_pretraining, 2)
        return acc[0], acc[1]


class DLoss(nn.Module):
    def forward(self, outputs, targets):
        return _BCEFn.apply(outputs, targets, False, False)


class VAELoss(nn.Module):
    """BCE + KL of the optional VAE head: old/graphs/losses/bar_loss.py:10-18 (KL of note and pre_note averaged)."""

    def forward(self, recon, labels, mu, logvar, pre_mu=None, pre_logvar=None):
        from ..model import kl_divergence
        loss = _BCEFn.apply(recon, labels, False, False)
        kl = kl_divergence(mu, logvar)
        if pre_mu is not None:
            kl = (kl + kl_divergence(pre_mu, pre_logvar)) / 2
        return loss + kl
