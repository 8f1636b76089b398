"""Shared autograd node for the fully connected discriminators (graph/z_discriminator.py,
graph/bar_discriminator_with_feature.py): a stack of ``nn.Linear`` (+ optional ReLU) run as tcgen05 GEMMs through
libbarvae.so (bias + ReLU in the epilogue; weight / bias / data gradients by the same kernels as the generator's Linear
layers), followed by the 512 -> 1 logit and the sigmoid, which stay PyTorch glue on a [B,512] tensor like the
generator's embedding lookup and dropout masks (512 MACs per sample; a 1-wide output is not a tensor-core shape)."""
import torch
import torch.nn as nn

from ..engine import BF16, Act
from .encodingBlock import gemm_of


class _HiddenFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, stack, x, *params):
        B = x.shape[0]
        h = Act(x.detach().to(BF16).contiguous(), B, 1, 1, x.shape[1])
        acts = [h]
        for lin, relu in stack:
            out = Act.empty(B, 1, 1, lin.out_features)
            gemm_of(lin).forward(h, out, act=relu, slope=0.0)
            acts.append(out)
            h = out
        ctx.stack, ctx.acts = stack, acts
        return h.t.view(B, -1).float()

    @staticmethod
    def backward(ctx, dh):
        stack, acts = ctx.stack, ctx.acts
        B = dh.shape[0]
        d = dh.contiguous().float()
        for i in range(len(stack) - 1, -1, -1):
            lin, relu = stack[i]
            g = gemm_of(lin)
            if relu:                                              # ReLU backward on a [B,512] tensor (glue)
                d = d * (acts[i + 1].t.view(B, -1) > 0)
            dy = Act(d.to(BF16).contiguous(), B, 1, 1, lin.out_features)
            g.wgrad(acts[i], dy)
            g.bias_grad(dy)
            dx = Act.empty(B, 1, 1, lin.in_features)
            g.dgrad(dy, dx)
            d = dx.t.view(B, -1).float()
        ctx.acts = None
        return (None, d) + (None,) * (len(ctx.needs_input_grad) - 2)


class MLPDiscriminator(nn.Module):
    """``_stack()`` -> [(nn.Linear, relu?) ...] hidden layers, ``_head()`` -> the final nn.Linear(., 1)"""

    def _stack(self):
        raise NotImplementedError

    def _head(self):
        raise NotImplementedError

    def _run(self, x):
        if not x.is_cuda:
            raise RuntimeError("the B200 path needs CUDA tensors; there is no CPU fallback")
        stack = self._stack()
        params = [p for lin, _ in stack for p in lin.parameters()]
        h = _HiddenFn.apply(stack, x, *params)
        return torch.sigmoid(self._head()(h))
