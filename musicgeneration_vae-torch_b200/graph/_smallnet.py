"""Operator-level autograd nodes for the two small networks of the adversarial phase -- the convolutional piano-roll
discriminator (graph/bar_discriminator.py) and the Refiner (graph/refiner.py): a convolution / transposed convolution /
Linear through bvae_conv_gemm + bvae_wgrad_gemm (tcgen05 when the channel counts allow it, the CUDA-core implicit-GEMM
kernels for the 1..16-channel layers) and nn.BatchNorm2d (+ activation) through bvae_bn_forward / bvae_bn_backward.

Unlike the generator (one autograd node per trunk) these nets are 0.3 % of the step's FLOPs on tensors of <= 64 channels,
so each operator is its own node over NHWC torch tensors and the cheap glue in between (2x2 max-pool, adds, sigmoid, the
final 192 -> 1 logit) stays PyTorch, like the generator's embedding lookup and dropout masks.  Parameter gradients are
accumulated by the kernels straight into ``.grad`` (engine.grad_ptr); frozen parameters (``requires_grad=False`` while the
generator trains against the discriminator, agent/barGen_with_gan.py:506-511) skip the weight gradient and still propagate
the input gradient."""
import ctypes as C

import torch

from .. import _lib
from ..engine import BF16, Act, grad_ptr
from .encodingBlock import gemm_of


class _ConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, layer, act, slope, out_f32, use_bias, bias_grad, *params):
        N, H, W, Cin = x.shape
        xa = Act(x.detach(), N, H, W, Cin)
        oh, ow = layer.out_hw(H, W)
        y = Act.empty(N, oh, ow, layer.Cout, dtype=torch.float32 if out_f32 else BF16, device=x.device)
        layer.forward(xa, y, act=act, slope=slope, use_bias=use_bias)
        ctx.layer, ctx.act, ctx.slope, ctx.use_bias, ctx.bias_grad = layer, act, slope, use_bias, bias_grad
        ctx.save_for_backward(x.detach(), y.t if act else None)
        return y.t

    @staticmethod
    def backward(ctx, dy):
        x, y = ctx.saved_tensors
        layer = ctx.layer
        N, H, W, Cin = x.shape
        d = dy.float()
        if ctx.act:                                     # (Leaky)ReLU backward on the conv output (glue: <= 64 channels)
            d = d * torch.where(y > 0, 1.0, ctx.slope)
        d = d.to(BF16).contiguous()
        dya = Act(d, d.shape[0], d.shape[1], d.shape[2], d.shape[3])
        xa = Act(x, N, H, W, Cin)
        layer.wgrad(xa, dya)
        if ctx.use_bias and ctx.bias_grad:
            layer.bias_grad(dya)
        elif ctx.use_bias:
            layer.zero_bias_grad()      # bias in front of a batch-statistics norm: the gradient is analytically zero
        dx = None
        if ctx.needs_input_grad[0]:
            dxa = Act.empty(N, H, W, Cin, device=x.device)
            layer.dgrad(dya, dxa)
            dx = dxa.t
        return (dx, None, None, None, None, None, None) + (None,) * (len(ctx.needs_input_grad) - 7)


def conv(x, module, act=False, slope=0.0, out_f32=True, use_bias=True, bias_grad=True):
    """x: [N,H,W,Cin] (any float dtype; stored as bf16 for the GEMM).  Returns [N,OH,OW,Cout] fp32 (raw output in front
    of a BatchNorm) or bf16 (with the activation fused in the epilogue).  ``bias_grad=False``: the bias feeds a BatchNorm
    running on batch statistics, which removes any per-channel constant -- its gradient is exactly zero (the reference
    yields ~1e-9 of fp32 noise there; summing the bf16-stored BatchNorm gradient would give 1e-2 of rounding noise, which
    Adam's normalisation would turn into a random walk of the bias)."""
    if not x.is_cuda:
        raise RuntimeError("the B200 path needs CUDA tensors; there is no CPU fallback")
    layer = gemm_of(module)
    params = [p for p in (module.weight, module.bias) if p is not None]
    return _ConvFn.apply(x.to(BF16).contiguous(), layer, act, slope, out_f32, use_bias, bias_grad, *params)


def linear(x, module, act=False):
    """x: [N, in] -> [N, out] fp32 (bias + optional ReLU in the GEMM epilogue)"""
    N = x.shape[0]
    y = conv(x.reshape(N, 1, 1, -1), module, act=act, slope=0.0, out_f32=not act)
    return y.reshape(N, -1)


class _BNFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, bn, act, slope, gamma, beta):
        N, H, W, Cc = x.shape
        dev = x.device
        x = x.contiguous()
        y = torch.empty((N, H, W, Cc), dtype=BF16, device=dev)
        save = torch.empty(4 * Cc, dtype=torch.float32, device=dev)
        lib = _lib.lib()
        scratch = torch.empty(lib.bvae_bn_scratch_floats(Cc), dtype=torch.float32, device=dev)
        d = _lib.BnDesc()
        d.P, d.C, d.x_pitch, d.y_pitch = N * H * W, Cc, Cc, Cc
        d.x_f32, d.y_f32, d.act, d.slope = int(x.dtype == torch.float32), 0, int(act), float(slope)
        d.training = int(bn.training or not bn.track_running_stats)
        d.eps, d.momentum = float(bn.eps), float(bn.momentum if bn.momentum is not None else 0.1)
        d.x, d.y, d.gamma, d.beta = x.data_ptr(), y.data_ptr(), gamma.data_ptr(), beta.data_ptr()
        if bn.track_running_stats:
            d.running_mean, d.running_var = bn.running_mean.data_ptr(), bn.running_var.data_ptr()
        d.save, d.scratch = save.data_ptr(), scratch.data_ptr()
        _lib.check(lib.bvae_bn_forward(C.byref(d), _lib.stream_ptr()), "bn_forward")
        if bn.training and bn.track_running_stats:
            bn.num_batches_tracked += 1
        ctx.bn, ctx.desc, ctx.keep = bn, d, (x, y, save, scratch)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y, save, scratch = ctx.keep
        bn, d = ctx.bn, ctx.desc
        dy = dy.contiguous()
        if dy.dtype not in (torch.float32, BF16):
            dy = dy.float()
        dx = torch.empty(x.shape, dtype=BF16, device=x.device)
        d.dy, d.dy_pitch, d.dy_f32 = dy.data_ptr(), d.C, int(dy.dtype == torch.float32)
        d.dx, d.dx_pitch = dx.data_ptr(), d.C
        d.dgamma, d.dbeta = grad_ptr(bn.weight), grad_ptr(bn.bias)
        _lib.check(_lib.lib().bvae_bn_backward(C.byref(d), _lib.stream_ptr()), "bn_backward")
        ctx.keep = None
        return dx, None, None, None, None, None


def batch_norm(x, bn, act=False, slope=0.0):
    """nn.BatchNorm2d (batch statistics in training mode, running statistics in eval) + optional (Leaky)ReLU on an NHWC
    tensor; returns bf16 NHWC.  The running statistics / num_batches_tracked of ``bn`` are updated like the module does."""
    if not x.is_cuda:
        raise RuntimeError("the B200 path needs CUDA tensors; there is no CPU fallback")
    if x.dtype not in (torch.float32, BF16):
        x = x.float()
    return _BNFn.apply(x, bn, act, slope, bn.weight, bn.bias)
