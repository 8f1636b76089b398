"""Encoder blocks (reference: graph/encodingBlock.py).  Each block owns the same sub-modules / parameter names as
the reference and exposes ``fwd`` / ``bwd`` in terms of engine.Act views; the arithmetic runs in libbarvae.so."""
import torch.nn as nn

from ..engine import Act, GemmLayer, NormBlock, raw_dtype
from .cbam import CBAM
from .weights_initializer import weights_init


def _norm(c):
    return nn.InstanceNorm2d(c, eps=1e-5, momentum=0.01, affine=True)


def gemm_of(m) -> GemmLayer:
    """engine.GemmLayer for an nn.Conv2d / nn.ConvTranspose2d / nn.Linear parameter holder (cached on the module)."""
    g = getattr(m, "_bvae_gemm", None)
    if g is None:
        if isinstance(m, nn.Linear):
            g = GemmLayer("linear", m.weight, m.bias)
        else:
            kind = "convT" if isinstance(m, nn.ConvTranspose2d) else "conv"
            op = tuple(m.output_padding) if kind == "convT" else (0, 0)
            g = GemmLayer(kind, m.weight, m.bias, tuple(m.kernel_size), tuple(m.stride), tuple(m.padding), op)
        m._bvae_gemm = g
    return g


def norm_block(bn, cbam, res_mode, slope) -> NormBlock:
    return NormBlock(bn.weight.shape[0], bn.weight, bn.bias, cbam.params() if cbam is not None else None, res_mode,
                     slope)


class _StemModule(nn.Module):
    """conv(1->32) -> LeakyReLU -> conv(32->32) -> IN -> out + CBAM(out) -> LeakyReLU  (encodingBlock.py:25-36,56-67)"""
    SLOPE = 0.01
    FIRST = SECOND = None

    def _build(self, first, second):
        spec = {"time": dict(kernel_size=(4, 1), stride=(2, 1), padding=[1, 0]),
                "pitch": dict(kernel_size=(1, 4), stride=(1, 2), padding=[0, 1])}
        setattr(self, first, nn.Conv2d(1, 32, bias=False, **spec[first]))
        setattr(self, second, nn.Conv2d(32, 32, bias=False, **spec[second]))
        self.bn = _norm(32)
        self.cbam = CBAM(32)
        self.apply(weights_init)

    def fwd(self, x: Act, out: Act):
        g1, g2 = gemm_of(getattr(self, self.FIRST)), gemm_of(getattr(self, self.SECOND))
        h1, w1 = g1.out_hw(x.H, x.W)
        t1 = Act.empty(x.N, h1, w1, 32)
        g1.forward(x, t1, act=True, slope=self.SLOPE)
        h2, w2 = g2.out_hw(h1, w1)
        y = Act.empty(x.N, h2, w2, 32, dtype=raw_dtype())
        st = g2.forward(t1, y, want_stats=True)
        nb = norm_block(self.bn, self.cbam, 1, self.SLOPE)
        return (x, t1, nb, nb.forward(y, out, stats=st))

    def bwd(self, ctx, dout: Act):
        x, t1, nb, nctx = ctx
        g1, g2 = gemm_of(getattr(self, self.FIRST)), gemm_of(getattr(self, self.SECOND))
        dy = Act.empty(dout.N, dout.H, dout.W, 32)
        nb.backward(nctx, dout, dy)
        g2.wgrad(t1, dy)
        dt1 = Act.empty(t1.N, t1.H, t1.W, 32)
        g2.dgrad(dy, dt1, mask=t1, mask_slope=self.SLOPE)
        g1.wgrad(x, dt1)


class TimePitchModule(_StemModule):
    FIRST, SECOND = "time", "pitch"

    def __init__(self):
        super().__init__()
        self._build("time", "pitch")


class PitchTimeModule(_StemModule):
    FIRST, SECOND = "pitch", "time"

    def __init__(self):
        super().__init__()
        self._build("pitch", "time")


class ResidualModule(nn.Module):
    """relu(x + CBAM(IN(conv2(relu(conv1(x))))))  (encodingBlock.py:87-100)"""

    def __init__(self, channel):
        super().__init__()
        self.conv1 = nn.Conv2d(channel, channel, 3, 1, 1, bias=False)
        self.conv2 = nn.Conv2d(channel, channel, 3, 1, 1, bias=False)
        self.bn = _norm(channel)
        self.cbam = CBAM(channel)
        self.apply(weights_init)

    def out_shape(self, H, W):
        return H, W, self.conv1.out_channels

    def fwd(self, x: Act, out: Act):
        g1, g2 = gemm_of(self.conv1), gemm_of(self.conv2)
        c1 = Act.empty(x.N, x.H, x.W, x.C)
        g1.forward(x, c1, act=True, slope=0.0)
        y = Act.empty(x.N, x.H, x.W, x.C, dtype=raw_dtype())
        st = g2.forward(c1, y, want_stats=True)
        nb = norm_block(self.bn, self.cbam, 2, 0.0)
        return (x, c1, nb, nb.forward(y, out, res=x, stats=st))

    def bwd(self, ctx, dout: Act) -> Act:
        x, c1, nb, nctx = ctx
        g1, g2 = gemm_of(self.conv1), gemm_of(self.conv2)
        dy = Act.empty(x.N, x.H, x.W, x.C)
        dres = Act.empty(x.N, x.H, x.W, x.C)
        nb.backward(nctx, dout, dy, dres)
        g2.wgrad(c1, dy)
        dc1 = Act.empty(x.N, x.H, x.W, x.C)
        g2.dgrad(dy, dc1, mask=c1, mask_slope=0.0)
        g1.wgrad(x, dc1)
        g1.dgrad(dc1, dres, addend=dres)        # dx = conv1^T(dc1) + skip gradient, written over dres
        return dres


class PoolingModule(nn.Module):
    """relu(y + CBAM(y)), y = IN(conv3x3 s2 (x))  (encodingBlock.py:118-126)"""

    def __init__(self, in_channel, out_channel):
        super().__init__()
        self.conv = nn.Conv2d(in_channel, out_channel, 3, 2, 1, bias=False)
        self.bn = _norm(out_channel)
        self.cbam = CBAM(out_channel)
        self.apply(weights_init)

    def out_shape(self, H, W):
        h, w = gemm_of(self.conv).out_hw(H, W)
        return h, w, self.conv.out_channels

    def fwd(self, x: Act, out: Act):
        g = gemm_of(self.conv)
        y = Act.empty(out.N, out.H, out.W, out.C, dtype=raw_dtype())
        st = g.forward(x, y, want_stats=True)
        nb = norm_block(self.bn, self.cbam, 1, 0.0)
        return (x, nb, nb.forward(y, out, stats=st))

    def bwd(self, ctx, dout: Act) -> Act:
        x, nb, nctx = ctx
        g = gemm_of(self.conv)
        dy = Act.empty(dout.N, dout.H, dout.W, dout.C)
        nb.backward(nctx, dout, dy)
        g.wgrad(x, dy)
        dx = Act.empty(x.N, x.H, x.W, x.C)
        g.dgrad(dy, dx)
        return dx
