"""Storage-precision emulation of the CUDA path on the CPU  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

``barvae_oracle.py`` is the reference algorithm in fp32.  The CUDA path computes the same algorithm with bf16 GEMM
operands and bf16 *stored* activations / inter-layer gradients, fp32 accumulation, fp32 raw convolution outputs in front
of every InstanceNorm, fp32 statistics, gates, losses and parameters.  This network's backward
pass amplifies forward perturbations ~1e4x (max-pool / arg-max routing, ReLU masks), so against the fp32 oracle a
gradient tensor legitimately moves by tens of per cent.  This file restates the oracle with the roundings put at
exactly the storage points of the kernels, so that the masks and arg-max routes agree and gradients can be held
tightly (tests/test_gpu_model.py::test_all_gradients_vs_storage_emulation):

  q(x)   value rounded to bf16, incoming gradient rounded to bf16   -- an activation the kernels store as bf16
  gq(x)  value untouched, incoming gradient rounded to bf16         -- a raw conv output (fp32) whose gradient is bf16
  wq(w)  value rounded to bf16, gradient passed through             -- a packed bf16 GEMM operand of an fp32 master

The norm block (InstanceNorm [+CBAM] [+residual] + activation, csrc/normblock.cu) is one autograd Function: its forward
is the oracle's arithmetic on the fp32 raw output; its backward differentiates the same expression rebuilt from what the
kernels SAVE -- the bf16 normalised activation, the bf16 block output (activation mask), the fp32-chosen arg-max
positions -- and then applies the InstanceNorm projection by hand.  With ``exact=True`` every rounding is the identity
and the arg-max / mask choices are the fp32 ones, so the whole file reduces to the fp32 oracle: that mode is checked
against ``barvae_oracle`` on the CPU (tests/test_oracle_golden.py::test_emulation_exact_mode_is_the_oracle), which pins
the hand-written backward.

Teacher forcing (``TEACH``).  Two implementations that both store bf16 activations drift apart layer by layer even with
identical rounding points: an fp32 summation-order difference of 1e-7 flips a bf16 rounding here and there, the flipped
values feed the next layer, and after a few layers a large fraction of the elements differ by one bf16 unit (measured on
the B200: 5e-5 of the elements after the stems, 11 % after the third block, 48 % after the seventh; z differs by 3e-3 --
the bf16 floor itself).  The backward pass then sees different ReLU masks / arg-max routes and single gradient tensors
move by tens of per cent.  To hold the BACKWARD composition tightly all the same, the test hands this file the stored
forward state of the CUDA run (``TEACH``: every stored activation, every norm block's saved tensors and arg-max
positions).  Each layer is then evaluated here from the CUDA path's own inputs -- its local result is compared with what
the CUDA path stored (``LOCAL``: per-layer forward parity, independent of drift) -- and replaced by the stored value, so
the backward pass runs on exactly the masks, routes and operands the kernels used.

Follows the same reference lines as barvae_oracle.py (cited there per function)."""
from __future__ import annotations

from collections import OrderedDict
from typing import Optional, Tuple

import torch
import torch.nn.functional as F

import barvae_oracle as O

Tensor = torch.Tensor
BF16 = torch.bfloat16

_EXACT = [False]
TRACE = None        # list: record (tag, tensor) of every stored activation, in forward order (debugging)
TEACH = None        # dict tag -> tensor (stored activation) or dict (norm-block state) taken from the CUDA run
LOCAL = None        # list of (tag, what, fraction of elements off by more than one bf16 unit, max |diff| in units)


class exact:
    """``with exact():`` -- all roundings off (the file then IS the fp32 oracle; used to pin the hand-written backward)"""

    def __enter__(self):
        self.prev, _EXACT[0] = _EXACT[0], True

    def __exit__(self, *a):
        _EXACT[0] = self.prev
        return False


def _r(t: Tensor) -> Tensor:
    return t if _EXACT[0] else t.to(BF16).float()


class _Q(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t):
        return _r(t)

    @staticmethod
    def backward(ctx, g):
        return _r(g)


class _GQ(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t):
        return t.view_as(t)

    @staticmethod
    def backward(ctx, g):
        return _r(g)


def q(t):
    return _Q.apply(t)


def gq(t):
    return _GQ.apply(t)


def wq(w):
    return w + (_r(w) - w).detach()


def _local(tag, what, mine, theirs):
    """per-layer forward parity under teacher forcing: how far is this file's result, computed from the CUDA path's own
    inputs, from what the CUDA path stored?  Unit = one bf16 step of the larger magnitude (2^-8 relative) plus 1e-6 of the
    tensor's rms (sums that cancel to ~0 carry the absolute fp32 noise of their terms)."""
    if LOCAL is None:
        return
    a, b = mine.detach().float(), theirs.detach().float()
    unit = torch.maximum(a.abs(), b.abs()) * 2.0 ** -8 + 1e-6 * float(b.pow(2).mean().sqrt()) + 1e-30
    r = (a - b).abs() / unit
    LOCAL.append((tag, what, float((r > 1.001).float().mean()), float(r.max())))


def _store(tag, t):
    """a stored activation: record it; under teacher forcing compare and substitute the CUDA path's value (the
    gradient still flows into this file's expression)"""
    if TRACE is not None:
        TRACE.append((tag, t.detach()))
    if TEACH is not None and tag in TEACH:
        _local(tag, "value", t, TEACH[tag])
        t = t + (TEACH[tag].to(t.dtype) - t).detach()
    return t


# ---------------------------------------------------------------------------------------------
# norm block: InstanceNorm (+CBAM) (+residual) + activation, forward as the kernels evaluate it, backward from what
# they save (graph/cbam.py:22-68, graph/encodingBlock.py:30-36,92-100,120-126, graph/decoder.py:30-36,93-109,137-154)
# ---------------------------------------------------------------------------------------------

def _fix(t, fixed):
    """value ``fixed`` (a saved forward quantity), gradient of ``t``"""
    return t if fixed is None else t + (fixed - t).detach()


class _SigmoidSaved(torch.autograd.Function):
    """sigmoid whose value AND derivative come from the gate the forward pass saved: s, s(1-s)"""

    @staticmethod
    def forward(ctx, x, s):
        ctx.save_for_backward(s)
        return s.clone()

    @staticmethod
    def backward(ctx, g):
        s, = ctx.saved_tensors
        return g * s * (1.0 - s), None


def _sigmoid(x, saved):
    return torch.sigmoid(x) if saved is None else _SigmoidSaved.apply(x, saved)


def _cbam_graph(u, avg, w1, w2, wsp, idx_hw, idx_c, gc_fix=None, gs_fix=None, mx_fix=None):
    """u*gc*gs with the pooled maxima taken at GIVEN positions (idx_hw [N,C] flat pixel of the channel max-pool,
    idx_c [N,1,H,W] channel of the spatial max), so that the route is a saved decision, not re-derived"""
    mx = _fix(u.flatten(2).gather(2, idx_hw.unsqueeze(-1)).squeeze(-1), mx_fix)   # [N,C]
    hid = F.relu(F.linear(avg, w1.flatten(1))) + F.relu(F.linear(mx, w1.flatten(1)))
    gc = _sigmoid(F.linear(hid, w2.flatten(1)), gc_fix)             # W2 is linear: W2 h_a + W2 h_m
    u1 = u * gc[:, :, None, None]
    m = u1.mean(1, keepdim=True)
    mc = u1.gather(1, idx_c)
    gs = _sigmoid(F.conv2d(torch.cat([m, mc], 1), wsp, padding=1), gs_fix)
    return u1 * gs, gc, gs


class _NormBlock(torch.autograd.Function):
    """mode 0: act(u) (no CBAM); 1: act(u + cbam(u)); 2: act(res + cbam(u)); u = IN(y)*gamma + beta"""

    @staticmethod
    def forward(ctx, y, res, gamma, beta, w1, w2, wsp, mode, slope, tag):
        N, C, H, W = y.shape
        mean = y.mean((2, 3), keepdim=True)
        var = y.var((2, 3), unbiased=False, keepdim=True)
        rstd = torch.rsqrt(var + 1e-5)
        uhat = (y - mean) * rstd
        g, b = gamma.view(1, C, 1, 1), beta.view(1, C, 1, 1)
        u = uhat * g + b
        idx_hw = idx_c = gc = gs = mxv = None
        if mode == 0:
            pre = u
        else:
            idx_hw = u.flatten(2).argmax(2)                                          # first index on ties
            avg = beta.view(1, C).expand(N, C)
            mxv = u.flatten(2).gather(2, idx_hw.unsqueeze(-1)).squeeze(-1)
            hid = F.relu(F.linear(avg, w1.flatten(1))) + F.relu(F.linear(mxv, w1.flatten(1)))
            gc0 = torch.sigmoid(F.linear(hid, w2.flatten(1)))
            idx_c = (u * gc0[:, :, None, None]).argmax(1, keepdim=True)              # spatial max is over u*gc
            cb, gc, gs = _cbam_graph(u, avg, w1, w2, wsp, idx_hw, idx_c)
            pre = (u + cb) if mode == 1 else (res + cb)
        out = _r(F.leaky_relu(pre, slope) if slope != 0.0 else F.relu(pre))
        uhat_r = _r(uhat)
        teach = TEACH.get(tag) if TEACH is not None else None
        if TRACE is not None:
            TRACE.append((tag, out.detach()))
        if teach is not None:
            # the CUDA path's saved state for this site replaces ours (after measuring how far ours is from it)
            _local(tag, "out", out, teach["out"])
            _local(tag, "uhat", uhat_r, teach["uhat"])
            _local(tag, "rstd", rstd, teach["rstd"])
            if mode != 0:
                _local(tag, "gate_c", gc, teach["gc"])
                _local(tag, "gate_s", gs, teach["gs"])
                if LOCAL is not None:
                    LOCAL.append((tag, "argmax_hw", float((idx_hw != teach["idx_hw"]).float().mean()), 0.0))
                    LOCAL.append((tag, "argmax_c", float((idx_c != teach["idx_c"]).float().mean()), 0.0))
                _local(tag, "pooled_max", mxv, teach["mx"])
                idx_hw, idx_c, gc, gs, mxv = teach["idx_hw"], teach["idx_c"], teach["gc"], teach["gs"], teach["mx"]
            out, uhat_r, rstd = teach["out"].clone(), teach["uhat"], teach["rstd"]
        ctx.mode, ctx.slope, ctx.teach = mode, slope, teach is not None
        ctx.save_for_backward(uhat_r, out, rstd, idx_hw, idx_c, gamma, beta, w1, w2, wsp,
                              pre if _EXACT[0] else None, gc, gs, mxv)
        return out

    @staticmethod
    def backward(ctx, dout):
        uhat_r, out, rstd, idx_hw, idx_c, gamma, beta, w1, w2, wsp, pre_exact, gc_s, gs_s, mx_s = ctx.saved_tensors
        mode, slope = ctx.mode, ctx.slope
        N, C, H, W = uhat_r.shape
        sign_src = pre_exact if pre_exact is not None else out       # kernels: sign bit of the stored bf16 output
        amask = torch.where(sign_src > 0, torch.ones_like(out), torch.full_like(out, slope))
        dpre = dout * amask
        with torch.enable_grad():
            ur = uhat_r.detach().requires_grad_(True)
            leaves = [t.detach().requires_grad_(True) for t in (gamma, beta)]
            gm, bt = leaves
            u = ur * gm.view(1, C, 1, 1) + bt.view(1, C, 1, 1)
            if mode == 0:
                pre = u
                wl = []
            else:
                wl = [t.detach().requires_grad_(True) for t in (w1, w2, wsp)]
                # the kernels use the gates and the max-pooled value SAVED by the forward pass (computed from the fp32
                # raw output): the hidden units of the channel MLP keep the ReLU state they had in the forward pass
                fix = (gc_s, gs_s, mx_s) if not _EXACT[0] else (None, None, None)
                cb = _cbam_graph(u, bt.view(1, C).expand(N, C), wl[0], wl[1], wl[2], idx_hw, idx_c, *fix)[0]
                pre = (u + cb) if mode == 1 else cb
            grads = torch.autograd.grad(pre, [ur] + leaves + wl, dpre)
        du = grads[0]
        # InstanceNorm backward on the saved bf16 normalised activation: dy = rstd (du - mean(du) - uhat mean(du uhat))
        m1 = du.mean((2, 3), keepdim=True)
        m2 = (du * uhat_r).mean((2, 3), keepdim=True)
        dy = rstd * (du - m1 - uhat_r * m2)
        dres = dpre if mode == 2 else None
        dw = list(grads[3:]) if mode != 0 else [None, None, None]
        return dy, dres, grads[1], grads[2], dw[0], dw[1], dw[2], None, None, None


def norm_block(y, sd, bn: str, cbam: Optional[str], mode: int, slope: float = 0.0, res: Optional[Tensor] = None,
               tag: Optional[str] = None):
    w1 = w2 = wsp = None
    if cbam is not None:
        w1, w2 = sd[cbam + "channel_attention.conv1.weight"], sd[cbam + "channel_attention.conv2.weight"]
        wsp = sd[cbam + "spatial_attention.conv.weight"]
    out = _NormBlock.apply(gq(y), res, sd[bn + "weight"], sd[bn + "bias"], w1, w2, wsp, mode, slope, tag or bn + "nb")
    return gq(out)          # the gradient of a stored bf16 activation is written once, as bf16, by the consumer kernels


# ---------------------------------------------------------------------------------------------
# encoder (graph/encoder.py:26-40, graph/encodingBlock.py, graph/phrase_encoder.py:27-41)
# ---------------------------------------------------------------------------------------------

def _enc_stem(x, sd, p, first, second):
    spec = {"time": dict(stride=(2, 1), padding=(1, 0)), "pitch": dict(stride=(1, 2), padding=(0, 1))}
    t1 = _store(p + "t1", q(F.leaky_relu(F.conv2d(x, wq(sd[p + first + ".weight"]), **spec[first]), 0.01)))
    y = F.conv2d(t1, wq(sd[p + second + ".weight"]), **spec[second])
    return norm_block(y, sd, p + "bn.", p + "cbam.", 1, 0.01, tag=p + "nb")


def residual_module(x, sd, p):
    c1 = _store(p + "c1", q(F.relu(F.conv2d(x, wq(sd[p + "conv1.weight"]), padding=1))))
    y = F.conv2d(c1, wq(sd[p + "conv2.weight"]), padding=1)
    return norm_block(y, sd, p + "bn.", p + "cbam.", 2, 0.0, res=gq(x), tag=p + "nb")


def pooling_module(x, sd, p):
    y = F.conv2d(x, wq(sd[p + "conv.weight"]), stride=2, padding=1)
    return norm_block(y, sd, p + "bn.", p + "cbam.", 1, 0.0, tag=p + "nb")


def encoder_forward(x, sd, p="", phrase=False):
    time = _enc_stem(x, sd, p + "time_pitch.", "time", "pitch")
    pitch = _enc_stem(x, sd, p + "pitch_time.", "pitch", "time")
    o = torch.cat((pitch, time), 1)
    for i in range(4):
        o = residual_module(o, sd, p + "layers.%d." % (2 * i))
        o = pooling_module(o, sd, p + "layers.%d." % (2 * i + 1))
    pooled = _store(p + "pooled", q(o.mean((2, 3))))               # AvgPool2d over the whole last map, stored bf16
    z = gq(F.linear(pooled, wq(sd[p + "linear.weight"])))           # dz is rounded for the GEMMs, not for the bias sum
    b = sd.get(p + "linear.bias")
    return _store(p + "z", z if b is None else z + b)


# ---------------------------------------------------------------------------------------------
# decoder (graph/decoder.py:192-222)
# ---------------------------------------------------------------------------------------------

def _dec_stem(x, sd, p, first, second):
    spec = {"time": dict(stride=(6, 1)), "pitch": dict(stride=(1, 3))}
    t1 = _store(p + "t1", q(F.relu(F.conv_transpose2d(x, wq(sd[p + first + ".weight"]), **spec[first]))))
    y = F.conv_transpose2d(t1, wq(sd[p + second + ".weight"]), **spec[second])
    return norm_block(y, sd, p + "bn.", p + "cbam.", 1, 0.0, tag=p + "nb")


def _up_block(x, sd, p, padded: bool):
    if padded:          # DeConvPitchPadding: bn2 on both branches (decoder.py:137,142), CBAM on the first
        kw = dict(stride=2, padding=1, output_padding=(0, 1))
        y1 = F.conv_transpose2d(gq(x), wq(sd[p + "deConv1.weight"]), sd[p + "deConv1.bias"], **kw)
        o1 = norm_block(y1, sd, p + "bn2.", p + "cbam1.", 1, tag=p + "o1.nb")
        y2 = F.conv_transpose2d(x, wq(sd[p + "deConv2.weight"]), sd[p + "deConv2.bias"], **kw)
        o2 = norm_block(y2, sd, p + "bn2.", None, 0, tag=p + "o2.nb")
        last = p + "cbam2."
    else:               # DeConvModule (decoder.py:91-109)
        y1 = F.conv_transpose2d(gq(x), wq(sd[p + "deConv1.weight"]), None, stride=2, padding=1)
        o1 = norm_block(y1, sd, p + "bn1.", None, 0, tag=p + "o1.nb")
        y2 = F.conv_transpose2d(x, wq(sd[p + "deConv2.weight"]), sd[p + "deConv2.bias"], stride=2, padding=1,
                                output_padding=1)
        o2 = norm_block(y2, sd, p + "bn2.", None, 0, tag=p + "o2.nb")
        last = p + "cbam."
    cat = gq(torch.cat((o1, o2), 1))
    y3 = F.conv2d(cat, wq(sd[p + "conv.weight"]))
    return norm_block(y3, sd, p + "bn3.", last, 1, tag=p + "o3.nb")


def decoder_forward(z, pre_z, pf, position, sd, p="", drop_masks=None):
    emb = F.embedding(position, sd[p + "position_embedding.weight"])
    pcat = _store(p + "pcat", q(torch.cat((pf, emb), 1)))
    bcat = _store(p + "bcat", q(torch.cat((z, pre_z), 1)))
    lb = q(F.relu(F.linear(bcat, wq(sd[p + "bar_linear.weight"]), sd[p + "bar_linear.bias"])))
    lp = q(F.relu(F.linear(pcat, wq(sd[p + "phrase_linear.weight"]), sd[p + "phrase_linear.bias"])))
    lin = _store(p + "lin", torch.cat((lb, lp), 1))
    if drop_masks is not None:
        keep = torch.cat((drop_masks[1], drop_masks[0]), 1) * (1.0 / 0.7)
        lin = _store(p + "x", q(lin * keep))
    x = lin.view(-1, 2304, 1, 1)
    pitch = _dec_stem(gq(x), sd, p + "pitch.", "pitch", "time")
    time = _dec_stem(x, sd, p + "time.", "time", "pitch")
    hcat = gq(torch.cat((pitch, time), 1))
    y = F.conv2d(hcat, wq(sd[p + "fit1.weight"]))
    o = norm_block(y, sd, p + "bn.", p + "cbam.", 1, tag=p + "fit1.nb")
    o = _up_block(o, sd, p + "layers.0.", True)
    o = _up_block(o, sd, p + "layers.1.", True)
    o = _up_block(o, sd, p + "layers.2.", False)
    o = _up_block(o, sd, p + "layers.3.", False)
    return _store(p + "recon", torch.sigmoid(F.conv2d(o, sd[p + "fit2.weight"])))   # fit2 reads the fp32 master weight


def model_forward(note, pre_note, phrase, position, sd, drop_masks=None):
    """graph/model.py:22-33 (training composition), storage precision of the CUDA path.  encoder(note) and
    encoder(pre_note) share weights and have no batch-coupled op: one pass over both (as the CUDA path does)."""
    B = note.shape[0]
    pf = encoder_forward(phrase, sd, "phrase_encoder.phrase_encoder.", phrase=True)
    zz = encoder_forward(torch.cat((note, pre_note), 0), sd, "encoder.")
    z, pre_z = zz[:B], zz[B:]
    gen = decoder_forward(z, pre_z, pf, position, sd, "decoder.", drop_masks)
    return gen, z, pre_z, pf


def train_grads(sd, batch, drop_masks=None, is_pretraining=True, bce_only_rows=None):
    """forward + Loss + backward (agent/barGen.py:302-335 without the optimiser): (loss, gen, z, grads).
    ``bce_only_rows``: restrict the BCE mean to those samples (used by the 512-bar test)."""
    note, pre_note, phrase, position = batch
    leaves = OrderedDict((k, t.detach().clone().requires_grad_(True)) for k, t in sd.items())
    gen, z, pre_z, pf = model_forward(note, pre_note, phrase, position, leaves, drop_masks)
    if bce_only_rows is None:
        loss = O.loss_forward(gen, note, is_pretraining)
    else:
        loss = O.bce_mean(gen[bce_only_rows], note[bce_only_rows])
    loss.backward()
    return loss.detach(), gen.detach(), z.detach(), OrderedDict((k, t.grad) for k, t in leaves.items())
