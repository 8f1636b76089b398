"""CPU oracle for the fully connected discriminators and the adversarial-phase generator composition  --  TEST
INFRASTRUCTURE, NOT PRODUCT CODE (only tests/ import it).

Plain-PyTorch fp32 functional restatements, each citing the reference lines it follows; weights are addressed through
state_dicts with the reference's key names.  Pinning: the reference ships no golden vectors, so
``oracle/gen_golden_disc.py`` runs the UNMODIFIED reference classes (importable on CPU in the build container) on seeded
inputs / weights and commits their outputs and gradients as ``tests/golden/golden_disc_v1.pt``;
``tests/test_oracle_golden.py::test_disc_*`` holds this file to them.
"""
from collections import OrderedDict

import torch
import torch.nn.functional as F

import barvae_oracle as O

Z_DIM = 1152


def z_disc_spec():
    """graph/z_discriminator.py:13-24 (PhraseZDiscriminator) == :38-49 (BarZDiscriminator)"""
    s, dims = OrderedDict(), [(Z_DIM, 512), (512, 512), (512, 512), (512, 512), (512, 1)]
    for i, (din, dout) in zip((0, 2, 4, 6, 8), dims):
        s["net.%d.weight" % i] = (dout, din)
        s["net.%d.bias" % i] = (dout,)
    return s


def feature_disc_spec():
    """graph/bar_discriminator_with_feature.py:10-11"""
    return OrderedDict([("linear1.weight", (512, Z_DIM)), ("linear2.weight", (1, 512))])


def make_disc_state_dict(spec, seed: int, kind: str = "lively"):
    """kind='reference': Linear weights ~ N(-1, 1) (graph/weights_initializer.py:19-23), biases PyTorch-default scale;
    kind='lively': fan-in scaled weights so that the sigmoid is not saturated and every ReLU is mixed."""
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    for k, shp in spec.items():
        fan_in = shp[1] if len(shp) == 2 else None
        if k.endswith(".bias"):
            sd[k] = (torch.rand(shp, generator=g) * 2 - 1) * (0.03 if kind == "reference" else 0.3)
        elif kind == "reference":
            sd[k] = torch.randn(shp, generator=g) - 1.0
        else:
            sd[k] = torch.randn(shp, generator=g) * (1.5 / fan_in ** 0.5)
    return sd


def z_disc_forward(x, sd):
    """graph/z_discriminator.py:28-29,53-54: Linear-ReLU x4, Linear, Sigmoid"""
    h = x
    for i in (0, 2, 4, 6):
        h = F.relu(F.linear(h, sd["net.%d.weight" % i], sd["net.%d.bias" % i]))
    return torch.sigmoid(F.linear(h, sd["net.8.weight"], sd["net.8.bias"]))


def feature_disc_forward(x, sd):
    """graph/bar_discriminator_with_feature.py:17-24"""
    x = x.view(-1, Z_DIM)
    return torch.sigmoid(F.linear(F.linear(x, sd["linear1.weight"]), sd["linear2.weight"]))


def fake_note(gen_note):
    """graph/model_with_gan.py:29,36: torch.gt(gen_note, 0.3) as float (non-differentiable)"""
    return (gen_note > 0.3).float()


def model_with_gan_forward(note, pre_note, phrase, position, sd, is_note=True, drop_masks=None):
    """graph/model_with_gan.py:20-38"""
    if is_note:
        gen, z, pre_z, pf = O.model_forward(note, pre_note, phrase, position, sd, True, drop_masks)
        return gen, z, pre_z, pf, O.encoder_forward(fake_note(gen), sd, "encoder.")
    gen = O.model_forward(note, pre_note, phrase, position, sd, False, drop_masks)
    return gen, O.encoder_forward(fake_note(gen), sd, "encoder.")


# ---------------------------------------------------------------------------------------------
# convolutional piano-roll discriminator (graph/bar_discriminator.py) and Refiner (graph/refiner.py)
# ---------------------------------------------------------------------------------------------
# Both contain nn.BatchNorm2d: batch statistics in training mode (the mode the GAN phase runs them in,
# agent/barGen_with_gan.py:464-466), running statistics in eval mode.  ``buffers`` (running_mean / running_var /
# num_batches_tracked under the reference's key names) are updated in place in training mode, exactly as the module does.

def _bn_spec(p, c):
    return [(p + "weight", (c,)), (p + "bias", (c,)), (p + "running_mean", (c,)), (p + "running_var", (c,)),
            (p + "num_batches_tracked", ())]


def bar_disc_spec():
    """graph/bar_discriminator.py:7-217, state_dict order (parameters and buffers interleaved per module)"""
    s = []
    p = "chord."                                                            # ChordFeature :11-24
    s += [(p + "chord_conv1.weight", (8, 1, 3, 1)), (p + "chord_conv2.weight", (16, 8, 3, 1)),
          (p + "chord_fit.weight", (16, 16, 1, 1)), (p + "chord_conv3.weight", (32, 16, 3, 3)),
          (p + "chord_conv4.weight", (64, 32, 3, 3))]
    for i, c in zip(range(1, 6), (8, 16, 16, 32, 64)):
        s += _bn_spec(p + "batch_norm%d." % i, c)
    p = "onoff."                                                            # OnOffFeature :65-79
    s += [(p + "onoff_conv1.weight", (8, 1, 3, 3)), (p + "onoff_conv2.weight", (8, 8, 3, 3))]
    s += _bn_spec(p + "batch_norm2.", 8)
    s += [(p + "onoff_conv3.weight", (16, 8, 3, 3)), (p + "onoff_conv4.weight", (32, 16, 3, 3)),
          (p + "onoff_fit.weight", (32, 32, 1, 1)), (p + "onoff_conv5.weight", (64, 32, 3, 3))]
    p = "basic."                                                            # BasicFeature :141-161
    s += [(p + "pitch1.weight", (8, 1, 1, 4)), (p + "pitch2.weight", (8, 8, 4, 1)), (p + "time1.weight", (8, 1, 4, 1)),
          (p + "time2.weight", (8, 8, 1, 4)), (p + "fit.weight", (8, 16, 1, 1))]
    s += _bn_spec(p + "bn.", 8)
    for i, (cin, cout, basic) in enumerate(((8, 16, False), (16, 32, False), (32, 64, True))):   # ConvModule :107-118
        q = p + "layers.%d." % i
        if not basic:
            s += [(q + "conv1.weight", (cin, cin, 3, 3))]
        s += [(q + "conv2.weight", (cout, cin, 3, 3))]
        s += _bn_spec(q + "bn1.", cin) + _bn_spec(q + "bn2.", cout)
    s += [("linear.weight", (1, 192))]
    return OrderedDict(s)


def refiner_spec():
    """graph/refiner.py:11-44 with layer2's Conv2d taking the 2 channels layer1 produces (:19 declares 1 and cannot run)"""
    s = [("layer1.0.weight", (2, 1, 4, 4)), ("layer1.0.bias", (2,))] + _bn_spec("layer1.1.", 2)
    s += [("layer2.0.weight", (8, 2, 4, 4)), ("layer2.0.bias", (8,))] + _bn_spec("layer2.1.", 8)
    s += [("layer3.0.weight", (1024, 2880)), ("layer3.0.bias", (1024,)), ("layer4.0.weight", (2880, 1024)),
          ("layer4.0.bias", (2880,))]
    s += [("layer5.0.weight", (8, 2, 4, 4))] + _bn_spec("layer5.1.", 2)
    s += [("layer6.0.weight", (2, 1, 4, 4))] + _bn_spec("layer6.1.", 1)
    return OrderedDict(s)


def make_conv_state_dict(spec, seed: int, kind: str = "lively"):
    """kind='reference': Conv2d / Linear weights ~ N(-1,1), BatchNorm weight ~ N(-1,1) too (graph/weights_initializer.py:12-17
    matches 'BatchNorm' and redraws the WEIGHT twice), ConvTranspose2d / biases at PyTorch-default scale, fresh running
    statistics; kind='lively': fan-in scaled weights, gamma ~ 1 +- 0.3, beta ~ +-0.3, non-trivial running statistics."""
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    for k, shp in spec.items():
        leaf = k.split(".")[-1]
        n = 1
        for d in shp:
            n *= d
        is_bn = any(k.endswith(s) for s in ("running_mean", "running_var", "num_batches_tracked")) or \
            (len(shp) == 1 and (".bn" in k or "batch_norm" in k or k.split(".")[-2] == "1"))
        is_convT = k.startswith("layer5.0") or k.startswith("layer6.0")
        if leaf == "num_batches_tracked":
            sd[k] = torch.tensor(0 if kind == "reference" else 3, dtype=torch.long)
        elif leaf == "running_mean":
            sd[k] = torch.zeros(shp) if kind == "reference" else torch.randn(shp, generator=g) * 0.2
        elif leaf == "running_var":
            sd[k] = torch.ones(shp) if kind == "reference" else torch.rand(shp, generator=g) + 0.5
        elif is_bn and leaf == "weight":
            sd[k] = (torch.randn(shp, generator=g) - 1.0) if kind == "reference" else 1.0 + 0.3 * torch.randn(shp, generator=g)
        elif is_bn and leaf == "bias":
            sd[k] = torch.zeros(shp) if kind == "reference" else 0.3 * torch.randn(shp, generator=g)
        elif leaf == "bias":
            sd[k] = (torch.rand(shp, generator=g) * 2 - 1) * 0.05
        else:
            fan_in = n // shp[0] if not is_convT else n // shp[1]
            if kind == "reference" and not is_convT:
                sd[k] = torch.randn(shp, generator=g) - 1.0
            else:
                sd[k] = torch.randn(shp, generator=g) * (1.4 / fan_in ** 0.5)
    return sd


def batch_norm(x, sd, p, training, momentum, eps=1e-5):
    """nn.BatchNorm2d.forward: F.batch_norm with the module's buffers (updated in place when training)"""
    if training:
        sd[p + "num_batches_tracked"] += 1
    return F.batch_norm(x, sd[p + "running_mean"], sd[p + "running_var"], sd[p + "weight"], sd[p + "bias"], training,
                        momentum, eps)


class _Ident:
    """precision policy of the fp32 oracle: nothing is rounded.  tests/test_gpu_convdisc.py passes barvae_emul's q / gq / wq
    (value+gradient, gradient-only, value-only bf16 rounding) to evaluate the SAME functions at the CUDA path's storage
    precision: bf16 GEMM operands and stored activations, fp32 raw conv outputs in front of a BatchNorm, bf16 gradients."""
    q = gq = wq = staticmethod(lambda t: t)


def bar_disc_forward(x, sd, training=True, prec=_Ident):
    """graph/bar_discriminator.py:199-217.  Note OnOffFeature.forward (:84-85): ``x[:, :-1]`` slices the CHANNEL axis of a
    one-channel tensor (empty) and the pad restores one channel of zeros, so ``x - onoff_x`` is x itself: the on/off
    feature is the per-step sum over pitches (replicated, not repaired)."""
    q, gq, wq = prec.q, prec.gq, prec.wq
    x = x.view(-1, 1, 192, 60)
    bn = lambda t, p, m: q(F.relu(batch_norm(gq(t), sd, p, training, m)))            # raw conv output fp32 -> BN+ReLU -> bf16
    cv = lambda t, k, **kw: F.conv2d(t, wq(sd[k]), **kw)
    rl = lambda t: q(F.relu(t))                                                      # ReLU fused in the conv epilogue -> bf16
    # ChordFeature :29-58
    c = q(x.view(-1, 1, 192, 12, 5).sum(4, keepdim=True).view(-1, 1, 192, 12))
    c = bn(cv(c, "chord.chord_conv1.weight", stride=(2, 1), padding=(1, 0)), "chord.batch_norm1.", 0.01)
    c = bn(cv(c, "chord.chord_conv2.weight", stride=(2, 1), padding=(1, 0)), "chord.batch_norm2.", 0.01)
    c = bn(cv(c, "chord.chord_fit.weight"), "chord.batch_norm3.", 0.01)
    c = bn(cv(c, "chord.chord_conv3.weight", stride=2, padding=1), "chord.batch_norm4.", 0.01)
    c = bn(cv(c, "chord.chord_conv4.weight", stride=2, padding=1), "chord.batch_norm5.", 0.01)
    c = F.avg_pool2d(c, (12, 3))
    # OnOffFeature :83-100
    shifted = F.pad(x[:, :-1], (0, 0, 0, 0, 1, 0))
    o = q(torch.sum(x - shifted, 3, keepdim=True))
    o = rl(cv(o, "onoff.onoff_conv1.weight", stride=(2, 1), padding=1))
    o = rl(cv(o, "onoff.onoff_conv2.weight", stride=(2, 1), padding=1))
    o = q(batch_norm(o, sd, "onoff.batch_norm2.", training, 0.1))
    o = rl(cv(o, "onoff.onoff_conv3.weight", stride=(2, 1), padding=1))
    o = rl(cv(o, "onoff.onoff_conv4.weight", stride=(2, 1), padding=1))
    o = rl(cv(o, "onoff.onoff_fit.weight"))
    o = rl(cv(o, "onoff.onoff_conv5.weight", stride=(2, 1), padding=1))
    o = F.avg_pool2d(o, (6, 1))
    # BasicFeature :163-183
    xq = q(x)
    pitch = rl(cv(xq, "basic.pitch1.weight", stride=(1, 2), padding=(0, 1)))
    pitch = rl(cv(pitch, "basic.pitch2.weight", stride=(2, 1), padding=(1, 0)))
    time = rl(cv(xq, "basic.time1.weight", stride=(2, 1), padding=(1, 0)))
    time = rl(cv(time, "basic.time2.weight", stride=(1, 2), padding=(0, 1)))
    b = torch.cat((pitch, time), 1)
    b = bn(cv(b, "basic.fit.weight"), "basic.bn.", 0.01)
    for i, basic in enumerate((False, False, True)):                        # ConvModule.forward :122-134
        p = "basic.layers.%d." % i
        if not basic:
            b = bn(cv(b, p + "conv1.weight", padding=1), p + "bn1.", 0.01)
        b = bn(cv(b, p + "conv2.weight", stride=2, padding=1), p + "bn2.", 0.01)
    b = F.avg_pool2d(b, (12, 4))
    out = torch.cat((c, o, b), 1).view(-1, 192)
    return torch.sigmoid(F.linear(out, sd["linear.weight"]))


def refiner_forward(x, sd, training=True, prec=_Ident):
    """graph/refiner.py:49-58 (with the layer2 fix of refiner_spec)"""
    q, gq, wq = prec.q, prec.gq, prec.wq
    bn = lambda t, p: batch_norm(gq(t), sd, p, training, 0.1)
    x2 = F.max_pool2d(q(F.leaky_relu(bn(F.conv2d(q(x), wq(sd["layer1.0.weight"]), sd["layer1.0.bias"], padding=2), "layer1.1."), 0.2)), 2)
    x8 = F.max_pool2d(q(F.leaky_relu(bn(F.conv2d(x2, wq(sd["layer2.0.weight"]), sd["layer2.0.bias"], padding=2), "layer2.1."), 0.2)), 2)
    f = q(F.relu(F.linear(x8.reshape(-1, 2880), wq(sd["layer3.0.weight"]), sd["layer3.0.bias"])))
    f = q(F.relu(F.linear(f, wq(sd["layer4.0.weight"]), sd["layer4.0.bias"])))
    x8t = x8 + f.view(-1, 8, 24, 15)
    x2t = x2 + q(F.relu(bn(F.conv_transpose2d(q(x8t), wq(sd["layer5.0.weight"]), stride=2, padding=1), "layer5.1.")))
    y = q(bn(F.conv_transpose2d(q(x2t), wq(sd["layer6.0.weight"]), stride=2, padding=1), "layer6.1."))
    return (x + torch.sigmoid(y)) * 0.5
