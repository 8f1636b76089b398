"""CPU oracle for the bar-VAE hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain-PyTorch fp32 *functional* restatement of the reference algorithm for the
path BASELINE.json's north_star names.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may import this file; the
product package (``musicgeneration_vae-torch_b200``) never does and fails loudly
when its CUDA library is missing.

Pinning: the reference ships no tests / golden vectors, so
this restatement is pinned against *outputs of the reference itself*: the
reference modules are importable on CPU in the build container, and
``oracle/gen_golden.py`` runs them on seeded inputs/weights and commits the
results under ``tests/golden/``.  ``tests/test_oracle_golden.py`` (CPU, ``-m "not
gpu"``) checks every function below against those vectors.

Every function cites the reference file:line it follows (paths relative to the
reference repository root).  Weights are addressed through a ``state_dict`` with the
reference's own key names, so a reference checkpoint can be fed in unchanged.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, Iterable, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]

ENC_LAYERS = [64, 128, 256, 512, 1024]          # graph/model.py:15,17
DEC_LAYERS = [1024, 512, 256, 128, 64]          # graph/model.py:16
LATENT = 1152                                   # graph/encoder.py:22
N_POSITIONS = 332                               # graph/decoder.py:187
BAR_H, BAR_W = 96, 60                           # agent/barGen.py:347
PHRASE_H = 384                                  # maker_bar.py:32

# graph/loss/bar_loss.py:10-17 (the 60-entry pitch prior; scaled by 0.08 at :17)
PITCH_PRIOR = [
    0.0079033, 0.00712255, 0.01189558, 0.00953322, 0.01102056, 0.01156428, 0.01136433, 0.01637716, 0.01211462,
    0.01776168, 0.01644157, 0.0171948, 0.01922302, 0.01582762, 0.02385192, 0.02001634, 0.02312213, 0.02348127,
    0.02263083, 0.0268141, 0.02373071, 0.02942328, 0.0272045, 0.0304963, 0.03032582, 0.02782333, 0.03458292,
    0.03230801, 0.03388906, 0.03283811, 0.03093611, 0.03616363, 0.03006419, 0.03296618, 0.02867032, 0.02654072,
    0.02609579, 0.01954488, 0.02251165, 0.01813882, 0.01599178, 0.01313839, 0.01104167, 0.01169814, 0.00756204,
    0.00793332, 0.00601032, 0.00540243, 0.00512497, 0.00286655, 0.00308927, 0.00260029, 0.00184589, 0.00166959,
    0.00103728, 0.00112497, 0.00071164, 0.00052543, 0.00072274, 0.00038808]


# ---------------------------------------------------------------------------------------------
# state_dict specification (key -> shape), reference key names
# ---------------------------------------------------------------------------------------------

def _cbam_spec(p: str, c: int) -> List[Tuple[str, Tuple[int, ...]]]:
    # graph/cbam.py:14-15 (C -> C//16 -> C, 1x1, no bias), :36 (2->1 3x3, no bias)
    return [(p + "channel_attention.conv1.weight", (c // 16, c, 1, 1)),
            (p + "channel_attention.conv2.weight", (c, c // 16, 1, 1)),
            (p + "spatial_attention.conv.weight", (1, 2, 3, 3))]


def _in_spec(p: str, c: int):
    return [(p + "weight", (c,)), (p + "bias", (c,))]


def encoder_spec(prefix: str = "", linear_bias: bool = True) -> "OrderedDict[str, Tuple[int, ...]]":
    """graph/encoder.py:8-24 + graph/encodingBlock.py (module registration order)."""
    s: List[Tuple[str, Tuple[int, ...]]] = []
    p = prefix + "time_pitch."                       # encodingBlock.py:12-19
    s += [(p + "time.weight", (32, 1, 4, 1)), (p + "pitch.weight", (32, 32, 1, 4))]
    s += _in_spec(p + "bn.", 32) + _cbam_spec(p + "cbam.", 32)
    p = prefix + "pitch_time."                       # encodingBlock.py:43-50
    s += [(p + "pitch.weight", (32, 1, 1, 4)), (p + "time.weight", (32, 32, 4, 1))]
    s += _in_spec(p + "bn.", 32) + _cbam_spec(p + "cbam.", 32)
    for i in range(1, len(ENC_LAYERS)):              # encoder.py:14-18
        cin, cout = ENC_LAYERS[i - 1], ENC_LAYERS[i]
        p = prefix + "layers.%d." % (2 * (i - 1))    # ResidualModule encodingBlock.py:74-81
        s += [(p + "conv1.weight", (cin, cin, 3, 3)), (p + "conv2.weight", (cin, cin, 3, 3))]
        s += _in_spec(p + "bn.", cin) + _cbam_spec(p + "cbam.", cin)
        p = prefix + "layers.%d." % (2 * (i - 1) + 1)  # PoolingModule encodingBlock.py:107-112
        s += [(p + "conv.weight", (cout, cin, 3, 3))]
        s += _in_spec(p + "bn.", cout) + _cbam_spec(p + "cbam.", cout)
    s += [(prefix + "linear.weight", (LATENT, 1024))]
    if linear_bias:                                  # encoder.py:22 vs phrase_encoder.py:23
        s += [(prefix + "linear.bias", (LATENT,))]
    return OrderedDict(s)


def phrase_model_spec(prefix: str = "") -> "OrderedDict[str, Tuple[int, ...]]":
    """graph/phrase_encoder.py:44-49: PhraseModel wraps PhraseEncoder (doubled prefix)."""
    return encoder_spec(prefix + "phrase_encoder.", linear_bias=False)


def decoder_spec(prefix: str = "") -> "OrderedDict[str, Tuple[int, ...]]":
    """graph/decoder.py:157-190 (registration order of Decoder.__init__)."""
    s: List[Tuple[str, Tuple[int, ...]]] = []
    s += [(prefix + "bar_linear.weight", (LATENT, 2 * LATENT)), (prefix + "bar_linear.bias", (LATENT,)),
          (prefix + "phrase_linear.weight", (LATENT, 2 * LATENT)), (prefix + "phrase_linear.bias", (LATENT,))]
    p = prefix + "time."                             # decoder TimePitchModule :12-19
    s += [(p + "time.weight", (2304, 1024, 6, 1)), (p + "pitch.weight", (1024, 1024, 1, 3))]
    s += _in_spec(p + "bn.", 1024) + _cbam_spec(p + "cbam.", 1024)
    p = prefix + "pitch."                            # decoder PitchTimeModule :43-49
    s += [(p + "pitch.weight", (2304, 1024, 1, 3)), (p + "time.weight", (1024, 1024, 6, 1))]
    s += _in_spec(p + "bn.", 1024) + _cbam_spec(p + "cbam.", 1024)
    s += [(prefix + "fit1.weight", (1024, 2048, 1, 1))]
    s += _in_spec(prefix + "bn.", 1024)
    s += [(prefix + "fit2.weight", (1, 64, 1, 1))]
    for i in range(1, len(DEC_LAYERS)):
        cin, cout = DEC_LAYERS[i - 1], DEC_LAYERS[i]
        p = prefix + "layers.%d." % (i - 1)
        if i < 3:                                    # DeConvPitchPadding :116-130
            s += [(p + "deConv1.weight", (cin, cout, 4, 4)), (p + "deConv1.bias", (cout,)),
                  (p + "deConv2.weight", (cin, cout, 4, 4)), (p + "deConv2.bias", (cout,)),
                  (p + "conv.weight", (cout, cin, 1, 1))]
            s += _in_spec(p + "bn1.", cout) + _in_spec(p + "bn2.", cout) + _in_spec(p + "bn3.", cout)
            s += _cbam_spec(p + "cbam1.", cout) + _cbam_spec(p + "cbam2.", cout)
        else:                                        # DeConvModule :73-87
            s += [(p + "deConv1.weight", (cin, cout, 4, 4)),
                  (p + "deConv2.weight", (cin, cout, 3, 3)), (p + "deConv2.bias", (cout,)),
                  (p + "conv.weight", (cout, cin, 1, 1))]
            s += _in_spec(p + "bn1.", cout) + _in_spec(p + "bn2.", cout) + _in_spec(p + "bn3.", cout)
            s += _cbam_spec(p + "cbam.", cout)
    s += _cbam_spec(prefix + "cbam.", 1024)
    s += [(prefix + "position_embedding.weight", (N_POSITIONS, LATENT))]
    return OrderedDict(s)


def generator_spec() -> "OrderedDict[str, Tuple[int, ...]]":
    """graph/model.py:15-17 (encoder, decoder, phrase_encoder; refiner excluded: it cannot execute)."""
    s = OrderedDict()
    s.update(encoder_spec("encoder."))
    s.update(decoder_spec("decoder."))
    s.update(phrase_model_spec("phrase_encoder."))
    return s


def make_state_dict(spec: "OrderedDict[str, Tuple[int, ...]]", seed: int, kind: str = "lively") -> SD:
    """Deterministic weights from a CPU generator (identical on every box with this torch build).

    kind = "reference": the distribution graph/weights_initializer.py:5-23 produces -- Conv2d / Linear
            weights ~ N(-1, 1); ConvTranspose2d, biases: small uniform (PyTorch default scale);
            InstanceNorm gamma=1, beta=0; position embedding ~ U(-1, 1) (decoder.py:188).
    kind = "lively": every path numerically active: fan-in scaled normal weights, gamma ~ 1 +- 0.3,
            beta ~ +-0.3, CBAM MLP weights with mixed sign.
    The exact RNG stream of the reference constructors is deliberately NOT reproduced:
    parity always copies one state_dict into both sides.
    """
    g = torch.Generator().manual_seed(seed)
    sd: SD = OrderedDict()
    for k, shp in spec.items():
        n = 1
        for d in shp:
            n *= d
        leaf = k.split(".")[-2]
        is_bias = k.endswith(".bias")
        is_norm = leaf.startswith("bn")
        # ConvTranspose2d weights ([Cin, Cout, kh, kw]): decoder.layers.*.deConv*, and the decoder head
        # (2304->1024 and 1024->1024 with k == stride); no Conv2d of the model has those shapes.
        is_convT = leaf.startswith("deConv") or (len(shp) == 4 and (
            shp[0] == 2304 or tuple(shp) in ((1024, 1024, 1, 3), (1024, 1024, 6, 1))))
        if leaf == "position_embedding":
            t = torch.rand(shp, generator=g) * 2 - 1
        elif is_norm:
            if kind == "reference":
                t = torch.zeros(shp) if is_bias else torch.ones(shp)
            else:
                t = (torch.randn(shp, generator=g) * 0.3) + (0.0 if is_bias else 1.0)
        elif is_bias:
            t = (torch.rand(shp, generator=g) * 2 - 1) * 0.02
        else:
            fan_in = n // shp[0] if not is_convT else n // shp[1]
            if kind == "reference" and not is_convT:
                t = torch.randn(shp, generator=g) - 1.0
            elif kind == "reference":
                t = (torch.rand(shp, generator=g) * 2 - 1) / math.sqrt(max(fan_in, 1))
            else:
                t = torch.randn(shp, generator=g) * (1.4 / math.sqrt(max(fan_in, 1)))
        sd[k] = t.float().contiguous()
    return sd


def make_inputs(batch: int, seed: int, density: float = 0.05):
    """seeded synthetic binary piano-roll bars."""
    g = torch.Generator().manual_seed(seed)
    note = (torch.rand(batch, 1, BAR_H, BAR_W, generator=g) < density).float()
    pre_note = (torch.rand(batch, 1, BAR_H, BAR_W, generator=g) < density).float()
    phrase = (torch.rand(batch, 1, PHRASE_H, BAR_W, generator=g) < density).float()
    position = torch.randint(0, N_POSITIONS, (batch,), generator=g)
    return note, pre_note, phrase, position


# ---------------------------------------------------------------------------------------------
# building blocks
# ---------------------------------------------------------------------------------------------

def instance_norm(x: Tensor, sd: SD, p: str) -> Tensor:
    """nn.InstanceNorm2d(C, eps=1e-5, affine=True), no running stats: per-(n,c) biased variance in
    train AND eval (graph/encodingBlock.py:17, graph/decoder.py:17,81-83)."""
    return F.instance_norm(x, None, None, sd[p + "weight"], sd[p + "bias"], True, 0.01, 1e-5)


def channel_attention(x: Tensor, sd: SD, p: str) -> Tensor:
    """graph/cbam.py:22-29."""
    w1, w2 = sd[p + "conv1.weight"], sd[p + "conv2.weight"]
    avg = F.adaptive_avg_pool2d(x, 1)
    mx = F.adaptive_max_pool2d(x, 1)
    a = F.conv2d(F.relu(F.conv2d(avg, w1)), w2)
    m = F.conv2d(F.relu(F.conv2d(mx, w1)), w2)
    return x * torch.sigmoid(a + m)


def spatial_attention(x: Tensor, sd: SD, p: str) -> Tensor:
    """graph/cbam.py:43-52."""
    avg = torch.mean(x, dim=1, keepdim=True)
    mx, _ = torch.max(x, dim=1, keepdim=True)
    q = F.conv2d(torch.cat([avg, mx], dim=1), sd[p + "conv.weight"], padding=1)
    return x * torch.sigmoid(q)


def cbam(x: Tensor, sd: SD, p: str) -> Tensor:
    """graph/cbam.py:64-68."""
    return spatial_attention(channel_attention(x, sd, p + "channel_attention."), sd, p + "spatial_attention.")


def enc_time_pitch(x, sd, p):
    """graph/encodingBlock.py:25-36."""
    o = F.conv2d(x, sd[p + "time.weight"], stride=(2, 1), padding=(1, 0))
    o = F.leaky_relu(o, 0.01)
    o = F.conv2d(o, sd[p + "pitch.weight"], stride=(1, 2), padding=(0, 1))
    o = instance_norm(o, sd, p + "bn.")
    o = o + cbam(o, sd, p + "cbam.")
    return F.leaky_relu(o, 0.01)


def enc_pitch_time(x, sd, p):
    """graph/encodingBlock.py:56-67."""
    o = F.conv2d(x, sd[p + "pitch.weight"], stride=(1, 2), padding=(0, 1))
    o = F.leaky_relu(o, 0.01)
    o = F.conv2d(o, sd[p + "time.weight"], stride=(2, 1), padding=(1, 0))
    o = instance_norm(o, sd, p + "bn.")
    o = o + cbam(o, sd, p + "cbam.")
    return F.leaky_relu(o, 0.01)


def residual_module(x, sd, p):
    """graph/encodingBlock.py:87-100 (CBAM(out), not out+CBAM(out))."""
    o = F.relu(F.conv2d(x, sd[p + "conv1.weight"], padding=1))
    o = F.conv2d(o, sd[p + "conv2.weight"], padding=1)
    o = instance_norm(o, sd, p + "bn.")
    o = cbam(o, sd, p + "cbam.")
    return F.relu(x + o)


def pooling_module(x, sd, p):
    """graph/encodingBlock.py:118-126."""
    o = F.conv2d(x, sd[p + "conv.weight"], stride=2, padding=1)
    o = instance_norm(o, sd, p + "bn.")
    o = o + cbam(o, sd, p + "cbam.")
    return F.relu(o)


def encoder_forward(x: Tensor, sd: SD, p: str = "", phrase: bool = False) -> Tensor:
    """graph/encoder.py:26-40 (phrase=False) / graph/phrase_encoder.py:27-41 (phrase=True)."""
    time = enc_time_pitch(x, sd, p + "time_pitch.")
    pitch = enc_pitch_time(x, sd, p + "pitch_time.")
    o = torch.cat((pitch, time), dim=1)
    for i in range(4):
        o = residual_module(o, sd, p + "layers.%d." % (2 * i))
        o = pooling_module(o, sd, p + "layers.%d." % (2 * i + 1))
    o = F.avg_pool2d(o, kernel_size=(12, 2) if phrase else (3, 2))
    o = o.view(-1, 1024)
    return F.linear(o, sd[p + "linear.weight"], sd.get(p + "linear.bias"))


def phrase_model_forward(x: Tensor, sd: SD, p: str = "") -> Tensor:
    """graph/phrase_encoder.py:52-55."""
    return encoder_forward(x, sd, p + "phrase_encoder.", phrase=True)


def dec_time_pitch(x, sd, p):
    """graph/decoder.py:25-36."""
    o = F.relu(F.conv_transpose2d(x, sd[p + "time.weight"], stride=(6, 1)))
    o = F.conv_transpose2d(o, sd[p + "pitch.weight"], stride=(1, 3))
    o = instance_norm(o, sd, p + "bn.")
    o = o + cbam(o, sd, p + "cbam.")
    return F.relu(o)


def dec_pitch_time(x, sd, p):
    """graph/decoder.py:55-66."""
    o = F.relu(F.conv_transpose2d(x, sd[p + "pitch.weight"], stride=(1, 3)))
    o = F.conv_transpose2d(o, sd[p + "time.weight"], stride=(6, 1))
    o = instance_norm(o, sd, p + "bn.")
    o = o + cbam(o, sd, p + "cbam.")
    return F.relu(o)


def deconv_pitch_padding(x, sd, p):
    """graph/decoder.py:135-154: bn2 on BOTH branches, bn1 unused (defect replicated)."""
    o1 = F.conv_transpose2d(x, sd[p + "deConv1.weight"], sd[p + "deConv1.bias"], stride=2, padding=1,
                            output_padding=(0, 1))
    o1 = instance_norm(o1, sd, p + "bn2.")
    o1 = F.relu(o1 + cbam(o1, sd, p + "cbam1."))
    o2 = F.conv_transpose2d(x, sd[p + "deConv2.weight"], sd[p + "deConv2.bias"], stride=2, padding=1,
                            output_padding=(0, 1))
    o2 = F.relu(instance_norm(o2, sd, p + "bn2."))
    o = F.conv2d(torch.cat((o1, o2), dim=1), sd[p + "conv.weight"])
    o = instance_norm(o, sd, p + "bn3.")
    o = o + cbam(o, sd, p + "cbam2.")
    return F.relu(o)


def deconv_module(x, sd, p):
    """graph/decoder.py:91-109."""
    o1 = F.conv_transpose2d(x, sd[p + "deConv1.weight"], None, stride=2, padding=1)
    o1 = F.relu(instance_norm(o1, sd, p + "bn1."))
    o2 = F.conv_transpose2d(x, sd[p + "deConv2.weight"], sd[p + "deConv2.bias"], stride=2, padding=1,
                            output_padding=1)
    o2 = F.relu(instance_norm(o2, sd, p + "bn2."))
    o = F.conv2d(torch.cat((o1, o2), dim=1), sd[p + "conv.weight"])
    o = instance_norm(o, sd, p + "bn3.")
    o = o + cbam(o, sd, p + "cbam.")
    return F.relu(o)


def decoder_forward(z, pre_z, phrase_feature, position, sd: SD, p: str = "",
                    drop_masks: Optional[Tuple[Tensor, Tensor]] = None, return_logits: bool = False):
    """graph/decoder.py:192-222.

    drop_masks = (phrase_mask, bar_mask): 0/1 keep masks for Dropout(p=0.3) in train mode, drawn
    phrase-branch first (decoder.py:196 then :201); None = eval mode (identity).  Kept values are
    scaled by 1/(1-0.3) exactly as nn.Dropout does.
    """
    pf = torch.cat((phrase_feature, F.embedding(position, sd[p + "position_embedding.weight"])), dim=1)
    pf = F.relu(F.linear(pf, sd[p + "phrase_linear.weight"], sd[p + "phrase_linear.bias"]))
    if drop_masks is not None:
        pf = pf * drop_masks[0] / 0.7
    bf = torch.cat((z, pre_z), dim=1)
    bf = F.relu(F.linear(bf, sd[p + "bar_linear.weight"], sd[p + "bar_linear.bias"]))
    if drop_masks is not None:
        bf = bf * drop_masks[1] / 0.7
    x = torch.cat((bf, pf), dim=1).view(-1, 2304, 1, 1)
    pitch = dec_pitch_time(x, sd, p + "pitch.")
    time = dec_time_pitch(x, sd, p + "time.")
    o = torch.cat((pitch, time), dim=1)
    o = F.conv2d(o, sd[p + "fit1.weight"])
    o = instance_norm(o, sd, p + "bn.")
    o = F.relu(o + cbam(o, sd, p + "cbam."))
    o = deconv_pitch_padding(o, sd, p + "layers.0.")
    o = deconv_pitch_padding(o, sd, p + "layers.1.")
    o = deconv_module(o, sd, p + "layers.2.")
    o = deconv_module(o, sd, p + "layers.3.")
    logits = F.conv2d(o, sd[p + "fit2.weight"])
    if return_logits:
        return torch.sigmoid(logits), logits
    return torch.sigmoid(logits)


def draw_dropout_masks(batch: int, seed: int) -> Tuple[Tensor, Tensor]:
    """Keep-masks for the two Dropout(0.3) calls of Decoder.forward, phrase branch first."""
    g = torch.Generator().manual_seed(seed)
    m_phrase = (torch.rand(batch, LATENT, generator=g) >= 0.3).float()
    m_bar = (torch.rand(batch, LATENT, generator=g) >= 0.3).float()
    return m_phrase, m_bar


def model_forward(note, pre_note, phrase, position, sd: SD, is_train: bool = True,
                  drop_masks=None):
    """graph/model.py:22-41 with the (non-executable) refiner call removed -- i.e. exactly
    graph/model_with_gan.py:20-38 without the re-encode of the thresholded output."""
    pf = phrase_model_forward(phrase, sd, "phrase_encoder.")
    if is_train:
        z = encoder_forward(note, sd, "encoder.")
        pre_z = encoder_forward(pre_note, sd, "encoder.")
        gen = decoder_forward(z, pre_z, pf, position, sd, "decoder.", drop_masks)
        return gen, z, pre_z, pf
    pre_z = encoder_forward(pre_note, sd, "encoder.")
    return decoder_forward(note, pre_z, pf, position, sd, "decoder.", drop_masks)


# ---------------------------------------------------------------------------------------------
# losses
# ---------------------------------------------------------------------------------------------

def bce_mean(p: Tensor, t: Tensor) -> Tensor:
    """nn.BCELoss (mean): each log term clamped at -100 (graph/loss/bar_loss.py:19,25,29)."""
    lp = torch.clamp(torch.log(p), min=-100.0)
    l1p = torch.clamp(torch.log(1.0 - p), min=-100.0)
    return -(t * lp + (1.0 - t) * l1p).mean()


def loss_forward(probs: Tensor, labels: Tensor, is_pretraining: bool = False) -> Tensor:
    """graph/loss/bar_loss.py:23-33, device-agnostic."""
    if is_pretraining:
        recon = bce_mean(probs, labels)
    else:
        prior = torch.tensor(PITCH_PRIOR, dtype=torch.float32, device=probs.device) * 0.08
        default = torch.tensor([0.1 / 60], dtype=torch.float32, device=probs.device)
        recon = bce_mean(probs, labels * 0.82 + default + prior)
    out = torch.gt(probs, 0.3).float()
    additional = torch.gt(labels - out, 0.0001).float().sum() * 0.005
    return recon + additional


def reparameterize(mean: Tensor, logvar: Tensor, eps: Tensor) -> Tensor:
    """old/graphs/models/bar_v1/encoder.py:60-63 with eps supplied (randn_like in the reference)."""
    return mean + eps * torch.exp(0.5 * logvar)


def kl_sum(mean: Tensor, logvar: Tensor) -> Tensor:
    """old/graphs/losses/loss.py:16 : -0.5 * sum(1 + var - mean^2 - exp(var)) (sum over batch AND dims)."""
    return -0.5 * torch.sum(1 + logvar - mean.pow(2) - logvar.exp())


def kl_pair(mean, logvar, pre_mean, pre_logvar) -> Tensor:
    """old/graphs/losses/bar_loss.py:12,18: mean of the note / pre_note KL sums."""
    elbo = (torch.sum(1 + logvar - mean.pow(2) - logvar.exp()) +
            torch.sum(1 + pre_logvar - pre_mean.pow(2) - pre_logvar.exp())) / 2
    return -0.5 * elbo


# ---------------------------------------------------------------------------------------------
# optimiser / training step / sampling loop
# ---------------------------------------------------------------------------------------------

def adam_step(params: SD, grads: SD, m: SD, v: SD, step: int, lr: float = 0.002,
              b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8) -> None:
    """torch.optim.Adam defaults as used at agent/barGen.py:61-62 (no weight decay / amsgrad).
    In place; ``step`` is the 1-based count after this update.  Parameters with no gradient are skipped."""
    bc1 = 1.0 - b1 ** step
    bc2 = 1.0 - b2 ** step
    for k, g in grads.items():
        if g is None:
            continue
        m[k].mul_(b1).add_(g, alpha=1 - b1)
        v[k].mul_(b2).addcmul_(g, g, value=1 - b2)
        denom = (v[k].sqrt() / math.sqrt(bc2)).add_(eps)
        params[k].addcdiv_(m[k], denom, value=-(lr / bc1))


def train_step(sd: SD, batch, m: SD, v: SD, step: int, lr: float = 0.002, drop_masks=None,
               is_pretraining: bool = True):
    """agent/barGen.py:249-252,302-335 (pre-training branch): zero_grad, forward, Loss, backward, Adam.
    Returns (loss, grads) and updates sd/m/v in place."""
    note, pre_note, phrase, position = batch
    leaves = OrderedDict((k, t.detach().clone().requires_grad_(True)) for k, t in sd.items())
    gen, z, pre_z, pf = model_forward(note, pre_note, phrase, position, leaves, True, drop_masks)
    loss = loss_forward(gen, note, is_pretraining)
    loss.backward()
    grads = OrderedDict((k, t.grad) for k, t in leaves.items())
    adam_step(sd, grads, m, v, step, lr)
    return loss.detach(), grads, gen.detach()


def sample_song(sd: SD, latents: Tensor, music_length: int) -> Tensor:
    """maker_bar.py:32-44: sequential bar generation.  ``latents`` is [music_length*4, S, 1152]
    (the randn draws, one per bar, for S songs generated in lock-step; S=1 in the reference).
    Returns the binarised piano-roll [S, music_length*4*96, 60]."""
    S = latents.shape[1]
    pre_phrase = torch.zeros(S, 1, PHRASE_H, BAR_W)
    pre_bar = torch.zeros(S, 1, BAR_H, BAR_W)
    phrase_idx = [330] + [i for i in range(music_length - 2, -1, -1)]
    outputs = []
    k = 0
    for idx in range(music_length):
        bar_set = []
        pos = torch.full((S,), phrase_idx[idx], dtype=torch.long)
        for _ in range(4):
            gen = model_forward(latents[k], pre_bar, pre_phrase, pos, sd, is_train=False)
            k += 1
            pre_bar = torch.gt(gen, 0.3).float()
            bar_set.append(pre_bar.reshape(S, BAR_H, BAR_W))
        phrase = torch.cat(bar_set, dim=1)
        outputs.append(phrase)
        pre_phrase = phrase.reshape(S, 1, PHRASE_H, BAR_W)
    return torch.cat(outputs, dim=1)


def grad_digest(grads: SD, n_samples: int = 8) -> Dict[str, Tensor]:
    """Compact, order-independent fingerprint of a gradient/parameter dict: per tensor
    [sum, abs-sum, sum of squares, n_samples entries at fixed pseudo-random flat indices]."""
    out = {}
    for k, g in grads.items():
        if g is None:
            continue
        f = g.detach().double().reshape(-1)
        n = f.numel()
        idx = torch.tensor([(i * 2654435761 + 12345) % n for i in range(n_samples)], dtype=torch.long)
        out[k] = torch.cat([torch.stack([f.sum(), f.abs().sum(), (f * f).sum()]), f[idx]]).float()
    return out
