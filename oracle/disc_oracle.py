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
