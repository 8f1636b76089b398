#!/usr/bin/env python
"""Generate tests/golden/golden_disc_v1.pt by running the UNMODIFIED reference discriminators and
graph/model_with_gan.py on CPU.  TEST INFRASTRUCTURE; build container only (needs /root/reference):

    python oracle/gen_golden_disc.py

Injected without editing any reference source: ``Tensor.type('torch.cuda.FloatTensor')`` is shimmed to CPU while
model_with_gan.Model runs (graph/model_with_gan.py:29,36 hard-code CUDA); the model runs in eval mode (no dropout draw).
"""
import os
import sys
from collections import OrderedDict

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("BARVAE_REFERENCE", "/root/reference")
sys.path.insert(0, HERE)
sys.path.insert(0, REF)

import barvae_oracle as O  # noqa: E402
import disc_oracle as D  # noqa: E402
from gen_golden import cpu_cuda_shim  # noqa: E402

torch.set_num_threads(8)


def run_disc(module, sd, x):
    module.load_state_dict(sd)
    module.zero_grad()
    x = x.clone().requires_grad_(True)
    out = module(x)
    loss = torch.nn.functional.binary_cross_entropy(out, torch.ones_like(out))      # DLoss (bar_loss.py:36-42)
    loss.backward()
    grads = OrderedDict((k, p.grad.clone()) for k, p in module.named_parameters())
    return {"out": out.detach().clone(), "loss": loss.detach().clone(), "dx": x.grad.clone(),
            "grad_digest": O.grad_digest(grads), "grads_small": {k: v for k, v in grads.items() if v.numel() <= 1024}}


def main():
    from graph.z_discriminator import BarZDiscriminator, PhraseZDiscriminator
    from graph.bar_discriminator_with_feature import BarFeatureDiscriminator
    from graph.model_with_gan import Model

    G = OrderedDict()
    G["meta"] = {"torch": torch.__version__, "reference": "KMU-AELAB-MusicProject/MusicGeneration_VAE-torch"}
    g = torch.Generator().manual_seed(77)
    x = torch.randn(6, D.Z_DIM, generator=g)
    G["x"] = x
    for name, cls, spec in (("bar_z", BarZDiscriminator, D.z_disc_spec()), ("phrase_z", PhraseZDiscriminator, D.z_disc_spec()),
                            ("feature", BarFeatureDiscriminator, D.feature_disc_spec())):
        m = cls()
        assert list(m.state_dict().keys()) == list(spec.keys()), name
        assert all(tuple(v.shape) == tuple(spec[k]) for k, v in m.state_dict().items()), name
        for kind, seed in (("lively", 5), ("reference", 6)):
            G["%s/%s" % (name, kind)] = run_disc(m, D.make_disc_state_dict(spec, seed, kind), x)
    # adversarial-phase generator composition
    sd = O.make_state_dict(O.generator_spec(), 11, "lively")
    batch = O.make_inputs(2, 31)
    model = Model()
    model.load_state_dict(sd)
    model.eval()
    lat = torch.randn(2, 1152, generator=g)
    with torch.no_grad(), cpu_cuda_shim():
        gen, z, pre_z, pf, zf = model(*batch)
        gen2, zf2 = model(lat, batch[1], batch[2], batch[3], False)
    G["gan/train"] = {"gen": gen, "z": z, "pre_z": pre_z, "pf": pf, "z_fake": zf}
    G["gan/sample"] = {"latent": lat, "gen": gen2, "z_fake": zf2}
    out = os.path.join(os.path.dirname(HERE), "tests", "golden", "golden_disc_v1.pt")
    torch.save(G, out)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
