"""Layer-by-layer comparison of the CUDA encoder / decoder forward with oracle/barvae_emul.py (storage-precision emulation):
prints, per stored activation, the fraction of elements that differ, the largest difference in bf16 ulps and rel-Frobenius.
Debug tool (GPU box):  python tools/debug_emul.py"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import barvae_emul as E  # noqa: E402
import barvae_oracle as O  # noqa: E402

PKG = "musicgeneration_vae-torch_b200"
Model = importlib.import_module(PKG + ".graph.model").Model


def nchw(act):
    return act.dense().permute(0, 3, 1, 2).float().cpu()


def cmp(tag, mine, want):
    d = (mine - want).abs()
    nz = float((d > 0).float().mean())
    rel = float(d.norm() / (want.norm() + 1e-30))
    ulp = want.abs().clamp_min(1e-30) * 2.0 ** -8
    print("%-46s differ %.5f  max %.2f ulp  rel-fro %.2e  shape %s" % (tag, nz, float((d / ulp).max()), rel, tuple(want.shape)))


def main():
    sd = O.make_state_dict(O.generator_spec(), 11, "lively")
    batch = O.make_inputs(2, 21)
    model = Model()
    model.load_state_dict(sd)
    model = model.cuda().train()
    x = batch[0]
    E.TRACE = []
    with torch.no_grad():
        ez = E.encoder_forward(x, sd, "encoder.")
    tr = dict(E.TRACE)
    E.TRACE = None
    enc = model.encoder
    with torch.no_grad():
        z, saved = enc._fwd(x.cuda(), True)
    torch.cuda.synchronize()
    c_pt, c_tp, ctxs, pa, _, cat = saved
    for name, c in (("encoder.time_pitch.", c_tp), ("encoder.pitch_time.", c_pt)):
        cmp(name + "t1", nchw(c[1]), tr[name + "t1"])
        cmp(name + "bn.out", nchw(c[3]["out"]), tr[name + "bn.out"])
    for i, c in enumerate(ctxs):
        p = "encoder.layers.%d." % i
        if i % 2 == 0:
            cmp(p + "c1", nchw(c[1]), tr[p + "c1"])
            cmp(p + "bn.out", nchw(c[3]["out"]), tr[p + "bn.out"])
        else:
            cmp(p + "bn.out", nchw(c[2]["out"]), tr[p + "bn.out"])
    cmp("z", z.float().cpu(), ez)


if __name__ == "__main__":
    main()
