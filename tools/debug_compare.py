"""Layer-by-layer forward / per-tensor gradient comparison of the CUDA path against the CPU oracle (debug aid)."""
import importlib
import os
import sys
from collections import OrderedDict

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import barvae_oracle as O  # noqa: E402

PKG = "musicgeneration_vae-torch_b200"
kind = sys.argv[1] if len(sys.argv) > 1 else "lively"
B = 2
sd = O.make_state_dict(O.generator_spec(), 11, kind)
batch = O.make_inputs(B, 21)
masks = O.draw_dropout_masks(B, 77)

# ---- oracle with recorded block outputs
rec_o = []
for name in ("enc_time_pitch", "enc_pitch_time", "residual_module", "pooling_module", "dec_time_pitch",
             "dec_pitch_time", "deconv_pitch_padding", "deconv_module"):
    fn = getattr(O, name)

    def wrap(fn=fn, name=name):
        def f(x, sd_, p):
            out = fn(x, sd_, p)
            rec_o.append((p, out.detach()))
            return out
        return f
    setattr(O, name, wrap())
leaves = OrderedDict((k, v.clone().requires_grad_(True)) for k, v in sd.items())
og, oz, opz, opf = O.model_forward(*batch, leaves, True, masks)
oloss = O.loss_forward(og, batch[0], True)
oloss.backward()

# ---- CUDA path with recorded block outputs
eb = importlib.import_module(PKG + ".graph.encodingBlock")
db = importlib.import_module(PKG + ".graph.decoder")
Model = importlib.import_module(PKG + ".graph.model").Model
Loss = importlib.import_module(PKG + ".graph.loss.bar_loss").Loss
rec_c = []
for cls in (eb._StemModule, eb.ResidualModule, eb.PoolingModule, db._HeadModule, db._UpBlock):
    orig = cls.fwd

    def fwd(self, x, out, orig=orig):
        ctx = orig(self, x, out)
        rec_c.append((type(self).__name__, out.dense().float().permute(0, 3, 1, 2).detach().cpu()))
        return ctx
    cls.fwd = fwd
model = Model()
model.load_state_dict(sd)
model = model.cuda().train()
cb = tuple(t.cuda() for t in batch)
gen, z, pz, pf = model(*cb, True, tuple(m.cuda() for m in masks))
loss = Loss()(gen, cb[0], True)
loss.backward()
torch.cuda.synchronize()


def rf(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


print("loss", float(loss), float(oloss))
print("gen maxabs %.4g meanabs %.4g | z %.4g pz %.4g pf %.4g" % (
    float((gen.cpu() - og).abs().max()), float((gen.cpu() - og).abs().mean()), rf(z.cpu(), oz), rf(pz.cpu(), opz),
    rf(pf.cpu(), opf)))
# oracle order: phrase(stem tp, stem pt, 8 blocks), enc(note), enc(pre), decoder; cuda order: phrase, enc(2B), decoder
o_phrase, o_note, o_pre, o_dec = rec_o[0:10], rec_o[10:20], rec_o[20:30], rec_o[30:]
c_phrase, c_enc, c_dec = rec_c[0:10], rec_c[10:20], rec_c[20:]
print("--- phrase encoder blocks")
# cuda stem order: pitch_time then time_pitch; oracle: time_pitch then pitch_time
perm = [1, 0] + list(range(2, 10))
for i in range(10):
    p, t = o_phrase[perm[i]]
    print("%-40s %.4g" % (p, rf(c_phrase[i][1], t)))
print("--- encoder blocks (note | pre_note)")
for i in range(10):
    p, t = o_note[perm[i]]
    p2, t2 = o_pre[perm[i]]
    print("%-40s %.4g %.4g" % (p, rf(c_enc[i][1][:B], t), rf(c_enc[i][1][B:], t2)))
print("--- decoder blocks")
# oracle: pitch(dec_pitch_time) then time ; cuda same
for i in range(len(o_dec)):
    p, t = o_dec[i]
    print("%-40s %.4g" % (p, rf(c_dec[i][1], t)))
print("--- gradients (rel-fro, cos, |g| share)")
tot = sum(float(v.grad.double().norm() ** 2) for v in leaves.values() if v.grad is not None) ** 0.5
rows = []
for k, p in model.named_parameters():
    g0 = leaves[k].grad
    if g0 is None:
        continue
    a, b = p.grad.detach().cpu().double().flatten(), g0.double().flatten()
    cos = float((a @ b) / (a.norm() * b.norm() + 1e-30))
    rows.append((float((a - b).norm() / (b.norm() + 1e-30)), cos, float(b.norm() / tot), float(a.norm() / (b.norm() + 1e-30)), k))
for r in rows:
    print("%8.4f cos %7.4f share %.2e ratio %.3f %s" % r)
