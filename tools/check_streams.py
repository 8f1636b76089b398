"""Stream-overlap race check at the benchmark batch: gradients of one training step with the branch / weight-gradient
streams ON must match the fully serial path to within the run-to-run floor of the serial path itself (fp32 atomics).
usage: python tools/check_streams.py [B]"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "musicgeneration_vae-torch_b200"
mdl = importlib.import_module(PKG + ".graph.model")
lossm = importlib.import_module(PKG + ".graph.loss.bar_loss")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dev = torch.device("cuda")
torch.manual_seed(3)
model = mdl.Model().to(dev).train()
model.decoder.dropout.p = 0.0
flat = model.flatten_parameters()
g = torch.Generator(device="cpu").manual_seed(5)
note = (torch.rand(B, 1, 96, 60, generator=g) < 0.05).float().to(dev)
pre = (torch.rand(B, 1, 96, 60, generator=g) < 0.05).float().to(dev)
phr = (torch.rand(B, 1, 384, 60, generator=g) < 0.05).float().to(dev)
pos = torch.randint(0, 4, (B,), generator=g).to(dev)
loss_fn = lossm.Loss()


def grads(streams):
    os.environ["BVAE_STREAMS"] = os.environ["BVAE_WGRAD_STREAM"] = "1" if streams else "0"
    flat.attach_grads(zero=True)
    gen = model(note, pre, phr, pos, True)[0]
    loss = loss_fn(gen, note, True)
    loss.backward()
    torch.cuda.synchronize()
    return flat.grad.clone(), float(loss)


def rel(a, b):
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


grads(True)                                    # warm-up (lazy operand packing, allocator pools)
s1, l1 = grads(False)
floor = worst = 0.0
for trial in range(4):
    s2, l2 = grads(False)
    floor = max(floor, rel(s2, s1))
for trial in range(4):
    o, lo = grads(True)
    worst = max(worst, rel(o, s1))
print("B=%d  serial-vs-serial (worst of 4) %.3e   overlap-vs-serial (worst of 4) %.3e   losses %.6f %.6f %.6f" % (B, floor, worst, l1, l2, lo))
ok = worst < 2 * floor + 1e-3 and abs(lo - l1) < 1e-3 * abs(l1)
print("STREAM CHECK " + ("OK" if ok else "FAILED"))
sys.exit(0 if ok else 1)
