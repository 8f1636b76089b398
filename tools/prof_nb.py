"""Run one norm block (forward + backward) a few times -- target of an ncu capture / CUDA-event timing.
usage: prof_nb.py C H W B mode(plain|self|ext)"""
import importlib
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "musicgeneration_vae-torch_b200"
eng = importlib.import_module(PKG + ".engine")
C, H, W, B = map(int, sys.argv[1:5])
mode = sys.argv[5]
dev = "cuda"
gamma = torch.ones(C, device=dev, requires_grad=True)
beta = torch.zeros(C, device=dev, requires_grad=True)
cb = None
if mode != "plain":
    cb = tuple(t.requires_grad_(True) for t in (torch.randn(C // 16, C, 1, 1, device=dev) / math.sqrt(C),
                                                torch.randn(C, C // 16, 1, 1, device=dev), torch.randn(1, 2, 3, 3, device=dev)))
nb = eng.NormBlock(C, gamma, beta, cb, {"plain": 0, "self": 1, "ext": 2}[mode], 0.0)
y = eng.Act(torch.randn(B, H, W, C, device=dev), B, H, W, C)
res = eng.Act(torch.randn(B, H, W, C, device=dev).to(torch.bfloat16), B, H, W, C) if mode == "ext" else None
out = eng.Act.empty(B, H, W, C)
dout = eng.Act(torch.randn(B, H, W, C, device=dev).to(torch.bfloat16), B, H, W, C)
dy = eng.Act.empty(B, H, W, C)
dres = eng.Act.empty(B, H, W, C) if mode == "ext" else None
for it in range(3):
    eng.profile_begin()
    ctx = nb.forward(y, out, res)
    nb.backward(ctx, dout, dy, dres)
    p = eng.profile_end()
elems = B * H * W * C
print("fwd %.3f ms (%.1f B/elem-equivalent at 6.5TB/s), bwd %.3f ms (%.1f)" % (
    p["nb_forward"], p["nb_forward"] * 1e-3 * 6.5e12 / elems, p["nb_backward"], p["nb_backward"] * 1e-3 * 6.5e12 / elems))
