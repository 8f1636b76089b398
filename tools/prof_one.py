"""Run ONE contraction (fwd|dgrad|wgrad) of one layer shape a few times -- the target of an ncu capture.
usage: prof_one.py kind cin cout kh kw sy sx py px opy opx H W B mode"""
import importlib
import os
import sys

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "musicgeneration_vae-torch_b200"
eng = importlib.import_module(PKG + ".engine")
eb = importlib.import_module(PKG + ".graph.encodingBlock")
a = sys.argv[1:]
kind = a[0]
cin, cout, kh, kw, sy, sx, py, px, opy, opx, H, W, B = map(int, a[1:14])
mode = a[14]
m = (nn.Conv2d(cin, cout, (kh, kw), (sy, sx), (py, px), bias=False) if kind == "conv"
     else nn.ConvTranspose2d(cin, cout, (kh, kw), (sy, sx), (py, px), output_padding=(opy, opx), bias=False)).cuda()
g = eb.gemm_of(m)
OH, OW = g.out_hw(H, W)
x = eng.Act(torch.randn(B, H, W, cin, device="cuda").to(torch.bfloat16), B, H, W, cin)
y = eng.Act.empty(B, OH, OW, cout, dtype=torch.float32)
dy = eng.Act(torch.randn(B, OH, OW, cout, device="cuda").to(torch.bfloat16), B, OH, OW, cout)
dx = eng.Act.empty(B, H, W, cin)
m.weight.grad = torch.zeros_like(m.weight)
fn = {"fwd": lambda: g.forward(x, y), "dgrad": lambda: g.dgrad(dy, dx), "wgrad": lambda: g.wgrad(x, dy)}[mode]
for _ in range(4):
    fn()
torch.cuda.synchronize()
print("done")
