"""Per-layer micro-benchmark of the contraction kernels (CUDA events, B bars): TFLOP/s of forward / dgrad / wgrad
for the model's layer shapes.  Usage: python tools/bench_layers.py [B] [filter-substring]"""
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
BF16 = torch.bfloat16

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
flt = sys.argv[2] if len(sys.argv) > 2 else ""
# name, kind, cin, cout, k, s, p, op, H, W, batch multiplier
LAYERS = [
    ("enc.stem2 (1,4)", "conv", 32, 32, (1, 4), (1, 2), (0, 1), (0, 0), 48, 60, 2),
    ("phr.stem2 (4,1)", "conv", 32, 32, (4, 1), (2, 1), (1, 0), (0, 0), 384, 30, 1),
    ("enc.l0 64 3x3", "conv", 64, 64, (3, 3), (1, 1), (1, 1), (0, 0), 48, 30, 2),
    ("phr.l0 64 3x3", "conv", 64, 64, (3, 3), (1, 1), (1, 1), (0, 0), 192, 30, 1),
    ("phr.l1 64->128 s2", "conv", 64, 128, (3, 3), (2, 2), (1, 1), (0, 0), 192, 30, 1),
    ("phr.l2 128 3x3", "conv", 128, 128, (3, 3), (1, 1), (1, 1), (0, 0), 96, 15, 1),
    ("phr.l4 256 3x3", "conv", 256, 256, (3, 3), (1, 1), (1, 1), (0, 0), 48, 8, 1),
    ("phr.l6 512 3x3", "conv", 512, 512, (3, 3), (1, 1), (1, 1), (0, 0), 24, 4, 1),
    ("phr.l7 512->1024 s2", "conv", 512, 1024, (3, 3), (2, 2), (1, 1), (0, 0), 24, 4, 1),
    ("dec.time.time", "convT", 2304, 1024, (6, 1), (6, 1), (0, 0), (0, 0), 1, 1, 1),
    ("dec.fit1", "conv", 2048, 1024, (1, 1), (1, 1), (0, 0), (0, 0), 6, 3, 1),
    ("dec.l0.deConv1", "convT", 1024, 512, (4, 4), (2, 2), (1, 1), (0, 1), 6, 3, 1),
    ("dec.l1.deConv1", "convT", 512, 256, (4, 4), (2, 2), (1, 1), (0, 1), 12, 7, 1),
    ("dec.l2.deConv1", "convT", 256, 128, (4, 4), (2, 2), (1, 1), (0, 0), 24, 15, 1),
    ("dec.l3.deConv1", "convT", 128, 64, (4, 4), (2, 2), (1, 1), (0, 0), 48, 30, 1),
    ("dec.l3.deConv2", "convT", 128, 64, (3, 3), (2, 2), (1, 1), (1, 1), 48, 30, 1),
    ("dec.l3.conv 1x1", "conv", 128, 64, (1, 1), (1, 1), (0, 0), (0, 0), 96, 60, 1),
    ("dec.l2.conv 1x1", "conv", 256, 128, (1, 1), (1, 1), (0, 0), (0, 0), 48, 30, 1),
]


def timeit(fn, iters=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


print("%-22s %9s | %8s %7s | %8s %7s | %8s %7s" % ("layer", "GFLOP", "fwd ms", "TF/s", "dgrad ms", "TF/s", "wgrad ms", "TF/s"))
tot = [0.0, 0.0, 0.0, 0.0]
for name, kind, cin, cout, k, s, p, op, H, W, mult in LAYERS:
    if flt and flt not in name:
        continue
    n = B * mult
    m = (nn.Conv2d(cin, cout, k, s, p, bias=False) if kind == "conv"
         else nn.ConvTranspose2d(cin, cout, k, s, p, output_padding=op, bias=False)).cuda()
    g = eb.gemm_of(m)
    OH, OW = g.out_hw(H, W)
    x = eng.Act(torch.randn(n, H, W, cin, device="cuda").to(BF16), n, H, W, cin)
    y = eng.Act.empty(n, OH, OW, cout, dtype=torch.float32)
    dy = eng.Act(torch.randn(n, OH, OW, cout, device="cuda").to(BF16), n, OH, OW, cout)
    dx = eng.Act.empty(n, H, W, cin)
    m.weight.grad = torch.zeros_like(m.weight)
    if kind == "conv":
        gf = 2.0 * n * OH * OW * cout * cin * k[0] * k[1] / 1e9
    else:
        gf = 2.0 * n * H * W * cout * cin * k[0] * k[1] / 1e9
    tf = timeit(lambda: g.forward(x, y))
    td = timeit(lambda: g.dgrad(dy, dx))
    tw = timeit(lambda: g.wgrad(x, dy))
    print("%-22s %9.1f | %8.3f %7.0f | %8.3f %7.0f | %8.3f %7.0f" % (name, gf, tf, gf / tf, td, gf / td, tw, gf / tw))
    tot[0] += gf; tot[1] += tf; tot[2] += td; tot[3] += tw
    del x, y, dy, dx, m, g
print("%-22s %9.1f | %8.3f %7.0f | %8.3f %7.0f | %8.3f %7.0f" % ("TOTAL", tot[0], tot[1], tot[0] / tot[1], tot[2], tot[0] / tot[2],
                                                             tot[3], tot[0] / tot[3]))
