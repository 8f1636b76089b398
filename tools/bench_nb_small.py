"""Norm sites with <= 1440 pixels: forward / backward time of one site with the per-sample nbs_*
kernels (BVAE_NB_SMALL=1 with BVAE_NB_SMALL_HW=1440: every site here takes them) against the round-2 kernels
(BVAE_NB_SMALL=0: nb_cl_* up to 128 pixels, the tiled nb_* / nbf_* sweeps above), CUDA events, one JSON line per (site, mode).
usage: bench_nb_small.py [B [CxHxW ...]]   (B = bars per step)"""
import importlib
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "musicgeneration_vae-torch_b200"
eng = importlib.import_module(PKG + ".engine")
lib = importlib.import_module(PKG + "._lib")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dev = "cuda"
# (C, H, W, samples): the norm sites of the model with <= 1440 pixels (tools/prof_sites.py lists them)
SITES = [(256, 12, 8, B), (512, 6, 4, B), (1024, 3, 2, B), (512, 24, 4, B), (1024, 12, 2, B), (512, 12, 7, B), (1024, 6, 3, B),
         (128, 24, 15, B), (256, 24, 15, B), (256, 48, 8, B), (32, 48, 30, B), (64, 48, 30, B), (128, 48, 30, B), (128, 96, 15, B)]
if len(sys.argv) > 2:
    SITES = [tuple(int(v) for v in a.split("x")) + (B,) for a in sys.argv[2:]]


def run(C, H, W, N, mode, reps=20):
    torch.manual_seed(0)
    gamma = torch.ones(C, device=dev, requires_grad=True)
    beta = torch.zeros(C, device=dev, requires_grad=True)
    cb = None
    if mode != "plain":
        cb = tuple(t.requires_grad_(True) for t in (torch.randn(C // 16, C, 1, 1, device=dev) / math.sqrt(C),
                                                    torch.randn(C, C // 16, 1, 1, device=dev), torch.randn(1, 2, 3, 3, device=dev)))
    nb = eng.NormBlock(C, gamma, beta, cb, {"plain": 0, "self": 1, "ext": 2}[mode], 0.0)
    y = eng.Act(torch.randn(N, H, W, C, device=dev), N, H, W, C)
    res = eng.Act(torch.randn(N, H, W, C, device=dev).to(torch.bfloat16), N, H, W, C) if mode == "ext" else None
    out = eng.Act.empty(N, H, W, C)
    dout = eng.Act(torch.randn(N, H, W, C, device=dev).to(torch.bfloat16), N, H, W, C)
    dy = eng.Act.empty(N, H, W, C)
    dres = eng.Act.empty(N, H, W, C) if mode == "ext" else None
    flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)      # 256 MB > L2
    tf = tb = 0.0
    for it in range(reps + 3):
        flush.zero_()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        ctx = nb.forward(y, out, res)
        e[1].record()
        nb.backward(ctx, dout, dy, dres)
        e[2].record()
        torch.cuda.synchronize()
        if it >= 3:
            tf += e[0].elapsed_time(e[1])
            tb += e[1].elapsed_time(e[2])
    return tf / reps * 1e3, tb / reps * 1e3, out.t.float().clone(), dy.t.float().clone()


for C, H, W, N in SITES:
    for mode in ("plain", "ext"):
        with lib.option("BVAE_NB_SMALL", 0):
            f0, b0, o0, d0 = run(C, H, W, N, mode)
        with lib.option("BVAE_NB_SMALL_HW", 1440):
            f1, b1, o1, d1 = run(C, H, W, N, mode)
        elems = N * H * W * C
        print(json.dumps({"site": "C%d %dx%d N%d %s" % (C, H, W, N, mode),
                          "fwd_us": {"nb_cl": round(f0, 1), "nbs": round(f1, 1)},
                          "bwd_us": {"nb_cl": round(b0, 1), "nbs": round(b1, 1)},
                          "fwd_floor_us_16B_per_elem": round(elems * 16 / 6547.8e9 * 1e6, 1),
                          "out_maxdiff": float((o0 - o1).abs().max()), "dy_maxdiff": float((d0 - d1).abs().max()),
                          "dy_scale": float(d0.abs().mean())}), flush=True)
