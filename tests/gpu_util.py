"""Helpers shared by the GPU parity tests."""
import importlib
import json
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = "musicgeneration_vae-torch_b200"
REPORT = os.path.join(ROOT, "gpurun_out", "parity_report.jsonl")


def pkg(sub=""):
    return importlib.import_module(PKG + (("." + sub) if sub else ""))


def report(**kw):
    """Append one line of measured parity numbers (kept as evidence under gpurun_out/)."""
    try:
        os.makedirs(os.path.dirname(REPORT), exist_ok=True)
        with open(REPORT, "a") as f:
            f.write(json.dumps(kw) + "\n")
    except OSError:
        pass


def rel_fro(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def nhwc(t):
    """NCHW fp32 -> contiguous NHWC"""
    return t.permute(0, 2, 3, 1).contiguous()


def nchw(t):
    return t.permute(0, 3, 1, 2).contiguous()
