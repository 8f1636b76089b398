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


# ---- stored forward state of the CUDA path, keyed like oracle/barvae_emul.py's tags (teacher forcing) -------------------
def _sel(t, rows):
    return t if rows is None else t.index_select(0, rows.to(t.device))


def _act_nchw(act, rows=None):
    return _sel(act.dense(), rows).permute(0, 3, 1, 2).float().cpu().contiguous()


def _nb_state(nctx, rows=None):
    """one norm-block site: what bvae_nb_forward saved for bvae_nb_backward"""
    H, W = nctx["H"], nctx["W"]
    nc = _sel(nctx["nc"], rows).float().cpu()                       # [N,C,8]: mean, rstd, a, b, gate_c, ...
    N = nc.shape[0]
    st = {"out": _act_nchw(nctx["out"], rows),
          "uhat": _sel(nctx["uhat"], rows).permute(0, 3, 1, 2).float().cpu().contiguous(),
          "rstd": nc[:, :, 1].reshape(N, -1, 1, 1).contiguous()}
    if "gs" in nctx:
        st["gc"] = nc[:, :, 4].contiguous()
        st["mx"] = nc[:, :, 5].contiguous()                         # affine-transformed max over H*W (channel max-pool)
        st["gs"] = _sel(nctx["gs"], rows).float().cpu().view(N, 1, H, W)
        st["idx_hw"] = _sel(nctx["nc_idx"], rows).long().cpu()
        st["idx_c"] = _sel(nctx["cidx"], rows).long().cpu().view(N, 1, H, W)
    return st


def cuda_forward_state(model, rows=None):
    """after ONE training-mode forward with ``_bvae_keep_state`` set on model.encoder / .decoder / the phrase trunk.
    ``rows``: keep only these samples (the bar encoder runs note and pre_note as one 2B batch: rows and rows + B)."""
    import torch
    t = {}
    rows_enc = None
    if rows is not None:
        rows = torch.as_tensor(rows, dtype=torch.long)
        Bfull = model.decoder._bvae_state[0].shape[0]
        rows_enc = torch.cat((rows, rows + Bfull))
    for prefix, mod, r in (("encoder.", model.encoder, rows_enc),
                           ("phrase_encoder.phrase_encoder.", model.phrase_encoder.phrase_encoder, rows)):
        z, (c_pt, c_tp, ctxs, pa, _, cat) = mod._bvae_state
        for name, c in (("time_pitch.", c_tp), ("pitch_time.", c_pt)):
            t[prefix + name + "t1"] = _act_nchw(c[1], r)
            t[prefix + name + "nb"] = _nb_state(c[3], r)
        for i, c in enumerate(ctxs):
            p = prefix + "layers.%d." % i
            if i % 2 == 0:
                t[p + "c1"] = _act_nchw(c[1], r)
                t[p + "nb"] = _nb_state(c[3], r)
            else:
                t[p + "nb"] = _nb_state(c[2], r)
        t[prefix + "pooled"] = _sel(pa.t.view(pa.N, 1024), r).float().cpu()
        t[prefix + "z"] = _sel(z.detach(), r).float().cpu()
    recon, (position, pcat, bcat, lin, keep, x, hcat, c_p, c_t, nb, c_f, ctxs, h, _) = model.decoder._bvae_state
    B = recon.shape[0]
    p = "decoder."
    t[p + "pcat"] = _sel(pcat.t.view(B, 2304), rows).float().cpu()
    t[p + "bcat"] = _sel(bcat.t.view(B, 2304), rows).float().cpu()
    t[p + "lin"] = _sel(lin.t.view(B, 2304), rows).float().cpu()
    if keep is not None:
        t[p + "x"] = _sel(x.t.view(B, 2304), rows).float().cpu()
    for name, c in (("pitch.", c_p), ("time.", c_t)):
        t[p + name + "t1"] = _act_nchw(c[1], rows)
        t[p + name + "nb"] = _nb_state(c[3], rows)
    t[p + "fit1.nb"] = _nb_state(c_f, rows)
    for i, c in enumerate(ctxs):
        bctx = c[2]
        t[p + "layers.%d.o1.nb" % i] = _nb_state(bctx[0][2], rows)
        t[p + "layers.%d.o2.nb" % i] = _nb_state(bctx[1][2], rows)
        t[p + "layers.%d.o3.nb" % i] = _nb_state(c[5], rows)
    t[p + "recon"] = _sel(recon.detach(), rows).float().cpu()
    return t


def keep_forward_state(model, flag=True):
    for m in (model.encoder, model.decoder, model.phrase_encoder.phrase_encoder):
        m._bvae_keep_state = flag
        if not flag and hasattr(m, "_bvae_state"):
            del m._bvae_state
