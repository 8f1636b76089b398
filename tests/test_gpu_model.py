"""GPU (B200) model-level parity: the CUDA path behind the reference's nn.Module interface vs (a) the CPU oracle on
the same seeded inputs and weights and (b) the committed golden vectors produced by the reference itself.

Arithmetic of the CUDA path: bf16 GEMM operands and stored activations, fp32 accumulation, fp32 raw conv outputs in
front of every InstanceNorm, fp32 statistics / losses / parameters.  Stated tolerances:
  forward  : recon |err| max 6e-2, mean 1e-2 ; z / pre_z / phrase_feature rel-Frobenius 2e-2 ; loss rel 1.5e-2
  backward : this network's backward pass amplifies perturbations by ~1e4 (fp32 vs fp64 on CPU already differ by
             7e-4 per tensor; rounding ONLY the weights to bf16 in the fp32 oracle moves the gradients by a median
             30 % per tensor -- max-pool/argmax routing and ReLU masks switch).  The test therefore measures that
             floor (oracle with bf16-rounded weights vs oracle) and requires, per tensor carrying >= 1 % of the
             gradient norm (the contraction weights), rel-Frobenius error <= 0.15 + 3 x floor; global gradient
             cosine >= 0.95; every tensor's norm within [0.02, 50] x the oracle's.  A wiring bug shows up as >= 100 %.
The per-kernel tests (tests/test_gpu_kernels.py) hold every block to 2e-3 where no such amplification exists.
With the reference initialisation N(-1,1) (graph/weights_initializer.py) the fp32 CPU oracle itself is not
reproducible across thread counts in backward (tests/test_oracle_golden.py), so only forward quantities are held."""
import os
from collections import OrderedDict

import pytest
import torch

from gpu_util import pkg, rel_fro, report

pytestmark = pytest.mark.gpu


def _load(model, sd):
    model.load_state_dict(sd)
    return model.cuda()


def _grads_vs(model, ref_grads):
    got = OrderedDict((k, p.grad) for k, p in model.named_parameters())
    errs, dots, na, nb = {}, 0.0, 0.0, 0.0
    for k, rg in ref_grads.items():
        if rg is None:
            assert got[k] is None or float(got[k].abs().max()) == 0.0, k      # unused bn1 gamma/beta
            continue
        a, b = got[k].detach().double().cpu(), rg.double()
        dots += float((a * b).sum())
        na += float((a * a).sum())
        nb += float((b * b).sum())
        errs[k] = (float((a - b).norm()), float(b.norm()))
    cos = dots / ((na ** 0.5) * (nb ** 0.5) + 1e-30)
    return cos, errs, nb ** 0.5


@pytest.mark.parametrize("kind", ["lively", "reference"])
def test_model_train_step_vs_oracle(golden, oracle, kind):
    O, c = oracle, golden[kind]
    Model = pkg("graph.model").Model
    Loss = pkg("graph.loss.bar_loss").Loss
    sd = O.make_state_dict(O.generator_spec(), c["seed_w"], kind)
    batch = O.make_inputs(c["B"], c["seed_x"])
    masks = O.draw_dropout_masks(c["B"], 77)
    model = _load(Model(), sd)
    model.train()
    note, pre_note, phrase, position = (t.cuda() for t in batch)
    gen, z, pre_z, pf = model(note, pre_note, phrase, position, True, tuple(m.cuda() for m in masks))
    loss = Loss()(gen, note, True)
    loss.backward()
    want = c["train_pre"]
    m = dict(test="model_train", kind=kind,
             gen_maxabs=float((gen.cpu() - want["gen"]).abs().max()), z=rel_fro(z, want["z"]),
             pre_z=rel_fro(pre_z, want["pre_z"]), pf=rel_fro(pf, want["pf"]),
             loss=float(loss), loss_want=float(want["loss"]))
    # full gradients from the oracle (CPU, fp32) on the same inputs
    leaves = OrderedDict((k, v.clone().requires_grad_(True)) for k, v in sd.items())
    og, oz, _, _ = O.model_forward(*batch, leaves, True, masks)
    O.loss_forward(og, batch[0], True).backward()
    ograds = OrderedDict((k, v.grad) for k, v in leaves.items())
    cos, errs, gnorm = _grads_vs(model, ograds)
    m.update(grad_cos=cos, gen_meanabs=float((gen.detach().cpu() - want["gen"]).abs().mean()))
    if kind == "lively":
        # perturbation floor: the oracle's own gradients when only the weights are rounded to bf16
        l2 = OrderedDict((k, v.to(torch.bfloat16).float().requires_grad_(True)) for k, v in sd.items())
        g2 = O.model_forward(*batch, l2, True, masks)[0]
        O.loss_forward(g2, batch[0], True).backward()
        bad, worst = [], 0.0
        for k, (e, n) in errs.items():
            share = n / gnorm
            mine = float(model.get_parameter(k).grad.double().norm())
            if n > 1e-6 * gnorm and not (0.02 < mine / n < 50):      # gross sanity on EVERY tensor that carries signal
                bad.append((k, "norm ratio", round(mine / n, 4)))
            if share < 1e-2:
                # small tensors (CBAM MLPs, norm affine) are single ReLU-boundary / argmax decisions away from O(1)
                # changes at batch 2 (e.g. encoder.pitch_time.cbam: a hidden pre-activation of +0.0097 in the oracle)
                continue
            floor = float((l2[k].grad - ograds[k]).norm()) / (n + 1e-30)
            rel = e / (n + 1e-30)
            worst = max(worst, rel / (0.15 + 3 * floor))
            if rel > 0.15 + 3 * floor:
                bad.append((k, round(rel, 3), round(floor, 3)))
        m.update(worst_ratio=worst, bad=bad[:8])
    report(**m)
    if kind == "lively":
        assert m["gen_maxabs"] < 6e-2 and m["gen_meanabs"] < 1e-2, m
        assert m["z"] < 2e-2 and m["pre_z"] < 2e-2 and m["pf"] < 2e-2, m
        assert abs(m["loss"] - m["loss_want"]) < 1.5e-2 * abs(m["loss_want"]), m
        assert not m["bad"], m
        assert cos > 0.95, m
    else:
        assert m["gen_meanabs"] < 5e-3, m
        assert m["z"] < 1e-2 and m["pre_z"] < 1e-2 and m["pf"] < 1e-2, m
        assert abs(m["loss"] - m["loss_want"]) < 1e-2 * abs(m["loss_want"]), m


def _emul_compare(model, egrads, tol, label, tol_cbam=None):
    """every gradient tensor of the model vs the storage-precision emulation: rel-Frobenius per tensor, no skipping.
    Biases in front of an InstanceNorm have an analytically zero gradient (our kernels write exactly 0, autograd
    yields rounding noise): held to an absolute bound instead.  ``tol_cbam`` applies to the CBAM attention weights
    (18-element spatial kernels, C/16-hidden-unit channel MLPs: sums of a few terms per sample that partly cancel).
    The bounds are set from the SPREAD of this comparison over repeated runs, not from one run: round 3 sampled it 26 times
    on the B200 (identical inputs; what moves between runs is the order of the fp32 atomics, amplified through bf16
    re-rounding -- two runs of the CUDA path differ from EACH OTHER by a median 33 % per tensor because ReLU masks and
    arg-max routes flip, which teacher forcing removes from this comparison but not from the stored state it starts from).
    Worst non-attention tensor per run: 1.9e-2 .. 4.1e-2 (the first encoder block's InstanceNorm gamma, the end of a
    ~100-stage backward chain); worst attention tensor per run: 3e-2 .. 1e-1, once 2.1e-1 (the channel-MLP weight of that
    block).  The round-2 bounds (3e-2 / 1.2e-1) sat inside that spread and failed one run in four."""
    gnorm = sum(float(g.double().norm()) ** 2 for g in egrads.values() if g is not None) ** 0.5
    rows, bad = [], []
    for k, p in model.named_parameters():
        eg = egrads[k]
        if eg is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k             # unused bn1 gamma / beta
            continue
        mine = p.grad.detach().double().cpu()
        n = float(eg.double().norm())
        if ".deConv" in k and k.endswith(".bias"):
            # (the emulation's value is what rounding the saved normalised activation to bf16 leaves of an exact zero)
            assert float(mine.abs().max()) <= 1e-6 * gnorm and n <= 1e-4 * gnorm, (k, float(mine.abs().max()), n)
            continue
        rel = float((mine - eg.double()).norm()) / (n + 1e-30)
        rows.append((rel, k))
        lim = tol_cbam if (tol_cbam is not None and "attention" in k) else tol
        if rel > lim:
            bad.append((k, round(rel, 4)))
    report(test=label + "_per_tensor", rel=[(k, round(r, 5)) for r, k in rows])
    rows.sort(reverse=True)
    rels = [r for r, _ in rows]
    report(test=label, n_tensors=len(rows), worst=rows[:6], median=rels[len(rels) // 2],
           worst_non_cbam=max(r for r, k in rows if "attention" not in k),
           n_over_1e2=sum(r > 1e-2 for r in rels), n_over_2e2=sum(r > 2e-2 for r in rels), n_over_tol=len(bad), tol=tol,
           tol_cbam=tol_cbam)
    return rows, bad


def _teacher_forced_grads(model, sd, batch, masks, rows=None, label=""):
    """oracle/barvae_emul.py run on the CUDA path's stored forward state (see its docstring): returns the emulation's
    gradients and checks the per-layer forward parity it measured on the way"""
    import barvae_emul as E
    from gpu_util import cuda_forward_state
    E.TEACH, E.LOCAL = cuda_forward_state(model, rows), []
    try:
        eloss, egen, ez, egrads = E.train_grads(sd, batch, masks, bce_only_rows=None)
        local = E.LOCAL
    finally:
        E.TEACH = E.LOCAL = None
    # per-layer forward parity: every stored tensor, recomputed on the CPU from the CUDA path's own inputs, is within one
    # bf16 unit of what the CUDA path stored for all but <= 5e-3 of the elements (measured worst: 3.8e-3, the normalised
    # activation of the 32-channel stem) and never further than 16 units (measured 5.2, on the 3x2 maps whose statistics are
    # sums of 6 values); arg-max routes agree to the same fraction
    worst = sorted(local, key=lambda r: -r[2])[:6]
    report(test=label + "_layer_forward", n_checks=len(local), worst_fraction_off=worst,
           max_units=max(r[3] for r in local))
    offenders = [(tag, what, frac, mx) for tag, what, frac, mx in local if frac > 5e-3 or mx > 16.0]
    return eloss, egen, egrads, offenders


def test_all_gradients_vs_storage_emulation(golden, oracle):
    """ALL gradient tensors of one training step against oracle/barvae_emul.py -- the fp32 oracle with the kernels'
    storage roundings (bf16 operands, stored activations and inter-layer gradients, the saved bf16 normalised activation in
    the norm-block backward) -- evaluated on the CUDA path's own stored forward state (teacher forcing: free-running, two
    bf16-storage implementations drift to the bf16 floor within a few layers and the backward pass, which routes through
    arg-max positions and ReLU masks, then differs by tens of per cent per tensor).
    Held: every layer's forward result within one bf16 unit of the stored one (<= 5e-3 of the elements off); every gradient
    tensor -- all 211 that carry signal, no share-based skipping: contraction weights, norm affine parameters, embedding
    within 6e-2 rel-Frobenius (measured over 16 runs: median 1.0e-2 per run, worst tensor 1.9e-2 .. 4.1e-2; 0.03 % at the last
    decoder block growing to 1.5-4 % at the first encoder block ~100 bf16 gradient roundings further back -- each rounding is
    a non-linear op, so even this linear backward pass decorrelates at 2^-9/sqrt(3) per stage), the CBAM attention weights
    within 3e-1 (measured worst per run 3e-2 .. 1e-1, once 2.1e-1: _emul_compare); the 4 analytically-zero biases exactly
    zero, the 4 unused bn1 affine parameters without gradient."""
    from gpu_util import keep_forward_state
    O, c = oracle, golden["lively"]
    Model = pkg("graph.model").Model
    Loss = pkg("graph.loss.bar_loss").Loss
    sd = O.make_state_dict(O.generator_spec(), c["seed_w"], "lively")
    batch = O.make_inputs(c["B"], c["seed_x"])
    masks = O.draw_dropout_masks(c["B"], 77)
    model = _load(Model(), sd).train()
    keep_forward_state(model)
    note, pre_note, phrase, position = (t.cuda() for t in batch)
    gen, z, pre_z, pf = model(note, pre_note, phrase, position, True, tuple(m.cuda() for m in masks))
    loss = Loss()(gen, note, True)
    loss.backward()
    torch.cuda.synchronize()
    eloss, egen, egrads, offenders = _teacher_forced_grads(model, sd, batch, masks, label="all_grads")
    rows, bad = _emul_compare(model, egrads, 6e-2, "all_grads_vs_emulation", 3e-1)
    assert not offenders, offenders[:10]
    assert abs(float(loss.detach()) - float(eloss)) < 1e-4 * float(eloss), (float(loss.detach()), float(eloss))
    assert not bad, bad[:10]


def test_b512_matches_golden_and_emulation(golden, oracle):
    """The benchmarked size (512 bars: TMA boxes spanning many samples, persistent-tile wrap-around, split-K waves).
    The generator has no batch-coupled op, so with the golden pair placed at samples 0 and 511 (random bars in between)
    recon / z / pf of those samples must match the reference-generated golden values to the B=2 tolerances; and with the
    loss restricted to those two samples the parameter gradients of the whole 512-bar backward pass must equal the
    storage-precision emulation's gradients of the pair alone (teacher-forced on the pair's rows of the 512-bar forward
    state; the other 510 bars pass through every kernel and contribute exact zeros)."""
    from gpu_util import keep_forward_state
    O, c = oracle, golden["lively"]
    Model = pkg("graph.model").Model
    sd = O.make_state_dict(O.generator_spec(), c["seed_w"], "lively")
    pair = O.make_inputs(c["B"], c["seed_x"])
    pmask = O.draw_dropout_masks(c["B"], 77)
    assert c["B"] == 2
    B = 512
    fill = O.make_inputs(B, 4242)
    fmask = O.draw_dropout_masks(B, 99)
    sel = [0, B - 1]
    batch = [t.clone() for t in fill]
    mk = [m.clone() for m in fmask]
    for t, pt in zip(batch, pair):
        t[sel] = pt
    for m, pm in zip(mk, pmask):
        m[sel] = pm
    model = _load(Model(), sd).train()
    keep_forward_state(model)
    note, pre_note, phrase, position = (t.cuda() for t in batch)
    gen, z, pre_z, pf = model(note, pre_note, phrase, position, True, tuple(m.cuda() for m in mk))
    want = c["train_pre"]
    m = dict(test="b512_forward", gen_maxabs=float((gen[sel].detach().cpu() - want["gen"]).abs().max()),
             gen_meanabs=float((gen[sel].detach().cpu() - want["gen"]).abs().mean()),
             z=rel_fro(z[sel], want["z"]), pre_z=rel_fro(pre_z[sel], want["pre_z"]), pf=rel_fro(pf[sel], want["pf"]))
    report(**m)
    assert m["gen_maxabs"] < 6e-2 and m["gen_meanabs"] < 1e-2, m
    assert m["z"] < 2e-2 and m["pre_z"] < 2e-2 and m["pf"] < 2e-2, m
    # BCE over the two golden samples only == Loss()(gen_pair, note_pair, True) up to its non-differentiable count term
    idx = torch.tensor(sel, device="cuda")
    loss = torch.nn.functional.binary_cross_entropy(gen[idx], note[idx])
    loss.backward()
    torch.cuda.synchronize()
    eloss, egen, egrads, offenders = _teacher_forced_grads(model, sd, pair, pmask, rows=sel, label="b512")
    keep_forward_state(model, False)
    ebce = float(O.bce_mean(egen, pair[0]))
    rows, bad = _emul_compare(model, egrads, 6e-2, "b512_grads_vs_emulation", 3e-1)
    assert not offenders, offenders[:10]
    assert abs(float(loss.detach()) - ebce) < 1e-4 * ebce, (float(loss.detach()), ebce)
    assert not bad, bad[:10]


def test_model_eval_and_sampling_vs_golden(golden, oracle):
    """is_train=False path (graph/model.py:34-41) and the maker_bar.py:32-44 sampling loop."""
    O, c = oracle, golden["lively"]
    Model = pkg("graph.model").Model
    sd = O.make_state_dict(O.generator_spec(), c["seed_w"], "lively")
    model = _load(Model(), sd).eval()
    _, pre_note, phrase, position = O.make_inputs(c["B"], c["seed_x"])
    g = torch.Generator().manual_seed(5)
    torch.randn(c["B"], O.LATENT, generator=g)
    zz = torch.randn(c["B"], O.LATENT, generator=g)
    with torch.no_grad():
        out = model(zz.cuda(), pre_note.cuda(), phrase.cuda(), position.cuda(), False)
    e = float((out.cpu() - c["model_eval"]).abs().max())
    em = float((out.cpu() - c["model_eval"]).abs().mean())
    report(test="model_eval", maxabs=e, meanabs=em)
    assert e < 6e-2 and em < 1e-2, (e, em)
    # sampling loop: binarised bars must match except where the reference probability is within 0.02 of the threshold
    s = golden["sample"]
    maker = pkg("maker_bar")
    roll, probs = maker.sample_songs(model, s["latents"].cuda(), s["music_length"], return_first_probs=True)
    e0 = float((probs.cpu() - s["first_probs"]).abs().max())
    mism = float((roll[0].cpu().to(torch.uint8) != s["roll"]).float().mean())
    report(test="sampling", first_probs_maxabs=e0, roll_mismatch=mism)
    # bars feed back into the next bar / phrase (maker_bar.py:38-44): a cell that flips near the 0.3 threshold changes
    # everything after it, so the roll as a whole only gets a loose bound; the first bar is held tight
    first_bar_mism = float((roll[0, :96].cpu().to(torch.uint8) != s["roll"][:96]).float().mean())
    assert e0 < 6e-2 and first_bar_mism < 1e-2 and mism < 8e-2, (e0, first_bar_mism, mism)


def test_adam_two_steps_vs_golden(golden, oracle):
    """two full training steps (forward, Loss, backward, fused flat Adam) vs the reference's torch.optim.Adam run."""
    O, c = oracle, golden["lively"]
    eng = pkg("engine")
    Model = pkg("graph.model").Model
    Loss = pkg("graph.loss.bar_loss").Loss
    sd = O.make_state_dict(O.generator_spec(), c["seed_w"], "lively")
    model = _load(Model(), sd).train()
    flat = model.flatten_parameters()
    batch = tuple(t.cuda() for t in O.make_inputs(c["B"], c["seed_x"]))
    masks = tuple(m.cuda() for m in O.draw_dropout_masks(c["B"], 77))
    losses = []
    for step in (1, 2):
        flat.attach_grads()
        gen = model(*batch, True, masks)[0]
        loss = Loss()(gen, batch[0], True)
        loss.backward()
        eng.adam_step(flat, 0.002, step)
        losses.append(float(loss))
    want = [float(v) for v in c["adam2"]["losses"]]
    report(test="adam2", losses=losses, want=want)
    assert abs(losses[0] - want[0]) < 2e-2 * want[0]
    # Adam's first steps move every weight by ~lr regardless of gradient scale: the second loss is sensitive to sign
    # flips of tiny gradients, so it gets a looser bound
    assert abs(losses[1] - want[1]) < 0.15 * want[1], (losses, want)
    dg = O.grad_digest(OrderedDict((k, v.detach().cpu()) for k, v in model.state_dict().items()))
    worst = max(abs(float(dg[k][1]) - float(w[1])) / (float(w[1]) + 1e-12) for k, w in c["adam2"]["param_digest"].items())
    report(test="adam2_params", worst_abs_sum_rel=worst)
    assert worst < 5e-2, worst


def test_state_dict_roundtrip_and_freeze(oracle):
    """reference key names (68 + 86 + 67 = 221, no buffers); requires_grad toggling (agent/barGen.py:143-149)."""
    O = oracle
    Model = pkg("graph.model").Model
    model = Model()
    assert list(model.state_dict().keys()) == list(O.generator_spec().keys())
    model = model.cuda()
    for p in model.parameters():
        p.requires_grad = False
    note, pre_note, phrase, position = (t.cuda() for t in O.make_inputs(1, 3))
    out = model(note, pre_note, phrase, position)[0]
    assert not out.requires_grad
    for p in model.parameters():
        p.requires_grad = True
    out = model(note, pre_note, phrase, position)[0]
    assert out.requires_grad


def test_fused_statistics_path_matches(oracle):
    """opt-in fusion of the InstanceNorm statistics into the convolution epilogue (bvae_conv_desc.stats): same encoder
    output as the separate statistics pass up to bf16 re-rounding downstream (measured 3e-3 on z; bound 1e-2)."""
    O = oracle
    eng = pkg("engine")
    Model = pkg("graph.model").Model
    sd = O.make_state_dict(O.generator_spec(), 11, "lively")
    model = _load(Model(), sd).eval()
    note = O.make_inputs(3, 5)[0].cuda()
    with torch.no_grad():
        z0 = model.encoder(note)
        eng.set_fuse_stats(True)
        try:
            z1 = model.encoder(note)
        finally:
            eng.set_fuse_stats(False)
    e = rel_fro(z1, z0)
    report(test="fused_stats", z_rel=e)
    assert e < 1e-2, e


def test_trainer_stream_and_segment_paths_agree(oracle, monkeypatch):
    """GeneratorTrainer.step with the branch / weight-gradient streams (the default) must give the same parameters as
    the fully serial path: the overlap only reorders independent launches.  The fp32 atomics make the weight gradients
    differ in summation order from run to run and Adam's normalisation turns that into update-sized differences on
    near-zero gradients, so the comparison is against the run-to-run floor of the SERIAL path (two serial runs), in
    units of the mean parameter update; `step_from_host` must match as well."""
    O = oracle
    Model = pkg("graph.model").Model
    Trainer = pkg("trainer").GeneratorTrainer
    sd = O.make_state_dict(O.generator_spec(), 21, "lively")
    batch = tuple(t.cuda() for t in O.make_inputs(4, 9))
    hbatch = tuple(t.cpu().pin_memory() for t in batch)
    masks = tuple(m.cuda() for m in O.draw_dropout_masks(4, 5))

    def run(env, from_host=False):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        model = _load(Model(), sd).train()
        tr = Trainer(model, lr=0.002)
        start = tr.flat.data.clone()
        for _ in range(2):
            loss = tr.step_from_host(*hbatch, masks) if from_host else tr.step(*batch, masks)
        torch.cuda.synchronize()
        return tr.flat.data.clone(), start, float(loss)

    off = {"BVAE_STREAMS": "0", "BVAE_WGRAD_STREAM": "0"}
    on = {"BVAE_STREAMS": "1", "BVAE_WGRAD_STREAM": "1"}
    serial, start, l0 = run(off)
    serial2, _, _ = run(off)
    overlap, _, l1 = run(on)
    host, _, l2 = run(on, from_host=True)
    upd = (serial - start).abs().mean().item()
    floor = (serial2 - serial).abs().mean().item() / upd          # run-to-run: fp32 atomics order x Adam's normalisation
    e1 = (overlap - serial).abs().mean().item() / upd
    e2 = (host - serial).abs().mean().item() / upd
    report(test="trainer_paths", floor=floor, overlap_vs_serial=e1, host_vs_serial=e2, losses=[l0, l1, l2])
    assert e1 < 2 * floor + 0.05 and e2 < 2 * floor + 0.05, (floor, e1, e2)
    # loss of the SECOND step: after one Adam step every weight has moved by ~lr whatever its gradient, so it inherits the
    # same run-to-run sensitivity (measured up to 2 %; test_adam_two_steps_vs_golden allows 15 % for the same reason)
    assert abs(l1 - l0) < 0.1 * abs(l0) and abs(l2 - l0) < 0.1 * abs(l0), (l0, l1, l2)


def test_decoder_branch_streams_agree(oracle, monkeypatch):
    """decoder branches on a companion stream (engine.fork, the default) must give the same parameters after two steps as
    the one-stream launch order (BVAE_DEC_STREAMS=0), within the run-to-run floor of the latter (same protocol as
    test_trainer_stream_and_segment_paths_agree)"""
    O = oracle
    Model = pkg("graph.model").Model
    Trainer = pkg("trainer").GeneratorTrainer
    sd = O.make_state_dict(O.generator_spec(), 21, "lively")
    batch = tuple(t.cuda() for t in O.make_inputs(4, 9))
    masks = tuple(m.cuda() for m in O.draw_dropout_masks(4, 5))

    def run(flag):
        monkeypatch.setenv("BVAE_DEC_STREAMS", flag)
        model = _load(Model(), sd).train()
        tr = Trainer(model, lr=0.002)
        start = tr.flat.data.clone()
        for _ in range(2):
            loss = tr.step(*batch, masks)
        torch.cuda.synchronize()
        return tr.flat.data.clone(), start, float(loss)

    a, start, l0 = run("0")
    a2, _, _ = run("0")
    b, _, l1 = run("1")
    upd = (a - start).abs().mean().item()
    floor = (a2 - a).abs().mean().item() / upd
    e = (b - a).abs().mean().item() / upd
    report(test="decoder_branch_streams", floor=floor, forked_vs_default=e, losses=[l0, l1])
    assert e < 2 * floor + 0.05 and abs(l1 - l0) < 0.1 * abs(l0), (floor, e, l0, l1)


def test_deterministic_mode_forward_is_bit_identical(oracle, monkeypatch):
    """bvae_set_deterministic(1): every forward reduction runs in a fixed order (include/barvae.h), so recon / z / pf and the
    loss are BIT-identical between runs and between the stream-overlapped and the serial schedule -- any difference would be a
    race, not summation order.  The default mode's spread (fp32 atomics order, amplified layer by layer through bf16
    re-rounding) is measured beside it and only reported."""
    O = oracle
    lib = pkg("_lib")
    Model = pkg("graph.model").Model
    Loss = pkg("graph.loss.bar_loss").Loss
    sd = O.make_state_dict(O.generator_spec(), 11, "lively")
    batch = tuple(t.cuda() for t in O.make_inputs(3, 21))
    masks = tuple(m.cuda() for m in O.draw_dropout_masks(3, 77))
    model = _load(Model(), sd).train()

    def run(streams):
        monkeypatch.setenv("BVAE_STREAMS", "1" if streams else "0")
        with torch.no_grad():
            gen, z, pre_z, pf = model(*batch, True, masks)
            loss = Loss()(gen, batch[0], True)
        torch.cuda.synchronize()
        return gen.clone(), z.clone(), pf.clone(), loss.clone()

    spread = []
    base = run(True)
    for _ in range(3):
        r = run(True)
        spread.append(max(float((a - b).abs().max()) for a, b in zip(base[:3], r[:3])))
    lib.set_deterministic(True)
    try:
        ref = run(True)
        for streams in (True, False, True, False):
            r = run(streams)
            for a, b in zip(ref, r):
                assert torch.equal(a, b), "deterministic mode: forward results differ between runs (streams=%s)" % streams
    finally:
        lib.set_deterministic(False)
    report(test="deterministic_forward", default_mode_maxabs_spread=spread, det_loss=float(ref[3]), default_loss=float(base[3]))
    assert abs(float(ref[3]) - float(base[3])) < 2e-2 * abs(float(base[3]))


def test_graph_replayed_step_matches_eager_steps(oracle):
    """GeneratorTrainer with CUDA-graph replay (the default after 3 eager steps per input signature; forced after 1 here):
    four steps -- one eager, one that captures, two replays -- must leave the same parameters as four eager steps, within the
    run-to-run floor of the eager path (the replay runs exactly the captured launches: forward + backward on all streams,
    Adam with its step-dependent scalars read from device memory, operand repack), and the replayed losses must follow
    the eager ones.  Injected dropout masks make the two runs comparable."""
    O = oracle
    Model = pkg("graph.model").Model
    Trainer = pkg("trainer").GeneratorTrainer
    sd = O.make_state_dict(O.generator_spec(), 21, "lively")
    batch = tuple(t.cuda() for t in O.make_inputs(4, 9))
    masks = tuple(m.cuda() for m in O.draw_dropout_masks(4, 5))

    def run(graph):
        model = Model()
        model.load_state_dict(sd)
        tr = Trainer(model.cuda().train(), lr=0.002, use_graph=graph)
        tr.graph_after = 1
        start = tr.flat.data.clone()
        losses = [float(tr.step(*batch, masks)) for _ in range(4)]
        torch.cuda.synchronize()
        return tr.flat.data.clone(), start, losses, tr

    a, start, la, _ = run(False)
    a2, _, _, _ = run(False)
    b, _, lb, tr = run(True)
    assert len(tr._graphs) == 1 and tr.use_graph and tr.step_count == 4           # captured once, replayed, counter in step
    upd = (a - start).abs().mean().item()
    floor = (a2 - a).abs().mean().item() / upd
    e = (b - a).abs().mean().item() / upd
    report(test="graph_step", floor=floor, graph_vs_eager=e, losses_eager=la, losses_graph=lb)
    assert e < 2 * floor + 0.05, (floor, e)
    assert all(abs(x - y) < 0.1 * abs(x) for x, y in zip(la, lb)), (la, lb)
    # a different batch size is a different signature: eager again, then its own graph
    small = tuple(t[:2] for t in batch)
    for _ in range(3):
        tr.step(*small, tuple(m[:2] for m in masks))
    assert len(tr._graphs) == 2
