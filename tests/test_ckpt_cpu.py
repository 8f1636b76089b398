"""CPU: reference-format checkpoints -- torch.optim.Adam state_dict <-> the flat Adam state (agent/barGen.py:151-197)."""
import torch

from gpu_util import pkg


def _tiny_adam_state(params, steps=2, seed=0):
    """what the reference stores under 'gen_optimizer1': torch.optim.Adam(generator.parameters()).state_dict()"""
    g = torch.Generator().manual_seed(seed)
    clones = [torch.nn.Parameter(p.detach().clone()) for p in params]
    opt = torch.optim.Adam(clones, lr=0.002)
    for _ in range(steps):
        for c in clones:
            c.grad = torch.randn(c.shape, generator=g) * 1e-2
        opt.step()
    return opt.state_dict(), clones


def test_torch_adam_state_roundtrip():
    Model = pkg("graph.model").Model
    Trainer = pkg("trainer").GeneratorTrainer
    model = Model()
    tr = Trainer(model, lr=0.1)
    params = list(model.parameters())
    sd, _ = _tiny_adam_state(params[:40])                 # a reference checkpoint may cover a prefix ...
    tr.load_state_dict(sd)
    assert tr.step_count == 2 and tr.lr == 0.002 and tr.betas == (0.9, 0.999)
    out = tr.torch_state_dict()
    assert out["param_groups"][0]["params"] == list(range(len(params)))
    for i in range(40):
        assert torch.equal(out["state"][i]["exp_avg"], sd["state"][i]["exp_avg"])
        assert torch.equal(out["state"][i]["exp_avg_sq"], sd["state"][i]["exp_avg_sq"])
        assert int(out["state"][i]["step"]) == 2
    assert float(out["state"][40]["exp_avg"].abs().sum()) == 0.0
    # ... or trailing entries this model does not have (the reference's Refiner parameters): ignored
    extra = torch.nn.Parameter(torch.zeros(3, 5))
    sd2, _ = _tiny_adam_state(params + [extra], steps=1)
    tr.load_state_dict(sd2)
    assert tr.step_count == 1
    last = len(params) - 1
    assert torch.equal(tr.torch_state_dict()["state"][last]["exp_avg"], sd2["state"][last]["exp_avg"])
    # a shape mismatch inside the covered range is an error, not a silent skip
    bad = {"state": {0: {"step": torch.tensor(1.0), "exp_avg": torch.zeros(7), "exp_avg_sq": torch.zeros(7)}},
           "param_groups": [{"lr": 0.002, "params": [0]}]}
    try:
        tr.load_state_dict(bad)
    except ValueError:
        pass
    else:
        raise AssertionError("shape mismatch accepted")
    # torch.optim.Adam itself accepts the exported dict (a maintainer can go back to the stock optimiser)
    opt = torch.optim.Adam(params, lr=0.5)
    tr.load_state_dict(sd)
    opt.load_state_dict(tr.torch_state_dict())
    assert opt.param_groups[0]["lr"] == 0.002
    assert torch.equal(opt.state[params[5]]["exp_avg"], sd["state"][5]["exp_avg"])


def test_algorithmic_flops_convention():
    """engine.GemmLayer.algorithmic_flops follows SURVEY.md section 8: Conv 2*Cout*Hout*Wout*Cin*kh*kw, ConvT
    2*Cin*Hin*Win*Cout*kh*kw, Linear 2*in*out (what bench.py's per-class tensor roofline divides by time)"""
    import torch.nn as nn
    gemm_of = pkg("graph.encodingBlock").gemm_of
    conv = gemm_of(nn.Conv2d(64, 128, 3, 2, 1, bias=False))                 # 24x15 -> 12x8
    assert conv.algorithmic_flops(5, 24 * 15, 12 * 8) == 5 * 2 * 128 * 12 * 8 * 64 * 9 and conv.channel_class() == "ch<=64"
    convt = gemm_of(nn.ConvTranspose2d(1024, 512, 4, 2, 1, output_padding=(0, 1)))   # 6x3 -> 12x7
    assert convt.algorithmic_flops(2, 6 * 3, 12 * 7) == 2 * 2 * 1024 * 6 * 3 * 512 * 16 and convt.channel_class() == "ch>=256"
    lin = gemm_of(nn.Linear(1024, 1152))
    assert lin.algorithmic_flops(7, 1, 1) == 7 * 2 * 1024 * 1152
    assert gemm_of(nn.Conv2d(256, 128, 1)).channel_class() == "ch128"
    # the encoder's contraction layers at 96x60 add up to SURVEY's 1.1291 GFLOP per bar (forward)
    Model = pkg("graph.model").Model
    enc = Model().encoder
    total, h, w = 0.0, 96, 60
    for stem, first, second in ((enc.time_pitch, "time", "pitch"), (enc.pitch_time, "pitch", "time")):
        g1, g2 = gemm_of(getattr(stem, first)), gemm_of(getattr(stem, second))
        h1, w1 = g1.out_hw(96, 60)
        h2, w2 = g2.out_hw(h1, w1)
        total += g1.algorithmic_flops(1, 96 * 60, h1 * w1) + g2.algorithmic_flops(1, h1 * w1, h2 * w2)
    h, w = 48, 30
    for layer in enc.layers:
        for m in layer.modules():
            if isinstance(m, nn.Conv2d) and m.kernel_size == (3, 3):
                g = gemm_of(m)
                oh, ow = g.out_hw(h, w)
                total += g.algorithmic_flops(1, h * w, oh * ow)
                h, w = oh, ow
    total += gemm_of(enc.linear).algorithmic_flops(1, 1, 1)
    assert abs(total / 1e9 - 1.1291) < 0.004, total / 1e9            # the CBAM 1x1s / 3x3 gate make up the rest


def test_profile_aggregation_by_class(monkeypatch):
    """engine.profile_begin / _timed / profile_end with stub events: per-tag totals, per-detail totals, per-class
    [ms, flops, launches]; outside a profiling window _timed is a plain call"""
    eng = pkg("engine")

    class Ev:
        def __init__(self, enable_timing=False):
            pass

        def record(self, *a):
            pass

        def elapsed_time(self, other):
            return 2.0

    monkeypatch.setattr(torch.cuda, "Event", Ev)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    calls = []
    assert eng._timed("conv_gemm", lambda a: calls.append(a) or 7, 1, detail="x", flops=5.0, cls="ch128") == 7
    assert not eng.profiling()
    eng.profile_begin()
    assert eng.profiling()
    eng._timed("conv_gemm", lambda: 0, detail="f conv", flops=10.0, cls="ch>=256")
    eng._timed("conv_gemm", lambda: 0, detail="f conv", flops=30.0, cls="ch>=256")
    eng._timed("wgrad_gemm", lambda: 0, detail="w conv", flops=4.0, cls="ch<=64")
    eng._timed("nb_forward", lambda: 0, detail="C64")
    out = eng.profile_end()
    assert not eng.profiling()
    assert out["conv_gemm"] == 4.0 and out["n_conv_gemm"] == 2 and out["n_gemm"] == 3 and out["nb_forward"] == 2.0
    assert out["detail"]["conv_gemm:f conv"] == (4.0, 2)
    assert out["classes"] == {"ch>=256": [4.0, 40.0, 2], "ch<=64": [2.0, 4.0, 1]}
