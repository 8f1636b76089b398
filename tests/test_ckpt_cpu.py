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
