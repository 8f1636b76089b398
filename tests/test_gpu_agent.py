"""GPU (B200): the entry points above the hot path -- BarGen trainer (agent/barGen.py surface), checkpoints with the
reference's keys, the sampling loop, and the optional VAE head (reparameterise + KL, old/ semantics)."""
import os

import pytest
import torch
import torch.nn.functional as F

from gpu_util import pkg, rel_fro, report

pytestmark = pytest.mark.gpu


def test_bargen_trains_checkpoints_and_samples(tmp_path):
    Config = pkg("config").Config
    BarGen = pkg("agent.barGen").BarGen
    ds = pkg("data.bar_dataset").SyntheticBars(n_items=4, bars_per_item=2, batch_size=2, seed=3)

    class Cfg(Config):
        root_path = str(tmp_path)
        batch_size = 2
        epoch = 2
        pretraining_step_size = 0          # checkpoint condition epoch > pretraining + 50 is exercised explicitly below

    agent = BarGen(Cfg(), dataset=ds)
    note, pre_note, pre_phrase, position = agent.make_batch([ds[0], ds[1]])      # agent/barGen.py:134-141
    assert note.shape == (4, 1, 96, 60) and pre_phrase.shape == (4, 1, 384, 60) and position.dtype == torch.long
    # Eight epochs over the same 8 bars (2 optimiser steps each).  The reference's training loss is BCE + 0.005 x (number of
    # notes the thresholded output misses) and the second term is not differentiable -- it RISES while the BCE falls (a model
    # that learns "mostly silence" misses every note: measured 5.27 -> 5.96 in total) -- so the criterion is the
    # reconstruction BCE of the training bars in eval mode, before vs after.  From the reference's N(-1,1) initialisation
    # (not reproducible through torch.manual_seed: the eval BCE before training already ranges 1.45 .. 1.93) the drop after 8
    # epochs was measured at 21 % .. 53 % over 12 runs on the B200 (profiles/r3/agent_test_seed_scan.jsonl; the slow runs
    # plateau at 0.77 and stay there for 16 epochs): the test asks for more than 10 %, a quarter failed one run in four.
    Loss = pkg("graph.loss.bar_loss").Loss
    dev = agent.device
    full = tuple(t.to(dev) for t in agent.make_batch([ds[i] for i in range(4)]))

    def bce():
        agent.generator.eval()
        with torch.no_grad():
            gen = agent.generator(*full)[0]
            val = float(Loss().parts(gen, full[0], True)[0])
        agent.generator.train()
        return val

    b0 = bce()
    losses = []
    for _ in range(8):
        agent.epoch += 1
        losses.append(agent.train_epoch())
    b1 = bce()
    report(test="bargen", epoch_losses=losses, bce_before=b0, bce_after=b1)
    assert all(l == l for l in losses) and b1 < 0.9 * b0, (b0, b1, losses)
    agent.save_checkpoint(Cfg.checkpoint_file, 1)
    ck = torch.load(os.path.join(str(tmp_path), Cfg.checkpoint_dir, "checkpoint.pth.tar"), weights_only=False)
    keys = list(ck["generator_state_dict"].keys())
    assert all(k.startswith("module.") for k in keys) and len(keys) == 221      # nn.DataParallel prefix, barGen.py:180
    before = {k: v.clone() for k, v in agent.generator.state_dict().items()}
    agent2 = BarGen(Cfg(), dataset=ds)                                            # picks the checkpoint up (barGen.py:108)
    for k, v in agent2.generator.state_dict().items():
        assert torch.equal(v, before[k]), k
    roll = agent2.generate(music_length=1, songs=3)                               # maker_bar.py:32-44
    assert roll.shape == (3, 4 * 96, 60) and set(roll.unique().tolist()) <= {0.0, 1.0}


def test_vae_head_forward_backward(oracle):
    """Model(vae_head=True): (recon, mu, logvar) with z = mu + eps*exp(0.5*logvar) (old/.../bar_v1/encoder.py:60-63) and
    the KL of old/graphs/losses/bar_loss.py:10-18; gradients of the head vs PyTorch autograd on the same latents."""
    O = oracle
    M = pkg("graph.model")
    VAELoss = pkg("graph.loss.bar_loss").VAELoss
    torch.manual_seed(0)
    model = M.Model(vae_head=True)
    sd = O.make_state_dict(O.generator_spec(), 11, "lively")
    model.load_state_dict(sd, strict=False)
    model = model.cuda().eval()
    with torch.no_grad():
        model.logvar_head.weight.normal_(0, 0.01)
    B = 2
    note, pre_note, phrase, position = (t.cuda() for t in O.make_inputs(B, 21))
    g = torch.Generator(device="cuda").manual_seed(1)
    eps = (torch.randn(B, 1152, device="cuda", generator=g), torch.randn(B, 1152, device="cuda", generator=g))
    recon, mu, logvar = model(note, pre_note, phrase, position, True, None, eps)
    pre_mu, pre_logvar = model.last_pre
    loss = VAELoss()(recon, note, mu, logvar, pre_mu, pre_logvar)
    loss.backward()
    assert recon.shape == (B, 1, 96, 60) and mu.shape == logvar.shape == (B, 1152)
    # reference composition on the same mu / logvar: KL average of note and pre_note (bar_loss.py:12,18)
    kl = (O.kl_sum(mu.detach().cpu(), logvar.detach().cpu()) + O.kl_sum(pre_mu.detach().cpu(), pre_logvar.detach().cpu())) / 2
    bce = O.bce_mean(recon.detach().cpu(), note.cpu())
    want = float(bce + kl)
    report(test="vae_head", loss=float(loss), want=want)
    assert abs(float(loss) - want) < 1e-3 * abs(want)
    gw = model.logvar_head.weight.grad
    assert gw is not None and torch.isfinite(gw).all() and float(gw.abs().sum()) > 0
    assert all(torch.isfinite(p.grad).all() for n, p in model.named_parameters() if p.grad is not None)
