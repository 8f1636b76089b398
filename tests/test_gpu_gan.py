"""GPU (B200): the adversarial-phase trainer (agent/barGen_with_gan.py: train_pretrain / train_wae / train_gan schedules on
the generator + four discriminators -- BASELINE config 4), the Refiner wired into Model(refiner=True), and the
gradient-accumulated (micro-batched) training step of BASELINE config 3."""
import os

import pytest
import torch

from gpu_util import pkg, rel_fro, report

pytestmark = pytest.mark.gpu


def _snap(agent):
    return {k: o.flat.data.clone() for k, o in agent._opts.items()}


def _moved(a, b):
    return {k: float((a[k] - b[k]).abs().max()) for k in a}


def test_gan_agent_schedules_and_checkpoint(tmp_path):
    """one iteration of each schedule: exactly the modules the reference steps (agent/barGen_with_gan.py:351-537) change,
    the frozen ones stay bit-identical, losses are finite; the checkpoint carries the reference's ten keys and round-trips"""
    Config = pkg("config").Config
    G = pkg("agent.barGen_with_gan")
    ds = pkg("data.bar_dataset").SyntheticBars(n_items=4, bars_per_item=2, batch_size=2, seed=3)

    class Cfg(Config):
        root_path = str(tmp_path)
        batch_size = 2
        epoch = 1
        pretraining_step_size = 1

    agent = G.BarGen(Cfg(), dataset=ds)
    dev = agent.device
    # "lively" weights everywhere: with the reference initialisation N(-1,1) every ReLU of the z discriminators is dead
    # (z ~ -500) and their gradients are legitimately all zero
    import barvae_oracle as O
    import disc_oracle as D
    agent.generator.load_state_dict(O.make_state_dict(O.generator_spec(), 11, "lively"))
    agent.discriminator.load_state_dict(D.make_conv_state_dict(D.bar_disc_spec(), 7, "lively"))
    agent.discriminator_feature.load_state_dict(D.make_disc_state_dict(D.feature_disc_spec(), 5, "lively"))
    agent.z_discriminator_bar.load_state_dict(D.make_disc_state_dict(D.z_disc_spec(), 5, "lively"))
    agent.z_discriminator_phrase.load_state_dict(D.make_disc_state_dict(D.z_disc_spec(), 6, "lively"))
    pkg("engine").bump_param_epoch()
    batch = tuple(t.to(dev) for t in agent.make_batch([ds[0], ds[1]]))
    valid, fake = torch.ones(4, device=dev), torch.zeros(4, device=dev)
    losses = {}
    rec = lambda name: (lambda l: losses.__setitem__(name, float(l)))

    s0 = _snap(agent)
    agent.epoch = 1
    agent.train_pretrain(*batch, rec("pre_gen"))
    m = _moved(_snap(agent), s0)
    assert m["generator"] > 0 and all(m[k] == 0 for k in m if k != "generator"), m

    s0 = _snap(agent)
    agent.epoch = 2                                    # (epoch + curr_it) % 2 == 1 with curr_it = 1: discriminator step too
    agent.train_wae(*batch, rec("wae_gen"), rec("wae_barz"), rec("wae_phrasez"), fake, valid, 1)
    m = _moved(_snap(agent), s0)
    assert m["generator"] > 0 and m["z_discriminator_bar"] > 0 and m["z_discriminator_phrase"] > 0, m
    assert m["discriminator"] == 0 and m["discriminator_feature"] == 0, m

    s0 = _snap(agent)
    bn_before = agent.discriminator.chord.batch_norm1.running_mean.clone()
    agent.train_gan(*batch, rec("gan_gen"), rec("gan_disc"), rec("gan_feat"), fake, valid, 1)
    m = _moved(_snap(agent), s0)
    assert m["generator"] > 0 and m["discriminator"] > 0 and m["discriminator_feature"] > 0, m
    assert m["z_discriminator_bar"] == 0 and m["z_discriminator_phrase"] == 0, m
    assert not torch.equal(agent.discriminator.chord.batch_norm1.running_mean, bn_before)     # train-mode BatchNorm
    report(test="gan_agent", losses=losses)
    assert all(v == v and abs(v) < 1e4 for v in losses.values()), losses
    # the generator step of train_gan reaches decoder, encoder (through the decoder and through the re-encoded bar) and
    # the phrase encoder; a skipped iteration ((epoch + curr_it) % 2 == 0) leaves the discriminators alone
    s0 = _snap(agent)
    agent.train_gan(*batch, rec("gan_gen2"), rec("x"), rec("y"), fake, valid, 0)
    m = _moved(_snap(agent), s0)
    assert m["generator"] > 0 and m["discriminator"] == 0 and m["discriminator_feature"] == 0, m

    # the per-iteration schedule of agent/barGen_horovod.py: all four discriminators every iteration, then the generator --
    # in the GAN phase with two forwards and one backward (train_add_gan)
    s0 = _snap(agent)
    agent.train_discriminator(*batch, rec("h_barz"), rec("h_phrasez"), rec("h_disc"), rec("h_feat"), fake, valid)
    m = _moved(_snap(agent), s0)
    assert m["generator"] == 0 and all(m[k] > 0 for k in m if k != "generator"), m
    s0 = _snap(agent)
    agent.train_add_gan(*batch, rec("h_gen_gan"), valid)
    m = _moved(_snap(agent), s0)
    assert m["generator"] > 0 and all(m[k] == 0 for k in m if k != "generator"), m
    s0 = _snap(agent)
    agent.train_wae_only(*batch, rec("h_gen_wae"), valid)
    m = _moved(_snap(agent), s0)
    assert m["generator"] > 0 and all(m[k] == 0 for k in m if k != "generator"), m
    assert all(v == v and abs(v) < 1e4 for v in losses.values()), losses
    report(test="gan_agent_horovod_schedule", losses={k: v for k, v in losses.items() if k.startswith("h_")})

    means = agent.train_epoch()                        # a whole epoch through the dispatcher (:284-296)
    agent.config.gan_schedule = "horovod"
    agent.epoch = 3
    means_h = agent.train_epoch()                      # and through the other one (barGen_horovod.py:312-324)
    assert set(means_h) == set(agent._opts) and all(v == v for v in means_h.values())
    agent.config.gan_schedule = "with_gan"
    assert set(means) == set(agent._opts)
    agent.save_checkpoint(Cfg.checkpoint_file, 3)
    ck = torch.load(os.path.join(str(tmp_path), Cfg.checkpoint_dir, "checkpoint.pth.tar"), weights_only=False)
    for key in ("generator_state_dict", "generator_optimizer", "discriminator_state_dict", "disc_optimizer",
                "discriminator_feature_state_dict", "disc_feature_optimizer", "z_discriminator_bar_state_dict",
                "opt_Zdiscriminator_bar_optimizer", "z_discriminator_phrase_state_dict",
                "opt_Zdiscriminator_phrase_optimizer"):                                          # :195-213
        assert key in ck, key
    assert len(ck["discriminator_state_dict"]) == len(D.bar_disc_spec()) and all(k.startswith("module.") for k in ck["generator_state_dict"])
    # stock torch.optim.Adam accepts the stored optimiser state of the convolutional discriminator
    ref_like = pkg("graph.bar_discriminator").BarDiscriminator()
    torch.optim.Adam(ref_like.parameters(), lr=0.002).load_state_dict(ck["disc_optimizer"])
    agent2 = G.BarGen(Cfg(), dataset=ds)
    for name in agent._opts:
        for (k, a), (_, b) in zip(getattr(agent, name).state_dict().items(), getattr(agent2, name).state_dict().items()):
            assert torch.equal(a, b), (name, k)
        assert torch.equal(agent._opts[name].flat.exp_avg, agent2._opts[name].flat.exp_avg), name
    roll = agent2.generate(music_length=1, songs=2)
    assert roll.shape == (2, 4 * 96, 60)


def test_model_with_refiner_matches_oracle_composition(oracle):
    """Model(refiner=True): graph/model.py:24-31,35-41 with the Refiner applied to the decoder output in both branches
    (reference keys: encoder.*, decoder.*, phrase_encoder.*, refiner.*), against oracle model_forward -> refiner_forward"""
    import disc_oracle as D
    O = oracle
    Model = pkg("graph.model").Model
    sd = O.make_state_dict(O.generator_spec(), 11, "lively")
    rsd = D.make_conv_state_dict(D.refiner_spec(), 9, "lively")
    full = dict(sd)
    full.update({"refiner." + k: v for k, v in rsd.items()})
    model = Model(refiner=True)
    assert list(model.state_dict().keys()) == list(sd.keys()) + ["refiner." + k for k in rsd]
    model.load_state_dict(full)
    model = model.cuda().eval()
    batch = O.make_inputs(3, 21)
    with torch.no_grad():
        gen, z, pre_z, pf = model(*(t.cuda() for t in batch))
        ogen = O.model_forward(*batch, sd, True, None)[0]
        want = D.refiner_forward(ogen, {k: v.clone() for k, v in rsd.items()}, False)
        g = torch.Generator().manual_seed(5)
        lat = torch.randn(3, 1152, generator=g)
        gen2 = model(lat.cuda(), batch[1].cuda(), batch[2].cuda(), batch[3].cuda(), False)
        want2 = D.refiner_forward(O.model_forward(lat, batch[1], batch[2], batch[3], sd, False, None),
                                  {k: v.clone() for k, v in rsd.items()}, False)
    e1, e2 = float((gen.cpu() - want).abs().max()), float((gen2.cpu() - want2).abs().max())
    report(test="model_refiner", train_branch_maxabs=e1, eval_branch_maxabs=e2)
    assert e1 < 6e-2 and e2 < 6e-2, (e1, e2)
    # and it trains: one step moves refiner and generator parameters
    model.train()
    tr = pkg("trainer").GeneratorTrainer(model, lr=0.002)
    before = tr.flat.data.clone()
    loss = tr.step(*(t.cuda() for t in batch))
    assert float(loss) == float(loss)
    moved = (tr.flat.data - before).abs()
    off = model.refiner.layer3[0].weight._bvae_off
    assert float(moved[off:off + 100].max()) > 0 and float(moved[:100].max()) > 0


def test_micro_batched_step_equals_full_step(oracle):
    """GeneratorTrainer(micro_bars=2) on 6 bars (chunks 2+2+2, and 4+2 with micro_bars=4) vs the one-pass step: same loss
    (deterministic forward: bit-identical BCE up to the chunk-weighted sum) and the same parameters after two steps within
    the run-to-run floor of the one-pass step (the generator has no batch-coupled op -- SURVEY.md section 8e)"""
    O = oracle
    lib = pkg("_lib")
    Model = pkg("graph.model").Model
    Trainer = pkg("trainer").GeneratorTrainer
    sd = O.make_state_dict(O.generator_spec(), 21, "lively")
    batch = tuple(t.cuda() for t in O.make_inputs(6, 9))
    masks = tuple(m.cuda() for m in O.draw_dropout_masks(6, 5))

    def run(micro):
        model = Model()
        model.load_state_dict(sd)
        tr = Trainer(model.cuda().train(), lr=0.002, micro_bars=micro)
        start = tr.flat.data.clone()
        losses = [float(tr.step(*batch, masks)) for _ in range(2)]
        torch.cuda.synchronize()
        return tr.flat.data.clone(), start, losses

    lib.set_deterministic(True)
    try:
        a, start, la = run(0)
        a2, _, _ = run(0)
        b, _, lb = run(2)
        c, _, lc = run(4)
    finally:
        lib.set_deterministic(False)
    upd = (a - start).abs().mean().item()
    floor = (a2 - a).abs().mean().item() / upd
    e2, e4 = (b - a).abs().mean().item() / upd, (c - a).abs().mean().item() / upd
    report(test="micro_batch", floor=floor, micro2=e2, micro4=e4, losses=[la, lb, lc])
    assert abs(lb[0] - la[0]) < 1e-5 * abs(la[0]) and abs(lc[0] - la[0]) < 1e-5 * abs(la[0]), (la, lb, lc)
    assert e2 < 2 * floor + 0.05 and e4 < 2 * floor + 0.05, (floor, e2, e4)
