"""Sampling loop with the semantics of the reference's maker_bar.py:32-44, batched over S independent songs.

Per phrase index the phrase encoder runs ONCE (its input is constant across the 4 inner bar iterations; the
reference recomputes it 4x), each bar is thresholded at 0.3 on the device, and the 4 bars of a phrase become the next
``pre_phrase``.  Songs are independent, so S songs advance in lock-step as one batch (BASELINE config 5: S = 8192)."""
import torch


@torch.no_grad()
def sample_songs(model, latents, music_length, return_first_probs=False, threshold=0.3):
    """latents: [music_length*4, S, 1152] (the randn draws, one per bar).  Returns uint8-valued float piano-roll
    [S, music_length*4*96, 60] (and optionally the first bar's probabilities)."""
    S = latents.shape[1]
    dev = latents.device
    pre_phrase = torch.zeros(S, 1, 384, 60, device=dev)
    pre_bar = torch.zeros(S, 1, 96, 60, device=dev)
    phrase_idx = [330] + [i for i in range(music_length - 2, -1, -1)]     # maker_bar.py:34
    out = torch.empty(S, music_length * 4, 96, 60, device=dev)
    was_training = model.training
    model.eval()
    first = None
    k = 0
    for idx in range(music_length):
        pos = torch.full((S,), phrase_idx[idx], dtype=torch.long, device=dev)
        phrase_feature = model.phrase_encoder(pre_phrase)               # once per phrase
        for _ in range(4):
            pre_z = model.encoder(pre_bar)
            probs = model.decoder(latents[k], pre_z, phrase_feature, pos)
            if getattr(model, "refiner", None) is not None:
                probs = model.refiner(probs)                             # graph/model.py:41 (Model(refiner=True))
            if first is None:
                first = probs.clone()
            pre_bar = (probs > threshold).float()                        # maker_bar.py:39
            out[:, k] = pre_bar[:, 0]
            k += 1
        pre_phrase = out[:, k - 4:k].reshape(S, 1, 384, 60)
    model.train(was_training)
    roll = out.reshape(S, music_length * 4 * 96, 60)
    return (roll, first) if return_first_probs else roll


def songs_to_host(roll):
    """Device roll [S, T, 60] of {0,1} floats -> numpy float32 on the host, crossing PCIe as bits: one
    bvae_threshold_pack launch (720 B per bar instead of the 23 KB the reference reads back per bar at
    maker_bar.py:40), expanded again with numpy on the host."""
    from .data.packed import threshold_pack, unpack_cells_host
    bits, _ = threshold_pack(roll, 0.5)
    return unpack_cells_host(bits.cpu().numpy(), tuple(roll.shape))


def load_generator(model, filename, device="cuda"):
    """maker_bar.py:25-30: ``checkpoint['generator_state_dict']`` saved from nn.DataParallel (``module.`` prefix); the
    reference's Refiner entries are ignored (graph/refiner.py cannot execute)."""
    ck = torch.load(filename, map_location=device, weights_only=False)
    sd = {(k[len("module."):] if k.startswith("module.") else k): v for k, v in ck["generator_state_dict"].items()}
    own = model.state_dict()
    # Refiner entries load only into Model(refiner=True), and only where the shapes agree: the reference's
    # refiner.layer2.0.weight [8,1,4,4] IS its defect (graph/refiner.py:19)
    sd = {k: v for k, v in sd.items() if not k.startswith("refiner.") or (k in own and tuple(own[k].shape) == tuple(v.shape))}
    model.load_state_dict(sd, strict=False)
    return model


def main(argv=None):
    """The reference script (maker_bar.py:10-54): load ``model/checkpoint_{model_number}.pth.tar``, sample
    ``music_length`` phrases of 4 bars, write ``./test.mid``."""
    import argparse
    from .graph.model import Model
    from .midi import write_midi
    ap = argparse.ArgumentParser()
    ap.add_argument("--model_number", type=int, default=10, help="Music length that want to make.")
    ap.add_argument("--music_length", type=int, default=10, help="Music length that want to make.")
    ap.add_argument("--songs", type=int, default=1, help="independent songs sampled in lock-step (test_<i>.mid)")
    ap.add_argument("--out", default="./test.mid")
    ap.add_argument("--seed", type=int, default=None)
    args = ap.parse_args(argv)
    dev = torch.device("cuda", torch.cuda.current_device())
    model = load_generator(Model().to(dev), "model/checkpoint_{}.pth.tar".format(args.model_number), dev)
    g = torch.Generator(device=dev)
    if args.seed is not None:
        g.manual_seed(args.seed)
    lat = torch.randn(args.music_length * 4, args.songs, 1152, device=dev, generator=g)
    rolls = songs_to_host(sample_songs(model, lat, args.music_length))
    for i, r in enumerate(rolls):
        path = args.out if i == 0 else args.out.replace(".mid", "_%d.mid" % i)
        print(r.shape, write_midi(r, path), "notes ->", path)


if __name__ == "__main__":
    main()
