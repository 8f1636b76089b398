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
            if first is None:
                first = probs.clone()
            pre_bar = (probs > threshold).float()                        # maker_bar.py:39
            out[:, k] = pre_bar[:, 0]
            k += 1
        pre_phrase = out[:, k - 4:k].reshape(S, 1, 384, 60)
    model.train(was_training)
    roll = out.reshape(S, music_length * 4 * 96, 60)
    return (roll, first) if return_first_probs else roll
