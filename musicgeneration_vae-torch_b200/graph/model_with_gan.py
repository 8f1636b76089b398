"""Generator composition of the adversarial phase (reference: graph/model_with_gan.py:10-38): the same three trunks as
graph/model.py, returning additionally the encoder feature of the THRESHOLDED generated bar
(``torch.gt(gen_note, 0.3)``, non-differentiable, :29,36) for the feature discriminator.

``forward(note, pre_note, phrase, position, is_note=True)`` -> ``(gen_note, z, pre_z, phrase_feature,
encoder(fake_note))``; with ``is_note=False`` the ``note`` slot carries the latent and the result is
``(gen_note, encoder(fake_note))``."""
from . import model as _base
from ..data.packed import threshold_pack


class Model(_base.Model):
    def __init__(self):
        super().__init__(vae_head=False)

    def _fake_feature(self, gen_note):
        _, fake = threshold_pack(gen_note.detach(), 0.3, want_bits=False, want_float=True)   # bvae_threshold_pack
        return self.encoder(fake.view_as(gen_note))

    def forward(self, note, pre_note, phrase, position, is_note=True, dropout_masks=None):
        if is_note:
            gen_note, z, pre_z, pf = super().forward(note, pre_note, phrase, position, True, dropout_masks)
            return gen_note, z, pre_z, pf, self._fake_feature(gen_note)
        gen_note = super().forward(note, pre_note, phrase, position, False, dropout_masks)
        return gen_note, self._fake_feature(gen_note)
