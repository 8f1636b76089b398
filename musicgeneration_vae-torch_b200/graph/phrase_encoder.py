"""Phrase encoder 1x384x60 -> 1152 (reference: graph/phrase_encoder.py:8-55): the encoder trunk on a 4-bar input,
global average over the final 12x2 map, bias-free Linear.  PhraseModel wraps PhraseEncoder, which is where the
doubled ``phrase_encoder.phrase_encoder.`` state_dict prefix comes from."""
import torch.nn as nn

from .encoder import _EncoderTrunk
from .weights_initializer import weights_init


class PhraseEncoder(_EncoderTrunk):
    def __init__(self, layers):
        super().__init__()
        self._build(layers, linear_bias=False)


class PhraseModel(nn.Module):
    def __init__(self, layers):
        super().__init__()
        self.phrase_encoder = PhraseEncoder(layers)
        self.apply(weights_init)

    def forward(self, phrase):
        return self.phrase_encoder(phrase)
