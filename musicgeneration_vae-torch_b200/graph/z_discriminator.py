"""Latent discriminators of the adversarial phase (reference: graph/z_discriminator.py:7-54): 1152 -> 512 -> 512 ->
512 -> 512 -> 1 with ReLU between and a sigmoid at the end; same ``net.{0,2,4,6,8}`` parameter names."""
import torch.nn as nn

from ._mlp import MLPDiscriminator
from .weights_initializer import weights_init


class _ZDiscriminator(MLPDiscriminator):
    def __init__(self, z_dim=1152):
        super().__init__()
        self.z_dim = z_dim
        self.net = nn.Sequential(nn.Linear(z_dim, 512), nn.ReLU(True), nn.Linear(512, 512), nn.ReLU(True),
                                 nn.Linear(512, 512), nn.ReLU(True), nn.Linear(512, 512), nn.ReLU(True),
                                 nn.Linear(512, 1), nn.Sigmoid())
        self.apply(weights_init)

    def _stack(self):
        return [(self.net[i], True) for i in (0, 2, 4, 6)]

    def _head(self):
        return self.net[8]

    def forward(self, x):
        return self._run(x)


class PhraseZDiscriminator(_ZDiscriminator):
    pass


class BarZDiscriminator(_ZDiscriminator):
    pass
