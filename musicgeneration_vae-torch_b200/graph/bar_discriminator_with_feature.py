"""Feature discriminator (reference: graph/bar_discriminator_with_feature.py:6-24): sigmoid(linear2(linear1(x))) on
the 1152-wide encoder feature of a thresholded bar, both Linears without bias and without a non-linearity between."""
import torch.nn as nn

from ._mlp import MLPDiscriminator
from .weights_initializer import weights_init


class BarFeatureDiscriminator(MLPDiscriminator):
    def __init__(self):
        super().__init__()
        self.linear1 = nn.Linear(1152, 512, bias=False)
        self.linear2 = nn.Linear(512, 1, bias=False)
        self.sigmoid = nn.Sigmoid()
        self.apply(weights_init)

    def _stack(self):
        return [(self.linear1, False)]

    def _head(self):
        return self.linear2

    def forward(self, x):
        return self._run(x.view(-1, 1152))
