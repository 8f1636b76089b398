"""CBAM parameter containers (reference: graph/cbam.py:7-68).

In this implementation the attention is never a stand-alone pass: channel attention (avg/max pool -> shared MLP ->
sigmoid gate) and spatial attention (mean/max over C -> 3x3 conv -> sigmoid gate) are fused with the
InstanceNorm, residual add and activation of the parent block inside bvae_nb_forward / bvae_nb_backward.  The
classes below exist so that ``state_dict`` keys (``...cbam.channel_attention.conv1.weight`` etc.) and
initialisation match the reference.
"""
import torch.nn as nn

from .weights_initializer import weights_init


class ChannelAttention(nn.Module):
    def __init__(self, channel):
        super().__init__()
        self.conv1 = nn.Conv2d(channel, channel // 16, 1, bias=False)     # cbam.py:14
        self.conv2 = nn.Conv2d(channel // 16, channel, 1, bias=False)     # cbam.py:15
        self.apply(weights_init)


class SpatialAttention(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv = nn.Conv2d(2, 1, 3, padding=1, bias=False)             # cbam.py:36
        self.apply(weights_init)


class CBAM(nn.Module):
    def __init__(self, channel):
        super().__init__()
        self.channel = channel
        self.channel_attention = ChannelAttention(channel)
        self.spatial_attention = SpatialAttention()
        self.apply(weights_init)

    def params(self):
        """(W1 [C/16,C,1,1], W2 [C,C/16,1,1], Wsp [1,2,3,3]) as consumed by engine.NormBlock."""
        return (self.channel_attention.conv1.weight, self.channel_attention.conv2.weight,
                self.spatial_attention.conv.weight)

    def forward(self, x):
        raise RuntimeError("CBAM is fused into its parent block (bvae_nb_forward); it has no stand-alone kernel")
