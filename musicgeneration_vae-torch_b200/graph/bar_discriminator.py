"""Convolutional piano-roll discriminator of the adversarial phase (reference: graph/bar_discriminator.py:7-217): three
feature towers over the two-bar roll ``cat((pre_note, note), dim=2)`` [B,1,192,60] -- chord (pitch classes summed over
octaves), on/off (per-step note count) and a generic tower -- each ending in a 64-vector, concatenated into one logit.
Same sub-module / parameter / buffer names as the reference (state_dict interchangeable: 87 keys, BatchNorm buffers included).  Convolutions run
through bvae_conv_gemm / bvae_wgrad_gemm, every nn.BatchNorm2d (batch statistics of THIS rank in training mode, as the
reference: no SyncBN) through bvae_bn_forward / bvae_bn_backward (graph/_smallnet.py).

Replicated as written: OnOffFeature slices the channel axis (``x[:, :-1]`` on a one-channel tensor, :84) so its "shifted"
roll is all zeros and the feature is the per-step sum over pitches; ConvModule owns a ``bn1`` even when ``isBasic`` never
uses it (:113); ``weights_init`` draws BatchNorm weights from N(-1, 1) too (graph/weights_initializer.py:12-17)."""
import torch
import torch.nn as nn

from ._smallnet import batch_norm, conv
from .weights_initializer import weights_init


def _bn(c, momentum=0.01):
    return nn.BatchNorm2d(c, eps=1e-5, momentum=momentum, affine=True)


class ChordFeature(nn.Module):
    def __init__(self):
        super().__init__()
        self.chord_conv1 = nn.Conv2d(1, 8, kernel_size=(3, 1), stride=(2, 1), padding=(1, 0), bias=False)
        self.chord_conv2 = nn.Conv2d(8, 16, kernel_size=(3, 1), stride=(2, 1), padding=(1, 0), bias=False)
        self.chord_fit = nn.Conv2d(16, 16, kernel_size=1, stride=1, bias=False)
        self.chord_conv3 = nn.Conv2d(16, 32, kernel_size=3, stride=2, padding=1, bias=False)
        self.chord_conv4 = nn.Conv2d(32, 64, kernel_size=3, stride=2, padding=1, bias=False)
        self.batch_norm1, self.batch_norm2, self.batch_norm3 = _bn(8), _bn(16), _bn(16)
        self.batch_norm4, self.batch_norm5 = _bn(32), _bn(64)
        self.apply(weights_init)

    def forward(self, x):                       # x: [B,192,60,1] NHWC
        B = x.shape[0]
        c = x.reshape(B, 192, 12, 5).sum(3).unsqueeze(-1)                  # pitch classes over the 5 octaves (:30-32)
        for cv, bn in ((self.chord_conv1, self.batch_norm1), (self.chord_conv2, self.batch_norm2),
                       (self.chord_fit, self.batch_norm3), (self.chord_conv3, self.batch_norm4),
                       (self.chord_conv4, self.batch_norm5)):
            c = batch_norm(conv(c, cv), bn, act=True)
        return c.float().mean((1, 2))                                      # AvgPool2d((12, 3)) == the whole 12x3 map


class OnOffFeature(nn.Module):
    def __init__(self):
        super().__init__()
        self.onoff_conv1 = nn.Conv2d(1, 8, kernel_size=3, stride=(2, 1), padding=1, bias=False)
        self.onoff_conv2 = nn.Conv2d(8, 8, kernel_size=3, stride=(2, 1), padding=1, bias=False)
        self.batch_norm2 = nn.BatchNorm2d(8)
        self.onoff_conv3 = nn.Conv2d(8, 16, kernel_size=3, stride=(2, 1), padding=1, bias=False)
        self.onoff_conv4 = nn.Conv2d(16, 32, kernel_size=3, stride=(2, 1), padding=1, bias=False)
        self.onoff_fit = nn.Conv2d(32, 32, kernel_size=1, stride=1, bias=False)
        self.onoff_conv5 = nn.Conv2d(32, 64, kernel_size=3, stride=(2, 1), padding=1, bias=False)
        self.apply(weights_init)

    def forward(self, x):
        o = x.sum(2, keepdim=True)                                         # [B,192,1,1] (see the module docstring)
        o = conv(o, self.onoff_conv1, act=True, out_f32=False)
        o = conv(o, self.onoff_conv2, act=True, out_f32=False)
        o = batch_norm(o, self.batch_norm2)
        for cv in (self.onoff_conv3, self.onoff_conv4, self.onoff_fit, self.onoff_conv5):
            o = conv(o, cv, act=True, out_f32=False)
        return o.float().mean((1, 2))                                      # AvgPool2d((6, 1))


class ConvModule(nn.Module):
    def __init__(self, in_channel, out_channel, isBasic=True):
        super().__init__()
        if not isBasic:
            self.conv1 = nn.Conv2d(in_channel, in_channel, kernel_size=3, stride=1, padding=1, bias=False)
        self.conv2 = nn.Conv2d(in_channel, out_channel, kernel_size=3, stride=2, padding=1, bias=False)
        self.bn1, self.bn2 = _bn(in_channel), _bn(out_channel)
        self.isBasic = isBasic
        self.apply(weights_init)

    def forward(self, x):
        if not self.isBasic:
            x = batch_norm(conv(x, self.conv1), self.bn1, act=True)
        return batch_norm(conv(x, self.conv2), self.bn2, act=True)


class BasicFeature(nn.Module):
    def __init__(self, layers):
        super().__init__()
        self.pitch1 = nn.Conv2d(1, 8, kernel_size=(1, 4), stride=(1, 2), padding=[0, 1], bias=False)
        self.pitch2 = nn.Conv2d(8, 8, kernel_size=(4, 1), stride=(2, 1), padding=[1, 0], bias=False)
        self.time1 = nn.Conv2d(1, 8, kernel_size=(4, 1), stride=(2, 1), padding=[1, 0], bias=False)
        self.time2 = nn.Conv2d(8, 8, kernel_size=(1, 4), stride=(1, 2), padding=[0, 1], bias=False)
        self.fit = nn.Conv2d(16, 8, kernel_size=1, stride=1, bias=False)
        self.bn = _bn(8)
        self.layers = nn.ModuleList([ConvModule(layers[i - 1], layers[i], False if i < 3 else True)
                                     for i in range(1, len(layers))])
        self.apply(weights_init)

    def forward(self, x):
        pitch = conv(conv(x, self.pitch1, act=True, out_f32=False), self.pitch2, act=True, out_f32=False)
        time = conv(conv(x, self.time1, act=True, out_f32=False), self.time2, act=True, out_f32=False)
        out = batch_norm(conv(torch.cat((pitch, time), 3), self.fit), self.bn, act=True)
        for layer in self.layers:
            out = layer(out)
        return out.float().mean((1, 2))                                    # AvgPool2d((12, 4))


class BarDiscriminator(nn.Module):
    def __init__(self):
        super().__init__()
        self.chord = ChordFeature()
        self.onoff = OnOffFeature()
        self.basic = BasicFeature([8, 16, 32, 64])
        self.linear = nn.Linear(64 * 3, 1, bias=False)
        self.apply(weights_init)

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("the B200 path needs CUDA tensors; there is no CPU fallback")
        x = x.reshape(-1, 192, 60, 1)                                      # [B,1,192,60] NCHW == NHWC for one channel
        feat = torch.cat((self.chord(x), self.onoff(x), self.basic(x)), 1)
        return torch.sigmoid(self.linear(feat))                            # 192 MACs per sample: glue, like the z heads
