"""Refiner (reference: graph/refiner.py:7-58): a small convolutional auto-encoder that post-processes the generated bar,
``(x + refine(x)) / 2``.  graph/model.py:31,41 calls it in both branches, but the reference's module cannot execute:
``layer1`` produces 2 channels and ``layer2``'s convolution is declared ``Conv2d(1, 8, ...)`` (:12 vs :19) while the
comments next to both layers (:16, :23) give the intended shapes.  This module carries the one-line fix
(``Conv2d(2, 8, ...)``) and is therefore OFF by default: ``graph.model.Model(refiner=True)`` enables it (SURVEY.md
section 8f N1).  Parameter / buffer names are the reference's (``layer1.0.weight`` ... ``layer6.1.running_var``); a reference
checkpoint loads except for ``layer2.0.weight``, whose reference shape [8,1,4,4] is the defect itself.

BatchNorm2d here is batch-coupled (statistics over the local batch in training mode, per rank -- no SyncBN in the
reference either); in eval mode (sampling, maker_bar.py) it uses the running statistics."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from ._smallnet import batch_norm, conv, linear
from .weights_initializer import weights_init


class Refiner(nn.Module):
    def __init__(self):
        super().__init__()
        self.layer1 = nn.Sequential(nn.Conv2d(1, 2, kernel_size=4, padding=2), nn.BatchNorm2d(2), nn.LeakyReLU(.2),
                                    nn.MaxPool2d(kernel_size=2))            # [96, 60] -> [48, 30]
        self.layer2 = nn.Sequential(nn.Conv2d(2, 8, kernel_size=4, padding=2), nn.BatchNorm2d(8), nn.LeakyReLU(.2),
                                    nn.MaxPool2d(kernel_size=2))            # [48, 30] -> [24, 15]   (the fix: 2 input channels)
        self.layer3 = nn.Sequential(nn.Linear(2880, 1024), nn.ReLU())
        self.layer4 = nn.Sequential(nn.Linear(1024, 2880), nn.ReLU())
        self.layer5 = nn.Sequential(nn.ConvTranspose2d(8, 2, kernel_size=4, stride=2, bias=False, padding=1),
                                    nn.BatchNorm2d(2), nn.ReLU())
        self.layer6 = nn.Sequential(nn.ConvTranspose2d(2, 1, kernel_size=4, stride=2, bias=False, padding=1),
                                    nn.BatchNorm2d(1), nn.Sigmoid())
        self.apply(weights_init)

    @staticmethod
    def _pool(t):                                # MaxPool2d(2) on an NHWC tensor (glue: <= 8 channels)
        return F.max_pool2d(t.permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1)

    def forward(self, x):                        # x [B,1,96,60] fp32 in (0,1)
        if not x.is_cuda:
            raise RuntimeError("the B200 path needs CUDA tensors; there is no CPU fallback")
        B = x.shape[0]
        xh = x.reshape(B, 96, 60, 1)
        bg = not self.layer1[1].training        # conv biases in front of a batch-statistics BatchNorm: zero gradient
        x_2 = self._pool(batch_norm(conv(xh, self.layer1[0], bias_grad=bg), self.layer1[1], act=True, slope=0.2))   # [B,48,30,2]
        x_8 = self._pool(batch_norm(conv(x_2, self.layer2[0], bias_grad=bg), self.layer2[1], act=True, slope=0.2))  # [B,24,15,8]
        flat = x_8.permute(0, 3, 1, 2).reshape(B, 2880)         # the reference flattens NCHW (:52)
        f = linear(linear(flat, self.layer3[0], act=True), self.layer4[0], act=True)
        x_8_t = x_8.float() + f.float().view(B, 8, 24, 15).permute(0, 2, 3, 1)
        x_2_t = x_2.float() + batch_norm(conv(x_8_t, self.layer5[0]), self.layer5[1], act=True).float()
        y = torch.sigmoid(batch_norm(conv(x_2_t, self.layer6[0]), self.layer6[1]).float())                # [B,96,60,1]
        return (x + y.reshape(B, 1, 96, 60)) * 0.5
