"""Initialiser with the semantics of the reference's graph/weights_initializer.py:5-23.

Matching is by class-name substring, exactly as the reference does, which has three consequences that are kept
on purpose: 'Conv2' matches nn.Conv2d but not nn.ConvTranspose2d; 'BatchNorm' never matches
nn.InstanceNorm2d; the "bias" branch re-draws the *weight*, so biases keep PyTorch's defaults.
"""
_MATCH = ("Conv2", "BatchNorm", "Linear")


def weights_init(m):
    name = type(m).__name__
    if any(tag in name for tag in _MATCH) and getattr(m, "weight", None) is not None:
        draws = 2 if getattr(m, "bias", None) is not None else 1
        for _ in range(draws):          # the second draw replaces the first (reference :10-11,16-17,22-23)
            m.weight.data.normal_(-1.0, 1.0)
