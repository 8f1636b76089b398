"""B200-native bar-VAE train / decode hot path.

The package directory name contains a hyphen, so import it with
``importlib.import_module("musicgeneration_vae-torch_b200")`` or through the ``barvae_b200`` shim at the repo root.
``install_dropin()`` registers the reference's top-level names (``graph``, ``agent``, ``config``) in ``sys.modules``
so that ``from graph.model import Model`` resolves to this implementation.
"""
import importlib
import sys

from . import _lib, engine  # noqa: F401
from ._lib import launch_count, reset_launch_count  # noqa: F401


def install_dropin():
    pkg = __name__
    for name in ("graph", "graph.model", "graph.encoder", "graph.decoder", "graph.phrase_encoder", "graph.cbam",
                 "graph.encodingBlock", "graph.weights_initializer", "graph.loss", "graph.loss.bar_loss",
                 "graph.model_with_gan", "graph.z_discriminator", "graph.bar_discriminator_with_feature",
                 "graph.bar_discriminator", "graph.refiner",
                 "data", "data.bar_dataset", "config", "agent", "agent.barGen", "agent.barGen_with_gan", "maker_bar", "main"):
        try:
            sys.modules[name] = importlib.import_module(pkg + "." + name)
        except ImportError:
            pass
