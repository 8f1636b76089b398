"""Entry point with the shape of the reference's main.py:7-12 (``agent = BarGen(Config()); agent.run()``).

    python -m musicgeneration_vae-torch_b200.main                       # one GPU
    torchrun --nproc-per-node 8 -m musicgeneration_vae-torch_b200.main  # one process per GPU, NCCL gradient all-reduce

``--agent gan`` runs the adversarial trainer (agent/barGen_with_gan.py; ``--gan-schedule horovod`` for the per-iteration
schedule of agent/barGen_horovod.py) instead of the generator-only one (agent/barGen.py).  Attributes of ``Config`` can be
overridden as ``--set name=value`` (the reference has no command line: config.py:1-20 is edited by hand)."""
import argparse
import ast

from .config import Config


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--agent", default="generator", choices=["generator", "gan"])
    ap.add_argument("--gan-schedule", default=None, choices=["with_gan", "horovod"])
    ap.add_argument("--set", action="append", default=[], metavar="NAME=VALUE", help="override a Config attribute")
    args = ap.parse_args(argv)
    config = Config()
    for item in args.set:
        name, _, value = item.partition("=")
        if not hasattr(config, name):
            raise SystemExit("Config has no attribute %r" % name)
        try:
            value = ast.literal_eval(value)
        except (ValueError, SyntaxError):
            pass
        setattr(config, name, value)
    if args.gan_schedule:
        config.gan_schedule = args.gan_schedule
    if args.agent == "gan":
        from .agent.barGen_with_gan import BarGen
    else:
        from .agent.barGen import BarGen
    agent = BarGen(config)
    agent.run()


if __name__ == "__main__":
    main()
