"""Import shim: ``import barvae_b200 as bv`` gives the package whose directory name
(``musicgeneration_vae-torch_b200``) is not a Python identifier."""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
_pkg = importlib.import_module("musicgeneration_vae-torch_b200")
sys.modules[__name__] = _pkg
