"""Import shim: the package directory is named ``svs-unet-pytorch_b200`` (a hyphen is not a
valid Python identifier), so ``import svs_unet_pytorch_b200`` loads it from that directory."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "svs-unet-pytorch_b200")
_spec = importlib.util.spec_from_file_location(
    "svs_unet_pytorch_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["svs_unet_pytorch_b200"] = _mod
_spec.loader.exec_module(_mod)
