"""Import alias: ``import dirb200`` -> the package directory ``deep-image-retrieval_b200/``.

The package directory name is fixed by the project layout and is not a valid
Python identifier, so this one-file shim registers it under the importable name
``dirb200`` (sub-modules resolve through ``__path__`` as usual:
``import dirb200.nets``, ``from dirb200 import lib`` ...).
"""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "deep-image-retrieval_b200")
_spec = importlib.util.spec_from_file_location(
    "dirb200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["dirb200"] = _mod
_spec.loader.exec_module(_mod)
