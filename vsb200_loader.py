"""Imports the package directory `hai-25-rag-on-edge_b200/` (not a valid Python identifier) under the name `vsb200`."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hai-25-rag-on-edge_b200")


def load():
    if "vsb200" in sys.modules:
        return sys.modules["vsb200"]
    spec = importlib.util.spec_from_file_location("vsb200", os.path.join(_PKG_DIR, "__init__.py"),
                                                  submodule_search_locations=[_PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["vsb200"] = mod
    spec.loader.exec_module(mod)
    return mod
