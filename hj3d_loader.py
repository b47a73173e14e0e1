"""Import the package directory ``3d-hashjoin_b200/`` (not a valid Python identifier) as ``hj3d_b200``."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(ROOT, "3d-hashjoin_b200")


def load():
    if "hj3d_b200" in sys.modules:
        return sys.modules["hj3d_b200"]
    spec = importlib.util.spec_from_file_location("hj3d_b200", os.path.join(PKG_DIR, "__init__.py"),
                                                  submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["hj3d_b200"] = mod
    spec.loader.exec_module(mod)
    return mod
