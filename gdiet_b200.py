"""Import shim: ``import gdiet_b200`` loads the package in ./genome-on-diet_b200 (hyphenated dir)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "genome-on-diet_b200")
_spec = importlib.util.spec_from_file_location("gdiet_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["gdiet_b200"] = _mod
_spec.loader.exec_module(_mod)
