"""Import shim: `import enzo_e_b200` loads the package in ./enzo-e_b200/.

The package directory keeps the project's name (with its hyphen), which is not
a valid Python identifier; this one-file module makes it importable under the
name `enzo_e_b200` (and its submodules as `enzo_e_b200.abi`, ...).
"""
import importlib.util
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_pkg_dir = os.path.join(_here, "enzo-e_b200")
_spec = importlib.util.spec_from_file_location(
    "enzo_e_b200", os.path.join(_pkg_dir, "__init__.py"),
    submodule_search_locations=[_pkg_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["enzo_e_b200"] = _mod
_spec.loader.exec_module(_mod)
