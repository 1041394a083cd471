"""Import alias: `import tnqs_b200` loads the package that lives in the directory
`tensornetworkquantumsimulator.jl_b200/` (a name Python cannot import directly because of the dot)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tensornetworkquantumsimulator.jl_b200")
_spec = importlib.util.spec_from_file_location(
    "tnqs_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["tnqs_b200"] = _mod
_spec.loader.exec_module(_mod)
