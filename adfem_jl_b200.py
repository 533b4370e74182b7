"""Import shim: the package directory is named `adfem.jl_b200/` (not an importable identifier), so
`import adfem_jl_b200` loads it from there and replaces this module with the real package."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "adfem.jl_b200")
_spec = importlib.util.spec_from_file_location("adfem_jl_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["adfem_jl_b200"] = _mod
_spec.loader.exec_module(_mod)
