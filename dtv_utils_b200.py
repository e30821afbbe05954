"""Import shim: the package directory is `dtv-utils_b200/` (the reference's name has a hyphen, which
Python cannot import); `import dtv_utils_b200` loads it under an importable name."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dtv-utils_b200")
_spec = importlib.util.spec_from_file_location("dtv_utils_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["dtv_utils_b200"] = _mod
_spec.loader.exec_module(_mod)
