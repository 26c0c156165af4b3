"""Import shim: the package directory is `sci-solver_fem_b200/` (hyphenated, fixed by the project
layout) which Python cannot import by name; this module loads it under `sci_solver_fem_b200`."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sci-solver_fem_b200")
_spec = importlib.util.spec_from_file_location("sci_solver_fem_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["sci_solver_fem_b200"] = _mod
_spec.loader.exec_module(_mod)
