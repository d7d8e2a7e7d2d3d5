"""Import shim: the package directory is `linear-programming_b200/` (the name the build contract
fixes), which is not a valid Python identifier.  Importing this module loads that directory as
the package `linear_programming_b200` (submodules included) from the repo root."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "linear-programming_b200")
_spec = _ilu.spec_from_file_location(__name__, _os.path.join(_dir, "__init__.py"),
                                     submodule_search_locations=[_dir])
_mod = _ilu.module_from_spec(_spec)
_sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
