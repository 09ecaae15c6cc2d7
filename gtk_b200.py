"""Import shim: the package directory is named ``galerkintoolkit.jl_b200`` (not a valid Python
identifier), so load it by path and expose it as ``gtk_b200``."""
import importlib.util
import os
import sys

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "galerkintoolkit.jl_b200")
_NAME = "galerkintoolkit_jl_b200"
if _NAME not in sys.modules:
    _spec = importlib.util.spec_from_file_location(_NAME, os.path.join(_DIR, "__init__.py"),
                                                   submodule_search_locations=[_DIR])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_NAME] = _mod
    _spec.loader.exec_module(_mod)
_pkg = sys.modules[_NAME]
hostprep = _pkg.hostprep
engine = _pkg.engine
gt = GT = _pkg.gt
build_library = _pkg.build_library
PACKAGE_DIR = _DIR
