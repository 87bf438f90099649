"""Import shim: the product package lives in ``differentialriccatiequations.jl_b200/`` (the
directory name contains a dot, so it cannot be imported by name).  ``import dre_b200`` loads that
directory as the package ``dre_b200``."""
import importlib.util
import os
import sys

_pkg_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "differentialriccatiequations.jl_b200")
_spec = importlib.util.spec_from_file_location(
    "dre_b200", os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["dre_b200"] = _mod
_spec.loader.exec_module(_mod)
