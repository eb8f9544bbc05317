"""Import alias for the package directory ``particlemdi.jl_b200/``.

The directory name mandated for this project contains a dot, which the ``import`` statement
cannot spell; this shim loads it under the importable name ``pmdi_b200`` (sub-modules resolve
as ``pmdi_b200.capi``, ``pmdi_b200.pmdi`` ...).
"""
import importlib.util
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_pkg_dir = os.path.join(_here, "particlemdi.jl_b200")
_spec = importlib.util.spec_from_file_location(
    __name__, os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
