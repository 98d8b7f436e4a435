"""Import shim: the package directory is named ``drone-collision-rp_b200`` (not a
valid Python identifier), so ``import drone_collision_rp_b200`` loads it from
there."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_here = _os.path.dirname(_os.path.abspath(__file__))
_pkg_dir = _os.path.join(_here, "drone-collision-rp_b200")
_spec = _ilu.spec_from_file_location(
    __name__, _os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir]
)
_mod = _ilu.module_from_spec(_spec)
_sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
