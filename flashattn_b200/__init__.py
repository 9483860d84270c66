"""Import shim: ``flashattention-from-scratch-with-triton_b200/`` is the package, but a hyphenated
directory name cannot be written in an ``import`` statement.  Importing ``flashattn_b200``
loads that directory as this module (sub-modules included)."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_pkg_dir = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                         "flashattention-from-scratch-with-triton_b200")
_spec = _ilu.spec_from_file_location(__name__, _os.path.join(_pkg_dir, "__init__.py"),
                                     submodule_search_locations=[_pkg_dir])
_mod = _ilu.module_from_spec(_spec)
_sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
