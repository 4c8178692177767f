"""Import alias: `import poppunk_mod_b200` loads the package in ./poppunk-mod_b200/ (a hyphen
cannot appear in an import statement).  The module object is replaced by the real package."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "poppunk-mod_b200")
_spec = _ilu.spec_from_file_location(__name__, _os.path.join(_dir, "__init__.py"),
                                     submodule_search_locations=[_dir])
_pkg = _ilu.module_from_spec(_spec)
_sys.modules[__name__] = _pkg
_spec.loader.exec_module(_pkg)
