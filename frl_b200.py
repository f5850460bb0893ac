"""Import alias for the hyphenated package directory.

The package lives in ``frl-distributed-ml-scaffold_b200/`` (the directory name the
build contract fixes); a hyphen is not importable, so ``import frl_b200`` loads that
directory as the package ``frl_b200`` and replaces this stub in ``sys.modules``.
"""
import importlib.util as _ilu
import os as _os
import sys as _sys

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)),
                         "frl-distributed-ml-scaffold_b200")
_spec = _ilu.spec_from_file_location(
    "frl_b200", _os.path.join(_PKG_DIR, "__init__.py"),
    submodule_search_locations=[_PKG_DIR])
_pkg = _ilu.module_from_spec(_spec)
_sys.modules["frl_b200"] = _pkg
_spec.loader.exec_module(_pkg)
