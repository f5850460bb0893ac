"""Import the UNMODIFIED reference (``/root/reference``) as the package ``frldistml.scaffold``.

TEST INFRASTRUCTURE.  Works only where ``/root/reference`` exists (the build container, not
the GPU box); used by ``oracle/make_golden.py`` to generate the committed fixtures and by the
``-m "not gpu"`` tests that pin ``oracle/ref_loop.py`` against the live reference.

What the shim does, without touching the reference tree:
  * a temp dir with ``frldistml/__init__.py`` and a symlink ``frldistml/scaffold -> /root/reference``
    (the reference uses relative imports and its tests use that package name);
  * ``sys.modules`` stubs for two absent visualisation deps: ``plotly.graph_objs`` (types.py:20)
    and ``nbformat`` (local_solver.py:16);
  * ``TORCH_FORCE_NO_WEIGHTS_ONLY_LOAD=1`` for the whole-module ``torch.load`` at solver.py:604.
"""
import os
import sys
import tempfile
import types

REFERENCE_DIR = "/root/reference"
_state = {}


def reference_available() -> bool:
    return os.path.isdir(REFERENCE_DIR) and os.path.exists(os.path.join(REFERENCE_DIR, "solver.py"))


def _stub_module(name: str, **attrs) -> types.ModuleType:
    mod = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(mod, k, v)
    sys.modules[name] = mod
    return mod


def import_reference():
    """Returns the imported ``frldistml.scaffold`` package (cached)."""
    if "pkg" in _state:
        return _state["pkg"]
    if not reference_available():
        raise RuntimeError("/root/reference is not present on this machine")
    os.environ["TORCH_FORCE_NO_WEIGHTS_ONLY_LOAD"] = "1"

    if "plotly" not in sys.modules:
        class Figure(dict):
            pass
        plotly = _stub_module("plotly")
        plotly.graph_objs = _stub_module("plotly.graph_objs", Figure=Figure)
    if "nbformat" not in sys.modules:
        nb = _stub_module("nbformat", write=lambda *a, **k: None)
        nb.notebooknode = _stub_module("nbformat.notebooknode", NotebookNode=dict)
        nb.v4 = _stub_module("nbformat.v4", new_markdown_cell=lambda s: {"md": s},
                             new_code_cell=lambda s: {"code": s},
                             new_notebook=lambda cells=None: {"cells": cells})
    if "mock" not in sys.modules:
        import unittest.mock
        sys.modules["mock"] = unittest.mock

    # an frl_b200.install_reference_alias() made earlier in this process must not shadow the
    # real reference
    for k in [k for k in sys.modules if k == "frldistml" or k.startswith("frldistml.")]:
        del sys.modules[k]
    root = tempfile.mkdtemp(prefix="frl_ref_shim_")
    os.makedirs(os.path.join(root, "frldistml"))
    open(os.path.join(root, "frldistml", "__init__.py"), "w").close()
    os.symlink(REFERENCE_DIR, os.path.join(root, "frldistml", "scaffold"))
    sys.path.insert(0, root)
    sys.dont_write_bytecode = True         # /root/reference is read-only
    import importlib
    pkg = importlib.import_module("frldistml.scaffold")
    for name in ("types", "criteria", "model", "lr_scheduler", "sampler", "transform", "task",
                 "problem", "multitask_problem", "solver_worker", "solver", "local_solver"):
        importlib.import_module("frldistml.scaffold." + name)
    _state["pkg"] = pkg
    return pkg
