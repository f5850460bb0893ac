"""The C-ABI library loads and exports every symbol include/frl_b200.h declares (no compute)."""
import ctypes
import os
import re

import frl_b200  # noqa: F401
from frl_b200 import _native

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    text = open(os.path.join(REPO, "include", "frl_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(frl_[a-z0-9_]+)\s*\(", text)))


def test_library_is_built_in_tree():
    assert os.path.exists(_native.LIB_PATH), "run `python __graft_entry__.py build` first"


def test_every_declared_symbol_is_exported():
    handle = ctypes.CDLL(_native.LIB_PATH)
    names = _header_functions()
    assert len(names) >= 14
    for name in names:
        assert hasattr(handle, name), f"{name} declared in the header but not exported"


def test_binding_covers_the_header():
    assert sorted(_native.SIGNATURES) == _header_functions()


def test_abi_version_and_struct_layout():
    lib = _native.lib()
    assert lib.frl_abi_version() == 1
    # frl_task_desc: 4 x int32, 4 x pointer, 3 x int64, 2 x float
    assert ctypes.sizeof(_native.TaskDesc) == 16 + 32 + 24 + 8
    assert lib.frl_criteria_scratch_bytes(2) > 0
    assert lib.frl_reduce_scratch_bytes() > 0


def test_argument_errors_are_reported_without_a_gpu():
    lib = _native.lib()
    rc = lib.frl_sgd_momentum(None, None, None, None, 16, 0.1, 0.9, 0.0, 0.0, 1.0, None, None, 0, 0, None)
    assert rc < 0
    assert b"frl_sgd_momentum" in lib.frl_last_error()
    rc = lib.frl_adam(None, None, None, None, None, None, 16, 0.1, 0.9, 0.999, 1e-8, 0.0, 0, 1.0, None, None, 0, None)
    assert rc < 0
