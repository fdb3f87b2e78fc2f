"""CPU: the C-ABI library loads without a GPU and exports exactly what include/mlbsim.h declares."""
import ctypes
import os
import re
import subprocess

import pytest

from mlb_odds_engine_b200 import _abi, _lib
from mlb_odds_engine_b200 import build as b

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mlbsim.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mlbsim_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    b.build_library()
    return _lib.load()


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(_lib.SYMBOLS)


def test_library_exports_every_declared_symbol(lib):
    for name in declared_symbols():
        assert hasattr(lib, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\bT (mlbsim_[a-z0-9_]+)\b", out))
    assert set(declared_symbols()) <= exported


def test_struct_sizes_match_numpy_mirrors(lib):
    for which, dt in _abi.STRUCTS.items():
        assert lib.mlbsim_struct_size(which) == dt.itemsize
    assert lib.mlbsim_struct_size(99) == 0
    assert lib.mlbsim_abi_version() == _abi.ABI_VERSION
    assert _abi.HIST_WORDS * 8 == _abi.HIST_DTYPE.itemsize


def test_library_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_device_is_a_loud_error(lib):
    """Without a GPU mlbsim_init must fail with NO_DEVICE — there is no CPU path to fall back to."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = ctypes.c_void_p()
    rc = lib.mlbsim_init(0, ctypes.byref(h))
    assert rc == 4 and not h.value
    assert b"no CUDA device" in lib.mlbsim_last_error(None)
    with pytest.raises(_lib.MlbsimError):
        _lib.Context(0)
    # every entry point rejects a NULL context instead of crashing
    assert lib.mlbsim_synchronize(None) == 2
    assert lib.mlbsim_upload_tables(None, None, 0) == 2
    assert lib.mlbsim_launch_count(None) == 0
    lib.mlbsim_destroy(None)
