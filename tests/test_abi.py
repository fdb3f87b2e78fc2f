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


def test_replay_kernel_dispatch_rule(lib):
    """Host-only: which replay-kernel instance a parameter set selects (include/mlbsim.h, mlbsim_replay_kernel_kind).  The
    specialised instance has no draw-by-draw path and no rates past 138 pitches, so the rule must send everything that
    could need either to the general one."""
    from helpers import REPLAY_SCENARIOS, load_replay
    from mlb_odds_engine_b200 import synth, tables

    kinds = {}
    for name in REPLAY_SCENARIOS:
        g = load_replay(name)
        kinds[name] = _lib.replay_kernel_kind(tables.build_params(g["matchup"], use_noise=g["use_noise"]))
    assert kinds["bench_texhou"] == "regular" and kinds["hetero_bullpen_noise"] == "regular"
    assert kinds["hetero_bullpen_nonoise"] == "nonoise" and kinds["nonoise_nobullpen"] == "nonoise"
    assert kinds["hetero_nobullpen_noise"] == "general"      # a starter without a bullpen pitches past the tabulated range
    m = synth.make_matchup("2025-04-04-TEX@HOU", n_relievers=(6, 4))
    p = tables.build_params(m, use_noise=True)
    assert _lib.replay_kernel_kind(p) == "regular"
    for field, value in (("pit_k", 0.0), ("pit_bb", 2.5), ("bat_k", 2.5), ("bat_bb", -1.0), ("k_mod", 0.0), ("bb_mod", float("nan"))):
        q = p.copy()
        if q[field].ndim:
            q[field][..., 0, 0] = value    # (side 0, first pitcher / batter)
            if field.startswith("bat"):
                q["pit_" + field[4:]][..., 0, :] = value   # the average of batter and pitcher must leave (0, 1)
            else:
                q["bat_" + field[4:]][..., 0, :] = value
        else:
            q[field] = value
        assert _lib.replay_kernel_kind(q) == "general", field
    q = p.copy()
    q["bullpen_len"][..., 1] = 0
    assert _lib.replay_kernel_kind(q) == "general"
    # a rate that only leaves (0, 1) on the fatigue ramp (walk rate x 1.0756 at 138 pitches)
    q = p.copy()
    q["pit_bb"][..., 1, :] = 0.96
    q["bat_bb"][..., 1, :] = 0.97
    assert _lib.replay_kernel_kind(q) == "general"
    q = p.copy()
    q["lineup_len"][..., 0] = 0
    with pytest.raises(_lib.MlbsimError):
        _lib.replay_kernel_kind(q)
