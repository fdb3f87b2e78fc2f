"""TEST INFRASTRUCTURE — ctypes binding of oracle/_build/libmlbsim_oracle.so (the CPU restatement).

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs.  Builds the library on demand with gcc (oracle/Makefile).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# MLBSIM_ORACLE_LIB: an experiment build of the twin (e.g. another NATIVE_PHILOX_ROUNDS), paired with MLBSIM_LIB
LIB_PATH = os.environ.get("MLBSIM_ORACLE_LIB") or os.path.join(HERE, "_build", "libmlbsim_oracle.so")

_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "mlbsim_oracle.c")
    hdr = os.path.join(HERE, "..", "include", "mlbsim.h")
    stale = (not os.path.exists(LIB_PATH)) or any(
        os.path.exists(p) and os.path.getmtime(p) > os.path.getmtime(LIB_PATH) for p in (src, hdr)
    )
    if os.environ.get("MLBSIM_ORACLE_LIB"):
        return LIB_PATH          # experiment builds are made by hand (make -C oracle OUT=... EXTRA=...)
    if force or stale:
        subprocess.run(["make", "-C", HERE, "-B"], check=True, capture_output=True)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(LIB_PATH)
        vp, u64, i64, u32 = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int64, ctypes.c_uint32
        L.oracle_replay.argtypes = [vp, vp, vp, i64, vp]
        L.oracle_replay.restype = ctypes.c_int
        L.oracle_grammar.argtypes = [vp, u64, i64, i64, vp, vp, vp, i64, vp]
        L.oracle_grammar.restype = ctypes.c_int
        L.oracle_native.argtypes = [vp, u32, u64, u64, u64, vp, u32]
        L.oracle_native.restype = ctypes.c_int
        L.oracle_native_persim.argtypes = [vp, u32, u64, u64, u64, vp]
        L.oracle_native_persim.restype = ctypes.c_int
        L.oracle_philox4x32_10.argtypes = [vp, vp, vp]
        L.oracle_philox4x32_10.restype = None
        L.oracle_philox2x32_10.argtypes = [vp, u32, vp]
        L.oracle_philox2x32_10.restype = None
        L.oracle_native_key.argtypes = [u64]
        L.oracle_native_key.restype = u32
        L.oracle_set_threads.argtypes = [ctypes.c_int]
        L.oracle_get_threads.restype = ctypes.c_int
        for f in ("params", "table", "hist", "replay_game", "summary"):
            getattr(L, f"oracle_sizeof_{f}").restype = ctypes.c_size_t
        _lib = L
        from mlb_odds_engine_b200 import _abi

        _abi.check_sizes(lambda w: [L.oracle_sizeof_params, L.oracle_sizeof_table, L.oracle_sizeof_hist,
                                    L.oracle_sizeof_replay_game, L.oracle_sizeof_summary][w]())
    return _lib


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def set_threads(n: int) -> None:
    lib().oracle_set_threads(int(n))


def get_threads() -> int:
    return int(lib().oracle_get_threads())


def philox4x32_10(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    lib().oracle_philox4x32_10(_ptr(c), _ptr(k), _ptr(out))
    return out


def philox2x32_10(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32)
    out = np.zeros(2, dtype=np.uint32)
    lib().oracle_philox2x32_10(_ptr(c), int(key), _ptr(out))
    return out


def native_key(seed: int) -> int:
    return int(lib().oracle_native_key(int(seed)))


def pack_tapes(tapes):
    """list of float64 arrays -> (flat tape, uint64 offsets[n+1])."""
    offsets = np.zeros(len(tapes) + 1, dtype=np.uint64)
    if len(tapes):
        offsets[1:] = np.cumsum([len(t) for t in tapes])
    flat = np.concatenate(tapes).astype(np.float64) if len(tapes) else np.zeros(0)
    return np.ascontiguousarray(flat), offsets


def replay(params: np.ndarray, tape: np.ndarray, offsets: np.ndarray) -> np.ndarray:
    from mlb_odds_engine_b200 import _abi

    n = len(offsets) - 1
    out = np.zeros(n, dtype=_abi.REPLAY_GAME_DTYPE)
    p = np.ascontiguousarray(params)
    rc = lib().oracle_replay(_ptr(p), _ptr(tape), _ptr(offsets), n, _ptr(out))
    assert rc == 0
    return out


def grammar(params: np.ndarray, seed: int, game_lo: int, game_hi: int, want_games=False, tape_stride=0):
    """Reference-grammar games with the CPU RNG.  Returns (hist, games|None, tapes|None, tape_len|None)."""
    from mlb_odds_engine_b200 import _abi

    n = game_hi - game_lo
    hist = np.zeros((), dtype=_abi.HIST_DTYPE)
    games = np.zeros(n, dtype=_abi.REPLAY_GAME_DTYPE) if want_games else None
    tape = np.zeros((n, tape_stride), dtype=np.float64) if tape_stride else None
    tlen = np.zeros(n, dtype=np.int64) if tape_stride else None
    p = np.ascontiguousarray(params)
    rc = lib().oracle_grammar(_ptr(p), seed, game_lo, game_hi, _ptr(hist),
                              _ptr(games) if want_games else None,
                              _ptr(tape) if tape_stride else None, tape_stride,
                              _ptr(tlen) if tape_stride else None)
    if rc != 0:
        raise RuntimeError("oracle_grammar: tape_stride too small for at least one game")
    return hist, games, tape, tlen


def native(table: np.ndarray, matchup: int, sim_lo: int, sim_hi: int, seed: int, flags: int = 0) -> np.ndarray:
    from mlb_odds_engine_b200 import _abi

    hist = np.zeros((), dtype=_abi.HIST_DTYPE)
    t = np.ascontiguousarray(table)
    rc = lib().oracle_native(_ptr(t), matchup, sim_lo, sim_hi, seed, _ptr(hist), flags)
    if rc != 0:
        raise RuntimeError(f"oracle_native failed with code {rc}")
    return hist


def native_persim(table: np.ndarray, matchup: int, sim_lo: int, sim_hi: int, seed: int) -> np.ndarray:
    from mlb_odds_engine_b200 import _abi

    out = np.zeros(sim_hi - sim_lo, dtype=_abi.REPLAY_GAME_DTYPE)
    t = np.ascontiguousarray(table)
    rc = lib().oracle_native_persim(_ptr(t), matchup, sim_lo, sim_hi, seed, _ptr(out))
    if rc != 0:
        raise RuntimeError(f"oracle_native_persim failed with code {rc}")
    return out
