"""ctypes binding of libmlbsim_b200.so (include/mlbsim.h).  No fallback: if the library is not
built, or there is no sm_100 device, the call raises."""
from __future__ import annotations

import ctypes
import os

import numpy as np

from . import _abi

# MLBSIM_LIB: load a tuning variant of the same library instead (scripts/build_variants.py); never a different backend
LIB_PATH = os.environ.get("MLBSIM_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libmlbsim_b200.so")

# every symbol include/mlbsim.h declares (tests/test_abi.py checks header <-> list <-> library)
SYMBOLS = [
    "mlbsim_init", "mlbsim_destroy", "mlbsim_last_error", "mlbsim_abi_version", "mlbsim_struct_size",
    "mlbsim_device_info", "mlbsim_set_stream", "mlbsim_synchronize", "mlbsim_upload_tables",
    "mlbsim_run_native", "mlbsim_run_native_host", "mlbsim_run_replay", "mlbsim_run_replay_dev",
    "mlbsim_run_persim", "mlbsim_malloc", "mlbsim_free", "mlbsim_memset", "mlbsim_memcpy_h2d", "mlbsim_memcpy_d2h",
    "mlbsim_launch_count", "mlbsim_set_launch_config", "mlbsim_last_kernel_ms",
    "mlbsim_summarize", "mlbsim_run_native_summary_host", "mlbsim_allreduce", "mlbsim_set_matchup_base",
    "mlbsim_alias_rows", "mlbsim_build_tables", "mlbsim_download_tables", "mlbsim_replay_kernel_kind",
]

ERR_NAMES = {1: "CUDA", 2: "ARG", 3: "TABLE", 4: "NO_DEVICE", 5: "STATE", 6: "CANNOT_SCORE"}


class MlbsimError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"mlbsim error {code} ({ERR_NAMES.get(code, '?')}): {msg}")
        self.code = code


class GameDidNotTerminate(RuntimeError):
    """A simulated game was still tied after MLBSIM_MAX_INNINGS innings (include/mlbsim.h): neither side can score in
    this matchup.  The reference loops forever there (core/game_simulator.py:110-111); this build stops and reports."""


def check_terminated(rec: np.ndarray) -> None:
    """Raise if any histogram / summary record counts truncated games."""
    t = np.asarray(rec["n_truncated"]).reshape(-1)
    if t.any():
        bad = np.nonzero(t)[0]
        raise GameDidNotTerminate(
            f"{int(t.sum())} simulated games were still tied after {_abi.MAX_INNINGS} innings "
            f"(matchup index {bad[:8].tolist()}{'...' if len(bad) > 8 else ''}): a matchup in which neither side can score")


_lib = None


def load() -> ctypes.CDLL:
    """Load the CUDA library (loudly: a missing .so is an ImportError, never a CPU path)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing — build it with `python -m mlb_odds_engine_b200.build` "
            "(nvcc, sm_100a). There is no CPU fallback."
        )
    L = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, u32, u64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_uint32, ctypes.c_uint64
    sigs = {
        "mlbsim_init": (i32, [i32, ctypes.POINTER(vp)]),
        "mlbsim_destroy": (None, [vp]),
        "mlbsim_last_error": (ctypes.c_char_p, [vp]),
        "mlbsim_abi_version": (i32, []),
        "mlbsim_struct_size": (ctypes.c_size_t, [i32]),
        "mlbsim_device_info": (i32, [vp, ctypes.POINTER(i32), ctypes.POINTER(i32), ctypes.c_char_p, ctypes.c_size_t]),
        "mlbsim_set_stream": (i32, [vp, ctypes.c_size_t]),
        "mlbsim_synchronize": (i32, [vp]),
        "mlbsim_upload_tables": (i32, [vp, vp, i32]),
        "mlbsim_run_native": (i32, [vp, i32, i32, u64, u64, u64, vp, u32]),
        "mlbsim_run_native_host": (i32, [vp, i32, i32, u64, u64, u64, vp, u32]),
        "mlbsim_run_replay": (i32, [vp, vp, vp, vp, i64, vp]),
        "mlbsim_run_replay_dev": (i32, [vp, vp, vp, vp, i64, vp]),
        "mlbsim_run_persim": (i32, [vp, i32, u64, u64, u64, vp]),
        "mlbsim_malloc": (i32, [vp, ctypes.c_size_t, ctypes.POINTER(vp)]),
        "mlbsim_free": (i32, [vp, vp]),
        "mlbsim_memset": (i32, [vp, vp, i32, ctypes.c_size_t]),
        "mlbsim_memcpy_h2d": (i32, [vp, vp, vp, ctypes.c_size_t]),
        "mlbsim_memcpy_d2h": (i32, [vp, vp, vp, ctypes.c_size_t]),
        "mlbsim_launch_count": (u64, [vp]),
        "mlbsim_set_launch_config": (i32, [vp, i32, i32]),
        "mlbsim_last_kernel_ms": (i32, [vp, ctypes.POINTER(ctypes.c_float)]),
        "mlbsim_summarize": (i32, [vp, vp, i32, vp]),
        "mlbsim_run_native_summary_host": (i32, [vp, i32, i32, u64, u64, u64, vp, u32]),
        "mlbsim_allreduce": (i32, [vp, vp, vp, i32]),
        "mlbsim_set_matchup_base": (i32, [vp, i32]),
        "mlbsim_alias_rows": (i32, [vp, vp, i32, i32, vp]),
        "mlbsim_build_tables": (i32, [vp, vp, i32, u32]),
        "mlbsim_download_tables": (i32, [vp, i32, i32, vp]),
        "mlbsim_replay_kernel_kind": (i32, [vp]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(L, name)  # AttributeError if the library lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    if L.mlbsim_abi_version() != _abi.ABI_VERSION:
        raise ImportError(f"ABI version {L.mlbsim_abi_version()} != {_abi.ABI_VERSION}")
    _abi.check_sizes(L.mlbsim_struct_size)
    _lib = L
    return L


REPLAY_KINDS = {0: "general", 1: "regular", 2: "nonoise"}


def replay_kernel_kind(params: np.ndarray) -> str:
    """Which replay-kernel instance `mlbsim_run_replay*` picks for these parameters (host-only; include/mlbsim.h)."""
    params = np.ascontiguousarray(params)
    rc = load().mlbsim_replay_kernel_kind(_ptr(params))
    if rc < 0:
        raise MlbsimError(-rc, "mlbsim_replay_kernel_kind: bad parameters")
    return REPLAY_KINDS[rc]


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(ctypes.c_void_p)
    return ctypes.c_void_p(int(a))


class Context:
    """One `mlbsim_ctx` (device + stream).  Methods raise MlbsimError on any non-zero return."""

    def __init__(self, device: int = 0):
        self.L = load()
        h = ctypes.c_void_p()
        rc = self.L.mlbsim_init(int(device), ctypes.byref(h))
        if rc != 0:
            raise MlbsimError(rc, (self.L.mlbsim_last_error(None) or b"").decode())
        self.h = h
        self.device = int(device)
        self.n_tables = 0

    def close(self):
        if getattr(self, "h", None):
            self.L.mlbsim_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise MlbsimError(rc, (self.L.mlbsim_last_error(self.h) or b"").decode())

    # ---- facts / knobs ----
    def device_info(self):
        sm, khz = ctypes.c_int(), ctypes.c_int()
        name = ctypes.create_string_buffer(256)
        self._ck(self.L.mlbsim_device_info(self.h, ctypes.byref(sm), ctypes.byref(khz), name, 256))
        return {"sm_count": sm.value, "sm_clock_khz": khz.value, "name": name.value.decode()}

    def set_stream(self, stream: int):
        self._ck(self.L.mlbsim_set_stream(self.h, int(stream)))

    def synchronize(self):
        self._ck(self.L.mlbsim_synchronize(self.h))

    def set_launch_config(self, threads_per_block=0, blocks_per_sm=0):
        self._ck(self.L.mlbsim_set_launch_config(self.h, int(threads_per_block), int(blocks_per_sm)))

    def set_matchup_base(self, base: int):
        """Global matchup id of uploaded table 0 (matchup sharding: include/mlbsim.h)."""
        self._ck(self.L.mlbsim_set_matchup_base(self.h, int(base)))

    def launch_count(self) -> int:
        return int(self.L.mlbsim_launch_count(self.h))

    def last_kernel_ms(self) -> float:
        ms = ctypes.c_float()
        self._ck(self.L.mlbsim_last_kernel_ms(self.h, ctypes.byref(ms)))
        return float(ms.value)

    # ---- memory ----
    def malloc(self, nbytes: int) -> int:
        p = ctypes.c_void_p()
        self._ck(self.L.mlbsim_malloc(self.h, int(nbytes), ctypes.byref(p)))
        return int(p.value)

    def free(self, dev_ptr: int):
        self._ck(self.L.mlbsim_free(self.h, ctypes.c_void_p(dev_ptr)))

    def memset(self, dev_ptr: int, value: int, nbytes: int):
        self._ck(self.L.mlbsim_memset(self.h, ctypes.c_void_p(dev_ptr), value, nbytes))

    def h2d(self, dev_ptr: int, host: np.ndarray):
        host = np.ascontiguousarray(host)
        self._ck(self.L.mlbsim_memcpy_h2d(self.h, ctypes.c_void_p(dev_ptr), _ptr(host), host.nbytes))

    def d2h(self, host: np.ndarray, dev_ptr: int):
        assert host.flags.c_contiguous
        self._ck(self.L.mlbsim_memcpy_d2h(self.h, _ptr(host), ctypes.c_void_p(dev_ptr), host.nbytes))

    # ---- compute ----
    def alias_rows(self, probs: np.ndarray) -> np.ndarray:
        """Device form of `tables.alias_rows` ([N, 7 or 8] float64 -> [N, 8] uint32), bit-identical to it."""
        probs = np.ascontiguousarray(probs, dtype=np.float64)
        n, n_own = probs.shape
        out = np.zeros((n, _abi.ROW_WORDS), dtype=np.uint32)
        self._ck(self.L.mlbsim_alias_rows(self.h, _ptr(probs), n, n_own, _ptr(out)))
        return out

    def upload_tables(self, tables: np.ndarray):
        tables = np.ascontiguousarray(np.atleast_1d(tables))
        assert tables.dtype == _abi.TABLE_DTYPE
        self._ck(self.L.mlbsim_upload_tables(self.h, _ptr(tables), len(tables)))
        self.n_tables = len(tables)

    def build_tables(self, params: np.ndarray, validate: bool = True):
        """Build and install the tables of a batch of matchups ON THE DEVICE from their raw numbers (PARAMS_DTYPE records:
        `tables.build_params` / `tables.params_from_arrays`).  Raises `MatchupCannotScore` like the host builder."""
        params = np.ascontiguousarray(np.atleast_1d(params))
        assert params.dtype == _abi.PARAMS_DTYPE
        rc = self.L.mlbsim_build_tables(self.h, _ptr(params), len(params), 0 if validate else _abi.BUILD_NO_VALIDATE)
        if rc == _abi.ERR_CANNOT_SCORE:
            self.n_tables = 0
            raise _abi.MatchupCannotScore((self.L.mlbsim_last_error(self.h) or b"").decode())
        self._ck(rc)
        self.n_tables = len(params)

    def download_tables(self, lo: int = 0, hi: int | None = None) -> np.ndarray:
        """The installed tables [lo, hi) as TABLE_DTYPE records (the very bytes the kernels read)."""
        hi = self.n_tables if hi is None else hi
        out = np.zeros(hi - lo, dtype=_abi.TABLE_DTYPE)
        self._ck(self.L.mlbsim_download_tables(self.h, int(lo), int(hi), _ptr(out)))
        return out

    def run_native_host(self, matchup_lo, matchup_hi, sim_lo, sim_hi, seed, flags=0, out=None) -> np.ndarray:
        n = matchup_hi - matchup_lo
        if out is None:
            out = np.zeros(n, dtype=_abi.HIST_DTYPE)
        assert out.dtype == _abi.HIST_DTYPE and len(out) == n and out.flags.c_contiguous
        self._ck(self.L.mlbsim_run_native_host(self.h, matchup_lo, matchup_hi, sim_lo, sim_hi, seed, _ptr(out), flags))
        check_terminated(out)
        return out

    def run_native_summary_host(self, matchup_lo, matchup_hi, sim_lo, sim_hi, seed, flags=0, out=None) -> np.ndarray:
        n = matchup_hi - matchup_lo
        if out is None:
            out = np.zeros(n, dtype=_abi.SUMMARY_DTYPE)
        assert out.dtype == _abi.SUMMARY_DTYPE and len(out) == n and out.flags.c_contiguous
        self._ck(self.L.mlbsim_run_native_summary_host(self.h, matchup_lo, matchup_hi, sim_lo, sim_hi, seed, _ptr(out), flags))
        check_terminated(out)
        return out

    def allreduce(self, nccl_comm: int, hist_dev: int, n_matchups: int):
        """In-place sum of device histograms over the caller's ncclComm_t, on the context's stream."""
        self._ck(self.L.mlbsim_allreduce(self.h, ctypes.c_void_p(nccl_comm), ctypes.c_void_p(hist_dev), n_matchups))

    def summarize(self, hist_dev: int, n_matchups: int, summary_dev: int):
        self._ck(self.L.mlbsim_summarize(self.h, ctypes.c_void_p(hist_dev), n_matchups, ctypes.c_void_p(summary_dev)))

    def run_native(self, matchup_lo, matchup_hi, sim_lo, sim_hi, seed, hist_dev: int, flags=0):
        self._ck(self.L.mlbsim_run_native(self.h, matchup_lo, matchup_hi, sim_lo, sim_hi, seed, ctypes.c_void_p(hist_dev), flags))

    def run_replay(self, params: np.ndarray, tape: np.ndarray, offsets: np.ndarray) -> np.ndarray:
        params = np.ascontiguousarray(params)
        assert params.dtype == _abi.PARAMS_DTYPE
        tape = np.ascontiguousarray(tape, dtype=np.float64)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(offsets) - 1
        out = np.zeros(n, dtype=_abi.REPLAY_GAME_DTYPE)
        self._ck(self.L.mlbsim_run_replay(self.h, _ptr(params), _ptr(tape), _ptr(offsets), n, _ptr(out)))
        return out

    def run_replay_dev(self, params: np.ndarray, tape_dev: int, offsets_dev: int, n_games: int, out_dev: int):
        params = np.ascontiguousarray(params)
        self._ck(self.L.mlbsim_run_replay_dev(self.h, _ptr(params), ctypes.c_void_p(tape_dev), ctypes.c_void_p(offsets_dev),
                                              n_games, ctypes.c_void_p(out_dev)))

    def run_persim(self, matchup, sim_lo, sim_hi, seed) -> np.ndarray:
        n = sim_hi - sim_lo
        out = np.zeros(n, dtype=_abi.REPLAY_GAME_DTYPE)
        self._ck(self.L.mlbsim_run_persim(self.h, matchup, sim_lo, sim_hi, seed, _ptr(out)))
        return out
