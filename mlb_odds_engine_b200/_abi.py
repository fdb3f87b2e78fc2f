"""numpy mirrors of the C structs in include/mlbsim.h (single source of truth: the header).

`check_sizes(lib)` compares every itemsize with the library's own sizeof so a drifting header
fails loudly at import time instead of corrupting memory.
"""
from __future__ import annotations

import numpy as np

ABI_VERSION = 6

MAX_LINEUP = 9
MAX_BULLPEN = 6
MAX_PITCHERS = 7
N_OUTCOMES = 7
N_TIERS = 4
N_BIP = 4
ROW_WORDS = 8
HIST_BINS = 64
N_CAPS = 5
INN_BINS = 32
REPLAY_MAX_INNINGS = 40
REPLAY_MAX_USED = 16

OUTCOMES = ("K", "BB", "1B", "2B", "3B", "HR", "OUT")
CAPS = (1, 3, 5, 7, None)  # innings cap of joint[c]; None = full game

TABLE_MAGIC = 0x4D4C4233
FLAG_FATIGUE = 1

ST_INNINGS_OVERFLOW = 1
ST_USED_OVERFLOW = 2
ST_TAPE_UNDERRUN = 4
ST_TRUNCATED = 8
MAX_INNINGS = 128

PARAMS_DTYPE = np.dtype(
    [
        ("lineup_len", "<i4", (2,)),
        ("bullpen_len", "<i4", (2,)),
        ("use_noise", "<i4"),
        ("reserved", "<i4"),
        ("k_mod", "<f8"),
        ("bb_mod", "<f8"),
        ("bat_k", "<f8", (2, MAX_LINEUP)),
        ("bat_bb", "<f8", (2, MAX_LINEUP)),
        ("pit_k", "<f8", (2, MAX_PITCHERS)),
        ("pit_bb", "<f8", (2, MAX_PITCHERS)),
        ("pit_hr", "<f8", (2, MAX_PITCHERS)),
        ("hit_prob", "<f8", (2, MAX_PITCHERS, MAX_LINEUP, N_BIP)),
        ("cdf_bip", "<f8", (4,)),
        ("cdf_hit", "<f8", (3,)),
        ("cdf_fb", "<f8", (2,)),
    ],
    align=True,
)

REPLAY_GAME_DTYPE = np.dtype(
    [
        ("home_score", "<i4"),
        ("away_score", "<i4"),
        ("n_innings", "<i4"),
        ("tape_used", "<i4"),
        ("n_used", "<i4", (2,)),
        ("status", "<i4"),
        ("reserved", "<i4"),
        ("away_runs", "u1", (REPLAY_MAX_INNINGS,)),
        ("home_runs", "u1", (REPLAY_MAX_INNINGS,)),
        ("used_home", "u1", (REPLAY_MAX_USED,)),
        ("used_away", "u1", (REPLAY_MAX_USED,)),
        ("pa", "<u2", (2, 8)),
    ],
    align=True,
)

TABLE_DTYPE = np.dtype(
    [
        ("magic", "<u4"),
        ("flags", "<u4"),
        ("lineup_len", "<u4", (2,)),
        ("bullpen_len", "<u4", (2,)),
        ("reserved", "<u4", (2,)),
        ("alias", "<u4", (2, MAX_PITCHERS, N_TIERS, MAX_LINEUP, ROW_WORDS)),
        ("fthr", "<u4", (2, N_TIERS, MAX_LINEUP, ROW_WORDS)),
        ("fslope", "<i4", (2, N_TIERS, MAX_LINEUP, ROW_WORDS)),
    ],
    align=True,
)

HIST_DTYPE = np.dtype(
    [
        ("joint", "<u8", (N_CAPS, HIST_BINS, HIST_BINS)),
        ("innings", "<u8", (INN_BINS,)),
        ("pa", "<u8", (2, 8)),
        ("rel_games", "<u8", (2, 8)),
        ("rel_picks", "<u8", (2, 8)),
        ("n_games", "<u8"),
        ("n_truncated", "<u8"),
        ("reserved", "<u8", (6,)),
    ],
    align=True,
)

HIST_WORDS = HIST_DTYPE.itemsize // 8

SUMMARY_DTYPE = np.dtype(
    [
        ("n_games", "<u8"),
        ("pa", "<u8", (2, 8)),
        ("sum_runs", "<u8", (N_CAPS, 2)),   # [cap][0 away, 1 home]
        ("sum_sq", "<u8", (2,)),
        ("sum_cross", "<u8"),
        ("home_wins", "<u8"),
        ("away_wins", "<u8"),
        ("sum_innings", "<u8"),
        ("n_truncated", "<u8"),
        ("reserved", "<u8", (6,)),
    ],
    align=True,
)

ERR_CANNOT_SCORE = 6
BUILD_NO_VALIDATE = 1


class MatchupCannotScore(ValueError):
    """Every reachable plate appearance of BOTH sides is a strikeout or an out: the game can never leave 0-0.  The
    reference loops forever on such input (core/game_simulator.py:110-111); tables are refused instead."""


STRUCTS = {0: PARAMS_DTYPE, 1: TABLE_DTYPE, 2: HIST_DTYPE, 3: REPLAY_GAME_DTYPE, 4: SUMMARY_DTYPE}


def summarize_hist(rec: np.ndarray) -> np.ndarray:
    """Host statement of `mlbsim_summarize` (csrc/summary_kernel.cu) for one HIST_DTYPE record: what the device
    summary must equal, integer for integer.  Used by the tests and by callers that already hold a histogram."""
    out = np.zeros((), dtype=SUMMARY_DTYPE)
    j = rec["joint"].astype(object)
    a = np.arange(HIST_BINS, dtype=object)[:, None]
    h = np.arange(HIST_BINS, dtype=object)[None, :]
    out["n_games"] = rec["n_games"]
    out["n_truncated"] = rec["n_truncated"]
    out["pa"] = rec["pa"]
    for c in range(N_CAPS):
        out["sum_runs"][c, 0] = int((j[c] * a).sum())
        out["sum_runs"][c, 1] = int((j[c] * h).sum())
    full = j[N_CAPS - 1]
    out["sum_sq"][0], out["sum_sq"][1] = int((full * a * a).sum()), int((full * h * h).sum())
    out["sum_cross"] = int((full * a * h).sum())
    out["home_wins"] = int(np.where(h > a, full, 0).sum())
    out["away_wins"] = int(np.where(a > h, full, 0).sum())
    out["sum_innings"] = int((rec["innings"].astype(object) * np.arange(INN_BINS, dtype=object)).sum())
    return out


def check_sizes(sizeof) -> None:
    """`sizeof(which) -> int` from the loaded library; raises on any mismatch."""
    names = {0: "mlbsim_params", 1: "mlbsim_table", 2: "mlbsim_hist", 3: "mlbsim_replay_game", 4: "mlbsim_summary"}
    for which, dt in STRUCTS.items():
        got = int(sizeof(which))
        if got != dt.itemsize:
            raise ImportError(f"ABI mismatch: sizeof({names[which]}) = {got} in the library, {dt.itemsize} in _abi.py")
