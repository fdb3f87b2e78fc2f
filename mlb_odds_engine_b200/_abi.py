"""numpy mirrors of the C structs in include/mlbsim.h (single source of truth: the header).

`check_sizes(lib)` compares every itemsize with the library's own sizeof so a drifting header
fails loudly at import time instead of corrupting memory.
"""
from __future__ import annotations

import numpy as np

ABI_VERSION = 2

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

TABLE_MAGIC = 0x4D4C4254
FLAG_FATIGUE = 1

ST_INNINGS_OVERFLOW = 1
ST_USED_OVERFLOW = 2
ST_TAPE_UNDERRUN = 4

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
        ("reserved", "<u8", (7,)),
    ],
    align=True,
)

HIST_WORDS = HIST_DTYPE.itemsize // 8

STRUCTS = {0: PARAMS_DTYPE, 1: TABLE_DTYPE, 2: HIST_DTYPE, 3: REPLAY_GAME_DTYPE}


def check_sizes(sizeof) -> None:
    """`sizeof(which) -> int` from the loaded library; raises on any mismatch."""
    names = {0: "mlbsim_params", 1: "mlbsim_table", 2: "mlbsim_hist", 3: "mlbsim_replay_game"}
    for which, dt in STRUCTS.items():
        got = int(sizeof(which))
        if got != dt.itemsize:
            raise ImportError(f"ABI mismatch: sizeof({names[which]}) = {got} in the library, {dt.itemsize} in _abi.py")
