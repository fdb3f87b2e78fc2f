"""The BASELINE.json workloads as data: which matchups, how many sims, how they shard over GPUs, and the small
fixed-(seed, sims) run whose result digest must be the same for ANY number of GPUs.  Shared by bench.py (measurement),
oracle/make_golden.py --digests (the oracle twin's digests, committed as tests/golden/bench_digests.json) and
tests/test_bench_digests.py (CPU: twin == committed file).  No CUDA in here."""
from __future__ import annotations

import hashlib

import numpy as np

SEED = 20250404
GAME_ID = "2025-04-04-TEX@HOU"

# name -> (BASELINE.json config, scaling over GPUs, shard axis, sims per matchup of one FULL job, digest sims per matchup)
SPEC = {
    "single": dict(config=1, scaling="strong", shard="sims", sims=10_000_000, digest_sims=1_000_000,
                   what="single game 2025-04-04-TEX@HOU x 10M sims, split over the ranks by sim index"),
    "slate": dict(config=2, scaling="strong", shard="sims", sims=1_000_000, digest_sims=100_000,
                  what="15-game slate x 1M sims per game, every game split over the ranks by sim index"),
    "fatigue": dict(config=3, scaling="strong", shard="sims", sims=100_000_000, digest_sims=1_000_000,
                    what="fatigue-heavy game (no bullpens: TTO 4+, pitch-count ramp, extras) x 100M sims, split by sim index"),
    "sweep": dict(config=4, scaling="strong", shard="matchups", sims=100_000, digest_sims=2_000,
                  what="1024-point sensitivity sweep (k_rate x bb_rate x hr_pa) x 100k sims, grid points split over the ranks"),
}


def count(n: int) -> str:
    return f"{n // 1_000_000}M" if n % 1_000_000 == 0 and n else f"{n // 1000}k" if n % 1000 == 0 and n else str(n)


def sweep_grid():
    from mlb_odds_engine_b200 import synth

    base = synth.make_matchup(GAME_ID, bullpens=False)   # the tools/ scripts simulate without bullpens
    axes = dict(k_rate=np.linspace(0.14, 0.34, 16), bb_rate=np.linspace(0.04, 0.12, 8), hr_pa_projected=np.linspace(0.015, 0.06, 8))
    return base, axes


def matchups(name: str):
    """`simulate_game` keyword dicts of a workload (the sweep's 1 024 dicts are only needed by the CPU-reference legs)."""
    from mlb_odds_engine_b200 import sweep, synth

    if name == "single":
        return [synth.make_matchup(GAME_ID)]
    if name == "fatigue":
        return [synth.make_matchup(GAME_ID, bullpens=False)]   # starters go the distance
    if name == "slate":
        return synth.make_slate("2025-04-04", 15)
    base, axes = sweep_grid()
    return sweep.grid_matchups(base, "home", **axes)[0]


def build_tables(name: str, alias_fn=None) -> np.ndarray:
    """TABLE_DTYPE array of a workload, built the way a user would (the sweep through sweep.grid_tables).  `alias_fn`:
    the device builder of the alias rows (`Context.alias_rows`) where a GPU is at hand — same bytes either way."""
    from mlb_odds_engine_b200 import sweep, tables

    if name == "sweep":
        base, axes = sweep_grid()
        return sweep.grid_tables(base, "home", alias_fn=alias_fn, **axes)[0]
    return tables.build_tables(matchups(name), alias_fn=alias_fn)


def build_params(name: str) -> np.ndarray:
    """PARAMS_DTYPE array of a workload — the input of the DEVICE table builder (`ShardedSimulator.load_params`): the raw
    numbers go up, the tables are built in HBM (same bytes as `build_tables` for every workload here; a test pins that)."""
    from mlb_odds_engine_b200 import sweep, tables

    if name == "sweep":
        base, axes = sweep_grid()
        return sweep.grid_params(base, "home", **axes)[0]
    return tables.params_batch(matchups(name))


def digest(records: np.ndarray) -> str:
    """sha256 over the raw bytes of HIST_DTYPE (sim shards) or SUMMARY_DTYPE (matchup shards) records, in matchup order."""
    return hashlib.sha256(np.ascontiguousarray(records).tobytes()).hexdigest()
