"""TEST INFRASTRUCTURE — not product code.  Drives the UNMODIFIED reference (imported read-only
from /root/reference in the build container; on the GPU box, where that path does not exist, from the byte code
oracle/make_ref.py compiled into oracle/_ref/) to

  * record value-level replay tapes of `simulate_game` (SURVEY.md §8(c) "Replay tape"), and
  * batch-run seeded games over worker processes for statistical oracles,

and writes the results as fixtures under tests/golden/ (see `oracle/make_golden.py`).

Nothing under the product package imports this file.  Reference call sites it hooks, without
modifying any reference file:
  core/pa_simulator.py:32,48,54,72,74,79,122   (np.random.beta / random / choice)
  core/bip_resolution.py:62                     (stdlib random.random)
  core/half_inning_simulator.py:15,23,35,76,103,105,120
  assets/bullpen_utils.py:142,144,146          (random.choice / random.choices)
"""
from __future__ import annotations

import importlib.machinery
import os
import random
import sys
import types

import numpy as np

_BYTECODE_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
REFERENCE_ROOT = os.environ.get("MLBSIM_REFERENCE_ROOT") or (
    "/root/reference" if os.path.isdir("/root/reference/core") else _BYTECODE_ROOT)
OUTCOMES = ["K", "BB", "1B", "2B", "3B", "HR", "OUT"]

_mods = None


_BYTECODE_EXT = ".refbc"


class _BytecodeFinder:
    """Imports the reference's modules from oracle/_ref/<package>/<module>.refbc (pyc-format byte code compiled from
    the unmodified checkout by oracle/make_ref.py)."""

    def __init__(self, root):
        self.root = root

    def find_spec(self, fullname, path=None, target=None):
        import importlib.util

        base = os.path.join(self.root, *fullname.split("."))
        init = os.path.join(base, "__init__" + _BYTECODE_EXT)
        if os.path.isfile(init):
            loader = importlib.machinery.SourcelessFileLoader(fullname, init)
            return importlib.util.spec_from_file_location(fullname, init, loader=loader, submodule_search_locations=[base])
        if os.path.isfile(base + _BYTECODE_EXT):
            loader = importlib.machinery.SourcelessFileLoader(fullname, base + _BYTECODE_EXT)
            return importlib.util.spec_from_file_location(fullname, base + _BYTECODE_EXT, loader=loader)
        if os.path.isdir(base) and any(f.endswith(_BYTECODE_EXT) for f in os.listdir(base)):   # namespace package (assets/)
            spec = importlib.machinery.ModuleSpec(fullname, None, is_package=True)
            spec.submodule_search_locations = [base]
            return spec
        return None


def reference_available() -> bool:
    if reference_kind() == "bytecode":
        return os.path.isfile(os.path.join(REFERENCE_ROOT, "core", "game_simulator" + _BYTECODE_EXT))
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "core"))


def reference_kind() -> str:
    """'source' (the checkout) or 'bytecode' (oracle/_ref, compiled from it by oracle/make_ref.py)."""
    return "bytecode" if os.path.realpath(REFERENCE_ROOT) == os.path.realpath(_BYTECODE_ROOT) else "source"


def load_reference():
    """Import the reference's hot-path modules (SURVEY.md Appendix B recipe)."""
    global _mods
    if _mods is not None:
        return _mods
    if not reference_available():
        raise RuntimeError(f"reference checkout not found at {REFERENCE_ROOT}")
    import pandas  # noqa: F401  must be imported before the pytz stub exists

    if "pytz" not in sys.modules:
        m = types.ModuleType("pytz")
        m.__spec__ = importlib.machinery.ModuleSpec("pytz", None)
        sys.modules["pytz"] = m
    sys.dont_write_bytecode = True
    if reference_kind() == "bytecode":
        if not any(isinstance(f, _BytecodeFinder) for f in sys.meta_path):
            sys.meta_path.insert(0, _BytecodeFinder(REFERENCE_ROOT))
    elif REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import core.game_simulator as gs
    import core.pa_simulator as pa
    import core.bip_resolution as bip
    import core.half_inning_simulator as hi
    import core.fatigue_modeling as fm
    import assets.bullpen_utils as bu
    import core.project_hr_pa as php

    _mods = types.SimpleNamespace(gs=gs, pa=pa, bip=bip, hi=hi, fm=fm, bu=bu, php=php)
    return _mods


class _NpTape:
    """Stands in for `np.random` inside core.pa_simulator; same stream consumption, values taped."""

    def __init__(self, tape):
        self.tape = tape

    def beta(self, a, b):
        v = np.random.beta(a, b)
        self.tape.append(float(v))
        return v

    def random(self):
        v = np.random.random()
        self.tape.append(float(v))
        return v

    def choice(self, seq, p=None):
        # legacy RandomState.choice(p=...) = one uniform + searchsorted(cdf, u, 'right')
        u = np.random.random_sample()
        self.tape.append(float(u))
        cdf = np.cumsum(np.asarray(p, dtype=np.float64))
        cdf /= cdf[-1]
        return seq[int(np.searchsorted(cdf, u, side="right"))]


class _PyTape:
    """Stands in for stdlib `random` in bip_resolution / half_inning_simulator / bullpen_utils."""

    def __init__(self, tape, bullpens):
        self.tape = tape
        self.bullpens = bullpens  # list of the full bullpen lists (to map picks to full-list index)

    def random(self):
        v = random.random()
        self.tape.append(float(v))
        return v

    def _index(self, pick):
        for bp in self.bullpens:
            if bp:
                for i, r in enumerate(bp):
                    if r is pick:
                        return i
        raise RuntimeError("picked reliever not found in any bullpen")

    def choice(self, seq):
        pick = random.choice(seq)
        self.tape.append(float(self._index(pick)))
        return pick

    def choices(self, seq, weights=None, k=1):
        picks = random.choices(seq, weights=weights, k=k)
        for pick in picks:
            self.tape.append(float(self._index(pick)))
        return picks


def saturate_reliever_counts(matchup):
    """Pre-set RELIEVER_USAGE_COUNTS["home"][name] = 3 so picks follow the stationary law
    (uniform over the bullpen with replacement; assets/bullpen_utils.py:124,141-142)."""
    ref = load_reference()
    for key in ("home_bullpen", "away_bullpen"):
        for rp in matchup.get(key) or []:
            ref.bu.RELIEVER_USAGE_COUNTS["home"][rp.get("name", "Unknown")] = 3


def reset_reliever_counts():
    ref = load_reference()
    ref.bu.RELIEVER_USAGE_COUNTS["home"].clear()
    ref.bu.RELIEVER_USAGE_COUNTS["away"].clear()


def _game_kwargs(matchup, use_noise):
    kw = {k: matchup[k] for k in ("home_lineup", "away_lineup", "home_pitcher", "away_pitcher", "env")}
    kw["home_bullpen"] = matchup.get("home_bullpen")
    kw["away_bullpen"] = matchup.get("away_bullpen")
    kw["use_noise"] = use_noise
    return kw


def summarize_result(result, matchup):
    """The integer outputs replay mode must match (core/game_simulator.py:134-144)."""
    def idx(names, bullpen):
        lut = {r.get("name", "Unknown"): i for i, r in enumerate(bullpen or [])}
        return [lut[n] for n in names]

    innings = result["innings"]
    return {
        "home_score": int(result["home_score"]),
        "away_score": int(result["away_score"]),
        "n_innings": len(innings),
        "away_runs": [int(i["away_runs"]) for i in innings],
        "home_runs": [int(i["home_runs"]) for i in innings],
        "used_home": idx(result["used_home_relievers"], matchup.get("home_bullpen")),
        "used_away": idx(result["used_away_relievers"], matchup.get("away_bullpen")),
        "recap": [int(result["recap"][o]) for o in OUTCOMES],
    }


def record_games(matchup, n_games, seed, use_noise=True, saturate=False):
    """Run `n_games` reference games with taped RNG.  Returns (tapes, summaries, pa_recap) where
    pa_recap[g] = per-game [AWAY[7], HOME[7]] outcome counts from PA_RECAP_LOG deltas."""
    ref = load_reference()
    tape: list[float] = []
    saved = (ref.pa.np, ref.bip.random, ref.hi.random, ref.bu.random)
    pyt = _PyTape(tape, [matchup.get("home_bullpen"), matchup.get("away_bullpen")])
    ref.pa.np = types.SimpleNamespace(random=_NpTape(tape))
    ref.bip.random = ref.hi.random = ref.bu.random = pyt
    reset_reliever_counts()
    if saturate:
        saturate_reliever_counts(matchup)
    random.seed(seed)
    np.random.seed(seed)
    tapes, sums, recaps = [], [], []
    try:
        for _ in range(n_games):
            tape.clear()
            for t in ("HOME", "AWAY"):
                for o in OUTCOMES:
                    ref.pa.PA_RECAP_LOG[t][o] = 0
            res = ref.gs.simulate_game(**_game_kwargs(matchup, use_noise))
            tapes.append(np.asarray(tape, dtype=np.float64))
            sums.append(summarize_result(res, matchup))
            recaps.append([[ref.pa.PA_RECAP_LOG[t][o] for o in OUTCOMES] for t in ("AWAY", "HOME")])
    finally:
        ref.pa.np, ref.bip.random, ref.hi.random, ref.bu.random = saved
        reset_reliever_counts()
    return tapes, sums, recaps


# ---------------------------------------------------------------------------------------------
# statistical oracle: joint (away, home) counts at caps {1,3,5,7,full} etc.
# ---------------------------------------------------------------------------------------------
HB = 64          # bins per team in the joint histograms (matches include/mlbsim.h MLBSIM_HIST_BINS)
INN_BINS = 32


def new_stats():
    return {
        "joint": np.zeros((5, HB, HB), dtype=np.int64),     # [cap index][away][home]
        "innings": np.zeros(INN_BINS, dtype=np.int64),
        "pa": np.zeros((2, 7), dtype=np.int64),             # [AWAY, HOME][outcome]
        "rel_games": np.zeros((2, 8), dtype=np.int64),      # [away, home] sims in which reliever i appeared
        "rel_picks": np.zeros((2, 8), dtype=np.int64),      # total picks
        "n": 0,
    }


def accumulate(stats, result, matchup):
    innings = result["innings"]
    a = h = 0
    caps = {}
    for inn in innings:
        a += inn["away_runs"]
        h += inn["home_runs"]
        if inn["inning"] in (1, 3, 5, 7):
            caps[inn["inning"]] = (a, h)
    for ci, cap in enumerate((1, 3, 5, 7)):
        ca, ch = caps[cap]
        stats["joint"][ci, min(ca, HB - 1), min(ch, HB - 1)] += 1
    stats["joint"][4, min(result["away_score"], HB - 1), min(result["home_score"], HB - 1)] += 1
    stats["innings"][min(len(innings), INN_BINS - 1)] += 1
    for side, key, bp in ((0, "used_away_relievers", "away_bullpen"), (1, "used_home_relievers", "home_bullpen")):
        lut = {r.get("name", "Unknown"): i for i, r in enumerate(matchup.get(bp) or [])}
        for n in result[key]:
            stats["rel_picks"][side, lut[n]] += 1
        for n in set(result[key]):
            stats["rel_games"][side, lut[n]] += 1
    stats["n"] += 1


def _stat_worker(args):
    matchup, n_games, seed, use_noise = args
    ref = load_reference()
    reset_reliever_counts()
    saturate_reliever_counts(matchup)
    random.seed(seed)
    np.random.seed(seed % (2**32))
    for t in ("HOME", "AWAY"):
        for o in OUTCOMES:
            ref.pa.PA_RECAP_LOG[t][o] = 0
    st = new_stats()
    kw = _game_kwargs(matchup, use_noise)
    for _ in range(n_games):
        accumulate(st, ref.gs.simulate_game(**kw), matchup)
    st["pa"][0] = [ref.pa.PA_RECAP_LOG["AWAY"][o] for o in OUTCOMES]
    st["pa"][1] = [ref.pa.PA_RECAP_LOG["HOME"][o] for o in OUTCOMES]
    return st


def stat_batch(matchup, n_games, seed, use_noise=True, procs=None):
    """Reference games over `procs` worker processes with distinct seeds; summed statistics."""
    import multiprocessing as mp

    procs = procs or os.cpu_count() or 1
    per = [n_games // procs + (1 if i < n_games % procs else 0) for i in range(procs)]
    jobs = [(matchup, per[i], seed * 1000 + i, use_noise) for i in range(procs) if per[i]]
    with mp.get_context("fork").Pool(len(jobs)) as pool:
        parts = pool.map(_stat_worker, jobs)
    tot = new_stats()
    for p in parts:
        for k in tot:
            tot[k] = tot[k] + p[k]
    return tot


# ---------------------------------------------------------------------------------------------
# timed batch: the reference's own CPU path on all host cores for a bounded time (bench.py's cpu_baseline)
# ---------------------------------------------------------------------------------------------
def _timed_worker(args):
    import time

    matchups, seconds, seed, use_noise, want_stats = args
    ref = load_reference()
    reset_reliever_counts()
    for m in matchups:
        saturate_reliever_counts(m)
    random.seed(seed)
    np.random.seed(seed % (2**32))
    kws = [_game_kwargs(m, use_noise) for m in matchups]
    stats = [new_stats() for _ in matchups] if want_stats else None
    sim = ref.gs.simulate_game
    n = 0
    t0 = time.perf_counter()
    deadline = t0 + seconds
    while True:
        for i, kw in enumerate(kws):          # the loop body of cli/run_distribution_simulator.py:368-382
            if want_stats:
                for t in ("HOME", "AWAY"):
                    for o in OUTCOMES:
                        ref.pa.PA_RECAP_LOG[t][o] = 0
            res = sim(**kw)
            if want_stats:
                accumulate(stats[i], res, matchups[i])
                stats[i]["pa"][0] += [ref.pa.PA_RECAP_LOG["AWAY"][o] for o in OUTCOMES]
                stats[i]["pa"][1] += [ref.pa.PA_RECAP_LOG["HOME"][o] for o in OUTCOMES]
        n += len(kws)
        if time.perf_counter() >= deadline:
            break
    return n, time.perf_counter() - t0, stats


def _fixed_worker(args):
    """`n_games` reference games round-robin over `matchups` (bench.py --impl reference: one step of one worker)."""
    matchups, n_games, seed, use_noise = args
    ref = load_reference()
    for m in matchups:
        saturate_reliever_counts(m)
    random.seed(seed)
    np.random.seed(seed % (2**32))
    kws = [_game_kwargs(m, use_noise) for m in matchups]
    sim = ref.gs.simulate_game
    done = 0
    while done < n_games:
        for kw in kws:
            sim(**kw)
        done += len(kws)
    return done


def timed_batch(matchups, seconds, seed=1, use_noise=True, procs=None, want_stats=False):
    """Run the unmodified `simulate_game` round-robin over `matchups` in `procs` forked worker processes (default: every
    host core) for about `seconds` each.  Returns (games, wall_seconds, procs, stats | None); games / wall_seconds is the
    reference's CPU throughput on this host."""
    import multiprocessing as mp
    import time

    procs = procs or os.cpu_count() or 1
    load_reference()                      # import once; the workers inherit it through fork
    jobs = [(matchups, seconds, seed * 1000 + i, use_noise, want_stats) for i in range(procs)]
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(procs) as pool:
        parts = pool.map(_timed_worker, jobs)
    wall = time.perf_counter() - t0
    games = sum(p[0] for p in parts)
    busy = max(p[1] for p in parts)       # workers start together; the slowest one bounds the batch
    stats = None
    if want_stats:
        stats = [new_stats() for _ in matchups]
        for p in parts:
            for i in range(len(matchups)):
                for k in stats[i]:
                    stats[i][k] = stats[i][k] + p[2][i][k]
    return games, busy, procs, stats, wall
