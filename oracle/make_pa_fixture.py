"""TEST INFRASTRUCTURE — plate-appearance-level fixture from the UNMODIFIED reference (build container only).

    python oracle/make_pa_fixture.py          -> tests/golden/pa_outcome_counts.json

For each probe point: `apply_fatigue_modifiers(pitcher, {pitch_count, tto_count})` (core/fatigue_modeling.py:7-54) followed by
`simulate_pa(batter, adjusted, umpire, weather_hr, use_noise=...)` (core/pa_simulator.py:91-143), N = 400 000 times with
seeded generators, outcome counts stored.  The game-level fixtures resolve a probability to ~3.5e-4; this pins the
per-PA law (`tables.outcome_probabilities`) itself at ~7e-4 per point and tier, including one extreme (k ~ 0.5,
bb ~ 0.3: the clip term of the noise integral matters there) and pitch counts beyond 75.
"""
from __future__ import annotations

import copy
import json
import multiprocessing as mp
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import ref_harness as rh  # noqa: E402
from mlb_odds_engine_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "pa_outcome_counts.json")
N = 400_000
OUTCOMES = ["K", "BB", "1B", "2B", "3B", "HR", "OUT"]


def points():
    """(label, matchup, slot, tto_count, pitch_count, use_noise).  The batter is away_lineup[slot], the pitcher the home
    starter (side 0, pitcher 0 of the tables)."""
    base = synth.make_matchup("2025-04-04-TEX@HOU")
    pts = []
    for noise in (True, False):
        for tto in (1, 2, 3, 4):
            pts.append((f"typical_tto{tto}", base, 2, tto, 10 * tto, noise))
        ext = copy.deepcopy(base)
        ext["away_lineup"][0].update(k_rate=0.52, bb_rate=0.31, speed=72)
        ext["home_pitcher"].update(k_rate=0.48, bb_rate=0.29, exit_velocity_avg=93.0, launch_angle_avg=22.0)
        ext["env"] = dict(ext["env"], umpire={"k_mod": 1.05, "bb_mod": 1.04}, weather_hr=1.15)
        pts.append(("extreme_k_bb", ext, 0, 2, 40, noise))
        pts.append(("fatigued_pc100", base, 5, 3, 100, noise))
        pts.append(("fatigued_pc140_tto5", base, 7, 5, 140, noise))
    return pts


def _run(args):
    idx, (label, m, slot, tto, pc, noise) = args
    ref = rh.load_reference()
    random.seed(1000 + idx)
    np.random.seed(2000 + idx)
    batter, pitcher, env = m["away_lineup"][slot], m["home_pitcher"], m["env"]
    counts = dict.fromkeys(OUTCOMES, 0)
    sim_pa, fat = ref.pa.simulate_pa, ref.fm.apply_fatigue_modifiers
    state = {"batters_faced": pc, "pitch_count": pc, "tto_count": tto}
    for _ in range(N):
        adj = fat(pitcher, state)                                     # half_inning_simulator.py:180
        out = sim_pa(batter, adj, env.get("umpire", {}), env.get("weather_hr", 1.0), pc, use_noise=noise)   # :182-194
        counts[out] += 1
    return {"label": label, "matchup": m, "slot": slot, "tto_count": tto, "pitch_count": pc, "use_noise": noise,
            "n": N, "counts": [counts[o] for o in OUTCOMES]}


if __name__ == "__main__":
    rh.load_reference()
    jobs = list(enumerate(points()))
    with mp.get_context("fork").Pool(min(len(jobs), os.cpu_count() or 1)) as pool:
        res = pool.map(_run, jobs)
    with open(OUT, "w") as f:
        json.dump({"outcomes": OUTCOMES, "points": res}, f)
    for r in res:
        print(r["label"], r["use_noise"], [round(c / r["n"], 4) for c in r["counts"]])
    print("wrote", OUT)
