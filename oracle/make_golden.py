"""TEST INFRASTRUCTURE.  Generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, build container only).  Re-run:  python oracle/make_golden.py [--stats]

replay_<scenario>.npz  value-level tapes + the integer results the reference produced on them
stats_<scenario>.npz   joint run histograms etc. of seeded reference batches (statistical oracle)
The matchup dicts themselves are stored as JSON inside each file so the tests need nothing else.
"""
from __future__ import annotations

import argparse
import copy
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import ref_harness as rh  # noqa: E402
from mlb_odds_engine_b200 import synth  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def scenarios():
    base = synth.make_matchup("2025-04-04-TEX@HOU", n_relievers=(6, 4))
    out = {}
    out["hetero_bullpen_noise"] = (base, True, 150)
    out["hetero_bullpen_nonoise"] = (base, False, 100)
    nobp = copy.deepcopy(base)
    nobp["home_bullpen"] = None
    nobp["away_bullpen"] = []
    out["hetero_nobullpen_noise"] = (nobp, True, 100)
    out["neutral_calibration"] = (synth.neutral_calibration_matchup(), True, 60)
    # one-sided bullpen, pitcher-friendly symmetric teams (extra innings), umpire/weather keys that
    # the hot path really reads (k_mod / bb_mod / weather_hr), EV/LA None, extreme speed/fielder
    odd = copy.deepcopy(synth.make_matchup("2025-04-05-NYY@BOS", n_relievers=(3, 5)))
    odd["away_bullpen"] = None
    odd["env"] = {"umpire": {"k_mod": 1.04, "bb_mod": 0.95}, "weather_hr": 1.12}
    for i, b in enumerate(odd["home_lineup"]):
        b["speed"] = [30, 70, 50][i % 3]
        b["k_rate"] = 0.30
    for i, b in enumerate(odd["away_lineup"]):
        b.pop("speed", None)
        b["k_rate"] = 0.30
    odd["home_pitcher"]["exit_velocity_avg"] = None
    odd["home_pitcher"]["launch_angle_avg"] = 4.0
    odd["home_pitcher"]["fielder_rating"] = 65
    odd["away_pitcher"]["exit_velocity_avg"] = 96.5
    odd["away_pitcher"].pop("launch_angle_avg")
    odd["away_pitcher"]["fielder_rating"] = 35
    odd["home_bullpen"][1]["exit_velocity_avg"] = 84.0
    odd["home_bullpen"][2]["IP"] = 3.0
    out["odd_onesided"] = (odd, True, 60)
    # three-batter lineups (the demo in half_inning_simulator.py:265-269) and degenerate rates
    short = copy.deepcopy(synth.make_matchup("2025-04-06-LAD@SF", n_relievers=(2, 2), lineup_len=3))
    short["home_lineup"][0]["k_rate"] = 0.0
    short["home_lineup"][0]["bb_rate"] = 0.0
    short["home_pitcher"]["k_rate"] = 0.0          # k == 0 for slot 0 -> beta_noise returns 0.0 without a draw
    short["away_lineup"][1]["bb_rate"] = 1.9
    short["away_pitcher"]["bb_rate"] = 0.15        # bb >= 1 for one slot -> no draw, walk unless K
    out["short_degenerate"] = (short, True, 60)
    return out


def jsonable(m):
    return json.dumps(m, sort_keys=True)


def make_replay(name, matchup, use_noise, n_games, seed):
    tapes, sums, recaps = rh.record_games(matchup, n_games, seed, use_noise=use_noise)
    lens = np.array([len(t) for t in tapes], dtype=np.int64)
    max_inn = max(s["n_innings"] for s in sums)
    max_used = max(1, max(max(len(s["used_home"]), len(s["used_away"])) for s in sums))
    away_runs = np.full((n_games, max_inn), -1, dtype=np.int16)
    home_runs = np.full((n_games, max_inn), -1, dtype=np.int16)
    used_home = np.full((n_games, max_used), -1, dtype=np.int16)
    used_away = np.full((n_games, max_used), -1, dtype=np.int16)
    for g, s in enumerate(sums):
        away_runs[g, : s["n_innings"]] = s["away_runs"]
        home_runs[g, : s["n_innings"]] = s["home_runs"]
        used_home[g, : len(s["used_home"])] = s["used_home"]
        used_away[g, : len(s["used_away"])] = s["used_away"]
    np.savez_compressed(
        os.path.join(GOLDEN, f"replay_{name}.npz"),
        matchup=jsonable(matchup), use_noise=use_noise, seed=seed,
        tape=np.concatenate(tapes), tape_len=lens,
        home_score=np.array([s["home_score"] for s in sums], dtype=np.int32),
        away_score=np.array([s["away_score"] for s in sums], dtype=np.int32),
        n_innings=np.array([s["n_innings"] for s in sums], dtype=np.int32),
        away_runs=away_runs, home_runs=home_runs, used_home=used_home, used_away=used_away,
        recap=np.array([s["recap"] for s in sums], dtype=np.int32),
        pa_recap=np.array(recaps, dtype=np.int32),
    )
    print(f"replay_{name}: {n_games} games, {lens.sum()} draws, mean tape {lens.mean():.1f}, max innings {max_inn}")


def make_stats(name, matchup, use_noise, n_games, seed):
    t = time.time()
    st = rh.stat_batch(matchup, n_games, seed, use_noise=use_noise)
    dt = time.time() - t
    np.savez_compressed(
        os.path.join(GOLDEN, f"stats_{name}.npz"),
        matchup=jsonable(matchup), use_noise=use_noise, seed=seed, n=st["n"],
        joint=st["joint"].astype(np.int32), innings=st["innings"], pa=st["pa"],
        rel_games=st["rel_games"], rel_picks=st["rel_picks"],
        ref_games_per_s=n_games / dt, ref_procs=os.cpu_count(),
    )
    print(f"stats_{name}: {n_games} games in {dt:.1f}s ({n_games / dt:.0f} games/s on {os.cpu_count()} procs)")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--stats", action="store_true", help="also regenerate the statistical fixtures (minutes)")
    ap.add_argument("--stats-games", type=int, default=200000)
    ap.add_argument("--only", default=None)
    args = ap.parse_args()
    os.makedirs(GOLDEN, exist_ok=True)
    sc = scenarios()
    for i, (name, (m, noise, n)) in enumerate(sc.items()):
        if args.only and args.only != name:
            continue
        make_replay(name, m, noise, n, seed=1000 + i)
    if args.stats:
        for i, name in enumerate(["hetero_bullpen_noise", "hetero_nobullpen_noise", "neutral_calibration", "odd_onesided"]):
            if args.only and args.only != name:
                continue
            m, noise, _ = sc[name]
            make_stats(name, m, noise, args.stats_games, seed=77 + i)


if __name__ == "__main__":
    main()
