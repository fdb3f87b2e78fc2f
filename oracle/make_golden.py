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
    # the 30-PA cap (half_inning_simulator.py:160,174): the home lineup never makes an out (HR every PA),
    # so every bottom half ends on the cap; the away side plays normally
    cap = copy.deepcopy(synth.make_matchup("2025-04-08-ATL@PHI", n_relievers=(2, 3)))
    for b in cap["home_lineup"]:
        b["k_rate"] = 0.0
        b["bb_rate"] = 0.0
    for p in [cap["away_pitcher"]] + cap["away_bullpen"]:
        p["k_rate"] = 0.0
        p["bb_rate"] = 0.0
        p["hr_pa"] = {"hr_pa_projected": 1.0}
    out["pa_cap"] = (cap, True, 20)
    # no noise and no bullpens together (fatigue ramp with deterministic K/BB probabilities)
    nn = copy.deepcopy(synth.make_matchup("2025-04-09-MIN@DET"))
    nn["home_bullpen"] = None
    nn["away_bullpen"] = None
    out["nonoise_nobullpen"] = (nn, False, 60)
    # the matchup bench.py times (BASELINE configs[0]/[1]): lets the bench line carry its own parity figures
    out["bench_texhou"] = (synth.make_matchup("2025-04-04-TEX@HOU"), True, 40)
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


# Fixtures that are the sum of several independent seeded batches: name -> (seeds, games per batch).  The batches are
# merged HERE, so `--stats --only <name>` reproduces the committed file (8 worker processes: the per-process seeds
# derive from the batch seed and the process index, ref_harness.stat_batch).
MERGED_STATS = {"bench_texhou": ((83, 991), 1_000_000)}
STATS_PROCS = 8


def make_stats(name, matchup, use_noise, n_games, seed, out_dir=None):
    seeds, per_batch = MERGED_STATS.get(name, ((seed,), n_games))
    t = time.time()
    st = None
    for sd in seeds:
        part = rh.stat_batch(matchup, per_batch, sd, use_noise=use_noise, procs=STATS_PROCS)
        st = part if st is None else {k: st[k] + part[k] for k in st}
    dt = time.time() - t
    extra = {"seeds": np.array(seeds, dtype=np.int64)} if len(seeds) > 1 else {}
    np.savez_compressed(
        os.path.join(out_dir or GOLDEN, f"stats_{name}.npz"),
        matchup=jsonable(matchup), use_noise=use_noise, seed=seeds[0], n=st["n"],
        joint=st["joint"].astype(np.int32), innings=st["innings"], pa=st["pa"],
        rel_games=st["rel_games"], rel_picks=st["rel_picks"],
        ref_games_per_s=st["n"] / dt, ref_procs=STATS_PROCS, **extra,
    )
    print(f"stats_{name}: {st['n']} games in {dt:.1f}s ({st['n'] / dt:.0f} games/s on {STATS_PROCS} procs)")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--stats", action="store_true", help="also regenerate the statistical fixtures (minutes)")
    ap.add_argument("--stats-games", type=int, default=200000)
    ap.add_argument("--only", default=None)
    ap.add_argument("--no-replay", action="store_true", help="leave the replay fixtures alone (with --stats)")
    ap.add_argument("--out-dir", default=None, help="write the stats fixtures somewhere else (to compare with the committed ones)")
    args = ap.parse_args()
    os.makedirs(GOLDEN, exist_ok=True)
    sc = scenarios()
    for i, (name, (m, noise, n)) in enumerate(sc.items()):
        if args.no_replay or (args.only and args.only != name):
            continue
        make_replay(name, m, noise, n, seed=1000 + i)
    if args.stats:
        # (stats_bench_texhou.npz holds 2 x 10^6 games: two seeded batches merged in make_stats, see MERGED_STATS)
        for i, name in enumerate(["hetero_bullpen_noise", "hetero_nobullpen_noise", "neutral_calibration", "odd_onesided",
                                  "short_degenerate", "hetero_bullpen_nonoise", "bench_texhou"]):
            if args.only and args.only != name:
                continue
            m, noise, _ = sc[name]
            make_stats(name, m, noise, args.stats_games, seed=77 + i, out_dir=args.out_dir)


if __name__ == "__main__" and not any(a in sys.argv for a in ("--distribution", "--distribution2", "--assets")):
    main()


# ---------------------------------------------------------------------------------------------
# N1 fixture: the reference's own reduction / pricing chain on a fixed set of per-sim results
# ---------------------------------------------------------------------------------------------
def make_distribution_fixture(n_sims=10000, seed=20250404, game_id="2025-04-04-TEX@HOU", with_calibration=True,
                              out_name="distribution_TEXHOU.json", matchup_seed=None):
    """Runs the UNMODIFIED `simulate_distribution` (cli/run_distribution_simulator.py:195) with its
    module-global `simulate_game` replaced by an iterator over per-sim results produced by the
    oracle's native twin (deterministic: the test regenerates them), and stores what it priced."""
    import importlib.machinery
    import shutil
    import types

    import oracle as orc
    from mlb_odds_engine_b200 import tables

    rh.load_reference()

    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        m.__spec__ = importlib.machinery.ModuleSpec(name, None)
        sys.modules[name] = m
        return m

    mp = stub("matplotlib"); mp.pyplot = stub("matplotlib.pyplot"); stub("bs4", BeautifulSoup=object)
    sel = stub("selenium"); sel.webdriver = stub("selenium.webdriver"); stub("selenium.webdriver.chrome")
    stub("selenium.webdriver.chrome.options", Options=object); stub("selenium.webdriver.chrome.service", Service=object)
    for extra in ("selenium.webdriver.common", "selenium.webdriver.common.by", "selenium.webdriver.support",
                  "selenium.webdriver.support.ui", "selenium.webdriver.support.expected_conditions", "dotenv"):
        if extra not in sys.modules:
            stub(extra, By=object, WebDriverWait=object, load_dotenv=lambda *a, **k: None)
    import cli.run_distribution_simulator as rds

    matchup = synth.make_matchup(game_id) if matchup_seed is None else synth.make_matchup(game_id, seed=matchup_seed)
    table = tables.build_table(matchup)
    games = orc.native_persim(table, 0, 0, n_sims, seed)
    names = [[p["name"] for p in matchup["home_bullpen"]], [p["name"] for p in matchup["away_bullpen"]]]

    def results():
        for g in games:
            k = int(g["n_innings"])
            yield {
                "home_score": int(g["home_score"]), "away_score": int(g["away_score"]),
                "innings": [{"inning": i + 1, "away_runs": int(g["away_runs"][i]), "home_runs": int(g["home_runs"][i])} for i in range(k)],
                "used_home_relievers": [names[0][int(x)] for x in g["used_home"][: int(g["n_used"][0])]],
                "used_away_relievers": [names[1][int(x)] for x in g["used_away"][: int(g["n_used"][1])]],
            }

    it = results()
    rds.simulate_game = lambda **kw: next(it)
    rds.load_all_stats = lambda: ({}, {})
    starters = {}
    for side in ("home", "away"):
        p = dict(matchup[f"{side}_pitcher"])
        starters[side] = p
    rds.build_game_assets = lambda gid, bs, ps: {
        "lineups": {"home": matchup["home_lineup"], "away": matchup["away_lineup"]},
        "pitchers": starters,
        "bullpens": {"home": matchup["home_bullpen"], "away": matchup["away_bullpen"]},
    }
    import tempfile
    os.chdir(tempfile.mkdtemp(prefix="mlbsim_golden_"))   # the reference reads and writes CWD-relative paths
    os.makedirs("logs", exist_ok=True)
    if with_calibration:
        shutil.copy(os.path.join(rh.REFERENCE_ROOT, "logs", "calibration_offset.json"), "logs/calibration_offset.json")
    # (without the file the reference prices with factors 1.0 and leaves the segments unscaled:
    #  cli/run_distribution_simulator.py:351-354, 657-674)
    rds.simulate_distribution(game_id, 9.5, no_weather=True, export_json="./dist_out.json", n_simulations=n_sims)
    doc = json.load(open("./dist_out.json"))
    snapshot = json.load(open(rds.SNAPSHOT_PATH))      # the side file of :853-859
    for block in ("raw_distributions", "scaled_distributions"):
        for v in doc[block].values():
            v.pop("values", None)
    fixture = {
        "game_id": game_id, "n_sims": n_sims, "seed": seed, "matchup_seed": matchup_seed,
        "calibration": json.load(open("logs/calibration_offset.json")) if with_calibration else None,
        "doc": doc, "table_snapshot": snapshot,
    }
    with open(os.path.join(GOLDEN, out_name), "w") as f:
        json.dump(fixture, f, indent=1, sort_keys=True)
    print(f"distribution fixture: {len(doc['markets'])} market entries, home {doc['home_score']:.3f} away {doc['away_score']:.3f}")


if __name__ == "__main__" and "--distribution" in sys.argv:
    make_distribution_fixture()
if __name__ == "__main__" and "--distribution2" in sys.argv:
    # second pin of N1: no calibration file in the working directory, a game id with a start-time suffix
    make_distribution_fixture(seed=77, game_id="2025-06-17-ARI@TOR-T1907", with_calibration=False,
                              out_name="distribution_ARITOR_nocal.json", matchup_seed=5150)


# ---------------------------------------------------------------------------------------------
# N3 fixture: the reference's HR/PA projection and bullpen builder on seeded stats rows
# ---------------------------------------------------------------------------------------------
def make_assets_fixture(n_pitchers=300, seed=20250405):
    """Calls the UNMODIFIED `project_hr_pa` (core/project_hr_pa.py:46) and `build_bullpen_for_team`
    (assets/bullpen_utils.py:19; its network lookup of today's starters replaced by an empty dict) on
    seeded synthetic stats rows and stores inputs + outputs as JSON (floats round-trip exactly)."""
    import copy, contextlib, io, json, math
    mods = rh.load_reference()
    rng = np.random.default_rng(seed)
    opt = {"xSLG": (0.3, 0.5), "xwOBAcon": (0.3, 0.45), "sweet_spot_pct": (0.25, 0.4), "xSLG_diff": (-0.05, 0.05),
           "wOBAdiff": (-0.04, 0.04), "K_pct": (0.12, 0.36), "BB_pct": (0.03, 0.14), "FIP": (2.5, 5.5),
           "location_plus": (80.0, 120.0)}
    rows = []
    for i in range(n_pitchers):
        tbf = float(rng.choice([0, 1, 37, 180, 499, 500, 650, 699, 700, 899, 900, 1100]))
        p = {"name": f"Pitcher {i}", "role": "RP" if rng.random() < 0.5 else "SP",
             "stuff_plus": float(rng.choice([70, 90, 90.5, 95, 95.5, 100, 104.2, 110, 115, 120, 131.7])),
             "HR": float(rng.integers(0, 40)), "TBF": tbf, "IP": float(rng.choice([0, 1, 55.3, 180.1])),
             "barrel_batted_rate": float(rng.uniform(0.02, 0.12)), "exit_velocity_avg": float(rng.normal(88.5, 2)),
             "launch_angle_avg": float(rng.normal(13, 4))}
        for k, (lo, hi) in opt.items():
            if rng.random() < 0.5:
                p[k] = float(rng.uniform(lo, hi))
        quirk = rng.integers(0, 12)
        if quirk == 0:
            p["exit_velocity_avg"] = "N/A"
        elif quirk == 1:
            p["launch_angle_avg"] = None
        elif quirk == 2:
            p["barrel_batted_rate"] = float("nan")
        elif quirk == 3:
            del p["role"], p["stuff_plus"]
        elif quirk == 4:
            p["league_avg_hr_pa"] = 0.0275
        elif quirk == 5:
            del p["TBF"], p["IP"], p["HR"]
        rows.append(p)
    proj = []
    with contextlib.redirect_stdout(io.StringIO()):
        for p in rows:
            proj.append(mods.php.project_hr_pa(copy.deepcopy(p)))

    # bullpen builder: 3 teams x 10 candidate rows, some unusable
    bu = mods.bu
    bu.fetch_probable_pitchers = lambda: {"g": {"home": {"name": "AAA Arm 0"}, "away": {"name": "Nobody"}}}
    stats = {}
    for team in ("AAA", "BBB", "CCC"):
        for j in range(10):
            r = {"name": f"{team} Arm {j}", "team_abbr": team, "throws": "L" if j % 3 == 0 else "R",
                 "k_rate": float(rng.uniform(0.15, 0.35)), "bb_rate": float(rng.uniform(0.04, 0.14)),
                 "hr_fb_rate": float(rng.choice([0.09, 11.5, 14.0, 0.13])), "stuff_plus": float(rng.normal(100, 12)),
                 "command_plus": float(rng.normal(100, 8)), "location_plus": float(rng.normal(100, 8)),
                 "barrel_batted_rate": float(rng.uniform(0.03, 0.1)), "exit_velocity_avg": float(rng.normal(88, 2)),
                 "launch_angle_avg": float(rng.normal(13, 3)), "HR": float(rng.integers(0, 12)),
                 "TBF": float(rng.integers(0, 320))}
            if j == 4:
                r["stuff_plus"] = None
            if j == 5:
                r["hr_fb_rate"] = float("nan")
            if j == 6:
                del r["k_rate"], r["exit_velocity_avg"], r["TBF"]
            if j == 7:
                r["bb_rate"] = "bad"
            stats[bu.normalize_name(r["name"])] = r
    pens = {}
    with contextlib.redirect_stdout(io.StringIO()):
        for team in ("AAA", "BBB", "CCC"):
            pens[team] = bu.build_bullpen_for_team(team, copy.deepcopy(stats), None)
        pens["AAA_top3"] = bu.build_bullpen_for_team("AAA", copy.deepcopy(stats), None, max_relievers=3)
    out = {"generator": "oracle/make_golden.py --assets", "seed": seed, "pitchers": rows, "projections": proj,
           "pitcher_stats": stats, "excluded_starters": [bu.normalize_name("AAA Arm 0")], "bullpens": pens}
    path = os.path.join(GOLDEN, "assets_hr_pa_bullpen.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=0)
    print(f"wrote {path}: {len(rows)} projections, {sum(len(v) for v in pens.values())} bullpen entries")


if __name__ == "__main__" and "--assets" in sys.argv:
    make_assets_fixture()
