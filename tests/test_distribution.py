"""CPU: histogram-fed reductions (N1) against what the UNMODIFIED reference priced on the same games
(tests/golden/distribution_TEXHOU.json, produced by oracle/make_golden.py --distribution)."""
import json
import math
import os

import numpy as np
import pytest

import oracle as orc
from mlb_odds_engine_b200 import _abi, distribution, synth, tables
from mlb_odds_engine_b200.simulator import RunHistogram, result_dicts

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
# two pins of the same chain: the calibration file present / a plain id; no calibration file (factors 1.0, segments
# unscaled: cli/run_distribution_simulator.py:351-354, 657-674) / an id with a -T1907 start-time suffix
FIXTURES = ["distribution_TEXHOU.json", "distribution_ARITOR_nocal.json"]


@pytest.fixture(scope="module", params=FIXTURES)
def case(request):
    fx = json.load(open(os.path.join(GOLDEN_DIR, request.param)))
    matchup = synth.make_matchup(fx["game_id"]) if fx.get("matchup_seed") is None else synth.make_matchup(fx["game_id"], seed=fx["matchup_seed"])
    table = tables.build_table(matchup)
    hist = orc.native(table, 0, 0, fx["n_sims"], fx["seed"])
    games = orc.native_persim(table, 0, 0, fx["n_sims"], fx["seed"])
    return fx, matchup, RunHistogram(hist, tables.as_matchup(matchup)), games


def test_markets_equal_the_reference_pricing(case):
    fx, matchup, hist, _ = case
    doc = distribution.price_from_histogram(hist, fx["game_id"], fx["calibration"])
    want, got = fx["doc"]["markets"], doc["markets"]
    assert len(want) == 186 and len(got) == len(want)
    for w, g in zip(want, got):
        assert (g["market"], g["side"]) == (w["market"], w["side"])
        assert g["sim_prob"] == pytest.approx(w["sim_prob"], abs=1e-12), (w, g)
        if math.isinf(w["fair_odds"]):
            assert g["fair_odds"] == w["fair_odds"]
        else:
            assert g["fair_odds"] == pytest.approx(w["fair_odds"], abs=1e-9), (w, g)


def test_moments_pmfs_and_summaries_equal_the_reference(case):
    fx, matchup, hist, _ = case
    doc = distribution.price_from_histogram(hist, fx["game_id"], fx["calibration"])
    ref = fx["doc"]
    assert doc["home_score"] == pytest.approx(ref["home_score"], rel=1e-13)
    assert doc["away_score"] == pytest.approx(ref["away_score"], rel=1e-13)
    for block in ("raw_distributions", "scaled_distributions"):
        for name, d in ref[block].items():
            for k, v in d.items():
                assert doc[block][name][k] == pytest.approx(v, rel=1e-11, abs=1e-11), (block, name, k)
    for seg, d in ref["summary_metrics"].items():
        for k, v in d.items():
            assert doc["summary_metrics"][seg][k] == pytest.approx(v, rel=1e-11, abs=1e-11), (seg, k)

    def same_pmf(a, b):
        a = {float(k): v for k, v in a.items()}
        b = {float(k): v for k, v in b.items()}
        assert a.keys() == b.keys()
        for k in a:
            assert a[k] == pytest.approx(b[k], abs=1e-12)

    same_pmf(doc["run_distribution"], ref["run_distribution"])
    same_pmf(doc["run_distribution_raw"], ref["run_distribution_raw"])
    same_pmf(doc["run_diff_distribution"], ref["run_diff_distribution"])
    for kind in ("totals", "spreads"):
        for seg, d in ref["pmfs"][kind].items():
            same_pmf(doc["pmfs"][kind][seg]["raw"], d["raw"])
            same_pmf(doc["pmfs"][kind][seg]["scaled"], d["scaled"])


def test_histogram_is_the_sufficient_statistic_of_the_per_sim_records(case):
    fx, matchup, hist, games = case
    dicts = result_dicts(games, tables.as_matchup(matchup))
    assert len(dicts) == fx["n_sims"]
    for ci, cap in enumerate(_abi.CAPS):
        j = np.zeros((64, 64), dtype=np.uint64)
        for r in dicts:
            inn = r["innings"] if cap is None else [i for i in r["innings"] if i["inning"] <= cap]
            j[min(sum(i["away_runs"] for i in inn), 63), min(sum(i["home_runs"] for i in inn), 63)] += 1
        np.testing.assert_array_equal(j, hist.rec["joint"][ci])
    usage = hist.reliever_usage()
    for side in ("home", "away"):
        want = {}
        for r in dicts:
            for name in set(r[f"used_{side}_relievers"]):
                want[name] = want.get(name, 0) + 1
        assert usage[side] == want
    recap = hist.pa_recap()
    assert sum(recap["HOME"].values()) + sum(recap["AWAY"].values()) == sum(sum(r["recap"].values()) for r in dicts)
    assert dicts[0]["game_type"] in {"Pitcher's Duel", "Blowout", "Tight Offensive Battle", "Balanced"}


def test_weighted_matches_numpy_on_expanded_lists():
    rng = np.random.default_rng(3)
    x = rng.integers(0, 20, 50)
    w = rng.integers(1, 9, 50)
    full = np.repeat(x, w).astype(float)
    v = distribution.Weighted(x, w)
    assert v.mean() == pytest.approx(full.mean(), rel=1e-14)
    assert v.std() == pytest.approx(full.std(), rel=1e-13)
    s = distribution.scale_distribution(v, 9.0, 4.3)
    assert s.mean() == pytest.approx(9.0, abs=1e-12) and s.std() == pytest.approx(4.3, rel=1e-12)
    assert distribution.to_american_odds(0.5) == 100.0 and distribution.to_american_odds(0.6) == -150.0
    assert distribution.to_american_odds(0.25) == 300.0 and distribution.to_american_odds(1.0) == -float("inf")
    assert distribution.adjust_for_push(0.45, 0.45) == (0.5, 0.5)


def test_start_time_and_table_snapshot_follow_the_reference(tmp_path, case):
    """N2: `start_time_iso` derived from the game id when no odds-file entry exists (run_distribution_simulator.py:243-250,
    core/utils.py:1009-1027) and the backtest/last_table_snapshot.json side file (:853-859), both byte-for-byte what the
    unmodified reference wrote for the same games."""
    fx, matchup, hist, _ = case
    doc = distribution.price_from_histogram(hist, fx["game_id"], fx["calibration"])
    assert doc["start_time_iso"] == fx["doc"]["start_time_iso"] and doc["start_time_iso"] is not None
    if "-T1907" in fx["game_id"]:
        assert doc["start_time_iso"].startswith("2025-06-17T19:07:00")
    snap = tmp_path / "backtest" / "last_table_snapshot.json"
    distribution.write_table_snapshot(doc, str(snap))
    assert json.load(open(snap)) == fx["table_snapshot"]
    assert distribution.start_time_iso_from_game_id("not-a-game-id") is None


def test_sim_json_roundtrip(tmp_path, case):
    fx, matchup, hist, _ = case
    doc = distribution.price_from_histogram(hist, fx["game_id"], fx["calibration"], start_time_iso="2025-04-04T19:10:00")
    path = tmp_path / "backtest" / "sims" / "2025-04-04" / f"{fx['game_id']}.json"
    distribution.write_sim_json(doc, str(path))
    back = json.load(open(path))
    # what core/snapshot_core.py:670-683 and cli/log_betting_evals.py:1994-2079 read
    assert back["start_time_iso"] == "2025-04-04T19:10:00"
    assert [(m["market"], m["side"], m["sim_prob"], m["fair_odds"]) for m in back["markets"]] == \
           [(m["market"], m["side"], m["sim_prob"], m["fair_odds"]) for m in doc["markets"]]
