"""Histogram-fed neutral-game calibration vs the list-based arithmetic of tools/simulate_calibration_game.py
(:94-143, :229-258), on games played by the CPU oracle (checker only)."""
import csv
import json
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
import oracle as orc  # noqa: E402

from mlb_odds_engine_b200 import _abi, calibration, synth, tables  # noqa: E402
from mlb_odds_engine_b200.simulator import RunHistogram  # noqa: E402

N = 20_000


@pytest.fixture(scope="module")
def played():
    t = tables.build_tables([synth.neutral_calibration_matchup()])
    hist = RunHistogram(orc.native(t, 0, 0, N, 11).reshape(-1)[0])
    games = orc.native_persim(t, 0, 0, N, 11)
    return hist, games


def test_logit_fit_is_the_published_one():
    # logs/calibration_offset.json "logit_win_pct_calibration": {"a": -0.0827, "b": 0.8327}
    a, b = calibration.logit_calibration()
    assert (round(a, 4), round(b, 4)) == (-0.0827, 0.8327)


def test_factors_equal_list_based_tool_arithmetic(played):
    hist, games = played
    doc = calibration.calibrate(hist)
    home = [int(x) for x in games["home_score"]]
    away = [int(x) for x in games["away_score"]]
    total = [h + a for h, a in zip(home, away)]
    diff = [h - a for h, a in zip(home, away)]
    T = calibration.TARGETS
    want = {
        "run_scaling_factor": round(T["mean"] / np.mean(total), 4),
        "stddev_scaling_factor": round(T["total_sd"] / round(np.std(total), 3), 4),
        "run_diff_scaling_factor": round(T["diff_sd"] / np.std(diff), 4),
    }
    off = doc["calibration_offset"]
    for k, v in want.items():
        assert off[k] == v, k
    tt = off["team_total_scaling"]
    assert tt["home_mean_factor"] == round(T["home_mean"] / np.mean(home), 4)
    assert tt["away_mean_factor"] == round(T["away_mean"] / np.mean(away), 4)
    assert tt["home_std_factor"] == round(T["home_sd"] / np.std(home), 4)
    assert tt["away_std_factor"] == round(T["away_sd"] / np.std(away), 4)
    assert doc["summary"]["home_win_pct"] == sum(h > a for h, a in zip(home, away)) / N
    # segments (:96-105, :238-258)
    for cap, key in ((1, "f1"), (3, "f3"), (5, "f5"), (7, "f7")):
        hs = games["home_runs"][:, :cap].astype(np.int64).sum(axis=1)
        as_ = games["away_runs"][:, :cap].astype(np.int64).sum(axis=1)
        tgt = calibration.SEGMENT_TARGETS[key]
        seg = doc["segments"][key]
        assert seg["run_sd_scaling_factor"] == round(tgt["run_sd"] / np.std(hs + as_), 4)
        assert seg["run_mean_shift"] == round(tgt["run_mean"] - np.mean(hs + as_), 4)
        assert seg["diff_sd_scaling_factor"] == round(tgt["diff_sd"] / np.std(hs - as_), 4)
        assert off["segment_scaling"][key] == {"run_mean": tgt["run_mean"], "run_sd": tgt["run_sd"], "diff_sd": tgt["diff_sd"]}


def test_outcome_rows_and_files(played, tmp_path):
    hist, games = played
    doc = calibration.calibrate(hist)
    counts = games["pa"].astype(np.int64).sum(axis=(0, 1))[:7]
    assert [r["Count"] for r in doc["outcomes"]] == [int(x) for x in counts]
    assert [r["Outcome"] for r in doc["outcomes"]] == list(_abi.OUTCOMES)
    # the reference's published neutral-game mix (logs/calibration_outcomes.csv), within MC error at N = 20 000
    pct = {r["Outcome"]: float(r["Percent"].rstrip("%")) for r in doc["outcomes"]}
    for o, want in {"K": 22.23, "BB": 8.36, "1B": 19.58, "2B": 4.77, "3B": 0.40, "HR": 2.52, "OUT": 42.14}.items():
        assert abs(pct[o] - want) < 0.15, (o, pct[o])
    jpath, cpath = calibration.write_files(doc, str(tmp_path / "logs"))
    back = json.load(open(jpath))
    assert back == doc["calibration_offset"] and list(back) == [
        "run_scaling_factor", "stddev_scaling_factor", "run_diff_scaling_factor", "team_total_scaling",
        "logit_win_pct_calibration", "segment_scaling"]
    rows = list(csv.reader(open(cpath)))
    assert rows[0] == ["Outcome", "Count", "Percent", "Per Game", "MLB Benchmark"] and len(rows) == 8
    sd = calibration.per_game_outcome_sd(games)
    assert 3.5 < sd["K"] < 5.0 and 0.4 < sd["3B"] < 0.8
    with pytest.raises(ValueError):
        calibration.calibrate(RunHistogram(np.zeros((), dtype=_abi.HIST_DTYPE)))


def test_distribution_accepts_the_generated_calibration(played):
    """The document written here is the one `simulate_distribution` reads (cli/run_distribution_simulator.py:351)."""
    from mlb_odds_engine_b200 import distribution

    hist, _ = played
    off = calibration.calibrate(hist)["calibration_offset"]
    doc = distribution.price_from_histogram(hist, "2025-04-04-NEA@NEH", calibration=off)
    assert len(doc["markets"]) == 186
