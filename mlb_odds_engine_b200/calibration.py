"""Neutral-game calibration on the GPU path — `tools/simulate_calibration_game.py` fed from histograms.

The reference tool plays 100 000 neutral games through `simulate_game` (:73-92), reduces per-sim lists to scaling
factors (:94-143, :229-258) and writes `logs/calibration_offset.json` (:266-294) — the file
`simulate_distribution` reads back (cli/run_distribution_simulator.py:351) — and `logs/calibration_outcomes.csv`
(:218-226).  Every number in both files is a function of the joint (away, home) run counts at caps {1,3,5,7,full}
and of the outcome counters, so here N is free to be 10⁷–10⁸: `calibrate(hist)` computes the same documents from one
`RunHistogram`.  Per-game outcome SDs (:190-194, printed only) need per-sim records: `per_game_outcome_sd`.

Moments are population moments (`np.mean`, `np.std` with ddof 0, as the tool uses) evaluated over weighted cells.
"""
from __future__ import annotations

import csv
import json
import math
import os

import numpy as np

from . import _abi, synth
from .simulator import GpuSimulator, RunHistogram

# tools/simulate_calibration_game.py:56-64
MLB_BENCHMARKS = {"K": 22.5, "BB": 8.2, "1B": 21.5, "2B": 5.2, "3B": 0.5, "HR": 3.0, "OUT": 39.1}
# :110-116
TARGETS = {"mean": 8.85, "total_sd": 4.65, "diff_sd": 3.0, "home_mean": 4.62, "away_mean": 4.35,
           "home_sd": 3.35, "away_sd": 3.20}
# :231-236
SEGMENT_TARGETS = {
    "f1": {"run_mean": 1.05, "run_sd": 1.55, "diff_sd": 1.55},
    "f3": {"run_mean": 3.30, "run_sd": 2.70, "diff_sd": 2.70},
    "f5": {"run_mean": 5.40, "run_sd": 3.35, "diff_sd": 3.35},
    "f7": {"run_mean": 7.25, "run_sd": 3.85, "diff_sd": 3.85},
}
# :130-137: the tool regresses logit(observed) on logit(simulated) over this fixed four-point set
_LOGIT_SIM = (0.52, 0.60, 0.67, 0.75)
_LOGIT_OBS = (0.50, 0.56, 0.62, 0.70)


def _logit(p):
    return math.log(p / (1.0 - p))


def logit_calibration() -> tuple[float, float]:
    """Intercept and slope of the ordinary least-squares line the tool fits (:130-137)."""
    x = np.array([_logit(p) for p in _LOGIT_SIM])
    y = np.array([_logit(p) for p in _LOGIT_OBS])
    xm, ym = x.mean(), y.mean()
    b = float(((x - xm) * (y - ym)).sum() / ((x - xm) ** 2).sum())
    return float(ym - b * xm), b


def _moments(values: np.ndarray, counts: np.ndarray) -> tuple[float, float]:
    n = counts.sum()
    mean = float((values * counts).sum() / n)
    var = float((((values - mean) ** 2) * counts).sum() / n)
    return mean, math.sqrt(var)


def calibrate(hist: RunHistogram) -> dict:
    """{"calibration_offset": <logs/calibration_offset.json document>, "outcomes": <rows of
    logs/calibration_outcomes.csv>, "summary": raw means / SDs / win % the tool prints, "segments": per-cap factors}."""
    n = hist.n_games
    if n <= 0:
        raise ValueError("calibration needs at least one simulated game")
    h, a, c = hist.score_counts()
    c = c.astype(np.float64)
    total_mean, total_sd = _moments((h + a).astype(np.float64), c)
    diff_mean, diff_sd = _moments((h - a).astype(np.float64), c)
    home_mean, home_sd = _moments(h.astype(np.float64), c)
    away_mean, away_sd = _moments(a.astype(np.float64), c)

    actual_std = round(total_sd, 3)                                         # :119
    a_, b_ = logit_calibration()
    home_win = float(c[h > a].sum() / c.sum())
    calibrated = 1.0 / (1.0 + math.exp(-(a_ + b_ * _logit(home_win)))) if 0.0 < home_win < 1.0 else home_win

    segments = {}
    for cap, key in ((1, "f1"), (3, "f3"), (5, "f5"), (7, "f7")):
        hs, as_, cs = hist.score_counts(cap)
        cs = cs.astype(np.float64)
        m, sd = _moments((hs + as_).astype(np.float64), cs)
        _, dsd = _moments((hs - as_).astype(np.float64), cs)
        tgt = SEGMENT_TARGETS[key]
        segments[key] = {
            "run_mean": round(tgt["run_mean"], 2), "run_sd": round(tgt["run_sd"], 2), "diff_sd": round(tgt["diff_sd"], 2),
            "run_sd_scaling_factor": round(tgt["run_sd"] / sd, 4) if sd > 0 else 1.0,
            "run_mean_shift": round(tgt["run_mean"] - m, 4),
            "diff_sd_scaling_factor": round(tgt["diff_sd"] / dsd, 4) if dsd > 0 else 1.0,
            "actual_mean": m, "actual_sd": sd, "actual_diff_sd": dsd,
        }

    offset = {
        "run_scaling_factor": round(TARGETS["mean"] / total_mean, 4),            # :264-265
        "stddev_scaling_factor": round(TARGETS["total_sd"] / actual_std, 4),      # :120
        "run_diff_scaling_factor": round(TARGETS["diff_sd"] / diff_sd, 4),        # :124
        "team_total_scaling": {
            "home_mean_factor": round(TARGETS["home_mean"] / home_mean, 4),
            "home_std_factor": round(TARGETS["home_sd"] / home_sd, 4),
            "away_mean_factor": round(TARGETS["away_mean"] / away_mean, 4),
            "away_std_factor": round(TARGETS["away_sd"] / away_sd, 4),
        },
        "logit_win_pct_calibration": {"a": round(a_, 4), "b": round(b_, 4)},
        "segment_scaling": {k: {f: v[f] for f in ("run_mean", "run_sd", "diff_sd")} for k, v in segments.items()},
    }

    pa = hist.rec["pa"][:, :7].astype(np.int64).sum(axis=0)
    total_pa = int(pa.sum())
    outcomes = [{"Outcome": o, "Count": int(pa[i]),
                 "Percent": f"{(100 * int(pa[i]) / total_pa if total_pa else 0.0):.2f}%",
                 "Per Game": f"{int(pa[i]) / n:.2f}", "MLB Benchmark": MLB_BENCHMARKS.get(o, "")}
                for i, o in enumerate(_abi.OUTCOMES)]
    return {
        "calibration_offset": offset,
        "outcomes": outcomes,
        "segments": segments,
        "summary": {"n_games": n, "total_mean": total_mean, "total_sd": total_sd, "diff_mean": diff_mean,
                    "diff_sd": diff_sd, "home_mean": home_mean, "away_mean": away_mean, "home_sd": home_sd,
                    "away_sd": away_sd, "home_win_pct": home_win, "calibrated_home_win_pct": calibrated,
                    "total_pa": total_pa, "pa_per_game": total_pa / n},
    }


def per_game_outcome_sd(games: np.ndarray) -> dict:
    """Sample SD (ddof 1, `statistics.stdev`) of each outcome's per-game count over per-sim records (:190-194)."""
    per_game = games["pa"].astype(np.int64).sum(axis=1)[:, :7]
    if len(per_game) < 2:
        return {o: 0.0 for o in _abi.OUTCOMES}
    sd = per_game.std(axis=0, ddof=1)
    return {o: float(sd[i]) for i, o in enumerate(_abi.OUTCOMES)}


def write_files(doc: dict, out_dir: str = "logs") -> tuple[str, str]:
    """`logs/calibration_offset.json` and `logs/calibration_outcomes.csv` as the tool writes them (:218-226, :292-294)."""
    os.makedirs(out_dir, exist_ok=True)
    jpath, cpath = os.path.join(out_dir, "calibration_offset.json"), os.path.join(out_dir, "calibration_outcomes.csv")
    with open(jpath, "w") as f:
        json.dump(doc["calibration_offset"], f, indent=2)
    with open(cpath, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["Outcome", "Count", "Percent", "Per Game", "MLB Benchmark"])
        for r in doc["outcomes"]:
            w.writerow([r["Outcome"], r["Count"], r["Percent"], r["Per Game"], r["MLB Benchmark"]])
    return jpath, cpath


def run_calibration(n_sims: int = 100_000, seed: int = 0, sim: GpuSimulator | None = None, matchup: dict | None = None,
                    out_dir: str | None = None) -> dict:
    """Simulate the neutral game (or `matchup`) on the GPU and reduce it; writes the two files when `out_dir` is set."""
    own = sim is None
    sim = sim or GpuSimulator(0)
    try:
        sim.load_matchups([matchup or synth.neutral_calibration_matchup()])
        doc = calibrate(sim.simulate(n_sims, seed=seed)[0])
    finally:
        if own:
            sim.close()
    if out_dir is not None:
        doc["files"] = write_files(doc, out_dir)
    return doc
