"""Histogram-fed version of the reductions in `simulate_distribution`
(cli/run_distribution_simulator.py:391-794 + extract_universal_markets :96-192).

Every per-sim quantity the reference derives is a function of the (home, away) run pair after 1, 3,
5, 7 innings or the full game and of global moments of those pairs, and the reductions are
permutation-invariant (SURVEY.md Appendix B addendum), so the joint counts of `mlbsim_hist.joint`
are sufficient statistics: this module evaluates the same chain over the non-empty cells with
their counts as weights — O(cells) instead of O(sims), so 10^7-10^8 sims per game never
materialise per-sim lists.  Formulas, rounding and label formats follow the reference line by line
(cited inline); the result is the reference's JSON document minus the per-sim `values` arrays
(nothing downstream reads those: core/snapshot_core.py:670-683, cli/log_betting_evals.py:1994-2079
use only `markets` and `start_time_iso`).
"""
from __future__ import annotations

import bisect
import functools
import json
import math
import os
import tempfile

import numpy as np

from . import _abi

# run_distribution_simulator.py:198-229
BENCHMARK_TOTALS = {
    "full_game": {"mean_total": 9.00, "std_total": 4.30, "mean_diff": 0.00, "std_diff": 4.30},
    "f1": {"mean_total": 1.05, "std_total": 1.55, "mean_diff": 0.00, "std_diff": 1.55},
    "f3": {"mean_total": 3.30, "std_total": 2.70, "mean_diff": 0.00, "std_diff": 2.70},
    "f5": {"mean_total": 5.40, "std_total": 3.35, "mean_diff": 0.00, "std_diff": 3.35},
    "f7": {"mean_total": 7.25, "std_total": 3.85, "mean_diff": 0.00, "std_diff": 3.85},
}
# run_distribution_simulator.py:622-652
SEGMENT_CONFIGS = {
    "F1": {"label": "1st Inning", "innings": 1, "total_lines": [0.5]},
    "F3": {"label": "First 3 Innings", "innings": 3, "total_lines": [2.5, 3.5, 4.5],
           "spread_lines": [0.5, 1.5, 2.5], "team_total_lines": [0.5, 1.5, 2.5, 3.5]},
    "F5": {"label": "First 5 Innings", "innings": 5, "total_lines": [3.5, 4.5, 5.5, 6.5],
           "spread_lines": [0.5, 1.5, 2.5], "team_total_lines": [1.5, 2.5, 3.5, 4.5, 5.5]},
    "F7": {"label": "First 7 Innings", "innings": 7, "total_lines": [5.5, 6.5, 7.5, 8.5],
           "spread_lines": [0.5, 1.5, 2.5], "team_total_lines": [2.5, 3.5, 4.5, 5.5, 6.5]},
}
SEG_NAME = {"f1": "1st_1_innings", "f3": "1st_3_innings", "f5": "1st_5_innings", "f7": "1st_7_innings"}
SEGMENT_TO_MARKET_KEY = {"1st Inning": "1st_1_innings", "First 3 Innings": "1st_3_innings",
                         "First 5 Innings": "1st_5_innings", "First 7 Innings": "1st_7_innings"}


# ---------------------------------------------------------------------------------------------
# small restatements of the helpers the reference chain calls
# ---------------------------------------------------------------------------------------------
def to_american_odds(prob):
    """core/market_pricer.py:9-22."""
    if prob >= 1.0:
        return -float("inf")
    if prob <= 0.0:
        return float("inf")
    decimal = 1 / prob
    if decimal >= 2:
        return round((decimal - 1) * 100, 2)
    return round(-100 / (decimal - 1), 2)


def adjust_for_push(p_win, p_loss):
    """core/market_pricer.py:84-88."""
    p_push = max(0.0, 1.0 - (p_win + p_loss))
    denom = max(1.0 - p_push, 1e-8)
    return p_win / denom, p_loss / denom


class Weighted:
    """A per-sim list in run-length form: `x[i]` occurs `w[i]` times."""

    def __init__(self, x, w):
        self.x = np.asarray(x, dtype=np.float64)
        self.w = np.asarray(w, dtype=np.float64)
        self.n = float(self.w.sum())

    def mean(self):  # np.mean
        return float(np.dot(self.x, self.w) / self.n)

    def std(self):   # np.std, ddof = 0
        m = self.mean()
        return float(math.sqrt(np.dot((self.x - m) ** 2, self.w) / self.n))

    def map(self, f):
        return Weighted(f(self.x), self.w)

    def prob(self, mask):
        return float(self.w[mask].sum() / self.n)

    def pmf_rounded(self):
        """summarize_pmf(np.round(values).astype(int)) — core/stats_tools.py:40-51 (np.round = half to even).  One
        bincount over the cells instead of one masked sum per key (the weights are integer counts: the sums are exact
        in any order)."""
        keys, inverse = np.unique(np.round(self.x).astype(np.int64), return_inverse=True)
        mass = np.bincount(inverse.reshape(-1), weights=self.w, minlength=len(keys)) / self.n
        return Pmf(zip(keys.astype(np.float64).tolist(), mass.tolist()))


class Pmf(dict):
    """{value: probability} in ascending value order (what `json.dump` writes), with the sorted columns cached so that
    a tail probability is a bisection and one C-level sum over a slice instead of a Python loop over the items."""

    def columns(self):
        c = self.__dict__.get("_cols")
        if c is None or len(c[0]) != len(self):
            c = self.__dict__["_cols"] = (list(self.keys()), list(self.values()))
        return c


def scale_distribution(v: Weighted, target_mean=None, target_sd=None) -> Weighted:
    """core/scaling_utils.py:18-28, same operation order."""
    raw_mean, raw_sd = v.mean(), v.std()
    scaled = v.x
    if target_sd is not None and raw_sd > 0:
        scaled = (scaled - raw_mean) * (target_sd / raw_sd) + raw_mean
    if target_mean is not None:
        scaled = scaled + (target_mean - raw_mean)
    return Weighted(scaled, v.w)


def tail(pmf, threshold, direction):
    """core/stats_tools.py:13-28."""
    if isinstance(pmf, Pmf):      # ascending keys: the same left-to-right sums as the loops below, without the loop
        keys, probs = pmf.columns()
        if direction == "over":
            return sum(probs[bisect.bisect_right(keys, threshold):])
        if direction == "under":
            return sum(probs[:bisect.bisect_left(keys, threshold)])
        return pmf.get(float(int(threshold)), 0.0)
    if direction == "over":
        return sum(p for x, p in pmf.items() if x > threshold)
    if direction == "under":
        return sum(p for x, p in pmf.items() if x < threshold)
    return pmf.get(float(int(threshold)), 0.0)


def team_total_scaling(v: Weighted, mean_factor, std_factor) -> Weighted:
    """core/pricing_engine.py:78-84."""
    mean_score = v.mean()
    return Weighted(((v.x - mean_score) * std_factor + mean_score) * mean_factor, v.w)


def _round1(x):
    """[round(v, 1) for v in x] with Python's float rounding (run_distribution_simulator.py:489-490)."""
    return np.array([round(float(v), 1) for v in x], dtype=np.float64)


def _entry(market, side, prob, odds):
    """build_entry, run_distribution_simulator.py:101-108."""
    return {"market": market, "side": side, "sim_prob": round(prob, 4), "fair_odds": round(odds, 2), "source": "simulator"}


SNAPSHOT_PATH = os.path.join("backtest", "last_table_snapshot.json")   # cli/run_distribution_simulator.py:25


def parse_game_id(game_id: str) -> dict:
    """{"date", "away", "home", "time"} of 'YYYY-MM-DD-AWAY@HOME[-T1905[-DH1]]' (core/utils.py:920-939, including its
    fallback for strings it cannot split)."""
    try:
        parts = game_id.split("-")
        if len(parts) < 4:
            raise ValueError("invalid game_id")
        away, home = parts[3].split("@")
        return {"date": "-".join(parts[:3]), "away": away, "home": home, "time": "-".join(parts[4:])}
    except Exception:
        return {"date": game_id, "away": "", "home": "", "time": ""}


@functools.lru_cache(maxsize=1)
def _eastern():
    from zoneinfo import ZoneInfo, ZoneInfoNotFoundError

    for name in ("US/Eastern", "America/New_York", "UTC"):      # core/utils.py:29-35
        try:
            return ZoneInfo(name)
        except ZoneInfoNotFoundError:
            continue


def start_time_iso_from_game_id(game_id: str):
    """What `simulate_distribution` stores as `start_time_iso` when the market-odds file has no entry for the game
    (cli/run_distribution_simulator.py:243-250): the Eastern start time encoded in the id's -T suffix, or midnight of
    its date (core/utils.py:1009-1027 `game_id_to_dt`), ISO formatted; None if the id carries no date."""
    from datetime import datetime

    parts = parse_game_id(game_id)
    date, token = parts.get("date"), parts.get("time", "")
    dt = None
    if token.startswith("T"):
        digits = "".join(c for c in token.split("-")[0][1:] if c.isdigit())[:4]
        if len(digits) == 4:
            try:
                dt = datetime.strptime(f"{date} {digits}", "%Y-%m-%d %H%M")
            except Exception:
                dt = None
    if dt is None and date:
        try:
            dt = datetime.strptime(date, "%Y-%m-%d")
        except Exception:
            dt = None
    return dt.replace(tzinfo=_eastern()).isoformat() if dt else None


def _teams(game_id):
    """(away, home) abbreviations from 'YYYY-MM-DD-AWAY@HOME[-T....]' (core/utils.py:920-948)."""
    left, right = game_id.split("@", 1)
    return left.split("-")[-1].upper(), right.split("-")[0].upper()


def _cells(joint):
    """Non-empty cells of an [away][home] count matrix -> (home, away, count)."""
    a, h = np.nonzero(joint)
    return h.astype(np.float64), a.astype(np.float64), joint[a, h].astype(np.float64)


# ---------------------------------------------------------------------------------------------
# the chain
# ---------------------------------------------------------------------------------------------
def price_from_histogram(hist, game_id: str, calibration: dict | None = None, start_time_iso=None) -> dict:
    """`hist`: RunHistogram or a HIST_DTYPE record.  Returns the `simulate_distribution` document
    (run_distribution_simulator.py:825-843) computed from joint counts."""
    rec = hist.rec if hasattr(hist, "rec") else hist
    joint = rec["joint"]
    calibration = calibration or {}
    if not start_time_iso:      # no odds-file entry: derive it from the id, as the reference does (:248-250)
        start_time_iso = start_time_iso_from_game_id(game_id)
    team_scaling = calibration.get("team_total_scaling", {})
    hm, hsd = team_scaling.get("home_mean_factor", 1.0), team_scaling.get("home_std_factor", 1.0)
    am, asd = team_scaling.get("away_mean_factor", 1.0), team_scaling.get("away_std_factor", 1.0)
    segment_scaling = calibration.get("segment_scaling", {})
    away_abbr, home_abbr = _teams(game_id)

    # ---- full game ---------------------------------------------------------------------------
    H, A, W = _cells(joint[_abi.CAPS.index(None)])
    raw_totals, raw_diffs = Weighted(H + A, W), Weighted(H - A, W)          # :427-428
    raw_distributions = {
        "totals": {"mean": raw_totals.mean(), "std": raw_totals.std()},      # :429-432
        "run_diffs": {"std": raw_diffs.std()},
    }
    fg = BENCHMARK_TOTALS["full_game"]
    scaled_totals = scale_distribution(raw_totals, target_mean=fg["mean_total"], target_sd=fg["std_total"])  # :452-456
    scaled_diffs = scale_distribution(raw_diffs, target_sd=fg["std_total"])                                   # :457-460
    scaled_distributions = {
        "totals": {"mean": scaled_totals.mean(), "std": scaled_totals.std()},
        "run_diffs": {"std": scaled_diffs.std()},
    }
    home_scores = Weighted((scaled_totals.x + scaled_diffs.x) / 2, W)       # :485
    away_scores = Weighted((scaled_totals.x - scaled_diffs.x) / 2, W)       # :486
    home_scores = team_total_scaling(home_scores, hm, hsd)                   # :487
    away_scores = team_total_scaling(away_scores, am, asd)                   # :488
    home_scores = Weighted(_round1(home_scores.x), W)                        # :489
    away_scores = Weighted(_round1(away_scores.x), W)                        # :490

    total = Weighted(home_scores.x + away_scores.x, W)                       # :499
    run_pmf_rounded = total.pmf_rounded()                                    # :501
    run_pmf_raw = raw_totals.pmf_rounded()                                   # :502
    run_diff_pmf = scaled_diffs.pmf_rounded()                                # :503
    pmfs = {
        "totals": {"full_game": {"raw": raw_totals.pmf_rounded(), "scaled": run_pmf_rounded}},
        "spreads": {"full_game": {"raw": raw_diffs.pmf_rounded(), "scaled": run_diff_pmf}},
    }

    runline = {}                                                             # :521-548
    for line in [-2.5, -1.5, -0.5, 0.5, 1.5, 2.5]:
        p_home_minus = tail(run_diff_pmf, line, "over")
        runline[f"{home_abbr} -{line}"] = {"prob": round(p_home_minus, 4), "odds": to_american_odds(p_home_minus)}
        runline[f"{away_abbr} +{line}"] = {"prob": round(1 - p_home_minus, 4), "odds": to_american_odds(1 - p_home_minus)}
        p_away_minus = tail(run_diff_pmf, -line, "under")
        runline[f"{away_abbr} -{line}"] = {"prob": round(p_away_minus, 4), "odds": to_american_odds(p_away_minus)}
        runline[f"{home_abbr} +{line}"] = {"prob": round(1 - p_away_minus, 4), "odds": to_american_odds(1 - p_away_minus)}

    # compute_moneyline on the scaled, rounded scores (core/market_pricer.py:148-165): ties -> away
    p_home = home_scores.prob(home_scores.x > away_scores.x)
    ml_home, ml_away = round(p_home, 4), round(1 - p_home, 4)
    moneyline = {away_abbr: {"prob": ml_away, "odds": to_american_odds(ml_away)},   # :571-581
                 home_abbr: {"prob": ml_home, "odds": to_american_odds(ml_home)}}

    team_totals = {}                                                         # :584-603
    home_pmf, away_pmf = home_scores.pmf_rounded(), away_scores.pmf_rounded()
    for line in [1.5, 2.5, 3.5, 4.5, 5.5, 6.5]:
        for team, pmf in ((home_abbr, home_pmf), (away_abbr, away_pmf)):
            p_over = tail(pmf, line, "over")
            team_totals[f"{team} Over {line:.1f}"] = {"prob": round(p_over, 4), "odds": to_american_odds(p_over)}
            team_totals[f"{team} Under {line:.1f}"] = {"prob": round(1 - p_over, 4), "odds": to_american_odds(1 - p_over)}

    # ---- segments ----------------------------------------------------------------------------
    derivative_segments = {}
    summary_metrics = {"full_game": _summary(Weighted(H, W), Weighted(A, W))}   # raw scores, :810
    for seg_key, config in SEGMENT_CONFIGS.items():
        seg_id = seg_key.lower()
        cap = config["innings"]
        h, a, w = _cells(joint[_abi.CAPS.index(cap)])
        seg_total, seg_diff = Weighted(h + a, w), Weighted(h - a, w)
        name = SEG_NAME[seg_id]
        raw_distributions[f"totals_{name}"] = {"mean": seg_total.mean(), "std": seg_total.std()}   # :440-450
        raw_distributions[f"run_diffs_{name}"] = {"std": seg_diff.std()}
        seg_cal = segment_scaling.get(seg_id, {})
        tot_scaled = scale_distribution(seg_total, seg_cal.get("run_mean"), seg_cal.get("run_sd"))  # :660-664
        diff_scaled = scale_distribution(seg_diff, None, seg_cal.get("diff_sd"))                    # :665-669
        scaled_distributions[f"totals_{name}"] = {"mean": tot_scaled.mean(), "std": tot_scaled.std()}
        scaled_distributions[f"run_diffs_{name}"] = {"std": diff_scaled.std()}
        pmf_total_seg, pmf_diff_seg = tot_scaled.pmf_rounded(), diff_scaled.pmf_rounded()           # :685-686
        pmfs["totals"][name] = {"raw": seg_total.pmf_rounded(), "scaled": pmf_total_seg}
        pmfs["spreads"][name] = {"raw": seg_diff.pmf_rounded(), "scaled": pmf_diff_seg}

        seg = {"label": config["label"], "markets": {}}
        totals = {}
        for line in config.get("total_lines", []):                                                  # :712-717
            p = tail(pmf_total_seg, line, "over")
            totals[f"Over {line}"] = {"prob": round(p, 4), "fair_odds": to_american_odds(p)}
            totals[f"Under {line}"] = {"prob": round(1 - p, 4), "fair_odds": to_american_odds(1 - p)}
        seg["markets"]["totals"] = totals
        if seg_key == "F1":                                                                         # :719-725
            p = tail(pmf_total_seg, 0.5, "over")
            seg["markets"]["totals"] = {"Over 0.5": {"prob": p, "fair_odds": to_american_odds(p)},
                                        "Under 0.5": {"prob": 1 - p, "fair_odds": to_american_odds(1 - p)}}
        else:
            # compute_partial_derivatives (:965-982): raw segment scores, 3-dp rounding, pushes excluded
            n = float(w.sum())
            ml_h = round(float(w[h > a].sum()) / n, 3)
            ml_a = round(float(w[a > h].sum()) / n, 3)
            seg["markets"]["moneyline"] = {home_abbr: {"prob": ml_h, "fair_odds": to_american_odds(ml_h)},   # :728-733
                                           away_abbr: {"prob": ml_a, "fair_odds": to_american_odds(ml_a)}}
            spreads = {}
            for line in config.get("spread_lines", []):                                             # :736-762
                p_hm = tail(pmf_diff_seg, line, "over")
                spreads[f"{home_abbr} -{line}"] = {"prob": round(p_hm, 4), "fair_odds": to_american_odds(p_hm)}
                spreads[f"{away_abbr} +{line}"] = {"prob": round(1 - p_hm, 4), "fair_odds": to_american_odds(1 - p_hm)}
                p_am = tail(pmf_diff_seg, -line, "under")
                spreads[f"{away_abbr} -{line}"] = {"prob": round(p_am, 4), "fair_odds": to_american_odds(p_am)}
                spreads[f"{home_abbr} +{line}"] = {"prob": round(1 - p_am, 4), "fair_odds": to_american_odds(1 - p_am)}
            seg["markets"]["runline"] = spreads
            hs_seg = team_total_scaling(Weighted(h, w), hm, hsd)                                    # :767-772
            as_seg = team_total_scaling(Weighted(a, w), am, asd)
            seg_tt = {}
            hp, ap = hs_seg.pmf_rounded(), as_seg.pmf_rounded()
            for line in config.get("team_total_lines", []):                                         # :774-789
                for team, pmf in ((home_abbr, hp), (away_abbr, ap)):
                    p_over = tail(pmf, line, "over")
                    seg_tt[f"{team} Over {line:.1f}"] = {"prob": round(p_over, 4), "fair_odds": to_american_odds(p_over)}
                    seg_tt[f"{team} Under {line:.1f}"] = {"prob": round(1 - p_over, 4), "fair_odds": to_american_odds(1 - p_over)}
            seg["markets"]["team_totals"] = seg_tt
        derivative_segments[config["label"]] = seg
        summary_metrics[seg_id] = _summary(team_total_scaling(Weighted(h, w), hm, hsd),
                                           team_total_scaling(Weighted(a, w), am, asd))            # :802-808

    # ---- extract_universal_markets (:96-192) ---------------------------------------------------
    entries = []
    for team, obj in moneyline.items():
        entries.append(_entry("h2h", team, obj["prob"], obj["odds"]))
    for team in (away_abbr, home_abbr):
        for line in [-2.5, -1.5, -0.5, 0.5, 1.5, 2.5]:
            data = runline.get(f"{team} {'+' if line > 0 else ''}{line}", {})
            if "prob" in data and "odds" in data:
                entries.append(_entry("spreads", f"{team} {line:+.1f}", data["prob"], data["odds"]))
    for line in [x / 2 for x in range(13, 26)]:
        for side in ("Over", "Under"):
            prob = tail(run_pmf_rounded, line, side.lower())
            if line % 1 == 0:
                adj_over, adj_under = adjust_for_push(tail(run_pmf_rounded, line, "over"), tail(run_pmf_rounded, line, "under"))
                prob = adj_over if side == "Over" else adj_under
            entries.append(_entry("totals", f"{side} {line}", prob, to_american_odds(prob)))
    for label, obj in team_totals.items():
        entries.append(_entry("team_totals", label, obj["prob"], obj["odds"]))
    for seg_label, seg in derivative_segments.items():
        suffix = SEGMENT_TO_MARKET_KEY[seg_label]
        for mkt_type, options in seg["markets"].items():
            prefix = {"moneyline": "h2h", "runline": "spreads", "total": "totals", "totals": "totals",
                      "team_totals": "team_totals"}.get(mkt_type)
            if not prefix:
                continue
            for raw_label, sim in options.items():
                entries.append(_entry(f"{prefix}_{suffix}", _normalise_segment_label(raw_label, prefix), sim["prob"], sim["fair_odds"]))

    return {
        "home_score": float(np.dot(H, W) / W.sum()),
        "away_score": float(np.dot(A, W) / W.sum()),
        "start_time_iso": start_time_iso,
        "run_distribution": run_pmf_rounded,
        "run_distribution_raw": run_pmf_raw,
        "run_diff_distribution": run_diff_pmf,
        "pmfs": pmfs,
        "raw_distributions": raw_distributions,
        "scaled_distributions": scaled_distributions,
        "markets": entries,
        "summary_metrics": summary_metrics,
        "n_simulations": int(W.sum()),
    }


def _normalise_segment_label(raw_label, prefix):
    """normalize_label_for_odds (core/utils.py:446-479) restricted to abbreviated team labels."""
    raw = raw_label.strip()
    if prefix == "spreads":
        team, point = raw.replace("+", " +").replace("-", " -").split()[:2]
        return f"{team} {float(point):+.1f}"
    if prefix == "totals":
        side, point = raw.split()[:2]
        return f"{side.title()} {float(point):.1f}"
    return raw


def _summary(home: Weighted, away: Weighted):
    """summarize_distribution, run_distribution_simulator.py:879-905."""
    diffs, totals = Weighted(home.x - away.x, home.w), Weighted(home.x + away.x, home.w)
    return {"mean_total": totals.mean(), "std_total": totals.std(), "mean_diff": diffs.mean(), "std_diff": diffs.std()}


def write_sim_json(doc: dict, target_path: str) -> None:
    """Atomic write, run_distribution_simulator.py:846-850 (json cannot carry float keys: PMF keys
    become strings exactly as json.dump renders the reference's float keys)."""
    os.makedirs(os.path.dirname(target_path) or ".", exist_ok=True)
    with tempfile.NamedTemporaryFile("w", dir=os.path.dirname(target_path) or ".", delete=False, suffix=".tmp") as f:
        json.dump(doc, f, indent=2)
        tmp = f.name
    os.replace(tmp, target_path)


def write_table_snapshot(doc: dict, path: str = SNAPSHOT_PATH) -> None:
    """The side file `simulate_distribution` rewrites after every game (cli/run_distribution_simulator.py:853-859):
    {"<market>:<side>": {"fair_odds", "ev_percent"}} of the game just priced."""
    snap = {f"{e['market']}:{e['side']}": {"fair_odds": e.get("fair_odds"), "ev_percent": e.get("ev_percent")} for e in doc["markets"]}
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    with open(path, "w") as f:
        json.dump(snap, f, indent=2)
