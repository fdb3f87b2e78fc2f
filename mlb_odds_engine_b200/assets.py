"""Pitcher-side asset arithmetic that feeds the sampling tables.

The hot path reads `pitcher["hr_pa"]["hr_pa_projected"]` (core/pa_simulator.py:126, default 0.046); Stuff+,
Location+, barrel rate, exit velocity, launch angle, K%, BB%, FIP and TBF act on the simulation ONLY through that
number, produced upstream by `project_hr_pa` (core/project_hr_pa.py:46-130).  A sweep over those knobs
(BASELINE config 5) therefore needs the projection on the host, vectorised over grid points.  This module restates

  * `project_hr_pa`            core/project_hr_pa.py:46-130   (scalar, same dict in / dict out)
  * `project_hr_pa_batch`      the same arithmetic over numpy columns (float64, same operation order, so the two
                               agree bit for bit; tests/test_assets.py pins both against reference outputs)
  * `reliever_from_stats`      assets/bullpen_utils.py:62-98  (one bullpen entry from a stats row)
  * `select_bullpen`           assets/bullpen_utils.py:100-104 (score sort, top `max_relievers`)
  * `structure_pitcher`        core/game_asset_builder.py:206-240 (starter dict; un-enriched starters get the string
                               "N/A" for EV / LA, which the table builder rejects exactly where the reference raises)

Name normalisation, depth-chart files, scraping of probable pitchers and lineups are the reference's data plane and
stay there: callers pass already-keyed stats rows.
"""
from __future__ import annotations

import math

import numpy as np

DEFAULT_SHRINK_BASE = 300.0      # core/project_hr_pa.py:4
REFERENCE_TBF = 650.0            # core/project_hr_pa.py:5

# weights of the linear model, in the reference's summation order (core/project_hr_pa.py:91-107)
_BASE = 0.0011
_W = (("barrel", 0.020), ("nev", 0.018), ("nla", 0.018), ("xSLG", 0.010), ("xwOBAcon", 0.006),
      ("sweet_spot_pct", 0.010), ("xSLG_diff", 0.005), ("wOBAdiff", 0.005), ("one_minus_k", 0.008),
      ("one_minus_bb", 0.005), ("norm_stuff", 0.004), ("location", 0.010), ("norm_fip", 0.005))


# League prior of HR/PA by role: (base, ((Stuff+ bound, is the bound a floor?, multiplier), ... first match wins)).
# core/project_hr_pa.py:17-37.  One table drives both the scalar and the batched projection.
_LEAGUE_PRIOR = {
    "RP": (0.020, ((120, True, 0.85), (110, True, 0.90), (95, False, 1.10))),
    "SP": (0.030, ((110, True, 0.90), (100, True, 0.95), (90, False, 1.10))),
}


def infer_league_avg_hr_pa(role: str = "SP", stuff_plus: float = 100.0) -> float:
    """League-average HR/PA for the pitcher's role, nudged by his Stuff+ band (core/project_hr_pa.py:17-37)."""
    base, bands = _LEAGUE_PRIOR["RP" if role.upper() == "RP" else "SP"]
    for bound, is_floor, mult in bands:
        if (stuff_plus >= bound) if is_floor else (stuff_plus <= bound):
            return base * mult
    return base


def dynamic_shrink_n(tbf) -> float:
    """Pseudo-count of the shrinkage towards the league prior: 300 batters faced at the reference workload of 650,
    inversely proportional to the pitcher's own TBF; a TBF that is not a number keeps the base, a non-positive one
    counts as a single batter (core/project_hr_pa.py:7-15)."""
    try:
        faced = float(tbf)
    except (TypeError, ValueError):
        return DEFAULT_SHRINK_BASE
    return DEFAULT_SHRINK_BASE * (REFERENCE_TBF / (faced if faced > 0 else 1))


def _float_or(value, fallback: float) -> float:
    """core/project_hr_pa.py:39-44 (`safe_float`): NaN and unparsable values fall back."""
    try:
        v = float(value)
    except (ValueError, TypeError):
        return fallback
    return fallback if math.isnan(v) else v


def project_hr_pa(pitcher: dict, mutate: bool = True) -> dict:
    """HR/PA projection of one pitcher dict (core/project_hr_pa.py:46-130).

    With `mutate` (the reference's behaviour) the cleaned `barrel_batted_rate`, `exit_velocity`, `launch_angle` are
    written back into `pitcher`.  Raises TypeError for non-dict input or list-valued contact fields, like the
    reference (:47-55)."""
    if not isinstance(pitcher, dict):
        raise TypeError("Expected dict for pitcher_data")
    for key in ("exit_velocity_avg", "launch_angle_avg", "barrel_batted_rate"):
        if isinstance(pitcher.get(key), list):
            raise TypeError(f"Invalid list found in field '{key}'")

    hr = pitcher.get("HR", 0)
    tbf = pitcher.get("TBF", 1)
    ip = pitcher.get("IP", 1)
    empirical = hr / tbf if tbf else 0.0

    barrel = _float_or(pitcher.get("barrel_batted_rate"), 0.06)
    ev = _float_or(pitcher.get("exit_velocity_avg"), 88.5)
    la = _float_or(pitcher.get("launch_angle_avg"), 12.5)
    if mutate:
        pitcher["barrel_batted_rate"], pitcher["exit_velocity"], pitcher["launch_angle"] = barrel, ev, la

    stuff = pitcher.get("stuff_plus", 100.0)
    role = pitcher.get("role", "SP")
    league = pitcher.get("league_avg_hr_pa", infer_league_avg_hr_pa(role, stuff))
    terms = {
        "barrel": barrel,
        "nev": (ev - 85) / 1000,
        "nla": (la - 12) / 100,
        "xSLG": pitcher.get("xSLG", 0.0),
        "xwOBAcon": pitcher.get("xwOBAcon", 0.0),
        "sweet_spot_pct": pitcher.get("sweet_spot_pct", 0.32),
        "xSLG_diff": pitcher.get("xSLG_diff", 0.0),
        "wOBAdiff": pitcher.get("wOBAdiff", 0.0),
        "one_minus_k": 1 - pitcher.get("K_pct", 0.0),
        "one_minus_bb": 1 - pitcher.get("BB_pct", 0.0),
        "norm_stuff": stuff / 100.0,
        "location": pitcher.get("location_plus", 100.0) / 100.0,
        "norm_fip": 1 - pitcher.get("FIP", 4.0) / 10.0,
    }
    model = _BASE
    for name, w in _W:
        model = model + w * terms[name]
    model *= 1.15
    if tbf >= 900:
        model *= 0.87
    elif tbf >= 700:
        model *= 0.93
    elif tbf >= 500:
        model *= 0.97
    model = min(max(model, 0.005), 0.07)

    shrink = dynamic_shrink_n(tbf)
    smoothed = (tbf * model + shrink * league) / (tbf + shrink)
    return {
        "pitcher_id": pitcher.get("pitcher_id", None),
        "role": role,
        "hr_pa_projected": smoothed,
        "empirical_hr_pa": empirical,
        "model_estimate_hr_pa": model,
        "hr_per_9": smoothed * (tbf / ip) * 9 if ip else 0.0,
        "league_avg_hr_pa": league,
    }


_BATCH_DEFAULTS = {
    "barrel_batted_rate": 0.06, "exit_velocity_avg": 88.5, "launch_angle_avg": 12.5, "xSLG": 0.0, "xwOBAcon": 0.0,
    "sweet_spot_pct": 0.32, "xSLG_diff": 0.0, "wOBAdiff": 0.0, "K_pct": 0.0, "BB_pct": 0.0, "stuff_plus": 100.0,
    "location_plus": 100.0, "FIP": 4.0, "TBF": 1.0,
}


def project_hr_pa_batch(role: str = "SP", **columns) -> np.ndarray:
    """`hr_pa_projected` for a whole grid at once.  `columns` are broadcastable float64 arrays named like the
    pitcher-dict keys `project_hr_pa` reads (missing ones take the reference's defaults; NaN in the three contact
    fields falls back like `safe_float`).  Same operations in the same order as the scalar function, so every
    element equals `project_hr_pa({...})["hr_pa_projected"]` exactly."""
    unknown = set(columns) - set(_BATCH_DEFAULTS) - {"league_avg_hr_pa"}
    if unknown:
        raise KeyError(f"unknown pitcher field(s) {sorted(unknown)}")
    c = {k: np.asarray(columns.get(k, d), dtype=np.float64) for k, d in _BATCH_DEFAULTS.items()}
    for k in ("barrel_batted_rate", "exit_velocity_avg", "launch_angle_avg"):
        c[k] = np.where(np.isnan(c[k]), _BATCH_DEFAULTS[k], c[k])
    shape = np.broadcast_shapes(*(v.shape for v in c.values()),
                                *([np.shape(columns["league_avg_hr_pa"])] if "league_avg_hr_pa" in columns else []))
    c = {k: np.broadcast_to(v, shape) for k, v in c.items()}
    stuff, tbf = c["stuff_plus"], c["TBF"]

    if "league_avg_hr_pa" in columns:
        league = np.broadcast_to(np.asarray(columns["league_avg_hr_pa"], dtype=np.float64), shape)
    else:
        base, bands = _LEAGUE_PRIOR["RP" if role.upper() == "RP" else "SP"]
        league = np.select([(stuff >= b) if floor else (stuff <= b) for b, floor, _ in bands], [base * m for _, _, m in bands], base)

    terms = {
        "barrel": c["barrel_batted_rate"], "nev": (c["exit_velocity_avg"] - 85) / 1000,
        "nla": (c["launch_angle_avg"] - 12) / 100, "xSLG": c["xSLG"], "xwOBAcon": c["xwOBAcon"],
        "sweet_spot_pct": c["sweet_spot_pct"], "xSLG_diff": c["xSLG_diff"], "wOBAdiff": c["wOBAdiff"],
        "one_minus_k": 1 - c["K_pct"], "one_minus_bb": 1 - c["BB_pct"], "norm_stuff": stuff / 100.0,
        "location": c["location_plus"] / 100.0, "norm_fip": 1 - c["FIP"] / 10.0,
    }
    model = np.full(shape, _BASE)
    for name, w in _W:
        model = model + w * terms[name]
    model = model * 1.15
    model = model * np.select([tbf >= 900, tbf >= 700, tbf >= 500], [0.87, 0.93, 0.97], 1.0)
    model = np.minimum(np.maximum(model, 0.005), 0.07)
    with np.errstate(divide="ignore"):
        shrink = np.where(tbf <= 0, DEFAULT_SHRINK_BASE * (REFERENCE_TBF / 1),
                          DEFAULT_SHRINK_BASE * (REFERENCE_TBF / np.where(tbf <= 0, 1.0, tbf)))
    return (tbf * model + shrink * league) / (tbf + shrink)


def _safe_float(val, fallback=0.0) -> float:
    """assets/bullpen_utils.py:9-16: lists contribute their first item, anything unparsable the fallback
    (NaN passes through, unlike `project_hr_pa`'s helper)."""
    try:
        if isinstance(val, list):
            val = val[0] if val else fallback
        return float(val)
    except Exception:
        return fallback


_RELIEVER_FIELDS = (("k_rate", 0.22), ("bb_rate", 0.08), ("hr_fb_rate", 0.10), ("stuff_plus", 100), ("command_plus", 100),
                    ("location_plus", 100), ("barrel_batted_rate", 0.06), ("exit_velocity_avg", 88.0),
                    ("launch_angle_avg", 13.0), ("HR", 0), ("TBF", 1))


def reliever_from_stats(name: str, stats: dict) -> dict | None:
    """One bullpen entry from a pitcher-stats row (assets/bullpen_utils.py:62-98).  Returns None where the
    reference skips the pitcher (no row, missing / NaN Stuff+ or HR/FB)."""
    if not isinstance(stats, dict) or not stats:
        return None
    hr_fb = stats.get("hr_fb_rate", 9.5)
    if hr_fb is not None and hr_fb > 1:
        hr_fb = hr_fb / 100.0
    stuff = stats.get("stuff_plus")
    if stuff is None or (isinstance(stuff, float) and math.isnan(stuff)):
        return None
    if hr_fb is None or (isinstance(hr_fb, float) and math.isnan(hr_fb)):
        return None
    # fields copied from the stats row with the reference's fallbacks, in its key order (assets/bullpen_utils.py:77-92)
    r = {"name": name, "throws": stats.get("throws", "R")}
    for key, fallback in _RELIEVER_FIELDS:
        r[key] = _safe_float({"hr_fb_rate": hr_fb, "stuff_plus": stuff}.get(key, stats.get(key)), fallback)
    r["TBF"] = max(r["TBF"], 1)
    r["role"] = "RP"
    r["hr_pa"] = project_hr_pa(r)
    r["score"] = r["stuff_plus"] + r["k_rate"] * 100 - r["bb_rate"] * 100
    return r


def select_bullpen(entries, max_relievers: int = 6, exclude=()) -> list[dict]:
    """(name, stats) pairs → the team's bullpen: build entries, drop today's starters (`exclude`, already-normalised
    names compared verbatim), stable sort by leverage score descending, keep `max_relievers`
    (assets/bullpen_utils.py:50-104)."""
    skip = set(exclude)
    pen = [r for name, stats in entries if name not in skip and (r := reliever_from_stats(name, stats)) is not None]
    pen.sort(key=lambda r: r["score"], reverse=True)
    return pen[:max_relievers]


def structure_pitcher(name: str, stats: dict | None) -> dict:
    """Starter dict as `build_game_assets` makes it (core/game_asset_builder.py:206-240).  Without a stats row the
    projection is `{"hr_pa_projected": None}` and EV / LA are "N/A": `simulate_game` then raises TypeError, and so
    does `tables.build_table`."""
    if isinstance(stats, list):
        stats = None
    hr_fb = stats.get("hr_fb_rate") if stats else None
    if hr_fb is not None and hr_fb > 1:
        hr_fb = hr_fb / 100.0
    missing = hr_fb is None or (isinstance(hr_fb, float) and math.isnan(hr_fb))

    def g(key, default):
        return stats.get(key, default) if stats else default

    return {
        "name": name,
        "throws": "R",
        "stuff_plus": g("stuff_plus", 100), "command_plus": g("command_plus", 100),
        "location_plus": g("location_plus", 100),
        "k_rate": g("k_rate", 0.22), "bb_rate": g("bb_rate", 0.08),
        "hr_fb_rate": "N/A" if missing else f"{hr_fb:.1%}",
        "hr_pa": project_hr_pa(stats) if stats else {"hr_pa_projected": None},
        "iso_allowed": g("iso_allowed", 0.140),
        "exit_velocity_avg": g("exit_velocity_avg", "N/A"), "launch_angle_avg": g("launch_angle_avg", "N/A"),
        "barrel_batted_rate": g("barrel_batted_rate", "N/A"),
        "HR": g("HR", "N/A"), "TBF": g("TBF", "N/A"), "IP": g("IP", "N/A"),
        "enriched": g("enriched", False),
    }
