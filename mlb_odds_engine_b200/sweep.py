"""Parameter-sensitivity sweeps on the GPU path — the shape of `tools/test_stuff_plus_sensitivity.py:51-77`
and its siblings (one pitcher dict per grid value, N games without bullpens, outcome rates read from
`PA_RECAP_LOG`) as ONE multi-matchup launch: every grid point is a matchup, its outcome counters are
`mlbsim_hist.pa` (BASELINE config 5: 1 024 points x 10^5 sims).

Which knobs move results: the hot path reads pitcher `k_rate, bb_rate, hr_pa.hr_pa_projected,
exit_velocity_avg, launch_angle_avg, fielder_rating` and batter `k_rate, bb_rate, speed` only; ISO,
HR/FB and Stuff+ act solely through `hr_pa_projected` upstream, so a sweep over
them changes nothing unless the caller maps them to `hr_pa_projected` first.
"""
from __future__ import annotations

import copy
import itertools

import numpy as np

from . import _abi, assets, tables
from .simulator import GpuSimulator

# pitcher fields that reach the simulation only through `project_hr_pa` (core/project_hr_pa.py:57-89)
UPSTREAM_FIELDS = frozenset({"stuff_plus", "location_plus", "barrel_batted_rate", "xSLG", "xwOBAcon", "sweet_spot_pct",
                             "xSLG_diff", "wOBAdiff", "K_pct", "BB_pct", "FIP", "TBF", "league_avg_hr_pa"})


def grid_matchups(base: dict, side: str = "home", project: bool = False, **axes) -> tuple[list[dict], list[dict]]:
    """Cartesian grid over pitcher fields of `base[f"{side}_pitcher"]`; `hr_pa_projected` addresses the
    nested `hr_pa` dict.  With `project=True` every grid point's `hr_pa` is recomputed with
    `assets.project_hr_pa` after the fields are set — what `build_game_assets` does for a real pitcher
    (core/game_asset_builder.py:214) — so upstream knobs (Stuff+, Location+, barrel rate, ...) move the
    tables; EV / LA then act twice, through the projection and through BIP resolution, as in the reference.
    Returns (matchups, points)."""
    names = list(axes)
    if project and "hr_pa_projected" in names:
        raise ValueError("hr_pa_projected cannot be a grid axis when project=True (it is recomputed)")
    out, points = [], []
    for combo in itertools.product(*(axes[n] for n in names)):
        m = copy.deepcopy(base)
        p = m[f"{side}_pitcher"]
        for n, v in zip(names, combo):
            if n == "hr_pa_projected":
                p.setdefault("hr_pa", {})["hr_pa_projected"] = float(v)
            else:
                p[n] = float(v)
        if project:
            p["hr_pa"] = assets.project_hr_pa(p, mutate=False)
        out.append(m)
        points.append(dict(zip(names, (float(v) for v in combo))))
    return out, points


def grid_tables(base: dict, side: str = "home", project: bool = False, use_noise: bool = True, alias_fn=None, **axes):
    """The device tables of `grid_matchups(base, side, project, **axes)` without materialising one dict per grid
    point: the base matchup is parsed once, the swept pitcher's numbers are written into stacked arrays and
    `tables.tables_from_arrays` builds every table in one pass (1 024 points: ~0.35 s instead of ~1.2 s).
    Byte-identical to `tables.build_tables(grid_matchups(...)[0])`.  `alias_fn`: e.g. `Context.alias_rows`, the device
    builder of the alias rows (same bytes).  Returns (TABLE_DTYPE[n], points)."""
    A, points = _grid_arrays(base, side, project, use_noise, **axes)
    return tables.tables_from_arrays(A, alias_fn=alias_fn), points


def grid_params(base: dict, side: str = "home", project: bool = False, use_noise: bool = True, **axes):
    """The same grid as `mlbsim_params` records for the DEVICE table builder (`GpuSimulator.load_params`): a few array
    copies on the host, the tables are built in HBM.  Returns (PARAMS_DTYPE[n], points)."""
    A, points = _grid_arrays(base, side, project, use_noise, **axes)
    return tables.params_from_arrays(A), points


def _grid_arrays(base: dict, side: str, project: bool, use_noise: bool, **axes):
    """Stacked per-matchup arrays (`tables.stack_matchups` layout) of a cartesian grid over one pitcher's fields."""
    names = list(axes)
    if project and "hr_pa_projected" in names:
        raise ValueError("hr_pa_projected cannot be a grid axis when project=True (it is recomputed)")
    combos = list(itertools.product(*(axes[n] for n in names)))
    points = [dict(zip(names, (float(v) for v in c))) for c in combos]
    G = len(combos)
    M = tables.as_matchup(base, use_noise)
    A = tables.stack_matchups([M])
    A = {k: np.repeat(v, G, axis=0) for k, v in A.items()}
    s = 0 if side == "home" else 1            # the home staff pitches to the away lineup: batting side 0
    pit = dict(M.staffs[s][0])
    col = {n: np.array([c[j] for c in combos], dtype=np.float64) for j, n in enumerate(names)}
    if "k_rate" in col:
        A["pit_k"][:, s, 0] = col["k_rate"]
    if "bb_rate" in col:
        A["pit_bb"][:, s, 0] = col["bb_rate"]
    if project:
        cols = {}
        for f in assets._BATCH_DEFAULTS:
            if f in col:
                cols[f] = col[f]
            elif f in pit:
                cols[f] = np.full(G, assets._float_or(pit[f], float("nan")) if f in ("barrel_batted_rate", "exit_velocity_avg", "launch_angle_avg")
                                  else pit[f], dtype=np.float64)
        if "league_avg_hr_pa" in col or "league_avg_hr_pa" in pit:
            cols["league_avg_hr_pa"] = col.get("league_avg_hr_pa", np.full(G, pit.get("league_avg_hr_pa", 0.0)))
        A["pit_hr"][:, s, 0] = assets.project_hr_pa_batch(role=pit.get("role", "SP"), **cols) * M.weather_hr
    elif "hr_pa_projected" in col:
        A["pit_hr"][:, s, 0] = col["hr_pa_projected"] * M.weather_hr
    contact = [f for f in ("exit_velocity_avg", "launch_angle_avg", "fielder_rating") if f in col]
    if contact:   # the BIP hit probabilities of this pitcher change too (core/bip_resolution.py:20-57)
        L = M.lineup_len[s]
        cache = {}
        for g in range(G):
            ev = col["exit_velocity_avg"][g] if "exit_velocity_avg" in col else pit.get("exit_velocity_avg")
            la = col["launch_angle_avg"][g] if "launch_angle_avg" in col else pit.get("launch_angle_avg")
            fr = col["fielder_rating"][g] if "fielder_rating" in col else pit.get("fielder_rating", 50)
            key = (ev, la, fr)
            if key not in cache:
                cache[key] = np.array([[tables.bip_hit_probability(t, ev, la, M.speed[s, b], fr) for t in tables.BIP_TYPES]
                                       for b in range(L)])
            A["hit_prob"][g, s, 0, :L] = cache[key]
    # (axes the hot path never reads — Stuff+, ISO, HR/FB without project=True — leave every table equal to the base
    # table, as in the reference)
    return A, points


def sensitivity_matchups(field: str, values, project: bool = False) -> list[dict]:
    """The matchups of `tools/test_<field>_sensitivity.py:10-66`: nine copies of the league-average batter on
    both sides, the same template pitcher (K 22.5 %, BB 8.2 %, no `hr_pa` ⇒ the 0.046 default, no EV / LA) on both
    mounds with `field` set to each value, neutral env, no bullpens.  In the reference these sweeps are flat for
    stuff_plus / command_plus / location_plus / hr_fb_rate / iso_allowed because nothing on the hot path reads
    them; `project=True` routes them through the HR/PA projection instead."""
    batter = {"name": "avg_batter", "handedness": "R", "k_rate": 0.225, "bb_rate": 0.082, "iso": 0.145,
              "avg": 0.245, "woba": 0.320}
    env = {"park_hr_mult": 1.0, "single_mult": 1.0, "weather_hr_mult": 1.0, "adi_mult": 1.0,
           "umpire": {"K": 1.0, "BB": 1.0}}
    out = []
    for v in values:
        pit = {"name": f"{field}_{v}", "throws": "R", "stuff_plus": 100, "command_plus": 100, "location_plus": 100,
               "k_rate": 0.225, "bb_rate": 0.082, "hr_fb_rate": 0.115, "iso_allowed": 0.140}
        pit[field] = v
        if project:
            pit["hr_pa"] = assets.project_hr_pa(pit, mutate=False)
        out.append({"game_id": f"{field}={v}", "home_lineup": [dict(batter) for _ in range(9)],
                    "away_lineup": [dict(batter) for _ in range(9)], "home_pitcher": dict(pit),
                    "away_pitcher": dict(pit), "env": env, "home_bullpen": None, "away_bullpen": None})
    return out


def pooled_rates(row: dict) -> dict:
    """HOME + AWAY outcome percentages as the tools print them (tools/test_stuff_plus_sensitivity.py:71-85)."""
    n = {o: row["counts"]["HOME"][o] + row["counts"]["AWAY"][o] for o in _abi.OUTCOMES}
    total = sum(n.values())
    return {o: (100.0 * c / total if total else 0.0) for o, c in n.items()}


def run_sweep(matchups, n_sims: int = 100_000, seed: int = 0, sim: GpuSimulator | None = None,
              use_noise: bool = True) -> list[dict]:
    """Per grid point: outcome counts and rates per batting side (what the tools print from PA_RECAP_LOG,
    tools/test_stuff_plus_sensitivity.py:71-77), mean runs, run SDs and home-win probability — from the device-side
    summaries (`mlbsim_summarize`), so a 1 024-point sweep moves 320 KB over PCIe instead of 168 MB."""
    own = sim is None
    sim = sim or GpuSimulator(0)
    if isinstance(matchups, np.ndarray) and matchups.dtype == _abi.TABLE_DTYPE:   # prebuilt tables (`grid_tables`)
        sim.load_tables(matchups)
    elif isinstance(matchups, np.ndarray) and matchups.dtype == _abi.PARAMS_DTYPE:  # raw numbers (`grid_params`): built in HBM
        sim.load_params(matchups)
    else:
        sim.load_matchups(matchups, use_noise=use_noise)
    rows = [summary_row(r) for r in sim.simulate_summary(n_sims, seed=seed)]
    if own:
        sim.close()
    return rows


def summary_row(r) -> dict:
    """One SUMMARY_DTYPE record -> the row `run_sweep` reports."""
    n = max(int(r["n_games"]), 1)
    pa = r["pa"][:, :7].astype(np.float64)
    away_mean, home_mean = (float(r["sum_runs"][_abi.N_CAPS - 1][s]) / n for s in (0, 1))
    away_sd, home_sd = (max(float(r["sum_sq"][s]) / n - m * m, 0.0) ** 0.5 for s, m in ((0, away_mean), (1, home_mean)))
    return {
        "counts": {team: {o: int(pa[s, i]) for i, o in enumerate(_abi.OUTCOMES)} for s, team in ((0, "AWAY"), (1, "HOME"))},
        "rates": {team: {o: float(pa[s, i] / max(pa[s].sum(), 1.0)) for i, o in enumerate(_abi.OUTCOMES)}
                  for s, team in ((0, "AWAY"), (1, "HOME"))},
        "pa_per_game": float(pa.sum() / n),
        "home_runs_mean": home_mean, "away_runs_mean": away_mean, "home_runs_sd": home_sd, "away_runs_sd": away_sd,
        "home_win": float(int(r["home_wins"]) / n), "away_win": float(int(r["away_wins"]) / n),
        "innings_mean": float(int(r["sum_innings"]) / n),
        "segment_runs_mean": {f"f{cap}": {"away": float(r["sum_runs"][c][0]) / n, "home": float(r["sum_runs"][c][1]) / n}
                              for c, cap in enumerate(_abi.CAPS[:-1])},
    }
