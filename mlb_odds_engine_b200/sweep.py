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

from . import _abi
from .simulator import GpuSimulator


def grid_matchups(base: dict, side: str = "home", **axes) -> tuple[list[dict], list[dict]]:
    """Cartesian grid over pitcher fields of `base[f"{side}_pitcher"]`; `hr_pa_projected` addresses the
    nested `hr_pa` dict.  Returns (matchups, points)."""
    names = list(axes)
    out, points = [], []
    for combo in itertools.product(*(axes[n] for n in names)):
        m = copy.deepcopy(base)
        p = m[f"{side}_pitcher"]
        for n, v in zip(names, combo):
            if n == "hr_pa_projected":
                p.setdefault("hr_pa", {})["hr_pa_projected"] = float(v)
            else:
                p[n] = float(v)
        out.append(m)
        points.append(dict(zip(names, (float(v) for v in combo))))
    return out, points


def run_sweep(matchups: list[dict], n_sims: int = 100_000, seed: int = 0, sim: GpuSimulator | None = None,
              use_noise: bool = True) -> list[dict]:
    """Per grid point: outcome rates per batting side (what the tools print from PA_RECAP_LOG,
    tools/test_stuff_plus_sensitivity.py:71-77), mean runs and home-win probability."""
    own = sim is None
    sim = sim or GpuSimulator(0)
    sim.load_matchups(matchups, use_noise=use_noise)
    hists = sim.simulate(n_sims, seed=seed)
    rows = []
    for h in hists:
        pa = h.rec["pa"][:, :7].astype(np.float64)
        home_mean, away_mean = h.mean_scores()
        rows.append({
            "rates": {team: {o: float(pa[s, i] / max(pa[s].sum(), 1.0)) for i, o in enumerate(_abi.OUTCOMES)}
                      for s, team in ((0, "AWAY"), (1, "HOME"))},
            "pa_per_game": float(pa.sum() / h.n_games),
            "home_runs_mean": home_mean, "away_runs_mean": away_mean, "home_win": h.home_win_prob(),
        })
    if own:
        sim.close()
    return rows
