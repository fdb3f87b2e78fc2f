"""Synthetic lineups / pitchers / bullpens in the reference's dict schema.

There is no network and `data/Pitchers.csv` is absent from the reference checkout, so every
benchmark and parity case runs on seeded synthetic players.  The dict keys follow what the
reference's asset builders emit (`core/game_asset_builder.py:206-267`,
`assets/bullpen_utils.py:77-95`, `assets/stats_loader.py:119-145`) and the distributions follow
SURVEY.md §8(d); only the keys the hot path reads influence results (batter `k_rate, bb_rate, speed`;
pitcher `k_rate, bb_rate, hr_pa.hr_pa_projected, exit_velocity_avg, launch_angle_avg, fielder_rating`).
"""
from __future__ import annotations

import numpy as np

SEED = 20250404

TEAMS = [
    "ARI", "ATL", "BAL", "BOS", "CHC", "CWS", "CIN", "CLE", "COL", "DET",
    "HOU", "KC", "LAA", "LAD", "MIA", "MIL", "MIN", "NYM", "NYY", "OAK",
    "PHI", "PIT", "SD", "SF", "SEA", "STL", "TB", "TEX", "TOR", "WSH",
]


def _clip(x, lo, hi):
    return float(min(max(x, lo), hi))


def make_batter(rng: np.random.Generator, name: str) -> dict:
    k = _clip(rng.normal(0.225, 0.045), 0.08, 0.40)
    bb = _clip(rng.normal(0.082, 0.025), 0.02, 0.18)
    iso = _clip(rng.normal(0.150, 0.045), 0.03, 0.33)
    avg = _clip(rng.normal(0.248, 0.025), 0.17, 0.33)
    return {
        "name": name,
        "handedness": "R" if rng.random() < 0.6 else "L",
        "k_rate": k,
        "bb_rate": bb,
        "iso": iso,
        "avg": avg,
        "woba": _clip(0.32 + 0.9 * (avg - 0.248) + 0.5 * (iso - 0.15) + 0.4 * (bb - 0.082), 0.24, 0.44),
        "speed": float(rng.uniform(25.0, 75.0)),
    }


def make_pitcher(rng: np.random.Generator, name: str, role: str = "SP") -> dict:
    k = _clip(rng.normal(0.232 if role == "SP" else 0.245, 0.04), 0.10, 0.40)
    bb = _clip(rng.normal(0.075 if role == "SP" else 0.088, 0.02), 0.02, 0.16)
    hr_pa = _clip(rng.normal(0.030, 0.008), 0.005, 0.07)
    tbf = float(rng.integers(500, 800)) if role == "SP" else float(rng.integers(120, 300))
    return {
        "name": name,
        "throws": "R" if rng.random() < 0.7 else "L",
        "role": role,
        "k_rate": k,
        "bb_rate": bb,
        "stuff_plus": float(rng.normal(100, 10)),
        "command_plus": float(rng.normal(100, 8)),
        "location_plus": float(rng.normal(100, 8)),
        "hr_fb_rate": _clip(rng.normal(0.115, 0.02), 0.05, 0.2),
        "iso_allowed": _clip(rng.normal(0.140, 0.02), 0.08, 0.22),
        "barrel_batted_rate": _clip(rng.normal(0.065, 0.015), 0.02, 0.12),
        "exit_velocity_avg": float(rng.normal(88.5, 1.8)),
        "launch_angle_avg": float(rng.normal(13.0, 3.5)),
        "fielder_rating": float(rng.uniform(30.0, 70.0)),
        "HR": float(round(hr_pa * tbf)),
        "TBF": tbf,
        "IP": tbf / 4.2,
        "hr_pa": {
            "hr_pa_projected": hr_pa,
            "empirical_hr_pa": hr_pa,
            "model_estimate_hr_pa": hr_pa,
            "hr_per_9": hr_pa * 4.2 * 9,
        },
    }


def make_team(rng: np.random.Generator, abbr: str, n_relievers: int = 6, lineup_len: int = 9) -> dict:
    lineup = [make_batter(rng, f"{abbr} Batter {i + 1}") for i in range(lineup_len)]
    starter = make_pitcher(rng, f"{abbr} Starter", "SP")
    bullpen = []
    for i in range(n_relievers):
        rp = make_pitcher(rng, f"{abbr} Reliever {i + 1}", "RP")
        rp.pop("IP")  # reference reliever dicts carry no IP key (assets/bullpen_utils.py:77-92)
        bullpen.append(rp)
    return {"lineup": lineup, "starter": starter, "bullpen": bullpen}


def neutral_env() -> dict:
    """The env dict `cli/run_distribution_simulator.py:335-341` builds with --no-weather."""
    return {
        "park_hr_mult": 1.0,
        "single_mult": 1.0,
        "weather_hr_mult": 1.0,
        "adi_mult": 1.0,
        "umpire": {"K": 1.0, "BB": 1.0},
    }


def make_matchup(game_id: str = "2025-04-04-TEX@HOU", seed: int = SEED, n_relievers=(6, 6),
                 lineup_len: int = 9, bullpens: bool = True) -> dict:
    """Keyword arguments of `simulate_game` (core/game_simulator.py:14-25) plus `game_id`."""
    date_away, home = game_id.rsplit("@", 1)
    away = date_away.rsplit("-", 1)[1]
    home = home.split("-")[0]
    # zlib.crc32 rather than hash(): stable across interpreter runs
    import zlib
    rng = np.random.default_rng([seed, zlib.crc32(game_id.encode())])
    a = make_team(rng, away, n_relievers[0], lineup_len)
    h = make_team(rng, home, n_relievers[1], lineup_len)
    return {
        "game_id": game_id,
        "home_lineup": h["lineup"],
        "away_lineup": a["lineup"],
        "home_pitcher": h["starter"],
        "away_pitcher": a["starter"],
        "env": neutral_env(),
        "home_bullpen": h["bullpen"] if bullpens else None,
        "away_bullpen": a["bullpen"] if bullpens else None,
    }


def make_slate(date: str = "2025-04-04", n_games: int = 15, seed: int = SEED) -> list[dict]:
    """A full day: `n_games` matchups over distinct teams (cli/full_slate_runner.py:124-128)."""
    rng = np.random.default_rng(seed)
    order = rng.permutation(len(TEAMS))
    out = []
    for g in range(n_games):
        away, home = TEAMS[order[2 * g]], TEAMS[order[2 * g + 1]]
        out.append(make_matchup(f"{date}-{away}@{home}", seed=seed))
    return out


def neutral_calibration_matchup() -> dict:
    """The reference's own neutral game (tools/simulate_calibration_game.py:19-68): nine copies of
    an average batter, one average pitcher for both sides, no bullpens.  `hr_pa_projected` is the
    value `project_hr_pa` returns for that pitcher dict (0.036288)."""
    avg_batter = {"name": "avg_batter", "handedness": "R", "k_rate": 0.225, "bb_rate": 0.082,
                  "iso": 0.125, "avg": 0.230, "woba": 0.305}
    avg_pitcher = {"name": "avg_pitcher", "throws": "R", "stuff_plus": 100, "command_plus": 100,
                   "location_plus": 100, "k_rate": 0.225, "bb_rate": 0.082, "hr_fb_rate": 0.115,
                   "iso_allowed": 0.140, "HR": 20, "TBF": 650, "IP": 180, "barrel_batted_rate": 0.065,
                   "exit_velocity": 88.5, "launch_angle": 13, "role": "SP",
                   "hr_pa": {"hr_pa_projected": 0.036288}}
    return {
        "game_id": "neutral-calibration",
        "home_lineup": [dict(avg_batter) for _ in range(9)],
        "away_lineup": [dict(avg_batter) for _ in range(9)],
        "home_pitcher": dict(avg_pitcher),
        "away_pitcher": dict(avg_pitcher),
        "env": neutral_env(),
        "home_bullpen": None,
        "away_bullpen": None,
    }
