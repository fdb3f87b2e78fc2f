"""Public host API over the C ABI: `GpuSimulator` (native / per-sim / replay modes) and
`RunHistogram` (the sufficient statistics the pricing code needs).

Reference seam this replaces: the loop `for i in range(n_simulations): simulate_game(...)`
(cli/run_distribution_simulator.py:368-382) and what it collects (:380-424).
"""
from __future__ import annotations

import numpy as np

from . import _abi, tables
from ._lib import Context


class RunHistogram:
    """Integer sufficient statistics of N simulated games of one matchup (one `mlbsim_hist`).

    joint[c][a][h] = number of sims with away = a, home = h runs after CAPS[c] innings
    (c = 4: final score).  SURVEY.md Appendix B addendum: these joint counts plus the
    reliever-appearance counts are sufficient for every market `simulate_distribution` prices.
    """

    def __init__(self, rec: np.ndarray, matchup: tables.Matchup | None = None):
        assert rec.dtype == _abi.HIST_DTYPE
        self.rec = rec.copy()
        self.matchup = matchup

    # ---- raw views ----
    @property
    def n_games(self) -> int:
        return int(self.rec["n_games"])

    def joint(self, cap=None) -> np.ndarray:
        """[away][home] counts after `cap` innings (1, 3, 5, 7) or the full game (None)."""
        return self.rec["joint"][_abi.CAPS.index(cap)]

    @property
    def innings(self) -> np.ndarray:
        return self.rec["innings"]

    def __add__(self, other: "RunHistogram") -> "RunHistogram":
        out = RunHistogram(self.rec, self.matchup)
        a = out.rec.reshape(1).view(np.uint64)
        a += other.rec.reshape(1).view(np.uint64)
        return out

    def __eq__(self, other):
        return isinstance(other, RunHistogram) and self.rec.tobytes() == other.rec.tobytes()

    # ---- adapters to the reference's vocabulary ----
    def pa_recap(self) -> dict:
        """Same shape as core/pa_simulator.py:10-13 PA_RECAP_LOG."""
        return {team: {o: int(self.rec["pa"][s][i]) for i, o in enumerate(_abi.OUTCOMES)}
                for s, team in ((1, "HOME"), (0, "AWAY"))}

    def reliever_usage(self, bullpen_names=None) -> dict:
        """{"home": {name: sims in which he appeared}, "away": {...}} —
        cli/run_distribution_simulator.py:391-397.  `bullpen_names`: [home names, away names] when the histogram
        travels without its Matchup."""
        out = {"home": {}, "away": {}}
        for s, side in ((0, "home"), (1, "away")):
            n = 8
            names = (bullpen_names[s] if bullpen_names is not None else
                     self.matchup.bullpen_names[s] if self.matchup is not None else [f"#{i}" for i in range(n)])
            for i, name in enumerate(names):
                c = int(self.rec["rel_games"][s][i])
                if c:
                    out[side][name] = out[side].get(name, 0) + c
        return out

    def score_counts(self, cap=None):
        """(home, away, count) arrays of the non-empty joint cells — weighted stand-in for the
        per-sim lists raw_home_scores / raw_away_scores (run_distribution_simulator.py:368-381)."""
        j = self.joint(cap)
        a, h = np.nonzero(j)
        return h.astype(np.int64), a.astype(np.int64), j[a, h].astype(np.int64)

    def home_win_prob(self) -> float:
        h, a, c = self.score_counts()
        return float(c[h > a].sum() / c.sum())

    def mean_scores(self, cap=None):
        h, a, c = self.score_counts(cap)
        n = c.sum()
        return float((h * c).sum() / n), float((a * c).sum() / n)


class _DeviceTables:
    """Placeholder for tables that exist only in HBM (built by `Context.build_tables`) until somebody asks for them."""

    def __init__(self, ctx):
        self.ctx = ctx

    def fetch(self) -> np.ndarray:
        return self.ctx.download_tables()


class GpuSimulator:
    """Owns one device context and the uploaded matchup tables."""

    def __init__(self, device: int = 0):
        self.ctx = Context(device)
        self.matchups: list[tables.Matchup] = []
        self._tables = None

    def close(self):
        self.ctx.close()

    # ---- inputs ----
    def load_matchups(self, matchups, use_noise: bool = True, builder: str = "device"):
        """`matchups`: iterable of `simulate_game` keyword dicts (or tables.Matchup).  `builder="device"` (default): the
        raw numbers go up as `mlbsim_params` and the sampling tables are built in HBM (csrc/table_kernel.cu);
        `builder="host"`: numpy / scipy build them (tables.build_tables, alias rows on the device) and they are
        uploaded — same tables up to one unit of a 29/31-bit entry in rare rows (the noise quadrature)."""
        self.matchups = [tables.as_matchup(m, use_noise) for m in matchups]
        if builder == "device":
            return self.load_params(tables.params_batch(self.matchups), self.matchups)
        if builder != "host":
            raise ValueError(f"builder must be 'device' or 'host', not {builder!r}")
        self._tables = tables.build_tables(self.matchups, alias_fn=self.ctx.alias_rows)   # alias rows on the device (bit-identical)
        self.ctx.upload_tables(self._tables)
        return self

    def load_params(self, params: np.ndarray, matchups=None, validate: bool = True):
        """Build and install tables on the device from PARAMS_DTYPE records (`tables.params_batch`, `sweep.grid_params`)."""
        self.ctx.build_tables(params, validate=validate)
        self.matchups = list(matchups) if matchups is not None else [None] * len(np.atleast_1d(params))
        self._tables = _DeviceTables(self.ctx)
        return self

    def load_tables(self, tbl: np.ndarray, matchups=None):
        """Upload prebuilt tables (e.g. from pinned memory in the e2e benchmark)."""
        self._tables = tbl
        self.matchups = list(matchups) if matchups is not None else [None] * len(np.atleast_1d(tbl))
        self.ctx.upload_tables(tbl)
        return self

    @property
    def tables(self) -> np.ndarray:
        """The installed tables as TABLE_DTYPE records on the host (downloaded once when they were built on the device)."""
        if isinstance(self._tables, _DeviceTables):
            self._tables = self._tables.fetch()
        return self._tables

    # ---- native-RNG mode ----
    def simulate(self, n_sims: int, seed: int = 0, sim_lo: int = 0, matchup_lo: int = 0, matchup_hi: int | None = None,
                 count_pa: bool = True) -> list[RunHistogram]:
        if self._tables is None:
            raise RuntimeError("load_matchups() first")
        hi = self.ctx.n_tables if matchup_hi is None else matchup_hi
        rec = self.ctx.run_native_host(matchup_lo, hi, sim_lo, sim_lo + int(n_sims), int(seed), 0 if count_pa else 1)
        return [RunHistogram(rec[i], self.matchups[matchup_lo + i]) for i in range(hi - matchup_lo)]

    def simulate_summary(self, n_sims: int, seed: int = 0, sim_lo: int = 0, matchup_lo: int = 0,
                         matchup_hi: int | None = None, count_pa: bool = True) -> np.ndarray:
        """Same games as `simulate`, reduced on the device to one SUMMARY_DTYPE record per matchup (outcome counts,
        run moments, win counts): 320 B per matchup come back instead of 165 KB — the shape of a sweep."""
        if self._tables is None:
            raise RuntimeError("load_matchups() first")
        hi = self.ctx.n_tables if matchup_hi is None else matchup_hi
        return self.ctx.run_native_summary_host(matchup_lo, hi, sim_lo, sim_lo + int(n_sims), int(seed), 0 if count_pa else 1)

    def simulate_persim(self, n_sims: int, seed: int = 0, sim_lo: int = 0, matchup: int = 0) -> np.ndarray:
        """One record per sim (REPLAY_GAME_DTYPE) — the same games `simulate` counts."""
        return self.ctx.run_persim(matchup, sim_lo, sim_lo + int(n_sims), int(seed))

    # ---- replay mode ----
    def replay(self, matchup, tape: np.ndarray, offsets: np.ndarray, use_noise: bool = True) -> np.ndarray:
        params = tables.build_params(matchup, use_noise=use_noise)
        return self.ctx.run_replay(params, tape, offsets)


def result_dicts(games: np.ndarray, matchup: tables.Matchup | None = None) -> list[dict]:
    """Per-sim records -> the dicts `simulate_game` returns, restricted to the fields the driver
    consumes (core/game_simulator.py:134-144; SURVEY.md Appendix B addendum).  Columns are converted to Python
    lists once, so 10 000 records cost tens of milliseconds."""
    names = matchup.bullpen_names if matchup is not None else None
    home, away = games["home_score"].tolist(), games["away_score"].tolist()
    n_inn = np.minimum(games["n_innings"], _abi.REPLAY_MAX_INNINGS).tolist()
    a_runs, h_runs = games["away_runs"].tolist(), games["home_runs"].tolist()
    n_used = np.minimum(games["n_used"], _abi.REPLAY_MAX_USED).tolist()
    used_cols = (games["used_home"].tolist(), games["used_away"].tolist())
    recaps = games["pa"].astype(np.int64).sum(axis=1)[:, :7].tolist()
    out = []
    for g in range(len(games)):
        ar, hr = a_runs[g], h_runs[g]
        innings = [{"inning": i + 1, "away_runs": ar[i], "home_runs": hr[i]} for i in range(n_inn[g])]
        used = []
        for s in (0, 1):
            idx = used_cols[s][g][: n_used[g][s]]
            used.append([names[s][i] for i in idx] if names is not None else idx)
        total, margin = home[g] + away[g], abs(home[g] - away[g])
        # game_type: core/game_simulator.py:123-132
        game_type = ("Pitcher's Duel" if total <= 3 else "Blowout" if margin >= 7
                     else "Tight Offensive Battle" if margin == 1 and total >= 6 else "Balanced")
        out.append({
            "home_score": home[g], "away_score": away[g], "innings": innings,
            "used_home_relievers": used[0], "used_away_relievers": used[1], "game_type": game_type,
            "recap": dict(zip(_abi.OUTCOMES, recaps[g])),
        })
    return out
