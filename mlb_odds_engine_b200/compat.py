"""Drop-in callables for the reference's module-global seam:

    import cli.run_distribution_simulator as rds
    rds.simulate_game = GpuSimulateGame(n_prefetch=10_000, seed=0)

`simulate_distribution` calls `simulate_game(**kw)` once per sim with keyword arguments only
(cli/run_distribution_simulator.py:370-379); the callable simulates a batch on the GPU the first
time it sees an argument set and then hands out one result dict per call, so the reference's own
reduction / pricing code runs unchanged on GPU-produced games.
"""
from __future__ import annotations

from . import tables
from .simulator import GpuSimulator, result_dicts


class ResultFeeder:
    """Hands out precomputed result dicts in order; refills through `refill(n_done) -> list[dict]`."""

    def __init__(self, refill):
        self._refill = refill
        self._buf = []
        self._pos = 0
        self._done = 0

    def next(self):
        if self._pos >= len(self._buf):
            self._buf = self._refill(self._done)
            self._pos = 0
            if not self._buf:
                raise RuntimeError("result feeder ran dry")
        r = self._buf[self._pos]
        self._pos += 1
        self._done += 1
        return r


def argument_fingerprint(kw: dict) -> tuple:
    """Every value of a `simulate_game` argument set that the hot path reads (the fields `tables.Matchup` parses),
    as one hashable tuple.  Two argument sets with the same fingerprint simulate the same games; an argument set that
    is edited in place between calls, or a new set that happens to live at a recycled address, changes it."""
    def batter(b):
        return (b.get("k_rate", 0.22), b.get("bb_rate", 0.08), b.get("speed", 50))

    def pitcher(p):
        hr = p.get("hr_pa", {})
        return (p.get("k_rate"), p.get("bb_rate"), hr.get("hr_pa_projected", 0.046) if isinstance(hr, dict) else repr(hr),
                p.get("exit_velocity_avg"), p.get("launch_angle_avg"), p.get("fielder_rating", 50), p.get("name", "Unknown"))

    env = kw.get("env") or {}
    ump = env.get("umpire", {}) or {}
    return (
        tuple(batter(b) for b in kw["home_lineup"]), tuple(batter(b) for b in kw["away_lineup"]),
        pitcher(kw["home_pitcher"]), pitcher(kw["away_pitcher"]),
        tuple(pitcher(p) for p in (kw.get("home_bullpen") or [])), tuple(pitcher(p) for p in (kw.get("away_bullpen") or [])),
        ump.get("k_mod", 1.0), ump.get("bb_mod", 1.0), env.get("weather_hr", 1.0), bool(kw.get("use_noise", True)),
    )


class GpuSimulateGame:
    """`simulate_game` replacement backed by `mlbsim_run_persim` (same signature, core/game_simulator.py:14-25).
    `debug` / `return_inning_scores` are honoured where they have a GPU-native meaning.

    Prefetched results are keyed by the CONTENT of the arguments (`argument_fingerprint`), never by object identity:
    the next game of a slate, or the same dicts edited in place, get their own simulation.  `reset()` drops the
    prefetched games explicitly (e.g. to re-simulate the same matchup with the next block of sim indices)."""

    def __init__(self, n_prefetch: int = 10_000, seed: int = 0, device: int = 0):
        self.n_prefetch = int(n_prefetch)
        self.seed = int(seed)
        self.sim = GpuSimulator(device)
        self._key = None
        self._feeder = None

    def reset(self):
        self._key = None
        self._feeder = None

    def __call__(self, home_lineup, away_lineup, home_pitcher, away_pitcher, env, home_bullpen=None, away_bullpen=None,
                 debug=False, return_inning_scores=False, use_noise=True):
        kw = dict(home_lineup=home_lineup, away_lineup=away_lineup, home_pitcher=home_pitcher, away_pitcher=away_pitcher,
                  env=env, home_bullpen=home_bullpen, away_bullpen=away_bullpen, use_noise=use_noise)
        key = argument_fingerprint(kw)
        if key != self._key:
            m = tables.Matchup(**kw)
            self.sim.load_matchups([m])

            def refill(done, m=m):
                games = self.sim.simulate_persim(self.n_prefetch, seed=self.seed, sim_lo=done)
                return result_dicts(games, m)

            self._feeder = ResultFeeder(refill)
            self._key = key
        r = self._feeder.next()
        if return_inning_scores:   # core/game_simulator.py:146-150
            r = dict(r)
            r["inning_scores"] = {i["inning"]: {"home": i["home_runs"], "away": i["away_runs"]} for i in r["innings"]}
        return r
