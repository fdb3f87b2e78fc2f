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


class GpuSimulateGame:
    """`simulate_game` replacement backed by `mlbsim_run_persim` (same signature, core/game_simulator.py:14-25).
    `debug` / `return_inning_scores` are honoured where they have a GPU-native meaning."""

    def __init__(self, n_prefetch: int = 10_000, seed: int = 0, device: int = 0):
        self.n_prefetch = int(n_prefetch)
        self.seed = int(seed)
        self.sim = GpuSimulator(device)
        self._key = None
        self._feeder = None

    def _arg_key(self, kw):
        return tuple(id(kw.get(k)) for k in ("home_lineup", "away_lineup", "home_pitcher", "away_pitcher", "env",
                                             "home_bullpen", "away_bullpen")) + (bool(kw.get("use_noise", True)),)

    def __call__(self, home_lineup, away_lineup, home_pitcher, away_pitcher, env, home_bullpen=None, away_bullpen=None,
                 debug=False, return_inning_scores=False, use_noise=True):
        kw = dict(home_lineup=home_lineup, away_lineup=away_lineup, home_pitcher=home_pitcher, away_pitcher=away_pitcher,
                  env=env, home_bullpen=home_bullpen, away_bullpen=away_bullpen, use_noise=use_noise)
        key = self._arg_key(kw)
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
