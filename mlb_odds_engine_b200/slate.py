"""Slate runner on the GPU path: replaces the serial per-game loop of `cli/full_slate_runner.py:135-162`
(one `simulate_distribution` per game id, 10 000 sims each) by ONE multi-matchup launch followed by
per-game histogram-fed pricing and the same per-game JSON files (`backtest/sims/<date>/<game_id>.json`,
`cli/run_distribution_simulator.py:797-850`).

Asset building (lineup scraping, probable pitchers, park/weather) stays with the reference's
network-bound builders (out of scope): callers pass the dicts those
builders produce (`core/game_asset_builder.py:263-267` + the env dict of
`cli/run_distribution_simulator.py:335-341`).
"""
from __future__ import annotations

import os

import numpy as np

from . import _abi, distribution, tables
from .simulator import GpuSimulator, RunHistogram

_POOL = None


def pricing_pool(workers: int):
    """A persistent pool of worker PROCESSES for the per-game host work (pricing a histogram and writing its JSON is
    ~5 ms of pure Python per game: after a one-launch slate it, not the GPU, is the critical path).  Spawned, not forked
    — the parent holds a CUDA context — and created once per process: the ~0.5 s of interpreter start-up is paid on the
    first slate only."""
    global _POOL
    if _POOL is None or _POOL[0] != workers:
        import multiprocessing as mp
        from concurrent.futures import ProcessPoolExecutor

        if _POOL is not None:
            _POOL[1].shutdown(wait=False)
        import atexit

        ex = ProcessPoolExecutor(workers, mp_context=mp.get_context("spawn"))
        list(ex.map(_warm, range(workers)))          # import numpy / the package in every worker now
        atexit.register(ex.shutdown, wait=True, cancel_futures=True)   # before the interpreter tears modules down
        _POOL = (workers, ex)
    return _POOL[1]


def _warm(i):
    return i


def _price_one(job):
    """Worker side: histogram bytes -> priced document (+ files).  Returns (game_id, doc | None, error | None)."""
    gid, rec_bytes, bullpen_names, calibration, start_time, json_path, snapshot_path = job
    try:
        rec = np.frombuffer(rec_bytes, dtype=_abi.HIST_DTYPE)[0]
        h = RunHistogram(rec, None)
        doc = distribution.price_from_histogram(h, gid, calibration, start_time)
        doc["reliever_usage"] = h.reliever_usage(bullpen_names)
        doc["pa_recap"] = h.pa_recap()
        if json_path:
            distribution.write_sim_json(doc, json_path)
        if snapshot_path:
            distribution.write_table_snapshot(doc, snapshot_path)
        return gid, doc, None
    except Exception as e:
        return gid, None, f"{type(e).__name__}: {e}"


def matchup_from_assets(game_id: str, assets: dict, env: dict | None = None, use_noise: bool = True) -> dict:
    """{"lineups","pitchers","bullpens"} (build_game_assets) -> simulate_game keyword dict."""
    return {
        "game_id": game_id,
        "home_lineup": assets["lineups"]["home"], "away_lineup": assets["lineups"]["away"],
        "home_pitcher": assets["pitchers"]["home"], "away_pitcher": assets["pitchers"]["away"],
        "env": env or {},
        "home_bullpen": assets.get("bullpens", {}).get("home", []),
        "away_bullpen": assets.get("bullpens", {}).get("away", []),
        "use_noise": use_noise,
    }


def run_slate(matchups: list[dict], n_sims: int = 10_000, seed: int = 0, calibration: dict | None = None,
              out_dir: str | None = None, safe: bool = False, sim: GpuSimulator | None = None,
              start_times: dict | None = None, snapshot_path: str | None = None, workers: int = 0) -> dict:
    """Simulate every game of the slate in one launch and price it.  `matchups`: simulate_game keyword
    dicts carrying `game_id` ('YYYY-MM-DD-AWAY@HOME[-Thhmm]').  Returns {game_id: document}; with `out_dir` also writes
    `<out_dir>/<date>/<game_id>.json`, and with `snapshot_path` the `last_table_snapshot.json` side file the reference
    rewrites after every game (cli/run_distribution_simulator.py:853-859; the last game of the slate wins, as in the
    serial loop).  A game that fails — malformed assets, an id without AWAY@HOME, an unwritable path — raises, or with
    `safe=True` is reported under "failed" and skipped: the semantics of `--safe` (cli/full_slate_runner.py:157-162).
    `workers` > 0 prices and writes the games in that many worker processes (`pricing_pool`)."""
    own = sim is None
    sim = sim or GpuSimulator(0)
    built, failed = [], {}
    for i, m in enumerate(matchups):
        gid = m.get("game_id", f"game{i}")
        try:
            if not all(distribution.parse_game_id(gid)[k] for k in ("date", "away", "home")):
                raise ValueError(f"game_id {gid!r} is not of the form YYYY-MM-DD-AWAY@HOME[-Thhmm]")
            built.append((gid, tables.as_matchup(m)))
        except Exception as e:  # what simulate_game would raise on the first PA / simulate_distribution on the id
            if not safe:
                raise
            failed[gid] = f"{type(e).__name__}: {e}"
    docs = {}
    try:
        if built:
            sim.load_matchups([m for _, m in built])
            hists = sim.simulate(n_sims, seed=seed)
            jobs = []
            for (gid, m), h in zip(built, hists):
                path = os.path.join(out_dir, distribution.parse_game_id(gid)["date"], f"{gid}.json") if out_dir else None
                jobs.append((gid, h.rec.tobytes(), m.bullpen_names, calibration, (start_times or {}).get(gid), path, None))
            results = list(pricing_pool(workers).map(_price_one, jobs)) if workers > 0 and len(jobs) > 1 else [_price_one(j) for j in jobs]
            for gid, doc, err in results:
                if err is not None:
                    if not safe:
                        raise RuntimeError(f"{gid}: {err}")
                    failed[gid] = err
                else:
                    docs[gid] = doc
            if snapshot_path and docs:      # the serial reference loop leaves the LAST game's snapshot behind
                last = [gid for gid, _ in built if gid in docs][-1]
                distribution.write_table_snapshot(docs[last], snapshot_path)
    finally:
        if own:
            sim.close()
    if failed:
        docs["failed"] = failed
    return docs
