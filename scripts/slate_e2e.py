"""Wall-clock breakdown of a whole slate run on the GPU path (BASELINE configs[2]: 15 games x 10^6 sims):
dict parsing, table build + upload, simulate + D2H, pricing, JSON write; then `slate.run_slate` end to end, serial and
with worker processes for the per-game host work, with a cold and a warm noise-integral cache.
Usage: python scripts/slate_e2e.py [n_games] [n_sims] [workers]"""
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from mlb_odds_engine_b200 import distribution, slate, synth, tables  # noqa: E402
from mlb_odds_engine_b200.simulator import GpuSimulator  # noqa: E402


def main():
    n_games = int(sys.argv[1]) if len(sys.argv) > 1 else 15
    n_sims = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
    workers = int(sys.argv[3]) if len(sys.argv) > 3 else min(8, os.cpu_count() or 1)
    games = synth.make_slate("2025-04-04", n_games)
    sim = GpuSimulator(0)
    with tempfile.TemporaryDirectory() as d:                        # warm-up: context, scipy import, kernels, worker processes
        slate.run_slate(games, n_sims=1000, sim=sim, out_dir=d, workers=workers)
    out = {"n_games": n_games, "n_sims": n_sims, "workers": workers, "host_cores": os.cpu_count()}
    tables._CLIP_CACHE.clear()
    t0 = time.perf_counter()
    ms = [tables.as_matchup(g) for g in games]
    t1 = time.perf_counter()
    sim.load_matchups(ms)
    t2 = time.perf_counter()
    hists = sim.simulate(n_sims, seed=1)
    t3 = time.perf_counter()
    docs = [distribution.price_from_histogram(h, g["game_id"]) for h, g in zip(hists, games)]
    t4 = time.perf_counter()
    with tempfile.TemporaryDirectory() as d:
        for doc, g in zip(docs, games):
            distribution.write_sim_json(doc, os.path.join(d, g["game_id"] + ".json"))
    t5 = time.perf_counter()
    out.update(parse_dicts_s=t1 - t0, build_and_upload_tables_cold_s=t2 - t1, simulate_and_d2h_s=t3 - t2, kernel_ms=sim.ctx.last_kernel_ms(),
               pricing_serial_s=t4 - t3, json_write_serial_s=t5 - t4)
    for label, w, cold in (("serial_cold", 0, True), ("serial_warm", 0, False), ("pooled_cold", workers, True), ("pooled_warm", workers, False)):
        best = None
        for _ in range(3):
            if cold:
                tables._CLIP_CACHE.clear()
            with tempfile.TemporaryDirectory() as d:
                t = time.perf_counter()
                slate.run_slate(games, n_sims=n_sims, seed=1, out_dir=d, sim=sim, workers=w)
                dt = time.perf_counter() - t
            best = dt if best is None else min(best, dt)
        out[f"run_slate_{label}_s"] = best
    out["games_per_s_whole_pipeline_best"] = n_games * n_sims / min(out["run_slate_pooled_warm_s"], out["run_slate_serial_warm_s"])
    print(json.dumps(out))


if __name__ == "__main__":
    main()
