"""Wall-clock breakdown of a whole slate run on the GPU path (BASELINE configs[2]: 15 games x 10^6 sims):
table build, upload + simulate + D2H, pricing, JSON write.  Usage: python scripts/slate_e2e.py [n_games] [n_sims]"""
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from mlb_odds_engine_b200 import distribution, slate, synth, tables  # noqa: E402
from mlb_odds_engine_b200.simulator import GpuSimulator  # noqa: E402

n_games = int(sys.argv[1]) if len(sys.argv) > 1 else 15
n_sims = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
games = synth.make_slate("2025-04-04", n_games)
sim = GpuSimulator(0)
slate.run_slate(games[:2], n_sims=1000, sim=sim)               # warm-up: context, scipy import, kernels
tables._CLIP_CACHE.clear()
out = {"n_games": n_games, "n_sims": n_sims}
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
with tempfile.TemporaryDirectory() as d:
    slate.run_slate(games, n_sims=n_sims, seed=1, out_dir=d, sim=sim)
t6 = time.perf_counter()
out.update(parse_dicts_s=t1 - t0, build_and_upload_tables_s=t2 - t1, simulate_and_d2h_s=t3 - t2, pricing_s=t4 - t3,
           json_write_s=t5 - t4, run_slate_total_s=t6 - t5, games_per_s_whole_pipeline=n_games * n_sims / (t6 - t5),
           kernel_ms=sim.ctx.last_kernel_ms() if hasattr(sim.ctx, "last_kernel_ms") else None)
print(json.dumps(out))
