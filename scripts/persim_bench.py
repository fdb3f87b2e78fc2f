"""Per-sim path timing (small-N compatibility): mlbsim_run_persim records/s and the cost of turning records into the
reference's result dicts.  usage: python scripts/persim_bench.py"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from mlb_odds_engine_b200 import synth  # noqa: E402
from mlb_odds_engine_b200.simulator import GpuSimulator, result_dicts  # noqa: E402

sim = GpuSimulator(0).load_matchups([synth.make_matchup()])
sim.simulate_persim(1000, seed=1)
out = {}
for n in (10_000, 1_000_000):
    t = time.perf_counter()
    g = sim.simulate_persim(n, seed=2)
    out[f"persim_{n}_s"] = time.perf_counter() - t
    out[f"persim_{n}_games_per_s"] = n / out[f"persim_{n}_s"]
t = time.perf_counter()
d = result_dicts(sim.simulate_persim(10_000, seed=3), sim.matchups[0])
out["result_dicts_10000_s"] = time.perf_counter() - t
print(json.dumps(out))
