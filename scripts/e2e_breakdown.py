#!/usr/bin/env python
"""Where the end-to-end step's time goes beyond the kernel (measurement aid; run on the GPU box).
One bench-sized call of `ShardedSimulator.simulate` (pinned host tables in, host histogram out), split into its parts with
host timers; every part is followed by a stream synchronize in the split run, so the parts add up to more than the fused call."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mlb_odds_engine_b200 import _abi, synth, tables  # noqa: E402
from mlb_odds_engine_b200.dist import ShardedSimulator  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    S = ShardedSimulator(0)
    m = synth.make_matchup("2025-04-04-TEX@HOU")
    S.load_matchups([m])
    tbl = S.sim.tables
    pinned = torch.from_numpy(tbl.view(np.uint8).reshape(-1).copy()).pin_memory()
    tbl_host = pinned.numpy().view(_abi.TABLE_DTYPE)
    S.load_tables(tbl_host, [m])
    sync = torch.cuda.synchronize
    for i in range(20):
        S.simulate(n, 1, sim_lo=i * n, tables_host=tbl_host)
    sync()
    K = 30
    t0 = time.perf_counter()
    for i in range(K):
        S.simulate(n, 1, sim_lo=(100 + i) * n, tables_host=tbl_host)
    sync()
    fused = (time.perf_counter() - t0) / K
    parts = {"load_tables": 0.0, "run_device": 0.0, "d2h": 0.0, "host_copy_check": 0.0}
    kern = 0.0
    for i in range(K):
        t = time.perf_counter(); S.load_tables(tbl_host, S.sim.matchups); sync(); parts["load_tables"] += time.perf_counter() - t
        t = time.perf_counter(); S.run_device(n, 1, (200 + i) * n); sync(); parts["run_device"] += time.perf_counter() - t
        kern += S.ctx.last_kernel_ms() * 1e-3
        t = time.perf_counter(); S.hist_pinned.copy_(S.hist_dev, non_blocking=True); sync(); parts["d2h"] += time.perf_counter() - t
        t = time.perf_counter()
        out = S.hist_pinned.numpy().view(np.uint64).view(_abi.HIST_DTYPE).copy()
        from mlb_odds_engine_b200._lib import check_terminated
        check_terminated(out)
        parts["host_copy_check"] += time.perf_counter() - t
    print(json.dumps({"sims": n, "fused_call_us": fused * 1e6, "kernel_us": kern / K * 1e6,
                      "overhead_us": (fused - kern / K) * 1e6,
                      "split_parts_us": {k: v / K * 1e6 for k, v in parts.items()}}))


if __name__ == "__main__":
    main()
