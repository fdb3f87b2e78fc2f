"""TEST INFRASTRUCTURE — the oracle twin's digests of bench.py's cross-N check, written to
tests/golden/bench_digests.json.

    python oracle/make_digests.py

For every workload of bench_workloads.SPEC: build the tables exactly as bench.py does, let the CPU integer twin
(oracle_native) play sims [0, digest_sims) of every matchup with the bench seed, and hash the records — the full
histograms for sim-sharded workloads, `_abi.summarize_hist` of them for the matchup-sharded sweep.  bench.py hashes what
the GPUs return for the same job; the two must be equal for 1, 2, 4 and 8 GPUs.
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

import bench_workloads as bw  # noqa: E402
import oracle as orc  # noqa: E402
from mlb_odds_engine_b200 import _abi  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "bench_digests.json")


def twin_records(name: str) -> np.ndarray:
    spec = bw.SPEC[name]
    tbl = bw.build_tables(name)
    n = spec["digest_sims"]
    hist = np.zeros(len(tbl), dtype=_abi.HIST_DTYPE)
    for m in range(len(tbl)):
        hist[m] = orc.native(tbl[m], m, 0, n, bw.SEED)
    if spec["shard"] == "matchups":
        out = np.zeros(len(tbl), dtype=_abi.SUMMARY_DTYPE)
        for m in range(len(tbl)):
            out[m] = _abi.summarize_hist(hist[m])
        return out
    return hist


def compute() -> dict:
    res = {}
    for name, spec in bw.SPEC.items():
        t = time.perf_counter()
        rec = twin_records(name)
        res[name] = {"sha256": bw.digest(rec), "matchups": int(len(rec)), "sims_per_matchup": spec["digest_sims"], "seed": bw.SEED,
                     "records": "SUMMARY_DTYPE" if spec["shard"] == "matchups" else "HIST_DTYPE",
                     "n_games": int(rec["n_games"].sum())}
        print(f"{name}: {res[name]['sha256'][:16]}...  {time.perf_counter() - t:.1f} s", file=sys.stderr)
    return res


if __name__ == "__main__":
    d = compute()
    with open(OUT, "w") as f:
        json.dump(d, f, indent=1)
        f.write("\n")
    print("wrote", OUT)
