"""CPU: the digests bench.py's `workloads` block must reproduce on 1, 2, 4 and 8 GPUs (tests/golden/bench_digests.json)
are what the oracle twin computes for the same fixed (seed, sims) jobs — and a sim-index split of such a job adds up
to the same bytes, which is why the digest cannot depend on the GPU count."""
import json
import os

import numpy as np

import bench_workloads as bw
import oracle as orc
from mlb_odds_engine_b200 import _abi
from mlb_odds_engine_b200.dist import shard_range

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_committed_digests_are_the_oracle_twins():
    import make_digests

    want = json.load(open(os.path.join(ROOT, "tests", "golden", "bench_digests.json")))
    assert set(want) == set(bw.SPEC)
    got = make_digests.compute()
    for name in bw.SPEC:
        assert got[name]["sha256"] == want[name]["sha256"], name
        assert got[name]["n_games"] == want[name]["matchups"] * bw.SPEC[name]["digest_sims"]


def test_digest_does_not_depend_on_how_the_sims_are_split():
    tbl = bw.build_tables("single")
    n = 60_000
    whole = np.zeros(1, dtype=_abi.HIST_DTYPE)
    whole[0] = orc.native(tbl[0], 0, 0, n, bw.SEED)
    for world in (2, 3, 8):
        tot = np.zeros(_abi.HIST_WORDS, dtype=np.uint64)
        for r in range(world):
            lo, hi = shard_range(0, n, r, world)
            tot += np.atleast_1d(orc.native(tbl[0], 0, lo, hi, bw.SEED)).view(np.uint64).reshape(-1)
        assert bw.digest(tot.view(_abi.HIST_DTYPE)) == bw.digest(whole), world
