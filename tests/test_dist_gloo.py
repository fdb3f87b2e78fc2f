"""CPU, world_size 2, gloo: the host-side sharding logic and the histogram all-reduce.  Each rank
computes its shard with the oracle's integer twin (no GPU here); the all-reduced totals must be
bit-identical to the single-rank run — the property the NCCL path relies on."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "dist_worker.py")


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_allreduce_equals_single_rank(world, tmp_path):
    import oracle as orc
    from mlb_odds_engine_b200 import _abi, synth, tables

    n, seed = 20_001, 17
    port = 29500 + (os.getpid() % 2000) + world
    out = str(tmp_path / "tot.bin")
    procs = [subprocess.Popen([sys.executable, WORKER, str(r), str(world), str(port), str(n), str(seed), out])
             for r in range(world)]
    try:
        for p in procs:
            assert p.wait(timeout=240) == 0
    finally:
        for p in procs:          # never leave a rank blocked in a collective behind
            if p.poll() is None:
                p.kill()
    got = np.frombuffer(open(out, "rb").read(), dtype=_abi.HIST_DTYPE)
    tb = tables.build_tables(synth.make_slate(n_games=3))
    summ = np.frombuffer(open(out + ".summ", "rb").read(), dtype=_abi.SUMMARY_DTYPE)
    assert len(summ) == 3
    for m in range(3):
        want = orc.native(tb[m], m, 1000, 1000 + n, seed)
        assert got[m].tobytes() == want.tobytes()
        # the matchup-sharded path (blocks of grid points + all-gather of summaries) sees the same games
        assert summ[m].tobytes() == _abi.summarize_hist(want).tobytes()


def test_shard_range_tiles_exactly():
    from mlb_odds_engine_b200.dist import shard_range

    for n in (0, 1, 7, 10_000_000, 2**40 + 13):
        for w in (1, 2, 3, 8):
            edges = [shard_range(5, n, r, w) for r in range(w)]
            assert edges[0][0] == 5 and edges[-1][1] == 5 + n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in edges) - min(h - l for l, h in edges) <= 1
