"""Worker of tests/test_dist_gloo.py (one process per rank, gloo on 127.0.0.1)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    rank, world, port, n, seed, out_path = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], int(sys.argv[4]), int(sys.argv[5]), sys.argv[6]
    import numpy as np
    import torch.distributed as dist

    import oracle as orc
    from mlb_odds_engine_b200 import _abi, dist as mdist, synth, tables

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    tb = tables.build_tables(synth.make_slate(n_games=3))
    orc.set_threads(2)
    rec = np.zeros(3, dtype=_abi.HIST_DTYPE)
    lo, hi = mdist.shard_range(1000, n, rank, world)
    for m in range(3):
        rec[m] = orc.native(tb[m], m, lo, hi, seed)
    tot = mdist.allreduce_hist_numpy(rec)
    # matchup-sharded sweep: this rank simulates only its block of the 3 grid points, all sims; summaries are gathered
    m_lo, m_hi = mdist.shard_range(0, 3, rank, world)
    local = np.zeros(m_hi - m_lo, dtype=_abi.SUMMARY_DTYPE)
    for m in range(m_lo, m_hi):
        local[m - m_lo] = _abi.summarize_hist(orc.native(tb[m], m, 1000, 1000 + n, seed))
    summ = mdist.allgather_summaries_numpy(local, 3)
    if rank == 0:
        with open(out_path, "wb") as f:
            f.write(tot.tobytes())
        with open(out_path + ".summ", "wb") as f:
            f.write(summ.tobytes())
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
