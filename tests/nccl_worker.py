"""Worker of tests/test_gpu_multi.py: one process per GPU (torchrun), NCCL all-reduce of histograms."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    import numpy as np
    import torch
    import torch.distributed as dist

    import oracle as orc
    from mlb_odds_engine_b200 import _abi, synth
    from mlb_odds_engine_b200.dist import ShardedSimulator

    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    S = ShardedSimulator(local)
    slate = synth.make_slate(n_games=3)
    slate[1]["home_bullpen"] = None
    S.load_matchups(slate)
    n, seed = 300_001, 5
    tot = S.simulate(n, seed, sim_lo=77)
    ok = True
    if dist.get_rank() == 0:
        for m in range(3):
            want = orc.native(S.sim.tables[m], m, 77, 77 + n, seed)
            if tot[m].tobytes() != want.tobytes():
                ok = False
                print(f"matchup {m}: all-reduced histogram differs from the oracle", flush=True)
    # sweep-shaped call: all-reduced histograms summarised on the device, every rank gets the job totals
    summ = S.simulate_summary(n, seed, sim_lo=77)
    for m in range(3):
        if summ[m].tobytes() != _abi.summarize_hist(tot[m]).tobytes():
            ok = False
            print(f"rank {dist.get_rank()} matchup {m}: device summary differs from the reduced histogram", flush=True)
    # matchup-sharded sweep: every rank owns a block of the 5 grid points, summaries all-gathered; the result must be
    # what one GPU computes alone (same games: the Philox counter carries the global matchup id)
    five = synth.make_slate(n_games=5)
    five[2]["away_bullpen"] = None
    from mlb_odds_engine_b200 import tables

    tb = tables.build_tables(five)
    B = ShardedSimulator(local, shard="matchups")
    B.load_tables(tb)
    got = B.simulate_summary(40_000, seed, sim_lo=5)
    for m in range(5):
        want = _abi.summarize_hist(orc.native(tb[m], m, 5, 5 + 40_000, seed))
        if got[m].tobytes() != want.tobytes():
            ok = False
            print(f"rank {dist.get_rank()} grid point {m}: matchup-sharded summary differs from the oracle", flush=True)
    B.close()
    flag = torch.tensor([0 if ok else 1], device=f"cuda:{local}")
    dist.all_reduce(flag)
    ok = int(flag.item()) == 0
    if dist.get_rank() == 0:
        print("nccl worker:", "OK" if ok else "MISMATCH", "world", dist.get_world_size(), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
