"""Worker of tests/test_gpu_multi.py::test_c_abi_allreduce: one process per GPU, NO torch on the data path.
The communicator is made with raw NCCL calls through ctypes (unique id exchanged through a file), the shard is
simulated with mlbsim_run_native and summed with mlbsim_allreduce on the library's own stream.
usage: nccl_capi_worker.py <rank> <world> <id_file>"""
import ctypes
import glob
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


class NcclUniqueId(ctypes.Structure):
    _fields_ = [("internal", ctypes.c_ubyte * 128)]   # (c_char arrays read back NUL-terminated: the id is binary)


def load_nccl():
    """libnccl.so.2: the wheel-bundled copy if there is one (no `import torch`), else the system library."""
    cands = []
    for p in sys.path:
        cands += glob.glob(os.path.join(p, "nvidia", "nccl", "lib", "libnccl.so.2"))
    for path in cands + ["libnccl.so.2"]:
        try:
            return ctypes.CDLL(path, mode=ctypes.RTLD_GLOBAL), path
        except OSError:
            continue
    raise OSError("libnccl.so.2 not found")


def main():
    rank, world, id_file = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
    import faulthandler

    faulthandler.dump_traceback_later(60, exit=True)   # a hung collective must not outlive the test

    def say(msg):
        print(f"[rank {rank}] {msg}", flush=True)

    import numpy as np

    import oracle as orc
    from mlb_odds_engine_b200 import _abi, synth, tables
    from mlb_odds_engine_b200._lib import Context
    from mlb_odds_engine_b200.dist import shard_range

    nccl, path = load_nccl()
    os.environ["MLBSIM_NCCL_LIB"] = path
    nccl.ncclGetUniqueId.argtypes = [ctypes.POINTER(NcclUniqueId)]
    nccl.ncclCommInitRank.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, NcclUniqueId, ctypes.c_int]
    nccl.ncclCommDestroy.argtypes = [ctypes.c_void_p]

    say(f"nccl from {path}")
    ctx = Context(rank)          # binds cuda device `rank` (mlbsim_init), own stream
    say("context up")
    uid = NcclUniqueId()
    if rank == 0:
        assert nccl.ncclGetUniqueId(ctypes.byref(uid)) == 0
        with open(id_file + ".tmp", "wb") as f:
            f.write(ctypes.string_at(ctypes.byref(uid), 128))
        os.replace(id_file + ".tmp", id_file)
    else:
        t0 = time.time()
        while not os.path.exists(id_file):
            if time.time() - t0 > 120:
                raise TimeoutError("no NCCL id from rank 0")
            time.sleep(0.05)
        raw = open(id_file, "rb").read()
        assert len(raw) == 128, len(raw)
        ctypes.memmove(ctypes.byref(uid), raw, 128)
    comm = ctypes.c_void_p()
    rc = nccl.ncclCommInitRank(ctypes.byref(comm), world, uid, rank)
    assert rc == 0, f"ncclCommInitRank -> {rc}"
    say("communicator up")

    slate = synth.make_slate(n_games=3)
    slate[2]["away_bullpen"] = None
    tbl = tables.build_tables(slate)
    ctx.upload_tables(tbl)
    n, seed, lo0 = 200_003, 9, 1234
    nbytes = 3 * _abi.HIST_DTYPE.itemsize
    hist_dev = ctx.malloc(nbytes)
    ctx.memset(hist_dev, 0, nbytes)
    lo, hi = shard_range(lo0, n, rank, world)
    ctx.run_native(0, 3, lo, hi, seed, hist_dev)
    say("kernel enqueued")
    ctx.allreduce(comm.value, hist_dev, 3)          # same stream, no host sync in between
    say("all-reduce enqueued")
    out = np.zeros(3, dtype=_abi.HIST_DTYPE)
    ctx.d2h(out, hist_dev)
    say("totals on host")
    ok = True
    for m in range(3):
        want = orc.native(tbl[m], m, lo0, lo0 + n, seed)
        if out[m].tobytes() != want.tobytes():
            ok = False
            print(f"rank {rank} matchup {m}: all-reduced histogram differs from the oracle", flush=True)
    ctx.free(hist_dev)
    nccl.ncclCommDestroy(comm)
    ctx.close()
    print(f"nccl c-abi worker rank {rank}:", "OK" if ok else "MISMATCH", flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
