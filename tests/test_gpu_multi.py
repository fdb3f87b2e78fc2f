"""GPU, >= 2 devices: the NCCL path (one process per GPU, sim-index shards, uint64 histogram
all-reduce) reproduces the single-stream oracle totals bit for bit."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_nccl_allreduce_matches_oracle():
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs (the 1-GPU box covers sharding through sim-range splits)")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29611", os.path.join(ROOT, "tests", "nccl_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "nccl worker: OK" in r.stdout


def test_c_abi_allreduce(tmp_path):
    """mlbsim_allreduce: torch-free ranks (raw NCCL communicator through ctypes) shard a slate by sim index and sum
    the device histograms inside the library; every rank ends with the oracle's totals."""
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4
    id_file = str(tmp_path / "nccl_id")
    env = dict(os.environ)   # NCCL_DEBUG stays whatever the caller set
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "nccl_capi_worker.py"), str(r), str(world), id_file],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env) for r in range(world)]
    outs = []
    import time
    deadline = time.time() + 150
    try:
        for p in procs:
            try:
                out, _ = p.communicate(timeout=max(1.0, deadline - time.time()))
            except subprocess.TimeoutExpired:
                p.kill()
                out, _ = p.communicate()
                out += "\n[killed: deadline]"
            outs.append(out)
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
    for r, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {r}:\n{out[-2000:]}"
        assert f"nccl c-abi worker rank {r}: OK" in out
