"""Sim-index sharding over the GPUs of one box + one all-reduce of the integer histograms.

Every (matchup, sim index) is independent and its random stream is keyed by (seed, matchup, sim
index) only, so rank r of W simply takes the contiguous slice of the sim range
[sim_lo + r*n/W, sim_lo + (r+1)*n/W) for every matchup; the only exchange is a sum of
`n_matchups * HIST_WORDS` uint64 counters (164 KB per matchup) — `torch.distributed.all_reduce`
(NCCL over NVLink on GPUs, gloo in the CPU tests).  Integer sums => totals are bit-identical for
1/2/4/8 GPUs.  torch is plumbing here: device memory, the stream, the collective.
"""
from __future__ import annotations

import numpy as np

from . import _abi
from ._lib import check_terminated


def shard_range(sim_lo: int, n_sims: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous slice of [sim_lo, sim_lo + n_sims) owned by `rank`; slices tile the range exactly."""
    lo = sim_lo + (n_sims * rank) // world
    hi = sim_lo + (n_sims * (rank + 1)) // world
    return lo, hi


def allreduce_hist_numpy(rec: np.ndarray) -> np.ndarray:
    """Sum HIST_DTYPE records over all ranks of the default process group (any backend that can
    reduce int64 on CPU, i.e. gloo); returns a new array.  Used by the CPU tests of the host logic."""
    import torch
    import torch.distributed as dist

    flat = torch.from_numpy(np.ascontiguousarray(rec).reshape(-1).view(np.int64).copy())
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    return flat.numpy().view(np.uint64).view(_abi.HIST_DTYPE).reshape(np.shape(rec))


class ShardedSimulator:
    """One rank's simulator + the all-reduce.  Requires torch.cuda; the kernel runs on torch's
    current stream so CUDA events and NCCL order correctly around it."""

    def __init__(self, device: int | None = None):
        import torch
        import torch.distributed as dist

        from .simulator import GpuSimulator

        self.torch = torch
        self.dist = dist
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.device = torch.cuda.current_device() if device is None else device
        torch.cuda.set_device(self.device)
        self.sim = GpuSimulator(self.device)
        self.ctx = self.sim.ctx
        self._bind_stream()
        self.hist_dev = None
        self.hist_pinned = None
        self.n_matchups = 0

    def close(self):
        self.sim.close()

    def _bind_stream(self):
        # torch's default stream has handle 0, which mlbsim_set_stream reads as "own stream":
        # name the legacy default stream explicitly (cudaStreamLegacy == 1)
        self.ctx.set_stream(self.torch.cuda.current_stream().cuda_stream or 1)

    def _ensure(self, n_matchups: int):
        torch = self.torch
        if self.n_matchups != n_matchups:
            words = n_matchups * _abi.HIST_WORDS
            self.hist_dev = torch.zeros(words, dtype=torch.int64, device=f"cuda:{self.device}")
            self.hist_pinned = torch.zeros(words, dtype=torch.int64).pin_memory()
            swords = n_matchups * (_abi.SUMMARY_DTYPE.itemsize // 8)
            self.summary_dev = torch.zeros(swords, dtype=torch.int64, device=f"cuda:{self.device}")
            self.summary_pinned = torch.zeros(swords, dtype=torch.int64).pin_memory()
            self.n_matchups = n_matchups

    def load_matchups(self, matchups, use_noise=True):
        self.sim.load_matchups(matchups, use_noise)
        self._ensure(self.ctx.n_tables)
        return self

    def load_tables(self, tbl, matchups=None):
        """`tbl`: TABLE_DTYPE array (may live in pinned memory); one H2D copy."""
        self.sim.load_tables(tbl, matchups)
        self._ensure(self.ctx.n_tables)
        return self

    # ---- device-resident step: tables already in HBM, result stays in HBM ----
    def run_device(self, n_sims_total: int, seed: int, sim_lo: int = 0, count_pa: bool = True):
        """Zero the device histograms, simulate this rank's shard, all-reduce.  Asynchronous."""
        self._bind_stream()
        lo, hi = shard_range(sim_lo, n_sims_total, self.rank, self.world)
        self.hist_dev.zero_()
        self.ctx.run_native(0, self.n_matchups, lo, hi, seed, self.hist_dev.data_ptr(), 0 if count_pa else 1)
        if self.world > 1:
            self.dist.all_reduce(self.hist_dev, op=self.dist.ReduceOp.SUM)
        return self.hist_dev

    # ---- end-to-end call: host tables in, host histograms out ----
    def simulate(self, n_sims_total: int, seed: int, sim_lo: int = 0, tables_host=None, count_pa: bool = True) -> np.ndarray:
        """The call a user makes: (optionally) upload host tables, simulate, all-reduce, copy the
        totals back.  Returns HIST_DTYPE[n_matchups] (every rank gets the full-job totals)."""
        if tables_host is not None:
            self.load_tables(tables_host, self.sim.matchups if len(self.sim.matchups) == len(np.atleast_1d(tables_host)) else None)
        self.run_device(n_sims_total, seed, sim_lo, count_pa)
        self.hist_pinned.copy_(self.hist_dev, non_blocking=True)
        self.torch.cuda.current_stream().synchronize()
        out = self.hist_pinned.numpy().view(np.uint64).view(_abi.HIST_DTYPE).copy()
        check_terminated(out)
        return out

    def simulate_summary(self, n_sims_total: int, seed: int, sim_lo: int = 0, tables_host=None, count_pa: bool = True) -> np.ndarray:
        """Sweep-shaped end-to-end call: as `simulate`, but the all-reduced histograms are reduced on the device
        (`mlbsim_summarize`) and only SUMMARY_DTYPE[n_matchups] (320 B each) crosses PCIe."""
        if tables_host is not None:
            self.load_tables(tables_host, self.sim.matchups if len(self.sim.matchups) == len(np.atleast_1d(tables_host)) else None)
        self.run_device(n_sims_total, seed, sim_lo, count_pa)
        self.ctx.summarize(self.hist_dev.data_ptr(), self.n_matchups, self.summary_dev.data_ptr())
        self.summary_pinned.copy_(self.summary_dev, non_blocking=True)
        self.torch.cuda.current_stream().synchronize()
        out = self.summary_pinned.numpy().view(np.uint64).view(_abi.SUMMARY_DTYPE).copy()
        check_terminated(out)
        return out
