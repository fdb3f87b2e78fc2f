"""Sharding over the GPUs of one box.  Every (matchup, sim index) is independent and its random
stream is keyed by (seed, matchup id, sim index) only, so the work can be cut either way:

* `shard="sims"` (single games, slates): rank r of W takes the contiguous slice of the sim range
  [sim_lo + r*n/W, sim_lo + (r+1)*n/W) of every matchup; the only exchange is a sum of
  `n_matchups * HIST_WORDS` uint64 counters (164 KB per matchup) — `torch.distributed.all_reduce`
  (NCCL over NVLink on GPUs, gloo in the CPU tests).
* `shard="matchups"` (sweeps: many matchups, few sims each): rank r owns the block of grid points
  [r*M/W, (r+1)*M/W), uploads only those tables (`mlbsim_set_matchup_base` keeps their global ids in the Philox
  counter), simulates all of their sims, reduces each histogram to its 328-byte summary on the device and the
  ranks all-gather the summaries: no histogram ever crosses a link and per-rank buffers are 1/W of the sweep.

Integer sums / disjoint blocks => results are bit-identical for 1/2/4/8 GPUs.  torch is plumbing here: device
memory, the stream, the collectives.
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


def allgather_summaries_numpy(local: np.ndarray, n_total: int) -> np.ndarray:
    """Matchup-sharded sweeps on the default process group (CPU / gloo tests of the host logic): every rank passes the
    SUMMARY_DTYPE records of ITS block (`shard_range(0, n_total, rank, world)`), gets all `n_total` back in order."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(), dist.get_rank()
    words = _abi.SUMMARY_DTYPE.itemsize // 8
    cap = -(-n_total // world)                      # blocks differ by at most one matchup: pad to the largest
    buf = torch.zeros(cap * words, dtype=torch.int64)
    flat = np.ascontiguousarray(local).reshape(-1).view(np.int64)
    buf[: flat.size] = torch.from_numpy(flat.copy())
    parts = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    out = np.zeros(n_total, dtype=_abi.SUMMARY_DTYPE)
    for r in range(world):
        lo, hi = shard_range(0, n_total, r, world)
        out[lo:hi] = parts[r].numpy()[: (hi - lo) * words].view(np.uint64).view(_abi.SUMMARY_DTYPE)
    assert shard_range(0, n_total, rank, world)[1] - shard_range(0, n_total, rank, world)[0] == len(np.atleast_1d(local))
    return out


def allreduce_hist_numpy(rec: np.ndarray) -> np.ndarray:
    """Sum HIST_DTYPE records over all ranks of the default process group (any backend that can
    reduce int64 on CPU, i.e. gloo); returns a new array.  Used by the CPU tests of the host logic."""
    import torch
    import torch.distributed as dist

    flat = torch.from_numpy(np.ascontiguousarray(rec).reshape(-1).view(np.int64).copy())
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    return flat.numpy().view(np.uint64).view(_abi.HIST_DTYPE).reshape(np.shape(rec))


class ShardedSimulator:
    """One rank's simulator + the exchange step.  Requires torch.cuda; the kernel runs on torch's
    current stream so CUDA events and NCCL order correctly around it."""

    def __init__(self, device: int | None = None, shard: str = "sims"):
        import torch
        import torch.distributed as dist

        from .simulator import GpuSimulator

        assert shard in ("sims", "matchups")
        self.shard = shard
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
        self.n_matchups = 0        # matchups THIS rank simulates (all of them, or its block)
        self.n_total = 0           # matchups of the whole job
        self.m_lo = 0

    def close(self):
        self.sim.close()

    def _bind_stream(self):
        # torch's default stream has handle 0, which mlbsim_set_stream reads as "own stream":
        # name the legacy default stream explicitly (cudaStreamLegacy == 1)
        self.ctx.set_stream(self.torch.cuda.current_stream().cuda_stream or 1)

    def _ensure(self, n_matchups: int, n_total: int):
        torch = self.torch
        if self.n_matchups != n_matchups or self.n_total != n_total:
            dev = f"cuda:{self.device}"
            words = n_matchups * _abi.HIST_WORDS
            self.hist_dev = torch.zeros(words, dtype=torch.int64, device=dev)
            swords = _abi.SUMMARY_DTYPE.itemsize // 8
            self.summary_dev = torch.zeros(max(1, n_matchups) * swords, dtype=torch.int64, device=dev)
            if self.shard == "sims":
                self.hist_pinned = torch.zeros(words, dtype=torch.int64).pin_memory()
                self._hist_host = self.hist_pinned.numpy().view(np.uint64).view(_abi.HIST_DTYPE)   # (one wrapper, not one per call)
                self.summary_pinned = torch.zeros(n_matchups * swords, dtype=torch.int64).pin_memory()
            else:
                # all-gather buffers: one padded block per rank (blocks differ by at most one matchup)
                self.block_cap = -(-n_total // self.world)
                self.gather_send = torch.zeros(self.block_cap * swords, dtype=torch.int64, device=dev)
                self.gather_recv = torch.zeros(self.world * self.block_cap * swords, dtype=torch.int64, device=dev)
                self.summary_pinned = torch.zeros(self.world * self.block_cap * swords, dtype=torch.int64).pin_memory()
            self.n_matchups, self.n_total = n_matchups, n_total

    def load_matchups(self, matchups, use_noise=True, builder="device"):
        from . import tables

        ms = [tables.as_matchup(m, use_noise) for m in matchups]
        if builder == "device":
            return self.load_params(tables.params_batch(ms), ms)
        return self.load_tables(tables.build_tables(ms, alias_fn=self.ctx.alias_rows), ms)

    def load_params(self, params, matchups=None):
        """`params`: PARAMS_DTYPE array of the WHOLE job (`tables.params_batch`, `sweep.grid_params`); this rank's tables
        are built on its own device (csrc/table_kernel.cu) — a matchup shard builds only its block."""
        params = np.atleast_1d(params)
        n_total = len(params)
        if self.shard == "matchups":
            self.m_lo, m_hi = shard_range(0, n_total, self.rank, self.world)
            if m_hi > self.m_lo:
                self.sim.load_params(params[self.m_lo:m_hi], None if matchups is None else list(matchups)[self.m_lo:m_hi])
            self.ctx.set_matchup_base(self.m_lo)
            self._ensure(m_hi - self.m_lo, n_total)
        else:
            self.sim.load_params(params, matchups)
            self._ensure(n_total, n_total)
        return self

    def load_tables(self, tbl, matchups=None):
        """`tbl`: TABLE_DTYPE array of the WHOLE job (may live in pinned memory); one H2D copy of what this rank needs."""
        tbl = np.atleast_1d(tbl)
        n_total = len(tbl)
        if self.shard == "matchups":
            self.m_lo, m_hi = shard_range(0, n_total, self.rank, self.world)
            if m_hi > self.m_lo:
                self.sim.load_tables(tbl[self.m_lo:m_hi], None if matchups is None else list(matchups)[self.m_lo:m_hi])
            self.ctx.set_matchup_base(self.m_lo)
            self._ensure(m_hi - self.m_lo, n_total)
        else:
            self.sim.load_tables(tbl, matchups)
            self._ensure(n_total, n_total)
        return self

    # ---- device-resident step: tables already in HBM, result stays in HBM ----
    def run_device(self, n_sims_total: int, seed: int, sim_lo: int = 0, count_pa: bool = True):
        """Zero the device histograms, simulate this rank's shard, all-reduce (sim shards only).  Asynchronous."""
        self._bind_stream()
        self.hist_dev.zero_()
        if self.shard == "matchups":
            if self.n_matchups:
                self.ctx.run_native(0, self.n_matchups, sim_lo, sim_lo + n_sims_total, seed, self.hist_dev.data_ptr(), 0 if count_pa else 1)
            return self.hist_dev
        lo, hi = shard_range(sim_lo, n_sims_total, self.rank, self.world)
        self.ctx.run_native(0, self.n_matchups, lo, hi, seed, self.hist_dev.data_ptr(), 0 if count_pa else 1)
        if self.world > 1:
            self.dist.all_reduce(self.hist_dev, op=self.dist.ReduceOp.SUM)
        return self.hist_dev

    def summarize_device(self):
        """Per-matchup summaries of the histograms of the last `run_device`, for the WHOLE job, on the device:
        sim shards summarise the all-reduced histograms; matchup shards summarise their block and all-gather
        (328 bytes per grid point is all that crosses a link).  Returns an int64 device tensor whose first
        n_total records (rank-major for matchup shards: see `_unpack_summaries`) are SUMMARY_DTYPE."""
        if self.n_matchups:
            self.ctx.summarize(self.hist_dev.data_ptr(), self.n_matchups, self.summary_dev.data_ptr())
        if self.shard == "sims":
            return self.summary_dev
        swords = _abi.SUMMARY_DTYPE.itemsize // 8
        self.gather_send.zero_()
        self.gather_send[: self.n_matchups * swords].copy_(self.summary_dev[: self.n_matchups * swords])
        if self.world > 1:
            self.dist.all_gather_into_tensor(self.gather_recv, self.gather_send)
        else:
            self.gather_recv.copy_(self.gather_send)
        return self.gather_recv

    def _unpack_summaries(self, flat: np.ndarray) -> np.ndarray:
        if self.shard == "sims":
            return flat.view(np.uint64).view(_abi.SUMMARY_DTYPE).copy()
        swords = _abi.SUMMARY_DTYPE.itemsize // 8
        out = np.zeros(self.n_total, dtype=_abi.SUMMARY_DTYPE)
        for r in range(self.world):
            lo, hi = shard_range(0, self.n_total, r, self.world)
            blk = flat[r * self.block_cap * swords: r * self.block_cap * swords + (hi - lo) * swords]
            out[lo:hi] = blk.view(np.uint64).view(_abi.SUMMARY_DTYPE)
        return out

    # ---- end-to-end calls: host tables in, host results out ----
    def simulate(self, n_sims_total: int, seed: int, sim_lo: int = 0, tables_host=None, count_pa: bool = True) -> np.ndarray:
        """The call a user makes: (optionally) upload host tables, simulate, all-reduce, copy the
        totals back.  Returns HIST_DTYPE[n_matchups] (every rank gets the full-job totals).  Sim shards only."""
        assert self.shard == "sims", "full histograms are exchanged by sim shards; matchup shards return summaries"
        if tables_host is not None:
            self.load_tables(tables_host, self.sim.matchups if len(self.sim.matchups) == len(np.atleast_1d(tables_host)) else None)
        self.run_device(n_sims_total, seed, sim_lo, count_pa)
        self.hist_pinned.copy_(self.hist_dev, non_blocking=True)
        self.torch.cuda.current_stream().synchronize()
        out = self._hist_host.copy()
        check_terminated(out)
        return out

    def simulate_summary(self, n_sims_total: int, seed: int, sim_lo: int = 0, tables_host=None, count_pa: bool = True) -> np.ndarray:
        """Sweep-shaped end-to-end call: upload, simulate, reduce every histogram to its summary on the device
        (`mlbsim_summarize`), exchange, and only SUMMARY_DTYPE[n_total] (328 B each) crosses PCIe."""
        if tables_host is not None:
            keep = self.sim.matchups if self.shard == "sims" and len(self.sim.matchups) == len(np.atleast_1d(tables_host)) else None
            self.load_tables(tables_host, keep)
        self.run_device(n_sims_total, seed, sim_lo, count_pa)
        dev = self.summarize_device()
        self.summary_pinned.copy_(dev[: self.summary_pinned.numel()], non_blocking=True)
        self.torch.cuda.current_stream().synchronize()
        out = self._unpack_summaries(self.summary_pinned.numpy())
        check_terminated(out)
        return out
