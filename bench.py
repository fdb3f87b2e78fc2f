#!/usr/bin/env python
"""bench.py — simulated full games / second of the Monte Carlo game sampler on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload single|slate]

One step = one pass of the hot path over one batch: `--sims` simulations (default 10 000 000) of
every matchup of the workload PER GPU (weak scaling: rank r simulates its own slice of the sim-index
range; histograms are combined with one NCCL all-reduce).  Default workload = BASELINE.json
configs[1]: the single game 2025-04-04-TEX@HOU on synthetic Batters/Pitchers.csv-schema players,
joint runs PMF at caps 1/3/5/7/full + PA outcome counters + reliever usage.  Prints ONE JSON line.

Behind the headline the same line carries a `workloads` block (skipped with --no-workloads): BASELINE configs[1-4] as
FIXED-SIZE jobs split over the N ranks (strong scaling: single game 10M, 15-game slate x 1M, fatigue game 100M by sim
index; the 1 024-point sweep x 100k by grid-point block), each with games/s, ms per step and the sha256 of a small
fixed-(seed, sims) result — the digest is the same for every N (results are keyed by (seed, matchup, sim)) and is
pinned against the CPU oracle twin in tests/golden/bench_digests.json.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import bench_workloads as bw  # noqa: E402

METRIC = "simulated_games_per_sec"
UNIT = "games/s"
I_ALG = 7150.0  # algorithmic thread-instructions per simulated game


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="single", choices=["single", "slate", "sweep", "fatigue"])
    ap.add_argument("--sims", type=int, default=None, help="sims per matchup per GPU per step")
    ap.add_argument("--seed", type=int, default=20250404)
    ap.add_argument("--threads-per-block", type=int, default=0)
    ap.add_argument("--blocks-per-sm", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-workloads", action="store_true", help="skip the extra BASELINE workloads block (tuning runs)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    return ap.parse_args()


def workload(args):
    ms = bw.matchups(args.workload)
    if args.workload == "single":
        sims = args.sims or 10_000_000
        name = f"single game {bw.GAME_ID} x {bw.count(sims)} sims per GPU per step (BASELINE configs[{0 if sims <= 10_000 else 1}])"
    elif args.workload == "fatigue":
        sims = args.sims or 10_000_000
        name = f"fatigue-heavy: single game without bullpens x {bw.count(sims)} sims per GPU per step (BASELINE configs[3] shape)"
    elif args.workload == "slate":
        sims = args.sims or 1_000_000
        name = f"15-game slate x {bw.count(sims)} sims per game per GPU per step (BASELINE configs[2])"
    else:
        args._grid = bw.sweep_grid()                         # the GPU arm builds its tables with sweep.grid_tables
        sims = args.sims or 100_000
        name = f"sensitivity sweep: 1024 grid points (k_rate x bb_rate x hr_pa) x {bw.count(sims)} sims per GPU per step (BASELINE configs[4])"
    return ms, sims, name


class ClockSampler:
    """SM clock, power and throttle reasons sampled DURING the timed region (B200_PROFILING.md): an
    in-process NVML thread at 5 ms (nvidia_ml_py), falling back to `nvidia-smi -lms`."""
    REASONS = (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20))

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.samples = []       # (sm_mhz, power_w, reasons_mask)
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        self._mode = None
        self._proc = None

    def _nvml_loop(self, h, nv):
        while not self._stop.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.samples.append((float(sm), float(pw), int(rs)))
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            # torch's device index follows CUDA_VISIBLE_DEVICES; NVML wants the physical device
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = self.gpu
            if vis:
                ids = [x for x in vis.split(",") if x.strip() != ""]
                if self.gpu < len(ids) and ids[self.gpu].strip().isdigit():
                    phys = int(ids[self.gpu])
            h = nv.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self._mode = "nvml"
            self._thread = threading.Thread(target=self._nvml_loop, args=(h, nv), daemon=True)
            self._thread.start()
            return
        except Exception:
            pass
        try:
            q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_power_cap,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown")
            self._proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "20",
                                           "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self._mode = "smi"
            self._thread = threading.Thread(target=self._smi_loop, daemon=True)
            self._thread.start()
        except Exception:
            self._mode = None

    def _smi_loop(self):
        for line in self._proc.stdout:
            f = [x.strip() for x in line.split(",")]
            try:
                mask = 0
                for (name, bit), val in zip(self.REASONS, f[3:7]):
                    if val.lower().startswith("active"):
                        mask |= bit
                self.samples.append((float(f[0]), float(f[2]), mask))
                self.max_mhz = float(f[1])
            except Exception:
                continue

    def mark(self):
        """Index of the next sample: bracket the timed region with two marks."""
        return len(self.samples)

    def stop(self, lo=0, hi=None):
        self._stop.set()
        if self._proc:
            self._proc.terminate()
        if self._thread:
            self._thread.join(timeout=1.0)
        if self._mode is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"]}
        window = self.samples[lo:hi] or self.samples[max(0, lo - 1):]   # a region shorter than one period: nearest sample
        if not window:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["no samples"]}
        mask = 0
        for _, _, r in window:
            mask |= r
        return {"sm_mhz": statistics.median(s for s, _, _ in window), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(name for name, bit in self.REASONS if mask & bit),
                "power_w_max": max(p for _, p, _ in window), "samples": len(window), "source": self._mode}


def cpu_port_throughput(matchups, seconds, use_noise=True):
    """The oracle's reference-grammar engine (C port of the reference algorithm incl. the Beta noise
    draws) on all host cores, bounded to ~`seconds` of wall time."""
    import oracle as orc
    from mlb_odds_engine_b200 import tables

    params = [tables.build_params(m, use_noise=use_noise) for m in matchups]
    cores = orc.get_threads()
    batch = 50_000 * max(1, cores // 2)
    done, t0, lo = 0, time.perf_counter(), 0
    orc.grammar(params[0], 1, 0, 20_000)  # warm-up
    t0 = time.perf_counter()
    i = 0
    while True:
        orc.grammar(params[i % len(params)], 12345, lo, lo + batch)
        lo += batch; done += batch; i += 1
        el = time.perf_counter() - t0
        if el >= seconds:
            break
    return done / el, cores, done, el


def cpu_reference_throughput(matchups, seconds, procs=None):
    """The UNMODIFIED reference `simulate_game` (core/game_simulator.py:14-156: the loop body of
    cli/run_distribution_simulator.py:368-382) on every host core for ~`seconds`: from /root/reference in the build
    container, from the byte code oracle/make_ref.py compiled into oracle/_ref/ on the GPU box.  None if neither exists."""
    import ref_harness as rh

    if not rh.reference_available():
        return None
    try:
        games, busy, procs, _, wall = rh.timed_batch(matchups, seconds, seed=1, procs=procs)
    except Exception as e:   # a baseline that cannot run must not take the measurement down with it
        sys.stderr.write(f"bench.py: the reference could not be run ({type(e).__name__}: {e}); reporting the C port only\n")
        return None
    return {"value": games / busy, "unit": UNIT, "cores": procs, "kind": "reference",
            "sample": f"{games} games of the same workload in {busy:.1f} s: the reference's own simulate_game "
                      f"(core/game_simulator.py, {rh.reference_kind()} of the unmodified checkout) in {procs} processes, "
                      f"reliever usage counts pre-saturated; {games / busy / procs:.0f} games/s/core"}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on all host cores — the real Python
    `simulate_game` when it is available (oracle/_ref byte code on the GPU box), else the C port of its algorithm.
    Each step is a bounded sample of the workload (a fixed number of games per worker process)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp

    import oracle as orc
    import ref_harness as rh
    from mlb_odds_engine_b200 import tables

    ms, sims, name = workload(args)
    # the C port beside it (a few seconds): what the algorithm costs without the interpreter
    pv, pcores, pdone, pel = cpu_port_throughput(ms, 3.0)
    cpu_port = {"value": pv, "unit": UNIT, "cores": pcores, "kind": "port",
                "sample": f"{pdone} games in {pel:.1f} s: C port of the reference algorithm (oracle_grammar) on {pcores} threads"}
    if not rh.reference_available():
        # fall back to the C port as the reference arm (oracle/_ref was not built: no reference checkout at build time)
        params = [tables.build_params(m) for m in ms]
        cores = orc.get_threads()
        probe_n = 20_000
        t = time.perf_counter(); orc.grammar(params[0], 1, 0, probe_n); rate = probe_n / (time.perf_counter() - t)
        step_seconds = min(2.0, 120.0 / max(1, args.steps + args.warmup))
        per_step = max(2_000, int(rate * step_seconds / len(params)))
        lo = 0
        for _ in range(args.warmup):
            for p in params:
                orc.grammar(p, args.seed, lo, lo + per_step)
            lo += per_step
        t0 = time.perf_counter()
        for _ in range(args.steps):
            for p in params:
                orc.grammar(p, args.seed, lo, lo + per_step)
            lo += per_step
        dt = time.perf_counter() - t0
        games = per_step * len(params) * args.steps
        kind, cores_used = "port", cores
        sample = (f"{games} games: C port of the reference algorithm (oracle_grammar, Beta noise sampled), {cores} threads; "
                  "oracle/_ref (the reference's byte code) was not available")
    else:
        procs = os.cpu_count() or 1
        rh.load_reference()
        for m in ms:
            rh.saturate_reliever_counts(m)
        # games per worker per step: ~step_seconds of work at the probed single-process rate
        step_seconds = min(2.0, 120.0 / max(1, args.steps + args.warmup))
        t = time.perf_counter()
        rh._fixed_worker((ms[:1], 40, 1, True))
        rate1 = 40 / (time.perf_counter() - t)
        per_worker = max(len(ms), int(rate1 * step_seconds * 0.8))
        per_worker -= per_worker % len(ms)
        per_worker = max(per_worker, len(ms))
        # a step = 4 tasks per worker process handed out dynamically, so that one slow core does not hold the step up
        slices = 4
        per_task = max(len(ms), (per_worker // slices) - (per_worker // slices) % len(ms))
        per_worker = per_task * slices
        with mp.get_context("fork").Pool(procs) as pool:
            def step(i):
                jobs = [(ms, per_task, args.seed * 100_003 + (i * procs + w) * slices + t, True) for w in range(procs) for t in range(slices)]
                return sum(pool.imap_unordered(rh._fixed_worker, jobs))
            for i in range(args.warmup):
                step(i)
            t0 = time.perf_counter()
            games = sum(step(args.warmup + i) for i in range(args.steps))
            dt = time.perf_counter() - t0
        kind, cores_used = "reference", procs
        sample = (f"{games} games ({per_worker} per worker process per step, {procs} processes): the reference's own simulate_game "
                  f"(core/game_simulator.py, {rh.reference_kind()} of the unmodified checkout), reliever usage counts pre-saturated; "
                  f"{games / dt / procs:.0f} games/s/core")
    v = games / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": name, "matchups": len(ms), "sims_per_matchup_per_gpu_per_step": sims,
                   "sample_per_step": f"{games // max(1, args.steps)} games per step (bounded sample of the workload)"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores_used, "kind": kind, "sample": sample},
        "cpu_port": cpu_port,
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def extra_workloads(args, S_sims, local, world, rank, torch, dist):
    """BASELINE configs[1-4] as fixed-size jobs over the N ranks (strong scaling), device-timed with CUDA events, max
    over ranks; plus, per workload, the digest of a small fixed-(seed, sims) result, which must not depend on N."""
    import numpy as np

    from mlb_odds_engine_b200 import _abi
    from mlb_odds_engine_b200.dist import ShardedSimulator

    try:
        with open(os.path.join(ROOT, "tests", "golden", "bench_digests.json")) as f:
            pinned = json.load(f)
    except Exception:
        pinned = {}
    S_blk = None
    out = {}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local}")
    for name, spec in bw.SPEC.items():
        if spec["shard"] == "matchups":
            S_blk = S_blk or ShardedSimulator(local, shard="matchups")
            S = S_blk
        else:
            S = S_sims
        t_tbl = time.perf_counter()          # dicts / grid axes -> mlbsim_params on the host -> tables built in HBM
        params = bw.build_params(name)
        S.load_params(params)
        t_tbl = time.perf_counter() - t_tbl
        n_m = len(params)
        sims = spec["sims"]
        steps, warm = (5, 2) if n_m * sims <= 20_000_000 else (3, 1)

        def one(lo):
            S.run_device(sims, bw.SEED, sim_lo=lo)
            if spec["shard"] == "matchups":
                S.summarize_device()

        for i in range(warm):
            one(i * sims)
            flush.fill_(1)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        launches0 = S.ctx.launch_count()
        evs = []
        for i in range(steps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            one((warm + i) * sims)
            e1.record()
            evs.append((e0, e1))
            flush.fill_(i & 0xFF)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms_total = torch.tensor([sum(a.elapsed_time(b) for a, b in evs)], dtype=torch.float64, device=f"cuda:{local}")
        if world > 1:
            dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
        ms_total = float(ms_total.item())
        # digest step: small fixed job, full result back on the host
        if spec["shard"] == "matchups":
            rec = S.simulate_summary(spec["digest_sims"], bw.SEED, sim_lo=0)
        else:
            rec = S.simulate(spec["digest_sims"], bw.SEED, sim_lo=0)
        assert int(rec["n_games"].sum()) == n_m * spec["digest_sims"], name
        d = bw.digest(rec)
        out[name] = {
            "baseline_config": spec["config"], "what": spec["what"], "scaling": spec["scaling"], "shard": spec["shard"],
            "matchups": n_m, "sims_per_matchup": sims, "steps": steps, "warmup": warm, "table_build_ms_once": 1e3 * t_tbl,
            "ms_per_step": ms_total / steps, "games_per_s": n_m * sims * steps / (ms_total * 1e-3),
            "gpu_launches": int(S.ctx.launch_count() - launches0),
            "exchange": ("all-gather of %d B summaries" % (n_m * _abi.SUMMARY_DTYPE.itemsize) if spec["shard"] == "matchups"
                         else "all-reduce of %d B histograms" % (n_m * _abi.HIST_DTYPE.itemsize)) if world > 1 else "none (1 GPU)",
            "digest": {"sims_per_matchup": spec["digest_sims"], "seed": bw.SEED, "sha256": d,
                       "matches_oracle_twin": (d == pinned[name]["sha256"]) if name in pinned else None},
        }
    if S_blk is not None:
        S_blk.close()
    return out


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL_DEBUG is left exactly as the caller set it; its log lines go to stderr (unless the caller chose a file)
        # so that stdout stays the one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))

    from mlb_odds_engine_b200 import _abi, tables
    from mlb_odds_engine_b200.dist import ShardedSimulator

    ms, sims_per_gpu, wname = workload(args)
    n_m = len(ms)
    import scipy.special, scipy.stats  # noqa: F401,E401  (one-off import cost kept out of the build time below)
    S = ShardedSimulator(local)
    if args.threads_per_block or args.blocks_per_sm:
        S.ctx.set_launch_config(args.threads_per_block, args.blocks_per_sm)
    S.ctx.build_tables(tables.params_batch(ms[:1]))     # (first use of the device table builder: module load kept out of the timing)
    t_build = time.perf_counter()               # dicts / grid axes -> mlbsim_params on the host -> tables built in HBM
    if getattr(args, "_grid", None):
        from mlb_odds_engine_b200 import sweep
        params = sweep.grid_params(args._grid[0], "home", **args._grid[1])[0]
    else:
        params = tables.params_batch(ms)
    S.load_params(params, ms)
    t_build = time.perf_counter() - t_build
    tbl = S.sim.tables                          # (downloaded: the e2e leg below uploads prebuilt tables from pinned memory)
    tbl_pinned = torch.from_numpy(tbl.view(np.uint8).reshape(-1).copy()).pin_memory()
    tbl_host = tbl_pinned.numpy().view(_abi.TABLE_DTYPE)

    S.load_tables(tbl_host, ms)
    info = S.ctx.device_info()

    total_per_step = sims_per_gpu * world           # sims per matchup per step over all ranks
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local}")  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    step_no = [0]

    def device_step(ev=None):
        lo = step_no[0] * total_per_step
        step_no[0] += 1
        if ev:
            ev[0].record()
        S.run_device(total_per_step, args.seed, sim_lo=lo)
        if ev:
            ev[1].record()

    # ---------------- device-resident timing ----------------
    sampler = ClockSampler(local)
    sampler.start()          # already running through the warm-up, so the timed region is covered from its first ms
    # The W warm-up steps of the contract, preceded by up to 0.3 s of the same steps when W of them would be shorter than
    # that: an idle GPU needs tens of milliseconds to reach its boost clock, and W = 3 steps of this kernel are 10 ms.
    # (a step COUNT, the same on every rank — the steps hold a collective — sized to ~0.3 s at 3 x 10^9 games/s)
    prewarm_steps = min(400, max(3, int(1e9 / max(1, n_m * sims_per_gpu))))
    for _ in range(prewarm_steps):
        device_step()
        flush.fill_(1)
    for _ in range(args.warmup):
        device_step()
        flush.fill_(1)
    barrier()
    mark_lo = sampler.mark()
    launches0 = S.ctx.launch_count()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kev = []
    t_wall = time.perf_counter()
    for i in range(args.steps):
        device_step(evs[i])
        flush.fill_(i & 0xFF)   # evict the 126 MB L2 between timed iterations
    barrier()
    t_wall = time.perf_counter() - t_wall
    launches = S.ctx.launch_count() - launches0
    clocks = sampler.stop(mark_lo, sampler.mark())
    step_ms = [a.elapsed_time(b) for a, b in evs]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    games = n_m * total_per_step * args.steps
    value = games / (total_ms * 1e-3)

    # kernel-only duration (events recorded by the library around the kernel on the same stream)
    device_step()
    kern_ms = S.ctx.last_kernel_ms()
    per_gpu_rate_kernel = n_m * sims_per_gpu / (kern_ms * 1e-3)

    # sanity: totals of the last step
    hist = S.hist_dev.cpu().numpy().view(np.uint64).view(_abi.HIST_DTYPE)
    assert int(hist["n_games"].sum()) == n_m * total_per_step, "histogram does not hold every simulated game"

    # ---------------- end-to-end timing (host tables in, host histograms out) ----------------
    # (the sweep's public call returns per-grid-point summaries reduced on the device; the others full histograms)
    e2e_call = S.simulate_summary if args.workload == "sweep" else S.simulate
    e2e_out_bytes = (_abi.SUMMARY_DTYPE if args.workload == "sweep" else _abi.HIST_DTYPE).itemsize
    for _ in range(args.warmup):
        e2e_call(total_per_step, args.seed, sim_lo=step_no[0] * total_per_step, tables_host=tbl_host)
        step_no[0] += 1
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = e2e_call(total_per_step, args.seed, sim_lo=step_no[0] * total_per_step, tables_host=tbl_host)
        step_no[0] += 1
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = games / float(e2e_s.item())
    assert int(out["n_games"].sum()) == n_m * total_per_step

    # ---------------- parity figures of BASELINE.json's metric (untimed): the last e2e step vs the reference ----------------
    parity = None
    fixture = os.path.join(ROOT, "tests", "golden", "stats_bench_texhou.npz")
    if rank == 0 and args.workload == "single" and os.path.exists(fixture):
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from helpers import chi2_two_sample, load_stats, win_over_stats, z_two_proportions

        g = load_stats("bench_texhou")
        n_gpu = int(out["n_games"][0])
        parity = {"vs": f"tests/golden/stats_bench_texhou.npz ({g['n']} games of the unmodified reference, same matchup)",
                  "gpu_games": n_gpu, "ref_games": g["n"], "joint_runs_pmf_chi2": {}, "probabilities": {}}
        for c, cap in enumerate(("f1", "f3", "f5", "f7", "full")):
            chi2, dof, pval = chi2_two_sample(g["joint"][c], out["joint"][0][c])
            parity["joint_runs_pmf_chi2"][cap] = {"chi2": round(chi2, 1), "dof": dof, "p": round(pval, 4)}
        ref_ev, gpu_ev = win_over_stats(g["joint"][4]), win_over_stats(out["joint"][0][4])
        for k in ("home_win", "home_minus_1.5", "over_7.5", "over_8.5", "over_9.5"):
            pg, pr = gpu_ev[k] / n_gpu, ref_ev[k] / g["n"]
            parity["probabilities"][k] = {"gpu": round(pg, 5), "ref": round(pr, 5), "delta": round(pg - pr, 5),
                                          "z": round(float(z_two_proportions(gpu_ev[k], n_gpu, ref_ev[k], g["n"])), 2)}

    # ---------------- the other BASELINE configs as fixed-size jobs over the N ranks (strong scaling) + cross-N digests ----------------
    workloads = None
    if not args.no_workloads and args.workload == "single":
        workloads = extra_workloads(args, S, local, world, rank, torch, dist)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- roofline: instruction issue (no HBM traffic, no dense contraction) ----------------
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    sm_max_mhz = float(peaks.get("sm_max_mhz") or clocks.get("sm_max_mhz") or 1965.0)
    peak_tinst = info["sm_count"] * 4 * 32 * sm_max_mhz * 1e6 / 1e12   # thread-instr/s, all 4 schedulers issuing every cycle
    achieved_tinst = per_gpu_rate_kernel * I_ALG / 1e12
    hbm_bytes_per_launch = n_m * (_abi.TABLE_DTYPE.itemsize + _abi.HIST_DTYPE.itemsize)
    traffic = None
    try:   # DRAM bytes of one launch from the committed ncu --set full capture (independent of the sim count: tables + flush)
        with open(os.path.join(ROOT, "profiles", "native_ncu_traffic.json")) as f:
            traffic = json.load(f)["dram_bytes_per_launch"]
    except Exception:
        pass
    roofline = {
        "bound": "issue", "achieved": achieved_tinst, "peak": peak_tinst, "unit": "Tinst/s (thread instructions)",
        "frac": achieved_tinst / peak_tinst, "traffic": traffic,
        "kernel": "mlbsim_native_kernel", "kernel_ms": kern_ms, "games_per_launch": n_m * sims_per_gpu,
        "i_alg_thread_instr_per_game": I_ALG,
        "peak_source": f"{info['sm_count']} SMs x 4 schedulers x 32 lanes x {sm_max_mhz:.0f} MHz (MEASURED_PEAKS.json sm_max_mhz)",
        "hbm": {"algorithmic_bytes_per_launch": hbm_bytes_per_launch,
                "achieved_gbs": hbm_bytes_per_launch / (kern_ms * 1e-3) / 1e9, "peak_gbs": peaks.get("hbm_gbs"),
                "note": "tables (18 KB/matchup) and histograms (165 KB/matchup) only; the path is not HBM-bound"},
    }

    # CPU baselines on this box's host cores (rank 0): the reference's own Python simulate_game (oracle/_ref byte code
    # here, where /root/reference does not exist), and beside it the C port of the same algorithm
    cpu_baseline = cpu_port = None
    if not args.no_cpu_baseline:
        v, cores, done, el = cpu_port_throughput(ms, min(4.0, args.cpu_seconds))
        cpu_port = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                    "sample": f"{done} games of the same workload in {el:.1f} s: C port of the reference algorithm "
                              f"(oracle_grammar: reference draw grammar, Beta noise sampled) on {cores} threads"}
        cpu_baseline = cpu_reference_throughput(ms, args.cpu_seconds) or cpu_port

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic",
        "config": {"workload": wname, "matchups": n_m, "sims_per_matchup_per_gpu_per_step": sims_per_gpu,
                   "rng": "Philox2x32-10, one block per plate appearance, counter=(sim index, PA number, matchup), key=f(seed)",
                   "parallelism": f"sim-index shards x{world}, uint64 histogram all-reduce" if world > 1 else "1 GPU",
                   "l2": "256 MB device buffer rewritten between timed iterations (L2 flush); every step simulates a fresh sim-index range",
                   "threads_per_block": args.threads_per_block or 640, "table_build_ms_once": 1e3 * t_build,
                   "prewarm": f"{prewarm_steps} untimed steps (~0.3 s) before the {args.warmup} warm-up steps: clock ramp"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(tbl.nbytes),
                "d2h_bytes_per_step": int(n_m * e2e_out_bytes)},
        "parity": parity,
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
        "cpu_port": cpu_port,
        "workloads": workloads,
        "wall_s_timed_region": t_wall,
        "device": info["name"],
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
