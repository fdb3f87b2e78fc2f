#!/usr/bin/env python
"""bench.py — simulated full games / second of the Monte Carlo game sampler on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload single|slate]

One step = one pass of the hot path over one batch: `--sims` simulations (default 10 000 000) of
every matchup of the workload PER GPU (weak scaling: rank r simulates its own slice of the sim-index
range; histograms are combined with one NCCL all-reduce).  Default workload = BASELINE.json
configs[1]: the single game 2025-04-04-TEX@HOU on synthetic Batters/Pitchers.csv-schema players,
joint runs PMF at caps 1/3/5/7/full + PA outcome counters + reliever usage.  Prints ONE JSON line.
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
    from mlb_odds_engine_b200 import synth

    if args.workload == "single":
        ms = [synth.make_matchup("2025-04-04-TEX@HOU")]
        sims = args.sims or 10_000_000
        name = f"single game 2025-04-04-TEX@HOU x {_count(sims)} sims per GPU per step (BASELINE configs[{0 if sims <= 10_000 else 1}])"
    elif args.workload == "fatigue":
        ms = [synth.make_matchup("2025-04-04-TEX@HOU", bullpens=False)]   # starters go the distance: TTO 4+, pitch-count ramp
        sims = args.sims or 10_000_000
        name = f"fatigue-heavy: single game without bullpens x {_count(sims)} sims per GPU per step (BASELINE configs[3] shape)"
    elif args.workload == "slate":
        ms = synth.make_slate("2025-04-04", 15)
        sims = args.sims or 1_000_000
        name = f"15-game slate x {_count(sims)} sims per game per GPU per step (BASELINE configs[2])"
    else:
        import numpy as np
        from mlb_odds_engine_b200 import sweep

        base = synth.make_matchup("2025-04-04-TEX@HOU", bullpens=False)   # the tools/ scripts simulate without bullpens
        axes = dict(k_rate=np.linspace(0.14, 0.34, 16), bb_rate=np.linspace(0.04, 0.12, 8), hr_pa_projected=np.linspace(0.015, 0.06, 8))
        ms, _ = sweep.grid_matchups(base, "home", **axes)    # dicts: only the CPU-baseline legs read them
        args._grid = (base, axes)                            # the GPU arm builds its tables with sweep.grid_tables
        sims = args.sims or 100_000
        name = f"sensitivity sweep: 1024 grid points (k_rate x bb_rate x hr_pa) x {_count(sims)} sims per GPU per step (BASELINE configs[4])"
    return ms, sims, name


def _count(n: int) -> str:
    return f"{n // 1_000_000}M" if n % 1_000_000 == 0 and n else f"{n // 1000}k" if n % 1000 == 0 and n else str(n)


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


def run_reference(args):
    """--impl reference: the reference's CPU algorithm for this path (oracle port: the reference is
    Python and /root/reference does not exist on the GPU box) on all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle as orc
    from mlb_odds_engine_b200 import tables

    ms, sims, name = workload(args)
    cores = orc.get_threads()
    params = [tables.build_params(m) for m in ms]
    # each step = a bounded sample of the workload: ~2 s of CPU work, less when K is large so that the
    # whole run (warm-up + K steps) stays around two minutes
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
    v = games / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": name, "sample_per_step": f"{per_step} games per matchup per step ({len(params)} matchups)"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{games} games: C port of the reference algorithm (oracle_grammar, Beta noise sampled), "
                                   f"{cores} threads; the Python reference itself runs ~450 games/s/core (BASELINE.md)"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


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
        os.environ["NCCL_DEBUG"] = os.environ.get("MLBSIM_NCCL_DEBUG", "WARN")  # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))

    from mlb_odds_engine_b200 import _abi, tables
    from mlb_odds_engine_b200.dist import ShardedSimulator

    ms, sims_per_gpu, wname = workload(args)
    n_m = len(ms)
    import scipy.special, scipy.stats  # noqa: F401,E401  (one-off import cost kept out of the build time below)
    tables._CLIP_CACHE.clear()
    t_build = time.perf_counter()
    if getattr(args, "_grid", None):
        from mlb_odds_engine_b200 import sweep
        tbl = sweep.grid_tables(args._grid[0], "home", **args._grid[1])[0]
    else:
        tbl = tables.build_tables(ms)
    t_build = time.perf_counter() - t_build
    tbl_pinned = torch.from_numpy(tbl.view(np.uint8).reshape(-1).copy()).pin_memory()
    tbl_host = tbl_pinned.numpy().view(_abi.TABLE_DTYPE)

    S = ShardedSimulator(local)
    if args.threads_per_block or args.blocks_per_sm:
        S.ctx.set_launch_config(args.threads_per_block, args.blocks_per_sm)
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

    cpu_baseline = None
    if not args.no_cpu_baseline:
        v, cores, done, el = cpu_port_throughput(ms, args.cpu_seconds)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"{done} games of the same workload in {el:.1f} s: C port of the reference algorithm "
                                  f"(oracle_grammar: reference draw grammar, Beta noise sampled) on {cores} threads; "
                                  "the Python reference itself measured 417-485 games/s/core (BASELINE.md)"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic",
        "config": {"workload": wname, "matchups": n_m, "sims_per_matchup_per_gpu_per_step": sims_per_gpu,
                   "rng": "Philox2x32-10, one block per plate appearance, counter=(sim index, PA number, matchup), key=f(seed)",
                   "parallelism": f"sim-index shards x{world}, uint64 histogram all-reduce" if world > 1 else "1 GPU",
                   "l2": "256 MB device buffer rewritten between timed iterations (L2 flush); every step simulates a fresh sim-index range",
                   "threads_per_block": args.threads_per_block or 640, "table_build_ms_once": 1e3 * t_build},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(tbl.nbytes),
                "d2h_bytes_per_step": int(n_m * e2e_out_bytes)},
        "parity": parity,
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
        "wall_s_timed_region": t_wall,
        "device": info["name"],
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
