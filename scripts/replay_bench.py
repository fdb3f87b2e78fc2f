#!/usr/bin/env python
"""Replay-mode throughput on the GPU (tapes resident in HBM): games/s and achieved GB/s.
Tapes come from the oracle's reference-grammar engine (test infrastructure) — this script is a
measurement aid, not part of the product path."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import oracle as orc  # noqa: E402
from mlb_odds_engine_b200 import _abi, synth, tables  # noqa: E402
from mlb_odds_engine_b200._lib import Context  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000
    bullpens = "--no-bullpens" not in sys.argv   # no bullpens: starters pitch on past 75 pitches (fatigue ramp, TTO 4+)
    m = synth.make_matchup("2025-04-04-TEX@HOU", bullpens=bullpens)
    params = tables.build_params(m)
    t0 = time.time()
    cache = f"/tmp/replay_bench_tapes_{n}_{int(bullpens)}.npz"   # (A/B runs of several libraries in one GPU call: generate once)
    if os.path.exists(cache):
        z = np.load(cache)
        flat, offsets, games = z["flat"], z["offsets"], z["games"]
    else:
        parts, lens, recs = [], [], []
        chunk = 25_000
        for lo in range(0, n, chunk):          # chunked: a 4096-draw stride covers any plausible game
            hi = min(n, lo + chunk)
            _, g, tape, tlen = orc.grammar(params, seed=3, game_lo=lo, game_hi=hi, want_games=True, tape_stride=4096)
            mask = np.arange(4096)[None, :] < tlen[:, None]
            parts.append(tape[mask])
            lens.append(tlen)
            recs.append(g)
        tlen = np.concatenate(lens)
        games = np.concatenate(recs)
        flat = np.concatenate(parts)
        offsets = np.zeros(n + 1, dtype=np.uint64)
        offsets[1:] = np.cumsum(tlen)
        try:
            np.savez(cache, flat=flat, offsets=offsets, games=games)
        except OSError:
            pass
    t_gen = time.time() - t0
    ctx = Context(0)
    d_tape = ctx.malloc(flat.nbytes); ctx.h2d(d_tape, flat)
    d_off = ctx.malloc(offsets.nbytes); ctx.h2d(d_off, offsets)
    d_out = ctx.malloc(n * _abi.REPLAY_GAME_DTYPE.itemsize)
    ms = []
    for it in range(8):
        ctx.run_replay_dev(params, d_tape, d_off, n, d_out)
        ms.append(ctx.last_kernel_ms())
    out = np.zeros(n, dtype=_abi.REPLAY_GAME_DTYPE)
    ctx.d2h(out, d_out)
    ok = all((out[f] == games[f]).all() for f in out.dtype.names)
    best = min(ms[2:])
    bytes_alg = flat.nbytes + offsets.nbytes + out.nbytes
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    print(json.dumps({
        "mode": "replay", "bullpens": bullpens, "games": n, "draws": int(len(flat)), "bit_exact_vs_oracle": bool(ok), "kernel_ms": best,
        "games_per_s": n / (best * 1e-3), "algorithmic_bytes": int(bytes_alg), "achieved_gbs": bytes_alg / (best * 1e-3) / 1e9,
        "hbm_peak_gbs": peaks.get("hbm_gbs"), "frac_of_hbm": (bytes_alg / (best * 1e-3) / 1e9) / peaks["hbm_gbs"] if peaks.get("hbm_gbs") else None,
        "tape_generation_s_cpu": t_gen,
    }))


if __name__ == "__main__":
    main()
