"""Application-level half of the generator battery: does a build of the library with fewer Philox rounds
produce the same DISTRIBUTION of games as the 10-round product, at a resolution far beyond what pricing needs?

    python scripts/rng_equivalence.py run --lib variants/libmlbsim_r7.so --sims 20000000000 --seed 1 --out gpurun_out/eq_r7_s1.npz
    python scripts/rng_equivalence.py compare gpurun_out/eq_r10_s1.npz gpurun_out/eq_r7_s1.npz [...]

`run` (one process per library: MLBSIM_LIB is read at import) plays `--sims` games of the bench matchup and of the
fatigue-heavy matchup in launches of 2 x 10^8 and stores the summed histograms; `compare` prints, for each pair
(first file = baseline), two-sample chi-square p-values over the joint (away, home) run cells of every innings cap, the
innings histogram, the plate-appearance outcome counts and the reliever usage, plus the largest |z| of the headline
probabilities (home win, over 8.5, first-5 home lead).  Two different SEEDS of the same library give the null calibration.
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CHUNK = 200_000_000
FIELDS = ("joint", "innings", "pa", "rel_games", "rel_picks", "n_games", "n_truncated")


def run(args) -> int:
    if args.lib:
        os.environ["MLBSIM_LIB"] = os.path.abspath(args.lib)
    import bench_workloads as bw
    from mlb_odds_engine_b200.simulator import GpuSimulator

    out = {}
    for name in ("single", "fatigue"):
        sim = GpuSimulator(0).load_matchups(bw.matchups(name))
        acc = None
        done = 0
        while done < args.sims:
            n = min(CHUNK, args.sims - done)
            rec = sim.ctx.run_native_host(0, 1, done, done + n, args.seed, 0)[0]
            if acc is None:
                acc = {f: np.array(rec[f], dtype=np.uint64) for f in FIELDS}
            else:
                for f in FIELDS:
                    acc[f] += rec[f]
            done += n
        assert int(acc["n_games"]) == args.sims and int(acc["n_truncated"]) == 0
        for f in FIELDS:
            out[f"{name}_{f}"] = acc[f]
    np.savez_compressed(args.out, **out)
    print("wrote", args.out, {k: int(v) for k, v in out.items() if k.endswith("n_games")})
    return 0


def _two_sample(a, b, min_count: float = 40.0):
    """chi-square homogeneity test of two count vectors (cells with fewer than `min_count` games in total are pooled)."""
    from scipy import stats

    a, b = np.asarray(a, dtype=np.float64).ravel(), np.asarray(b, dtype=np.float64).ravel()
    tot = a + b
    small = tot < min_count
    if small.any():
        a = np.concatenate([a[~small], [a[small].sum()]])
        b = np.concatenate([b[~small], [b[small].sum()]])
        tot = a + b
    keep = tot > 0
    a, b, tot = a[keep], b[keep], tot[keep]
    na, nb = a.sum(), b.sum()
    ea, eb = tot * na / (na + nb), tot * nb / (na + nb)
    c = float(((a - ea) ** 2 / ea + (b - eb) ** 2 / eb).sum())
    dof = len(tot) - 1
    return c, dof, float(stats.chi2.sf(c, dof))


def _prob_z(ka, na, kb, nb):
    pa, pb, p = ka / na, kb / nb, (ka + kb) / (na + nb)
    return float((pa - pb) / np.sqrt(p * (1 - p) * (1 / na + 1 / nb))), float(pa), float(pb)


def compare(args) -> int:
    base = np.load(args.files[0])
    for path in args.files[1:]:
        other = np.load(path)
        print(f"== {os.path.basename(args.files[0])}  vs  {os.path.basename(path)}")
        for name in ("single", "fatigue"):
            n = int(base[f"{name}_n_games"])
            row = [f"{name} ({n:.3g} games each)"]
            ja, jb = base[f"{name}_joint"], other[f"{name}_joint"]
            for cap, label in enumerate(("f1", "f3", "f5", "f7", "full")):
                c, dof, p = _two_sample(ja[cap], jb[cap])
                row.append(f"{label}: chi2 {c:.0f}/{dof} p={p:.3g}")
            for f in ("innings", "pa", "rel_games", "rel_picks"):
                if base[f"{name}_{f}"].sum() == 0:
                    continue
                c, dof, p = _two_sample(base[f"{name}_{f}"], other[f"{name}_{f}"])
                row.append(f"{f}: chi2 {c:.1f}/{dof} p={p:.3g}")
            full_a, full_b = ja[4].astype(np.float64), jb[4].astype(np.float64)
            away, home = np.indices(full_a.shape)
            probes = {"home_win": home > away, "over_8.5": (home + away) > 8.5, "total_even": ((home + away) % 2) == 0}
            f5a, f5b = ja[2].astype(np.float64), jb[2].astype(np.float64)
            zs = []
            for label, mask in probes.items():
                z, pa, pb = _prob_z(full_a[mask].sum(), full_a.sum(), full_b[mask].sum(), full_b.sum())
                zs.append(f"{label} {pa:.6f} / {pb:.6f} z={z:+.2f}")
            z, pa, pb = _prob_z(f5a[home > away].sum(), f5a.sum(), f5b[home > away].sum(), f5b.sum())
            zs.append(f"f5_home_lead {pa:.6f} / {pb:.6f} z={z:+.2f}")
            print("  " + "\n    ".join(row))
            print("    " + "; ".join(zs))
    return 0


def main() -> int:
    ap = argparse.ArgumentParser()
    sub = ap.add_subparsers(dest="cmd", required=True)
    r = sub.add_parser("run")
    r.add_argument("--lib", default=None)
    r.add_argument("--sims", type=int, default=10_000_000_000)
    r.add_argument("--seed", type=int, default=1)
    r.add_argument("--out", required=True)
    c = sub.add_parser("compare")
    c.add_argument("files", nargs="+")
    args = ap.parse_args()
    return run(args) if args.cmd == "run" else compare(args)


if __name__ == "__main__":
    sys.exit(main())
