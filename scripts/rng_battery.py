"""p-values for the raw statistics of scripts/rng_battery.cu (the generator battery of DESIGN.md §10).

    python scripts/rng_battery.py [--scale S] [--seed N] [--rounds 3 4 5 6 7 8 10] [--out FILE]

Builds `variants/rng_battery` with nvcc when it is missing, runs it on cuda:0 and prints one row per (rounds, test) with the
two-sided-safe p-value (the upper tail of the chi-square statistic; `FAIL` below 1e-6, `weak` below 1e-3).  Ten rounds is the
published generator and serves as the control row: a test that fails THERE is a broken test, not a broken generator.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys

import numpy as np
from scipy import stats

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "variants", "rng_battery")
SRC = os.path.join(ROOT, "scripts", "rng_battery.cu")


def build(force: bool = False) -> str:
    if force or not os.path.exists(BIN) or os.path.getmtime(BIN) < os.path.getmtime(SRC):
        os.makedirs(os.path.dirname(BIN), exist_ok=True)
        subprocess.check_call(["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-o", BIN, SRC])
    return BIN


def _pooled_chi2(obs, expected, min_expected: float = 10.0):
    """chi-square of observed counts against expected counts, neighbouring cells pooled from both ends until every
    cell expects at least `min_expected`."""
    obs, expected = list(map(float, obs)), list(map(float, expected))
    while len(expected) > 2 and expected[0] < min_expected:
        e0, o0 = expected.pop(0), obs.pop(0)
        expected[0] += e0
        obs[0] += o0
    while len(expected) > 2 and expected[-1] < min_expected:
        e0, o0 = expected.pop(), obs.pop()
        expected[-1] += e0
        obs[-1] += o0
    o, e = np.array(obs), np.array(expected)
    return float(((o - e) ** 2 / e).sum()), len(e) - 1


def p_values(rec: dict) -> dict:
    """{test name: (statistic, dof, p)} for one JSON line of the battery."""
    out = {}
    for name, t in rec.items():
        if name == "rounds":
            continue
        if name == "avalanche":
            out["avalanche_chi2"] = (t["chi2"], t["dof"], float(stats.chi2.sf(t["chi2"], t["dof"])))
            p_one = 2.0 * stats.norm.sf(t["zmax"])
            p = -math.expm1(t["cells"] * math.log1p(-p_one)) if p_one < 1.0 else 1.0   # the largest of `cells` |z| values
            out["avalanche_zmax"] = (t["zmax"], t["cells"], float(p))
        elif name.startswith("hamming"):
            n = t["n"]
            e = n * stats.binom.pmf(np.arange(65), 64, 0.5)
            c, dof = _pooled_chi2(t["hist"], e)
            out[name] = (c, dof, float(stats.chi2.sf(c, dof)))
        elif name == "runs_up":
            h = np.array(t["hist"], dtype=np.float64)
            k = np.arange(1, 8)
            p = np.array([1.0 / math.factorial(i) - 1.0 / math.factorial(i + 1) for i in k] + [1.0 / math.factorial(8)])
            c, dof = _pooled_chi2(h, h.sum() * p)
            out[name] = (c, dof, float(stats.chi2.sf(c, dof)))
        elif name.startswith("birthday"):
            h = np.array(t["hist"], dtype=np.float64)
            e = h.sum() * np.concatenate([stats.poisson.pmf(np.arange(31), t["lambda"]), [stats.poisson.sf(30, t["lambda"])]])
            c, dof = _pooled_chi2(h, e)
            out[name] = (c, dof, float(stats.chi2.sf(c, dof)))
        else:
            out[name] = (t["chi2"], t["dof"], float(stats.chi2.sf(t["chi2"], t["dof"])))
    return out


def verdict(p: float) -> str:
    return "FAIL" if p < 1e-6 or p > 1 - 1e-6 else "weak" if p < 1e-3 or p > 1 - 1e-3 else "ok"


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=int, default=4)
    ap.add_argument("--seed", type=int, default=20250404)
    ap.add_argument("--rounds", type=int, nargs="*", default=[3, 4, 5, 6, 7, 8, 10])
    ap.add_argument("--out", default=None, help="also write {rounds: {test: [stat, dof, p]}} as JSON")
    args = ap.parse_args()
    exe = build()
    proc = subprocess.run([exe, str(args.scale), str(args.seed)] + [str(r) for r in args.rounds], capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stderr)
        return proc.returncode
    table = {}
    for line in proc.stdout.splitlines():
        rec = json.loads(line)
        table[rec["rounds"]] = p_values(rec)
    tests = list(next(iter(table.values())))
    print(f"seed {args.seed}, scale {args.scale}; p-values (upper tail), FAIL: p < 1e-6 or > 1 - 1e-6")
    print("test".ljust(24) + "".join(f"R={r}".rjust(12) for r in table))
    for tname in tests:
        print(tname.ljust(24) + "".join((f"{table[r][tname][2]:.3g}" + ("!" if verdict(table[r][tname][2]) == "FAIL" else
                                                                         "?" if verdict(table[r][tname][2]) == "weak" else "")).rjust(12)
                                        for r in table))
    print("failed tests".ljust(24) + "".join(str(sum(verdict(v[2]) == "FAIL" for v in table[r].values())).rjust(12) for r in table))
    if args.out:
        with open(args.out, "w") as f:
            json.dump({str(r): {k: list(v) for k, v in t.items()} for r, t in table.items()}, f, indent=1)
    return 0


if __name__ == "__main__":
    sys.exit(main())
