"""Shared test helpers: golden fixture loading and comparisons (test infrastructure)."""
from __future__ import annotations

import glob
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

REPLAY_SCENARIOS = sorted(os.path.basename(p)[len("replay_"):-4] for p in glob.glob(os.path.join(GOLDEN, "replay_*.npz")))
STATS_SCENARIOS = sorted(os.path.basename(p)[len("stats_"):-4] for p in glob.glob(os.path.join(GOLDEN, "stats_*.npz")))


def load_replay(name):
    z = np.load(os.path.join(GOLDEN, f"replay_{name}.npz"), allow_pickle=False)
    d = {k: z[k] for k in z.files}
    d["matchup"] = json.loads(str(d["matchup"]))
    d["use_noise"] = bool(d["use_noise"])
    offsets = np.zeros(len(d["tape_len"]) + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum(d["tape_len"])
    d["offsets"] = offsets
    d["tape"] = np.ascontiguousarray(d["tape"], dtype=np.float64)
    return d


def load_stats(name):
    z = np.load(os.path.join(GOLDEN, f"stats_{name}.npz"), allow_pickle=False)
    d = {k: z[k] for k in z.files}
    d["matchup"] = json.loads(str(d["matchup"]))
    d["use_noise"] = bool(d["use_noise"])
    d["n"] = int(d["n"])
    return d


def assert_replay_matches_golden(out, g):
    """`out`: REPLAY_GAME_DTYPE array from an implementation; `g`: golden dict from the reference."""
    n = len(g["home_score"])
    assert len(out) == n
    assert (out["status"] == 0).all(), f"status flags set: {np.unique(out['status'])}"
    np.testing.assert_array_equal(out["home_score"], g["home_score"])
    np.testing.assert_array_equal(out["away_score"], g["away_score"])
    np.testing.assert_array_equal(out["n_innings"], g["n_innings"])
    np.testing.assert_array_equal(out["tape_used"], g["tape_len"])
    for i in range(n):
        k = int(g["n_innings"][i])
        np.testing.assert_array_equal(out["away_runs"][i, :k], g["away_runs"][i, :k])
        np.testing.assert_array_equal(out["home_runs"][i, :k], g["home_runs"][i, :k])
        uh = g["used_home"][i][g["used_home"][i] >= 0]
        ua = g["used_away"][i][g["used_away"][i] >= 0]
        assert out["n_used"][i, 0] == len(uh) and out["n_used"][i, 1] == len(ua)
        np.testing.assert_array_equal(out["used_home"][i, : len(uh)], uh)
        np.testing.assert_array_equal(out["used_away"][i, : len(ua)], ua)
    # per-game PA_RECAP_LOG deltas [AWAY, HOME][7] and recap (both sides summed)
    np.testing.assert_array_equal(out["pa"][:, :, :7].astype(np.int64), g["pa_recap"])
    np.testing.assert_array_equal(out["pa"][:, :, :7].sum(axis=1).astype(np.int64), g["recap"])


def chi2_two_sample(c1, c2, min_expected=5.0):
    """Two-sample chi-square on count vectors, pooling sparse cells.  Returns (chi2, dof, p)."""
    from scipy import stats

    c1 = np.asarray(c1, dtype=np.float64).ravel()
    c2 = np.asarray(c2, dtype=np.float64).ravel()
    n1, n2 = c1.sum(), c2.sum()
    tot = c1 + c2
    exp_small = tot * min(n1, n2) / (n1 + n2)
    order = np.argsort(-exp_small)
    c1, c2, exp_small = c1[order], c2[order], exp_small[order]
    keep = exp_small >= min_expected
    a = np.concatenate([c1[keep], [c1[~keep].sum()]])
    b = np.concatenate([c2[keep], [c2[~keep].sum()]])
    if a[-1] + b[-1] == 0:
        a, b = a[:-1], b[:-1]
    k1, k2 = np.sqrt(n2 / n1), np.sqrt(n1 / n2)
    chi2 = float(np.sum((k1 * a - k2 * b) ** 2 / (a + b)))
    dof = len(a) - 1
    return chi2, dof, float(stats.chi2.sf(chi2, dof))


def z_two_proportions(k1, n1, k2, n2):
    p = (k1 + k2) / (n1 + n2)
    se = np.sqrt(max(p * (1 - p), 1e-300) * (1.0 / n1 + 1.0 / n2))
    return (k1 / n1 - k2 / n2) / se


def win_over_stats(joint_full):
    """From a [away][home] full-game count matrix: dict of event counts used by the 4-sigma gate."""
    j = np.asarray(joint_full, dtype=np.float64)
    a = np.arange(j.shape[0])[:, None]
    h = np.arange(j.shape[1])[None, :]
    ev = {"home_win": (h > a), "home_minus_1.5": (h - a > 1.5)}
    for line in (6.5, 7.5, 8.5, 9.5, 10.5, 11.5, 12.5):
        ev[f"over_{line}"] = (a + h > line)
    return {k: float(j[m].sum()) for k, m in ev.items()}
