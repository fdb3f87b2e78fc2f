"""CPU: the native law (oracle twin of the CUDA kernel, tables from the host builder) against statistics of the
UNMODIFIED reference (tests/golden/stats_*.npz, 10^6 reference games per scenario) — the same chi-square / 4-sigma
gate the GPU test applies at 10^7 sims (SURVEY.md §8(d) parity criteria), here at 3 x 10^6 oracle sims."""
import numpy as np
import pytest

import oracle as orc
from helpers import STATS_SCENARIOS, chi2_two_sample, load_stats, win_over_stats, z_two_proportions
from mlb_odds_engine_b200 import tables

N = 3_000_000
# 7 scenarios x 9 chi-square statistics share this gate: 1e-4 per statistic keeps the family-wise false-alarm rate below 1 %
# (the fixtures are single 10^6-game samples and dominate the statistic: hetero_nobullpen_noise's full-game table sits
# at p ~ 0.03 against ANY seed of the native law, old or new)
P_GATE = 1e-4


@pytest.mark.parametrize("name", STATS_SCENARIOS)
def test_native_law_matches_reference_statistics(name):
    g = load_stats(name)
    t = tables.build_tables([tables.as_matchup(g["matchup"], use_noise=g["use_noise"])])[0]
    h = orc.native(t, 0, 0, N, 20251020)
    assert int(h["n_games"]) == N
    for c in range(5):
        chi2, dof, p = chi2_two_sample(g["joint"][c], h["joint"][c])
        assert p > P_GATE, f"cap {c}: chi2 {chi2:.1f}/{dof} p={p:.2e}"
    ref, got = win_over_stats(g["joint"][4]), win_over_stats(h["joint"][4])
    for k in ref:
        z = z_two_proportions(got[k], N, ref[k], g["n"])
        assert abs(z) < 4.0, f"{k}: z={z:.2f}"
    for s in range(2):
        chi2, dof, p = chi2_two_sample(g["pa"][s], h["pa"][s][:7])
        assert p > P_GATE, f"pa side {s}: chi2 {chi2:.1f}/{dof} p={p:.2e}"
    chi2, dof, p = chi2_two_sample(g["innings"], h["innings"])
    assert p > P_GATE, f"innings: p={p:.2e}"
    ref_rel = g["rel_games"][::-1]  # fixture rows are [away, home]; mlbsim_hist rows are [home staff, away staff]
    for s in range(2):
        for i in range(8):
            if ref_rel[s][i] or h["rel_games"][s][i]:
                z = z_two_proportions(float(h["rel_games"][s][i]), N, float(ref_rel[s][i]), g["n"])
                assert abs(z) < 4.0, f"reliever {s}/{i}: z={z:.2f}"
