"""CPU: the per-plate-appearance law the tables encode (`tables.outcome_probabilities`: Beta noise marginalised, clip
term integrated, TTO tiers, pitch-count fatigue) against outcome counts of the UNMODIFIED reference's own
`apply_fatigue_modifiers` + `simulate_pa` (core/fatigue_modeling.py:7-54, core/pa_simulator.py:91-143): 400 000 draws
per probe point, tests/golden/pa_outcome_counts.json from oracle/make_pa_fixture.py.  The game-level fixtures cannot
see a per-PA bias below ~1e-3; this gate is |z| < 4.5 per outcome at sigma ~ 7e-4 plus a chi-square per point."""
import json
import os

import numpy as np
import pytest
from scipy import stats

from mlb_odds_engine_b200 import tables

FIXTURE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pa_outcome_counts.json")
POINTS = json.load(open(FIXTURE))["points"]


@pytest.mark.parametrize("pt", POINTS, ids=[f"{p['label']}-{'noise' if p['use_noise'] else 'nonoise'}" for p in POINTS])
def test_outcome_law_matches_reference_simulate_pa(pt):
    m = tables.as_matchup(pt["matchup"], pt["use_noise"])
    tier = min(pt["tto_count"], 4) - 1
    # side 0 = away bats against the home staff; pitcher 0 = the starter
    p = tables.outcome_probabilities(m, 0, 0, tier, pt["slot"], pitch_count=pt["pitch_count"], use_noise=pt["use_noise"])
    counts = np.array(pt["counts"], dtype=np.float64)
    n = counts.sum()
    assert n == pt["n"] and abs(p.sum() - 1.0) < 1e-12
    z = (counts - n * p) / np.sqrt(np.maximum(n * p * (1 - p), 1e-12))
    assert np.abs(z).max() < 4.5, dict(zip(("K", "BB", "1B", "2B", "3B", "HR", "OUT"), np.round(z, 2)))
    keep = n * p >= 5
    chi2 = float((((counts - n * p) ** 2) / (n * p))[keep].sum())
    assert stats.chi2.sf(chi2, keep.sum() - 1) > 1e-4, chi2


def test_probe_points_cover_tiers_noise_extremes_and_fatigue():
    labels = {(p["label"], p["use_noise"]) for p in POINTS}
    for noise in (True, False):
        for t in (1, 2, 3, 4):
            assert (f"typical_tto{t}", noise) in labels
        assert ("extreme_k_bb", noise) in labels and ("fatigued_pc100", noise) in labels
    assert any(p["pitch_count"] > 75 and p["tto_count"] > 4 for p in POINTS)
    # the extreme point is where the clip term of the noise integral is visible (k + 1.01 bb > 0.8)
    ext = next(p for p in POINTS if p["label"] == "extreme_k_bb" and p["use_noise"])
    assert ext["counts"][0] / ext["n"] > 0.5 and ext["counts"][1] / ext["n"] > 0.3
