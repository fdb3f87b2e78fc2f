"""CPU: host-side table builders against the reference's documented behaviour."""
import copy

import numpy as np
import pytest

import oracle as orc
from helpers import load_replay
from mlb_odds_engine_b200 import _abi, synth, tables


def test_contact_cdfs_are_numpys_own_doubles():
    cdf_bip, cdf_hit, cdf_fb = tables.contact_cdfs()
    assert cdf_bip.tolist() == [0.27999999999999997, 0.6, 0.8999999999999999, 1.0]      # SURVEY §8(a) A6 [probe]
    assert cdf_hit.tolist() == [0.7618102421595871, 0.9801508535132989, 1.0]
    assert cdf_fb.tolist() == [0.9207292905271502, 1.0]


def test_bip_hit_probability_literals():
    f = tables.bip_hit_probability
    assert f("LD", None, None, 50, 50) == 0.55                      # 0.65 * 1.03 clamped to 0.55
    assert f("POP", 88, 8.0, 50, 50) == 0.05                        # floor
    assert f("GB", 88, 8.0, 50, 50) == 0.305
    assert f("GB", 96, 12, 70, 30) == 0.305 * 1.025 * 1.03 * min(1.03 * 1.03, 1.06)
    assert f("FB", 84, 5.0, 30, 65) == 0.10 * 0.97 * 0.97 * (0.97 * 0.96)
    with pytest.raises(TypeError):
        f("GB", "N/A", 12, 50, 50)                                  # un-enriched starter: bip_resolution.py:40


def test_reference_error_behaviour():
    m = synth.make_matchup()
    bad = copy.deepcopy(m); del bad["home_pitcher"]["k_rate"]
    with pytest.raises(KeyError):
        tables.build_table(bad)
    bad = copy.deepcopy(m); bad["away_pitcher"]["hr_pa"]["hr_pa_projected"] = None
    with pytest.raises(TypeError):
        tables.build_table(bad)
    bad = copy.deepcopy(m); bad["home_bullpen"] = bad["home_bullpen"] + bad["away_bullpen"]
    with pytest.raises(ValueError):
        tables.build_table(bad)


def test_inert_inputs_do_not_change_the_table():
    """env multipliers under the driver's key names, ISO/AVG/wOBA, Stuff+ ... are unread by the hot path (A7)."""
    m = synth.make_matchup()
    t0 = tables.build_table(m)
    m2 = copy.deepcopy(m)
    m2["env"].update({"park_hr_mult": 1.3, "weather_hr_mult": 1.2, "adi_mult": 0.9, "single_mult": 1.1})
    m2["env"]["umpire"] = {"K": 1.2, "BB": 0.8}
    for b in m2["home_lineup"]:
        b.update({"iso": 0.3, "avg": 0.33, "woba": 0.44})
    m2["home_pitcher"].update({"stuff_plus": 140, "hr_fb_rate": 0.3, "iso_allowed": 0.3})
    assert tables.build_table(m2).tobytes() == t0.tobytes()
    m3 = copy.deepcopy(m)
    m3["env"]["weather_hr"] = 1.2          # the key the hot path does read (half_inning_simulator.py:186)
    assert tables.build_table(m3).tobytes() != t0.tobytes()


def test_outcome_law_matches_survey_probe():
    """A9: probabilities sum to 1, thresholds are monotone, tiers shift K down / BB up."""
    m = tables.as_matchup(synth.make_matchup())
    for side in range(2):
        for tier in range(4):
            p = tables.outcome_probabilities(m, side, 0, tier, 3)
            assert abs(p.sum() - 1.0) < 1e-12 and (p >= 0).all()
    p1 = tables.outcome_probabilities(m, 0, 0, 0, 0)
    p4 = tables.outcome_probabilities(m, 0, 0, 3, 0)
    assert p4[0] < p1[0] and p4[1] > p1[1]
    k = (m.bat_k[0, 0] + m.pit_k[0, 0]) / 2
    bb = (m.bat_bb[0, 0] + m.pit_bb[0, 0]) / 2
    assert abs(p1[0] - k) < 1e-15 and abs(p1[1] - 1.01 * bb) < 1e-9   # clip term of the Beta noise ~4e-11
    t = tables.build_table(m)
    # every alias row encodes its target law to within the 2^-32 quantisation of the cuts
    for side, pit, tier, slot in ((0, 0, 0, 0), (1, 3, 2, 8), (0, 6, 3, 4), (1, 0, 1, 5)):
        want = tables.outcome_probabilities(m, side, pit, tier, slot)
        got = tables.alias_probabilities(t["alias"][side, pit, tier, slot])
        assert np.abs(got - want).max() < 2.0**-30
        # the single is split 0.2 / 0.8 by "runner on first takes second" (half_inning_simulator.py:105)
        var = tables.alias_variant_probabilities(t["alias"][side, pit, tier, slot])
        assert abs(var[7] - 0.8 * want[2]) < 2.0**-30 and abs(var[2] - 0.2 * want[2]) < 2.0**-30


def test_alias_rows_degenerate_and_uniform():
    for p in ([1, 0, 0, 0, 0, 0, 0], [0, 0, 0, 0, 0, 1, 0], [0, 1, 0, 0, 0, 0, 0]):
        row = tables.alias_row(p)
        assert tables.alias_probabilities(row).tolist() == [float(x) for x in p]   # exactly, no 2^-32 leak
    row = tables.alias_row([1 / 7] * 7)
    assert np.abs(tables.alias_probabilities(row) - 1 / 7).max() < 2.0**-31
    assert (row >> 3).max() < 2**29 and (row[7] >> 3) == 2**29 - 1      # column 7 of a 7-outcome row never keeps itself (complemented cut)
    # eight variants: all columns carry mass; batch builder = row builder
    rng = np.random.default_rng(5)
    P = tables.outcome_variants(rng.dirichlet(np.ones(7), size=500))
    assert np.allclose(P.sum(axis=1), 1.0) and np.allclose(P[:, 7], 4 * P[:, 2])
    rows = tables.alias_rows(P)
    assert np.array_equal(rows, np.stack([tables.alias_row(p) for p in P]))
    assert max(np.abs(tables.alias_variant_probabilities(r) - p).max() for r, p in zip(rows, P)) < 2.0**-30   # <= 7 cuts x 2^-33
    one_single = tables.alias_row(tables.outcome_variants([0, 0, 1, 0, 0, 0, 0]))
    assert tables.alias_probabilities(one_single).tolist() == [0, 0, 1, 0, 0, 0, 0]


def test_noise_clip_quadrature_against_sampling():
    """Appendix A caveat: E[min(1, X + 1.01 Y)] must be integrated when k + 1.01 bb is near 1."""
    rng = np.random.default_rng(0)
    k, bb = 0.5, 0.3
    x = rng.beta(30 * k, 30 * (1 - k), 2_000_000)
    y = rng.beta(30 * bb, 30 * (1 - bb), 2_000_000)
    mc = np.minimum(1.0, x + 1.01 * y).mean()
    pk, pkbb = tables._p_k_or_bb(k, bb, True)
    assert pk == k and abs(pkbb - mc) < 3e-4 and pkbb < k + 1.01 * bb - 0.002
    assert abs(tables._p_k_or_bb(0.22, 0.08, True)[1] - (0.22 + 1.01 * 0.08)) < 1e-9
    assert tables._p_k_or_bb(0.9, 0.3, False) == (0.9, 1.0)
    assert tables._p_k_or_bb(1.2, 0.1, True) == (1.0, 1.0)
    assert tables._p_k_or_bb(0.3, -0.1, True) == (0.3, 0.3)


def test_fatigue_slope_reproduces_exact_law():
    """thr + (pc - 75) * slope tracks the exactly recomputed CDF to < 2^-24 over a starter's range."""
    m = synth.make_matchup(bullpens=False)
    mm = tables.as_matchup(m)
    t = tables.build_table(m)
    assert t["flags"] & _abi.FLAG_FATIGUE
    for pc in (76, 90, 120, 175):
        for tier in (0, 3):
            exact = tables.outcome_cdf(mm, 1, 0, tier, 4, pitch_count=pc)
            approx = (t["fthr"][1, tier, 4, :6].astype(np.int64) + (pc - 75) * t["fslope"][1, tier, 4, :6].astype(np.int64)) / 2.0**31
            assert np.abs(exact - approx).max() < 2.0**-24
    t2 = tables.build_table(synth.make_matchup())
    assert not (t2["flags"] & _abi.FLAG_FATIGUE) and not t2["fslope"].any()


def test_oracle_twin_pa_mix_matches_table_law():
    """The integer twin's outcome frequencies converge to the analytic per-PA law (neutral game)."""
    m = synth.neutral_calibration_matchup()
    mm = tables.as_matchup(m)
    t = tables.build_table(m)
    n = 200_000
    h = orc.native(t, 0, 0, n, 3)
    per_game = orc.native_persim(t, 0, 0, 2000, 3)
    assert int(h["n_games"]) == n
    # first-time-through, fresh starter: every batter identical => tier-1 law for the first 9 PAs
    p = tables.outcome_probabilities(mm, 0, 0, 0, 0)
    assert abs(p[0] - 0.225) < 1e-12
    tot = h["pa"][:, :7].sum()
    mix = h["pa"][:, :7].sum(axis=0) / tot
    published = np.array([22.23, 8.36, 19.58, 4.77, 0.40, 2.52, 42.14]) / 100   # logs/calibration_outcomes.csv
    assert np.abs(mix - published).max() < 0.002
    assert abs(tot / n - 81.38) < 0.2
    assert (per_game["pa"][:, :, :7].sum(axis=(1, 2)) == per_game["tape_used"]).all()


def test_params_roundtrip_fields():
    g = load_replay("odd_onesided")
    p = tables.build_params(g["matchup"], use_noise=True)
    assert p["k_mod"] == 1.04 and p["bb_mod"] == 0.95
    assert p["bullpen_len"].tolist() == [5, 0] and p["lineup_len"].tolist() == [9, 9]
    hr = g["matchup"]["home_pitcher"]["hr_pa"]["hr_pa_projected"] * 1.12
    assert p["pit_hr"][0, 0] == hr


def test_summary_host_statement_on_oracle_histogram():
    """`_abi.summarize_hist` (what mlbsim_summarize must equal) against per-sim records of the same games."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
    import oracle as orc
    from mlb_odds_engine_b200 import _abi, sweep

    t = tables.build_tables([synth.make_matchup()])
    n = 5000
    rec = orc.native(t, 0, 0, n, 3).reshape(-1)[0]
    games = orc.native_persim(t, 0, 0, n, 3)
    s = _abi.summarize_hist(rec)
    a, h = games["away_score"].astype(np.int64), games["home_score"].astype(np.int64)
    assert int(s["n_games"]) == n
    assert int(s["sum_runs"][4][0]) == a.sum() and int(s["sum_runs"][4][1]) == h.sum()
    assert int(s["sum_sq"][0]) == (a * a).sum() and int(s["sum_sq"][1]) == (h * h).sum() and int(s["sum_cross"]) == (a * h).sum()
    assert int(s["home_wins"]) == (h > a).sum() and int(s["away_wins"]) == (a > h).sum()
    assert int(s["sum_innings"]) == games["n_innings"].astype(np.int64).sum()
    for c, cap in enumerate((1, 3, 5, 7)):
        assert int(s["sum_runs"][c][0]) == games["away_runs"][:, :cap].astype(np.int64).sum()
        assert int(s["sum_runs"][c][1]) == games["home_runs"][:, :cap].astype(np.int64).sum()
    assert np.array_equal(s["pa"], rec["pa"])
    row = sweep.summary_row(s)
    assert abs(row["home_runs_sd"] - h.std()) < 1e-9 and abs(row["innings_mean"] - games["n_innings"].mean()) < 1e-12
    assert abs(sum(row["rates"]["HOME"].values()) - 1.0) < 1e-12


def test_matchup_that_cannot_score_is_refused():
    """Every row of both sides a certain out => the game can never end (core/game_simulator.py:110-111 loops while
    tied); the builder raises instead of handing the kernel a game that only its inning guard would stop."""
    from mlb_odds_engine_b200 import synth

    m = synth.make_matchup("2025-04-04-TEX@HOU")
    for side in ("home", "away"):
        for b in m[f"{side}_lineup"]:
            b["k_rate"], b["bb_rate"] = 1.0, 0.0
        for p in [m[f"{side}_pitcher"]] + list(m[f"{side}_bullpen"]):
            p["k_rate"], p["bb_rate"] = 1.0, 0.0
    m["env"] = dict(m["env"], umpire={"k_mod": 1.2, "bb_mod": 1.0})
    with pytest.raises(tables.MatchupCannotScore):
        tables.build_tables([m], use_noise=False)
    assert len(tables.build_tables([m], use_noise=False, validate=False)) == 1
    # one side able to reach base is enough for the game to end
    for b in m["home_lineup"]:
        b["k_rate"] = 0.2
    assert len(tables.build_tables([m], use_noise=False)) == 1


def test_alias_draw_rule_is_exact_on_the_tie():
    """include/mlbsim.h: variant = ((w0 << 3) > e) ? col : (e & 7) with e = (2^29 - 1 - cut) << 3 | alias.  Exactly `cut` of
    a column's 2^29 draws keep the column — checked on the draws either side of the cut — and a column whose own
    outcome has probability 0 never returns it, not even on the draw whose fraction ties with the stored cut (the
    alias bits sit BELOW the cut in the compared word; storing the cut complemented makes that tie fall to the alias)."""
    one = 1 << 29

    def variant(w0, row):
        col = w0 >> 29
        e = int(row[col])
        fast = col if ((w0 << 3) & 0xFFFFFFFF) > e else e & 7
        assert fast == (col if (w0 & (one - 1)) > (e >> 3) else e & 7)      # the two statements of the header agree
        return fast

    rng = np.random.default_rng(3)
    P = tables.outcome_variants(rng.dirichlet(np.full(7, 0.5), size=200))
    P[:50, 0] = 0.0                      # K impossible
    P[:50] /= P[:50].sum(axis=1, keepdims=True)
    for p, row in zip(P, tables.alias_rows(P)):
        for col in range(8):
            cut = one - 1 - (int(row[col]) >> 3)            # draws of this column that keep it
            alias = int(row[col]) & 7
            keep = [variant((col << 29) | f, row) == col for f in (0, 1, one - cut - 1, one - cut, one - 2, one - 1) if 0 <= f < one]
            fr = [f for f in (0, 1, one - cut - 1, one - cut, one - 2, one - 1) if 0 <= f < one]
            for f, k in zip(fr, keep):
                if alias != col:
                    assert k == (f >= one - cut), (col, f, cut)     # the top `cut` fractions keep the column, nothing else
            if p[col] == 0.0:
                assert cut == 0 or alias == col
                assert all(variant((col << 29) | f, row) != col or alias == col for f in (0, 1, 2, one - 1))
        got = tables.alias_variant_probabilities(row)
        assert got[0] == 0.0 if p[0] == 0.0 else True
