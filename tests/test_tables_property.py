"""Property tests (hypothesis) for the host table builder: whatever rates a caller supplies, every alias row is a
valid encoding of the law `outcome_cdf_batch` states, and the batched builder equals the row-at-a-time one."""
import numpy as np
from hypothesis import given, settings, strategies as st

from mlb_odds_engine_b200 import synth, tables

rate = st.floats(min_value=0.0, max_value=0.6, allow_nan=False)


@settings(max_examples=25, deadline=None)
@given(k=st.lists(rate, min_size=4, max_size=4), bb=st.lists(rate, min_size=4, max_size=4),
       hr=st.floats(min_value=0.0, max_value=0.2), L=st.integers(min_value=1, max_value=9), noise=st.booleans(),
       pens=st.booleans())
def test_any_matchup_builds_valid_rows(k, bb, hr, L, noise, pens):
    m = synth.make_matchup("2025-04-10-LAD@SFG", bullpens=pens)
    m["home_lineup"], m["away_lineup"] = m["home_lineup"][:L], m["away_lineup"][:max(1, 9 - L + 1)]
    m["home_pitcher"].update(k_rate=k[0], bb_rate=bb[0])
    m["away_pitcher"].update(k_rate=k[1], bb_rate=bb[1])
    m["away_pitcher"]["hr_pa"] = {"hr_pa_projected": hr}
    m["home_lineup"][0].update(k_rate=k[2], bb_rate=bb[2])
    m["away_lineup"][0].update(k_rate=k[3], bb_rate=bb[3])
    M = tables.as_matchup(m, use_noise=noise)
    t = tables.build_tables([M])[0]
    assert t["lineup_len"].tolist() == [len(m["away_lineup"]), len(m["home_lineup"])]
    S, P, T, B = tables._table_rows(M)
    c = tables.outcome_cdf_batch(M, S, P, T, B, np.zeros(len(S)))
    assert (np.diff(c, axis=1) >= 0).all() and c.min() >= 0.0 and c.max() <= 1.0
    want = np.diff(np.concatenate([np.zeros((len(S), 1)), c, np.ones((len(S), 1))], axis=1), axis=1)
    rows = t["alias"][S, P, T, B]
    assert ((rows >> 3) < (1 << 29)).all()
    for i in (0, len(S) // 2, len(S) - 1):
        got = tables.alias_probabilities(rows[i])
        assert np.abs(got - want[i]).max() < 2.0**-29
        var = tables.alias_variant_probabilities(rows[i])
        assert abs(var.sum() - 1.0) < 1e-12 and (var >= 0).all()
        assert np.array_equal(rows[i], tables.alias_row(tables.outcome_variants(want[i])))
    # fatigue rows exist exactly for staffs without a bullpen
    assert bool(t["flags"] & 1) == (not pens)
    if not pens:
        f = t["fthr"][0, 0, 0, :6].astype(np.int64)
        assert (np.diff(f) >= 0).all() and f.max() <= 2**31
