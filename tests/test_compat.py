"""CPU: the prefetch cache of compat.GpuSimulateGame is keyed by argument CONTENT (id()-keyed caches serve
the previous game's results when an address is recycled or a dict is edited in place)."""
import copy
import timeit

from mlb_odds_engine_b200 import synth
from mlb_odds_engine_b200.compat import argument_fingerprint


def _kw(m, use_noise=True):
    return dict(home_lineup=m["home_lineup"], away_lineup=m["away_lineup"], home_pitcher=m["home_pitcher"],
                away_pitcher=m["away_pitcher"], env=m["env"], home_bullpen=m["home_bullpen"], away_bullpen=m["away_bullpen"],
                use_noise=use_noise)


def test_fingerprint_follows_content_not_identity():
    m = synth.make_matchup("2025-04-04-TEX@HOU")
    fp = argument_fingerprint(_kw(m))
    assert argument_fingerprint(_kw(copy.deepcopy(m))) == fp            # a copy at other addresses: same games
    assert argument_fingerprint(_kw(m, use_noise=False)) != fp
    m["home_lineup"][3]["k_rate"] += 0.01                                # edited in place: ids unchanged
    assert argument_fingerprint(_kw(m)) != fp
    fp = argument_fingerprint(_kw(m))
    m["away_bullpen"][2]["hr_pa"]["hr_pa_projected"] *= 1.1
    assert argument_fingerprint(_kw(m)) != fp
    fp = argument_fingerprint(_kw(m))
    m["env"]["umpire"] = {"k_mod": 1.03, "bb_mod": 1.0}
    assert argument_fingerprint(_kw(m)) != fp
    other = synth.make_matchup("2025-04-04-NYY@BOS")
    assert argument_fingerprint(_kw(other)) != argument_fingerprint(_kw(synth.make_matchup("2025-04-04-TEX@HOU")))
    no_bp = dict(_kw(m), home_bullpen=None)
    assert argument_fingerprint(no_bp) == argument_fingerprint(dict(no_bp, home_bullpen=[]))   # both: no bullpen


def test_fingerprint_is_cheap_enough_for_one_call_per_sim():
    kw = _kw(synth.make_matchup("2025-04-04-TEX@HOU"))
    per_call = min(timeit.repeat(lambda: argument_fingerprint(kw), number=200, repeat=3)) / 200
    assert per_call < 2e-4   # 10 000 sims: well under the cost of materialising their result dicts
