"""CPU: the C oracle's replay engine against tapes recorded from the unmodified reference."""
import numpy as np
import pytest

import oracle as orc
from helpers import REPLAY_SCENARIOS, assert_replay_matches_golden, load_replay
from mlb_odds_engine_b200 import tables


def test_golden_fixtures_present():
    assert len(REPLAY_SCENARIOS) >= 8


@pytest.mark.parametrize("name", REPLAY_SCENARIOS)
def test_oracle_replay_bit_exact(name):
    g = load_replay(name)
    params = tables.build_params(g["matchup"], use_noise=g["use_noise"])
    out = orc.replay(params, g["tape"], g["offsets"])
    assert_replay_matches_golden(out, g)


def test_oracle_replay_flags_short_tape():
    g = load_replay(REPLAY_SCENARIOS[0])
    params = tables.build_params(g["matchup"], use_noise=g["use_noise"])
    off = g["offsets"][:2].copy()
    off[1] -= 5
    out = orc.replay(params, g["tape"], off)
    assert out["status"][0] & 4


def test_philox_known_answers():
    # Random123 known-answer vectors for philox4x32-10
    kat = [
        ([0, 0, 0, 0], [0, 0], [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]),
        ([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2, [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]),
        ([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0],
         [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]),
    ]
    for ctr, key, want in kat:
        got = orc.philox4x32_10(ctr, key)
        assert [int(x) for x in got] == want


def test_philox2x32_known_answers():
    # Random123 known-answer vectors for philox2x32-10
    kat = [
        ([0, 0], 0, [0xFF1DAE59, 0x6CD10DF2]),
        ([0xFFFFFFFF, 0xFFFFFFFF], 0xFFFFFFFF, [0x2C3F628B, 0xAB4FD7AD]),
        ([0x243F6A88, 0x85A308D3], 0x13198A2E, [0xDD7CE038, 0xF62A4C12]),
    ]
    for ctr, key, want in kat:
        got = orc.philox2x32_10(ctr, key)
        assert [int(x) for x in got] == want
