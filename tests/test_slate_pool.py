"""CPU: slate.run_slate with worker processes (pricing + JSON off the critical path) writes byte-identical files and
returns identical documents to the serial path; `--safe` semantics hold in both.  The simulator is replaced by the oracle
twin (no GPU here) — only the host-side pipeline is under test."""
import json
import os

import oracle as orc
from mlb_odds_engine_b200 import slate, synth, tables
from mlb_odds_engine_b200.simulator import RunHistogram


class TwinSim:
    """Stands in for GpuSimulator: same interface, histograms from the CPU twin."""

    def load_matchups(self, ms):
        self.ms = ms
        self.tb = tables.build_tables(ms)

    def simulate(self, n, seed=0):
        return [RunHistogram(orc.native(self.tb[i], i, 0, n, seed), self.ms[i]) for i in range(len(self.ms))]

    def close(self):
        pass


def test_pooled_slate_equals_serial_slate(tmp_path):
    games = synth.make_slate("2025-04-04", 6)
    games[2]["game_id"] += "-T1907"
    a = slate.run_slate(games, n_sims=5_000, seed=3, sim=TwinSim(), out_dir=str(tmp_path / "a"), snapshot_path=str(tmp_path / "a" / "snap.json"))
    b = slate.run_slate(games, n_sims=5_000, seed=3, sim=TwinSim(), out_dir=str(tmp_path / "b"), snapshot_path=str(tmp_path / "b" / "snap.json"),
                        workers=2)
    assert list(a) == list(b) == [g["game_id"] for g in games]
    for g in games:
        gid = g["game_id"]
        assert json.dumps(a[gid], sort_keys=True) == json.dumps(b[gid], sort_keys=True)
        fa = open(tmp_path / "a" / "2025-04-04" / f"{gid}.json").read()
        assert fa == open(tmp_path / "b" / "2025-04-04" / f"{gid}.json").read()
        assert json.loads(fa)["start_time_iso"] == a[gid]["start_time_iso"] is not None
    assert a[games[2]["game_id"]]["start_time_iso"].startswith("2025-04-04T19:07:00")
    # the snapshot side file is the LAST game's, as the serial reference loop leaves it (run_distribution_simulator.py:853-859)
    snap = json.load(open(tmp_path / "a" / "snap.json"))
    assert snap == json.load(open(tmp_path / "b" / "snap.json"))
    last = a[games[-1]["game_id"]]
    assert snap == {f"{e['market']}:{e['side']}": {"fair_odds": e["fair_odds"], "ev_percent": None} for e in last["markets"]}


def test_safe_mode_reports_bad_games_instead_of_raising(tmp_path):
    import pytest

    games = synth.make_slate("2025-04-04", 3)
    games[1]["game_id"] = "no-teams-here"             # pricing cannot name the teams
    del games[2]["home_pitcher"]["k_rate"]            # simulate_game would raise KeyError on the first PA
    with pytest.raises(Exception):
        slate.run_slate(games, n_sims=2_000, sim=TwinSim())
    for workers in (0, 2):
        out = slate.run_slate(games, n_sims=2_000, sim=TwinSim(), out_dir=str(tmp_path / f"w{workers}"), safe=True, workers=workers)
        assert set(out) == {games[0]["game_id"], "failed"}
        assert set(out["failed"]) == {"no-teams-here", games[2]["game_id"]}
        assert os.path.exists(tmp_path / f"w{workers}" / "2025-04-04" / f"{games[0]['game_id']}.json")
