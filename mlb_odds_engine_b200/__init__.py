"""mlb_odds_engine_b200 — B200-native Monte Carlo game sampler behind the call surface of
BLON333/MLB_Odds_Engine's hot path (core/game_simulator.py -> half_inning_simulator.py ->
pa_simulator.py + bip_resolution, fatigue_modeling, bullpen_utils.simulate_reliever_chain).

Host Python builds per-matchup tables (`tables`), the compute is hand-written sm_100a CUDA behind
the C ABI of include/mlbsim.h (`_lib`).  There is no CPU fallback: every compute entry point
raises if the CUDA library is missing or no B200-class device is present.
"""
__version__ = "0.1.0"
