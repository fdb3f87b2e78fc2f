#!/usr/bin/env python
"""Warp-schedule model of the round structure both game kernels share (no GPU needed).

    python scripts/schedule_model.py > profiles/r02_schedule_model.log

A warp of 32 lanes plays one game per lane.  A half inning lasts 3 + NegBin(3, p_out) plate appearances, a game 17 / 18
halves (+ extras).  The kernels run ROUNDS: U plate-appearance steps in which only the lanes whose half is still running
are active, then one boundary pass for the lanes whose half ended (refill of finished games at the next round's start).
The model replays that schedule and reports lanes per step, plate appearances per round and — with per-region
instruction counts taken from the ncu source pages (profiles/r02_v19_native_sass_hot.txt: PA step 43, the rest of a round
156 executed warp instructions; profiles/r02_v26_replay_sass.txt: step 260, boundary + refill 226) — warp instructions
per plate appearance of a lane, the quantity the issue-bound kernels' throughput is inversely proportional to.

It is calibrated on what was MEASURED: lanes per step of the native kernel's five steps (ncu: 31.0 / 28.3 / 27.0 / 17.9 /
10.2, of which ~1 lane is a finished game waiting for its refill) and the measured ranking of 3 / 4 / 5 / 6 steps per
round (2.90 / 3.05 / 3.085 / 3.00 x 10^9 games/s).  It is then used for the schedules that were NOT built, to say why:
  * two games per lane with a switch at any step,
  * the common-case boundary inline in the step, rare cases parked for a round-end pass (native and replay).
A model is not a measurement: the numbers below are the reason these variants were not given GPU time, nothing more.
(The replay kernel's inline boundary WAS built afterwards — REPLAY_INLINE_BOUNDARY, profiles/r02_rp18.log: the block came out at
~96 instructions per step and measured +0.9 / +1.0 % at 3 / 4 steps per round, where the "+80" rows below say +1 %.)
"""
import numpy as np

rng = np.random.default_rng(1)
P_OUT = 0.685          # P(a plate appearance is an out): 4.35 plate appearances per half inning, as measured
W, L, ROUNDS = 2000, 32, 400


def half_len(shape):
    return 3 + rng.negative_binomial(3, P_OUT, size=shape)


def game_halves(shape):
    h = np.where(rng.random(shape) < 0.47, 17, 18)          # bottom of the 9th skipped when the home team leads
    extras = rng.random(shape) < 0.09
    return h + 2 * extras * (1 + (rng.random(shape) < 0.45))


def one_game(U):
    """The kernels' schedule: U steps, then a boundary pass for every lane whose half ended."""
    r, hl = half_len((W, L)), game_halves((W, L))
    act, pa = np.zeros(U), 0
    for _ in range(ROUNDS):
        for s in range(U):
            live = r > 0
            act[s] += live.sum(); pa += live.sum(); r = r - live
        over = r == 0
        hl = hl - over
        fin = over & (hl == 0)
        r = np.where(over, half_len((W, L)), r)
        hl = np.where(fin, game_halves((W, L)), hl)
    n = W * ROUNDS
    return act / n, pa / n


def two_games(U, boundaries_per_lane):
    """Two games per lane; a lane whose active half ends switches to its other game at the next step if that one can
    play.  boundaries_per_lane: 2 = the round-end pass handles both games of a lane (a second, thinly occupied pass),
    1 = one game per lane per round, the other waits a round."""
    r, hl = half_len((W, L, 2)), game_halves((W, L, 2))
    a = np.zeros((W, L), dtype=int)
    act, pa, second = np.zeros(U), 0, 0
    I, J = np.arange(W)[:, None], np.arange(L)[None, :]
    for _ in range(ROUNDS):
        for s in range(U):
            switch = (r[I, J, a] == 0) & (r[I, J, 1 - a] > 0)
            a = np.where(switch, 1 - a, a)
            ra = r[I, J, a]
            live = ra > 0
            act[s] += live.sum(); pa += live.sum()
            r[I, J, a] = ra - live
        over = r == 0
        if boundaries_per_lane == 1:
            keep = np.zeros_like(over)
            keep[I, J, a] = over[:, :, 0] & over[:, :, 1]
            over = over & ~keep
        second += (over.sum(axis=2) > 1).sum()
        hl = hl - over
        fin = over & (hl == 0)
        r = np.where(over, half_len((W, L, 2)), r)
        hl = np.where(fin, game_halves((W, L, 2)), hl)
    n = W * ROUNDS
    return act / n, pa / n, second / n


def inline_boundary(U, p_slow):
    """The common boundary (no game over, no pitcher change) inline in the step; a lane with a rare boundary parks until
    the round-end pass.  p_slow: share of boundaries that are rare (game over 1/18 + pitcher changes + ...)."""
    r, hl = half_len((W, L)), game_halves((W, L))
    parked = np.zeros((W, L), bool)
    act, pa = np.zeros(U), 0
    for _ in range(ROUNDS):
        for s in range(U):
            live = ~parked
            act[s] += live.sum(); pa += live.sum(); r = r - live
            over = live & (r == 0)
            hl = hl - over
            slow = (over & (hl == 0)) | (over & (rng.random((W, L)) < p_slow - 1 / 18))
            parked |= slow
            r = np.where(over & ~slow, half_len((W, L)), r)
        fin = parked & (hl == 0)
        r = np.where(parked, half_len((W, L)), r)
        hl = np.where(fin, game_halves((W, L)), hl)
        parked[:] = False
    n = W * ROUNDS
    return act / n, pa / n


def main():
    print("# native kernel: PA step 43 warp instructions, rest of a round (refill, wrap fix-up, boundary pass, loop) 156")
    STEP, REST = 43, 156
    base = None
    for U, measured in ((3, 2.90), (4, 3.05), (5, 3.085), (6, 3.00)):
        act, pa = one_game(U)
        cost = (U * STEP + REST) / pa
        if U == 5:
            base = cost
        print(f"one game per lane, {U} steps: lanes/step {act.round(1)}  PA/round {pa:6.1f}  warp instr per lane-PA {cost:5.2f}"
              f"   (measured {measured} x 10^9 games/s)")
    print("  -> model ranking 5 > 6 > 4 > 3 steps as measured (it overstates the loss of short rounds: their boundary pass is thinner)")
    print("# two games per lane, switch at any step: +8 instructions per step (128-bit save / restore of the hot words through")
    print("#   shared memory, predicated), +7 wrap fix-up for the second game, boundary pass on shared-memory state +16")
    for U in (4, 5, 6):
        for b in (2, 1):
            act, pa, second = two_games(U, b)
            rest = REST + 7 + 16 + (60 if b == 2 else 0)      # b == 2: a second boundary pass at `second` lanes per round
            cost = (U * (STEP + 8) + rest) / pa
            print(f"two games per lane, {U} steps, {b} boundary pass(es): lanes/step {act.round(1)}  PA/round {pa:6.1f}  lanes in 2nd pass {second:4.1f}"
                  f"  warp instr per lane-PA {cost:5.2f}  ({100 * (base / cost - 1):+5.1f} % vs today)")
    print("# boundary inline in the step (+29: runs, score, cap commit, tests, side swap) + first-PA capture at any step (+2);")
    print("#   rare boundaries (game over, pitcher change, misc run: ~1/3) parked for a round-end pass of 54")
    for U in (4, 5, 6, 8):
        act, pa = inline_boundary(U, 1 / 3)
        cost = (U * (STEP + 31) + 54 + 26 + 14 + 10) / pa
        print(f"inline boundary, {U} steps: lanes/step {act.round(1)}  PA/round {pa:6.1f}  warp instr per lane-PA {cost:5.2f}  ({100 * (base / cost - 1):+5.1f} % vs today)")

    print()
    print("# replay kernel: PA step 260 warp instructions, boundary pass 165 (with its rare tails), refill + loop 61")
    STEP, BND, REST = 260, 165, 61
    act, pa = one_game(3)
    base = (3 * STEP + BND + REST) / pa
    print(f"one game per lane, 3 steps (the product): lanes/step {act.round(1)}  PA/round {pa:6.1f}  warp instr per lane-PA {base:5.2f}")
    for U in (3, 4, 5, 6):
        for fast in (60, 80):
            act, pa = inline_boundary(U, 3 / 18)
            cost = (U * (STEP + fast) + 110 + REST) / pa
            print(f"inline common boundary (+{fast} per step), rare ones (game over, pitcher change) parked, {U} steps: lanes/step {act.round(1)}"
                  f"  PA/round {pa:6.1f}  warp instr per lane-PA {cost:5.2f}  ({100 * (base / cost - 1):+5.1f} % vs today)")


if __name__ == "__main__":
    main()
