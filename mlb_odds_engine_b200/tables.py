"""Host-side builders: reference argument dicts -> device tables (float64 numpy).

Inputs are exactly the keyword arguments of `simulate_game` (core/game_simulator.py:14-25).  Two
products per matchup:

* `build_params`  -> `mlbsim_params`  raw float64 numbers for REPLAY mode.  Every quantity the
  reference computes per PA from dict values is either copied verbatim or precomputed here with
  the same float64 operation order, so the device compares the same doubles
  (core/bip_resolution.py:20-57, core/pa_simulator.py:46,54,68-79).
* `build_table`   -> `mlbsim_table`   integer sampling tables for NATIVE mode: the per-PA law of
  core/pa_simulator.py:91-143 with the Beta(30p, 30(1-p)) noise of :23-32 integrated out (it is
  redrawn independently every PA and only the outcome is consumed), one Walker alias row per
  (side, pitcher, TTO tier of core/fatigue_modeling.py:25-35, batter), and for starters without
  a bullpen 31-bit cumulative thresholds with an integer slope per pitch above 75 (:38-43).

Only keys the reference's hot path reads are used; everything else in the
dicts (iso, avg, woba, stuff_plus, park multipliers, ...) is inert there and therefore here.
"""
from __future__ import annotations

import math

import copy
import os

import numpy as np

from . import _abi

# ---- model literals (reference file:line) ---------------------------------------------------
BIP_TYPES = ("GB", "LD", "FB", "POP")            # pa_simulator.py:54
BIP_WEIGHTS = (0.28, 0.32, 0.30, 0.10)           # pa_simulator.py:54
BASE_BABIP = {"LD": 0.65, "GB": 0.305, "FB": 0.10, "POP": 0.02}  # bip_resolution.py:13-18
TTO_PENALTY = (0.0, 0.015, 0.035, 0.060)         # fatigue_modeling.py:25-32 (tto 1, 2, 3, >=4)
NOISE_WEIGHT = 30                                # pa_simulator.py:23
INFIELD_FALLBACK = 0.10                          # pa_simulator.py:74
FATIGUE_START = 75                               # fatigue_modeling.py:38
SLOPE_SPAN = 100                                 # pitches over which the integer slope is fitted


def _normalised_cdf(p):
    """legacy `np.random.choice(p=...)`: cdf = cumsum(p); cdf /= cdf[-1]."""
    cdf = np.cumsum(np.asarray(p, dtype=np.float64))
    cdf /= cdf[-1]
    return cdf


def contact_cdfs():
    """The three categorical CDFs of resolve_contact (pa_simulator.py:54, :68-72, :75-79)."""
    cdf_bip = _normalised_cdf(BIP_WEIGHTS)
    probs = [0.76, 0.22, 0.02]
    probs[0] *= 1.01
    total = sum(probs)
    cdf_hit = _normalised_cdf([p / total for p in probs])
    fb = [0.92, 0.08]
    fb[0] *= 1.01
    total = sum(fb)
    cdf_fb = _normalised_cdf([p / total for p in fb])
    return cdf_bip, cdf_hit, cdf_fb


# resolve_bip's modifiers as data: (below, factor if below, above, factor if above); inside the band the factor is 1.0
_SPEED_BAND = (40, 0.97, 65, 1.025)        # core/bip_resolution.py:23-26  batter speed
_RANGE_BAND = (40, 1.03, 60, 0.97)         # :29-32  fielder rating (50 = neutral)
_EV_BAND = (85, 0.97, 95, 1.03)            # :40-43  exit velocity, default 88
_LA_BAND = (6.0, 0.96, 10.0, 1.03)         # :45-48  launch angle, default 12
_CONTACT_CAP = 1.06                        # :51
_BABIP_LIMITS = (0.05, 0.55)               # :57


def _band_factor(x, band):
    """1.0 inside [below, above], else the band's factor.  Comparing a non-number raises TypeError exactly where the
    reference does (the "N/A" strings of un-enriched starters, core/bip_resolution.py:40)."""
    below, f_below, above, f_above = band
    return f_above if x > above else f_below if x < below else 1.0


def bip_hit_probability(bip_type, ev, la, batter_speed, fielder_rating):
    """P(hit | ball in play of `bip_type`), core/bip_resolution.py:20-57.  Floating-point products are taken in the
    reference's order — ((base x speed) x range) x min(ev x la, 1.06) — so replay mode compares the same doubles;
    a factor of 1.0 is exact, which is what lets the bands be data."""
    prob = BASE_BABIP[bip_type] * _band_factor(batter_speed, _SPEED_BAND) * _band_factor(fielder_rating, _RANGE_BAND)
    contact = _band_factor(88 if ev is None else ev, _EV_BAND) * _band_factor(12 if la is None else la, _LA_BAND)
    prob *= min(contact, _CONTACT_CAP)
    return max(min(prob, _BABIP_LIMITS[1]), _BABIP_LIMITS[0])


class Matchup:
    """Numbers the hot path reads from one `simulate_game(**kwargs)` argument set.

    Raises what the reference would raise on the same malformed input: KeyError
    for a pitcher without k_rate/bb_rate (fatigue_modeling.py:34-35), TypeError for a non-numeric
    exit velocity / launch angle (bip_resolution.py:40) or hr_pa_projected=None (pa_simulator.py:46).
    """

    def __init__(self, home_lineup, away_lineup, home_pitcher, away_pitcher, env=None,
                 home_bullpen=None, away_bullpen=None, use_noise=True, game_id=None, **_ignored):
        env = env or {}
        self.game_id = game_id
        self.use_noise = bool(use_noise)
        umpire = env.get("umpire", {}) or {}
        # pa_simulator.py:112-114 — only read when the umpire dict is truthy; x * 1.0 is exact
        self.k_mod = float(umpire.get("k_mod", 1.0))
        self.bb_mod = float(umpire.get("bb_mod", 1.0))
        self.weather_hr = env.get("weather_hr", 1.0)  # half_inning_simulator.py:186
        # side 0 = away bats vs the home staff; side 1 = home bats vs the away staff
        self.lineups = [list(away_lineup), list(home_lineup)]
        self.staffs = [[home_pitcher] + list(home_bullpen or []), [away_pitcher] + list(away_bullpen or [])]
        for s in range(2):
            if not 1 <= len(self.lineups[s]) <= _abi.MAX_LINEUP:
                raise ValueError(f"lineup length {len(self.lineups[s])} outside 1..{_abi.MAX_LINEUP}")
            if len(self.staffs[s]) > _abi.MAX_PITCHERS:
                raise ValueError(f"bullpen of {len(self.staffs[s]) - 1} exceeds {_abi.MAX_BULLPEN} relievers")
        self.bullpen_names = [[p.get("name", "Unknown") for p in st[1:]] for st in self.staffs]  # [home, away]

        L, P = _abi.MAX_LINEUP, _abi.MAX_PITCHERS
        self.bat_k = np.zeros((2, L))
        self.bat_bb = np.zeros((2, L))
        self.speed = np.full((2, L), 50.0)
        self.pit_k = np.zeros((2, P))
        self.pit_bb = np.zeros((2, P))
        self.pit_hr = np.zeros((2, P))
        self.hit_prob = np.zeros((2, P, L, _abi.N_BIP))
        for s in range(2):
            for b, bat in enumerate(self.lineups[s]):
                self.bat_k[s, b] = bat.get("k_rate", 0.22)      # pa_simulator.py:108
                self.bat_bb[s, b] = bat.get("bb_rate", 0.08)    # pa_simulator.py:109
                self.speed[s, b] = bat.get("speed", 50)         # pa_simulator.py:62
            for p, pit in enumerate(self.staffs[s]):
                self.pit_k[s, p] = pit["k_rate"]
                self.pit_bb[s, p] = pit["bb_rate"]
                hr = pit.get("hr_pa", {}).get("hr_pa_projected", 0.046) * self.weather_hr  # pa_simulator.py:46
                self.pit_hr[s, p] = hr
                ev = pit.get("exit_velocity_avg")
                la = pit.get("launch_angle_avg")
                fr = pit.get("fielder_rating", 50)
                for b in range(len(self.lineups[s])):
                    for t, name in enumerate(BIP_TYPES):
                        self.hit_prob[s, p, b, t] = bip_hit_probability(name, ev, la, self.speed[s, b], fr)

    @property
    def lineup_len(self):
        return [len(self.lineups[0]), len(self.lineups[1])]

    @property
    def bullpen_len(self):
        return [len(self.staffs[0]) - 1, len(self.staffs[1]) - 1]


def as_matchup(m, use_noise=True) -> Matchup:
    if isinstance(m, Matchup):
        return m
    kw = dict(m)
    kw.setdefault("use_noise", use_noise)
    return Matchup(**kw)


# ---------------------------------------------------------------------------------------------
# replay parameters
# ---------------------------------------------------------------------------------------------
def build_params(m, use_noise=None) -> np.ndarray:
    m = as_matchup(m) if use_noise is None else as_matchup(m, use_noise)
    out = np.zeros((), dtype=_abi.PARAMS_DTYPE)
    out["lineup_len"] = m.lineup_len
    out["bullpen_len"] = m.bullpen_len
    out["use_noise"] = int(m.use_noise if use_noise is None else use_noise)
    out["k_mod"], out["bb_mod"] = m.k_mod, m.bb_mod
    for f in ("bat_k", "bat_bb", "pit_k", "pit_bb", "pit_hr", "hit_prob"):
        out[f] = getattr(m, f)
    out["cdf_bip"], out["cdf_hit"], out["cdf_fb"] = contact_cdfs()
    return out


# ---------------------------------------------------------------------------------------------
# native law
# ---------------------------------------------------------------------------------------------
def _p_k_or_bb(k, bb, use_noise):
    """(P(K), P(K or BB)) for effective rates k, bb (pa_simulator.py:117-124, :35-41)."""
    pk, pkbb = _p_k_or_bb_batch(np.array([k], dtype=np.float64), np.array([bb], dtype=np.float64), use_noise)
    return float(pk[0]), float(pkbb[0])


def _p_k_or_bb_batch(k, bb, use_noise):
    """Vectorised over rows.  With noise the reference draws kp ~ Beta(30k, 30(1-k)) (0.0 / 1.0 without
    a draw when k <= 0 / k >= 1), bp likewise, then bp *= 1.01, and compares one uniform u:
    K iff u < kp, BB iff kp <= u < kp + bp.  Hence P(K) = E[kp] and P(K or BB) = E[min(1, kp + bp)]."""
    k = np.asarray(k, dtype=np.float64)
    bb = np.asarray(bb, dtype=np.float64)
    pk = np.clip(k, 0.0, 1.0)
    if not use_noise:
        return pk, np.maximum(pk, np.clip(k + bb * 1.01, 0.0, 1.0))
    pkbb = np.empty_like(pk)
    no_walk = bb <= 0.0
    sure = (~no_walk) & ((bb >= 1.0) | (k >= 1.0))        # kp + 1.01 >= 1, or kp = 1
    no_k = (~no_walk) & (~sure) & (k <= 0.0)               # only the walk term is random
    both = ~(no_walk | sure | no_k)
    pkbb[no_walk] = pk[no_walk]
    pkbb[sure] = 1.0
    if no_k.any():
        pkbb[no_k] = _expect_clip(None, bb[no_k])
    if both.any():
        pkbb[both] = _expect_clip_cached(k[both], bb[both])
    return pk, pkbb


_CLIP_CACHE: dict = {}


def _expect_clip_cached(k, bb):
    """`_expect_clip` with a (k, bb) -> value memo: slates and sweeps repeat most (batter, pitcher, tier)
    combinations (a sweep over one pitcher leaves the other 13 pitchers' rows untouched).  Distinct pairs of the
    batch are found first, so a repeated pair is integrated once."""
    z = np.empty(len(k), dtype=np.complex128)   # (k, bb) as one sortable scalar: both doubles are kept exactly
    z.real, z.imag = k, bb
    pairs, inverse = np.unique(z, return_inverse=True)
    uk, ub = np.ascontiguousarray(pairs.real), np.ascontiguousarray(pairs.imag)
    vals = np.empty(len(pairs))
    miss = []
    for i, key in enumerate(zip(uk.tolist(), ub.tolist())):
        v = _CLIP_CACHE.get(key)
        if v is None:
            miss.append(i)
        else:
            vals[i] = v
    if miss:
        mi = np.array(miss)
        new = _expect_clip(uk[mi], ub[mi])
        vals[mi] = new
        if len(_CLIP_CACHE) > 2_000_000:
            _CLIP_CACHE.clear()
        for i, v in zip(miss, new.tolist()):
            _CLIP_CACHE[(float(uk[i]), float(ub[i]))] = v
    return vals[inverse.reshape(-1)]


_GL = None
_CLIP_THREADS = max(1, min(16, (os.cpu_count() or 1)))


def _gauss_legendre_01():
    global _GL
    if _GL is None:
        nodes, weights = np.polynomial.legendre.leggauss(48)   # agrees with 400 nodes to 7e-15 (32: 2e-12)
        _GL = (0.5 * (nodes + 1.0), 0.5 * weights)
    return _GL


def _expect_clip(k, bb):
    """`_expect_clip_rows` over the host's cores for large batches (the special-function loops release the GIL)."""
    n = len(np.atleast_1d(bb))
    workers = min(_CLIP_THREADS, n // 128)
    if workers < 2:
        return _expect_clip_rows(k, bb)
    from concurrent.futures import ThreadPoolExecutor

    bb = np.atleast_1d(np.asarray(bb, dtype=np.float64))
    k = None if k is None else np.atleast_1d(np.asarray(k, dtype=np.float64))
    cuts = np.linspace(0, n, workers + 1).astype(int)
    with ThreadPoolExecutor(workers) as pool:
        parts = pool.map(lambda i: _expect_clip_rows(None if k is None else k[cuts[i]:cuts[i + 1]],
                                                     bb[cuts[i]:cuts[i + 1]]), range(workers))
        return np.concatenate(list(parts))


def _expect_clip_rows(k, bb):
    """E[min(1, X + 1.01*Y)] for arrays k, bb: X ~ Beta(30k, 30(1-k)) (X = 0 if k is None),
    Y ~ Beta(30bb, 30(1-bb)).  For realistic rates the clip term is ~1e-11 and the result equals
    k + 1.01*bb to ten digits; for extreme grid points (k = 0.5, bb = 0.3: 0.4 %) it matters
    (SURVEY.md Appendix A caveat).  Gauss-Legendre over y, closed form in x."""
    from scipy import special, stats

    bb = np.atleast_1d(np.asarray(bb, dtype=np.float64))
    y, w = _gauss_legendre_01()
    ay, by = (NOISE_WEIGHT * bb)[:, None], (NOISE_WEIGHT * (1 - bb))[:, None]
    wy = w[None, :] * stats.beta.pdf(y[None, :], ay, by)
    c = (1.0 - 1.01 * y)[None, :]            # X must exceed c for the clip to bite
    if k is None:
        excess = np.maximum(-c, 0.0)
        mean = 1.01 * bb
    else:
        k = np.atleast_1d(np.asarray(k, dtype=np.float64))
        ax, bx = (NOISE_WEIGHT * k)[:, None], (NOISE_WEIGHT * (1 - k))[:, None]
        cc = np.clip(c, 0.0, 1.0)
        # E[(X - c)^+] = k * (1 - I_c(a+1, b)) - c * (1 - I_c(a, b)); for c < 0 it is k - c
        excess = k[:, None] * (1.0 - special.betainc(ax + 1, bx, cc)) - cc * (1.0 - special.betainc(ax, bx, cc))
        excess = np.where(c < 0.0, k[:, None] - c, excess)
        mean = k + 1.01 * bb
    return np.clip(mean - np.sum(wy * excess, axis=1), 0.0, 1.0)


def _effective_rates(m: Matchup, side, pit, tier, slot, pitch_count):
    """Per-PA K and BB rates before noise: fatigue-adjusted pitcher rate averaged with the batter's, times the
    umpire modifiers (core/fatigue_modeling.py:23-43 then core/pa_simulator.py:108-114)."""
    pc = np.asarray(pitch_count, dtype=np.float64)
    pen = np.asarray(TTO_PENALTY)[tier]
    fl = np.maximum(0.0, (pc - FATIGUE_START) / 25)
    ak = m.pit_k[side, pit] * (1 - pen) * (1.0 - 0.02 * fl)
    ab = m.pit_bb[side, pit] * (1 + pen) * (1.0 + 0.03 * fl)
    return (m.bat_k[side, slot] + ak) / 2 * m.k_mod, (m.bat_bb[side, slot] + ab) / 2 * m.bb_mod


def outcome_cdf_batch(m: Matchup, side, pit, tier, slot, pitch_count, use_noise=None):
    """Rows of cumulative P over [K, BB, 1B, 2B, 3B, HR] (OUT is the remainder) for arrays of PA
    contexts — the marginal law of core/pa_simulator.py:91-143 + core/bip_resolution.py +
    core/fatigue_modeling.py.  Returns float64 [N, 6]."""
    use_noise = m.use_noise if use_noise is None else use_noise
    side, pit, tier, slot = (np.asarray(x, dtype=np.int64) for x in (side, pit, tier, slot))
    k, bb = _effective_rates(m, side, pit, tier, slot, pitch_count)
    pk, pkbb = _p_k_or_bb_batch(k, bb, use_noise)
    contact = 1.0 - pkbb
    cdf_bip, cdf_hit, cdf_fb = contact_cdfs()
    w_bip = np.diff(np.concatenate([[0.0], cdf_bip]))
    h = np.diff(np.concatenate([[0.0], cdf_hit]))
    f = np.diff(np.concatenate([[0.0], cdf_fb]))
    hr = np.clip(m.pit_hr[side, pit], 0.0, 1.0)
    hp = m.hit_prob[side, pit, slot]
    ph = hp[..., 0] * w_bip[0] + hp[..., 1] * w_bip[1] + hp[..., 2] * w_bip[2] + hp[..., 3] * w_bip[3]   # fixed order
    s1 = (1 - hr) * (ph * h[0] + (1 - ph) * INFIELD_FALLBACK * f[0])
    s2 = (1 - hr) * (ph * h[1] + (1 - ph) * INFIELD_FALLBACK * f[1])
    s3 = (1 - hr) * ph * h[2]
    c = np.stack([pk, pkbb, pkbb + contact * s1, pkbb + contact * (s1 + s2), pkbb + contact * (s1 + s2 + s3),
                  pkbb + contact * (s1 + s2 + s3 + hr)], axis=1)
    return np.minimum(np.maximum.accumulate(c, axis=1), 1.0)


def outcome_cdf(m: Matchup, side, pit, tier, slot, pitch_count=0, use_noise=None):
    return outcome_cdf_batch(m, [side], [pit], [tier], [slot], [pitch_count], use_noise)[0]


def outcome_probabilities(m: Matchup, side, pit, tier, slot, pitch_count=0, use_noise=None):
    """P over [K, BB, 1B, 2B, 3B, HR, OUT]."""
    c = outcome_cdf(m, side, pit, tier, slot, pitch_count, use_noise)
    return np.diff(np.concatenate([[0.0], c, [1.0]]))


def _quantise(cdf):
    q = np.floor(np.asarray(cdf) * 2147483648.0 + 0.5)
    return np.clip(q, 0, 2147483648).astype(np.uint64)


ALIAS_COLS = 8
ALIAS_FRAC_BITS = 29
VARIANT_1B_ADVANCE = 7      # include/mlbsim.h MLBSIM_V_1B_ADV: single on which a runner on 1B takes second
P_FIRST_TO_SECOND = 0.8     # core/half_inning_simulator.py:105


def outcome_variants(probs) -> np.ndarray:
    """[..., 7] outcome probabilities -> [..., 8] alias-draw variants: the single is split by the independent
    0.8 draw "runner on first takes second" (half_inning_simulator.py:105), so that draw costs the kernel nothing:
    variant 2 = single, runner holds (0.2), variant 7 = single, runner advances (0.8)."""
    probs = np.asarray(probs, dtype=np.float64)
    out = np.empty(probs.shape[:-1] + (8,))
    out[..., :7] = probs
    adv = probs[..., 2] * P_FIRST_TO_SECOND
    out[..., 7] = adv
    out[..., 2] = probs[..., 2] - adv
    return out


def alias_row(p) -> np.ndarray:
    """Walker/Vose alias table for 7 or 8 probabilities `p` on 8 columns (with 7, column 7 carries
    no mass of its own).  Entry = (2^29 - 1 - cut) << 3 | alias with cut in [0, 2^29): a draw w0 picks column
    w0 >> 29 and keeps the column's own outcome iff (w0 << 3) > entry, i.e. for exactly `cut` of the column's 2^29
    draws (the complemented form makes the tie with the alias bits fall to the alias: an outcome of probability 0
    never occurs).  Deterministic."""
    n_own = len(p)
    q = np.zeros(ALIAS_COLS)
    q[:n_own] = np.asarray(p, dtype=np.float64) * ALIAS_COLS / float(np.sum(p))
    cut = np.ones(ALIAS_COLS)
    alias = np.arange(ALIAS_COLS)
    small = [i for i in range(ALIAS_COLS) if q[i] < 1.0]
    large = [i for i in range(ALIAS_COLS) if q[i] >= 1.0]
    while small and large:
        s = small.pop()
        g = large.pop()
        cut[s] = q[s]
        alias[s] = g
        q[g] = q[g] - (1.0 - q[s])
        (small if q[g] < 1.0 else large).append(g)
    for i in small + large:      # leftovers are full columns up to rounding
        cut[i] = 1.0
        alias[i] = i if i < n_own else int(np.argmax(p))
    one = 1 << ALIAS_FRAC_BITS
    out = np.zeros(ALIAS_COLS, dtype=np.uint32)
    for i in range(ALIAS_COLS):
        c = int(math.floor(cut[i] * one + 0.5))     # how many of the column's 2^29 draws keep the column's own outcome
        a = int(alias[i])
        if c >= one:             # full column: encode as "always alias" with alias = own outcome
            c, a = 0, (i if i < n_own else a)
        if i >= n_own:
            c = 0
        c = max(c, 0)
        out[i] = ((one - 1 - c) << 3) | a
    return out


def alias_rows(P) -> np.ndarray:
    """`alias_row` for N rows at once ([N, 7 or 8] probabilities -> [N, 8] uint32): the same stack discipline
    (highest column first) run in lock step over all rows, so the result is identical entry for entry."""
    P = np.ascontiguousarray(P, dtype=np.float64)   # (C order: the row sums below then round like the device builder's)
    N, n_own = P.shape
    q = np.zeros((N, ALIAS_COLS))
    q[:, :n_own] = P * ALIAS_COLS / P.sum(axis=1)[:, None]
    cut = np.ones((N, ALIAS_COLS))
    alias = np.tile(np.arange(ALIAS_COLS), (N, 1))
    small = np.zeros((N, ALIAS_COLS), dtype=np.int64)
    large = np.zeros((N, ALIAS_COLS), dtype=np.int64)
    ns = np.zeros(N, dtype=np.int64)
    nl = np.zeros(N, dtype=np.int64)
    for i in range(ALIAS_COLS):
        lt = q[:, i] < 1.0
        r = np.nonzero(lt)[0]
        small[r, ns[r]] = i
        ns[r] += 1
        r = np.nonzero(~lt)[0]
        large[r, nl[r]] = i
        nl[r] += 1
    for _ in range(ALIAS_COLS):
        r = np.nonzero((ns > 0) & (nl > 0))[0]
        if len(r) == 0:
            break
        sm, lg = small[r, ns[r] - 1], large[r, nl[r] - 1]
        ns[r] -= 1
        nl[r] -= 1
        cut[r, sm] = q[r, sm]
        alias[r, sm] = lg
        q[r, lg] = q[r, lg] - (1.0 - q[r, sm])
        lt = q[r, lg] < 1.0
        rs, rl = r[lt], r[~lt]
        small[rs, ns[rs]] = lg[lt]
        ns[rs] += 1
        large[rl, nl[rl]] = lg[~lt]
        nl[rl] += 1
    top = np.argmax(P, axis=1)
    for stack, depth in ((small, ns), (large, nl)):      # leftovers are full columns up to rounding
        for j in range(ALIAS_COLS):
            r = np.nonzero(depth > j)[0]
            col = stack[r, j]
            cut[r, col] = 1.0
            alias[r, col] = np.where(col < n_own, col, top[r])
    one = 1 << ALIAS_FRAC_BITS
    c = np.floor(cut * one + 0.5).astype(np.int64)
    cols = np.arange(ALIAS_COLS)[None, :]
    full = c >= one
    a = np.where(full & (cols < n_own), cols, alias)
    c = np.where(full | (cols >= n_own), 0, np.maximum(c, 0))
    return (((one - 1 - c) << 3) | a).astype(np.uint32)


def alias_variant_probabilities(row) -> np.ndarray:
    """Exact probabilities of the 8 draw variants an alias row encodes (for tests)."""
    one = 1 << ALIAS_FRAC_BITS
    p = np.zeros(8)
    for col in range(ALIAS_COLS):
        c, a = one - 1 - (int(row[col]) >> 3), int(row[col]) & 7      # exactly c of the column's 2^29 draws keep `col`
        p[col] += c / float(one) / ALIAS_COLS
        p[a] += (1.0 - c / float(one)) / ALIAS_COLS
    return p


def alias_probabilities(row) -> np.ndarray:
    """Exact probabilities of the 7 outcomes a table row encodes (both single variants count as 1B)."""
    p = alias_variant_probabilities(row)
    p[2] += p[VARIANT_1B_ADVANCE]
    return p[:7]


def _table_rows(m: Matchup):
    """(side, pitcher, tier, slot) of every row the kernel can reach for this matchup."""
    idx = [(s, p, tier, b) for s in range(2) for p in range(m.bullpen_len[s] + 1)
           for tier in range(_abi.N_TIERS) for b in range(m.lineup_len[s])]
    return tuple(np.array(x, dtype=np.int64) for x in zip(*idx))


def _fatigue_rows(m: Matchup, S, P):
    """Rows that need pitch-count slopes: both starters of a matchup in which some staff has no bullpen (see
    `tables_from_arrays`)."""
    return np.nonzero((P == 0) & (min(m.bullpen_len) == 0))[0]


def _warm_clip_cache(ms) -> None:
    """One threaded pass of the noise integral over every distinct (k, bb) of a batch of matchups; the per-matchup
    builds that follow are cache hits.  Pure optimisation: values are the ones `build_table` would compute."""
    ks, bs = [], []
    for m in ms:
        if not m.use_noise:
            continue
        S, P, T, B = _table_rows(m)
        k, bb = _effective_rates(m, S, P, T, B, np.zeros(len(S)))
        ks.append(k), bs.append(bb)
        fat = _fatigue_rows(m, S, P)
        if len(fat):
            k, bb = _effective_rates(m, S[fat], P[fat], T[fat], B[fat], np.full(len(fat), FATIGUE_START + SLOPE_SPAN))
            ks.append(k), bs.append(bb)
    if not ks:
        return
    kb = np.unique(np.stack([np.concatenate(ks), np.concatenate(bs)], axis=1), axis=0)
    k, bb = kb[:, 0], kb[:, 1]
    both = (bb > 0.0) & (bb < 1.0) & (k > 0.0) & (k < 1.0)      # the branch of _p_k_or_bb_batch that integrates
    if both.any():
        _expect_clip_cached(k[both], bb[both])


def build_table(m, use_noise=None) -> np.ndarray:
    """One `mlbsim_table` (include/mlbsim.h) from a matchup: alias rows for every reachable (side, pitcher, TTO tier,
    lineup slot) and, for starters that are never replaced, 31-bit thresholds with pitch-count slopes."""
    m = as_matchup(m) if use_noise is None else as_matchup(m, use_noise)
    if use_noise is not None and bool(use_noise) != bool(m.use_noise):
        m = copy.copy(m)
        m.use_noise = bool(use_noise)
    return build_tables([m])[0]


def stack_matchups(ms) -> dict:
    """Per-matchup numbers of `Matchup` objects stacked along a leading axis: the input of `tables_from_arrays`."""
    return {
        "lineup_len": np.array([m.lineup_len for m in ms], dtype=np.int64).reshape(len(ms), 2),
        "bullpen_len": np.array([m.bullpen_len for m in ms], dtype=np.int64).reshape(len(ms), 2),
        "use_noise": np.array([bool(m.use_noise) for m in ms], dtype=bool),
        "k_mod": np.array([m.k_mod for m in ms], dtype=np.float64),
        "bb_mod": np.array([m.bb_mod for m in ms], dtype=np.float64),
        "bat_k": np.stack([m.bat_k for m in ms]), "bat_bb": np.stack([m.bat_bb for m in ms]),
        "pit_k": np.stack([m.pit_k for m in ms]), "pit_bb": np.stack([m.pit_bb for m in ms]),
        "pit_hr": np.stack([m.pit_hr for m in ms]), "hit_prob": np.stack([m.hit_prob for m in ms]),
    }


def params_from_arrays(A: dict) -> np.ndarray:
    """`mlbsim_params` records (PARAMS_DTYPE) of a batch of matchups given as stacked arrays (`stack_matchups`): the input
    of the DEVICE table builder (`Context.build_tables`, csrc/table_kernel.cu) — and of replay mode, one record at a
    time.  A few array copies; the tables themselves are then built in HBM."""
    n = len(A["use_noise"])
    out = np.zeros(n, dtype=_abi.PARAMS_DTYPE)
    out["lineup_len"], out["bullpen_len"] = A["lineup_len"], A["bullpen_len"]
    out["use_noise"] = A["use_noise"]
    for f in ("k_mod", "bb_mod", "bat_k", "bat_bb", "pit_k", "pit_bb", "pit_hr", "hit_prob"):
        out[f] = A[f]
    out["cdf_bip"], out["cdf_hit"], out["cdf_fb"] = contact_cdfs()
    return out


def params_batch(matchups, use_noise=True) -> np.ndarray:
    """`params_from_arrays` for `simulate_game` keyword dicts / Matchup objects."""
    ms = [m if isinstance(m, Matchup) else as_matchup(m, use_noise) for m in matchups]
    if not ms:
        return np.zeros(0, dtype=_abi.PARAMS_DTYPE)
    return params_from_arrays(stack_matchups(ms))


def _cdf_tensor(A, valid, fl):
    """Cumulative outcome law of every valid (matchup, side, pitcher, tier, slot) row at fatigue level `fl`
    (0 = at most 75 pitches): the arithmetic of `outcome_cdf_batch`, same operations in the same order, evaluated
    once per DISTINCT row of inputs.  Returns (float64 [n_distinct, 6], inverse | None, row coordinates): row i of the
    valid rows has the law `c[inverse[i]]` (`c[i]` when nothing repeats)."""
    M, S, P, T, B = np.nonzero(valid)
    X = np.empty((len(M), 13))
    X[:, 0], X[:, 1], X[:, 2] = A["pit_k"][M, S, P], A["pit_bb"][M, S, P], A["pit_hr"][M, S, P]
    X[:, 3], X[:, 4] = A["bat_k"][M, S, B], A["bat_bb"][M, S, B]
    X[:, 5], X[:, 6], X[:, 7] = A["k_mod"][M], A["bb_mod"][M], A["use_noise"][M]
    X[:, 8] = T
    X[:, 9:13] = A["hit_prob"][M, S, P, B]
    u = _unique_rows(X) if len(M) >= 4096 else None
    if u is None:
        return _cdf_rows(X, fl), None, (M, S, P, T, B)
    index, inverse = u
    return _cdf_rows(X[index], fl), inverse, (M, S, P, T, B)


def _cdf_rows(X, fl):
    """Rows of [pit_k, pit_bb, pit_hr, bat_k, bat_bb, k_mod, bb_mod, use_noise, tier, hit_prob x 4] -> cumulative
    outcome law [n, 6] (K, K+BB, +1B, +2B, +3B, +HR)."""
    pen = np.asarray(TTO_PENALTY)[X[:, 8].astype(np.int64)]
    ak = X[:, 0] * (1 - pen) * (1.0 - 0.02 * fl)
    ab = X[:, 1] * (1 + pen) * (1.0 + 0.03 * fl)
    k = (X[:, 3] + ak) / 2 * X[:, 5]
    bb = (X[:, 4] + ab) / 2 * X[:, 6]
    noise = X[:, 7] != 0.0
    pk = np.empty(len(X))
    pkbb = np.empty(len(X))
    for flag in (True, False):
        sel = noise == flag
        if sel.any():
            pk[sel], pkbb[sel] = _p_k_or_bb_batch(k[sel], bb[sel], flag)
    contact = 1.0 - pkbb
    cdf_bip, cdf_hit, cdf_fb = contact_cdfs()
    w_bip = np.diff(np.concatenate([[0.0], cdf_bip]))
    h = np.diff(np.concatenate([[0.0], cdf_hit]))
    f = np.diff(np.concatenate([[0.0], cdf_fb]))
    hr = np.clip(X[:, 2], 0.0, 1.0)
    hp = X[:, 9:13]
    ph = hp[..., 0] * w_bip[0] + hp[..., 1] * w_bip[1] + hp[..., 2] * w_bip[2] + hp[..., 3] * w_bip[3]   # fixed order
    s1 = (1 - hr) * (ph * h[0] + (1 - ph) * INFIELD_FALLBACK * f[0])
    s2 = (1 - hr) * (ph * h[1] + (1 - ph) * INFIELD_FALLBACK * f[1])
    s3 = (1 - hr) * ph * h[2]
    c = np.stack([pk, pkbb, pkbb + contact * s1, pkbb + contact * (s1 + s2), pkbb + contact * (s1 + s2 + s3),
                  pkbb + contact * (s1 + s2 + s3 + hr)], axis=1)
    return np.minimum(np.maximum.accumulate(c, axis=1), 1.0)


_HASH_MULT = None


def _unique_rows(X: np.ndarray):
    """(index, inverse) such that X[index][inverse] == X, found through a 64-bit multiplicative hash of the rows' bit
    patterns (one pass over the data instead of a lexicographic sort of 100-byte rows) and verified exactly; None when
    nothing repeats or — never observed — two different rows collide."""
    global _HASH_MULT
    n, k = X.shape
    if _HASH_MULT is None or len(_HASH_MULT) < k:
        _HASH_MULT = np.random.default_rng(0x6D6C62).integers(1, 2**63, size=max(k, 32), dtype=np.uint64) | np.uint64(1)
    bits = np.ascontiguousarray(X, dtype=np.float64).view(np.uint64)
    with np.errstate(over="ignore"):
        h = (bits * _HASH_MULT[None, :k]).sum(axis=1, dtype=np.uint64)
    _, index, inverse = np.unique(h, return_index=True, return_inverse=True)
    if len(index) > 0.8 * n:
        return None
    if not np.array_equal(bits[index][inverse], bits):
        return None
    return index, inverse.reshape(-1)


MatchupCannotScore = _abi.MatchupCannotScore   # (defined beside the ABI: the device builder raises it too)


def tables_from_arrays(A: dict, validate: bool = True, alias_fn=None) -> np.ndarray:
    """`mlbsim_table` records (include/mlbsim.h) for a batch of matchups given as stacked arrays (`stack_matchups`):
    alias rows for every reachable (side, pitcher, TTO tier, lineup slot) and, for starters that are never replaced,
    31-bit thresholds with pitch-count slopes.  One pass of the noise integral and one of the alias construction
    over the DISTINCT rows of the whole batch (a sweep over one pitcher repeats the other side's rows in every grid
    point); no per-matchup Python.  `alias_fn`: a drop-in for `alias_rows`, e.g. `Context.alias_rows` (the device builder,
    bit-identical)."""
    alias_fn = alias_fn or alias_rows
    n = len(A["use_noise"])
    out = np.zeros(n, dtype=_abi.TABLE_DTYPE)
    if n == 0:
        return out
    P_, T_, B_ = _abi.MAX_PITCHERS, _abi.N_TIERS, _abi.MAX_LINEUP
    pit = np.arange(P_)[None, None, :, None, None]
    slot = np.arange(B_)[None, None, None, None, :]
    valid = (pit <= A["bullpen_len"][:, :, None, None, None]) & (slot < A["lineup_len"][:, :, None, None, None])
    valid = np.broadcast_to(valid, (n, 2, P_, T_, B_))
    alias = np.zeros((n,) + _abi.TABLE_DTYPE["alias"].shape, dtype=np.uint32)   # unused rows: zero words (never reached)
    fthr = np.full((n,) + _abi.TABLE_DTYPE["fthr"].shape, 2147483648, dtype=np.uint32)
    fslope = np.zeros((n,) + _abi.TABLE_DTYPE["fslope"].shape, dtype=np.int32)
    c0, inv0, where = _cdf_tensor(A, valid, 0.0)
    probs = np.diff(np.concatenate([np.zeros((len(c0), 1)), c0, np.ones((len(c0), 1))], axis=1), axis=1)   # distinct rows
    expand = (lambda x: x) if inv0 is None else (lambda x: x[inv0])
    if validate:
        # a matchup terminates with probability 1 iff somebody can reach base (K = column 0, OUT = column 6)
        can_reach = np.zeros(n, dtype=bool)
        np.logical_or.at(can_reach, where[0], expand((probs[:, 0] + probs[:, 6]) < 1.0))
        if not can_reach.all():
            raise MatchupCannotScore(f"matchup(s) {np.nonzero(~can_reach)[0][:8].tolist()}: no plate appearance of either side "
                                     "can put a runner on base, the game would never end")
    alias[where] = expand(alias_fn(outcome_variants(probs)))
    # A starter without a bullpen is never replaced: his pitch count passes 75 (core/fatigue_modeling.py:38).  Such a
    # matchup is flagged and BOTH of its starters get pitch-count rows: one who has a bullpen normally leaves at his third
    # time through the order, long before 75 pitches, but lineup wraps that coincide with the end of a half do not
    # count (half_inning_simulator.py:177-178), so once in ~10^4 games he gets there too.  (With bullpens on both sides
    # the flag stays off and that rare starter keeps his 75-pitch law for the few PAs until the hook at 90.)
    flagged = (A["bullpen_len"] == 0).any(axis=1)
    fat = valid & (pit == 0) & flagged[:, None, None, None, None]
    if fat.any():
        c1, inv1, (M, S, P, T, B) = _cdf_tensor(A, fat, max(0.0, SLOPE_SPAN / 25))
        if inv1 is not None:
            c1 = c1[inv1]
        c0f = expand(c0)[fat[where]]    # the 75-pitch law of the same rows, already computed above (same row order)
        q0, q1 = _quantise(c0f), _quantise(c1)
        fthr[M, S, T, B, :6] = q0
        fslope[M, S, T, B, :6] = np.round((q1.astype(np.int64) - q0.astype(np.int64)) / SLOPE_SPAN)
        out["flags"][np.unique(M)] |= _abi.FLAG_FATIGUE
    out["magic"] = _abi.TABLE_MAGIC
    out["lineup_len"], out["bullpen_len"] = A["lineup_len"], A["bullpen_len"]
    out["alias"], out["fthr"], out["fslope"] = alias, fthr, fslope
    return out


def build_tables(matchups, use_noise=True, validate: bool = True, alias_fn=None) -> np.ndarray:
    """Tables for a batch of matchups (dicts take `use_noise` as the default for a missing key).  `validate=False`
    skips the can-anybody-score check (tests of the kernel's own termination guard); `alias_fn`: see
    `tables_from_arrays`."""
    ms = [m if isinstance(m, Matchup) else as_matchup(m, use_noise) for m in matchups]
    if not ms:
        return np.zeros(0, dtype=_abi.TABLE_DTYPE)
    return tables_from_arrays(stack_matchups(ms), validate=validate, alias_fn=alias_fn)
