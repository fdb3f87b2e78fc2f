/* mlbsim.h — C ABI of the B200 Monte Carlo game sampler.
 *
 * The reference (BLON333/MLB_Odds_Engine) has no FFI: its seam is the Python call
 *   simulate_game(home_lineup, away_lineup, home_pitcher, away_pitcher, env, home_bullpen,
 *                 away_bullpen, ..., use_noise)              core/game_simulator.py:14-25
 * made once per simulation from cli/run_distribution_simulator.py:368-382, plus the process-global
 * outcome counters PA_RECAP_LOG (core/pa_simulator.py:10-20) that tools/ read.  This header is
 * what a binding for that seam links against (ctypes stub: INTEGRATION.md).  Plain C, no torch
 * or CUDA types; every pointer is caller-owned; nothing throws across the boundary.
 */
#ifndef MLBSIM_H
#define MLBSIM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MLBSIM_ABI_VERSION 6

/* ---- model dimensions ------------------------------------------------------------------- */
#define MLBSIM_MAX_LINEUP 9   /* len(lineup); 9 in every caller (half_inning_simulator.py:175) */
#define MLBSIM_MAX_BULLPEN 6  /* max_relievers=6, assets/bullpen_utils.py:19,104 */
#define MLBSIM_MAX_PITCHERS 7 /* starter (index 0) + bullpen (index 1 + position in the list) */
#define MLBSIM_N_OUTCOMES 7   /* K BB 1B 2B 3B HR OUT — order of PA_RECAP_LOG, pa_simulator.py:11 */
#define MLBSIM_N_TIERS 4      /* TTO 1, 2, 3, >=4 — fatigue_modeling.py:25-32 */
#define MLBSIM_N_BIP 4        /* GB LD FB POP — pa_simulator.py:54 */

enum { MLBSIM_K = 0, MLBSIM_BB, MLBSIM_1B, MLBSIM_2B, MLBSIM_3B, MLBSIM_HR, MLBSIM_OUT };

/* "side" everywhere = the BATTING side: 0 = away bats (top half, home staff pitches),
 * 1 = home bats (bottom half, away staff pitches).  pit_* [side] is the staff that FACES it. */

/* ---- raw float64 parameters of one matchup (replay mode) ---------------------------------- */
typedef struct mlbsim_params {
  int32_t lineup_len[2];  /* [side] */
  int32_t bullpen_len[2]; /* [side]: relievers of the staff facing `side`; 0 = None or [] */
  int32_t use_noise;      /* simulate_game(use_noise=...) */
  int32_t reserved;
  double k_mod, bb_mod; /* env["umpire"].get("k_mod"/"bb_mod", 1.0), pa_simulator.py:112-114 */
  double bat_k[2][MLBSIM_MAX_LINEUP];
  double bat_bb[2][MLBSIM_MAX_LINEUP];
  double pit_k[2][MLBSIM_MAX_PITCHERS];
  double pit_bb[2][MLBSIM_MAX_PITCHERS];
  double pit_hr[2][MLBSIM_MAX_PITCHERS]; /* hr_pa_projected * weather_hr, pa_simulator.py:46 */
  /* resolve_bip probability after all modifiers and the [0.05,0.55] clamp, computed on the host
   * in float64 in the reference's multiply order (core/bip_resolution.py:20-57) */
  double hit_prob[2][MLBSIM_MAX_PITCHERS][MLBSIM_MAX_LINEUP][MLBSIM_N_BIP];
  double cdf_bip[4]; /* cumsum([.28,.32,.30,.10]) / last       pa_simulator.py:54 */
  double cdf_hit[3]; /* normalised [0.76*1.01, 0.22, 0.02]     pa_simulator.py:68-72 */
  double cdf_fb[2];  /* normalised [0.92*1.01, 0.08]           pa_simulator.py:75-79 */
} mlbsim_params;

/* ---- per-game integers replay mode returns (core/game_simulator.py:134-144) ---------------- */
#define MLBSIM_REPLAY_MAX_INNINGS 40
#define MLBSIM_REPLAY_MAX_USED 16
#define MLBSIM_ST_INNINGS_OVERFLOW 1 /* more innings than recorded; scores still exact */
#define MLBSIM_ST_USED_OVERFLOW 2
#define MLBSIM_ST_TAPE_UNDERRUN 4 /* the tape ended before the game did */
#define MLBSIM_ST_TRUNCATED 8     /* native / per-sim mode: still tied after MLBSIM_MAX_INNINGS innings */

typedef struct mlbsim_replay_game {
  int32_t home_score, away_score;
  int32_t n_innings;  /* len(result["innings"]) */
  int32_t tape_used;  /* tape entries consumed; == tape length for a well-formed tape */
  int32_t n_used[2];  /* [0] len(used_home_relievers), [1] len(used_away_relievers) */
  int32_t status;
  int32_t reserved;
  uint8_t away_runs[MLBSIM_REPLAY_MAX_INNINGS]; /* innings[i]["away_runs"] */
  uint8_t home_runs[MLBSIM_REPLAY_MAX_INNINGS];
  uint8_t used_home[MLBSIM_REPLAY_MAX_USED]; /* indices into home_bullpen */
  uint8_t used_away[MLBSIM_REPLAY_MAX_USED];
  uint16_t pa[2][8]; /* [side][outcome] per-game PA_RECAP_LOG deltas; [7] unused */
} mlbsim_replay_game;

/* ---- native mode: per-matchup integer sampling table -----------------------------------------
 * One PA consumes one Philox2x32-10 block (w0, w1).  w0 draws one of EIGHT variants with Walker's
 * alias method on 8 columns — the seven outcomes K BB 1B 2B 3B HR OUT, with the single split by the
 * independent 0.8 draw "runner on first takes second" (core/half_inning_simulator.py:105):
 * variant 2 = single, runner holds; variant 7 (MLBSIM_V_1B_ADV) = single, runner advances:
 *     col = w0 >> 29;  e = alias[side][pit][tier][slot][col];      e = (2^29 - 1 - cut) << 3 | alias
 *     variant = ((w0 & 0x1FFFFFFF) > (e >> 3)) ? col : (e & 7)    (== (w0 << 3) > e: one shift, one compare)
 * i.e. one table word per PA, and exactly `cut` of the 2^29 draws of a column keep the column's own variant — a
 * variant of probability 0 never occurs (the complemented cut makes the tie with the alias bits fall to the alias).  Starters past 75 pitches (fatigue_modeling.py:38; only possible
 * for a staff without a bullpen, MLBSIM_FLAG_FATIGUE) use cumulative 31-bit thresholds instead:
 *     u = w0 >> 1;  outcome = #{ j < 6 : u >= fthr[..][j] + (pitch_count - 75) * fslope[..][j] }
 * and take the 0.8 draw from the low 16 bits of w1.                                              */
#define MLBSIM_TABLE_MAGIC 0x4d4c4233u /* "3BLM": alias rows hold the 8 draw variants, cuts stored complemented */
#define MLBSIM_V_1B_ADV 7
#define MLBSIM_ROW_WORDS 8
#define MLBSIM_FLAG_FATIGUE 1u /* some staff has no bullpen: pitch_count can exceed 75 */

typedef struct mlbsim_table {
  uint32_t magic;
  uint32_t flags;
  uint32_t lineup_len[2];
  uint32_t bullpen_len[2];
  uint32_t reserved[2];
  uint32_t alias[2][MLBSIM_MAX_PITCHERS][MLBSIM_N_TIERS][MLBSIM_MAX_LINEUP][MLBSIM_ROW_WORDS];
  uint32_t fthr[2][MLBSIM_N_TIERS][MLBSIM_MAX_LINEUP][MLBSIM_ROW_WORDS];  /* starters, at pitch_count 75 */
  int32_t fslope[2][MLBSIM_N_TIERS][MLBSIM_MAX_LINEUP][MLBSIM_ROW_WORDS]; /* per pitch above 75 */
} mlbsim_table;

/* ---- native mode: per-matchup output, all uint64 counters ---------------------------------- */
#define MLBSIM_HIST_BINS 64 /* runs per team, last bin = ">= 63" */
#define MLBSIM_N_CAPS 5     /* first 1, 3, 5, 7 innings, full game (run_distribution_simulator.py:414) */
#define MLBSIM_INN_BINS 32  /* innings played, last bin = ">= 31" */
/* Termination guard.  The reference loops while the score is tied (core/game_simulator.py:110-111: no inning cap), so
 * a matchup in which neither side can score never returns there.  Here a game that is still tied after
 * MLBSIM_MAX_INNINGS innings stops, is binned like any other game and is counted in mlbsim_hist.n_truncated (the Python
 * layer raises when that is non-zero).  128 innings x 2 halves x 30 plate appearances (half_inning_simulator.py:160)
 * = 7 680 < 2^13, so the plate-appearance field of the Philox counter cannot carry into the matchup field. */
#define MLBSIM_MAX_INNINGS 128

typedef struct mlbsim_hist {
  uint64_t joint[MLBSIM_N_CAPS][MLBSIM_HIST_BINS][MLBSIM_HIST_BINS]; /* [cap][away][home] */
  uint64_t innings[MLBSIM_INN_BINS];
  uint64_t pa[2][8];        /* [side][outcome] = PA_RECAP_LOG["AWAY"|"HOME"]; [7] unused */
  uint64_t rel_games[2][8]; /* [0] home bullpen, [1] away bullpen: sims in which reliever i appeared
                               (run_distribution_simulator.py:391-397) */
  uint64_t rel_picks[2][8]; /* total picks of reliever i */
  uint64_t n_games;
  uint64_t n_truncated;     /* games stopped by the termination guard below while still tied (counted in n_games too) */
  uint64_t reserved[6];
} mlbsim_hist;

#define MLBSIM_HIST_WORDS (sizeof(mlbsim_hist) / 8)

/* ---- native mode: per-matchup scalar summary (sweeps) ---------------------------------------
 * What tools/test_*_sensitivity.py and tools/simulate_calibration_game.py reduce N games to — outcome
 * counts (PA_RECAP_LOG), run means / SDs, win counts — as exact integer moments of one mlbsim_hist,
 * computed on the device so that a 1 024-point sweep returns 320 B per grid point instead of 165 KB.
 * Runs are clamped at 63 exactly as the histogram bins are. */
typedef struct mlbsim_summary {
  uint64_t n_games;
  uint64_t pa[2][8];          /* copy of mlbsim_hist.pa */
  uint64_t sum_runs[MLBSIM_N_CAPS][2]; /* [cap 1,3,5,7,full][0 away, 1 home]: sum over sims of runs */
  uint64_t sum_sq[2];         /* full game: sum of away^2, home^2 */
  uint64_t sum_cross;         /* full game: sum of away*home */
  uint64_t home_wins;         /* full game: sims with home > away */
  uint64_t away_wins;
  uint64_t sum_innings;       /* sum over sims of innings played (bin 31 counts as 31) */
  uint64_t n_truncated;       /* copy of mlbsim_hist.n_truncated */
  uint64_t reserved[6];
} mlbsim_summary;

/* ---- error codes --------------------------------------------------------------------------- */
#define MLBSIM_OK 0
#define MLBSIM_ERR_CUDA 1       /* a CUDA runtime call failed; see mlbsim_last_error */
#define MLBSIM_ERR_ARG 2        /* bad argument (NULL, range, size mismatch) */
#define MLBSIM_ERR_TABLE 3      /* table failed validation (magic, lengths) */
#define MLBSIM_ERR_NO_DEVICE 4  /* no CUDA device / not an sm_100 part */
#define MLBSIM_ERR_STATE 5      /* call order (e.g. run before upload) */
#define MLBSIM_ERR_CANNOT_SCORE 6 /* mlbsim_build_tables: a matchup in which nobody can reach base (the game never ends) */

typedef struct mlbsim_ctx mlbsim_ctx;

/* Creates a context bound to CUDA device `device` with its own stream.  Replaces nothing in the
 * reference (it has no device); one ctx per host thread. */
int mlbsim_init(int device, mlbsim_ctx** out);
void mlbsim_destroy(mlbsim_ctx* ctx);
const char* mlbsim_last_error(const mlbsim_ctx* ctx); /* ctx may be NULL: last init error */
int mlbsim_abi_version(void);
/* sizeof of the structs above as compiled into the library: 0 params, 1 table, 2 hist,
 * 3 replay_game, 4 summary (a binding checks its own mirrors against these at load time). */
size_t mlbsim_struct_size(int which);

/* Device facts for roofline arithmetic. */
int mlbsim_device_info(const mlbsim_ctx* ctx, int* sm_count, int* sm_clock_khz, char* name, size_t name_len);

/* Uses an externally owned stream (e.g. torch's current stream) for all later calls; 0 restores
 * the context's own stream.  `stream` is a cudaStream_t passed as an integer; pass 1
 * (cudaStreamLegacy) to name the legacy default stream, whose handle is otherwise 0. */
int mlbsim_set_stream(mlbsim_ctx* ctx, uintptr_t stream);
int mlbsim_synchronize(mlbsim_ctx* ctx);

/* Copies `n_matchups` host tables to the device (replaces the per-call dict arguments of
 * simulate_game, core/game_simulator.py:14-21, for native mode).  The host buffer may be reused on return.  Up to
 * 1 MiB of tables (50 matchups) are copied through a pinned buffer of the context and the call does not wait for the
 * device: the copy and the staging kernel are ordered on the context's stream ahead of whatever is launched next
 * (mlbsim_set_stream makes a new stream wait for them).  Larger uploads return when the copy has completed. */
int mlbsim_upload_tables(mlbsim_ctx* ctx, const mlbsim_table* tables, int n_matchups);

/* Matchup sharding ("for the sweep: by matchup block"): a rank that uploads only the tables of its
 * own block tells the library which global matchup id its table 0 has, so that the Philox counter of uploaded table i
 * carries id `base + i` and a grid point draws the same games whichever rank owns it.  Default 0; base + n_matchups
 * must stay <= 4096. */
int mlbsim_set_matchup_base(mlbsim_ctx* ctx, int base);

/* Table construction helper: Walker alias rows for `n_rows` probability rows of `n_own` (7 or 8) float64 entries each
 * (host pointers; out receives n_rows * 8 words, entry = (2^29 - 1 - cut29) << 3 | alias as in mlbsim_table.alias).  Bit-identical
 * to the host builder (mlb_odds_engine_b200/tables.py::alias_rows): one thread per row runs the same float64
 * operations in the same order.  Replaces nothing in the reference (it samples with np.random.choice per PA,
 * core/pa_simulator.py:54-79); it takes the per-row Python/numpy work out of building 10^3-point sweeps. */
int mlbsim_alias_rows(mlbsim_ctx* ctx, const double* probs_host, int n_rows, int n_own, uint32_t* out_host);

/* Table construction on the device: the sampling tables of `n_matchups` matchups from their raw float64 numbers (the
 * mlbsim_params of replay mode: what core/pa_simulator.py:91-143, core/bip_resolution.py and core/fatigue_modeling.py read
 * per plate appearance), installed like mlbsim_upload_tables — one block per matchup, one warp per (side, pitcher, tier,
 * slot) row: effective rates, the Beta-noise expectation E[min(1, kp + 1.01 bp)] by 48-node Gauss-Legendre quadrature
 * (lanes = nodes), the contact split, the Walker alias row, and for never-replaced starters the pitch-count thresholds.
 * Everything but the quadrature is the host builder's float64 arithmetic in the same order (bit-identical); the
 * quadrature agrees with the host's (scipy) to ~1e-16, so a 29/31-bit table entry can differ from the host-built one by
 * one unit, rarely.  A 1 024-point sweep builds in about a millisecond and its 20 MB of tables never exist on the host.
 * MLBSIM_ERR_CANNOT_SCORE (nothing installed; indices in mlbsim_last_error) when some matchup could never leave 0-0 —
 * the reference would loop forever there (core/game_simulator.py:110-111) — unless MLBSIM_BUILD_NO_VALIDATE is passed. */
#define MLBSIM_BUILD_NO_VALIDATE 1u
int mlbsim_build_tables(mlbsim_ctx* ctx, const mlbsim_params* params_host, int n_matchups, uint32_t flags);

/* Copy installed tables [matchup_lo, matchup_hi) back to the host (e.g. to hand the oracle the very bytes the device uses). */
int mlbsim_download_tables(mlbsim_ctx* ctx, int matchup_lo, int matchup_hi, mlbsim_table* out_host);

/* Native-RNG mode: for every uploaded matchup m in [matchup_lo, matchup_hi) simulate the sims
 * [sim_lo, sim_hi) — the loop of cli/run_distribution_simulator.py:368-382 — and ADD the results
 * into hist_dev[m - matchup_lo] (device memory, caller-allocated, caller-zeroed).  Sim i of
 * matchup m draws one Philox2x32-10 block per plate appearance with counter
 * (i[31:0], pa_number[12:0] | (base + m) << 14 | i[37:32] << 26) and a 32-bit key derived from `seed`
 * (word 0 of Philox4x32-10 over a fixed domain-separation counter), so its outcome does not depend
 * on the launch geometry or the number of GPUs.  Limits: m < 4096, i < 2^38, < 2^13 plate appearances per game
 * (guaranteed by MLBSIM_MAX_INNINGS).  Asynchronous on the
 * context's stream.  `flags` bit 0: 0 = also count PA outcomes (default), 1 = skip them. */
int mlbsim_run_native(mlbsim_ctx* ctx, int matchup_lo, int matchup_hi, uint64_t sim_lo, uint64_t sim_hi,
                      uint64_t seed, mlbsim_hist* hist_dev, uint32_t flags);

/* Host-buffer convenience used by the Python layer and the e2e benchmark: zero device
 * histograms, run, copy to `hist_host`, synchronise. */
int mlbsim_run_native_host(mlbsim_ctx* ctx, int matchup_lo, int matchup_hi, uint64_t sim_lo, uint64_t sim_hi,
                           uint64_t seed, mlbsim_hist* hist_host, uint32_t flags);

/* Reduce `n_matchups` device histograms to summaries (device pointers; asynchronous on the context's
 * stream).  Replaces the per-grid-point reductions of tools/test_stuff_plus_sensitivity.py:71-85 and
 * tools/simulate_calibration_game.py:94-128. */
int mlbsim_summarize(mlbsim_ctx* ctx, const mlbsim_hist* hist_dev, int n_matchups, mlbsim_summary* summary_dev);

/* Host-buffer convenience for sweeps: zero device histograms, run, summarise on the device, copy
 * `matchup_hi - matchup_lo` summaries to `summary_host`, synchronise.  The full histograms stay on the
 * device (scratch) and are not copied. */
int mlbsim_run_native_summary_host(mlbsim_ctx* ctx, int matchup_lo, int matchup_hi, uint64_t sim_lo, uint64_t sim_hi,
                                   uint64_t seed, mlbsim_summary* summary_host, uint32_t flags);

/* Multi-GPU exchange step: in-place ncclAllReduce(SUM, uint64) of `n_matchups` device
 * histograms over the caller's communicator, enqueued on the context's stream right behind the kernel (no host
 * synchronisation in between).  `nccl_comm` is an `ncclComm_t` the caller created (one rank per context / GPU);
 * the library resolves `ncclAllReduce` from libnccl.so.2 at first use (dlopen; override the path with the
 * environment variable MLBSIM_NCCL_LIB) and does not link NCCL.  Integer sums: totals are bit-identical for any
 * number of ranks.  Hosts that already run torch.distributed can all-reduce the same buffer there instead
 * (mlb_odds_engine_b200/dist.py). */
int mlbsim_allreduce(mlbsim_ctx* ctx, void* nccl_comm, mlbsim_hist* hist_dev, int n_matchups);

/* Per-sim variant of native mode for small N (the compatibility path under
 * cli/run_distribution_simulator.py:368-389, which reads result["home_score"], ["away_score"],
 * ["innings"][i]{away_runs,home_runs}, ["used_*_relievers"]): the same law, the same Philox
 * streams and therefore the same games as mlbsim_run_native, but one mlbsim_replay_game record
 * per sim (tape_used = plate appearances played) written to `out_host` (sim_hi - sim_lo records).
 * Synchronous. */
int mlbsim_run_persim(mlbsim_ctx* ctx, int matchup, uint64_t sim_lo, uint64_t sim_hi, uint64_t seed,
                      mlbsim_replay_game* out_host);

/* Replay mode: game g consumes tape[tape_offsets[g] .. tape_offsets[g+1]) (float64 values in the
 * reference's draw order) and writes out[g].  Host pointers; synchronous.
 * Integer results are bit-exact with core/game_simulator.py:14-156 on the same draws. */
int mlbsim_run_replay(mlbsim_ctx* ctx, const mlbsim_params* params, const double* tape,
                      const uint64_t* tape_offsets, int64_t n_games, mlbsim_replay_game* out);

/* Device-pointer variant (tape, offsets and out already resident in HBM); asynchronous.  The kernel streams every game's
 * tape into shared memory in aligned 32-byte granules (16-byte cp.async copies), so `tape_dev` must be 16-byte aligned
 * and READABLE for 16 entries (128 bytes) past the last draw (tape_offsets[n_games]); what the slack holds does not
 * matter.  Buffers from mlbsim_malloc qualify.  Per-game outcome counters are exact below 512 equal outcomes per side. */
int mlbsim_run_replay_dev(mlbsim_ctx* ctx, const mlbsim_params* params_host, const double* tape_dev,
                          const uint64_t* tape_offsets_dev, int64_t n_games, mlbsim_replay_game* out_dev);

/* Which instance of the replay kernel these parameters select (host-only, needs no device, ctx may be NULL):
 * MLBSIM_REPLAY_KIND_NONOISE (use_noise == 0), MLBSIM_REPLAY_KIND_REGULAR (noise on, a bullpen on both sides — so nobody
 * pitches past 120 plate appearances, core/game_simulator.py:8-12 — and every K / BB rate of core/fatigue_modeling.py:25-43 +
 * core/pa_simulator.py:108-114 strictly inside (0, 1) over the tabulated range: every plate appearance draws both Beta values,
 * the draw-by-draw path is not instantiated), else MLBSIM_REPLAY_KIND_GENERAL.  Same integers from every instance; the
 * environment variable MLBSIM_REPLAY_GENERAL=1 makes mlbsim_run_replay* use the general instance where it would pick the
 * regular one (tests, A/B).  Negative = MLBSIM_ERR_ARG. */
#define MLBSIM_REPLAY_KIND_GENERAL 0
#define MLBSIM_REPLAY_KIND_REGULAR 1
#define MLBSIM_REPLAY_KIND_NONOISE 2
int mlbsim_replay_kernel_kind(const mlbsim_params* params);

/* Device memory helpers so a ctypes-only caller needs no CUDA binding. */
int mlbsim_malloc(mlbsim_ctx* ctx, size_t bytes, void** dev_ptr);
int mlbsim_free(mlbsim_ctx* ctx, void* dev_ptr);
int mlbsim_memset(mlbsim_ctx* ctx, void* dev_ptr, int value, size_t bytes);
int mlbsim_memcpy_h2d(mlbsim_ctx* ctx, void* dev_dst, const void* host_src, size_t bytes);
int mlbsim_memcpy_d2h(mlbsim_ctx* ctx, void* host_dst, const void* dev_src, size_t bytes);

/* Kernel launches issued by this context since creation (bench.py's gpu_launches). */
uint64_t mlbsim_launch_count(const mlbsim_ctx* ctx);

/* Tuning knobs (0 = default): threads per block and blocks per SM of the native kernel. */
int mlbsim_set_launch_config(mlbsim_ctx* ctx, int threads_per_block, int blocks_per_sm);

/* Timing of the last mlbsim_run_native* kernel(s) in milliseconds, measured with CUDA events on
 * the launching stream (synchronises). */
int mlbsim_last_kernel_ms(mlbsim_ctx* ctx, float* ms);

#ifdef __cplusplus
}
#endif
#endif /* MLBSIM_H */
