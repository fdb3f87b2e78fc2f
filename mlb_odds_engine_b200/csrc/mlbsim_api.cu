// mlbsim_api.cu — the C ABI declared in include/mlbsim.h: context, uploads, launches.
// No CPU fallback: every compute entry point needs a CUDA device and fails with an error code
// (never an exception) when something is missing.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/mlbsim.h"
#include "native_kernel.h"

#ifndef NATIVE_FETCHES_PER_WARP
#define NATIVE_FETCHES_PER_WARP 32   /* measured: 8 -> 32 gives +2 % (single game) to +4.5 % (slate), profiles/r01_v14_fetch_variants.log */
#endif

struct mlbsim_ctx {
  int device = -1;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  cudaDeviceProp prop{};
  mlbsim_table* tables_dev = nullptr;          // as uploaded (the per-sim kernel reads these)
  mlbsim_staged_table* staged_dev = nullptr;   // shared-memory image of each table (the native kernel bulk-copies these)
  int n_tables = 0;
  int tables_cap = 0;
  int matchup_base = 0;   // global id of uploaded table 0 (matchup sharding)
  unsigned long long* counters_dev = nullptr;
  int counters_cap = 0;
  mlbsim_hist* hist_scratch = nullptr;  // used by *_host entry points
  int hist_scratch_cap = 0;
  mlbsim_summary* summary_scratch = nullptr;
  int summary_scratch_cap = 0;
  mlbsim_params* params_dev = nullptr;
  // small table uploads go through a context-owned pinned buffer: the caller's buffer is free on return and the call does not
  // wait for the device (the copy, the staging kernel and whatever is launched next queue up on the stream back to back)
  void* upload_stage = nullptr;
  cudaEvent_t upload_ev = nullptr;   // recorded behind the staging kernel of the last staged upload
  bool upload_pending = false;
  mlbsim_params params_stage{};   // context-owned pageable copy: the caller's buffer is free as soon as a replay call returns
  int threads_per_block = 0;
  int blocks_per_sm = 0;
  int occ_threads = -1, occ_blocks = 0;   // cached occupancy query of the native kernel
  bool replay_attr_set = false;           // the replay kernels have been granted their dynamic shared memory on this device
  bool nodes_set = false;                 // quadrature nodes of the table builder are in this device's constant memory
  uint64_t launches = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool timed = false;
  std::string err;
};

static thread_local std::string g_init_error;

#define CK(ctx, call)                                                                                   \
  do {                                                                                                  \
    cudaError_t e_ = (call);                                                                            \
    if (e_ != cudaSuccess) {                                                                            \
      (ctx)->err = std::string(#call) + ": " + cudaGetErrorName(e_) + " (" + cudaGetErrorString(e_) + ")"; \
      return MLBSIM_ERR_CUDA;                                                                           \
    }                                                                                                   \
  } while (0)

static int fail(mlbsim_ctx* ctx, int code, const char* msg) {
  if (ctx) ctx->err = msg;
  return code;
}

extern "C" {

int mlbsim_abi_version(void) { return MLBSIM_ABI_VERSION; }

size_t mlbsim_struct_size(int which) {
  switch (which) {
    case 0: return sizeof(mlbsim_params);
    case 1: return sizeof(mlbsim_table);
    case 2: return sizeof(mlbsim_hist);
    case 3: return sizeof(mlbsim_replay_game);
    case 4: return sizeof(mlbsim_summary);
    default: return 0;
  }
}

const char* mlbsim_last_error(const mlbsim_ctx* ctx) { return ctx ? ctx->err.c_str() : g_init_error.c_str(); }

int mlbsim_init(int device, mlbsim_ctx** out) {
  if (!out) { g_init_error = "mlbsim_init: out is NULL"; return MLBSIM_ERR_ARG; }
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    g_init_error = std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    return MLBSIM_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= n) { g_init_error = "mlbsim_init: device index out of range"; return MLBSIM_ERR_ARG; }
  mlbsim_ctx* ctx = new (std::nothrow) mlbsim_ctx();
  if (!ctx) { g_init_error = "out of host memory"; return MLBSIM_ERR_ARG; }
  ctx->device = device;
  auto bail = [&](const char* what, cudaError_t ce) {
    g_init_error = std::string(what) + ": " + cudaGetErrorString(ce);
    mlbsim_destroy(ctx);   // releases whatever was created so far
    return MLBSIM_ERR_CUDA;
  };
  if ((e = cudaSetDevice(device)) != cudaSuccess) return bail("cudaSetDevice", e);
  if ((e = cudaGetDeviceProperties(&ctx->prop, device)) != cudaSuccess) return bail("cudaGetDeviceProperties", e);
  if (ctx->prop.major != 10) {
    g_init_error = std::string("device '") + ctx->prop.name + "' is sm_" + std::to_string(ctx->prop.major) +
                   std::to_string(ctx->prop.minor) + "; this library is built for sm_100a only";
    delete ctx;
    return MLBSIM_ERR_NO_DEVICE;
  }
  if ((e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
  ctx->stream = ctx->own_stream;
  if ((e = cudaEventCreate(&ctx->ev0)) != cudaSuccess) return bail("cudaEventCreate", e);
  if ((e = cudaEventCreate(&ctx->ev1)) != cudaSuccess) return bail("cudaEventCreate", e);
  if ((e = cudaMalloc(&ctx->params_dev, sizeof(mlbsim_params))) != cudaSuccess) return bail("cudaMalloc", e);
  if ((e = cudaFuncSetAttribute(mlbsim_native_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)mlbsim_native_smem_bytes(MLBSIM_NATIVE_MAX_THREADS))) != cudaSuccess)
    return bail("cudaFuncSetAttribute", e);
  *out = ctx;
  return MLBSIM_OK;
}

void mlbsim_destroy(mlbsim_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->own_stream) cudaStreamSynchronize(ctx->own_stream);
  if (ctx->upload_pending) cudaEventSynchronize(ctx->upload_ev);   // (a copy out of the pinned staging buffer may still be in flight on a caller's stream)
  if (ctx->upload_ev) cudaEventDestroy(ctx->upload_ev);
  if (ctx->upload_stage) cudaFreeHost(ctx->upload_stage);
  cudaFree(ctx->tables_dev);
  cudaFree(ctx->staged_dev);
  cudaFree(ctx->counters_dev);
  cudaFree(ctx->hist_scratch);
  cudaFree(ctx->summary_scratch);
  cudaFree(ctx->params_dev);
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  delete ctx;
}

int mlbsim_device_info(const mlbsim_ctx* ctx, int* sm_count, int* sm_clock_khz, char* name, size_t name_len) {
  if (!ctx) return MLBSIM_ERR_ARG;
  if (sm_count) *sm_count = ctx->prop.multiProcessorCount;
  if (sm_clock_khz) *sm_clock_khz = ctx->prop.clockRate;
  if (name && name_len) {
    std::strncpy(name, ctx->prop.name, name_len - 1);
    name[name_len - 1] = 0;
  }
  return MLBSIM_OK;
}

int mlbsim_set_stream(mlbsim_ctx* ctx, uintptr_t stream) {
  if (!ctx) return MLBSIM_ERR_ARG;
  cudaStream_t s = stream ? reinterpret_cast<cudaStream_t>(stream) : ctx->own_stream;
  if (s != ctx->stream && ctx->upload_pending) {   // tables uploaded on the old stream must be in place for launches on the new one
    CK(ctx, cudaSetDevice(ctx->device));
    CK(ctx, cudaStreamWaitEvent(s, ctx->upload_ev, 0));
  }
  ctx->stream = s;
  return MLBSIM_OK;
}

int mlbsim_synchronize(mlbsim_ctx* ctx) {
  if (!ctx) return MLBSIM_ERR_ARG;
  CK(ctx, cudaSetDevice(ctx->device));
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  return MLBSIM_OK;
}

int mlbsim_set_launch_config(mlbsim_ctx* ctx, int threads_per_block, int blocks_per_sm) {
  if (!ctx) return MLBSIM_ERR_ARG;
  if (threads_per_block < 0 || threads_per_block > MLBSIM_NATIVE_MAX_THREADS || (threads_per_block % 32) != 0)
    return fail(ctx, MLBSIM_ERR_ARG, "threads_per_block must be a multiple of 32 in [0, 640]");
  if (blocks_per_sm < 0 || blocks_per_sm > 32) return fail(ctx, MLBSIM_ERR_ARG, "blocks_per_sm out of range");
  ctx->threads_per_block = threads_per_block;
  ctx->blocks_per_sm = blocks_per_sm;
  return MLBSIM_OK;
}

uint64_t mlbsim_launch_count(const mlbsim_ctx* ctx) { return ctx ? ctx->launches : 0; }

int mlbsim_set_matchup_base(mlbsim_ctx* ctx, int base) {
  if (!ctx) return MLBSIM_ERR_ARG;
  if (base < 0 || (unsigned)base >= RNG_MAX_MATCHUPS) return fail(ctx, MLBSIM_ERR_ARG, "matchup base outside [0, 4096)");
  ctx->matchup_base = base;
  return MLBSIM_OK;
}

// device buffers for n_matchups tables (raw + staged image) and their work counters
constexpr size_t UPLOAD_STAGE_BYTES = 1u << 20;   // 50 tables: single games and slates; sweeps upload from the caller's buffer

static int ensure_table_capacity(mlbsim_ctx* ctx, int n_matchups) {
  if (n_matchups > ctx->tables_cap) {
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->tables_dev);
    cudaFree(ctx->staged_dev);
    ctx->tables_dev = nullptr;
    ctx->staged_dev = nullptr;
    ctx->tables_cap = ctx->n_tables = 0;
    CK(ctx, cudaMalloc(&ctx->tables_dev, sizeof(mlbsim_table) * (size_t)n_matchups));
    CK(ctx, cudaMalloc(&ctx->staged_dev, sizeof(mlbsim_staged_table) * (size_t)n_matchups));
    ctx->tables_cap = n_matchups;
  }
  if (n_matchups > ctx->counters_cap) {
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->counters_dev);
    ctx->counters_dev = nullptr;
    CK(ctx, cudaMalloc(&ctx->counters_dev, sizeof(unsigned long long) * (size_t)n_matchups));
    ctx->counters_cap = n_matchups;
  }
  return MLBSIM_OK;
}

int mlbsim_upload_tables(mlbsim_ctx* ctx, const mlbsim_table* tables, int n_matchups) {
  if (!ctx) return MLBSIM_ERR_ARG;
  if (!tables || n_matchups <= 0) return fail(ctx, MLBSIM_ERR_ARG, "mlbsim_upload_tables: no tables");
  for (int i = 0; i < n_matchups; ++i) {
    const mlbsim_table& t = tables[i];
    if (t.magic != MLBSIM_TABLE_MAGIC) return fail(ctx, MLBSIM_ERR_TABLE, "table magic mismatch");
    for (int s = 0; s < 2; ++s) {
      if (t.lineup_len[s] < 1 || t.lineup_len[s] > MLBSIM_MAX_LINEUP) return fail(ctx, MLBSIM_ERR_TABLE, "lineup_len out of range");
      if (t.bullpen_len[s] > MLBSIM_MAX_BULLPEN) return fail(ctx, MLBSIM_ERR_TABLE, "bullpen_len out of range");
    }
  }
  CK(ctx, cudaSetDevice(ctx->device));
  if (int rc = ensure_table_capacity(ctx, n_matchups)) return rc;
  ctx->n_tables = n_matchups;
  const size_t bytes = sizeof(mlbsim_table) * (size_t)n_matchups;
  const bool staged = bytes <= UPLOAD_STAGE_BYTES;
  const void* src = tables;
  if (staged) {
    if (!ctx->upload_stage) {
      CK(ctx, cudaMallocHost(&ctx->upload_stage, UPLOAD_STAGE_BYTES));
      CK(ctx, cudaEventCreateWithFlags(&ctx->upload_ev, cudaEventDisableTiming));
    }
    if (ctx->upload_pending) CK(ctx, cudaEventSynchronize(ctx->upload_ev));   // the previous upload has left the staging buffer
    std::memcpy(ctx->upload_stage, tables, bytes);
    src = ctx->upload_stage;
  }
  CK(ctx, cudaMemcpyAsync(ctx->tables_dev, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  mlbsim_stage_tables_kernel<<<n_matchups, 256, 0, ctx->stream>>>(ctx->tables_dev, ctx->staged_dev, n_matchups);
  CK(ctx, cudaGetLastError());
  ctx->launches++;
  if (staged) {
    CK(ctx, cudaEventRecord(ctx->upload_ev, ctx->stream));
    ctx->upload_pending = true;
  } else {
    CK(ctx, cudaStreamSynchronize(ctx->stream));   // large uploads are read straight from the caller's buffer
  }
  return MLBSIM_OK;
}

static int launch_native(mlbsim_ctx* ctx, int matchup_lo, int matchup_hi, uint64_t sim_lo, uint64_t sim_hi, uint64_t seed,
                         mlbsim_hist* hist_dev, uint32_t flags) {
  if (!ctx->tables_dev) return fail(ctx, MLBSIM_ERR_STATE, "mlbsim_run_native before mlbsim_upload_tables");
  if (matchup_lo < 0 || matchup_hi > ctx->n_tables || matchup_lo >= matchup_hi) return fail(ctx, MLBSIM_ERR_ARG, "matchup range out of bounds");
  if (sim_hi < sim_lo) return fail(ctx, MLBSIM_ERR_ARG, "sim_hi < sim_lo");
  if (sim_hi > RNG_MAX_SIM_INDEX) return fail(ctx, MLBSIM_ERR_ARG, "sim index beyond 2^38 (Philox counter field)");
  if ((unsigned)(ctx->matchup_base + matchup_hi) > RNG_MAX_MATCHUPS) return fail(ctx, MLBSIM_ERR_ARG, "matchup id beyond 4096 (Philox counter field)");
  if (!hist_dev) return fail(ctx, MLBSIM_ERR_ARG, "hist pointer is NULL");
  const int n_m = matchup_hi - matchup_lo;
  const int threads = ctx->threads_per_block ? ctx->threads_per_block : MLBSIM_NATIVE_MAX_THREADS;
  const size_t smem = mlbsim_native_smem_bytes(threads);
  if (ctx->occ_threads != threads) {   // one occupancy query per block size, not one per launch
    int occ = 0;
    CK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, mlbsim_native_kernel, threads, smem));
    ctx->occ_threads = threads;
    ctx->occ_blocks = occ;
  }
  const int occ = ctx->occ_blocks;
  if (occ < 1) return fail(ctx, MLBSIM_ERR_CUDA, "native kernel does not fit on an SM");
  int bps = ctx->blocks_per_sm ? ctx->blocks_per_sm : occ;
  if (bps > occ) bps = occ;
  const int full_grid = ctx->prop.multiProcessorCount * bps;

  ctx->timed = false;
  bool first = true;
  // a launch covers < 2^32 sims per matchup; longer ranges are split (results are additive)
  const uint64_t MAX_PER_LAUNCH = 0xFFFF0000ull;
  uint64_t lo = sim_lo;
  do {
    const uint64_t n = (sim_hi - lo) < MAX_PER_LAUNCH ? (sim_hi - lo) : MAX_PER_LAUNCH;
    // grid sized to the work: a small call (the reference's default is 10 000 sims) does not pay table staging, LUT
    // construction and a 20 KB histogram flush in every one of 296 blocks for a handful of games each
    const uint64_t work_blocks = (n * (uint64_t)n_m + (uint64_t)threads - 1) / (uint64_t)threads;
    const int grid = (int)(work_blocks < (uint64_t)full_grid ? (work_blocks ? work_blocks : 1) : (uint64_t)full_grid);
    const uint64_t total_warps = (uint64_t)grid * (threads / 32);
    NativeArgs a;
    a.tables = ctx->staged_dev + matchup_lo;
    a.hist = reinterpret_cast<unsigned long long*>(hist_dev);
    a.counters = ctx->counters_dev;
    a.sim_lo = lo;
    a.n_sims = (uint32_t)n;
    a.n_matchups = (uint32_t)n_m;
    a.matchup_base = (uint32_t)(ctx->matchup_base + matchup_lo);
    // chunk: ~NATIVE_FETCHES_PER_WARP fetches per warp when all warps share one matchup, bounded to [32, 1024],
    // multiple of 32.  Smaller chunks shorten the tail (warps that run dry wait at the block barrier).
    uint64_t chunk = (n * (uint64_t)n_m) / (total_warps * NATIVE_FETCHES_PER_WARP + 1);
    chunk = (chunk / 32) * 32;
    if (chunk < 32) chunk = 32;
    if (chunk > 1024) chunk = 1024;
    a.chunk = (uint32_t)chunk;
    a.flags = flags;
    a.keys = philox2_keys_from_seed(seed);
    CK(ctx, cudaMemsetAsync(ctx->counters_dev, 0, sizeof(unsigned long long) * (size_t)n_m, ctx->stream));
    if (first) CK(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    if (n > 0) {
      mlbsim_native_kernel<<<grid, threads, smem, ctx->stream>>>(a);
      CK(ctx, cudaGetLastError());
      ctx->launches++;
    }
    first = false;
    lo += n;
  } while (lo < sim_hi);
  CK(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
  ctx->timed = true;
  return MLBSIM_OK;
}

int mlbsim_run_native(mlbsim_ctx* ctx, int matchup_lo, int matchup_hi, uint64_t sim_lo, uint64_t sim_hi, uint64_t seed,
                      mlbsim_hist* hist_dev, uint32_t flags) {
  if (!ctx) return MLBSIM_ERR_ARG;
  CK(ctx, cudaSetDevice(ctx->device));
  return launch_native(ctx, matchup_lo, matchup_hi, sim_lo, sim_hi, seed, hist_dev, flags);
}

int mlbsim_run_native_host(mlbsim_ctx* ctx, int matchup_lo, int matchup_hi, uint64_t sim_lo, uint64_t sim_hi, uint64_t seed,
                           mlbsim_hist* hist_host, uint32_t flags) {
  if (!ctx) return MLBSIM_ERR_ARG;
  if (!hist_host) return fail(ctx, MLBSIM_ERR_ARG, "hist pointer is NULL");
  CK(ctx, cudaSetDevice(ctx->device));
  const int n_m = matchup_hi - matchup_lo;
  if (n_m <= 0) return fail(ctx, MLBSIM_ERR_ARG, "matchup range out of bounds");
  if (n_m > ctx->hist_scratch_cap) {
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->hist_scratch);
    ctx->hist_scratch = nullptr;
    CK(ctx, cudaMalloc(&ctx->hist_scratch, sizeof(mlbsim_hist) * (size_t)n_m));
    ctx->hist_scratch_cap = n_m;
  }
  CK(ctx, cudaMemsetAsync(ctx->hist_scratch, 0, sizeof(mlbsim_hist) * (size_t)n_m, ctx->stream));
  int rc = launch_native(ctx, matchup_lo, matchup_hi, sim_lo, sim_hi, seed, ctx->hist_scratch, flags);
  if (rc != MLBSIM_OK) return rc;
  CK(ctx, cudaMemcpyAsync(hist_host, ctx->hist_scratch, sizeof(mlbsim_hist) * (size_t)n_m, cudaMemcpyDeviceToHost, ctx->stream));
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  return MLBSIM_OK;
}

int mlbsim_summarize(mlbsim_ctx* ctx, const mlbsim_hist* hist_dev, int n_matchups, mlbsim_summary* summary_dev) {
  if (!ctx) return MLBSIM_ERR_ARG;
  if (!hist_dev || !summary_dev || n_matchups < 0) return fail(ctx, MLBSIM_ERR_ARG, "mlbsim_summarize: bad argument");
  if (n_matchups == 0) return MLBSIM_OK;
  CK(ctx, cudaSetDevice(ctx->device));
  mlbsim_summary_kernel<<<n_matchups, MLBSIM_SUMMARY_THREADS, 0, ctx->stream>>>(hist_dev, summary_dev);
  CK(ctx, cudaGetLastError());
  ctx->launches++;
  return MLBSIM_OK;
}

int mlbsim_run_native_summary_host(mlbsim_ctx* ctx, int matchup_lo, int matchup_hi, uint64_t sim_lo, uint64_t sim_hi,
                                   uint64_t seed, mlbsim_summary* summary_host, uint32_t flags) {
  if (!ctx) return MLBSIM_ERR_ARG;
  if (!summary_host) return fail(ctx, MLBSIM_ERR_ARG, "summary pointer is NULL");
  CK(ctx, cudaSetDevice(ctx->device));
  const int n_m = matchup_hi - matchup_lo;
  if (n_m <= 0) return fail(ctx, MLBSIM_ERR_ARG, "matchup range out of bounds");
  if (n_m > ctx->hist_scratch_cap) {
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->hist_scratch);
    ctx->hist_scratch = nullptr;
    CK(ctx, cudaMalloc(&ctx->hist_scratch, sizeof(mlbsim_hist) * (size_t)n_m));
    ctx->hist_scratch_cap = n_m;
  }
  if (n_m > ctx->summary_scratch_cap) {
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->summary_scratch);
    ctx->summary_scratch = nullptr;
    CK(ctx, cudaMalloc(&ctx->summary_scratch, sizeof(mlbsim_summary) * (size_t)n_m));
    ctx->summary_scratch_cap = n_m;
  }
  CK(ctx, cudaMemsetAsync(ctx->hist_scratch, 0, sizeof(mlbsim_hist) * (size_t)n_m, ctx->stream));
  int rc = launch_native(ctx, matchup_lo, matchup_hi, sim_lo, sim_hi, seed, ctx->hist_scratch, flags);
  if (rc != MLBSIM_OK) return rc;
  rc = mlbsim_summarize(ctx, ctx->hist_scratch, n_m, ctx->summary_scratch);
  if (rc != MLBSIM_OK) return rc;
  CK(ctx, cudaMemcpyAsync(summary_host, ctx->summary_scratch, sizeof(mlbsim_summary) * (size_t)n_m, cudaMemcpyDeviceToHost, ctx->stream));
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  return MLBSIM_OK;
}

// ncclAllReduce(sendbuff, recvbuff, count, datatype, op, comm, stream); ncclUint64 = 5, ncclSum = 0 (nccl.h)
typedef int (*nccl_allreduce_fn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef const char* (*nccl_errstr_fn)(int);

int mlbsim_allreduce(mlbsim_ctx* ctx, void* nccl_comm, mlbsim_hist* hist_dev, int n_matchups) {
  if (!ctx) return MLBSIM_ERR_ARG;
  if (!nccl_comm || !hist_dev || n_matchups <= 0) return fail(ctx, MLBSIM_ERR_ARG, "mlbsim_allreduce: bad argument");
  // resolved once per process, whichever context (host thread) gets here first
  static std::once_flag once;
  static nccl_allreduce_fn allreduce = nullptr;
  static nccl_errstr_fn errstr = nullptr;
  static std::string load_error;
  std::call_once(once, [] {
    const char* path = std::getenv("MLBSIM_NCCL_LIB");
    void* h = dlopen(path && *path ? path : "libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
      load_error = std::string("mlbsim_allreduce: cannot load NCCL (") + dlerror() + "); set MLBSIM_NCCL_LIB";
      return;
    }
    allreduce = reinterpret_cast<nccl_allreduce_fn>(dlsym(h, "ncclAllReduce"));
    errstr = reinterpret_cast<nccl_errstr_fn>(dlsym(h, "ncclGetErrorString"));
    if (!allreduce) load_error = "mlbsim_allreduce: ncclAllReduce not found in the NCCL library";
  });
  if (!allreduce) return fail(ctx, MLBSIM_ERR_STATE, load_error.c_str());
  CK(ctx, cudaSetDevice(ctx->device));
  const int rc = allreduce(hist_dev, hist_dev, (size_t)n_matchups * MLBSIM_HIST_WORDS, /*ncclUint64*/ 5, /*ncclSum*/ 0, nccl_comm, ctx->stream);
  if (rc != 0) {
    ctx->err = std::string("ncclAllReduce: ") + (errstr ? errstr(rc) : "error ") + " (" + std::to_string(rc) + ")";
    return MLBSIM_ERR_CUDA;
  }
  return MLBSIM_OK;
}

int mlbsim_alias_rows(mlbsim_ctx* ctx, const double* probs_host, int n_rows, int n_own, uint32_t* out_host) {
  if (!ctx) return MLBSIM_ERR_ARG;
  if (!probs_host || !out_host || n_rows < 0 || (n_own != 7 && n_own != 8)) return fail(ctx, MLBSIM_ERR_ARG, "mlbsim_alias_rows: bad argument");
  if (n_rows == 0) return MLBSIM_OK;
  CK(ctx, cudaSetDevice(ctx->device));
  double* p_dev = nullptr;
  uint32_t* o_dev = nullptr;
  cudaError_t e;
  if ((e = cudaMalloc(&p_dev, sizeof(double) * (size_t)n_rows * n_own)) != cudaSuccess ||
      (e = cudaMalloc(&o_dev, sizeof(uint32_t) * (size_t)n_rows * MLBSIM_ROW_WORDS)) != cudaSuccess) {
    cudaFree(p_dev);
    ctx->err = std::string("cudaMalloc: ") + cudaGetErrorString(e);
    return MLBSIM_ERR_CUDA;
  }
  if ((e = cudaMemcpyAsync(p_dev, probs_host, sizeof(double) * (size_t)n_rows * n_own, cudaMemcpyHostToDevice, ctx->stream)) == cudaSuccess) {
    mlbsim_alias_rows_kernel<<<(n_rows + 127) / 128, 128, 0, ctx->stream>>>(p_dev, n_rows, n_own, o_dev);
    ctx->launches++;
    if ((e = cudaGetLastError()) == cudaSuccess &&
        (e = cudaMemcpyAsync(out_host, o_dev, sizeof(uint32_t) * (size_t)n_rows * MLBSIM_ROW_WORDS, cudaMemcpyDeviceToHost, ctx->stream)) == cudaSuccess)
      e = cudaStreamSynchronize(ctx->stream);
  }
  cudaFree(p_dev);
  cudaFree(o_dev);
  if (e != cudaSuccess) {
    ctx->err = std::string("mlbsim_alias_rows: ") + cudaGetErrorString(e);
    return MLBSIM_ERR_CUDA;
  }
  return MLBSIM_OK;
}

// Gauss-Legendre nodes and weights on (0, 1): Newton iteration on P_n (the table builder's quadrature)
static void gauss_legendre_01(int n, double* y, double* w) {
  const double pi = 3.14159265358979323846;
  for (int i = 0; i < n; ++i) {
    double x = std::cos(pi * (i + 0.75) / (n + 0.5)), dp = 1.0;
    for (int it = 0; it < 100; ++it) {
      double p0 = 1.0, p1 = x;
      for (int k = 2; k <= n; ++k) {
        const double p2 = ((2.0 * k - 1.0) * x * p1 - (k - 1.0) * p0) / k;
        p0 = p1;
        p1 = p2;
      }
      dp = n * (x * p1 - p0) / (x * x - 1.0);
      const double dx = p1 / dp;
      x -= dx;
      if (std::fabs(dx) < 1e-16) break;
    }
    y[n - 1 - i] = 0.5 * (x + 1.0);
    w[n - 1 - i] = 1.0 / ((1.0 - x * x) * dp * dp);
  }
}

int mlbsim_build_tables(mlbsim_ctx* ctx, const mlbsim_params* params_host, int n_matchups, uint32_t flags) {
  if (!ctx) return MLBSIM_ERR_ARG;
  if (!params_host || n_matchups <= 0) return fail(ctx, MLBSIM_ERR_ARG, "mlbsim_build_tables: no matchups");
  for (int i = 0; i < n_matchups; ++i) {
    const mlbsim_params& p = params_host[i];
    for (int s = 0; s < 2; ++s) {
      if (p.lineup_len[s] < 1 || p.lineup_len[s] > MLBSIM_MAX_LINEUP) return fail(ctx, MLBSIM_ERR_TABLE, "lineup_len out of range");
      if (p.bullpen_len[s] < 0 || p.bullpen_len[s] > MLBSIM_MAX_BULLPEN) return fail(ctx, MLBSIM_ERR_TABLE, "bullpen_len out of range");
    }
  }
  CK(ctx, cudaSetDevice(ctx->device));
  if (int rc = ensure_table_capacity(ctx, n_matchups)) return rc;
  if (!ctx->nodes_set) {
    double y[48], w[48];
    gauss_legendre_01(48, y, w);
    CK(ctx, mlbsim_table_set_nodes(y, w, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));   // y, w live on this stack frame
    ctx->nodes_set = true;
  }
  mlbsim_params* p_dev = nullptr;
  uint32_t* st_dev = nullptr;
  std::vector<uint32_t> status((size_t)n_matchups);
  cudaError_t e;
  if ((e = cudaMalloc(&p_dev, sizeof(mlbsim_params) * (size_t)n_matchups)) != cudaSuccess ||
      (e = cudaMalloc(&st_dev, sizeof(uint32_t) * (size_t)n_matchups)) != cudaSuccess) {
    cudaFree(p_dev);
    ctx->err = std::string("cudaMalloc: ") + cudaGetErrorString(e);
    return MLBSIM_ERR_CUDA;
  }
  ctx->n_tables = n_matchups;
  if ((e = cudaMemcpyAsync(p_dev, params_host, sizeof(mlbsim_params) * (size_t)n_matchups, cudaMemcpyHostToDevice, ctx->stream)) == cudaSuccess) {
    cudaMemsetAsync(st_dev, 0, sizeof(uint32_t) * (size_t)n_matchups, ctx->stream);
    mlbsim_build_tables_kernel<<<n_matchups * MLBSIM_TABLE_BUILD_SPLIT, 256, 0, ctx->stream>>>(p_dev, ctx->tables_dev, st_dev, n_matchups);
    mlbsim_stage_tables_kernel<<<n_matchups, 256, 0, ctx->stream>>>(ctx->tables_dev, ctx->staged_dev, n_matchups);
    ctx->launches += 2;
    if ((e = cudaGetLastError()) == cudaSuccess &&
        (e = cudaMemcpyAsync(status.data(), st_dev, sizeof(uint32_t) * (size_t)n_matchups, cudaMemcpyDeviceToHost, ctx->stream)) == cudaSuccess)
      e = cudaStreamSynchronize(ctx->stream);
  }
  cudaFree(p_dev);
  cudaFree(st_dev);
  if (e != cudaSuccess) {
    ctx->n_tables = 0;
    ctx->err = std::string("mlbsim_build_tables: ") + cudaGetErrorString(e);
    return MLBSIM_ERR_CUDA;
  }
  if (!(flags & MLBSIM_BUILD_NO_VALIDATE)) {
    std::string bad;
    int n_bad = 0;
    for (int i = 0; i < n_matchups; ++i)
      if (!status[(size_t)i]) {
        if (n_bad < 8) bad += (n_bad ? ", " : "") + std::to_string(i);
        ++n_bad;
      }
    if (n_bad) {
      ctx->n_tables = 0;   // nothing is installed
      ctx->err = "matchup(s) [" + bad + "]: no plate appearance of either side can put a runner on base, the game would never end";
      return MLBSIM_ERR_CANNOT_SCORE;
    }
  }
  return MLBSIM_OK;
}

int mlbsim_download_tables(mlbsim_ctx* ctx, int matchup_lo, int matchup_hi, mlbsim_table* out_host) {
  if (!ctx) return MLBSIM_ERR_ARG;
  if (!ctx->tables_dev || ctx->n_tables == 0) return fail(ctx, MLBSIM_ERR_STATE, "mlbsim_download_tables before tables were installed");
  if (!out_host || matchup_lo < 0 || matchup_hi > ctx->n_tables || matchup_lo > matchup_hi) return fail(ctx, MLBSIM_ERR_ARG, "mlbsim_download_tables: bad argument");
  if (matchup_lo == matchup_hi) return MLBSIM_OK;
  CK(ctx, cudaSetDevice(ctx->device));
  CK(ctx, cudaMemcpyAsync(out_host, ctx->tables_dev + matchup_lo, sizeof(mlbsim_table) * (size_t)(matchup_hi - matchup_lo), cudaMemcpyDeviceToHost, ctx->stream));
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  return MLBSIM_OK;
}

int mlbsim_last_kernel_ms(mlbsim_ctx* ctx, float* ms) {
  if (!ctx || !ms) return MLBSIM_ERR_ARG;
  if (!ctx->timed) return fail(ctx, MLBSIM_ERR_STATE, "no timed launch yet");
  CK(ctx, cudaSetDevice(ctx->device));
  CK(ctx, cudaEventSynchronize(ctx->ev1));
  CK(ctx, cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
  return MLBSIM_OK;
}

// Replay-kernel dispatch: true when every plate appearance of every game these parameters can produce draws both Beta values
// and reads its K / BB rates from the kernel's tables — noise on, a bullpen on both sides (nobody then passes 120 pitches:
// game_simulator.py:8-12 at every half-inning boundary, at most 30 plate appearances per half) and every rate
// (core/fatigue_modeling.py:25-43, core/pa_simulator.py:108-114) strictly inside (0, 1) from 0 to 138 pitches, the range the
// tables cover.  The rates are linear in the pitch count past 75, so the two ends decide; the margin makes the answer
// independent of the last bit of this host arithmetic.  Anything else runs the general kernel.
static bool replay_params_regular(const mlbsim_params* P) {
  if (!P->use_noise || P->bullpen_len[0] <= 0 || P->bullpen_len[1] <= 0) return false;
  static const double pen[MLBSIM_N_TIERS] = {0.0, 0.015, 0.035, 0.060};
  const double lo = 1e-9, hi = 1.0 - 1e-9;
  for (int side = 0; side < 2; ++side)
    for (int pit = 0; pit <= P->bullpen_len[side]; ++pit)
      for (int tier = 0; tier < MLBSIM_N_TIERS; ++tier)
        for (int slot = 0; slot < P->lineup_len[side]; ++slot)
          for (int end = 0; end < 2; ++end) {
            const double fl = end ? (138.0 - 75.0) / 25.0 : 0.0;
            const double ak = P->pit_k[side][pit] * (1.0 - pen[tier]) * (1.0 - 0.02 * fl);
            const double ab = P->pit_bb[side][pit] * (1.0 + pen[tier]) * (1.0 + 0.03 * fl);
            const double k = (P->bat_k[side][slot] + ak) / 2.0 * P->k_mod, bb = (P->bat_bb[side][slot] + ab) / 2.0 * P->bb_mod;
            if (!(k > lo && k < hi && bb > lo && bb < hi)) return false;
          }
  return true;
}

int mlbsim_replay_kernel_kind(const mlbsim_params* params) {
  if (!params) return -MLBSIM_ERR_ARG;
  if (!params->use_noise) return MLBSIM_REPLAY_KIND_NONOISE;
  for (int s = 0; s < 2; ++s) {
    if (params->lineup_len[s] < 1 || params->lineup_len[s] > MLBSIM_MAX_LINEUP) return -MLBSIM_ERR_ARG;
    if (params->bullpen_len[s] < 0 || params->bullpen_len[s] > MLBSIM_MAX_BULLPEN) return -MLBSIM_ERR_ARG;
  }
  return replay_params_regular(params) ? MLBSIM_REPLAY_KIND_REGULAR : MLBSIM_REPLAY_KIND_GENERAL;
}

int mlbsim_run_replay_dev(mlbsim_ctx* ctx, const mlbsim_params* params_host, const double* tape_dev,
                          const uint64_t* tape_offsets_dev, int64_t n_games, mlbsim_replay_game* out_dev) {
  if (!ctx) return MLBSIM_ERR_ARG;
  if (!params_host || !tape_offsets_dev || !out_dev || n_games < 0) return fail(ctx, MLBSIM_ERR_ARG, "mlbsim_run_replay: bad argument");
  for (int s = 0; s < 2; ++s) {
    if (params_host->lineup_len[s] < 1 || params_host->lineup_len[s] > MLBSIM_MAX_LINEUP) return fail(ctx, MLBSIM_ERR_ARG, "lineup_len out of range");
    if (params_host->bullpen_len[s] < 0 || params_host->bullpen_len[s] > MLBSIM_MAX_BULLPEN) return fail(ctx, MLBSIM_ERR_ARG, "bullpen_len out of range");
  }
  if (n_games == 0) return MLBSIM_OK;
  if (n_games >= (1ll << 32)) return fail(ctx, MLBSIM_ERR_ARG, "more than 2^32 - 1 games per replay call");
  CK(ctx, cudaSetDevice(ctx->device));
  // The caller may reuse params_host as soon as this returns, whatever kind of memory it is: copy it into the
  // context first.  The context's copy is pageable, so cudaMemcpyAsync returns only after it has been staged for DMA
  // and a second call cannot overwrite it under a copy still in flight.
  std::memcpy(&ctx->params_stage, params_host, sizeof(mlbsim_params));
  CK(ctx, cudaMemcpyAsync(ctx->params_dev, &ctx->params_stage, sizeof(mlbsim_params), cudaMemcpyHostToDevice, ctx->stream));
  ReplayArgs a{ctx->params_dev, tape_dev, tape_offsets_dev, out_dev, (long long)n_games, {}, {}, {}, {1.01, 0.10, 0.4, 0.14, 0.8}};
  std::memcpy(a.cdf_bip, params_host->cdf_bip, sizeof(a.cdf_bip));
  std::memcpy(a.cdf_hit, params_host->cdf_hit, sizeof(a.cdf_hit));
  std::memcpy(a.cdf_fb, params_host->cdf_fb, sizeof(a.cdf_fb));
  if (ctx->counters_cap < 1) {
    CK(ctx, cudaMalloc(&ctx->counters_dev, sizeof(unsigned long long)));
    ctx->counters_cap = 1;
  }
  CK(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
  CK(ctx, cudaMemsetAsync(out_dev, 0, sizeof(mlbsim_replay_game) * (size_t)n_games, ctx->stream));  // kernels write non-zero fields only
  {
    const int threads = MLBSIM_REPLAY_FLAT_THREADS;
    const int grid = ctx->prop.multiProcessorCount * MLBSIM_REPLAY_FLAT_BLOCKS;
    const long long warps = (long long)grid * (threads / 32);
    long long chunk = n_games / (warps * 4 + 1);
    chunk = (chunk / 32) * 32;
    if (chunk < 32) chunk = 32;
    if (chunk > 256) chunk = 256;
    CK(ctx, cudaMemsetAsync(ctx->counters_dev, 0, sizeof(unsigned long long), ctx->stream));
    const size_t dyn = MLBSIM_REPLAY_DYN_SMEM;
    if (dyn > 0 && !ctx->replay_attr_set) {   // the per-lane tape rings: more than the 48 KB a kernel gets without asking
      CK(ctx, cudaFuncSetAttribute(mlbsim_replay_flat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
      CK(ctx, cudaFuncSetAttribute(mlbsim_replay_flat_nonoise_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
      CK(ctx, cudaFuncSetAttribute(mlbsim_replay_flat_regular_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
      ctx->replay_attr_set = true;
    }
    // MLBSIM_REPLAY_GENERAL=1 (environment; tests and A/B): never pick the specialised instance
    const char* force_general = std::getenv("MLBSIM_REPLAY_GENERAL");
    const int kind = mlbsim_replay_kernel_kind(params_host);   // (arguments were range-checked above)
    if (kind == MLBSIM_REPLAY_KIND_NONOISE) mlbsim_replay_flat_nonoise_kernel<<<grid, threads, dyn, ctx->stream>>>(a, ctx->counters_dev, (uint32_t)chunk);
    else if (kind == MLBSIM_REPLAY_KIND_REGULAR && !(force_general && force_general[0] == '1'))
      mlbsim_replay_flat_regular_kernel<<<grid, threads, dyn, ctx->stream>>>(a, ctx->counters_dev, (uint32_t)chunk);
    else mlbsim_replay_flat_kernel<<<grid, threads, dyn, ctx->stream>>>(a, ctx->counters_dev, (uint32_t)chunk);
  }
  CK(ctx, cudaGetLastError());
  CK(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
  ctx->timed = true;
  ctx->launches++;
  return MLBSIM_OK;
}

int mlbsim_run_replay(mlbsim_ctx* ctx, const mlbsim_params* params, const double* tape, const uint64_t* tape_offsets,
                      int64_t n_games, mlbsim_replay_game* out) {
  if (!ctx) return MLBSIM_ERR_ARG;
  if (!params || !tape_offsets || !out || n_games < 0) return fail(ctx, MLBSIM_ERR_ARG, "mlbsim_run_replay: bad argument");
  if (n_games == 0) return MLBSIM_OK;
  CK(ctx, cudaSetDevice(ctx->device));
  const uint64_t n_draws = tape_offsets[n_games];
  if (n_draws && !tape) return fail(ctx, MLBSIM_ERR_ARG, "tape is NULL");
  double* tape_dev = nullptr;
  uint64_t* off_dev = nullptr;
  mlbsim_replay_game* out_dev = nullptr;
  int rc = MLBSIM_OK;
  cudaError_t e;
  // (rounded up to whole 128-byte chunks + one: the kernel loads aligned 16-byte units of whole windows)
  if ((e = cudaMalloc(&tape_dev, sizeof(double) * (size_t)(((n_draws + 15) / 16) * 16 + 16))) != cudaSuccess ||
      (e = cudaMalloc(&off_dev, sizeof(uint64_t) * (size_t)(n_games + 1))) != cudaSuccess ||
      (e = cudaMalloc(&out_dev, sizeof(mlbsim_replay_game) * (size_t)n_games)) != cudaSuccess) {
    ctx->err = std::string("cudaMalloc: ") + cudaGetErrorString(e);
    rc = MLBSIM_ERR_CUDA;
  }
  if (rc == MLBSIM_OK) {
    if ((e = cudaMemcpyAsync(tape_dev, tape, sizeof(double) * (size_t)n_draws, cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess ||
        (e = cudaMemcpyAsync(off_dev, tape_offsets, sizeof(uint64_t) * (size_t)(n_games + 1), cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess) {
      ctx->err = std::string("cudaMemcpyAsync: ") + cudaGetErrorString(e);
      rc = MLBSIM_ERR_CUDA;
    }
  }
  if (rc == MLBSIM_OK) rc = mlbsim_run_replay_dev(ctx, params, tape_dev, off_dev, n_games, out_dev);
  if (rc == MLBSIM_OK) {
    if ((e = cudaMemcpyAsync(out, out_dev, sizeof(mlbsim_replay_game) * (size_t)n_games, cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess ||
        (e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) {
      ctx->err = std::string("replay copy-back: ") + cudaGetErrorString(e);
      rc = MLBSIM_ERR_CUDA;
    }
  } else {
    cudaStreamSynchronize(ctx->stream);
  }
  cudaFree(tape_dev);
  cudaFree(off_dev);
  cudaFree(out_dev);
  return rc;
}

int mlbsim_run_persim(mlbsim_ctx* ctx, int matchup, uint64_t sim_lo, uint64_t sim_hi, uint64_t seed, mlbsim_replay_game* out_host) {
  if (!ctx) return MLBSIM_ERR_ARG;
  if (!ctx->tables_dev) return fail(ctx, MLBSIM_ERR_STATE, "mlbsim_run_persim before mlbsim_upload_tables");
  if (matchup < 0 || matchup >= ctx->n_tables) return fail(ctx, MLBSIM_ERR_ARG, "matchup out of range");
  if (sim_hi < sim_lo || !out_host) return fail(ctx, MLBSIM_ERR_ARG, "mlbsim_run_persim: bad argument");
  if (sim_hi > RNG_MAX_SIM_INDEX || (unsigned)(ctx->matchup_base + matchup) >= RNG_MAX_MATCHUPS) return fail(ctx, MLBSIM_ERR_ARG, "sim or matchup index beyond the Philox counter fields");
  const uint64_t n = sim_hi - sim_lo;
  if (n == 0) return MLBSIM_OK;
  CK(ctx, cudaSetDevice(ctx->device));
  // bounded device scratch: records leave in slices of at most 4 Mi sims (704 MB), whatever n is
  const uint64_t SLICE = 4ull << 20;
  const uint64_t cap_n = n < SLICE ? n : SLICE;
  mlbsim_replay_game* out_dev = nullptr;
  CK(ctx, cudaMalloc(&out_dev, sizeof(mlbsim_replay_game) * (size_t)cap_n));
  cudaError_t e = cudaSuccess;
  for (uint64_t done = 0; done < n && e == cudaSuccess; done += cap_n) {
    const uint64_t m = (n - done) < cap_n ? (n - done) : cap_n;
    PersimArgs a;
    a.table = ctx->tables_dev + matchup;
    a.out = out_dev;
    a.sim_lo = sim_lo + done;
    a.n_sims = m;
    a.matchup_id = (uint32_t)(ctx->matchup_base + matchup);
    a.keys = philox2_keys_from_seed(seed);
    const int threads = 128;
    uint64_t blocks = (m + threads - 1) / threads;
    const uint64_t cap = (uint64_t)ctx->prop.multiProcessorCount * 16;
    if (blocks > cap) blocks = cap;
    mlbsim_persim_kernel<<<(int)blocks, threads, 0, ctx->stream>>>(a);
    ctx->launches++;
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_host + done, out_dev, sizeof(mlbsim_replay_game) * (size_t)m, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  }
  cudaFree(out_dev);
  if (e != cudaSuccess) {
    ctx->err = std::string("mlbsim_run_persim: ") + cudaGetErrorString(e);
    return MLBSIM_ERR_CUDA;
  }
  return MLBSIM_OK;
}

int mlbsim_malloc(mlbsim_ctx* ctx, size_t bytes, void** dev_ptr) {
  if (!ctx || !dev_ptr) return MLBSIM_ERR_ARG;
  CK(ctx, cudaSetDevice(ctx->device));
  CK(ctx, cudaMalloc(dev_ptr, bytes + 256));   // slack: a tape buffer from here is readable to the end of its last 128-byte chunk
  return MLBSIM_OK;
}

int mlbsim_free(mlbsim_ctx* ctx, void* dev_ptr) {
  if (!ctx) return MLBSIM_ERR_ARG;
  CK(ctx, cudaSetDevice(ctx->device));
  CK(ctx, cudaFree(dev_ptr));
  return MLBSIM_OK;
}

int mlbsim_memset(mlbsim_ctx* ctx, void* dev_ptr, int value, size_t bytes) {
  if (!ctx || !dev_ptr) return MLBSIM_ERR_ARG;
  CK(ctx, cudaSetDevice(ctx->device));
  CK(ctx, cudaMemsetAsync(dev_ptr, value, bytes, ctx->stream));
  return MLBSIM_OK;
}

int mlbsim_memcpy_h2d(mlbsim_ctx* ctx, void* dev_dst, const void* host_src, size_t bytes) {
  if (!ctx || !dev_dst || !host_src) return MLBSIM_ERR_ARG;
  CK(ctx, cudaSetDevice(ctx->device));
  CK(ctx, cudaMemcpyAsync(dev_dst, host_src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  return MLBSIM_OK;
}

int mlbsim_memcpy_d2h(mlbsim_ctx* ctx, void* host_dst, const void* dev_src, size_t bytes) {
  if (!ctx || !host_dst || !dev_src) return MLBSIM_ERR_ARG;
  CK(ctx, cudaSetDevice(ctx->device));
  CK(ctx, cudaMemcpyAsync(host_dst, dev_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  return MLBSIM_OK;
}

}  // extern "C"
