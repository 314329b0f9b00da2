// c_api.cu -- the extern "C" surface declared in include/rbg_b200.h:
// argument validation, error reporting, launch sequencing, and the
// host-buffer variants (pipelined H2D -> kernels -> D2H over env slices).
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <chrono>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "rbg_host.h"

namespace rbg {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

int set_error(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int set_cuda_error(cudaError_t e, const char *what) {
  return set_error(RBG_ECUDA, "%s: %s", what, cudaGetErrorString(e));
}

int check_launch(const char *what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, what);
  return RBG_OK;
}

// ---- device properties, cached per device ---------------------------------
struct DevProps {
  int sms = 0;
  size_t smem_per_sm = 0;
};
static DevProps g_devprops[64];
static const DevProps &dev_props() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
  DevProps &d = g_devprops[dev];
  if (d.sms == 0) {
    int v = 0;
    d.sms = (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) ? v : 148;
    d.smem_per_sm = (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev) == cudaSuccess && v > 0) ? (size_t)v : (size_t)228 * 1024;
  }
  return d;
}
int device_sm_count() { return dev_props().sms; }
size_t device_smem_per_sm() { return dev_props().smem_per_sm; }

// ---- per-kernel event timing (rbg_kernel_timing / rbg_kernel_time) --------
struct EventPair {
  cudaEvent_t a, b;
};
static std::atomic<int> g_timing{0};
static std::mutex g_timing_mu;
static std::vector<EventPair> g_events[RBG_K_COUNT];

bool kernel_timing_on() { return g_timing.load(std::memory_order_relaxed) != 0; }

LaunchScope::LaunchScope(int kernel_id, cudaStream_t s) : id(kernel_id), stream(s), rec(nullptr) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (!g_timing.load(std::memory_order_relaxed) || id < 0 || id >= RBG_K_COUNT) return;
  EventPair *ep = new EventPair;
  if (cudaEventCreate(&ep->a) != cudaSuccess || cudaEventCreate(&ep->b) != cudaSuccess) {
    delete ep;
    return;
  }
  cudaEventRecord(ep->a, stream);
  rec = ep;
}

LaunchScope::~LaunchScope() {
  if (!rec) return;
  EventPair *ep = static_cast<EventPair *>(rec);
  cudaEventRecord(ep->b, stream);
  std::lock_guard<std::mutex> lock(g_timing_mu);
  g_events[id].push_back(*ep);
  delete ep;
}

static int check_dims(int64_t B, int G, int N, int need_cells_per_agent) {
  if (B < 0) return set_error(RBG_EINVAL, "B=%lld must be >= 0", (long long)B);
  if (B > 0x3fffffffLL) return set_error(RBG_EINVAL, "B=%lld too large (max 2^30-1 per call)", (long long)B);
  if (G < 2 || G > RBG_MAX_G) return set_error(RBG_EINVAL, "grid size G=%d outside [2,%d]", G, RBG_MAX_G);
  if (N < 1 || N > RBG_MAX_N) return set_error(RBG_EINVAL, "num_agents N=%d outside [1,%d]", N, RBG_MAX_N);
  if ((int64_t)need_cells_per_agent * N > (int64_t)G * G)
    return set_error(RBG_EINVAL, "N=%d agents need %d cells on a %dx%d grid (choice(replace=False) would raise)",
                     N, need_cells_per_agent * N, G, G);
  return RBG_OK;
}

static bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static bool aligned8(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 7u) == 0; }

// 16-byte alignment of the bulk arrays is needed by the 128-bit path only, i.e. when G*G % 4 == 0 (otherwise the
// kernels use scalar accesses, and e.g. the step slices of a stacked rollout are legitimately unaligned)
static int check_state(const rbg_state *s, const char *name, int G) {
  if (!s) return set_error(RBG_EINVAL, "%s is NULL", name);
  if (!s->grid || !s->step_count || !s->agent_id || !s->start || !s->target || !s->position || !s->key)
    return set_error(RBG_EINVAL, "%s has a NULL field", name);
  if ((G * G) % 4 == 0 && !aligned16(s->grid)) return set_error(RBG_EALIGN, "%s.grid not 16-byte aligned", name);
  if (!aligned8(s->start) || !aligned8(s->target) || !aligned8(s->position))
    return set_error(RBG_EALIGN, "%s.start/target/position not 8-byte aligned", name);
  return RBG_OK;
}

static int check_timestep(const rbg_timestep *t, int G) {
  if (!t) return set_error(RBG_EINVAL, "timestep is NULL");
  if (!t->obs_grid || !t->action_mask || !t->obs_step_count || !t->reward || !t->discount || !t->step_type ||
      !t->num_connections || !t->ratio_connections || !t->total_path_length)
    return set_error(RBG_EINVAL, "timestep has a NULL field");
  if ((G * G) % 4 == 0 && !aligned16(t->obs_grid)) return set_error(RBG_EALIGN, "timestep.obs_grid not 16-byte aligned");
  return RBG_OK;
}

static int check_dataset(const rbg_env_params *params) {
  if (!params->dataset_heads || !params->dataset_targets)
    return set_error(RBG_EINVAL, "autoreset_kind RBG_GEN_DATASET needs params->dataset_heads / dataset_targets (device int32[K,2,N])");
  if (params->dataset_K < 1 || params->dataset_K > 0x7fffffffLL)
    return set_error(RBG_EINVAL, "autoreset_kind RBG_GEN_DATASET: number of boards dataset_K=%lld", (long long)params->dataset_K);
  return RBG_OK;
}

static int debug_flags() {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("RBG_DEBUG_FLAGS");
    v = e ? atoi(e) : 0;
  }
  return v;
}
static int env_int(const char *name) {
  const char *e = getenv(name);
  return e ? atoi(e) : 0;
}

static int generator_state_impl(int kind, const uint32_t *keys, int64_t B, int G, int N, const rbg_state *out,
                                const rbg_timestep *ts, int extra_split, const int32_t *list,
                                const int32_t *list_count, cudaStream_t stream, int32_t *list_ticket = nullptr) {
  if (kind == RBG_GEN_PRW || kind == RBG_GEN_UNIFORM) {
    PrwParams p;
    memset(&p, 0, sizeof(p));
    p.keys = keys;
    p.B = B;
    p.G = G;
    p.N = N;
    p.mode = kind == RBG_GEN_PRW ? PRW_MODE_STATE : PRW_MODE_UNIFORM;
    // PRWG:52 key, pos_key = split(key): the board uses split(key)[0].
    // UG:76 the uniform generator consumes the split itself (both halves).
    p.extra_split = (kind == RBG_GEN_PRW ? 1 : 0) + extra_split;
    p.debug = debug_flags();
    p.st = *out;
    if (ts) {
      p.ts = *ts;
      p.observe = 1;
    }
    p.list = list;
    p.list_count = list_count;
    p.list_ticket = list_ticket;
    return launch_prw(p, B, env_int("RBG_PRW_M"), env_int("RBG_PRW_THREADS"), stream);
  }
  if (kind == RBG_GEN_SEEDEXT) {
    SeedExtParams p;
    memset(&p, 0, sizeof(p));
    p.keys = keys;
    p.B = B;
    p.G = G;
    p.N = N;
    p.randomness = 0.0f;  // RSG:30 jit(generate_starts_ends) with the defaults
    p.two_sided = 1;
    p.iterations = 1;
    p.ext_steps = -1;
    p.extra_split = 1 + extra_split;  // RSG:34 key, pos_key = split(key)
    p.mode = 2;
    p.st = *out;
    if (ts) {
      p.ts = *ts;
      p.observe = 1;
    }
    p.list = list;
    p.list_count = list_count;
    return launch_seedext(p, B, stream);
  }
  if (kind == RBG_GEN_SEQRW) {
    SeqRwParams p;
    memset(&p, 0, sizeof(p));
    p.keys = keys;
    p.B = B;
    p.G = G;
    p.N = N;
    p.extra_split = 1 + extra_split;  // sequential_random_walk_generator.py:42 key, pos_key = split(key)
    p.mode = 2;
    p.st = *out;
    if (ts) {
      p.ts = *ts;
      p.observe = 1;
    }
    p.list = list;
    p.list_count = list_count;
    return launch_seqrw(p, B, stream);
  }
  return set_error(RBG_EINVAL, "unknown generator kind %d", kind);
}

// ---- auto-reset workspace ------------------------------------------------------
// VmapAutoResetWrapper regenerates a board whenever an env finishes.  The reset key
// of an episode is known as soon as the episode starts (split(state.key)[0]), so the
// NEXT episode of every env that just reset is generated ahead of time on a side
// stream ("refill") into a per-env cache; when the env finishes, env_kernel swaps the
// cached episode in.  Only envs whose cache entry is not ready (first episode, or two
// terminations in consecutive steps) take the synchronous reset kernel.  Results are
// identical either way: a board is a pure function of its key.
struct WsLayout {
  size_t sync_list, refill_list[2], refill_keys[2], cache_tag, cache_key, cache_pins, group_done, group_pending, total;
};
static WsLayout ws_layout(int64_t B, int N) {
  auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
  WsLayout w;
  size_t off = 512;  // counters: sync @0, refill[0] @64, refill[1] @128; persistent rollout: 16 sets of 4 ints @256
  const size_t b = (size_t)(B > 0 ? B : 0);
  w.sync_list = off;
  off = up(off + 4 * b);
  for (int i = 0; i < 2; ++i) {
    w.refill_list[i] = off;
    off = up(off + 4 * b);
    w.refill_keys[i] = off;
    off = up(off + 8 * b);
  }
  w.cache_tag = off;
  off = up(off + 8 * b);
  w.cache_key = off;
  off = up(off + 8 * b);
  w.cache_pins = off;
  off = up(off + 4 * b * (size_t)N);
  w.group_done = off;  // persistent rollout: per env group, the last launch that wrote it / its requests in flight
  off = up(off + 4 * b);
  w.group_pending = off;
  off = up(off + 4 * b);
  w.total = off;
  return w;
}

struct AutoResetCtx {
  int64_t B = 0;
  int G = 0, N = 0, kind = -1;
  uint64_t step = 0;
  int persist_epoch = 0;  // launches of the persistent rollout on this workspace
  int persist_groups = 0; // group size the flags were laid out for
  cudaStream_t last_stream = nullptr;  // the stream the workspace was last used on
  bool tail_valid = false;
  cudaStream_t side = nullptr;
  cudaEvent_t env_done = nullptr;
  cudaEvent_t refill_done[2] = {nullptr, nullptr};
  bool refill_pending[2] = {false, false};
  // fused rollout: slices 1.. of the batch run on their own streams
  cudaStream_t slice_stream[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t slice_fork = nullptr, slice_done[3] = {nullptr, nullptr, nullptr};
};
static std::mutex g_ar_mu;
static std::unordered_map<void *, AutoResetCtx> g_ar;

// The next speculative call on `workspace` must adopt it afresh (counters zeroed, cache warm-started):
// its contents are no longer what the recorded context left there.  Pending refills are joined first.
static void invalidate_workspace(void *workspace, cudaStream_t stream, bool block = false) {
  std::lock_guard<std::mutex> lock(g_ar_mu);
  auto it = g_ar.find(workspace);
  if (it == g_ar.end()) return;
  for (int i = 0; i < 2; ++i)
    if (it->second.refill_pending[i]) {
      if (block)
        cudaEventSynchronize(it->second.refill_done[i]);
      else
        cudaStreamWaitEvent(stream, it->second.refill_done[i], 0);
      it->second.refill_pending[i] = false;
    }
  it->second.B = 0;
}

// Per-step auto-reset: where the cache refill of step t runs.  Default: on the CALLER'S stream, as a kernel that says
// launch_dependents at once, followed (next call) by an env kernel launched with programmatic stream serialization, so
// the refill of step t overlaps the env kernel of step t + 1 by construction and nothing else moves (the synchronous
// reset kernel of step t + 1 is a plain launch and waits for both).  RBG_STEP_SIDE_STREAM=1: the round-1 scheme, a
// low-priority side stream joined two steps later, whose overlap depended on which hardware queue the side stream
// happened to share (65.6 or 84.5 us per step).
static bool step_side_stream() {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("RBG_STEP_SIDE_STREAM");
    v = (e && atoi(e) != 0) ? 1 : 0;
  }
  return v == 1;
}

// The workspace moves to another stream (rare): its work on the old stream is awaited on the host.  Events per call
// would do, but an event record between two launches stops the second from starting under the first's tail.
static int workspace_follow_stream(AutoResetCtx *ctx, cudaStream_t stream) {
  if (ctx->tail_valid && ctx->last_stream != stream) {
    cudaStreamSynchronize(ctx->last_stream);
    cudaGetLastError();  // (the old stream may be gone)
  }
  ctx->last_stream = stream;
  ctx->tail_valid = true;
  return RBG_OK;
}

static bool speculative_enabled() {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("RBG_NO_SPECULATIVE_RESET");
    v = (e && atoi(e) != 0) ? 0 : 1;
  }
  return v == 1;
}

static int connector_step_impl(const rbg_state *in, const rbg_state *out, const int32_t *action, int32_t *action_out,
                               int random_policy, int64_t B, int G, int N, const rbg_env_params *params,
                               const rbg_timestep *ts, void *workspace, cudaStream_t stream) {
  int rc;
  if ((rc = check_dims(B, G, N, 1))) return rc;
  if (B == 0) return RBG_OK;  // an empty batch has no buffers to check
  if ((rc = check_state(in, "state_in", G))) return rc;
  if ((rc = check_state(out, "state_out", G))) return rc;
  if ((rc = check_timestep(ts, G))) return rc;
  if (!params) return set_error(RBG_EINVAL, "params is NULL");
  if (!random_policy && !action) return set_error(RBG_EINVAL, "action is NULL");
  const bool autoreset = params->autoreset_kind >= 0;
  const bool dataset = params->autoreset_kind == RBG_GEN_DATASET;
  if (autoreset && !dataset && !workspace) return set_error(RBG_EINVAL, "auto-reset needs a workspace (rbg_step_workspace_bytes)");
  if (autoreset && params->autoreset_kind > RBG_GEN_SEQRW)
    return set_error(RBG_EINVAL, "unknown autoreset generator kind %d", params->autoreset_kind);
  if (dataset && (rc = check_dataset(params))) return rc;
  if (B == 0) return RBG_OK;
  EnvParams p;
  memset(&p, 0, sizeof(p));
  p.in = *in;
  p.out = *out;
  p.ts = *ts;
  p.action = action;
  p.action_out = action_out;
  p.random_policy = random_policy;
  p.B = B;
  p.G = G;
  p.N = N;
  p.mode = ENV_MODE_STEP;
  p.env = *params;
  if (!autoreset || dataset) return launch_env(p, stream);  // dataset resets are a lookup inside the kernel

  uint8_t *ws = reinterpret_cast<uint8_t *>(workspace);
  if (!aligned16(ws)) return set_error(RBG_EALIGN, "workspace not 16-byte aligned");
  const WsLayout wl = ws_layout(B, N);
  const int kind = params->autoreset_kind;
  const bool speculative = speculative_enabled() && (kind == RBG_GEN_PRW || kind == RBG_GEN_UNIFORM);
  int32_t *sync_count = reinterpret_cast<int32_t *>(ws);
  int32_t *sync_list = reinterpret_cast<int32_t *>(ws + wl.sync_list);
  p.list = sync_list;
  p.list_count = sync_count;
  cudaError_t e;
  AutoResetCtx *ctx = nullptr;
  int par = 0;
  if (speculative) {
    std::lock_guard<std::mutex> lock(g_ar_mu);
    ctx = &g_ar[workspace];
    if (!step_side_stream()) {
      if ((rc = workspace_follow_stream(ctx, stream))) return rc;
    } else if (!ctx->side) {
      // lowest priority: the refill has a whole step of slack, env_kernel's CTAs go first
      int prio_lo = 0, prio_hi = 0;
      cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
      if ((e = cudaStreamCreateWithPriority(&ctx->side, cudaStreamNonBlocking, prio_lo)) != cudaSuccess) return set_cuda_error(e, "cudaStreamCreate(side)");
      if ((e = cudaEventCreateWithFlags(&ctx->env_done, cudaEventDisableTiming)) != cudaSuccess) return set_cuda_error(e, "cudaEventCreate");
      for (int i = 0; i < 2; ++i)
        if ((e = cudaEventCreateWithFlags(&ctx->refill_done[i], cudaEventDisableTiming)) != cudaSuccess) return set_cuda_error(e, "cudaEventCreate");
    }
    if (ctx->B != B || ctx->G != G || ctx->N != N || ctx->kind != kind) {
      // first use of this workspace (or a different batch in it): nothing cached yet.
      // Pending refills of the previous occupant must not land after the clear.
      for (int i = 0; i < 2; ++i)
        if (ctx->refill_pending[i]) {
          cudaStreamWaitEvent(stream, ctx->refill_done[i], 0);
          ctx->refill_pending[i] = false;
        }
      if ((e = cudaMemsetAsync(ws, 0, 256, stream)) != cudaSuccess) return set_cuda_error(e, "cudaMemsetAsync(counters)");
      // warm start: the next episode of EVERY env, one bulk generation keyed by the
      // current State.key (a cold cache would send each env's first reset down the
      // synchronous path)
      PrwParams q;
      memset(&q, 0, sizeof(q));
      q.keys = in->key;
      q.B = B;
      q.G = G;
      q.N = N;
      q.mode = kind == RBG_GEN_PRW ? PRW_MODE_STATE : PRW_MODE_UNIFORM;
      q.extra_split = (kind == RBG_GEN_PRW ? 1 : 0) + 1;
      q.debug = debug_flags();
      q.to_cache = 1;
      q.cache_tag = reinterpret_cast<unsigned long long *>(ws + wl.cache_tag);
      q.cache_key = reinterpret_cast<uint2 *>(ws + wl.cache_key);
      q.cache_pins = reinterpret_cast<uint32_t *>(ws + wl.cache_pins);
      if ((rc = launch_prw(q, B, env_int("RBG_PRW_M"), env_int("RBG_PRW_THREADS"), stream))) return rc;
      ctx->B = B;
      ctx->G = G;
      ctx->N = N;
      ctx->kind = kind;
      ctx->step = 0;
      ctx->persist_epoch = 0;
    }
    par = (int)(ctx->step & 1);
    // the refill launched two steps ago used this parity's list buffers
    if (ctx->refill_pending[par]) {
      if ((e = cudaStreamWaitEvent(stream, ctx->refill_done[par], 0)) != cudaSuccess) return set_cuda_error(e, "cudaStreamWaitEvent");
      ctx->refill_pending[par] = false;
    }
    p.cache_tag = reinterpret_cast<const unsigned long long *>(ws + wl.cache_tag);
    p.cache_key = reinterpret_cast<const uint2 *>(ws + wl.cache_key);
    p.cache_pins = reinterpret_cast<const uint32_t *>(ws + wl.cache_pins);
    p.refill_list = reinterpret_cast<int32_t *>(ws + wl.refill_list[par]);
    p.refill_keys = reinterpret_cast<uint32_t *>(ws + wl.refill_keys[par]);
    p.refill_count = reinterpret_cast<int32_t *>(ws + 64 + 64 * par);
  }
  // counters: with the speculative path the list kernels clear their own counter when
  // they finish (prw_kernel list_ticket); they were zeroed once when the workspace was adopted
  if (!speculative) {
    if ((e = cudaMemsetAsync(ws, 0, 64, stream)) != cudaSuccess) return set_cuda_error(e, "cudaMemsetAsync(reset counter)");
    // This path leaves its reset count in the counter.  If the same workspace has served (or will
    // serve) a speculative batch, that batch must adopt it afresh: its kernels rely on counters
    // that are zero between steps.
    invalidate_workspace(workspace, stream);
  }
  const bool chained = speculative && !step_side_stream() && !g_timing.load(std::memory_order_relaxed);
  if ((rc = launch_env(p, stream, chained))) return rc;
  // VmapAutoResetWrapper._auto_reset: key, _ = split(state.key); reset(key):
  // one more leading split()[0] than a plain generator call.
  if (chained) {
    // one launch: the synchronous resets (rare misses), launch_dependents, then the cache refill
    PrwParams a, q;
    memset(&a, 0, sizeof(a));
    a.keys = out->key;
    a.B = B;
    a.G = G;
    a.N = N;
    a.mode = kind == RBG_GEN_PRW ? PRW_MODE_STATE : PRW_MODE_UNIFORM;
    a.extra_split = (kind == RBG_GEN_PRW ? 1 : 0) + 1;
    a.debug = debug_flags();
    a.st = *out;
    a.ts = *ts;
    a.observe = 1;
    a.list = sync_list;
    a.list_count = sync_count;
    a.list_ticket = sync_count + 1;
    memset(&q, 0, sizeof(q));
    q.keys = p.refill_keys;
    q.keys_compact = 1;
    q.B = B;
    q.G = G;
    q.N = N;
    q.mode = a.mode;
    q.extra_split = a.extra_split;
    q.debug = a.debug;
    q.list = p.refill_list;
    q.list_count = p.refill_count;
    q.list_ticket = p.refill_count + 1;
    q.to_cache = 1;
    q.cache_tag = reinterpret_cast<unsigned long long *>(ws + wl.cache_tag);
    q.cache_key = reinterpret_cast<uint2 *>(ws + wl.cache_key);
    q.cache_pins = reinterpret_cast<uint32_t *>(ws + wl.cache_pins);
    if ((rc = launch_prw_pair(a, q, B, stream))) return rc;
    ctx->step++;
    return RBG_OK;
  }
  if ((rc = generator_state_impl(kind, out->key, B, G, N, out, ts, 1, sync_list, sync_count, stream, speculative ? sync_count + 1 : nullptr))) return rc;
  if (speculative) {
    const bool side = step_side_stream();
    if (side) {
      if ((e = cudaEventRecord(ctx->env_done, stream)) != cudaSuccess) return set_cuda_error(e, "cudaEventRecord");
      if ((e = cudaStreamWaitEvent(ctx->side, ctx->env_done, 0)) != cudaSuccess) return set_cuda_error(e, "cudaStreamWaitEvent(side)");
    }
    PrwParams q;
    memset(&q, 0, sizeof(q));
    q.keys = p.refill_keys;
    q.keys_compact = 1;
    q.B = B;
    q.G = G;
    q.N = N;
    q.mode = kind == RBG_GEN_PRW ? PRW_MODE_STATE : PRW_MODE_UNIFORM;
    q.extra_split = (kind == RBG_GEN_PRW ? 1 : 0) + 1;  // as the auto-reset call above
    q.debug = debug_flags();
    q.list = p.refill_list;
    q.list_count = p.refill_count;
    q.list_ticket = p.refill_count + 1;
    q.to_cache = 1;
    q.cache_tag = reinterpret_cast<unsigned long long *>(ws + wl.cache_tag);
    q.cache_key = reinterpret_cast<uint2 *>(ws + wl.cache_key);
    q.cache_pins = reinterpret_cast<uint32_t *>(ws + wl.cache_pins);
    if (side) {
      if ((rc = launch_prw(q, B, env_int("RBG_PRW_M"), env_int("RBG_PRW_THREADS"), ctx->side))) return rc;
      if ((e = cudaEventRecord(ctx->refill_done[par], ctx->side)) != cudaSuccess) return set_cuda_error(e, "cudaEventRecord(side)");
      ctx->refill_pending[par] = true;
    } else {
      q.trigger_dependents = 1;
      if ((rc = launch_prw(q, B, env_int("RBG_PRW_M"), env_int("RBG_PRW_THREADS"), stream))) return rc;
    }
    ctx->step++;
  }
  return RBG_OK;
}

// ---- device scratch for the _host variants --------------------------------
struct Scratch {
  void *ptr = nullptr;
  size_t bytes = 0;
  int device = -1;
};
static std::mutex g_scratch_mu;
// One scratch, one set of streams and one signature PER DEVICE: a process that drives several GPUs through the
// _host variants (one thread per device, or device switches between calls) keeps each device's workspaces alive.
constexpr int kMaxDevices = 64;
// Who laid the scratch out last (a step variant with auto-reset keeps its workspaces there between
// calls): any other use, shape or base pointer means those workspaces have been overwritten.
struct ScratchSig {
  int fn = 0;
  int64_t B = 0;
  int G = 0, N = 0, kind = -1;
  void *base = nullptr;
  bool operator==(const ScratchSig &o) const { return fn == o.fn && B == o.B && G == o.G && N == o.N && kind == o.kind && base == o.base; }
};
struct HostCtx {
  Scratch scratch;
  cudaStream_t streams[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t stream_ev[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t slice_ev[16] = {nullptr};
  cudaEvent_t copy_ev[16] = {nullptr};
  cudaEvent_t entry_ev = nullptr;
  uint8_t *stage8 = nullptr;  // pinned host staging of the byte observation (rbg_connector_step_host_io)
  size_t stage8_bytes = 0;
  // thread-count tuning of the packed transport: the first calls of a batch shape try 1/4, 1/2, 3/4 and all of the
  // pool's workers (four calls each, the last three timed; more threads must win by 1 %) and the fastest count stays
  int64_t tune_B = -1;
  int tune_call = 0, tune_best = 0;
  double tune_best_s = 0.0;
  // mixed transport: how many of the step's slices cross the bus as int32 straight into the caller's (pinned) buffer
  // instead of as bytes for the host threads to widen; moved by one slice per call towards the balance of bus and host
  int wide_slices = 0;
  int64_t wide_B = -1;
  ScratchSig sig;
};
// what the _host variants moved over the bus since the last reset (rbg_host_transfer_stats)
static std::atomic<long long> g_h2d_bytes{0}, g_d2h_bytes{0};
static bool host_io_wide() {  // RBG_HOST_IO_WIDE=1: the observation crosses the bus as int32 (round 1's transport)
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("RBG_HOST_IO_WIDE");
    v = e && atoi(e) != 0 ? 1 : 0;
  }
  return v == 1;
}
static HostCtx g_host[kMaxDevices];
// RBG_HOST_IO_SPLIT=k: k of a step's slices cross the bus as int32 straight into the caller's (pinned) buffer instead of as
// bytes for the host threads; RBG_HOST_IO_SPLIT=auto: adaptive (one slice per call towards whichever of bus and host
// finished last).  Default: none.  Measured (profiles/r02g_hostio_split.log): on the boxes of this pool the bus (38 MB of
// bytes per step land after 0.87 ms) and 8 widening threads (1.0 ms) are already balanced, so moving slices to the bus gains
// nothing (1 / 3 of 16 slices at the end of the step: 44.5 / 42.2 against 43.5 M env-steps/s on the same box) and a wide
// slice at the START of the step costs 0.7 ms.
static int host_io_split_env() {  // -2 off, -1 auto, >= 0 fixed
  static int v = -3;
  if (v == -3) {
    const char *e = getenv("RBG_HOST_IO_SPLIT");
    v = !e ? -2 : (strcmp(e, "auto") == 0 ? -1 : atoi(e));
    if (v < -2) v = -2;
  }
  return v;
}
static bool host_io_fixed_split() { return host_io_split_env() >= 0; }
static int host_io_split_init(int nslices) {
  const int v = host_io_split_env();
  if (v >= 0) return v > nslices ? nslices : v;
  return v == -1 ? nslices / 8 : 0;
}
static bool host_io_blocking_sync() {  // RBG_HOST_BLOCKING_SYNC=1: the calling thread sleeps on the copy events instead of spinning
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("RBG_HOST_BLOCKING_SYNC");
    v = (e && atoi(e) != 0) ? 1 : 0;
  }
  return v == 1;
}
static thread_local HostCtx *g_hc = &g_host[0];  // the device of the _host call in progress (set by use_device)
#define g_scratch (g_hc->scratch)
#define g_streams (g_hc->streams)
#define g_stream_ev (g_hc->stream_ev)
#define g_scratch_sig (g_hc->sig)

static int scratch_get(size_t bytes, int device, void **out) {
  if (g_scratch.ptr && (g_scratch.bytes < bytes || g_scratch.device != device)) {
    cudaFree(g_scratch.ptr);
    g_scratch = Scratch();
  }
  if (!g_scratch.ptr) {
    cudaError_t e = cudaMalloc(&g_scratch.ptr, bytes);
    if (e != cudaSuccess) {
      g_scratch = Scratch();
      return set_cuda_error(e, "cudaMalloc(host-variant scratch)");
    }
    g_scratch.bytes = bytes;
    g_scratch.device = device;
  }
  for (int i = 0; i < 3; ++i)
    if (!g_streams[i]) {
      cudaError_t e = cudaStreamCreateWithFlags(&g_streams[i], cudaStreamNonBlocking);
      if (e != cudaSuccess) return set_cuda_error(e, "cudaStreamCreate");
      if ((e = cudaEventCreateWithFlags(&g_stream_ev[i], cudaEventDisableTiming)) != cudaSuccess) return set_cuda_error(e, "cudaEventCreate");
    }
  *out = g_scratch.ptr;
  return RBG_OK;
}

struct Carver {
  uint8_t *base;
  size_t off = 0;
  template <typename T>
  T *take(size_t n) {
    off = (off + 255) & ~(size_t)255;
    T *p = base ? reinterpret_cast<T *>(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
};

static void carve_state(Carver &c, int64_t B, int G, int N, rbg_state *s) {
  s->grid = c.take<int32_t>((size_t)B * G * G);
  s->step_count = c.take<int32_t>((size_t)B);
  s->agent_id = c.take<int32_t>((size_t)B * N);
  s->start = c.take<int32_t>((size_t)B * N * 2);
  s->target = c.take<int32_t>((size_t)B * N * 2);
  s->position = c.take<int32_t>((size_t)B * N * 2);
  s->key = c.take<uint32_t>((size_t)B * 2);
}
static void carve_timestep(Carver &c, int64_t B, int G, int N, rbg_timestep *t) {
  t->obs_grid = c.take<int32_t>((size_t)B * N * G * G);
  t->action_mask = c.take<uint8_t>((size_t)B * N * 5);
  t->obs_step_count = c.take<int32_t>((size_t)B);
  t->reward = c.take<float>((size_t)B * N);
  t->discount = c.take<float>((size_t)B * N);
  t->step_type = c.take<int8_t>((size_t)B);
  t->num_connections = c.take<int32_t>((size_t)B);
  t->ratio_connections = c.take<float>((size_t)B);
  t->total_path_length = c.take<int32_t>((size_t)B);
}

#define RBG_CPY(dst, src, n, kind, st)                                                    \
  do {                                                                                    \
    cudaError_t e_ = cudaMemcpyAsync((dst), (src), (n), (kind), (st));                    \
    if (e_ != cudaSuccess) return set_cuda_error(e_, "cudaMemcpyAsync " #dst);            \
  } while (0)

static int copy_state(const rbg_state *dst, const rbg_state *src, int64_t off, int64_t n, int G, int N,
                      cudaMemcpyKind kind, int64_t dst_off, cudaStream_t st) {
  const size_t c = (size_t)G * G;
  RBG_CPY(dst->grid + dst_off * c, src->grid + off * c, n * c * 4, kind, st);
  RBG_CPY(dst->step_count + dst_off, src->step_count + off, n * 4, kind, st);
  RBG_CPY(dst->agent_id + dst_off * N, src->agent_id + off * N, n * N * 4, kind, st);
  RBG_CPY(dst->start + dst_off * N * 2, src->start + off * N * 2, n * N * 8, kind, st);
  RBG_CPY(dst->target + dst_off * N * 2, src->target + off * N * 2, n * N * 8, kind, st);
  RBG_CPY(dst->position + dst_off * N * 2, src->position + off * N * 2, n * N * 8, kind, st);
  RBG_CPY(dst->key + dst_off * 2, src->key + off * 2, n * 8, kind, st);
  return RBG_OK;
}
// `parts`: 1 = observation.grid (the bulk), 2 = every other leaf, 3 = both
static int copy_timestep(const rbg_timestep *dst, const rbg_timestep *src, int64_t off, int64_t n, int G, int N,
                         cudaMemcpyKind kind, int64_t dst_off, cudaStream_t st, int parts = 3) {
  const size_t c = (size_t)G * G * N;
  if (parts & 1) RBG_CPY(dst->obs_grid + dst_off * c, src->obs_grid + off * c, n * c * 4, kind, st);
  if (!(parts & 2)) return RBG_OK;
  RBG_CPY(dst->action_mask + dst_off * N * 5, src->action_mask + off * N * 5, n * N * 5, kind, st);
  RBG_CPY(dst->obs_step_count + dst_off, src->obs_step_count + off, n * 4, kind, st);
  RBG_CPY(dst->reward + dst_off * N, src->reward + off * N, n * N * 4, kind, st);
  RBG_CPY(dst->discount + dst_off * N, src->discount + off * N, n * N * 4, kind, st);
  RBG_CPY(dst->step_type + dst_off, src->step_type + off, n, kind, st);
  RBG_CPY(dst->num_connections + dst_off, src->num_connections + off, n * 4, kind, st);
  RBG_CPY(dst->ratio_connections + dst_off, src->ratio_connections + off, n * 4, kind, st);
  RBG_CPY(dst->total_path_length + dst_off, src->total_path_length + off, n * 4, kind, st);
  return RBG_OK;
}

static rbg_state state_at(const rbg_state &s, int64_t off, int G, int N) {
  rbg_state r;
  r.grid = s.grid + off * G * G;
  r.step_count = s.step_count + off;
  r.agent_id = s.agent_id + off * N;
  r.start = s.start + off * N * 2;
  r.target = s.target + off * N * 2;
  r.position = s.position + off * N * 2;
  r.key = s.key + off * 2;
  return r;
}
static rbg_timestep timestep_at(const rbg_timestep &t, int64_t off, int G, int N) {
  rbg_timestep r;
  r.obs_grid = t.obs_grid + off * N * G * G;
  r.action_mask = t.action_mask + off * N * 5;
  r.obs_step_count = t.obs_step_count + off;
  r.reward = t.reward + off * N;
  r.discount = t.discount + off * N;
  r.step_type = t.step_type + off;
  r.num_connections = t.num_connections + off;
  r.ratio_connections = t.ratio_connections + off;
  r.total_path_length = t.total_path_length + off;
  return r;
}

// slices are multiples of 64 envs so that every slice base keeps the 16-byte
// alignment of the bulk arrays for any (G, N)
static int64_t slice_size(int64_t B, int nslices) {
  int64_t s = (B + nslices - 1) / nslices;
  s = (s + 63) / 64 * 64;
  return s < 64 ? 64 : s;
}

}  // namespace rbg

using namespace rbg;

extern "C" {

int rbg_version(void) { return RBG_VERSION; }
const char *rbg_last_error(void) { return g_err; }

int rbg_device_info(int *sm_arch, int *sm_count) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return set_cuda_error(e, "cudaGetDevice");
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) return set_cuda_error(e, "cudaGetDeviceProperties");
  if (sm_arch) *sm_arch = prop.major * 10 + prop.minor;
  if (sm_count) *sm_count = prop.multiProcessorCount;
  return RBG_OK;
}

int64_t rbg_launch_count(int reset) {
  return reset ? g_launches.exchange(0) : g_launches.load();
}

int rbg_kernel_timing(int enable) {
  g_timing.store(enable ? 1 : 0);
  return RBG_OK;
}

int rbg_kernel_time(int kernel, int64_t *launches, double *total_ms) {
  if (kernel < 0 || kernel >= RBG_K_COUNT) return set_error(RBG_EINVAL, "rbg_kernel_time: unknown kernel id %d", kernel);
  std::vector<EventPair> evs;
  {
    std::lock_guard<std::mutex> lock(g_timing_mu);
    evs.swap(g_events[kernel]);
  }
  double ms = 0.0;
  int rc = RBG_OK;
  for (EventPair &ep : evs) {
    float t = 0.0f;
    cudaError_t e = cudaEventSynchronize(ep.b);
    if (e == cudaSuccess) e = cudaEventElapsedTime(&t, ep.a, ep.b);
    if (e != cudaSuccess) rc = set_cuda_error(e, "rbg_kernel_time");
    ms += t;
    cudaEventDestroy(ep.a);
    cudaEventDestroy(ep.b);
  }
  if (launches) *launches = (int64_t)evs.size();
  if (total_ms) *total_ms = ms;
  return rc;
}

int rbg_split_keys(const uint32_t key[2], int64_t B, int64_t offset, int64_t count, uint32_t *out, void *stream) {
  if (!key || !out) return set_error(RBG_EINVAL, "rbg_split_keys: NULL pointer");
  if (B < 1 || B > 0x7fffffffLL || offset < 0 || count < 0 || offset + count > B)
    return set_error(RBG_EINVAL, "rbg_split_keys: bad slice [%lld,+%lld) of B=%lld", (long long)offset,
                     (long long)count, (long long)B);
  return launch_split_keys(key[0], key[1], B, offset, count, out, (cudaStream_t)stream);
}

int rbg_prw_generate(const uint32_t *keys, int64_t B, int G, int N, int32_t *heads, int32_t *targets,
                     int32_t *solved, int32_t *stats, void *stream) {
  int rc;
  if ((rc = check_dims(B, G, N, 1))) return rc;
  if (B == 0) return RBG_OK;  // an empty batch has no buffers to check
  if (!keys || !heads || !targets || !solved) return set_error(RBG_EINVAL, "rbg_prw_generate: NULL pointer");
  if (!aligned16(solved)) return set_error(RBG_EALIGN, "rbg_prw_generate: solved not 16-byte aligned");
  PrwParams p;
  memset(&p, 0, sizeof(p));
  p.keys = keys;
  p.B = B;
  p.G = G;
  p.N = N;
  p.mode = PRW_MODE_BOARD;
  p.debug = debug_flags();
  p.heads = heads;
  p.targets = targets;
  p.solved = solved;
  p.stats = stats;
  return launch_prw(p, B, env_int("RBG_PRW_M"), env_int("RBG_PRW_THREADS"), (cudaStream_t)stream);
}

int rbg_generator_state(int kind, const uint32_t *keys, int64_t B, int G, int N, const rbg_state *out, void *stream) {
  int rc;
  // (the sequential random walk has no choice(replace=False) over the agents: more agents than cells is a failed generation)
  if ((rc = check_dims(B, G, N, kind == RBG_GEN_UNIFORM ? 2 : kind == RBG_GEN_SEQRW ? 0 : 1))) return rc;
  if (B == 0) return RBG_OK;  // an empty batch has no buffers to check
  if (!keys) return set_error(RBG_EINVAL, "keys is NULL");
  if ((rc = check_state(out, "state", G))) return rc;
  return generator_state_impl(kind, keys, B, G, N, out, nullptr, 0, nullptr, nullptr, (cudaStream_t)stream);
}

int rbg_dataset_state(const uint32_t *keys, int64_t B, int G, int N, const int32_t *heads, const int32_t *targets, int64_t K,
                      const rbg_state *out, void *stream) {
  int rc;
  if ((rc = check_dims(B, G, N, 1))) return rc;
  if (B == 0) return RBG_OK;
  if (!keys || !heads || !targets) return set_error(RBG_EINVAL, "rbg_dataset_state: NULL pointer");
  if (K < 1 || K > 0x7fffffffLL) return set_error(RBG_EINVAL, "rbg_dataset_state: number of boards K=%lld", (long long)K);
  if ((rc = check_state(out, "state", G))) return rc;
  return launch_dataset_state(keys, B, G, N, heads, targets, K, *out, (cudaStream_t)stream);
}

int rbg_seedext_solved(const uint32_t *keys, int64_t B, int G, int N, float randomness, int two_sided,
                       int iterations, int64_t extension_steps, int32_t *solved, void *stream) {
  int rc;
  if ((rc = check_dims(B, G, N, 1))) return rc;
  if (B == 0) return RBG_OK;  // an empty batch has no buffers to check
  if (!keys || !solved) return set_error(RBG_EINVAL, "rbg_seedext_solved: NULL pointer");
  if (!aligned16(solved)) return set_error(RBG_EALIGN, "rbg_seedext_solved: solved not 16-byte aligned");
  SeedExtParams p;
  memset(&p, 0, sizeof(p));
  p.keys = keys;
  p.B = B;
  p.G = G;
  p.N = N;
  p.randomness = randomness;
  p.two_sided = two_sided;
  p.iterations = iterations;
  p.ext_steps = extension_steps;
  p.mode = 0;
  p.solved = solved;
  return launch_seedext(p, B, (cudaStream_t)stream);
}

int rbg_seedext_starts_ends(const uint32_t *keys, int64_t B, int G, int N, float randomness, int two_sided,
                            int iterations, int64_t extension_steps, int32_t *starts, int32_t *ends, void *stream) {
  int rc;
  if ((rc = check_dims(B, G, N, 1))) return rc;
  if (B == 0) return RBG_OK;  // an empty batch has no buffers to check
  if (!keys || !starts || !ends) return set_error(RBG_EINVAL, "rbg_seedext_starts_ends: NULL pointer");
  SeedExtParams p;
  memset(&p, 0, sizeof(p));
  p.keys = keys;
  p.B = B;
  p.G = G;
  p.N = N;
  p.randomness = randomness;
  p.two_sided = two_sided;
  p.iterations = iterations;
  p.ext_steps = extension_steps;
  p.mode = 1;
  p.starts = starts;
  p.ends = ends;
  return launch_seedext(p, B, (cudaStream_t)stream);
}

int rbg_seqrw_generate(const uint32_t *keys, int64_t B, int G, int N, void *board, int as_float32, int32_t *stats, void *stream) {
  int rc;
  if ((rc = check_dims(B, G, N, 0))) return rc;
  if (B == 0) return RBG_OK;
  if (G < 3) return set_error(RBG_EINVAL, "SequentialRandomWalk: rows=%d (available_cells pads with jnp.full(rows - 3, -1): rows >= 3)", G);
  if (!keys || !board) return set_error(RBG_EINVAL, "rbg_seqrw_generate: NULL argument");
  if (!aligned16(board)) return set_error(RBG_EALIGN, "board not 16-byte aligned");
  SeqRwParams p;
  memset(&p, 0, sizeof(p));
  p.keys = keys;
  p.B = B;
  p.G = G;
  p.N = N;
  p.mode = 0;
  p.float_board = as_float32 ? 1 : 0;
  p.board = reinterpret_cast<int32_t *>(board);
  p.stats = stats;
  return launch_seqrw(p, B, (cudaStream_t)stream);
}

int rbg_seqrw_starts_ends(const uint32_t *keys, int64_t B, int G, int N, int32_t *starts, int32_t *ends, void *stream) {
  int rc;
  if ((rc = check_dims(B, G, N, 0))) return rc;
  if (B == 0) return RBG_OK;
  if (!keys || !starts || !ends) return set_error(RBG_EINVAL, "rbg_seqrw_starts_ends: NULL argument");
  SeqRwParams p;
  memset(&p, 0, sizeof(p));
  p.keys = keys;
  p.B = B;
  p.G = G;
  p.N = N;
  p.mode = 1;
  p.starts = starts;
  p.ends = ends;
  return launch_seqrw(p, B, (cudaStream_t)stream);
}

int rbg_connector_observe(const rbg_state *state, int64_t B, int G, int N, const rbg_timestep *ts, void *stream) {
  int rc;
  if ((rc = check_dims(B, G, N, 1))) return rc;
  if (B == 0) return RBG_OK;  // an empty batch has no buffers to check
  if ((rc = check_state(state, "state", G))) return rc;
  if ((rc = check_timestep(ts, G))) return rc;
  EnvParams p;
  memset(&p, 0, sizeof(p));
  p.in = *state;
  p.out = *state;
  p.ts = *ts;
  p.B = B;
  p.G = G;
  p.N = N;
  p.mode = ENV_MODE_OBSERVE;
  p.env.autoreset_kind = -1;
  return launch_env(p, (cudaStream_t)stream);
}

int rbg_connector_reset(int kind, const uint32_t *keys, int64_t B, int G, int N, const rbg_state *state,
                        const rbg_timestep *ts, void *stream) {
  int rc = rbg_generator_state(kind, keys, B, G, N, state, stream);
  if (rc) return rc;
  return rbg_connector_observe(state, B, G, N, ts, stream);
}

int rbg_connector_reset_dataset(const uint32_t *keys, int64_t B, int G, int N, const int32_t *heads, const int32_t *targets,
                                int64_t K, const rbg_state *state, const rbg_timestep *ts, void *stream) {
  int rc = rbg_dataset_state(keys, B, G, N, heads, targets, K, state, stream);
  if (rc) return rc;
  return rbg_connector_observe(state, B, G, N, ts, stream);
}

int rbg_split_each(const uint32_t *keys, int64_t B, int num, uint32_t *out, void *stream) {
  if (B < 0 || B > 0x3fffffffLL) return set_error(RBG_EINVAL, "rbg_split_each: B=%lld", (long long)B);
  if (num < 1 || num > 65536) return set_error(RBG_EINVAL, "rbg_split_each: num=%d outside [1,65536]", num);
  if (B == 0) return RBG_OK;
  if (!keys || !out) return set_error(RBG_EINVAL, "rbg_split_each: NULL pointer");
  return launch_split_each(keys, B, num, out, (cudaStream_t)stream);
}

int64_t rbg_step_workspace_bytes(int64_t B, int G, int N) {
  (void)G;
  return (int64_t)ws_layout(B, N).total;
}

int rbg_workspace_release(void *workspace) {
  std::lock_guard<std::mutex> lock(g_ar_mu);
  auto it = g_ar.find(workspace);
  if (it == g_ar.end()) return RBG_OK;
  AutoResetCtx &c = it->second;
  if (c.side) {
    cudaStreamSynchronize(c.side);
    cudaStreamDestroy(c.side);
    cudaEventDestroy(c.env_done);
    cudaEventDestroy(c.refill_done[0]);
    cudaEventDestroy(c.refill_done[1]);
  }
  for (int i = 0; i < 3; ++i)
    if (c.slice_stream[i]) {
      cudaStreamSynchronize(c.slice_stream[i]);
      cudaStreamDestroy(c.slice_stream[i]);
      cudaEventDestroy(c.slice_done[i]);
    }
  if (c.slice_fork) cudaEventDestroy(c.slice_fork);
  if (c.tail_valid) {
    cudaStreamSynchronize(c.last_stream);
    cudaGetLastError();
  }
  g_ar.erase(it);
  return RBG_OK;
}

int rbg_connector_step(const rbg_state *in, const rbg_state *out, const int32_t *action, int64_t B, int G, int N,
                       const rbg_env_params *params, const rbg_timestep *ts, void *workspace, void *stream) {
  return connector_step_impl(in, out, action, nullptr, 0, B, G, N, params, ts, workspace, (cudaStream_t)stream);
}

int rbg_connector_step_random(const rbg_state *in, const rbg_state *out, int32_t *action_out, int64_t B, int G, int N,
                              const rbg_env_params *params, const rbg_timestep *ts, void *workspace, void *stream) {
  return connector_step_impl(in, out, nullptr, action_out, 1, B, G, N, params, ts, workspace, (cudaStream_t)stream);
}

// Fused rollout (rollout_warp_kernel): one launch per chunk of steps plus one refill of
// the next-episode cache for the envs that reset during the chunk.
static int rollout_fused(const rbg_state *state, int32_t *action_out, int64_t T, int64_t B, int G, int N,
                         const rbg_env_params *params, const rbg_timestep *ts, void *workspace, cudaStream_t stream) {
  uint8_t *ws = reinterpret_cast<uint8_t *>(workspace);
  const WsLayout wl = ws_layout(B, N);
  const int kind = params->autoreset_kind;
  cudaError_t e;
  int rc;
  AutoResetCtx *ctx;
  {
    std::lock_guard<std::mutex> lock(g_ar_mu);
    ctx = &g_ar[workspace];
  }
  if ((rc = workspace_follow_stream(ctx, stream))) return rc;
  // refills launched by the step-wise path on the side stream must have landed
  for (int i = 0; i < 2; ++i)
    if (ctx->refill_pending[i]) {
      if ((e = cudaStreamWaitEvent(stream, ctx->refill_done[i], 0)) != cudaSuccess) return set_cuda_error(e, "cudaStreamWaitEvent");
      ctx->refill_pending[i] = false;
    }
  auto cache_params = [&](PrwParams &q) {
    memset(&q, 0, sizeof(q));
    q.B = B;
    q.G = G;
    q.N = N;
    q.mode = kind == RBG_GEN_PRW ? PRW_MODE_STATE : PRW_MODE_UNIFORM;
    q.extra_split = (kind == RBG_GEN_PRW ? 1 : 0) + 1;
    q.debug = debug_flags();
    q.to_cache = 1;
    q.cache_tag = reinterpret_cast<unsigned long long *>(ws + wl.cache_tag);
    q.cache_key = reinterpret_cast<uint2 *>(ws + wl.cache_key);
    q.cache_pins = reinterpret_cast<uint32_t *>(ws + wl.cache_pins);
  };
  // Next-episode cache + refill kernel, or every reset generated inside the rollout kernel?  Inside,
  // a whole warp generates one board (prw_kernel packs 32 / Np per warp), but the work hides in the
  // kernel's idle issue slots and no refill kernel runs.  Measured: the cache wins while a step emits
  // little per env (10x10/5: 2.10 vs 1.45 G env-steps/s, 16x16/8: 678 vs 636 M), in-kernel generation
  // wins for large boards, where resets are rare per byte written (24x24/12: 214 vs 183 M, 32x32/16:
  // 95 vs 84 M, 40x40/32: 26.8 vs 23.3 M); 20x20/10 (16 000 B per env-step) is the break-even.
  static int no_cache_env = -2;
  if (no_cache_env == -2) {
    const char *ex = getenv("RBG_ROLLOUT_NO_CACHE");
    no_cache_env = ex ? (atoi(ex) != 0 ? 1 : 0) : -1;
  }
  // RBG_ROLLOUT_IMPL=legacy: rollout_warp_kernel + refill kernel (round 1); default: the persistent kernel whose
  // CTAs carry their own generator warps (rollout_persist_kernel), which always works from the cache
  static int persist = -1;
  if (persist < 0) {
    const char *ex = getenv("RBG_ROLLOUT_IMPL");
    persist = (ex && strcmp(ex, "legacy") == 0) ? 0 : 1;
  }
  const bool no_cache = persist ? false : (no_cache_env >= 0 ? no_cache_env == 1 : (size_t)N * G * G * 4 > 16384);
  if (no_cache) {
    ctx->B = 0;  // the cache is not maintained: a later step-wise call adopts the workspace afresh
  } else if (ctx->B != B || ctx->G != G || ctx->N != N || ctx->kind != kind) {
    if ((e = cudaMemsetAsync(ws, 0, 256, stream)) != cudaSuccess) return set_cuda_error(e, "cudaMemsetAsync(counters)");
    PrwParams q;  // warm start: the next episode of every env
    cache_params(q);
    q.keys = state->key;
    if ((rc = launch_prw(q, B, env_int("RBG_PRW_M"), env_int("RBG_PRW_THREADS"), stream))) return rc;
    ctx->B = B;
    ctx->G = G;
    ctx->N = N;
    ctx->kind = kind;
    ctx->step = 0;
    ctx->persist_epoch = 0;
  }
  static int chunk_env = -1;
  if (chunk_env < 0) {
    chunk_env = env_int("RBG_ROLLOUT_CHUNK");
    if (chunk_env <= 0) chunk_env = 20;  // the reference's n_steps (agent_training/configs/env/connector.yaml:27)
  }
  EnvParams p;
  memset(&p, 0, sizeof(p));
  p.in = *state;
  p.out = *state;
  p.B = B;
  p.G = G;
  p.N = N;
  p.mode = ENV_MODE_STEP;
  p.random_policy = 1;
  p.env = *params;
  p.cache_tag = reinterpret_cast<const unsigned long long *>(ws + wl.cache_tag);
  p.cache_key = reinterpret_cast<const uint2 *>(ws + wl.cache_key);
  p.cache_pins = reinterpret_cast<const uint32_t *>(ws + wl.cache_pins);
  if (no_cache) p.cache_tag = nullptr;
  // Slices of the batch on their own streams: each slice alternates rollout chunk -> cache
  // refill, and while one slice's (memory-bound) rollout drains, another slice's (issue-bound)
  // refill and the ramp of its next chunk fill the SMs.  Slices never share an env, a list or
  // a cache entry, so the result does not depend on how they interleave.
  static int slices_env = -1;
  if (slices_env < 0) {
    slices_env = env_int("RBG_ROLLOUT_SLICES");
    if (slices_env <= 0) slices_env = 2;
    if (slices_env > 4) slices_env = 4;
  }
  int nsl = persist ? 1 : slices_env;  // a persistent grid fills the machine by itself; its launches overlap head to tail instead
  while (nsl > 1 && B < (int64_t)nsl * 2048) --nsl;
  if (g_timing.load(std::memory_order_relaxed)) nsl = 1;  // event pairs must time one kernel at a time
  int64_t cuts[5];
  cuts[0] = 0;
  for (int s = 1; s < nsl; ++s) cuts[s] = ((B * s / nsl) + 127) & ~(int64_t)127;  // whole CTAs on either side
  cuts[nsl] = B;
  if (nsl > 1) {
    if (!ctx->slice_fork && (e = cudaEventCreateWithFlags(&ctx->slice_fork, cudaEventDisableTiming)) != cudaSuccess) return set_cuda_error(e, "cudaEventCreate");
    if ((e = cudaEventRecord(ctx->slice_fork, stream)) != cudaSuccess) return set_cuda_error(e, "cudaEventRecord(fork)");
    for (int s = 1; s < nsl; ++s) {
      if (!ctx->slice_stream[s - 1]) {
        if ((e = cudaStreamCreateWithFlags(&ctx->slice_stream[s - 1], cudaStreamNonBlocking)) != cudaSuccess) return set_cuda_error(e, "cudaStreamCreate(slice)");
        if ((e = cudaEventCreateWithFlags(&ctx->slice_done[s - 1], cudaEventDisableTiming)) != cudaSuccess) return set_cuda_error(e, "cudaEventCreate");
      }
      if ((e = cudaStreamWaitEvent(ctx->slice_stream[s - 1], ctx->slice_fork, 0)) != cudaSuccess) return set_cuda_error(e, "cudaStreamWaitEvent(fork)");
    }
  }
  for (int64_t t0 = 0; t0 < T; t0 += chunk_env) {
    const int n = (int)((T - t0) < chunk_env ? (T - t0) : chunk_env);
    p.ts = timestep_at(*ts, t0 * B, G, N);
    for (int s = 0; s < nsl; ++s) {
      cudaStream_t st = s == 0 ? stream : ctx->slice_stream[s - 1];
      p.env_lo = cuts[s];
      p.env_hi = cuts[s + 1];
      // the slice's share of the list buffers; counter + ticket pairs 16 bytes apart from ws + 64
      p.refill_list = reinterpret_cast<int32_t *>(ws + wl.refill_list[0]) + p.env_lo;
      p.refill_keys = reinterpret_cast<uint32_t *>(ws + wl.refill_keys[0]) + 2 * p.env_lo;
      p.refill_count = reinterpret_cast<int32_t *>(ws + 64 + 16 * s);
      if (persist) {
        static int overlap_env = -1;  // RBG_ROLLOUT_OVERLAP=0: plain stream order between launches
        if (overlap_env < 0) {
          const char *ex = getenv("RBG_ROLLOUT_OVERLAP");
          overlap_env = (ex && atoi(ex) == 0) ? 0 : 1;
        }
        if (ctx->persist_epoch == 0 || ctx->persist_epoch > 0x3fffffff) {  // flags start at 0 = "written by launch 0"
          if ((e = cudaMemsetAsync(ws + wl.group_done, 0, (wl.group_pending - wl.group_done) + 4 * (size_t)B, st)) != cudaSuccess) return set_cuda_error(e, "cudaMemsetAsync(group flags)");
          if ((e = cudaMemsetAsync(ws + 256, 0, 256, st)) != cudaSuccess) return set_cuda_error(e, "cudaMemsetAsync(persist counters)");
          ctx->persist_epoch = 0;
        }
        const int epoch = ++ctx->persist_epoch;
        const bool overlap = overlap_env == 1 && epoch > 1 && !g_timing.load(std::memory_order_relaxed);
        if ((rc = launch_rollout_persist(p, kind, n, action_out ? action_out + t0 * B * N : nullptr, reinterpret_cast<int32_t *>(ws + 256),
                                         reinterpret_cast<int32_t *>(ws + wl.group_done), reinterpret_cast<int32_t *>(ws + wl.group_pending), epoch, overlap,
                                         reinterpret_cast<uint64_t *>(ws + wl.cache_tag), reinterpret_cast<uint2 *>(ws + wl.cache_key),
                                         reinterpret_cast<uint32_t *>(ws + wl.cache_pins), st)))
          return rc;
        continue;
      }
      if (no_cache) p.refill_list = nullptr;
      if ((rc = launch_rollout(p, kind, n, action_out ? action_out + t0 * B * N : nullptr, st))) return rc;
      if (no_cache) continue;
      PrwParams q;  // refill the cache entries consumed in this chunk (the kernel clears the counter when done)
      cache_params(q);
      q.keys = p.refill_keys;
      q.keys_compact = 1;
      q.list = p.refill_list;
      q.list_count = p.refill_count;
      q.list_ticket = p.refill_count + 1;
      q.bulk_list = 1;
      if ((rc = launch_prw(q, p.env_hi - p.env_lo, env_int("RBG_PRW_M"), env_int("RBG_PRW_THREADS"), st))) return rc;
    }
  }
  for (int s = 1; s < nsl; ++s) {
    if ((e = cudaEventRecord(ctx->slice_done[s - 1], ctx->slice_stream[s - 1])) != cudaSuccess) return set_cuda_error(e, "cudaEventRecord(join)");
    if ((e = cudaStreamWaitEvent(stream, ctx->slice_done[s - 1], 0)) != cudaSuccess) return set_cuda_error(e, "cudaStreamWaitEvent(join)");
  }
  return RBG_OK;
}

int rbg_connector_rollout_random(const rbg_state *state, int32_t *action_out, int64_t T, int64_t B, int G, int N,
                                 const rbg_env_params *params, const rbg_timestep *ts, void *workspace, void *stream) {
  int rc;
  if (T < 0) return set_error(RBG_EINVAL, "rollout length T=%lld", (long long)T);
  if ((rc = check_dims(B, G, N, 1))) return rc;
  if (B == 0 || T == 0) return RBG_OK;  // an empty batch has no buffers to check
  if ((rc = check_state(state, "state", G))) return rc;
  if ((rc = check_timestep(ts, G))) return rc;
  if (!params) return set_error(RBG_EINVAL, "params is NULL");
  // (when G*G % 4 == 0 every step slice of the stacked arrays keeps the 16-byte alignment the
  // vector path needs; otherwise the kernels use scalar accesses)
  if (B == 0 || T == 0) return RBG_OK;
  const int kind = params->autoreset_kind;
  static int fused = -1;
  if (fused < 0) fused = env_int("RBG_NO_FUSED_ROLLOUT") ? 0 : 1;
  if (kind > RBG_GEN_SEQRW) return set_error(RBG_EINVAL, "unknown autoreset generator kind %d", kind);
  if (kind < 0) return set_error(RBG_EINVAL, "rollout needs an auto-reset generator kind (params->autoreset_kind >= 0)");
  if (kind == RBG_GEN_DATASET && (rc = check_dataset(params))) return rc;
  if (fused && kind == RBG_GEN_DATASET) {
    // resets are a table lookup inside the kernel: no cache, no refill, no workspace
    int chunk = env_int("RBG_ROLLOUT_CHUNK");
    if (chunk <= 0) chunk = 20;
    EnvParams p;
    memset(&p, 0, sizeof(p));
    p.in = *state;
    p.out = *state;
    p.B = B;
    p.G = G;
    p.N = N;
    p.mode = ENV_MODE_STEP;
    p.random_policy = 1;
    p.env = *params;
    for (int64_t t0 = 0; t0 < T; t0 += chunk) {
      const int n = (int)((T - t0) < chunk ? (T - t0) : chunk);
      p.ts = timestep_at(*ts, t0 * B, G, N);
      if ((rc = launch_rollout(p, kind, n, action_out ? action_out + t0 * B * N : nullptr, (cudaStream_t)stream))) return rc;
    }
    return RBG_OK;
  }
  if (fused && speculative_enabled() && workspace && (kind == RBG_GEN_PRW || kind == RBG_GEN_UNIFORM)) {
    if (!aligned16(workspace)) return set_error(RBG_EALIGN, "workspace not 16-byte aligned");
    return rollout_fused(state, action_out, T, B, G, N, params, ts, workspace, (cudaStream_t)stream);
  }
  for (int64_t t = 0; t < T; ++t) {  // step-wise: SeedExtension / SequentialRandomWalk resets, or the fused kernel switched off
    const rbg_timestep tt = timestep_at(*ts, t * B, G, N);
    rc = connector_step_impl(state, state, nullptr, action_out ? action_out + t * B * N : nullptr, 1, B, G, N, params, &tt, workspace,
                             (cudaStream_t)stream);
    if (rc) return rc;
  }
  return RBG_OK;
}

int rbg_random_actions(const rbg_state *state, int64_t B, int G, int N, int32_t *action, void *stream) {
  int rc;
  if ((rc = check_dims(B, G, N, 1))) return rc;
  if (B == 0) return RBG_OK;  // an empty batch has no buffers to check
  if ((rc = check_state(state, "state", G))) return rc;
  if (!action) return set_error(RBG_EINVAL, "action is NULL");
  return launch_random_actions(*state, B, G, N, action, (cudaStream_t)stream);
}

int rbg_validate(const int32_t *boards, int64_t B, int G, int N, int32_t *flags, void *stream) {
  int rc;
  if ((rc = check_dims(B, G, N, 0))) return rc;
  if (B == 0) return RBG_OK;  // an empty batch has no buffers to check
  if (!boards || !flags) return set_error(RBG_EINVAL, "rbg_validate: NULL pointer");
  return launch_validate(boards, B, G, N, flags, (cudaStream_t)stream);
}

int rbg_board_statistics(const int32_t *boards, int64_t B, int G, int count_current_wire, int32_t *scored, int32_t *detours, int32_t *diversity,
                         void *stream) {
  int rc;
  if ((rc = check_dims(B, G, 1, 0))) return rc;
  if (B == 0) return RBG_OK;
  if (!boards || !detours || !diversity) return set_error(RBG_EINVAL, "rbg_board_statistics: NULL pointer");
  return launch_board_stats(boards, B, G, count_current_wire, scored, detours, diversity, (cudaStream_t)stream);
}

// ---------------------------------------------------------------- _host
void *rbg_host_alloc(int64_t bytes) {
  void *p = nullptr;
  if (cudaMallocHost(&p, (size_t)bytes) != cudaSuccess) {
    set_error(RBG_ENOMEM, "cudaMallocHost(%lld) failed", (long long)bytes);
    return nullptr;
  }
  return p;
}
void rbg_host_free(void *p) {
  if (p) cudaFreeHost(p);
}

static int use_device(int device, int *cur) {
  cudaError_t e = cudaGetDevice(cur);
  if (e != cudaSuccess) return set_cuda_error(e, "cudaGetDevice (no CUDA device: this library has no CPU path)");
  if (device >= 0 && device != *cur) {
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaSetDevice");
    *cur = device;
  }
  if (*cur < 0 || *cur >= kMaxDevices) return set_error(RBG_EINVAL, "device %d out of range", *cur);
  g_hc = &g_host[*cur];
  return RBG_OK;
}

// A host-variant call is about to use the scratch as `sig` describes.  If that is not how it was
// used last, the auto-reset workspaces a step variant keeps in it (`nws` of them, `stride` bytes
// apart from `ws`) hold someone else's bytes: their contexts are invalidated.
static void scratch_claim(const ScratchSig &sig, uint8_t *ws, size_t stride, int64_t nws) {
  if (g_scratch_sig == sig) return;
  g_scratch_sig = sig;
  for (int64_t i = 0; i < nws; ++i) invalidate_workspace(ws + (size_t)i * stride, nullptr, true);
}

int rbg_prw_generate_host(const uint32_t *keys, int64_t B, int G, int N, int32_t *heads, int32_t *targets,
                          int32_t *solved, int device) {
  int rc, dev;
  if ((rc = check_dims(B, G, N, 1))) return rc;
  if (B == 0) return RBG_OK;  // an empty batch has no buffers to check
  if (!keys || !heads || !targets || !solved) return set_error(RBG_EINVAL, "rbg_prw_generate_host: NULL pointer");
  if ((rc = use_device(device, &dev))) return rc;
  std::lock_guard<std::mutex> lock(g_scratch_mu);
  Carver size{nullptr};
  size.take<uint32_t>((size_t)B * 2);
  size.take<int32_t>((size_t)B * 2 * N);
  size.take<int32_t>((size_t)B * 2 * N);
  size.take<int32_t>((size_t)B * G * G);
  void *base;
  if ((rc = scratch_get(size.off + 256, dev, &base))) return rc;
  scratch_claim(ScratchSig{1, B, G, N, -1, base}, nullptr, 0, 0);
  Carver c{reinterpret_cast<uint8_t *>(base)};
  uint32_t *dk = c.take<uint32_t>((size_t)B * 2);
  int32_t *dh = c.take<int32_t>((size_t)B * 2 * N);
  int32_t *dt = c.take<int32_t>((size_t)B * 2 * N);
  int32_t *ds = c.take<int32_t>((size_t)B * G * G);
  const int nsl = B >= 4096 ? 4 : 1;
  const int64_t sl = slice_size(B, nsl);
  int si = 0;
  for (int64_t off = 0; off < B; off += sl, ++si) {
    const int64_t n = (B - off) < sl ? (B - off) : sl;
    cudaStream_t st = g_streams[si % 3];
    RBG_CPY(dk + off * 2, keys + off * 2, n * 8, cudaMemcpyHostToDevice, st);
    rc = rbg_prw_generate(dk + off * 2, n, G, N, dh + off * 2 * N, dt + off * 2 * N, ds + off * G * G, nullptr, st);
    if (rc) return rc;
    RBG_CPY(heads + off * 2 * N, dh + off * 2 * N, n * 2 * N * 4, cudaMemcpyDeviceToHost, st);
    RBG_CPY(targets + off * 2 * N, dt + off * 2 * N, n * 2 * N * 4, cudaMemcpyDeviceToHost, st);
    RBG_CPY(solved + off * G * G, ds + off * G * G, n * G * G * 4, cudaMemcpyDeviceToHost, st);
  }
  for (int i = 0; i < 3; ++i) {
    cudaError_t e = cudaStreamSynchronize(g_streams[i]);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaStreamSynchronize");
  }
  return RBG_OK;
}

int rbg_connector_reset_host(int kind, const uint32_t *keys, int64_t B, int G, int N, const rbg_state *state,
                             const rbg_timestep *ts, int device) {
  int rc, dev;
  if ((rc = check_dims(B, G, N, kind == RBG_GEN_UNIFORM ? 2 : 1))) return rc;
  if (B == 0) return RBG_OK;  // an empty batch has no buffers to check
  if (!keys || !state || !ts) return set_error(RBG_EINVAL, "rbg_connector_reset_host: NULL pointer");
  if ((rc = use_device(device, &dev))) return rc;
  std::lock_guard<std::mutex> lock(g_scratch_mu);
  Carver size{nullptr};
  rbg_state ds;
  rbg_timestep dt;
  size.take<uint32_t>((size_t)B * 2);
  carve_state(size, B, G, N, &ds);
  carve_timestep(size, B, G, N, &dt);
  void *base;
  if ((rc = scratch_get(size.off + 256, dev, &base))) return rc;
  scratch_claim(ScratchSig{2, B, G, N, kind, base}, nullptr, 0, 0);
  Carver c{reinterpret_cast<uint8_t *>(base)};
  uint32_t *dk = c.take<uint32_t>((size_t)B * 2);
  carve_state(c, B, G, N, &ds);
  carve_timestep(c, B, G, N, &dt);
  const int nsl = B >= 4096 ? 4 : 1;
  const int64_t sl = slice_size(B, nsl);
  int si = 0;
  for (int64_t off = 0; off < B; off += sl, ++si) {
    const int64_t n = (B - off) < sl ? (B - off) : sl;
    cudaStream_t st = g_streams[si % 3];
    RBG_CPY(dk + off * 2, keys + off * 2, n * 8, cudaMemcpyHostToDevice, st);
    rbg_state dss = state_at(ds, off, G, N);
    rbg_timestep dts = timestep_at(dt, off, G, N);
    if ((rc = rbg_connector_reset(kind, dk + off * 2, n, G, N, &dss, &dts, st))) return rc;
    if ((rc = copy_state(state, &ds, off, n, G, N, cudaMemcpyDeviceToHost, off, st))) return rc;
    if ((rc = copy_timestep(ts, &dt, off, n, G, N, cudaMemcpyDeviceToHost, off, st))) return rc;
  }
  for (int i = 0; i < 3; ++i) {
    cudaError_t e = cudaStreamSynchronize(g_streams[i]);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaStreamSynchronize");
  }
  return RBG_OK;
}

int rbg_connector_step_host(const rbg_state *in, const rbg_state *out, const int32_t *action, int64_t B, int G,
                            int N, const rbg_env_params *params, const rbg_timestep *ts, int device) {
  int rc, dev;
  if ((rc = check_dims(B, G, N, 1))) return rc;
  if (B == 0) return RBG_OK;  // an empty batch has no buffers to check
  if (!in || !out || !action || !params || !ts) return set_error(RBG_EINVAL, "rbg_connector_step_host: NULL pointer");
  if ((rc = use_device(device, &dev))) return rc;
  std::lock_guard<std::mutex> lock(g_scratch_mu);
  const int nsl = B >= 4096 ? 8 : 1;
  const int64_t sl = slice_size(B, nsl);
  const int64_t nslices = (B + sl - 1) / sl;
  Carver size{nullptr};
  rbg_state ds;
  rbg_timestep dt;
  carve_state(size, B, G, N, &ds);
  carve_timestep(size, B, G, N, &dt);
  size.take<int32_t>((size_t)B * N);
  const size_t ws_bytes = (size_t)rbg_step_workspace_bytes(sl, G, N);
  size.take<uint8_t>((size_t)nslices * ws_bytes);
  void *base;
  if ((rc = scratch_get(size.off + 256, dev, &base))) return rc;
  Carver c{reinterpret_cast<uint8_t *>(base)};
  carve_state(c, B, G, N, &ds);
  carve_timestep(c, B, G, N, &dt);
  int32_t *da = c.take<int32_t>((size_t)B * N);
  uint8_t *ws = c.take<uint8_t>((size_t)nslices * ws_bytes);
  scratch_claim(ScratchSig{3, B, G, N, params->autoreset_kind, base}, ws, ws_bytes, nslices);
  int si = 0;
  for (int64_t off = 0; off < B; off += sl, ++si) {
    const int64_t n = (B - off) < sl ? (B - off) : sl;
    cudaStream_t st = g_streams[si % 3];
    if ((rc = copy_state(&ds, in, off, n, G, N, cudaMemcpyHostToDevice, off, st))) return rc;
    RBG_CPY(da + off * N, action + off * N, n * N * 4, cudaMemcpyHostToDevice, st);
    rbg_state dss = state_at(ds, off, G, N);
    rbg_timestep dts = timestep_at(dt, off, G, N);
    if ((rc = rbg_connector_step(&dss, &dss, da + off * N, n, G, N, params, &dts, ws + (size_t)si * ws_bytes, st)))
      return rc;
    if ((rc = copy_state(out, &ds, off, n, G, N, cudaMemcpyDeviceToHost, off, st))) return rc;
    if ((rc = copy_timestep(ts, &dt, off, n, G, N, cudaMemcpyDeviceToHost, off, st))) return rc;
  }
  for (int i = 0; i < 3; ++i) {
    cudaError_t e = cudaStreamSynchronize(g_streams[i]);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaStreamSynchronize");
  }
  return RBG_OK;
}

// Env-pool style step: the State stays on the device (updated in place), the actions come from and
// the TimeStep goes to HOST buffers, pipelined over env slices on three streams.
int rbg_connector_step_host_io(const rbg_state *state, const int32_t *action, int64_t B, int G, int N,
                               const rbg_env_params *params, const rbg_timestep *ts, int device) {
  int rc, dev;
  if ((rc = check_dims(B, G, N, 1))) return rc;
  if (B == 0) return RBG_OK;  // an empty batch has no buffers to check
  if (!action || !params || !ts) return set_error(RBG_EINVAL, "rbg_connector_step_host_io: NULL pointer");
  if ((rc = check_state(state, "state", G))) return rc;
  if ((rc = use_device(device, &dev))) return rc;
  std::lock_guard<std::mutex> lock(g_scratch_mu);
  cudaError_t e;
  const bool packed = !host_io_wide();
  const auto t_call0 = std::chrono::steady_clock::now();
  int tune_slot = -1;  // which candidate this call measures (odd calls of the tuning phase)
  bool in_tuning = false;  // the thread count is being tried out: the bus / host split stays put meanwhile
  if (packed && !host_pool_fixed() && host_pool_max_threads() >= 4) {
    const int64_t key = B * 4096 + (int64_t)G * 64 + N;
    if (g_hc->tune_B != key) {
      g_hc->tune_B = key;
      g_hc->tune_call = 0;
      g_hc->tune_best = 0;
      g_hc->tune_best_s = 0.0;
    }
    // candidates: a big pool (one rank per host) tries 1/4, 3/8 and 1/2 of the cores: beyond half of them the calling thread
    // (which polls the copy events), the driver's threads and whatever else the process runs collide with the workers and
    // the step becomes bimodal (16-core boxes: 8 threads 61-66 M env-steps/s; 12 threads 35-58 M; 16 threads 25-44 M).  A
    // small pool (several ranks share the host) tries 1/4, 1/2, 3/4 and all of its few workers.
    const bool big = host_pool_max_threads() >= 8;
    const int ncand = big ? 3 : 4;
    if (g_hc->tune_call < 4 * ncand) {
      in_tuning = true;
      const int cand = g_hc->tune_call / 4;  // four calls each, the last three timed
      int n = big ? host_pool_max_threads() * (cand + 2) / 8 : host_pool_max_threads() * (cand + 1) / 4;
      host_pool_set_threads(n < 1 ? 1 : n);
      if (g_hc->tune_call & 3) tune_slot = cand;
      g_hc->tune_call++;
    } else if (g_hc->tune_call == 4 * ncand) {
      host_pool_set_threads(g_hc->tune_best ? g_hc->tune_best : (host_pool_max_threads() + 1) / 2);
      g_hc->tune_call++;
    }
  }
  static int nsl_env = -1;
  if (nsl_env < 0) {
    const char *ex = getenv("RBG_HOST_IO_SLICES");
    nsl_env = ex ? atoi(ex) : 0;
    if (nsl_env < 0 || nsl_env > 16) nsl_env = 0;
  }
  // 16 slices for big batches: the host starts widening after 1/16 of the copies and ends 1/16 after them (measured at
  // 65 536 envs 10x10/5: 8 slices 1.27-1.6 ms per step, 16 slices 1.17-1.24 ms)
  const int nsl = B >= 4096 ? (nsl_env ? nsl_env : (packed && B >= 32768 ? 16 : 8)) : 1;
  const int64_t sl = slice_size(B, nsl);
  const int64_t nslices = (B + sl - 1) / sl;
  const size_t obs_n = (size_t)B * N * G * G;
  Carver size{nullptr};
  rbg_timestep dt;
  carve_timestep(size, B, G, N, &dt);
  size.take<int32_t>((size_t)B * N);
  const size_t ws_bytes = (size_t)rbg_step_workspace_bytes(sl, G, N);
  size.take<uint8_t>((size_t)nslices * ws_bytes);
  if (packed) size.take<uint8_t>(obs_n);
  void *base;
  if ((rc = scratch_get(size.off + 256, dev, &base))) return rc;
  Carver c{reinterpret_cast<uint8_t *>(base)};
  carve_timestep(c, B, G, N, &dt);
  int32_t *da = c.take<int32_t>((size_t)B * N);
  uint8_t *ws = c.take<uint8_t>((size_t)nslices * ws_bytes);
  uint8_t *obs8 = packed ? c.take<uint8_t>(obs_n) : nullptr;
  scratch_claim(ScratchSig{4, B, G, N, params->autoreset_kind, base}, ws, ws_bytes, nslices);
  if (packed && g_hc->stage8_bytes < obs_n) {
    if (g_hc->stage8) cudaFreeHost(g_hc->stage8);
    g_hc->stage8 = nullptr;
    g_hc->stage8_bytes = 0;
    if ((e = cudaHostAlloc(reinterpret_cast<void **>(&g_hc->stage8), obs_n, cudaHostAllocDefault)) != cudaSuccess) return set_cuda_error(e, "cudaHostAlloc(observation staging)");
    g_hc->stage8_bytes = obs_n;
  }
  // One compute stream runs the slices' kernels back to back (the whole batch is ~0.1 ms of kernels); the two
  // copy streams take the observation (96 % of the bytes) of each slice as soon as its kernel is done, so the
  // bus is busy from the first slice on.  The small leaves go once for the whole batch: 8 copies instead of 8 per
  // slice.  Packed transport (default): the observation's codes cross the bus as bytes into a pinned staging
  // buffer and the host pool widens slice s into the caller's int32 buffer while slices s+1.. are in flight; the
  // caller's observation buffer need not be pinned.
  cudaEvent_t *slice_ev = g_hc->slice_ev, *copy_ev = g_hc->copy_ev;
  cudaStream_t cs = g_streams[0];
  // The State was produced by the caller on the device's default stream (or a stream that stream synchronises
  // with: every blocking stream does): our compute stream is ordered behind it by an event, the host does not wait.
  if (!g_hc->entry_ev && (e = cudaEventCreateWithFlags(&g_hc->entry_ev, cudaEventDisableTiming)) != cudaSuccess) return set_cuda_error(e, "cudaEventCreate");
  if ((e = cudaEventRecord(g_hc->entry_ev, cudaStreamLegacy)) != cudaSuccess) return set_cuda_error(e, "cudaEventRecord(entry)");
  if ((e = cudaStreamWaitEvent(cs, g_hc->entry_ev, 0)) != cudaSuccess) return set_cuda_error(e, "cudaStreamWaitEvent(entry)");
  RBG_CPY(da, action, (size_t)B * N * 4, cudaMemcpyHostToDevice, cs);
  g_h2d_bytes.fetch_add((long long)B * N * 4, std::memory_order_relaxed);
  const size_t per_env = (size_t)N * G * G;
  // Observation codes are <= 3 N: with at most 5 agents they fit a nibble, two cells per byte over the bus (a slice is a
  // multiple of 64 envs, so every slice starts on a whole, 16-byte aligned byte when an env has an even number of cells).
  // RBG_HOST_IO_BITS=8 keeps a byte per cell.
  static const int bits_env = env_int("RBG_HOST_IO_BITS");
  const int obits = (packed && 3 * N <= 15 && (per_env & 1u) == 0 && bits_env != 8) ? 4 : 8;
  const int oshift = obits == 4 ? 1 : 0;
  // Mixed transport (off by default, see host_io_split_env): `nwide` of the step's slices go over the bus as int32, the
  // rest as bytes.
  bool mix = false;
  if (packed && nslices > 1 && host_io_split_env() == -1) {
    cudaPointerAttributes pa;
    if (cudaPointerGetAttributes(&pa, ts->obs_grid) == cudaSuccess && pa.type == cudaMemoryTypeHost) mix = true;
    (void)cudaGetLastError();
  }
  const int64_t mix_key = B * 4096 + (int64_t)G * 64 + N;
  if (g_hc->wide_B != mix_key) {
    g_hc->wide_B = mix_key;
    g_hc->wide_slices = host_io_split_init((int)nslices);
  }
  const int nwide = !packed ? (int)nslices : (mix || host_io_fixed_split()) ? (g_hc->wide_slices > (int)nslices ? (int)nslices : g_hc->wide_slices) : 0;
  // evenly spread, the first slice first and never the last one (its widening is the host's tail either way)
  static const int split_pos = env_int("RBG_HOST_IO_SPLIT_POS");  // experiment: 1 = evenly spread from the first slice on, 2 = the first ones
  auto slice_is_wide = [&](int i) {
    if (!packed) return true;
    if (split_pos == 1) return ((int64_t)i * nwide) % nslices < nwide;
    if (split_pos == 2) return i < nwide;
    return i >= (int)nslices - nwide;  // the last ones: the host threads get their bytes first
  };
  size_t packed_cells = 0, wide_cells = 0;
  int si = 0;
  for (int64_t off = 0; off < B; off += sl, ++si) {
    const int64_t n = (B - off) < sl ? (B - off) : sl;
    const bool wide = slice_is_wide(si);
    rbg_state dss = state_at(*state, off, G, N);
    rbg_timestep dts = timestep_at(dt, off, G, N);
    if ((rc = rbg_connector_step(&dss, &dss, da + off * N, n, G, N, params, &dts, ws + (size_t)si * ws_bytes, cs))) return rc;
    if (!wide && (rc = launch_narrow_codes(dts.obs_grid, obs8 + (((size_t)off * per_env) >> oshift), (int64_t)((size_t)n * per_env), cs, obits))) return rc;
    if (!slice_ev[si] && (e = cudaEventCreateWithFlags(&slice_ev[si], cudaEventDisableTiming)) != cudaSuccess) return set_cuda_error(e, "cudaEventCreate");
    if ((e = cudaEventRecord(slice_ev[si], cs)) != cudaSuccess) return set_cuda_error(e, "cudaEventRecord");
    cudaStream_t xs = g_streams[1 + (si & 1)];
    if ((e = cudaStreamWaitEvent(xs, slice_ev[si], 0)) != cudaSuccess) return set_cuda_error(e, "cudaStreamWaitEvent");
    (wide ? wide_cells : packed_cells) += (size_t)n * per_env;
    if (!wide) {
      RBG_CPY(g_hc->stage8 + (((size_t)off * per_env) >> oshift), obs8 + (((size_t)off * per_env) >> oshift), ((size_t)n * per_env) >> oshift, cudaMemcpyDeviceToHost, xs);
      if (!copy_ev[si] && (e = cudaEventCreateWithFlags(&copy_ev[si], cudaEventDisableTiming | (host_io_blocking_sync() ? cudaEventBlockingSync : 0))) != cudaSuccess)
        return set_cuda_error(e, "cudaEventCreate");
      if ((e = cudaEventRecord(copy_ev[si], xs)) != cudaSuccess) return set_cuda_error(e, "cudaEventRecord(copy)");
    } else if ((rc = copy_timestep(ts, &dt, off, n, G, N, cudaMemcpyDeviceToHost, off, xs, 1))) {
      return rc;
    }
  }
  if ((rc = copy_timestep(ts, &dt, 0, B, G, N, cudaMemcpyDeviceToHost, 0, cs, 2))) return rc;
  const auto t_issued = std::chrono::steady_clock::now();
  g_d2h_bytes.fetch_add((long long)((packed_cells >> oshift) + 4 * wide_cells) + (long long)B * (N * 5 + 4 + N * 8 + 1 + 12), std::memory_order_relaxed);
  if (packed) {
    si = 0;
    for (int64_t off = 0; off < B; off += sl, ++si) {
      const int64_t n = (B - off) < sl ? (B - off) : sl;
      if (slice_is_wide(si)) continue;
      if ((e = cudaEventSynchronize(copy_ev[si])) != cudaSuccess) {
        host_pool_wait();
        return set_cuda_error(e, "cudaEventSynchronize(copy)");
      }
      if (obits == 4)
        host_pool_widen4(g_hc->stage8 + (((size_t)off * per_env) >> 1), ts->obs_grid + (size_t)off * per_env, (size_t)n * per_env);
      else
        host_pool_widen(g_hc->stage8 + (size_t)off * per_env, ts->obs_grid + (size_t)off * per_env, (size_t)n * per_env);
    }
  }
  int rc_sync = RBG_OK;
  for (int i = 0; i < 3; ++i) {
    e = cudaStreamSynchronize(g_streams[i]);
    if (e != cudaSuccess) rc_sync = set_cuda_error(e, "cudaStreamSynchronize");
  }
  const auto t_bus_done = std::chrono::steady_clock::now();
  if (packed) host_pool_wait();
  {  // RBG_HOST_IO_TRACE=1: where a call's time goes (issue of the launches and copies / until the last copy landed / until
     // the last slice is widened), averaged over 32 calls, on stderr
    static const int trace = env_int("RBG_HOST_IO_TRACE");
    if (trace) {
      static double acc[3] = {0, 0, 0};
      static int ncalls = 0;
      const auto t_now = std::chrono::steady_clock::now();
      acc[0] += std::chrono::duration<double>(t_issued - t_call0).count();
      acc[1] += std::chrono::duration<double>(t_bus_done - t_call0).count();
      acc[2] += std::chrono::duration<double>(t_now - t_call0).count();
      if (++ncalls == 32) {
        fprintf(stderr, "[rbg host_io] threads %d wide slices %d/%d: issued %.3f ms, copies landed %.3f ms, widened %.3f ms\n", host_pool_threads(), nwide, (int)nslices,
                acc[0] / 32 * 1e3, acc[1] / 32 * 1e3, acc[2] / 32 * 1e3);
        acc[0] = acc[1] = acc[2] = 0;
        ncalls = 0;
      }
    }
  }
  if (mix && rc_sync == RBG_OK) {
    // who finished last?  The host threads still busy well after the last copy landed: the host is the bound, one more
    // slice goes over the bus as int32.  The pool idle when the copies ended (their last slice is widened within a
    // twentieth of the call): the bus is the bound, one slice fewer.
    const auto t_end = std::chrono::steady_clock::now();
    const double total = std::chrono::duration<double>(t_end - t_call0).count();
    const double host_tail = std::chrono::duration<double>(t_end - t_bus_done).count();
    // (the widening of the last byte slice always trails the copies by about 1 / nslices of the call: the dead band)
    if (!in_tuning) {
      if (host_tail > 1.6 / (double)nslices * total && g_hc->wide_slices < (int)nslices - 1)
        g_hc->wide_slices++;
      else if (host_tail < 0.6 / (double)nslices * total && g_hc->wide_slices > 0)
        g_hc->wide_slices--;
    }
  }
  if (tune_slot >= 0 && rc_sync == RBG_OK) {
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_call0).count();
    // the fastest call decides; a larger thread count must beat a smaller one by 1 % to replace it (the candidates of a
    // single-rank host stop at half of the cores, where more threads still help: 6 / 8 threads 65.4 / 68.1 M env-steps/s)
    const int nthr = host_pool_threads();
    if (g_hc->tune_best == 0 || (nthr == g_hc->tune_best ? dt < g_hc->tune_best_s : dt < 0.99 * g_hc->tune_best_s)) {
      g_hc->tune_best = nthr;
      g_hc->tune_best_s = dt;
    }
  }
  return rc_sync;
}

int rbg_host_widen(const uint8_t *src, int32_t *dst, int64_t n) {
  if (n < 0 || (n > 0 && (!src || !dst))) return set_error(RBG_EINVAL, "rbg_host_widen: n=%lld, src=%p, dst=%p", (long long)n, (const void *)src, (void *)dst);
  if (n == 0) return RBG_OK;
  host_pool_widen(src, dst, (size_t)n);
  host_pool_wait();
  return RBG_OK;
}

int rbg_host_widen4(const uint8_t *src, int32_t *dst, int64_t n) {
  if (n < 0 || (n & 1) || (n > 0 && (!src || !dst))) return set_error(RBG_EINVAL, "rbg_host_widen4: n=%lld (even), src=%p, dst=%p", (long long)n, (const void *)src, (void *)dst);
  if (n == 0) return RBG_OK;
  host_pool_widen4(src, dst, (size_t)n);
  host_pool_wait();
  return RBG_OK;
}

int rbg_host_transfer_stats(int64_t *h2d_bytes, int64_t *d2h_bytes, int *host_threads, int reset) {
  if (h2d_bytes) *h2d_bytes = reset ? g_h2d_bytes.exchange(0) : g_h2d_bytes.load();
  if (d2h_bytes) *d2h_bytes = reset ? g_d2h_bytes.exchange(0) : g_d2h_bytes.load();
  if (host_threads) *host_threads = host_io_wide() ? 0 : host_pool_threads();
  return RBG_OK;
}

}  // extern "C"
