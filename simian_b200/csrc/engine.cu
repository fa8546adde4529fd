// engine.cu -- host side of libsimian_b200.so: the C-ABI of include/simian_b200.h
// over the kernels of kernels.cuh.  No CPU fallback: without a CUDA device
// simian_create fails with SIMIAN_ERR_NO_DEVICE.
//
// Replaces the JS engine object of SimianJS/simian.js (constructor :36-84,
// run :99-168, schedService :170-190, addEntity :211-229, attachService
// :207-209) and Entity.reqService (SimianJS/entity.js:35-68, on the device).
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <chrono>
#include <map>
#include <mutex>
#include <random>
#include <string>
#include <vector>

#include "../../include/simian_b200.h"
#include "kernels.cuh"

#ifndef SIMIAN_SPLIT_SUMS_DEFAULT
#define SIMIAN_SPLIT_SUMS_DEFAULT 1 // option "split_sums" when the model does not say
#endif
#ifndef SIMIAN_SPLIT_DEFAULT
#define SIMIAN_SPLIT_DEFAULT 0 // option "split_handlers" when the model does not say
#endif

extern "C" void simian_trim_cache(void);

// window_loop.cu: the persistent window loop, compiled in its own translation unit
int simian_loop_blocks_per_sm(size_t smem_bytes);
cudaError_t simian_loop_launch(const DevCfg &cfg, uint32_t win_ix, uint32_t max_windows,
                               size_t smem_bytes, cudaStream_t stream);
size_t simian_loop_smem_bytes(uint32_t epb);

namespace {

std::string g_create_error;

// Process-wide cache of device allocations, keyed by (device, bytes): an engine
// built with the same shape as a destroyed one reuses its HBM instead of paying
// cudaFree + cudaMalloc of multi-GB pools (tens of ms per engine).  Everything
// an engine reads is (re)initialised by k_init / setup, never assumed zero.
struct CacheKey {
  int device;
  size_t bytes;
  bool operator<(const CacheKey &o) const {
    return device != o.device ? device < o.device : bytes < o.bytes;
  }
};
std::multimap<CacheKey, void *> g_alloc_cache;
size_t g_alloc_cache_bytes = 0;
std::mutex g_alloc_cache_mu;

// pinned host bounce buffers for result read-back (pageable D2H runs at a few GB/s)
std::multimap<size_t, void *> g_pinned_cache;

void *pinned_take(size_t bytes) {
  {
    std::lock_guard<std::mutex> g(g_alloc_cache_mu);
    auto it = g_pinned_cache.lower_bound(bytes);
    if (it != g_pinned_cache.end() && it->first <= 2 * bytes + 4096) {
      void *p = it->second;
      g_pinned_cache.erase(it);
      return p;
    }
  }
  void *p = nullptr;
  if (cudaMallocHost(&p, bytes) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return p;
}

void pinned_give(size_t bytes, void *p) {
  std::lock_guard<std::mutex> g(g_alloc_cache_mu);
  if (g_pinned_cache.size() >= 8) {
    cudaFreeHost(p);
    return;
  }
  g_pinned_cache.emplace(bytes, p);
}

// CUDA-IPC mappings of peers' buffers, keyed by the 64 handle bytes: engines of
// the same shape reuse the same device buffers (allocation cache), so their
// handles repeat and cudaIpcOpenMemHandle (~ms each) is paid once per peer
std::map<std::string, void *> g_ipc_cache;

cudaError_t ipc_open_cached(void **out, const cudaIpcMemHandle_t &h) {
  std::string key((const char *)&h, sizeof h);
  {
    std::lock_guard<std::mutex> g(g_alloc_cache_mu);
    auto it = g_ipc_cache.find(key);
    if (it != g_ipc_cache.end()) {
      *out = it->second;
      return cudaSuccess;
    }
  }
  cudaError_t ce = cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess);
  if (ce == cudaSuccess) {
    std::lock_guard<std::mutex> g(g_alloc_cache_mu);
    g_ipc_cache[key] = *out;
  }
  return ce;
}

size_t alloc_cache_limit() {
  const char *s = getenv("SIMIAN_B200_CACHE_GB");
  double gb = s ? atof(s) : 64.0;
  return (size_t)(gb * 1073741824.0);
}

void *cache_take(int device, size_t bytes) {
  std::lock_guard<std::mutex> g(g_alloc_cache_mu);
  auto it = g_alloc_cache.find(CacheKey{device, bytes});
  if (it == g_alloc_cache.end()) return nullptr;
  void *p = it->second;
  g_alloc_cache.erase(it);
  g_alloc_cache_bytes -= bytes;
  return p;
}

bool cache_give(int device, size_t bytes, void *p) {
  std::lock_guard<std::mutex> g(g_alloc_cache_mu);
  if (g_alloc_cache_bytes + bytes > alloc_cache_limit()) return false;
  g_alloc_cache.emplace(CacheKey{device, bytes}, p);
  g_alloc_cache_bytes += bytes;
  return true;
}

struct HostClass {
  std::string name;
  uint32_t first_id = 0, count = 0, state_bytes = 0;
  double name_hash = 0;
  int base_rank_override = -1; // simian_set_base_rank: a user-defined getBaseRank
  std::vector<uint8_t> params;
  std::vector<uint8_t> init_state; // count * state_bytes when given
  bool has_init = false;
  uint8_t fn_of_service[SIMIAN_MAX_SERVICES];
  // resolved at setup
  uint32_t base_rank = 0, rank_off = 0, slot_base = 0, n_local = 0;
  uint8_t *d_state = nullptr;
  void *d_params = nullptr;
};

struct SchedOp { // a recorded schedService call, replayed on the device at run
  int kind = 0;  // 0 bulk, 1 host records, 2 entity constructors
  double time = 0;
  uint32_t handler = 0;
  uint64_t data = 0;
  int cls = 0;
  uint32_t first = 0, count = 0, repeat = 0, seq_base = 0;
  simian_event_t *d_recs = nullptr; // kind 1: uploaded at call time
  uint64_t n_recs = 0, d_bytes = 0;
  bool ingested = false; // kind 1: already in the calendar (early layout)
};

int builtin_fn_by_name(const char *n) { return simian_fn_by_name(n); } // simian_handlers.def

} // namespace

struct simian_engine {
  std::string name;
  double start, end, min_delay, inf_time;
  int rank, size, device;
  // options
  uint64_t seed = 1;
  double bucket_width = 0;
  uint32_t ring = 1024;
  uint64_t pool_events = 0, outbox_events = 0;
  int launch_batch = 16, tie_check = 1, force_sequential = 0;
  uint32_t epb_override = 0; // option "block_entities"
  int pdl = 1;               // option "pdl": window launches overlap the previous one's drain
  // option "persistent_loop": the window loop inside one cooperative launch when
  // every block's CTA is resident at once.  Off by default: measured (B200, r2_h) a
  // window of a small model costs ~18.5 us with either driver -- it is bound by its
  // chain of dependent memory round trips, not by the launch (PDL hides that)
  int loop = 0;
  int deferred_pull = 0; // option "deferred_pull": two-ended outboxes (size > 1, device-driven exchange)
  // option "split_handlers": a window is two launches -- k_window_stage (dequeue, rank,
  // trace chain, recycle) and k_emit (handlers + appends at a higher occupancy).  Pure
  // handlers only; -1 = the default for the model (see setup_layout)
  int split = -1;
  // option "pdl_exchange": the three launches of a device-driven multi-GPU window
  // (k_window -> k_pull -> k_advance) are programmatic dependents of each other
  int pdl_exchange = 0;
  // option "split_sums": the sums stateful handlers defer (simian_defer_weighted_sum, the
  // LANL receive handler) are computed by k_sums after the window kernel instead of
  // inside it; -1 = the default for the model
  int split_sums = -1;
  uint64_t job_cap = 0; // option "sum_jobs": descriptors per window (0: four per entity)
  uint32_t sums_ctas = 888; // grid of k_sums: the CTAs resident at once
  void *hot_base = nullptr; // the hot-metadata slab (setup_layout) and its size
  size_t hot_bytes = 0;
  int fused_advance = -1; // option "fused_advance": the last CTA of k_window does the LBTS step (-1: by size)
  bool l2_applied = false;
  int l2_persist = 2; // option "l2_persist": persisting L2 window over the cell head words
  int payloads = 0; // option "payloads": handlers may send / read payload words (continuation records)
  bool split_on = false; // in force for this layout
  uint32_t rest_ctas = 444; // grid of k_window_rest: the CTAs resident at once
  uint32_t emit_ctas = 740; // grid of k_emit: the CTAs resident at once
  bool loop_ok = false;      // every entity block's CTA is resident at once (set at layout)
  // model
  std::vector<HostClass> classes;
  std::map<std::string, int> class_ix;
  std::map<std::string, int> service_ix;
  std::vector<SchedOp> ops;
  uint32_t n_ids = 0, sched_seq = 0;
  uint64_t initial_events = 0;
  uint64_t record_events = 0; // the part of initial_events that came as host records
  // state
  bool running = false, set_up = false, stepping = false;
  bool laid_out = false;            // HBM layout allocated + initialised (may precede set_up)
  uint64_t laid_out_initial = 0;    // initial_events the pool was sized for
  cudaStream_t copy_stream = nullptr; // H2D of simian_sched_events, overlapped with ingest
  std::string err;
  int err_code = 0;
  // device
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  DevCfg cfg;
  std::vector<std::pair<void *, size_t>> allocs;
  uint32_t win_ix = 0;
  uint64_t launches = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  unsigned long long *d_digest = nullptr;
  DevAcc *h_acc = nullptr; // pinned
  WinState *h_win = nullptr;
  uint32_t *h_counts = nullptr;
  std::chrono::steady_clock::time_point t_begin;
  size_t smem_bytes = WINDOW_SMEM_BYTES(SIMIAN_EPB);
  void *ipc_slab = nullptr; // outbox_count at +0, exchange slots at +4096
  uint64_t salt = 0;
  bool peer_seen[SIMIAN_MAX_RANKS] = {};

  int fail(int code, const std::string &m) {
    if (!err_code) {
      err_code = code;
      err = m;
    }
    return code;
  }
};

#define CK(call)                                                                       \
  do {                                                                                 \
    cudaError_t _e = (call);                                                           \
    if (_e != cudaSuccess)                                                             \
      return e->fail(SIMIAN_ERR_CUDA, std::string(#call) + " (engine.cu:" +              \
                                          std::to_string(__LINE__) + "): " +                   \
                                          cudaGetErrorString(_e));                             \
  } while (0)

namespace {

const char *status_text(int code) {
  switch (code) {
    case SIMIAN_ERR_LOOKAHEAD: return "attempted to send with too little delay";
    case SIMIAN_ERR_IDLE_SEND: return "Cannot send an event when Simian is idle";
    case SIMIAN_ERR_OUT_OF_ORDER: return "Out of order event";
    case SIMIAN_ERR_RUNNING: return "Cannot add new entity when Simian is running!";
    case SIMIAN_ERR_UNKNOWN_NAME: return "unknown entity / service / device function";
    case SIMIAN_ERR_POOL_EXHAUSTED: return "event pool exhausted: raise option pool_events";
    case SIMIAN_ERR_WINDOW_OVERFLOW: return "too many due pages for one entity block in one window";
    case SIMIAN_ERR_ENTITY_BURST: return "one entity has too many due events in one window";
    case SIMIAN_ERR_SELF_PENDING: return "too many same-window self-sends pending on one entity";
    case SIMIAN_ERR_OUTBOX_FULL: return "cross-GPU outbox full: raise option outbox_events";
    case SIMIAN_ERR_RING_SPAN: return "window spans too many calendar sub-buckets: raise option bucket_width";
    case SIMIAN_ERR_STALLED: return "a device-side wait timed out (page install or peer exchange flag)";
    case SIMIAN_ERR_PAYLOAD: return "event payload: word missing, too long, or not available on this path";
    default: return "error";
  }
}

void trim_alloc_cache() { // cached device allocations only
  std::lock_guard<std::mutex> g(g_alloc_cache_mu);
  int cur = 0;
  cudaGetDevice(&cur);
  for (auto &kv : g_alloc_cache) {
    cudaSetDevice(kv.first.device);
    cudaFree(kv.second);
  }
  cudaSetDevice(cur);
  g_alloc_cache.clear();
  g_alloc_cache_bytes = 0;
}

template <class T>
int dev_alloc(simian_engine *e, T **p, size_t n) {
  size_t bytes = (n ? n : 1) * sizeof(T);
  bytes = (bytes + 511) & ~(size_t)511;
  void *q = cache_take(e->device, bytes);
  if (!q) {
    cudaError_t ce = cudaMalloc(&q, bytes);
    if (ce != cudaSuccess) { // make room: drop the cached ALLOCATIONS and retry once
      // (not the CUDA-IPC mappings of peers' buffers: live engines of this process
      // may still be running kernels that read through them)
      cudaGetLastError();
      trim_alloc_cache();
      CK(cudaMalloc(&q, bytes));
    }
  }
  e->allocs.push_back({q, bytes});
  *p = (T *)q;
  return 0;
}

// The parameter blob of a class must be what its attached device function
// reads (the reference throws inside the handler on a bad model; here a bad
// blob would index past HBM buffers).  No blob at all: the device copy is zeros.
int check_class_params(simian_engine *e, const HostClass &h, int fn) {
  const size_t need = simian_fn_params_bytes(fn); // the struct the handler reads (registry)
  if (need <= sizeof(simian_no_params_t)) return 0;
  if (h.params.empty()) return 0;
  if (h.params.size() < need)
    return e->fail(SIMIAN_ERR_BAD_ARGUMENT,
                   "class " + h.name + ": parameter blob smaller than its device function reads");
  if (fn == SIMIAN_FN_PHOLD_GENERATE) {
    simian_phold_params_t p;
    memcpy(&p, h.params.data(), sizeof p);
    if (p.n_entities == 0 || (uint64_t)p.first_id + p.n_entities > e->n_ids)
      return e->fail(SIMIAN_ERR_BAD_ARGUMENT, "phold: first_id + n_entities exceeds the entities added");
    if (p.mode == SIMIAN_PHOLD_OTHER_GROUP && (p.groups < 2 || p.n_entities < p.groups))
      return e->fail(SIMIAN_ERR_BAD_ARGUMENT, "phold OTHER_GROUP needs 2 <= groups <= n_entities");
  }
  return 0;
}

uint32_t local_count(uint32_t count, uint32_t rank_off, uint32_t g) {
  return count > rank_off ? (count - rank_off + g - 1) / g : 0;
}

// allocate the HBM layout, initialise it, replay schedService, set up window 0
// give everything the layout allocated back (a model-changing call arrived
// after an early layout): the scheduled records are still on the device
void teardown_layout(simian_engine *e) {
  if (!e->laid_out || e->set_up) return;
  cudaStreamSynchronize(e->stream);
  for (auto &a : e->allocs)
    if (!cache_give(e->device, a.second, a.first)) cudaFree(a.first);
  e->allocs.clear();
  for (SchedOp &op : e->ops) op.ingested = false;
  for (HostClass &h : e->classes) {
    h.d_state = nullptr;
    h.d_params = nullptr;
  }
  e->d_digest = nullptr;
  e->hot_base = nullptr;
  e->hot_bytes = 0;
  e->laid_out = false;
}

int setup_layout(simian_engine *e) {
  if (e->laid_out) return 0;
  DevCfg &c = e->cfg;
  CK(cudaFuncSetAttribute(k_window, cudaFuncAttributeMaxDynamicSharedMemorySize,
                          (int)WINDOW_SMEM_BYTES(SIMIAN_EPB_MAX)));
  CK(cudaFuncSetAttribute(k_window_blobs, cudaFuncAttributeMaxDynamicSharedMemorySize,
                          (int)WINDOW_SMEM_BYTES(SIMIAN_EPB_MAX)));
  CK(cudaFuncSetAttribute(k_local_min, cudaFuncAttributeMaxDynamicSharedMemorySize,
                          (int)WINDOW_SMEM_BYTES(SIMIAN_EPB_MAX)));
  CK(cudaFuncSetAttribute(k_window_stage, cudaFuncAttributeMaxDynamicSharedMemorySize,
                          (int)WINDOW_SMEM_BYTES(SIMIAN_EPB_MAX)));
  CK(cudaFuncSetAttribute(k_window_rest, cudaFuncAttributeMaxDynamicSharedMemorySize,
                          (int)WINDOW_SMEM_BYTES(SIMIAN_EPB_MAX)));
  memset(&c, 0, sizeof c);
  c.start = e->start;
  c.end = e->end;
  c.min_delay = e->min_delay;
  c.inf_time = e->inf_time;
  c.seed = e->seed;
  c.rank = e->rank;
  c.size = e->size;
  c.tie_check = e->tie_check;
  c.n_classes = (int)e->classes.size();
  c.n_ids = e->n_ids;
  c.all_pure = e->force_sequential ? 0 : 1;
  if (c.n_classes == 0) return e->fail(SIMIAN_ERR_BAD_ARGUMENT, "no entities added");
  uint32_t g = (uint32_t)e->size, slot = 0;
  for (int k = 0; k < c.n_classes; k++) {
    HostClass &h = e->classes[k];
    h.base_rank = h.base_rank_override >= 0
                      ? (uint32_t)h.base_rank_override % (uint32_t)e->size
                      : (uint32_t)fmod(h.name_hash, (double)e->size); // simian.js:192-194
    h.rank_off = ((uint32_t)e->rank + g - h.base_rank % g) % g;
    h.n_local = local_count(h.count, h.rank_off, g);
    h.slot_base = slot;
    slot += h.n_local;
    DevClass &d = c.cls[k];
    d.first_id = h.first_id;
    d.count = h.count;
    d.rank_off = h.rank_off;
    d.slot_base = h.slot_base;
    d.n_local = h.n_local;
    d.base_rank = h.base_rank;
    d.state_bytes = h.state_bytes;
    d.params_bytes = (uint32_t)h.params.size();
    memcpy(d.fn_of_service, h.fn_of_service, sizeof d.fn_of_service);
    for (int sidx = 0; sidx < SIMIAN_MAX_SERVICES; sidx++) {
      const int fn = h.fn_of_service[sidx];
      if (!fn) continue;
      if (!simian_fn_is_pure(fn)) c.all_pure = 0;
      if (int rc = check_class_params(e, h, fn)) return rc;
    }
    if (dev_alloc(e, &h.d_state, (size_t)h.n_local * h.state_bytes)) return e->err_code;
    if (h.state_bytes && h.n_local) {
      if (h.has_init) {
        std::vector<uint8_t> loc((size_t)h.n_local * h.state_bytes);
        for (uint32_t j = 0; j < h.n_local; j++)
          memcpy(&loc[(size_t)j * h.state_bytes],
                 &h.init_state[(size_t)(j * g + h.rank_off) * h.state_bytes], h.state_bytes);
        CK(cudaMemcpyAsync(h.d_state, loc.data(), loc.size(), cudaMemcpyHostToDevice, e->stream));
        CK(cudaStreamSynchronize(e->stream));
      } else {
        CK(cudaMemsetAsync(h.d_state, 0, (size_t)h.n_local * h.state_bytes, e->stream));
      }
    }
    d.state = h.d_state;
    // the device copy is at least as large as any built-in parameter struct and
    // zero-filled beyond what was given: a handler never reads recycled memory
    uint8_t *dp = nullptr;
    const size_t pbytes = h.params.size() > 256 ? h.params.size() : 256;
    if (dev_alloc(e, &dp, pbytes)) return e->err_code;
    CK(cudaMemsetAsync(dp, 0, pbytes, e->stream));
    if (!h.params.empty())
      CK(cudaMemcpyAsync(dp, h.params.data(), h.params.size(), cudaMemcpyHostToDevice, e->stream));
    h.d_params = dp;
    d.params = dp;
  }
  c.n_slots = slot;
  // Entities per block (option "block_entities", <= SIMIAN_EPB).  Shrinking the
  // block so that the grid fills whole waves of resident CTAs was measured on
  // config 2 (480 instead of 512: 4.92 instead of 4.61 waves) and gained nothing:
  // CTAs do not run in lock-step waves, and smaller blocks pay more per-CTA overhead.
  c.epb = e->epb_override ? e->epb_override : EPB;
  if (c.epb > SIMIAN_EPB_MAX || c.epb < 1)
    return e->fail(SIMIAN_ERR_BAD_ARGUMENT, "bad block_entities");
  e->smem_bytes = WINDOW_SMEM_BYTES(c.epb);
  c.epb_inv = ~0ull / c.epb + 1ull; // ceil(2^64 / epb): x / epb == umul64hi(x, epb_inv) for x < 2^32
  c.n_blocks = (slot + c.epb - 1) / c.epb;
  if (c.n_blocks == 0) c.n_blocks = 1;
  double w = e->bucket_width > 0 ? e->bucket_width : e->min_delay / 8.0;
  c.inv_w = 1.0 / w;
  uint32_t ring = 256;
  while (ring < e->ring) ring <<= 1;
  c.ring = ring;
  c.ring_mask = ring - 1;
  // events this rank will hold at the start: script loops (schedService /
  // constructors) run on every rank and only owners keep their share; host
  // records may already be the rank's own (all counted: an upper bound)
  uint64_t local_initial = (e->initial_events - e->record_events) / g + e->record_events + 1;
  // default pool: twice the initial events + one (partial) page for every cell
  // of the ring, the upper bound on simultaneously open pages
  uint64_t pool_events = e->pool_events
                             ? e->pool_events
                             : 2 * local_initial +
                                   (uint64_t)PAGE * ((uint64_t)(ring + 2) * c.n_blocks + 1024ull +
                                                     (uint64_t)SIMIAN_BC * c.n_blocks);
  // The safe default (one page per cell) grows with entities x ring: 9.5 GB for
  // 2^20 entities, 152 GB for 2^24.  Beyond 16 GB the default becomes ten times the
  // initial records (measured need of the PHOLD configs: ~7.5x)
  // or 16 GB, whichever is larger; SIMIAN_ERR_POOL_EXHAUSTED is loud, "pool_events"
  // overrides.
  if (!e->pool_events) {
    const uint64_t cap16 = (16ull << 30) / sizeof(simian_event_t);
    if (pool_events > cap16) {
      uint64_t heur = 10 * local_initial + (uint64_t)PAGE * (2ull * SIMIAN_BC * c.n_blocks + 1024ull);
      if (heur < cap16) heur = cap16;
      if (heur < pool_events) pool_events = heur;
    }
  }
  uint64_t n_pages = (pool_events + PAGE - 1) / PAGE;
  if (n_pages > 0xFFFFFFFFull / PAGE) n_pages = 0xFFFFFFFFull / PAGE;
  c.n_pages = (uint32_t)n_pages;
  size_t n_cells = (size_t)(ring + 2) * c.n_blocks;
  // the hot metadata -- cell head words (the target of every append's claim), entity
  // core states, block page stacks -- in ONE allocation, so that one persisting L2
  // access window covers it (option "l2_persist")
  {
    auto up = [](size_t b) { return (b + 511) & ~(size_t)511; };
    const size_t b_cells = up(n_cells * sizeof(unsigned long long));
    const size_t b_core = up((size_t)c.n_slots * sizeof(EntityCore));
    const size_t b_bc = up((size_t)c.n_blocks * SIMIAN_BC * sizeof(uint32_t));
    const size_t b_cnt = up((size_t)c.n_blocks * sizeof(uint32_t));
    const size_t b_snap = up((size_t)c.n_blocks * sizeof(unsigned long long));
    uint8_t *slab = nullptr;
    e->hot_bytes = b_cells + b_core + b_bc + b_cnt + b_snap;
    if (dev_alloc(e, &slab, e->hot_bytes)) return e->err_code;
    e->hot_base = slab;
    c.cell_state = (unsigned long long *)slab;
    c.core = (EntityCore *)(slab + b_cells);
    c.bcache = (uint32_t *)(slab + b_cells + b_core);
    c.bcount = (uint32_t *)(slab + b_cells + b_core + b_bc);
    c.snap = (unsigned long long *)(slab + b_cells + b_core + b_bc + b_cnt);
  }
  if (dev_alloc(e, &c.pool, (size_t)c.n_pages * PAGE)) return e->err_code;
  if (dev_alloc(e, &c.page_next, c.n_pages)) return e->err_code;
  if (dev_alloc(e, &c.free_ring, c.n_pages)) return e->err_code;
  if (dev_alloc(e, &c.win, 2)) return e->err_code;
  if (dev_alloc(e, &c.acc, 1)) return e->err_code;
  if (dev_alloc(e, &c.part, SIMIAN_PART_SLOTS)) return e->err_code;
  if (dev_alloc(e, &e->d_digest, 2)) return e->err_code;
  // split window: only models of pure handlers, not beside the deferred pull (its
  // pull CTAs ride in k_window) or the persistent loop
  c.blobs = e->payloads ? 1 : 0;
  e->split_on = c.all_pure && !e->deferred_pull && !e->loop && !e->payloads &&
                (e->split < 0 ? SIMIAN_SPLIT_DEFAULT != 0 : e->split != 0);
  c.stage = nullptr;
  c.stage_n = nullptr;
  if (e->split_on) {
    int sms = 0, per_sm = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, e->device);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_window_rest, NTHREADS,
                                                      e->smem_bytes) != cudaSuccess || per_sm < 1) {
      cudaGetLastError();
      per_sm = 1;
    }
    e->rest_ctas = (uint32_t)(per_sm * (sms > 0 ? sms : 148));
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_emit, NTHREADS, 0) != cudaSuccess ||
        per_sm < 1) {
      cudaGetLastError();
      per_sm = 1;
    }
    e->emit_ctas = (uint32_t)(per_sm * (sms > 0 ? sms : 148));
    if (const char *env = getenv("SIMIAN_B200_EMIT_CTAS")) e->emit_ctas = (uint32_t)atoi(env); // (tuning)
    if (dev_alloc(e, &c.stage, (size_t)c.n_blocks * NREFS)) return e->err_code;
    if (dev_alloc(e, &c.stage_n, c.n_blocks)) return e->err_code;
    CK(cudaMemsetAsync(c.stage_n, 0, sizeof(uint32_t) * (size_t)c.n_blocks, e->stream));
  }
  c.jobs = nullptr;
  c.job_cap = 0;
  if (!c.all_pure && !e->loop &&
      (e->split_sums < 0 ? SIMIAN_SPLIT_SUMS_DEFAULT != 0 : e->split_sums != 0)) {
    // descriptors of one window, all blocks together: by default four per entity (the
    // LANL benchmark parks about one per entity and window), at least 1 M
    uint64_t cap = e->job_cap ? e->job_cap : 4ull * c.n_slots;
    if (!e->job_cap && cap < (1ull << 20)) cap = 1ull << 20;
    if (cap > 0xFFFFFFF0ull) cap = 0xFFFFFFF0ull;
    c.job_cap = (uint32_t)cap;
    if (dev_alloc(e, &c.jobs, (size_t)c.job_cap)) return e->err_code;
    int sms = 0, per_sm = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, e->device);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sums, NTHREADS, 0) != cudaSuccess ||
        per_sm < 1) {
      cudaGetLastError();
      per_sm = 1;
    }
    e->sums_ctas = (uint32_t)(per_sm * (sms > 0 ? sms : 148));
  }
  if (e->size > 1) {
    // a window in which every local event sends one record to a uniformly chosen
    // peer (the t = 0 burst of PHOLD at 100 % remote), with a factor 2 of slack
    uint64_t cap = e->outbox_events ? e->outbox_events : 2 * local_initial / (g - 1) + 65536;
    c.outbox_cap = (uint32_t)cap;
    c.split_outbox = e->deferred_pull ? 1u : 0u;
    if (dev_alloc(e, &c.outbox, 2 * (size_t)e->size * cap)) return e->err_code;
    // the two small peer-visible arrays live in ONE allocation of their own, large
    // enough (2 MB) not to be carved out of a slab CUDA shares between small
    // cudaMallocs: an IPC handle names a whole allocation
    uint8_t *slab = nullptr;
    if (dev_alloc(e, &slab, (size_t)2 << 20)) return e->err_code;
    e->ipc_slab = slab;
    c.outbox_count = (uint32_t *)slab;
    c.xch = (MinSlot *)(slab + 4096);
    c.peer_xch[e->rank] = c.xch;
  }
  if (!e->h_acc) {
    // pinned mirrors of the device flags, from the process-wide pinned pool (a
    // cudaMallocHost / cudaFreeHost pair per engine costs about a millisecond)
    e->h_acc = (DevAcc *)pinned_take(sizeof(DevAcc));
    e->h_win = (WinState *)pinned_take(2 * sizeof(WinState));
    e->h_counts = (uint32_t *)pinned_take(sizeof(uint32_t) * SIMIAN_MAX_RANKS);
    if (!e->h_acc || !e->h_win || !e->h_counts) return e->fail(SIMIAN_ERR_CUDA, "cudaMallocHost failed");
    CK(cudaEventCreate(&e->ev0));
    CK(cudaEventCreate(&e->ev1));
  }
  // persistent window loop: only when every block's CTA fits on the GPU at once
  e->loop_ok = false;
  if (e->loop && e->size == 1 && !e->payloads) {
    int sms = 0, coop = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, e->device);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, e->device);
    const int per_sm = simian_loop_blocks_per_sm(simian_loop_smem_bytes(e->cfg.epb));
    e->loop_ok = coop && per_sm > 0 && (uint64_t)c.n_blocks <= (uint64_t)per_sm * (uint64_t)sms;
  }
  k_init<<<296, 256, 0, e->stream>>>(c, (uint32_t)n_cells);
  CK(cudaGetLastError());
  e->laid_out = true;
  e->laid_out_initial = e->initial_events;
  return 0;
}

int ingest_records(simian_engine *e, const simian_event_t *d, uint64_t n, cudaStream_t st) {
  if (!n) return 0;
  int grid = (int)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184);
  k_ingest<<<grid, 256, 0, st>>>(e->cfg, d, n, 1, 0, 0);
  CK(cudaGetLastError());
  return 0;
}

// option "l2_persist" (default 2): keep the hot metadata of the calendar resident in L2 --
// a persisting access-policy window on the engine's stream.  2: the whole hot slab (cell
// head words = the target of every append's claim, entity core states, block page
// stacks, row snapshot; 34 MB for config 2) when it fits the budget below, else as 1;
// 1: the head words only; 0: off.  Measured on config 2: 15.76 -> 15.32 (1) -> 15.08 ms (2)
// per step.  Best effort: a failure here is
// not a failure of the run.  The window is removed from the stream when the engine lets
// go of it (the stream may be the caller's).
void clear_l2_policy(simian_engine *e) {
  if (!e->l2_applied || !e->stream) return;
  cudaStreamAttrValue av = {};
  av.accessPolicyWindow.num_bytes = 0; // (disables the window)
  cudaStreamSetAttribute(e->stream, cudaStreamAttributeAccessPolicyWindow, &av);
  cudaCtxResetPersistingL2Cache();
  cudaGetLastError();
  e->l2_applied = false;
}

void apply_l2_policy(simian_engine *e) {
  if (e->l2_persist <= 0 || !e->hot_base || !e->stream) return;
  const DevCfg &c = e->cfg;
  int max_win = 0, l2 = 0;
  cudaDeviceGetAttribute(&max_win, cudaDevAttrMaxAccessPolicyWindowSize, e->device);
  cudaDeviceGetAttribute(&l2, cudaDevAttrMaxPersistingL2CacheSize, e->device);
  const size_t heads = (size_t)(c.ring + 2) * c.n_blocks * sizeof(unsigned long long);
  // At most ~40 MB of the 126 MB L2 are set aside: measured at 2^21 entities (heads 34 MB,
  // slab 68 MB) the head words alone gain 2.3 % and the whole slab LOSES 2 % -- the
  // records streaming through need the rest.  The whole slab when it fits that budget,
  // else the head words when they do, else nothing.
  size_t budget = (size_t)40 << 20;
  if (const char *env = getenv("SIMIAN_B200_L2_PERSIST_MB")) budget = (size_t)atoi(env) << 20; // (tuning)
  if (max_win <= 0 || l2 <= 0) return;
  if (budget > (size_t)l2) budget = (size_t)l2;
  if (budget > (size_t)max_win) budget = (size_t)max_win;
  size_t bytes = 0;
  if (e->l2_persist > 1 && e->hot_bytes <= budget) bytes = e->hot_bytes;
  else if (heads <= budget) bytes = heads;
  if (!bytes) return;
  const size_t set_aside = bytes;
  cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, set_aside);
  cudaStreamAttrValue av = {};
  av.accessPolicyWindow.base_ptr = e->hot_base; // (the cell head words come first)
  av.accessPolicyWindow.num_bytes = bytes;
  av.accessPolicyWindow.hitRatio = set_aside >= bytes ? 1.0f : (float)set_aside / (float)bytes;
  av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
  av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
  if (cudaStreamSetAttribute(e->stream, cudaStreamAttributeAccessPolicyWindow, &av) == cudaSuccess)
    e->l2_applied = true;
  cudaGetLastError();
}

// allocate the HBM layout, initialise it, replay schedService, set up window 0
int setup(simian_engine *e) {
  if (e->set_up) return 0;
  // an early layout sized for far fewer initial events than were scheduled since
  if (e->laid_out && e->initial_events > 2 * e->laid_out_initial + 1024) teardown_layout(e);
  int rc = setup_layout(e);
  if (rc) return rc;
  DevCfg &c = e->cfg;
  // replay schedService (simian.js:170-190)
  for (SchedOp &op : e->ops) {
    if (op.kind == 0) {
      unsigned long long total = (unsigned long long)op.count * op.repeat;
      int grid = (int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
      if (grid < 1) grid = 1;
      k_sched_bulk<<<grid, 256, 0, e->stream>>>(c, op.time, op.handler, op.data, op.cls, op.first,
                                                op.count, op.repeat, op.seq_base);
      CK(cudaGetLastError());
    } else if (op.kind == 2) {
      uint32_t nl = e->classes[op.cls].n_local;
      int grid = (int)((nl + 127) / 128 < 2368 ? (nl + 127) / 128 : 2368);
      if (grid < 1) grid = 1;
      k_construct<<<grid, 128, 0, e->stream>>>(c, op.cls, op.handler);
      CK(cudaGetLastError());
    } else if (op.n_recs) {
      if (!op.ingested && ingest_records(e, op.d_recs, op.n_recs, e->stream)) return e->err_code;
      CK(cudaStreamSynchronize(e->stream));
      if (!cache_give(e->device, op.d_bytes, op.d_recs)) cudaFree(op.d_recs);
      op.d_recs = nullptr;
    }
  }
  e->ops.clear();
  k_begin<<<1, NTHREADS, 0, e->stream>>>(c);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(e->stream));
  e->win_ix = 0;
  e->launches = 0;
  e->set_up = true;
  apply_l2_policy(e);
  return 0;
}

int poll(simian_engine *e) { // device flags -> pinned host copies
  CK(cudaMemcpyAsync(e->h_acc, e->cfg.acc, sizeof(DevAcc), cudaMemcpyDeviceToHost, e->stream));
  CK(cudaMemcpyAsync(e->h_win, e->cfg.win, 2 * sizeof(WinState), cudaMemcpyDeviceToHost,
                     e->stream));
  CK(cudaStreamSynchronize(e->stream));
  if (e->h_acc->error) return e->fail((int)e->h_acc->error, status_text((int)e->h_acc->error));
  return 0;
}

int collect(simian_engine *e, simian_stats *out) {
  CK(cudaMemsetAsync(e->d_digest, 0, 2 * sizeof(unsigned long long), e->stream));
  k_digest<<<592, 256, 0, e->stream>>>(e->cfg, e->d_digest);
  CK(cudaGetLastError());
  unsigned long long dg[2];
  CK(cudaMemcpyAsync(dg, e->d_digest, sizeof dg, cudaMemcpyDeviceToHost, e->stream));
  int rc = poll(e);
  if (out) {
    memset(out, 0, sizeof *out);
    const DevAcc &a = *e->h_acc;
    out->total_events = a.committed;
    out->windows = a.windows;
    out->synch_events = e->size > 1 ? 2 * a.windows : 0; // simian.js:145
    out->digest = dg[0];
    out->state_checksum = dg[1];
    out->ties = a.ties;
    out->distinguishable_ties = a.dist_ties;
    out->emitted = a.emitted;
    out->dropped_past_end = a.dropped;
    out->last_time = a.max_key ? simian_key_time(a.max_key) : e->start;
    out->run_seconds =
        std::chrono::duration<double>(std::chrono::steady_clock::now() - e->t_begin).count();
    float ms = 0;
    if (cudaEventElapsedTime(&ms, e->ev0, e->ev1) == cudaSuccess) out->device_seconds = ms * 1e-3;
    out->kernel_launches = e->launches;
    out->redistributions = a.redistributions;
    out->far_appends = a.far_appends;
    out->pages_allocated = a.page_allocs;
    out->error = e->err_code;
  }
  return rc ? rc : e->err_code;
}

} // namespace

extern "C" {

const char *simian_version(void) { return "simian-b200 0.1.0 (sm_100a)"; }

void simian_trim_cache(void) {
  std::lock_guard<std::mutex> g(g_alloc_cache_mu);
  int cur = 0;
  cudaGetDevice(&cur);
  for (auto &kv : g_alloc_cache) {
    cudaSetDevice(kv.first.device);
    cudaFree(kv.second);
  }
  cudaSetDevice(cur);
  g_alloc_cache.clear();
  g_alloc_cache_bytes = 0;
  for (auto &kv : g_pinned_cache) cudaFreeHost(kv.second);
  g_pinned_cache.clear();
  for (auto &kv : g_ipc_cache) cudaIpcCloseMemHandle(kv.second);
  g_ipc_cache.clear();
}

double simian_hash_name(const char *str) { // hash.js:21-26
  double res = 5381;
  size_t i = strlen(str);
  while (i--) res = 33 * res + (double)(unsigned char)str[i];
  return res;
}

simian_engine *simian_create(const char *simName, double startTime, double endTime,
                             double minDelay, int rank, int size, int device) {
  if (size < 1 || rank < 0 || rank >= size) {
    g_create_error = "bad rank/size";
    return nullptr;
  }
  if (size > SIMIAN_MAX_RANKS) { // peer tables of the engine (the GPUs of one NVLink domain)
    g_create_error = "size exceeds SIMIAN_MAX_RANKS (16): one engine per GPU of one box";
    return nullptr;
  }
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0) {
    g_create_error = std::string("no CUDA device (") +
                     (ce == cudaSuccess ? "device count 0" : cudaGetErrorString(ce)) +
                     "): simian-b200 has no CPU fallback";
    return nullptr;
  }
  simian_engine *e = new simian_engine();
  e->name = simName ? simName : "simian";
  e->start = startTime;
  e->end = (endTime != 0.0) ? endTime : 1e100;  // simian.js:41
  e->min_delay = (minDelay != 0.0) ? minDelay : 1; // simian.js:42
  e->inf_time = simian_inf_time(endTime, minDelay); // simian.js:64 (raw arguments)
  e->rank = rank;
  e->size = size;
  e->device = device >= 0 ? device : rank % ndev;
  if (cudaSetDevice(e->device) != cudaSuccess ||
      cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess) {
    g_create_error = "cudaSetDevice / cudaStreamCreate failed";
    delete e;
    return nullptr;
  }
  e->own_stream = true;
  return e;
}

void simian_destroy(simian_engine *e) {
  if (!e) return;
  cudaSetDevice(e->device);
  if (e->stream) cudaStreamSynchronize(e->stream);
  clear_l2_policy(e);
  // peer mappings stay open in the process-wide cache (simian_trim_cache closes them)
  for (auto &a : e->allocs)
    if (!cache_give(e->device, a.second, a.first)) cudaFree(a.first);
  for (SchedOp &op : e->ops)
    if (op.d_recs && !cache_give(e->device, op.d_bytes, op.d_recs)) cudaFree(op.d_recs);
  if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
  if (e->h_acc) pinned_give(sizeof(DevAcc), e->h_acc);
  if (e->h_win) pinned_give(2 * sizeof(WinState), e->h_win);
  if (e->h_counts) pinned_give(sizeof(uint32_t) * SIMIAN_MAX_RANKS, e->h_counts);
  if (e->ev0) cudaEventDestroy(e->ev0);
  if (e->ev1) cudaEventDestroy(e->ev1);
  if (e->own_stream && e->stream) cudaStreamDestroy(e->stream);
  delete e;
}

const char *simian_last_error(simian_engine *e) {
  return e ? e->err.c_str() : g_create_error.c_str();
}

int simian_set_option(simian_engine *e, const char *key, double v) {
  if (e->set_up) return e->fail(SIMIAN_ERR_RUNNING, "options must be set before the first run");
  teardown_layout(e);
  std::string k = key;
  if (k == "seed") e->seed = (uint64_t)v;
  else if (k == "bucket_width") e->bucket_width = v;
  else if (k == "ring_buckets") e->ring = (uint32_t)v;
  else if (k == "pool_events") e->pool_events = (uint64_t)v;
  else if (k == "outbox_events") e->outbox_events = (uint64_t)v;
  else if (k == "launch_batch") e->launch_batch = v < 1 ? 1 : (int)v;
  else if (k == "tie_check") e->tie_check = v != 0;
  else if (k == "force_sequential") e->force_sequential = v != 0;
  else if (k == "block_entities") e->epb_override = (uint32_t)v;
  else if (k == "pdl") e->pdl = v != 0;
  else if (k == "persistent_loop") e->loop = v != 0;
  else if (k == "deferred_pull") e->deferred_pull = v != 0;
  else if (k == "split_handlers") e->split = v != 0;
  else if (k == "payloads") e->payloads = v != 0;
  else if (k == "fused_advance") e->fused_advance = v != 0;
  else if (k == "l2_persist") e->l2_persist = (int)v; // 1: cell head words, 2: the whole hot slab
  else if (k == "pdl_exchange") e->pdl_exchange = v != 0;
  else if (k == "split_sums") e->split_sums = v != 0;
  else if (k == "sum_jobs") e->job_cap = v < 1 ? 1 : (uint64_t)v;
  else return e->fail(SIMIAN_ERR_BAD_ARGUMENT, "unknown option " + k);
  return 0;
}

int simian_set_stream(simian_engine *e, void *s) {
  // work already queued on the old stream (early layout + ingest of
  // simian_sched_events) must be complete before anything runs on the new one
  if (e->stream) {
    cudaSetDevice(e->device);
    cudaStreamSynchronize(e->stream);
    clear_l2_policy(e);
  }
  if (e->own_stream && e->stream) cudaStreamDestroy(e->stream);
  e->stream = (cudaStream_t)s;
  e->own_stream = false;
  if (e->set_up) apply_l2_policy(e);
  return 0;
}

int simian_define_class(simian_engine *e, const char *className, uint32_t stateBytes) {
  auto it = e->class_ix.find(className);
  if (it != e->class_ix.end()) return it->second;
  if (e->set_up) return -e->fail(SIMIAN_ERR_RUNNING, "classes must be defined before the first run");
  teardown_layout(e);
  if ((int)e->classes.size() >= SIMIAN_MAX_CLASSES || (stateBytes & 7u))
    return -e->fail(SIMIAN_ERR_BAD_ARGUMENT, "too many classes or stateBytes not a multiple of 8");
  HostClass h;
  h.name = className;
  h.state_bytes = stateBytes;
  h.name_hash = simian_hash_name(className);
  h.first_id = e->n_ids;
  memset(h.fn_of_service, 0, sizeof h.fn_of_service);
  e->classes.push_back(h);
  e->class_ix[className] = (int)e->classes.size() - 1;
  return (int)e->classes.size() - 1;
}

int simian_set_class_params(simian_engine *e, const char *className, const void *p,
                            uint32_t bytes) {
  auto it = e->class_ix.find(className);
  if (it == e->class_ix.end()) return e->fail(SIMIAN_ERR_UNKNOWN_NAME, "unknown class");
  if (e->set_up) return e->fail(SIMIAN_ERR_RUNNING, "parameters must be set before the first run");
  teardown_layout(e);
  e->classes[it->second].params.assign((const uint8_t *)p, (const uint8_t *)p + bytes);
  return 0;
}

int simian_add_entities(simian_engine *e, const char *className, uint32_t firstNum,
                        uint32_t count, const void *initState) {
  if (e->running || e->set_up) // simian.js:215
    return e->fail(SIMIAN_ERR_RUNNING, status_text(SIMIAN_ERR_RUNNING));
  teardown_layout(e);
  int ci = simian_define_class(e, className, 0);
  if (ci < 0) return -ci;
  HostClass &h = e->classes[ci];
  if (h.count == 0) h.first_id = e->n_ids; // ids are dealt when the first instance arrives
  if (firstNum != h.count || h.first_id + h.count != e->n_ids)
    return e->fail(SIMIAN_ERR_BAD_ARGUMENT,
                   "instances of a class must be added contiguously from 0, class after class");
  if (initState && h.state_bytes) {
    h.init_state.resize((size_t)(h.count + count) * h.state_bytes, 0);
    memcpy(&h.init_state[(size_t)h.count * h.state_bytes], initState,
           (size_t)count * h.state_bytes);
    h.has_init = true;
  } else if (h.has_init) {
    h.init_state.resize((size_t)(h.count + count) * h.state_bytes, 0);
  }
  h.count += count;
  e->n_ids += count;
  return 0;
}

int simian_intern_service(simian_engine *e, const char *serviceName) {
  auto it = e->service_ix.find(serviceName);
  if (it != e->service_ix.end()) return it->second;
  int id = (int)e->service_ix.size();
  if (id >= SIMIAN_MAX_SERVICES) return -e->fail(SIMIAN_ERR_BAD_ARGUMENT, "too many services");
  e->service_ix[serviceName] = id;
  return id;
}

int simian_attach_service(simian_engine *e, const char *className, const char *serviceName,
                          const char *deviceFnName) {
  if (e->set_up) return e->fail(SIMIAN_ERR_RUNNING, "services must be attached before the first run");
  teardown_layout(e);
  int ci = simian_define_class(e, className, 0);
  if (ci < 0) return -ci;
  int fn = builtin_fn_by_name(deviceFnName);
  if (!fn)
    return e->fail(SIMIAN_ERR_UNKNOWN_NAME, std::string("no device function named ") + deviceFnName);
  int sid = simian_intern_service(e, serviceName);
  if (sid < 0) return -sid;
  e->classes[ci].fn_of_service[sid] = (uint8_t)fn;
  return 0;
}

int simian_device_function_id(const char *deviceFnName) {
  return deviceFnName ? builtin_fn_by_name(deviceFnName) : 0;
}

int64_t simian_first_id(simian_engine *e, const char *className) {
  auto it = e->class_ix.find(className);
  return it == e->class_ix.end() ? -1 : (int64_t)e->classes[it->second].first_id;
}

int simian_base_rank(simian_engine *e, const char *className) {
  auto it = e->class_ix.find(className);
  if (it == e->class_ix.end()) return -1;
  const HostClass &h = e->classes[it->second];
  if (h.base_rank_override >= 0) return h.base_rank_override % e->size;
  return (int)fmod(h.name_hash, (double)e->size);
}

int simian_set_base_rank(simian_engine *e, const char *className, int baseRank) {
  auto it = e->class_ix.find(className);
  if (it == e->class_ix.end()) return e->fail(SIMIAN_ERR_UNKNOWN_NAME, "unknown entity class");
  if (baseRank < 0) return e->fail(SIMIAN_ERR_BAD_ARGUMENT, "base rank must be >= 0");
  if (e->set_up || e->classes[it->second].count)
    return e->fail(SIMIAN_ERR_RUNNING, "the base rank of a class is set before its entities are added");
  teardown_layout(e);
  e->classes[it->second].base_rank_override = baseRank;
  return 0;
}

int simian_sched_service_bulk(simian_engine *e, double time, const char *serviceName,
                              uint64_t data, const char *rx, uint32_t firstId, uint32_t count,
                              uint32_t repeat) {
  if (e->set_up) return e->fail(SIMIAN_ERR_RUNNING, "schedService after the run started");
  auto it = e->class_ix.find(rx);
  if (it == e->class_ix.end()) return e->fail(SIMIAN_ERR_UNKNOWN_NAME, "unknown entity class");
  HostClass &h = e->classes[it->second];
  if ((uint64_t)firstId + count > h.count) return e->fail(SIMIAN_ERR_UNKNOWN_NAME, "no such entity");
  int sid = simian_intern_service(e, serviceName);
  if (sid < 0) return -sid;
  uint32_t seq_base = e->sched_seq;
  e->sched_seq += count * repeat;
  if (time > e->end) return 0; // simian.js:173
  if (time < e->start) return e->fail(SIMIAN_ERR_OUT_OF_ORDER, "schedService before startTime");
  SchedOp op;
  op.kind = 0;
  op.time = time;
  op.handler = (uint32_t)sid;
  op.data = data;
  op.cls = it->second;
  op.first = firstId;
  op.count = count;
  op.repeat = repeat;
  op.seq_base = seq_base;
  e->ops.push_back(std::move(op));
  e->initial_events += (uint64_t)count * repeat;
  return 0;
}

int simian_sched_service(simian_engine *e, double time, const char *serviceName, uint64_t data,
                         const char *rx, uint32_t rxId) {
  return simian_sched_service_bulk(e, time, serviceName, data, rx, rxId, 1, 1);
}

int simian_sched_events(simian_engine *e, const simian_event_t *ev, uint64_t n) {
  if (e->set_up) return e->fail(SIMIAN_ERR_RUNNING, "schedService after the run started");
  if (!n) return 0;
  // Host -> HBM from the caller's buffer (fast when it is pinned), in chunks on
  // a copy stream.  When the model is already described (entities added) the
  // HBM layout is allocated now and every chunk goes into the calendar while
  // the next one is still crossing PCIe; otherwise the records wait on the
  // device for simian_run.  Ids and times are validated by the ingest kernel.
  CK(cudaSetDevice(e->device));
  SchedOp op;
  op.kind = 1;
  op.n_recs = n;
  op.d_bytes = (n * sizeof(simian_event_t) + 511) & ~(size_t)511;
  op.d_recs = (simian_event_t *)cache_take(e->device, op.d_bytes);
  if (!op.d_recs) CK(cudaMalloc((void **)&op.d_recs, op.d_bytes));
  e->initial_events += n;
  e->record_events += n;
  // (worth it for bulk inputs only: small calls just park their records)
  bool early = !e->classes.empty() && e->n_ids > 0 && e->err_code == 0 &&
               (n >= 65536 || e->laid_out);
  if (early && setup_layout(e)) return e->err_code;
  if (!e->copy_stream) CK(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
  const uint64_t chunk = 1ull << 20; // records (32 MB)
  std::vector<cudaEvent_t> evs;
  cudaError_t ce = cudaSuccess;
  const char *what = "";
#define TRY(call) if (ce == cudaSuccess) { ce = (call); what = #call; }
  for (uint64_t off = 0; off < n && ce == cudaSuccess && !e->err_code; off += chunk) {
    uint64_t m = n - off < chunk ? n - off : chunk;
    TRY(cudaMemcpyAsync(op.d_recs + off, ev + off, m * sizeof(simian_event_t),
                        cudaMemcpyHostToDevice, e->copy_stream));
    if (early) {
      cudaEvent_t done = nullptr;
      TRY(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
      if (done) evs.push_back(done);
      TRY(cudaEventRecord(done, e->copy_stream));
      TRY(cudaStreamWaitEvent(e->stream, done, 0));
      if (ce == cudaSuccess) ingest_records(e, op.d_recs + off, m, e->stream);
    }
  }
  TRY(cudaStreamSynchronize(e->copy_stream)); // the host buffer is only borrowed
#undef TRY
  for (cudaEvent_t d : evs) cudaEventDestroy(d);
  if (ce != cudaSuccess || e->err_code) { // nothing leaks on a failed call
    cudaStreamSynchronize(e->stream);
    if (!cache_give(e->device, op.d_bytes, op.d_recs)) cudaFree(op.d_recs);
    if (ce != cudaSuccess) return e->fail(SIMIAN_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(ce));
    return e->err_code;
  }
  op.ingested = early;
  e->ops.push_back(op);
  return 0;
}

int simian_construct_entities(simian_engine *e, const char *className, const char *deviceFnName,
                              uint64_t expectedEvents) {
  if (e->set_up) return e->fail(SIMIAN_ERR_RUNNING, status_text(SIMIAN_ERR_RUNNING));
  auto it = e->class_ix.find(className);
  if (it == e->class_ix.end()) return e->fail(SIMIAN_ERR_UNKNOWN_NAME, "unknown entity class");
  int fn = builtin_fn_by_name(deviceFnName);
  if (!fn)
    return e->fail(SIMIAN_ERR_UNKNOWN_NAME, std::string("no device function named ") + deviceFnName);
  SchedOp op;
  op.kind = 2;
  op.cls = it->second;
  op.handler = (uint32_t)fn;
  e->ops.push_back(op);
  e->initial_events += expectedEvents;
  return 0;
}

int simian_setup(simian_engine *e) {
  if (e->err_code) return e->err_code;
  CK(cudaSetDevice(e->device));
  return setup(e);
}

namespace {
// this engine's share of the run nonce (the nonce is the sum over all ranks)
void ensure_salt(simian_engine *e) {
  if (e->salt) return;
  std::random_device rd;
  e->salt = (((uint64_t)rd() << 32) ^ (uint64_t)rd() ^
             (uint64_t)std::chrono::steady_clock::now().time_since_epoch().count())
            << 24; // the low 24 bits stay free for the window number
  if (!e->salt) e->salt = 1ull << 24;
  e->cfg.run_nonce += e->salt;
}

struct IpcBlob { // what one rank publishes to its peers
  cudaIpcMemHandle_t outbox, slab;
  uint64_t cap;
  uint64_t salt; // random per engine; the run nonce is the sum over all ranks
  int32_t rank, size;
};
} // namespace

int simian_ipc_export(simian_engine *e, void *blob, uint64_t blobBytes, uint64_t *written) {
  if (written) *written = sizeof(IpcBlob);
  if (!blob) return 0; // size query
  if (blobBytes < sizeof(IpcBlob)) return e->fail(SIMIAN_ERR_BAD_ARGUMENT, "ipc blob too small");
  if (e->size < 2) return e->fail(SIMIAN_ERR_BAD_ARGUMENT, "ipc export needs size > 1");
  int rc = simian_setup(e);
  if (rc) return rc;
  IpcBlob b;
  memset(&b, 0, sizeof b);
  CK(cudaIpcGetMemHandle(&b.outbox, e->cfg.outbox));
  CK(cudaIpcGetMemHandle(&b.slab, e->ipc_slab));
  // k_init has zeroed the exchange slots before any peer can write into them
  CK(cudaStreamSynchronize(e->stream));
  b.cap = e->cfg.outbox_cap;
  ensure_salt(e);
  b.salt = e->salt;
  b.rank = e->rank;
  b.size = e->size;
  memcpy(blob, &b, sizeof b);
  return 0;
}

int simian_ipc_import(simian_engine *e, int peerRank, const void *blob, uint64_t blobBytes) {
  if (blobBytes < sizeof(IpcBlob) || peerRank < 0 || peerRank >= e->size ||
      peerRank >= SIMIAN_MAX_RANKS || peerRank == e->rank)
    return e->fail(SIMIAN_ERR_BAD_ARGUMENT, "bad ipc import");
  int rc = simian_setup(e);
  if (rc) return rc;
  IpcBlob b;
  memcpy(&b, blob, sizeof b);
  if (b.rank != peerRank || b.size != e->size || b.cap != e->cfg.outbox_cap)
    return e->fail(SIMIAN_ERR_BAD_ARGUMENT, "ipc blob does not match this world");
  void *po = nullptr, *ps = nullptr;
  CK(ipc_open_cached(&po, b.outbox));
  CK(ipc_open_cached(&ps, b.slab));
  void *pc = ps;
  e->cfg.peer_xch[peerRank] = (MinSlot *)((uint8_t *)ps + 4096);
  if (!e->peer_seen[peerRank]) { // nonce = sum of every rank's salt (own one added at export)
    e->cfg.run_nonce += b.salt;
    e->peer_seen[peerRank] = true;
  }
  e->cfg.peer_outbox[peerRank] = (const simian_event_t *)po;
  e->cfg.peer_count[peerRank] = (const uint32_t *)pc;
  return 0;
}

// Peer connection inside ONE process (all ranks of a world on one device, or on
// devices with peer access enabled by the caller): the same wiring as
// simian_ipc_export / simian_ipc_import, with plain device pointers.
int simian_connect_local(simian_engine *e, int peerRank, simian_engine *peer) {
  if (!peer || peerRank < 0 || peerRank >= e->size || peerRank >= SIMIAN_MAX_RANKS ||
      peerRank == e->rank || peer->rank != peerRank || peer->size != e->size)
    return e->fail(SIMIAN_ERR_BAD_ARGUMENT, "bad local peer");
  int rc = simian_setup(e);
  if (rc) return rc;
  if (simian_setup(peer)) return e->fail(SIMIAN_ERR_BAD_ARGUMENT, "peer engine failed to set up");
  if (peer->cfg.outbox_cap != e->cfg.outbox_cap)
    return e->fail(SIMIAN_ERR_BAD_ARGUMENT, "peer outbox does not match this world");
  CK(cudaStreamSynchronize(e->stream)); // k_init has zeroed the exchange slots
  ensure_salt(e);
  ensure_salt(peer);
  if (!e->peer_seen[peerRank]) {
    e->cfg.run_nonce += peer->salt;
    e->peer_seen[peerRank] = true;
  }
  e->cfg.peer_xch[peerRank] = peer->cfg.xch;
  e->cfg.peer_outbox[peerRank] = peer->cfg.outbox;
  e->cfg.peer_count[peerRank] = peer->cfg.outbox_count;
  return 0;
}

int simian_window_publish_min(simian_engine *e) {
  k_local_min<<<1, NTHREADS, e->smem_bytes, e->stream>>>(e->cfg, e->win_ix, 1);
  CK(cudaGetLastError());
  e->launches++;
  return 0;
}

int simian_window_pull(simian_engine *e, int waitForPeers) {
  if (e->size < 2) return 0;
  uint32_t bpp = 148; // blocks per peer
  if (const char *env = getenv("SIMIAN_B200_PULL_BPP")) bpp = (uint32_t)atoi(env); // (tuning)
  if (bpp < 1) bpp = 1;
  // with two-ended outboxes: the part that must be in before the next window
  {
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3((e->size - 1) * bpp);
    lc.blockDim = dim3(256);
    lc.stream = e->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = at;
    lc.numAttrs = e->pdl_exchange ? 1 : 0;
    CK(cudaLaunchKernelEx(&lc, k_pull, e->cfg, e->win_ix, bpp, waitForPeers,
                          e->cfg.split_outbox ? 1 : 0));
  }
  CK(cudaGetLastError());
  e->launches++;
  return 0;
}

int simian_minvec_ptr(simian_engine *e, void **devMinVec) {
  if (!e->set_up) return e->fail(SIMIAN_ERR_BAD_ARGUMENT, "call simian_setup first");
  *devMinVec = (void *)e->cfg.acc->minvec;
  return 0;
}

int simian_window_advance_async(simian_engine *e, int minFromPeers) {
  {
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(1);
    lc.blockDim = dim3(NTHREADS);
    lc.stream = e->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = at;
    lc.numAttrs = e->pdl_exchange ? 1 : 0;
    CK(cudaLaunchKernelEx(&lc, k_advance, e->cfg, e->win_ix, minFromPeers));
  }
  CK(cudaGetLastError());
  e->launches++;
  e->win_ix++;
  return 0;
}

int simian_poll(simian_engine *e, int *done) {
  int rc = poll(e);
  if (rc) return rc;
  if (done) *done = (int)e->h_win[e->win_ix & 1u].done;
  return 0;
}

int simian_run_begin(simian_engine *e) {
  if (e->err_code) return e->err_code;
  CK(cudaSetDevice(e->device));
  e->t_begin = std::chrono::steady_clock::now();
  int rc = setup(e);
  if (rc) return rc;
  e->running = true;
  CK(cudaEventRecord(e->ev0, e->stream));
  return 0;
}

// One window's launch(es).  fuse: what the last CTA does (k_window); pull_blocks: CTAs
// of the deferred pull riding in front of the entity blocks.  With programmatic
// dependent launch the CTAs of a launch become resident while the last wave of the
// one before drains (they wait inside the kernel).
static int launch_window(simian_engine *e, int fuse, uint32_t pull_blocks, bool pdl) {
  cudaLaunchConfig_t lc = {};
  lc.blockDim = dim3(NTHREADS);
  lc.stream = e->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  lc.attrs = at;
  lc.numAttrs = pdl ? 1 : 0;
  if (e->split_on && pull_blocks == 0) {
    lc.gridDim = dim3(e->cfg.n_blocks);
    lc.dynamicSmemBytes = e->smem_bytes;
    CK(cudaLaunchKernelEx(&lc, k_window_stage, e->cfg, e->win_ix));
    // the blocks k_window_stage left alone (bursts, redistribution passes): a grid of
    // resident CTAs looping over the blocks; nothing to do in a steady window
    lc.gridDim = dim3(e->cfg.n_blocks < e->rest_ctas ? e->cfg.n_blocks : e->rest_ctas);
    CK(cudaLaunchKernelEx(&lc, k_window_rest, e->cfg, e->win_ix));
    // k_emit: a persistent grid of warps striding over the blocks' chunks
    const uint32_t want = (e->cfg.n_blocks * EMIT_CHUNKS_PER_BLOCK + EMIT_WARPS - 1) / EMIT_WARPS;
    lc.gridDim = dim3(want < e->emit_ctas ? want : e->emit_ctas);
    lc.dynamicSmemBytes = 0;
    CK(cudaLaunchKernelEx(&lc, k_emit, e->cfg, e->win_ix, fuse));
    e->launches += 3;
    return 0;
  }
  lc.gridDim = dim3(e->cfg.n_blocks + pull_blocks);
  lc.dynamicSmemBytes = e->smem_bytes;
  // size == 1: the LBTS step by the last CTA of k_window (fence + ticket + barrier in every
  // CTA) or as a one-CTA launch of its own (k_lbts).  Measured: config 2 (2048 CTAs per
  // window) 15.07 -> 14.77 ms, config 5 (4096) 40.1 -> 38.9 ms with the launch of its own;
  // config 1 (2 CTAs) 2.13 -> 2.30 s.  Default: its own launch from 592 entity blocks on.
  const bool own_lbts = fuse == 1 && (e->fused_advance < 0 ? e->cfg.n_blocks >= 592u
                                                           : e->fused_advance == 0);
  const int fuse_k = own_lbts ? 0 : fuse;
  if (e->cfg.blobs) // (its own instantiation: the ordinary kernel never looks at record flags)
    CK(cudaLaunchKernelEx(&lc, k_window_blobs, e->cfg, e->win_ix, fuse_k, pull_blocks));
  else
    CK(cudaLaunchKernelEx(&lc, k_window, e->cfg, e->win_ix, fuse_k, pull_blocks));
  e->launches++;
  if (own_lbts) { // the LBTS step as a launch of its own
    lc.gridDim = dim3(1);
    lc.dynamicSmemBytes = 0;
    CK(cudaLaunchKernelEx(&lc, k_lbts, e->cfg, e->win_ix));
    e->launches++;
  }
  if (e->cfg.jobs) { // the sums the window's handlers deferred
    lc.gridDim = dim3(e->sums_ctas);
    lc.dynamicSmemBytes = 0;
    CK(cudaLaunchKernelEx(&lc, k_sums, e->cfg, e->win_ix));
    e->launches++;
  }
  return 0;
}

int simian_window_process(simian_engine *e) {
  return launch_window(e, e->size == 1, 0u, false);
}

int simian_window_process_lbts(simian_engine *e, int publish) {
  // deferred pull: the far records of the previous window are pulled by the first
  // CTAs of this launch, beside the entity blocks (k_window, pull_blocks)
  uint32_t pull_blocks = 0;
  if (e->cfg.split_outbox && e->size > 1) {
    const char *env = getenv("SIMIAN_B200_PULL_CTAS"); // total pull CTAs per launch (tuning)
    uint32_t total = env ? (uint32_t)atoi(env) : 48u;
    uint32_t bpp = total / (uint32_t)(e->size - 1);
    if (bpp < 4u) bpp = 4u;
    pull_blocks = bpp * (uint32_t)(e->size - 1);
  }
  // as for size == 1 (launch_window): from 592 entity blocks on, the LBTS candidate is
  // computed (and published) by a one-CTA launch of its own instead of the last CTA of
  // the window kernel -- the CTAs then end without fence + ticket + barrier
  const bool own_lbts = pull_blocks == 0 && (e->fused_advance < 0 ? e->cfg.n_blocks >= 592u
                                                                  : e->fused_advance == 0);
  if (own_lbts) {
    int rc = launch_window(e, 0, 0u, e->pdl_exchange != 0);
    if (rc) return rc;
    k_local_min<<<1, NTHREADS, e->smem_bytes, e->stream>>>(e->cfg, e->win_ix, publish ? 1 : 0);
    CK(cudaGetLastError());
    e->launches++;
    return 0;
  }
  return launch_window(e, publish ? 3 : 2, pull_blocks, e->pdl_exchange != 0);
}

int simian_window_outbox(simian_engine *e, uint32_t *hostCounts, void **devOutbox,
                         uint64_t *cap) {
  if (e->size == 1) {
    hostCounts[0] = 0;
  } else {
    if (e->cfg.split_outbox)
      return e->fail(SIMIAN_ERR_BAD_ARGUMENT,
                     "deferred_pull engines are driven by the device-driven exchange only");
    CK(cudaMemcpyAsync(e->h_counts, e->cfg.outbox_count + (size_t)(e->win_ix & 1u) * e->size,
                       sizeof(uint32_t) * (size_t)e->size, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    memcpy(hostCounts, e->h_counts, sizeof(uint32_t) * (size_t)e->size);
  }
  if (devOutbox) // the outbox this window filled (double buffered by window parity)
    *devOutbox = e->cfg.outbox + (size_t)(e->win_ix & 1u) * e->size * e->cfg.outbox_cap;
  if (cap) *cap = e->cfg.outbox_cap;
  return 0;
}

int simian_window_ingest(simian_engine *e, const void *devRecords, uint64_t n) {
  if (!n) return 0;
  int grid = (int)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184);
  k_ingest<<<grid, 256, 0, e->stream>>>(e->cfg, (const simian_event_t *)devRecords, n, 0, 1,
                                        e->win_ix);
  CK(cudaGetLastError());
  e->launches++;
  return 0;
}

int simian_window_local_min(simian_engine *e, void **devMinVec) {
  k_local_min<<<1, NTHREADS, e->smem_bytes, e->stream>>>(e->cfg, e->win_ix, 0);
  CK(cudaGetLastError());
  e->launches++;
  if (devMinVec) *devMinVec = (void *)e->cfg.acc->minvec;
  return 0;
}

int simian_window_advance(simian_engine *e, int *done) {
  k_advance<<<1, NTHREADS, 0, e->stream>>>(e->cfg, e->win_ix, 0);
  CK(cudaGetLastError());
  e->launches++;
  e->win_ix++;
  int rc = poll(e);
  if (rc) return rc;
  if (done) *done = (int)e->h_win[e->win_ix & 1u].done;
  return 0;
}

int simian_run_end(simian_engine *e, simian_stats *out) {
  cudaEventRecord(e->ev1, e->stream);
  e->running = false;
  return collect(e, out);
}

int simian_run(simian_engine *e, simian_stats *out) {
  if (e->size != 1)
    return e->fail(SIMIAN_ERR_BAD_ARGUMENT,
                   "size > 1: drive the windows with the stepping calls (the host runs the collectives)");
  int rc = simian_run_begin(e);
  if (rc) return rc;
  int batch = e->launch_batch;
  if (e->loop_ok) {
    // the whole window loop inside cooperative launches of k_window_loop (window_loop.cu): a grid
    // barrier per window instead of a kernel launch (models whose blocks are all
    // resident at once; tiny windows are launch-latency bound otherwise)
    for (;;) {
      CK(simian_loop_launch(e->cfg, e->win_ix, 1u << 16, simian_loop_smem_bytes(e->cfg.epb), e->stream));
      e->launches++;
      rc = poll(e);
      if (rc) break;
      if (e->h_acc->window_index == e->win_ix) { // the launch made no progress
        rc = e->fail(SIMIAN_ERR_CUDA, "window kernel did not advance");
        break;
      }
      e->win_ix = e->h_acc->window_index;
      if (e->h_win[e->win_ix & 1u].done) break;
    }
    int rc2 = simian_run_end(e, out);
    return rc ? rc : rc2;
  }
  for (;;) {
    for (int i = 0; i < batch; i++) {
      if ((rc = launch_window(e, 1, 0u, e->pdl != 0))) break;
      e->win_ix++;
    }
    if (rc) break;
    CK(cudaGetLastError());
    rc = poll(e);
    if (rc) break;
    if (e->h_win[e->win_ix & 1u].done) break;
    if (e->h_acc->window_index != e->win_ix) { // a launch made no progress
      rc = e->fail(SIMIAN_ERR_CUDA, "window kernel did not advance");
      break;
    }
    if (batch < 256) batch *= 2;
  }
  int rc2 = simian_run_end(e, out);
  return rc ? rc : rc2;
}

// the caller's buffer is page-locked host memory: results are copied straight into it
static bool is_pinned_host(const void *p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}

static int read_core(simian_engine *e, const char *className, uint32_t first, uint32_t count,
                     uint64_t *out, int which) {
  auto it = e->class_ix.find(className);
  if (it == e->class_ix.end()) return e->fail(SIMIAN_ERR_UNKNOWN_NAME, "unknown class");
  if (!e->set_up) return e->fail(SIMIAN_ERR_BAD_ARGUMENT, "nothing has run yet");
  HostClass &h = e->classes[it->second];
  if ((uint64_t)first + count > h.count) return e->fail(SIMIAN_ERR_UNKNOWN_NAME, "no such entity");
  if (!h.n_local || !count) return 0;
  CK(cudaSetDevice(e->device));
  // one field of the 16-byte entity records, contiguous, then ONE copy
  unsigned long long *tmp = nullptr;
  if (dev_alloc(e, &tmp, (size_t)h.n_local)) return e->err_code;
  k_gather_core<<<592, 256, 0, e->stream>>>(e->cfg, h.slot_base, h.n_local, which, tmp);
  CK(cudaGetLastError());
  uint32_t g = (uint32_t)e->size;
  if (g == 1 && is_pinned_host(out)) { // local index == num: one DMA into the caller's buffer
    CK(cudaMemcpyAsync(out, tmp + first, sizeof(uint64_t) * (size_t)count, cudaMemcpyDeviceToHost,
                       e->stream));
    CK(cudaStreamSynchronize(e->stream));
  } else {
    size_t pin_bytes = sizeof(uint64_t) * (size_t)h.n_local;
    uint64_t *loc = (uint64_t *)pinned_take(pin_bytes);
    if (!loc) return e->fail(SIMIAN_ERR_CUDA, "cudaMallocHost failed");
    CK(cudaMemcpyAsync(loc, tmp, pin_bytes, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    if (g == 1) {
      memcpy(out, loc + first, sizeof(uint64_t) * (size_t)count);
    } else {
      for (uint32_t j = 0; j < h.n_local; j++) {
        uint32_t num = j * g + h.rank_off;
        if (num >= first && num < first + count) out[num - first] = loc[j];
      }
    }
    pinned_give(pin_bytes, loc);
  }
  // hand the scratch buffer straight back to the allocation cache
  auto a = e->allocs.back();
  e->allocs.pop_back();
  if (!cache_give(e->device, a.second, a.first)) cudaFree(a.first);
  return 0;
}

// this rank's own instances of a class, in local order (num = rank_off + j*size)
int simian_read_local(simian_engine *e, const char *className, int which, uint64_t *out,
                      uint64_t capacity, uint64_t *nLocal, uint32_t *firstNum, uint32_t *stride) {
  auto it = e->class_ix.find(className);
  if (it == e->class_ix.end()) return e->fail(SIMIAN_ERR_UNKNOWN_NAME, "unknown class");
  if (!e->set_up) return e->fail(SIMIAN_ERR_BAD_ARGUMENT, "nothing has run yet");
  HostClass &h = e->classes[it->second];
  if (nLocal) *nLocal = h.n_local;
  if (firstNum) *firstNum = h.rank_off;
  if (stride) *stride = (uint32_t)e->size;
  if (!out) return 0; // size query
  if (capacity < h.n_local) return e->fail(SIMIAN_ERR_BAD_ARGUMENT, "buffer too small");
  if (!h.n_local) return 0;
  CK(cudaSetDevice(e->device));
  unsigned long long *tmp = nullptr;
  if (dev_alloc(e, &tmp, (size_t)h.n_local)) return e->err_code;
  k_gather_core<<<592, 256, 0, e->stream>>>(e->cfg, h.slot_base, h.n_local, which, tmp);
  CK(cudaGetLastError());
  size_t bytes = sizeof(uint64_t) * (size_t)h.n_local;
  if (is_pinned_host(out)) { // one DMA into the caller's page-locked buffer
    CK(cudaMemcpyAsync(out, tmp, bytes, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
  } else {
    uint64_t *loc = (uint64_t *)pinned_take(bytes);
    if (!loc) return e->fail(SIMIAN_ERR_CUDA, "cudaMallocHost failed");
    CK(cudaMemcpyAsync(loc, tmp, bytes, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    memcpy(out, loc, bytes);
    pinned_give(bytes, loc);
  }
  auto a = e->allocs.back();
  e->allocs.pop_back();
  if (!cache_give(e->device, a.second, a.first)) cudaFree(a.first);
  return 0;
}

// diagnostic: summed SM cycles per k_window phase (SIMIAN_PROFILE_PHASES builds)
int simian_debug_phase_cycles(simian_engine *e, uint64_t *out12) {
  if (!e->h_acc) return SIMIAN_ERR_BAD_ARGUMENT;
  for (int i = 0; i < 12; i++) out12[i] = e->h_acc->phase_cycles[i];
  return 0;
}

int simian_read_counts(simian_engine *e, const char *className, uint32_t first, uint32_t count,
                       uint64_t *out) {
  return read_core(e, className, first, count, out, 0);
}

int simian_read_trace_hashes(simian_engine *e, const char *className, uint32_t first,
                             uint32_t count, uint64_t *out) {
  return read_core(e, className, first, count, out, 1);
}

int simian_read_state(simian_engine *e, const char *className, uint32_t first, uint32_t count,
                      void *out) {
  auto it = e->class_ix.find(className);
  if (it == e->class_ix.end()) return e->fail(SIMIAN_ERR_UNKNOWN_NAME, "unknown class");
  if (!e->set_up) return e->fail(SIMIAN_ERR_BAD_ARGUMENT, "nothing has run yet");
  HostClass &h = e->classes[it->second];
  if ((uint64_t)first + count > h.count) return e->fail(SIMIAN_ERR_UNKNOWN_NAME, "no such entity");
  if (!h.state_bytes) return 0;
  std::vector<uint8_t> loc((size_t)h.n_local * h.state_bytes);
  CK(cudaSetDevice(e->device));
  CK(cudaMemcpy(loc.data(), h.d_state, loc.size(), cudaMemcpyDeviceToHost));
  uint32_t g = (uint32_t)e->size;
  for (uint32_t j = 0; j < h.n_local; j++) {
    uint32_t num = j * g + h.rank_off;
    if (num >= first && num < first + count)
      memcpy((uint8_t *)out + (size_t)(num - first) * h.state_bytes,
             &loc[(size_t)j * h.state_bytes], h.state_bytes);
  }
  return 0;
}

} // extern "C"
