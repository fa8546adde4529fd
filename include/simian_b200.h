/*
 * simian_b200.h -- C-ABI of libsimian_b200.so, the B200-native drop-in for the
 * event-oriented service path of SimianJS (eventQ.js + simian.js + entity.js).
 *
 * Plain C: pointers, sizes, doubles.  No torch / CUDA types.  Every entry point
 * names the reference interface it replaces (paths relative to /root/reference).
 * The reference's own native boundary is the `masala` object of
 * SimianJS/MasalaChai/masala.cpp:39-146 (natives return false + pending
 * exception on failure); here every call returns an int status (0 = ok) and
 * simian_last_error() yields the message the JS facade rethrows.
 *
 * Ownership: the engine owns all device memory.  Host pointers are borrowed
 * for the duration of the call; results are copied into caller buffers.
 * Threading: one host thread per engine; calls are not re-entrant; one engine
 * = one MPI rank of the reference = one GPU (simian.js:71-79).
 */
#ifndef SIMIAN_B200_H
#define SIMIAN_B200_H

#include <stdint.h>

#include "simian_spec.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct simian_engine simian_engine;

/* status codes; 1..4 mirror the reference's thrown errors */
enum simian_status {
  SIMIAN_OK = 0,
  SIMIAN_ERR_LOOKAHEAD = 1,    /* entity.js:44 "attempted to send with too little delay" */
  SIMIAN_ERR_IDLE_SEND = 2,    /* entity.js:42 "Cannot send an event when Simian is idle" */
  SIMIAN_ERR_OUT_OF_ORDER = 3, /* simian.js:125 "Out of order event"                      */
  SIMIAN_ERR_RUNNING = 4,      /* simian.js:215 "Cannot add new entity when Simian is running!" */
  SIMIAN_ERR_UNKNOWN_NAME = 5, /* entity / service / device function not registered       */
  SIMIAN_ERR_BAD_ARGUMENT = 6,
  SIMIAN_ERR_POOL_EXHAUSTED = 7, /* event pool (pages) full: raise option "pool_events"   */
  SIMIAN_ERR_WINDOW_OVERFLOW = 8, /* one entity block has more due pages than a CTA lists */
  SIMIAN_ERR_ENTITY_BURST = 9,   /* one entity has more due events than fit on chip       */
  SIMIAN_ERR_SELF_PENDING = 10,  /* too many same-window self-sends pending on an entity  */
  SIMIAN_ERR_OUTBOX_FULL = 11,   /* cross-GPU staging buffer full: raise "outbox_events"  */
  SIMIAN_ERR_RING_SPAN = 12,     /* window spans the whole calendar ring: raise "bucket_width" */
  SIMIAN_ERR_CUDA = 13,
  SIMIAN_ERR_NO_DEVICE = 14, /* no CUDA device: there is NO CPU fallback                  */
  SIMIAN_ERR_STALLED = 15,   /* a device-side wait (page install, peer flag) timed out     */
  SIMIAN_ERR_PAYLOAD = 16    /* payload word missing / too long / not available on this path */
};

/* What Simian.run prints (simian.js:159-164) plus the parity quantities. */
typedef struct simian_stats {
  uint64_t total_events;   /* "SIMULATED EVENTS" of this rank (simian.js:133)          */
  uint64_t windows;        /* iterations of the window loop (simian.js:119)            */
  uint64_t synch_events;   /* simian.js:145                                            */
  uint64_t digest;         /* sum over local entities of simian_digest_term            */
  uint64_t state_checksum; /* sum over local entities of folded user state             */
  uint64_t ties;           /* same-entity equal-time commits seen                      */
  uint64_t distinguishable_ties;
  uint64_t emitted;          /* events created by reqService and not dropped           */
  uint64_t dropped_past_end; /* time > endTime (entity.js:48)                          */
  double last_time;          /* engine.now after run (max committed time)              */
  double run_seconds;        /* host wall clock around run ("SIMULATION FINISHED IN")  */
  double device_seconds;     /* CUDA-event time of the window loop on the engine stream */
  uint64_t kernel_launches;  /* kernels of this library launched inside run            */
  uint64_t redistributions;  /* far-future list redistribution passes                  */
  uint64_t far_appends;
  uint64_t pages_allocated;
  int32_t error;
  int32_t pad_;
} simian_stats;

const char *simian_version(void);

/* new Simian(simName, startTime, endTime, minDelay, useMPI)   simian.js:36-84
 * rank/size play MPI.rank()/MPI.size() (simian.js:73-75): one engine per GPU.
 * size <= 16 (SIMIAN_MAX_RANKS: the GPUs of one NVLink domain; the peer tables
 * of the device-driven exchange are that long) -- larger worlds are refused.
 * device < 0 selects cuda device `rank % deviceCount`.  Returns NULL on failure
 * (simian_last_error(NULL) tells why; in particular when no GPU is present). */
simian_engine *simian_create(const char *simName, double startTime,
                             double endTime, double minDelay, int rank,
                             int size, int device);

/* Simian.exit                                                  simian.js:90-97 */
void simian_destroy(simian_engine *e);

const char *simian_last_error(simian_engine *e);

/* Engine tunables, all optional, before the first run:
 *   "seed"            engine seed of the shared counter-based RNG (default 1)
 *   "bucket_width"    calendar sub-bucket width in simulated time (default minDelay/8)
 *   "ring_buckets"    calendar ring length, power of two (default 1024)
 *   "pool_events"     event pool capacity in records (default 2x initial + one page per calendar cell)
 *   "outbox_events"   per-peer cross-GPU staging capacity in records
 *   "launch_batch"    window kernels enqueued between host polls
 *   "tie_check"       1: count same-entity time ties (default 1)
 *   "force_sequential" 1: always run handlers one thread per entity (testing)
 *   "block_entities"  entities per entity block = per CTA of the window kernel (default and max 512)
 *   "deferred_pull"   1: two-ended cross-GPU outboxes, see simian_window_pull (default 0)
 *   "persistent_loop" 1: when every entity block's CTA is resident at once (small and
 *                     medium models, size == 1) the whole window loop runs inside
 *                     cooperative launches of one kernel, a grid barrier per window
 *                     instead of a kernel launch; 0 (default): one launch per window
 *                     (measured equal: DESIGN.md section 7)
 *   "pdl"             1: window kernels are launched as programmatic dependents of
 *                     each other, the next window's CTAs become resident while the
 *                     previous one drains (default 1; size == 1)
 *   "payloads"        1: handlers may send events whose `data` is longer than 8 bytes
 *                     (req_service_blob / payload(k) of the handler context; the words
 *                     travel as continuation records, include/simian_spec.h).  Default 0:
 *                     the window kernel then never looks at a record's flags
 *   "fused_advance"   size == 1: 1 = the last CTA of the window kernel computes the next window,
 *                     0 = a one-CTA launch of its own does (k_lbts): the window kernel's CTAs
 *                     then end without fence + ticket + barrier.  Default: by the number of
 *                     entity blocks (own launch from 592 on: config 2 15.07 -> 14.77 ms)
 *   "l2_persist"      2 (default): the hot metadata of the calendar -- cell head words, entity
 *                     core states, block page stacks -- is one allocation kept resident in
 *                     L2 by a persisting access-policy window on the engine's stream
 *                     (removed when the engine lets go of the stream); 1: the head words
 *                     only; 0: off.  Config 2: 15.76 -> 15.08 ms per step
 *   "split_sums"      models with stateful handlers: the sums a handler defers
 *                     (simian_defer_weighted_sum, the LANL receive handler) are computed by
 *                     k_sums behind the window kernel -- one global job array, a persistent
 *                     grid of warps -- instead of inside it (default 1; "sum_jobs" = the
 *                     array's capacity per window, default four per entity; passes that
 *                     find no room compute their sums in place)
 *   "pdl_exchange"    1: the launches of a device-driven multi-GPU window are programmatic
 *                     dependents of each other (default 0: measured equal)
 *   "split_handlers"  1: models of pure handlers run a window as three launches --
 *                     k_window_stage (dequeue, rank, park the events with their commit
 *                     indices, trace chain, recycle; 48 registers), k_window_rest
 *                     (bursts and redistribution passes in the one-kernel form) and
 *                     k_emit (handlers + appends, a persistent grid of warps) -- instead
 *                     of one; same results bit for bit.  Default 0: measured equal to
 *                     the one-kernel window on config 2         */
int simian_set_option(simian_engine *e, const char *key, double value);

/* Entity classes.  An entity class = the `name` of addEntity plus the JS class
 * object (simian.js:211-229).  stateBytes (multiple of 8) of per-entity user
 * state live in HBM, zero-initialised unless initState is given.            */
int simian_define_class(simian_engine *e, const char *className,
                        uint32_t stateBytes);
int simian_set_class_params(simian_engine *e, const char *className,
                            const void *params, uint32_t bytes);

/* addEntity(name, klass, num) for num = firstNum .. firstNum+count-1
 * (simian.js:211-229).  Only entities with getOffsetRank == rank are
 * instantiated (simian.js:221-223).  Instances of a class are added
 * contiguously from 0; SIMIAN_ERR_RUNNING while running (simian.js:215).     */
int simian_add_entities(simian_engine *e, const char *className,
                        uint32_t firstNum, uint32_t count,
                        const void *initState /* nullable, count*stateBytes */);

/* attachService(klass, name, fun)   simian.js:207-209 / entity.js:70-72 with
 * fun = a __device__ handler compiled into the library, looked up by name
 * ("phold_generate", "noop", ...).                                           */
int simian_attach_service(simian_engine *e, const char *className,
                          const char *serviceName, const char *deviceFnName);

/* id of a device function in the library's handler registry
 * (include/simian_handlers.def; a model author's own handlers are compiled in
 * with make HANDLERS=...), 0 when no handler of that name is registered.  Needs
 * no engine and no GPU.                                                       */
int simian_device_function_id(const char *deviceFnName);

/* interned id of a service name (the `name` field of an event)              */
int simian_intern_service(simian_engine *e, const char *serviceName);

/* global entity id of className(0); ids of a class are contiguous           */
int64_t simian_first_id(simian_engine *e, const char *className);

/* getBaseRank(name) = hash(name) % size   simian.js:192-194, hash.js:21-26  */
int simian_base_rank(simian_engine *e, const char *className);
double simian_hash_name(const char *name);
/* A model script that overrides Simian.getBaseRank (documented extension point,
 * Docs/README.Simian:92-100, simian.js:192-194): instances of className live on
 * rank (baseRank + num) % size.  Before the class's entities are added; the
 * same value on every rank.  getOffsetRank keeps the reference form
 * (simian.js:196-198), which the device-side routing implements.             */
int simian_set_base_rank(simian_engine *e, const char *className, int baseRank);

/* schedService(time, eventName, data, rx, rxId)             simian.js:170-190 */
int simian_sched_service(simian_engine *e, double time, const char *serviceName,
                         uint64_t data, const char *rx, uint32_t rxId);

/* the model-script loop
 *   for k < repeat: for i < count: schedService(time, name, data, rx, first+i)
 * (phold-noop-noMPI.js:47) built on the device, no host loop.               */
int simian_sched_service_bulk(simian_engine *e, double time,
                              const char *serviceName, uint64_t data,
                              const char *rx, uint32_t firstId, uint32_t count,
                              uint32_t repeat);

/* Batch injection from a HOST buffer of POD records (dst/src = global ids).
 * Same meaning as n schedService / pre-run reqService calls; this is the call
 * the end-to-end benchmark times (host->device copy included).              */
int simian_sched_events(simian_engine *e, const simian_event_t *hostEvents,
                        uint64_t n);

/* `new entityClass(name, out, engine, num)` bodies (simian.js:227): entity
 * constructors may run handler code and call reqService before run()
 * (pdes_lanl_benchmarkV8.js:255-256).  Runs the named device function once per
 * local instance at now = startTime; it is not an event.  expectedEvents: how
 * many events the constructors will send in total (pool sizing hint).       */
int simian_construct_entities(simian_engine *e, const char *className,
                              const char *deviceFnName, uint64_t expectedEvents);

/* Simian.run()                                               simian.js:99-168
 * size == 1: the whole window loop runs here.  size > 1: use the stepping
 * calls below (the host runs the collectives, mpi.cpp:364-476).             */
int simian_run(simian_engine *e, simian_stats *out);

/* ---- stepping interface for size > 1 (one window = simian.js:120-148) ---- */
/* allocate calendar / pools, ingest scheduled events, compute window 0      */
int simian_run_begin(simian_engine *e);
/* drain every local event below the window bound (simian.js:122-134)        */
int simian_window_process(simian_engine *e);
/* per-peer staged record counts of this window, to host (sndCounts,
 * mpi.cpp:21,461); also returns the device buffers to exchange.             */
int simian_window_outbox(simian_engine *e, uint32_t *hostCounts /* [size] */,
                         void **devOutbox, uint64_t *recordsPerPeerCapacity);
/* received records (device pointer, contiguous) -> local calendar
 * (simian.js:138-142)                                                        */
int simian_window_ingest(simian_engine *e, const void *devRecords, uint64_t n);
/* local minLeft candidate (simian.js:143) written to a device buffer of two
 * doubles {minLeft, -needRedistribution}; allreduce(MIN) it (simian.js:144) */
int simian_window_local_min(simian_engine *e, void **devMinVec);
/* consume the allreduced vector, set up the next window; *done = 1 when
 * globalMinLeft > endTime (simian.js:119)                                    */
int simian_window_advance(simian_engine *e, int *done);
int simian_run_end(simian_engine *e, simian_stats *out);

/* ---- device-driven exchange: peers' outboxes mapped over NVLink ----------
 * Replaces sendAndCount + alltoallSum + probe/recv (mpi.cpp:418-476,318-362)
 * without a host round trip: each rank publishes CUDA-IPC handles of its
 * (double-buffered) outbox, maps its peers', and after the window's
 * allreduce(MIN) (mpi.cpp:380 -- it also orders the exchange) PULLS the
 * records staged for it straight out of the peers' memory into its calendar.
 * Per window: simian_window_process, simian_window_publish_min,
 * simian_window_pull(wait), simian_window_advance_async(fromPeers) -- no
 * collective library in the loop -- or simian_window_local_min, an external
 * allreduce(MIN), simian_window_pull(0), simian_window_advance_async(0);
 * simian_poll once per batch of windows.  One process per GPU, all GPUs on one
 * NVLink/NVSwitch box.                                                        */
/* allocate the HBM layout and replay schedService now (idempotent)          */
int simian_setup(simian_engine *e);
/* blob == NULL: *written = blob size.  Otherwise fills the blob to publish. */
int simian_ipc_export(simian_engine *e, void *blob, uint64_t blobBytes, uint64_t *written);
int simian_ipc_import(simian_engine *e, int peerRank, const void *blob, uint64_t blobBytes);

/* The same connection for ranks that live in ONE process (every rank's engine on
 * one device: the multi-rank paths of a one-GPU test run): peer buffers are
 * wired by plain device pointers, no CUDA IPC.  Call on every ordered pair.   */
int simian_connect_local(simian_engine *e, int peerRank, simian_engine *peer);
/* simian_window_process with the LBTS candidate computed by the last CTA of
 * the same launch (no separate simian_window_local_min / publish_min launch);
 * publish != 0: and written to the peers' exchange slots                     */
int simian_window_process_lbts(simian_engine *e, int publish);
/* waitForPeers != 0: first wait (on the device) until every rank has
 * published this window's LBTS candidate, i.e. finished its window           */
int simian_window_pull(simian_engine *e, int waitForPeers);
/* Deferred pull (option "deferred_pull" = 1 before setup; device-driven exchange
 * only).  Senders stage a record that cannot be due in the receiver's next window
 * (time >= the sending window's bound + 2 minDelay) from the BACK of the outbox;
 * simian_window_pull then takes the front end (plus, after a jump over idle time,
 * the far records below the next bound) before the advance step, and the first
 * CTAs of the next simian_window_process_lbts launch pull the rest over NVLink
 * beside that window's entity blocks: same calls, same three launches per window. */
/* simian_window_local_min + allreduce(MIN), send half: the candidate is also
 * written into every peer's exchange slots over NVLink (mpi.cpp:364-390
 * without a collective library or the host)                                  */
int simian_window_publish_min(simian_engine *e);
/* device address of the two doubles simian_window_local_min writes           */
int simian_minvec_ptr(simian_engine *e, void **devMinVec);
/* simian_window_advance without the host read-back.  minFromPeers != 0: the
 * global minimum is taken over the exchange slots (receive half of the
 * peer-memory allreduce) instead of the vector at simian_minvec_ptr          */
int simian_window_advance_async(simian_engine *e, int minFromPeers);
/* device flags -> host: sticky error as return value, *done (simian.js:119) */
int simian_poll(simian_engine *e, int *done);

/* stream the engine launches on (a cudaStream_t); pass torch's current
 * stream so collectives and kernels are ordered                             */
int simian_set_stream(simian_engine *e, void *cudaStream);

/* ---- results: what the parity tests compare ------------------------------ */
/* `out` buffers in page-locked host memory (cudaHostAlloc / cudaHostRegister,
 * torch pin_memory) are filled by one DMA; pageable ones through a pinned
 * bounce buffer and a host copy.                                              */
/* per-entity committed-event counts of className(first .. first+count-1);
 * entries of entities living on other ranks are left untouched              */
int simian_read_counts(simian_engine *e, const char *className, uint32_t first,
                       uint32_t count, uint64_t *out);
int simian_read_trace_hashes(simian_engine *e, const char *className,
                             uint32_t first, uint32_t count, uint64_t *out);
/* The same for THIS rank's own instances only, compact and in local order:
 * out[j] belongs to instance *firstNum + j * *stride.  which: 0 counts,
 * 1 trace hashes.  out == NULL: size query (*nLocal).                        */
int simian_read_local(simian_engine *e, const char *className, int which, uint64_t *out,
                      uint64_t capacity, uint64_t *nLocal, uint32_t *firstNum,
                      uint32_t *stride);
/* getEntity(name, num) state read-back (simian.js:200-205)                  */
int simian_read_state(simian_engine *e, const char *className, uint32_t first,
                      uint32_t count, void *out);

/* Engines of the same shape reuse the HBM of destroyed ones (process-wide
 * cache, SIMIAN_B200_CACHE_GB, default 64); this returns it all to CUDA and
 * closes the cached CUDA-IPC mappings of peers' buffers: call it only while no
 * multi-rank engine of this process is between run_begin and run_end.  (When an
 * allocation fails the library drops its cached allocations by itself, never
 * the mappings.)                                                             */
void simian_trim_cache(void);

/* diagnostic only: summed SM cycles spent in each phase of the window kernel
 * by all CTAs (zeros unless the library was built with SIMIAN_PROFILE_PHASES) */
int simian_debug_phase_cycles(simian_engine *e, uint64_t *out12);

#ifdef __cplusplus
}
#endif
#endif /* SIMIAN_B200_H */
