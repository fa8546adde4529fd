/*
 * simian_oracle.cpp -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * A line-by-line CPU restatement of the reference's event-oriented service
 * path, used as the checker for the CUDA engine.  Only tests/, the smoke test
 * in __graft_entry__.py and bench.py's cpu_baseline / --impl reference legs may
 * build, load or call this file.  The product (simian_b200/) never does.
 *
 * Reference files restated (paths relative to /root/reference):
 *   SimianJS/eventQ.js:29-43   eventQ.push   -> EventQ::push
 *   SimianJS/eventQ.js:53-92   eventQ.pop    -> EventQ::pop
 *   SimianJS/simian.js:36-84   constructor   -> oracle_create / Rank
 *   SimianJS/simian.js:99-168  run           -> oracle_run
 *   SimianJS/simian.js:170-190 schedService  -> oracle_sched_service
 *   SimianJS/simian.js:192-198 getBaseRank/getOffsetRank -> World::offset_rank
 *   SimianJS/simian.js:211-229 addEntity     -> oracle_add_entities
 *   SimianJS/entity.js:35-68   reqService    -> HostCtx::req_service
 *   SimianJS/hash.js:21-26     hash          -> oracle_hash_name
 *   SimianJS/MasalaChai/mpi.cpp:418-476,364-390 (alltoallSum / sendAndCount /
 *       allreduce) -> the in-process multi-rank exchange in oracle_run.
 *
 * Third-party arithmetic absent from /root/reference: Math.log of the JS engine
 * (mozjs-50a1 per SimianJS/MasalaChai/Makefile:14) and libc rand()
 * (random.cpp:19-46).  Both are replaced, in the model, by the shared portable
 * log and counter-based RNG of include/simian_spec.h.
 *
 * PARITY PINNING: the reference's own tests hold no golden vectors for this
 * path (the scripts under SimianJS/Tests are timing scripts).  The oracle is pinned instead
 * against (a) the only fixtures the reference has -- the 6-element heap input
 * of SimianLua/eventQ.lua:62 and the sortedness check of
 * SimianJS/Tests/test.Q.js:14-21 -- and (b) outputs of the UNMODIFIED reference
 * engine SimianPie/simian.py run in the build container by
 * oracle/pin_simianpie.py, committed under tests/golden/.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <chrono>
#include <map>
#include <mutex>
#include <thread>
#include <string>
#include <vector>

#include "../include/simian_models.h"
#include "../include/simian_spec.h"

namespace {

enum {
  ORACLE_OK = 0,
  ORACLE_ERR_LOOKAHEAD = 1,    /* entity.js:44  */
  ORACLE_ERR_IDLE_SEND = 2,    /* entity.js:42  */
  ORACLE_ERR_OUT_OF_ORDER = 3, /* simian.js:125 */
  ORACLE_ERR_RUNNING = 4,      /* simian.js:215 */
  ORACLE_ERR_UNKNOWN = 5,
};

/* ---- SimianJS/eventQ.js : binary heap, head = least time ----------------- */
struct QItem {
  simian_event_t e;
  uint64_t ec; /* push counter: only tie_mode 2 looks at it */
};

struct EventQ {
  std::vector<QItem> list;
  uint64_t ec = 0;
  /* 0: literal JS (time only, eventQ.js:36,62,66); 1: the engine's defined order
   * (time, handler, src, seq); 2: FIFO among equal times, the order of the
   * reference's Python twin (heapq of (time, ec, e), SimianPie/simian.py:286-287) */
  int tie_mode;

  /* `a.time < b.time` (eventQ.js:36) */
  bool lt(const QItem &a, const QItem &b) const {
    if (tie_mode == 0) return a.e.time < b.e.time;
    if (tie_mode == 2) return a.e.time != b.e.time ? a.e.time < b.e.time : a.ec < b.ec;
    return simian_event_before(a.e.time, a.e.handler, a.e.src, a.e.seq, b.e.time,
                               b.e.handler, b.e.src, b.e.seq) != 0;
  }

  void push(const simian_event_t &ev) { /* eventQ.js:29-43 */
    QItem item{ev, ec++};
    size_t curPos, parentPos = list.size();
    list.push_back(item);
    while (parentPos > 0) {
      curPos = parentPos;
      parentPos = parentPos / 2; /* floor(parentPos/2) on a 0-based array */
      if (lt(list[curPos], list[parentPos])) {
        QItem temp = list[curPos];
        list[curPos] = list[parentPos];
        list[parentPos] = temp;
      } else
        break;
    }
  }

  double peek() const { return list[0].e.time; } /* eventQ.js:45-47 */
  size_t length() const { return list.size(); }   /* eventQ.js:49-51 */

  simian_event_t pop() { /* eventQ.js:53-92 */
    size_t length = list.size() - 1;
    QItem temp = list[0];
    list[0] = list[length];
    list[length] = temp;
    QItem topItem = list.back();
    list.pop_back();
    size_t curPos = 0, left = 1, right = 2, min = 0;
    int state = 0;
    while (true) {
      /* `list[min].time > list[left].time` (eventQ.js:62,66) */
      if ((left < length) && lt(list[left], list[min])) {
        state = 1;
        min = left;
      }
      if ((right < length) && lt(list[right], list[min])) {
        state = 2;
        min = right;
      }
      if (state == 0)
        break;
      else if (state == 1) {
        temp = list[curPos];
        list[curPos] = list[left];
        list[left] = temp;
        curPos = left;
        left *= 2;
        right = left + 1;
        state = 0;
      } else {
        temp = list[curPos];
        list[curPos] = list[right];
        list[right] = temp;
        curPos = right;
        left = right * 2;
        right = left + 1;
        state = 0;
      }
    }
    return topItem.e;
  }
};

struct ClassInfo {
  std::string name;
  uint32_t first_id = 0; /* global id of instance 0 */
  uint32_t count = 0;
  uint32_t state_bytes = 0;
  double name_hash = 0; /* hash.js */
  int base_rank_override = -1; /* >= 0: the model script's own getBaseRank */
  std::vector<uint8_t> params;
  std::map<uint16_t, int> fn_of_service; /* service id -> builtin fn */
  std::vector<uint8_t> state;            /* count * state_bytes */
};

struct Last {
  double time;
  uint32_t handler, src;
  uint64_t data;
  bool valid;
};

struct World;

/* one MPI rank of the reference = one Simian instance (simian.js:36-84) */
struct Rank {
  EventQ q;
  double now;
  uint64_t numEvents = 0, synchEvents = 0;
  /* parity statistics, per rank so that ranks can run on separate threads */
  uint64_t ties = 0, dist_ties = 0, emitted = 0, dropped = 0;
  int error = 0;
  std::string errmsg;
  std::vector<uint32_t> sndCounts;                 /* mpi.cpp:21 */
  std::vector<std::vector<simian_event_t>> outbox; /* MPI_Send per event */
};

struct World {
  std::string name;
  double startTime, endTime, minDelay, infTime;
  int size, tie_mode;
  uint64_t seed = 1;
  bool running = false;
  int error = 0;
  std::string errmsg;
  std::vector<Rank> ranks;
  std::vector<ClassInfo> classes;
  std::map<std::string, int> class_ix;
  std::map<std::string, uint16_t> service_ix;
  /* per global entity id */
  std::vector<uint64_t> count, hash;
  std::vector<Last> last;
  std::vector<int> class_of;
  uint32_t n_ids = 0;
  uint32_t sched_seq = 0;
  /* stats */
  uint64_t windows = 0, ties = 0, dist_ties = 0, emitted = 0, dropped = 0;
  double last_time = 0, run_seconds = 0;
  int threads = 1; /* ranks stepped concurrently by oracle_run (1 = sequential) */
  /* payload words of the events in flight, keyed (src << 32 | seq); shared by the
   * ranks of this process (one rank per PROCESS -- the gloo form -- has no such
   * table: payloads are not supported there) */
  std::map<uint64_t, std::vector<uint64_t>> blobs;
  std::mutex blob_mu;

  /* fold the per-rank statistics into the world's (after a window / at the end) */
  void gather_rank_stats() {
    for (auto &r : ranks) {
      ties += r.ties;
      dist_ties += r.dist_ties;
      emitted += r.emitted;
      dropped += r.dropped;
      r.ties = r.dist_ties = r.emitted = r.dropped = 0;
      if (r.error && !error) {
        error = r.error;
        errmsg = r.errmsg;
      }
    }
  }

  uint16_t intern_service(const char *s) {
    auto it = service_ix.find(s);
    if (it != service_ix.end()) return it->second;
    uint16_t id = (uint16_t)service_ix.size();
    service_ix[s] = id;
    return id;
  }
  /* getBaseRank: hash(name) % size  (simian.js:192-194) */
  int base_rank(const ClassInfo &c) const {
    if (c.base_rank_override >= 0) return c.base_rank_override % size; /* user getBaseRank */
    return (int)fmod(c.name_hash, (double)size);
  }
  /* getOffsetRank: (baseRanks[name] + num) % size  (simian.js:196-198) */
  int offset_rank(uint32_t gid) const {
    const ClassInfo &c = classes[class_of[gid]];
    return (int)(((uint64_t)base_rank(c) + (gid - c.first_id)) % (uint64_t)size);
  }
  void fail(int code, const std::string &m) {
    if (!error) {
      error = code;
      errmsg = m;
    }
  }
};

/* what a handler reaches through `self` (see simian_models.h) */
struct HostCtx {
  World *w;
  Rank *r;
  const simian_event_t *ev;
  uint32_t self;
  uint64_t commit_idx, key;
  uint32_t n_emit;
  ClassInfo *cls;

  double now() const { return r->now; }
  uint32_t self_id() const { return self; }
  uint64_t draw(uint32_t i) const { return simian_rng_draw(key, i); }
  uint16_t handler() const { return ev ? ev->handler : 0; }
  const void *params() const { return cls->params.data(); }
  void *state() const {
    return cls->state_bytes
               ? (void *)(cls->state.data() +
                          (size_t)(self - cls->first_id) * cls->state_bytes)
               : nullptr;
  }
  uint32_t src() const { return ev ? ev->src : SIMIAN_NULL_ENTITY; }
  uint64_t data() const { return ev ? ev->data : 0; }

  /* payloads beyond 8 bytes (entity.js:52-60: `data` is any object): the event
   * carries the number of words, the words stay in a table of the world keyed by
   * (src, seq) -- the names the engine's continuation records carry too */
  uint32_t payload_words() const {
    return (ev && (ev->flags & SIMIAN_EVF_BLOB)) ? (uint32_t)ev->data : 0u;
  }
  uint64_t payload(uint32_t k) {
    if (k < payload_words()) {
      std::lock_guard<std::mutex> g(w->blob_mu);
      auto it = w->blobs.find(((uint64_t)ev->src << 32) | ev->seq);
      if (it != w->blobs.end() && k < it->second.size()) return it->second[k];
    }
    if (!r->error) {
      r->error = ORACLE_ERR_UNKNOWN;
      r->errmsg = "payload word missing";
    }
    return 0;
  }
  void req_service_blob(double offset, uint16_t service, const uint64_t *words, uint32_t nwords,
                        uint32_t rx) {
    req_service_blob_fn(offset, service, nwords, rx, [words](uint32_t k) { return words[k]; });
  }
  template <class F>
  void req_service_blob_fn(double offset, uint16_t service, uint32_t nwords, uint32_t rx, F word) {
    const uint32_t seq = simian_emit_seq(commit_idx, n_emit);
    {
      std::vector<uint64_t> v(nwords);
      for (uint32_t k = 0; k < nwords; k++) v[k] = word(k);
      std::lock_guard<std::mutex> g(w->blob_mu);
      w->blobs[((uint64_t)self << 32) | seq] = std::move(v);
    }
    send_(offset, service, (uint64_t)nwords, rx, SIMIAN_EVF_BLOB);
  }

  /* Entity.reqService  (entity.js:35-68) */
  void req_service(double offset, uint16_t service, uint64_t data, uint32_t rx) {
    send_(offset, service, data, rx, 0);
  }
  void send_(double offset, uint16_t service, uint64_t data, uint32_t rx, uint16_t flags) {
    uint32_t k = n_emit++;
    if ((rx != SIMIAN_NULL_ENTITY) && (offset < w->minDelay)) { /* :41 */
      if (!r->error) {
        r->error = w->running ? ORACLE_ERR_LOOKAHEAD : ORACLE_ERR_IDLE_SEND;
        r->errmsg = w->running ? "attempted to send with too little delay"
                               : "Cannot send an event when Simian is idle";
      }
      return;
    }
    double time = r->now + offset; /* :47 */
    if (time > w->endTime) {       /* :48 */
      r->dropped++;
      return;
    }
    if (rx == SIMIAN_NULL_ENTITY) rx = self; /* :50-51 */
    simian_event_t e;                        /* :52-60 */
    e.time = time;
    e.dst = rx;
    e.src = self;
    e.handler = service;
    e.flags = flags;
    e.seq = simian_emit_seq(commit_idx, k);
    e.data = data;
    r->emitted++;
    int recvRank = w->offset_rank(rx); /* :62 */
    int myRank = (int)(r - &w->ranks[0]);
    if (recvRank == myRank)
      r->q.push(e); /* :64 */
    else {          /* :66 sendAndCount -> mpi.cpp:439-476 */
      r->outbox[recvRank].push_back(e);
      r->sndCounts[recvRank]++;
    }
  }
};

void dispatch(World *w, Rank *r, const simian_event_t &ev) {
  uint32_t gid = ev.dst;
  ClassInfo &c = w->classes[w->class_of[gid]];
  /* `entity[curEvent.name]` (simian.js:130) */
  auto it = c.fn_of_service.find(ev.handler);
  if (it == c.fn_of_service.end()) {
    if (!r->error) {
      r->error = ORACLE_ERR_UNKNOWN;
      r->errmsg = "service is not a function";
    }
    return;
  }
  /* commit bookkeeping: what parity is measured on */
  uint64_t idx = w->count[gid];
  Last &l = w->last[gid];
  if (l.valid && l.time == ev.time) {
    r->ties++;
    if (l.handler != ev.handler || l.src != ev.src || l.data != ev.data)
      r->dist_ties++;
  }
  l.time = ev.time;
  l.handler = ev.handler;
  l.src = ev.src;
  l.data = ev.data;
  l.valid = true;
  w->hash[gid] =
      simian_trace_step(w->hash[gid], ev.time, ev.handler, ev.src, ev.data);
  w->count[gid] = idx + 1;

  HostCtx ctx;
  ctx.w = w;
  ctx.r = r;
  ctx.ev = &ev;
  ctx.self = gid;
  ctx.commit_idx = idx;
  ctx.key = simian_rng_key(w->seed, gid, idx);
  ctx.n_emit = 0;
  ctx.cls = &c;
  switch (it->second) { // generated from include/simian_handlers.def
#define SIMIAN_HANDLER(id, name, func, pure, ptype) \
  case id:                                          \
    func(ctx);                                      \
    break;
#include "../include/simian_handlers.def"
#undef SIMIAN_HANDLER
    default:
      if (!r->error) {
        r->error = ORACLE_ERR_UNKNOWN;
        r->errmsg = "unknown builtin function";
      }
  }
}

int fn_by_name(const char *n) { return simian_fn_by_name(n); }

}  // namespace

extern "C" {

typedef struct oracle_stats {
  uint64_t total_events; /* simian.js:153 */
  uint64_t windows;      /* iterations of simian.js:119 */
  uint64_t synch_events; /* simian.js:145,154 */
  uint64_t digest, state_checksum;
  uint64_t ties, distinguishable_ties;
  uint64_t emitted, dropped_past_end;
  double last_time; /* engine.now after run */
  double run_seconds;
  int32_t error;
  int32_t pad_;
} oracle_stats;

/* hash.js:21-26 -- djb2 in doubles, iterating the string backwards */
double oracle_hash_name(const char *str) {
  double res = 5381;
  size_t i = strlen(str);
  while (i--) res = 33 * res + (double)(unsigned char)str[i];
  return res;
}

void *oracle_create(const char *simName, double startTime, double endTime,
                    double minDelay, int size, int tie_mode) {
  World *w = new World();
  w->name = simName;
  w->startTime = startTime;
  w->endTime = (endTime != 0.0) ? endTime : 1e100; /* simian.js:41 */
  w->minDelay = (minDelay != 0.0) ? minDelay : 1;  /* simian.js:42 */
  w->infTime = simian_inf_time(endTime, minDelay); /* simian.js:64 raw args */
  w->size = size < 1 ? 1 : size;
  w->tie_mode = tie_mode;
  w->ranks.resize(w->size);
  for (auto &r : w->ranks) {
    r.q.tie_mode = tie_mode;
    r.now = startTime; /* simian.js:44 */
    r.sndCounts.assign(w->size, 0);
    r.outbox.resize(w->size);
  }
  return w;
}

void oracle_destroy(void *o) { delete (World *)o; }
void oracle_set_seed(void *o, uint64_t seed) { ((World *)o)->seed = seed; }
/* host threads oracle_run steps the ranks on (default 1) */
void oracle_set_threads(void *o, int n) { ((World *)o)->threads = n < 1 ? 1 : n; }
const char *oracle_last_error(void *o) { return ((World *)o)->errmsg.c_str(); }

int oracle_define_class(void *o, const char *className, uint32_t stateBytes) {
  World *w = (World *)o;
  if (w->class_ix.count(className)) return w->class_ix[className];
  ClassInfo c;
  c.name = className;
  c.state_bytes = stateBytes;
  c.name_hash = oracle_hash_name(className);
  c.first_id = w->n_ids;
  w->classes.push_back(c);
  w->class_ix[className] = (int)w->classes.size() - 1;
  return (int)w->classes.size() - 1;
}

int oracle_set_class_params(void *o, const char *className, const void *p,
                            uint32_t bytes) {
  World *w = (World *)o;
  if (!w->class_ix.count(className)) return ORACLE_ERR_UNKNOWN;
  ClassInfo &c = w->classes[w->class_ix[className]];
  c.params.assign((const uint8_t *)p, (const uint8_t *)p + bytes);
  return 0;
}

/* addEntity for num = first .. first+count-1 (simian.js:211-229).  Instances
 * of one class must be added contiguously and classes one after another.    */
int oracle_add_entities(void *o, const char *className, uint32_t first,
                        uint32_t count) {
  World *w = (World *)o;
  if (w->running) { /* simian.js:215 */
    w->fail(ORACLE_ERR_RUNNING, "Cannot add new entity when Simian is running!");
    return ORACLE_ERR_RUNNING;
  }
  if (!w->class_ix.count(className)) oracle_define_class(o, className, 0);
  int ci = w->class_ix[className];
  ClassInfo &c = w->classes[ci];
  if (c.count == 0) c.first_id = w->n_ids; /* ids are dealt when the first instance arrives */
  if (first != c.count || c.first_id + c.count != w->n_ids) {
    w->fail(ORACLE_ERR_UNKNOWN, "entities must be added contiguously");
    return ORACLE_ERR_UNKNOWN;
  }
  c.count += count;
  w->n_ids += count;
  w->count.resize(w->n_ids, 0);
  w->hash.resize(w->n_ids, 0);
  w->last.resize(w->n_ids, Last{0, 0, 0, 0, false});
  w->class_of.resize(w->n_ids, ci);
  c.state.resize((size_t)c.count * c.state_bytes, 0);
  return 0;
}

int oracle_first_id(void *o, const char *className) {
  World *w = (World *)o;
  if (!w->class_ix.count(className)) return -1;
  return (int)w->classes[w->class_ix[className]].first_id;
}

/* a model script overriding Simian.getBaseRank (simian.js:192-194) */
int oracle_set_base_rank(void *o, const char *className, int baseRank) {
  World *w = (World *)o;
  if (!w->class_ix.count(className) || baseRank < 0) return -1;
  w->classes[w->class_ix[className]].base_rank_override = baseRank;
  return 0;
}

int oracle_base_rank(void *o, const char *className) {
  World *w = (World *)o;
  if (!w->class_ix.count(className)) return -1;
  return w->base_rank(w->classes[w->class_ix[className]]);
}

int oracle_intern_service(void *o, const char *serviceName) {
  return ((World *)o)->intern_service(serviceName);
}

/* attachService(klass, name, fun)  (simian.js:207-209) with fun = a named
 * builtin function */
int oracle_attach_service(void *o, const char *className,
                          const char *serviceName, const char *fnName) {
  World *w = (World *)o;
  if (!w->class_ix.count(className)) oracle_define_class(o, className, 0);
  int fn = fn_by_name(fnName);
  if (!fn) return ORACLE_ERR_UNKNOWN;
  w->classes[w->class_ix[className]].fn_of_service[w->intern_service(
      serviceName)] = fn;
  return 0;
}

/* schedService(time, eventName, data, rx, rxId)  (simian.js:170-190) */
int oracle_sched_service(void *o, double time, const char *serviceName,
                         uint64_t data, const char *rx, uint32_t rxId) {
  World *w = (World *)o;
  uint32_t seq = w->sched_seq++;
  if (time > w->endTime) return 0; /* :173 */
  if (!w->class_ix.count(rx)) return ORACLE_ERR_UNKNOWN;
  ClassInfo &c = w->classes[w->class_ix[rx]];
  if (rxId >= c.count) return ORACLE_ERR_UNKNOWN;
  uint32_t gid = c.first_id + rxId;
  int recvRank = w->offset_rank(gid); /* :175 */
  simian_event_t e;                   /* :178-186 */
  e.time = time;
  e.dst = gid;
  e.src = SIMIAN_NULL_ENTITY;
  e.handler = w->intern_service(serviceName);
  e.flags = SIMIAN_EVF_SCHED;
  e.seq = seq;
  e.data = data;
  /* every rank runs the same script; only the owner pushes (:177,188) */
  w->ranks[recvRank].q.push(e);
  return 0;
}

/* `new entityClass(...)` bodies (simian.js:227): entity constructors may run
 * handler code and call reqService before run() (pdes_lanl_benchmarkV8.js:
 * 255-256).  Runs the named constructor once per instance, in num order, on
 * the owning rank, at now = startTime; it is not an event (no commit).       */
int oracle_construct_entities(void *o, const char *className, const char *fnName) {
  World *w = (World *)o;
  if (!w->class_ix.count(className)) return ORACLE_ERR_UNKNOWN;
  int fn = fn_by_name(fnName);
  ClassInfo &c = w->classes[w->class_ix[className]];
  for (uint32_t i = 0; i < c.count && !w->error; i++) {
    uint32_t gid = c.first_id + i;
    HostCtx ctx;
    ctx.w = w;
    ctx.r = &w->ranks[w->offset_rank(gid)];
    ctx.ev = nullptr;
    ctx.self = gid;
    ctx.commit_idx = SIMIAN_CTOR_COMMIT;
    ctx.key = simian_rng_key(w->seed, gid, SIMIAN_CTOR_COMMIT);
    ctx.n_emit = 0;
    ctx.cls = &c;
    switch (fn) {
      case SIMIAN_FN_LANL_CONSTRUCT:
        simian_lanl_construct(ctx);
        break;
      default:
        w->fail(ORACLE_ERR_UNKNOWN, "unknown constructor function");
    }
    w->gather_rank_stats();
  }
  return w->error;
}

/* the model-script loop `for i: for k<repeat: schedService(time,...,rx,i)` */
int oracle_sched_service_bulk(void *o, double time, const char *serviceName,
                              uint64_t data, const char *rx, uint32_t first,
                              uint32_t count, uint32_t repeat) {
  for (uint32_t k = 0; k < repeat; k++)
    for (uint32_t i = 0; i < count; i++) {
      int rc = oracle_sched_service(o, time, serviceName, data, rx, first + i);
      if (rc) return rc;
    }
  return 0;
}

/* inject pre-built records (same meaning as schedService / a pre-run
 * reqService; used for batch inputs) */
int oracle_sched_events(void *o, const simian_event_t *ev, uint64_t n) {
  World *w = (World *)o;
  for (uint64_t i = 0; i < n; i++) {
    if (ev[i].time > w->endTime) continue;
    if (ev[i].dst >= w->n_ids) return ORACLE_ERR_UNKNOWN;
    w->ranks[w->offset_rank(ev[i].dst)].q.push(ev[i]);
  }
  return 0;
}

/* simian.js:122-134 for one rank: drain every event below the epoch */
static void rank_window(World *w, Rank &r, double epoch) {
  while (r.q.length() && (r.q.peek() < epoch) && !r.error) { /* :122 */
    simian_event_t cur = r.q.pop();                          /* :123 */
    if (r.now > cur.time) {                                  /* :124-125 */
      r.error = ORACLE_ERR_OUT_OF_ORDER;
      r.errmsg = "Out of order event";
      break;
    }
    r.now = cur.time; /* :126 */
    dispatch(w, &r, cur);
    if (r.error) break;
    r.numEvents++; /* :133 */
  }
}

/* simian.js:137-143 for one rank: receive what the others sent, then minLeft.
 * alltoallSum (mpi.cpp:418-437) then probe/recv/push (:138-142).
 * MPI_Probe(ANY_SOURCE) order is unspecified in the reference; here: by source
 * rank, then send order.                                                     */
static double rank_exchange(World *w, int ri) {
  Rank &r = w->ranks[ri];
  for (int src = 0; src < w->size; src++) {
    std::vector<simian_event_t> &box = w->ranks[src].outbox[ri];
    for (const simian_event_t &e : box) r.q.push(e);
    box.clear();
    w->ranks[src].sndCounts[ri] = 0;
  }
  r.synchEvents += 2;                               /* :145 */
  return r.q.length() ? r.q.peek() : w->infTime;   /* :143 */
}

namespace {
/* sense-reversing barrier for the rank threads (the MPI collectives of the
 * reference are the synchronisation points of its ranks, mpi.cpp:422,380) */
struct SpinBarrier {
  std::atomic<int> count{0}, sense{0};
  int n = 1;
  void wait() {
    int s = sense.load(std::memory_order_acquire);
    if (count.fetch_add(1, std::memory_order_acq_rel) == n - 1) {
      count.store(0, std::memory_order_relaxed);
      sense.store(s ^ 1, std::memory_order_release);
    } else {
      int spins = 0;
      while (sense.load(std::memory_order_acquire) == s)
        if (++spins > 2000) std::this_thread::yield();
    }
  }
};
}  // namespace

/* Simian.run  (simian.js:99-168).  All `size` ranks in lock-step: one after the
 * other on the calling thread (threads == 1), or -- like the reference under
 * mpirun, one process per core -- on `threads` host threads, ranks dealt
 * round-robin, meeting at the two collectives of the window.  Same results.  */
int oracle_run(void *o, oracle_stats *out) {
  World *w = (World *)o;
  auto t0 = std::chrono::steady_clock::now();
  const double endTime = w->endTime, minDelay = w->minDelay,
               infTime = w->infTime;
  const int size = w->size;
  w->running = true;                   /* :117 */
  double globalMinLeft = w->startTime; /* :118 */
  int nthreads = w->threads < size ? w->threads : size;
  if (nthreads <= 1 || size == 1) {
    while (globalMinLeft <= endTime && !w->error) { /* :119 */
      double epoch = globalMinLeft + minDelay;      /* :120 */
      w->windows++;
      for (int ri = 0; ri < size; ri++) rank_window(w, w->ranks[ri], epoch);
      w->gather_rank_stats();
      if (w->error) break;
      if (size > 1) { /* :136-145 */
        double gmin = infTime;
        for (int ri = 0; ri < size; ri++) {
          double minLeft = rank_exchange(w, ri);
          if (minLeft < gmin) gmin = minLeft; /* :144 MIN */
        }
        globalMinLeft = gmin;
      } else {
        Rank &r = w->ranks[0];
        globalMinLeft = r.q.length() ? r.q.peek() : infTime; /* :147 */
      }
    }
  } else {
    SpinBarrier bar;
    bar.n = nthreads;
    std::vector<double> minLeft(size, infTime);
    std::atomic<int> stop{0};
    double gml = globalMinLeft;
    auto worker = [&](int t) {
      for (;;) {
        double epoch = gml + minDelay; /* :120 */
        for (int ri = t; ri < size; ri += nthreads) rank_window(w, w->ranks[ri], epoch);
        bar.wait(); /* every MPI_Send of the window has been posted */
        for (int ri = t; ri < size; ri += nthreads) minLeft[ri] = rank_exchange(w, ri);
        bar.wait(); /* MPI_Allreduce(MIN) */
        if (t == 0) {
          w->windows++;
          w->gather_rank_stats();
          double gmin = infTime;
          for (int ri = 0; ri < size; ri++)
            if (minLeft[ri] < gmin) gmin = minLeft[ri];
          gml = gmin;
          if (!(gml <= endTime) || w->error) stop.store(1); /* :119 */
        }
        bar.wait();
        if (stop.load()) break;
      }
    };
    if (globalMinLeft <= endTime) {
      std::vector<std::thread> th;
      for (int t = 1; t < nthreads; t++) th.emplace_back(worker, t);
      worker(0);
      for (auto &x : th) x.join();
    }
  }
  w->running = false;
  w->run_seconds =
      std::chrono::duration<double>(std::chrono::steady_clock::now() - t0)
          .count();
  if (out) {
    memset(out, 0, sizeof(*out));
    double last = w->startTime;
    for (auto &r : w->ranks) {
      out->total_events += r.numEvents; /* :153 allreduce SUM */
      out->synch_events += r.synchEvents;
      if (r.now > last) last = r.now;
    }
    out->windows = w->windows;
    uint64_t dg = 0, sc = 0;
    for (uint32_t g = 0; g < w->n_ids; g++)
      dg += simian_digest_term(g, w->count[g], w->hash[g]);
    for (auto &c : w->classes) {
      for (uint32_t i = 0; i < c.count; i++) {
        uint64_t acc = simian_mix64((uint64_t)(c.first_id + i) + 1);
        for (uint32_t b = 0; b + 8 <= c.state_bytes; b += 8) {
          uint64_t word;
          memcpy(&word, &c.state[(size_t)i * c.state_bytes + b], 8);
          acc = simian_state_fold(acc, word);
        }
        sc += acc;
      }
    }
    out->digest = dg;
    out->state_checksum = sc;
    out->ties = w->ties;
    out->distinguishable_ties = w->dist_ties;
    out->emitted = w->emitted;
    out->dropped_past_end = w->dropped;
    out->last_time = last;
    out->run_seconds = w->run_seconds;
    out->error = w->error;
  }
  return w->error;
}

int oracle_read_counts(void *o, const char *className, uint64_t *out) {
  World *w = (World *)o;
  if (!w->class_ix.count(className)) return ORACLE_ERR_UNKNOWN;
  ClassInfo &c = w->classes[w->class_ix[className]];
  memcpy(out, &w->count[c.first_id], (size_t)c.count * 8);
  return 0;
}

int oracle_read_trace_hashes(void *o, const char *className, uint64_t *out) {
  World *w = (World *)o;
  if (!w->class_ix.count(className)) return ORACLE_ERR_UNKNOWN;
  ClassInfo &c = w->classes[w->class_ix[className]];
  memcpy(out, &w->hash[c.first_id], (size_t)c.count * 8);
  return 0;
}

int oracle_read_state(void *o, const char *className, void *out) {
  World *w = (World *)o;
  if (!w->class_ix.count(className)) return ORACLE_ERR_UNKNOWN;
  ClassInfo &c = w->classes[w->class_ix[className]];
  memcpy(out, c.state.data(), c.state.size());
  return 0;
}

/* ---- one rank at a time, collectives done by the caller -------------------
 * The same window loop as oracle_run (simian.js:118-149), cut at the places
 * where the reference crosses into MPI (simian.js:137,140,144), so that the
 * test suite can run `size` processes over gloo, each stepping ITS rank of an
 * identically constructed world, and exercise the product's host-side window
 * driver (simian_b200/distributed.py) without a GPU.                          */
void oracle_step_begin(void *o) {
  World *w = (World *)o;
  w->running = true; /* simian.js:117 */
}

/* simian.js:120-134 for one rank */
int oracle_step_process(void *o, int rank, double globalMinLeft) {
  World *w = (World *)o;
  Rank &r = w->ranks[rank];
  double epoch = globalMinLeft + w->minDelay; /* :120 */
  w->windows++; /* each process steps exactly one rank of its world */
  rank_window(w, r, epoch);
  w->gather_rank_stats();
  return w->error;
}

/* sndCounts[peer] and the staged events (mpi.cpp:439-476); clears them */
uint64_t oracle_step_outbox(void *o, int rank, int peer, simian_event_t *out,
                            uint64_t cap) {
  World *w = (World *)o;
  std::vector<simian_event_t> &box = w->ranks[rank].outbox[peer];
  uint64_t n = box.size();
  if (out) {
    if (n > cap) n = cap;
    if (n) memcpy(out, box.data(), n * sizeof(simian_event_t));
    box.clear();
    w->ranks[rank].sndCounts[peer] = 0;
  }
  return n;
}

/* probe/recv/push (simian.js:138-142) */
void oracle_step_ingest(void *o, int rank, const simian_event_t *ev, uint64_t n) {
  World *w = (World *)o;
  for (uint64_t i = 0; i < n; i++) w->ranks[rank].q.push(ev[i]);
}

/* minLeft (simian.js:143) */
double oracle_step_min_left(void *o, int rank) {
  World *w = (World *)o;
  Rank &r = w->ranks[rank];
  r.synchEvents += 2; /* :145 */
  return r.q.length() ? r.q.peek() : w->infTime;
}

/* per-rank results after the loop */
int oracle_step_end(void *o, int rank, oracle_stats *out) {
  World *w = (World *)o;
  w->running = false;
  memset(out, 0, sizeof(*out));
  Rank &r = w->ranks[rank];
  out->total_events = r.numEvents;
  out->synch_events = r.synchEvents;
  out->windows = w->windows;
  uint64_t dg = 0;
  for (uint32_t g = 0; g < w->n_ids; g++)
    if (w->offset_rank(g) == rank) dg += simian_digest_term(g, w->count[g], w->hash[g]);
  out->digest = dg;
  out->ties = w->ties;
  out->distinguishable_ties = w->dist_ties;
  out->emitted = w->emitted;
  out->dropped_past_end = w->dropped;
  out->last_time = r.now;
  out->error = w->error;
  return w->error;
}

double oracle_start_time(void *o) { return ((World *)o)->startTime; }
double oracle_end_time(void *o) { return ((World *)o)->endTime; }
int oracle_owner_rank(void *o, uint32_t gid) { return ((World *)o)->offset_rank(gid); }

/* ---- bare eventQ access, for pinning the heap against the reference's
 * fixtures (SimianLua/eventQ.lua:62, SimianJS/Tests/test.Q.js:14-21) ------- */
void *oracle_q_new(int tie_mode) {
  EventQ *q = new EventQ();
  q->tie_mode = tie_mode;
  return q;
}
void oracle_q_free(void *q) { delete (EventQ *)q; }
void oracle_q_push(void *q, double time, uint64_t tag) {
  simian_event_t e;
  memset(&e, 0, sizeof e);
  e.time = time;
  e.data = tag;
  ((EventQ *)q)->push(e);
}
uint64_t oracle_q_len(void *q) { return ((EventQ *)q)->length(); }
double oracle_q_peek(void *q) { return ((EventQ *)q)->peek(); }
double oracle_q_pop(void *q, uint64_t *tag) {
  simian_event_t e = ((EventQ *)q)->pop();
  if (tag) *tag = e.data;
  return e.time;
}

/* ---- spec probes (so Python tests can pin the header's arithmetic) ------- */
double oracle_spec_log(double x) { return simian_log(x); }
uint64_t oracle_spec_mix64(uint64_t x) { return simian_mix64(x); }
uint64_t oracle_spec_rng_key(uint64_t s, uint32_t e, uint64_t c) {
  return simian_rng_key(s, e, c);
}
uint64_t oracle_spec_rng_draw(uint64_t k, uint32_t i) {
  return simian_rng_draw(k, i);
}
double oracle_spec_u01(uint64_t x) { return simian_u01(x); }
double oracle_spec_u01_open(uint64_t x) { return simian_u01_open(x); }
double oracle_spec_powi(double b, uint64_t n) { return simian_powi(b, n); }
double oracle_spec_exponential(uint64_t d, double l) {
  return simian_exponential(d, l);
}
uint64_t oracle_spec_trace_step(uint64_t h, double t, uint32_t hd, uint32_t s,
                                uint64_t d) {
  return simian_trace_step(h, t, hd, s, d);
}
uint64_t oracle_spec_digest_term(uint32_t e, uint64_t c, uint64_t h) {
  return simian_digest_term(e, c, h);
}

} /* extern "C" */
