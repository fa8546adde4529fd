/*
 * simian_models.h -- the built-in service handlers ("user code" of the
 * reference's model scripts) written ONCE as C++ templates over a handler
 * context, and instantiated by both the CUDA engine (device context) and the
 * CPU oracle (host context).  The engine semantics (queue, window loop, emit
 * path) are NOT here: those are implemented separately on each side.
 *
 * Context concept (what a reference handler can reach through `self`):
 *   double   now()                      engine.now            simian.js:126
 *   uint32_t self_id()                  global id of `self`   entity.js:24-27
 *   uint64_t draw(uint32_t i)           i-th RNG draw of this committed event,
 *                                       = simian_rng_draw(simian_rng_key(engine
 *                                       seed, self_id, commit index), i)
 *   uint16_t handler()                  id of the service being executed
 *   const void* params()                per-class model parameters
 *   void*    state()                    per-entity user state (may be null)
 *   uint32_t src()                      tx/txId of the event  simian.js:131
 *   uint64_t data()                     the event's payload   simian.js:131
 *   void req_service(double offset, uint16_t service, uint64_t data,
 *                    uint32_t rx)       self.reqService(...)  entity.js:35-68
 *        rx == SIMIAN_NULL_ENTITY  <=>  rx omitted (send to self, no lookahead
 *        check, entity.js:41,50-51)
 *   payloads beyond the 8-byte word (entity.js:52-60: `data` is any object):
 *   void req_service_blob(double offset, uint16_t service, const uint64_t *words,
 *                         uint32_t nwords, uint32_t rx)
 *   void req_service_blob_fn(offset, service, nwords, rx, word)   word(k) -> uint64_t
 *   uint32_t payload_words()            words the event being executed carries
 *   uint64_t payload(uint32_t k)        word k of them
 */
#ifndef SIMIAN_MODELS_H
#define SIMIAN_MODELS_H

#include "simian_spec.h"

/* built-in device function ids (what attachService binds a service name to) */
enum simian_builtin_fn {
  SIMIAN_FN_NONE = 0,
  SIMIAN_FN_PHOLD_GENERATE = 1, /* "phold_generate" */
  SIMIAN_FN_LANL_SEND = 2,      /* "lanl_send"      */
  SIMIAN_FN_LANL_RECEIVE = 3,   /* "lanl_receive"   */
  SIMIAN_FN_LANL_FINISH = 4,    /* "lanl_finish"    */
  SIMIAN_FN_COUNT_ONLY = 5,     /* "noop"           */
  SIMIAN_FN_LANL_CONSTRUCT = 6, /* "lanl_construct": entity constructor, not a service */
  SIMIAN_FN_HELLO_GENERATE = 7, /* "hello_generate" */
  SIMIAN_FN_HELLO_SQUARE = 8,   /* "hello_square"   */
  SIMIAN_FN_HELLO_SQRT = 9,     /* "hello_sqrt"     */
  SIMIAN_FN_HELLO_RESULT = 10,  /* "hello_result"   */
  SIMIAN_FN_BLOB_SEND = 11,     /* "blob_send"      */
  SIMIAN_FN_BLOB_RECV = 12,     /* "blob_recv"      */
  SIMIAN_FN_BLOB_ACK = 13,      /* "blob_ack"       */
  SIMIAN_FN_LANL_OUTPUT = 14,   /* "lanl_output"    */
  SIMIAN_FN__END,
  SIMIAN_FN_USER = 32 /* first id of handlers a model author registers (simian_handlers.def) */
};

typedef struct simian_no_params { uint32_t unused_; } simian_no_params_t;

/* A handler is PURE when it neither touches state() nor sends to itself with
 * rx omitted below the window bound: its effect depends only on (self id,
 * commit index, now, handler, src, data).  The engine then runs it with one
 * thread per EVENT instead of one thread per entity.  (Table: simian_handlers.def;
 * the function itself is defined at the end of this header.)                    */
SIMIAN_HD int simian_fn_is_pure(int fn);

/* -------------------------------------------------------------- PHOLD ---- */

/* destination modes */
#define SIMIAN_PHOLD_UNIFORM 0u     /* always uniform over all entities
                                       (phold-noop-noMPI.js:38)              */
#define SIMIAN_PHOLD_ROSS_REMOTE 1u /* w.p. p_remote uniform over all, else
                                       self (ROSS phold.c percent_remote)    */
#define SIMIAN_PHOLD_OTHER_GROUP 2u /* w.p. p_remote uniform over entities
                                       whose (num % groups) differs, else
                                       self (BASELINE config 3)              */

typedef struct simian_phold_params {
  uint32_t first_id;   /* global id of Node(0)                               */
  uint32_t n_entities; /* `count`                   phold-noop-noMPI.js:24   */
  double p_remote;
  double lambda;    /* exponential(lambda)          phold-noop-noMPI.js:28,39 */
  double lookahead; /* = minDelay                   phold-noop-noMPI.js:25   */
  uint32_t mode;
  uint32_t groups; /* model-level group count for OTHER_GROUP               */
  uint32_t pad_;
} simian_phold_params_t; /* 48 bytes */

/* Node.generate  (SimianJS/Examples/phold-noop-noMPI.js:36-42):
 *   targetId = floor(range(count)); offset = exponential(1) + lookahead;
 *   self.reqService(offset, "generate", null, "Node", targetId)
 * draw 0 = remote coin, draw 1 = destination, draw 2 = delay.               */
template <class Ctx>
SIMIAN_HD void simian_phold_generate(Ctx &ctx) {
  const simian_phold_params_t *p = (const simian_phold_params_t *)ctx.params();
  uint32_t me = ctx.self_id() - p->first_id;
  uint32_t target = me;
  if (p->mode == SIMIAN_PHOLD_UNIFORM) {
    target = (uint32_t)(simian_u01(ctx.draw(1)) * (double)p->n_entities);
  } else if (simian_u01(ctx.draw(0)) < p->p_remote) {
    if (p->mode == SIMIAN_PHOLD_ROSS_REMOTE) {
      target = (uint32_t)(simian_u01(ctx.draw(1)) * (double)p->n_entities);
    } else {
      /* uniform over the entities of the other (groups-1) residue classes;
       * n_entities is a multiple of groups */
      uint32_t g = p->groups, per = p->n_entities / g;
      uint64_t k = (uint64_t)(simian_u01(ctx.draw(1)) *
                              (double)((uint64_t)per * (g - 1u)));
      uint32_t hop = 1u + (uint32_t)(k % (g - 1u));
      uint32_t row = (uint32_t)(k / (g - 1u));
      target = row * g + (me % g + hop) % g;
    }
  }
  double offset = simian_exponential(ctx.draw(2), p->lambda) + p->lookahead;
  ctx.req_service(offset, ctx.handler(), 0ull, p->first_id + target);
}

/* "noop": commit only (counts + trace hash), emits nothing. */
template <class Ctx>
SIMIAN_HD void simian_noop(Ctx &) {}

/* ------------------------------------------- LANL PDES benchmark (V8) ---- */
/* SimianJS/Examples/pdes_lanl_benchmarkV8.js:99-343.  Followed: the documented
 * intent (:16-85) and the handler bodies (:259-309) with the shared RNG in
 * place of masala.random (each handler invocation draws draw(0), draw(1), ...
 * of its own committed event; the reference re-seeds the global stream from
 * the event payload at handler entry, :260,:291 -- same idea).  Deviations,
 * all stated in DESIGN.md: the geometric source / list-size distributions
 * (p_send, invert, p_list: :209-221) use simian_powi, a defined sequence of
 * multiplications, instead of Math.pow (r_min, :196, is passed in as a
 * parameter); every entity's list has room for the largest one (entity 0's), an
 * entity uses its own list_size of it; an entity whose `ops` is 0 (deep in the
 * geometric tail) does no list work on a receive (JS: NaN loop bounds; the
 * Python twin divides by zero there); the weighted sum `value` is
 * stored in the entity state instead of being discarded (:308) so the compute
 * loop is observable; the statistics fan-in to entity 0 (:311-342, an N-way
 * distinguishable time tie) is by default done off-path by reading
 * entity state (FinishHandler only commits); with stats_on_path FinishHandler
 * sends the reference's message -- 7 + time_bins words, a payload (simian_spec.h)
 * -- to OutputHandler of entity 0, which accumulates the global statistics in
 * its state (the on-disk table is still written by the host from that state).
 * All n_ent messages are due at entity 0 in ONE window: models of a few dozen
 * entities (an entity's due records must fit a CTA's slots).                   */

#define SIMIAN_LANL_TIME_BINS 16
#define SIMIAN_CTOR_COMMIT 0xFFFFFFFFFFFFFFFFull /* commit index of a constructor call */

typedef struct simian_lanl_params {
  uint32_t first_id; /* global id of PDES_LANL_Node(0) */
  uint32_t n_ent;    /* :99  */
  double s_ent, q_avg, p_receive;       /* :100-102 */
  double m_ent, ops_ent, ops_sigma;     /* :106,108,109 */
  double cache_friendliness;            /* :110 */
  double end_time;  /* the MODEL's endTime (:115); engine end = max(end,2)+3*minDelay (:349) */
  double min_delay; /* :116 */
  double r_min;     /* pow(1-p_receive, n_ent)  (:196), computed by the host */
  double p_send, p_list; /* :103,107: 0 = uniform */
  uint32_t time_bins; /* :112, <= SIMIAN_LANL_TIME_BINS */
  uint16_t svc_send, svc_receive; /* interned ids of "SendHandler" / "ReceiveHandler" */
  uint16_t svc_finish;
  uint16_t svc_output;    /* "OutputHandler" (stats_on_path) */
  uint16_t invert;        /* :104 */
  uint16_t stats_on_path; /* FinishHandler sends the statistics message to entity 0 (:311-315) */
} simian_lanl_params_t; /* 120 bytes */

/* per-entity state (:202-252); the list of list_size doubles follows it */
typedef struct simian_lanl_state {
  double local_intersend_delay, q_target, last_scheduled; /* :215-217,228,230 */
  double ops, ops_max, ops_min, ops_mean, value_sink;     /* :222,236-238,308 */
  int64_t q_size;                                         /* :229 */
  uint64_t send_count, receive_count, ops_count;          /* :233-235 */
  uint32_t list_size, active_elements;                    /* :221,223 */
  uint32_t time_sends[SIMIAN_LANL_TIME_BINS];             /* :239-240 */
  /* the global statistics entity 0 collects in OutputHandler (:241-251, :321-330);
   * present in every entity's state, used by entity 0 only (stats_on_path) */
  uint64_t stats_received, gsend_count, greceive_count, gops_count;
  double gops_max, gops_min, gops_mean;
  uint32_t gtime_sends[SIMIAN_LANL_TIME_BINS];
} simian_lanl_state_t; /* 288 bytes */

SIMIAN_HD double simian_floor(double x) { /* Math.floor for |x| < 2^52 */
  double t = (double)(long long)x;
  return (t > x) ? t - 1.0 : t;
}
SIMIAN_HD double simian_ceil(double x) { return -simian_floor(-x); }

#if defined(__CUDA_ARCH__)
#define SIMIAN_SQRT(x) __dsqrt_rn(x)
#else
#define SIMIAN_SQRT(x) __builtin_sqrt(x)
#endif

/* SendHandler body (:259-288); also run by the constructor (:256) */
template <class Ctx>
SIMIAN_HD void simian_lanl_send_body(Ctx &ctx, uint32_t &di) {
  const simian_lanl_params_t *p = (const simian_lanl_params_t *)ctx.params();
  simian_lanl_state_t *st = (simian_lanl_state_t *)ctx.state();
  const double now = ctx.now(), endTime = p->end_time;
  st->send_count++;
  st->q_size--;
  uint32_t bin = (uint32_t)simian_floor(now / (endTime + 0.0001) * (double)p->time_bins); /* :263 */
  if (bin < SIMIAN_LANL_TIME_BINS) st->time_sends[bin]++;
  /* reschedule myself until the queue is full or time has run out (:267-274) */
  while (((double)st->q_size < st->q_target) && !(st->last_scheduled > endTime)) {
    double own_delay = simian_exponential(ctx.draw(di++), 1.0 / st->local_intersend_delay);
    st->last_scheduled = st->last_scheduled + own_delay;
    if (st->last_scheduled < endTime) {
      st->q_size++;
      ctx.req_service(st->last_scheduled - now, p->svc_send, ctx.draw(di++), SIMIAN_NULL_ENTITY);
    }
  }
  /* computation request to the destination entity (:276-287) */
  double dest;
  if (p->p_receive == 1.0)
    dest = 0;
  else if (p->p_receive == 0)
    dest = simian_floor((double)p->n_ent * simian_u01(ctx.draw(di++)));
  else {
    double U = (1.0 - p->r_min) * simian_u01(ctx.draw(di++)) + p->r_min;
    if (!(U > 0.0)) U = 2.2250738585072014e-308;
    dest = simian_floor(simian_ceil(simian_log(U) / simian_log(1.0 - p->p_receive))) - 1;
    if (dest < 0) dest = 0;
    if (dest > (double)(p->n_ent - 1)) dest = (double)(p->n_ent - 1);
  }
  uint64_t new_seed = ctx.draw(di++);
  if (now + p->min_delay < endTime)
    ctx.req_service(p->min_delay, p->svc_receive, new_seed, p->first_id + (uint32_t)dest);
}

template <class Ctx>
SIMIAN_HD void simian_lanl_send(Ctx &ctx) {
  uint32_t di = 0;
  simian_lanl_send_body(ctx, di);
}

/* gauss(mu, sigma) (:180-193): Box-Muller polar method */
template <class Ctx>
SIMIAN_HD double simian_lanl_gauss(Ctx &ctx, uint32_t &di, double mu, double sigma) {
  double x1, x2, w;
  do {
    x1 = 2.0 * simian_u01(ctx.draw(di++)) - 1.0;
    x2 = 2.0 * simian_u01(ctx.draw(di++)) - 1.0;
    w = (x1 * x1) + (x2 * x2);
  } while (w >= 1.0 || w == 0.0);
  w = SIMIAN_SQRT((-2.0 * simian_log(w)) / w);
  return ((x1 * w) * sigma) + mu;
}

/* The receive handler's randomly weighted subset sum (:304-307):
 *   value = sum over i < na, j < nj of list[i] * uniform()
 * Term k = i * nj + j uses draw di0 + k.  The ORDER OF THE ADDITIONS is part of the
 * specification (f64 addition does not associate) and is chosen so that 32 lanes
 * can share one sum: 32 partial sums, partial l adding the terms k = l, l + 32,
 * l + 64, ... in increasing k, then the butterfly p[l] += p[l ^ o] for
 * o = 16, 8, 4, 2, 1 (every p[l] ends with the same bits; value = p[0]).       */
template <class Ctx>
SIMIAN_HD double simian_lanl_weighted_sum(Ctx &ctx, const double *list, uint32_t na,
                                          uint32_t nj, uint32_t di0) {
  double p[32], q[32];
  for (int l = 0; l < 32; l++) p[l] = 0.0;
  const uint32_t total = na * nj;
  for (uint32_t k = 0; k < total; k++)
    p[k & 31u] += list[k / nj] * simian_u01(ctx.draw(di0 + k));
  for (int o = 16; o; o >>= 1) {
    for (int l = 0; l < 32; l++) q[l] = p[l] + p[l ^ o];
    for (int l = 0; l < 32; l++) p[l] = q[l];
  }
  return p[0];
}

/* An engine may take the sum over (a warp per sum, after the entity's events of
 * the window): it then returns true and adds the value to *sink itself, later,
 * in commit order.  Default: not taken over.                                     */
template <class Ctx>
SIMIAN_HD bool simian_defer_weighted_sum(Ctx &, const double *, uint32_t, uint32_t,
                                         uint32_t, double *) {
  return false;
}

/* ReceiveHandler (:290-309): randomly weighted subset sum over the list */
template <class Ctx>
SIMIAN_HD void simian_lanl_receive(Ctx &ctx) {
  const simian_lanl_params_t *p = (const simian_lanl_params_t *)ctx.params();
  simian_lanl_state_t *st = (simian_lanl_state_t *)ctx.state();
  const double *list = (const double *)(st + 1);
  uint32_t di = 0;
  double r_ops = simian_floor(simian_lanl_gauss(ctx, di, st->ops, st->ops * p->ops_sigma));
  if (!(r_ops > 1.0)) r_ops = 1.0;                                       /* max(1, ..) */
  double r_active = 1.0, r_ops_per_element = 0.0;
  if (st->ops > 0.0) {
    r_active = simian_floor((double)st->active_elements * (r_ops / st->ops));
    if (r_active > (double)st->list_size) r_active = (double)st->list_size; /* :295 */
    if (!(r_active > 1.0)) r_active = 1.0;                                  /* :296 */
    r_ops_per_element = simian_floor(r_ops / r_active);                     /* :297 */
  } /* ops == 0 (p_list tail): no list work, see the note at the top */
  st->receive_count++;
  if (r_ops > st->ops_max) st->ops_max = r_ops;
  if (r_ops < st->ops_min) st->ops_min = r_ops;
  st->ops_mean = (st->ops_mean * (double)(st->receive_count - 1) + r_ops) / (double)st->receive_count;
  uint32_t na = (uint32_t)r_active, nj = (uint32_t)r_ops_per_element;
  /* :304-307, in the defined order of simian_lanl_weighted_sum */
  if (!simian_defer_weighted_sum(ctx, list, na, nj, di, &st->value_sink))
    st->value_sink += simian_lanl_weighted_sum(ctx, list, na, nj, di);
  st->ops_count += (uint64_t)na * nj;
}

/* FinishHandler (:311-315): msg = [num, send_count, receive_count, ops_min, ops_mean,
 * ops_max, time_sends[0 .. time_bins), opsCount] to OutputHandler of entity 0 */
template <class Ctx>
SIMIAN_HD void simian_lanl_finish(Ctx &ctx) {
  const simian_lanl_params_t *p = (const simian_lanl_params_t *)ctx.params();
  if (!p->stats_on_path) return; /* the fan-in is done off the event path */
  const simian_lanl_state_t *st = (const simian_lanl_state_t *)ctx.state();
  const uint32_t bins = p->time_bins < SIMIAN_LANL_TIME_BINS ? p->time_bins : SIMIAN_LANL_TIME_BINS;
  const uint32_t num = ctx.self_id() - p->first_id;
  ctx.req_service_blob_fn(p->min_delay, p->svc_output, 7u + bins, p->first_id + 0u,
                          [st, num, bins](uint32_t k) -> uint64_t {
                            if (k == 0) return (uint64_t)num;
                            if (k == 1) return st->send_count;
                            if (k == 2) return st->receive_count;
                            if (k == 3) return simian_f64_bits(st->ops_min);
                            if (k == 4) return simian_f64_bits(st->ops_mean);
                            if (k == 5) return simian_f64_bits(st->ops_max);
                            if (k < 6u + bins) return (uint64_t)st->time_sends[k - 6u];
                            return st->ops_count;
                          });
}

/* OutputHandler (:317-342), the accumulation of the global statistics (:321-330);
 * the table itself is written by the host from entity 0's state */
template <class Ctx>
SIMIAN_HD void simian_lanl_output(Ctx &ctx) {
  const simian_lanl_params_t *p = (const simian_lanl_params_t *)ctx.params();
  simian_lanl_state_t *st = (simian_lanl_state_t *)ctx.state();
  const uint32_t bins = p->time_bins < SIMIAN_LANL_TIME_BINS ? p->time_bins : SIMIAN_LANL_TIME_BINS;
  if (ctx.payload_words() != 7u + bins) return; /* (not a statistics message) */
  const uint64_t sends = ctx.payload(1), recvs = ctx.payload(2);
  const double ops_min = simian_bits_f64(ctx.payload(3)), ops_mean = simian_bits_f64(ctx.payload(4)),
               ops_max = simian_bits_f64(ctx.payload(5));
  st->stats_received++;
  double den = (double)(st->greceive_count + recvs);
  if (!(den > 1.0)) den = 1.0; /* max(1, .) */
  st->gops_mean = ((double)recvs * ops_mean + st->gops_mean * (double)st->greceive_count) / den; /* :322 */
  st->gsend_count += sends;
  st->greceive_count += recvs;
  st->gops_count += ctx.payload(6u + bins);
  if (ops_min < st->gops_min) st->gops_min = ops_min;
  if (ops_max > st->gops_max) st->gops_max = ops_max;
  for (uint32_t i = 0; i < bins; i++) st->gtime_sends[i] += (uint32_t)ctx.payload(6u + i);
}

/* `num` of the entity a handler runs for (class-local index) */
template <class Ctx>
SIMIAN_HD uint32_t simian_lanl_num(Ctx &ctx, const simian_lanl_params_t *p) {
  return ctx.self_id() - p->first_id;
}

/* constructor PDES_LANL_Node (:203-257) */
template <class Ctx>
SIMIAN_HD void simian_lanl_construct(Ctx &ctx) {
  const simian_lanl_params_t *p = (const simian_lanl_params_t *)ctx.params();
  simian_lanl_state_t *st = (simian_lanl_state_t *)ctx.state();
  double *list = (double *)(st + 1);
  uint32_t di = 0;
  const uint32_t num = simian_lanl_num(ctx, p);
  double prob = 1.0 / (double)p->n_ent;                                    /* :208 */
  if (p->p_send != 0)                                                      /* :209-212 */
    prob = p->p_send * simian_powi(1.0 - p->p_send, p->invert ? (uint64_t)(p->n_ent - num) : num);
  double target_global_sends = (double)p->n_ent * p->s_ent;                /* :198 */
  double target_sends = simian_floor(prob * target_global_sends);          /* :213 */
  st->local_intersend_delay =
      (target_sends > 0) ? p->end_time / target_sends : 10 * p->end_time;  /* :214-216 */
  prob = 1.0 / (double)p->n_ent;                                           /* :219 */
  if (p->p_list != 0) prob = p->p_list * simian_powi(1.0 - p->p_list, num); /* :220 */
  st->list_size = (uint32_t)simian_floor(prob * (double)p->n_ent * p->m_ent);   /* :221 */
  st->ops = simian_floor(prob * (double)p->n_ent * p->ops_ent);                 /* :222 */
  st->active_elements = (uint32_t)simian_floor(p->cache_friendliness * (double)st->list_size);
  for (uint32_t i = 0; i < st->list_size; i++) list[i] = simian_u01(ctx.draw(di++)); /* :226 */
  st->q_target = (p->q_avg / p->s_ent) * target_sends;                     /* :228 */
  st->q_size = 1;
  st->last_scheduled = ctx.now();
  st->send_count = st->receive_count = st->ops_count = 0;
  st->ops_max = 0;
  st->ops_min = 1e100;
  st->ops_mean = 0.0;
  st->value_sink = 0.0;
  for (uint32_t i = 0; i < SIMIAN_LANL_TIME_BINS; i++) st->time_sends[i] = 0;
  st->stats_received = st->gsend_count = st->greceive_count = st->gops_count = 0; /* :241-251 */
  st->gops_max = 0;
  st->gops_min = 1e100;
  st->gops_mean = 0.0;
  for (uint32_t i = 0; i < SIMIAN_LANL_TIME_BINS; i++) st->gtime_sends[i] = 0;
  /* :255 FinishUp at end of time, :256 first SendHandler call */
  ctx.req_service(p->end_time - ctx.now(), p->svc_finish, 0ull, SIMIAN_NULL_ENTITY);
  simian_lanl_send_body(ctx, di);
}

/* ------------------------------------------------- hello.js (two classes) */
/* SimianJS/Examples/hello.js:27-93: Alice and Bob entities; `generate` sends a
 * number to a random Alice ("square") or Bob ("sqrt") and reschedules itself
 * with rx omitted; the receiver replies "result" to tx/txId.  The payload is
 * the f64 as its 64 bits.  Integer delays: many same-entity time ties.
 * draw 0 = target class, draw 1 = target id, draw 2 = data.                  */
typedef struct simian_hello_params {
  uint32_t alice_first, bob_first; /* global ids of Alice(0) / Bob(0) */
  uint32_t count, pad_;            /* hello.js:25 */
  uint16_t svc_generate, svc_square, svc_sqrt, svc_result;
} simian_hello_params_t; /* 24 bytes */

template <class Ctx>
SIMIAN_HD void simian_hello_generate(Ctx &ctx) { /* hello.js:32-48 / :68-84 */
  const simian_hello_params_t *p = (const simian_hello_params_t *)ctx.params();
  uint32_t target = (uint32_t)(simian_u01(ctx.draw(0)) * 2.0);          /* :36 */
  uint32_t targetId = (uint32_t)(simian_u01(ctx.draw(1)) * (double)p->count); /* :37 */
  double data = simian_u01(ctx.draw(2)) * 100.0;                        /* :39 */
  ctx.req_service(10.0, target ? p->svc_sqrt : p->svc_square, simian_f64_bits(data),
                  (target ? p->bob_first : p->alice_first) + targetId); /* :46 */
  ctx.req_service(25.0, p->svc_generate, 0ull, SIMIAN_NULL_ENTITY);      /* :47 */
}
template <class Ctx>
SIMIAN_HD void simian_hello_square(Ctx &ctx) { /* hello.js:50-52 */
  const simian_hello_params_t *p = (const simian_hello_params_t *)ctx.params();
  double d = simian_bits_f64(ctx.data());
  ctx.req_service(10.0, p->svc_result, simian_f64_bits(d * d), ctx.src());
}
template <class Ctx>
SIMIAN_HD void simian_hello_sqrt(Ctx &ctx) { /* hello.js:86-88 */
  const simian_hello_params_t *p = (const simian_hello_params_t *)ctx.params();
  double d = simian_bits_f64(ctx.data());
  ctx.req_service(10.0, p->svc_result, simian_f64_bits(SIMIAN_SQRT(d)), ctx.src());
}
template <class Ctx>
SIMIAN_HD void simian_hello_result(Ctx &) {} /* hello.js:54-59: logs only */

/* ------------------------------------- payloads beyond 8 bytes ("blobs") ---- */
/* entity.js:52-60: `data` of an event is any JS object.  Here: any number of 64-bit
 * words (<= SIMIAN_BLOB_MAX_WORDS), sent with ctx.req_service_blob and read with
 * ctx.payload_words() / ctx.payload(k) (simian_spec.h: continuation records).  The
 * test model: every entity keeps sending a random-length vector of random words to
 * a random entity ("recv"), which folds the words in order and acknowledges the
 * fold to the sender ("ack": the sender's trace hash then covers every word).
 * draw 0 = target, 1 = length, 2 = delay to the receiver, 3 = own delay, 4.. words. */
#define SIMIAN_BLOB_MODEL_MAX_WORDS 24
typedef struct simian_blob_params {
  uint32_t first_id, n_entities;
  double lookahead;
  uint32_t max_words, pad_;
  uint16_t svc_send, svc_recv, svc_ack, pad2_;
} simian_blob_params_t; /* 32 bytes */

template <class Ctx>
SIMIAN_HD void simian_blob_send(Ctx &ctx) {
  const simian_blob_params_t *p = (const simian_blob_params_t *)ctx.params();
  uint32_t target = (uint32_t)(simian_u01(ctx.draw(0)) * (double)p->n_entities);
  uint32_t mw = p->max_words < SIMIAN_BLOB_MODEL_MAX_WORDS ? p->max_words : SIMIAN_BLOB_MODEL_MAX_WORDS;
  uint32_t nw = (uint32_t)(ctx.draw(1) % (uint64_t)(mw + 1u)); /* 0 .. max_words */
  /* (the words are produced on demand: no staging array in the handler) */
  ctx.req_service_blob_fn(simian_exponential(ctx.draw(2), 1.0) + p->lookahead, p->svc_recv, nw,
                          p->first_id + target, [&ctx](uint32_t k) { return ctx.draw(4 + k); });
  ctx.req_service(simian_exponential(ctx.draw(3), 1.0) + p->lookahead, p->svc_send, 0ull,
                  SIMIAN_NULL_ENTITY);
}
template <class Ctx>
SIMIAN_HD void simian_blob_recv(Ctx &ctx) {
  const simian_blob_params_t *p = (const simian_blob_params_t *)ctx.params();
  uint32_t nw = ctx.payload_words();
  uint64_t fold = simian_mix64((uint64_t)nw + 1ull);
  for (uint32_t k = 0; k < nw; k++) fold = simian_mix64(fold ^ ctx.payload(k));
  ctx.req_service(p->lookahead, p->svc_ack, fold, ctx.src());
}
template <class Ctx>
SIMIAN_HD void simian_blob_ack(Ctx &) {} /* commits: the trace hash covers the fold */

/* ------------------------------------------------ the handler registry ---- */
#ifdef SIMIAN_USER_HANDLERS_H
#include SIMIAN_USER_HANDLERS_H /* a model author's handlers (make HANDLERS=...) */
#endif

SIMIAN_HD int simian_fn_is_pure(int fn) {
  switch (fn) {
#define SIMIAN_HANDLER(id, name, func, pure, ptype) \
  case id:                                          \
    return pure;
#include "simian_handlers.def"
#undef SIMIAN_HANDLER
    default:
      return 0;
  }
}

/* name -> id (0: not registered); bytes of the parameter struct the handler reads */
static inline int simian_fn_by_name(const char *n) {
#define SIMIAN_HANDLER(id, name, func, pure, ptype) \
  if (!strcmp(n, name)) return id;
#include "simian_handlers.def"
#undef SIMIAN_HANDLER
  return SIMIAN_FN_NONE;
}
static inline size_t simian_fn_params_bytes(int fn) {
  switch (fn) {
#define SIMIAN_HANDLER(id, name, func, pure, ptype) \
  case id:                                          \
    return sizeof(ptype);
#include "simian_handlers.def"
#undef SIMIAN_HANDLER
    default:
      return 0;
  }
}

#endif /* SIMIAN_MODELS_H */
