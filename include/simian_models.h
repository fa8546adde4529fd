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
 *   void req_service(double offset, uint16_t service, uint64_t data,
 *                    uint32_t rx)       self.reqService(...)  entity.js:35-68
 *        rx == SIMIAN_NULL_ENTITY  <=>  rx omitted (send to self, no lookahead
 *        check, entity.js:41,50-51)
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
  SIMIAN_FN__END
};

/* A handler is PURE when it neither touches state() nor sends to itself with
 * rx omitted below the window bound: its effect depends only on (self id,
 * commit index, now, handler, src, data).  The engine then runs it with one
 * thread per EVENT instead of one thread per entity.                         */
SIMIAN_HD int simian_fn_is_pure(int fn) {
  return fn == SIMIAN_FN_PHOLD_GENERATE || fn == SIMIAN_FN_COUNT_ONLY;
}

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

#endif /* SIMIAN_MODELS_H */
