/* tests/user_handlers/bounce.h -- a model author's own handlers, compiled into the
 * engine and the oracle with  make HANDLERS=<repo>/tests/user_handlers/bounce
 * (include/simian_handlers.def explains the mechanism).  The JS equivalent
 * (SimianJS model script):
 *     function bounce(self, data, tx, txId) {        // attachService(Ring, "bounce", bounce)
 *       self.reqService(lookahead + uniform(), "bounce", data + 1, "Ring", (self.num + 1) % n)
 *       if (data % 7 == 0) self.reqService(lookahead * 3, "tick", data, "Ring", self.num)
 *     }
 *     function tick(self, data, tx, txId) {}          // counts only                 */
#ifndef USER_BOUNCE_H
#define USER_BOUNCE_H

typedef struct user_bounce_params {
  uint32_t first_id, n; /* global id of Ring(0), number of Ring entities */
  double lookahead;
  uint16_t svc_bounce, svc_tick;
  uint32_t pad_;
} user_bounce_params_t; /* 24 bytes */

template <class Ctx>
SIMIAN_HD void user_bounce(Ctx &ctx) {
  const user_bounce_params_t *p = (const user_bounce_params_t *)ctx.params();
  uint32_t me = ctx.self_id() - p->first_id;
  double off = p->lookahead + simian_u01(ctx.draw(0));
  ctx.req_service(off, p->svc_bounce, ctx.data() + 1ull, p->first_id + (me + 1u) % p->n);
  if (ctx.data() % 7ull == 0ull)
    ctx.req_service(p->lookahead * 3.0, p->svc_tick, ctx.data(), p->first_id + me);
}

template <class Ctx>
SIMIAN_HD void user_tick(Ctx &) {}

#endif
