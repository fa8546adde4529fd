/* tests/user_handlers/bounce.h -- a model author's own handlers, compiled into the
 * engine and the oracle with  make HANDLERS=<repo>/tests/user_handlers/bounce
 * (include/simian_handlers.def explains the mechanism).  The JS equivalent
 * (SimianJS model script):
 *     function bounce(self, data, tx, txId) {        // attachService(Ring, "bounce", bounce)
 *       self.reqService(lookahead + uniform(), "bounce", data + 1, "Ring", (self.num + 1) % n)
 *       if (data % 7 == 0) self.reqService(lookahead * 3, "tick", data, "Ring", self.num)
 *     }
 *     function tick(self, data, tx, txId) {}          // counts only
 * plus, every fifth bounce, a three-word payload to "tick" of the entity two places on,
 * which folds it and bounces the fold back: payloads on the thread-per-event path.     */
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
  /* a payload longer than 8 bytes from a PURE handler (one thread per event): three
   * words to the entity two places on; "tick" folds them and bounces the fold back */
  if (ctx.data() % 5ull == 0ull) {
    const uint64_t d = ctx.data();
    ctx.req_service_blob_fn(p->lookahead * 2.0, p->svc_tick, 3u, p->first_id + (me + 2u) % p->n,
                            [&ctx, d](uint32_t k) -> uint64_t { return ctx.draw(1 + k) ^ d; });
  }
}

template <class Ctx>
SIMIAN_HD void user_tick(Ctx &ctx) {
  const user_bounce_params_t *p = (const user_bounce_params_t *)ctx.params();
  const uint32_t nw = ctx.payload_words();
  if (nw == 0) return; /* the plain tick: counts only */
  uint64_t fold = 0;
  for (uint32_t k = 0; k < nw; k++) fold = simian_mix64(fold ^ ctx.payload(k));
  ctx.req_service(p->lookahead, p->svc_bounce, (fold & 0xFFFull) * 5ull + 1ull, ctx.src());
}

#endif
