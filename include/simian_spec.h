/*
 * simian_spec.h -- the ONE shared specification header of simian-b200.
 *
 * Everything that decides a result bit lives here and is compiled, unchanged,
 * into (a) the CUDA engine (nvcc, -fmad=false), (b) the CPU oracle
 * (g++, -ffp-contract=off) and is transcribed 1:1 into the pure-Python model used
 * to pin the oracle against the unmodified SimianPie engine
 * (oracle/pin_simianpie.py).
 *
 *   - the fixed-size POD event record that replaces the reference's JS event
 *     object  {tx, txId, rx, rxId, name, data, time}
 *     (reference: SimianJS/entity.js:52-60, SimianJS/simian.js:178-186);
 *   - the counter-based RNG that replaces the single global libc rand() stream
 *     (reference: SimianJS/MasalaChai/random.cpp:19-63);
 *   - a portable natural log built from IEEE +,-,*,/ and bit operations only, so
 *     that exponential delays are bit-identical on host and device
 *     (reference call site: SimianJS/Examples/phold-noop-noMPI.js:28);
 *   - the per-entity order-sensitive trace hash and the digests built from it;
 *   - the engine's defined tie-break key (time, handler, src, seq)
 *     (the reference heap compares .time only: SimianJS/eventQ.js:36,62,66).
 *
 * Plain C99 / C++ / CUDA. No FMA contraction may be applied to this file.
 */
#ifndef SIMIAN_SPEC_H
#define SIMIAN_SPEC_H

#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define SIMIAN_HD __host__ __device__ __forceinline__
#else
#define SIMIAN_HD static inline
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ events */

#define SIMIAN_NULL_ENTITY 0xFFFFFFFFu /* tx/txId == null (simian.js:179-180) */

/* flags */
#define SIMIAN_EVF_SCHED 0x0001u /* injected by schedService (tx = null)      */
/* Payloads larger than the 8-byte `data` word (entity.js:52-60: `data` is any
 * object).  The event itself is the HEADER: flags & BLOB, data = number of 64-bit
 * payload words.  Word k travels as a CONTINUATION record -- same time, dst, src,
 * handler and seq as its header, flags = CONT | k, data = the word -- through the
 * calendar, the outboxes and the exchange like any record, but is never committed:
 * it is what the header's handler reads with payload(k).  (src, seq) names the
 * header: seq = simian_emit_seq(commit index of the sender, emission number).   */
#define SIMIAN_EVF_BLOB 0x0002u
#define SIMIAN_EVF_CONT 0x8000u          /* | word index, 15 bits */
#define SIMIAN_BLOB_MAX_WORDS 0x7FFFu

/* 32 bytes = exactly one DRAM sector; a scattered append costs one sector.  */
typedef struct
#if defined(__CUDACC__)
    __align__(32)
#elif defined(__GNUC__)
    __attribute__((aligned(32)))
#endif
        simian_event {
  double time;      /* absolute simulated time            (entity.js:47,59)  */
  uint32_t dst;     /* receiver: global entity id         (rx, rxId)         */
  uint32_t src;     /* sender: global entity id or NULL   (tx, txId)         */
  uint16_t handler; /* interned service name              (name)             */
  uint16_t flags;
  uint32_t seq;     /* per-source emission counter: tie-break only           */
  uint64_t data;    /* 8-byte payload, or the word count of a longer one (data) */
} simian_event_t;

/* ------------------------------------------------------------ bit helpers */

SIMIAN_HD uint64_t simian_f64_bits(double x) {
  uint64_t u;
#if defined(__CUDA_ARCH__)
  u = (uint64_t)__double_as_longlong(x);
#else
  memcpy(&u, &x, 8);
#endif
  return u;
}

SIMIAN_HD double simian_bits_f64(uint64_t u) {
  double x;
#if defined(__CUDA_ARCH__)
  x = __longlong_as_double((long long)u);
#else
  memcpy(&x, &u, 8);
#endif
  return x;
}

/* Monotone map f64 -> u64: a < b  <=>  key(a) < key(b) (no NaNs).           */
SIMIAN_HD uint64_t simian_time_key(double t) {
  uint64_t u = simian_f64_bits(t);
  return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}

SIMIAN_HD double simian_key_time(uint64_t k) {
  uint64_t u = (k & 0x8000000000000000ull) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
  return simian_bits_f64(u);
}

/* ------------------------------------------------------------------- RNG  */

#define SIMIAN_GOLD 0x9E3779B97F4A7C15ull
#define SIMIAN_C1 0xD1342543DE82EF95ull
#define SIMIAN_C2 0xA0761D6478BD642Full

/* splitmix64 finaliser */
SIMIAN_HD uint64_t simian_mix64(uint64_t x) {
  x ^= x >> 30;
  x *= 0xBF58476D1CE4E5B9ull;
  x ^= x >> 27;
  x *= 0x94D049BB133111EBull;
  x ^= x >> 31;
  return x;
}

/* Stream key of one committed event: depends only on (seed, entity, the
 * entity's commit index) -- never on global execution order, so the draw
 * sequence is identical under any placement / parallel schedule.            */
SIMIAN_HD uint64_t simian_rng_key(uint64_t seed, uint32_t entity,
                                  uint64_t commit_idx) {
  uint64_t k = simian_mix64(seed ^ (((uint64_t)entity + 1ull) * SIMIAN_C1));
  return simian_mix64(k + commit_idx * SIMIAN_C2);
}

/* draw_idx-th 64-bit draw of that event */
SIMIAN_HD uint64_t simian_rng_draw(uint64_t key, uint32_t draw_idx) {
  return simian_mix64(key + ((uint64_t)draw_idx + 1ull) * SIMIAN_GOLD);
}

/* u in [0, 1): 53 bits.  floor(u*n) < n for every integer n < 2^53.         */
SIMIAN_HD double simian_u01(uint64_t x) {
  return (double)(x >> 11) * 1.1102230246251565e-16; /* 2^-53 */
}

/* u in (0, 1): never 0, never 1 (contrast random.cpp:27-29, closed [0,1]).  */
SIMIAN_HD double simian_u01_open(uint64_t x) {
  return ((double)(x >> 12) + 0.5) * 2.220446049250313e-16; /* 2^-52 */
}

/* -------------------------------------------------------- portable log()  */

/* Natural log for finite normal x > 0, from IEEE basic operations and integer
 * bit manipulation only (argument reduction x = 2^k * m, m in [sqrt2/2,sqrt2),
 * then log(1+f) = 2s + s*R(s^2), s = f/(2+f), with the classical degree-14
 * minimax coefficients).  Every operation is a separately rounded IEEE op:
 * this file must be compiled without FMA contraction.                        */
SIMIAN_HD double simian_log(double x) {
  const double ln2_hi = 6.93147180369123816490e-01;
  const double ln2_lo = 1.90821492927058770002e-10;
  const double L1 = 6.666666666666735130e-01, L2 = 3.999999999940941908e-01,
               L3 = 2.857142874366239149e-01, L4 = 2.222219843214978396e-01,
               L5 = 1.818357216161805012e-01, L6 = 1.531383769920937332e-01,
               L7 = 1.479819860511658591e-01;
  uint64_t ix = simian_f64_bits(x);
  int k = (int)(ix >> 52) - 1023;
  uint64_t man = ix & 0x000FFFFFFFFFFFFFull;
  /* m in [1,2); fold the upper part down so that m in [sqrt2/2, sqrt2) */
  if (man >= 0x6A09E667F3BCDull) { /* m >= sqrt(2) (to 52 bits) */
    k += 1;
    ix = man | 0x3FE0000000000000ull; /* m/2 in [sqrt2/2, 1) */
  } else {
    ix = man | 0x3FF0000000000000ull;
  }
  double m = simian_bits_f64(ix);
  double f = m - 1.0;
  double dk = (double)k;
  double s = f / (2.0 + f);
  double z = s * s;
  double w = z * z;
  double t1 = w * (L2 + w * (L4 + w * L6));
  double t2 = z * (L1 + w * (L3 + w * (L5 + w * L7)));
  double R = t2 + t1;
  double hfsq = (0.5 * f) * f;
  return dk * ln2_hi - ((hfsq - (s * (hfsq + R) + dk * ln2_lo)) - f);
}

/* exponential(lambda) = -ln(uniform())/lambda   (phold-noop-noMPI.js:28)    */
/* b^n for an integer exponent, as a FIXED sequence of f64 multiplications (square
 * and multiply, low bit first): Math.pow / ** of the reference scripts
 * (pdes_lanl_benchmarkV8.js:211-212,220) is libm and not reproducible bit for bit
 * across platforms; this is -- on host, device and in oracle/pyspec.py.          */
SIMIAN_HD double simian_powi(double b, uint64_t n) {
  double r = 1.0;
  while (n) {
    if (n & 1ull) r = r * b;
    b = b * b;
    n >>= 1;
  }
  return r;
}

SIMIAN_HD double simian_exponential(uint64_t draw, double lambda) {
  double e = -simian_log(simian_u01_open(draw));
  return (lambda == 1.0) ? e : e / lambda; /* x/1.0 == x: same bits */
}

/* ------------------------------------------------------------ trace hash  */

/* Per-entity order-sensitive chain over the events the entity commits, in
 * commit order.  Covers what the reference handler sees: (time, service
 * name, tx/txId, data) -- simian.js:126-131.  `seq` is NOT hashed.          */
/* per-event part (order independent, computable in parallel) */
SIMIAN_HD uint64_t simian_trace_event(double time, uint32_t handler, uint32_t src,
                                      uint64_t data) {
  uint64_t x = simian_mix64(simian_f64_bits(time) +
                            SIMIAN_GOLD * ((((uint64_t)handler) << 32) | src));
  return simian_mix64(x ^ data);
}
/* chain part (sequential per entity, a few integer operations) */
SIMIAN_HD uint64_t simian_trace_chain(uint64_t h, uint64_t x) {
  h = (h ^ x) * SIMIAN_GOLD;
  return h ^ (h >> 29);
}
SIMIAN_HD uint64_t simian_trace_step(uint64_t h, double time, uint32_t handler,
                                     uint32_t src, uint64_t data) {
  return simian_trace_chain(h, simian_trace_event(time, handler, src, data));
}

/* order-insensitive across entities, order-sensitive within each entity     */
SIMIAN_HD uint64_t simian_digest_term(uint32_t entity, uint64_t count,
                                      uint64_t hash) {
  return simian_mix64(hash ^ simian_mix64(((uint64_t)entity << 1) | 1ull)) +
         simian_mix64(count + SIMIAN_C1 * ((uint64_t)entity + 1ull));
}

/* fold one 8-byte word of user entity state into the state checksum         */
SIMIAN_HD uint64_t simian_state_fold(uint64_t acc, uint64_t word) {
  return simian_mix64(acc ^ word) + SIMIAN_C2;
}

/* ------------------------------------------------------ defined tie-break */

/* Engine order among events of ONE destination entity:
 * (time, handler, src, seq) ascending.  Equals the SimianJS pop order
 * whenever the oracle reports zero distinguishable same-entity time ties.   */
SIMIAN_HD int simian_event_before(double ta, uint32_t ha, uint32_t sa,
                                  uint32_t qa, double tb, uint32_t hb,
                                  uint32_t sb, uint32_t qb) {
  if (ta != tb) return ta < tb;
  if (ha != hb) return ha < hb;
  if (sa != sb) return sa < sb;
  return qa < qb;
}

/* seq of the k-th emission made while committing the entity's commit_idx-th
 * event: stateless per-source monotone counter.                             */
SIMIAN_HD uint32_t simian_emit_seq(uint64_t commit_idx, uint32_t k) {
  return (uint32_t)(commit_idx * 16ull + (uint64_t)k);
}

/* ------------------------------------------------ window arithmetic (A5)  */

/* infTime = endTime + 2*minDelay                         (simian.js:64)     */
SIMIAN_HD double simian_inf_time(double end_time, double min_delay) {
  return end_time + 2.0 * min_delay;
}

#ifdef __cplusplus
}
#endif
#endif /* SIMIAN_SPEC_H */
