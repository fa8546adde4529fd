// kernels.cuh -- hand-written sm_100a kernels of the simian-b200 window loop.
//
// One lookahead window of the reference (SimianJS/simian.js:119-148) is ONE
// launch of k_window.  An entity GROUP (64 consecutive local entities) is owned
// by ONE WARP; the warps of a CTA never wait for each other between the first
// and the last instruction of a window (no block barrier on the hot path):
//   K2  dequeue: the pages of the group's calendar cells in the window's rows
//       are bulk-copied into the warp's shared-memory slots (cp.async.bulk +
//       mbarrier, one copy per page)                 (eventQ.pop, eventQ.js:53-92)
//   K3  the due records are counting-sorted by entity inside the warp
//       (histogram + scan -> contiguous per-entity segments), ranked under
//       (time, handler, src, seq) and their handlers run, one lane per event
//                                                     (simian.js:124-133)
//   K4  emit: lock-free paged append of 32-B records into the destination
//       cell, or into the per-peer outbox             (entity.js:47-67, eventQ.push)
//   K1  LBTS: warp min-reduction of leftover + emitted times, one result-less
//       atomic per CTA, last CTA computes the next window (simian.js:120,143-147)
// Windows whose records overflow the warp's slots (bursts) are processed in
// several passes over entity sub-ranges (slow_passes).
// Compiled with -fmad=false: include/simian_spec.h must see no FMA contraction.
#pragma once
#include "device_types.cuh"

#define GE SIMIAN_GE
#define GE_SHIFT SIMIAN_GE_SHIFT
#define PAGE SIMIAN_PAGE
#define NSLOT SIMIAN_NSLOT
#define PGMAX SIMIAN_PGMAX
#define NROUND ((NSLOT + 31) / 32)
#define FULL 0xffffffffu
static_assert(NSLOT <= 255, "slot indices and arrival counts are kept in bytes");
static_assert(NROUND <= 8, "per-lane ranks / slots are packed 8 bits each into 64 bits");
static_assert(PAGE <= 31 && (PAGE & (PAGE - 1)) == 0, "page record count: 5 bits of pg_cnt");
static_assert(SIMIAN_MAX_CELLS <= 96, "ring span");

// page-list entry flags (pg_cnt)
#define PG_COUNT(x) ((uint32_t)(x) & 31u)
#define PG_LINKED 0x4000u // the page's page_next entry is set (reset it when the page is freed)
#define PG_FREE_FLAG 0x8000u

// per-phase SM-cycle accounting of k_window (diagnostic builds only)
#ifdef SIMIAN_PROFILE_PHASES
#define PHASE(k)                                                                 \
  do {                                                                           \
    __syncwarp();                                                                \
    if ((threadIdx.x & 31) == 0) {                                               \
      long long _t = clock64();                                                  \
      atomicAdd(&c.acc->phase_cycles[k], (unsigned long long)(_t - ws.t_last));  \
      ws.t_last = _t;                                                            \
    }                                                                            \
  } while (0)
#else
#define PHASE(k) do { } while (0)
#endif

// ---------------------------------------------------------------- helpers --

struct __align__(16) RecBits { // the 32-byte record as four 64-bit words
  unsigned long long q0, q1, q2, q3;
  __device__ __forceinline__ double time() const { return __longlong_as_double((long long)q0); }
  __device__ __forceinline__ uint32_t dst() const { return (uint32_t)q1; }
  __device__ __forceinline__ uint32_t src() const { return (uint32_t)(q1 >> 32); }
  __device__ __forceinline__ uint32_t handler() const { return (uint32_t)(q2 & 0xFFFFu); }
  __device__ __forceinline__ uint32_t flags() const { return (uint32_t)((q2 >> 16) & 0xFFFFu); }
  __device__ __forceinline__ uint32_t seq() const { return (uint32_t)(q2 >> 32); }
  __device__ __forceinline__ unsigned long long data() const { return q3; }
};

__device__ __forceinline__ RecBits make_rec(double time, uint32_t dst, uint32_t src,
                                            uint32_t handler, uint32_t flags,
                                            uint32_t seq, unsigned long long data) {
  RecBits r;
  r.q0 = (unsigned long long)__double_as_longlong(time);
  r.q1 = (unsigned long long)dst | ((unsigned long long)src << 32);
  r.q2 = (unsigned long long)(handler & 0xFFFFu) | ((unsigned long long)(flags & 0xFFFFu) << 16) |
         ((unsigned long long)seq << 32);
  r.q3 = data;
  return r;
}

// one 256-bit transaction per record: a scattered append is exactly one sector
__device__ __forceinline__ void st_rec(simian_event_t *p, const RecBits &r) {
  asm volatile("st.global.v4.u64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(r.q0), "l"(r.q1),
               "l"(r.q2), "l"(r.q3)
               : "memory");
}
// Same record as two 128-bit stores.  Used inside __noinline__ functions: ptxas
// 12.9 turns the 256-bit form into a single 64-bit store there (seen in SASS:
// LDL.64 + STG.E.64), silently dropping 24 of the 32 bytes.
__device__ __forceinline__ void st_rec_2x16(simian_event_t *p, unsigned long long q0,
                                            unsigned long long q1, unsigned long long q2,
                                            unsigned long long q3) {
  asm volatile("st.global.v2.u64 [%0], {%1,%2};" ::"l"(p), "l"(q0), "l"(q1) : "memory");
  asm volatile("st.global.v2.u64 [%0], {%1,%2};" ::"l"((char *)p + 16), "l"(q2), "l"(q3)
               : "memory");
}
__device__ __forceinline__ RecBits ld_rec(const simian_event_t *p) {
  RecBits r;
  asm volatile("ld.global.v4.u64 {%0,%1,%2,%3}, [%4];"
               : "=l"(r.q0), "=l"(r.q1), "=l"(r.q2), "=l"(r.q3)
               : "l"(p));
  return r;
}
// the record as two 128-bit loads (the __noinline__ twin of st_rec_2x16)
__device__ __forceinline__ RecBits ld_rec_2x16(const simian_event_t *p) {
  RecBits r;
  asm volatile("ld.global.v2.u64 {%0,%1}, [%2];" : "=l"(r.q0), "=l"(r.q1) : "l"(p));
  asm volatile("ld.global.v2.u64 {%0,%1}, [%2];" : "=l"(r.q2), "=l"(r.q3) : "l"((const char *)p + 16));
  return r;
}
// first half only: {time, dst|src}
__device__ __forceinline__ void ld_rec_head(const simian_event_t *p, unsigned long long &q0,
                                            unsigned long long &q1) {
  asm volatile("ld.global.v2.u64 {%0,%1}, [%2];" : "=l"(q0), "=l"(q1) : "l"(p));
}
// L2-coherent loads (bypass L1) for data other CTAs wrote in this launch
__device__ __forceinline__ uint32_t ld_cg(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ unsigned long long ld_cg64(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}

// ---- bulk asynchronous copies (TMA unit, no tensor map) + mbarrier ---------
// global -> shared, 16-byte granules, completion counted in bytes on an
// mbarrier in shared memory (SASS: UBLKCP + SYNCS).

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t arrivals) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals)
               : "memory");
  // make the initialised barrier visible to the async proxy (the copy engine)
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// one arrival that also announces `bytes` of bulk-copy traffic still to land
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gsrc, uint32_t bytes,
                                         unsigned long long *bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(smem_dst)),
      "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

__device__ __forceinline__ long long tb_of(const DevCfg &c, double t) {
  double x = (t - c.start) * c.inv_w;
  if (!(x < 4.0e18)) return 4000000000000000000LL; // +inf / NaN / huge
  if (x < -4.0e18) return -4000000000000000000LL;
  return __double2ll_rd(x);
}

__device__ __forceinline__ void set_error(const DevCfg &c, unsigned int code) {
  atomicCAS(&c.acc->error, 0u, code);
}

// entity id <-> (rank, slot) under the reference placement
//   getOffsetRank(name, num) = (baseRanks[name] + num) % size   simian.js:196-198
__device__ __forceinline__ int class_of_gid(const DevCfg &c, uint32_t gid) {
  int k = 0;
#pragma unroll 1
  for (int i = 1; i < c.n_classes; i++)
    if (gid >= c.cls[i].first_id) k = i;
  return k;
}
__device__ __forceinline__ void route_gid(const DevCfg &c, uint32_t gid, int &rank,
                                          uint32_t &slot) {
  int k = class_of_gid(c, gid);
  const DevClass &cl = c.cls[k];
  uint32_t num = gid - cl.first_id;
  if (c.size == 1) {
    rank = 0;
    slot = cl.slot_base + num;
  } else {
    uint32_t g = (uint32_t)c.size;
    rank = (int)((cl.base_rank + num) % g);
    // slot on the OWNING rank: classes are laid out in the same order there
    // and class k's slot_base on rank r is sum over earlier classes of their
    // local counts on r -- recomputed here because it differs per rank.
    uint32_t sb = 0;
    for (int i = 0; i < k; i++) {
      const DevClass &ci = c.cls[i];
      uint32_t ro = ((uint32_t)rank + g - ci.base_rank % g) % g;
      sb += (ci.count > ro) ? (ci.count - ro + g - 1) / g : 0;
    }
    slot = sb + num / g;
  }
}
__device__ __forceinline__ uint32_t gid_of_slot(const DevCfg &c, uint32_t slot, int &k) {
  k = 0;
#pragma unroll 1
  for (int i = 1; i < c.n_classes; i++)
    if (slot >= c.cls[i].slot_base) k = i;
  const DevClass &cl = c.cls[k];
  uint32_t j = slot - cl.slot_base;
  return cl.first_id + (c.size == 1 ? j : j * (uint32_t)c.size + cl.rank_off);
}

// 64-bit min / max over the warp with redux.sync on the two halves
__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v) {
  uint32_t hi = (uint32_t)(v >> 32), lo = (uint32_t)v;
  uint32_t mh = __reduce_min_sync(FULL, hi);
  uint32_t ml = __reduce_min_sync(FULL, hi == mh ? lo : 0xFFFFFFFFu);
  return ((unsigned long long)mh << 32) | ml;
}
__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
  uint32_t hi = (uint32_t)(v >> 32), lo = (uint32_t)v;
  uint32_t mh = __reduce_max_sync(FULL, hi);
  uint32_t ml = __reduce_max_sync(FULL, hi == mh ? lo : 0u);
  return ((unsigned long long)mh << 32) | ml;
}

// ------------------------------------------------------ paged cell append --

// In shared memory: one per WARP in k_window (the group's page cache, bulk
// copied from bcache[g]), one per CTA in the other appending kernels (no cache).
struct __align__(16) AppendCtx {
  // [0, SIMIAN_BC): the group's free pages, a stack popped from the top with a
  // shared-memory atomic; [SIMIAN_BC]: the stack height as stored in HBM
  uint32_t pgc[SIMIAN_BC_ROW];
  uint32_t far_appends, page_allocs;
  uint32_t pgc_take, pgc_have;
};

// head word of an empty cell: no page, "full" so the first append installs one
#define CELL_EMPTY (((unsigned long long)SIMIAN_NONE << 32) | (unsigned long long)PAGE)
// Bit 31 of the head word's count half (and of a page_next entry): the page has
// an OLDER page behind it.
#define CELL_OLDER 0x80000000u
#define CELL_COUNT(lo32) ((uint32_t)(lo32) & 0x7FFFFFFFu)

__device__ __forceinline__ void append_ctx_init(AppendCtx &ax) {
  if (threadIdx.x == 0) ax.far_appends = ax.page_allocs = ax.pgc_take = ax.pgc_have = 0;
}

// Free-page invariant for a page whose contents are unknown (never used since
// the pool was allocated, or a far-row page whose records lie in the future).
__device__ __forceinline__ void scrub_page(const DevCfg &c, uint32_t p) {
  unsigned long long *w = reinterpret_cast<unsigned long long *>(c.pool + (size_t)p * PAGE);
#pragma unroll 4
  for (uint32_t s = 0; s < (uint32_t)PAGE; s++) w[4 * s] = SIMIAN_TIME_DEAD;
}

// Pop a page from the free ring.  Pages pushed during a launch become
// allocatable only after the next advance step raises alloc_limit, so a launch
// never pops a ring slot that is being written.
__device__ __forceinline__ uint32_t page_get(const DevCfg &c, AppendCtx &ax,
                                             unsigned long long alloc_limit) {
  atomicAdd(&ax.page_allocs, 1u);
  if (ax.pgc_have) { // group cache first: a shared-memory atomic, no L2 round trip
    uint32_t k = atomicAdd(&ax.pgc_take, 1u);
    if (k < ax.pgc_have) return ax.pgc[ax.pgc_have - 1u - k];
  }
  unsigned long long i = atomicAdd(&c.acc->alloc_head, 1ull);
  if (i >= alloc_limit) {
    set_error(c, SIMIAN_ERR_POOL_EXHAUSTED);
    return SIMIAN_NONE;
  }
  const uint32_t p = c.free_ring[i % c.n_pages];
  if (i < c.n_pages) { // first lap of the ring: a page nobody has used yet
    scrub_page(c, p);
    __threadfence(); // dead times are visible before the page is linked into a cell
  }
  return p;
}

// Return a page to the ring.  `linked`: its page_next entry was set.
__device__ __forceinline__ void page_free(const DevCfg &c, uint32_t p, bool linked) {
  if (linked) c.page_next[p] = SIMIAN_NONE;
  unsigned long long pos = atomicAdd(&c.acc->free_tail, 1ull);
  c.free_ring[pos % c.n_pages] = p;
}

// call from all threads at the end of a kernel that appended (not k_window,
// which reports through its BlockPart)
__device__ __forceinline__ void append_ctx_flush(const DevCfg &c, AppendCtx &ax) {
  __syncthreads();
  if (threadIdx.x == 0) {
    if (ax.far_appends) atomicAdd(&c.acc->far_appends, (unsigned long long)ax.far_appends);
    if (ax.page_allocs) atomicAdd(&c.acc->page_allocs, (unsigned long long)ax.page_allocs);
  }
}

// sum a per-thread count of appended records into acc->appended (one atomic per warp)
__device__ __forceinline__ void appended_flush(const DevCfg &c, uint32_t n) {
  n = __reduce_add_sync(FULL, n);
  if ((threadIdx.x & 31) == 0 && n) atomicAdd(&c.acc->appended, (unsigned long long)n);
}

// eventQ.push (eventQ.js:29-43) as an append to the cell's page chain.
// ONE 64-bit atomicAdd on the cell's head word claims a slot and names the
// page in the same L2 round trip.  Exactly one thread installs a new page: the
// one whose claim returned PAGE (first overflow; an empty cell is born "full");
// later claimants wait for the page field to change (starvation-free under
// independent thread scheduling).
// Everything after a claim that did not land inside the head page.
__device__ __noinline__ void cell_append_slow(const DevCfg &c, AppendCtx &ax,
                                              unsigned long long alloc_limit,
                                              unsigned long long *hp, unsigned long long old,
                                              unsigned long long q0, unsigned long long q1,
                                              unsigned long long q2, unsigned long long q3) {
#pragma unroll 1
  for (;;) {
    const uint32_t p = (uint32_t)(old >> 32), i = CELL_COUNT(old);
    if (i < (uint32_t)PAGE) {
      st_rec_2x16(c.pool + ((size_t)p * PAGE + i), q0, q1, q2, q3);
      return;
    }
    if (i == (uint32_t)PAGE) { // first overflow: install the next page
      uint32_t np = page_get(c, ax, alloc_limit);
      if (np == SIMIAN_NONE) return; // sticky error set
      if (p != SIMIAN_NONE) {
        // the link carries the old head's own "older page" bit.  A reader that
        // sees the new head word before this store (no fence here) finds NONE
        // -- page_next of a free page -- and re-reads (list_walk).
        c.page_next[np] = p | ((uint32_t)old & CELL_OLDER);
      } else {
        // the cell's first page: mark its ring row as occupied (LBTS search)
        const uint32_t row = (uint32_t)((unsigned long long)(hp - c.cell_state) % c.rows);
        if (row < c.ring) {
          uint32_t *w = c.row_bits + (row >> 5);
          const uint32_t bit = 1u << (row & 31u);
          if (!(ld_cg(w) & bit)) atomicOr(w, bit);
        }
      }
      st_rec_2x16(c.pool + (size_t)np * PAGE, q0, q1, q2, q3); // slot 0 is ours
      // no fence: appenders of this launch only write into the page; records
      // are read after a launch boundary or after the ticket fence
      atomicExch(hp, ((unsigned long long)np << 32) |
                         (unsigned long long)(p == SIMIAN_NONE ? 0u : CELL_OLDER) | 1ull);
      return;
    }
    uint32_t spins = 0;
#pragma unroll 1
    while ((uint32_t)(ld_cg64(hp) >> 32) == p) { // wait for the installer
      if (ld_cg(&c.acc->error)) return;
      if (++spins > (1u << 24)) { // seconds: never legitimately; fail, do not hang
        set_error(c, SIMIAN_ERR_STALLED);
        return;
      }
      __nanosleep(40);
    }
    old = atomicAdd(hp, 1ull);
  }
}

__device__ __forceinline__ void cell_append(const DevCfg &c, AppendCtx &ax,
                                            unsigned long long alloc_limit,
                                            unsigned long long *hp, const RecBits &r) {
  unsigned long long old = atomicAdd(hp, 1ull);
  if (CELL_COUNT(old) < (uint32_t)PAGE)
    st_rec_2x16(c.pool + ((size_t)(uint32_t)(old >> 32) * PAGE + CELL_COUNT(old)), r.q0, r.q1,
                r.q2, r.q3);
  else
    cell_append_slow(c, ax, alloc_limit, hp, old, r.q0, r.q1, r.q2, r.q3);
}

// calendar cell (index into cell_state) of a record whose dst is a LOCAL SLOT;
// SIMIAN_NONE: error set
__device__ __forceinline__ uint32_t cell_of_local(const DevCfg &c, AppendCtx &ax,
                                                  long long tb_clean, uint32_t far_row,
                                                  double time, uint32_t dst_slot) {
  const uint32_t g = dst_slot >> GE_SHIFT;
  if (g >= c.n_groups) { // an id the model made up (reference: throws at dispatch)
    set_error(c, SIMIAN_ERR_UNKNOWN_NAME);
    return SIMIAN_NONE;
  }
  long long tb = tb_of(c, time);
  if (tb < tb_clean) {
    set_error(c, SIMIAN_ERR_OUT_OF_ORDER);
    return SIMIAN_NONE;
  }
  if (tb - tb_clean < (long long)c.ring) return g * c.rows + (uint32_t)(tb & (long long)c.ring_mask);
  atomicMin(&c.acc->far_min[far_row], simian_time_key(time));
  atomicAdd(&ax.far_appends, 1u);
  return g * c.rows + c.ring + far_row;
}

// route a record whose dst is a LOCAL SLOT into the calendar or the far row
__device__ __forceinline__ void append_local(const DevCfg &c, AppendCtx &ax, long long tb_clean,
                                             uint32_t far_row, unsigned long long alloc_limit,
                                             const RecBits &r) {
  uint32_t cell = cell_of_local(c, ax, tb_clean, far_row, r.time(), r.dst());
  if (cell != SIMIAN_NONE) cell_append(c, ax, alloc_limit, c.cell_state + cell, r);
}

// ------------------------------------------------------------ init / load --

__global__ void k_init(const __grid_constant__ DevCfg c, unsigned long long n_cells) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t j = i; j < c.n_pages; j += stride) {
    c.free_ring[j] = (uint32_t)j;
    c.page_next[j] = SIMIAN_NONE;
  }
  for (size_t j = i; j < n_cells; j += stride) c.cell_state[j] = CELL_EMPTY;
  for (size_t j = i; j < SIMIAN_PART_SLOTS; j += stride) {
    BlockPart z;
    z.min_key = SIMIAN_KEY_NONE;
    z.max_key = 0;
    z.committed = z.emitted = z.dropped = z.ties = z.dties = z.appended = z.far_appends =
        z.page_allocs = 0;
    c.part[j] = z;
  }
  // group page caches start with SIMIAN_BC_FILL pages each when the pool allows
  const uint32_t fill =
      ((unsigned long long)c.n_groups * SIMIAN_BC_FILL * 2ull <= c.n_pages) ? SIMIAN_BC_FILL : 0u;
  for (size_t j = i; j < (size_t)c.n_groups * SIMIAN_BC_ROW; j += stride) {
    uint32_t g = (uint32_t)(j / SIMIAN_BC_ROW), k = (uint32_t)(j % SIMIAN_BC_ROW);
    c.bcache[j] = k < fill ? g * fill + k : (k == SIMIAN_BC ? fill : SIMIAN_NONE);
  }
  // Free-page invariant: pages the caches start with get dead time words here;
  // every other page is scrubbed when it leaves the ring for the first time
  // (page_get, cache refill), so an engine pays for the pages it uses, not for
  // the pool it reserved.
  {
    RecBits dead;
    dead.q0 = SIMIAN_TIME_DEAD;
    dead.q1 = 0xFFFFFFFFFFFFFFFFull;
    dead.q2 = dead.q3 = 0;
    const size_t nrec = (size_t)c.n_groups * fill * PAGE;
    for (size_t j = i; j < nrec; j += stride) st_rec(c.pool + j, dead);
  }
  for (size_t j = i; j < (size_t)c.n_groups * GE; j += stride) {
    EntityCore z;
    z.count = 0;
    z.hash = 0;
    c.core[j] = z;
  }
  for (size_t j = i; j < c.ring / 32; j += stride) c.row_bits[j] = 0;
  if (i == 0) {
    DevAcc *a = c.acc;
    a->min_key = SIMIAN_KEY_NONE;
    a->far_min[0] = a->far_min[1] = SIMIAN_KEY_NONE;
    a->alloc_head = (unsigned long long)c.n_groups * fill; // ring slots [0, alloc_head) went to the caches
    a->free_tail = c.n_pages;
    a->committed = a->emitted = a->dropped = a->ties = a->dist_ties = 0;
    a->max_key = 0;
    a->windows = a->redistributions = a->far_appends = a->page_allocs = a->leaked_pages = 0;
    a->appended = 0;
    a->ticket = 0;
    a->error = 0;
    a->window_index = 0;
    a->loop_arrive = 0;
    a->minvec[0] = a->minvec[1] = 0;
    for (int q = 0; q < 12; q++) a->phase_cycles[q] = 0;
    WinState w;
    w.gmin = c.start;
    w.hi = c.start + c.min_delay;
    w.tb_lo = 0;
    w.tb_hi = 0;
    w.tb_clean = tb_of(c, c.start);
    w.tb_clean_new = w.tb_clean;
    w.done = 0;
    w.redist = 0;
    w.far_row = 0;
    w.parity = 0;
    w.alloc_limit = c.n_pages;
    c.win[0] = w;
    c.win[1] = w;
  }
  for (size_t j = i; j < 2 * (size_t)c.size; j += stride) {
    if (c.outbox_count) c.outbox_count[j] = 0;
    if (c.xch) {
      MinSlot z;
      z.min_left = z.far_min = 0;
      z.seq = z.pad_ = 0;
      c.xch[j] = z;
    }
  }
}

// schedService loop on the device (simian.js:170-190 for every (k, i))
__global__ void k_sched_bulk(const __grid_constant__ DevCfg c, double time, uint32_t handler, unsigned long long data,
                             int cls, uint32_t first, uint32_t count, uint32_t repeat,
                             uint32_t seq_base) {
  __shared__ AppendCtx ax;
  append_ctx_init(ax);
  __syncthreads();
  const WinState &W = c.win[0];
  unsigned long long total = (unsigned long long)count * repeat;
  unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  uint32_t napp = 0;
  for (unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
       idx < total; idx += stride) {
    uint32_t i = (uint32_t)(idx % count);
    uint32_t gid = c.cls[cls].first_id + first + i;
    int rank;
    uint32_t slot;
    route_gid(c, gid, rank, slot);
    if (rank != c.rank) continue; // only the owner pushes (simian.js:177)
    RecBits r = make_rec(time, slot, SIMIAN_NULL_ENTITY, handler, SIMIAN_EVF_SCHED,
                         seq_base + (uint32_t)idx, data);
    append_local(c, ax, W.tb_clean, W.far_row, W.alloc_limit, r);
    napp++;
  }
  appended_flush(c, napp);
  append_ctx_flush(c, ax);
}

// Records -> local calendar.  dst_is_gid: records come from the host API
// (global ids, owner filter); otherwise from a peer's outbox (dst = our slot).
__global__ void k_ingest(const __grid_constant__ DevCfg c, const simian_event_t *recs, unsigned long long n,
                         int dst_is_gid, int track_min, uint32_t win_ix) {
  __shared__ AppendCtx ax;
  __shared__ unsigned long long smin;
  append_ctx_init(ax);
  if (threadIdx.x == 0) smin = SIMIAN_KEY_NONE;
  __syncthreads();
  const WinState &W = c.win[win_ix & 1u];
  unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  unsigned long long mn = SIMIAN_KEY_NONE;
  uint32_t napp = 0;
  for (unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
       idx < n; idx += stride) {
    RecBits r = ld_rec(recs + idx);
    if (dst_is_gid) {
      if (r.time() > c.end) continue; // simian.js:173
      if (r.dst() >= c.n_ids) {
        set_error(c, SIMIAN_ERR_UNKNOWN_NAME);
        continue;
      }
      if (!(r.time() >= c.start)) { // also NaN
        set_error(c, SIMIAN_ERR_OUT_OF_ORDER);
        continue;
      }
      int rank;
      uint32_t slot;
      route_gid(c, r.dst(), rank, slot);
      if (rank != c.rank) continue;
      r.q1 = (unsigned long long)slot | (r.q1 & 0xFFFFFFFF00000000ull);
    } else if (r.dst() >= c.n_slots) { // a peer sent a slot this rank does not have
      set_error(c, SIMIAN_ERR_UNKNOWN_NAME);
      continue;
    }
    unsigned long long k = simian_time_key(r.time());
    mn = k < mn ? k : mn;
    append_local(c, ax, W.tb_clean, W.far_row, W.alloc_limit, r);
    napp++;
  }
  appended_flush(c, napp);
  if (track_min) {
    mn = warp_min_u64(mn);
    if ((threadIdx.x & 31) == 0 && mn != SIMIAN_KEY_NONE) atomicMin(&smin, mn);
    __syncthreads();
    if (threadIdx.x == 0 && smin != SIMIAN_KEY_NONE) atomicMin(&c.acc->min_key, smin);
  }
  append_ctx_flush(c, ax);
}

// Cross-GPU exchange, receiver side (simian.js:138-142): the records the peers
// staged for this rank in window win_ix are read straight out of THEIR
// outboxes over NVLink (coalesced 32-byte records) and appended to the local
// calendar -- alltoallv and ingest in one kernel, counts never visit the host.
// Runs after the window's allreduce, which orders it behind every peer's
// k_window.  gridDim.x = (size - 1) * blocks_per_peer.
__global__ void k_pull(const __grid_constant__ DevCfg c, uint32_t win_ix, uint32_t blocks_per_peer,
                       int wait_flags) {
  __shared__ AppendCtx ax;
  append_ctx_init(ax);
  const WinState &W = c.win[win_ix & 1u];
  if (wait_flags && !W.done && threadIdx.x < (unsigned)c.size) {
    // every rank has published window win_ix's candidate, i.e. finished its
    // k_window: the outboxes are complete (the barrier MPI_Alltoall is, mpi.cpp:422)
    const volatile unsigned long long *f =
        &c.xch[W.parity * (uint32_t)c.size + threadIdx.x].seq;
    const unsigned long long want = c.run_nonce + (unsigned long long)win_ix + 1ull;
    uint32_t spins = 0;
    while (*f != want) {
      if (ld_cg(&c.acc->error)) break;
      if (++spins > (1u << 27)) { // > 10 s: a peer died or never launched this window
        set_error(c, SIMIAN_ERR_STALLED);
        break;
      }
      __nanosleep(100);
    }
    __threadfence_system();
  }
  __syncthreads();
  uint32_t napp = 0;
  if (!W.done) {
    const uint32_t q = blockIdx.x / blocks_per_peer, sub = blockIdx.x % blocks_per_peer;
    const uint32_t p = ((uint32_t)c.rank + 1u + q) % (uint32_t)c.size; // never c.rank
    const uint32_t box = W.parity * (uint32_t)c.size + (uint32_t)c.rank;
    const uint32_t *pc = c.peer_count[p];
    const simian_event_t *src = c.peer_outbox[p];
    if (pc && src) {
      uint32_t n = *(const volatile uint32_t *)(pc + box);
      n = n < c.outbox_cap ? n : c.outbox_cap;
      src += (size_t)box * c.outbox_cap;
      // four 32-byte NVLink reads in flight per thread, then the four appends
      const uint32_t stride = blocks_per_peer * blockDim.x;
      for (uint32_t idx = sub * blockDim.x + threadIdx.x; idx < n; idx += 4 * stride) {
        RecBits r[4];
#pragma unroll
        for (int u = 0; u < 4; u++)
          if (idx + u * stride < n) r[u] = ld_rec(src + idx + u * stride);
#pragma unroll
        for (int u = 0; u < 4; u++)
          if (idx + u * stride < n) {
            if (r[u].dst() >= c.n_slots) {
              set_error(c, SIMIAN_ERR_UNKNOWN_NAME);
              continue;
            }
            append_local(c, ax, W.tb_clean, W.far_row, W.alloc_limit, r[u]);
            napp++;
          }
      }
    } else {
      set_error(c, SIMIAN_ERR_BAD_ARGUMENT); // peers were never connected
    }
  }
  appended_flush(c, napp);
  append_ctx_flush(c, ax);
}

// ---------------------------------------------------------- window kernel --
// (k_construct, which needs the handler context, is defined after DevCtx)

// everything one warp (= one entity group) keeps on chip for a window
struct __align__(32) WarpShared {
  simian_event_t slots[NSLOT];  // the window's records; the pure path overwrites slot s
                                // with the record event s emits (staged, then appended)
  EntityCore core_s[GE];        // {commit count, trace hash} of the group's entities
  AppendCtx ax;                 // the group's page cache + append counters
  unsigned long long xs[NSLOT]; // trace words in commit order, entity after entity
  WinState W;
  unsigned long long mbar;      // completion barrier of the bulk copies
  long long t_last;
  // the warp's results of this window (K1), updated by lane 0 after warp-level
  // reductions: no lane keeps accumulators in registers across the phases
  unsigned long long acc_min, acc_max;
  uint32_t acc_committed, acc_emitted, acc_dropped, acc_ties, acc_dties, acc_appended;
  uint32_t cnt[GE];             // due events per entity
  uint32_t pg_id[PGMAX];        // pages of the group's cells in the window rows (one chunk)
  uint32_t npg, total;
  uint16_t pg_cnt[PGMAX];       // records in the page | PG_LINKED | PG_FREE_FLAG
  uint16_t pg_off[PGMAX];       // records in the pages listed before this one
  uint16_t base[GE];            // first sorted position of each entity's events
  uint16_t ent[NSLOT];          // per slot: entity | arrival index << 8 (0xFFFF: not due)
  uint8_t ord[NSLOT];           // slots grouped by entity (arrival order inside an entity)
  uint8_t srt[NSLOT];           // slots in commit order: entity after entity, key order inside
};

struct CtaShared { // per CTA: result accumulators + scratch of the advance step
  unsigned long long min_key, max_key;
  uint32_t committed, emitted, dropped, ties, dties, appended, far_appends, page_allocs;
  uint32_t is_last, found;
  long long scan_j;
  unsigned long long red[10];
  WinState nw;
};

// dynamic shared memory of k_window for wpc warps per CTA
#define WINDOW_SMEM_BYTES(wpc) (sizeof(CtaShared) + (size_t)(wpc) * sizeof(WarpShared))

// per-thread accumulators of the handler phase
struct ThreadAcc {
  unsigned long long min_key, max_key;
  uint32_t committed, emitted, dropped, ties, dties, appended;
};

__device__ __forceinline__ void thread_acc_init(ThreadAcc &ta) {
  ta.min_key = SIMIAN_KEY_NONE;
  ta.max_key = 0;
  ta.committed = ta.emitted = ta.dropped = ta.ties = ta.dties = ta.appended = 0;
}

// fold warp-level results into the warp's accumulators (call with all lanes;
// the values are warp-uniform or reduced here; lane 0 writes)
__device__ __forceinline__ void wacc_min(WarpShared &ws, unsigned long long lane_min) {
  const unsigned long long mn = warp_min_u64(lane_min);
  if ((threadIdx.x & 31u) == 0 && mn < ws.acc_min) ws.acc_min = mn;
}
__device__ __forceinline__ void wacc_max(WarpShared &ws, unsigned long long lane_max) {
  const unsigned long long mx = warp_max_u64(lane_max);
  if ((threadIdx.x & 31u) == 0 && mx > ws.acc_max) ws.acc_max = mx;
}
// per-lane ThreadAcc of the per-entity path -> the warp's accumulators
__device__ __forceinline__ void wacc_thread(WarpShared &ws, const ThreadAcc &ta) {
  wacc_min(ws, ta.min_key);
  wacc_max(ws, ta.max_key);
  uint32_t s0 = __reduce_add_sync(FULL, ta.committed);
  uint32_t s1 = __reduce_add_sync(FULL, ta.emitted);
  uint32_t s2 = __reduce_add_sync(FULL, ta.dropped);
  uint32_t s3 = __reduce_add_sync(FULL, ta.ties);
  uint32_t s4 = __reduce_add_sync(FULL, ta.dties);
  uint32_t s5 = __reduce_add_sync(FULL, ta.appended);
  if ((threadIdx.x & 31u) == 0) {
    ws.acc_committed += s0;
    ws.acc_emitted += s1;
    ws.acc_dropped += s2;
    ws.acc_ties += s3;
    ws.acc_dties += s4;
    ws.acc_appended += s5;
  }
}

// what a handler reaches through `self` on the device (see simian_models.h)
// SEQ: one thread per entity (stateful handlers, same-window self-sends);
// !SEQ: one thread per event (pure handlers).
template <bool SEQ>
struct DevCtx {
  const DevCfg &c;
  const WinState &W; // window bound / recycle frontier / far row in force
  AppendCtx &ax;
  // what the handler did, read by the caller afterwards (kept here, by value,
  // so that no per-thread accumulator struct has its address taken)
  unsigned long long acc_min; // min time key of what it sent (LBTS candidate)
  uint32_t acc_emitted, acc_dropped, acc_appended;
  const DevClass *cl;
  uint32_t self_gid, self_slot;
  unsigned long long commit_idx, key;
  uint32_t n_emit;
  double now_;
  uint32_t handler_, src_;
  unsigned long long data_;
  // !SEQ: where the first locally routed record is staged (a shared-memory
  // slot) instead of being appended right away; null = append directly
  simian_event_t *stage;
  // fold the times of LOCALLY routed records into the LBTS candidate (window
  // kernel: yes; constructors: no -- the calendar search finds those, and a
  // candidate that a later window commits would move the window bounds)
  bool local_min;
  // same-window self-sends (rx omitted, offset < minDelay allowed: entity.js:41)
  int npend;
  double pend_time[SEQ ? SIMIAN_SELF_MAX : 1];
  unsigned long long pend_q2[SEQ ? SIMIAN_SELF_MAX : 1], pend_data[SEQ ? SIMIAN_SELF_MAX : 1];

  __device__ __forceinline__ DevCtx(const DevCfg &c_, const WinState &W_, AppendCtx &ax_)
      : c(c_), W(W_), ax(ax_), acc_min(SIMIAN_KEY_NONE), acc_emitted(0), acc_dropped(0),
        acc_appended(0), stage(nullptr), local_min(true), npend(0) {}

  __device__ __forceinline__ double now() const { return now_; }
  __device__ __forceinline__ uint32_t self_id() const { return self_gid; }
  __device__ __forceinline__ unsigned long long draw(uint32_t i) const {
    return simian_rng_draw(key, i);
  }
  __device__ __forceinline__ uint16_t handler() const { return (uint16_t)handler_; }
  __device__ __forceinline__ const void *params() const { return cl->params; }
  __device__ __forceinline__ void *state() const {
    return cl->state_bytes
               ? (void *)(cl->state + (size_t)(self_slot - cl->slot_base) * cl->state_bytes)
               : nullptr;
  }
  __device__ __forceinline__ uint32_t src() const { return src_; }
  __device__ __forceinline__ unsigned long long data() const { return data_; }

  __device__ __forceinline__ void put_local(const RecBits &r) {
    if (local_min) {
      unsigned long long tk = simian_time_key(r.time());
      acc_min = tk < acc_min ? tk : acc_min;
    }
    acc_appended++;
    if (!SEQ && stage) { // K4 is done by the append phase of k_window
      ulonglong2 *d = reinterpret_cast<ulonglong2 *>(stage);
      d[0] = make_ulonglong2(r.q0, r.q1);
      d[1] = make_ulonglong2(r.q2, r.q3);
      stage = nullptr;
    } else {
      append_local(c, ax, W.tb_clean, W.far_row, W.alloc_limit, r);
    }
  }

  // Entity.reqService (entity.js:35-68)
  __device__ __forceinline__ void req_service(double offset, uint16_t service,
                                              unsigned long long data, uint32_t rx) {
    uint32_t k = n_emit++;
    if (rx != SIMIAN_NULL_ENTITY && offset < c.min_delay) { // entity.js:41-45
      set_error(c, SIMIAN_ERR_LOOKAHEAD);
      return;
    }
    double time = now_ + offset; // entity.js:47
    if (time > c.end) {          // entity.js:48
      acc_dropped++;
      return;
    }
    uint32_t seq = simian_emit_seq(commit_idx, k);
    acc_emitted++;
    if (rx == SIMIAN_NULL_ENTITY) { // entity.js:50-51
      if (time < now_) {
        set_error(c, SIMIAN_ERR_OUT_OF_ORDER); // simian.js:124-125 at pop time
        return;
      }
      if (time < W.hi) { // lands inside this window: stays with the entity
        if (!SEQ || npend >= SIMIAN_SELF_MAX) {
          set_error(c, SIMIAN_ERR_SELF_PENDING);
          return;
        }
        pend_time[npend] = time;
        pend_q2[npend] = (unsigned long long)service | ((unsigned long long)seq << 32);
        pend_data[npend] = data;
        npend++;
        return;
      }
      put_local(make_rec(time, self_slot, self_gid, service, 0, seq, data));
      return;
    }
    if (rx >= c.n_ids) { // the reference throws at dispatch (simian.js:128)
      set_error(c, SIMIAN_ERR_UNKNOWN_NAME);
      return;
    }
    int rank;
    uint32_t slot;
    route_gid(c, rx, rank, slot); // getOffsetRank, entity.js:62
    RecBits r = make_rec(time, slot, self_gid, service, 0, seq, data);
    if (rank == c.rank) { // entity.js:64
      put_local(r);
    } else { // entity.js:66: staged for the per-window exchange
      // the receiver has not seen it when the LBTS candidates are reduced:
      // the sender's candidate covers it (simian.js:143-144)
      unsigned long long tk = simian_time_key(time);
      acc_min = tk < acc_min ? tk : acc_min;
      // warp-aggregated claim: one atomic per (warp, destination rank)
      const uint32_t box = W.parity * (uint32_t)c.size + (uint32_t)rank;
      unsigned act = __activemask();
      unsigned peers = __match_any_sync(act, rank);
      int leader = __ffs(peers) - 1, ln = threadIdx.x & 31;
      uint32_t i = 0;
      if (ln == leader) i = atomicAdd(&c.outbox_count[box], (uint32_t)__popc(peers));
      i = __shfl_sync(peers, i, leader) + __popc(peers & ((1u << ln) - 1u));
      if (i >= c.outbox_cap) {
        set_error(c, SIMIAN_ERR_OUTBOX_FULL);
        return;
      }
      if (SEQ) // (inside the __noinline__ per-entity function: see st_rec_2x16)
        st_rec_2x16(c.outbox + ((size_t)box * c.outbox_cap + i), r.q0, r.q1, r.q2, r.q3);
      else
        st_rec(c.outbox + ((size_t)box * c.outbox_cap + i), r);
    }
  }
};

// the pure handlers only (thread-per-event path: simian_fn_is_pure)
__device__ __forceinline__ void run_pure_handler(DevCtx<false> &ctx, uint32_t fn) {
  if (fn == SIMIAN_FN_PHOLD_GENERATE)
    simian_phold_generate(ctx);
  else if (fn == SIMIAN_FN_COUNT_ONLY)
    simian_noop(ctx);
  else if (fn == SIMIAN_FN_HELLO_GENERATE)
    simian_hello_generate(ctx);
  else if (fn == SIMIAN_FN_HELLO_SQUARE)
    simian_hello_square(ctx);
  else if (fn == SIMIAN_FN_HELLO_SQRT)
    simian_hello_sqrt(ctx);
  else if (fn == SIMIAN_FN_HELLO_RESULT)
    simian_hello_result(ctx);
  else
    set_error(ctx.c, SIMIAN_ERR_UNKNOWN_NAME);
}

template <bool SEQ>
__device__ __forceinline__ void run_handler(DevCtx<SEQ> &ctx, uint32_t fn) {
  switch (fn) {
    case SIMIAN_FN_PHOLD_GENERATE:
      simian_phold_generate(ctx);
      break;
    case SIMIAN_FN_COUNT_ONLY:
      simian_noop(ctx);
      break;
    case SIMIAN_FN_LANL_SEND:
      simian_lanl_send(ctx);
      break;
    case SIMIAN_FN_LANL_RECEIVE:
      simian_lanl_receive(ctx);
      break;
    case SIMIAN_FN_LANL_FINISH:
      simian_lanl_finish(ctx);
      break;
    case SIMIAN_FN_LANL_CONSTRUCT:
      simian_lanl_construct(ctx);
      break;
    case SIMIAN_FN_HELLO_GENERATE:
      simian_hello_generate(ctx);
      break;
    case SIMIAN_FN_HELLO_SQUARE:
      simian_hello_square(ctx);
      break;
    case SIMIAN_FN_HELLO_SQRT:
      simian_hello_sqrt(ctx);
      break;
    case SIMIAN_FN_HELLO_RESULT:
      simian_hello_result(ctx);
      break;
    default:
      set_error(ctx.c, SIMIAN_ERR_UNKNOWN_NAME);
  }
}

// `new entityClass(...)` bodies (simian.js:227) on the device: the named
// constructor runs once per local instance at now = startTime; not an event.
// Everything it sends goes to the calendar (no window is open yet).
__global__ void k_construct(const __grid_constant__ DevCfg c, int cls, uint32_t fn) {
  __shared__ AppendCtx ax;
  __shared__ WinState W;
  append_ctx_init(ax);
  if (threadIdx.x == 0) {
    W = c.win[0];
    W.hi = __longlong_as_double(0xFFF0000000000000LL); // -inf: nothing is "inside the window"
  }
  __syncthreads();
  const DevClass &cl = c.cls[cls];
  ThreadAcc ta;
  thread_acc_init(ta);
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < cl.n_local;
       j += gridDim.x * blockDim.x) {
    DevCtx<false> ctx(c, W, ax);
    ctx.local_min = false;
    int k;
    ctx.self_slot = cl.slot_base + j;
    ctx.self_gid = gid_of_slot(c, ctx.self_slot, k);
    ctx.cl = &cl;
    ctx.commit_idx = SIMIAN_CTOR_COMMIT;
    ctx.key = simian_rng_key(c.seed, ctx.self_gid, SIMIAN_CTOR_COMMIT);
    ctx.n_emit = 0;
    ctx.now_ = c.start; // simian.js:44
    ctx.handler_ = 0;
    ctx.src_ = SIMIAN_NULL_ENTITY;
    ctx.data_ = 0;
    run_handler(ctx, fn);
    ta.emitted += ctx.acc_emitted;
    ta.dropped += ctx.acc_dropped;
    ta.appended += ctx.acc_appended;
    ta.min_key = ctx.acc_min < ta.min_key ? ctx.acc_min : ta.min_key;
  }
  if (ta.emitted) atomicAdd(&c.acc->emitted, (unsigned long long)ta.emitted);
  if (ta.dropped) atomicAdd(&c.acc->dropped, (unsigned long long)ta.dropped);
  if (ta.min_key != SIMIAN_KEY_NONE) atomicMin(&c.acc->min_key, ta.min_key);
  appended_flush(c, ta.appended);
  append_ctx_flush(c, ax);
}

__device__ __forceinline__ ulonglong2 slot_lo(const WarpShared &ws, uint32_t s) {
  return reinterpret_cast<const ulonglong2 *>(&ws.slots[s])[0]; // {time, dst | src << 32}
}
__device__ __forceinline__ ulonglong2 slot_hi(const WarpShared &ws, uint32_t s) {
  return reinterpret_cast<const ulonglong2 *>(&ws.slots[s])[1]; // {handler | flags | seq, data}
}

// All due events of one entity, in commit order (simian.js:122-134 restricted
// to one entity; entities never interact inside a window because cross-entity
// sends carry offset >= minDelay, entity.js:41).  Its calendar events are
// srt[b .. b+n) (sorted by the rank phase); same-window self-sends are merged
// in by time.  One thread.
// `carry` (a hot entity whose window is cut into time slices, split passes):
// the same-window self-sends still pending and the tie state travel from slice
// to slice through ws.xs (unused by this path); a pending self-send is
// committed in the slice its time bin belongs to (bin < bin_hi).
__device__ __forceinline__ uint32_t time_bin(const WinState &W, double t) {
  const double x = (t - W.gmin) * (64.0 / (W.hi - W.gmin));
  return x < 63.0 ? (uint32_t)x : 63u;
}
#define CARRY_NPEND 0
#define CARRY_TIME 1
#define CARRY_Q2 (1 + SIMIAN_SELF_MAX)
#define CARRY_DATA (1 + 2 * SIMIAN_SELF_MAX)
#define CARRY_LAST (1 + 3 * SIMIAN_SELF_MAX) // have_last, last_time, last_q, last_data
static_assert(CARRY_LAST + 4 <= NSLOT, "carry area lives in ws.xs");

__device__ __noinline__ void process_entity(const DevCfg &c, WarpShared &ws, ThreadAcc &ta,
                                            uint32_t slot, uint32_t le, uint32_t b, uint32_t n,
                                            bool carry, uint32_t bin_hi) {
  int k;
  DevCtx<true> ctx(c, ws.W, ws.ax);
  ctx.self_slot = slot;
  ctx.self_gid = gid_of_slot(c, slot, k);
  ctx.cl = &c.cls[k];
  EntityCore st = ws.core_s[le];
  bool have_last = false;
  double last_time = 0;
  unsigned long long last_q = 0, last_data = 0;
  if (carry) {
    ctx.npend = (int)ws.xs[CARRY_NPEND];
    for (int j = 0; j < ctx.npend; j++) {
      ctx.pend_time[j] = __longlong_as_double((long long)ws.xs[CARRY_TIME + j]);
      ctx.pend_q2[j] = ws.xs[CARRY_Q2 + j];
      ctx.pend_data[j] = ws.xs[CARRY_DATA + j];
    }
    have_last = ws.xs[CARRY_LAST] != 0;
    last_time = __longlong_as_double((long long)ws.xs[CARRY_LAST + 1]);
    last_q = ws.xs[CARRY_LAST + 2];
    last_data = ws.xs[CARRY_LAST + 3];
  }
  uint32_t taken = 0;
#pragma unroll 1
  for (;;) {
    const bool have_g = taken < n;
    const uint32_t gs = have_g ? ws.srt[b + taken] : 0u;
    // earliest pending same-window self-send (of this time slice)
    int pj = -1;
    for (int j = 0; j < ctx.npend; j++) {
      if (carry && time_bin(ws.W, ctx.pend_time[j]) >= bin_hi) continue;
      if (pj < 0 || ctx.pend_time[j] < ctx.pend_time[pj] ||
          (ctx.pend_time[j] == ctx.pend_time[pj] && ctx.pend_q2[j] < ctx.pend_q2[pj]))
        pj = j;
    }
    double time;
    uint32_t handler, src;
    unsigned long long data;
    if (pj >= 0 &&
        (!have_g || ctx.pend_time[pj] < __longlong_as_double((long long)slot_lo(ws, gs).x))) {
      time = ctx.pend_time[pj];
      handler = (uint32_t)(ctx.pend_q2[pj] & 0xFFFFu);
      src = ctx.self_gid;
      data = ctx.pend_data[pj];
      ctx.npend--;
      ctx.pend_time[pj] = ctx.pend_time[ctx.npend];
      ctx.pend_q2[pj] = ctx.pend_q2[ctx.npend];
      ctx.pend_data[pj] = ctx.pend_data[ctx.npend];
    } else if (have_g) {
      const ulonglong2 a = slot_lo(ws, gs), bh = slot_hi(ws, gs);
      time = __longlong_as_double((long long)a.x);
      handler = (uint32_t)(bh.x & 0xFFFFull);
      src = (uint32_t)(a.y >> 32);
      data = bh.y;
      taken++;
    } else
      break;

    // commit (simian.js:124-133)
    if (c.tie_check && have_last && time == last_time) {
      ta.ties++;
      unsigned long long q = ((unsigned long long)handler << 32) | src;
      if (q != last_q || data != last_data) ta.dties++;
    }
    last_time = time;
    last_q = ((unsigned long long)handler << 32) | src;
    last_data = data;
    have_last = true;
    st.hash = simian_trace_step(st.hash, time, handler, src, data);
    unsigned long long idx = st.count;
    st.count = idx + 1;
    ta.committed++;
    unsigned long long tk = simian_time_key(time);
    ta.max_key = tk > ta.max_key ? tk : ta.max_key;

    ctx.commit_idx = idx;
    ctx.key = simian_rng_key(c.seed, ctx.self_gid, idx);
    ctx.n_emit = 0;
    ctx.now_ = time; // simian.js:126
    ctx.handler_ = handler;
    ctx.src_ = src;
    ctx.data_ = data;
    uint32_t fn = handler < SIMIAN_MAX_SERVICES ? ctx.cl->fn_of_service[handler] : 0;
    run_handler(ctx, fn); // simian.js:129-131
  }
  c.core[slot] = st;
  ta.emitted += ctx.acc_emitted;
  ta.dropped += ctx.acc_dropped;
  ta.appended += ctx.acc_appended;
  ta.min_key = ctx.acc_min < ta.min_key ? ctx.acc_min : ta.min_key;
  if (carry) {
    ws.core_s[le] = st;
    ws.xs[CARRY_NPEND] = (unsigned long long)ctx.npend;
    for (int j = 0; j < ctx.npend; j++) {
      ws.xs[CARRY_TIME + j] = (unsigned long long)__double_as_longlong(ctx.pend_time[j]);
      ws.xs[CARRY_Q2 + j] = ctx.pend_q2[j];
      ws.xs[CARRY_DATA + j] = ctx.pend_data[j];
    }
    ws.xs[CARRY_LAST] = have_last ? 1ull : 0ull;
    ws.xs[CARRY_LAST + 1] = (unsigned long long)__double_as_longlong(last_time);
    ws.xs[CARRY_LAST + 2] = last_q;
    ws.xs[CARRY_LAST + 3] = last_data;
  }
}

// Rank phase, one lane per EVENT: the event in slot s sits at position p of the
// entity-grouped order; its rank inside its entity's segment under (time,
// handler, src, seq) gives its commit position q = base + rank (srt[q] = s).
// STATS (thread-per-event path): tie statistics against its predecessor,
// and the order-independent part of the trace hash parked at xs[q] so that the
// chain phase reads the words in commit order.  Returns the rank.
template <bool STATS>
__device__ __forceinline__ uint32_t rank_event(const DevCfg &c, WarpShared &ws, uint32_t slot0,
                                               uint32_t p, uint32_t s, bool &tie, bool &dtie) {
  const ulonglong2 a = slot_lo(ws, s);
  const ulonglong2 bh = slot_hi(ws, s);
  const uint32_t le = (uint32_t)a.y - slot0;
  const uint32_t n = ws.cnt[le], b = ws.base[le];
  const double tr = __longlong_as_double((long long)a.x);
  // tie-break words of this event: (handler, src) and seq
  const unsigned long long my2 = ((bh.x & 0xFFFFull) << 32) | (a.y >> 32);
  const uint32_t myseq = (uint32_t)(bh.x >> 32);
  uint32_t rank = 0;
  // tie statistics: is there an equal-time event before this one, and does
  // any of those differ from it in what the handler sees
  bool has_pred = false, any_diff = false;
  if (n > 1) {
#pragma unroll 1
    for (uint32_t j = b; j < b + n; j++) {
      if (j == p) continue;
      const uint32_t s2 = ws.ord[j];
      const ulonglong2 o = slot_lo(ws, s2);
      const double to = __longlong_as_double((long long)o.x);
      if (to < tr) {
        rank++;
      } else if (to == tr) { // a true same-entity time tie: the defined order decides
        const ulonglong2 oh = slot_hi(ws, s2);
        const unsigned long long o2 = ((oh.x & 0xFFFFull) << 32) | (o.y >> 32);
        const uint32_t oseq = (uint32_t)(oh.x >> 32);
        // identical records (same key down to seq) are ordered by their slot: any
        // order of indistinguishable events commits the same thing
        const bool before = o2 != my2 ? o2 < my2 : (oseq != myseq ? oseq < myseq : s2 < s);
        if (before) {
          rank++;
          has_pred = true;
          any_diff |= (o2 != my2) | (oh.y != bh.y);
        }
      }
    }
  }
  ws.srt[b + rank] = (uint8_t)s;
  if (STATS) {
    if (c.tie_check && has_pred) {
      tie = true;
      // distinguishable tie: the LATEST equal-time predecessor differs in (handler,
      // src) or data.  When every equal-time predecessor is indistinguishable from
      // this event (bursts of identical schedService calls) the answer is no;
      // otherwise the latest one is looked for (rare: models with real ties).
      if (any_diff) {
        uint32_t ps = SIMIAN_NONE, pseq = 0;
        unsigned long long p2 = 0, pdata = 0;
#pragma unroll 1
        for (uint32_t j = b; j < b + n; j++) {
          if (j == p) continue;
          const uint32_t s2 = ws.ord[j];
          const ulonglong2 o = slot_lo(ws, s2);
          if (__longlong_as_double((long long)o.x) != tr) continue;
          const ulonglong2 oh = slot_hi(ws, s2);
          const unsigned long long o2 = ((oh.x & 0xFFFFull) << 32) | (o.y >> 32);
          const uint32_t oseq = (uint32_t)(oh.x >> 32);
          const bool before = o2 != my2 ? o2 < my2 : (oseq != myseq ? oseq < myseq : s2 < s);
          // keep the latest predecessor: (key, seq, slot) maximal
          if (before &&
              (ps == SIMIAN_NONE || (p2 != o2 ? p2 < o2 : (pseq != oseq ? pseq < oseq : ps < s2)))) {
            ps = s2;
            p2 = o2;
            pseq = oseq;
            pdata = oh.y;
          }
        }
        if (p2 != my2 || pdata != bh.y) dtie = true;
      }
    }
    ws.xs[b + rank] =
        simian_trace_event(tr, (uint32_t)(bh.x & 0xFFFFull), (uint32_t)(a.y >> 32), bh.y);
  }
  return rank;
}

// Handler phase, one lane per EVENT (pure handlers): the handler itself
// (simian.js:129-131): RNG, log, destination.  Commit index = the entity's
// committed count + rank.  The record it sends to a local entity replaces the
// consumed one in slot s.
// `ecnt` += emitted | dropped << 8 | appended << 16; `mn` = min time key sent.
__device__ __forceinline__ void run_event(const DevCfg &c, WarpShared &ws, uint32_t slot0,
                                          uint32_t s, uint32_t rank, uint32_t &ecnt,
                                          unsigned long long &mn) {
  const ulonglong2 a = slot_lo(ws, s), bh = slot_hi(ws, s);
  const uint32_t slot = (uint32_t)a.y;
  int k;
  DevCtx<false> ctx(c, ws.W, ws.ax);
  ctx.stage = &ws.slots[s];
  ctx.self_slot = slot;
  ctx.self_gid = gid_of_slot(c, slot, k);
  ctx.cl = &c.cls[k];
  ctx.commit_idx = ws.core_s[slot - slot0].count + rank;
  ctx.key = simian_rng_key(c.seed, ctx.self_gid, ctx.commit_idx);
  ctx.n_emit = 0;
  ctx.now_ = __longlong_as_double((long long)a.x); // simian.js:126
  ctx.handler_ = (uint32_t)(bh.x & 0xFFFFull);
  ctx.src_ = (uint32_t)(a.y >> 32);
  ctx.data_ = bh.y;
  const uint32_t h = ctx.handler_;
  run_pure_handler(ctx, h < SIMIAN_MAX_SERVICES ? ctx.cl->fn_of_service[h] : 0);
  if (ctx.stage) // nothing staged: mark the slot
    reinterpret_cast<unsigned long long *>(&ws.slots[s])[1] = 0xFFFFFFFFFFFFFFFFull;
  ecnt += ctx.acc_emitted | (ctx.acc_dropped << 8) | (ctx.acc_appended << 16);
  mn = ctx.acc_min < mn ? ctx.acc_min : mn;
}

// K3 + K4 + trace chain for the records in slots [0, nvalid): classification
// (due / leftover), counting sort by entity, rank, handlers, appends, chain.
// Entities [e_lo, e_hi) of the group are committed (every due record in the
// slots belongs to one of them).  `first`: leftover times feed the LBTS minimum.
// `carry`: one time slice of a hot entity (split passes): the entity's state in
// shared memory is kept current for the next slice; bin_hi bounds the slice.
// Warp-uniform.  Results go to the warp's accumulators (ws.acc_*).
__device__ __forceinline__ void warp_pass(const DevCfg &c, WarpShared &ws, uint32_t slot0,
                                          uint32_t nvalid, uint32_t e_lo, uint32_t e_hi,
                                          bool first, bool carry = false, uint32_t bin_hi = 64u) {
  const uint32_t lane = threadIdx.x & 31u;
  ws.cnt[lane] = 0;
  ws.cnt[lane + 32] = 0;
  __syncwarp();
  // ---- classify: due records join their entity's histogram bin (the value the
  // shared-memory atomic returns is the record's arrival index inside the bin)
  {
    const double w_hi = ws.W.hi, w_gmin = ws.W.gmin;
    unsigned long long mn = SIMIAN_KEY_NONE, mx = 0;
    for (uint32_t s = lane; s < nvalid; s += 32) {
      const ulonglong2 a = slot_lo(ws, s);
      const double t = __longlong_as_double((long long)a.x);
      uint32_t e = 0xFFFFu;
      if (t < w_hi) {
        if (t >= w_gmin) {
          const uint32_t le = (uint32_t)a.y - slot0;
          if (le < GE) {
            e = le | (atomicAdd(&ws.cnt[le], 1u) << 8);
            const unsigned long long tk = simian_time_key(t);
            mx = tk > mx ? tk : mx;
          } else {
            set_error(c, SIMIAN_ERR_UNKNOWN_NAME); // a record in another group's cell
          }
        }
      } else {
        const unsigned long long tk = simian_time_key(t);
        mn = tk < mn ? tk : mn;
      }
      ws.ent[s] = (uint16_t)e;
    }
    if (first) wacc_min(ws, mn); // leftovers: K1
    if (c.all_pure) wacc_max(ws, mx); // (the per-entity path tracks what it commits)
  }
  __syncwarp();
  // ---- exclusive scan of the 64 bins (two per lane, packed 16 + 16 bits)
  uint32_t m;
  {
    const uint32_t c0 = ws.cnt[lane], c1 = ws.cnt[lane + 32];
    uint32_t x = c0 | (c1 << 16);
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t y = __shfl_up_sync(FULL, x, o);
      if (lane >= (uint32_t)o) x += y;
    }
    const uint32_t tot = __shfl_sync(FULL, x, 31);
    const uint32_t tot0 = tot & 0xFFFFu;
    m = tot0 + (tot >> 16);
    ws.base[lane] = (uint16_t)((x & 0xFFFFu) - c0);
    ws.base[lane + 32] = (uint16_t)(tot0 + (x >> 16) - c1);
  }
  __syncwarp();
  if (m == 0 && !(carry && !c.all_pure)) return;
  // ---- place: slots grouped by entity
  for (uint32_t s = lane; s < nvalid; s += 32) {
    const uint32_t e = ws.ent[s];
    if (e != 0xFFFFu) ws.ord[ws.base[e & (GE - 1u)] + (e >> 8)] = (uint8_t)s;
  }
  __syncwarp();
  PHASE(3);
  const uint32_t nr = (m + 31u) >> 5;
  if (c.all_pure) {
    // ---- rank, one lane per EVENT; the rank is parked in ent[slot]
    {
      uint32_t nt = 0, nd = 0;
#pragma unroll 1
      for (uint32_t v = 0; v < nr; v++) {
        const uint32_t p = lane + 32u * v;
        bool tie = false, dtie = false;
        if (p < m) {
          const uint32_t s = ws.ord[p];
          ws.ent[s] = (uint16_t)rank_event<true>(c, ws, slot0, p, s, tie, dtie);
        }
        if (c.tie_check) {
          nt += (uint32_t)__popc(__ballot_sync(FULL, tie));
          nd += (uint32_t)__popc(__ballot_sync(FULL, dtie));
        }
      }
      if (lane == 0) {
        ws.acc_committed += m;
        ws.acc_ties += nt;
        ws.acc_dties += nd;
      }
    }
    __syncwarp();
    PHASE(4);
    // ---- handlers (K3), one lane per EVENT; the sent record is staged in the
    // consumed record's slot
    {
      uint32_t ecnt = 0;
      unsigned long long mn = SIMIAN_KEY_NONE;
#pragma unroll 1
      for (uint32_t v = 0; v < nr; v++) {
        const uint32_t p = lane + 32u * v;
        if (p < m) {
          const uint32_t s = ws.ord[p];
          run_event(c, ws, slot0, s, ws.ent[s], ecnt, mn);
        }
      }
      wacc_min(ws, mn);
      const uint32_t s1 = __reduce_add_sync(FULL, ecnt & 255u);
      const uint32_t s2 = __reduce_add_sync(FULL, (ecnt >> 8) & 255u);
      const uint32_t s5 = __reduce_add_sync(FULL, ecnt >> 16);
      if (lane == 0) {
        ws.acc_emitted += s1;
        ws.acc_dropped += s2;
        ws.acc_appended += s5;
      }
    }
    PHASE(5); // (no barrier: the append phase reads only the slots its own lane staged)
    // ---- append (K4), one lane per EVENT; the claims of all of a lane's
    // events are in flight together
    {
      const WinState &W = ws.W;
      unsigned long long old[NROUND];
      uint32_t cell[NROUND];
      uint32_t spack[(NROUND + 3) / 4]; // the slot of each round, 8 bits each
#pragma unroll
      for (int q = 0; q < (NROUND + 3) / 4; q++) spack[q] = 0;
#pragma unroll
      for (int v = 0; v < NROUND; v++) {
        cell[v] = SIMIAN_NONE;
        if ((uint32_t)v < nr && lane + 32u * v < m) {
          const uint32_t s = ws.ord[lane + 32u * v];
          spack[v / 4] |= s << (8 * (v % 4));
          const ulonglong2 a = slot_lo(ws, s);
          if ((uint32_t)a.y != SIMIAN_NONE)
            cell[v] = cell_of_local(c, ws.ax, W.tb_clean, W.far_row,
                                    __longlong_as_double((long long)a.x), (uint32_t)a.y);
        }
      }
#pragma unroll
      for (int v = 0; v < NROUND; v++)
        if (cell[v] != SIMIAN_NONE) old[v] = atomicAdd(c.cell_state + cell[v], 1ull);
      // ---- chain, one lane per entity, while the claims are in flight: fold
      // the trace words (in commit order at xs[base ..]) and publish the count
      {
        if (carry) __syncwarp(); // the handlers have read the entity's old count
#pragma unroll
        for (int h = 0; h < 2; h++) {
          const uint32_t le = lane + 32u * h;
          const uint32_t n = ws.cnt[le];
          if (le >= e_lo && le < e_hi && n) {
            const uint32_t b = ws.base[le];
            EntityCore st = ws.core_s[le];
#pragma unroll 1
            for (uint32_t j = 0; j < n; j++) st.hash = simian_trace_chain(st.hash, ws.xs[b + j]);
            st.count += n;
            c.core[slot0 + le] = st;
            if (carry) ws.core_s[le] = st;
          }
        }
      }
      // Claims are serviced in two rounds so that no lane ever waits for a
      // page while it still owes one: first every claim that landed in a
      // page or must INSTALL one (never blocks); only then the claims that
      // have to wait for another thread's install.  (Waiting first can
      // deadlock: two threads holding crossed install/wait claims.)
      uint32_t waiting = 0;
#pragma unroll
      for (int v = 0; v < NROUND; v++)
        if (cell[v] != SIMIAN_NONE) {
          const uint32_t idx = CELL_COUNT(old[v]);
          if (idx <= (uint32_t)PAGE) { // lands in a page, or installs one
            const uint32_t s = (spack[v / 4] >> (8 * (v % 4))) & 255u;
            const ulonglong2 a = slot_lo(ws, s), bh = slot_hi(ws, s);
            if (idx < (uint32_t)PAGE) {
              RecBits r;
              r.q0 = a.x, r.q1 = a.y, r.q2 = bh.x, r.q3 = bh.y;
              st_rec(c.pool + ((size_t)(uint32_t)(old[v] >> 32) * PAGE + idx), r);
            } else {
              cell_append_slow(c, ws.ax, W.alloc_limit, c.cell_state + cell[v], old[v], a.x, a.y,
                               bh.x, bh.y);
            }
          } else {
            waiting |= 1u << v;
          }
        }
      __syncwarp(); // every install of this warp is done before any of its lanes waits
      if (waiting) { // rare
#pragma unroll
        for (int v = 0; v < NROUND; v++)
          if (waiting & (1u << v)) {
            const uint32_t s = (spack[v / 4] >> (8 * (v % 4))) & 255u;
            const ulonglong2 a = slot_lo(ws, s), bh = slot_hi(ws, s);
            cell_append_slow(c, ws.ax, W.alloc_limit, c.cell_state + cell[v], old[v], a.x, a.y,
                             bh.x, bh.y);
          }
      }
    }
    __syncwarp();
    PHASE(6);
  } else {
    // ---- stateful handlers: sort each entity's events (rank), then one lane
    // per entity commits them in order, same-window self-sends merged in
#pragma unroll 1
    for (uint32_t v = 0; v < nr; v++) {
      const uint32_t p = lane + 32u * v;
      bool t0 = false, t1 = false;
      if (p < m) rank_event<false>(c, ws, slot0, p, ws.ord[p], t0, t1);
    }
    __syncwarp();
    // (process_entity is not inlined: its accumulators live in local memory)
    ThreadAcc t2;
    thread_acc_init(t2);
    // (a time slice of a hot entity runs even without calendar events: pending
    // self-sends may fall into it)
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const uint32_t le = lane + 32u * h;
      const uint32_t n = ws.cnt[le];
      if (le >= e_lo && le < e_hi && (n || carry))
        process_entity(c, ws, t2, slot0 + le, le, ws.base[le], n, carry, bin_hi);
    }
    __syncwarp();
    wacc_thread(ws, t2);
  }
}

// ---- page listing ------------------------------------------------------------
// A lane owns one calendar cell of the window (row tb_lo + lane) and walks its
// page chain through a cursor it keeps in registers; the pages go into the
// warp's page list, at most PGMAX per chunk.

struct Cursor {
  uint32_t page; // SIMIAN_NONE: exhausted
  uint32_t cnt;  // records in it | CELL_OLDER
};

__device__ __forceinline__ Cursor cursor_of_head(unsigned long long hw) {
  Cursor k;
  k.page = (uint32_t)(hw >> 32);
  uint32_t n = CELL_COUNT(hw);
  k.cnt = (n < (uint32_t)PAGE ? n : (uint32_t)PAGE) | ((uint32_t)hw & CELL_OLDER);
  return k;
}

// page_next of a page known to have an older page.  The installer publishes the
// cell's new head word without a fence after writing the link, so a reader in
// the same launch may still see NONE (the value of a free page): re-read.
__device__ __forceinline__ uint32_t link_of(const DevCfg &c, uint32_t p) {
  uint32_t raw = ld_cg(&c.page_next[p]);
  uint32_t spins = 0;
  while (raw == SIMIAN_NONE) {
    if (++spins > (1u << 22)) {
      set_error(c, SIMIAN_ERR_STALLED);
      break;
    }
    __nanosleep(20);
    raw = ld_cg(&c.page_next[p]);
  }
  return raw;
}

// Lists the next chunk of pages.  Returns true when every cursor is exhausted
// (the list is complete); ws.npg / pg_* hold the chunk, nvalid its records.
__device__ __forceinline__ bool list_chunk(const DevCfg &c, WarpShared &ws, Cursor &cur,
                                           uint32_t free_flag, uint32_t &npg_out,
                                           uint32_t &nvalid_out) {
  const uint32_t lane = threadIdx.x & 31u;
  if (lane == 0) ws.npg = 0;
  __syncwarp();
#pragma unroll 1
  while (cur.page != SIMIAN_NONE) {
    const uint32_t idx = atomicAdd(&ws.npg, 1u);
    if (idx >= (uint32_t)PGMAX) break; // resume here in the next chunk
    const uint32_t older = cur.cnt & CELL_OLDER;
    ws.pg_id[idx] = cur.page;
    ws.pg_cnt[idx] = (uint16_t)(CELL_COUNT(cur.cnt) | (older ? PG_LINKED : 0u) | free_flag);
    if (!older) {
      cur.page = SIMIAN_NONE;
      break;
    }
    const uint32_t raw = link_of(c, cur.page);
    if (raw == SIMIAN_NONE) { // error set
      cur.page = SIMIAN_NONE;
      break;
    }
    cur.page = raw & 0x7FFFFFFFu;
    cur.cnt = (uint32_t)PAGE | (raw & CELL_OLDER); // older pages are always full
  }
  __syncwarp();
  uint32_t npg = ws.npg;
  npg = npg < (uint32_t)PGMAX ? npg : (uint32_t)PGMAX;
  // record offsets: exclusive prefix of the page counts
  uint32_t carry = 0;
  for (uint32_t i0 = 0; i0 < npg; i0 += 32) {
    const uint32_t i = i0 + lane;
    const uint32_t n = i < npg ? PG_COUNT(ws.pg_cnt[i]) : 0u;
    uint32_t x = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t y = __shfl_up_sync(FULL, x, o);
      if (lane >= (uint32_t)o) x += y;
    }
    if (i < npg) ws.pg_off[i] = (uint16_t)(carry + x - n);
    carry += __shfl_sync(FULL, x, 31);
  }
  __syncwarp();
  npg_out = npg;
  nvalid_out = carry;
  return __ballot_sync(FULL, cur.page != SIMIAN_NONE) == 0u;
}

// One scan over the pages of the current chunk (slow path).  Every due record
// has a KEY in [0, 64): its local entity (BYTIME == false), or, for the records
// of entity `ent` only, its time bin inside the window (BYTIME == true).
// COLLECT == false: histogram of the keys (ws.cnt) (+ the leftover minimum);
// COLLECT == true: the due records with key in [e_lo, e_hi) are copied into the
// slots (ws.total counts them).
template <bool COLLECT, bool BYTIME>
__device__ __forceinline__ void scan_chunk(const DevCfg &c, WarpShared &ws,
                                           unsigned long long &left_min, uint32_t slot0,
                                           uint32_t npg, uint32_t e_lo, uint32_t e_hi,
                                           uint32_t ent) {
  const uint32_t lane = threadIdx.x & 31u;
  const double w_hi = ws.W.hi, w_gmin = ws.W.gmin;
  constexpr int U = 4;
  const uint32_t lim = npg * PAGE;
  for (uint32_t base = 0; base < lim; base += U * 32) {
    unsigned long long q0[U], q1[U];
    const simian_event_t *src[U];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const uint32_t i = base + u * 32 + lane;
      const uint32_t pg = i / PAGE, sl = i % PAGE;
      ok[u] = i < lim && sl < PG_COUNT(ws.pg_cnt[pg < (uint32_t)PGMAX ? pg : 0]);
      src[u] = c.pool + ((size_t)ws.pg_id[pg < (uint32_t)PGMAX ? pg : 0] * PAGE + sl);
      if (ok[u]) ld_rec_head(src[u], q0[u], q1[u]);
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      bool due = false;
      uint32_t le = 0;
      if (ok[u]) {
        const double t = __longlong_as_double((long long)q0[u]);
        if (t < w_hi) {
          le = (uint32_t)q1[u] - slot0;
          due = t >= w_gmin;
          if (due && le >= GE) {
            set_error(c, SIMIAN_ERR_UNKNOWN_NAME);
            due = false;
          }
          if (BYTIME) {
            due = due && le == ent;
            if (due) le = time_bin(ws.W, t); // the key
          }
          if (COLLECT) due = due && le >= e_lo && le < e_hi;
        } else if (!COLLECT && !BYTIME) {
          unsigned long long tk = simian_time_key(t);
          left_min = tk < left_min ? tk : left_min;
        }
      }
      if (!COLLECT) {
        if (due) atomicAdd(&ws.cnt[le], 1u);
      } else {
        const unsigned mk = __ballot_sync(FULL, due);
        if (mk) {
          const uint32_t p0 = ws.total; // warp-uniform: updated by lane 0 below
          if (due) {
            const uint32_t pos = p0 + __popc(mk & ((1u << lane) - 1u));
            if (pos < (uint32_t)NSLOT) {
              unsigned long long q2, q3;
              ld_rec_head((const simian_event_t *)((const char *)src[u] + 16), q2, q3);
              ulonglong2 *d = reinterpret_cast<ulonglong2 *>(&ws.slots[pos]);
              d[0] = make_ulonglong2(q0[u], q1[u]);
              d[1] = make_ulonglong2(q2, q3);
            } else {
              set_error(c, SIMIAN_ERR_WINDOW_OVERFLOW); // cannot happen: ranges are sized from the histogram
            }
          }
          __syncwarp();
          if (lane == 0) ws.total = p0 + __popc(mk);
          __syncwarp();
        }
      }
    }
  }
}

// The uncommon path of a window: the pages of its rows hold more records than
// the warp has slots (bursts such as the 16 events per entity all scheduled at
// t = 0), or more pages than the list holds.  A first scan builds the histogram
// of due records per entity; the group's entities are then cut into ranges of
// at most NSLOT due records, and each range is one filtered scan + one ordinary
// pass.  Page lists longer than PGMAX are re-walked chunk by chunk in every scan.
// Kept out of line so that the common single-pass window pays nothing for it.
// BYTIME: the same for ONE entity (`ent`) with more due events than slots: its
// window is cut into time slices along a 64-bin time histogram, each slice one
// pass; the entity's state and pending self-sends carry over (warp_pass carry).
template <bool BYTIME>
__device__ __noinline__ void slow_passes(const DevCfg &c, WarpShared &ws, uint32_t slot0,
                                         unsigned long long hw, uint32_t free_flag, bool complete,
                                         uint32_t npg0, uint32_t ent) {
  const uint32_t lane = threadIdx.x & 31u;
  unsigned long long left_min = SIMIAN_KEY_NONE;
  ws.cnt[lane] = 0;
  ws.cnt[lane + 32] = 0;
  __syncwarp();
  // ---- scan 1: histogram (the first chunk is already listed)
  {
    uint32_t npg = npg0, nv;
    Cursor cur;
    cur.page = SIMIAN_NONE;
    cur.cnt = 0;
    bool done = complete;
    if (!complete) { // re-walk from the heads: the cursors of the first listing are gone
      cur = cursor_of_head(hw);
      done = list_chunk(c, ws, cur, free_flag, npg, nv);
    }
#pragma unroll 1
    for (;;) {
      scan_chunk<false, BYTIME>(c, ws, left_min, slot0, npg, 0, GE, ent);
      if (done) break;
      done = list_chunk(c, ws, cur, free_flag, npg, nv);
    }
  }
  __syncwarp();
  if (!BYTIME) wacc_min(ws, left_min); // leftovers: K1
  // exclusive prefix of the histogram, kept in registers (warp_pass reuses ws.cnt)
  const uint32_t h0 = ws.cnt[lane], h1 = ws.cnt[lane + 32];
  uint32_t x = h0, y = h1;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t a = __shfl_up_sync(FULL, x, o), b = __shfl_up_sync(FULL, y, o);
    if (lane >= (uint32_t)o) x += a, y += b;
  }
  const uint32_t t0 = __shfl_sync(FULL, x, 31);
  const uint32_t in0 = x, in1 = t0 + y; // inclusive prefix at entities lane, lane + 32
  const uint32_t total_due = __shfl_sync(FULL, in1, 31);
  if (total_due == 0) return;
  if (BYTIME) { // nothing pending, no tie state yet
    if (lane == 0) ws.xs[CARRY_NPEND] = 0, ws.xs[CARRY_LAST] = 0;
    __syncwarp();
  }
  uint32_t e_lo = 0;
#pragma unroll 1
  while (e_lo < GE) {
    // inclusive prefix just before e_lo
    const uint32_t before =
        e_lo == 0 ? 0u
                  : (e_lo <= 32 ? __shfl_sync(FULL, in0, (e_lo - 1u) & 31u)
                                : __shfl_sync(FULL, in1, (e_lo - 33u) & 31u));
    // entities e >= e_lo whose inclusive prefix still fits: a prefix of [e_lo, GE)
    const unsigned f0 = __ballot_sync(FULL, lane >= e_lo && in0 - before <= (uint32_t)NSLOT);
    const unsigned f1 = __ballot_sync(FULL, lane + 32u >= e_lo && in1 - before <= (uint32_t)NSLOT);
    const uint32_t fit = (uint32_t)__popc(f0) + (uint32_t)__popc(f1);
    uint32_t e_hi = e_lo + fit;
    if (fit == 0) { // one entity with more due events than the warp has slots
      if (BYTIME) { // ... inside 1/64 of the window: give up
        if (lane == 0) set_error(c, SIMIAN_ERR_ENTITY_BURST);
        break;
      }
      slow_passes<true>(c, ws, slot0, hw, free_flag, complete, npg0, e_lo);
      e_lo++;
      continue;
    }
    const uint32_t upto = e_hi <= 32 ? __shfl_sync(FULL, in0, (e_hi - 1u) & 31u)
                                     : __shfl_sync(FULL, in1, (e_hi - 33u) & 31u);
    if (upto != before || (BYTIME && !c.all_pure)) {
      // ---- scan 2: the due records of [e_lo, e_hi) -> slots
      if (lane == 0) ws.total = 0;
      __syncwarp();
      uint32_t npg = npg0, nv;
      Cursor cur;
      cur.page = SIMIAN_NONE;
      cur.cnt = 0;
      bool done = complete;
      if (!complete) {
        cur = cursor_of_head(hw);
        done = list_chunk(c, ws, cur, free_flag, npg, nv);
      }
#pragma unroll 1
      for (;;) {
        scan_chunk<true, BYTIME>(c, ws, left_min, slot0, npg, e_lo, e_hi, ent);
        if (done) break;
        done = list_chunk(c, ws, cur, free_flag, npg, nv);
      }
      __syncwarp();
      uint32_t m = ws.total;
      m = m < (uint32_t)NSLOT ? m : (uint32_t)NSLOT;
      if (BYTIME)
        warp_pass(c, ws, slot0, m, ent, ent + 1u, false, true, e_hi);
      else
        warp_pass(c, ws, slot0, m, e_lo, e_hi, false);
    }
    e_lo = e_hi;
  }
}

// free a whole cell (pages -> ring, head cleared); one thread.  `scrub`: the
// records may lie in the future (far rows): restore the free-page invariant.
__device__ __forceinline__ void free_cell(const DevCfg &c, unsigned long long *hp, bool scrub) {
  const unsigned long long hw = *hp;
  uint32_t p = (uint32_t)(hw >> 32), older = (uint32_t)hw & CELL_OLDER;
  while (p != SIMIAN_NONE) {
    const uint32_t raw = older ? c.page_next[p] : SIMIAN_NONE;
    if (scrub) scrub_page(c, p);
    page_free(c, p, older != 0);
    if (raw == SIMIAN_NONE) break;
    p = raw & 0x7FFFFFFFu;
    older = raw & CELL_OLDER;
  }
  *hp = CELL_EMPTY;
}

// Push the pages of the current chunk that are marked free onto the group's
// page stack in HBM (what does not fit: to the global ring).  `height`: the
// stack height, warp-uniform, updated.
__device__ __forceinline__ void recycle_chunk(const DevCfg &c, WarpShared &ws, uint32_t *stack,
                                              uint32_t npg, uint32_t &height) {
  const uint32_t lane = threadIdx.x & 31u;
  for (uint32_t i0 = 0; i0 < npg; i0 += 32) {
    const uint32_t i = i0 + lane;
    const uint32_t pc = i < npg ? ws.pg_cnt[i] : 0u;
    const bool f = (pc & PG_FREE_FLAG) != 0;
    const unsigned mk = __ballot_sync(FULL, f);
    const uint32_t room = (uint32_t)SIMIAN_BC - height;
    if (f) {
      const uint32_t k = __popc(mk & ((1u << lane) - 1u));
      const uint32_t pid = ws.pg_id[i];
      if (k < room) {
        if (pc & PG_LINKED) c.page_next[pid] = SIMIAN_NONE;
        stack[height + k] = pid;
      } else {
        page_free(c, pid, (pc & PG_LINKED) != 0);
      }
    }
    const uint32_t n = (uint32_t)__popc(mk);
    height += n < room ? n : room;
  }
}

// Far-row redistribution for group g (own cells only: no cross-warp traffic).
__device__ __noinline__ void redistribute_group(const DevCfg &c, WarpShared &ws, uint32_t g) {
  const WinState &W = ws.W;
  const uint32_t lane = threadIdx.x & 31u;
  unsigned long long *cells = c.cell_state + (size_t)g * c.rows;
  // 1. recycle rows [tb_clean, tb_clean_new): only dead records remain there
  long long nrec = W.tb_clean_new - W.tb_clean;
  if (nrec > (long long)c.ring) nrec = c.ring;
  for (long long i = lane; i < nrec; i += 32) {
    long long tb = W.tb_clean + i;
    free_cell(c, cells + (uint32_t)(tb & (long long)c.ring_mask), false);
  }
  __syncwarp();
  // 2. move the far row: into the calendar when within the new horizon, else
  //    into the other far row
  unsigned long long *far_cell = cells + c.ring + W.far_row;
  const uint32_t other = W.far_row ^ 1u;
  const unsigned long long hw = *far_cell;
  Cursor cur;
  cur.page = SIMIAN_NONE;
  cur.cnt = 0;
  if (lane == 0) cur = cursor_of_head(hw);
  bool done;
#pragma unroll 1
  do {
    uint32_t npg, nv;
    done = list_chunk(c, ws, cur, 0u, npg, nv);
    for (uint32_t i = lane; i < npg * PAGE; i += 32) {
      const uint32_t pg = i / PAGE, s = i % PAGE;
      if (s >= PG_COUNT(ws.pg_cnt[pg])) continue;
      RecBits r = ld_rec_2x16(c.pool + ((size_t)ws.pg_id[pg] * PAGE + s));
      append_local(c, ws.ax, W.tb_clean_new, other, W.alloc_limit, r);
    }
    __syncwarp();
  } while (!done);
  // 3. release the far row's pages; their records lie in the future: scrub
  if (lane == 0) free_cell(c, far_cell, true);
  __syncwarp();
}

// ------- LBTS candidate + next-window set-up (one CTA, all threads) --------

// local minLeft candidate (simian.js:143): min over every pending local event.
// acc->min_key already holds leftovers of row tb_hi + everything appended since
// the last advance; rows above tb_hi are searched only when that candidate does
// not already lie in row tb_hi (sparse case: the window jumped over idle time).
__device__ void local_min_candidate(const DevCfg &c, const WinState &W, CtaShared &sh,
                                    unsigned long long &cand_out,
                                    unsigned long long &far_out) {
  const uint32_t tid = threadIdx.x, nt = blockDim.x;
  DevAcc *a = c.acc;
  unsigned long long far = ld_cg64(&a->far_min[W.far_row]);
  unsigned long long cand = ld_cg64(&a->min_key);
  cand = far < cand ? far : cand;
  long long jmax = W.tb_clean + (long long)c.ring - 1;
  // nothing pending at all (every appended record has been committed): the
  // rows need not be searched (end of the run, simian.js:147 `infTime`)
  if (ld_cg64(&a->appended) == ld_cg64(&a->committed)) jmax = W.tb_hi;
  if (cand != SIMIAN_KEY_NONE) {
    long long tc = tb_of(c, simian_key_time(cand));
    jmax = tc < jmax ? tc : jmax;
  }
  const long long j0 = W.tb_hi + 1;
  // first occupied sub-bucket row in [j0, jmax]: one bit per ring row
  if (tid == 0) sh.scan_j = 0x7FFFFFFFFFFFFFFFLL;
  __syncthreads();
  if (j0 <= jmax) {
    long long nrows = jmax - j0 + 1;
    if (nrows > (long long)c.ring) nrows = c.ring;
    long long best = 0x7FFFFFFFFFFFFFFFLL;
    for (long long k = tid; k < nrows; k += nt) {
      const long long j = j0 + k;
      const uint32_t row = (uint32_t)(j & (long long)c.ring_mask);
      if ((ld_cg(&c.row_bits[row >> 5]) >> (row & 31u)) & 1u) {
        best = j;
        break; // this thread's rows are visited in increasing order
      }
    }
    if (best != 0x7FFFFFFFFFFFFFFFLL)
      atomicMin((unsigned long long *)&sh.scan_j, (unsigned long long)best);
  }
  __syncthreads();
  const long long jf = sh.scan_j;
  __syncthreads();
  if (jf != 0x7FFFFFFFFFFFFFFFLL) {
    if (tid == 0) sh.min_key = SIMIAN_KEY_NONE;
    __syncthreads();
    unsigned long long mn = SIMIAN_KEY_NONE;
    const uint32_t row = (uint32_t)(jf & (long long)c.ring_mask);
    for (uint32_t g = tid; g < c.n_groups; g += nt) {
      unsigned long long hw = ld_cg64(&c.cell_state[(size_t)g * c.rows + row]);
      uint32_t p = (uint32_t)(hw >> 32), f = CELL_COUNT(hw), older = (uint32_t)hw & CELL_OLDER;
      while (p != SIMIAN_NONE) {
        f = f < PAGE ? f : PAGE;
        for (uint32_t s = 0; s < f; s++) {
          unsigned long long q0 =
              ld_cg64((const unsigned long long *)(c.pool + ((size_t)p * PAGE + s)));
          unsigned long long kk = simian_time_key(__longlong_as_double((long long)q0));
          mn = kk < mn ? kk : mn;
        }
        if (!older) break;
        const uint32_t raw = ld_cg(&c.page_next[p]);
        if (raw == SIMIAN_NONE) break;
        p = raw & 0x7FFFFFFFu;
        older = raw & CELL_OLDER;
        f = PAGE;
      }
    }
    if (mn != SIMIAN_KEY_NONE) atomicMin(&sh.min_key, mn);
    __syncthreads();
    unsigned long long m2 = sh.min_key;
    cand = m2 < cand ? m2 : cand;
    __syncthreads();
  }
  cand_out = cand;
  far_out = far;
}

// Write window k+1 given globalMinLeft (simian.js:119-120,144-147).
__device__ void setup_next_window(const DevCfg &c, const WinState &W, WinState *Wn, double gmin,
                                  double far_min_time, bool after_redist, WinState &nw) {
  const uint32_t tid = threadIdx.x, nt = blockDim.x;
  DevAcc *a = c.acc;
  if (tid == 0) {
    nw = W;
    nw.redist = 0;
    if (after_redist) {
      nw.tb_clean = W.tb_clean_new;
      nw.far_row = W.far_row ^ 1u;
      a->far_min[W.far_row] = SIMIAN_KEY_NONE;
      a->redistributions++;
      gmin = W.gmin; // the window that was postponed
    } else {
      nw.tb_clean = W.tb_hi; // rows below tb_hi were recycled by this window
      a->windows++;
    }
    nw.gmin = gmin;
    nw.hi = gmin + c.min_delay; // epoch, simian.js:120
    nw.done = !(gmin <= c.end); // simian.js:119
    nw.tb_lo = tb_of(c, nw.gmin);
    nw.tb_hi = tb_of(c, nw.hi);
    if (nw.tb_lo < nw.tb_clean) nw.tb_lo = nw.tb_clean;
    nw.tb_clean_new = nw.tb_clean;
    if (!nw.done) {
      // Spend the next launch on a recycle + far-row redistribution pass when a
      // far-future event comes due, or when the window jumped so far over idle
      // time that un-recycled rows would alias the rows it reads.
      if (!after_redist &&
          (far_min_time < nw.hi || nw.tb_hi - nw.tb_clean >= (long long)(c.ring / 2))) {
        nw.redist = 1;
        nw.tb_clean_new = nw.tb_lo;
      }
      if (nw.tb_hi - nw.tb_lo + 1 > 32)
        atomicCAS(&a->error, 0u, (unsigned int)SIMIAN_ERR_RING_SPAN);
    }
    nw.alloc_limit = ld_cg64(&a->free_tail);
#ifdef SIMIAN_PROFILE_PHASES
#ifndef SIMIAN_PROFILE_KEEP_FIRST
    if (a->windows == 1) // steady-state windows only: drop the t = 0 burst window
      for (int q = 0; q < 12; q++) a->phase_cycles[q] = 0;
#endif
#endif
    a->min_key = SIMIAN_KEY_NONE;
    a->ticket = 0;
    a->window_index++;
    nw.parity = a->window_index & 1u;
  }
  __syncthreads();
  // Occupancy bits of recycled rows are cleared while nothing can append to the
  // ring rows they alias: after a window, rows [old tb_clean, tb_hi) (recycled by
  // it); before a redistribution pass, the rows it is going to recycle (the pass
  // then appends into rows up to a ring above its NEW frontier, whose bits must
  // survive -- so nothing is cleared after it).
  {
    long long lo = W.tb_clean, n = after_redist ? 0 : nw.tb_clean - lo;
    if (nw.redist) {
      if (n == 0) lo = nw.tb_clean;
      n = nw.tb_clean_new - lo;
    }
    if (n >= (long long)c.ring) {
      for (uint32_t k = tid; k < c.ring / 32; k += nt) c.row_bits[k] = 0;
    } else {
      for (long long k = tid; k < n; k += nt) {
        const uint32_t row = (uint32_t)((lo + k) & (long long)c.ring_mask);
        atomicAnd(&c.row_bits[row >> 5], ~(1u << (row & 31u)));
      }
    }
  }
  __syncthreads();
  if (tid == 0) {
    *Wn = nw;
    if (nw.done) { // later launches of either parity must see the end
      c.win[0] = nw;
      c.win[1] = nw;
    }
  }
}

// Sum the launch's results into the global accumulators (one CTA, all threads;
// runs after the ticket fence, so every CTA's contribution is visible).  The
// CTAs add into SIMIAN_PART_SLOTS accumulator slots (block index mod slots), so
// this is ONE round trip, not a loop over the blocks; the slots are cleared
// for the next launch.
__device__ void reduce_parts(const DevCfg &c, CtaShared &sh) {
  const uint32_t tid = threadIdx.x, nt = blockDim.x;
  if (tid < 10) sh.red[tid] = tid == 0 ? SIMIAN_KEY_NONE : 0ull;
  __syncthreads();
  {
    unsigned long long mn = SIMIAN_KEY_NONE, mx = 0;
    uint32_t v[8];
#pragma unroll
    for (int q = 0; q < 8; q++) v[q] = 0;
    for (uint32_t s = tid; s < SIMIAN_PART_SLOTS; s += nt) {
      BlockPart *p = &c.part[s];
      unsigned long long a = ld_cg64(&p->min_key), b = ld_cg64(&p->max_key);
      mn = a < mn ? a : mn;
      mx = b > mx ? b : mx;
      const uint32_t *w = &p->committed;
#pragma unroll
      for (int q = 0; q < 8; q++) v[q] += ld_cg(w + q);
      p->min_key = SIMIAN_KEY_NONE;
      p->max_key = 0;
#pragma unroll
      for (int q = 0; q < 8; q++) (&p->committed)[q] = 0;
    }
    mn = warp_min_u64(mn);
    mx = warp_max_u64(mx);
#pragma unroll
    for (int q = 0; q < 8; q++) v[q] = __reduce_add_sync(FULL, v[q]);
    if ((tid & 31) == 0) {
      if (mn != SIMIAN_KEY_NONE) atomicMin(&sh.red[0], mn);
      if (mx) atomicMax(&sh.red[1], mx);
#pragma unroll
      for (int q = 0; q < 8; q++)
        if (v[q]) atomicAdd(&sh.red[2 + q], (unsigned long long)v[q]);
    }
  }
  __syncthreads();
  if (tid == 0) {
    DevAcc *a = c.acc;
    unsigned long long m0 = ld_cg64(&a->min_key); // ingested / scheduled since the last advance
    a->min_key = sh.red[0] < m0 ? sh.red[0] : m0;
    unsigned long long m1 = ld_cg64(&a->max_key);
    a->max_key = sh.red[1] > m1 ? sh.red[1] : m1;
    a->committed += sh.red[2];
    a->emitted += sh.red[3];
    a->dropped += sh.red[4];
    a->ties += sh.red[5];
    a->dist_ties += sh.red[6];
    a->appended += sh.red[7];
    a->far_appends += sh.red[8];
    a->page_allocs += sh.red[9];
    __threadfence();
  }
  __syncthreads();
}

// size > 1: this rank's minLeft candidate (simian.js:143) -> minvec, and, with
// `publish`, the send half of the allreduce(MIN) over peer memory: one thread
// per peer writes the candidate into the peer's slot array -- data, fence, flag.
// One CTA, all threads; the launch's results must be visible (ticket fence or
// a kernel boundary).
__device__ void lbts_candidate(const DevCfg &c, const WinState &W, CtaShared &sh, uint32_t win_ix,
                               int publish) {
  unsigned long long cand, far;
  if (!W.done) reduce_parts(c, sh);
  if (W.redist || W.done) {
    cand = far = SIMIAN_KEY_NONE;
  } else {
    local_min_candidate(c, W, sh, cand, far);
  }
  const double inf = __longlong_as_double(0x7FF0000000000000LL);
  const double m0 = cand == SIMIAN_KEY_NONE ? c.inf_time : simian_key_time(cand);
  const double m1 = far == SIMIAN_KEY_NONE ? inf : simian_key_time(far);
  if (threadIdx.x == 0) {
    c.acc->minvec[0] = m0;
    c.acc->minvec[1] = m1;
    c.acc->ticket = 0;
  }
  if (publish && !W.done && threadIdx.x < (unsigned)c.size) {
    MinSlot *dst = c.peer_xch[threadIdx.x];
    if (dst) {
      dst += W.parity * (uint32_t)c.size + (uint32_t)c.rank;
      *(volatile double *)&dst->min_left = m0;
      *(volatile double *)&dst->far_min = m1;
      __threadfence_system();
      *(volatile unsigned long long *)&dst->seq = c.run_nonce + (unsigned long long)win_ix + 1ull;
    } else {
      set_error(c, SIMIAN_ERR_BAD_ARGUMENT); // peers were never connected
    }
  }
}

// One lookahead window (or one redistribution pass) for entity group g, by the
// calling WARP.  Adds the warp's results into the CTA accumulators.
__device__ __forceinline__ void window_warp(const DevCfg &c, const WinState *Wg, WarpShared &ws,
                                            CtaShared &sh, uint32_t g, uint32_t mbar_parity) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t slot0 = g << GE_SHIFT;
  // ---- bulk copies that do not depend on the window: the group's entity
  // states and its page cache row (arrival 1 of 2 on the warp's barrier)
  if (lane == 0) {
    mbar_arrive_expect_tx(&ws.mbar, (uint32_t)(GE * sizeof(EntityCore) + SIMIAN_BC_ROW * 4));
    bulk_g2s(ws.core_s, c.core + slot0, (uint32_t)(GE * sizeof(EntityCore)), &ws.mbar);
    bulk_g2s(ws.ax.pgc, c.bcache + (size_t)g * SIMIAN_BC_ROW, SIMIAN_BC_ROW * 4, &ws.mbar);
#ifdef SIMIAN_PROFILE_PHASES
    ws.t_last = clock64();
#endif
    ws.acc_min = SIMIAN_KEY_NONE;
    ws.acc_max = 0;
    ws.acc_committed = ws.acc_emitted = ws.acc_dropped = 0;
    ws.acc_ties = ws.acc_dties = ws.acc_appended = 0;
  }
  // the window descriptor: nine 8-byte words
  if (lane < 9)
    reinterpret_cast<unsigned long long *>(&ws.W)[lane] =
        reinterpret_cast<const unsigned long long *>(Wg)[lane];
  __syncwarp();
  const WinState &W = ws.W;
  unsigned long long *cells = c.cell_state + (size_t)g * c.rows;
  PHASE(0);

  if (W.redist) {
    if (lane == 0) {
      mbar_arrive(&ws.mbar);
      ws.ax.far_appends = ws.ax.page_allocs = ws.ax.pgc_take = ws.ax.pgc_have = 0;
    }
    __syncwarp();
    uint32_t spins = 0;
    while (!mbar_try_wait(&ws.mbar, mbar_parity))
      if (++spins > (1u << 24)) {
        set_error(c, SIMIAN_ERR_STALLED);
        break;
      }
    // (the page cache is not used by a redistribution pass: pages come from the ring)
    redistribute_group(c, ws, g);
  } else {
    // ---- K2a: this lane's cell of the window: row tb_lo + lane.  Row tb_hi may
    // be appended to during this launch (see SIMIAN_TIME_DEAD); the rows below
    // it are recycled after this window.
    const int ncell = (int)(W.tb_hi - W.tb_lo + 1);
    unsigned long long hw = CELL_EMPTY;
    uint32_t free_flag = 0;
    if ((int)lane < ncell) {
      const long long tb = W.tb_lo + lane;
      hw = ld_cg64(cells + (uint32_t)(tb & (long long)c.ring_mask));
      free_flag = tb == W.tb_hi ? 0u : PG_FREE_FLAG;
    }
    Cursor cur = cursor_of_head(hw);
    uint32_t npg, nvalid;
    const bool complete = list_chunk(c, ws, cur, free_flag, npg, nvalid);
    PHASE(1);
    const bool fast = complete && nvalid <= (uint32_t)NSLOT;
    // ---- K2: one bulk copy per listed page into the slots (arrival 2 of 2)
    if (lane == 0) {
      if (fast && nvalid)
        mbar_arrive_expect_tx(&ws.mbar, nvalid * 32u);
      else
        mbar_arrive(&ws.mbar);
    }
    __syncwarp();
    if (fast)
      for (uint32_t i = lane; i < npg; i += 32) {
        const uint32_t n = PG_COUNT(ws.pg_cnt[i]);
        if (n)
          bulk_g2s(&ws.slots[ws.pg_off[i]], c.pool + (size_t)ws.pg_id[i] * PAGE, n * 32u,
                   &ws.mbar);
      }
    {
      uint32_t spins = 0;
      while (!mbar_try_wait(&ws.mbar, mbar_parity))
        if (++spins > (1u << 24)) { // never legitimately: fail, do not hang
          set_error(c, SIMIAN_ERR_STALLED);
          break;
        }
    }
    if (lane == 0) {
      ws.ax.far_appends = ws.ax.page_allocs = ws.ax.pgc_take = 0;
      ws.ax.pgc_have = ws.ax.pgc[SIMIAN_BC];
    }
    __syncwarp();
    PHASE(2);
    if (nvalid == 0 && complete) {
      // nothing in the window's rows for this group
    } else if (fast) {
      warp_pass(c, ws, slot0, nvalid, 0, GE, true);
    } else {
      slow_passes<false>(c, ws, slot0, hw, free_flag, complete, npg, 0u);
    }
    __syncwarp();
    // ---- recycle: the pages of rows [tb_lo, tb_hi) go onto the group's own
    // page stack in HBM (what does not fit: to the global ring), heads are
    // cleared.  The entries below the ones popped in this launch stay where they
    // are, the recycled pages are written above them.
    {
      uint32_t *stack = c.bcache + (size_t)g * SIMIAN_BC_ROW;
      const uint32_t have = ws.ax.pgc_have, take = ws.ax.pgc_take;
      uint32_t height = have - (take < have ? take : have);
      if (complete) {
        // (a slow window whose list was complete still holds it: npg of the first listing)
        recycle_chunk(c, ws, stack, npg, height);
      } else {
        Cursor k2 = cursor_of_head(hw);
        bool done;
#pragma unroll 1
        do {
          uint32_t n2, nv2;
          done = list_chunk(c, ws, k2, free_flag, n2, nv2);
          recycle_chunk(c, ws, stack, n2, height);
        } while (!done);
      }
      if ((int)lane < ncell - 1) {
        const long long tb = W.tb_lo + lane;
        cells[(uint32_t)(tb & (long long)c.ring_mask)] = CELL_EMPTY;
      }
      // rows the window jumped over (dead records only)
      long long nskip = W.tb_lo - W.tb_clean;
      if (nskip > (long long)c.ring) nskip = c.ring;
      for (long long i = lane; i < nskip; i += 32) {
        const long long tb = W.tb_clean + i;
        free_cell(c, cells + (uint32_t)(tb & (long long)c.ring_mask), false);
      }
      // low stack: take pages from the ring for the next window (one atomic)
      if (height < (uint32_t)SIMIAN_BC_LOW) {
        const uint32_t need = (uint32_t)SIMIAN_BC_FILL - height;
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(&c.acc->alloc_head, (unsigned long long)need);
        base = __shfl_sync(FULL, base, 0);
        // ring slots beyond alloc_limit are not valid yet: take only what is
        const unsigned long long lim = W.alloc_limit;
        const uint32_t okn =
            base >= lim ? 0u : (lim - base < need ? (uint32_t)(lim - base) : need);
        if (lane < okn) {
          const uint32_t p = c.free_ring[(base + lane) % c.n_pages];
          if (base + lane < c.n_pages) scrub_page(c, p); // first lap of the ring
          stack[height + lane] = p;
        }
        if (lane == 0 && okn < need)
          atomicAdd(&c.acc->leaked_pages, (unsigned long long)(need - okn));
        height += okn;
      }
      if (lane == 0) stack[SIMIAN_BC] = height;
    }
    PHASE(7);
  }

  // ---- K1: the warp's results -> the CTA's accumulators
  __syncwarp();
  if (lane == 0) {
    if (ws.acc_min != SIMIAN_KEY_NONE) atomicMin(&sh.min_key, ws.acc_min);
    if (ws.acc_max) atomicMax(&sh.max_key, ws.acc_max);
    if (ws.acc_committed) atomicAdd(&sh.committed, ws.acc_committed);
    if (ws.acc_emitted) atomicAdd(&sh.emitted, ws.acc_emitted);
    if (ws.acc_dropped) atomicAdd(&sh.dropped, ws.acc_dropped);
    if (ws.acc_ties) atomicAdd(&sh.ties, ws.acc_ties);
    if (ws.acc_dties) atomicAdd(&sh.dties, ws.acc_dties);
    if (ws.acc_appended) atomicAdd(&sh.appended, ws.acc_appended);
    if (ws.ax.far_appends) atomicAdd(&sh.far_appends, ws.ax.far_appends);
    if (ws.ax.page_allocs) atomicAdd(&sh.page_allocs, ws.ax.page_allocs);
  }
  PHASE(8);
}

__device__ __forceinline__ void cta_acc_reset(CtaShared &sh) {
  sh.min_key = SIMIAN_KEY_NONE;
  sh.max_key = 0;
  sh.committed = sh.emitted = sh.dropped = sh.ties = sh.dties = sh.appended = 0;
  sh.far_appends = sh.page_allocs = 0;
  sh.is_last = 0;
}

// the CTA's accumulators -> its BlockPart slot: one thread per field,
// independent, result-less atomics
__device__ __forceinline__ void cta_acc_flush(const DevCfg &c, CtaShared &sh, uint32_t tid) {
  if (tid < 10) {
    BlockPart *p = &c.part[blockIdx.x % SIMIAN_PART_SLOTS];
    switch (tid) {
      case 0: if (sh.min_key != SIMIAN_KEY_NONE) atomicMin(&p->min_key, sh.min_key); break;
      case 1: if (sh.max_key) atomicMax(&p->max_key, sh.max_key); break;
      case 2: if (sh.committed) atomicAdd(&p->committed, sh.committed); break;
      case 3: if (sh.emitted) atomicAdd(&p->emitted, sh.emitted); break;
      case 4: if (sh.dropped) atomicAdd(&p->dropped, sh.dropped); break;
      case 5: if (sh.ties) atomicAdd(&p->ties, sh.ties); break;
      case 6: if (sh.dties) atomicAdd(&p->dties, sh.dties); break;
      case 7: if (sh.appended) atomicAdd(&p->appended, sh.appended); break;
      case 8: if (sh.far_appends) atomicAdd(&p->far_appends, sh.far_appends); break;
      default: if (sh.page_allocs) atomicAdd(&p->page_allocs, sh.page_allocs); break;
    }
  }
}

// the step after the last CTA of a window: this rank's next window (size == 1)
__device__ void advance_single(const DevCfg &c, const WinState &W, CtaShared &sh, uint32_t win_ix) {
  reduce_parts(c, sh);
  WinState *Wn = &c.win[(win_ix + 1u) & 1u];
  if (W.redist) {
    setup_next_window(c, W, Wn, W.gmin, 0.0, true, sh.nw);
  } else {
    unsigned long long cand, far;
    local_min_candidate(c, W, sh, cand, far);
    double gmin = cand == SIMIAN_KEY_NONE ? c.inf_time : simian_key_time(cand); // simian.js:147
    double farT = far == SIMIAN_KEY_NONE ? __longlong_as_double(0x7FF0000000000000LL)
                                         : simian_key_time(far);
    setup_next_window(c, W, Wn, gmin, farT, false, sh.nw);
  }
}

// One lookahead window (or one redistribution pass): warp w of CTA b owns
// entity group b * wpc + w (wpc = blockDim.x / 32).
// fuse_advance: 1 (size == 1): the last CTA to finish computes window k+1;
// 2 / 3 (size > 1): the last CTA computes this rank's LBTS candidate (3: and
// publishes it to the peers' exchange slots); 0: neither.
__global__ void __launch_bounds__(SIMIAN_THREADS, SIMIAN_MINBLOCKS)
    k_window(const __grid_constant__ DevCfg c, uint32_t win_ix, int fuse_advance) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5, wpc = blockDim.x >> 5;
  WarpShared &ws = reinterpret_cast<WarpShared *>(smem_raw)[warp];
  CtaShared &sh = *reinterpret_cast<CtaShared *>(smem_raw + (size_t)wpc * sizeof(WarpShared));
  // the warp's copy barrier: two arrivals per window (state + cache, records)
  if (lane == 0) mbar_init(&ws.mbar, 2u);
  if (tid == 0) cta_acc_reset(sh);
  // Programmatic dependent launch: every resident CTA of the previous window
  // has started, so this grid's CTAs may take the SM slots that launch's last
  // wave frees -- and wait here until it has completed and its writes are
  // visible.  (Both calls are no-ops for an ordinary launch.)
  cudaTriggerProgrammaticLaunchCompletion();
  cudaGridDependencySynchronize();
  const WinState *Wg = &c.win[win_ix & 1u];
  if (Wg->done) return; // simian.js:119: enqueued past the end of the run
  if (tid == 0) sh.found = ld_cg(&c.acc->error);
  __syncthreads();
  if (sh.found) return; // sticky error: CTA-uniform exit
  const uint32_t g = blockIdx.x * wpc + warp;
  if (g < c.n_groups) window_warp(c, Wg, ws, sh, g, 0u);
  __syncthreads();
  cta_acc_flush(c, sh, tid);
  if (!fuse_advance) return;
  // ---- last CTA computes the next window (simian.js:147)
  __threadfence();
  __syncthreads();
  if (tid == 0) sh.is_last = (atomicAdd(&c.acc->ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!sh.is_last) return;
  __threadfence();
  const WinState W = *Wg;
  if (fuse_advance >= 2) {
    lbts_candidate(c, W, sh, win_ix, fuse_advance == 3);
    return;
  }
  advance_single(c, W, sh, win_ix);
}

__global__ void __launch_bounds__(SIMIAN_THREADS) k_local_min(const __grid_constant__ DevCfg c, uint32_t win_ix,
                                                            int publish) {
  __shared__ CtaShared sh;
  const WinState W = c.win[win_ix & 1u];
  lbts_candidate(c, W, sh, win_ix, publish);
}

// size > 1: consume the allreduced {globalMinLeft, min far} (simian.js:144)
__global__ void __launch_bounds__(SIMIAN_THREADS) k_advance(const __grid_constant__ DevCfg c, uint32_t win_ix,
                                                          int from_xch) {
  __shared__ WinState nw;
  const WinState W = c.win[win_ix & 1u];
  WinState *Wn = &c.win[(win_ix + 1u) & 1u];
  if (W.done) {
    if (threadIdx.x == 0) *Wn = W;
    return;
  }
  double g0 = c.acc->minvec[0], g1 = c.acc->minvec[1];
  if (from_xch) { // allreduce(MIN), receive half: every flag was seen by k_pull
    g0 = g1 = __longlong_as_double(0x7FF0000000000000LL);
    for (int p = 0; p < c.size; p++) {
      const MinSlot *sl = &c.xch[W.parity * (uint32_t)c.size + p];
      double a = *(const volatile double *)&sl->min_left;
      double b2 = *(const volatile double *)&sl->far_min;
      g0 = a < g0 ? a : g0;
      g1 = b2 < g1 ? b2 : g1;
    }
  }
  if (W.redist)
    setup_next_window(c, W, Wn, W.gmin, 0.0, true, nw);
  else
    setup_next_window(c, W, Wn, g0, g1, false, nw);
  // The outbox the NEXT window fills was sent two windows ago; every peer has
  // read it (its pull precedes the allreduce this advance follows).  The one
  // just sent may still be being read: it is left alone.
  if (threadIdx.x < c.size && c.outbox_count)
    c.outbox_count[((W.parity ^ 1u) * (uint32_t)c.size) + threadIdx.x] = 0;
}

// first window: gmin = startTime even when the queue is empty (simian.js:118)
__global__ void __launch_bounds__(SIMIAN_THREADS) k_begin(const __grid_constant__ DevCfg c) {
  WinState W = c.win[0];
  __shared__ WinState w0, nw;
  if (threadIdx.x == 0) {
    w0 = W;
    w0.tb_hi = w0.tb_clean; // nothing recycled yet
  }
  __syncthreads();
  const double inf = __longlong_as_double(0x7FF0000000000000LL);
  unsigned long long far = c.acc->far_min[0];
  // what entity constructors sent to other ranks waits in outbox 0 until the
  // first window's exchange: only this rank's candidate covers it
  const unsigned long long keep_min = c.acc->min_key;
  setup_next_window(c, w0, &c.win[0], c.start, far == SIMIAN_KEY_NONE ? inf : simian_key_time(far),
                    false, nw);
  if (threadIdx.x == 0) {
    c.acc->windows = 0;
    c.acc->window_index = 0;
    c.win[0].parity = 0;
    c.acc->min_key = keep_min;
  }
}

// ------------------------------------------------------------- reductions --

// one field of the entity records of a class, contiguous (result read-back)
__global__ void k_gather_core(const __grid_constant__ DevCfg c, uint32_t slot_base, uint32_t n,
                              int which, unsigned long long *out) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
    EntityCore st = c.core[slot_base + j];
    out[j] = which ? st.hash : st.count;
  }
}

// digest = sum of simian_digest_term over local entities; state checksum
__global__ void k_digest(const __grid_constant__ DevCfg c, unsigned long long *out /* [2] */) {
  unsigned long long dg = 0, sc = 0;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x; s < c.n_slots; s += stride) {
    int k;
    uint32_t gid = gid_of_slot(c, (uint32_t)s, k);
    EntityCore st = c.core[s];
    dg += simian_digest_term(gid, st.count, st.hash);
    const DevClass &cl = c.cls[k];
    unsigned long long acc = simian_mix64((unsigned long long)gid + 1ull);
    const unsigned long long *w =
        (const unsigned long long *)(cl.state + (size_t)(s - cl.slot_base) * cl.state_bytes);
    for (uint32_t i = 0; i < cl.state_bytes / 8; i++) acc = simian_state_fold(acc, w[i]);
    sc += acc;
  }
  for (int o = 16; o; o >>= 1) {
    dg += __shfl_xor_sync(0xffffffffu, dg, o);
    sc += __shfl_xor_sync(0xffffffffu, sc, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&out[0], dg);
    atomicAdd(&out[1], sc);
  }
}
