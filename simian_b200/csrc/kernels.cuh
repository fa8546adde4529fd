// kernels.cuh -- hand-written sm_100a kernels of the simian-b200 window loop.
//
// One lookahead window of the reference (SimianJS/simian.js:119-148) is ONE
// launch of k_window:
//   K2  coalesced dequeue of every event below the window bound from the
//       block's calendar cells: the records of the window's pages are copied
//       into shared-memory slots with cp.async in one round trip and the due
//       ones linked per entity                      (eventQ.pop, eventQ.js:53-92)
//   K3  handlers of all entities in parallel, each entity's events in
//       (time, handler, src, seq) order               (simian.js:124-133)
//   K4  emit: lock-free paged append of 32-B records into the destination
//       cell, or into the per-peer outbox             (entity.js:47-67, eventQ.push)
//   K1  LBTS: block min-reduction of leftover + emitted times, one atomicMin
//       per CTA, last CTA computes the next window    (simian.js:120,143-147)
// The common window is a straight line of phases (collect_all, link_due,
// process_pass); windows whose records overflow the CTA's 1024 slots (bursts)
// go through ONE out-of-line function (filtered_passes / partition_burst).
// Compiled with -fmad=false: include/simian_spec.h must see no FMA contraction.
#pragma once
#include "device_types.cuh"

#define EPB SIMIAN_EPB
#define NTHREADS SIMIAN_THREADS
#define PAGE SIMIAN_PAGE
#define NREFS SIMIAN_REFS
#define PGMAX SIMIAN_PGMAX
#define PG_FREE_FLAG 0x8000u

// per-phase SM-cycle accounting of k_window (diagnostic builds only)
#ifdef SIMIAN_PROFILE_PHASES
#define PHASE(k)                                                                 \
  do {                                                                           \
    if (threadIdx.x == 0) {                                                      \
      long long _t = clock64();                                                  \
      atomicAdd(&c.acc->phase_cycles[k], (unsigned long long)(_t - sh.t_last));  \
      sh.t_last = _t;                                                            \
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
// Same record as two 128-bit stores.  Used wherever the store may end up inside a
// function that is NOT inlined into a kernel body (__noinline__ functions and
// everything they inline): ptxas 12.9 can turn the 256-bit form into a single
// 64-bit store there (seen in SASS: STG.E.64 for a st.global.v4.u64 of the PTX),
// silently dropping 24 of the 32 bytes -- and whether it does depends on unrelated
// code (round 2: adding a second kernel that shares the functions broke 25 sites
// that had compiled correctly).  Hence the WIDE template flag below: 256-bit stores
// only on call chains that are force-inlined into a __global__ function.
__device__ __forceinline__ void st_rec_2x16(simian_event_t *p, unsigned long long q0,
                                            unsigned long long q1, unsigned long long q2,
                                            unsigned long long q3) {
  asm volatile("st.global.v2.u64 [%0], {%1,%2};" ::"l"(p), "l"(q0), "l"(q1) : "memory");
  asm volatile("st.global.v2.u64 [%0], {%1,%2};" ::"l"((char *)p + 16), "l"(q2), "l"(q3)
               : "memory");
}
template <bool WIDE>
__device__ __forceinline__ void st_rec_sel(simian_event_t *p, const RecBits &r) {
  if constexpr (WIDE) st_rec(p, r);
  else st_rec_2x16(p, r.q0, r.q1, r.q2, r.q3);
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
// the flags half-word of a record in the pool
__device__ __forceinline__ uint32_t ld_rec_flags(const simian_event_t *p) {
  return (uint32_t)*reinterpret_cast<const unsigned short *>(reinterpret_cast<const char *>(p) + 18);
}
// 16 bytes global -> shared without passing through registers (LDGSTS)
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *src) {
  uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.wait_all;" ::: "memory");
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
  uint32_t mh = __reduce_min_sync(0xffffffffu, hi);
  uint32_t ml = __reduce_min_sync(0xffffffffu, hi == mh ? lo : 0xFFFFFFFFu);
  return ((unsigned long long)mh << 32) | ml;
}
__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
  uint32_t hi = (uint32_t)(v >> 32), lo = (uint32_t)v;
  uint32_t mh = __reduce_max_sync(0xffffffffu, hi);
  uint32_t ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
  return ((unsigned long long)mh << 32) | ml;
}

// ------------------------------------------------------ paged cell append --

struct AppendCtx { // per-CTA, in shared memory
  uint32_t far_appends, page_allocs;
  // the entity block's own free pages (k_window only; pgc_have == 0 elsewhere):
  // a stack, popped from the top with a shared-memory atomic
  uint32_t pgc_take, pgc_have;
  uint32_t pgc[SIMIAN_BC];
  // k_emit: the page stack of an entity block popped IN PLACE (HBM) with an atomic on
  // its height -- several warps of several CTAs may work for the same block there;
  // null elsewhere
  uint32_t *gstack_count;
  const uint32_t *gstack;
};

// head word of an empty cell: no page, "full" so the first append installs one
#define CELL_EMPTY (((unsigned long long)SIMIAN_NONE << 32) | (unsigned long long)PAGE)
// Bit 31 of the head word's count half (and of a page_next entry): the page has
// an OLDER page behind it.  Cells of one page -- almost all of them -- are then
// listed without reading page_next at all; claims on a chained cell miss the
// `count < PAGE` fast path and are serviced by cell_append_slow.
#define CELL_OLDER 0x80000000u
#define CELL_COUNT(lo32) ((uint32_t)(lo32) & 0x7FFFFFFFu)

__device__ __forceinline__ void append_ctx_init(AppendCtx &ax) {
  if (threadIdx.x == 0) {
    ax.far_appends = ax.page_allocs = ax.pgc_take = ax.pgc_have = 0;
    ax.gstack_count = nullptr;
    ax.gstack = nullptr;
  }
}

// Pop a page from the free ring.  Pages pushed during a launch become
// allocatable only after the next advance step raises alloc_limit, so a launch
// never pops a ring slot that is being written.
__device__ __forceinline__ uint32_t page_get(const DevCfg &c, AppendCtx &ax,
                                             unsigned long long alloc_limit) {
  atomicAdd(&ax.page_allocs, 1u);
  if (ax.pgc_have) { // block cache first: a shared-memory atomic, no L2 round trip
    uint32_t k = atomicAdd(&ax.pgc_take, 1u);
    if (k < ax.pgc_have) return ax.pgc[ax.pgc_have - 1u - k];
  }
  if (ax.gstack_count) { // the block's stack in HBM: one atomic on its height
    const int old = atomicSub((int *)ax.gstack_count, 1);
    if (old > 0) return ld_cg(&ax.gstack[old - 1]);
    atomicAdd((int *)ax.gstack_count, 1); // empty: undo, take from the ring
  }
  unsigned long long i = atomicAdd(&c.acc->alloc_head, 1ull);
  if (i >= alloc_limit) {
    set_error(c, SIMIAN_ERR_POOL_EXHAUSTED);
    return SIMIAN_NONE;
  }
  return c.free_ring[i % c.n_pages];
}

__device__ __forceinline__ void page_free(const DevCfg &c, uint32_t p) {
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
  for (int o = 16; o; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
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
    if (i < (uint32_t)PAGE) { // (a chained cell: the fast path does not take these)
      st_rec_2x16(c.pool + ((size_t)p * PAGE + i), q0, q1, q2, q3);
      return;
    }
    if (i == (uint32_t)PAGE) { // first overflow: install the next page
      uint32_t np = page_get(c, ax, alloc_limit);
      if (np == SIMIAN_NONE) return; // sticky error set
      // the link carries the old head's own "older page" bit
      c.page_next[np] = p == SIMIAN_NONE ? SIMIAN_NONE : (p | ((uint32_t)old & CELL_OLDER));
      st_rec_2x16(c.pool + (size_t)np * PAGE, q0, q1, q2, q3); // slot 0 is ours
      // no fence: appenders of this launch only write into the page; chains
      // and records are read after a launch boundary or after the ticket fence
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

template <bool WIDE = false>
__device__ __forceinline__ void cell_append(const DevCfg &c, AppendCtx &ax,
                                            unsigned long long alloc_limit, uint32_t cell,
                                            const RecBits &r) {
  unsigned long long *hp = c.cell_state + cell;
  unsigned long long old = atomicAdd(hp, 1ull);
  if ((uint32_t)old < (uint32_t)PAGE)
    st_rec_sel<WIDE>(c.pool + ((size_t)(uint32_t)(old >> 32) * PAGE + (uint32_t)old), r);
  else
    cell_append_slow(c, ax, alloc_limit, hp, old, r.q0, r.q1, r.q2, r.q3);
}

// calendar cell of a record whose dst is a LOCAL SLOT (SIMIAN_NONE: error set)
__device__ __forceinline__ uint32_t cell_of_local(const DevCfg &c, AppendCtx &ax,
                                                  long long tb_clean, uint32_t far_row,
                                                  double time, uint32_t dst_slot) {
  uint32_t b = (uint32_t)__umul64hi((unsigned long long)dst_slot, c.epb_inv); // dst_slot / epb
  long long tb = tb_of(c, time);
  if (tb < tb_clean) {
    set_error(c, SIMIAN_ERR_OUT_OF_ORDER);
    return SIMIAN_NONE;
  }
  if (tb - tb_clean < (long long)c.ring) return (uint32_t)(tb & (long long)c.ring_mask) * c.n_blocks + b;
  atomicMin(&c.acc->far_min[far_row], simian_time_key(time));
  atomicAdd(&ax.far_appends, 1u);
  return (c.ring + far_row) * c.n_blocks + b;
}

// route a record whose dst is a LOCAL SLOT into the calendar or the far row
template <bool WIDE = false>
__device__ __forceinline__ void append_local(const DevCfg &c, AppendCtx &ax, long long tb_clean,
                                             uint32_t far_row, unsigned long long alloc_limit,
                                             const RecBits &r) {
  uint32_t cell = cell_of_local(c, ax, tb_clean, far_row, r.time(), r.dst());
  if (cell != SIMIAN_NONE) cell_append<WIDE>(c, ax, alloc_limit, cell, r);
}

// ------------------------------------------------------------ init / load --

__global__ void k_init(const __grid_constant__ DevCfg c, uint32_t n_cells) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t j = i; j < c.n_pages; j += stride) {
    c.free_ring[j] = (uint32_t)j;
    c.page_next[j] = SIMIAN_NONE;
  }
  for (size_t j = i; j < n_cells; j += stride) c.cell_state[j] = CELL_EMPTY;
  for (size_t j = i; j < c.n_blocks; j += stride) c.snap[j] = CELL_EMPTY;
  for (size_t j = i; j < SIMIAN_PART_SLOTS; j += stride) {
    BlockPart z;
    z.min_key = SIMIAN_KEY_NONE;
    z.max_key = 0;
    z.committed = z.emitted = z.dropped = z.ties = z.dties = z.appended = z.far_appends =
        z.page_allocs = 0;
    c.part[j] = z;
  }
  // block page caches start with SIMIAN_BC_FILL pages each when the pool allows
  const uint32_t fill =
      ((unsigned long long)c.n_blocks * SIMIAN_BC_FILL * 2ull <= c.n_pages) ? SIMIAN_BC_FILL : 0u;
  for (size_t j = i; j < (size_t)c.n_blocks * SIMIAN_BC; j += stride) {
    uint32_t bb = (uint32_t)(j / SIMIAN_BC), k = (uint32_t)(j % SIMIAN_BC);
    c.bcache[j] = k < fill ? bb * fill + k : SIMIAN_NONE;
    if (k == 0) c.bcount[bb] = fill;
  }
  for (size_t j = i; j < c.n_slots; j += stride) {
    EntityCore z;
    z.count = 0;
    z.hash = 0;
    c.core[j] = z;
  }
  if (i == 0) {
    DevAcc *a = c.acc;
    a->min_key = SIMIAN_KEY_NONE;
    a->far_min[0] = a->far_min[1] = SIMIAN_KEY_NONE;
    a->alloc_head = (unsigned long long)c.n_blocks * fill; // ring slots [0, alloc_head) went to the caches
    a->free_tail = c.n_pages;
    a->committed = a->emitted = a->dropped = a->ties = a->dist_ties = 0;
    a->max_key = 0;
    a->windows = a->redistributions = a->far_appends = a->page_allocs = a->leaked_pages = 0;
    a->appended = 0;
    a->ticket = 0;
    a->error = 0;
    a->window_index = 0;
    a->loop_seq = 0;
    a->job_total[0] = a->job_total[1] = 0;
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
    if (c.outbox_count) c.outbox_count[j] = c.outbox_count[SIMIAN_FAR_COUNT_OFF + j] = 0;
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
    append_local<true>(c, ax, W.tb_clean, W.far_row, W.alloc_limit, r);
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
    append_local<true>(c, ax, W.tb_clean, W.far_row, W.alloc_limit, r);
    if (!(r.flags() & SIMIAN_EVF_CONT)) napp++; // (payload words are never committed)
  }
  appended_flush(c, napp);
  if (track_min) {
    for (int o = 16; o; o >>= 1) {
      unsigned long long x = __shfl_xor_sync(0xffffffffu, mn, o);
      mn = x < mn ? x : mn;
    }
    if ((threadIdx.x & 31) == 0 && mn != SIMIAN_KEY_NONE) atomicMin(&smin, mn);
    __syncthreads();
    if (threadIdx.x == 0 && smin != SIMIAN_KEY_NONE) atomicMin(&c.acc->min_key, smin);
  }
  append_ctx_flush(c, ax);
}

// The record loop of the cross-GPU pull for one CTA: peer number q (of size - 1),
// CTA `sub` of blocks_per_peer.  W: the window the records were sent in; a_*: the
// recycle frontier / far row / page limit the appends obey.  See k_pull for `mode`.
__device__ __forceinline__ void pull_records(const DevCfg &c, const WinState &W,
                                             long long a_tb_clean, uint32_t a_far_row,
                                             unsigned long long a_alloc_limit, AppendCtx &ax,
                                             uint32_t q,
                                             uint32_t sub, uint32_t blocks_per_peer, int mode,
                                             double hi_next, uint32_t &napp,
                                             unsigned long long &mn) {
  const double thr = W.hi + 2.0 * c.min_delay; // the senders' near / far threshold
  const uint32_t p = ((uint32_t)c.rank + 1u + q) % (uint32_t)c.size; // never c.rank
  const uint32_t box = W.parity * (uint32_t)c.size + (uint32_t)c.rank;
  const uint32_t *pc = c.peer_count[p];
  const simian_event_t *src = c.peer_outbox[p];
  if (!pc || !src) {
    set_error(c, SIMIAN_ERR_BAD_ARGUMENT); // peers were never connected
    return;
  }
  uint32_t n_near = *(const volatile uint32_t *)(pc + box);
  uint32_t n_far = *(const volatile uint32_t *)(pc + SIMIAN_FAR_COUNT_OFF + box);
  n_near = n_near < c.outbox_cap ? n_near : c.outbox_cap;
  n_far = n_far < c.outbox_cap - n_near ? n_far : c.outbox_cap - n_near;
  src += (size_t)box * c.outbox_cap;
  const uint32_t stride = blocks_per_peer * blockDim.x;
  // which ends this launch reads: [0, n_near) and / or the n_far records at the back
  const bool do_near = mode != 2;
  const bool do_far = mode == 2 || (mode == 1 && hi_next > thr);
  for (int end = 0; end < 2; end++) {
    if (end == 0 ? !do_near : !do_far) continue;
    const uint32_t n = end == 0 ? n_near : n_far;
    const simian_event_t *base = end == 0 ? src : src + (c.outbox_cap - n);
#ifndef SIMIAN_PULL_ILP
    // four 32-byte NVLink reads in flight per thread, then the four appends one
    // after the other (the variant below, -DSIMIAN_PULL_ILP, keeps the four cell
    // claims in flight as well: measured slower, 23.65 against 23.07 ms per step of
    // the config-3 shape on 2 GPUs)
    for (uint32_t idx = sub * blockDim.x + threadIdx.x; idx < n; idx += 4 * stride) {
      RecBits r[4];
#pragma unroll
      for (int u = 0; u < 4; u++)
        if (idx + u * stride < n) r[u] = ld_rec(base + idx + u * stride);
#pragma unroll
      for (int u = 0; u < 4; u++)
        if (idx + u * stride < n) {
          if (r[u].dst() >= c.n_slots) { // a slot this rank does not have
            set_error(c, SIMIAN_ERR_UNKNOWN_NAME);
            continue;
          }
          if (end == 1) { // far records: below the next bound now, the rest later
            const bool below = r[u].time() < hi_next;
            if (mode == 1 ? !below : below) continue;
          }
          if (mode == 2) {
            unsigned long long k = simian_time_key(r[u].time());
            mn = k < mn ? k : mn;
          }
          append_local<true>(c, ax, a_tb_clean, a_far_row, a_alloc_limit, r[u]);
          if (!(r[u].flags() & SIMIAN_EVF_CONT)) napp++; // (payload words are never committed)
        }
    }
  }
}
#else
    // four 32-byte NVLink reads in flight per thread, then the four cell claims in
    // flight together, then the four stores
    for (uint32_t idx = sub * blockDim.x + threadIdx.x; idx < n; idx += 4 * stride) {
      RecBits r[4];
      uint32_t cell[4];
      unsigned long long old[4];
#pragma unroll
      for (int u = 0; u < 4; u++)
        if (idx + u * stride < n) r[u] = ld_rec(base + idx + u * stride);
#pragma unroll
      for (int u = 0; u < 4; u++) {
        cell[u] = SIMIAN_NONE;
        if (idx + u * stride < n) {
          if (r[u].dst() >= c.n_slots) { // a slot this rank does not have
            set_error(c, SIMIAN_ERR_UNKNOWN_NAME);
            continue;
          }
          if (end == 1) { // far records: below the next bound now, the rest later
            const bool below = r[u].time() < hi_next;
            if (mode == 1 ? !below : below) continue;
          }
          if (mode == 2) {
            unsigned long long k = simian_time_key(r[u].time());
            mn = k < mn ? k : mn;
          }
          cell[u] = cell_of_local(c, ax, a_tb_clean, a_far_row, r[u].time(), r[u].dst());
          if (cell[u] != SIMIAN_NONE) napp++;
        }
      }
#pragma unroll
      for (int u = 0; u < 4; u++)
        if (cell[u] != SIMIAN_NONE) old[u] = atomicAdd(c.cell_state + cell[u], 1ull);
      // (as in the append phase of k_window: every claim that lands in a page or
      // installs one first, the claims that wait for another thread's install last)
      uint32_t waiting = 0;
#pragma unroll
      for (int u = 0; u < 4; u++)
        if (cell[u] != SIMIAN_NONE) {
          if (CELL_COUNT(old[u]) <= (uint32_t)PAGE) {
            if ((uint32_t)old[u] < (uint32_t)PAGE)
              st_rec(c.pool + ((size_t)(uint32_t)(old[u] >> 32) * PAGE + (uint32_t)old[u]), r[u]);
            else
              cell_append_slow(c, ax, a_alloc_limit, c.cell_state + cell[u], old[u], r[u].q0,
                               r[u].q1, r[u].q2, r[u].q3);
          } else {
            waiting |= 1u << u;
          }
        }
      if (waiting) {
#pragma unroll
        for (int u = 0; u < 4; u++)
          if (waiting & (1u << u))
            cell_append_slow(c, ax, a_alloc_limit, c.cell_state + cell[u], old[u], r[u].q0,
                             r[u].q1, r[u].q2, r[u].q3);
      }
    }
  }
}
#endif

// Cross-GPU exchange, receiver side (simian.js:138-142): the records the peers
// staged for this rank in window win_ix are read straight out of THEIR
// outboxes over NVLink (coalesced 32-byte records) and appended to the local
// calendar -- alltoallv and ingest in one kernel, counts never visit the host.
// Runs after the window's allreduce, which orders it behind every peer's
// k_window.  gridDim.x = (size - 1) * blocks_per_peer.
// mode 0: everything the peers staged (front end of the boxes).
// Deferred pull (two-ended boxes, DevCfg::split_outbox):
// mode 1 (before the advance step): the front end -- what may be due in the next
//   window; and, when the next window turns out to reach beyond the senders'
//   near / far threshold (a jump over idle time), the far records below its bound.
// mode 2 (after the advance step, while the next window runs -- normally by the
//   first CTAs of that window's own k_window launch, see there): the back end --
//   far records at or above the next window's bound; their times go into the
//   LBTS candidate of that window.
__global__ void k_pull(const __grid_constant__ DevCfg c, uint32_t win_ix, uint32_t blocks_per_peer,
                       int wait_flags, int mode) {
  __shared__ AppendCtx ax;
  __shared__ double s_hi_next;
  __shared__ unsigned long long smin;
  // (programmatic dependent launch, option "pdl_exchange": no-ops otherwise)
  cudaTriggerProgrammaticLaunchCompletion();
  cudaGridDependencySynchronize();
  append_ctx_init(ax);
  const WinState &W = c.win[win_ix & 1u];
  if (threadIdx.x == 0) smin = SIMIAN_KEY_NONE;
  // The CTAs of peer p wait for p's flag only (mode 0): records of the ranks that
  // finished their window early are pulled while the late ones are still running.
  // Every peer has CTAs here, so when the kernel ends all flags have been seen
  // (k_advance reads every slot).  The deferred modes need every candidate at once.
  const uint32_t my_peer =
      ((uint32_t)c.rank + 1u + blockIdx.x / blocks_per_peer) % (uint32_t)c.size;
  if (wait_flags && !W.done &&
      (mode == 0 ? threadIdx.x == my_peer : threadIdx.x < (unsigned)c.size)) {
    // a rank that has published window win_ix's candidate has finished its
    // k_window: its outboxes are complete (the barrier MPI_Alltoall is, mpi.cpp:422)
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
  // the next window's bound: mode 1 computes it the way k_advance will (minimum
  // over the exchange slots, simian.js:144-147); mode 2 reads what k_advance wrote
  if (threadIdx.x == 0) {
    double hn = __longlong_as_double(0x7FF0000000000000LL);
    if (mode == 1 && !W.done) {
      double g0 = hn;
      for (int p = 0; p < c.size; p++) {
        double a = *(const volatile double *)&c.xch[W.parity * (uint32_t)c.size + p].min_left;
        g0 = a < g0 ? a : g0;
      }
      hn = (W.redist ? W.gmin : g0) + c.min_delay;
    } else if (mode == 2) {
      hn = c.win[(win_ix + 1u) & 1u].hi;
    }
    s_hi_next = hn;
  }
  __syncthreads();
  const double hi_next = s_hi_next;
  // appends of mode 2 run beside the NEXT window: its recycle frontier and limits
  const WinState &Wa = mode == 2 ? c.win[(win_ix + 1u) & 1u] : W;
  uint32_t napp = 0;
  unsigned long long mn = SIMIAN_KEY_NONE;
  if (!W.done)
    pull_records(c, W, Wa.tb_clean, Wa.far_row, Wa.alloc_limit, ax, blockIdx.x / blocks_per_peer,
                 blockIdx.x % blocks_per_peer, blocks_per_peer, mode, hi_next, napp, mn);
  appended_flush(c, napp);
  if (mode == 2) { // pending records the next window's LBTS step must see
    mn = warp_min_u64(mn);
    if ((threadIdx.x & 31) == 0 && mn != SIMIAN_KEY_NONE) atomicMin(&smin, mn);
    __syncthreads();
    if (threadIdx.x == 0 && smin != SIMIAN_KEY_NONE) atomicMin(&c.acc->min_key, smin);
  }
  append_ctx_flush(c, ax);
}

// ---------------------------------------------------------- window kernel --
// (k_construct, which needs the handler context, is defined after it below)

#define LIST_END 0xFFFFu
// stage_n[b] of a block whose window the split kernels leave to k_window_rest (a
// burst that overflows the slots, or a redistribution pass)
#define STAGE_REST 0xFFFFFFFFu
#define EV_PER_THREAD ((NREFS + NTHREADS - 1) / NTHREADS)
static_assert(EV_PER_THREAD <= 8, "ranks are packed 8 bits each into one 64-bit register");
#define RSTACK_MAX 80 // entity sub-ranges a CTA can have pending
#define NBINS 16      // histogram bins over the block's entities (burst partition)

// K2 page copy: 16-byte cp.async per half record (default), or -- build with
// -DSIMIAN_USE_BULK_SCAN -- one cp.async.bulk per page completing on an mbarrier
// (UBLKCP + SYNCS in SASS).  Measured on config 2 (tools/ab_run.py, median of 5,
// twice): steady windows 13.22 ms against 13.25 ms, the t = 0 burst window 1.85 ms
// against 1.75 ms (whole records in shared memory cost the 16-way tie ranking
// two-way bank conflicts): 15.07 ms against 15.00 ms per step.  Not faster: the
// kernel is not bound by the copy (profiles/README.md, r2_e), so the default stays.
#ifdef SIMIAN_USE_BULK_SCAN
#define SIMIAN_BULK_SCAN 1
#endif
struct __align__(32) SlotRec {
  ulonglong2 lo; // {time, dst | src << 32}
  ulonglong2 hi; // {handler | flags << 16 | seq << 32, data}
};
#ifdef SIMIAN_BULK_SCAN
#define SLOT_LO(sh, i) ((sh).rec[i].lo)
#define SLOT_HI(sh, i) ((sh).rec[i].hi)
#else
#define SLOT_LO(sh, i) ((sh).lo[i])
#define SLOT_HI(sh, i) ((sh).hi[i])
#endif

// {commit count, trace hash} of the block's entities with due events: c.epb entries
// behind the WindowShared struct (the block size is a run-time option)
#define SH_CORE(sh) \
  (reinterpret_cast<EntityCore *>(reinterpret_cast<unsigned char *>(&(sh)) + sizeof(WindowShared)))
#define WINDOW_SMEM_BYTES(epb) (sizeof(WindowShared) + (size_t)(epb) * sizeof(EntityCore))

struct __align__(32) WindowShared {
  WinState W;
  AppendCtx ax;
  unsigned long long min_key, max_key, free_base, ring_base;
  uint32_t committed, emitted, dropped, ties, dties, appended;
  uint32_t npg, nvalid, nfree, free_pos, total, is_last, found, stack_n, pass_lo, pass_hi, pass_pg_lo,
      pass_pg_hi;
  long long scan_j, t_last;
  // entity sub-ranges still to process: (lo, hi, first page-list index, end page-list index)
  uint32_t rstack[4 * RSTACK_MAX];
  // burst partition (a window whose due events overflow the slots): due records
  // per 1/16th of the block's entities, the parts they are grouped into, and
  // the scratch pages (listed behind the cells' pages in pg_id) of each part
  uint32_t hist[NBINS], pfill[NBINS], pcnt[NBINS], plo[NBINS], phi[NBINS], ppage0[NBINS + 1];
  uint32_t part_of_bin[NBINS];
  uint32_t nparts, bin_shift;
  // due continuation records of the block (payload words, SIMIAN_EVF_CONT): one list
  // through nxt[], searched by payload(k); they are in no entity's list
  uint32_t cont_head;
  // split sums: jobs the handlers of this pass parked, the range of c.jobs reserved for
  // them (SIMIAN_NONE: no room, computed here) and the descriptors written so far
  uint32_t pass_jobs, job_base, job_count;
  uint32_t head[SIMIAN_EPB_MAX];   // per local entity: index of its first due event
  uint32_t pg_id[PGMAX];           // pages of this block's cells in the window rows
  uint32_t pg_off[PGMAX];          // records in the pages listed before this one
  uint16_t pg_cnt[PGMAX];
  uint16_t nxt[NREFS];          // next due event of the same entity
  unsigned long long xs[NREFS]; // trace words, stored in commit order along the entity list
  // (the entities' {commit count, trace hash}, 16 bytes x entities per block, follow
  // the struct in dynamic shared memory: SH_CORE)
  // due events: the 32-byte records themselves.  The thread-per-event path
  // overwrites slot i with the record event i emits (staged, then appended).
#ifdef SIMIAN_BULK_SCAN
  // whole records, so that a page is ONE bulk copy (cp.async.bulk) into the slots
  SlotRec rec[NREFS];
  unsigned long long mbar; // completion barrier of the bulk copies
#else
  // as two 16-byte halves (no bank conflicts); filled with 16-byte cp.async
  ulonglong2 lo[NREFS]; // {time, dst | src << 32}
  ulonglong2 hi[NREFS]; // {handler | flags << 16 | seq << 32, data}
#endif
};

template <class SH>
__device__ __forceinline__ RecBits slot_get(const SH &sh, uint32_t i) {
  ulonglong2 a = SLOT_LO(sh, i), b = SLOT_HI(sh, i);
  RecBits r;
  r.q0 = a.x;
  r.q1 = a.y;
  r.q2 = b.x;
  r.q3 = b.y;
  return r;
}
template <class SH>
__device__ __forceinline__ void slot_put(SH &sh, uint32_t i, const RecBits &r) {
  SLOT_LO(sh, i) = make_ulonglong2(r.q0, r.q1);
  SLOT_HI(sh, i) = make_ulonglong2(r.q2, r.q3);
}

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

// what a handler reaches through `self` on the device (see simian_models.h)
// SEQ: one thread per entity (stateful handlers, same-window self-sends);
// !SEQ: one thread per event (pure handlers).
// WIDE: records are written with one 256-bit store (only where the context lives in
// code that is force-inlined into a kernel body -- see st_rec_2x16).
template <bool SEQ, bool WIDE = false>
struct DevCtx {
  const DevCfg &c;
  const WinState &W; // window bound / recycle frontier / far row in force
  AppendCtx &ax;
  ThreadAcc &ta;
  const DevClass *cl;
  uint32_t self_gid, self_slot;
  unsigned long long commit_idx, key;
  uint32_t n_emit;
  double now_;
  uint32_t handler_, src_;
  unsigned long long data_;
  // payload of the event being committed (SIMIAN_EVF_BLOB): its flags and seq, and
  // the CTA's slots, where the continuation records of the window are listed
  uint32_t flags_, seq_;
  const WindowShared *pl_sh;
  // !SEQ: where the first locally routed record is staged (the two halves of a
  // shared-memory slot) instead of being appended right away; null = append directly
  ulonglong2 *stage_lo, *stage_hi;
  // fold the times of LOCALLY routed records into the LBTS candidate (window
  // kernel: yes; constructors: no -- the calendar search finds those, and a
  // candidate that a later window commits would move the window bounds)
  bool local_min;
  bool split_ok; // remote records may be classified near / far (not in constructors)
  // SEQ: where a handler may park a deferred warp-wide sum (simian_defer_weighted_sum):
  // the shared-memory slot of the calendar event being committed (NONE: a pending
  // self-send, or not inside the window kernel) and the entity's count at window start
  WindowShared *job_sh;
  uint32_t job_slot;
  unsigned long long job_count0;
  // same-window self-sends (rx omitted, offset < minDelay allowed: entity.js:41)
  int npend;
  double pend_time[SEQ ? SIMIAN_SELF_MAX : 1];
  unsigned long long pend_q2[SEQ ? SIMIAN_SELF_MAX : 1], pend_data[SEQ ? SIMIAN_SELF_MAX : 1];

  __device__ __forceinline__ DevCtx(const DevCfg &c_, const WinState &W_, AppendCtx &ax_,
                                    ThreadAcc &ta_)
      : c(c_), W(W_), ax(ax_), ta(ta_), flags_(0), seq_(0), pl_sh(nullptr), stage_lo(nullptr),
        stage_hi(nullptr), local_min(true), split_ok(true), job_sh(nullptr),
        job_slot(SIMIAN_NONE), job_count0(0), npend(0) {}

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
      ta.min_key = tk < ta.min_key ? tk : ta.min_key;
    }
    ta.appended++;
    if (!SEQ && stage_lo) { // K4 is done by the append phase of k_window / k_emit
      *stage_lo = make_ulonglong2(r.q0, r.q1);
      *stage_hi = make_ulonglong2(r.q2, r.q3);
      stage_lo = nullptr;
    } else {
      append_local<WIDE>(c, ax, W.tb_clean, W.far_row, W.alloc_limit, r);
    }
  }

  // ---- payloads beyond the 8-byte data word (entity.js:52-60) ----
  __device__ __forceinline__ uint32_t payload_words() const {
    return (flags_ & SIMIAN_EVF_BLOB) ? (uint32_t)data_ : 0u;
  }
  // word k of the payload of the event being committed: its continuation record is
  // due in the same window and sits in the CTA's slots
  __device__ __forceinline__ unsigned long long payload(uint32_t k) {
    if (pl_sh && k < payload_words()) {
      const unsigned long long want1 = (unsigned long long)self_slot | ((unsigned long long)src_ << 32);
      const unsigned long long want2 = (unsigned long long)(handler_ & 0xFFFFu) |
                                       ((unsigned long long)(SIMIAN_EVF_CONT | k) << 16) |
                                       ((unsigned long long)seq_ << 32);
      const unsigned long long t = (unsigned long long)__double_as_longlong(now_);
      for (uint32_t j = pl_sh->cont_head; j != LIST_END; j = pl_sh->nxt[j]) {
        const ulonglong2 a = SLOT_LO(*pl_sh, j);
        if (a.x != t || a.y != want1) continue;
        const ulonglong2 b = SLOT_HI(*pl_sh, j);
        if (b.x == want2) return b.y;
      }
    }
    set_error(c, SIMIAN_ERR_PAYLOAD);
    return 0ull;
  }

  // a record for another rank: staged for the per-window exchange (entity.js:66)
  __device__ __forceinline__ void put_remote(int rank, double time, const RecBits &r) {
    // warp-aggregated claim: one atomic per (warp, destination rank)
    const uint32_t box = W.parity * (uint32_t)c.size + (uint32_t)rank;
    // Two-ended outbox (c.split_outbox, the deferred pull): records that cannot
    // be due in the receiver's next window -- time >= this window's bound + 2
    // minDelay -- are staged from the BACK of the box and pulled while that next
    // window runs; the others from the front, pulled before it.
    const uint32_t far = (c.split_outbox && split_ok &&
                          time >= W.hi + 2.0 * c.min_delay) ? 1u : 0u;
    const int key = rank * 2 + (int)far;
    unsigned act = __activemask();
    unsigned peers = __match_any_sync(act, key);
    int leader = __ffs(peers) - 1, ln = threadIdx.x & 31;
    uint32_t i = 0;
    if (ln == leader)
      i = atomicAdd(&c.outbox_count[far * SIMIAN_FAR_COUNT_OFF + box], (uint32_t)__popc(peers));
    i = __shfl_sync(peers, i, leader) + __popc(peers & ((1u << ln) - 1u));
    // (the two ends meet: the other end's count is only a lower bound here, the
    // receiver checks the sum again)
    if (i >= c.outbox_cap ||
        i + ld_cg(&c.outbox_count[(far ^ 1u) * SIMIAN_FAR_COUNT_OFF + box]) >= c.outbox_cap) {
      set_error(c, SIMIAN_ERR_OUTBOX_FULL);
      return;
    }
    simian_event_t *dst_rec =
        c.outbox + ((size_t)box * c.outbox_cap + (far ? c.outbox_cap - 1u - i : i));
    st_rec_sel<WIDE>(dst_rec, r);
  }

  // Entity.reqService (entity.js:35-68)
  struct NoWords {
    __device__ __forceinline__ unsigned long long operator()(uint32_t) const { return 0ull; }
  };
  struct WordsAt {
    const unsigned long long *p;
    __device__ __forceinline__ unsigned long long operator()(uint32_t k) const { return p[k]; }
  };
  __device__ __forceinline__ void req_service(double offset, uint16_t service,
                                              unsigned long long data, uint32_t rx) {
    send_<false>(offset, service, data, rx, NoWords(), 0u);
  }
  // ... with `data` = nwords 64-bit words (any JS object of the reference, flattened)
  __device__ __forceinline__ void req_service_blob(double offset, uint16_t service,
                                                   const uint64_t *words_, uint32_t nwords,
                                                   uint32_t rx) {
    WordsAt src{reinterpret_cast<const unsigned long long *>(words_)};
    req_service_blob_fn(offset, service, nwords, rx, src);
  }
  // ... the words produced on demand: word(k) -> uint64 (no staging array)
  template <class F>
  __device__ __forceinline__ void req_service_blob_fn(double offset, uint16_t service,
                                                      uint32_t nwords, uint32_t rx, F word) {
    if (!c.blobs || nwords > SIMIAN_BLOB_MAX_WORDS) { // option "payloads" not set
      set_error(c, SIMIAN_ERR_PAYLOAD);
      return;
    }
    send_<true>(offset, service, (unsigned long long)nwords, rx, word, nwords);
  }

  template <bool BLOB, class F>
  __device__ __forceinline__ void send_(double offset, uint16_t service, unsigned long long data,
                                        uint32_t rx, F words, uint32_t nwords) {
    uint32_t k = n_emit++;
    const uint32_t fl = BLOB ? SIMIAN_EVF_BLOB : 0u;
    if (rx != SIMIAN_NULL_ENTITY && offset < c.min_delay) { // entity.js:41-45
      set_error(c, SIMIAN_ERR_LOOKAHEAD);
      return;
    }
    double time = now_ + offset; // entity.js:47
    if (time > c.end) {          // entity.js:48
      ta.dropped++;
      return;
    }
    uint32_t seq = simian_emit_seq(commit_idx, k);
    ta.emitted++;
    if (rx == SIMIAN_NULL_ENTITY) { // entity.js:50-51
      if (time < now_) {
        set_error(c, SIMIAN_ERR_OUT_OF_ORDER); // simian.js:124-125 at pop time
        return;
      }
      if (time < W.hi) { // lands inside this window: stays with the entity
        if (BLOB) { // (a pending self-send carries no payload)
          set_error(c, SIMIAN_ERR_PAYLOAD);
          return;
        }
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
      put_local(make_rec(time, self_slot, self_gid, service, fl, seq, data));
      if (BLOB)
        for (uint32_t w = 0; w < nwords; w++) // the payload words: never committed, not counted
          append_local<WIDE>(c, ax, W.tb_clean, W.far_row, W.alloc_limit,
                             make_rec(time, self_slot, self_gid, service, SIMIAN_EVF_CONT | w, seq,
                                      words(w)));
      return;
    }
    if (rx >= c.n_ids) { // the reference throws at dispatch (simian.js:128)
      set_error(c, SIMIAN_ERR_UNKNOWN_NAME);
      return;
    }
    int rank;
    uint32_t slot;
    route_gid(c, rx, rank, slot); // getOffsetRank, entity.js:62
    RecBits r = make_rec(time, slot, self_gid, service, fl, seq, data);
    if (rank == c.rank) { // entity.js:64
      put_local(r);
      if (BLOB)
        for (uint32_t w = 0; w < nwords; w++)
          append_local<WIDE>(c, ax, W.tb_clean, W.far_row, W.alloc_limit,
                             make_rec(time, slot, self_gid, service, SIMIAN_EVF_CONT | w, seq,
                                      words(w)));
    } else { // entity.js:66: staged for the per-window exchange
      // the receiver has not seen it when the LBTS candidates are reduced:
      // the sender's candidate covers it (simian.js:143-144)
      unsigned long long tk = simian_time_key(time);
      ta.min_key = tk < ta.min_key ? tk : ta.min_key;
      put_remote(rank, time, r);
      if (BLOB)
        for (uint32_t w = 0; w < nwords; w++)
          put_remote(rank, time,
                     make_rec(time, slot, self_gid, service, SIMIAN_EVF_CONT | w, seq, words(w)));
    }
  }
};

// Dispatch: generated from include/simian_handlers.def (attachService binds a
// service name to one of these by name; simian.js:129-131 calls it).
// The pure handlers only (thread-per-event path: simian_fn_is_pure).
template <class Ctx>
__device__ __forceinline__ void run_pure_handler(Ctx &ctx, uint32_t fn) {
  switch (fn) {
#define SIMIAN_HANDLER(id, name, func, pure, ptype) \
  case id:                                          \
    if constexpr (pure) func(ctx);                  \
    else set_error(ctx.c, SIMIAN_ERR_UNKNOWN_NAME); \
    break;
#include "../../include/simian_handlers.def"
#undef SIMIAN_HANDLER
    default:
      set_error(ctx.c, SIMIAN_ERR_UNKNOWN_NAME);
  }
}

template <class Ctx>
__device__ __forceinline__ void run_handler(Ctx &ctx, uint32_t fn) {
  switch (fn) {
#define SIMIAN_HANDLER(id, name, func, pure, ptype) \
  case id:                                          \
    func(ctx);                                      \
    break;
#include "../../include/simian_handlers.def"
#undef SIMIAN_HANDLER
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
    DevCtx<false, true> ctx(c, W, ax, ta);
    ctx.local_min = false;
    ctx.split_ok = false; // no window is open: everything is pulled before the first one
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
  }
  if (ta.emitted) atomicAdd(&c.acc->emitted, (unsigned long long)ta.emitted);
  if (ta.dropped) atomicAdd(&c.acc->dropped, (unsigned long long)ta.dropped);
  if (ta.min_key != SIMIAN_KEY_NONE) atomicMin(&c.acc->min_key, ta.min_key);
  appended_flush(c, ta.appended);
  append_ctx_flush(c, ax);
}

// key order (time, handler, src, seq), then the slot index so that identical
// records are all taken, in a fixed order: a < b
__device__ __forceinline__ bool key_lt(unsigned long long a1, unsigned long long a2, uint32_t a3,
                                       uint32_t a4, unsigned long long b1,
                                       unsigned long long b2, uint32_t b3, uint32_t b4) {
  if (a1 != b1) return a1 < b1;
  if (a2 != b2) return a2 < b2;
  if (a3 != b3) return a3 < b3;
  return a4 < b4;
}

// ---- deferred warp-wide sums (the LANL receive handler's weighted subset sum) ----
// The per-entity path runs one thread per entity; a handler's long sum would make
// one lane work while 31 wait.  simian_defer_weighted_sum parks the sum as a JOB on
// the consumed event's shared-memory slot -- xs[slot] = na | nj << 20 | di0 << 40 |
// commit delta << 45 | sink offset << 56, the record's data word = the list
// pointer -- and after the entities' events of the pass a WARP computes each job in
// the order simian_lanl_weighted_sum defines and adds it to the sink, an entity's
// jobs in commit order.  (Every event of an entity in a pass has a commit delta
// below NREFS = 1024, so a model's jobs are either all parked or all computed in
// line: the additions into the sink keep their commit order.)
#define JOB_NA(x) ((uint32_t)(x) & 0xFFFFFu)
#define JOB_NJ(x) ((uint32_t)((x) >> 20) & 0xFFFFFu)
#define JOB_DI0(x) ((uint32_t)((x) >> 40) & 0x1Fu)
#define JOB_DELTA(x) ((uint32_t)((x) >> 45) & 0x7FFu)
#define JOB_SINK_OFF(x) ((int)(signed char)((x) >> 56))
static_assert(NREFS <= 2048, "commit delta of a parked job: 11 bits");

__device__ __forceinline__ bool simian_defer_weighted_sum(DevCtx<true> &ctx, const double *list,
                                                          uint32_t na, uint32_t nj, uint32_t di0,
                                                          double *sink) {
  if (!ctx.job_sh || ctx.job_slot == SIMIAN_NONE) return false;
  const unsigned long long delta = ctx.commit_idx - ctx.job_count0;
  const long long off = sink - list; // in doubles
  if (na == 0 || na >= (1u << 20) || nj >= (1u << 20) || di0 >= 32u || delta >= 2048ull ||
      off < -128 || off > 127 || (unsigned long long)na * nj >= (1ull << 32))
    return false;
  WindowShared &sh = *ctx.job_sh;
  SLOT_HI(sh, ctx.job_slot).y = (unsigned long long)list; // (the event's data has been read)
  sh.xs[ctx.job_slot] = (unsigned long long)na | ((unsigned long long)nj << 20) |
                        ((unsigned long long)di0 << 40) | (delta << 45) |
                        ((unsigned long long)(unsigned char)(signed char)off << 56);
  if (ctx.c.jobs) atomicAdd(&sh.pass_jobs, 1u);
  return true;
}

#ifndef SIMIAN_WSUM_UNROLL
#define SIMIAN_WSUM_UNROLL 1
#endif
// one job, all 32 lanes: partial sums by lane, then the butterfly (the defined order)
__device__ __forceinline__ double warp_weighted_sum(unsigned long long key, const double *list,
                                                    uint32_t na, uint32_t nj, uint32_t di0) {
  const uint32_t lane = threadIdx.x & 31u, total = na * nj;
  double p = 0.0;
  // list index i = k / nj, advanced incrementally (k grows by 32 per step).  The
  // loop is bound by its arithmetic (a splitmix64 draw + u64 -> f64 per term), not
  // by the list loads: loading eight elements ahead made it slower (747 -> 899 ms
  // per config-4 run).
  const uint32_t qi = 32u / nj, ri = 32u % nj;
  uint32_t i = lane / nj, j = lane % nj;
  constexpr int wsum_unroll = SIMIAN_WSUM_UNROLL;
#pragma unroll wsum_unroll
  for (uint32_t k = lane; k < total; k += 32u) {
    p += list[i] * simian_u01(simian_rng_draw(key, di0 + k));
    i += qi;
    j += ri;
    if (j >= nj) j -= nj, i++;
  }
  for (int o = 16; o; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
  return p;
}

// All due events of one entity, in order (simian.js:122-134 restricted to one
// entity; entities never interact inside a window because cross-entity sends
// carry offset >= minDelay, entity.js:41).  The records are in shared memory.
__device__ __noinline__ void process_entity(const DevCfg &c, WindowShared &sh, ThreadAcc &ta,
                                            uint32_t slot, uint32_t first) {
  uint32_t n = 0;
  for (uint32_t j = first; j != LIST_END; j = sh.nxt[j]) n++;
  int k;
  DevCtx<true> ctx(c, sh.W, sh.ax, ta);
  ctx.self_slot = slot;
  ctx.self_gid = gid_of_slot(c, slot, k);
  ctx.cl = &c.cls[k];
  EntityCore st = SH_CORE(sh)[slot % c.epb]; // (local entity: slot - block * epb)
  ctx.job_sh = &sh;
  ctx.job_count0 = st.count;
  unsigned long long p1 = 0, p2 = 0;
  uint32_t p3 = 0, p4 = 0;
  bool have_prev = false, have_last = false;
  double last_time = 0;
  unsigned long long last_q = 0, last_data = 0;
  uint32_t taken = 0;
#pragma unroll 1
  for (;;) {
    // next event of the static group: min key strictly above the previous one
    bool have_g = false;
    uint32_t gj = 0;
    unsigned long long g1 = 0, g2 = 0;
    uint32_t g3 = 0;
    if (taken < n) {
#pragma unroll 1
      for (uint32_t j = first; j != LIST_END; j = sh.nxt[j]) {
        const RecBits r = slot_get(sh, j);
        unsigned long long k1 = simian_time_key(r.time());
        unsigned long long k2 = ((unsigned long long)r.handler() << 32) | r.src();
        uint32_t k3 = r.seq();
        if (have_prev && !key_lt(p1, p2, p3, p4, k1, k2, k3, j)) continue;
        if (!have_g || key_lt(k1, k2, k3, j, g1, g2, g3, gj)) {
          gj = j;
          g1 = k1;
          g2 = k2;
          g3 = k3;
          have_g = true;
        }
      }
    }
    // earliest pending same-window self-send
    int pj = -1;
    for (int j = 0; j < ctx.npend; j++)
      if (pj < 0 || ctx.pend_time[j] < ctx.pend_time[pj] ||
          (ctx.pend_time[j] == ctx.pend_time[pj] && ctx.pend_q2[j] < ctx.pend_q2[pj]))
        pj = j;
    double time;
    uint32_t handler, src, ev_flags = 0, ev_seq = 0;
    unsigned long long data;
    if (pj >= 0 &&
        (!have_g || ctx.pend_time[pj] < __longlong_as_double((long long)SLOT_LO(sh, gj).x))) {
      time = ctx.pend_time[pj];
      handler = (uint32_t)(ctx.pend_q2[pj] & 0xFFFFu);
      src = ctx.self_gid;
      data = ctx.pend_data[pj];
      ctx.npend--;
      ctx.pend_time[pj] = ctx.pend_time[ctx.npend];
      ctx.pend_q2[pj] = ctx.pend_q2[ctx.npend];
      ctx.pend_data[pj] = ctx.pend_data[ctx.npend];
      ctx.job_slot = SIMIAN_NONE; // no slot to park a deferred sum on
    } else if (have_g) {
      ctx.job_slot = gj;
      const RecBits g = slot_get(sh, gj);
      ev_flags = g.flags();
      ev_seq = g.seq();
      time = g.time();
      handler = g.handler();
      src = g.src();
      data = g.data();
      p1 = g1;
      p2 = g2;
      p3 = g3;
      p4 = gj;
      have_prev = true;
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
    ctx.flags_ = ev_flags;
    ctx.seq_ = ev_seq;
    ctx.pl_sh = &sh;
    uint32_t fn = handler < SIMIAN_MAX_SERVICES ? ctx.cl->fn_of_service[handler] : 0;
    run_handler(ctx, fn); // simian.js:129-131
  }
  c.core[slot] = st;
}

// ---- thread-per-event path (pure handlers) --------------------------------
// Phase A1, one thread per EVENT: rank of the event inside its entity's list
// under (time, handler, src, seq), tie check against its predecessor, and the
// order-independent part of the trace hash, which is parked on the rank-th
// node of the entity's list so that the chain phase reads it in commit order.
// Returns the rank.
__device__ __forceinline__ uint32_t rank_event(const DevCfg &c, WindowShared &sh, ThreadAcc &ta,
                                               uint32_t slot0, uint32_t i) {
  const ulonglong2 a = SLOT_LO(sh, i);
  const uint32_t le = (uint32_t)a.y - slot0;
  const double tr = __longlong_as_double((long long)a.x);
  uint32_t rank = 0;
  const uint32_t first = sh.head[le];
  const ulonglong2 bh = SLOT_HI(sh, i);
  // tie-break words of this event: (handler, src) and seq
  const unsigned long long my2 = ((bh.x & 0xFFFFull) << 32) | (a.y >> 32);
  const uint32_t myseq = (uint32_t)(bh.x >> 32);
  // tie statistics: is there an equal-time event before this one, and does
  // any of those differ from it in what the handler sees
  bool has_pred = false, any_diff = false;
  // t = the rank-th node of the list: it trails the walk (rank so far <= nodes
  // visited so far), one step forward for every event found to precede this one
  uint32_t t = first;
#pragma unroll 1
  for (uint32_t j = first; j != LIST_END; j = sh.nxt[j]) {
    if (j == i) continue;
    const ulonglong2 o = SLOT_LO(sh, j);
    const double to = __longlong_as_double((long long)o.x);
    if (to < tr) {
      rank++;
      t = sh.nxt[t];
    } else if (to == tr) { // a true same-entity time tie: the defined order decides
      const ulonglong2 oh = SLOT_HI(sh, j);
      const unsigned long long o2 = ((oh.x & 0xFFFFull) << 32) | (o.y >> 32);
      const uint32_t oseq = (uint32_t)(oh.x >> 32);
      // identical records (same key down to seq) are ordered by their slot: any
      // order of indistinguishable events commits the same thing
      const bool before = o2 != my2 ? o2 < my2 : (oseq != myseq ? oseq < myseq : j < i);
      if (before) {
        rank++;
        t = sh.nxt[t];
        has_pred = true;
        any_diff |= (o2 != my2) | (oh.y != bh.y);
      }
    }
  }
  if (c.tie_check && has_pred) {
    ta.ties++;
    // distinguishable tie: the LATEST equal-time predecessor differs in (handler,
    // src) or data.  When every equal-time predecessor is indistinguishable from
    // this event (bursts of identical schedService calls) the answer is no;
    // otherwise the latest one is looked for (rare: models with real ties).
    if (any_diff) {
      uint32_t pj = LIST_END, pseq = 0;
      unsigned long long p2 = 0;
#pragma unroll 1
      for (uint32_t j = first; j != LIST_END; j = sh.nxt[j]) {
        if (j == i || __longlong_as_double((long long)SLOT_LO(sh, j).x) != tr) continue;
        const unsigned long long ox = SLOT_HI(sh, j).x;
        const unsigned long long o2 = ((ox & 0xFFFFull) << 32) | (SLOT_LO(sh, j).y >> 32);
        const uint32_t oseq = (uint32_t)(ox >> 32);
        const bool before = o2 != my2 ? o2 < my2 : (oseq != myseq ? oseq < myseq : j < i);
        // keep the latest predecessor: (key, seq, slot) maximal
        if (before &&
            (pj == LIST_END || (p2 != o2 ? p2 < o2 : (pseq != oseq ? pseq < oseq : pj < j)))) {
          pj = j;
          p2 = o2;
          pseq = oseq;
        }
      }
      if (p2 != my2 || SLOT_HI(sh, pj).y != bh.y) ta.dties++;
    }
  }
  sh.xs[t] = simian_trace_event(tr, (uint32_t)(bh.x & 0xFFFFull), (uint32_t)(a.y >> 32), bh.y);
  unsigned long long tk = simian_time_key(tr);
  ta.max_key = tk > ta.max_key ? tk : ta.max_key;
  ta.committed++;
  return rank;
}

// Phase A2, one thread per EVENT: the handler itself (simian.js:129-131): RNG,
// log, destination.  Commit index = the entity's committed count + rank.  The
// record it sends to a local entity replaces the consumed one in slot i.
template <bool WIDE>
__device__ __forceinline__ void run_event(const DevCfg &c, WindowShared &sh, ThreadAcc &ta,
                                          uint32_t slot0, uint32_t i, uint32_t rank) {
  const RecBits r = slot_get(sh, i);
  const uint32_t slot = r.dst();
  int k;
  DevCtx<false, WIDE> ctx(c, sh.W, sh.ax, ta);
  ctx.stage_lo = &SLOT_LO(sh, i);
  ctx.stage_hi = &SLOT_HI(sh, i);
  ctx.self_slot = slot;
  ctx.self_gid = gid_of_slot(c, slot, k);
  ctx.cl = &c.cls[k];
  ctx.commit_idx = SH_CORE(sh)[slot - slot0].count + rank;
  ctx.key = simian_rng_key(c.seed, ctx.self_gid, ctx.commit_idx);
  ctx.n_emit = 0;
  ctx.now_ = r.time(); // simian.js:126
  ctx.handler_ = r.handler();
  ctx.src_ = r.src();
  ctx.data_ = r.data();
  ctx.flags_ = r.flags();
  ctx.seq_ = r.seq();
  ctx.pl_sh = &sh;
  uint32_t h = r.handler();
  run_pure_handler(ctx, h < SIMIAN_MAX_SERVICES ? ctx.cl->fn_of_service[h] : 0);
  if (ctx.stage_lo) SLOT_LO(sh, i).y = 0xFFFFFFFFFFFFFFFFull; // nothing staged
}

// Phase C, one thread per entity: fold the trace words, already in commit
// order along the list, and publish the new commit count.
__device__ __forceinline__ void chain_entity(const DevCfg &c, WindowShared &sh, uint32_t slot,
                                             uint32_t le, uint32_t first) {
  EntityCore st = SH_CORE(sh)[le];
  uint32_t n = 0;
#pragma unroll 1
  for (uint32_t j = first; j != LIST_END; j = sh.nxt[j]) {
    st.hash = simian_trace_chain(st.hash, sh.xs[j]);
    n++;
  }
  st.count += n;
  c.core[slot] = st;
}

// K2: ONE coalesced pass over the block's pages.  Every record whose time is
// inside [gmin, hi) and whose entity is inside [e_lo, e_hi) is copied into a
// shared-memory slot (cp.async, no registers) and pushed onto its entity's
// list (eventQ.pop for all of them at once, eventQ.js:53-92); the first event
// of an entity also fetches the entity's {count, hash}.  On the first pass the
// minimum time >= hi is tracked (K1).  Slots are claimed with one
// shared-memory atomic per warp.
template <bool BLOBS>
__device__ __forceinline__ void collect_due(const DevCfg &c, WindowShared &sh, ThreadAcc &ta,
                                            uint32_t slot0, uint32_t pg_lo, uint32_t pg_hi,
                                            uint32_t e_lo, uint32_t e_hi, bool first) {
  const int tid = threadIdx.x, lane = tid & 31;
  const double w_hi = sh.W.hi, w_gmin = sh.W.gmin;
  constexpr int U = 4;
  // thread -> (page, slot): the slot inside the page is fixed per thread, the
  // page advances by NTHREADS / PAGE per step (PAGE divides NTHREADS or vice versa)
  static_assert(NTHREADS % PAGE == 0 || PAGE % NTHREADS == 0, "page / CTA shape");
  const uint32_t lim = (pg_hi - pg_lo) * PAGE;
  for (uint32_t base = 0; base < lim; base += U * NTHREADS) {
    unsigned long long q0[U], q1[U];
    uint32_t ref[U];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const uint32_t i = base + u * NTHREADS + tid;
      const uint32_t pg = pg_lo + i / PAGE, sl = i % PAGE;
      ok[u] = false;
      ref[u] = 0;
      if (i < lim) {
        const uint32_t cnt = (uint32_t)(sh.pg_cnt[pg] & 0x7FFFu);
        ok[u] = sl < cnt;
        ref[u] = sh.pg_id[pg] * PAGE + sl;
      }
      if (ok[u]) ld_rec_head(c.pool + ref[u], q0[u], q1[u]);
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      bool due = false;
      uint32_t le = 0;
      if (ok[u]) {
        double t = __longlong_as_double((long long)q0[u]);
        if (t < w_hi) {
          le = (uint32_t)q1[u] - slot0;
          due = t >= w_gmin && le >= e_lo && le < e_hi;
        } else if (first) {
          unsigned long long tk = simian_time_key(t);
          ta.min_key = tk < ta.min_key ? tk : ta.min_key;
        }
      }
      unsigned m = __ballot_sync(0xffffffffu, due);
      if (m) {
        uint32_t p0 = 0;
        int leader = __ffs(m) - 1;
        if (lane == leader) p0 = atomicAdd(&sh.total, (uint32_t)__popc(m));
        p0 = __shfl_sync(0xffffffffu, p0, leader);
        if (due) {
          uint32_t pos = p0 + __popc(m & ((1u << lane) - 1u));
          if (pos < NREFS) {
            SLOT_LO(sh, pos) = make_ulonglong2(q0[u], q1[u]);
            cp_async16(&SLOT_HI(sh, pos), (const char *)(c.pool + ref[u]) + 16);
            if (BLOBS && (ld_rec_flags(c.pool + ref[u]) & SIMIAN_EVF_CONT)) {
              // a payload word: listed for payload(k), in no entity's list
              sh.nxt[pos] = (uint16_t)atomicExch(&sh.cont_head, pos);
            } else {
              uint32_t prev = atomicExch(&sh.head[le], pos);
              sh.nxt[pos] = (uint16_t)prev;
              if (prev == LIST_END) cp_async16(&SH_CORE(sh)[le], &c.core[slot0 + le]);
            }
          } else if (first) {
            // overflow (a burst): where in the block the records without a slot
            // belong -- the partition step sizes its parts from this
            atomicAdd(&sh.hist[le >> sh.bin_shift], 1u);
          }
        }
      }
    }
  }
  cp_async_wait_all();
}

// K2, common case: the pages of the window's rows hold no more records than
// the CTA has slots.  ALL of them are copied into the shared-memory slots in
// one round trip (cp.async, no registers, every load of the CTA in flight at
// once); slot = records in the pages listed before + index in the page.
#ifdef SIMIAN_BULK_SCAN
// ONE bulk copy per listed page (its records are contiguous: <= 4 KB), issued by
// one thread each; the bytes land on the CTA's mbarrier, which every thread then
// waits on itself -- no block barrier between the copy and the classification.
__device__ __forceinline__ void collect_all(const DevCfg &c, WindowShared &sh, uint32_t npg,
                                            uint32_t nvalid) {
  if (threadIdx.x == 0) mbar_arrive_expect_tx(&sh.mbar, nvalid * 32u);
  for (uint32_t i = threadIdx.x; i < npg; i += NTHREADS) {
    const uint32_t n = (uint32_t)(sh.pg_cnt[i] & 0x7FFFu);
    if (n)
      bulk_g2s(&sh.rec[sh.pg_off[i]], c.pool + (size_t)sh.pg_id[i] * PAGE, n * 32u, &sh.mbar);
  }
  uint32_t spins = 0;
  while (!mbar_try_wait(&sh.mbar, 0u))
    if (++spins > (1u << 24)) { // never legitimately: fail, do not hang
      set_error(c, SIMIAN_ERR_STALLED);
      break;
    }
}
#else
__device__ __forceinline__ void collect_all(const DevCfg &c, WindowShared &sh, uint32_t npg,
                                            uint32_t /*nvalid*/) {
  const uint32_t lim = npg * PAGE;
  for (uint32_t i = threadIdx.x; i < lim; i += NTHREADS) {
    const uint32_t pg = i / PAGE, sl = i % PAGE;
    if (sl < (uint32_t)(sh.pg_cnt[pg] & 0x7FFFu)) {
      const simian_event_t *src = c.pool + ((size_t)sh.pg_id[pg] * PAGE + sl);
      const uint32_t pos = sh.pg_off[pg] + sl;
      cp_async16(&SLOT_LO(sh, pos), src);
      cp_async16(&SLOT_HI(sh, pos), (const char *)src + 16);
    }
  }
  cp_async_wait_all();
}
#endif

// ... then every thread classifies the records in its slots (the same slots it
// handles in the later phases): due events are pushed onto their entity's list
// and fetch the entity's {count, hash} (waited for after the rank phase); the
// others only feed the LBTS minimum (K1).  Returns the mask of its due slots.
// BLOBS (option "payloads", its own instantiation of the window kernel so that the
// ordinary one never looks at a record's flags): continuation records are listed for
// payload(k) instead of being linked as events.
template <bool BLOBS>
__device__ __forceinline__ uint32_t link_due(const DevCfg &c, WindowShared &sh, ThreadAcc &ta,
                                             uint32_t slot0, uint32_t nvalid) {
  const double w_hi = sh.W.hi, w_gmin = sh.W.gmin;
  uint32_t live = 0;
#pragma unroll
  for (int u = 0; u < EV_PER_THREAD; u++) {
    const uint32_t i = threadIdx.x + u * NTHREADS;
    if (i < nvalid) {
      const ulonglong2 a = SLOT_LO(sh, i);
      const double t = __longlong_as_double((long long)a.x);
      if (t < w_hi) {
        if (t >= w_gmin) {
          if (BLOBS && ((uint32_t)(SLOT_HI(sh, i).x >> 16) & SIMIAN_EVF_CONT)) {
            // a payload word: listed for payload(k), in no entity's list, not an event
            sh.nxt[i] = (uint16_t)atomicExch(&sh.cont_head, i);
          } else {
            const uint32_t le = (uint32_t)a.y - slot0;
            const uint32_t prev = atomicExch(&sh.head[le], i);
            sh.nxt[i] = (uint16_t)prev;
            if (prev == LIST_END) cp_async16(&SH_CORE(sh)[le], &c.core[slot0 + le]);
            live |= 1u << u;
          }
        }
      } else {
        unsigned long long tk = simian_time_key(t);
        ta.min_key = tk < ta.min_key ? tk : ta.min_key;
      }
    }
  }
  return live;
}

// Burst partition.  The first pass found more due events than the CTA has
// slots for (e.g. the 16 events per entity all scheduled at t = 0).  Instead
// of re-scanning every page of the window once per entity sub-range, the due
// records are copied ONCE into scratch pages grouped by entity sub-range
// ("parts", sized from the histogram the overflowing pass built); each part is
// then one ordinary pass over its own few pages, all of whose records are due.
// The scratch pages come from the block's page stack / the free ring, are
// listed behind the cells' pages in pg_id and are recycled with them at the end
// of the launch.  Returns (CTA-uniform) false when the page list has no room:
// the caller falls back to splitting the entity range and re-scanning.
__device__ __noinline__ bool partition_burst(const DevCfg &c, WindowShared &sh, uint32_t slot0,
                                            uint32_t npg_scan) {
  const int tid = threadIdx.x;
  const uint32_t shift = sh.bin_shift;
  // the records that did get a slot in the overflowing pass
  for (uint32_t i = tid; i < NREFS; i += NTHREADS)
    atomicAdd(&sh.hist[((uint32_t)SLOT_LO(sh, i).y - slot0) >> shift], 1u);
  __syncthreads();
  if (tid == 0) {
    const uint32_t nb = ((c.epb - 1u) >> shift) + 1u; // bins in use
    uint32_t q = 0, cur = 0, bin0 = 0, pages = sh.npg;
    bool fits = true;
    for (uint32_t bin = 0; bin <= nb; bin++) {
      const uint32_t h = bin < nb ? sh.hist[bin] : 0u;
      if (bin == nb || (cur != 0 && cur + h > (uint32_t)NREFS)) { // close part q
        const uint32_t np = (cur + PAGE - 1) / PAGE;
        if (pages + np > (uint32_t)PGMAX) {
          fits = false;
          break;
        }
        sh.pcnt[q] = cur;
        sh.pfill[q] = 0;
        sh.plo[q] = bin0 << shift;
        sh.phi[q] = (bin << shift) < c.epb ? (bin << shift) : c.epb;
        sh.ppage0[q] = pages;
        for (uint32_t k = 0; k < np; k++)
          sh.pg_cnt[pages + k] =
              (uint16_t)((k + 1 < np ? (uint32_t)PAGE : cur - k * PAGE) | PG_FREE_FLAG);
        pages += np;
        q++;
        bin0 = bin;
        cur = 0;
      }
      if (bin < nb) {
        sh.part_of_bin[bin] = q;
        cur += h;
      }
    }
    sh.ppage0[q] = pages;
    sh.nparts = fits ? q : 0u;
    if (fits) sh.nfree += pages - sh.npg; // the scratch pages are recycled with the others
  }
  __syncthreads();
  const uint32_t nparts = sh.nparts;
  if (nparts == 0) return false;
  const uint32_t pg_end = sh.ppage0[nparts];
  for (uint32_t i = npg_scan + tid; i < pg_end; i += NTHREADS)
    sh.pg_id[i] = page_get(c, sh.ax, sh.W.alloc_limit); // SIMIAN_NONE: sticky error set
  __syncthreads();
  if (tid == 0) {
    sh.npg = pg_end; // recycled with the cells' pages
    sh.found = ld_cg(&c.acc->error);
  }
  __syncthreads();
  if (sh.found) return true; // pool exhausted: nothing is pushed, the launch ends
  // one pass over the window's pages: every due record -> its part's scratch pages
  const double w_hi = sh.W.hi, w_gmin = sh.W.gmin;
  constexpr int U = 4;
  const uint32_t lim = npg_scan * PAGE;
  for (uint32_t base = 0; base < lim; base += U * NTHREADS) {
    RecBits r[U];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const uint32_t i = base + u * NTHREADS + tid;
      const uint32_t pg = i / PAGE, sl = i % PAGE;
      ok[u] = i < lim && sl < (uint32_t)(sh.pg_cnt[pg < PGMAX ? pg : 0] & 0x7FFFu);
      if (ok[u]) r[u] = ld_rec_2x16(c.pool + ((size_t)sh.pg_id[pg] * PAGE + sl));
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      if (!ok[u]) continue;
      const double t = r[u].time();
      if (!(t < w_hi && t >= w_gmin)) continue;
      const uint32_t q = sh.part_of_bin[(r[u].dst() - slot0) >> shift];
      const uint32_t pos = atomicAdd(&sh.pfill[q], 1u);
      if (pos >= sh.pcnt[q]) { // cannot happen: both passes see the same records
        set_error(c, SIMIAN_ERR_WINDOW_OVERFLOW);
        continue;
      }
      st_rec_2x16(c.pool + ((size_t)sh.pg_id[sh.ppage0[q] + pos / PAGE] * PAGE + pos % PAGE),
                  r[u].q0, r[u].q1, r[u].q2, r[u].q3);
    }
  }
  __syncthreads();
  if (tid == 0) {
    for (uint32_t q = nparts; q-- > 0;) { // lowest sub-range on top of the stack
      if (sh.pcnt[q] == 0) continue;
      uint32_t *st = &sh.rstack[4 * sh.stack_n];
      st[0] = sh.plo[q];
      st[1] = sh.phi[q];
      st[2] = sh.ppage0[q];
      st[3] = sh.ppage0[q + 1];
      sh.stack_n++;
    }
  }
  __syncthreads();
  return true;
}

// walk one cell's page chain into the CTA's page list.  `lo32`: the count half
// of the cell's head word (records in the head page, CELL_OLDER)
__device__ __forceinline__ void list_cell_pages(const DevCfg &c, WindowShared &sh, uint32_t head,
                                                uint32_t lo32, uint32_t flag) {
  uint32_t p = head, cntp = CELL_COUNT(lo32), older = lo32 & CELL_OLDER;
  if (cntp > (uint32_t)PAGE) cntp = PAGE;
  if (p == SIMIAN_NONE) return;
#pragma unroll 1
  for (;;) {
    uint32_t idx = atomicAdd(&sh.npg, 1u);
    uint32_t off = atomicAdd(&sh.nvalid, cntp);
    if (idx < PGMAX) {
      sh.pg_id[idx] = p;
      sh.pg_cnt[idx] = (uint16_t)(cntp | flag);
      sh.pg_off[idx] = off;
      if (flag) atomicAdd(&sh.nfree, 1u); // recycled after this window
    }
    if (!older) break; // the common case: no link is read
    const uint32_t raw = c.page_next[p];
    if (raw == SIMIAN_NONE) break;
    p = raw & 0x7FFFFFFFu;
    older = raw & CELL_OLDER;
    cntp = PAGE; // older pages are always full
  }
}

// ------- LBTS candidate + next-window set-up (one CTA, all threads) --------

// local minLeft candidate (simian.js:143): min over every pending local event.
// acc->min_key already holds leftovers of row tb_hi + everything appended since
// the last advance; rows above tb_hi are searched only when that candidate does
// not already lie in row tb_hi (sparse case: the window jumped over idle time).
template <class SH>
__device__ void local_min_candidate(const DevCfg &c, const WinState &W, SH &sh,
                                    unsigned long long &cand_out,
                                    unsigned long long &far_out) {
  const int tid = threadIdx.x;
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
  long long j0 = W.tb_hi + 1;
  // first non-empty sub-bucket row in [j0, jmax], (row, block) pairs in row-major order
  if (tid == 0) sh.scan_j = 0x7FFFFFFFFFFFFFFFLL;
  __syncthreads();
  if (j0 <= jmax) {
    unsigned long long total = (unsigned long long)(jmax - j0 + 1) * c.n_blocks;
    for (unsigned long long base = 0; base < total; base += (unsigned long long)NTHREADS * 8) {
#pragma unroll 1
      for (int u = 0; u < 8; u++) {
        unsigned long long i = base + (unsigned long long)u * NTHREADS + tid;
        if (i < total) {
          long long j = j0 + (long long)(i / c.n_blocks);
          uint32_t b = (uint32_t)(i % c.n_blocks);
          if ((uint32_t)(ld_cg64(&c.cell_state[(uint32_t)(j & (long long)c.ring_mask) *
                                                   c.n_blocks + b]) >> 32) != SIMIAN_NONE)
            atomicMin((unsigned long long *)&sh.scan_j, (unsigned long long)j);
        }
      }
      __syncthreads();
      if (sh.scan_j != 0x7FFFFFFFFFFFFFFFLL) break;
    }
  }
  __syncthreads();
  long long jf = sh.scan_j;
  __syncthreads();
  if (jf != 0x7FFFFFFFFFFFFFFFLL) {
    if (tid == 0) sh.min_key = SIMIAN_KEY_NONE;
    __syncthreads();
    unsigned long long mn = SIMIAN_KEY_NONE;
    uint32_t row = (uint32_t)(jf & (long long)c.ring_mask) * c.n_blocks;
    for (uint32_t b = tid; b < c.n_blocks; b += NTHREADS) {
      unsigned long long hw = ld_cg64(&c.cell_state[row + b]);
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
                                  double far_min_time, bool after_redist) {
  const int tid = threadIdx.x;
  DevAcc *a = c.acc;
  __shared__ WinState nw;
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
      if (nw.tb_hi - nw.tb_lo + 1 > SIMIAN_MAX_CELLS)
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
  // snapshot the heads of row tb_hi: appends of the next launch may target it
  if (!nw.done) {
    uint32_t row = (uint32_t)(nw.tb_hi & (long long)c.ring_mask) * c.n_blocks;
    for (uint32_t b0 = 0; b0 < c.n_blocks; b0 += 8 * blockDim.x) { // 8 loads in flight per thread
      unsigned long long v[8];
#pragma unroll
      for (int q = 0; q < 8; q++) {
        uint32_t b = b0 + q * blockDim.x + tid;
        if (b < c.n_blocks) v[q] = ld_cg64(&c.cell_state[row + b]);
      }
#pragma unroll
      for (int q = 0; q < 8; q++) {
        uint32_t b = b0 + q * blockDim.x + tid;
        if (b < c.n_blocks) c.snap[b] = v[q];
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

// free a whole cell (pages -> ring, head cleared); one thread
__device__ __forceinline__ void free_cell(const DevCfg &c, uint32_t cell) {
  const unsigned long long hw = c.cell_state[cell];
  uint32_t p = (uint32_t)(hw >> 32), older = (uint32_t)hw & CELL_OLDER;
  while (p != SIMIAN_NONE) {
    const uint32_t raw = older ? c.page_next[p] : SIMIAN_NONE;
    page_free(c, p);
    if (raw == SIMIAN_NONE) break;
    p = raw & 0x7FFFFFFFu;
    older = raw & CELL_OLDER;
  }
  c.cell_state[cell] = CELL_EMPTY;
}

// Far-row redistribution for block b (own cells only: no cross-CTA traffic).
__device__ void redistribute_block(const DevCfg &c, WindowShared &sh, uint32_t b) {
  const WinState &W = sh.W;
  const int tid = threadIdx.x;
  // 1. recycle rows [tb_clean, tb_clean_new): only dead records remain there
  long long nrec = W.tb_clean_new - W.tb_clean;
  if (nrec > (long long)c.ring) nrec = c.ring;
  for (long long i = tid; i < nrec; i += NTHREADS) {
    long long tb = W.tb_clean + i;
    free_cell(c, (uint32_t)(tb & (long long)c.ring_mask) * c.n_blocks + b);
  }
  __syncthreads();
  // 2. move the far row: into the calendar when within the new horizon, else
  //    into the other far row
  uint32_t far_cell = (c.ring + W.far_row) * c.n_blocks + b;
  uint32_t other = W.far_row ^ 1u;
  if (tid == 0) sh.npg = 0;
  __syncthreads();
  if (tid == 0) {
    unsigned long long hw = c.cell_state[far_cell];
    list_cell_pages(c, sh, (uint32_t)(hw >> 32), (uint32_t)hw, PG_FREE_FLAG);
  }
  __syncthreads();
  uint32_t npg = sh.npg;
  if (npg > PGMAX) {
    if (tid == 0) set_error(c, SIMIAN_ERR_WINDOW_OVERFLOW);
    npg = PGMAX;
  }
  for (uint32_t i = tid; i < npg * PAGE; i += NTHREADS) {
    uint32_t pg = i / PAGE, s = i % PAGE;
    if (s >= (uint32_t)(sh.pg_cnt[pg] & 0x7FFFu)) continue;
    RecBits r = ld_rec(c.pool + ((size_t)sh.pg_id[pg] * PAGE + s));
    append_local(c, sh.ax, W.tb_clean_new, other, W.alloc_limit, r);
  }
  __syncthreads();
  for (uint32_t i = tid; i < npg; i += NTHREADS) page_free(c, sh.pg_id[i]);
  if (tid == 0) c.cell_state[far_cell] = CELL_EMPTY;
}

// Sum the launch's results into the global accumulators (one CTA, all threads;
// runs after the ticket fence, so every CTA's contribution is visible).  The
// CTAs add into SIMIAN_PART_SLOTS accumulator slots (block index mod slots), so
// this is ONE round trip, not a loop over the blocks; the slots are cleared
// for the next launch.
template <class SH>
__device__ void reduce_parts(const DevCfg &c, SH &sh) {
  const int tid = threadIdx.x;
  __shared__ unsigned long long red[10];
  if (tid < 10) red[tid] = tid == 0 ? SIMIAN_KEY_NONE : 0ull;
  __syncthreads();
  if (tid < SIMIAN_PART_SLOTS) {
    BlockPart *p = &c.part[tid];
    unsigned long long mn = ld_cg64(&p->min_key), mx = ld_cg64(&p->max_key);
    uint32_t v[8];
    const uint32_t *w = &p->committed;
#pragma unroll
    for (int q = 0; q < 8; q++) v[q] = ld_cg(w + q);
    p->min_key = SIMIAN_KEY_NONE;
    p->max_key = 0;
#pragma unroll
    for (int q = 0; q < 8; q++) (&p->committed)[q] = 0;
    mn = warp_min_u64(mn);
    mx = warp_max_u64(mx);
#pragma unroll
    for (int q = 0; q < 8; q++) v[q] = __reduce_add_sync(0xffffffffu, v[q]);
    if ((tid & 31) == 0) {
      if (mn != SIMIAN_KEY_NONE) atomicMin(&red[0], mn);
      if (mx) atomicMax(&red[1], mx);
#pragma unroll
      for (int q = 0; q < 8; q++)
        if (v[q]) atomicAdd(&red[2 + q], (unsigned long long)v[q]);
    }
  }
  __syncthreads();
  if (tid == 0) {
    DevAcc *a = c.acc;
    unsigned long long m0 = ld_cg64(&a->min_key); // ingested / scheduled since the last advance
    a->min_key = red[0] < m0 ? red[0] : m0;
    unsigned long long m1 = ld_cg64(&a->max_key);
    a->max_key = red[1] > m1 ? red[1] : m1;
    a->committed += red[2];
    a->emitted += red[3];
    a->dropped += red[4];
    a->ties += red[5];
    a->dist_ties += red[6];
    a->appended += red[7];
    a->far_appends += red[8];
    a->page_allocs += red[9];
    __threadfence();
  }
  __syncthreads();
}

// size > 1: this rank's minLeft candidate (simian.js:143) -> minvec, and, with
// `publish`, the send half of the allreduce(MIN) over peer memory: one thread
// per peer writes the candidate into the peer's slot array -- data, fence, flag.
// One CTA, all threads; the launch's results must be visible (ticket fence or
// a kernel boundary).
template <class SH>
__device__ void lbts_candidate(const DevCfg &c, const WinState &W, SH &sh, uint32_t win_ix,
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

// Phase B (K4), one thread per EVENT: append the records the handlers staged in the
// shared-memory slots (slot i = threadIdx.x + u * NTHREADS for bit u of `live`; a
// slot whose dst is NONE staged nothing).  The claims of all of a thread's events
// are in flight together; `between` runs while they are.
template <bool WIDE, int NEV, class SH, class F>
__device__ __forceinline__ void append_staged(const DevCfg &c, SH &sh, AppendCtx &ax,
                                              const WinState &W, uint32_t live, uint32_t base,
                                              uint32_t stride, F between) {
  unsigned long long old[NEV];
  uint32_t cell[NEV];
#pragma unroll
  for (int u = 0; u < NEV; u++) {
    uint32_t i = base + u * stride;
    cell[u] = SIMIAN_NONE;
    if (live & (1u << u)) {
      ulonglong2 a = SLOT_LO(sh, i);
      if ((uint32_t)a.y != SIMIAN_NONE)
        cell[u] = cell_of_local(c, ax, W.tb_clean, W.far_row,
                                __longlong_as_double((long long)a.x), (uint32_t)a.y);
    }
  }
#pragma unroll
  for (int u = 0; u < NEV; u++)
    if (cell[u] != SIMIAN_NONE) old[u] = atomicAdd(c.cell_state + cell[u], 1ull);
  between();
  // Claims are serviced in two rounds so that no thread ever waits for a
  // page while it still owes one: first every claim that landed in a
  // page or must INSTALL one (never blocks); only then the claims that
  // have to wait for another thread's install.  (Waiting first can
  // deadlock: two threads holding crossed install/wait claims.)
  uint32_t waiting = 0;
#pragma unroll
  for (int u = 0; u < NEV; u++)
    if (cell[u] != SIMIAN_NONE) {
      if (CELL_COUNT(old[u]) <= (uint32_t)PAGE) { // lands in a page, or installs one
        const RecBits r = slot_get(sh, base + u * stride);
        if ((uint32_t)old[u] < (uint32_t)PAGE)
          st_rec_sel<WIDE>(
              c.pool + ((size_t)(uint32_t)(old[u] >> 32) * PAGE + (uint32_t)old[u]), r);
        else
          cell_append_slow(c, ax, W.alloc_limit, c.cell_state + cell[u], old[u], r.q0, r.q1,
                           r.q2, r.q3);
      } else {
        waiting |= 1u << u;
      }
    }
  if (waiting) { // rare
#pragma unroll
    for (int u = 0; u < NEV; u++)
      if (waiting & (1u << u)) {
        const RecBits r = slot_get(sh, base + u * stride);
        cell_append_slow(c, ax, W.alloc_limit, c.cell_state + cell[u], old[u], r.q0, r.q1,
                         r.q2, r.q3);
      }
  }
}

// K3 + K4 + trace chain for the m due events one pass collected into the
// shared-memory slots (entities [e_lo, e_hi) of the block).  CTA-uniform.
// `live`: bit u = slot (threadIdx.x + u * NTHREADS) holds a due event of this thread.
// STAGE (split window, k_window_stage): after the rank phase the events are parked
// in HBM for k_emit -- compacted, each with its commit index -- instead of running
// their handlers here; the trace chain follows at once.
// WIDE: the pass is inlined into a kernel body (256-bit record stores allowed).
template <bool STAGE, bool WIDE>
__device__ __forceinline__ void process_pass(const DevCfg &c, WindowShared &sh, ThreadAcc &ta,
                                             uint32_t slot0, uint32_t live, uint32_t e_lo,
                                             uint32_t e_hi) {
  const int tid = threadIdx.x;
  const WinState &W = sh.W;
  if (c.all_pure) {
    // A1, one thread per EVENT: rank (kept in a register), tie check, trace word
    unsigned long long rkpack = 0; // EV_PER_THREAD ranks, 8 bits each
#pragma unroll 1
    for (int u = 0; u < EV_PER_THREAD; u++) {
      uint32_t i = tid + u * NTHREADS;
      if (live & (1u << u)) {
        uint32_t r = rank_event(c, sh, ta, slot0, i);
        if (r > 255u) set_error(c, SIMIAN_ERR_ENTITY_BURST);
        rkpack |= (unsigned long long)(r & 255u) << (8 * u);
      }
    }
    cp_async_wait_all(); // the entities' {count, hash} (link_due)
    __syncthreads();
    PHASE(3);
    if constexpr (STAGE) {
      simian_event_t *out = c.stage + (size_t)blockIdx.x * NREFS;
      const uint32_t lane = (uint32_t)tid & 31u;
#pragma unroll
      for (int u = 0; u < EV_PER_THREAD; u++) {
        const bool lv = (live >> u) & 1u;
        const unsigned m = __ballot_sync(0xffffffffu, lv);
        if (!m) continue; // (warp-uniform)
        uint32_t p0 = 0;
        if (lane == 0) p0 = atomicAdd(&sh.total, (uint32_t)__popc(m));
        p0 = __shfl_sync(0xffffffffu, p0, 0);
        if (lv) {
          RecBits r = slot_get(sh, tid + u * NTHREADS);
          const unsigned long long ci =
              SH_CORE(sh)[r.dst() - slot0].count + (uint32_t)((rkpack >> (8 * u)) & 255ull);
          if (ci >> 48) set_error(c, SIMIAN_ERR_ENTITY_BURST); // (2.8e14 commits of one entity)
          r.q2 = (r.q2 & 0xFFFFull) | (ci << 16);
          st_rec(out + p0 + __popc(m & ((1u << lane) - 1u)), r);
        }
      }
      PHASE(4);
      for (uint32_t le = e_lo + tid; le < e_hi; le += NTHREADS) {
        uint32_t f = sh.head[le];
        if (f != LIST_END) chain_entity(c, sh, slot0 + le, le, f);
      }
      PHASE(5);
    } else {
    // A2 (K3), one thread per EVENT: handler; the sent record is staged in
    // the consumed record's slot
#pragma unroll 1
    for (int u = 0; u < EV_PER_THREAD; u++) {
      uint32_t i = tid + u * NTHREADS;
      if (live & (1u << u)) run_event<WIDE>(c, sh, ta, slot0, i, (uint32_t)((rkpack >> (8 * u)) & 255ull));
    }
    PHASE(4); // (no barrier: phase B reads only the slots its own thread staged)
    // B (K4) with C (trace chain) folded while the claims are in flight
    append_staged<WIDE, EV_PER_THREAD>(c, sh, sh.ax, W, live, (uint32_t)tid, (uint32_t)NTHREADS, [&]() {
      // C, one thread per entity: fold the trace hash chain, publish the count
      // (reads what phase A1 parked; A2 / B only touch the record slots)
      for (uint32_t le = e_lo + tid; le < e_hi; le += NTHREADS) {
        uint32_t f = sh.head[le];
        if (f != LIST_END) chain_entity(c, sh, slot0 + le, le, f);
      }
    });
    PHASE(5);
    } // (!STAGE)
  } else if constexpr (!STAGE) { // (the split window is for models of pure handlers only)
    cp_async_wait_all(); // the entities' {count, hash} (link_due)
    // no deferred sum is parked on any slot yet (xs is otherwise unused on this path)
#pragma unroll
    for (int u = 0; u < EV_PER_THREAD; u++)
      if (live & (1u << u)) sh.xs[tid + u * NTHREADS] = 0ull;
    if (tid == 0) sh.pass_jobs = sh.job_count = 0;
    __syncthreads();
    // K3 + K4: stateful handlers, one thread per entity.  A separate
    // accumulator keeps `ta` of the thread-per-event path in registers
    // (process_entity is not inlined).
    ThreadAcc t2;
    thread_acc_init(t2);
    for (uint32_t le = e_lo + tid; le < e_hi; le += NTHREADS) {
      uint32_t f = sh.head[le];
      if (f != LIST_END) process_entity(c, sh, t2, slot0 + le, f);
    }
    // ---- the sums the handlers deferred (simian_defer_weighted_sum): one WARP per
    // entity, the entity's jobs in commit order, each in the defined order
    __syncthreads();
    PHASE(4);
    if (c.jobs) { // split sums: one range of the global job array for this pass
      if (tid == 0) {
        sh.job_base = SIMIAN_NONE;
        const uint32_t n = sh.pass_jobs;
        if (n) {
          unsigned long long *tot = &c.acc->job_total[W.parity & 1u];
          const unsigned long long base = atomicAdd(tot, (unsigned long long)n);
          if (base + n <= (unsigned long long)c.job_cap)
            sh.job_base = (uint32_t)base;
          else // no room: give the range back, the sums of this pass are computed here
            atomicAdd(tot, 0ull - (unsigned long long)n);
        }
      }
      __syncthreads();
    }
    for (uint32_t le = e_lo + (uint32_t)(tid >> 5); le < e_hi; le += NTHREADS / 32) {
      const uint32_t first = sh.head[le];
      if (first == LIST_END) continue;
      // split sums: the entity's jobs become one RUN of descriptors in the block's
      // segment of c.jobs, computed by k_sums; when the segment is full the warp
      // computes them here as without the option
      uint32_t run_base = SIMIAN_NONE, run_len = 0;
      if (c.jobs && sh.job_base != SIMIAN_NONE) {
        for (uint32_t j = first; j != LIST_END; j = sh.nxt[j])
          if (sh.xs[j]) run_len++;
        if (run_len == 0) continue;
        uint32_t base = 0;
        if ((tid & 31) == 0) base = atomicAdd(&sh.job_count, run_len);
        run_base = sh.job_base + __shfl_sync(0xffffffffu, base, 0);
      }
      int last = -1;
      uint32_t emitted_jobs = 0;
#pragma unroll 1
      for (;;) {
        // the next job of this entity (warp-uniform walk of its short list)
        uint32_t bj = LIST_END, bd = 2048u;
        for (uint32_t j = first; j != LIST_END; j = sh.nxt[j]) {
          const unsigned long long x = sh.xs[j];
          if (x) {
            const uint32_t d = JOB_DELTA(x);
            if ((int)d > last && d < bd) bd = d, bj = j;
          }
        }
        if (bj == LIST_END) break;
        last = (int)bd;
        const unsigned long long x = sh.xs[bj];
        const double *list = (const double *)SLOT_HI(sh, bj).y;
        int k;
        const uint32_t gid = gid_of_slot(c, slot0 + le, k);
        const unsigned long long key = simian_rng_key(c.seed, gid, SH_CORE(sh)[le].count + bd);
        if (run_base != SIMIAN_NONE) {
          if ((tid & 31) == 0)
            st_rec_2x16(c.jobs + ((size_t)run_base + emitted_jobs),
                        (unsigned long long)list, key,
                        x & ~(0x7FFull << 45), // (na, nj, di0, sink offset; the commit delta is spent)
                        emitted_jobs == 0 ? (unsigned long long)run_len : 0ull);
          emitted_jobs++;
          continue;
        }
        const double v = warp_weighted_sum(key, list, JOB_NA(x), JOB_NJ(x), JOB_DI0(x));
        if ((tid & 31) == 0) {
          double *sink = const_cast<double *>(list) + JOB_SINK_OFF(x);
          *sink += v;
        }
      }
    }
    __syncthreads();
    PHASE(5);
    ta.min_key = t2.min_key < ta.min_key ? t2.min_key : ta.min_key;
    ta.max_key = t2.max_key > ta.max_key ? t2.max_key : ta.max_key;
    ta.committed += t2.committed;
    ta.emitted += t2.emitted;
    ta.dropped += t2.dropped;
    ta.ties += t2.ties;
    ta.dties += t2.dties;
    ta.appended += t2.appended;
  }
}

// per-thread accumulators -> the CTA's accumulators in shared memory (K1)
template <class SH>
__device__ __forceinline__ void accumulate_block(SH &sh, const ThreadAcc &ta) {
  unsigned long long mn = warp_min_u64(ta.min_key), mx = warp_max_u64(ta.max_key);
  uint32_t s0 = __reduce_add_sync(0xffffffffu, ta.committed);
  uint32_t s1 = __reduce_add_sync(0xffffffffu, ta.emitted);
  uint32_t s2 = __reduce_add_sync(0xffffffffu, ta.dropped);
  uint32_t s3 = __reduce_add_sync(0xffffffffu, ta.ties);
  uint32_t s4 = __reduce_add_sync(0xffffffffu, ta.dties);
  uint32_t s5 = __reduce_add_sync(0xffffffffu, ta.appended);
  if ((threadIdx.x & 31) == 0) {
    if (mn != SIMIAN_KEY_NONE) atomicMin(&sh.min_key, mn);
    if (mx) atomicMax(&sh.max_key, mx);
    if (s0) atomicAdd(&sh.committed, s0);
    if (s1) atomicAdd(&sh.emitted, s1);
    if (s2) atomicAdd(&sh.dropped, s2);
    if (s3) atomicAdd(&sh.ties, s3);
    if (s4) atomicAdd(&sh.dties, s4);
    if (s5) atomicAdd(&sh.appended, s5);
  }
}

// slots [0, m) are all live: this thread's mask
__device__ __forceinline__ uint32_t live_below(uint32_t m) {
  uint32_t live = 0;
#pragma unroll
  for (int u = 0; u < EV_PER_THREAD; u++)
    if (threadIdx.x + u * NTHREADS < m) live |= 1u << u;
  return live;
}

// The uncommon path of a window: the pages of its rows hold more records than
// the CTA has slots.  A filtered scan gives slots to the DUE records only; when
// even those overflow (bursts such as the 16 events per entity all scheduled at
// t = 0) they are partitioned into scratch pages by entity sub-range
// (partition_burst) -- or, when the page list has no room, the entity range is
// split and the same pages are re-scanned -- and the sub-ranges are processed
// one pass each.  Kept out of line, with its own accumulators, so that the
// common single-pass window pays nothing for it.
template <bool BLOBS>
__device__ __noinline__ void filtered_passes(const DevCfg &c, WindowShared &sh, uint32_t slot0,
                                             uint32_t npg_scan) {
  const int tid = threadIdx.x;
  ThreadAcc ta;
  thread_acc_init(ta);
  collect_due<BLOBS>(c, sh, ta, slot0, 0, npg_scan, 0, c.epb, true);
  __syncthreads();
  const uint32_t m0 = sh.total;
  uint32_t m = m0, e_lo = 0, e_hi = c.epb, pg_lo = 0, pg_hi = npg_scan;
  bool part_ok = false;
  if (m0 > NREFS) part_ok = partition_burst(c, sh, slot0, npg_scan);
#pragma unroll 1
  for (;;) {
    if (m > NREFS && !part_ok) { // split the entity range and re-scan the same pages
      if (e_hi - e_lo <= 1) {
        if (tid == 0) set_error(c, SIMIAN_ERR_ENTITY_BURST);
        break;
      }
      __syncthreads();
      if (tid == 0) {
        uint32_t parts = m / (NREFS * 3u / 4u) + 1u, span = e_hi - e_lo;
        if (parts > span) parts = span;
        if (parts > 16u) parts = 16u;
        if (sh.stack_n + parts > RSTACK_MAX) parts = 2;
        for (uint32_t q = parts; q-- > 0;) { // lowest sub-range on top of the stack
          uint32_t *st = &sh.rstack[4 * sh.stack_n];
          st[0] = e_lo + (uint32_t)((unsigned long long)span * q / parts);
          st[1] = e_lo + (uint32_t)((unsigned long long)span * (q + 1) / parts);
          st[2] = pg_lo;
          st[3] = pg_hi;
          sh.stack_n++;
        }
      }
    } else if (m && m <= NREFS) {
      uint32_t lv = live_below(m);
      if (BLOBS) { // payload words hold slots but are not events
#pragma unroll
        for (int u = 0; u < EV_PER_THREAD; u++)
          if ((lv & (1u << u)) &&
              ((uint32_t)(SLOT_HI(sh, tid + u * NTHREADS).x >> 16) & SIMIAN_EVF_CONT))
            lv &= ~(1u << u);
      }
      process_pass<false, false>(c, sh, ta, slot0, lv, e_lo, e_hi);
    }
    part_ok = false; // only the first (whole-block) overflow is partitioned
    __syncthreads();
    if (sh.stack_n == 0) break;
    __syncthreads();
    if (tid == 0) {
      sh.stack_n--;
      sh.pass_lo = sh.rstack[4 * sh.stack_n];
      sh.pass_hi = sh.rstack[4 * sh.stack_n + 1];
      sh.pass_pg_lo = sh.rstack[4 * sh.stack_n + 2];
      sh.pass_pg_hi = sh.rstack[4 * sh.stack_n + 3];
      sh.total = 0;
      sh.cont_head = LIST_END;
    }
    for (uint32_t i = tid; i < c.epb; i += NTHREADS) sh.head[i] = LIST_END;
    __syncthreads();
    e_lo = sh.pass_lo, e_hi = sh.pass_hi;
    pg_lo = sh.pass_pg_lo, pg_hi = sh.pass_pg_hi;
    collect_due<BLOBS>(c, sh, ta, slot0, pg_lo, pg_hi, e_lo, e_hi, false);
    __syncthreads();
    m = sh.total;
  }
  accumulate_block(sh, ta);
}

// Per-window set-up of a CTA: sticky error flag, the block's own free pages
// (written by this block's CTA of the previous window), accumulators, window
// descriptor.  All threads; ends with a block barrier.  Returns the error flag.
__device__ __forceinline__ uint32_t window_cta_begin(const DevCfg &c, WindowShared &sh,
                                                     const WinState *Wg, uint32_t b) {
  const int tid = threadIdx.x;
  if (tid == 0) sh.found = ld_cg(&c.acc->error);
  if (tid < SIMIAN_BC) sh.ax.pgc[tid] = ld_cg(&c.bcache[(size_t)b * SIMIAN_BC + tid]);
  if (tid == 0) {
#ifdef SIMIAN_BULK_SCAN
    mbar_init(&sh.mbar, 1u);
#endif
    sh.t_last = clock64();
    { // L2-coherent copy: in the loop the descriptor was written by another CTA of this launch
      const unsigned long long *src = reinterpret_cast<const unsigned long long *>(Wg);
      unsigned long long *dst = reinterpret_cast<unsigned long long *>(&sh.W);
      for (int q = 0; q < (int)(sizeof(WinState) / 8); q++) dst[q] = ld_cg64(src + q);
    }
    sh.min_key = SIMIAN_KEY_NONE;
    sh.max_key = 0;
    sh.committed = sh.emitted = sh.dropped = sh.ties = sh.dties = sh.appended = 0;
    sh.npg = 0;
    sh.nvalid = 0;
    sh.nfree = 0;
    sh.free_pos = 0;
    sh.is_last = 0;
    sh.job_count = sh.pass_jobs = 0;
    sh.job_base = SIMIAN_NONE;
    sh.ax.far_appends = sh.ax.page_allocs = sh.ax.pgc_take = 0;
    sh.ax.pgc_have = ld_cg(&c.bcount[b]);
    sh.ax.gstack_count = nullptr;
    sh.ax.gstack = nullptr;
  }
  __syncthreads();
  return sh.found;
}

// One lookahead window (or one redistribution pass) for entity block b: K2a
// list -> K2 dequeue -> K3 handlers -> K4 emit -> recycle -> the CTA's results
// into its BlockPart slot.  All threads of the CTA.
template <bool SPLIT, bool BLOBS = false>
__device__ __forceinline__ void window_cta_body(const DevCfg &c, WindowShared &sh, uint32_t b) {
  const int tid = threadIdx.x;
  const WinState &W = sh.W;
  PHASE(0);
  bool staged = false; // SPLIT: this window's events were parked for k_emit

  ThreadAcc ta;
  thread_acc_init(ta);
  uint32_t npg = 0;
  int ncell = 0;

  if (W.redist) {
    if constexpr (SPLIT) { // left to k_window_rest
      if (tid == 0) c.stage_n[b] = STAGE_REST;
      return;
    } else {
      redistribute_block(c, sh, b);
    }
  } else {
    // ---- K2a: list the pages of this block's cells in rows [tb_lo, tb_hi]
    ncell = (int)(W.tb_hi - W.tb_lo + 1);
    if (tid < ncell) {
      long long tb = W.tb_lo + tid;
      uint32_t cell = (uint32_t)(tb & (long long)c.ring_mask) * c.n_blocks + b;
      // row tb_hi may be appended to during this launch: use the snapshot; rows
      // fully below the bound are recycled after this window
      unsigned long long hw = (tb == W.tb_hi) ? c.snap[b] : c.cell_state[cell];
      list_cell_pages(c, sh, (uint32_t)(hw >> 32), (uint32_t)hw,
                      (tb == W.tb_hi) ? 0u : PG_FREE_FLAG);
    }
    // set-up of the first (normally only) pass, under the same barrier
    for (uint32_t i = tid; i < c.epb; i += NTHREADS) sh.head[i] = LIST_END;
    if (tid >= NTHREADS - NBINS) sh.hist[tid - (NTHREADS - NBINS)] = 0;
    if (tid == NTHREADS - 1) {
      sh.total = 0;
      sh.cont_head = LIST_END;
      sh.stack_n = 0;
      sh.pass_lo = 0;
      sh.pass_hi = c.epb;
      sh.pass_pg_lo = 0; // (pass_pg_hi of the first pass = the number of listed pages)
      // histogram bin = entity >> bin_shift: the smallest power of two that
      // covers the block with NBINS bins
      uint32_t sft = 0;
      while (((c.epb - 1u) >> sft) >= (uint32_t)NBINS) sft++;
      sh.bin_shift = sft;
    }
    __syncthreads();
    PHASE(1);
    npg = sh.npg;
    if (npg > PGMAX) {
      if (tid == 0) set_error(c, SIMIAN_ERR_WINDOW_OVERFLOW);
      npg = 0;
    }
    // ---- K2..K4: normally ONE pass over the whole block.  A window whose due
    // events overflow the shared-memory slots (bursts such as the 16 events per
    // entity at t = 0) takes the out-of-line multi-pass path.
    const uint32_t slot0 = b * c.epb;
    const uint32_t nvalid = sh.nvalid;
    if (nvalid == 0) {
      // nothing in the window's rows for this block
    } else if (nvalid <= NREFS) {
      // all records of the window's pages fit the slots: one bulk copy, then
      // every thread links the due ones among its slots
      collect_all(c, sh, npg, nvalid);
#ifndef SIMIAN_BULK_SCAN
      __syncthreads(); // (bulk copies: every thread has waited on the mbarrier itself)
#endif
      const uint32_t live = link_due<BLOBS>(c, sh, ta, slot0, nvalid);
      __syncthreads();
      PHASE(2);
      process_pass<SPLIT, true>(c, sh, ta, slot0, live, 0, c.epb);
      staged = SPLIT;
    } else {
      if constexpr (SPLIT) { // a burst: left to k_window_rest (nothing was touched yet)
        if (tid == 0) c.stage_n[b] = STAGE_REST;
        return;
      } else {
        filtered_passes<BLOBS>(c, sh, slot0, npg);
        npg = sh.npg; // scratch pages of a partition are recycled with the others
        if (npg > PGMAX) npg = 0;
      }
    }
    PHASE(6);
  }
  // ---- K1: CTA reduction of the accumulators into this CTA's BlockPart: the warps add
  // their sums BEFORE the barrier the recycle phase needs anyway, one thread per field
  // sends them off behind it (independent, result-less atomics) -- no barrier of its own
  accumulate_block(sh, ta);
  __syncthreads();
  if (SPLIT && tid == 0) c.stage_n[b] = staged ? sh.total : 0u;
  if (tid >= NTHREADS - 10) {
    BlockPart *p = &c.part[b % SIMIAN_PART_SLOTS];
    switch (tid - (NTHREADS - 10)) {
      case 0: if (sh.min_key != SIMIAN_KEY_NONE) atomicMin(&p->min_key, sh.min_key); break;
      case 1: if (sh.max_key) atomicMax(&p->max_key, sh.max_key); break;
      case 2: if (sh.committed) atomicAdd(&p->committed, sh.committed); break;
      case 3: if (sh.emitted) atomicAdd(&p->emitted, sh.emitted); break;
      case 4: if (sh.dropped) atomicAdd(&p->dropped, sh.dropped); break;
      case 5: if (sh.ties) atomicAdd(&p->ties, sh.ties); break;
      case 6: if (sh.dties) atomicAdd(&p->dties, sh.dties); break;
      case 7: if (sh.appended) atomicAdd(&p->appended, sh.appended); break;
      case 8: if (sh.ax.far_appends) atomicAdd(&p->far_appends, sh.ax.far_appends); break;
      default: if (sh.ax.page_allocs) atomicAdd(&p->page_allocs, sh.ax.page_allocs); break;
    }
  }
  // ---- recycle: the pages of rows [tb_clean, tb_hi) go onto the block's own
  // page stack (what does not fit: to the global ring), heads are cleared.  The
  // stack lives in HBM (bcache[b]); the entries below the ones popped in this
  // launch stay where they are, the recycled pages are written above them.
  {
    const uint32_t have = sh.ax.pgc_have;
    const uint32_t rem = have - (sh.ax.pgc_take < have ? sh.ax.pgc_take : have);
    const uint32_t nfree = W.redist ? 0u : sh.nfree;       // counted while the pages were listed
    const uint32_t room = SIMIAN_BC - rem;                 // free pages the stack still takes
    const uint32_t over = nfree > room ? nfree - room : 0; // the rest goes to the ring
    uint32_t cnt = rem + (nfree - over);                   // stack height after recycling
    // low stack: take pages from the ring for the next window (one atomic)
    const uint32_t need = (cnt < SIMIAN_BC_LOW) ? SIMIAN_BC_FILL - cnt : 0u;
    if (over || need) {
      if (tid == 0) {
        if (over) sh.free_base = atomicAdd(&c.acc->free_tail, (unsigned long long)over);
        if (need) sh.ring_base = atomicAdd(&c.acc->alloc_head, (unsigned long long)need);
      }
      __syncthreads();
    }
    if (!W.redist) {
      uint32_t *stack = c.bcache + (size_t)b * SIMIAN_BC;
      for (uint32_t i = tid; i < npg; i += NTHREADS)
        if (sh.pg_cnt[i] & PG_FREE_FLAG) {
          const uint32_t k = atomicAdd(&sh.free_pos, 1u);
          if (k < nfree - over)
            stack[rem + k] = sh.pg_id[i];
          else if (k < nfree)
            c.free_ring[(sh.free_base + (k - (nfree - over))) % c.n_pages] = sh.pg_id[i];
        }
      if (tid < ncell - 1) {
        long long tb = W.tb_lo + tid;
        c.cell_state[(uint32_t)(tb & (long long)c.ring_mask) * c.n_blocks + b] = CELL_EMPTY;
      }
      // rows the window jumped over (dead records only)
      long long nskip = W.tb_lo - W.tb_clean;
      if (nskip > (long long)c.ring) nskip = c.ring;
      for (long long i = tid; i < nskip; i += NTHREADS) {
        long long tb = W.tb_clean + i;
        if (tb >= W.tb_lo) break;
        free_cell(c, (uint32_t)(tb & (long long)c.ring_mask) * c.n_blocks + b);
      }
    }
    if (need) {
      // ring slots beyond alloc_limit are not valid yet: take only what is
      unsigned long long lim = W.alloc_limit, base = sh.ring_base;
      uint32_t okn = base >= lim ? 0u : (lim - base < need ? (uint32_t)(lim - base) : need);
      if ((uint32_t)tid < okn)
        c.bcache[(size_t)b * SIMIAN_BC + cnt + tid] = c.free_ring[(base + tid) % c.n_pages];
      if (tid == 0 && okn < need)
        atomicAdd(&c.acc->leaked_pages, (unsigned long long)(need - okn));
      cnt += okn;
    }
    if (tid == 0) c.bcount[b] = cnt;
  }

  PHASE(7);
  PHASE(8);
}

// End of a window launch: the last CTA to finish does the LBTS step
// (simian.js:143-147).  fuse_advance: 1 (size == 1): it computes window k+1;
// 2 / 3 (size > 1): this rank's LBTS candidate (3: and publishes it to the peers'
// exchange slots); 0: nothing.  All threads of every CTA of the launch.
template <class SH>
__device__ __forceinline__ void window_tail(const DevCfg &c, SH &sh, uint32_t win_ix,
                                            int fuse_advance) {
  const int tid = threadIdx.x;
  const WinState &W = sh.W;
  if (!fuse_advance) return;
  __threadfence();
  __syncthreads();
  if (tid == 0) sh.is_last = (atomicAdd(&c.acc->ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!sh.is_last) return;
  __threadfence();
  if (fuse_advance >= 2) {
    lbts_candidate(c, W, sh, win_ix, fuse_advance == 3);
    return;
  }
  reduce_parts(c, sh);
  WinState *Wn = &c.win[(win_ix + 1u) & 1u];
  if (W.redist) {
    setup_next_window(c, W, Wn, W.gmin, 0.0, true);
  } else {
    unsigned long long cand, far;
    local_min_candidate(c, W, sh, cand, far);
    double gmin = cand == SIMIAN_KEY_NONE ? c.inf_time : simian_key_time(cand); // simian.js:147
    double farT = far == SIMIAN_KEY_NONE ? __longlong_as_double(0x7FF0000000000000LL)
                                         : simian_key_time(far);
    setup_next_window(c, W, Wn, gmin, farT, false);
  }
  PHASE(9);
}

// One lookahead window (or one redistribution pass) for entity block blockIdx.x.
// fuse_advance: 1 (size == 1): the last CTA to finish computes window k+1;
// 2 / 3 (size > 1): the last CTA computes this rank's LBTS candidate (3: and
// publishes it to the peers' exchange slots); 0: neither.
// pull_blocks > 0 (size > 1, two-ended outboxes): the FIRST pull_blocks CTAs of the
// grid are not entity blocks -- they read the far end of the peers' outboxes of
// the PREVIOUS window over NVLink and append the records to the calendar while
// the other CTAs run this window (those records lie at or above this window's
// bound); their times join the LBTS candidate like everything else a CTA reports.
template <bool BLOBS>
__global__ void __launch_bounds__(NTHREADS, SIMIAN_MINBLOCKS)
    k_window_t(const __grid_constant__ DevCfg c, uint32_t win_ix, int fuse_advance,
               uint32_t pull_blocks) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  WindowShared &sh = *reinterpret_cast<WindowShared *>(smem_raw);
  const int tid = threadIdx.x;
  const bool puller = blockIdx.x < pull_blocks;
  const uint32_t b = puller ? 0u : blockIdx.x - pull_blocks;
  // Programmatic dependent launch: every resident CTA of the previous window
  // has started, so this grid's CTAs may take the SM slots that launch's last
  // wave frees -- and wait here until it has completed and its writes are
  // visible.  (Both calls are no-ops for an ordinary launch.)
  cudaTriggerProgrammaticLaunchCompletion();
  cudaGridDependencySynchronize();
  const WinState *Wg = &c.win[win_ix & 1u];
  if (Wg->done) return; // simian.js:119: enqueued past the end of the run
  if (puller) {
    // (no entity block: only the accumulators and the window descriptor)
    if (tid == 0) {
      sh.found = ld_cg(&c.acc->error);
      sh.W = *Wg;
      sh.min_key = SIMIAN_KEY_NONE;
      sh.max_key = 0;
      sh.committed = sh.emitted = sh.dropped = sh.ties = sh.dties = sh.appended = 0;
      sh.is_last = 0;
      sh.ax.far_appends = sh.ax.page_allocs = sh.ax.pgc_take = sh.ax.pgc_have = 0;
      sh.ax.gstack_count = nullptr;
      sh.ax.gstack = nullptr;
    }
    __syncthreads();
    if (sh.found) return;
    if (win_ix > 0) {
      const WinState &Wp = c.win[(win_ix - 1u) & 1u]; // the window the records were sent in
      const uint32_t bpp = pull_blocks / (uint32_t)(c.size - 1);
      uint32_t napp = 0;
      unsigned long long mn = SIMIAN_KEY_NONE;
      // Beside a redistribution pass the appends keep the OLD recycle frontier (the
      // rows the pass frees are then never aliased by a row appended to) and go to
      // the far row the pass itself fills, not the one it drains.
      pull_records(c, Wp, sh.W.tb_clean, sh.W.redist ? sh.W.far_row ^ 1u : sh.W.far_row,
                   sh.W.alloc_limit, sh.ax, blockIdx.x / bpp, blockIdx.x % bpp, bpp, 2, sh.W.hi,
                   napp, mn);
      napp = __reduce_add_sync(0xffffffffu, napp);
      mn = warp_min_u64(mn);
      if ((tid & 31) == 0) {
        if (napp) atomicAdd(&sh.appended, napp);
        if (mn != SIMIAN_KEY_NONE) atomicMin(&sh.min_key, mn);
      }
    }
    __syncthreads();
    if (tid < 10) { // the CTA's results -> its BlockPart slot (as window_cta_body does)
      BlockPart *p = &c.part[blockIdx.x % SIMIAN_PART_SLOTS];
      switch (tid) {
        case 0: if (sh.min_key != SIMIAN_KEY_NONE) atomicMin(&p->min_key, sh.min_key); break;
        case 7: if (sh.appended) atomicAdd(&p->appended, sh.appended); break;
        case 8: if (sh.ax.far_appends) atomicAdd(&p->far_appends, sh.ax.far_appends); break;
        case 9: if (sh.ax.page_allocs) atomicAdd(&p->page_allocs, sh.ax.page_allocs); break;
        default: break;
      }
    }
  } else {
    if (window_cta_begin(c, sh, Wg, b)) return; // sticky error: CTA-uniform exit
    window_cta_body<false, BLOBS>(c, sh, b);
  }
  window_tail(c, sh, win_ix, fuse_advance);
}
// the two instantiations: the ordinary window kernel and the one for models whose
// events carry payload words (option "payloads")
#define k_window k_window_t<false>
#define k_window_blobs k_window_t<true>

#ifndef SIMIAN_LOOP_TU
// ---- the split window (option "split_handlers"; models of pure handlers) -------
// One window = three launches, chained by programmatic dependent launch:
//   k_window_stage  K2a list -> K2 dequeue -> A1 rank -> park the ranked events in HBM
//                   (c.stage: {time, dst | src << 32, handler | commit index << 16, data},
//                   compacted) -> C trace chain -> recycle.  The common window only:
//                   no handler, no burst path, no redistribution in its call graph, so
//                   it needs 60 registers instead of the 102 the one-kernel window wants.
//   k_window_rest   the one-kernel window (window_cta_body<false>) for the blocks
//                   k_window_stage flagged (STAGE_REST): bursts that overflow the slots,
//                   redistribution passes.  A few CTAs looping over the blocks; they
//                   find nothing to do in a steady window.
//   k_emit          K3 + K4: the handlers of the parked events, their records appended;
//                   then the LBTS step (last CTA).
__global__ void __launch_bounds__(NTHREADS, SIMIAN_STAGE_MINBLOCKS)
    k_window_stage(const __grid_constant__ DevCfg c, uint32_t win_ix) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  WindowShared &sh = *reinterpret_cast<WindowShared *>(smem_raw);
  cudaTriggerProgrammaticLaunchCompletion();
  cudaGridDependencySynchronize();
  const WinState *Wg = &c.win[win_ix & 1u];
  if (Wg->done) return; // simian.js:119
  if (window_cta_begin(c, sh, Wg, blockIdx.x)) return; // sticky error
  window_cta_body<true>(c, sh, blockIdx.x);
}

__global__ void __launch_bounds__(NTHREADS, SIMIAN_MINBLOCKS)
    k_window_rest(const __grid_constant__ DevCfg c, uint32_t win_ix) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  WindowShared &sh = *reinterpret_cast<WindowShared *>(smem_raw);
  cudaTriggerProgrammaticLaunchCompletion();
  cudaGridDependencySynchronize();
  const WinState *Wg = &c.win[win_ix & 1u];
  if (Wg->done) return;
#pragma unroll 1
  for (uint32_t b = blockIdx.x; b < c.n_blocks; b += gridDim.x) {
    if (c.stage_n[b] != STAGE_REST) continue; // (CTA-uniform)
    __syncthreads();                          // every thread has read the flag
    if (threadIdx.x == 0) c.stage_n[b] = 0u;  // nothing parked: k_emit skips the block
    if (window_cta_begin(c, sh, Wg, b)) return; // sticky error
    window_cta_body<false>(c, sh, b);
    __syncthreads(); // the shared struct is reused by the next block
  }
}

#ifndef SIMIAN_EMIT_MINBLOCKS
#define SIMIAN_EMIT_MINBLOCKS 5
#endif
#define EMIT_WARPS (NTHREADS / 32)
#ifndef EMIT_EV
#define EMIT_EV 2                     // events per lane and chunk
#endif
#define EMIT_CHUNK (32 * EMIT_EV)     // parked events a warp takes at a time
#define EMIT_CHUNKS_PER_BLOCK (NREFS / EMIT_CHUNK)
struct __align__(32) EmitShared {
  WinState W;
  unsigned long long min_key, max_key;
  uint32_t committed, emitted, dropped, ties, dties, appended;
  uint32_t found, is_last, far_appends, page_allocs;
  long long scan_j, t_last;
  AppendCtx ax[EMIT_WARPS]; // per warp: the page stack of the block it works for
  // a warp's chunk of parked events; a handler's first local record replaces the
  // event it consumed (staged, then appended: append_staged)
#ifdef SIMIAN_BULK_SCAN
  SlotRec rec[EMIT_WARPS * EMIT_CHUNK];
#else
  ulonglong2 lo[EMIT_WARPS * EMIT_CHUNK];
  ulonglong2 hi[EMIT_WARPS * EMIT_CHUNK];
#endif
};

// k_emit: K3 + K4 of the split window.  A persistent grid of WARPS: every warp takes
// chunks of 64 parked events (static stride over all blocks' chunks), fetches them
// into its own shared-memory slots with cp.async, runs the handlers (simian.js:129-131)
// with the commit indices k_window_stage gave the events -- two per lane -- stages
// what they send in the consumed slots and appends it with both cell claims in
// flight.  No block barrier between the CTA's set-up and its end; a page a claim has
// to install comes from the source block's stack in HBM (an atomic on its height).
// The last CTA does the LBTS step (fuse_advance as in k_window).
__global__ void __launch_bounds__(NTHREADS, SIMIAN_EMIT_MINBLOCKS)
    k_emit(const __grid_constant__ DevCfg c, uint32_t win_ix, int fuse_advance) {
  __shared__ EmitShared sh;
  const int tid = threadIdx.x;
  const uint32_t lane = (uint32_t)tid & 31u, warp = (uint32_t)tid >> 5;
  cudaTriggerProgrammaticLaunchCompletion();
  cudaGridDependencySynchronize();
  const WinState *Wg = &c.win[win_ix & 1u];
  if (Wg->done) return;
  // the CTA's set-up, one independent load per thread
  if (tid < (int)(sizeof(WinState) / 8))
    reinterpret_cast<unsigned long long *>(&sh.W)[tid] =
        reinterpret_cast<const unsigned long long *>(Wg)[tid];
  if (tid == 32) sh.found = ld_cg(&c.acc->error);
  if (tid == 33) {
    sh.min_key = SIMIAN_KEY_NONE;
    sh.max_key = 0;
    sh.committed = sh.emitted = sh.dropped = sh.ties = sh.dties = sh.appended = 0;
    sh.is_last = 0;
    sh.far_appends = sh.page_allocs = 0;
    sh.t_last = clock64();
  }
  if (lane == 0) {
    AppendCtx &a = sh.ax[warp];
    a.far_appends = a.page_allocs = a.pgc_take = a.pgc_have = 0;
    a.gstack_count = nullptr;
    a.gstack = nullptr;
  }
  __syncthreads();
  if (sh.found) return; // sticky error: CTA-uniform exit, the window does not advance
  ThreadAcc ta;
  thread_acc_init(ta);
  AppendCtx &ax = sh.ax[warp];
  const uint32_t slot_base = warp * EMIT_CHUNK + lane;
  const uint32_t n_chunks = c.n_blocks * EMIT_CHUNKS_PER_BLOCK;
  const uint32_t total_warps = gridDim.x * EMIT_WARPS;
  if (!sh.W.redist) {
#pragma unroll 1
    for (uint32_t ch = blockIdx.x * EMIT_WARPS + warp; ch < n_chunks; ch += total_warps) {
      const uint32_t b = ch / EMIT_CHUNKS_PER_BLOCK;
      const uint32_t off = (ch % EMIT_CHUNKS_PER_BLOCK) * EMIT_CHUNK;
      uint32_t n = c.stage_n[b];
      n = n <= (uint32_t)NREFS ? n : 0u; // (STAGE_REST never survives k_window_rest)
      if (off >= n) continue;            // (warp-uniform)
      const simian_event_t *in = c.stage + (size_t)b * NREFS + off;
      const uint32_t m = n - off < (uint32_t)EMIT_CHUNK ? n - off : (uint32_t)EMIT_CHUNK;
      uint32_t live = 0;
#pragma unroll
      for (int u = 0; u < EMIT_EV; u++) {
        const uint32_t j = lane + 32u * u;
        if (j < m) {
          cp_async16(&SLOT_LO(sh, slot_base + 32u * u), in + j);
          cp_async16(&SLOT_HI(sh, slot_base + 32u * u), (const char *)(in + j) + 16);
          live |= 1u << u;
        }
      }
      if (lane == 0) { // the source block's page stack serves this chunk's installs
        ax.gstack_count = c.bcount + b;
        ax.gstack = c.bcache + (size_t)b * SIMIAN_BC;
      }
      cp_async_wait_all(); // (a lane reads only the slots it fetched itself)
      __syncwarp();
#ifdef SIMIAN_EMIT_UNROLL
#pragma unroll
#else
#pragma unroll 1
#endif
      for (int u = 0; u < EMIT_EV; u++) {
        if (!(live & (1u << u))) continue;
        const uint32_t i = slot_base + 32u * u;
        const RecBits r = slot_get(sh, i);
        int k;
        DevCtx<false, true> ctx(c, sh.W, ax, ta);
        ctx.stage_lo = &SLOT_LO(sh, i);
        ctx.stage_hi = &SLOT_HI(sh, i);
        ctx.self_slot = r.dst();
        ctx.self_gid = gid_of_slot(c, ctx.self_slot, k);
        ctx.cl = &c.cls[k];
        ctx.commit_idx = r.q2 >> 16;
        ctx.key = simian_rng_key(c.seed, ctx.self_gid, ctx.commit_idx);
        ctx.n_emit = 0;
        ctx.now_ = r.time(); // simian.js:126
        ctx.handler_ = r.handler();
        ctx.src_ = r.src();
        ctx.data_ = r.data();
        const uint32_t h = r.handler();
        run_pure_handler(ctx, h < SIMIAN_MAX_SERVICES ? ctx.cl->fn_of_service[h] : 0);
        if (ctx.stage_lo) SLOT_LO(sh, i).y = 0xFFFFFFFFFFFFFFFFull; // nothing staged
      }
      append_staged<true, EMIT_EV>(c, sh, ax, sh.W, live, slot_base, 32u, []() {});
      __syncwarp(); // the slots and the stack pointers are reused by the next chunk
    }
  }
  accumulate_block(sh, ta);
  if (lane == 0) {
    if (ax.far_appends) atomicAdd(&sh.far_appends, ax.far_appends);
    if (ax.page_allocs) atomicAdd(&sh.page_allocs, ax.page_allocs);
  }
  __syncthreads();
  if (tid < 10) { // one thread per field: independent, result-less atomics
    BlockPart *p = &c.part[blockIdx.x % SIMIAN_PART_SLOTS];
    switch (tid) {
      case 0: if (sh.min_key != SIMIAN_KEY_NONE) atomicMin(&p->min_key, sh.min_key); break;
      case 3: if (sh.emitted) atomicAdd(&p->emitted, sh.emitted); break;
      case 4: if (sh.dropped) atomicAdd(&p->dropped, sh.dropped); break;
      case 7: if (sh.appended) atomicAdd(&p->appended, sh.appended); break;
      case 8: if (sh.far_appends) atomicAdd(&p->far_appends, sh.far_appends); break;
      case 9: if (sh.page_allocs) atomicAdd(&p->page_allocs, sh.page_allocs); break;
      default: break;
    }
  }
  window_tail(c, sh, win_ix, fuse_advance);
}

// k_sums (option "split_sums"): the weighted subset sums the stateful handlers of a
// window deferred (the LANL receive handler, pdes_lanl_benchmarkV8.js:304-307), one
// WARP per run of an entity's jobs, each sum in the order simian_lanl_weighted_sum
// defines, the runs' values added to the entity's sink in commit order.  The list
// elements a sum reads are contiguous (list[0 .. na)): the warp fetches them into its
// own 4 KB of shared memory with cp.async -- every load of the job in flight at once --
// instead of one dependent HBM round trip per 32 terms (inside k_window, whose shared
// memory is full, 67 % of the stall samples of the sum are that load), 512 elements
// at a time; then the terms are pure arithmetic on 40 registers.
#ifndef SIMIAN_SUMS_MINBLOCKS
#define SIMIAN_SUMS_MINBLOCKS 4
#endif
#define SUMS_CHUNK 512 // list elements a warp stages at a time
#ifndef SIMIAN_SUMS_UNROLL
#define SIMIAN_SUMS_UNROLL 4
#endif

// one job, all 32 lanes, the defined order (simian_lanl_weighted_sum): lane l adds the
// terms k = l, l + 32, ... in increasing k; the list comes through `buf` in chunks
__device__ __forceinline__ double warp_weighted_sum_staged(unsigned long long key,
                                                           const double *list, uint32_t na,
                                                           uint32_t nj, uint32_t di0,
                                                           double *buf) {
  const uint32_t lane = threadIdx.x & 31u;
  double p = 0.0;
  if (na * nj != 0u) {
    const uint32_t qi = 32u / nj, ri = 32u % nj;
#pragma unroll 1
    for (uint32_t i0 = 0; i0 < na; i0 += SUMS_CHUNK) {
      const uint32_t cnt = na - i0 < (uint32_t)SUMS_CHUNK ? na - i0 : (uint32_t)SUMS_CHUNK;
      for (uint32_t q = lane; q < cnt; q += 32u) { // (lists are 8-byte aligned only)
        uint32_t d = (uint32_t)__cvta_generic_to_shared(buf + q);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(list + i0 + q)
                     : "memory");
      }
      cp_async_wait_all();
      __syncwarp();
      const uint32_t k_lo = i0 * nj, k_hi = (i0 + cnt) * nj;
      uint32_t k = k_lo + ((lane - k_lo) & 31u); // first k >= k_lo with k = lane (mod 32)
      uint32_t i = k / nj - i0, j = k % nj;
      constexpr int sums_unroll = SIMIAN_SUMS_UNROLL;
#pragma unroll sums_unroll
      for (; k < k_hi; k += 32u) {
        p += buf[i] * simian_u01(simian_rng_draw(key, di0 + k));
        i += qi;
        j += ri;
        if (j >= nj) j -= nj, i++;
      }
      __syncwarp(); // the chunk is overwritten next
    }
  }
  for (int o = 16; o; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
  return p;
}
__global__ void __launch_bounds__(NTHREADS, SIMIAN_SUMS_MINBLOCKS)
    k_sums(const __grid_constant__ DevCfg c, uint32_t win_ix) {
  __shared__ double sbuf[NTHREADS / 32][SUMS_CHUNK];
  cudaTriggerProgrammaticLaunchCompletion();
  cudaGridDependencySynchronize();
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  // the descriptors window win_ix wrote; the counter of the other parity is cleared for
  // window win_ix + 1 (its last reader, k_sums(win_ix - 1), is long done) -- also when
  // this launch was enqueued past the end of the run: such a launch then finds zero
  unsigned long long total = c.acc->job_total[win_ix & 1u];
  if (blockIdx.x == 0 && threadIdx.x == 0) c.acc->job_total[(win_ix + 1u) & 1u] = 0ull;
  if (total > c.job_cap) total = c.job_cap;
  const uint32_t n = (uint32_t)total;
  // a persistent grid of warps over ALL jobs: an entity block that is a hot spot of the
  // model (geometric destinations) spreads over the whole GPU
  for (uint32_t j0 = blockIdx.x * (NTHREADS / 32) + warp; j0 < n; j0 += gridDim.x * (NTHREADS / 32)) {
    const RecBits h = ld_rec_2x16(c.jobs + j0);
    const uint32_t run = (uint32_t)h.q3;
    if (run == 0) continue; // inside another warp's run
    for (uint32_t r = 0; r < run && j0 + r < n; r++) {
      const RecBits d = r == 0 ? h : ld_rec_2x16(c.jobs + j0 + r);
      const double *list = (const double *)d.q0;
      const double v = warp_weighted_sum_staged(d.q1, list, JOB_NA(d.q2), JOB_NJ(d.q2),
                                                JOB_DI0(d.q2), sbuf[warp]);
      if (lane == 0) {
        double *sink = const_cast<double *>(list) + JOB_SINK_OFF(d.q2);
        *sink += v;
      }
    }
  }
}
#endif

#ifdef SIMIAN_LOOP_TU
// The window loop of simian.js:118-149 INSIDE one launch, for models whose entity
// blocks are all resident at once (gridDim.x = n_blocks <= the CTAs the GPU holds;
// launched cooperatively): a window costs a grid barrier instead of a kernel
// launch.  After each window the last CTA to arrive computes the next window and
// releases the others (acc->loop_seq = windows done so far); the __threadfence()
// every thread then executes also invalidates the SM's L1 (CCTL.IVALL in SASS), so
// what other CTAs wrote in the previous window is read from L2.  size == 1 only.
// Runs at most max_windows windows; the host polls in between.
// Compiled in its own translation unit (window_loop.cu): inside k_window's the
// loop costs the per-window kernel registers (config 2: 15.0 -> 16.6 ms per step).
__global__ void __launch_bounds__(NTHREADS, SIMIAN_MINBLOCKS)
    k_window_loop(const __grid_constant__ DevCfg c, uint32_t win_ix0, uint32_t max_windows) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  WindowShared &sh = *reinterpret_cast<WindowShared *>(smem_raw);
  const int tid = threadIdx.x;
  const uint32_t b = blockIdx.x;
#pragma unroll 1
  for (uint32_t w = 0; w < max_windows; w++) {
    const uint32_t win_ix = win_ix0 + w;
    const WinState *Wg = &c.win[win_ix & 1u];
    // (the descriptor was written by another CTA of this launch: read through L2)
    if (ld_cg(&Wg->done)) return; // simian.js:119
    if (window_cta_begin(c, sh, Wg, b)) return; // sticky error: CTA-uniform exit
    window_cta_body<false>(c, sh, b);
    // ---- grid barrier; the last CTA to arrive computes the next window
    __threadfence();
    __syncthreads();
    if (tid == 0) sh.is_last = (atomicAdd(&c.acc->ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (sh.is_last) {
      __threadfence();
      reduce_parts(c, sh);
      const WinState &W = sh.W;
      WinState *Wn = &c.win[(win_ix + 1u) & 1u];
      if (W.redist) {
        setup_next_window(c, W, Wn, W.gmin, 0.0, true);
      } else {
        unsigned long long cand, far;
        local_min_candidate(c, W, sh, cand, far);
        double gmin = cand == SIMIAN_KEY_NONE ? c.inf_time : simian_key_time(cand); // simian.js:147
        double farT = far == SIMIAN_KEY_NONE ? __longlong_as_double(0x7FF0000000000000LL)
                                             : simian_key_time(far);
        setup_next_window(c, W, Wn, gmin, farT, false);
      }
      __syncthreads();
      if (tid == 0) {
        __threadfence();
        atomicExch(&c.acc->loop_seq, win_ix + 1u); // release the other CTAs
      }
    } else if (tid == 0) {
      uint32_t spins = 0;
      while (ld_cg(&c.acc->loop_seq) != win_ix + 1u) {
        if (ld_cg(&c.acc->error)) break;
        if (++spins > (1u << 26)) { // seconds: a CTA of this grid never became resident
          set_error(c, SIMIAN_ERR_STALLED);
          break;
        }
        __nanosleep(20);
      }
    }
    __syncthreads();
    __threadfence(); // (also drops the SM's L1 lines: the next window reads L2)
  }
}
#endif

// size == 1, option "fused_advance" = 0: the LBTS step of window win_ix (what the last CTA
// of k_window does otherwise: simian.js:143-147) as a launch of its own, one CTA.  The
// CTAs of k_window then end without the fence + ticket + barrier of the last-CTA
// pattern -- a fixed cost per CTA that counts when a window is many waves of short CTAs.
struct __align__(16) LbtsShared {
  WinState W;
  unsigned long long min_key;
  long long scan_j, t_last;
};
__global__ void __launch_bounds__(NTHREADS) k_lbts(const __grid_constant__ DevCfg c, uint32_t win_ix) {
  __shared__ LbtsShared sh;
  cudaTriggerProgrammaticLaunchCompletion();
  cudaGridDependencySynchronize();
  const WinState *Wg = &c.win[win_ix & 1u];
  if (Wg->done) return;
  if (ld_cg(&c.acc->error)) return; // sticky error: the window does not advance
  if (threadIdx.x < (int)(sizeof(WinState) / 8))
    reinterpret_cast<unsigned long long *>(&sh.W)[threadIdx.x] =
        reinterpret_cast<const unsigned long long *>(Wg)[threadIdx.x];
  if (threadIdx.x == 0) sh.t_last = clock64();
  __syncthreads();
  const WinState &W = sh.W;
  reduce_parts(c, sh);
  WinState *Wn = &c.win[(win_ix + 1u) & 1u];
  if (W.redist) {
    setup_next_window(c, W, Wn, W.gmin, 0.0, true);
  } else {
    unsigned long long cand, far;
    local_min_candidate(c, W, sh, cand, far);
    double gmin = cand == SIMIAN_KEY_NONE ? c.inf_time : simian_key_time(cand); // simian.js:147
    double farT = far == SIMIAN_KEY_NONE ? __longlong_as_double(0x7FF0000000000000LL)
                                         : simian_key_time(far);
    setup_next_window(c, W, Wn, gmin, farT, false);
  }
}

__global__ void __launch_bounds__(NTHREADS) k_local_min(const __grid_constant__ DevCfg c, uint32_t win_ix,
                                                        int publish) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  WindowShared &sh = *reinterpret_cast<WindowShared *>(smem_raw);
  lbts_candidate(c, c.win[win_ix & 1u], sh, win_ix, publish);
}

// size > 1: consume the allreduced {globalMinLeft, min far} (simian.js:144)
__global__ void __launch_bounds__(NTHREADS) k_advance(const __grid_constant__ DevCfg c, uint32_t win_ix,
                                                      int from_xch) {
  cudaTriggerProgrammaticLaunchCompletion();
  cudaGridDependencySynchronize();
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
    setup_next_window(c, W, Wn, W.gmin, 0.0, true);
  else
    setup_next_window(c, W, Wn, g0, g1, false);
  // The outbox the NEXT window fills was sent two windows ago; every peer has
  // read it (its pull precedes the allreduce this advance follows).  The one
  // just sent may still be being read: it is left alone.
  if (threadIdx.x < c.size && c.outbox_count) {
    const uint32_t j = ((W.parity ^ 1u) * (uint32_t)c.size) + threadIdx.x;
    c.outbox_count[j] = c.outbox_count[SIMIAN_FAR_COUNT_OFF + j] = 0;
  }
}

// first window: gmin = startTime even when the queue is empty (simian.js:118)
__global__ void __launch_bounds__(NTHREADS) k_begin(const __grid_constant__ DevCfg c) {
  WinState W = c.win[0];
  __shared__ WinState w0;
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
                    false);
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
