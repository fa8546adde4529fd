// device_types.cuh -- HBM-resident data layout of the simian-b200 engine.
//
// Replaces (paths relative to /root/reference):
//   SimianJS/eventQ.js        one binary heap of boxed event objects per rank
//   SimianJS/simian.js:53-59  entities{} / eventList[] / eventQueue
// with
//   * a paged CALENDAR in HBM: cell (entity group, time sub-bucket) -> chain of
//     512-byte pages of 32-byte POD event records (one DRAM sector per record);
//     an entity GROUP is 64 consecutive local entities, owned by ONE WARP of
//     the window kernel;
//   * per-entity core state {commit count, trace hash} (16 B) + per-class user
//     state, indexed by local slot;
//   * a double-buffered device-resident window descriptor (the LBTS state of
//     simian.js:118-149) so the host never has to see a timestamp.
#pragma once
#include <stdint.h>

#include "../../include/simian_models.h"
#include "../../include/simian_spec.h"

#ifndef SIMIAN_GE_SHIFT
#define SIMIAN_GE_SHIFT 6 // log2(entities per group)
#endif
#define SIMIAN_GE (1u << SIMIAN_GE_SHIFT) // entities per group (one warp owns one group)
static_assert(SIMIAN_GE == 64, "a lane of the owning warp handles entities lane and lane + 32");
#ifndef SIMIAN_WPC
#define SIMIAN_WPC 4 // default warps (= entity groups) per CTA of the window kernel
#endif
#define SIMIAN_WPC_MAX 8
#define SIMIAN_THREADS (32 * SIMIAN_WPC_MAX) // max threads per CTA of the window kernel
#ifndef SIMIAN_PAGE
#define SIMIAN_PAGE 16 // records per page (512 B)
#endif
#ifndef SIMIAN_NSLOT
#define SIMIAN_NSLOT 144 // records a warp holds on chip per pass
#endif
#ifndef SIMIAN_PGMAX
#define SIMIAN_PGMAX 96 // pages a warp lists on chip; longer lists are walked in chunks
#endif
#ifndef SIMIAN_BC
#define SIMIAN_BC 32 // free pages an entity group keeps for itself (group page cache)
#endif
#define SIMIAN_BC_ROW (SIMIAN_BC + 4) // words per cache row in HBM: pages, then the count
#define SIMIAN_BC_FILL 16 // refill target / initial fill of a group page cache
#define SIMIAN_BC_LOW 4   // refill from the global ring below this

#ifndef SIMIAN_MINBLOCKS
#define SIMIAN_MINBLOCKS 3 // __launch_bounds__(256, 3): 80 registers; 6 CTAs of 4 warps per SM
#endif

#define SIMIAN_MAX_RANKS 16 // GPUs of one box (peer memory over NVLink)
#define SIMIAN_MAX_CLASSES 8
#define SIMIAN_MAX_SERVICES 32
#define SIMIAN_MAX_CELLS 96 // sub-buckets one window may span
#define SIMIAN_NONE 0xFFFFFFFFu
#define SIMIAN_KEY_NONE 0xFFFFFFFFFFFFFFFFull
#define SIMIAN_SELF_MAX 8 // same-window self-sends pending per entity
// Time word of a record slot nobody has written in the page's current life.
// FREE-PAGE INVARIANT: every record slot of a free page holds a time below the
// lower bound of every window that can ever read the page again (-inf after
// k_init and after a far-row page is released; otherwise the time of a record
// a past window consumed).  A reader may therefore copy a page beyond the
// records that were complete when the launch began: what it finds there is
// either such a dead record or one being appended right now (time >= the
// window's upper bound) -- never a due one.
#define SIMIAN_TIME_DEAD 0xFFF0000000000000ull

struct __align__(16) EntityCore {
  uint64_t count; // committed events                (numEvents, per entity)
  uint64_t hash;  // order-sensitive trace hash
};

struct DevClass {
  uint32_t first_id;  // global id of instance 0
  uint32_t count;     // instances over all ranks
  uint32_t rank_off;  // (rank - baseRank) mod size: first local num
  uint32_t slot_base; // first local slot of this class on this rank
  uint32_t n_local;   // local instances
  uint32_t base_rank; // getBaseRank(name)                 simian.js:192-194
  uint32_t state_bytes;
  uint32_t params_bytes;
  uint8_t *state;     // n_local * state_bytes
  const void *params; // model parameters (device copy)
  uint8_t fn_of_service[SIMIAN_MAX_SERVICES]; // service id -> device function id
};

// LBTS / window descriptor (simian.js:118-120,143-147), device resident.
struct WinState {
  double gmin; // globalMinLeft: every event below it is committed
  double hi;   // epoch = gmin + minDelay; events with time < hi run now
  long long tb_lo, tb_hi; // calendar sub-buckets the window touches
  long long tb_clean;     // every sub-bucket below it has been recycled
  long long tb_clean_new; // redistribution pass: recycle up to here first
  uint32_t done;          // gmin > endTime                  simian.js:119
  uint32_t redist;        // this launch is a far-list redistribution pass
  uint32_t far_row;       // which of the two far rows receives appends
  uint32_t parity;        // window index & 1: which outbox this window fills
  unsigned long long alloc_limit; // pages allocatable in this launch
};
static_assert(sizeof(WinState) == 72, "copied into shared memory as nine 8-byte words");

// Global accumulators (atomics), reset by the advance step.
struct DevAcc {
  unsigned long long min_key;    // min time key of leftover/emitted/ingested
  unsigned long long far_min[2]; // min time key in far row 0 / 1
  unsigned long long alloc_head, free_tail; // page ring cursors
  unsigned long long committed, emitted, dropped, ties, dist_ties;
  unsigned long long max_key; // max committed time key
  unsigned long long windows, redistributions, far_appends, page_allocs,
      leaked_pages;
  unsigned long long appended; // records put into the local calendar, ever
  unsigned int ticket;
  unsigned int error; // sticky, first error wins
  unsigned int window_index;
  unsigned int loop_arrive; // persistent window loop: CTAs that finished the current window
  double minvec[2]; // {local minLeft candidate, local far min} for allreduce
  unsigned long long phase_cycles[12]; // SIMIAN_PROFILE_PHASES builds only
};

// results of one k_window launch: every CTA adds into slot (block index mod
// SIMIAN_PART_SLOTS) with result-less atomics -- 64 cache lines instead of one
// -- and the last CTA sums the slots in a single round trip
#define SIMIAN_PART_SLOTS 64
struct __align__(64) BlockPart {
  unsigned long long min_key, max_key;
  uint32_t committed, emitted, dropped, ties, dties, appended, far_appends, page_allocs;
};

// one rank's LBTS candidate as published to a peer (allreduce(MIN) over NVLink
// peer memory, mpi.cpp:364-390): data first, then seq = window number + 1
struct __align__(32) MinSlot {
  double min_left, far_min;
  unsigned long long seq;
  unsigned long long pad_;
};

struct DevCfg {
  double start, end, min_delay, inf_time;
  uint64_t seed;
  int rank, size;
  int n_classes, tie_check;
  int all_pure, pad2_; // every attached handler is stateless: thread-per-event path
  DevClass cls[SIMIAN_MAX_CLASSES];
  // entities: core[] is padded to whole groups
  EntityCore *core;
  uint32_t n_slots, n_groups;
  // calendar
  double inv_w;
  uint32_t ring, ring_mask;
  uint32_t rows, pad3_; // ring + 2: head words per group (the last two are the far-future rows)
  // cell head word = (head page << 32) | older-page bit << 31 | records claimed
  // in the head page; [n_groups][rows]: the sub-buckets of one group are
  // contiguous, so the warp that owns the group lists a window's cells from a
  // few sectors
  unsigned long long *cell_state;
  simian_event_t *pool;
  uint32_t *page_next; // older page of the same cell (always full); NONE while the page
                       // is free or has no older page
  uint32_t *free_ring;
  uint32_t n_pages, n_ids;
  // one bit per ring row: some group has (had) a page in that row since the row
  // was last recycled; the sparse-case LBTS search reads ring/32 words instead of
  // every head word
  uint32_t *row_bits; // [ring / 32]
  WinState *win;      // [2]
  DevAcc *acc;
  BlockPart *part; // [SIMIAN_PART_SLOTS]
  // group page caches: pages an entity group recycled and will reuse, so that
  // steady-state windows never touch the global free ring.  Row g =
  // {SIMIAN_BC page ids (a stack), count, 3 pad words}: one 144-byte bulk copy
  uint32_t *bcache; // [n_groups][SIMIAN_BC_ROW]
  // cross-GPU staging (size > 1), double buffered by window parity: the
  // records window k sends stay readable by the peers until window k+2
  simian_event_t *outbox; // [2][size][outbox_cap]
  uint32_t *outbox_count; // [2][size]
  uint32_t outbox_cap, pad1_;
  // the peers' outboxes, mapped over NVLink (CUDA IPC); null = not connected
  const simian_event_t *peer_outbox[SIMIAN_MAX_RANKS];
  const uint32_t *peer_count[SIMIAN_MAX_RANKS];
  // LBTS exchange slots: every rank WRITES its {minLeft, min far time, window
  // number} into slot [parity][its rank] of every peer's array (and its own)
  unsigned long long run_nonce;          // same on every rank, new for every engine: flags of an
                                         // earlier run over the same (cached) buffers never match
  MinSlot *xch;                          // [2][size], local
  MinSlot *peer_xch[SIMIAN_MAX_RANKS];   // peers' arrays (peer_xch[rank] == xch)
};
