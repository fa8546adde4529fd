// device_types.cuh -- HBM-resident data layout of the simian-b200 engine.
//
// Replaces (paths relative to /root/reference):
//   SimianJS/eventQ.js        one binary heap of boxed event objects per rank
//   SimianJS/simian.js:53-59  entities{} / eventList[] / eventQueue
// with
//   * a paged CALENDAR in HBM: cell (time sub-bucket, entity block) -> chain of
//     4-KB pages of 32-byte POD event records (one DRAM sector per record);
//   * per-entity core state {commit count, trace hash} (16 B) + per-class user
//     state, indexed by local slot;
//   * a double-buffered device-resident window descriptor (the LBTS state of
//     simian.js:118-149) so the host never has to see a timestamp.
#pragma once
#include <stdint.h>

#include "../../include/simian_models.h"
#include "../../include/simian_spec.h"

#ifndef SIMIAN_EPB
#define SIMIAN_EPB 512 // entities per block by default (one CTA owns one block)
#endif
#ifndef SIMIAN_EPB_MAX
#define SIMIAN_EPB_MAX 1024 // option "block_entities": models with few due events per window
#endif
#ifndef SIMIAN_THREADS
#define SIMIAN_THREADS 256 // threads per CTA of the window kernel
#endif
#ifndef SIMIAN_PAGE
#define SIMIAN_PAGE 128 // records per page (4 KB)
#endif
#ifndef SIMIAN_REFS
#define SIMIAN_REFS (4 * SIMIAN_THREADS) // due events a CTA holds on chip per pass
#endif
#ifndef SIMIAN_PGMAX
#define SIMIAN_PGMAX 640 // pages a CTA can list per window
#endif
#ifndef SIMIAN_BC
#define SIMIAN_BC 64 // free pages an entity block keeps for itself (block page cache)
#endif
#define SIMIAN_BC_FILL 32 // refill target / initial fill of a block page cache
#define SIMIAN_BC_LOW 8   // refill from the global ring below this

#ifndef SIMIAN_MINBLOCKS
#define SIMIAN_MINBLOCKS 3 // __launch_bounds__ min CTAs/SM of the window kernel
#endif

#ifndef SIMIAN_STAGE_MINBLOCKS
#define SIMIAN_STAGE_MINBLOCKS 3 // min CTAs/SM of k_window_stage (the split window)
#endif

#define SIMIAN_MAX_RANKS 16 // GPUs of one box (peer memory over NVLink)
#define SIMIAN_FAR_COUNT_OFF 64 // uint32 offset of the far-end counters behind outbox_count
#define SIMIAN_MAX_CLASSES 8
#define SIMIAN_MAX_SERVICES 32
#define SIMIAN_MAX_CELLS 96 // sub-buckets one window may span
#define SIMIAN_NONE 0xFFFFFFFFu
#define SIMIAN_KEY_NONE 0xFFFFFFFFFFFFFFFFull
#define SIMIAN_SELF_MAX 8 // same-window self-sends pending per entity

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
  uint8_t fn_of_service[SIMIAN_MAX_SERVICES]; // service id -> builtin fn
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
  unsigned int loop_seq; // k_window_loop: windows completed in the current launch sequence
  unsigned int pad_jobs_;
  // split sums: descriptors written to c.jobs by window k (k & 1); k_sums(k) clears the other
  unsigned long long job_total[2];
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
  int all_pure; // every attached handler is stateless: thread-per-event path
  int blobs;    // option "payloads": events may carry continuation records (SIMIAN_EVF_CONT)
  DevClass cls[SIMIAN_MAX_CLASSES];
  // entities
  EntityCore *core;
  uint32_t n_slots, n_blocks;
  // entities per block, <= SIMIAN_EPB: chosen at setup so that the grid of the
  // window kernel fills whole waves of resident CTAs (epb_inv = ceil(2^64/epb))
  uint32_t epb, pad3_;
  unsigned long long epb_inv;
  // calendar
  double inv_w;
  uint32_t ring, ring_mask;
  // cell head word = (head page << 32) | records claimed in the head page;
  // [(ring+2) * n_blocks], row-major by sub-bucket slot; the last two rows are
  // the far-future rows
  unsigned long long *cell_state;
  simian_event_t *pool;
  uint32_t *page_next; // older page of the same cell (always full)
  uint32_t *free_ring;
  uint32_t n_pages, n_ids;
  unsigned long long *snap; // [n_blocks] head words of row tb_hi at window start
  WinState *win;                  // [2]
  DevAcc *acc;
  BlockPart *part; // [SIMIAN_PART_SLOTS]
  // block page caches: pages an entity block recycled and will reuse, so that
  // steady-state windows never touch the global free ring
  uint32_t *bcache; // [n_blocks][SIMIAN_BC]
  uint32_t *bcount; // [n_blocks]
  // cross-GPU staging (size > 1), double buffered by window parity: the
  // records window k sends stay readable by the peers until window k+2
  simian_event_t *outbox; // [2][size][outbox_cap]
  uint32_t *outbox_count; // [2][size]
  uint32_t outbox_cap;
  // two-ended outboxes (deferred pull): far records are staged from the back of a
  // box, counted in outbox_count[SIMIAN_FAR_COUNT_OFF + ...]
  uint32_t split_outbox;
  // the peers' outboxes, mapped over NVLink (CUDA IPC); null = not connected
  const simian_event_t *peer_outbox[SIMIAN_MAX_RANKS];
  const uint32_t *peer_count[SIMIAN_MAX_RANKS];
  // LBTS exchange slots: every rank WRITES its {minLeft, min far time, window
  // number} into slot [parity][its rank] of every peer's array (and its own)
  unsigned long long run_nonce;          // same on every rank, new for every engine: flags of an
                                         // earlier run over the same (cached) buffers never match
  MinSlot *xch;                          // [2][size], local
  MinSlot *peer_xch[SIMIAN_MAX_RANKS];   // peers' arrays (peer_xch[rank] == xch)
  // split window (option "split_handlers", pure handlers only): k_window_stage ranks a
  // block's due events and parks them here -- {time, dst | src << 32, handler |
  // commit index << 16, data}, compacted -- and k_emit runs their handlers and appends
  // what they send at its own (higher) occupancy.  null: one kernel does both.
  simian_event_t *stage; // [n_blocks][SIMIAN_REFS]
  uint32_t *stage_n;     // [n_blocks] records parked by the block in this window
  // deferred sums in a kernel of their own (option "split_sums", stateful handlers):
  // k_window writes the jobs its handlers parked (simian_defer_weighted_sum) as 32-byte
  // descriptors -- {list pointer, rng key, na | nj << 20 | di0 << 40 | sink offset << 56,
  // run length (first job of an entity's run) or 0} -- and k_sums computes them at its
  // own, higher occupancy.  null: the window kernel computes them itself.
  simian_event_t *jobs; // [job_cap]: ONE array for all blocks (k_sums balances over it);
                        // a CTA reserves the range of a pass with one atomic on acc->job_total
  uint32_t job_cap, pad4_;
};
