"""Host-side window driver for size > 1: one process per GPU, the collectives
of the reference's MPI loop done with torch.distributed (NCCL over NVLink on
GPUs, gloo in the CPU tests).

Reference protocol replaced (SimianJS/simian.js:136-145, MasalaChai/mpi.cpp):
  alltoallSum()  MPI_Alltoall of per-peer send counts  mpi.cpp:418-437
      -> all_gather of the `size` outbox counters
  sendAndCount / probe / recv  one blocking MPI_Send of a msgpack blob per
  remote event + MPI_Probe/MPI_Recv per event           mpi.cpp:439-476,318-362
      -> ONE alltoallv of fixed 32-byte POD records per window
  allreduce(minLeft, MIN)  MPI_Allreduce on one double   mpi.cpp:364-390
      -> all_reduce(MIN) of {minLeft, min far-future time} straight from the
         engine's device buffer
The engine side of each step is a C-ABI call (simian_window_*), see
include/simian_b200.h.  Nothing here touches event payloads on the host.

RankEngine protocol (what drive_windows needs from a rank):
  begin(); process(); outbox() -> (counts, [per-peer tensor views (n,4) int64]);
  new_inbox(n) -> tensor; ingest(inbox, n); local_min() -> float64[2] tensor;
  advance() -> done; end() -> stats dict
"""
import torch
import torch.distributed as dist


class _DevBuf:
    """zero-copy torch view of engine-owned device memory"""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {
            "shape": (nbytes,), "typestr": "|u1", "data": (ptr, False),
            "version": 2}


def _device_view(ptr, nbytes, dtype, device):
    t = torch.as_tensor(_DevBuf(ptr, nbytes), device=device)
    return t.view(dtype)


class GpuRankEngine:
    """simian_b200.capi.Engine (one GPU = one rank) behind the RankEngine
    protocol; every method is one C-ABI call plus zero-copy tensor views."""

    def __init__(self, engine, device):
        self.e, self.device = engine, torch.device(device)
        self.size = engine.size
        self._outbox = {}  # device address -> view (double buffered by window parity)
        self._minvec = None
        # engine kernels and NCCL collectives are ordered on torch's stream
        engine.set_stream(torch.cuda.current_stream(self.device).cuda_stream)

    def begin(self):
        self.e.run_begin()

    def process(self):
        self.e.window_process()

    def outbox(self):
        counts, ptr, cap = self.e.window_outbox()  # D2H of `size` counters
        box = self._outbox.get(ptr)
        if box is None:
            box = self._outbox[ptr] = _device_view(
                ptr, self.size * cap * 32, torch.int64, self.device).view(self.size, cap, 4)
        return counts, [box[p, :counts[p]] for p in range(self.size)]

    def new_inbox(self, n):
        if getattr(self, "_inbox", None) is None or self._inbox.shape[0] < n:
            self._inbox = torch.empty((max(n, 1024) * 5 // 4, 4), dtype=torch.int64,
                                      device=self.device)
        return self._inbox

    def ingest(self, inbox, n):
        if n:
            self.e.window_ingest(inbox.data_ptr(), n)

    def local_min(self):
        ptr = self.e.window_local_min()
        if self._minvec is None:
            self._minvec = _device_view(ptr, 16, torch.float64, self.device)
        return self._minvec

    def advance(self):
        return self.e.window_advance()

    def end(self):
        return self.e.run_end()


def connect_peers(engine, rank, size, group=None):
    """Publish this rank's outbox (CUDA-IPC handles) and map every peer's:
    one all_gather of a ~150-byte blob per rank, once per engine."""
    blob = engine.ipc_export()
    blobs = [None] * size
    dist.all_gather_object(blobs, blob, group=group)
    for p in range(size):
        if p != rank:
            engine.ipc_import(p, blobs[p])


def drive_windows_p2p(engine, rank, size, device, group=None, batch=8,
                      max_batch=64, profile=False, allreduce="peer"):
    """Simian.run for size > 1 (simian.js:118-149) with the device-driven
    exchange: no host round trip inside a batch of windows.  Per window:
      k_window           events below the bound; remote sends -> outbox[parity],
                         and into this rank's LBTS candidate       :122-134, entity.js:66
        (its last CTA)   minLeft candidate                          :143
      all_reduce(MIN)    {minLeft, min far time}; orders the pull   :144 (mpi.cpp:380)
      k_pull             peers' outboxes -> local calendar, NVLink  :137-142
      k_advance          next window                                :119-120
    The host polls `done`/error once per batch; windows launched past the end
    are no-ops on every rank (the window sequence is global).
    allreduce="peer" (default): the MIN-allreduce itself runs over NVLink peer
    memory -- the last CTA of k_window writes this rank's candidate into every
    peer's exchange slots, k_pull waits for all of them, k_advance takes the
    minimum -- so a window is three kernel launches and no collective-library call;
    allreduce="nccl": torch.distributed all_reduce as drawn above.
    Engines created with option deferred_pull = 1 stage the records that cannot be
    due in the receiver's next window from the back of their outboxes; k_pull then
    takes only the front end, and the first CTAs of the NEXT k_window launch pull
    the rest over NVLink beside that window's entity blocks (same three launches)."""
    engine.set_stream(torch.cuda.current_stream(device).cuda_stream)
    connect_peers(engine, rank, size, group)
    engine.run_begin()
    ptr = engine.window_local_min_ptr()
    mv = _device_view(ptr, 16, torch.float64, device)
    windows = 0
    marks = []  # profile: CUDA events around every step of every window

    def mark():
        if profile:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            marks.append(ev)

    while True:
        for _ in range(batch):
            mark()
            if allreduce == "peer":
                engine.window_process_lbts(True)   # k_window + candidate + publish
                mark()
                mark()
                mark()
                engine.window_pull(True)
                mark()
                engine.window_advance_async(True)
            else:
                engine.window_process_lbts(False)  # k_window + candidate
                mark()
                mark()
                dist.all_reduce(mv, op=dist.ReduceOp.MIN, group=group)
                mark()
                engine.window_pull(False)
                mark()
                engine.window_advance_async(False)
            mark()
            windows += 1
        if engine.poll():
            break
        batch = min(2 * batch, max_batch)
    st = engine.run_end()
    # This rank is done (poll synchronised its stream), but a peer may still be
    # pulling the last window's records out of this rank's outbox: nobody may
    # tear its engine down -- or build the next one over the same buffers --
    # before every rank got here (MPI_Barrier at the end of run, simian.js:152).
    dist.barrier(group=group)
    st["driver_windows"] = windows
    if profile:
        names = ["k_window", "k_local_min", "all_reduce", "k_pull", "k_advance"]
        tot = dict.fromkeys(names, 0.0)
        first = dict.fromkeys(names, 0.0)
        for w in range(windows):
            for j, nm in enumerate(names):
                ms = marks[6 * w + j].elapsed_time(marks[6 * w + j + 1])
                tot[nm] += ms
                if w == 0:
                    first[nm] = ms
        st["phase_ms"] = tot
        st["phase_ms_first_window"] = first
    return st


def _exchange(rank, size, sends, recv_counts, inbox, group=None):
    """alltoallv of (n,4) int64 record tensors; returns #records received."""
    offs, total = [], 0
    for p in range(size):
        offs.append(total)
        total += recv_counts[p]
    if dist.get_backend(group) == "nccl":
        outs = [inbox[offs[p]:offs[p] + recv_counts[p]] for p in range(size)]
        dist.all_to_all(outs, [s.contiguous() for s in sends], group=group)
    else:  # gloo has no all_to_all: point-to-point pairs
        ops = []
        for p in range(size):
            if p == rank:
                continue
            if sends[p].shape[0]:
                ops.append(dist.P2POp(dist.isend, sends[p].contiguous(), p, group))
            if recv_counts[p]:
                ops.append(dist.P2POp(dist.irecv, inbox[offs[p]:offs[p] + recv_counts[p]],
                                      p, group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
    return total


def drive_windows(eng, rank, size, group=None, max_windows=None):
    """Simian.run for size > 1 (simian.js:118-149): window loop with the
    per-window exchange and min-allreduce.  Returns the rank's stats dict with
    the window count."""
    eng.begin()
    windows = 0
    while True:
        eng.process()                                   # simian.js:122-134
        counts, sends = eng.outbox()
        counts[rank] = 0
        dev = sends[0].device
        mine = torch.tensor(counts, dtype=torch.int64, device=dev)
        table = [torch.empty_like(mine) for _ in range(size)]
        dist.all_gather(table, mine, group=group)       # alltoallSum, :137
        recv_counts = [int(x) for x in torch.stack(table)[:, rank].tolist()]
        recv_counts[rank] = 0
        n_in = sum(recv_counts)
        inbox = eng.new_inbox(n_in)
        _exchange(rank, size, sends, recv_counts, inbox, group)   # :138-142
        eng.ingest(inbox, n_in)
        mv = eng.local_min()                            # :143
        dist.all_reduce(mv, op=dist.ReduceOp.MIN, group=group)    # :144
        windows += 1
        if eng.advance():                               # :119
            break
        if max_windows and windows >= max_windows:
            break
    st = eng.end()
    st["driver_windows"] = windows
    return st


def drive_windows_loopback(ranks, max_windows=None):
    """The same window loop for ALL ranks of a world inside ONE process (the
    ranks' engines may share one GPU): the exchange and the min-reduction of
    simian.js:136-145 are done with tensor copies instead of collectives.  Used
    to test the multi-rank engine paths (outboxes, ingest, LBTS candidate,
    advance) where a second GPU is not available.  `ranks`: RankEngine objects,
    index = rank.  Returns the per-rank stats dicts."""
    size = len(ranks)
    for r in ranks:
        r.begin()
    windows = 0
    while True:
        for r in ranks:
            r.process()                                     # simian.js:122-134
        boxes = [r.outbox()[1] for r in ranks]              # [src][dst] -> records
        for dst, r in enumerate(ranks):                     # :137-142
            parts = [boxes[src][dst] for src in range(size) if src != dst]
            n_in = sum(int(p.shape[0]) for p in parts)
            inbox = r.new_inbox(n_in)
            if n_in:
                torch.cat(parts, out=inbox[:n_in])
            r.ingest(inbox, n_in)
        mvs = [r.local_min() for r in ranks]                # :143
        g = torch.stack([m.to(mvs[0].device) for m in mvs]).min(dim=0).values
        for m in mvs:                                       # :144
            m.copy_(g)
        windows += 1
        dones = [r.advance() for r in ranks]                # :119
        if all(dones):
            break
        if max_windows and windows >= max_windows:
            break
    out = []
    for r in ranks:
        st = r.end()
        st["driver_windows"] = windows
        out.append(st)
    return out


def drive_windows_p2p_loopback(engines, device, batch=8, max_batch=64):
    """drive_windows_p2p for ALL ranks of a world inside ONE process on ONE
    device: the same three kernels per rank and window -- k_window (fused LBTS
    candidate, published into every peer's exchange slots), k_pull (flags, then
    the peers' outboxes -> calendar), k_advance (minimum over the slots) -- with
    the peers wired by plain device pointers (simian_connect_local) instead of
    CUDA IPC.  Everything runs on ONE stream, phase by phase over the ranks, so
    every flag a k_pull waits for was written by a kernel that has already
    completed: no kernel ever waits for a kernel that has not run.  This is how
    the device-driven exchange is covered by a one-GPU test run."""
    size = len(engines)
    stream = torch.cuda.current_stream(device).cuda_stream
    for e in engines:
        e.set_stream(stream)
    for r, e in enumerate(engines):
        for p, pe in enumerate(engines):
            if p != r:
                e.connect_local(p, pe)
    for e in engines:
        e.run_begin()
    windows = 0
    while True:
        for _ in range(batch):
            for e in engines:
                e.window_process_lbts(True)    # k_window + candidate + publish
            for e in engines:
                e.window_pull(True)            # all flags are already there
            for e in engines:
                e.window_advance_async(True)
            windows += 1
        dones = [e.poll() for e in engines]
        if all(dones):
            break
        assert not any(dones), "the window sequence is global: ranks end together"
        batch = min(2 * batch, max_batch)
    out = []
    for e in engines:
        st = e.run_end()
        st["driver_windows"] = windows
        out.append(st)
    return out


def run_distributed_model(build, name, end, min_delay, rank, world, local_rank, seed=1,
                          driver="p2p", profile=False, **options):
    """One run of a model on `world` GPUs (one process each); returns this
    rank's stats with the engine under "_engine".  build(engine): the model script
    (addEntity / attachService / schedService loops).  driver: "p2p"
    (device-driven pull over NVLink) or "nccl" (host-driven alltoallv)."""
    from . import capi
    device = torch.device("cuda", local_rank)
    e = capi.Engine(name, 0.0, end, min_delay, rank=rank, size=world, device=local_rank,
                    seed=seed, **options)
    build(e)
    if driver == "p2p":
        st = drive_windows_p2p(e, rank, world, device, profile=profile)
    else:
        st = drive_windows(GpuRankEngine(e, device), rank, world)
    st["_engine"] = e
    return st


def run_distributed_phold(cfg, rank, world, local_rank, seed=1, driver="p2p",
                          **options):
    """One PHOLD step on `world` GPUs (one process each); returns this rank's
    stats.  cfg: a dict of simian_b200.workloads.CONFIGS shape."""
    from .workloads import build_phold
    return run_distributed_model(
        lambda e: build_phold(e, cfg["n"], cfg["per_entity"], cfg["lookahead"],
                              cfg["p_remote"], cfg.get("lam", 1.0), cfg["mode"], cfg["groups"]),
        "phold", cfg["end"], cfg["min_delay"], rank, world, local_rank, seed=seed,
        driver=driver, **options)
