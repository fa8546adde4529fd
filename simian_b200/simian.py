"""Host-side mirror of the reference's model-script API, over the C-ABI.

Same names, argument meaning and error behaviour as SimianJS/simian.js and
SimianJS/entity.js (and their Python twin SimianPie/simian.py):

    Simian(simName, startTime, endTime, minDelay, useMPI)     simian.js:36-84
      .addEntity(name, entityClass, num, *args)               simian.js:211-229
      .schedService(time, eventName, data, rx, rxId)          simian.js:170-190
      .attachService(klass, name, fun)                        simian.js:207-209
      .getEntity(name, num)                                   simian.js:200-205
      .getBaseRank(name) / .getOffsetRank(name, num)          simian.js:192-198
      .run() / .exit()                                        simian.js:99-168,90-97
    Entity(name, out, engine, num)                            entity.js:21-29
      .reqService(offset, eventName, data, rx, rxId)          entity.js:35-68
      .attachService(name, fun)                               entity.js:70-72

Extension points kept: a subclass may override getBaseRank (simian.js:192-194,
Docs/README.Simian:92-100) -- the value is handed to the engine
(simian_set_base_rank); getOffsetRank keeps the reference form, which is what the
device-side routing implements.  The per-rank log "<simName>.<rank>.out"
(simian.js:83,225) and the run banner / statistics printed by rank 0
(simian.js:104-113,159-165) are opt-in (`log=True`, `verbose=True`): a 16 M
entity model would write 16 M lines.

What changes: service handlers are not host closures but CUDA device functions
compiled into libsimian_b200.so, bound BY NAME (`device_fn("phold_generate")`),
and `data` is an 8-byte payload (an int; longer payloads: handler-side req_service_blob).  Process-oriented coroutines
(entity.js:74-205) are out of scope.
"""
import struct

from . import capi


class DeviceFn:
    """A service handler that lives on the GPU: the name of a __device__
    function registered in the library (simian_attach_service)."""

    def __init__(self, name):
        self.name = name

    def __repr__(self):
        return "device_fn(%r)" % self.name


def device_fn(name):
    return DeviceFn(name)


class SimianError(capi.SimianError):
    pass


class Entity(object):
    """Entity base class (entity.js:21-29).  Subclasses declare services as
    class attributes:  generate = device_fn("phold_generate")
    and optionally  state_bytes = N  and  params = bytes(...)  (model
    parameters the device handlers read)."""
    state_bytes = 0
    params = b""

    def __init__(self, name, out, engine, num, *args):
        self.name = name        # entity.js:23
        self.out = out          # entity.js:24
        self.engine = engine    # entity.js:25
        self.num = num          # entity.js:26

    def __str__(self):
        return "%s(%d)" % (self.name, self.num)  # entity.js:31-33

    def reqService(self, offset, eventName, data, rx=None, rxId=None):
        """entity.js:35-68.  While the engine runs, reqService is executed by
        the device handlers; from the host it is only legal before run()
        (constructor-time sends, e.g. pdes_lanl_benchmarkV8.js:255)."""
        eng = self.engine
        if rx is not None and offset < eng.minDelay:        # entity.js:41
            if not eng.running:
                raise SimianError(2, "Cannot send an event when Simian is idle")
            raise SimianError(1, "%s attempted to send with too little delay" % self)
        time = eng.now + offset                             # entity.js:47
        if time > eng.endTime:                              # entity.js:48
            return
        rx = self.name if rx is None else rx                # entity.js:50-51
        rxId = self.num if rxId is None else rxId
        eng._inject(time, eventName, data, rx, rxId, self.name, self.num)

    def attachService(self, name, fun):                     # entity.js:70-72
        self.engine.attachService(type(self), name, fun, _class_name=self.name)

    @property
    def state(self):
        """the entity's user state (state_bytes of the class) read back from HBM:
        what `engine.getEntity(name, num).<field>` reads in the reference"""
        nb = type(self).state_bytes
        return bytes(self.engine._h.read_state(self.name, 1, nb, first=self.num)) if nb else b""

    @property
    def count(self):
        """committed events of this entity (read back from HBM)"""
        return int(self.engine._h.read_counts(self.name, 1, first=self.num)[0])

    @property
    def traceHash(self):
        return int(self.engine._h.read_trace_hashes(self.name, 1, first=self.num)[0])


class Simian(object):
    version = "simian-b200 (SimianJS event-oriented service API on B200)"

    def __init__(self, simName, startTime=0.0, endTime=1e100, minDelay=1,
                 useMPI=False, rank=None, size=None, device=-1, seed=1,
                 log=False, verbose=False, **options):
        self.name = simName
        self.startTime = startTime
        self.endTime = endTime or 1e100          # simian.js:41
        self.minDelay = minDelay or 1            # simian.js:42
        self.now = startTime                     # simian.js:44
        self.running = False                     # simian.js:50
        self.entities = {}                       # simian.js:53
        self.baseRanks = {}                      # simian.js:67
        self.useMPI = bool(useMPI)
        if useMPI:                               # simian.js:71-75: MPI rank/size
            import torch.distributed as dist
            self.rank = dist.get_rank() if rank is None else rank
            self.size = dist.get_world_size() if size is None else size
        else:                                    # simian.js:77-79
            self.rank, self.size = 0, 1
        self._h = capi.Engine(simName, startTime, endTime, minDelay,
                              rank=self.rank, size=self.size, device=device,
                              seed=seed, **options)
        self._classes = {}
        self._added = {}     # class -> instances announced by addEntity, not yet in the engine
        self._pending = []   # host records injected before run
        self._seq = 0
        # One output file per rank (simian.js:83), opt-in: log=True, a path
        # prefix, or an open file object
        self.verbose = bool(verbose)
        self._own_out = False
        if log is True:
            log = simName
        if isinstance(log, str):
            self.out = open("%s.%d.out" % (log, self.rank), "w")
            self._own_out = True
        else:
            self.out = log or None
        self.stats = None
        if type(self).getOffsetRank is not Simian.getOffsetRank:
            raise SimianError(6, "getOffsetRank cannot be overridden: the device-side routing "
                                 "implements (baseRank + num) % size; override getBaseRank")

    def __str__(self):                           # simian.js:86-88
        return self.version

    # ---- placement ------------------------------------------------------
    def getBaseRank(self, name):                 # simian.js:192-194
        return int(capi.lib().simian_hash_name(name.encode()) % self.size)

    def getOffsetRank(self, name, num):          # simian.js:196-198
        return (self.baseRanks[name] + num) % self.size

    def getEntity(self, name, num):              # simian.js:200-205
        ents = self.entities.get(name)
        if ents is not None:
            return ents.get(num)

    # ---- model construction -----------------------------------------------
    def _register_class(self, name, entityClass):
        if name in self._classes:
            return
        self._classes[name] = entityClass
        self._h.define_class(name, entityClass.state_bytes)
        # a subclass's own getBaseRank (simian.js:192-194) goes to the engine
        if type(self).getBaseRank is not Simian.getBaseRank:
            self._h.set_base_rank(name, int(self.getBaseRank(name)) % self.size)
        for attr in dir(entityClass):
            fn = getattr(entityClass, attr, None)
            if isinstance(fn, DeviceFn):
                self._h.attach_service(name, attr, fn.name)
        if entityClass.params:
            self._h.set_class_params(name, bytes(entityClass.params))

    def attachService(self, klass, name, fun, _class_name=None):  # simian.js:207-209
        if not isinstance(fun, DeviceFn):
            raise SimianError(5, "service handlers are device functions: "
                                 "use device_fn('name')")
        setattr(klass, name, fun)
        for cname, k in self._classes.items():
            if k is klass or cname == _class_name:
                self._h.attach_service(cname, name, fun.name)

    def addEntity(self, name, entityClass, num, *theArgs):        # simian.js:211-229
        if self.running:
            raise SimianError(4, "Cannot add new entity when Simian is running!")
        self._register_class(name, entityClass)
        ents = self.entities.setdefault(name, {})
        self.baseRanks[name] = self.getBaseRank(name)
        # The engine numbers entities class after class; model scripts may
        # interleave classes (hello.js:98-101), so the instances are counted
        # here and handed over in one piece per class (_flush_entities).
        first, cnt = self._added.get(name, (num, 0))
        if num != first + cnt:
            raise SimianError(6, "instances of %s must be added as 0, 1, 2, ..." % name)
        self._added[name] = (first, cnt + 1)
        computedRank = self.getOffsetRank(name, num)
        if computedRank == self.rank:                             # simian.js:223
            if self.out is not None:                              # simian.js:225
                self.out.write("%s(%d) Running on rank %d\n" % (name, num, computedRank))
            ents[num] = entityClass(name, self.out, self, num, *theArgs)

    def _flush_entities(self):
        for name in self._classes:            # registration order = id order
            if name in self._added:
                first, cnt = self._added.pop(name)
                self._h.add_entities(name, first, cnt)

    def addEntities(self, name, entityClass, first, count):
        """Bulk form of the model-script loop `for i: addEntity(name, K, i)`
        (phold-noop-noMPI.js:46): no per-entity host object is created;
        getEntity() builds them lazily."""
        if self.running:
            raise SimianError(4, "Cannot add new entity when Simian is running!")
        self._register_class(name, entityClass)
        self.baseRanks[name] = self.getBaseRank(name)
        self._flush_entities()
        self._h.add_entities(name, first, count)
        self.entities.setdefault(name, _LazyEntities(self, name, entityClass))
        if self.out is not None:                                  # simian.js:225
            base = self.baseRanks[name]
            self.out.writelines("%s(%d) Running on rank %d\n" % (name, num, self.rank)
                                for num in range(first, first + count)
                                if (base + num) % self.size == self.rank)

    def setClassParams(self, name, blob):
        self._h.set_class_params(name, bytes(blob))

    def firstId(self, name):
        """global id of instance 0 (needs every entity to have been added)"""
        self._flush_entities()
        return self._h.first_id(name)

    def internService(self, name):
        return self._h.intern_service(name)

    def schedService(self, time, eventName, data, rx, rxId):      # simian.js:170-190
        if time > self.endTime:
            return
        self._flush_entities()
        self._h.sched_service(time, eventName, int(data or 0), rx, rxId)

    def schedServiceBulk(self, time, eventName, data, rx, first, count, repeat=1):
        self._flush_entities()
        self._h.sched_service_bulk(time, eventName, int(data or 0), rx, first,
                                   count, repeat)

    def _inject(self, time, eventName, data, rx, rxId, tx, txId):
        import numpy as np
        self._flush_entities()
        ev = np.zeros(1, dtype=capi.EVENT_DTYPE)
        ev["time"] = time
        ev["dst"] = self._h.first_id(rx) + rxId
        ev["src"] = self._h.first_id(tx) + txId
        ev["handler"] = self._h.intern_service(eventName)
        ev["seq"] = self._seq
        ev["data"] = int(data or 0)
        self._seq += 1
        self._h.sched_events(ev)

    # ---- run ------------------------------------------------------------------
    def run(self):                                                # simian.js:99-168
        self._flush_entities()
        if self.verbose and not self.rank:                        # simian.js:104-113
            print(self)
            print("===========================================")
            print("----------SIMIAN-B200 PDES ENGINE----------")
            print("===========================================")
            print("GPU: ON")
            print("MPI: ON" if self.useMPI else "MPI: OFF")
        self.running = True
        try:
            if self.size == 1:
                self.stats = self._h.run()
            else:
                import torch
                from .distributed import drive_windows_p2p
                dev = torch.device("cuda", torch.cuda.current_device())
                self.stats = drive_windows_p2p(self._h, self.rank, self.size, dev)
        finally:
            self.running = False
        self.now = self.stats["last_time"]
        if self.stats.get("distinguishable_ties"):
            # equal-time events of one entity that differ in what the handler sees
            # commit in the engine's order (time, handler, src, seq); the reference
            # engines use heap shape (SimianJS) / FIFO (SimianPie) there, so
            # per-entity traces can differ from theirs on such models
            import warnings
            warnings.warn("%d distinguishable same-entity time ties: their commit order is the "
                          "engine's (time, handler, src, seq), not the reference heap's"
                          % self.stats["distinguishable_ties"], RuntimeWarning, stacklevel=2)
        if self.verbose:
            for line in self.report_lines():
                print(line)
        return self.stats

    def report_lines(self):
        """What rank 0 prints after a run (simian.js:151-165); [] on other ranks.
        size > 1: the event counts are summed over the ranks (MPI allreduce SUM,
        simian.js:153-154)."""
        st = self.stats
        total, synch = st["total_events"], st["synch_events"]
        if self.size > 1:
            import torch
            import torch.distributed as dist
            t = torch.tensor([total, synch], dtype=torch.int64)
            if dist.get_backend() == "nccl":
                t = t.cuda()
            dist.all_reduce(t)
            total, synch = int(t[0]), int(t[1])
        if self.rank:
            return []
        sec = st["run_seconds"] or float("nan")
        return ["SIMULATION FINISHED IN : %s (s)" % sec,
                "SIMULATED EVENTS       : %d" % total,
                "SYNCHRONIZATION EVENTS : %d" % synch,
                "MODEL-EVENTS/SECOND    : %s" % (total / sec),
                "TOTAL-EVENTS/SECOND    : %s" % ((total + synch) / sec),
                "==========================================="]

    def exit(self):                                               # simian.js:90-97
        if self.out is not None and self._own_out:
            self.out.flush()
            self.out.close()
        self.out = None
        if self._h is not None:
            self._h.close()
            self._h = None


class _LazyEntities(dict):
    def __init__(self, engine, name, klass):
        super().__init__()
        self._e, self._n, self._k = engine, name, klass

    def get(self, num, default=None):
        if num not in self:
            if self._e.getOffsetRank(self._n, num) != self._e.rank:
                return default
            self[num] = self._k(self._n, self._e.out, self._e, num)
        return self[num]


def phold_params(first_id, count, lookahead, p_remote=0.0, lam=1.0, mode=0, groups=1):
    """simian_phold_params_t of include/simian_models.h"""
    return struct.pack("<IIdddIIII", first_id, count, p_remote, lam, lookahead,
                       mode, groups, 0, 0)
