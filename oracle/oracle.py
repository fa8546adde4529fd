"""ctypes binding of the CPU oracle (oracle/simian_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, by __graft_entry__.smoke() and by
bench.py's cpu_baseline / --impl reference legs -- never by simian_b200/.
The method names mirror the product's low-level binding
(simian_b200/capi.py) so one model-builder function can drive both.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# SIMIAN_ORACLE_LIB: a variant built with a model author's handlers
# (make -C oracle HANDLERS=... NAME=...), see include/simian_handlers.def
LIB_PATH = os.environ.get("SIMIAN_ORACLE_LIB") or os.path.join(HERE, "_build",
                                                               "libsimian_oracle.so")


class OracleStats(C.Structure):
    _fields_ = [
        ("total_events", C.c_uint64),
        ("windows", C.c_uint64),
        ("synch_events", C.c_uint64),
        ("digest", C.c_uint64),
        ("state_checksum", C.c_uint64),
        ("ties", C.c_uint64),
        ("distinguishable_ties", C.c_uint64),
        ("emitted", C.c_uint64),
        ("dropped_past_end", C.c_uint64),
        ("last_time", C.c_double),
        ("run_seconds", C.c_double),
        ("error", C.c_int32),
        ("pad_", C.c_int32),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "pad_"}


EVENT_DTYPE = np.dtype(
    [("time", "<f8"), ("dst", "<u4"), ("src", "<u4"), ("handler", "<u2"),
     ("flags", "<u2"), ("seq", "<u4"), ("data", "<u8")], align=False)
assert EVENT_DTYPE.itemsize == 32


def build(force=False):
    """Compile the oracle with the recipe in oracle/Makefile."""
    if os.environ.get("SIMIAN_ORACLE_LIB"):
        return LIB_PATH
    if force or not os.path.exists(LIB_PATH) or any(
            os.path.getmtime(os.path.join(HERE, p)) > os.path.getmtime(LIB_PATH)
            for p in ("simian_oracle.cpp", "../include/simian_spec.h",
                      "../include/simian_models.h", "../include/simian_handlers.def")):
        subprocess.check_call(["make", "-s", "-C", HERE])
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        vp, cp, d, u64, u32, i = (C.c_void_p, C.c_char_p, C.c_double,
                                  C.c_uint64, C.c_uint32, C.c_int)
        sig = {
            "oracle_hash_name": (d, [cp]),
            "oracle_create": (vp, [cp, d, d, d, i, i]),
            "oracle_destroy": (None, [vp]),
            "oracle_set_seed": (None, [vp, u64]),
            "oracle_last_error": (cp, [vp]),
            "oracle_define_class": (i, [vp, cp, u32]),
            "oracle_set_class_params": (i, [vp, cp, vp, u32]),
            "oracle_add_entities": (i, [vp, cp, u32, u32]),
            "oracle_first_id": (i, [vp, cp]),
            "oracle_base_rank": (i, [vp, cp]),
            "oracle_set_base_rank": (i, [vp, cp, i]),
            "oracle_intern_service": (i, [vp, cp]),
            "oracle_attach_service": (i, [vp, cp, cp, cp]),
            "oracle_sched_service": (i, [vp, d, cp, u64, cp, u32]),
            "oracle_sched_service_bulk": (i, [vp, d, cp, u64, cp, u32, u32, u32]),
            "oracle_sched_events": (i, [vp, vp, u64]),
            "oracle_construct_entities": (i, [vp, cp, cp]),
            "oracle_set_threads": (None, [vp, i]),
            "oracle_run": (i, [vp, C.POINTER(OracleStats)]),
            "oracle_read_counts": (i, [vp, cp, vp]),
            "oracle_read_trace_hashes": (i, [vp, cp, vp]),
            "oracle_read_state": (i, [vp, cp, vp]),
            "oracle_step_begin": (None, [vp]),
            "oracle_step_process": (i, [vp, i, d]),
            "oracle_step_outbox": (u64, [vp, i, i, vp, u64]),
            "oracle_step_ingest": (None, [vp, i, vp, u64]),
            "oracle_step_min_left": (d, [vp, i]),
            "oracle_step_end": (i, [vp, i, C.POINTER(OracleStats)]),
            "oracle_start_time": (d, [vp]),
            "oracle_end_time": (d, [vp]),
            "oracle_owner_rank": (i, [vp, u32]),
            "oracle_q_new": (vp, [i]),
            "oracle_q_free": (None, [vp]),
            "oracle_q_push": (None, [vp, d, u64]),
            "oracle_q_len": (u64, [vp]),
            "oracle_q_peek": (d, [vp]),
            "oracle_q_pop": (d, [vp, C.POINTER(u64)]),
            "oracle_spec_log": (d, [d]),
            "oracle_spec_mix64": (u64, [u64]),
            "oracle_spec_rng_key": (u64, [u64, u32, u64]),
            "oracle_spec_rng_draw": (u64, [u64, u32]),
            "oracle_spec_u01": (d, [u64]),
            "oracle_spec_u01_open": (d, [u64]),
            "oracle_spec_exponential": (d, [u64, d]),
            "oracle_spec_powi": (d, [d, u64]),
            "oracle_spec_trace_step": (u64, [u64, d, u32, u32, u64]),
            "oracle_spec_digest_term": (u64, [u32, u64, u64]),
        }
        for name, (res, args) in sig.items():
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        _lib = L
    return _lib


class OracleError(RuntimeError):
    pass


class Oracle:
    """Low-level handle; same method names as simian_b200.capi.Engine."""

    def __init__(self, name, start, end, min_delay, size=1, tie_mode=0, seed=1,
                 threads=1):
        self.L = lib()
        self.h = self.L.oracle_create(name.encode(), start, end, min_delay,
                                      size, tie_mode)
        self.L.oracle_set_seed(self.h, seed)
        self.L.oracle_set_threads(self.h, threads)
        self.classes = {}

    def close(self):
        if self.h:
            self.L.oracle_destroy(self.h)
            self.h = None

    __del__ = close

    def _ck(self, rc):
        if rc:
            raise OracleError("oracle error %d: %s" % (
                rc, self.L.oracle_last_error(self.h).decode()))

    def define_class(self, name, state_bytes=0):
        self.classes[name] = state_bytes
        return self.L.oracle_define_class(self.h, name.encode(), state_bytes)

    def set_class_params(self, name, blob):
        buf = (C.c_char * len(blob)).from_buffer_copy(blob)
        self._ck(self.L.oracle_set_class_params(self.h, name.encode(), buf,
                                                len(blob)))

    def add_entities(self, name, first, count):
        self._ck(self.L.oracle_add_entities(self.h, name.encode(), first, count))

    def first_id(self, name):
        return self.L.oracle_first_id(self.h, name.encode())

    def base_rank(self, name):
        return self.L.oracle_base_rank(self.h, name.encode())

    def set_base_rank(self, name, base_rank):
        self._ck(self.L.oracle_set_base_rank(self.h, name.encode(), int(base_rank)))

    def intern_service(self, name):
        return self.L.oracle_intern_service(self.h, name.encode())

    def attach_service(self, cls, service, fn):
        self._ck(self.L.oracle_attach_service(self.h, cls.encode(),
                                              service.encode(), fn.encode()))

    def sched_service(self, time, service, data, rx, rx_id):
        self._ck(self.L.oracle_sched_service(self.h, time, service.encode(),
                                             data, rx.encode(), rx_id))

    def sched_service_bulk(self, time, service, data, rx, first, count, repeat=1):
        self._ck(self.L.oracle_sched_service_bulk(
            self.h, time, service.encode(), data, rx.encode(), first, count,
            repeat))

    def sched_events(self, events):
        ev = np.ascontiguousarray(events, dtype=EVENT_DTYPE)
        self._ck(self.L.oracle_sched_events(self.h, ev.ctypes.data, len(ev)))

    def construct_entities(self, cls, fn, expected_events=0):
        self._ck(self.L.oracle_construct_entities(self.h, cls.encode(), fn.encode()))

    def read_state(self, name, n, state_bytes):
        out = np.zeros(n * state_bytes, dtype=np.uint8)
        self._ck(self.L.oracle_read_state(self.h, name.encode(), out.ctypes.data))
        return out

    def run(self, check=True):
        st = OracleStats()
        rc = self.L.oracle_run(self.h, C.byref(st))
        if check:
            self._ck(rc)
        return st.as_dict()

    def _read(self, fn, name, n):
        out = np.zeros(n, dtype=np.uint64)
        self._ck(fn(self.h, name.encode(), out.ctypes.data))
        return out

    def read_counts(self, name, n):
        return self._read(self.L.oracle_read_counts, name, n)

    def read_trace_hashes(self, name, n):
        return self._read(self.L.oracle_read_trace_hashes, name, n)


class OracleRankEngine:
    """One rank of an oracle world behind the RankEngine protocol of
    simian_b200/distributed.py (CPU tensors).  Lets the gloo tests drive the
    product's window/exchange loop without a GPU."""

    def __init__(self, oracle, rank, size):
        import torch
        self.torch = torch
        self.o, self.rank, self.size = oracle, rank, size
        self.L = oracle.L
        self.gmin = self.L.oracle_start_time(oracle.h)
        self.end_time = self.L.oracle_end_time(oracle.h)
        self.minvec = torch.zeros(2, dtype=torch.float64)

    def begin(self):
        self.L.oracle_step_begin(self.o.h)

    def process(self):
        rc = self.L.oracle_step_process(self.o.h, self.rank, self.gmin)
        self.o._ck(rc)

    def outbox(self):
        counts, bufs = [], []
        for peer in range(self.size):
            n = self.L.oracle_step_outbox(self.o.h, self.rank, peer, None, 0)
            arr = np.zeros(max(n, 1), dtype=EVENT_DTYPE)
            got = self.L.oracle_step_outbox(self.o.h, self.rank, peer,
                                            arr.ctypes.data, n) if n else 0
            counts.append(int(got))
            bufs.append(self.torch.from_numpy(arr.view(np.int64).reshape(-1, 4))[:got])
        return counts, bufs

    def new_inbox(self, n):
        return self.torch.zeros((max(n, 1), 4), dtype=self.torch.int64)

    def ingest(self, inbox, n):
        if n:
            arr = np.ascontiguousarray(inbox[:n].numpy())
            self.L.oracle_step_ingest(self.o.h, self.rank, arr.ctypes.data, n)

    def local_min(self):
        self.minvec[0] = self.L.oracle_step_min_left(self.o.h, self.rank)
        self.minvec[1] = float("inf")
        return self.minvec

    def advance(self):
        self.gmin = float(self.minvec[0])  # allreduced by the driver
        return not (self.gmin <= self.end_time)  # simian.js:119

    def end(self):
        st = OracleStats()
        self.o._ck(self.L.oracle_step_end(self.o.h, self.rank, C.byref(st)))
        return st.as_dict()
