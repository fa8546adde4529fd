"""ctypes binding of libsimian_b200.so -- one Python method per C-ABI entry
point of include/simian_b200.h, nothing else.  There is no CPU fallback: when
the library is missing or no CUDA device is present this module raises."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SIMIAN_B200_LIB") or os.path.join(
    HERE, "libsimian_b200.so")  # the env override selects a tuning variant

EVENT_DTYPE = np.dtype(
    [("time", "<f8"), ("dst", "<u4"), ("src", "<u4"), ("handler", "<u2"),
     ("flags", "<u2"), ("seq", "<u4"), ("data", "<u8")], align=False)
assert EVENT_DTYPE.itemsize == 32
NULL_ENTITY = 0xFFFFFFFF


class Stats(C.Structure):
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
        ("device_seconds", C.c_double),
        ("kernel_launches", C.c_uint64),
        ("redistributions", C.c_uint64),
        ("far_appends", C.c_uint64),
        ("pages_allocated", C.c_uint64),
        ("error", C.c_int32),
        ("pad_", C.c_int32),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "pad_"}


# every symbol include/simian_b200.h declares: name -> (restype, argtypes)
_vp, _cp, _d, _u64, _u32, _i = (C.c_void_p, C.c_char_p, C.c_double, C.c_uint64,
                                C.c_uint32, C.c_int)
SIGNATURES = {
    "simian_version": (_cp, []),
    "simian_create": (_vp, [_cp, _d, _d, _d, _i, _i, _i]),
    "simian_destroy": (None, [_vp]),
    "simian_last_error": (_cp, [_vp]),
    "simian_set_option": (_i, [_vp, _cp, _d]),
    "simian_define_class": (_i, [_vp, _cp, _u32]),
    "simian_set_class_params": (_i, [_vp, _cp, _vp, _u32]),
    "simian_add_entities": (_i, [_vp, _cp, _u32, _u32, _vp]),
    "simian_attach_service": (_i, [_vp, _cp, _cp, _cp]),
    "simian_intern_service": (_i, [_vp, _cp]),
    "simian_first_id": (C.c_int64, [_vp, _cp]),
    "simian_base_rank": (_i, [_vp, _cp]),
    "simian_set_base_rank": (_i, [_vp, _cp, _i]),
    "simian_hash_name": (_d, [_cp]),
    "simian_sched_service": (_i, [_vp, _d, _cp, _u64, _cp, _u32]),
    "simian_sched_service_bulk": (_i, [_vp, _d, _cp, _u64, _cp, _u32, _u32, _u32]),
    "simian_sched_events": (_i, [_vp, _vp, _u64]),
    "simian_construct_entities": (_i, [_vp, _cp, _cp, _u64]),
    "simian_run": (_i, [_vp, C.POINTER(Stats)]),
    "simian_run_begin": (_i, [_vp]),
    "simian_window_process": (_i, [_vp]),
    "simian_window_outbox": (_i, [_vp, C.POINTER(_u32), C.POINTER(_vp),
                                   C.POINTER(_u64)]),
    "simian_window_ingest": (_i, [_vp, _vp, _u64]),
    "simian_window_local_min": (_i, [_vp, C.POINTER(_vp)]),
    "simian_window_advance": (_i, [_vp, C.POINTER(_i)]),
    "simian_run_end": (_i, [_vp, C.POINTER(Stats)]),
    "simian_setup": (_i, [_vp]),
    "simian_ipc_export": (_i, [_vp, _vp, _u64, C.POINTER(_u64)]),
    "simian_ipc_import": (_i, [_vp, _i, _vp, _u64]),
    "simian_connect_local": (_i, [_vp, _i, _vp]),
    "simian_window_pull": (_i, [_vp, _i]),
    "simian_window_publish_min": (_i, [_vp]),
    "simian_window_process_lbts": (_i, [_vp, _i]),
    "simian_window_advance_async": (_i, [_vp, _i]),
    "simian_poll": (_i, [_vp, C.POINTER(_i)]),
    "simian_minvec_ptr": (_i, [_vp, C.POINTER(_vp)]),
    "simian_set_stream": (_i, [_vp, _vp]),
    "simian_read_counts": (_i, [_vp, _cp, _u32, _u32, _vp]),
    "simian_read_trace_hashes": (_i, [_vp, _cp, _u32, _u32, _vp]),
    "simian_read_state": (_i, [_vp, _cp, _u32, _u32, _vp]),
    "simian_read_local": (_i, [_vp, _cp, _i, _vp, _u64, C.POINTER(_u64), C.POINTER(_u32),
                               C.POINTER(_u32)]),
    "simian_debug_phase_cycles": (_i, [_vp, C.POINTER(_u64)]),
    "simian_device_function_id": (_i, [_cp]),
    "simian_trim_cache": (None, []),
}

_lib = None


def lib():
    """Load libsimian_b200.so (built in-tree by __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "%s is missing: run `python -c 'import __graft_entry__ as g; "
                "g.build()'` (there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        _lib = L
    return _lib


class SimianError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("simian error %d: %s" % (code, msg))
        self.code = code


class Engine:
    """One engine = one rank = one GPU (simian.js:71-79)."""

    def __init__(self, name, start, end, min_delay, rank=0, size=1, device=-1,
                 seed=1, **options):
        self.L = lib()
        self.rank, self.size = rank, size
        self.h = self.L.simian_create(name.encode(), start, end, min_delay,
                                      rank, size, device)
        if not self.h:
            raise SimianError(14, self.L.simian_last_error(None).decode())
        self.set_option("seed", seed)
        for k, v in options.items():
            self.set_option(k, v)

    def close(self):
        if getattr(self, "h", None):
            self.L.simian_destroy(self.h)
            self.h = None

    __del__ = close

    def _ck(self, rc):
        if rc:
            raise SimianError(rc, self.L.simian_last_error(self.h).decode())

    def set_option(self, key, value):
        self._ck(self.L.simian_set_option(self.h, key.encode(), float(value)))

    def define_class(self, name, state_bytes=0):
        rc = self.L.simian_define_class(self.h, name.encode(), state_bytes)
        if rc < 0:
            self._ck(-rc)
        return rc

    def set_class_params(self, name, blob):
        buf = (C.c_char * len(blob)).from_buffer_copy(blob)
        self._ck(self.L.simian_set_class_params(self.h, name.encode(), buf,
                                                len(blob)))

    def add_entities(self, name, first, count, init_state=None):
        ptr = None
        if init_state is not None:
            init_state = np.ascontiguousarray(init_state)
            ptr = init_state.ctypes.data
        self._ck(self.L.simian_add_entities(self.h, name.encode(), first, count,
                                            ptr))

    def attach_service(self, cls, service, fn):
        self._ck(self.L.simian_attach_service(self.h, cls.encode(),
                                              service.encode(), fn.encode()))

    def intern_service(self, name):
        rc = self.L.simian_intern_service(self.h, name.encode())
        if rc < 0:
            self._ck(-rc)
        return rc

    def first_id(self, name):
        return int(self.L.simian_first_id(self.h, name.encode()))

    def base_rank(self, name):
        return self.L.simian_base_rank(self.h, name.encode())

    def set_base_rank(self, name, base_rank):
        self._ck(self.L.simian_set_base_rank(self.h, name.encode(), int(base_rank)))

    def sched_service(self, time, service, data, rx, rx_id):
        self._ck(self.L.simian_sched_service(self.h, time, service.encode(),
                                             data, rx.encode(), rx_id))

    def sched_service_bulk(self, time, service, data, rx, first, count, repeat=1):
        self._ck(self.L.simian_sched_service_bulk(
            self.h, time, service.encode(), data, rx.encode(), first, count,
            repeat))

    def sched_events(self, events):
        ev = np.ascontiguousarray(events, dtype=EVENT_DTYPE)
        self._ck(self.L.simian_sched_events(self.h, ev.ctypes.data, len(ev)))

    def construct_entities(self, cls, fn, expected_events=0):
        self._ck(self.L.simian_construct_entities(self.h, cls.encode(), fn.encode(),
                                                  int(expected_events)))

    def run(self, check=True):
        st = Stats()
        rc = self.L.simian_run(self.h, C.byref(st))
        if check:
            self._ck(rc)
        return st.as_dict()

    # ---- stepping (size > 1) ------------------------------------------
    def run_begin(self):
        self._ck(self.L.simian_run_begin(self.h))

    def window_process(self):
        self._ck(self.L.simian_window_process(self.h))

    def window_outbox(self):
        counts = (_u32 * self.size)()
        ptr, cap = _vp(), _u64()
        self._ck(self.L.simian_window_outbox(self.h, counts, C.byref(ptr),
                                             C.byref(cap)))
        return list(counts), ptr.value, cap.value

    def window_ingest(self, dev_ptr, n):
        self._ck(self.L.simian_window_ingest(self.h, dev_ptr, n))

    def window_local_min(self):
        ptr = _vp()
        self._ck(self.L.simian_window_local_min(self.h, C.byref(ptr)))
        return ptr.value

    def window_local_min_ptr(self):
        """device address of the {minLeft, min far time} pair (no launch)"""
        ptr = _vp()
        self._ck(self.L.simian_minvec_ptr(self.h, C.byref(ptr)))
        return ptr.value

    def window_advance(self):
        done = _i()
        self._ck(self.L.simian_window_advance(self.h, C.byref(done)))
        return bool(done.value)

    # ---- device-driven exchange over peer memory -------------------------
    def setup(self):
        self._ck(self.L.simian_setup(self.h))

    def ipc_export(self):
        n = _u64()
        self._ck(self.L.simian_ipc_export(self.h, None, 0, C.byref(n)))
        buf = C.create_string_buffer(n.value)
        self._ck(self.L.simian_ipc_export(self.h, buf, n.value, C.byref(n)))
        return buf.raw[:n.value]

    def ipc_import(self, peer_rank, blob):
        buf = C.create_string_buffer(bytes(blob), len(blob))
        self._ck(self.L.simian_ipc_import(self.h, peer_rank, buf, len(blob)))

    def connect_local(self, peer_rank, peer):
        """wire a peer engine of the SAME process by plain device pointers"""
        self._ck(self.L.simian_connect_local(self.h, peer_rank, peer.h))

    def window_pull(self, wait_for_peers=False):
        self._ck(self.L.simian_window_pull(self.h, int(wait_for_peers)))

    def window_process_lbts(self, publish=True):
        self._ck(self.L.simian_window_process_lbts(self.h, int(publish)))

    def window_publish_min(self):
        self._ck(self.L.simian_window_publish_min(self.h))

    def window_advance_async(self, min_from_peers=False):
        self._ck(self.L.simian_window_advance_async(self.h, int(min_from_peers)))

    def poll(self):
        done = _i()
        self._ck(self.L.simian_poll(self.h, C.byref(done)))
        return bool(done.value)

    def run_end(self, check=True):
        st = Stats()
        rc = self.L.simian_run_end(self.h, C.byref(st))
        if check:
            self._ck(rc)
        return st.as_dict()

    def set_stream(self, cuda_stream):
        self._ck(self.L.simian_set_stream(self.h, cuda_stream))

    def debug_phase_cycles(self):
        out = (_u64 * 12)()
        self._ck(self.L.simian_debug_phase_cycles(self.h, out))
        return list(out)

    # ---- results --------------------------------------------------------
    def _read(self, fn, name, n, first=0):
        out = np.zeros(n, dtype=np.uint64)
        self._ck(fn(self.h, name.encode(), first, n, out.ctypes.data))
        return out

    def read_counts(self, name, n, first=0):
        return self._read(self.L.simian_read_counts, name, n, first)

    def read_trace_hashes(self, name, n, first=0):
        return self._read(self.L.simian_read_trace_hashes, name, n, first)

    def read_local(self, name, which, out=None):
        """this rank's own instances: (values, first num, stride); which: 0
        counts, 1 trace hashes.  `out`: a uint64 array to fill (page-locked host
        memory is written by one DMA, without a bounce buffer)"""
        n, first, stride = _u64(), _u32(), _u32()
        self._ck(self.L.simian_read_local(self.h, name.encode(), which, None, 0,
                                          C.byref(n), C.byref(first), C.byref(stride)))
        if out is None:
            out = np.empty(n.value, dtype=np.uint64)
        elif out.dtype != np.uint64 or out.size < n.value or not out.flags.c_contiguous:
            raise ValueError("out: contiguous uint64 array of at least %d entries" % n.value)
        else:
            out = out[:n.value]
        self._ck(self.L.simian_read_local(self.h, name.encode(), which, out.ctypes.data,
                                          n.value, C.byref(n), C.byref(first), C.byref(stride)))
        return out, first.value, stride.value

    def read_state(self, name, n, state_bytes, first=0):
        out = np.zeros(n * state_bytes, dtype=np.uint8)
        self._ck(self.L.simian_read_state(self.h, name.encode(), first, n,
                                          out.ctypes.data))
        return out
