#!/usr/bin/env python
"""A model whose handlers are NOT part of the library (tests/user_handlers/bounce.h,
registered through bounce.def): engine against oracle, both built with
make HANDLERS=tests/user_handlers/bounce NAME=user.  Run with
SIMIAN_B200_LIB / SIMIAN_ORACLE_LIB pointing at the two variants."""
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as om  # noqa: E402
from simian_b200 import capi  # noqa: E402

N, END, MIN_DELAY = 5000, 12.0, 0.25


def build(eng):
    eng.define_class("Ring", 0)
    eng.add_entities("Ring", 0, N)
    eng.attach_service("Ring", "bounce", "user_bounce")
    eng.attach_service("Ring", "tick", "user_tick")
    ids = [eng.intern_service(s) for s in ("bounce", "tick")]
    eng.set_class_params("Ring", struct.pack("<IIdHHI", eng.first_id("Ring"), N, MIN_DELAY,
                                             ids[0], ids[1], 0))
    eng.sched_service_bulk(0.0, "bounce", 0, "Ring", 0, N, 2)


o = om.Oracle("bounce", 0.0, END, MIN_DELAY, seed=5, tie_mode=1)
build(o)
so = o.run()
co, ho = o.read_counts("Ring", N), o.read_trace_hashes("Ring", N)
e = capi.Engine("bounce", 0.0, END, MIN_DELAY, seed=5, payloads=1)  # (the ticks carry payloads)
build(e)
se = e.run()
ce, he = e.read_counts("Ring", N), e.read_trace_hashes("Ring", N)
e.close()
assert se["total_events"] == so["total_events"] and se["total_events"] > 10 * N, (se, so)
assert se["digest"] == so["digest"] and se["windows"] == so["windows"]
assert float(se["last_time"]).hex() == float(so["last_time"]).hex()
np.testing.assert_array_equal(ce, co)
np.testing.assert_array_equal(he, ho)
print("USER-OK events %d windows %d digest %016x" % (se["total_events"], se["windows"],
                                                     se["digest"]))
