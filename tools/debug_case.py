#!/usr/bin/env python
"""Debug aid: run small PHOLD cases on engine and oracle, print stats side by side."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from oracle import oracle as om
om.build()
from helpers import run_engine_phold, run_oracle_phold
import numpy as np
cases = [
    dict(name="tiny", n=100, per_entity=1, min_delay=0.1, lookahead=0.1, p_remote=0.0, mode=0, end=5.0, seed=1),
    dict(name="one_block", n=1000, per_entity=1, min_delay=0.05, lookahead=0.05, p_remote=0.0, mode=0, end=30.0, seed=1),
    dict(name="multi", n=5000, per_entity=16, min_delay=0.1, lookahead=0.1, p_remote=0.1, mode=1, end=6.0, seed=2),
]
sel = sys.argv[1:] or [c["name"] for c in cases]
for c in cases:
    if c["name"] not in sel: continue
    so, co, ho = run_oracle_phold(om, c)
    for seq in (0, 1):
        try:
            se, ce, he = run_engine_phold(c, force_sequential=seq)
        except Exception as ex:
            print(c["name"], "seq", seq, "ERROR", ex); continue
        keys = ("total_events", "windows", "emitted", "dropped_past_end", "ties", "last_time", "digest")
        print(c["name"], "seq", seq, "ok" if (se["digest"] == so["digest"] and se["total_events"] == so["total_events"]) else "MISMATCH")
        print("  oracle", {k: so[k] for k in keys})
        print("  engine", {k: se[k] for k in keys}, "pages", se["pages_allocated"], "far", se["far_appends"], "launches", se["kernel_launches"])
        bad = np.nonzero(ce != co)[0]
        print("  count mismatches", len(bad), bad[:10], ce[bad[:10]], co[bad[:10]])
        badh = np.nonzero(he != ho)[0]
        print("  hash mismatches", len(badh), badh[:10])
