#!/usr/bin/env python
"""A/B probe: median device time of the whole config run and of its t = 0 burst
window alone.  usage: python tools/ab_run.py <config> [reps] [opt=value ...]"""
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simian_b200 import capi  # noqa: E402
from simian_b200.workloads import CONFIGS, build_phold  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "config2_phold_1m"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
opts = {k: float(v) for k, v in (kv.split("=") for kv in sys.argv[3:])}
c = dict(CONFIGS[name])
opts = dict(c.pop("options", {}), **opts)


def run(end):
    ts = []
    for rep in range(reps + 1):
        e = capi.Engine(name, 0.0, end, c["min_delay"], seed=1, **opts)
        build_phold(e, c["n"], c["per_entity"], c["lookahead"], c["p_remote"], 1.0,
                    c["mode"], c["groups"])
        st = e.run()
        e.close()
        if rep:
            ts.append(st["device_seconds"] * 1e3)
    return st, ts


st1, t1 = run(c["min_delay"] * 0.5)
st, t = run(c["end"])
print("%s: full med %.3f min %.3f ms (%.3e ev/s) | burst med %.3f min %.3f ms | events %d windows %d" % (
    os.path.basename(os.environ.get("SIMIAN_B200_LIB", "default")), statistics.median(t), min(t),
    st["total_events"] / (statistics.median(t) * 1e-3), statistics.median(t1), min(t1),
    st["total_events"], st["windows"]))
