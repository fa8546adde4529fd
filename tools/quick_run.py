#!/usr/bin/env python
"""Quick throughput probe: python tools/quick_run.py <config> <scale> [opt=value ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simian_b200 import capi  # noqa: E402
from simian_b200.workloads import CONFIGS, build_phold  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "config2_phold_1m"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
opts = dict(kv.split("=") for kv in sys.argv[3:])
opts = {k: float(v) for k, v in opts.items()}
c = dict(CONFIGS[name])
opts = dict(c.pop("options", {}), **opts)
if c.get("kind") == "lanl":
    from simian_b200.workloads import build_lanl, lanl_engine_end  # noqa: E402
    c.pop("kind")
    c["n_ent"] = max(1, int(c["n_ent"] * scale))
    for rep in range(2):
        e = capi.Engine(name, 0.0, lanl_engine_end(c["end_time"], c["min_delay"]),
                        c["min_delay"], seed=1, **opts)
        build_lanl(e, **c)
        st = e.run()
        e.close()
else:
    c["n"] = max(c["groups"], int(c["n"] * scale) // c["groups"] * c["groups"])
    for rep in range(3):
        e = capi.Engine(name, 0.0, c["end"], c["min_delay"], seed=1, **opts)
        build_phold(e, c["n"], c["per_entity"], c["lookahead"], c["p_remote"], 1.0,
                    c["mode"], c["groups"])
        st = e.run()
        e.close()
print("%s x%g %s: events %d windows %d dev %.2f ms  %.3e ev/s  pages %d far %d redist %d" % (
    name, scale, opts, st["total_events"], st["windows"], st["device_seconds"] * 1e3,
    st["total_events"] / st["device_seconds"], st["pages_allocated"], st["far_appends"],
    st["redistributions"]))
