#!/usr/bin/env python
"""Per-phase SM-cycle breakdown of k_window on BASELINE config 2.
Needs a library built with -DSIMIAN_PROFILE_PHASES, e.g.
  make -C simian_b200/csrc variant NAME=prof DEFS="-DSIMIAN_PROFILE_PHASES"
  SIMIAN_B200_LIB=$PWD/simian_b200/libsimian_b200_prof.so python tools/phase_profile.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simian_b200 import capi  # noqa: E402
from simian_b200.workloads import CONFIGS, build_phold  # noqa: E402

c = dict(CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "config2_phold_1m"])
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
opts = c.pop("options", {})
for rep in range(2):
    if c.get("kind") == "lanl":
        from simian_b200.workloads import build_lanl, lanl_engine_end
        k = {x: c[x] for x in c if x != "kind"}
        k["n_ent"] = max(64, int(k["n_ent"] * scale))
        e = capi.Engine("p", 0.0, lanl_engine_end(k["end_time"], k["min_delay"]), k["min_delay"],
                        seed=1, **opts)
        build_lanl(e, **k)
    else:
        e = capi.Engine("p", 0.0, c["end"], c["min_delay"], seed=1, **opts)
        build_phold(e, max(1, int(c["n"] * scale)), c["per_entity"], c["lookahead"],
                    c["p_remote"], 1.0, c["mode"], c["groups"])
    st = e.run()
    ph = e.debug_phase_cycles()
    e.close()
tot = sum(ph) or 1
names = ["0 init/W/page cache", "1 list pages", "2 scan -> smem slots", "3 A1 rank",
         "4 A2 handler+stage", "5 B append", "6 C chain", "7 recycle", "8 reduce/part",
         "9 advance(last CTA)"]
print("events %d dev %.2f ms  ev/s %.3e  launches %d" % (
    st["total_events"], st["device_seconds"] * 1e3,
    st["total_events"] / st["device_seconds"], st["kernel_launches"]))
for n, v in zip(names, ph):
    print("%-22s %6.2f%%  %12d cycles" % (n, 100.0 * v / tot, v))
print("total cycles %d over %d windows" % (tot, st["windows"]))
