#!/usr/bin/env python
"""Phase breakdown of the t = 0 burst window alone (config 2 with endTime < minDelay).
Needs libsimian_b200_prof1.so (-DSIMIAN_PROFILE_PHASES -DSIMIAN_PROFILE_KEEP_FIRST)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simian_b200 import capi
from simian_b200.workloads import CONFIGS, build_phold
c = dict(CONFIGS["config2_phold_1m"])
for rep in range(2):
    e = capi.Engine("p", 0.0, 0.05, c["min_delay"], seed=1)
    build_phold(e, c["n"], c["per_entity"], c["lookahead"], c["p_remote"], 1.0, c["mode"], c["groups"])
    st = e.run()
    ph = e.debug_phase_cycles()
    e.close()
tot = sum(ph) or 1
names = ["0 init", "1 list pages", "2 scan", "3 A1 rank", "4 A2 handler", "5 B append", "6 C chain", "7 recycle", "8 reduce", "9 advance"]
print("events %d windows %d dev %.2f ms" % (st["total_events"], st["windows"], st["device_seconds"] * 1e3))
for n, v in zip(names, ph):
    print("%-16s %6.2f%% %14d" % (n, 100.0 * v / tot, v))
