#!/usr/bin/env python
"""One multi-rank case with every rank on ONE GPU (device-driven exchange, peers
wired by device pointers): python tools/loopback_case.py <case> <world> <deferred 0|1>
Cases: tests/dist_worker.py CASES.  Prints the per-rank statistics."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
from dist_worker import CASES, build, shape  # noqa: E402
from simian_b200 import capi  # noqa: E402
from simian_b200.distributed import drive_windows_p2p_loopback  # noqa: E402

name, world, deferred = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
c = CASES[name]
end, min_delay, classes, tie_mode = shape(c)
opts = {"bucket_width": 1.0 / 32} if min_delay < 0.05 else {}
if c.get("kind") == "hello":
    opts = {"bucket_width": 1.0}
engines = []
for rank in range(world):
    e = capi.Engine("p2p", 0.0, end, min_delay, rank=rank, size=world, device=0,
                    seed=c["seed"], deferred_pull=deferred, **opts)
    build(e, c, world)
    engines.append(e)
stats = drive_windows_p2p_loopback(engines, torch.device("cuda", 0))
for r, st in enumerate(stats):
    print(r, st["total_events"], st["windows"], "%016x" % st["digest"], st["error"])
