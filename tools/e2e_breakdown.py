#!/usr/bin/env python
"""Where the end-to-end step of bench.py spends its time (host wall clock with
a device sync after each stage)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from simian_b200 import capi
from simian_b200.workloads import CONFIGS, phold_params_blob
import bench
cfg = dict(CONFIGS["config2_phold_1m"])
buf, ev = bench.make_host_events(cfg, torch, np, capi)
n = cfg["n"]
for rep in range(4):
    T = {}
    def lap(k, t0):
        torch.cuda.synchronize(); T[k] = (time.perf_counter() - t0) * 1e3
    t = time.perf_counter(); e = capi.Engine("x", 0.0, cfg["end"], cfg["min_delay"], seed=1)
    e.define_class("Node", 0); e.add_entities("Node", 0, n); e.attach_service("Node", "generate", "phold_generate")
    e.set_class_params("Node", phold_params_blob(0, n, cfg["p_remote"], 1.0, cfg["lookahead"], cfg["mode"], cfg["groups"]))
    lap("create", t)
    t = time.perf_counter(); e.sched_events(ev); lap("sched_events(H2D)", t)
    t = time.perf_counter(); e.setup(); lap("setup(alloc+init+ingest)", t)
    t = time.perf_counter(); st = e.run(); lap("run", t)
    t = time.perf_counter(); c = e.read_counts("Node", n); lap("read_counts", t)
    t = time.perf_counter(); h = e.read_trace_hashes("Node", n); lap("read_hashes", t)
    t = time.perf_counter(); e.close(); lap("close", t)
    print(" ".join("%s=%.2fms" % kv for kv in T.items()), "| device run %.2f ms" % (st["device_seconds"] * 1e3), "total %.1f" % sum(T.values()))
