#!/usr/bin/env python
"""Debug: per-window outbox counts of the e2e-shaped and resident-shaped engines (torchrun)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from simian_b200 import capi
from simian_b200.distributed import connect_peers
from simian_b200.workloads import build_phold, phold_params_blob
import bench
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.0625
cfg = bench.workload_for(world, scale)
n = cfg["n"]
for shape in ("resident", "e2e"):
    e = capi.Engine("dbg", 0.0, cfg["end"], cfg["min_delay"], rank=rank, size=world, device=local, seed=1)
    if shape == "resident":
        build_phold(e, n, cfg["per_entity"], cfg["lookahead"], cfg["p_remote"], 1.0, cfg["mode"], cfg["groups"])
    else:
        buf, ev = bench.make_host_events(cfg, torch, np, capi, rank, world)
        e.define_class("Node", 0); e.add_entities("Node", 0, n); e.attach_service("Node", "generate", "phold_generate")
        e.set_class_params("Node", phold_params_blob(0, n, cfg["p_remote"], 1.0, cfg["lookahead"], cfg["mode"], cfg["groups"]))
        e.sched_events(ev)
    e.set_stream(torch.cuda.current_stream().cuda_stream)
    connect_peers(e, rank, world)
    e.run_begin()
    log = []
    for w in range(12):
        e.window_process_lbts(True)
        torch.cuda.synchronize()
        counts, ptr, cap = e.window_outbox()
        log.append((max(counts), sum(counts)))
        e.window_pull(True)
        e.window_advance_async(True)
        try:
            done = e.poll()
        except Exception as ex:
            log.append(("ERR", str(ex)[:60])); break
        if done: break
    st = e.run_end(check=False)
    dist.barrier()
    print("rank %d %s cap %d events %d windows %d err %d counts %s" % (rank, shape, cap, st["total_events"], st["windows"], st["error"], log[:8]), flush=True)
    e.close()
dist.destroy_process_group()
