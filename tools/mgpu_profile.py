#!/usr/bin/env python
"""Per-step device time of the multi-GPU window loop (torchrun, one rank per GPU)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from simian_b200 import capi
from simian_b200.distributed import drive_windows_p2p
from simian_b200.workloads import build_phold
import bench
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cfg = bench.workload_for(world, 1.0)
opts = {k: float(v) for k, v in (kv.split("=") for kv in sys.argv[1:])}
for rep in range(2):
    e = capi.Engine("p", 0.0, cfg["end"], cfg["min_delay"], rank=rank, size=world, device=local, seed=1, **opts)
    build_phold(e, cfg["n"], cfg["per_entity"], cfg["lookahead"], cfg["p_remote"], 1.0, cfg["mode"], cfg["groups"])
    st = drive_windows_p2p(e, rank, world, torch.device("cuda", local), profile=True,
                           allreduce=os.environ.get("SIMIAN_ALLREDUCE", "peer"))
    e.close()
if rank == 0:
    print(json.dumps({"windows": st["windows"], "driver_windows": st["driver_windows"], "events": st["total_events"],
                      "device_ms": st["device_seconds"] * 1e3, "phase_ms": st["phase_ms"],
                      "first_window_ms": st["phase_ms_first_window"]}))
dist.destroy_process_group()
