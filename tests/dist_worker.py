"""Worker of the multi-rank tests (launched by torch.distributed.run).
  --engine oracle --backend gloo : CPU, world_size 2+, the product's window
      driver (simian_b200/distributed.py) over the oracle's per-rank stepping
  --engine cuda   --backend nccl : one GPU per rank, the CUDA engine
Each rank checks ITS entities bit-for-bit against a single-process oracle run
of the same model (results are placement invariant)."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from oracle import oracle as om  # noqa: E402
from simian_b200.distributed import drive_windows  # noqa: E402
from simian_b200.workloads import build_phold  # noqa: E402

CASES = {
    "ross": dict(n=3000, per_entity=4, min_delay=0.1, lookahead=0.1, p_remote=0.3,
                 mode=1, groups=1, end=5.0, seed=4),
    "other_group": dict(n=4096, per_entity=8, min_delay=0.1, lookahead=0.1,
                        p_remote=0.5, mode=2, groups=0, end=4.0, seed=5),
    "uniform_sparse": dict(n=500, per_entity=1, min_delay=0.01, lookahead=0.01,
                           p_remote=0.0, mode=0, groups=1, end=3.0, seed=6),
    "big": dict(n=1 << 17, per_entity=16, min_delay=0.1, lookahead=0.1,
                p_remote=0.5, mode=2, groups=0, end=2.0, seed=7),
}


def build(eng, c, world):
    groups = c["groups"] or world
    build_phold(eng, c["n"], c["per_entity"], c["lookahead"], c["p_remote"], 1.0,
                c["mode"], groups)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--engine", default="oracle")
    ap.add_argument("--backend", default="gloo")
    ap.add_argument("--cases", default="ross,other_group,uniform_sparse")
    ap.add_argument("--driver", default="p2p", choices=["p2p", "p2p_nccl", "nccl"],
                    help="cuda engine: device-driven pull over NVLink, or the "
                         "host-driven NCCL alltoallv")
    a = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if a.backend == "nccl":
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist.init_process_group("gloo")
    om.build()
    for name in a.cases.split(","):
        c = CASES[name]
        # single-process reference run of the same model
        ref = om.Oracle("ref", 0.0, c["end"], c["min_delay"], size=1, seed=c["seed"])
        build(ref, c, world)
        st_ref = ref.run()
        cnt_ref = ref.read_counts("Node", c["n"])
        h_ref = ref.read_trace_hashes("Node", c["n"])
        assert st_ref["distinguishable_ties"] == 0
        if a.engine == "oracle":
            o = om.Oracle("dist", 0.0, c["end"], c["min_delay"], size=world,
                          seed=c["seed"])
            build(o, c, world)
            eng = om.OracleRankEngine(o, rank, world)
            st = drive_windows(eng, rank, world)
            cnt = o.read_counts("Node", c["n"])
            h = o.read_trace_hashes("Node", c["n"])
            mine = np.array([o.L.oracle_owner_rank(o.h, g) == rank
                             for g in range(c["n"])])
        else:
            from simian_b200 import capi
            from simian_b200.distributed import GpuRankEngine
            opts = {"bucket_width": 1.0 / 32} if c["min_delay"] < 0.05 else {}
            e = capi.Engine("dist", 0.0, c["end"], c["min_delay"], rank=rank,
                            size=world, device=local, seed=c["seed"], **opts)
            build(e, c, world)
            if a.driver in ("p2p", "p2p_nccl"):
                from simian_b200.distributed import drive_windows_p2p
                st = drive_windows_p2p(e, rank, world, torch.device("cuda", local),
                                       allreduce="peer" if a.driver == "p2p" else "nccl")
            else:
                st = drive_windows(GpuRankEngine(e, "cuda:%d" % local), rank, world)
            cnt = e.read_counts("Node", c["n"])
            h = e.read_trace_hashes("Node", c["n"])
            base = e.base_rank("Node")
            mine = (base + np.arange(c["n"])) % world == rank
            e.close()
        assert mine.sum() > 0
        np.testing.assert_array_equal(cnt[mine], cnt_ref[mine])
        np.testing.assert_array_equal(h[mine], h_ref[mine])
        assert st["windows"] == st_ref["windows"], (st["windows"], st_ref["windows"])
        assert st["driver_windows"] >= st_ref["windows"]
        tot = torch.tensor([st["total_events"], st["digest"] & 0x7FFFFFFFFFFFFFFF],
                           dtype=torch.int64)
        dig = torch.tensor([st["digest"] >> 32, st["digest"] & 0xFFFFFFFF],
                           dtype=torch.int64)
        if a.backend == "nccl":
            tot, dig = tot.cuda(), dig.cuda()
        dist.all_reduce(tot)
        assert int(tot[0]) == st_ref["total_events"], (int(tot[0]), st_ref["total_events"])
        # digest is a wrapping sum over entities: combine across ranks mod 2^64
        parts = [torch.zeros_like(dig) for _ in range(world)]
        dist.all_gather(parts, dig)
        full = sum((int(p[0]) << 32) | int(p[1]) for p in parts) & 0xFFFFFFFFFFFFFFFF
        assert full == st_ref["digest"], (hex(full), hex(st_ref["digest"]))
        if rank == 0:
            print("DIST-OK %s world=%d events=%d windows=%d" % (
                name, world, st_ref["total_events"], st_ref["windows"]), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
