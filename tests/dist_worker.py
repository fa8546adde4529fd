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


LANL = dict(kind="lanl", n_ent=2100, s_ent=12.0, q_avg=2.0, p_receive=0.002, m_ent=24.0,
            ops_ent=30.0, ops_sigma=0.2, cache_friendliness=0.5, end_time=8.0,
            min_delay=0.5, time_bins=16, seed=4)
HELLO = dict(kind="hello", count=300, end=200.0, min_delay=0.5, seed=3)
# payloads beyond 8 bytes: headers + continuation records through the outboxes
BLOBS = dict(kind="blobs", n=1500, end=6.0, min_delay=0.25, lookahead=0.25, max_words=12, seed=9)
CASES["lanl"] = LANL
CASES["hello"] = HELLO
CASES["blobs"] = BLOBS


def engine_options(c):
    """CUDA engine options a case needs"""
    if c.get("kind") == "hello":
        return {"bucket_width": 1.0}
    if c.get("kind") == "blobs":
        return {"payloads": 1}
    return {"bucket_width": 1.0 / 32} if c.get("min_delay", 1.0) < 0.05 else {}


def build(eng, c, world):
    if c.get("kind") == "lanl":
        from simian_b200.workloads import build_lanl
        keys = ("n_ent", "s_ent", "q_avg", "p_receive", "m_ent", "ops_ent", "ops_sigma",
                "cache_friendliness", "end_time", "min_delay", "time_bins")
        build_lanl(eng, **{k: c[k] for k in keys})
        return
    if c.get("kind") == "hello":
        from simian_b200.workloads import build_hello
        build_hello(eng, c["count"])
        return
    if c.get("kind") == "blobs":
        from simian_b200.workloads import build_blobs
        build_blobs(eng, c["n"], c["lookahead"], c["max_words"])
        return
    groups = c["groups"] or world
    build_phold(eng, c["n"], c["per_entity"], c["lookahead"], c["p_remote"], 1.0,
                c["mode"], groups)


def shape(c):
    """-> (engine end time, minDelay, [(class, count)], oracle tie mode)"""
    if c.get("kind") == "lanl":
        from simian_b200.workloads import LANL_CLASS, lanl_engine_end
        return (lanl_engine_end(c["end_time"], c["min_delay"]), c["min_delay"],
                [(LANL_CLASS, c["n_ent"])], 1)
    if c.get("kind") == "hello":
        return c["end"], c["min_delay"], [("Alice", c["count"]), ("Bob", c["count"])], 1
    if c.get("kind") == "blobs":
        return c["end"], c["min_delay"], [("Node", c["n"])], 0
    return c["end"], c["min_delay"], [("Node", c["n"])], 0


def read_all(h, classes):
    cnt = np.concatenate([h.read_counts(k, n) for k, n in classes])
    hsh = np.concatenate([h.read_trace_hashes(k, n) for k, n in classes])
    return cnt, hsh


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--engine", default="oracle")
    ap.add_argument("--backend", default="gloo")
    ap.add_argument("--cases", default="ross,other_group,uniform_sparse")
    ap.add_argument("--driver", default="p2p", choices=["p2p", "p2p_deferred", "p2p_nccl", "nccl"],
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
        end, min_delay, classes, tie_mode = shape(c)
        # single-process reference run of the same model
        ref = om.Oracle("ref", 0.0, end, min_delay, size=1, seed=c["seed"], tie_mode=tie_mode)
        build(ref, c, world)
        st_ref = ref.run()
        cnt_ref, h_ref = read_all(ref, classes)
        if tie_mode == 0:
            assert st_ref["distinguishable_ties"] == 0
        if a.engine == "oracle":
            o = om.Oracle("dist", 0.0, end, min_delay, size=world, seed=c["seed"],
                          tie_mode=tie_mode)
            build(o, c, world)
            eng = om.OracleRankEngine(o, rank, world)
            st = drive_windows(eng, rank, world)
            cnt, h = read_all(o, classes)
            gid0 = 0
            mine = []
            for k, n in classes:
                mine += [o.L.oracle_owner_rank(o.h, gid0 + g) == rank for g in range(n)]
                gid0 += n
            mine = np.array(mine)
        else:
            from simian_b200 import capi
            from simian_b200.distributed import GpuRankEngine
            opts = {"bucket_width": 1.0 / 32} if min_delay < 0.05 else {}
            if c.get("kind") == "hello":
                opts = {"bucket_width": 1.0}
            if a.driver == "p2p_deferred":
                opts["deferred_pull"] = 1
            e = capi.Engine("dist", 0.0, end, min_delay, rank=rank,
                            size=world, device=local, seed=c["seed"], **opts)
            build(e, c, world)
            if a.driver in ("p2p", "p2p_deferred", "p2p_nccl"):
                from simian_b200.distributed import drive_windows_p2p
                st = drive_windows_p2p(e, rank, world, torch.device("cuda", local),
                                       allreduce="nccl" if a.driver == "p2p_nccl" else "peer")
            else:
                st = drive_windows(GpuRankEngine(e, "cuda:%d" % local), rank, world)
            cnt, h = read_all(e, classes)
            mine = np.concatenate([(e.base_rank(k) + np.arange(n)) % world == rank
                                   for k, n in classes])
            # the compact per-rank read-back returns the same values
            off = 0
            for k, n in classes:
                loc, first, stride = e.read_local(k, 0)
                np.testing.assert_array_equal(loc, cnt[off:off + n][first::stride])
                loc, first, stride = e.read_local(k, 1)
                np.testing.assert_array_equal(loc, h[off:off + n][first::stride])
                off += n
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
