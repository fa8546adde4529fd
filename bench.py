#!/usr/bin/env python
"""bench.py -- committed PHOLD events/sec on N B200s (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this framework
  python bench.py --impl reference --gpus N --steps K ...  # reference CPU path

A step = one complete run of the window loop (Simian.run, simian.js:99-168)
over one freshly scheduled synthetic PHOLD workload:
  N=1 : BASELINE config 2 -- 2^20 entities x 16 events, Exp(1)+0.1 delays,
        10% remote, endTime 10 (~1.7e8 committed events per step)
  N>1 : BASELINE config 3 shape, weak scaling -- 2^21 entities per GPU x 16
        events, 50% of emitted events cross GPUs, endTime 5
`value`  : events / CUDA-event time of the window loop, initial events already
           resident in HBM (the reference times run() only, simian.js:100,159)
`e2e`    : the same through the C-ABI with HOST buffers: pinned host records
           -> simian_sched_events (H2D) -> run -> per-entity counts + trace
           hashes read back (D2H), wall clock.
One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BYTES_PER_EVENT = 96  # 32 read + 32 write + 16 + 16 state
# dram__bytes_read.sum + dram__bytes_write.sum of one steady-state k_window launch
# (config 2, ~1.68 M committed events): profiles/r1_s_k_window_ncu_full.csv
NCU_TRAFFIC_BYTES_PER_LAUNCH = 79651328 + 45706496


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"  # B200_PROFILING.md fallback


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def mark(self):
        """the timed region starts now (the sampler itself starts before the
        warm-up: nvidia-smi needs ~100 ms to deliver its first sample, more than
        a short timed region lasts)"""
        self.t_mark = time.perf_counter()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)  # one more sampling period
        self.proc.terminate()
        self.t.join(timeout=2)
        t0 = getattr(self, "t_mark", 0.0)
        rows = [r for t, r in self.rows if t >= t0]
        window = "timed"
        if not rows:  # timed region shorter than the sampling period
            rows, window = [r for _, r in self.rows], "warmup+timed"
        sm = [float(r[0]) for r in rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows if len(r) >= 6
                          for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm), "window": window}


def workload_for(n_gpus, scale):
    from simian_b200.workloads import CONFIGS, PHOLD_OTHER_GROUP
    if n_gpus == 1:
        c = dict(CONFIGS["config2_phold_1m"])
        c["label"] = "config2: PHOLD 2^20 entities x16 events, Exp(1)+0.1, 10% remote, endTime 10"
    else:
        c = dict(CONFIGS["config3_phold_16m"])
        c["n"] = (1 << 21) * n_gpus
        c["groups"] = n_gpus
        c["mode"] = PHOLD_OTHER_GROUP
        c["label"] = ("config3 shape (weak): PHOLD 2^21 entities/GPU x16 events, "
                      "50%% cross-GPU, endTime 5, %d GPUs" % n_gpus)
    if scale != 1.0:
        c["n"] = max(c["groups"], int(c["n"] * scale) // c["groups"] * c["groups"])
        c["label"] += " [scaled x%g]" % scale
    return c


def host_cores():
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    return max(1, min(n, 64))


def cpu_baseline_sample(cfg, cores=None):
    """The oracle (CPU port of the reference algorithm) on a bounded sample of
    the same workload shape -- fewer entities, same per-entity load and delays --
    run the way the reference runs on a multi-core host: one rank per core
    (mpirun -np <cores>; here the ranks of one oracle world on host threads,
    meeting at the window's two collectives, simian.js:136-145)."""
    from oracle import oracle as om
    from simian_b200.workloads import build_phold
    om.build()
    cores = cores or host_cores()
    groups = max(cfg["groups"], 1)
    n = min(cfg["n"], (1 << 16) * max(1, min(cores, 8)))
    unit = groups * cores // __import__("math").gcd(groups, cores)
    n = max(unit, n // unit * unit)
    o = om.Oracle("cpu_baseline", 0.0, cfg["end"], cfg["min_delay"], size=cores,
                  seed=1, threads=cores)
    build_phold(o, n, cfg["per_entity"], cfg["lookahead"], cfg["p_remote"], 1.0,
                cfg["mode"], cfg["groups"])
    st = o.run()
    o.close()
    return {"value": st["total_events"] / st["run_seconds"], "unit": "events/s",
            "cores": cores, "kind": "port",
            "sample": "%d entities x%d events, endTime %g: %d events in %.2f s "
                      "(oracle/simian_oracle.cpp, %d ranks on %d host threads)" % (
                          n, cfg["per_entity"], cfg["end"], st["total_events"],
                          st["run_seconds"], cores, cores),
            "events": st["total_events"], "seconds": st["run_seconds"]}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path (the
    oracle port; SimianJS/Lua cannot run in this image) on all host cores, one
    rank per core as under mpirun."""
    if rank != 0:
        return
    cfg = workload_for(args.gpus, args.scale)
    for _ in range(args.warmup):
        cpu_baseline_sample(cfg)
    ev = sec = 0.0
    last = None
    for _ in range(args.steps):
        last = cpu_baseline_sample(cfg)
        ev += last["events"]
        sec += last["seconds"]
    v = ev / sec
    line = {"impl": "reference", "metric": "committed_events_per_sec", "value": v,
            "unit": "events/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * sec / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": cfg["label"], "sample": last["sample"]},
            "cpu_baseline": {"value": v, "unit": "events/s", "cores": last["cores"],
                             "kind": "port", "sample": last["sample"]},
            "e2e": {"value": v, "unit": "events/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def make_host_events(cfg, torch, np, capi, rank=0, world=1):
    """The initial schedService calls of the model script as POD records in
    PINNED host memory (phold-noop-noMPI.js:47), seq in call order.  world > 1:
    the calls this rank's process acts on (only the owner pushes, simian.js:177)."""
    n, per = cfg["n"], cfg["per_entity"]
    base = int(capi.lib().simian_hash_name(b"Node") % world)  # simian.js:192-194
    ids = np.arange((rank - base) % world, n, world, dtype=np.uint32)
    total = len(ids) * per
    buf = torch.empty(total * 32, dtype=torch.uint8, pin_memory=True)
    ev = buf.numpy().view(capi.EVENT_DTYPE)
    ev["time"] = 0.0
    ev["dst"] = np.tile(ids, per)
    ev["src"] = capi.NULL_ENTITY
    ev["handler"] = 0
    ev["flags"] = 1
    ev["seq"] = (np.repeat(np.arange(per, dtype=np.uint32), len(ids)) * np.uint32(n)
                 + np.tile(ids, per))
    ev["data"] = 0
    return buf, ev


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--scale", type=float, default=1.0,
                    help="scale the entity count (parity/debug only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    from simian_b200 import capi
    from simian_b200.workloads import build_phold
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    # stdout carries ONE JSON line: whatever libraries print there while the
    # benchmark runs (NCCL's version banner at its first collective) goes to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)

    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    cfg = workload_for(world, args.scale)
    peak, peak_kind = load_peaks()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step_resident():
        """inputs resident: schedService loop built on the device, then run."""
        if world == 1:
            e = capi.Engine("bench", 0.0, cfg["end"], cfg["min_delay"], seed=1,
                            device=local_rank)
            build_phold(e, cfg["n"], cfg["per_entity"], cfg["lookahead"],
                        cfg["p_remote"], 1.0, cfg["mode"], cfg["groups"])
            st = e.run()
            e.close()
            return st
        from simian_b200.distributed import run_distributed_phold
        st = run_distributed_phold(cfg, rank, world, local_rank)
        st.pop("_engine").close()
        return st

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        one_step_resident()
    barrier()
    sampler.mark()
    dev_s, wall0, events, launches, windows = 0.0, time.perf_counter(), 0, 0, 0
    for _ in range(args.steps):
        st = one_step_resident()
        dev_s += st["device_seconds"]
        events += st["total_events"]
        launches += st["kernel_launches"]
        windows += st["windows"]
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop() if rank == 0 else None
    if dist is not None:
        t = torch.tensor([dev_s, float(events), float(launches)], dtype=torch.float64,
                         device="cuda")
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        dev_s, events, launches = float(tmax[0]), int(t[1]), int(t[2])
    value = events / dev_s

    # ---- end to end through the C-ABI with HOST buffers ----------------------
    # every step: pinned host records -> simian_sched_events (H2D) -> run ->
    # per-entity counts + trace hashes of the rank's entities read back (D2H)
    e2e = None
    if not args.no_e2e:
        buf, ev = make_host_events(cfg, torch, np, capi, rank, world)
        n = cfg["n"]
        n_mine = len(ev) // cfg["per_entity"]
        # the step's results land in page-locked host buffers the caller owns
        res = torch.empty((2, n_mine), dtype=torch.int64, pin_memory=True).numpy().view(np.uint64)

        def one_step_e2e():
            from simian_b200.workloads import phold_params_blob
            e = capi.Engine("bench_e2e", 0.0, cfg["end"], cfg["min_delay"], seed=1,
                            rank=rank, size=world, device=local_rank)
            e.define_class("Node", 0)
            e.add_entities("Node", 0, n)
            e.attach_service("Node", "generate", "phold_generate")
            e.set_class_params("Node", phold_params_blob(
                0, n, cfg["p_remote"], 1.0, cfg["lookahead"], cfg["mode"], cfg["groups"]))
            e.sched_events(ev)                      # H2D of the step's inputs
            if world == 1:
                st = e.run()
            else:
                from simian_b200.distributed import drive_windows_p2p
                st = drive_windows_p2p(e, rank, world, torch.device("cuda", local_rank))
            counts = e.read_local("Node", 0, out=res[0])[0]   # D2H of the step's results
            hashes = e.read_local("Node", 1, out=res[1])[0]
            e.close()
            return st, counts, hashes
        one_step_e2e()
        barrier()
        t0 = time.perf_counter()
        ev_e2e = 0
        ksteps = max(1, min(args.steps, 3))
        for _ in range(ksteps):
            st2, counts, hashes = one_step_e2e()
            ev_e2e += st2["total_events"]
        barrier()
        t_e2e = time.perf_counter() - t0
        assert int(counts.sum()) == st2["total_events"]
        n_local = len(ev) // cfg["per_entity"]
        if dist is not None:
            t = torch.tensor([t_e2e, float(ev_e2e)], dtype=torch.float64, device="cuda")
            tmax = t.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            t_e2e, ev_e2e = float(tmax[0]), int(t[1])
        e2e = {"value": ev_e2e / t_e2e, "unit": "events/s",
               "h2d_bytes_per_step": int(ev.nbytes + 48) * world,
               "d2h_bytes_per_step": int(16 * n_local + 160) * world,
               "steps": ksteps, "ms_per_step": 1e3 * t_e2e / ksteps}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    per_gpu = value / world
    achieved = per_gpu * BYTES_PER_EVENT / 1e9
    line = {
        "metric": "committed_events_per_sec", "value": value, "unit": "events/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dev_s / args.steps,
        "wall_ms_per_step": 1e3 * wall / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": cfg["label"], "entities": cfg["n"],
                   "events_per_entity": cfg["per_entity"], "min_delay": cfg["min_delay"],
                   "end_time": cfg["end"], "events_per_step": events // args.steps,
                   "windows_per_step": windows // args.steps,
                   "l2_policy": "inputs_larger_than_l2 (event pool >= 512 MB per GPU)",
                   "parity": "bit-exact vs oracle at reduced sizes (tests/), "
                             "0 distinguishable ties"},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak,
                     "traffic": NCU_TRAFFIC_BYTES_PER_LAUNCH if world == 1 else None,
                     "algorithmic_bytes_per_launch":
                         BYTES_PER_EVENT * (events // max(windows, 1) // world),
                     "kernel": "k_window", "peak_kind": peak_kind,
                     "bytes_per_event": BYTES_PER_EVENT,
                     "note": "achieved = committed events/s per GPU x 96 B over the whole "
                             "timed window loop (k_window is >99% of it); traffic = "
                             "dram read+write bytes of one steady-state k_window launch, "
                             "ncu --set full, profiles/r1_s_k_window_ncu_full.csv"},
    }
    if e2e is not None:
        line["e2e"] = e2e
    if not args.no_cpu_baseline and world == 1:
        cb = cpu_baseline_sample(cfg)
        cb.pop("events"), cb.pop("seconds")
        line["cpu_baseline"] = cb
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
