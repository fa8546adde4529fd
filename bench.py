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
`config.parity` : digest (wrapping sum over all entities of the per-entity
           checksum of committed count and order-sensitive trace hash) and total
           events of the last timed step, summed over the ranks.  The reference
           arm prints the same two numbers from the CPU oracle on the same
           configuration (N=1: the full config 2), so the two arms can be
           compared bit for bit.
`extra_configs` : the other BASELINE configs measured after the headline
           (config 5 tiny windows, config 1 reference's own case, config 4 LANL).
One JSON line on stdout (rank 0).
"""
import argparse
import glob
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BYTES_PER_EVENT = 96  # 32 read + 32 write + 16 + 16 state
NVLINK_PEAK_GBS = 770.0  # measured peer copy per direction (B200_PROFILING.md); nominal 900


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"  # B200_PROFILING.md fallback


def ncu_traffic():
    """dram read + write bytes of one steady-state k_window launch, from the
    newest ncu summary committed under profiles/ (not measured in this run: a
    run under ncu is never a bench value).  Returns (bytes or None, note)."""
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_k_window_ncu_full.csv")),
                   key=lambda f: (os.path.basename(f).split("_")[0:2]))
    files = [f for f in files if "warp_groups" not in f]
    if not files:
        return None, "no ncu summary under profiles/"
    f = files[-1]
    rd = wr = None
    for line in open(f):
        c = line.strip().split(",")
        if c[0] == "dram__bytes_read.sum":
            rd = float(c[2]) * (1e6 if c[1] == "Mbyte" else 1e9 if c[1] == "Gbyte" else 1.0)
        if c[0] == "dram__bytes_write.sum":
            wr = float(c[2]) * (1e6 if c[1] == "Mbyte" else 1e9 if c[1] == "Gbyte" else 1.0)
    if rd is None or wr is None:
        return None, "no dram byte counters in %s" % os.path.basename(f)
    return int(rd + wr), ("dram__bytes_read.sum + dram__bytes_write.sum of one steady-state "
                          "k_window launch of config 2 (ncu --set full), profiles/%s"
                          % os.path.basename(f))


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


# ---- workloads -----------------------------------------------------------------

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


def extra_workloads(n_gpus, scale):
    """the BASELINE configs beside the headline: (key, cfg) in measurement order"""
    from simian_b200.workloads import CONFIGS, PHOLD_OTHER_GROUP
    out = []
    if n_gpus == 1:
        # same shape as the N>1 workload on one GPU: the base of the weak-scaling curve
        a = dict(CONFIGS["config3_phold_16m"])
        a.update(n=1 << 21, groups=2, mode=PHOLD_OTHER_GROUP,
                 label="scale anchor: config3 shape on ONE GPU, 2^21 entities x16, endTime 5 "
                       "(50% of sends leave the sender's model-level group; all local)")
        out.append(("scale_anchor", a))
    c5 = dict(CONFIGS["config5_phold_4m_small_lookahead"])
    c5["label"] = ("config5: PHOLD 2^22 entities x16, minDelay 0.01 (tiny windows), 10%% remote, "
                   "endTime 2%s" % (", sharded over %d GPUs" % n_gpus if n_gpus > 1 else ""))
    out.append(("config5", c5))
    if n_gpus == 1:
        c1 = dict(CONFIGS["config1_phold_1k"])
        c1["label"] = "config1: PHOLD 1k entities x1 event, minDelay 1e-4, endTime 1000 (reference's own case)"
        c1["end"] = 10.0  # 100 k windows of 1e-4: the same windows/s figure in a bounded run
        c1["label"] += " [endTime 10 here]"
        out.append(("config1", c1))
    c4 = dict(CONFIGS["config4_lanl"])
    c4["n_ent"] = (1 << 20) * n_gpus if n_gpus > 1 else 1 << 20
    # weak scaling keeps the hot spots as hot per entity block: the geometric
    # destination parameter shrinks with the entity count (mean destination index
    # = n_ent / 10.5 at every N); with p_receive fixed, entity 0 alone would receive
    # N times the events per window and its block overflow the CTA's page list
    c4["p_receive"] = c4["p_receive"] / n_gpus
    c4["label"] = ("config4: LANL PDES benchmark V8-style, %d entities%s, 8-KB list per entity, "
                   "~1000 ops per receive, geometric destinations, endTime 100"
                   % (c4["n_ent"], " over %d GPUs" % n_gpus if n_gpus > 1 else ""))
    out.append(("config4", c4))
    if scale != 1.0:
        for _, c in out:
            if c.get("kind") == "lanl":
                c["n_ent"] = max(64, int(c["n_ent"] * scale))
            else:
                c["n"] = max(c["groups"], int(c["n"] * scale) // c["groups"] * c["groups"])
            c["label"] += " [scaled x%g]" % scale
    return out


def is_lanl(cfg):
    return cfg.get("kind") == "lanl"


def engine_times(cfg):
    if is_lanl(cfg):
        from simian_b200.workloads import lanl_engine_end
        return lanl_engine_end(cfg["end_time"], cfg["min_delay"]), cfg["min_delay"]
    return cfg["end"], cfg["min_delay"]


def build_model(eng, cfg):
    if is_lanl(cfg):
        from simian_b200.workloads import build_lanl
        keys = ("n_ent", "s_ent", "q_avg", "p_receive", "m_ent", "ops_ent", "ops_sigma",
                "cache_friendliness", "end_time", "min_delay", "time_bins")
        return build_lanl(eng, **{k: cfg[k] for k in keys})
    from simian_b200.workloads import build_phold
    return build_phold(eng, cfg["n"], cfg["per_entity"], cfg["lookahead"], cfg["p_remote"],
                       1.0, cfg["mode"], cfg["groups"])


def bytes_per_event(cfg):
    """algorithmic HBM bytes per committed event"""
    if not is_lanl(cfg):
        return BYTES_PER_EVENT
    # per event: record read + write + entity core; per receive (about half of the
    # events): 8 bytes for each active list element + the entity's 168-byte state
    from simian_b200.workloads import lanl_list_size
    active = cfg["cache_friendliness"] * lanl_list_size(cfg["n_ent"], cfg["m_ent"])
    return BYTES_PER_EVENT + 0.5 * (8.0 * active + 2 * 168.0)


def host_cores():
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    return max(1, min(n, 64))


def cpu_run(cfg, cores=None, full=False):
    """The oracle (CPU port of the reference algorithm) on the workload -- the
    whole of it (`full`) or a bounded sample of the same shape (fewer entities,
    same per-entity load and delays) -- run the way the reference runs on a
    multi-core host: one rank per core (mpirun -np <cores>; here the ranks of one
    oracle world on host threads, meeting at the window's two collectives,
    simian.js:136-145)."""
    from oracle import oracle as om
    om.build()
    cores = cores or host_cores()
    cfg = dict(cfg)
    end, min_delay = engine_times(cfg)
    if is_lanl(cfg):
        n_full = cfg["n_ent"]
        n = n_full if full else min(n_full, 1 << 12)
        cfg["n_ent"] = n
        per = "~200"
    else:
        groups = max(cfg["groups"], 1)
        n_full = cfg["n"]
        n = n_full if full else min(n_full, (1 << 16) * max(1, min(cores, 8)))
        if not full:
            unit = groups * cores // math.gcd(groups, cores)
            n = max(unit, n // unit * unit)
        cfg["n"] = n
        per = "x%d" % cfg["per_entity"]
    o = om.Oracle("cpu_baseline", 0.0, end, min_delay, size=cores, seed=1, threads=cores)
    build_model(o, cfg)
    st = o.run()
    o.close()
    whole = n == n_full
    return {"value": st["total_events"] / st["run_seconds"], "unit": "events/s",
            "cores": cores, "kind": "port", "same_config": whole,
            "sample": "%s: %d entities %s events, endTime %g: %d events in %.2f s "
                      "(oracle/simian_oracle.cpp, %d ranks on %d host threads)" % (
                          "the whole workload" if whole else "a sample of the workload",
                          n, per, end, st["total_events"], st["run_seconds"], cores, cores),
            "events": st["total_events"], "seconds": st["run_seconds"],
            "digest": "%016x" % st["digest"], "windows": st["windows"]}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path (the
    oracle port; SimianJS/Lua cannot run in this image) on all host cores, one
    rank per core as under mpirun.  N=1: the WHOLE of config 2, so that its digest
    and event count can be compared with the native arm's; N>1: a sample of the
    config-3 shape (the whole of it is minutes per step on the host)."""
    if rank != 0:
        return
    cfg = workload_for(args.gpus, args.scale)
    full = args.gpus == 1
    for _ in range(args.warmup):
        cpu_run(cfg, full=full)
    ev = sec = 0.0
    last = None
    for _ in range(args.steps):
        last = cpu_run(cfg, full=full)
        ev += last["events"]
        sec += last["seconds"]
    v = ev / sec
    line = {"impl": "reference", "metric": "committed_events_per_sec", "value": v,
            "unit": "events/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * sec / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": cfg["label"], "sample": last["sample"],
                       "same_config": last["same_config"],
                       "parity": {"digest": last["digest"], "total_events": last["events"],
                                  "windows": last["windows"]}},
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
    ap.add_argument("--pull", default="all", choices=["deferred", "all"],
                    help="N > 1: cross-GPU pull overlapped with the next window (two-ended "
                         "outboxes) or everything between two windows")
    ap.add_argument("--no-extra", action="store_true",
                    help="skip the extra_configs (configs 1, 4, 5, scale anchor)")
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

    def one_step_resident(c, profile=False):
        """inputs resident: schedService loop built on the device, then run."""
        end, min_delay = engine_times(c)
        opts = dict(c.get("options", {}))
        if world == 1:
            e = capi.Engine("bench", 0.0, end, min_delay, seed=1, device=local_rank, **opts)
            build_model(e, c)
            st = e.run()
            e.close()
            return st
        from simian_b200.distributed import run_distributed_model
        if args.pull == "deferred":
            opts["deferred_pull"] = 1
        st = run_distributed_model(lambda e: build_model(e, c), "bench", end, min_delay, rank,
                                   world, local_rank, profile=profile, **opts)
        st.pop("_engine").close()
        return st

    def measure(c, steps, warmup):
        """`steps` timed steps of workload c: device time (max over ranks), events,
        launches, windows, and the digest / event count of the last step, all ranks"""
        for _ in range(warmup):
            one_step_resident(c)
        barrier()
        dev_s, wall0, events, launches, windows = 0.0, time.perf_counter(), 0, 0, 0
        st = None
        for _ in range(steps):
            st = one_step_resident(c)
            dev_s += st["device_seconds"]
            events += st["total_events"]
            launches += st["kernel_launches"]
            windows += st["windows"]
        barrier()
        wall = time.perf_counter() - wall0
        digest, last_events = st["digest"], st["total_events"]
        ties = st["distinguishable_ties"]
        if dist is not None:
            t = torch.tensor([dev_s, float(events), float(launches)], dtype=torch.float64,
                             device="cuda")
            tmax = t.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            dev_s, events, launches = float(tmax[0]), int(t[1]), int(t[2])
            # wrapping 64-bit sums over the ranks
            d = torch.tensor([digest - (1 << 64) if digest >= (1 << 63) else digest,
                              last_events, ties], dtype=torch.int64, device="cuda")
            dist.all_reduce(d, op=dist.ReduceOp.SUM)
            digest, last_events, ties = int(d[0]) & 0xFFFFFFFFFFFFFFFF, int(d[1]), int(d[2])
        return {"dev_s": dev_s, "wall_s": wall, "events": events, "launches": launches,
                "windows": windows // 1, "steps": steps, "digest": digest,
                "last_events": last_events, "dist_ties": ties}

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        one_step_resident(cfg)
    barrier()
    sampler.mark()
    m = measure(cfg, args.steps, 0)
    clocks = sampler.stop() if rank == 0 else None
    dev_s, events, launches, windows = m["dev_s"], m["events"], m["launches"], m["windows"]
    value = events / dev_s

    # ---- N > 1: the same model on ONE GPU (rank 0), for the parity of the sharded run
    single = None
    nvlink = None
    if world > 1:
        # NVLink roofline of the exchange: one profiled step (CUDA events around
        # every kernel of every window)
        stp = one_step_resident(cfg, profile=True)
        ph = stp.get("phase_ms", {})
        t = torch.tensor([ph.get("k_pull", 0.0), ph.get("k_window", 0.0),
                          ph.get("k_advance", 0.0), float(stp["emitted"])],
                         dtype=torch.float64, device="cuda")
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        pull_ms = float(tmax[0])
        # records that cross GPUs: p_remote of what was emitted, 32 bytes each,
        # read by the receiver out of the sender's HBM
        cross_bytes_per_gpu = cfg["p_remote"] * float(t[3]) / world * 32.0
        nvlink = {"bound": "nvlink", "achieved": cross_bytes_per_gpu / (pull_ms * 1e-3) / 1e9
                  if pull_ms > 0 else None, "peak": NVLINK_PEAK_GBS, "unit": "GB/s",
                  "frac": cross_bytes_per_gpu / (pull_ms * 1e-3) / 1e9 / NVLINK_PEAK_GBS
                  if pull_ms > 0 else None, "kernel": "k_pull",
                  "k_pull_ms_per_step": pull_ms, "k_window_ms_per_step": float(tmax[1]),
                  "k_advance_ms_per_step": float(tmax[2]),
                  "bytes_per_gpu_per_step": cross_bytes_per_gpu,
                  "note": "bytes each GPU reads from its peers' outboxes per step (p_remote x "
                          "emitted records x 32 B / GPUs) over the summed k_pull time of one "
                          "profiled step (max over ranks; includes waiting for the slowest "
                          "peer's window); peak = measured peer copy, 770 GB/s per direction"}
        barrier()
        if rank == 0 and not args.no_extra:
            try:
                capi.lib().simian_trim_cache()
                e = capi.Engine("bench1", 0.0, cfg["end"], cfg["min_delay"], seed=1,
                                device=local_rank,
                                pool_events=10 * cfg["n"] * cfg["per_entity"])
                build_model(e, cfg)
                s1 = e.run()
                e.close()
                capi.lib().simian_trim_cache()
                single = {"digest": "%016x" % s1["digest"], "total_events": s1["total_events"],
                          "windows": s1["windows"],
                          "ms": 1e3 * s1["device_seconds"],
                          "matches_sharded_run": bool(s1["digest"] == m["digest"] and
                                                      s1["total_events"] == m["last_events"])}
            except Exception as ex:  # noqa: BLE001 -- e.g. the model does not fit one GPU
                single = {"skipped": str(ex)[:200]}
        barrier()

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
                            rank=rank, size=world, device=local_rank,
                            deferred_pull=int(world > 1 and args.pull == "deferred"))
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
        e2e_digest, e2e_events = st2["digest"], st2["total_events"]
        n_local = len(ev) // cfg["per_entity"]
        if dist is not None:
            t = torch.tensor([t_e2e, float(ev_e2e)], dtype=torch.float64, device="cuda")
            tmax = t.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            t_e2e, ev_e2e = float(tmax[0]), int(t[1])
            d = torch.tensor([e2e_digest - (1 << 64) if e2e_digest >= (1 << 63) else e2e_digest,
                              e2e_events], dtype=torch.int64, device="cuda")
            dist.all_reduce(d, op=dist.ReduceOp.SUM)
            e2e_digest, e2e_events = int(d[0]) & 0xFFFFFFFFFFFFFFFF, int(d[1])
        e2e = {"value": ev_e2e / t_e2e, "unit": "events/s",
               "h2d_bytes_per_step": int(ev.nbytes + 48) * world,
               "d2h_bytes_per_step": int(16 * n_local + 160) * world,
               "steps": ksteps, "ms_per_step": 1e3 * t_e2e / ksteps,
               "digest_matches_resident_run": bool(e2e_digest == m["digest"] and
                                                   e2e_events == m["last_events"])}
        del buf, ev, res

    # ---- the other BASELINE configs -------------------------------------------
    extras = {}
    if not args.no_extra:
        for key, c in extra_workloads(world, args.scale):
            if dist is not None and key == "scale_anchor":
                continue
            try:
                capi.lib().simian_trim_cache()
                steps = 1 if is_lanl(c) or key == "config1" else 3
                mm = measure(c, steps, 1)
                v = mm["events"] / mm["dev_s"]
                bpe = bytes_per_event(c)
                ent = c["n_ent"] if is_lanl(c) else c["n"]
                x = {"workload": c["label"], "value": v, "unit": "events/s", "n_gpus": world,
                     "entities": ent,
                     "ms_per_step": 1e3 * mm["dev_s"] / steps,
                     "events_per_step": mm["events"] // steps,
                     "windows_per_step": mm["windows"] // steps,
                     "windows_per_s": mm["windows"] / mm["dev_s"],
                     "gpu_launches": mm["launches"],
                     "roofline": {"bound": "hbm", "achieved": v / world * bpe / 1e9, "peak": peak,
                                  "unit": "GB/s", "frac": v / world * bpe / 1e9 / peak,
                                  "bytes_per_event": bpe, "peak_kind": peak_kind},
                     "parity": {"digest": "%016x" % mm["digest"],
                                "total_events": mm["last_events"],
                                "distinguishable_ties": mm["dist_ties"]}}
                if rank == 0 and not args.no_cpu_baseline:
                    cb = cpu_run(c)
                    for k in ("events", "seconds", "digest", "windows"):
                        cb.pop(k)
                    x["cpu_baseline"] = cb
                extras[key] = x
            except Exception as ex:  # noqa: BLE001
                extras[key] = {"workload": c["label"], "error": str(ex)[:300]}
            barrier()

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    per_gpu = value / world
    achieved = per_gpu * BYTES_PER_EVENT / 1e9
    traffic, traffic_note = ncu_traffic() if world == 1 else (None, "single-GPU capture only")
    line = {
        "metric": "committed_events_per_sec", "value": value, "unit": "events/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dev_s / args.steps,
        "wall_ms_per_step": 1e3 * m["wall_s"] / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": cfg["label"], "entities": cfg["n"],
                   "events_per_entity": cfg["per_entity"], "min_delay": cfg["min_delay"],
                   "end_time": cfg["end"], "events_per_step": events // args.steps,
                   "windows_per_step": windows // args.steps,
                   "l2_policy": "inputs_larger_than_l2 (event pool >= 512 MB per GPU)",
                   "exchange": None if world == 1 else (
                       "device-driven pull over NVLink peer memory, far records pulled while "
                       "the next window runs (two-ended outboxes)" if args.pull == "deferred"
                       else "device-driven pull over NVLink peer memory between two windows"),
                   "parity": {"digest": "%016x" % m["digest"], "total_events": m["last_events"],
                              "distinguishable_ties": m["dist_ties"],
                              "single_gpu_run_of_the_same_model": single,
                              "checked": "the reference arm (bench.py --impl reference) prints the "
                                         "oracle's digest and total_events of the same configuration; "
                                         "tests/test_full_size_gpu.py compares every entity of config 2 "
                                         "with the oracle; the oracle is pinned against the unmodified "
                                         "SimianPie engine and the reference's fixtures -- a SimianJS "
                                         "binary cannot run in this image (no JS engine)"}},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic,
                     "algorithmic_bytes_per_launch":
                         BYTES_PER_EVENT * (events // max(windows, 1) // world),
                     "kernel": "k_window", "peak_kind": peak_kind,
                     "bytes_per_event": BYTES_PER_EVENT,
                     "note": "achieved = committed events/s per GPU x 96 B over the whole "
                             "timed window loop (k_window is >99% of it); traffic = "
                             + traffic_note},
    }
    if nvlink is not None:
        line["roofline_nvlink"] = nvlink
    if e2e is not None:
        line["e2e"] = e2e
    if extras:
        line["extra_configs"] = extras
        if "scale_anchor" in extras and "value" in extras["scale_anchor"]:
            line["scale_anchor"] = {k: extras["scale_anchor"][k] for k in
                                    ("workload", "value", "unit", "ms_per_step", "entities")}
    if not args.no_cpu_baseline and world == 1:
        cb = cpu_run(cfg, full=True)
        cb["digest_matches_native"] = bool(cb["digest"] == "%016x" % m["digest"] and
                                           cb["events"] == m["last_events"])
        for k in ("events", "seconds", "windows"):
            cb.pop(k)
        line["cpu_baseline"] = cb
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
