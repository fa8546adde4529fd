"""-m gpu: the CUDA engine, called through the C-ABI, against the CPU oracle
and the committed SimianPie goldens.  Bit-exact on everything."""
import json
import os

import numpy as np
import pytest

from helpers import assert_parity, run_engine_phold, run_oracle_phold

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "phold_simianpie.json")


def golden_cases():
    with open(GOLD) as f:
        return json.load(f)["cases"]


@pytest.mark.parametrize("case", golden_cases(), ids=lambda c: c["name"])
def test_engine_matches_simianpie_goldens(case):
    opts = {}
    if case["min_delay"] < 1e-3:
        opts["bucket_width"] = 1.0 / 64
    st, cnt, h = run_engine_phold(case, **opts)
    assert st["total_events"] == case["total_events"]
    assert list(cnt) == case["counts"]
    assert ["%016x" % x for x in h] == case["hashes"]
    assert "%016x" % st["digest"] == case["digest"]
    assert float(st["last_time"]).hex() == case["last_time_hex"]


CASES = [
    dict(name="one_block", n=1000, per_entity=1, min_delay=0.05, lookahead=0.05,
         p_remote=0.0, mode=0, end=30.0, seed=1),
    dict(name="multi_block_ross", n=5000, per_entity=16, min_delay=0.1,
         lookahead=0.1, p_remote=0.1, mode=1, end=6.0, seed=2),
    dict(name="ragged_last_block", n=1025, per_entity=3, min_delay=0.1,
         lookahead=0.1, p_remote=0.3, mode=1, end=8.0, seed=3),
    dict(name="other_group", n=4096, per_entity=8, min_delay=0.1,
         lookahead=0.1, p_remote=0.5, mode=2, groups=8, end=5.0, seed=4),
    dict(name="small_lookahead", n=3000, per_entity=16, min_delay=0.01,
         lookahead=0.01, p_remote=0.1, mode=1, end=1.5, seed=5),
    dict(name="burst_64_per_entity", n=2048, per_entity=64, min_delay=0.1,
         lookahead=0.1, p_remote=0.1, mode=1, end=1.0, seed=6),
    dict(name="single_entity", n=1, per_entity=5, min_delay=0.1, lookahead=0.1,
         p_remote=0.0, mode=0, end=50.0, seed=7),
    dict(name="lambda_half", n=777, per_entity=4, min_delay=0.2, lookahead=0.25,
         p_remote=0.2, lam=0.5, mode=1, end=20.0, seed=8),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: c["name"])
@pytest.mark.parametrize("sequential", [0, 1], ids=["per_event", "per_entity"])
@pytest.mark.parametrize("loop", [1, 0], ids=["persistent_loop", "launch_per_window"])
def test_engine_matches_oracle(case, sequential, loop, oracle_mod):
    """both handler mappings: one thread per event (pure handlers) and one
    thread per entity (the general path); both window drivers: the loop inside
    one cooperative launch (k_window_loop, the default for models this small)
    and one launch of k_window per window"""
    so, co, ho = run_oracle_phold(oracle_mod, case)
    se, ce, he = run_engine_phold(case, force_sequential=sequential, persistent_loop=loop)
    assert_parity(so, co, ho, se, ce, he)
    if loop:  # the whole run in a handful of launches
        assert se["kernel_launches"] <= 4


BURST_CASES = [
    # 16 events per entity at t = 0, blocks of 200 entities: bins of 16 entities,
    # parts of several bins, a ragged last block
    (dict(name="burst_blocks_of_200", n=3001, per_entity=16, min_delay=0.1, lookahead=0.1,
          p_remote=0.2, mode=1, end=2.0, seed=21), dict(block_entities=200)),
    # a burst in a block of one bin per entity (block_entities <= 16)
    (dict(name="burst_blocks_of_13", n=200, per_entity=100, min_delay=0.1, lookahead=0.1,
          p_remote=0.5, mode=1, end=1.0, seed=22), dict(block_entities=13)),
    # one part = one histogram bin that overflows again: the partition is followed
    # by entity-range splits over the part's own pages
    (dict(name="burst_part_overflows", n=1024, per_entity=40, min_delay=0.1, lookahead=0.1,
          p_remote=0.1, mode=1, end=1.0, seed=23), dict()),
    # a mid-run burst: wide windows (minDelay = 2 mean delays) keep every window
    # above the slot count
    (dict(name="burst_every_window", n=2048, per_entity=8, min_delay=2.0, lookahead=2.0,
          p_remote=0.3, mode=1, end=12.0, seed=24), dict(bucket_width=1.0)),
]


@pytest.mark.parametrize("case,opts", BURST_CASES, ids=lambda c: c.get("name", "") if "n" in c else None)
@pytest.mark.parametrize("sequential", [0, 1], ids=["per_event", "per_entity"])
def test_burst_partition_matches_oracle(case, opts, sequential, oracle_mod):
    """windows whose due events overflow the CTA's slots: the due records are
    copied into scratch pages grouped by entity sub-range (partition_burst)"""
    so, co, ho = run_oracle_phold(oracle_mod, case)
    se, ce, he = run_engine_phold(case, force_sequential=sequential, **opts)
    assert_parity(so, co, ho, se, ce, he)


@pytest.mark.parametrize("sequential", [0, 1], ids=["per_event", "per_entity"])
@pytest.mark.parametrize("epb", [1024, 777, 64])
def test_entity_block_size_is_a_run_time_option(epb, sequential, oracle_mod):
    """block_entities up to 1024 (sparse windows: more due events per CTA): the
    per-block entity states are a run-time sized tail of the CTA's shared memory"""
    case = dict(name="blocks", n=5000, per_entity=3, min_delay=0.02, lookahead=0.02,
                p_remote=0.3, mode=1, end=1.5, seed=31)
    so, co, ho = run_oracle_phold(oracle_mod, case)
    se, ce, he = run_engine_phold(case, force_sequential=sequential, block_entities=epb,
                                  bucket_width=0.01)
    assert_parity(so, co, ho, se, ce, he)


def test_sparse_windows_and_far_list(oracle_mod):
    """config-1 shape: minDelay << mean delay, so almost every event goes
    through the far-future list and the window jumps over idle time."""
    case = dict(name="sparse", n=200, per_entity=1, min_delay=1e-3,
                lookahead=1e-3, p_remote=0.0, mode=0, end=4.0, seed=9)
    so, co, ho = run_oracle_phold(oracle_mod, case)
    # default bucket width (minDelay/4) -> horizon 0.5: far list exercised
    for loop in (1, 0):
        se, ce, he = run_engine_phold(case, ring_buckets=256, persistent_loop=loop)
        assert_parity(so, co, ho, se, ce, he)
        assert se["far_appends"] > 0 and se["redistributions"] > 0
    # tuned bucket width
    se, ce, he = run_engine_phold(case, bucket_width=1.0 / 32)
    assert_parity(so, co, ho, se, ce, he)


def test_empty_run_and_end_time_zero_events(oracle_mod):
    """no scheduled events: one window at startTime, then done (simian.js:118,147)"""
    from simian_b200 import capi
    e = capi.Engine("empty", 0.0, 5.0, 0.5)
    e.define_class("Node", 0)
    e.add_entities("Node", 0, 10)
    e.attach_service("Node", "generate", "phold_generate")
    st = e.run()
    assert st["total_events"] == 0 and st["windows"] == 1
    o = oracle_mod.Oracle("empty", 0.0, 5.0, 0.5)
    o.define_class("Node", 0)
    o.add_entities("Node", 0, 10)
    assert o.run()["windows"] == 1


def test_bad_model_parameters_are_errors_not_memory_faults():
    """a parameter blob shorter than the attached device function reads, an
    out-of-range destination the handler computes, PHOLD groups < 2: the
    reference throws inside the handler; here they are engine errors"""
    from simian_b200 import capi
    from simian_b200.workloads import phold_params_blob
    e = capi.Engine("short", 0.0, 1.0, 0.1)
    e.define_class("Node", 0)
    e.add_entities("Node", 0, 64)
    e.attach_service("Node", "generate", "phold_generate")
    e.set_class_params("Node", phold_params_blob(0, 64, 0.1, 1.0, 0.1, 1, 1)[:16])
    e.sched_service_bulk(0.0, "generate", 0, "Node", 0, 64, 1)
    with pytest.raises(capi.SimianError) as ei:
        e.run()
    assert ei.value.code == 6
    # destinations beyond the entities that exist: n_entities overstated
    e = capi.Engine("range", 0.0, 1.0, 0.1)
    e.define_class("Node", 0)
    e.add_entities("Node", 0, 64)
    e.attach_service("Node", "generate", "phold_generate")
    e.set_class_params("Node", phold_params_blob(0, 1 << 20, 0.1, 1.0, 0.1, 0, 1))
    e.sched_service_bulk(0.0, "generate", 0, "Node", 0, 64, 1)
    with pytest.raises(capi.SimianError) as ei:
        e.run()
    assert ei.value.code == 6
    # OTHER_GROUP with one group: a modulo by zero in the handler
    e = capi.Engine("groups", 0.0, 1.0, 0.1)
    e.define_class("Node", 0)
    e.add_entities("Node", 0, 64)
    e.attach_service("Node", "generate", "phold_generate")
    e.set_class_params("Node", phold_params_blob(0, 64, 0.5, 1.0, 0.1, 2, 1))
    e.sched_service_bulk(0.0, "generate", 0, "Node", 0, 64, 1)
    with pytest.raises(capi.SimianError) as ei:
        e.run()
    assert ei.value.code == 6


def test_lookahead_violation_is_an_error():
    """entity.js:41-45: rx given and offset < minDelay must throw"""
    from simian_b200 import capi
    from simian_b200.workloads import build_phold
    e = capi.Engine("bad", 0.0, 5.0, 0.5)
    build_phold(e, 64, 1, lookahead=0.0, lam=1000.0)  # offsets << minDelay
    with pytest.raises(capi.SimianError) as ei:
        e.run()
    assert ei.value.code == 1


# ---- LANL PDES benchmark (stateful handlers, constructors that send) ----------

LANL_GOLD = os.path.join(os.path.dirname(__file__), "golden", "lanl_simianpie.json")


def lanl_golden_cases():
    with open(LANL_GOLD) as f:
        return json.load(f)["cases"]


@pytest.mark.parametrize("split_sums", [0, 1], ids=["sums_in_window", "k_sums"])
@pytest.mark.parametrize("case", lanl_golden_cases(), ids=lambda c: c["name"])
def test_engine_lanl_matches_simianpie_goldens(case, split_sums):
    from helpers import assert_lanl_golden, run_engine_lanl
    out = run_engine_lanl(case, split_sums=split_sums)
    assert out[0]["error"] == 0
    assert_lanl_golden(case, *out)


LANL_CASES = [
    dict(name="lanl_multi_block", n_ent=3000, s_ent=10.0, q_avg=1.0,
         p_receive=0.0, m_ent=32.0, ops_ent=48.0, ops_sigma=0.1,
         cache_friendliness=0.5, end_time=10.0, min_delay=1.0, time_bins=10,
         seed=2),
    dict(name="lanl_hot_spots", n_ent=2500, s_ent=12.0, q_avg=2.0,
         p_receive=0.002, m_ent=24.0, ops_ent=30.0, ops_sigma=0.2,
         cache_friendliness=0.5, end_time=8.0, min_delay=0.5, time_bins=16,
         seed=4),
    dict(name="lanl_all_to_entity0", n_ent=300, s_ent=6.0, q_avg=1.0,
         p_receive=1.0, m_ent=8.0, ops_ent=10.0, ops_sigma=0.0,
         cache_friendliness=1.0, end_time=6.0, min_delay=1.0, time_bins=4,
         seed=6),
    dict(name="lanl_deep_timer_queue", n_ent=1100, s_ent=40.0, q_avg=6.0,
         p_receive=0.0, m_ent=16.0, ops_ent=16.0, ops_sigma=0.5,
         cache_friendliness=0.75, end_time=4.0, min_delay=0.25, time_bins=8,
         seed=8),
    # geometric source and list-size distributions (pdes_lanl_benchmarkV8.js:209-221),
    # several blocks; every entity keeps a list and ops >= 1
    dict(name="lanl_geometric_mild", n_ent=1400, s_ent=10.0, q_avg=1.0,
         p_receive=0.001, p_send=0.002, p_list=0.003, invert=False, m_ent=40.0,
         ops_ent=100.0, ops_sigma=0.2, cache_friendliness=0.5, end_time=6.0,
         min_delay=0.5, time_bins=8, seed=10),
    # a steep list-size distribution: the entities deep in the tail own no list and
    # have ops == 0 (no list work on a receive); inverted source distribution
    dict(name="lanl_geometric_steep_tail", n_ent=700, s_ent=12.0, q_avg=1.0,
         p_receive=0.0, p_send=0.004, p_list=0.05, invert=True, m_ent=30.0,
         ops_ent=60.0, ops_sigma=0.1, cache_friendliness=0.5, end_time=5.0,
         min_delay=0.5, time_bins=10, seed=12),
]


# the deferred weighted sums: computed by the window kernel itself, by k_sums (option
# split_sums: one global job array, a persistent grid of warps), and by both when the
# job array is too small for some passes (sum_jobs = 96: those passes fall back)
SUM_MODES = [dict(split_sums=0), dict(split_sums=1), dict(split_sums=1, sum_jobs=96)]


@pytest.mark.parametrize("sums", SUM_MODES, ids=["sums_in_window", "k_sums", "k_sums_overflowing"])
@pytest.mark.parametrize("case", LANL_CASES, ids=lambda c: c["name"])
def test_engine_lanl_matches_oracle(case, sums, oracle_mod):
    """bit-exact counts, trace hashes (engine tie order = oracle tie_mode 1),
    the whole per-entity state incl. the f64 lists, window and emit counts"""
    from helpers import run_engine_lanl, run_oracle_lanl
    so, co, ho, sto, lo = run_oracle_lanl(oracle_mod, case, tie_mode=1)
    se, ce, he, ste, le = run_engine_lanl(case, **sums)
    assert se["error"] == 0
    assert se["total_events"] == so["total_events"]
    np.testing.assert_array_equal(ce, co)
    np.testing.assert_array_equal(he, ho)
    assert ste.tobytes() == sto.tobytes()
    assert le.tobytes() == lo.tobytes()
    assert se["digest"] == so["digest"]
    assert se["state_checksum"] == so["state_checksum"]
    assert float(se["last_time"]).hex() == float(so["last_time"]).hex()
    assert se["windows"] == so["windows"]
    assert se["emitted"] == so["emitted"]
    assert se["dropped_past_end"] == so["dropped_past_end"]
    assert se["ties"] == so["ties"]
    assert se["distinguishable_ties"] == so["distinguishable_ties"]


def test_identical_records_are_all_committed(oracle_mod):
    """two schedService calls that produce byte-identical records (same time,
    service, data, seq): the reference executes both (eventQ holds both); the
    engine must not collapse them, on either handler mapping"""
    from simian_b200 import capi
    from simian_b200.workloads import phold_params_blob
    n = 300
    ev = np.zeros(2 * n, dtype=capi.EVENT_DTYPE)
    ev["time"] = 0.25
    ev["dst"] = np.tile(np.arange(n, dtype=np.uint32), 2)
    ev["src"] = capi.NULL_ENTITY
    ev["handler"] = 0
    ev["flags"] = 1
    ev["seq"] = 7
    outs = []
    for make in ("oracle", 0, 1):
        if make == "oracle":
            h = oracle_mod.Oracle("dup", 0.0, 3.0, 0.1, seed=5)
        else:
            h = capi.Engine("dup", 0.0, 3.0, 0.1, seed=5, force_sequential=make)
        h.define_class("Node", 0)
        h.add_entities("Node", 0, n)
        h.attach_service("Node", "generate", "phold_generate")
        h.set_class_params("Node", phold_params_blob(0, n, 0.2, 1.0, 0.1, 1, 1))
        h.sched_events(ev)
        st = h.run()
        outs.append((st, h.read_counts("Node", n), h.read_trace_hashes("Node", n)))
        h.close()
    so, co, ho = outs[0]
    assert so["total_events"] > 2 * n
    for se, ce, he in outs[1:]:
        assert se["total_events"] == so["total_events"]
        np.testing.assert_array_equal(ce, co)
        np.testing.assert_array_equal(he, ho)
        assert se["digest"] == so["digest"]


# ---- hello.js: two entity classes, payloads, replies to tx/txId ------------------

HELLO_GOLD = os.path.join(os.path.dirname(__file__), "golden", "hello_simianpie.json")


def hello_golden_cases():
    with open(HELLO_GOLD) as f:
        return json.load(f)["cases"]


HELLO_CASES = hello_golden_cases() + [
    dict(name="hello_two_blocks", count=700, end=260.0, min_delay=0.5, seed=3),
    dict(name="hello_sparse_default_buckets", count=40, end=120.0, min_delay=0.01, seed=4,
         default_buckets=True),
]


@pytest.mark.parametrize("case", HELLO_CASES, ids=lambda c: c["name"])
@pytest.mark.parametrize("sequential", [0, 1], ids=["per_event", "per_entity"])
def test_engine_hello_matches_oracle(case, sequential, oracle_mod):
    """engine order of equal-time events = oracle tie_mode 1: bit-exact counts,
    trace hashes, digest, tie statistics; entities of both classes share blocks"""
    from helpers import run_hello
    from simian_b200 import capi
    opts = {} if case.get("default_buckets") else {"bucket_width": 1.0}
    so, co, ho = run_hello(
        lambda: oracle_mod.Oracle(case["name"], 0.0, case["end"], case["min_delay"],
                                  tie_mode=1, seed=case["seed"]), case)
    se, ce, he = run_hello(
        lambda: capi.Engine(case["name"], 0.0, case["end"], case["min_delay"],
                            seed=case["seed"], force_sequential=sequential, **opts), case)
    assert se["error"] == 0
    assert se["total_events"] == so["total_events"]
    np.testing.assert_array_equal(ce, co)
    np.testing.assert_array_equal(he, ho)
    assert se["digest"] == so["digest"]
    assert float(se["last_time"]).hex() == float(so["last_time"]).hex()
    assert se["windows"] == so["windows"]
    assert se["emitted"] == so["emitted"]
    assert se["ties"] == so["ties"] > 0
    assert se["distinguishable_ties"] == so["distinguishable_ties"] > 0
    if "counts" in case:  # the unmodified SimianPie engine
        assert list(ce) == case["counts"]
        assert se["total_events"] == case["total_events"]


@pytest.mark.timeout(120)
def test_contended_cells_many_page_installs(oracle_mod):
    """integer delays: thousands of events per window land in the same few
    calendar cells, pages fill and are installed while many claims are in
    flight (the append phase holds up to four claims per thread).  Regression
    for a deadlock between crossed install/wait claims; repeated because it was
    timing dependent."""
    from helpers import run_hello
    from simian_b200 import capi
    case = dict(name="hello_contended", count=3000, end=180.0, min_delay=0.5, seed=11)
    so, co, ho = run_hello(
        lambda: oracle_mod.Oracle(case["name"], 0.0, case["end"], case["min_delay"],
                                  tie_mode=1, seed=case["seed"]), case)
    for rep in range(4):
        se, ce, he = run_hello(
            lambda: capi.Engine(case["name"], 0.0, case["end"], case["min_delay"],
                                seed=case["seed"], bucket_width=1.0), case)
        assert se["error"] == 0
        assert se["total_events"] == so["total_events"]
        np.testing.assert_array_equal(ce, co)
        np.testing.assert_array_equal(he, ho)


def test_read_back_into_pinned_and_pageable_buffers(oracle_mod):
    """results go straight into a caller's page-locked buffer (one DMA) or
    through the bounce buffer: same values"""
    import torch
    from simian_b200 import capi
    from simian_b200.workloads import build_phold
    case = CASES[1]
    e = capi.Engine("rb", 0.0, case["end"], case["min_delay"], seed=case["seed"])
    build_phold(e, case["n"], case["per_entity"], case["lookahead"], case["p_remote"], 1.0,
                case["mode"], 1)
    e.run()
    n = case["n"]
    pinned = torch.empty((2, n + 5), dtype=torch.int64, pin_memory=True).numpy().view(np.uint64)
    c0, first, stride = e.read_local("Node", 0)
    c1, _, _ = e.read_local("Node", 0, out=pinned[0])
    h0, _, _ = e.read_local("Node", 1)
    h1, _, _ = e.read_local("Node", 1, out=pinned[1])
    assert (first, stride, len(c1)) == (0, 1, n)
    np.testing.assert_array_equal(c0, c1)
    np.testing.assert_array_equal(h0, h1)
    np.testing.assert_array_equal(c0, e.read_counts("Node", n))
    with pytest.raises(ValueError):
        e.read_local("Node", 0, out=np.zeros(3, dtype=np.uint64))
    e.close()


SPLIT_CASES = [(c, {}) for c in CASES if c["name"] in
               ("multi_block_ross", "ragged_last_block", "other_group", "small_lookahead",
                "burst_64_per_entity", "lambda_half")] + BURST_CASES + [
    # config-1 shape: far list, redistribution passes (left to k_window_rest)
    (dict(name="sparse_far_list", n=200, per_entity=1, min_delay=1e-3, lookahead=1e-3,
          p_remote=0.0, mode=0, end=4.0, seed=9), dict(ring_buckets=256)),
]


@pytest.mark.parametrize("case,opts", SPLIT_CASES, ids=lambda c: c.get("name", "") if "n" in c else None)
def test_split_window_matches_oracle(case, opts, oracle_mod):
    """option split_handlers: a window is k_window_stage (dequeue, rank, park the
    events with their commit indices, trace chain, recycle) -> k_window_rest
    (bursts / redistribution passes, the one-kernel form) -> k_emit (handlers +
    appends, persistent warps).  Same bits as the one-kernel window."""
    so, co, ho = run_oracle_phold(oracle_mod, case)
    se, ce, he = run_engine_phold(case, split_handlers=1, persistent_loop=0, **opts)
    assert_parity(so, co, ho, se, ce, he)
    assert se["kernel_launches"] >= 3 * se["windows"]


# ---- payloads beyond 8 bytes (entity.js:52-60: `data` is any object) ----------------

BLOB_GOLD = os.path.join(os.path.dirname(__file__), "golden", "blob_simianpie.json")


def blob_golden_cases():
    with open(BLOB_GOLD) as f:
        return json.load(f)["cases"]


@pytest.mark.parametrize("case", blob_golden_cases(), ids=lambda c: c["name"])
def test_engine_payloads_match_simianpie_goldens(case):
    """events whose data is a list of words (SimianPie: a Python list): header +
    continuation records through the calendar; the acknowledged fold of every payload
    is covered by the senders' trace hashes"""
    from helpers import assert_blob_golden, run_blobs
    from simian_b200 import capi
    e = capi.Engine(case["name"], 0.0, case["end"], case["min_delay"], seed=case["seed"],
                    payloads=1)
    assert_blob_golden(case, *run_blobs(e, case))


BLOB_CASES = [
    (dict(name="blobs_multi_block", n=3000, end=5.0, min_delay=0.25, lookahead=0.25,
          max_words=12, seed=21), dict()),
    # wide windows and small entity blocks: the due records (events + payload words)
    # overflow the CTA's slots in every window -- filtered scans, burst partition
    (dict(name="blobs_overflowing_windows", n=1200, end=9.0, min_delay=1.5, lookahead=1.5,
          max_words=24, seed=22), dict(bucket_width=0.5)),
    (dict(name="blobs_small_blocks", n=900, end=6.0, min_delay=0.5, lookahead=0.5,
          max_words=7, seed=23), dict(block_entities=48)),
]


@pytest.mark.parametrize("case,opts", BLOB_CASES, ids=lambda c: c.get("name", "") if "n" in c else None)
def test_engine_payloads_match_oracle(case, opts, oracle_mod):
    from helpers import assert_parity, run_blobs
    from simian_b200 import capi
    so, co, ho = run_blobs(oracle_mod.Oracle(case["name"], 0.0, case["end"], case["min_delay"],
                                             seed=case["seed"]), case)
    se, ce, he = run_blobs(capi.Engine(case["name"], 0.0, case["end"], case["min_delay"],
                                       seed=case["seed"], payloads=1, **opts), case)
    assert_parity(so, co, ho, se, ce, he)


def test_payloads_need_the_option(oracle_mod):
    """without option payloads = 1 a handler that sends a payload is a loud error"""
    from helpers import run_blobs
    from simian_b200 import capi
    case = dict(name="b", n=64, end=2.0, min_delay=0.25, lookahead=0.25, max_words=4, seed=1)
    e = capi.Engine("b", 0.0, case["end"], case["min_delay"], seed=1)
    with pytest.raises(capi.SimianError) as ei:
        run_blobs(e, case)
    assert ei.value.code == 16
    e.close()


@pytest.mark.parametrize("case,opts", SPLIT_CASES, ids=lambda c: c.get("name", "") if "n" in c else None)
def test_lbts_step_as_its_own_launch_matches_oracle(case, opts, oracle_mod):
    """option fused_advance = 0 (the default from 592 entity blocks on): the window
    kernel's CTAs end without the last-CTA pattern and a one-CTA launch (k_lbts) computes
    the next window -- bursts, far list and redistribution passes included"""
    so, co, ho = run_oracle_phold(oracle_mod, case)
    se, ce, he = run_engine_phold(case, fused_advance=0, persistent_loop=0, **opts)
    assert_parity(so, co, ho, se, ce, he)
    assert se["kernel_launches"] >= 2 * se["windows"]
