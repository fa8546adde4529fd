"""-m "not gpu": the CPU oracle against the reference's own fixtures, the
committed SimianPie goldens and the pure-Python transcription of the spec."""
import json
import math
import os
import random

import numpy as np
import pytest

from helpers import run_oracle_phold

GOLD = os.path.join(os.path.dirname(__file__), "golden", "phold_simianpie.json")


def golden_cases():
    with open(GOLD) as f:
        return json.load(f)["cases"]


# ---- eventQ.js ------------------------------------------------------------

def test_heap_fixture_from_simianlua(oracle_mod):
    """the only heap fixture the reference holds: SimianLua/eventQ.lua:62"""
    L = oracle_mod.lib()
    q = L.oracle_q_new(0)
    data = [7878, 78217, 90, 1, 425, 10]
    for i, t in enumerate(data):
        L.oracle_q_push(q, float(t), i)
    assert L.oracle_q_peek(q) == 1.0
    out = [L.oracle_q_pop(q, None) for _ in data]
    assert out == sorted(map(float, data))
    assert L.oracle_q_len(q) == 0
    L.oracle_q_free(q)


def test_heap_sortedness_like_test_Q_js(oracle_mod):
    """SimianJS/Tests/test.Q.js:14-21: push random times, pop all, never
    'Out of order'; plus interleaved push/pop."""
    L = oracle_mod.lib()
    rnd = random.Random(5)
    q = L.oracle_q_new(0)
    n = 200000
    for i in range(n):
        L.oracle_q_push(q, rnd.random(), i)
    prev = -1.0
    for _ in range(n // 2):
        t = L.oracle_q_pop(q, None)
        assert t >= prev
        prev = t
    for i in range(n // 4):
        L.oracle_q_push(q, prev + rnd.random(), i)
    while L.oracle_q_len(q):
        t = L.oracle_q_pop(q, None)
        assert t >= prev
        prev = t
    L.oracle_q_free(q)


def test_heap_equal_time_order_is_shape_dependent(oracle_mod):
    """eventQ.js compares .time only (:36,:62,:66): ten equal-time pushes pop
    as 0,9,8,...,1 -- not FIFO."""
    import ctypes
    L = oracle_mod.lib()
    q = L.oracle_q_new(0)
    for i in range(10):
        L.oracle_q_push(q, 1.0, i)
    tags = []
    for _ in range(10):
        tag = ctypes.c_uint64()
        L.oracle_q_pop(q, ctypes.byref(tag))
        tags.append(tag.value)
    assert tags == [0, 9, 8, 7, 6, 5, 4, 3, 2, 1]
    L.oracle_q_free(q)


# ---- hash.js / placement ----------------------------------------------------

def test_hash_js_values(oracle_mod):
    L = oracle_mod.lib()
    assert L.oracle_hash_name(b"Node") == 6385183179.0
    # reference arithmetic: doubles, string iterated backwards
    res = 5381.0
    for ch in reversed("PDES_LANL_Node"):
        res = 33 * res + ord(ch)
    assert L.oracle_hash_name(b"PDES_LANL_Node") == res
    assert res > 2.0 ** 53  # inexact: % size is taken on the rounded double


@pytest.mark.parametrize("size,base", [(1, 0), (2, 1), (4, 3), (8, 3)])
def test_base_rank_node(oracle_mod, size, base):
    o = oracle_mod.Oracle("x", 0.0, 1.0, 0.1, size=size)
    o.define_class("Node", 0)
    assert o.base_rank("Node") == base


# ---- spec header vs its Python transcription --------------------------------

def test_spec_bit_exact_vs_pyspec(oracle_mod):
    from oracle import pyspec
    L = oracle_mod.lib()
    rnd = random.Random(11)
    for _ in range(20000):
        x = rnd.getrandbits(64)
        assert L.oracle_spec_mix64(x) == pyspec.mix64(x)
        assert L.oracle_spec_u01(x) == pyspec.u01(x)
        u = pyspec.u01_open(x)
        assert L.oracle_spec_u01_open(x) == u
        assert 0.0 < u < 1.0
        assert L.oracle_spec_log(u) == pyspec.log(u)
        assert L.oracle_spec_exponential(x, 2.0) == pyspec.exponential(x, 2.0)
        e, c = rnd.getrandbits(32), rnd.getrandbits(40)
        k = pyspec.rng_key(x, e, c)
        assert L.oracle_spec_rng_key(x, e, c) == k
        assert L.oracle_spec_rng_draw(k, 3) == pyspec.rng_draw(k, 3)
        t = rnd.random() * 100
        assert L.oracle_spec_trace_step(x, t, 3, e, c) == pyspec.trace_step(x, t, 3, e, c)
        assert L.oracle_spec_digest_term(e, c, x) == pyspec.digest_term(e, c, x)


def test_integer_power_is_bit_exact_and_close_to_libm(oracle_mod):
    """simian_powi (the geometric distributions of the LANL benchmark,
    pdes_lanl_benchmarkV8.js:211-212,220): same bits on host and in pyspec, and
    within a few ulp-per-multiplication of math.pow"""
    from oracle import pyspec
    L = oracle_mod.lib()
    rnd = random.Random(5)
    assert L.oracle_spec_powi(0.37, 0) == 1.0
    for _ in range(5000):
        b, n = rnd.random(), rnd.randrange(0, 5000)
        v = L.oracle_spec_powi(b, n)
        assert v == pyspec.powi(b, n)
        ref = math.pow(b, n)
        assert abs(v - ref) <= 1e-12 * ref + 5e-324


def test_portable_log_is_accurate(oracle_mod):
    L = oracle_mod.lib()
    rnd = random.Random(3)
    worst = 0.0
    for _ in range(50000):
        x = rnd.random() * 10 ** rnd.uniform(-300, 300)
        got, ref = L.oracle_spec_log(x), math.log(x)
        ulp = math.ulp(ref) if ref else 2.0 ** -1074
        worst = max(worst, abs(got - ref) / ulp)
    assert worst <= 1.0
    assert L.oracle_spec_log(1.0) == 0.0
    assert L.oracle_spec_log(2.0 ** -53) == math.log(2.0 ** -53)


def test_u01_index_never_reaches_n(oracle_mod):
    L = oracle_mod.lib()
    top = (1 << 64) - 1
    for n in (1, 2, 3, 10, 1000, 1 << 20, (1 << 24) - 1, 1 << 24, 1 << 31):
        assert int(L.oracle_spec_u01(top) * float(n)) < n


# ---- oracle vs the unmodified reference engine (SimianPie) -------------------

@pytest.mark.parametrize("case", golden_cases(), ids=lambda c: c["name"])
@pytest.mark.parametrize("size", [1, 2, 3, 8])
@pytest.mark.parametrize("tie_mode", [0, 1])
def test_oracle_matches_simianpie_goldens(oracle_mod, case, size, tie_mode):
    st, cnt, h = run_oracle_phold(oracle_mod, case, size=size, tie_mode=tie_mode)
    assert st["total_events"] == case["total_events"]
    assert st["distinguishable_ties"] == 0 == case["distinguishable_ties"]
    assert st["ties"] == case["ties"]
    assert list(cnt) == case["counts"]
    assert ["%016x" % x for x in h] == case["hashes"]
    assert "%016x" % st["digest"] == case["digest"]
    assert float(st["last_time"]).hex() == case["last_time_hex"]
    assert st["synch_events"] == (0 if size == 1 else 2 * size * st["windows"])


# ---- engine semantics -----------------------------------

def _noop_world(oracle_mod, end, min_delay, times, size=1):
    o = oracle_mod.Oracle("sem", 0.0, end, min_delay, size=size)
    o.define_class("Node", 0)
    o.add_entities("Node", 0, 4)
    o.attach_service("Node", "tick", "noop")
    for i, t in enumerate(times):
        o.sched_service(t, "tick", i, "Node", i % 4)
    return o


def test_event_at_exactly_end_time_runs_and_later_is_dropped(oracle_mod):
    o = _noop_world(oracle_mod, 5.0, 1.0, [0.0, 5.0, 5.0000001, 7.0])
    st = o.run()
    assert st["total_events"] == 2  # simian.js:119,173
    assert st["last_time"] == 5.0


def test_windows_skip_idle_time(oracle_mod):
    """gmin jumps to the next event: windows are not on a fixed grid (:147)"""
    o = _noop_world(oracle_mod, 100.0, 1.0, [0.5, 0.9, 50.0, 50.5, 99.0])
    st = o.run()
    # windows: [0,1) {0.5,0.9}; [50,51) {50,50.5}; [99,100) {99}; then inf
    assert st["total_events"] == 5 and st["windows"] == 3


def test_first_window_starts_at_start_time_even_if_empty(oracle_mod):
    o = _noop_world(oracle_mod, 10.0, 1.0, [])
    st = o.run()
    assert st["total_events"] == 0 and st["windows"] == 1  # :118


def test_lookahead_violation_throws(oracle_mod):
    from simian_b200.workloads import build_phold
    o = oracle_mod.Oracle("bad", 0.0, 5.0, 0.5)
    build_phold(o, 64, 1, lookahead=0.0, lam=1000.0)
    with pytest.raises(oracle_mod.OracleError) as ei:
        o.run()
    assert "too little delay" in str(ei.value)  # entity.js:44


def test_add_entities_while_running_is_refused(oracle_mod):
    # simian.js:215 is unreachable from outside run(); the contiguity rule of
    # the bulk addEntities is what the binding enforces
    o = oracle_mod.Oracle("x", 0.0, 1.0, 0.1)
    o.define_class("Node", 0)
    o.add_entities("Node", 0, 4)
    with pytest.raises(oracle_mod.OracleError):
        o.add_entities("Node", 10, 4)


def test_results_do_not_depend_on_rank_count(oracle_mod):
    case = dict(n=600, per_entity=4, min_delay=0.1, lookahead=0.1, p_remote=0.3,
                mode=1, end=6.0, seed=21)
    ref = run_oracle_phold(oracle_mod, case, size=1)
    for size in (2, 4, 7):
        st, cnt, h = run_oracle_phold(oracle_mod, case, size=size)
        assert st["digest"] == ref[0]["digest"]
        assert st["windows"] == ref[0]["windows"]
        np.testing.assert_array_equal(cnt, ref[1])
        np.testing.assert_array_equal(h, ref[2])


def test_sched_events_equals_sched_service(oracle_mod):
    from simian_b200.workloads import build_phold, phold_params_blob
    case = dict(n=300, per_entity=2, min_delay=0.1, lookahead=0.1, p_remote=0.1,
                mode=1, end=4.0, seed=2)
    ref = run_oracle_phold(oracle_mod, case)
    o = oracle_mod.Oracle("ev", 0.0, case["end"], case["min_delay"], seed=2)
    o.define_class("Node", 0)
    o.add_entities("Node", 0, case["n"])
    o.attach_service("Node", "generate", "phold_generate")
    o.set_class_params("Node", phold_params_blob(0, case["n"], 0.1, 1.0, 0.1, 1, 1))
    ev = np.zeros(case["n"] * 2, dtype=oracle_mod.EVENT_DTYPE)
    ev["dst"] = np.tile(np.arange(case["n"], dtype=np.uint32), 2)
    ev["src"] = 0xFFFFFFFF
    ev["flags"] = 1
    ev["seq"] = np.arange(len(ev), dtype=np.uint32)
    o.sched_events(ev)
    st = o.run()
    assert st["digest"] == ref[0]["digest"]


# ---- LANL PDES benchmark (pdes_lanl_benchmarkV8.js) ---------------------------

LANL_GOLD = os.path.join(os.path.dirname(__file__), "golden", "lanl_simianpie.json")


def lanl_golden_cases():
    with open(LANL_GOLD) as f:
        return json.load(f)["cases"]


@pytest.mark.parametrize("case", lanl_golden_cases(), ids=lambda c: c["name"])
def test_oracle_lanl_matches_simianpie_goldens(oracle_mod, case):
    """constructors that send (pdes_lanl_benchmarkV8.js:255-256), same-window
    self-timers, the geometric hot-spot destination and the ReceiveHandler
    compute loop, against the unmodified SimianPie engine."""
    from helpers import assert_lanl_golden, run_oracle_lanl
    out = run_oracle_lanl(oracle_mod, case, tie_mode=1)
    assert_lanl_golden(case, *out)


@pytest.mark.parametrize("case", [c for c in lanl_golden_cases() if not c.get("stats_on_path")],
                         ids=lambda c: c["name"])
def test_oracle_lanl_literal_js_heap_same_counts_and_state(oracle_mod, case):
    """the t = minDelay requests of the constructors are distinguishable
    same-entity ties: under the literal JS heap order (tie_mode 0)
    only the order-sensitive trace hashes of the tied entities may differ --
    counts and the full entity state are order independent."""
    from helpers import run_oracle_lanl
    s1, c1, h1, st1, l1 = run_oracle_lanl(oracle_mod, case, tie_mode=1)
    s0, c0, h0, st0, l0 = run_oracle_lanl(oracle_mod, case, tie_mode=0)
    assert s0["total_events"] == s1["total_events"]
    np.testing.assert_array_equal(c0, c1)
    assert st0.tobytes() == st1.tobytes() and l0.tobytes() == l1.tobytes()
    assert s0["state_checksum"] == s1["state_checksum"]
    assert s0["distinguishable_ties"] == case["distinguishable_ties"] > 0
    assert (h0 != h1).sum() <= 2 * case["distinguishable_ties"]


def test_oracle_lanl_stats_fan_in_depends_on_the_tie_order(oracle_mod):
    """the statistics fan-in on the event path (pdes_lanl_benchmarkV8.js:311-342) is an
    n_ent-way time tie at entity 0 and its running mean (:322) is order dependent in
    f64: under the literal JS heap order the integer totals are the same, the mean
    may differ in its last bits -- the reference's own table depends on the heap
    shape.  The goldens pin the engine's (= SimianPie's FIFO) order."""
    from helpers import run_oracle_lanl
    case = [c for c in lanl_golden_cases() if c.get("stats_on_path")][0]
    s1, c1, h1, st1, l1 = run_oracle_lanl(oracle_mod, case, tie_mode=1)
    s0, c0, h0, st0, l0 = run_oracle_lanl(oracle_mod, case, tie_mode=0)
    for k in ("stats_received", "gsend_count", "greceive_count", "gops_count"):
        assert int(st0[k][0]) == int(st1[k][0]) == case["global_stats"][k]
    assert abs(float(st0["gops_mean"][0]) - float(st1["gops_mean"][0])) < 1e-9
    np.testing.assert_array_equal(c0, c1)


def test_oracle_lanl_geometric_tail_runs_and_is_rank_invariant(oracle_mod):
    """p_list steep enough that entities deep in the tail have no list and
    ops == 0: a receive there does no list work (include/simian_models.h); the
    result does not depend on the number of ranks"""
    from helpers import run_oracle_lanl
    case = dict(name="tail", n_ent=300, s_ent=12.0, q_avg=1.0, p_receive=0.0,
                p_send=0.01, p_list=0.06, invert=True, m_ent=30.0, ops_ent=60.0,
                ops_sigma=0.1, cache_friendliness=0.5, end_time=5.0, min_delay=0.5,
                time_bins=10, seed=12)
    ref = run_oracle_lanl(oracle_mod, case, size=1)
    st = ref[3]
    assert (st["list_size"] == 0).sum() > 100 and st["list_size"][0] == 540
    assert int(st["receive_count"][st["ops"] == 0].sum()) > 0   # the guarded path ran
    assert float(st["value_sink"][st["ops"] == 0].sum()) == 0.0
    out = run_oracle_lanl(oracle_mod, case, size=3)
    assert out[0]["digest"] == ref[0]["digest"]
    assert out[0]["state_checksum"] == ref[0]["state_checksum"]


def test_oracle_lanl_multi_rank_equals_single_rank(oracle_mod):
    from helpers import run_oracle_lanl
    case = lanl_golden_cases()[1]
    ref = run_oracle_lanl(oracle_mod, case, size=1)
    for size in (2, 3):
        out = run_oracle_lanl(oracle_mod, case, size=size)
        assert out[0]["total_events"] == ref[0]["total_events"]
        assert out[0]["digest"] == ref[0]["digest"]
        assert out[0]["state_checksum"] == ref[0]["state_checksum"]
        np.testing.assert_array_equal(out[2], ref[2])


# ---- ranks on host threads (bench.py --impl reference) -------------------------

@pytest.mark.parametrize("size,threads", [(4, 4), (5, 2), (8, 3)])
def test_oracle_threaded_ranks_equal_sequential(oracle_mod, size, threads):
    """oracle_run steps the ranks on host threads exactly like the sequential
    lock-step loop: same counts, hashes, digest, windows, statistics."""
    case = dict(name="thr", n=4000, per_entity=8, min_delay=0.1, lookahead=0.1,
                p_remote=0.5, mode=2, groups=size, end=4.0, seed=3)
    n = case["n"] // size * size
    out = []
    for thr in (1, threads):
        o = oracle_mod.Oracle("thr", 0.0, case["end"], case["min_delay"], size=size,
                              seed=case["seed"], threads=thr)
        from simian_b200.workloads import build_phold
        build_phold(o, n, case["per_entity"], case["lookahead"], case["p_remote"], 1.0,
                    case["mode"], size)
        st = o.run()
        out.append((st, o.read_counts("Node", n), o.read_trace_hashes("Node", n)))
    (s1, c1, h1), (s2, c2, h2) = out
    for k in ("total_events", "windows", "digest", "ties", "distinguishable_ties",
              "emitted", "dropped_past_end", "synch_events"):
        assert s1[k] == s2[k], k
    assert float(s1["last_time"]).hex() == float(s2["last_time"]).hex()
    np.testing.assert_array_equal(c1, c2)
    np.testing.assert_array_equal(h1, h2)


# ---- hello.js: two classes, payloads, replies, integer delays --------------------

HELLO_GOLD = os.path.join(os.path.dirname(__file__), "golden", "hello_simianpie.json")


def hello_golden_cases():
    with open(HELLO_GOLD) as f:
        return json.load(f)["cases"]


@pytest.mark.parametrize("case", hello_golden_cases(), ids=lambda c: c["name"])
def test_oracle_hello_matches_simianpie(oracle_mod, case):
    """multi-class placement, f64 payloads, replies to tx/txId against the
    unmodified SimianPie engine.  With SimianPie's FIFO order of equal times
    (tie_mode 2) everything matches, trace hashes of all entities included;
    with the literal JS heap (0) and the engine's order (1) the counts match
    and the hashes of the entities that saw no distinguishable tie."""
    from helpers import run_hello
    for tie_mode in (2, 0, 1):
        st, counts, hashes = run_hello(
            lambda: oracle_mod.Oracle(case["name"], 0.0, case["end"], case["min_delay"],
                                      tie_mode=tie_mode, seed=case["seed"]), case)
        assert st["total_events"] == case["total_events"]
        assert list(counts) == case["counts"]
        assert float(st["last_time"]).hex() == case["last_time_hex"]
        assert st["ties"] == case["ties"]
        hx = ["%016x" % x for x in hashes]
        if tie_mode == 2:
            assert hx == case["hashes"]
            assert st["distinguishable_ties"] == case["distinguishable_ties"]
        else:
            for i, dep in enumerate(case["order_dependent"]):
                if not dep:
                    assert hx[i] == case["hashes"][i]


# ---- payloads beyond 8 bytes (entity.js:52-60) ------------------------------------

BLOB_GOLD = os.path.join(os.path.dirname(__file__), "golden", "blob_simianpie.json")


def blob_golden_cases():
    with open(BLOB_GOLD) as f:
        return json.load(f)["cases"]


@pytest.mark.parametrize("case", blob_golden_cases(), ids=lambda c: c["name"])
def test_oracle_payloads_match_simianpie_goldens(oracle_mod, case):
    """events whose data is a list of words: same counts, trace hashes (they cover
    the fold of every payload the receivers acknowledged), digest, final time as the
    unmodified SimianPie engine"""
    from helpers import assert_blob_golden, run_blobs
    o = oracle_mod.Oracle(case["name"], 0.0, case["end"], case["min_delay"], seed=case["seed"])
    assert_blob_golden(case, *run_blobs(o, case))


def test_oracle_payloads_are_rank_invariant(oracle_mod):
    from helpers import run_blobs
    case = dict(name="b", n=300, end=6.0, min_delay=0.25, lookahead=0.25, max_words=12, seed=5)
    ref = run_blobs(oracle_mod.Oracle("b", 0.0, case["end"], case["min_delay"], seed=5), case)
    for size in (2, 3):
        out = run_blobs(oracle_mod.Oracle("b", 0.0, case["end"], case["min_delay"], size=size,
                                          seed=5, threads=size), case)
        assert out[0]["digest"] == ref[0]["digest"]
        np.testing.assert_array_equal(out[1], ref[1])
        np.testing.assert_array_equal(out[2], ref[2])
