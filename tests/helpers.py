"""Shared helpers of the parity tests: run one PHOLD case on the oracle and on
the CUDA engine through the same model-builder (simian_b200.workloads)."""
import numpy as np

from simian_b200.workloads import build_phold


def run_oracle_phold(oracle_mod, case, size=1, tie_mode=0):
    o = oracle_mod.Oracle(case.get("name", "case"), 0.0, case["end"],
                          case["min_delay"], size=size, tie_mode=tie_mode,
                          seed=case.get("seed", 1))
    build_phold(o, case["n"], case["per_entity"], case["lookahead"],
                case.get("p_remote", 0.0), case.get("lam", 1.0),
                case.get("mode", 0), case.get("groups", 1))
    st = o.run()
    return st, o.read_counts("Node", case["n"]), o.read_trace_hashes(
        "Node", case["n"])


def run_engine_phold(case, **options):
    from simian_b200 import capi
    e = capi.Engine(case.get("name", "case"), 0.0, case["end"],
                    case["min_delay"], seed=case.get("seed", 1), **options)
    build_phold(e, case["n"], case["per_entity"], case["lookahead"],
                case.get("p_remote", 0.0), case.get("lam", 1.0),
                case.get("mode", 0), case.get("groups", 1))
    st = e.run()
    counts = e.read_counts("Node", case["n"])
    hashes = e.read_trace_hashes("Node", case["n"])
    e.close()
    return st, counts, hashes


def assert_parity(st_o, cnt_o, h_o, st_e, cnt_e, h_e):
    """bit-exact: per-entity counts, per-entity trace hashes, digest, total
    events, f64 last time, window count."""
    assert st_o["distinguishable_ties"] == 0, "oracle saw distinguishable ties"
    assert st_e["error"] == 0
    assert st_e["total_events"] == st_o["total_events"]
    np.testing.assert_array_equal(cnt_e, cnt_o)
    np.testing.assert_array_equal(h_e, h_o)
    assert st_e["digest"] == st_o["digest"]
    assert st_e["state_checksum"] == st_o["state_checksum"]
    assert float(st_e["last_time"]).hex() == float(st_o["last_time"]).hex()
    assert st_e["windows"] == st_o["windows"]
    assert st_e["emitted"] == st_o["emitted"]
    assert st_e["dropped_past_end"] == st_o["dropped_past_end"]
    assert st_e["ties"] == st_o["ties"]


# ---- LANL PDES benchmark ------------------------------------------------------

LANL_KEYS = ("n_ent", "s_ent", "q_avg", "p_receive", "m_ent", "ops_ent",
             "ops_sigma", "cache_friendliness", "end_time", "min_delay",
             "time_bins")


LANL_OPT_KEYS = ("p_send", "p_list", "invert", "stats_on_path")


def build_lanl_case(eng, case):
    from simian_b200.workloads import build_lanl
    kw = {k: case[k] for k in LANL_KEYS}
    kw.update({k: case[k] for k in LANL_OPT_KEYS if k in case})
    build_lanl(eng, **kw)
    return eng


def read_lanl(handle, case):
    """-> (counts, hashes, state struct array, lists)"""
    from simian_b200.workloads import (LANL_CLASS, lanl_state_bytes,
                                       lanl_states)
    n = case["n_ent"]
    pl = case.get("p_list", 0.0)
    raw = handle.read_state(LANL_CLASS, n, lanl_state_bytes(n, case["m_ent"], pl))
    st, lists = lanl_states(raw, n, case["m_ent"], pl)
    return (handle.read_counts(LANL_CLASS, n),
            handle.read_trace_hashes(LANL_CLASS, n), st, lists)


def run_oracle_lanl(oracle_mod, case, size=1, tie_mode=1):
    from simian_b200.workloads import lanl_engine_end
    o = oracle_mod.Oracle(case.get("name", "lanl"), 0.0,
                          lanl_engine_end(case["end_time"], case["min_delay"]),
                          case["min_delay"], size=size, tie_mode=tie_mode,
                          seed=case.get("seed", 1))
    build_lanl_case(o, case)
    st = o.run()
    return (st,) + read_lanl(o, case)


def run_engine_lanl(case, **options):
    from simian_b200 import capi
    from simian_b200.workloads import lanl_engine_end
    if case.get("stats_on_path"):
        options = dict(options, payloads=1)   # the statistics messages carry payloads
    e = capi.Engine(case.get("name", "lanl"), 0.0,
                    lanl_engine_end(case["end_time"], case["min_delay"]),
                    case["min_delay"], seed=case.get("seed", 1), **options)
    build_lanl_case(e, case)
    st = e.run()
    out = (st,) + read_lanl(e, case)
    e.close()
    return out


def assert_lanl_golden(case, st, counts, hashes, state, lists):
    """against tests/golden/lanl_simianpie.json (unmodified SimianPie engine)"""
    assert st["total_events"] == case["total_events"]
    assert list(counts) == case["counts"]
    assert ["%016x" % x for x in hashes] == case["hashes"]
    assert "%016x" % st["digest"] == case["digest"]
    assert float(st["last_time"]).hex() == case["last_time_hex"]
    assert st["ties"] == case["ties"]
    g = case["state"]
    for k in ("send_count", "receive_count", "ops_count", "q_size"):
        assert [int(x) for x in state[k]] == g[k], k
    for k in ("last_scheduled", "ops_max", "ops_min", "ops_mean", "value_sink"):
        assert [float(x).hex() for x in state[k]] == g[k], k
    assert [[int(x) for x in row] for row in state["time_sends"]] == g["time_sends"]
    assert [[float(x).hex() for x in row[:4]] for row in lists] == g["list_head"]
    if "global_stats" in case:  # the fan-in on the event path: what entity 0 collected
        gs, e0 = case["global_stats"], state[0]
        assert int(e0["stats_received"]) == gs["stats_received"] == case["n_ent"]
        for k in ("gsend_count", "greceive_count", "gops_count"):
            assert int(e0[k]) == gs[k], k
        for k in ("gops_max", "gops_min", "gops_mean"):
            assert float(e0[k]).hex() == gs[k], k
        assert [int(x) for x in e0["gtime_sends"]] == gs["gtime_sends"]
        assert int(e0["gsend_count"]) == int(state["send_count"].sum())


# ---- hello.js (two classes) ----------------------------------------------------

def run_hello(handle_factory, case):
    """-> (stats, counts[2*count], hashes[2*count]) in global id order"""
    from simian_b200.workloads import build_hello
    h = handle_factory()
    build_hello(h, case["count"])
    st = h.run()
    n = case["count"]
    counts = np.concatenate([h.read_counts("Alice", n), h.read_counts("Bob", n)])
    hashes = np.concatenate([h.read_trace_hashes("Alice", n), h.read_trace_hashes("Bob", n)])
    h.close()
    return st, counts, hashes


# ---- payloads beyond 8 bytes ("blobs" model) -------------------------------------

def run_blobs(handle, case):
    """-> (stats, counts, hashes)"""
    from simian_b200.workloads import build_blobs
    build_blobs(handle, case["n"], case["lookahead"], case["max_words"])
    st = handle.run()
    out = (st, handle.read_counts("Node", case["n"]), handle.read_trace_hashes("Node", case["n"]))
    handle.close()
    return out


def assert_blob_golden(case, st, counts, hashes):
    """against tests/golden/blob_simianpie.json (unmodified SimianPie engine, events
    whose data is a Python list)"""
    assert st["total_events"] == case["total_events"]
    assert list(counts) == case["counts"]
    assert ["%016x" % x for x in hashes] == case["hashes"]
    assert "%016x" % st["digest"] == case["digest"]
    assert float(st["last_time"]).hex() == case["last_time_hex"]
