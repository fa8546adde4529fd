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
