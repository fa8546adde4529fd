"""-m gpu: the reference's PHOLD model script (phold-noop-noMPI.js:23-49),
written against the host-side mirror of the Simian / Entity API."""
import pytest

pytestmark = pytest.mark.gpu


def test_phold_script_shape_matches_oracle(oracle_mod):
    from simian_b200.simian import Entity, Simian, device_fn, phold_params
    from simian_b200.workloads import build_phold

    simName, startTime, endTime, minDelay, useMPI = "PHOLD-NOOP-NOMPI", 0, 50, 0.05, False
    count, lookahead = 10, minDelay

    class Node(Entity):
        generate = device_fn("phold_generate")  # Node.generate, :36-42

    eng = Simian(simName, startTime, endTime, minDelay, useMPI)
    for i in range(count):
        eng.addEntity("Node", Node, i)
    eng.setClassParams("Node", phold_params(eng.firstId("Node"), count, lookahead))
    for i in range(count):
        eng.schedService(0, "generate", None, "Node", i)
    st = eng.run()

    o = oracle_mod.Oracle(simName, startTime, endTime, minDelay)
    build_phold(o, count, 1, lookahead)
    so = o.run()
    assert st["total_events"] == so["total_events"] > 100
    assert st["digest"] == so["digest"]
    assert eng.now == so["last_time"]
    cnt = o.read_counts("Node", count)
    assert [eng.getEntity("Node", i).count for i in range(count)] == list(cnt)
    assert str(eng.getEntity("Node", 3)) == "Node(3)"
    eng.exit()


def test_reference_error_behaviour():
    from simian_b200.simian import Entity, Simian, SimianError, device_fn
    from simian_b200 import capi

    class Node(Entity):
        generate = device_fn("phold_generate")

    eng = Simian("err", 0, 10, 0.5)
    eng.addEntity("Node", Node, 0)
    ent = eng.getEntity("Node", 0)
    with pytest.raises(capi.SimianError) as ei:   # entity.js:42
        ent.reqService(0.1, "generate", None, "Node", 0)
    assert "idle" in str(ei.value)
    ent.reqService(0.1, "generate", None)         # rx omitted: no lookahead check (:41)
    ent.reqService(100.0, "generate", None)       # beyond endTime: dropped (:48)
    with pytest.raises(capi.SimianError):
        eng.attachService(Node, "other", lambda *a: None)  # must be a device fn
    with pytest.raises(capi.SimianError):         # unknown device function name
        eng.attachService(Node, "other", device_fn("no_such_fn"))
    eng.exit()


def test_hello_script_shape_matches_simianpie_golden():
    """SimianJS/Examples/hello.js:96-109 against the mirror API: two classes
    added interleaved, schedService per entity, run; per-entity counts equal
    the unmodified SimianPie engine's (tests/golden/hello_simianpie.json)."""
    import json
    import os
    from simian_b200.simian import Entity, Simian, device_fn
    from simian_b200.workloads import HELLO_SERVICES, hello_params_blob
    gold = os.path.join(os.path.dirname(__file__), "golden", "hello_simianpie.json")
    with open(gold) as f:
        case = json.load(f)["cases"][0]
    count = case["count"]

    class Alice(Entity):
        generate = device_fn("hello_generate")
        square = device_fn("hello_square")
        result = device_fn("hello_result")

    class Bob(Entity):
        generate = device_fn("hello_generate")
        sqrt = device_fn("hello_sqrt")
        result = device_fn("hello_result")

    eng = Simian("HELLO", 0, case["end"], case["min_delay"], False, seed=case["seed"],
                 bucket_width=1.0)
    ids = [eng.internService(s) for s in HELLO_SERVICES]
    for i in range(count):                       # hello.js:98-101
        eng.addEntity("Alice", Alice, i)
        eng.addEntity("Bob", Bob, i)
    blob = hello_params_blob(eng.firstId("Alice"), eng.firstId("Bob"), count, ids)
    eng.setClassParams("Alice", blob)
    eng.setClassParams("Bob", blob)
    for i in range(count):                       # hello.js:103-106
        eng.schedService(0, "generate", None, "Alice", i)
        eng.schedService(50, "generate", None, "Bob", i)
    # integer delays: distinguishable same-entity ties, whose commit order is engine
    # specific -- run() says so
    with pytest.warns(RuntimeWarning, match="distinguishable same-entity time ties"):
        st = eng.run()
    assert st["distinguishable_ties"] > 0
    assert st["total_events"] == case["total_events"]
    got = ([eng.getEntity("Alice", i).count for i in range(count)] +
           [eng.getEntity("Bob", i).count for i in range(count)])
    assert got == case["counts"]
    assert float(eng.now).hex() == case["last_time_hex"]
    eng.exit()


def test_rank_log_run_banner_and_state_readback(tmp_path, capsys, oracle_mod):
    """the per-rank log "<simName>.<rank>.out" (simian.js:83,225), what rank 0
    prints around run (simian.js:104-113,159-165) and getEntity state read-back"""
    import numpy as np
    from simian_b200.simian import Entity, Simian
    from simian_b200.workloads import (LANL_CLASS, LANL_STATE_DTYPE, build_lanl, lanl_engine_end,
                                       lanl_report, lanl_state_bytes, lanl_states)

    class Node(Entity):
        pass

    eng = Simian("LOGTEST", 0, 5, 0.5, log=str(tmp_path / "LOGTEST"), verbose=True)
    for i in range(3):
        eng.addEntity("Node", Node, i)
    eng.addEntities("Other", Node, 0, 2)
    from simian_b200.simian import device_fn
    eng.attachService(Node, "tick", device_fn("noop"))
    eng.schedService(1.0, "tick", None, "Node", 1)
    eng.schedService(2.0, "tick", None, "Other", 0)
    st = eng.run()
    assert st["total_events"] == 2
    eng.exit()
    text = (tmp_path / "LOGTEST.0.out").read_text().splitlines()
    assert text == ["Node(0) Running on rank 0", "Node(1) Running on rank 0",
                    "Node(2) Running on rank 0", "Other(0) Running on rank 0",
                    "Other(1) Running on rank 0"]
    printed = capsys.readouterr().out.splitlines()
    assert printed[1] == "===========================================" and "MPI: OFF" in printed
    assert "SIMULATED EVENTS       : 2" in printed
    assert any(l.startswith("SIMULATION FINISHED IN : ") for l in printed)
    assert any(l.startswith("MODEL-EVENTS/SECOND    : ") for l in printed)

    # LANL statistics table from the entity state read back from HBM
    case = dict(n_ent=40, s_ent=10.0, q_avg=1.0, p_receive=0.05, m_ent=8.0, ops_ent=20.0,
                ops_sigma=0.1, cache_friendliness=0.5, end_time=6.0, min_delay=0.5, time_bins=8)
    from simian_b200 import capi
    e = capi.Engine("lanl", 0.0, lanl_engine_end(6.0, 0.5), 0.5, seed=3)
    build_lanl(e, **case)
    s2 = e.run()
    raw = e.read_state(LANL_CLASS, 40, lanl_state_bytes(40, 8.0))
    state, _ = lanl_states(raw, 40, 8.0)
    e.close()
    import io
    out = io.StringIO()
    tot = lanl_report(state, 8, out)
    assert len(out.getvalue().splitlines()) == 1 + 40 + 4
    assert tot["sends"] == int(state["send_count"].sum()) > 0
    assert tot["receives"] == int(state["receive_count"].sum())
    assert s2["total_events"] > tot["sends"]


def test_placement_extension_points():
    """getBaseRank may be overridden (Docs/README.Simian:92-100); getOffsetRank
    may not: the device routing implements the reference form"""
    from simian_b200.simian import Entity, Simian, SimianError, device_fn

    class Node(Entity):
        tick = device_fn("noop")

    class MySimian(Simian):
        def getBaseRank(self, name):
            return 7

    eng = MySimian("place", 0, 5, 0.5)
    eng.addEntity("Node", Node, 0)
    assert eng.baseRanks["Node"] == 7 and eng._h.base_rank("Node") == 0   # 7 % size(1)
    eng.schedService(1.0, "tick", None, "Node", 0)
    assert eng.run()["total_events"] == 1
    eng.exit()

    class Bad(Simian):
        def getOffsetRank(self, name, num):
            return 0

    with pytest.raises(SimianError):
        Bad("bad", 0, 5, 0.5)
