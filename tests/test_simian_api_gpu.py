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
