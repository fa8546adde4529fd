#!/usr/bin/env python
"""Generate golden vectors by running the UNMODIFIED reference engine.

TEST INFRASTRUCTURE ONLY.  Runs in the build container, where /root/reference
exists; the GPU box never runs this.  It imports SimianPie/simian.py (the
reference's own Python implementation of eventQ + simian + entity, see
SimianPie/simian.py:256-289,1584-1655) untouched, drives it with a PHOLD model
whose handler is the pure-Python transcription of include/simian_spec.h /
simian_models.h (oracle/pyspec.py), and writes per-entity committed counts,
per-entity order-sensitive trace hashes, the digest, the total event count and
the final engine time to tests/golden/phold_simianpie.json.

SimianPie differs from SimianJS only in the order of equal-time events (FIFO
`ec` counter, simian.py:286-287) -- irrelevant on inputs without
distinguishable same-entity ties, which the generated cases are (checked).

usage: python oracle/pin_simianpie.py [--reference /root/reference]
"""
import argparse
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import pyspec  # noqa: E402

CASES = [
    # name, n, per_entity, min_delay, lookahead, p_remote, lambda, mode, groups, end, seed
    dict(name="config1_shape", n=1000, per_entity=1, min_delay=1e-4,
         lookahead=1e-4, p_remote=0.0, lam=1.0, mode=0, groups=1, end=20.0,
         seed=1),
    dict(name="ross_remote_10pct", n=256, per_entity=4, min_delay=0.1,
         lookahead=0.1, p_remote=0.1, lam=1.0, mode=1, groups=1, end=10.0,
         seed=7),
    dict(name="other_group_50pct", n=64, per_entity=16, min_delay=0.1,
         lookahead=0.1, p_remote=0.5, lam=1.0, mode=2, groups=4, end=5.0,
         seed=3),
    dict(name="lambda2_small_lookahead", n=100, per_entity=3, min_delay=0.01,
         lookahead=0.01, p_remote=0.25, lam=2.0, mode=1, groups=1, end=8.0,
         seed=11),
]


def run_case(Simian, case):
    n = case["n"]
    eng = Simian("pin_" + case["name"], 0.0, case["end"], case["min_delay"],
                 False, silent=True)
    stats = {"ties": 0, "dist_ties": 0}

    class Node(eng.Entity):
        def __init__(self, baseInfo, *args):
            super(Node, self).__init__(baseInfo)
            self.count = 0
            self.hash = 0
            self.last = None

        def generate(self, data, tx, txId):
            now = self.engine.now
            src = pyspec.NULL_ENTITY if txId is None else txId
            d = 0 if data is None else data
            cur = (now, 0, src, d)
            if self.last is not None and self.last[0] == now:
                stats["ties"] += 1
                if self.last != cur:
                    stats["dist_ties"] += 1
            self.last = cur
            idx = self.count
            self.hash = pyspec.trace_step(self.hash, now, 0, src, d)
            self.count = idx + 1
            target, offset = pyspec.phold_generate(
                case["seed"], self.num, idx, n, case["p_remote"], case["lam"],
                case["lookahead"], case["mode"], case["groups"])
            self.reqService(offset, "generate", None, "Node", target)

    for i in range(n):
        eng.addEntity("Node", Node, i)
    for _ in range(case["per_entity"]):
        for i in range(n):
            eng.schedService(0.0, "generate", None, "Node", i)
    eng.run()
    ents = [eng.getEntity("Node", i) for i in range(n)]
    counts = [e.count for e in ents]
    hashes = [e.hash for e in ents]
    digest = 0
    for i in range(n):
        digest = (digest + pyspec.digest_term(i, counts[i], hashes[i])) & pyspec.M64
    out = dict(case)
    out.update(total_events=sum(counts), last_time_hex=float(eng.now).hex(),
               digest="%016x" % digest, counts=counts,
               hashes=["%016x" % h for h in hashes], ties=stats["ties"],
               distinguishable_ties=stats["dist_ties"])
    eng.exit()
    return out


LANL_CASES = [
    dict(name="lanl_uniform", n_ent=48, s_ent=20.0, q_avg=1.0, p_receive=0.0,
         m_ent=40.0, ops_ent=60.0, ops_sigma=0.1, cache_friendliness=0.5,
         end_time=20.0, min_delay=1.0, time_bins=10, seed=5),
    dict(name="lanl_hot_spots_q3", n_ent=40, s_ent=15.0, q_avg=3.0,
         p_receive=0.05, m_ent=64.0, ops_ent=100.0, ops_sigma=0.3,
         cache_friendliness=0.25, end_time=12.0, min_delay=0.5, time_bins=8,
         seed=9),
    # geometric source and list-size distributions (pdes_lanl_benchmarkV8.js:209-221):
    # entity 0 sends most and owns the longest list; every entity keeps ops >= 1
    dict(name="lanl_geometric_send_list", n_ent=36, s_ent=18.0, q_avg=2.0,
         p_receive=0.03, p_send=0.06, p_list=0.05, invert=False, m_ent=48.0,
         ops_ent=120.0, ops_sigma=0.2, cache_friendliness=0.5, end_time=14.0,
         min_delay=0.5, time_bins=7, seed=13),
    # inverted source distribution: the high ids send most, the low ids receive most
    dict(name="lanl_geometric_inverted", n_ent=30, s_ent=25.0, q_avg=1.0,
         p_receive=0.08, p_send=0.1, p_list=0.04, invert=True, m_ent=32.0,
         ops_ent=90.0, ops_sigma=0.1, cache_friendliness=0.75, end_time=10.0,
         min_delay=1.0, time_bins=10, seed=17),
    # the statistics fan-in ON the event path (pdes_lanl_benchmarkV8.js:311-342):
    # FinishHandler sends its message (a list) to OutputHandler of entity 0 -- n_ent
    # events at one instant at one entity, each with a payload of 7 + time_bins words
    dict(name="lanl_stats_on_path", n_ent=44, s_ent=16.0, q_avg=2.0, p_receive=0.04,
         m_ent=40.0, ops_ent=80.0, ops_sigma=0.2, cache_friendliness=0.5, end_time=9.0,
         min_delay=0.5, time_bins=9, stats_on_path=True, seed=21),
    dict(name="lanl_stats_on_path_geometric", n_ent=32, s_ent=20.0, q_avg=1.0,
         p_receive=0.06, p_send=0.07, p_list=0.05, invert=False, m_ent=36.0,
         ops_ent=110.0, ops_sigma=0.1, cache_friendliness=0.5, end_time=8.0,
         min_delay=1.0, time_bins=6, stats_on_path=True, seed=23),
]
LANL_SERVICES = {"SendHandler": 0, "ReceiveHandler": 1, "FinishHandler": 2, "OutputHandler": 3}


def run_lanl_case(Simian, case):
    """pdes_lanl_benchmarkV8.js "MAIN" (:346-356) on the unmodified SimianPie
    engine; handlers = oracle/pyspec.py transcription of simian_models.h."""
    n = case["n_ent"]
    P = pyspec.LanlParams(n, case["s_ent"], case["q_avg"], case["p_receive"],
                          case["m_ent"], case["ops_ent"], case["ops_sigma"],
                          case["cache_friendliness"], case["end_time"],
                          case["min_delay"], case["time_bins"],
                          case.get("p_send", 0.0), case.get("p_list", 0.0),
                          case.get("invert", False), case.get("stats_on_path", False))
    eng = Simian("pin_" + case["name"], 0.0,
                 max(case["end_time"], 2) + 3 * case["min_delay"],
                 case["min_delay"], False, silent=True)
    stats = {"ties": 0, "dist_ties": 0}

    class Ctx:
        def __init__(self, ent, commit_idx):
            self.ent = ent
            self.now = ent.engine.now
            self.key = pyspec.rng_key(case["seed"], ent.num, commit_idx)

        def draw(self, i):
            return pyspec.rng_draw(self.key, i)

        def req_service(self, offset, service, data, rx):
            if rx is None:
                self.ent.reqService(offset, service, data)
            else:
                self.ent.reqService(offset, service, data, "PDES_LANL_Node", rx)

    class PDES_LANL_Node(eng.Entity):
        def __init__(self, baseInfo, *args):
            super(PDES_LANL_Node, self).__init__(baseInfo)
            self.count = 0
            self.hash = 0
            self.last = None
            pyspec.lanl_construct(P, self, Ctx(self, pyspec.CTOR_COMMIT))

        def _commit(self, service, data, txId):
            now = self.engine.now
            src = pyspec.NULL_ENTITY if txId is None else txId
            h = LANL_SERVICES[service]
            cur = (now, h, src, data)
            if self.last is not None and self.last[0] == now:
                stats["ties"] += 1
                if self.last != cur:
                    stats["dist_ties"] += 1
            self.last = cur
            idx = self.count
            self.hash = pyspec.trace_step(self.hash, now, h, src, data)
            self.count = idx + 1
            return Ctx(self, idx)

        def SendHandler(self, data, tx, txId):
            pyspec.lanl_send(P, self, self._commit("SendHandler", data, txId))

        def ReceiveHandler(self, data, tx, txId):
            pyspec.lanl_receive(P, self, self._commit("ReceiveHandler", data, txId))

        def FinishHandler(self, data, tx, txId):
            self._commit("FinishHandler", data, txId)
            if P.stats_on_path:                                  # :311-315
                self.reqService(case["min_delay"], "OutputHandler",
                                pyspec.lanl_finish_msg(P, self), "PDES_LANL_Node", 0)

        def OutputHandler(self, msg, tx, txId):                  # :317-342
            # (the header record's data word = the number of payload words)
            self._commit("OutputHandler", len(msg), txId)
            pyspec.lanl_output(P, self, msg)

    for i in range(n):
        eng.addEntity("PDES_LANL_Node", PDES_LANL_Node, i)
    eng.run()
    ents = [eng.getEntity("PDES_LANL_Node", i) for i in range(n)]
    counts = [e.count for e in ents]
    hashes = [e.hash for e in ents]
    digest = 0
    for i in range(n):
        digest = (digest + pyspec.digest_term(i, counts[i], hashes[i])) & pyspec.M64
    out = dict(case)
    out.update(
        total_events=sum(counts), last_time_hex=float(eng.now).hex(),
        digest="%016x" % digest, counts=counts,
        hashes=["%016x" % h for h in hashes], ties=stats["ties"],
        distinguishable_ties=stats["dist_ties"],
        state=dict(
            send_count=[e.send_count for e in ents],
            receive_count=[e.receive_count for e in ents],
            ops_count=[e.ops_count for e in ents],
            q_size=[e.q_size for e in ents],
            time_sends=[e.time_sends for e in ents],
            last_scheduled=[float(e.last_scheduled).hex() for e in ents],
            ops_max=[float(e.ops_max).hex() for e in ents],
            ops_min=[float(e.ops_min).hex() for e in ents],
            ops_mean=[float(e.ops_mean).hex() for e in ents],
            value_sink=[float(e.value_sink).hex() for e in ents],
            list_head=[[float(x).hex() for x in e.list[:4]] for e in ents]))
    if P.stats_on_path:  # what entity 0 collected (:321-330)
        e0 = ents[0]
        out["global_stats"] = dict(
            stats_received=e0.stats_received, gsend_count=e0.gsend_count,
            greceive_count=e0.greceive_count, gops_count=e0.gops_count,
            gops_max=float(e0.gops_max).hex(), gops_min=float(e0.gops_min).hex(),
            gops_mean=float(e0.gops_mean).hex(), gtime_sends=list(e0.gtime_sends))
    eng.exit()
    return out


HELLO_CASES = [
    dict(name="hello_10", count=10, end=400.0, min_delay=1e-4, seed=2),
    dict(name="hello_37", count=37, end=300.0, min_delay=0.5, seed=6),
]
HELLO_SVC = {"generate": 0, "square": 1, "sqrt": 2, "result": 3}


def run_hello_case(Simian, case):
    """SimianJS/Examples/hello.js on the unmodified SimianPie engine; handlers =
    transcription of simian_hello_* (simian_models.h).  Global ids: Alice(i) = i,
    Bob(i) = count + i."""
    import math
    count = case["count"]
    eng = Simian("pin_" + case["name"], 0.0, case["end"], case["min_delay"], False,
                 silent=True)
    stats = {"ties": 0, "dist_ties": 0}
    first = {"Alice": 0, "Bob": count}

    class Base(eng.Entity):
        def __init__(self, baseInfo, *args):
            super(Base, self).__init__(baseInfo)
            self.gid = first[self.name] + self.num
            self.count = 0
            self.hash = 0
            self.last = None
            self.dist_tied = False

        def _commit(self, service, data, tx, txId):
            now = self.engine.now
            src = pyspec.NULL_ENTITY if txId is None else first[tx] + txId
            d = 0 if data is None else pyspec.f64_bits(data)
            h = HELLO_SVC[service]
            cur = (now, h, src, d)
            if self.last is not None and self.last[0] == now:
                stats["ties"] += 1
                if self.last != cur:
                    stats["dist_ties"] += 1
                    self.dist_tied = True
            self.last = cur
            idx = self.count
            self.hash = pyspec.trace_step(self.hash, now, h, src, d)
            self.count = idx + 1
            return idx

        def generate(self, data, tx, txId):                      # hello.js:32-48
            idx = self._commit("generate", data, tx, txId)
            t, tid, val = pyspec.hello_generate(case["seed"], self.gid, idx, count)
            self.reqService(10, "sqrt" if t else "square", val, "Bob" if t else "Alice", tid)
            self.reqService(25, "generate", None)

        def result(self, data, tx, txId):                        # hello.js:54-59
            self._commit("result", data, tx, txId)

    class Alice(Base):
        def square(self, data, tx, txId):                        # hello.js:50-52
            self._commit("square", data, tx, txId)
            self.reqService(10, "result", data * data, tx, txId)

    class Bob(Base):
        def sqrt(self, data, tx, txId):                          # hello.js:86-88
            self._commit("sqrt", data, tx, txId)
            self.reqService(10, "result", math.sqrt(data), tx, txId)

    for i in range(count):                                       # hello.js:98-101
        eng.addEntity("Alice", Alice, i)
        eng.addEntity("Bob", Bob, i)
    for i in range(count):                                       # hello.js:103-106
        eng.schedService(0, "generate", None, "Alice", i)
        eng.schedService(50, "generate", None, "Bob", i)
    eng.run()
    ents = ([eng.getEntity("Alice", i) for i in range(count)] +
            [eng.getEntity("Bob", i) for i in range(count)])
    out = dict(case)
    out.update(total_events=sum(e.count for e in ents),
               last_time_hex=float(eng.now).hex(),
               counts=[e.count for e in ents],
               hashes=["%016x" % e.hash for e in ents],
               order_dependent=[bool(e.dist_tied) for e in ents],
               ties=stats["ties"], distinguishable_ties=stats["dist_ties"])
    eng.exit()
    return out


BLOB_CASES = [
    dict(name="blobs_40", n=40, end=12.0, min_delay=0.25, lookahead=0.25, max_words=9, seed=3),
    dict(name="blobs_long_payloads", n=25, end=8.0, min_delay=0.5, lookahead=0.5,
         max_words=24, seed=8),
]
BLOB_SVC = {"send": 0, "recv": 1, "ack": 2}


def run_blob_case(Simian, case):
    """Events whose `data` is a LIST (entity.js:52-60: any object) on the unmodified
    SimianPie engine; handlers = transcription of simian_blob_* (simian_models.h).
    The commit bookkeeping hashes the 8-byte data word the engine's header record
    carries: the number of payload words ("recv"), the fold ("ack"), 0 ("send")."""
    n = case["n"]
    eng = Simian("pin_" + case["name"], 0.0, case["end"], case["min_delay"], False,
                 silent=True)
    stats = {"ties": 0, "dist_ties": 0}

    class Node(eng.Entity):
        def __init__(self, baseInfo, *args):
            super(Node, self).__init__(baseInfo)
            self.count = 0
            self.hash = 0
            self.last = None

        def _commit(self, service, word, txId):
            now = self.engine.now
            src = pyspec.NULL_ENTITY if txId is None else txId
            h = BLOB_SVC[service]
            cur = (now, h, src, word)
            if self.last is not None and self.last[0] == now:
                stats["ties"] += 1
                if self.last != cur:
                    stats["dist_ties"] += 1
            self.last = cur
            idx = self.count
            self.hash = pyspec.trace_step(self.hash, now, h, src, word)
            self.count = idx + 1
            return idx

        def send(self, data, tx, txId):
            idx = self._commit("send", 0, txId)
            target, words, d_recv, d_self = pyspec.blob_send(
                case["seed"], self.num, idx, n, case["max_words"], case["lookahead"])
            self.reqService(d_recv, "recv", words, "Node", target)
            self.reqService(d_self, "send", None)

        def recv(self, data, tx, txId):
            self._commit("recv", len(data), txId)
            self.reqService(case["lookahead"], "ack", pyspec.blob_fold(data), tx, txId)

        def ack(self, data, tx, txId):
            self._commit("ack", data, txId)

    for i in range(n):
        eng.addEntity("Node", Node, i)
    for i in range(n):
        eng.schedService(0, "send", None, "Node", i)
    eng.run()
    ents = [eng.getEntity("Node", i) for i in range(n)]
    counts = [e.count for e in ents]
    hashes = [e.hash for e in ents]
    digest = 0
    for i in range(n):
        digest = (digest + pyspec.digest_term(i, counts[i], hashes[i])) & pyspec.M64
    out = dict(case)
    out.update(total_events=sum(counts), last_time_hex=float(eng.now).hex(),
               digest="%016x" % digest, counts=counts,
               hashes=["%016x" % h for h in hashes], ties=stats["ties"],
               distinguishable_ties=stats["dist_ties"])
    eng.exit()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(HERE, "..", "tests", "golden",
                                                  "phold_simianpie.json"))
    a = ap.parse_args()
    sys.path.insert(0, os.path.join(a.reference, "SimianPie"))
    from simian import Simian  # the unmodified reference engine
    out_path = os.path.abspath(a.out)
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)  # SimianPie writes "<name>.<rank>.out" into the cwd
        try:
            res = [run_case(Simian, c) for c in CASES]
        finally:
            os.chdir(cwd)
    for r in res:
        assert r["distinguishable_ties"] == 0, r["name"]
        print(r["name"], "events", r["total_events"], "digest", r["digest"],
              "ties", r["ties"])
    with open(out_path, "w") as f:
        json.dump({"generator": "oracle/pin_simianpie.py",
                   "reference_engine": "SimianPie/simian.py (unmodified)",
                   "cases": res}, f, separators=(",", ":"))
    print("wrote", out_path, os.path.getsize(out_path), "bytes")
    # LANL PDES benchmark: the t = minDelay requests of the constructors tie at
    # hot destinations; SimianPie pops equal times FIFO (simian.py:286-287),
    # which for these equals the engine's (time, handler, src, seq) order.
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            lres = [run_lanl_case(Simian, c) for c in LANL_CASES]
        finally:
            os.chdir(cwd)
    for r in lres:
        print(r["name"], "events", r["total_events"], "digest", r["digest"],
              "ties", r["ties"], "distinguishable", r["distinguishable_ties"])
    lpath = os.path.join(os.path.dirname(out_path), "lanl_simianpie.json")
    with open(lpath, "w") as f:
        json.dump({"generator": "oracle/pin_simianpie.py",
                   "reference_engine": "SimianPie/simian.py (unmodified)",
                   "tie_order": "FIFO (SimianPie); equals the engine order "
                                "(oracle tie_mode=1) on these cases",
                   "cases": lres}, f, separators=(",", ":"))
    print("wrote", lpath, os.path.getsize(lpath), "bytes")
    # hello.js: integer delays -> many distinguishable same-entity ties, whose
    # order is engine specific (FIFO here, heap shape in SimianJS): counts are
    # pinned for every entity, trace hashes for the entities that saw none.
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            hres = [run_hello_case(Simian, c) for c in HELLO_CASES]
        finally:
            os.chdir(cwd)
    for r in hres:
        print(r["name"], "events", r["total_events"], "ties", r["ties"], "distinguishable",
              r["distinguishable_ties"], "order-dependent entities", sum(r["order_dependent"]))
    hpath = os.path.join(os.path.dirname(out_path), "hello_simianpie.json")
    with open(hpath, "w") as f:
        json.dump({"generator": "oracle/pin_simianpie.py",
                   "reference_engine": "SimianPie/simian.py (unmodified)",
                   "cases": hres}, f, separators=(",", ":"))
    print("wrote", hpath, os.path.getsize(hpath), "bytes")
    # payloads beyond 8 bytes (events whose data is a list)
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            bres = [run_blob_case(Simian, c) for c in BLOB_CASES]
        finally:
            os.chdir(cwd)
    for r in bres:
        assert r["distinguishable_ties"] == 0, r["name"]
        print(r["name"], "events", r["total_events"], "digest", r["digest"], "ties", r["ties"])
    bpath = os.path.join(os.path.dirname(out_path), "blob_simianpie.json")
    with open(bpath, "w") as f:
        json.dump({"generator": "oracle/pin_simianpie.py",
                   "reference_engine": "SimianPie/simian.py (unmodified)",
                   "cases": bres}, f, separators=(",", ":"))
    print("wrote", bpath, os.path.getsize(bpath), "bytes")


if __name__ == "__main__":
    main()
