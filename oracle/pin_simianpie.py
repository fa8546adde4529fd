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


if __name__ == "__main__":
    main()
