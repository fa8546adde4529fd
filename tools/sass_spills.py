#!/usr/bin/env python
"""Where does k_window spill?  Disassembles the built library (no GPU needed) and
attributes LDL/STL instructions of the kernel's main body to source lines; also
prints the opcode histogram that goes into profiles/ (UBLKCP / SYNCS = bulk copies
+ mbarrier, LDGSTS = cp.async, ATOMS / ATOMG / RED, BAR).
usage: python tools/sass_spills.py [lib.so] [--all-functions] [--kernel=<mangled prefix>]
(default kernel: the ordinary window kernel k_window_t<false>; e.g. --kernel=_Z14k_window_stage,
--kernel=_Z6k_emit)"""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = next((a for a in sys.argv[1:] if not a.startswith("--")),
           os.path.join(ROOT, "simian_b200", "libsimian_b200.so"))
allf = "--all-functions" in sys.argv
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp,
                      stdout=subprocess.DEVNULL)
kern = next((a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--kernel=")),
            "_Z10k_window_tILb0E")
cubin = sorted(f for f in os.listdir(tmp) if f.endswith(".cubin") and "window_loop" not in f)[0]
txt = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)],
                     capture_output=True, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(txt)
             if l.startswith(".text." + kern) or l.startswith(".text._Z8k_window"))
end = next((i for i in range(start + 1, len(txt)) if txt[i].startswith("//-------") and ".text." in txt[i]),
           len(txt))
cur, fn = None, None
spill, ops, per_fn = collections.Counter(), collections.Counter(), collections.Counter()
for l in txt[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r'^\s*\.type\s+(\S+),@function', l)
    if m:
        fn = m.group(1)
    m = re.search(r'/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)', l)
    if not m:
        continue
    op = m.group(2)
    per_fn[fn or "k_window (body)"] += 1
    if fn is None or allf:
        ops[op.split(".")[0]] += 1
        if op.startswith(("LDL", "STL")):
            spill[(cur, op[:3])] += 1
src = open(os.path.join(ROOT, "simian_b200", "csrc", "kernels.cuh")).read().split("\n")
print("instructions per function:")
for k, v in per_fn.most_common():
    print("  %6d  %s" % (v, k[-70:]))
print("opcode histogram (%s):" % ("all functions of k_window" if allf else "k_window body"))
for k, v in sorted(ops.items(), key=lambda x: -x[1]):
    print("  %-10s %d" % (k, v))
print("spill instructions by source line:")
for ((f, ln), op), v in sorted(spill.items(), key=lambda x: -x[1])[:40]:
    print("  %3d %s %s:%d  %s" % (v, op, f, ln, src[ln - 1].strip()[:80] if f == "kernels.cuh" else ""))
