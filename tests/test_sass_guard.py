"""Toolchain guard (no GPU needed): every 256-bit record store of the PTX must be a
256-bit store in the SASS.

ptxas 12.9 has been seen to turn `st.global.v4.u64` (simian_b200/csrc/kernels.cuh,
st_rec) into a single 64-bit store inside functions that are not inlined into a
kernel body -- silently dropping 24 of a record's 32 bytes -- and whether it does
depends on unrelated code.  The kernels therefore use the 256-bit form only on call
chains that are force-inlined into a __global__ function (WIDE template flag); this
test disassembles the built library and fails if any store attributed to the
st_rec source line is narrower than 256 bits."""
import glob
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "simian_b200", "libsimian_b200.so")
SRC = os.path.join(ROOT, "simian_b200", "csrc", "kernels.cuh")


@pytest.mark.skipif(not (shutil.which("cuobjdump") and shutil.which("nvdisasm")),
                    reason="CUDA binary utilities not installed")
def test_wide_record_stores_stay_wide(tmp_path):
    if not os.path.exists(LIB):
        import __graft_entry__ as g
        g.build()
    with open(SRC) as f:
        first = [i + 1 for i, l in enumerate(f) if "st.global.v4.u64" in l and "asm" in l]
    assert first, "st_rec not found"
    # (the line table names any line of the multi-line asm statement)
    asm_lines = {ln + k for ln in first for k in range(4)}
    subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp_path, check=True,
                   capture_output=True)
    cubins = glob.glob(str(tmp_path / "*.cubin"))
    assert cubins
    wide = narrow = 0
    bad = []
    for cb in cubins:
        out = subprocess.run(["nvdisasm", "-g", "-c", cb], capture_output=True, text=True).stdout
        fn, line = None, None
        for l in out.splitlines():
            m = re.match(r"\s*\.text\.(\S+):", l)
            if m:
                fn = m.group(1)
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', l)
            if m:
                line = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.search(r"\bSTG(\.[A-Z0-9_]+)*", l)
            if m and line and line[0] == "kernels.cuh" and line[1] in asm_lines:
                if ".256" in m.group(0):
                    wide += 1
                else:
                    narrow += 1
                    bad.append((fn, m.group(0)))
    assert wide > 0, "no 256-bit record store found in the SASS at all"
    assert narrow == 0, "record stores narrowed by ptxas: %r" % bad[:8]
