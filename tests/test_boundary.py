"""-m "not gpu": the C-ABI library loads on a CPU-only box, exports every
symbol include/simian_b200.h declares, and refuses to compute without a GPU."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "simian_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(simian_[a-z_0-9]+)\s*\(", src)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    g.build()
    from simian_b200 import capi
    return capi


def test_header_declares_the_expected_entry_points():
    syms = declared_symbols()
    for s in ("simian_create", "simian_add_entities", "simian_attach_service",
              "simian_sched_service", "simian_run", "simian_destroy",
              "simian_read_counts", "simian_read_trace_hashes"):
        assert s in syms


def test_library_exports_every_declared_symbol(built):
    L = ctypes.CDLL(built.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(L, s), s


def test_binding_covers_every_declared_symbol(built):
    assert sorted(built.SIGNATURES) == declared_symbols()


def test_library_is_sm100a_only(built):
    out = subprocess.run(["cuobjdump", "-lelf", built.LIB_PATH],
                         capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback(built):
    """without a CUDA device the product path fails loudly"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(built.SimianError) as ei:
        built.Engine("x", 0.0, 1.0, 0.1)
    assert ei.value.code == 14 and "no CPU fallback" in str(ei.value)


def test_world_larger_than_the_peer_tables_is_refused(built):
    """size > SIMIAN_MAX_RANKS (16) would overrun the engine's peer tables
    (include/simian_b200.h documents the limit): refused at create, GPU or not"""
    with pytest.raises(built.SimianError) as ei:
        built.Engine("x", 0.0, 1.0, 0.1, rank=0, size=17)
    assert "SIMIAN_MAX_RANKS" in str(ei.value)
    with pytest.raises(built.SimianError):
        built.Engine("x", 0.0, 1.0, 0.1, rank=3, size=2)


def test_hash_name_matches_hash_js(built):
    assert built.lib().simian_hash_name(b"Node") == 6385183179.0


def test_product_never_touches_the_oracle():
    """only tests/, __graft_entry__.smoke and bench.py may use oracle/"""
    pkg = os.path.join(ROOT, "simian_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".js", ".cc")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower(), os.path.join(dirpath, f)
