"""Handlers a model author registers without editing the library
(attachService(klass, name, fun) with `fun` = user code, SimianJS/simian.js:207-209):
tests/user_handlers/bounce.{h,def} compiled into a variant of the engine and of the
oracle with make HANDLERS=... (include/simian_handlers.def)."""
import ctypes
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PREFIX = os.path.join(ROOT, "tests", "user_handlers", "bounce")
ENGINE = os.path.join(ROOT, "simian_b200", "libsimian_b200_user.so")
ORACLE = os.path.join(ROOT, "oracle", "_build", "libsimian_oracle_user.so")


@pytest.fixture(scope="module")
def variants():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "simian_b200", "csrc"),
                           "HANDLERS=" + PREFIX, "NAME=user"])
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"),
                           "HANDLERS=" + PREFIX, "NAME=user"])
    return ENGINE, ORACLE


def test_user_handlers_are_in_the_registry(variants):
    L = ctypes.CDLL(variants[0])
    L.simian_device_function_id.argtypes = [ctypes.c_char_p]
    assert L.simian_device_function_id(b"user_bounce") == 32   # SIMIAN_FN_USER + 0
    assert L.simian_device_function_id(b"user_tick") == 33
    assert L.simian_device_function_id(b"phold_generate") == 1  # the built-ins stay
    assert L.simian_device_function_id(b"no_such_handler") == 0
    # the stock library does not know them
    from simian_b200 import capi
    import __graft_entry__ as g
    g.build()
    assert capi.lib().simian_device_function_id(b"user_bounce") == 0
    assert capi.lib().simian_device_function_id(b"lanl_receive") == 3


@pytest.mark.gpu
def test_user_handler_model_matches_the_oracle(variants):
    env = dict(os.environ, SIMIAN_B200_LIB=variants[0], SIMIAN_ORACLE_LIB=variants[1])
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "user_handlers",
                                                     "run_bounce.py")],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0 and "USER-OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
