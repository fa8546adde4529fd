"""Multi-rank path.  CPU: world_size 2 and 3 over gloo, the product's window
driver stepping the oracle's ranks.  GPU (needs >= 2 devices): the CUDA engine
over NCCL, one process per GPU."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _launch(nproc, *args, timeout=600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
           "--nproc-per-node", str(nproc), "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "dist_worker.py"), *args]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return r.stdout


@pytest.mark.parametrize("world", [2, 3])
def test_window_driver_over_gloo(world, oracle_mod):
    out = _launch(world, "--engine", "oracle", "--backend", "gloo", "--cases",
                  "ross,other_group,uniform_sparse,lanl,hello")
    assert out.count("DIST-OK") == 5


@pytest.mark.gpu
def test_cuda_engine_over_nccl_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run under gpurun --gpus 2)")
    world = min(torch.cuda.device_count(), 8)
    for w in sorted({2, world}):
        for driver in ("p2p", "p2p_deferred", "p2p_nccl", "nccl"):
            out = _launch(w, "--engine", "cuda", "--backend", "nccl", "--driver",
                          driver, "--cases",
                          "ross,other_group,uniform_sparse,big,lanl,hello")
            assert out.count("DIST-OK") == 6, driver
