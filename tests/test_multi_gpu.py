"""GPU tests of the multi-GPU paths (need >= 2 GPUs; skipped otherwise): one process per GPU under torchrun with
the NVLink peer-memory exchange and with the NCCL fallback, both against the single-GPU prices."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P1 = (100, 100, 0.05, 0.2, 1)


def rel(a, b):
    return abs(a - b) / max(abs(b), 1e-300)


def torchrun(nproc, env_extra):
    env = dict(os.environ, **env_extra)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "mgpu_worker.py")]
    p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("MGPU ")][0]
    return json.loads(line[5:])


@pytest.mark.parametrize("mode", ["peer", "nccl"])
def test_torchrun_two_ranks_match_single_gpu(gpu, mode):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    one = {
        "asia": gpu.mc_asia(*P1, 1_000_001, 252, "call", seed=31).price,
        "eur": gpu.mc_eur(*P1, 3_000_001, "put", seed=31).price,
        "basket": gpu.mc_eur_multi(*P1, 500_001, "call", 16, 0.5, seed=31).price,
        "amer": gpu.mc_amer(*P1, 1_000_002, 50, "put", seed=31).price,
        "amer_call": gpu.mc_amer(100, 110, 0.02, 0.75, 1, 200_000, 20, "call", seed=31).price,
        "binom": gpu.binom(*P1, 1_000_001, "call").price,
        "amer_tiny": gpu.mc_amer(*P1, 2, 5, "put", seed=31).price,
        "asia_tiny": gpu.mc_asia(*P1, 1, 7, "call", seed=31).price,
        "eur_tiny": gpu.mc_eur(*P1, 1, "put", seed=31).price,
        "binom_tiny": gpu.binom(*P1, 1, "call").price,
        "amer_lsm": gpu.mc_amer(*P1, 400_000, 50, "put", seed=31, lsm=True).price,
    }
    two = torchrun(2, {"PCF_NO_PEER": "1"} if mode == "nccl" else {})
    assert two["world"] == 2
    if mode == "nccl":
        assert two["peer"] is False
    for k, v in one.items():
        assert rel(two[k], v) < 1e-13 or abs(two[k] - v) < 1e-13, (k, two[k], v)   # identical normal stream, only summation order differs
    assert two["amer_repeat"] == two["amer"]
