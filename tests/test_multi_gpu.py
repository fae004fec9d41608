"""Sharded vs unsharded parity on real GPUs, inside the driver-run `-m gpu` suite: spawns tests/mgpu_check.py under
torch.distributed.run on every visible GPU (2, 4 or 8) and skips on a single-GPU box."""
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _run(world, extra=()):
    port = 29500 + (os.getpid() % 400)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "mgpu_check.py"), *extra]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    tail = (r.stdout + r.stderr)[-4000:]
    assert r.returncode == 0 and "MGPU_CHECK OK" in r.stdout, tail


def test_sharded_matches_unsharded_on_all_visible_gpus():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 8 if n >= 8 else (4 if n >= 4 else 2)
    _run(world, ("--c3",))


def test_sharded_matches_unsharded_world_2():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    if torch.cuda.device_count() < 4:
        pytest.skip("covered by the all-GPU test")
    _run(2)
