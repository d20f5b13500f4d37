"""The multi-GPU partitions on real devices: spawns ``torchrun --nproc-per-node 2`` over NCCL
when at least two GPUs are visible; the gloo / CPU version of the host logic is tests/test_distributed.py."""
import os
import socket
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2])
def test_partitions_over_nccl(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, {torch.cuda.device_count()} visible")
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1")
    res = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
         "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "mgpu_worker.py")],
        env=env, capture_output=True, text=True, timeout=900)
    tail = "\n".join((res.stdout + res.stderr).splitlines()[-30:])
    assert res.returncode == 0, tail
    assert "MGPU_OK" in res.stdout, tail
