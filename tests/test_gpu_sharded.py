"""GPU, >= 2 devices: the sharded reduce / whole-array prefix reduction / mkperm
histogram on real GPUs -- both the peer-mailbox path (csrc/sharded.cu, one C-ABI call
per rank) and the NCCL path of drjit-core_b200/sharded.py -- against the CPU oracle.
Spawns `torchrun --nproc-per-node 2 tests/sharded_worker.py`; skipped on a 1-GPU box
(run it with `gpurun --gpus 2`).  The reference has no counterpart (SURVEY.md
section 2: no NCCL anywhere), so this test is the parity evidence for section 8e."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_on_gpus(dr, world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, found {torch.cuda.device_count()}")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "sharded_worker.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert p.returncode == 0 and f"SHARDED_OK world={world}" in p.stdout, (p.stdout[-3000:], p.stderr[-3000:])
