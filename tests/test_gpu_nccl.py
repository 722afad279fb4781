"""N-GPU == 1-GPU on hardware (NCCL over NVLink): needs at least two CUDA devices, skipped otherwise
(`gpurun --gpus 2 -- python -m pytest tests/test_gpu_nccl.py -m gpu`).  The same check runs inside every
`bench.py --gpus N` (N > 1) before the warm-up and is reported as "multi_gpu_parity" in its JSON line."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 4])
def test_nccl_sharded_run_equals_single_gpu(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} CUDA devices, found {torch.cuda.device_count()}")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "dist_nccl_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert f"NCCL_PARITY_OK world={world}" in r.stdout
