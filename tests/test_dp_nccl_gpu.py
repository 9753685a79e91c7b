"""Data-parallel step over NCCL on real hardware (needs >= 2 GPUs; skipped otherwise).  ``gpurun --gpus 2 -- python -m
pytest tests/test_dp_nccl_gpu.py -m gpu``.  The checks themselves are in tests/dp_nccl_worker.py; the same semantics are
covered on CPU (gloo, world 2) by tests/test_data_parallel_gloo.py."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_stage1_data_parallel_nccl_world2(mode):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run under gpurun --gpus 2)")
    port = 29600 + os.getpid() % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "dp_nccl_worker.py"), mode]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"dp_nccl_{mode}.log"), "w") as f:
        f.write(r.stdout + "\n--- stderr ---\n" + r.stderr[-4000:])
    assert r.returncode == 0 and "DP_NCCL_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
