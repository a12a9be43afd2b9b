"""Multi-GPU data path on hardware: `sharding.solve_sharded_nccl` under torch.distributed.run with world size 2 (NCCL over NVLink),
the REAL solver on every rank, gathered result compared bit for bit with a single-GPU solve (tests/dist_worker_nccl.py).
Needs a box with >= 2 GPUs (`gpurun --gpus 2`); skipped on a single-GPU box.  The host-side logic of the same module runs on CPU
with gloo in tests/test_sharding_gloo.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_solve_sharded_over_nccl_world2_equals_single_gpu_solve_bitwise():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (NCCL refuses two ranks on one device)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tests", "dist_worker_nccl.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0 and "SHARDED_NCCL_OK" in r.stdout, (r.stdout[-2000:], r.stderr[-2000:])
