"""2-GPU check of the sharded driver (needs >= 2 GPUs: gpurun --gpus 2); skipped on one GPU."""
import os
import subprocess
import sys

import pytest

import pyammsb as A

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("mode,coll", [("partitioned", "peer"), ("replicated", "peer"), ("partitioned", "nccl"), ("columns", "peer")])
def test_sharded_driver_matches_single_gpu(mode, coll):
    n = A.device_count()
    if n < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    world = 2 if n < 4 else 4
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                        "--master-addr", "127.0.0.1", "--master-port", "29533",
                        os.path.join(ROOT, "tests", "run_dist_gpu.py"), mode, coll], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-3000:])
    assert "dist gpu check ok" in r.stdout
