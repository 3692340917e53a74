"""Multi-GPU equivalence of the data-parallel path (needs >= 2 GPUs; skipped on a 1-GPU box)."""
import os
import subprocess
import sys

import pytest

import b200

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_data_parallel_update_matches_single_gpu():
    L = b200.lib()
    n = L.ppo_b200_device_count()
    if n < 2:
        pytest.skip("needs 2 GPUs (run through `gpurun --gpus 2`)")
    world = 8 if n >= 8 else 4 if n >= 4 else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", "29511", os.path.join(HERE, "dist_equivalence.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    sys.stdout.write(res.stdout[-4000:])
    assert res.returncode == 0 and "DIST_EQUIVALENCE_OK" in res.stdout, res.stdout[-2000:] + res.stderr[-4000:]
