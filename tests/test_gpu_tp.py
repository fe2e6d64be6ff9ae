"""GPU (>= 2 devices): tensor-parallel shards + all-reduce reproduce the single-GPU logits
(BASELINE.json config 5).  Launches scripts/tp_check.py under torchrun; skipped on 1-GPU boxes
(the driver's round-end `pytest -m gpu` runs on one GPU; `gpurun --gpus 2` runs this for real)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world,shape", [(2, "512,1376,3,8,8,-1024,96"), (2, "256,688,2,8,4,512,64"),
                                         (4, "512,1376,2,8,8,-1024,64"), (8, "1024,2752,2,16,8,-2048,48")])
def test_tp_matches_single_gpu(world, shape):
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + world),
           os.path.join(ROOT, "scripts", "tp_check.py"), "--shape", shape, "--steps", "20"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "PASS" in r.stdout
