"""Multi-GPU cell shard of the PatchMatch phase on real peer memory (CUDA IPC + NVLink): needs >= 2 GPUs in the box, skipped
otherwise (the single-GPU box of the round-end run; the logic is covered there by the two-rank in-process test of test_gpu_pm.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cell_shard_over_peer_memory_equals_the_single_gpu_sweep():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = min(n, 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "multi_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and f"MULTI_GPU_CHECK world={world} ok=True" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
