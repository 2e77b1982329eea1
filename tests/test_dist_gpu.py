"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped on a single-GPU box)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_two_gpu_parity_vs_oracle():
    try:
        out = subprocess.check_output(["nvidia-smi", "-L"], text=True)
        n_gpus = len([l for l in out.splitlines() if l.startswith("GPU ")])
    except (OSError, subprocess.CalledProcessError):
        n_gpus = 0
    if n_gpus < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29611",
           os.path.join(ROOT, "tests", "dist_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0 and "DIST_GPU_CHECK PASS" in r.stdout, r.stdout[-4000:] + r.stderr[-4000:]
