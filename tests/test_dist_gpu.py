"""Multi-GPU parity: the sharded engine on 2, 4 and 8 GPUs of one box against the CPU oracle
(tests/dist_gpu_worker.py, one process per GPU under torchrun).  A world size is skipped when the
box has fewer GPUs; the driver's one-GPU test box skips all three -- `bench.py --gpus N` repeats a
sharded parity check in front of its timed region for that reason (sanity.sharded_parity)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count():
    try:
        out = subprocess.check_output(["nvidia-smi", "-L"], text=True)
        return len([l for l in out.splitlines() if l.startswith("GPU ")])
    except (OSError, subprocess.CalledProcessError):
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_parity_vs_oracle(world):
    if _gpu_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29611 + world),
           os.path.join(ROOT, "tests", "dist_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1500, cwd=ROOT)
    log = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(log):
        with open(os.path.join(log, f"dist_gpu_worker_{world}.log"), "w") as f:
            f.write(r.stdout[-20000:] + "\n--- stderr ---\n" + r.stderr[-5000:])
    assert r.returncode == 0 and "DIST_GPU_CHECK PASS" in r.stdout, r.stdout[-4000:] + r.stderr[-4000:]
