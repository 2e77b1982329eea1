"""N>1 host logic on CPU: world_size 2 and 4 over gloo (see tests/dist_worker.py)."""
import math
import os
import subprocess
import sys

import pytest

from oracle.pyoracle import random_circuit_script

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bv(n, secret):
    return [("bv", secret)]


CASES = {
    # name: (qubits, semantics, script)
    "h_on_every_qubit": (14, "corrected", [("h", q) for q in range(14)]),
    "global_controls_and_diagonals": (14, "corrected",
        [("h", q) for q in range(14)] + [("cnot", 13, 2), ("cnot", 12, 13), ("rz", 13, 0.3), ("z", 12),
                                          ("cphase", 13, 12, 0.7), ("cphase", 3, 13, -0.2), ("phase", 12, 1.1),
                                          ("cnot", 13, 12), ("ry", 13, 0.4), ("rx", 12, 1.3), ("cnot", 0, 13)]),
    "bv_ancilla_is_global": (14, "corrected", _bv(14, 0x1A5B)),
    "qft": (13, "corrected", [("x", 3), ("ry", 12, 0.6), ("qft",)]),
    "random_brickwork": (14, "corrected", random_circuit_script(14, 6, seed=77)),
    "random_brickwork_reference_semantics": (14, "reference", random_circuit_script(14, 4, seed=78)),
    "ghz_reference_semantics": (13, "reference", [("ghz",), ("h", 12), ("cnot", 12, 11)]),
    # math=fast: passes are subsequences of the queue (commuting gates reordered), swaps ride on them;
    # compared within 1e-12 instead of bit for bit (a 4th entry = extra engine options)
    "fast_reordered_brickwork": (14, "corrected", random_circuit_script(14, 8, seed=79), {"math": "fast", "tile_kernel": "ldg8"}),
    "fast_reordered_brickwork_10bit_tiles": (15, "corrected", random_circuit_script(15, 6, seed=80),
                                             {"math": "fast", "tile_kernel": "ldg8", "tile_bits": 10}),
    "fast_reordered_qft": (13, "corrected", [("h", q) for q in range(13)] + [("ry", 12, 0.6), ("qft",)],
                           {"math": "fast", "tile_kernel": "ldg8"}),
    "fast_in_order_brickwork": (14, "corrected", random_circuit_script(14, 5, seed=81),
                                {"math": "fast", "tile_kernel": "ldg8", "reorder": "off"}),
}


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_schedule_replays_to_the_oracle_state(world):
    from qcs_b200 import build
    build.build_all()
    env = dict(os.environ)
    env["MASTER_ADDR"] = "127.0.0.1"
    env["OMP_NUM_THREADS"] = "1"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + world),
           os.path.join(ROOT, "tests", "dist_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("ok ") == len(CASES), r.stdout[-3000:]
