"""Runs one small workload for ncu captures: python scripts/prof_case.py <case> <n> [tile_kernel] [math]."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qcs_b200 import Circuit
from qcs_b200 import workloads as po
case = sys.argv[1]; n = int(sys.argv[2]); tk = sys.argv[3] if len(sys.argv) > 3 else "tma"
scripts = {
    "rz_all": [("rz", q, 0.1 * q) for q in range(n)],
    "h_high": [("h", q) for q in range(5, 12)] * 2,
    "h_one": [("h", 8)],
    "qft": [("qft",)],
    "random": po.random_circuit_script(n, 2),
}
math = sys.argv[4] if len(sys.argv) > 4 else "exact"
c = Circuit(n, semantics="corrected", tile_kernel=tk, math=math)
for _ in range(2):
    po.replay(c, scripts[case]); c.flush()
print(c.stats())
