"""How far math=fast lands from the bit-exact oracle (test-side measurement: it runs the CPU oracle as the checker): largest component error
relative to the largest amplitude, on dense states.  usage: python tests/fast_math_error.py [qubits ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from qcs_b200 import Circuit
from oracle import pyoracle as po

for n in [int(a) for a in sys.argv[1:]] or [16, 20, 24]:
    rng = np.random.default_rng(n)
    init = rng.normal(size=2 ** n) + 1j * rng.normal(size=2 ** n)
    init /= np.linalg.norm(init)
    for name, script in (("qft", [("qft",)]), ("random_d20", po.random_circuit_script(n, 20)),
                         ("random_d20+qft", po.random_circuit_script(n, 20) + [("qft",)])):
        orc = po.Oracle(n, "corrected"); orc.load_state(init); po.replay(orc, script); want = orc.state(); orc.close()
        row = []
        for reorder in ("off", "on"):
            c = Circuit(n, semantics="corrected", tile_kernel="ldg8", math="fast", reorder=reorder)
            c.load_state(init); po.replay(c, script); got = c.state(); c.close()
            row.append(np.abs(got - want).max() / np.abs(want).max())
        print(f"n={n:2d} {name:16s} max |delta| / max |a|: in-order {row[0]:.2e}  reordered {row[1]:.2e}", flush=True)
