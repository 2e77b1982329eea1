import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from qcs_b200 import Circuit
from oracle import pyoracle as po
n = 14; tk = sys.argv[1] if len(sys.argv) > 1 else "tma"
script = po.random_circuit_script(n, 3, seed=3)
orc = po.Oracle(n, "corrected"); po.replay(orc, script)
c = Circuit(n, semantics="corrected", tile_kernel=tk); po.replay(c, script)
print("tma probe equal:", bool(np.all(c.state() == orc.state())), c.stats()["passes"], flush=True)
