"""31-qubit depth-40 random circuit under math=fast on one GPU (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qcs_b200 import Circuit
from qcs_b200 import workloads as po
for n, d in ((31, 40), (30, 8)):
    script = po.random_circuit_script(n, d)
    c = Circuit(n, semantics="corrected", tile_kernel="ldg8", math="fast")
    c.set_timing(True)
    po.replay(c, script[: len(script) // 8]); c.flush()
    c.reset_stats(); c.marker(0)
    po.replay(c, script); c.flush()
    c.marker(1)
    st = c.stats()
    print(f"n={n} depth={d} passes={st['passes']} segs={st['segments']} {c.marker_elapsed_ms(0, 1):.1f} ms", flush=True)
    c.close()
