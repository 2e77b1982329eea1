import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qcs_b200 import Circuit
from qcs_b200 import workloads as po
n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
for name, script in (("random_d8", po.random_circuit_script(n, 8)), ("qft", [("qft",)]), ("h_all", [("h", q) for q in range(n)])):
    for fl in (5, 4, 3, 2):
        c = Circuit(n, semantics="corrected", fixed_low=fl)
        c.set_timing(True)
        po.replay(c, script); c.flush(); c.reset_stats()
        t0 = time.perf_counter(); po.replay(c, script); c.flush(); dt = time.perf_counter() - t0
        st = c.stats()
        print(f"{name} n={n} fixed_low={fl}: {dt*1e3:.1f} ms passes={st['passes']} segs={st['segments']} "
              f"{st['gates_submitted']/dt:.0f} gates/s {st['pass_bytes']/st['pass_ms']/1e6:.0f} GB/s per pass", flush=True)
        c.close()
