"""Grover iteration rate (phase flip + diffusion with the exact sequential-sum replay), both semantics
and math=fast (tree sums); development aid.  usage: r2_grover_timing.py [qubits] [iterations]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qcs_b200 import Circuit

n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 400
sol = 0xABCDE % (1 << n)
for label, kw in (("corrected", dict(semantics="corrected")), ("reference", dict(semantics="reference")),
                  ("corrected math=fast", dict(semantics="corrected", math="fast"))):
    c = Circuit(n, **kw)
    for q in range(n):
        c.h(q)
    for _ in range(20):
        c.phase_flip(sol); c.diffusion()
    c.flush()
    t0 = time.perf_counter()
    for _ in range(iters):
        c.phase_flip(sol); c.diffusion()
    c.flush()
    dt = time.perf_counter() - t0
    print(f"{label:22s} n={n}: {iters / dt:8.1f} iterations/s, {1e3 * dt / iters:7.3f} ms per iteration, "
          f"p(sol)={c.get_probability(sol):.6g}", flush=True)
    c.close()
t0 = time.perf_counter()
c = Circuit(n, semantics="corrected"); c.grover_search(sol); p = c.get_probability(sol)
print(f"qc_grover_search({n} qubits) corrected: {time.perf_counter() - t0:.3f} s, p(sol)={p:.9f}", flush=True)
c.close()
