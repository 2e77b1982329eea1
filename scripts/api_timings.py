"""Wall-clock cost of each state-touching API call at a given width (development aid)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from qcs_b200 import Circuit
from qcs_b200 import workloads as po
n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
def t(label, f, reps=1):
    t0 = time.perf_counter()
    for _ in range(reps): r = f()
    dt = (time.perf_counter() - t0) / reps
    print(f"{label:40s} {dt*1e3:10.3f} ms", flush=True)
    return r
for sem in ("corrected", "reference"):
    print(f"== {n} qubits, semantics {sem}")
    c = t("qc_create (cold)", lambda: Circuit(n, semantics=sem)); c.close()
    c = t("qc_create (pooled buffers)", lambda: Circuit(n, semantics=sem))
    t("H on every qubit + flush", lambda: ([c.h(q) for q in range(n)], c.flush()))
    t("qc_get_probability", lambda: c.get_probability(5), 3)
    t("qc_find_most_likely_state", lambda: c.find_most_likely_state(), 3)
    t("prob0 (exact sequential sum)", lambda: c.prob0(3), 3)
    po.srand(1)
    t("qc_measure (1 qubit)", lambda: c.measure(n - 1))
    t("qc_measure (1 qubit)", lambda: c.measure(0))
    t("qc_run_shots_sparse 1e6", lambda: c.run_shots_sparse(1_000_000))
    if n <= 28:
        t("qc_run_shots 1e6 (dense histogram)", lambda: c.run_shots(1_000_000))
    t("qc_run (measure every qubit)", lambda: c.run())
    t("qc_destroy", lambda: c.close())
