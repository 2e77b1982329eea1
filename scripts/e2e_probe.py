"""Where the end-to-end time of one circuit goes, call by call (development aid):
    python scripts/e2e_probe.py [qubits]
create -> 465 gate calls (queued) -> first read (flush) -> second read -> destroy, for the read orders
and option settings that matter (lazy |0...0>, argmax fused into the last pass)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qcs_b200 import Circuit
n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
cases = [("argmax first", dict(), "argmax"), ("argmax first, fuse_argmax=off", dict(fuse_argmax="off"), "argmax"),
         ("probability first", dict(), "prob"), ("flush first", dict(), "flush"),
         ("argmax first, lazy_init=off", dict(lazy_init="off"), "argmax")]
for label, kw, first in cases:
    for rep in range(3):
        t = [time.perf_counter()]
        c = Circuit(n, semantics="corrected", **kw); t.append(time.perf_counter())
        c.qft(); t.append(time.perf_counter())
        if first == "argmax":
            c.find_most_likely_state(); t.append(time.perf_counter())
            c.get_probability(12345); t.append(time.perf_counter())
        elif first == "prob":
            c.get_probability(12345); t.append(time.perf_counter())
            c.find_most_likely_state(); t.append(time.perf_counter())
        else:
            c.flush(); t.append(time.perf_counter())
            c.find_most_likely_state(); t.append(time.perf_counter())
        c.close(); t.append(time.perf_counter())
        names = ["create", "submit", "first_read", "second_read", "destroy"]
        if True:
            print(f"{label:32s}", {k: round((t[i + 1] - t[i]) * 1e3, 2) for i, k in enumerate(names)},
                  "total", round((t[-1] - t[0]) * 1e3, 2), flush=True)
