"""Quick device timing of fused passes (development aid; bench.py is the contract)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qcs_b200 import Circuit
from qcs_b200 import workloads as po

def run(n, script, label, **kw):
    c = Circuit(n, semantics="corrected", **kw)
    c.set_timing(True)
    po.replay(c, script); c.flush()      # warm-up
    c.reset_stats()
    t0 = time.perf_counter()
    po.replay(c, script); c.flush()
    dt = time.perf_counter() - t0
    st = c.stats()
    gbs = st["pass_bytes"] / (st["pass_ms"] * 1e-3) / 1e9 if st["pass_ms"] else 0
    print(f"{label}: n={n} gates={st['gates_submitted']} passes={st['passes']} segs={st['segments']} "
          f"wall={dt*1e3:.1f} ms pass_ms={st['pass_ms']:.1f} -> {st['gates_submitted']/dt:.0f} gates/s, "
          f"{gbs:.0f} GB/s per pass ({gbs/6464.9*100:.0f}% of measured HBM)", flush=True)
    c.close()

n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
for tk in (sys.argv[2].split(",") if len(sys.argv) > 2 else ("tma", "tma16", "ldg")):
    print("== tile_kernel", tk, flush=True)
    run(n, [("h", 8)], "h_one", tile_kernel=tk)
    run(n, [("h", q) for q in range(5, 12)] * 2, "h_high_1seg", tile_kernel=tk)
    run(n, [("rz", q, 0.1 * q) for q in range(n)], "rz_all", tile_kernel=tk)
    run(n, [("h", q) for q in range(n)], "h_all", tile_kernel=tk)
    run(n, [("qft",)], "qft", tile_kernel=tk)
    run(n, [("qft",)], "qft_budget48", pass_flops=48, tile_kernel=tk)
    run(n, [("qft",)], "qft_budget200", pass_flops=200, tile_kernel=tk)
    run(n, po.random_circuit_script(n, 4), "random_d4", tile_kernel=tk)
sys.exit(0)
run(n, [("h", q) for q in range(5, 12)] * 2, "h_high_1seg")
run(n, [("rz", q, 0.1 * q) for q in range(n)], "rz_all")
run(n, [("h", q) for q in range(n)], "h_all")
run(n, [("qft",)], "qft")
run(n, [("qft",)], "qft_budget48", pass_flops=48)
run(n, [("qft",)], "qft_budget200", pass_flops=200)
run(n, po.random_circuit_script(n, 4), "random_d4")
run(n, [("h", q) for q in range(n)], "h_all_unfused", fusion="off")
