"""Round-2 kernel sweep: every plain-load tile kernel shape (8 or 16 amplitudes per thread, 10/11/12-bit
tiles) on the workloads of BASELINE configs 3 and 5, bit-exact and math=fast (development aid;
bench.py is the contract).   usage: r2_sweep.py [qubits] [what,...] [math,...] [shape,...] [option=value,...]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qcs_b200 import Circuit
from qcs_b200 import workloads as po

PEAK = 6550.1
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(p):
    PEAK = float(json.load(open(p))["hbm_gbs"])


def run(n, script, label, reps=3, **kw):
    c = Circuit(n, semantics="corrected", **kw)
    c.set_timing(True)
    po.replay(c, script); c.flush()
    c.reset_stats()
    c.marker(0)
    for _ in range(reps):
        po.replay(c, script); c.flush()
    c.marker(1)
    ms = c.marker_elapsed_ms(0, 1) / reps
    st = c.stats()
    gbs = st["pass_bytes"] / (st["pass_ms"] * 1e-3) / 1e9 if st["pass_ms"] else 0
    per_pass = ""
    if hasattr(c, "pass_times"):
        per_pass = " [" + " ".join(f"{t:.2f}" for t in c.pass_times()) + "]"
    print(f"{label:34s} n={n} passes={st['passes'] // reps:3d} segs={st['segments'] // reps:3d} {ms:8.2f} ms "
          f"{gbs:6.0f} GB/s/pass = {gbs / PEAK * 100:3.0f}% HBM  flops/amp={st['pass_flops_per_amp'] / reps:.0f}{per_pass}",
          flush=True)
    c.close()
    return ms


n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
what = sys.argv[2].split(",") if len(sys.argv) > 2 else ["qft", "random", "rz", "h"]
maths = sys.argv[3].split(",") if len(sys.argv) > 3 and sys.argv[3] else ["exact", "fast"]
shapes = [("ldg8", 10), ("ldg8", 11), ("ldg", 10), ("ldg", 11), ("ldg", 12)]
if len(sys.argv) > 4 and sys.argv[4]:
    shapes = [(s.split(":")[0], int(s.split(":")[1]) if ":" in s else None) for s in sys.argv[4].split(",")]
extra = dict(kv.split("=") for kv in sys.argv[5].split(",")) if len(sys.argv) > 5 and sys.argv[5] else {}
for math in maths:
    for tk, tb in shapes:
        tag = f"{math}/{tk}/t{tb}" + "".join(f"/{k}={v}" for k, v in extra.items())
        kw = dict(math=math, tile_kernel=tk, tile_bits=tb, **extra)   # tile_bits None = the engine's default
        if "qft" in what:
            run(n, [("qft",)], f"qft {tag}", **kw)
        if "random" in what:
            run(n, po.random_circuit_script(n, 8), f"random_d8 {tag}", reps=1, **kw)
        if "rz" in what:
            run(n, [("rz", q, 0.1 * q) for q in range(n)], f"rz_all {tag}", **kw)
        if "h" in what:
            run(n, [("h", q) for q in range(n)], f"h_all {tag}", **kw)
