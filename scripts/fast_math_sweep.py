"""math=exact vs math=fast on the workloads of BASELINE configs 3 and 5 (development aid; bench.py is the contract).
usage: fast_math_sweep.py [qubits] [tile_bits,...] [compute_bound_flops,...]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qcs_b200 import Circuit
from qcs_b200 import workloads as po

PEAK = 6550.1
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(p):
    PEAK = float(json.load(open(p))["hbm_gbs"])


def run(n, script, label, reps=3, **kw):
    kw.setdefault("tile_kernel", "ldg8")
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
    print(f"{label:28s} n={n} gates={st['gates_submitted'] // reps} passes={st['passes'] // reps} "
          f"segs={st['segments'] // reps} {ms:8.2f} ms  {st['gates_submitted'] / reps / ms * 1e3:9.0f} gates/s  "
          f"{gbs:6.0f} GB/s per pass = {gbs / PEAK * 100:3.0f}% HBM  flops/amp={st['pass_flops_per_amp'] / reps:.0f}",
          flush=True)
    c.close()


n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
tiles = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [11]
cbfs = [float(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [90.0]
for math in ("exact", "fast"):
    for tb in tiles:
        for cbf in cbfs:
            tag = f"{math}/t{tb}/cbf{cbf:g}"
            os.environ["QCS_CUDA_COMPUTE_BOUND_FLOPS"] = repr(cbf)
            run(n, [("qft",)], f"qft {tag}", math=math, tile_bits=tb, reorder="off")
            run(n, po.random_circuit_script(n, 8), f"random_d8 {tag}", reps=1, math=math, tile_bits=tb, reorder="off")
    run(n, [("rz", q, 0.1 * q) for q in range(n)], f"rz_all {math}", math=math)
    run(n, [("h", q) for q in range(n)], f"h_all {math}", math=math)
os.environ.pop("QCS_CUDA_COMPUTE_BOUND_FLOPS", None)
for tb in tiles:
    for segs in (4, 6, 8, 12):
        tag = f"fast+reorder/t{tb}/segs{segs}"
        run(n, po.random_circuit_script(n, 8), f"random_d8 {tag}", reps=1, math="fast", tile_bits=tb, reorder_segments=segs)
    run(n, [("qft",)], f"qft fast+reorder/t{tb}", math="fast", tile_bits=tb)

# 16 amplitudes per thread (tile_kernel=ldg) under math=fast: per-gate costs spread over twice the amplitudes
for tb in (11, 12):
    run(n, [("qft",)], f"qft fast/ldg16/t{tb}", math="fast", tile_kernel="ldg", tile_bits=tb)
    run(n, po.random_circuit_script(n, 8), f"random_d8 fast+reorder/ldg16/t{tb}", reps=1, math="fast",
        tile_kernel="ldg", tile_bits=tb)
    run(n, [("rz", q, 0.1 * q) for q in range(n)], f"rz_all fast/ldg16/t{tb}", math="fast", tile_kernel="ldg", tile_bits=tb)
