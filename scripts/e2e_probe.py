import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if os.environ.get("WITH_TORCH"):
    import torch; torch.cuda.set_device(0); torch.cuda.synchronize(); x = torch.zeros(8, device="cuda")
from qcs_b200 import Circuit
n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
for rep in range(3):
    t = [time.perf_counter()]
    c = Circuit(n, semantics="corrected"); t.append(time.perf_counter())
    c.qft(); t.append(time.perf_counter())
    c.flush(); t.append(time.perf_counter())
    p = c.get_probability(12345); t.append(time.perf_counter())
    b = c.find_most_likely_state(); t.append(time.perf_counter())
    c.close(); t.append(time.perf_counter())
    names = ["create", "submit", "flush", "get_prob", "argmax", "destroy"]
    print(rep, {k: round((t[i + 1] - t[i]) * 1e3, 2) for i, k in enumerate(names)}, flush=True)
for pf in ():
    c = Circuit(n, semantics="corrected", pass_flops=pf)
    c.qft(); c.flush()
    c.reset_stats(); c.set_timing(True)
    t0 = time.perf_counter(); c.qft(); c.flush(); dt = time.perf_counter() - t0
    st = c.stats()
    print(f"pass_flops={pf}: {dt*1e3:.1f} ms, passes={st['passes']}, {n*(n+1)//2/dt:.0f} gates/s", flush=True)
    c.close()
