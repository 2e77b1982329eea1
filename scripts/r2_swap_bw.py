"""Swap-carrying passes: TMA bulk stores vs 16-byte stores, per pass, or one configuration given as option=value,... (development aid; run under torchrun):
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/r2_swap_bw.py [local_qubits]"""
import ctypes, math, os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qcs_b200 import Circuit, _ffi
from qcs_b200 import workloads as po

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
_, C = _ffi.load()
uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    buf = ctypes.create_string_buffer(128)
    assert C.qcs_cuda_dist_unique_id(buf) == 0
    uid.copy_(torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8))
dist.broadcast(uid, 0)
assert C.qcs_cuda_dist_init(rank, world, bytes(uid.cpu().numpy().tobytes()), lr) == 0, _ffi.last_error()
nl = int(sys.argv[1]) if len(sys.argv) > 1 else 30
n = nl + int(math.log2(world))


def run(script, label, reps=2, **kw):
    c = Circuit(n, semantics="corrected", **kw)
    c.set_timing(True)
    po.replay(c, script); c.flush()
    dist.barrier(); torch.cuda.synchronize()
    c.reset_stats()
    c.marker(0)
    for _ in range(reps):
        po.replay(c, script); c.flush()
    c.marker(1)
    ms = c.marker_elapsed_ms(0, 1) / reps
    t = torch.tensor([ms], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    st = c.stats()
    if rank == 0:
        fr = st["fused_remaps"]
        fms = st["fused_remap_pass_ms"] / fr if fr else 0.0
        gbs = st["fused_remap_bytes"] / fr / (fms * 1e-3) / 1e9 if fms else 0.0
        print(f"{label:34s} n={n} {float(t[0]):8.2f} ms passes={st['passes'] // reps} remaps={st['remaps'] // reps} fused={fr // reps} "
              f"carrying pass {fms:6.2f} ms = {gbs:4.0f} GB/s/dir; plain {(st['pass_ms'] - st['fused_remap_pass_ms']) / max(1, st['passes'] - fr):6.2f} ms "
              f"[{' '.join(f'{x:.1f}' for x in c.pass_times())}]", flush=True)
    c.close()


extra = dict(kv.split("=") for kv in sys.argv[2].split(",")) if len(sys.argv) > 2 and sys.argv[2] else {}
if extra:   # one configuration, bit-exact only: r2_swap_bw.py 30 compute_bound_flops=1e9
    tag = ",".join(f"{k}={v}" for k, v in extra.items())
    run([("qft",)], f"qft exact {tag}", reps=4, **extra)
    run(po.random_circuit_script(n, 8), f"random_d8 exact {tag}", reps=1, **extra)
else:
    for store in ("bulk", "thread"):
        for math_ in ("exact", "fast"):
            run([("qft",)], f"qft {math_}/{store}", math=math_, swap_store=store)
            run(po.random_circuit_script(n, 8), f"random_d8 {math_}/{store}", reps=1, math=math_, swap_store=store)
C.qcs_cuda_dist_finalize()
dist.destroy_process_group()
