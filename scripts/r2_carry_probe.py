"""A compute-heavy pass that carries a remap, on 2 GPUs: the first pass of a QFT (targets 0..9, controls from
everywhere) followed by one gate on the top qubit, so that the only pass is also the carrying one.  Compared
with the same pass without the trailing gate (development aid; run under torchrun):
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/r2_carry_probe.py [local_qubits]
Environment: QCS_CUDA_STAGGER (kernels.h PassExtras)."""
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


def heavy(depth):
    s = []
    for t in range(depth):
        s.append(("h", t))
        for c in range(t + 1, n):
            s.append(("cphase", c, t, math.pi / (1 << (c - t))))
    return s


def run(script, label, reps=3, **kw):
    # a fresh circuit per repetition: the layout a remap leaves behind would make the next one unnecessary
    for rep in range(reps):
        c = Circuit(n, semantics="corrected", lazy_init="off", **kw)
        c.set_timing(True)
        dist.barrier(); torch.cuda.synchronize()
        po.replay(c, script); c.flush()
        torch.cuda.synchronize()
        st = c.stats()
        times = c.pass_times()
        c.close()
    t = torch.tensor(times, dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"{label:44s} n={n} passes={st['passes']} fused={st['fused_remaps']} "
              f"oop={st['out_of_place_remaps']} flops/amp={st['pass_flops_per_amp']:.0f} "
              f"[{' '.join(f'{x:.1f}' for x in t.tolist())}] (max over ranks)", flush=True)


for depth in (10, 4, 1):
    base = heavy(depth)
    run(base, f"depth {depth}: plain")
    for rb in ("double", "inplace"):
        for store in ("bulk", "thread"):
            run(base + [("h", n - 1)], f"depth {depth}: + h(top) {rb}/{store}", remap_buffer=rb, swap_store=store)
C.qcs_cuda_dist_finalize()
dist.destroy_process_group()
