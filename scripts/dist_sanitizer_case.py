"""One small sharded circuit with remaps fused into passes, for compute-sanitizer (memcheck / racecheck)
on a 2-GPU box; every rank runs under its own sanitizer:
  python -m torch.distributed.run --no-python --nproc-per-node 2 --master-addr 127.0.0.1 \
      compute-sanitizer --tool memcheck python scripts/dist_sanitizer_case.py"""
import ctypes, os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qcs_b200 import Circuit, _ffi
from qcs_b200 import workloads as po

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("gloo")   # rendezvous only: NCCL inside torch is not what is under test
_, C = _ffi.load()
buf = ctypes.create_string_buffer(128)
if rank == 0:
    assert C.qcs_cuda_dist_unique_id(buf) == 0, _ffi.last_error()
obj = [buf.raw]
dist.broadcast_object_list(obj, 0)
assert C.qcs_cuda_dist_init(rank, world, obj[0], lr) == 0, _ffi.last_error()
n = 15
for store in ("bulk", "thread"):
    c = Circuit(n, semantics="corrected", swap_store=store)
    po.replay(c, po.random_circuit_script(n, 6, seed=3) + [("qft",)])
    c.flush()
    st = c.stats()
    p = sum(c.get_probability(i) for i in (0, 1, 77, (1 << n) - 1))
    if rank == 0:
        print(f"swap_store={store}: passes={st['passes']} remaps={st['remaps']} fused={st['fused_remaps']} "
              f"multi={st['multi_remaps']} p4={p:.6f}", flush=True)
    c.close()
C.qcs_cuda_dist_finalize()
dist.destroy_process_group()
