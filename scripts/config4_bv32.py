"""BASELINE config 4: Bernstein-Vazirani on n qubits (default 32) + qc_run_shots with 10^6 shots,
state sharded over the ranks of a torchrun launch (2 or 4 GPUs):

  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/config4_bv32.py [n] [shots]

Checks what the width allows (the CPU reference cannot run >= 31 qubits): under
corrected semantics every shot must land on `secret` or `secret + 2^(n-1)` (the ancilla ends in
|->), roughly half each.  Prints one JSON line with timings."""
import ctypes, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from qcs_b200 import workloads as po

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
shots = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
from qcs_b200 import Circuit, _ffi
H, C = _ffi.load()
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        buf = ctypes.create_string_buffer(128)
        assert C.qcs_cuda_dist_unique_id(buf) == 0
        uid.copy_(torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    assert C.qcs_cuda_dist_init(rank, world, bytes(uid.cpu().numpy().tobytes()), lr) == 0, _ffi.last_error()
secret = 0x2AAAAAAA & ((1 << (n - 1)) - 1)
t0 = time.perf_counter()
c = Circuit(n, semantics="corrected")
c.set_timing(True)
t1 = time.perf_counter()
c.bv(secret); c.flush()
t2 = time.perf_counter()
po.srand(32)
idx = c.run_shots_sparse(shots)          # every rank draws the same rand() stream
t3 = time.perf_counter()
dense_s = None
if rank == 0 and os.environ.get("QCS_DENSE_HISTOGRAM", "1") == "1":
    po.srand(32)
    d0 = time.perf_counter()
    dense = c.run_shots(shots)           # the qcs.h contract: int[2^n] zeroed and filled by the callee
    dense_s = time.perf_counter() - d0
    hist_ok = bool(dense[secret] + dense[secret + (1 << (n - 1))] == shots)
    assert np.array_equal(np.bincount(idx[idx >= 0] >> (n - 1), minlength=2), [dense[secret], dense[secret + (1 << (n - 1))]])
    del dense
else:
    hist_ok = None
    if world > 1:
        po.srand(32); c.run_shots_sparse(shots)   # keep the ranks' collective calls aligned
st = c.stats()
c.close()
ok = bool(np.all((idx == secret) | (idx == secret + (1 << (n - 1)))))
ones = int(np.sum(idx == secret + (1 << (n - 1))))
if rank == 0:
    print(json.dumps({"config": f"{n}-qubit Bernstein-Vazirani + {shots} shots, {world} GPU(s)", "all_shots_on_secret": ok,
                      "dense_histogram_agrees": hist_ok, "ancilla_one_fraction": ones / shots,
                      "create_s": t1 - t0, "bv_s": t2 - t1, "shots_sparse_s": t3 - t2, "shots_dense_s": dense_s,
                      "passes": st["passes"], "remaps": st["remaps"], "fused_remaps": st["fused_remaps"]}), flush=True)
if world > 1:
    C.qcs_cuda_dist_finalize(); dist.destroy_process_group()
sys.exit(0 if ok else 1)
