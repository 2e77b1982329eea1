"""Multi-GPU parity worker (one process per GPU, NCCL): run under torchrun by
tests/test_dist_gpu.py or by hand:
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/dist_gpu_worker.py
Every rank drives the same circuit through the public C API on a sharded engine and compares
its shard (and every global read) with the CPU oracle run on the unsharded state."""
import ctypes
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    from qcs_b200 import Circuit, _ffi
    _, C = _ffi.load()
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        buf = ctypes.create_string_buffer(128)
        assert C.qcs_cuda_dist_unique_id(buf) == 0, _ffi.last_error()
        uid.copy_(torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    assert C.qcs_cuda_dist_init(rank, world, bytes(uid.cpu().numpy().tobytes()), lr) == 0, _ffi.last_error()
    from tests.test_dist_gloo import CASES
    ok = True

    def check(cond, msg):
        nonlocal ok
        if not cond:
            ok = False
            print(f"[rank {rank}] FAIL {msg}", flush=True)

    for fusion, exchange in (("on", "p2p"), ("on", "nccl"), ("off", "p2p")):
        for name, case in CASES.items():
            n, sem, script = case[:3]
            extra = case[3] if len(case) > 3 else {}
            if extra.get("math") == "fast":
                # opt-in tolerance mode (reordered schedule, FMA arithmetic): 1e-12, fused kernels only
                if fusion != "on":
                    continue
                c = Circuit(n, semantics=sem, fusion=fusion, exchange=exchange, **extra)
                orc = po.Oracle(n, sem)
                po.replay(c, script); po.replay(orc, script)
                c.flush(); st = c.stats()
                first, count = c._shard()
                got = c.state(); full = orc.state(); want = full[first:first + count]
                check(np.all(np.abs(got - want) <= 1e-12 * np.abs(full).max()),
                      f"{name}/{exchange}: max |delta| {np.abs(got - want).max():.3e}")
                check(abs(c.get_probability(5) - orc.get_probability(5)) <= 1e-12, f"{name}: prob")
                if rank == 0:
                    print(f"done {name}/{fusion}/{exchange}: passes={st['passes']} remaps={st['remaps']} "
                          f"fused={st['fused_remaps']}", flush=True)
                c.close(); orc.close()
                continue
            c = Circuit(n, semantics=sem, fusion=fusion, exchange=exchange)
            orc = po.Oracle(n, sem)
            po.replay(c, script); po.replay(orc, script)
            c.flush(); st = c.stats()
            first, count = c._shard()
            got = c.state(); want = orc.state()[first:first + count]
            check(np.all(got == want), f"{name}/{fusion}: {int(np.sum(got != want))} shard amplitudes differ")
            check(c.find_most_likely_state() == orc.find_most_likely_state(), f"{name}: argmax")
            for idx in (0, 5, (1 << n) - 1, (1 << (n - 1)) + 3):
                check(c.get_probability(idx) == orc.get_probability(idx), f"{name}: prob {idx}")
            if sem == "reference":
                orc.normalize(); c.normalize()
            for q in (0, n - 1, n - 2):
                check(c.prob0(q) == po.Oracle.lib().orc_prob0(orc.h_, q), f"{name}: prob0({q})")
            po.srand(99); want_sh = orc.run_shots(3000); want_m = [orc.measure(n - 1), orc.measure(0), orc.measure(n - 2)]
            po.srand(99); got_sh = c.run_shots(3000); got_m = [c.measure(n - 1), c.measure(0), c.measure(n - 2)]
            check(np.array_equal(got_sh, want_sh), f"{name}: shots")
            check(got_m == want_m, f"{name}: measure {got_m} vs {want_m}")
            got = c.state(); want = orc.state()[first:first + count]
            check(np.all(got == want), f"{name}/{fusion}: post-measurement shard differs")
            if rank == 0:
                print(f"done {name}/{fusion}/{exchange}: passes={st['passes']} remaps={st['remaps']}", flush=True)
            c.close(); orc.close()
    # Larger shards (many tiles per rank): swaps ride on the stores of fused passes (SwapStore);
    # the same circuits with stand-alone swap kernels must give the same bits.
    multi_seen = [0]
    big = {"random_19": (19, po.random_circuit_script(19, 10, seed=5)),
           "qft_20": (20, [("x", 1), ("ry", 19, 0.3), ("h", 18), ("qft",)]),
           "h_top_down_19": (19, [("h", q) for q in reversed(range(19))] + [("cnot", 3, 18), ("cnot", 18, 17)])}
    for name, (n, script) in big.items():
        orc = po.Oracle(n, "corrected")
        po.replay(orc, script)
        # swaps on the stores of a pass through TMA bulk stores (default) / 16-byte stores, or stand-alone
        # and with up to 3 positions traded on one pass (an all-to-all among 2^k ranks) or one at a time
        # ... into the ranks' second shard buffers (remap_buffer=double: no handshake, what `auto` picks
        # while memory allows) or in place behind the per-tile handshake (remap_buffer=inplace)
        for fuse, store, rbuf in (("on", "bulk", "double"), ("on", "thread", "double"), ("on", "bulk", "inplace"),
                                  ("on", "thread", "inplace"), ("off", None, None)):
            for tile_bits in ((10, 11, 12) if fuse == "on" else (11,)):
                for remap_max in ((3, 1) if fuse == "on" and world >= 4 else (3,)):
                    c = Circuit(n, semantics="corrected", fuse_swaps=fuse, swap_store=store, tile_bits=tile_bits,
                                remap_max=remap_max, remap_buffer=rbuf)
                    po.replay(c, script); c.flush(); st = c.stats()
                    first, count = c._shard()
                    got = c.state(); want = orc.state()[first:first + count]
                    tag = f"{name}/fuse_swaps={fuse}/{store}/{rbuf}/t{tile_bits}/remap_max={remap_max}"
                    check(np.all(got == want), f"{tag}: {int(np.sum(got != want))} shard amplitudes differ")
                    check((st["fused_remaps"] > 0) == (fuse == "on"), f"{tag}: fused_remaps={st['fused_remaps']}")
                    if fuse == "on":
                        check(st["out_of_place_remaps"] == (st["fused_remaps"] if rbuf == "double" else 0),
                              f"{tag}: out_of_place_remaps={st['out_of_place_remaps']} of {st['fused_remaps']}")
                    if remap_max == 1:
                        check(st["multi_remaps"] == 0, f"{tag}: multi_remaps={st['multi_remaps']}")
                    multi_seen[0] += st["multi_remaps"]
                    if rank == 0:
                        print(f"done {tag}: passes={st['passes']} remaps={st['remaps']} fused={st['fused_remaps']} "
                              f"multi={st['multi_remaps']}", flush=True)
                    c.close()
        # the same circuit in math=fast: whole-queue reordered schedule, swaps riding on its passes
        full = orc.state()
        for fuse in ("on", "off"):
            c = Circuit(n, semantics="corrected", fuse_swaps=fuse, math="fast", tile_kernel="ldg8")
            po.replay(c, script); c.flush(); st = c.stats()
            first, count = c._shard()
            got = c.state(); want = full[first:first + count]
            check(np.all(np.abs(got - want) <= 1e-12 * np.abs(full).max()),
                  f"{name}/fast/fuse_swaps={fuse}: max |delta| {np.abs(got - want).max():.3e}")
            if rank == 0:
                print(f"done {name}/fast/fuse_swaps={fuse}: passes={st['passes']} remaps={st['remaps']} "
                      f"fused={st['fused_remaps']}", flush=True)
            c.close()
        orc.close()
    if world >= 4:
        check(multi_seen[0] > 0, "no pass traded more than one position pair on >= 4 ranks")
    # Grover across ranks: the diffusion mean is the reference's sequential sum, continued from rank
    # to rank in basis-index order -- every amplitude equal
    for sem in ("corrected", "reference"):
        for n in (13, 16):
            c = Circuit(n, semantics=sem); orc = po.Oracle(n, sem)
            c.grover_search(5000); orc.grover_search(5000)
            first, count = c._shard()
            got = c.state(); want = orc.state()[first:first + count]
            check(np.all(got == want), f"grover/{sem}/{n}: {int(np.sum(got != want))} shard amplitudes differ")
            check(c.find_most_likely_state() == orc.find_most_likely_state(), f"grover/{sem}/{n}: argmax")
            c.close(); orc.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DIST_GPU_CHECK", "PASS" if int(flag) else "FAIL", flush=True)
    C.qcs_cuda_dist_finalize()
    dist.destroy_process_group()
    sys.exit(0 if int(flag) else 1)


if __name__ == "__main__":
    main()
