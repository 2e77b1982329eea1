"""Worker of tests/test_dist_gloo.py (one process per rank, gloo, CPU only).

Each rank asks a PLAN-ONLY engine (qcs_cuda_dist_init_plan_only + dry run) for the schedule it
would execute as rank r of `world` -- physical gates and position swaps -- and replays it on a
numpy shard, trading half-shards with its partner over gloo exactly as dist.cu does over NCCL.
Rank 0 gathers the shards, undoes the engine's qubit layout and compares every amplitude with
the CPU oracle run on the unsharded state.  This checks the host-side logic of the N>1 path
(layout tracking, victim choice, swap pairing, global controls / diagonals) without a GPU.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402

GK_GENERIC, GK_REAL, GK_HSYM, GK_SWAP, GK_DIAG, GK_NOP = range(6)
GF_ROW0_ONLY, GF_D0_IDENT, GF_D1_IDENT = 1, 2, 4


def cmul(gr, gi, v):
    """c_mul(g, v) with the reference's operation order (each numpy op rounds once, no FMA)."""
    return ((gr * v.real) - (gi * v.imag)) + 1j * ((gr * v.imag) + (gi * v.real))


def apply_gate(shard, nl, rank, kind, flags, tpos, cpos, m):
    if kind == GK_NOP:
        return
    idx = np.arange(shard.size, dtype=np.int64)
    sel = np.ones(shard.size, dtype=bool)
    if cpos >= 0:
        if cpos >= nl:
            if not (rank >> (cpos - nl)) & 1:
                return
        else:
            sel &= ((idx >> cpos) & 1) == 1
    if kind == GK_DIAG:
        tbit = ((idx >> tpos) & 1) if tpos < nl else np.full(shard.size, (rank >> (tpos - nl)) & 1)
        for v, (dr, di), ident in ((0, (m[0], m[1]), GF_D0_IDENT), (1, (m[6], m[7]), GF_D1_IDENT)):
            if flags & ident:
                continue
            s = sel & (tbit == v)
            shard[s] = cmul(dr, di, shard[s])
        return
    assert tpos < nl, "pairing gate on a global position must have been swapped local first"
    i0 = idx[sel & (((idx >> tpos) & 1) == 0)]
    i1 = i0 | (1 << tpos)
    v0, v1 = shard[i0].copy(), shard[i1].copy()
    a, b = cmul(m[0], m[1], v0), cmul(m[2], m[3], v1)
    shard[i0] = (a.real + b.real) + 1j * (a.imag + b.imag)
    if not flags & GF_ROW0_ONLY:
        a, b = cmul(m[4], m[5], v0), cmul(m[6], m[7], v1)
        shard[i1] = (a.real + b.real) + 1j * (a.imag + b.imag)


def swap_positions(shard, nl, rank, lpos, gpos):
    gbit = gpos - nl
    partner = rank ^ (1 << gbit)
    mybit = (rank >> gbit) & 1
    idx = np.arange(shard.size, dtype=np.int64)
    leaving = ((idx >> lpos) & 1) == (1 - mybit)
    send = torch.from_numpy(np.ascontiguousarray(shard[leaving]).view(np.float64).copy())
    recv = torch.empty_like(send)
    if rank < partner:
        dist.send(send, partner); dist.recv(recv, partner)
    else:
        dist.recv(recv, partner); dist.send(send, partner)
    shard[leaving] = recv.numpy().view(np.complex128)


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    from qcs_b200 import Circuit, _ffi
    _, C = _ffi.load()
    assert C.qcs_cuda_dist_init_plan_only(rank, world) == 0, _ffi.last_error()
    from tests.test_dist_gloo import CASES
    failures = []
    for name, case in CASES.items():
        n, sem, script = case[:3]
        extra = case[3] if len(case) > 3 else {}
        nl = n - (world.bit_length() - 1)
        c = Circuit(n, dryrun=True, semantics=sem, **extra)
        po.replay(c, script)
        c.flush()
        trace, perm = c.trace(), c.layout()
        shard = np.zeros(1 << nl, dtype=np.complex128)
        if rank == 0:
            shard[0] = 1.0
        n_swaps = 0
        for ent in trace:
            if ent[0] == "swap":
                swap_positions(shard, nl, rank, ent[1], ent[2]); n_swaps += 1
            elif ent[0] == "swap_local":
                idx = np.arange(shard.size, dtype=np.int64)
                a, b = ent[1], ent[2]
                ba, bb = (idx >> a) & 1, (idx >> b) & 1
                src = (idx & ~((1 << a) | (1 << b))) | (bb << a) | (ba << b)
                shard[:] = shard[src]
            else:
                apply_gate(shard, nl, rank, *ent[1:])
        parts = [torch.empty(2 << nl, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(parts, torch.from_numpy(shard.view(np.float64).copy()))
        perms = [None] * world
        dist.all_gather_object(perms, (perm, n_swaps))
        if rank == 0:
            assert all(p == perms[0] for p in perms), f"{name}: ranks disagree on the layout {perms}"
            phys = np.concatenate([p.numpy() for p in parts]).view(np.complex128)
            logical = np.arange(1 << n, dtype=np.int64)
            where = np.zeros_like(logical)
            for q in range(n):
                where |= ((logical >> q) & 1) << perm[q]
            got = phys[where]
            orc = po.Oracle(n, sem)
            po.replay(orc, script)
            want = orc.state()
            if extra.get("math") == "fast":  # gates run in a different (commuting) order: rounding differs
                same = bool(np.all(np.abs(got - want) <= 1e-12 * np.abs(want).max()))
            else:
                same = bool(np.all(got == want))
            if not same:
                failures.append(f"{name}: {int(np.sum(got != want))} amplitudes differ, max |delta| "
                                f"{np.abs(got - want).max():.3e} (swaps={n_swaps})")
            else:
                print(f"ok {name}: n={n} world={world} swaps={n_swaps} layout={perm}", flush=True)
        c.close()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and failures:
        print("\n".join(failures), flush=True)
        sys.exit(1)


if __name__ == "__main__":
    main()
