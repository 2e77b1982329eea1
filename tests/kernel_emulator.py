"""CPU model of the fused tile kernel, driven by the SAME tables the GPU kernel reads.

Test infrastructure only (no GPU, no product code path).  tests/plan_emulator.py checks WHAT a plan
computes from the gate records' documented fields (kind, tpos, cpos); this module checks that the
kernel would compute it: it moves amplitudes the way fused_body.inc does and decides the way the
generated interpreter (gen_fused_lists.py) does, reading only what they read --

  * tile_run            tile number -> tile origin              (tile_origin_of)
  * seg[].gb_lut, goff  thread id / register -> global address  (thread_phys_bits, gaddr)
  * seg[].tb_lut, toff  thread id / register -> tile index      (thread_tile_bits, tile_slot),
                        i.e. the transposition between segments; stoff must be swizzle(toff)
  * gate[].op           case label -> symbolic op (inverse of fused_ops.h): which registers pair,
                        which register bit is a control, which half of a diagonal applies
  * gate[].csel / tsel  thread-level control / diagonal target, tested on xfull
  * fan / uniform-fan headers (entry count, first control or per-entry cpos, target register),
    ufan_header / pad[0] for the prologue of the math=fast kernels

so a wrong role table, register index, label, csel or tsel shows up as a wrong amplitude here,
without a GPU.  Arithmetic is numpy complex128 (1e-12 bar); bit-exactness is the GPU tests' claim.
"""
import os
import re

import numpy as np

from tests.plan_emulator import GF_FAN_HEADER, GF_ROW0_ONLY  # noqa: F401

_OPS_H = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                      "qcs_b200", "csrc", "cuda", "fused_ops.h")
SYM_TCTL, SYM_NOP = 0x100, 255


def _label_tables():
    text = open(_OPS_H).read()
    out = {}
    for R in (3, 4):
        body = re.search(r"QCS%d_CASE_LABEL\[512\] = \{(.*?)\};" % R, text, re.S).group(1)
        table = [int(x) for x in re.findall(r"\d+", body)]
        assert len(table) == 512
        nop = table[SYM_NOP]
        inv = {}
        for sym, lab in enumerate(table):
            if lab != nop:
                assert lab not in inv, "a case label must belong to one symbolic op"
                inv[lab] = sym
        inv[nop] = SYM_NOP
        out[R] = inv
    return out


_INV = None


def _swz(slot):
    return slot ^ (((slot >> 3) ^ (slot >> 6) ^ (slot >> 9)) & 7)


def _cmul_diag(regs, sel, m_re, m_im):
    regs[sel] = regs[sel] * complex(m_re, m_im)


def run_plan(passes, n_qubits, fast, state=None):
    """Runs the passes the way the kernel would; returns the state vector."""
    if state is None:
        state = np.zeros(1 << n_qubits, dtype=np.complex128)
        state[0] = 1.0
    for p in passes:
        addr, regs = run_pass(p, state, n_qubits, fast)
        state = state.copy()
        state[addr] = regs
    return state


def run_sharded(plans, n_qubits, world, fast, shards=None):
    """The same for `world` ranks.  plans[rank] = [(PassParams, swap)], swap = None or a tuple of up to
    three (lpos, gpos) pairs: the pass trades local positions lpos for global positions gpos on its
    stores, all at once (fused_body.inc OP_STG_SWAP); an entry with no segments is a stand-alone swap.
    Returns the shards."""
    nl = n_qubits - (world.bit_length() - 1)
    if shards is None:
        shards = [np.zeros(1 << nl, dtype=np.complex128) for _ in range(world)]
        shards[0][0] = 1.0
    n_entries = len(plans[0])
    assert all(len(pl) == n_entries for pl in plans), "ranks must execute the same schedule"
    for k in range(n_entries):
        swap = plans[0][k][1]
        assert all(pl[k][1] == swap for pl in plans), "ranks disagree on a swap"
        outs = []
        for rank in range(world):
            p = plans[rank][k][0]
            if p.n_segments == 0:
                idx = np.arange(1 << nl, dtype=np.int64)
                outs.append((idx, shards[rank].copy()))
            else:
                assert p.shard_base == rank << nl
                outs.append(run_pass(p, shards[rank], nl, fast))
        new = [np.full(1 << nl, np.nan + 0j) for _ in range(world)]
        for rank in range(world):
            addr, regs = outs[rank]
            addr, regs = addr.ravel(), regs.ravel()
            if swap is None:
                new[rank][addr] = regs
                continue
            # the kernel's store rule (fused_body.inc OP_STG_SWAP, kernels.h SwapStore): the element's bits
            # at the traded local positions name the destination rank's traded rank bits; there it sits
            # at its own index with those bits replaced by MY traded rank bits
            dest_rank = np.full(addr.shape, rank, dtype=np.int64)
            dest_addr = addr.astype(np.int64).copy()
            for lpos, gpos in swap:
                gbit = gpos - nl
                b = (addr.astype(np.int64) >> lpos) & 1
                mine = (rank >> gbit) & 1
                dest_rank = (dest_rank & ~(1 << gbit)) | (b << gbit)
                dest_addr = (dest_addr & ~(1 << lpos)) | (mine << lpos)
            for r2 in range(world):
                sel = dest_rank == r2
                new[r2][dest_addr[sel]] = regs[sel]
        assert not any(np.isnan(s).any() for s in new), "every amplitude of every shard is written once"
        shards = new
    return shards


def run_pass(p, state, n_qubits, fast, tables=None):
    """One fused pass over a shard of 2^n_qubits amplitudes: returns (store addresses, values)."""
    global _INV
    if _INV is None:
        _INV = _label_tables()
    if tables is None and p.n_thread_tables:
        tables = p.thread_tables.reshape(p.n_thread_tables, -1)   # plan_emulator.read_plan attaches them
    if True:
        T, R = p.tile_bits, p.reg_bits
        CT, NREG, n_tiles = 1 << (T - R), 1 << R, 1 << (n_qubits - T)
        inv = _INV[R]
        tiles = np.arange(n_tiles, dtype=np.uint64)
        origin = np.zeros(n_tiles, dtype=np.uint64)
        for k in range(p.n_tile_runs):
            run = p.tile_run[k]
            origin |= ((tiles >> np.uint64(run.src)) & np.uint64((1 << run.len) - 1)) << np.uint64(run.dst)
        tile_full = np.uint64(p.shard_base) | origin                       # [tile]
        tid = np.arange(CT)

        def thread_lut(lut, dtype):
            a = np.array([[lut[g][v] for v in range(8)] for g in range(3)], dtype=dtype)
            return a[0][tid & 7] | a[1][(tid >> 3) & 7] | a[2][(tid >> 6) & 7]

        def reg_or(offs, dtype):
            return np.array([np.bitwise_or.reduce([dtype(offs[k]) for k in range(R) if (r >> k) & 1] or [dtype(0)])
                             for r in range(NREG)], dtype=dtype)

        def addresses(which):
            seg = p.seg[0 if which == 0 else p.n_segments - 1]
            gbase = thread_lut(seg.gb_lut, np.uint64)
            goff = reg_or(p.goff[which], np.uint64)
            return origin[:, None, None] | gbase[None, :, None] | goff[None, None, :]

        addr = addresses(0)
        assert np.array_equal(np.sort(addr.ravel()), np.arange(1 << n_qubits, dtype=np.uint64)), "loads must cover the shard once"
        regs = state[addr.astype(np.int64)]                                # [tile, thread, register]
        tile_data = np.zeros((n_tiles, 1 << T), dtype=np.complex128)
        tidx_prev = None
        tile_pos = [p.tile_pos[b] for b in range(T)]
        for s in range(p.n_segments):
            seg = p.seg[s]
            tbase = thread_lut(seg.tb_lut, np.int64)
            toff = reg_or(seg.toff, np.int64)
            assert all(seg.stoff[k] == _swz(seg.toff[k]) for k in range(R)), "swizzled offsets out of sync"
            tidx = tbase[:, None] | toff[None, :]                          # tile index of [thread, register]
            assert np.array_equal(np.sort(tidx.ravel()), np.arange(1 << T)), "roles must cover the tile once"
            # role table consistency: the role bits are what tb_lut / toff were built from
            thread_roles = T - R
            for k in range(R):
                assert seg.toff[k] == 1 << seg.role_tilebit[thread_roles + k]
            if s > 0:  # transposition through shared memory: old roles out, new roles in
                tile_data[:, tidx_prev.ravel()] = regs.reshape(n_tiles, -1)
                regs = tile_data[:, tidx.ravel()].reshape(n_tiles, CT, NREG)
            tidx_prev = tidx
            # the thread's physical index bits must be its tile bits deposited at the tile positions
            phys = thread_lut(seg.gb_lut, np.uint64)
            dep = np.zeros(CT, dtype=np.uint64)
            for b in range(T):
                dep |= ((tbase.astype(np.uint64) >> np.uint64(b)) & np.uint64(1)) << np.uint64(tile_pos[b])
            assert np.array_equal(phys, dep), "gb_lut disagrees with tb_lut"
            xfull = tile_full[:, None] | phys[None, :]                    # [tile, thread]
            # physical position of each register role bit (for fans whose target is a register bit etc.)
            rbit = np.arange(NREG)

            def xbit(pos):
                return ((xfull >> np.uint64(pos)) & np.uint64(1)).astype(bool)   # [tile, thread]

            def apply_factor(sel_thread, sel_reg, factor):
                """regs[tile, thread, r] *= factor where sel_thread[tile, thread] and sel_reg[r]"""
                mask = sel_thread[:, :, None] & sel_reg[None, None, :]
                if np.ndim(factor) == 0:
                    regs[mask] *= factor
                else:
                    regs[mask] *= np.broadcast_to(factor[:, :, None], regs.shape)[mask]

            def fan_tables(h, K):
                tabs = []
                for i in range((K + 3) // 4):
                    t = np.full(16, np.nan, dtype=np.complex128)
                    for j in range(16):
                        rec = h + 1 + 4 * i + (j >> 2)
                        if rec <= h + K:
                            t[j] = complex(p.gate[rec].m[2 * (j & 3)], p.gate[rec].m[2 * (j & 3) + 1])
                    tabs.append(t)
                return tabs

            def ufan_factor(h):
                """the kernel prologue's work for the uniform fan at record h: one factor per CTA from the tile's origin"""
                g = p.gate[h]
                K = g.tsel
                if g.pad[0]:
                    mask = (tile_full >> np.uint64(g.cpos)) & np.uint64((1 << K) - 1)
                else:
                    mask = np.zeros(n_tiles, dtype=np.uint64)
                    for k in range(K):
                        mask |= ((tile_full >> np.uint64(p.gate[h + 1 + k].cpos)) & np.uint64(1)) << np.uint64(k)
                factor = np.ones(n_tiles, dtype=np.complex128)
                for i, tab in enumerate(fan_tables(h, K)):
                    nib = ((mask >> np.uint64(4 * i)) & np.uint64(15)).astype(np.int64)
                    factor *= np.where(nib > 0, tab[nib], 1.0)
                return factor

            gi = seg.gate_begin
            pending = np.ones((n_tiles, CT), dtype=np.complex128)          # math=fast: (fr, fi)
            while gi < seg.gate_end:
                g = p.gate[gi]
                sym = inv[g.op]
                gi += 1
                if sym == SYM_NOP:
                    continue
                active = np.ones((n_tiles, CT), dtype=bool)
                if sym & SYM_TCTL:
                    active = xbit(g.csel)
                base = sym & 0xFF
                m = [complex(g.m[2 * k], g.m[2 * k + 1]) for k in range(4)]
                if base < 160:                                            # pairing op
                    c = base % 5 - 1
                    t = (base // 5) % 4
                    row0 = (base // 20) % 2
                    a_sel = ((rbit >> t) & 1) == 0
                    if c >= 0:
                        a_sel &= ((rbit >> c) & 1) == 1
                    ra = rbit[a_sel]
                    rb = ra | (1 << t)
                    v0, v1 = regs[:, :, ra].copy(), regs[:, :, rb].copy()
                    n0 = m[0] * v0 + m[1] * v1
                    n1 = v1 if row0 else m[2] * v0 + m[3] * v1
                    regs[:, :, ra] = np.where(active[:, :, None], n0, v0)
                    regs[:, :, rb] = np.where(active[:, :, None], n1, v1)
                elif base < 175:                                          # diagonal, target not a register bit
                    c = (base - 160) // 3 - 1
                    halves = (base - 160) % 3 + 1
                    tb = xbit(g.tsel)
                    reg_sel = np.ones(NREG, dtype=bool) if c < 0 else ((rbit >> c) & 1) == 1
                    if halves & 1:
                        if fast and c < 0:
                            pending[active & ~tb] *= m[0]
                        else:
                            apply_factor(active & ~tb, reg_sel, m[0])
                    if halves & 2:
                        if fast and c < 0:
                            pending[active & tb] *= m[3]
                        else:
                            apply_factor(active & tb, reg_sel, m[3])
                elif 235 <= base < 240:                                   # thread-table fan (math=fast)
                    assert fast and tables is not None, "thread-table fans exist under math=fast only"
                    t = base - 236
                    h = gi - 1
                    factor = np.broadcast_to(tables[g.tsel][None, :], (n_tiles, CT)).copy()
                    if g.csel != 0xFF:
                        factor *= ufan_factor(p.ufan_header[g.csel])[:, None]
                    if t >= 0:
                        apply_factor(np.ones((n_tiles, CT), dtype=bool), ((rbit >> t) & 1) == 1, factor)
                    else:
                        tb = xbit(g.tpos)
                        pending[tb] *= factor[tb]
                    gi += g.pad[1]
                elif base < 240:                                          # diagonal, target = register bit t
                    k = (base - 175) // 3
                    halves = (base - 175) % 3 + 1
                    t, c = k // 5, k % 5 - 1
                    ctl = np.ones(NREG, dtype=bool) if c < 0 else ((rbit >> c) & 1) == 1
                    if halves & 1:
                        apply_factor(active, ctl & (((rbit >> t) & 1) == 0), m[0])
                    if halves & 2:
                        apply_factor(active, ctl & (((rbit >> t) & 1) == 1), m[3])
                elif base < 248:                                          # fan header
                    t = base - 241
                    K, c0 = g.tsel, g.csel
                    h = gi - 1
                    part = [xbit(c0 + k) if c0 != 0xFF else xbit(p.gate[h + 1 + k].cpos) for k in range(K)]
                    tgt_thread = np.ones((n_tiles, CT), dtype=bool) if t >= 0 else xbit(g.tpos)
                    tgt_reg = ((rbit >> t) & 1) == 1 if t >= 0 else np.ones(NREG, dtype=bool)
                    if not fast:
                        for k in range(K):
                            e = p.gate[h + 1 + k]
                            apply_factor(part[k] & tgt_thread, tgt_reg, complex(e.m[6], e.m[7]))
                    else:
                        factor = np.ones((n_tiles, CT), dtype=np.complex128)
                        for i, tab in enumerate(fan_tables(h, K)):
                            nib = np.zeros((n_tiles, CT), dtype=np.int64)
                            for b in range(4):
                                if 4 * i + b < K:
                                    nib |= (part[4 * i + b] & tgt_thread).astype(np.int64) << b
                            factor *= tab[nib]
                        if t >= 0:
                            apply_factor(tgt_thread, tgt_reg, factor)
                        else:
                            pending *= factor
                    gi += K
                elif base < 255:                                          # uniform fan (math=fast)
                    assert fast, "uniform fans exist under math=fast only"
                    t = base - 249
                    K, slot = g.tsel, g.csel
                    h = gi - 1
                    assert p.ufan_header[slot] == h and slot < p.n_ufans
                    f2 = np.broadcast_to(ufan_factor(h)[:, None], (n_tiles, CT))
                    if t >= 0:
                        apply_factor(np.ones((n_tiles, CT), dtype=bool), ((rbit >> t) & 1) == 1, f2)
                    else:
                        tb = xbit(g.tpos)
                        pending[tb] *= f2[tb]
                    gi += K
                else:
                    raise AssertionError(f"unknown symbolic op {sym}")
            if fast:
                regs *= pending[:, :, None]                                # QCS3F_FLUSH_ASM
        addr = addresses(1)
        assert np.array_equal(np.sort(addr.ravel()), np.arange(1 << n_qubits, dtype=np.uint64)), "stores must cover the shard once"
        return addr.astype(np.int64), regs
