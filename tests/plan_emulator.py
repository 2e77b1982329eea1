"""CPU interpreter of the pass descriptors a flush hands to the GPU (PassParams, common.h).

Test infrastructure only.  It reads the raw kernel-parameter blocks of a plan-only (dry-run)
engine through qcs_cuda_last_plan_raw and applies them to a numpy state vector, record by record:
what the fused tile kernel would compute, minus the tiling (role assignment, transpositions and
addressing are the kernel's business and are covered by the GPU parity tests).  It checks the
part of a plan that can be wrong without a GPU noticing here: which records a pass carries, in
which order, with which matrices -- and, under math=fast, the product tables the planner lays
over the entries of a controlled-phase fan (planner.cpp write_fan_tables), walked exactly the way
the kernel walks them (four entries per lookup).

Arithmetic is numpy complex128, so results agree with the oracle to rounding (1e-12 bar), not bit
for bit; bit-exactness is the GPU tests' claim.
"""
import ctypes

import numpy as np

TILE_BITS, MAX_REG_BITS, MAX_SEGMENTS, MAX_GATES, MAX_FANS = 12, 4, 12, 240, 48
GK_GENERIC, GK_REAL, GK_HSYM, GK_SWAP, GK_DIAG, GK_NOP = range(6)
GF_ROW0_ONLY, GF_TFAN_HEADER, GF_FAN_HEADER = 1, 0x40, 0x80


class DGate(ctypes.Structure):
    _fields_ = [("m", ctypes.c_double * 8), ("op", ctypes.c_uint16), ("csel", ctypes.c_uint8),
                ("tsel", ctypes.c_uint8), ("kind", ctypes.c_uint8), ("flags", ctypes.c_uint8),
                ("tpos", ctypes.c_int8), ("cpos", ctypes.c_int8), ("treg_creg", ctypes.c_uint8),
                ("pad", ctypes.c_uint8 * 7)]


class DSegment(ctypes.Structure):
    _fields_ = [("role_tilebit", ctypes.c_uint8 * TILE_BITS), ("gate_begin", ctypes.c_uint16),
                ("gate_end", ctypes.c_uint16), ("tb_lut", (ctypes.c_uint16 * 8) * 3),
                ("gb_lut", (ctypes.c_uint64 * 8) * 3), ("toff", ctypes.c_uint16 * MAX_REG_BITS),
                ("stoff", ctypes.c_uint16 * MAX_REG_BITS)]


class DTileRun(ctypes.Structure):
    _fields_ = [("src", ctypes.c_uint8), ("len", ctypes.c_uint8), ("dst", ctypes.c_uint8),
                ("pad", ctypes.c_uint8)]


class PassParams(ctypes.Structure):
    _fields_ = [("shard_base", ctypes.c_uint64), ("tile_pos", ctypes.c_uint8 * TILE_BITS),
                ("tile_bits", ctypes.c_int32), ("nontile_mask", ctypes.c_uint64),
                ("n_segments", ctypes.c_int32), ("n_gates", ctypes.c_int32),
                ("reg_bits", ctypes.c_int32), ("n_tile_runs", ctypes.c_int32),
                ("tile_run", DTileRun * 12), ("goff", (ctypes.c_uint64 * MAX_REG_BITS) * 2),
                ("seg", DSegment * MAX_SEGMENTS), ("gate", DGate * (MAX_GATES + MAX_FANS)),
                ("ufan_header", ctypes.c_uint16 * MAX_FANS), ("n_ufans", ctypes.c_int32),
                ("n_thread_tables", ctypes.c_int32)]


def read_plan(circuit):
    """The PassParams blocks of the circuit's last flush."""
    C = circuit.C
    assert ctypes.sizeof(PassParams) == C.qcs_cuda_pass_descriptor_bytes(), "layout mirror out of date"
    n = C.qcs_cuda_last_plan_raw(circuit.e, -1, None, 0)
    out = []
    for k in range(n):
        p = PassParams()
        C.qcs_cuda_last_plan_raw(circuit.e, k, ctypes.byref(p), ctypes.sizeof(p))
        p.thread_tables = read_tables(circuit, k)   # rides along as a plain attribute
        check_thread_tables(p, p.thread_tables)
        out.append(p)
    return out


def read_tables(circuit, pass_index):
    """The thread-table fans that travel next to a pass (math=fast): array [table, thread id] of complex."""
    C = circuit.C
    n = C.qcs_cuda_last_plan_tables(circuit.e, pass_index, None, 0)
    if n == 0:
        return np.zeros((0, 0), dtype=np.complex128)
    buf = np.empty(n, dtype=np.float64)
    assert C.qcs_cuda_last_plan_tables(circuit.e, pass_index, buf.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), n) == n
    return buf.view(np.complex128)


def check_thread_tables(p, tables):
    """Every thread-table fan of pass `p` against its own entry records: table[tid] must be the product of
    the phases of the entries whose control the thread with id `tid` has set, `tid` mapped to tile-index
    bits through the segment's role table exactly as the kernel maps it."""
    T, R = p.tile_bits, p.reg_bits
    CT = 1 << (T - R)
    assert tables.size == p.n_thread_tables * CT
    tables = tables.reshape(p.n_thread_tables, CT) if p.n_thread_tables else tables
    tile_pos = [p.tile_pos[b] for b in range(T)]
    seen = 0
    for s in range(p.n_segments):
        seg = p.seg[s]
        for gi in range(seg.gate_begin, seg.gate_end):
            g = p.gate[gi]
            if not (g.flags & GF_TFAN_HEADER):
                continue
            assert g.tsel == seen, "tables are numbered in record order"
            seen += 1
            n_own = g.pad[2]
            want = np.ones(CT, dtype=np.complex128)
            for tid in range(CT):
                tb = 0
                for role in range(T - R):
                    if (tid >> role) & 1:
                        tb |= 1 << seg.role_tilebit[role]
                for e in range(n_own):
                    rec = p.gate[gi + 1 + e]
                    assert rec.tpos == g.tpos and rec.cpos in tile_pos and rec.kind == GK_DIAG
                    if (tb >> tile_pos.index(rec.cpos)) & 1:
                        want[tid] *= complex(rec.m[6], rec.m[7])
            assert np.abs(tables[g.tsel] - want).max() <= 1e-14, "thread table disagrees with its entries"
            if g.csel != 0xFF:
                h = gi + 1 + n_own
                assert p.ufan_header[g.csel] == h and (p.gate[h].flags & GF_FAN_HEADER) and g.pad[1] == n_own + 1 + p.gate[h].tsel
            else:
                assert g.pad[1] == n_own
    assert seen == p.n_thread_tables


def _bit(idx, pos):
    return (idx >> np.uint64(pos)) & np.uint64(1)


def _apply_gate(state, idx, g):
    m = [complex(g.m[2 * k], g.m[2 * k + 1]) for k in range(4)]
    ctl = np.ones(idx.shape, dtype=bool) if g.cpos < 0 else _bit(idx, g.cpos) == 1
    if g.kind == GK_NOP:
        return
    if g.kind == GK_DIAG:
        t1 = _bit(idx, g.tpos) == 1
        state[ctl & ~t1] *= m[0]
        state[ctl & t1] *= m[3]
        return
    i0 = idx[ctl & (_bit(idx, g.tpos) == 0)]
    i1 = i0 | np.uint64(1 << g.tpos)
    v0, v1 = state[i0].copy(), state[i1].copy()
    state[i0] = m[0] * v0 + m[1] * v1
    if not (g.flags & GF_ROW0_ONLY):
        state[i1] = m[2] * v0 + m[3] * v1


def _apply_fast_fan(state, idx, gates, h, uniform=False):
    """A fan the way the math=fast kernel walks it: the thread's participation mask, then one table
    lookup per group of four entries, then ONE multiplication of the amplitudes with target bit 1."""
    hd = gates[h]
    K = hd.tsel
    mask = np.zeros(idx.shape, dtype=np.uint64)
    for k in range(K):
        cpos = gates[h + 1 + k].cpos
        if hd.csel != 0xFF and not uniform:
            assert cpos == hd.csel + k, "consecutive-control fan with a gap"
        if uniform and hd.pad[0]:
            assert cpos == hd.cpos + k, "consecutive-control uniform fan with a gap"
        mask |= _bit(idx, cpos) << np.uint64(k)
    mask[_bit(idx, hd.tpos) == 0] = 0
    factor = np.ones(idx.shape, dtype=np.complex128)
    for i in range((K + 3) // 4):
        nib = ((mask >> np.uint64(4 * i)) & np.uint64(15)).astype(np.int64)
        table = np.empty(16, dtype=np.complex128)
        for j in range(16):
            rec = h + 1 + 4 * i + (j >> 2)
            table[j] = complex(gates[rec].m[2 * (j & 3)], gates[rec].m[2 * (j & 3) + 1]) if rec <= h + K else np.nan
        assert table[0] == 1.0
        factor *= table[nib]
    state *= factor


def run_plan(passes, n_qubits, fast, state=None):
    """Applies the passes to `state` (default |0..0>) and returns it."""
    if state is None:
        state = np.zeros(1 << n_qubits, dtype=np.complex128)
        state[0] = 1.0
    idx = np.arange(1 << n_qubits, dtype=np.uint64)
    for p in passes:
        covered = 0
        # uniform fans (math=fast): listed for the kernel prologue, every control outside the tile
        tile = set(p.tile_pos[b] for b in range(p.tile_bits))
        for k in range(p.n_ufans):
            hd = p.gate[p.ufan_header[k]]
            assert (hd.flags & GF_FAN_HEADER) and hd.csel == k
            assert all(p.gate[p.ufan_header[k] + 1 + e].cpos not in tile for e in range(hd.tsel))
        for s in range(p.n_segments):
            seg = p.seg[s]
            assert seg.gate_begin == covered, "segments must tile the gate list"
            gi = seg.gate_begin
            while gi < seg.gate_end:
                g = p.gate[gi]
                if g.flags & GF_TFAN_HEADER:
                    # thread-table fan (math=fast): the own entries behind the header are ordinary records --
                    # applied one by one here; that their product is what the table holds is
                    # check_thread_tables' business -- and a uniform fan may follow (handled below)
                    assert gi + g.pad[1] <= seg.gate_end - 1, "thread-table fan crosses its segment"
                    gi += 1
                    continue
                if g.flags & GF_FAN_HEADER:
                    assert gi + g.tsel <= seg.gate_end - 1, "fan crosses its segment"
                    if fast:
                        uniform = gi in [p.ufan_header[k] for k in range(p.n_ufans)]
                        _apply_fast_fan(state, idx, p.gate, gi, uniform)
                        gi += 1 + g.tsel
                    else:
                        gi += 1  # exact mode: the entries behind the header are ordinary gate records
                    continue
                _apply_gate(state, idx, g)
                gi += 1
            covered = seg.gate_end
        assert covered == p.n_gates
    return state
