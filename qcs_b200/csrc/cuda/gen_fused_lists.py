#!/usr/bin/env python3
"""Generates the gate interpreter of the fused tile kernel.

A thread holds NREG = 2^R amplitudes in named PTX registers qr0.. / qi0..; register index bit
k is register-role bit k.  A gate pairing across role bit T touches pairs (r, r | 1<<T) with
bit T of r clear, optionally only those with bit C set (control on a register bit).

The interpreter is ONE inline-PTX block per register-file shape: a `brx.idx` jump table indexed by
the gate's case label, every case a straight-line run of mul.rn / add.rn / sub.rn on the named
registers (never contracted, the reference's operation order: c_mul then c_add, reference
src/complex.c:23-57), matrix entries fetched from the pass descriptor in kernel-parameter space by
`ld.param` right where they are used.  Written as C++ `switch` the same table is lowered by the
compiler to a compare tree (~35 instructions per gate, profiles/r1_fused_kernel_history.md); the
jump table costs ~11.  FP64 instructions issue at half rate on sm_100, everything else at full
rate, so a gate's cost is 2 x (FP64 instructions) + (everything else): the dispatch is paid in
the same currency as the arithmetic.

Outputs
  gen_fused_lists.py lists > fused_lists.inc   QCS<R>_REGS_ALL(OP), QCS<R>_DISPATCH_ASM
  gen_fused_lists.py ops   > fused_ops.h       QCS<R>_CASE_LABEL[symbolic id] (planner side)
(build.py regenerates both when this file is newer.)

Asm operands of QCS<R>_DISPATCH_ASM:  %0 "+r" gate index (advanced past what was executed),
%1 "l" param-space address of the current DGate, %2 "l" xfull = the thread's basis index with the
register bits clear (rank bits included).  The named register `ufb` (declared with the amplitude
registers, set once per CTA) holds the shared-memory address of the uniform-fan factor table: as
an asm operand the compiler re-derived it (S2UR + ULEA) in front of every gate.
"""
import sys

KINDS = ["HSYM", "REAL", "SWAP", "GENERIC"]  # order = kind index in qcs_op_id_pair()

# byte offsets inside DGate (common.h)
OFF_M = lambda i: 8 * i       # noqa: E731  m[i]
OFF_W0 = 64                   # op (u16) | csel << 16 | tsel << 24
OFF_TPOS = 70
OFF_CPOS = 71
OFF_PAD1 = 74                 # pad[1]: thread-table fans, records to skip behind the header
SIZEOF_DGATE = 80

SYM_NOP = 255
SYM_TCTL = 0x100


def pairs(R, t, c):
    res = []
    for r in range(1 << R):
        if (r >> t) & 1:
            continue
        if c >= 0 and not (r >> c) & 1:
            continue
        res.append((r, r | (1 << t)))
    return res


def regs(R, t=-1, tval=0, c=-1):
    res = []
    for r in range(1 << R):
        if t >= 0 and ((r >> t) & 1) != tval:
            continue
        if c >= 0 and not (r >> c) & 1:
            continue
        res.append(r)
    return res


# ---- PTX snippets -----------------------------------------------------------------------------
def ld_m(indices):
    return [f"ld.param.f64 g{i}, [%1+{OFF_M(i)}];" for i in indices]


def cmul(xr, xi, gr, gi, v):
    """x = g * v in the reference's c_mul order: x.r = g.r*v.r - g.i*v.i ; x.i = g.r*v.i + g.i*v.r"""
    return [f"mul.rn.f64 {xr}, {gr}, qr{v};", f"mul.rn.f64 u0, {gi}, qi{v};", f"sub.rn.f64 {xr}, {xr}, u0;",
            f"mul.rn.f64 {xi}, {gr}, qi{v};", f"mul.rn.f64 u0, {gi}, qr{v};", f"add.rn.f64 {xi}, {xi}, u0;"]


def cmul_fast(xr, xi, gr, gi, ngi, v, acc):
    """fast mode: x (+)= g * v with fused multiply-adds; ngi = register holding -g.i.
    acc=False starts the chain with a plain product."""
    if acc:
        return [f"fma.rn.f64 {xr}, {gr}, qr{v}, {xr};", f"fma.rn.f64 {xr}, {ngi}, qi{v}, {xr};",
                f"fma.rn.f64 {xi}, {gr}, qi{v}, {xi};", f"fma.rn.f64 {xi}, {gi}, qr{v}, {xi};"]
    return [f"mul.rn.f64 {xr}, {gr}, qr{v};", f"fma.rn.f64 {xr}, {ngi}, qi{v}, {xr};",
            f"mul.rn.f64 {xi}, {gr}, qi{v};", f"fma.rn.f64 {xi}, {gi}, qr{v}, {xi};"]


def op_pair_fast(kind, row0, a, b):
    """math=fast: the same update with fused multiply-adds (results within a few ulp of the
    reference's separately rounded ones; never used in the bit-exact mode).  n* registers hold
    negated matrix entries (pair_loads_fast)."""
    if kind == "HSYM":
        o = [f"mul.rn.f64 t0, g0, qr{a};", f"mul.rn.f64 t1, g0, qi{a};",
             f"fma.rn.f64 qr{a}, g2, qr{b}, t0;", f"fma.rn.f64 qi{a}, g2, qi{b}, t1;"]
        if not row0:
            o += [f"fma.rn.f64 qr{b}, n2, qr{b}, t0;", f"fma.rn.f64 qi{b}, n2, qi{b}, t1;"]
        return o
    if kind == "REAL":
        o = [f"mul.rn.f64 t0, g0, qr{a};", f"mul.rn.f64 t1, g0, qi{a};"]
        if not row0:
            o += [f"mul.rn.f64 t4, g4, qr{a};", f"mul.rn.f64 t5, g4, qi{a};"]
        o += [f"fma.rn.f64 qr{a}, g2, qr{b}, t0;", f"fma.rn.f64 qi{a}, g2, qi{b}, t1;"]
        if not row0:
            o += [f"fma.rn.f64 qr{b}, g6, qr{b}, t4;", f"fma.rn.f64 qi{b}, g6, qi{b}, t5;"]
        return o
    if kind == "SWAP":
        return op_pair(kind, row0, a, b)
    o = cmul_fast("t0", "t1", "g0", "g1", "n1", a, False) + cmul_fast("t0", "t1", "g2", "g3", "n3", b, True)
    if not row0:
        o += cmul_fast("t4", "t5", "g4", "g5", "n5", a, False) + cmul_fast("t4", "t5", "g6", "g7", "n7", b, True)
    o += [f"mov.f64 qr{a}, t0;", f"mov.f64 qi{a}, t1;"]
    if not row0:
        o += [f"mov.f64 qr{b}, t4;", f"mov.f64 qi{b}, t5;"]
    return o


def pair_loads_fast(kind, row0):
    o = pair_loads(kind, row0)
    if kind == "HSYM" and not row0:
        o += ["neg.f64 n2, g2;"]
    if kind == "GENERIC":
        o += ["neg.f64 n1, g1;", "neg.f64 n3, g3;"]
        if not row0:
            o += ["neg.f64 n5, g5;", "neg.f64 n7, g7;"]
    return o


def op_diag_fast(r, er="dr", ei="di"):
    """math=fast: amplitude r *= (er, ei) with fused multiply-adds"""
    return [f"mul.rn.f64 t0, {ei}, qi{r};", f"mul.rn.f64 t1, {ei}, qr{r};", "neg.f64 t0, t0;",
            f"fma.rn.f64 qr{r}, {er}, qr{r}, t0;", f"fma.rn.f64 qi{r}, {er}, qi{r}, t1;"]


def acc_thread_factor(er="dr", ei="di"):
    """math=fast: (fr, fi) *= (er, ei).  (fr, fi) is the thread's pending scalar factor: a diagonal
    gate whose target and control are NOT register bits multiplies all amplitudes of a thread by the
    same number, which commutes with everything else a segment does to the thread's registers, so it
    is multiplied up here and applied once when the segment ends (fused_body.inc, QCS3F_FLUSH_ASM)."""
    return [f"mul.rn.f64 t0, fi, {ei};", f"mul.rn.f64 t1, fi, {er};", "neg.f64 t0, t0;",
            f"fma.rn.f64 t2, fr, {er}, t0;", f"fma.rn.f64 fi, fr, {ei}, t1;", "mov.f64 fr, t2;"]


def op_pair(kind, row0, a, b):
    """a = register with target bit 0, b = target bit 1.  Operation order = c_add(c_mul(g0,v0),
    c_mul(g1,v1)) (reference src/q_gates.c:140-141); row0 = reference-semantics controlled update."""
    if kind == "HSYM":  # real, U00==U10=m0, U01=m2, U11=-m2: the four products are shared
        o = [f"mul.rn.f64 t0, g0, qr{a};", f"mul.rn.f64 t1, g0, qi{a};",
             f"mul.rn.f64 t2, g2, qr{b};", f"mul.rn.f64 t3, g2, qi{b};",
             f"add.rn.f64 qr{a}, t0, t2;", f"add.rn.f64 qi{a}, t1, t3;"]
        if not row0:
            o += [f"sub.rn.f64 qr{b}, t0, t2;", f"sub.rn.f64 qi{b}, t1, t3;"]
        return o
    if kind == "REAL":  # all imaginary parts zero
        o = [f"mul.rn.f64 t0, g0, qr{a};", f"mul.rn.f64 t1, g0, qi{a};",
             f"mul.rn.f64 t2, g2, qr{b};", f"mul.rn.f64 t3, g2, qi{b};"]
        if not row0:
            o += [f"mul.rn.f64 t4, g4, qr{a};", f"mul.rn.f64 t5, g4, qi{a};",
                  f"mul.rn.f64 t6, g6, qr{b};", f"mul.rn.f64 t7, g6, qi{b};"]
        o += [f"add.rn.f64 qr{a}, t0, t2;", f"add.rn.f64 qi{a}, t1, t3;"]
        if not row0:
            o += [f"add.rn.f64 qr{b}, t4, t6;", f"add.rn.f64 qi{b}, t5, t7;"]
        return o
    if kind == "SWAP":  # exact X: moves only
        if row0:
            return [f"mov.f64 qr{a}, qr{b};", f"mov.f64 qi{a}, qi{b};"]
        return [f"mov.f64 t0, qr{a};", f"mov.f64 t1, qi{a};", f"mov.f64 qr{a}, qr{b};",
                f"mov.f64 qi{a}, qi{b};", f"mov.f64 qr{b}, t0;", f"mov.f64 qi{b}, t1;"]
    # GENERIC: the full 28-flop update
    o = cmul("t0", "t1", "g0", "g1", a) + cmul("t2", "t3", "g2", "g3", b)
    if not row0:
        o += cmul("t4", "t5", "g4", "g5", a) + cmul("t6", "t7", "g6", "g7", b)
    o += [f"add.rn.f64 qr{a}, t0, t2;", f"add.rn.f64 qi{a}, t1, t3;"]
    if not row0:
        o += [f"add.rn.f64 qr{b}, t4, t6;", f"add.rn.f64 qi{b}, t5, t7;"]
    return o


def pair_loads(kind, row0):
    if kind == "HSYM":
        return ld_m([0, 2])
    if kind == "REAL":
        return ld_m([0, 2] if row0 else [0, 2, 4, 6])
    if kind == "SWAP":
        return []
    return ld_m([0, 1, 2, 3] if row0 else range(8))


def op_diag(r, er="dr", ei="di"):
    """amplitude r *= (er, ei), c_mul order"""
    return [f"mul.rn.f64 t0, {er}, qr{r};", f"mul.rn.f64 t1, {ei}, qi{r};",
            f"mul.rn.f64 t2, {er}, qi{r};", f"mul.rn.f64 t3, {ei}, qr{r};",
            f"sub.rn.f64 qr{r}, t0, t1;", f"add.rn.f64 qi{r}, t2, t3;"]


def thread_bit_test(field_shift, pred="p"):
    """pred := bit (header byte at w0 >> field_shift) of xfull is SET"""
    return [f"bfe.u32 cs, w0, {field_shift}, 8;", "shr.u64 t64, %2, cs;", "and.b64 t64, t64, 1;",
            f"setp.ne.u64 {pred}, t64, 0;"]


def emit(R, fast=False):
    """fast=False: the bit-exact interpreter.  fast=True (math=fast): the same case
    labels and record layout, fused multiply-adds, and controlled-phase fans that multiply the
    entries' phases into ONE per-thread factor (looked up four entries at a time in tables the
    planner lays over the entries' matrix slots) before touching the amplitudes."""
    P = f"QCS{R}F" if fast else f"QCS{R}"
    pair_body = op_pair_fast if fast else op_pair
    pair_ld = pair_loads_fast if fast else pair_loads
    diag_body = op_diag_fast if fast else op_diag
    labels = {}          # symbolic id -> case label (consecutive)
    blocks = []          # (label name, [ptx lines])

    def new_case(sym):
        labels[sym] = len(labels)
        return f"C{labels[sym]}"

    def case(sym, body, thread_ctl):
        # a control that is not a register bit: the same block behind one test of the thread's own
        # basis index; the TCTL label sits right in front of the plain one and falls through
        if thread_ctl:
            blocks.append((new_case(sym | SYM_TCTL), thread_bit_test(16) + ["@!p bra DONE;"]))
        blocks.append((new_case(sym), body + ["bra DONE;"]))

    # pairing ops: symbolic id = ((kind*2 + row0)*4 + treg)*5 + (creg+1)
    for ki, kind in enumerate(KINDS):
        for row0 in (0, 1):
            for t in range(R):
                for c in range(-1, R):
                    if c == t:
                        continue
                    sym = ((ki * 2 + row0) * 4 + t) * 5 + (c + 1)
                    body = pair_ld(kind, row0)
                    for a, b in pairs(R, t, c):
                        body += pair_body(kind, row0, a, b)
                    case(sym, body, c < 0)
    # diagonal, target NOT a register bit: id = 160 + (creg+1)*3 + (halves-1); the thread's own target
    # bit picks the entry.  halves: 1 = only entry 0 is not the identity, 2 = only entry 1, 3 = both
    for c in range(-1, R):
        for halves in (1, 2, 3):
            sym = 160 + (c + 1) * 3 + (halves - 1)
            body = thread_bit_test(24)
            if halves == 3:
                body += ld_m([0, 1, 6, 7]) + ["selp.f64 dr, g6, g0, p;", "selp.f64 di, g7, g1, p;"]
            elif halves == 2:
                body += ["@!p bra DONE;", f"ld.param.f64 dr, [%1+{OFF_M(6)}];", f"ld.param.f64 di, [%1+{OFF_M(7)}];"]
            else:
                body += ["@p bra DONE;", f"ld.param.f64 dr, [%1+{OFF_M(0)}];", f"ld.param.f64 di, [%1+{OFF_M(1)}];"]
            if fast and c < 0:
                body += acc_thread_factor()
            else:
                for r in regs(R, c=c):
                    body += diag_body(r)
            case(sym, body, c < 0)
    # diagonal, target = register bit t: id = 175 + (t*5 + creg+1)*3 + (halves-1)
    for t in range(R):
        for c in range(-1, R):
            if c == t:
                continue
            for halves in (1, 2, 3):
                sym = 175 + (t * 5 + c + 1) * 3 + (halves - 1)
                body = []
                if halves & 1:
                    body += ld_m([0, 1])
                    for r in regs(R, t, 0, c):
                        body += diag_body(r, "g0", "g1")
                if halves & 2:
                    body += ld_m([6, 7])
                    for r in regs(R, t, 1, c):
                        body += diag_body(r, "g6", "g7")
                case(sym, body, c < 0)
    # controlled-phase fan (id 240 + treg + 1; header layout: common.h QCS_OP_FAN_BASE): K consecutive
    # gates diag(1, e^{i a_k}) on ONE target, each controlled by a bit that is not a register bit.
    # Bit k of `mask` = this thread takes part in entry k.  Entries no lane of the warp takes part in
    # are skipped without being touched; per entry: isolate the lowest pending bit, test the thread's
    # own bit, fetch the phase, multiply the registers whose target bit is 1.  Same arithmetic and
    # the same order per amplitude as gate by gate.
    for t in range(-1, R):
        sym = 240 + t + 1
        n = f"{t + 1}"
        E = SIZEOF_DGATE  # entries follow the header record
        body = ["bfe.u32 K, w0, 24, 8;", "bfe.u32 c0, w0, 16, 8;", "setp.eq.u32 p, c0, 255;",
                f"@p bra FG{n};",
                "shr.u64 t64, %2, c0;", "cvt.u32.u64 mask, t64;",
                f"FM{n}:", "bfe.u32 mask, mask, 0, K;"]
        if t < 0:  # the target is a thread-level bit too: all or none of this thread's amplitudes
            body += [f"ld.param.u8 cs, [%1+{OFF_TPOS}];", "shr.u64 t64, %2, cs;", "and.b64 t64, t64, 1;",
                     "setp.eq.u64 p, t64, 0;", "@p mov.u32 mask, 0;"]
        diag_regs = regs(R, t, 1) if t >= 0 else regs(R)
        if fast:
            # Table walk (planner.cpp write_fan_tables): group i = entries 4i..4i+3; its 16 products
            # sit in the m[] slots of those four records, product j at record 4i + (j >> 2), slot
            # j & 3.  One lookup + one complex multiply-add per group, whatever the entry count;
            # groups in which no lane of the warp takes part are skipped.
            body += ["mov.b32 tst, mask;",  # this thread's own mask, kept for the final test
                     "mov.f64 ar, 0d3FF0000000000000;", "mov.f64 ai, 0d0000000000000000;",
                     "redux.sync.or.b32 wm, mask, 0xffffffff;", f"add.u64 ea, %1, {E};",
                     f"FL{n}:", "setp.eq.u32 p, wm, 0;", f"@p bra FE{n};",
                     "and.b32 low, wm, 15;", "setp.eq.u32 pa, low, 0;", f"@pa bra FS{n};",
                     "and.b32 low, mask, 15;", "shr.u32 kidx, low, 2;", "and.b32 low, low, 3;",
                     f"mul.lo.u32 kidx, kidx, {E};", "mad.lo.u32 kidx, low, 16, kidx;",
                     "cvt.u64.u32 t64, kidx;", "add.u64 t64, t64, ea;",
                     "ld.param.v2.f64 {dr, di}, [t64];",
                     "mul.rn.f64 t0, ai, di;", "mul.rn.f64 t1, ai, dr;", "neg.f64 t0, t0;",
                     "fma.rn.f64 t2, ar, dr, t0;", "fma.rn.f64 ai, ar, di, t1;", "mov.f64 ar, t2;",
                     f"FS{n}:", "shr.u32 mask, mask, 4;", "shr.u32 wm, wm, 4;",
                     f"add.u64 ea, ea, {4 * E};", f"bra FL{n};",
                     f"FE{n}:", "setp.eq.u32 p, tst, 0;", f"@p bra FX{n};"]
            if t < 0:
                body += acc_thread_factor("ar", "ai")
            else:
                for r in diag_regs:
                    body += op_diag_fast(r, "ar", "ai")
            body += [f"FX{n}:", "add.s32 %0, %0, K;", "bra DONE;",
                     f"FG{n}:", "mov.u32 mask, 0;", "mov.u32 kidx, 0;", "mov.u64 ea, %1;",
                     f"FH{n}:", f"ld.param.u8 cs, [ea+{E + OFF_CPOS}];", "shr.u64 t64, %2, cs;",
                     "cvt.u32.u64 tst, t64;", "and.b32 tst, tst, 1;", "shl.b32 tst, tst, kidx;",
                     "or.b32 mask, mask, tst;", f"add.u64 ea, ea, {E};", "add.u32 kidx, kidx, 1;",
                     "setp.lt.u32 p, kidx, K;", f"@p bra FH{n};", f"bra FM{n};"]
            blocks.append((new_case(sym), body))
            continue
        entry = ["bfind.u32 kidx, low;", f"mul.wide.u32 ea, kidx, {E};", "add.u64 ea, ea, %1;",
                 f"ld.param.v2.f64 {{dr, di}}, [ea+{E + OFF_M(6)}];"]
        for r in diag_regs:
            entry += op_diag(r)
        body += ["redux.sync.or.b32 wm, mask, 0xffffffff;",
                 # all lanes of the warp take part in the same entries (controls on warp bits or
                 # outside the tile: the common case): no per-entry lane test, no reconvergence
                 "setp.eq.u32 pa, mask, wm;", "vote.sync.all.pred pa, pa, 0xffffffff;", f"@pa bra FU{n};",
                 f"FL{n}:", "setp.eq.u32 p, wm, 0;", f"@p bra FE{n};",
                 "neg.s32 low, wm;", "and.b32 low, low, wm;", "xor.b32 wm, wm, low;",
                 "and.b32 tst, mask, low;", "setp.eq.u32 pa, tst, 0;", f"@pa bra FS{n};"]
        body += entry
        body += [f"FS{n}:", f"bra FL{n};",
                 f"FU{n}:", "setp.eq.u32 p, wm, 0;", f"@p bra FE{n};",
                 "neg.s32 low, wm;", "and.b32 low, low, wm;", "xor.b32 wm, wm, low;"]
        body += entry
        # (Round 2: this loop was also tried software-pipelined -- the next entry's phase fetched into a
        # second register pair before the current entry's FP64 block, two entries per trip.  The ncu source
        # view had put 8 % of a QFT pass's stall samples on the first DMUL behind each LDCU pair, yet the
        # 30-qubit QFT went from 62.5 to 63.8 ms: the other seven warps of the scheduler already cover
        # that latency, and the USEL / predicated uniform instructions the pipelining adds are not free.)
        body += [f"bra.uni FU{n};",
                 f"FE{n}:", "add.s32 %0, %0, K;", "bra DONE;",
                 # controls at arbitrary positions: gather the mask from the entries' cpos bytes
                 f"FG{n}:", "mov.u32 mask, 0;", "mov.u32 kidx, 0;", "mov.u64 ea, %1;",
                 f"FH{n}:", f"ld.param.u8 cs, [ea+{E + OFF_CPOS}];", "shr.u64 t64, %2, cs;",
                 "cvt.u32.u64 tst, t64;", "and.b32 tst, tst, 1;", "shl.b32 tst, tst, kidx;",
                 "or.b32 mask, mask, tst;", f"add.u64 ea, ea, {E};", "add.u32 kidx, kidx, 1;",
                 "setp.lt.u32 p, kidx, K;", f"@p bra FH{n};", f"bra FM{n};"]
        blocks.append((new_case(sym), body))
    # uniform fan (id 248 + treg + 1, math=fast only; common.h QCS_OP_UFAN_BASE): the product of the
    # entries' phases is the same for the whole CTA and waits in shared-memory slot `csel` (ufb = base
    # address of the table).  The exact interpreter never sees one (label kept so that both
    # interpreters share one label table).
    for t in range(-1, R):
        sym = 248 + t + 1
        n = f"{t + 1}"
        if not fast:
            blocks.append((new_case(sym), ["bra DONE;"]))
            continue
        body = ["bfe.u32 K, w0, 24, 8;", "bfe.u32 c0, w0, 16, 8;", "add.s32 %0, %0, K;",
                "mad.lo.u32 c0, c0, 16, ufb;", "ld.shared.v2.f64 {ar, ai}, [c0];"]
        if t < 0:
            body += [f"ld.param.u8 cs, [%1+{OFF_TPOS}];", "shr.u64 t64, %2, cs;", "and.b64 t64, t64, 1;",
                     "setp.eq.u64 p, t64, 0;", "@p bra DONE;"]
            body += acc_thread_factor("ar", "ai")
        else:
            for r in regs(R, t, 1):
                body += op_diag_fast(r, "ar", "ai")
        blocks.append((new_case(sym), body + ["bra DONE;"]))
    # thread-table fan (id 235 + treg + 1, math=fast only; common.h QCS_OP_TFAN_BASE): the product of the
    # phases of the in-tile controls this thread has set comes out of a host-built table -- one load, the
    # threads of a warp read consecutive entries -- times the per-CTA factor of the out-of-tile controls
    # (uniform-fan slot csel, 0xFF = none).  tfb (named 64-bit register, set once per CTA) = address of
    # this thread's entry of table 0, tfs = bytes per table.  pad[1] records behind the header are skipped.
    for t in range(-1, R):
        sym = 235 + t + 1
        n = f"{t + 1}"
        if not fast:
            blocks.append((new_case(sym), ["bra DONE;"]))
            continue
        body = [f"ld.param.u8 cs, [%1+{OFF_PAD1}];", "add.s32 %0, %0, cs;",
                "bfe.u32 K, w0, 24, 8;", "bfe.u32 c0, w0, 16, 8;",
                "mul.wide.u32 ea, K, tfs;", "add.u64 ea, ea, tfb;",
                "ld.global.nc.v2.f64 {ar, ai}, [ea];",
                "setp.eq.u32 p, c0, 255;", f"@p bra TA{n};",
                "mad.lo.u32 c0, c0, 16, ufb;", "ld.shared.v2.f64 {dr, di}, [c0];",
                "mul.rn.f64 t0, ai, di;", "mul.rn.f64 t1, ai, dr;", "neg.f64 t0, t0;",
                "fma.rn.f64 t2, ar, dr, t0;", "fma.rn.f64 ai, ar, di, t1;", "mov.f64 ar, t2;",
                f"TA{n}:"]
        if t < 0:
            body += [f"ld.param.u8 cs, [%1+{OFF_TPOS}];", "shr.u64 t64, %2, cs;", "and.b64 t64, t64, 1;",
                     "setp.eq.u64 p, t64, 0;", "@p bra DONE;"]
            body += acc_thread_factor("ar", "ai")
        else:
            for r in regs(R, t, 1):
                body += op_diag_fast(r, "ar", "ai")
        blocks.append((new_case(sym), body + ["bra DONE;"]))
    blocks.append((new_case(SYM_NOP), ["bra DONE;"]))

    n_cases = len(labels)
    text = ["{", ".reg .b32 w0, op, cs, c0, K, mask, wm, low, tst, kidx;", ".reg .b64 t64, ea;",
            ".reg .pred p, pa;", ".reg .f64 g<8>, t<8>, n<8>, u0, dr, di, ar, ai;",
            f"ld.param.u32 w0, [%1+{OFF_W0}];", "add.s32 %0, %0, 1;", "and.b32 op, w0, 0xffff;",
            "TS: .branchtargets " + ", ".join(f"C{i}" for i in range(n_cases)) + ";",
            "brx.idx op, TS;"]
    for name, body in blocks:
        text.append(f"{name}:")
        text += body
    text += ["DONE:", "}"]
    out = [f"#define {P}_REGS_ALL(OP) " + " ".join(f"OP({r})" for r in regs(R))]
    if fast:
        flush = ["{", ".reg .pred p, q;", ".reg .f64 t0, t1;",
                 "setp.eq.f64 p, fr, 0d3FF0000000000000;", "setp.eq.f64 q, fi, 0d0000000000000000;",
                 "and.pred p, p, q;", "@p bra FLUSHED;"]
        for r in regs(R):
            flush += op_diag_fast(r, "fr", "fi")
        flush += ["FLUSHED:", "}"]
        out.append(f"#define {P}_FLUSH_ASM \\")
        for line in flush:
            sep = "\\n" if line.endswith(":") or line in ("{", "}") else "\\n\\t"
            out.append(f'  "{line}{sep}" \\')
        out[-1] = out[-1][:-2]
    out.append(f"#define {P}_DISPATCH_ASM \\")
    for line in text:
        sep = "\\n" if line.endswith(":") or line in ("{", "}") else "\\n\\t"
        out.append(f'  "{line}{sep}" \\')
    out[-1] = out[-1][:-2]
    table = [labels[SYM_NOP]] * 512  # anything unassigned executes nothing
    for sym, lab in labels.items():
        table[sym] = lab
    return out, table


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "lists"
    if what == "lists":
        out = ["// GENERATED by gen_fused_lists.py -- do not edit.", "#pragma once", ""]
        for R in (3, 4):
            out += emit(R)[0]
            out.append("")
        for R in (3, 4):
            out += emit(R, fast=True)[0]
            out.append("")
    else:
        out = ["// GENERATED by gen_fused_lists.py -- do not edit.",
               "// symbolic op id (common.h qcs_op_id_*, | QCS_OP_TCTL) -> case label of the generated jump table",
               "#pragma once", ""]
        assert all(emit(R, fast=True)[1] == emit(R)[1] for R in (3, 4))  # one label table serves both interpreters
        for R in (3, 4):
            table = emit(R)[1]
            out.append(f"static const unsigned short QCS{R}_CASE_LABEL[512] = {{")
            for i in range(0, 512, 16):
                out.append("  " + ", ".join(str(v) for v in table[i:i + 16]) + ",")
            out.append("};")
            out.append("")
    print("\n".join(out))


if __name__ == "__main__":
    main()
