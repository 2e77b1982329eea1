"""Gate-fusion scheduler invariants, checked on plan-only (dry-run) engines: no GPU needed."""
import re

import numpy as np
import pytest

from oracle import pyoracle as po


@pytest.fixture(scope="module", autouse=True)
def _built():
    from qcs_b200 import build
    build.build_all()


def _plan(n, script, **kw):
    from qcs_b200 import Circuit
    c = Circuit(n, dryrun=True, **kw)
    po.replay(c, script)
    c.flush()
    return c.describe_plan(), c.stats()


def _parse(text):
    passes = []
    for line in text.splitlines():
        m = re.match(r"pass (\d+): gates=(\d+) api_gates=(\d+) segments=(\d+) flops/amp=([\d.]+) tile=\[(.*)\]", line)
        if m:
            passes.append({"gates": int(m.group(2)), "api": int(m.group(3)), "nseg": int(m.group(4)),
                           "flops": float(m.group(5)), "tile": [int(x) for x in m.group(6).split(",")],
                           "segs": []})
            continue
        m = re.match(r"\s+seg (\d+): regs\(pos\)=\[(.*?)\] gates (\d+)\.\.(\d+):(.*)", line)
        if m:
            gates = re.findall(r"(\w+)\((?:c(\d+),)?t(\d+)\)", m.group(5))
            passes[-1]["segs"].append({"regs": [int(x) for x in m.group(2).split(",")],
                                       "range": (int(m.group(3)), int(m.group(4))),
                                       "gates": [(k, int(c) if c else -1, int(t)) for k, c, t in gates]})
    return passes


PAIRING = {"generic", "real", "hsym", "swap"}


@pytest.mark.parametrize("tile_kernel,reg_bits,tile_bits", [("tma", 3, 12), ("tma16", 4, 12), ("ldg", 4, 12),
                                                            ("ldg", 4, 11), ("ldg", 4, 10),
                                                            ("ldg8", 3, 12), ("ldg8", 3, 11), ("ldg8", 3, 10)])
@pytest.mark.parametrize("n", [12, 20, 30, 34])
def test_plan_invariants_random_circuit(n, tile_kernel, reg_bits, tile_bits):
    script = po.random_circuit_script(n, 6)
    text, st = _plan(n, script, semantics="corrected", tile_kernel=tile_kernel, tile_bits=tile_bits)
    passes = _parse(text)
    assert passes and st["passes"] == len(passes)
    assert sum(p["api"] for p in passes) == len(script)
    for p in passes:
        tile = p["tile"]
        # a pass runs on the smallest tile (>= 10 bits for the plain-load kernels, else 12) that holds its pairing qubits
        lo = 10 if tile_kernel in ("ldg8", "ldg") else 12
        assert lo <= len(tile) <= max(lo, tile_bits) and tile == sorted(set(tile)) and tile[:5] == [0, 1, 2, 3, 4]
        assert all(t < n for t in tile)
        assert p["nseg"] == len(p["segs"]) <= 12
        # first and last segment keep the low five bits on the lanes (coalesced 512-byte rows)
        assert all(r >= 5 for r in p["segs"][0]["regs"])
        assert all(r >= 5 for r in p["segs"][-1]["regs"])
        pos = 0
        for s in p["segs"]:
            assert len(set(s["regs"])) == reg_bits and set(s["regs"]) <= set(tile)
            assert s["range"][0] == pos
            pos = s["range"][1]
            for kind, ctl, tgt in s["gates"]:
                if kind in PAIRING:
                    assert tgt in s["regs"], (kind, tgt, s["regs"])
        assert pos == p["gates"]
    # fusion actually happens: far fewer passes than gates
    assert len(passes) * 3 < len(script)


def _order_case(n=16, depth=3, seed=5, **kw):
    script = po.random_circuit_script(n, depth, seed=seed)
    text, _ = _plan(n, script, semantics="corrected", **kw)
    flat = [g for p in _parse(text) for s in p["segs"] for g in s["gates"]]
    want = []
    for op in script:
        if op[0] == "h": want.append(("hsym", -1, op[1]))
        elif op[0] == "rz": want.append(("diag", -1, op[1]))
        elif op[0] == "cnot": want.append(("swap", op[1], op[2]))
    return flat, want


def test_gate_order_is_preserved():
    flat, want = _order_case(reorder="off")
    assert flat == want


def test_exact_mode_only_moves_gates_that_move_amplitudes():
    """Default bit-exact schedule (PlannerConfig::reorder_exact): gates that round keep their relative
    order, whatever qubits they act on; a CNOT (moves only) may trade places with a gate it commutes
    with -- never with one that pairs on one of its qubits or is diagonal on its target."""
    flat, want = _order_case(depth=6)
    assert sorted(flat) == sorted(want)
    assert flat != want, "the case is meant to exercise a reordered plan"
    assert [g for g in flat if g[0] != "swap"] == [g for g in want if g[0] != "swap"]
    # brickwork: every gate is unique per (layer, qubits) but tuples repeat across layers -- match the
    # k-th occurrence of a tuple in the plan with its k-th occurrence in the script (a gate never
    # overtakes an identical one: same qubits, at least one pairing use)
    seen, where = {}, {}
    for i, g in enumerate(flat):
        k = seen.get(g, 0); seen[g] = k + 1
        where[(g, k)] = i
    seen, pos = {}, []
    for g in want:
        k = seen.get(g, 0); seen[g] = k + 1
        pos.append(where[(g, k)])
    pairing = {"hsym", "swap", "generic", "real"}
    for a in range(len(want)):
        for b in range(a + 1, len(want)):
            ka, ca, ta = want[a]; kb, cb, tb = want[b]
            conflict = False
            for q in {ca, ta} & {cb, tb} - {-1}:
                a_pairs = ka in pairing and q == ta
                b_pairs = kb in pairing and q == tb
                conflict = conflict or a_pairs or b_pairs
            if conflict:
                assert pos[a] < pos[b], (want[a], want[b])


def test_qft30_plan_is_compact():
    text, st = _plan(30, [("qft",)], semantics="corrected")
    passes = _parse(text)
    assert st["gates_submitted"] == 465
    assert len(passes) <= 5, len(passes)            # 30 pairing targets, 12 then 7 new positions per pass
    assert all(p["api"] <= 240 for p in passes)     # QCS_MAX_PASS_GATES


def test_tile_cap_search_keeps_the_smallest_tiles_that_need_no_more_passes():
    """engine.cu plan_batch: a 30-qubit QFT needs 5 passes whether 11-bit tiles are allowed or not, so the
    plan on 10-bit tiles wins (targets 10+5+5+5+5); a brickwork circuit needs fewer passes on 11-bit tiles
    and keeps them; tile_search=off restores the greedy growth."""
    for math in ("exact", "fast"):
        on = _parse(_plan(30, [("qft",)], semantics="corrected", math=math)[0])
        off = _parse(_plan(30, [("qft",)], semantics="corrected", math=math, tile_search="off")[0])
        assert len(on) == len(off) == 5
        assert all(len(p["tile"]) == 10 for p in on), [len(p["tile"]) for p in on]
        assert any(len(p["tile"]) == 11 for p in off), [len(p["tile"]) for p in off]
    script = po.random_circuit_script(30, 8)
    on = _parse(_plan(30, script, semantics="corrected")[0])
    small = _parse(_plan(30, script, semantics="corrected", tile_bits=10)[0])
    assert len(on) < len(small) and any(len(p["tile"]) == 11 for p in on)


def test_remap_buffer_option_is_validated():
    from qcs_b200 import Circuit
    from qcs_b200.circuit import QcsError
    for ok in ("inplace", "auto"):
        Circuit(12, dryrun=True, semantics="corrected", remap_buffer=ok).close()
    with pytest.raises(QcsError):
        Circuit(12, dryrun=True, semantics="corrected", remap_buffer="triple")


def test_reference_semantics_cphase_is_a_value_noop():
    """As written in the reference a controlled phase changes no amplitude (defect D1): it is not scheduled."""
    text, st = _plan(14, [("h", 3), ("cphase", 1, 3, 0.4), ("cphase", 9, 3, 0.1), ("h", 8)],
                     semantics="reference")
    kinds = [g[0] for p in _parse(text) for s in p["segs"] for g in s["gates"]]
    assert kinds == ["hsym", "hsym"]
    assert st["gates_submitted"] == 4


def test_pass_flops_budget_splits_passes():
    script = [("rx", q % 12, 0.1 * (q + 1)) for q in range(48)]
    few = _parse(_plan(14, script, semantics="corrected", pass_flops=1000)[0])
    many = _parse(_plan(14, script, semantics="corrected", pass_flops=40)[0])
    assert len(many) > len(few)


def test_queue_peephole_drops_only_exact_pairs():
    """X.X / Y.Y / Z.Z / CNOT.CNOT vanish from the queue (corrected semantics) when
    nothing between them involves their qubits; H.H (identity only up to rounding) stays."""
    def executed(script, **kw):
        _, st = _plan(14, script, semantics="corrected", **kw)
        return st["gates_executed"], st["gates_cancelled"]
    assert executed([("x", 3), ("h", 5), ("x", 3)]) == (1, 2)
    assert executed([("y", 2), ("z", 4), ("rz", 7, 0.3), ("z", 4), ("y", 2)]) == (1, 4)
    assert executed([("cnot", 1, 6), ("h", 9), ("cnot", 1, 6)]) == (1, 2)
    assert executed([("x", 3), ("cnot", 3, 4), ("x", 3)]) == (3, 0)      # the CNOT reads qubit 3
    assert executed([("cnot", 1, 6), ("rz", 6, 0.1), ("cnot", 1, 6)]) == (3, 0)
    assert executed([("h", 3), ("h", 3)]) == (2, 0)
    assert executed([("x", 3), ("x", 3), ("x", 3)]) == (1, 2)
    assert executed([("x", 3), ("h", 5), ("x", 3)], peephole="off") == (3, 0)
    # reference semantics: never (a controlled X is not an involution there, the scratch buffer is observable)
    _, st = _plan(14, [("x", 3), ("x", 3)], semantics="reference")
    assert st["gates_executed"] == 2 and st["gates_cancelled"] == 0


# ---- the descriptors themselves, interpreted on the CPU (tests/plan_emulator.py) -------------
def _close(got, want, tol=1e-12):
    """SURVEY 8(c) metric: |delta| <= tol * max|want| everywhere, relative where amplitudes are not tiny."""
    scale = np.abs(want).max()
    err = np.abs(got - want)
    assert err.max() <= tol * scale, err.max() / scale
    big = np.abs(want) >= 1e-3 * scale
    assert (err[big] / np.abs(want[big])).max() <= tol


def _fan_script(n, seed):
    """Runs of controlled phases on one target with consecutive and scattered controls, H's in between
    (fans of every length 1..n-1, both mask paths of the kernel)."""
    rng = np.random.default_rng(seed)
    script = [("h", q) for q in range(n)]
    for t in range(n):
        ctrls = [q for q in range(n) if q != t]
        if t % 2:
            rng.shuffle(ctrls)
        ctrls = ctrls[: 1 + int(rng.integers(0, n - 1))]
        script += [("cphase", c, t, float(rng.uniform(-3, 3))) for c in ctrls]
        script += [("h", int(rng.integers(0, n))), ("rz", int(rng.integers(0, n)), float(rng.uniform(-3, 3)))]
    return script


@pytest.mark.parametrize("math", ["exact", "fast", "fast+reorder"])
@pytest.mark.parametrize("tile_kernel,tile_bits", [("ldg8", 10), ("ldg8", 11), ("ldg8", 12), ("ldg", 10), ("ldg", 11), ("ldg", 12)])
@pytest.mark.parametrize("case", ["qft", "random+qft", "fans", "generic"])
def test_descriptors_compute_the_circuit(case, tile_kernel, tile_bits, math):
    from qcs_b200 import Circuit
    from tests import plan_emulator as pe
    n = 13
    if case == "qft":
        # a dense start state: on a basis state the QFT's controls are all |0> and its fans do nothing
        script = [("h", q) for q in range(n)] + [("ry", q, 0.3 + 0.1 * q) for q in range(n)] + [("qft",)]
    elif case == "random+qft":
        script = po.random_circuit_script(n, 5, seed=99) + [("qft",)]
    elif case == "fans":
        script = _fan_script(n, 5)
    else:
        rng = np.random.default_rng(3)
        script = []
        for _ in range(60):
            q = int(rng.integers(0, n))
            script.append([("rx", q, float(rng.uniform(-3, 3))), ("ry", q, float(rng.uniform(-3, 3))),
                           ("y", q), ("phase", q, float(rng.uniform(-3, 3))), ("z", q),
                           ("cnot", q, (q + 1 + int(rng.integers(0, n - 1))) % n)][int(rng.integers(0, 6))])
    orc = po.Oracle(n, "corrected")
    po.replay(orc, script)
    want = orc.state()
    orc.close()
    reorder = "on" if math.endswith("+reorder") else "off"
    math = math.split("+")[0]
    c = Circuit(n, dryrun=True, semantics="corrected", tile_kernel=tile_kernel, tile_bits=tile_bits, math=math,
                reorder=reorder, peephole="off")
    po.replay(c, script)
    c.flush()
    passes = pe.read_plan(c)
    assert len(passes) == c.stats()["passes"] > 0
    got = pe.run_plan(passes, n, fast=(math == "fast"))
    c.close()
    _close(got, want)
    assert sum(p_.n_gates for p_ in passes) >= 1
    if math == "fast" and reorder == "off" and case != "generic":
        # the check has teeth: reading the product tables as plain phases gives a different state
        # (runs whose controls all sit in the tile become thread-table fans, whose own records are plain
        # phases -- tests/plan_emulator.check_thread_tables covers those)
        heads = [g.flags for p in passes for g in list(p.gate)[: p.n_gates]]
        n_fans = sum(1 for f in heads if f & pe.GF_FAN_HEADER)
        assert n_fans + sum(1 for f in heads if f & pe.GF_TFAN_HEADER) > 0
        if n_fans:
            wrong = pe.run_plan(passes, n, fast=False)
            assert np.abs(wrong - want).max() > 1e-6


def test_fast_math_needs_corrected_semantics_and_ldg8():
    from qcs_b200 import Circuit
    from qcs_b200.circuit import QcsError
    for kw in ({"semantics": "reference"}, {"semantics": "corrected", "tile_kernel": "tma"},
               {"semantics": "corrected", "tile_kernel": "tma16"}):
        with pytest.raises(QcsError):
            Circuit(12, dryrun=True, math="fast", **kw)
    with pytest.raises(QcsError):
        Circuit(12, dryrun=True, semantics="corrected", math="sloppy")


def _mixed_script(n, length, seed):
    """Every gate kind the host layer can queue, on random qubits (all unitary: the comparison is at 1e-12)."""
    rng = np.random.default_rng(seed)
    s = []
    for _ in range(length):
        q = int(rng.integers(0, n))
        c = int(rng.integers(0, n - 1)); c = c if c < q else c + 1
        a = float(rng.uniform(-3, 3))
        s.append([("h", q), ("x", q), ("y", q), ("z", q), ("phase", q, a), ("rx", q, a), ("ry", q, a), ("rz", q, a),
                  ("cnot", c, q), ("cphase", c, q, a), ("cphase", c, q, -a), ("rz", q, 0.0)][int(rng.integers(0, 12))])
    return s


@pytest.mark.parametrize("seed", range(12))
def test_reordered_plans_on_random_gate_soup(seed):
    """Commutation-aware scheduling must never trade two gates that do not commute: random mixtures of
    pairing, diagonal, controlled and value-preserving gates, 11-14 qubits, every tile size; the plan
    covers every gate exactly once and computes the oracle's state."""
    from qcs_b200 import Circuit
    from tests import plan_emulator as pe
    rng = np.random.default_rng(100 + seed)
    n = int(rng.integers(11, 15))
    script = _mixed_script(n, int(rng.integers(40, 400)), seed)
    tile_kernel, tile_bits = [("ldg8", 10), ("ldg8", 11), ("ldg8", 12), ("ldg", 11)][seed % 4]
    orc = po.Oracle(n, "corrected")
    po.replay(orc, script)
    want = orc.state()
    orc.close()
    c = Circuit(n, dryrun=True, semantics="corrected", tile_kernel=tile_kernel, tile_bits=tile_bits, math="fast",
                peephole="off")
    po.replay(c, script)
    c.flush()
    text, st = c.describe_plan(), c.stats()
    passes = pe.read_plan(c)
    c.close()
    assert sum(p["api"] for p in _parse(text)) == len(script), "every queued gate belongs to exactly one pass"
    _close(pe.run_plan(passes, n, fast=True), want)


@pytest.mark.parametrize("seed", range(12))
def test_exact_reordered_plans_are_bit_exact(seed):
    """The bit-exact mode's schedule (PlannerConfig::reorder_exact) may move X / CNOT past gates they commute
    with and nothing else.  The order the plan executes the gates in, replayed gate by gate on the CPU
    oracle, must give the SAME BITS as the order the circuit was written in -- and through the plan
    emulator the oracle's state."""
    from qcs_b200 import Circuit
    from tests import plan_emulator as pe
    rng = np.random.default_rng(300 + seed)
    n = int(rng.integers(11, 15))
    if seed % 3 == 0:
        n = 16 + seed // 3          # wider than a tile: the in-order plan needs more passes
        script = po.random_circuit_script(n, int(rng.integers(4, 12)), seed=seed)
    else:
        script = []
        for _ in range(int(rng.integers(60, 300))):
            q = int(rng.integers(0, n))
            c = int(rng.integers(0, n - 1)); c = c if c < q else c + 1
            a = float(rng.uniform(-3, 3))
            script.append([("h", q), ("x", q), ("cnot", c, q), ("cnot", c, q), ("rz", q, a), ("ry", q, a),
                           ("cphase", c, q, a)][int(rng.integers(0, 7))])
    kind_of = {"h": "hsym", "x": "swap", "cnot": "swap", "rz": "diag", "ry": "real", "cphase": "diag"}
    key = lambda op: (kind_of[op[0]], op[1] if op[0] in ("cnot", "cphase") else -1, op[2] if op[0] in ("cnot", "cphase") else op[1])
    tile_bits = [10, 11, 12][seed % 3]   # (brickwork cases: 10)
    c = Circuit(n, dryrun=True, semantics="corrected", tile_bits=tile_bits, peephole="off")
    po.replay(c, script)
    c.flush()
    flat = [g for p in _parse(c.describe_plan()) for s in p["segs"] for g in s["gates"]]
    passes = pe.read_plan(c)
    c.close()
    assert sorted(flat) == sorted(key(op) for op in script)
    # k-th occurrence of a (kind, control, target) in the plan = its k-th occurrence in the script
    queues = {}
    for op in script:
        queues.setdefault(key(op), []).append(op)
    taken = {}
    executed = []
    for g in flat:
        k = taken.get(g, 0); taken[g] = k + 1
        executed.append(queues[g][k])
    start = np.random.default_rng(seed).normal(size=2 << n).view(np.complex128)[: 1 << n].copy()
    start /= np.linalg.norm(start)
    states = []
    for order in (script, executed):
        orc = po.Oracle(n, "corrected")
        orc.load_state(start)
        po.replay(orc, order)
        states.append(orc.state())
        orc.close()
    assert np.array_equal(states[0].view(np.uint64), states[1].view(np.uint64)), "the executed order rounds differently"
    _close(pe.run_plan(passes, n, fast=False, state=start.copy()), states[0])
    if seed % 3 == 0:
        assert executed != list(script), "brickwork circuits are where the reordering pays: it must happen here"


# ---- the kernel's own view of a plan: role tables, addresses, case labels (tests/kernel_emulator.py) ---
@pytest.mark.parametrize("math", ["exact", "fast", "fast+reorder"])
@pytest.mark.parametrize("tile_kernel,tile_bits", [("ldg8", 10), ("ldg8", 11), ("ldg8", 12), ("ldg", 10), ("ldg", 11), ("ldg", 12),
                                                    ("tma", 12), ("tma16", 12)])
@pytest.mark.parametrize("case", ["qft", "random+qft", "fans", "generic"])
def test_kernel_model_computes_the_circuit(case, tile_kernel, tile_bits, math):
    """Moves amplitudes through the tile / thread / register roles and decodes case labels exactly as the
    GPU kernel does (from the same tables): a wrong role table, register index, label, csel or tsel is
    a wrong amplitude here."""
    from qcs_b200 import Circuit
    from tests import kernel_emulator as ke
    from tests import plan_emulator as pe
    if math != "exact" and tile_kernel.startswith("tma"):
        pytest.skip("math=fast exists for the plain-load kernels")
    n = 13
    rng = np.random.default_rng(11)
    if case == "qft":
        script = [("qft",)]
    elif case == "random+qft":
        script = po.random_circuit_script(n, 5, seed=99) + [("qft",)]
    elif case == "fans":
        script = _fan_script(n, 5)
    else:
        script = _mixed_script(n, 150, 3)
    init = rng.normal(size=2 ** n) + 1j * rng.normal(size=2 ** n)
    init /= np.linalg.norm(init)
    orc = po.Oracle(n, "corrected")
    orc.load_state(init)
    po.replay(orc, script)
    want = orc.state()
    orc.close()
    reorder = "on" if math.endswith("+reorder") else "off"
    math = math.split("+")[0]
    c = Circuit(n, dryrun=True, semantics="corrected", tile_kernel=tile_kernel, tile_bits=tile_bits, math=math,
                reorder=reorder, peephole="off")
    po.replay(c, script)
    c.flush()
    passes = pe.read_plan(c)
    c.close()
    _close(ke.run_plan(passes, n, fast=(math == "fast"), state=init.copy()), want)


def test_kernel_model_reference_semantics():
    """Row-0-only controlled updates (defect D1) through the kernel model: all but the last gate of a flush
    run fused (the last one alone, after the scratch snapshot), so the state must match the oracle's."""
    from qcs_b200 import Circuit
    from tests import kernel_emulator as ke
    from tests import plan_emulator as pe
    n = 12
    rng = np.random.default_rng(5)
    script = _mixed_script(n, 80, 8)
    init = rng.normal(size=2 ** n) + 1j * rng.normal(size=2 ** n)
    orc = po.Oracle(n, "reference")
    orc.load_state(init)
    po.replay(orc, script[:-1])
    want = orc.state()
    orc.close()
    c = Circuit(n, dryrun=True, semantics="reference", tile_kernel="ldg8")
    po.replay(c, script[:-1] + [("rz", 0, 0.0)])   # a value-preserving last gate keeps the fused part = script[:-1]
    c.flush()
    passes = pe.read_plan(c)
    c.close()
    got = ke.run_plan(passes, n, fast=False, state=init.copy())
    assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("mode", ["exact", "exact-single-swaps", "fast", "fast-in-order"])
@pytest.mark.parametrize("case", ["brickwork", "qft", "soup"])
def test_sharded_kernel_model(case, mode, world):
    """Every rank's plan run through the kernel model, position swaps on the stores of the passes that
    carry them (the element whose bit differs from the rank bit lands in the partner's shard at the
    flipped index): the shards, read under the engine's final layout, are the oracle's state."""
    import ctypes
    from qcs_b200 import Circuit, _ffi
    from tests import kernel_emulator as ke
    from tests import plan_emulator as pe
    _, C = _ffi.load()
    n = {2: 14, 4: 15, 8: 16}[world]
    script = {"brickwork": po.random_circuit_script(n, 7, seed=21),
              "qft": [("h", q) for q in range(n)] + [("ry", n - 1, 0.7), ("qft",)],
              "soup": _mixed_script(n, 160, 17)}[case]
    kw = {"exact": {}, "exact-single-swaps": {"remap_max": 1}, "fast": {"math": "fast"},
          "fast-in-order": {"math": "fast", "reorder": "off"}}[mode]
    plans, layouts = [], []
    try:
        for rank in range(world):
            C.qcs_cuda_dist_finalize()
            assert C.qcs_cuda_dist_init_plan_only(rank, world) == 0, _ffi.last_error()
            c = Circuit(n, dryrun=True, semantics="corrected", tile_kernel="ldg8", peephole="off", **kw)
            po.replay(c, script)
            c.flush()
            entries = []
            for k, p in enumerate(pe.read_plan(c)):
                lpos, gpos = (ctypes.c_int * 3)(-1, -1, -1), (ctypes.c_int * 3)(-1, -1, -1)
                cnt = C.qcs_cuda_last_plan_swap(c.e, k, lpos, gpos)
                entries.append((p, tuple((lpos[i], gpos[i]) for i in range(cnt)) if cnt else None))
            plans.append(entries)
            layouts.append(c.layout())
            c.close()
    finally:
        C.qcs_cuda_dist_finalize()
    assert all(l == layouts[0] for l in layouts), "ranks disagree on the layout"
    assert any(sw is not None for _, sw in plans[0]), "the case must exercise a position swap"
    most = max(len(sw) for _, sw in plans[0] if sw is not None)
    if mode == "exact-single-swaps" or mode == "fast":
        assert most == 1
    elif mode == "exact" and world >= 4 and case != "soup":
        # in-order schedules trade several positions on one pass (an all-to-all among 2^k ranks)
        assert most >= (3 if world == 8 and case == "brickwork" else 2), most
    shards = ke.run_sharded(plans, n, world, fast=not mode.startswith("exact"))
    phys = np.concatenate(shards)
    logical = np.arange(1 << n, dtype=np.int64)
    where = np.zeros_like(logical)
    for q in range(n):
        where |= ((logical >> q) & 1) << layouts[0][q]
    orc = po.Oracle(n, "corrected")
    po.replay(orc, script)
    want = orc.state()
    orc.close()
    _close(phys[where], want)
