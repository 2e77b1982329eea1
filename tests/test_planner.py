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
        # a pass runs on the smallest tile (>= 10 bits for ldg8, else 12) that holds its pairing qubits
        lo = 10 if tile_kernel == "ldg8" else 12
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


def test_gate_order_is_preserved():
    n = 16
    script = po.random_circuit_script(n, 3, seed=5)
    text, _ = _plan(n, script, semantics="corrected")
    flat = [g for p in _parse(text) for s in p["segs"] for g in s["gates"]]
    want = []
    for op in script:
        if op[0] == "h": want.append(("hsym", -1, op[1]))
        elif op[0] == "rz": want.append(("diag", -1, op[1]))
        elif op[0] == "cnot": want.append(("swap", op[1], op[2]))
    assert flat == want


def test_qft30_plan_is_compact():
    text, st = _plan(30, [("qft",)], semantics="corrected")
    passes = _parse(text)
    assert st["gates_submitted"] == 465
    assert len(passes) <= 5, len(passes)            # 30 pairing targets, 12 then 7 new positions per pass
    assert all(p["api"] <= 240 for p in passes)     # QCS_MAX_PASS_GATES


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
