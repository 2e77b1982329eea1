"""GPU parity of the opt-in math=fast mode (fused multiply-adds, controlled-phase fans collapsed
into one factor per thread) against the CPU oracle.

Bar: the north star's "within 1e-12 relative (fp64)" in the metric of SURVEY 8(c): per component
|delta| <= 1e-12 * max|a_ref|, and relative 1e-12 wherever |a_ref| >= 1e-3 * max|a_ref|.  The
default mode (math=exact) is held to bit-exactness by tests/test_gpu_parity.py; this mode trades
that for FP64 issue slots and says so (DESIGN.md "math=fast").
"""
import numpy as np
import pytest

from oracle import pyoracle as po
from tests.test_planner import _fan_script

pytestmark = pytest.mark.gpu

TOL = 1e-12


def Circuit(*a, **k):
    from qcs_b200 import Circuit as C
    k.setdefault("tile_kernel", "ldg8")
    return C(*a, semantics="corrected", math="fast", **k)


def _unitary(rng):
    """Random 2x2 unitary, row-major {re, im} x 4.  Tolerance tests use well-conditioned gates: a
    non-unitary matrix amplifies rounding differences by its condition number per application,
    which says nothing about the kernel."""
    q, r = np.linalg.qr(rng.normal(size=(2, 2)) + 1j * rng.normal(size=(2, 2)))
    q = q * (np.diag(r) / np.abs(np.diag(r)))
    return [float(v) for z in q.reshape(4) for v in (z.real, z.imag)]


def _random_script(rng, n, length):
    s = []
    for _ in range(length):
        k = int(rng.integers(0, 12))
        q = int(rng.integers(0, n))
        c = int(rng.integers(0, n - 1)); c = c if c < q else c + 1
        ang = float(rng.uniform(-3, 3))
        s.append([("h", q), ("x", q), ("y", q), ("z", q), ("phase", q, ang), ("rx", q, ang),
                  ("ry", q, ang), ("rz", q, ang), ("cnot", c, q), ("cphase", c, q, ang),
                  ("apply_1q", _unitary(rng), q), ("apply_c1q", _unitary(rng), c, q)][k])
    return s


def _check(got, want, what=""):
    scale = np.abs(want).max()
    err = np.abs(got - want)
    assert err.max() <= TOL * scale, f"{what}: abs err {err.max() / scale:.3e} of the largest amplitude"
    big = np.abs(want) >= 1e-3 * scale
    rel = (err[big] / np.abs(want[big])).max()
    assert rel <= TOL, f"{what}: rel err {rel:.3e}"
    return err.max() / scale


@pytest.mark.parametrize("reorder", ["on", "off"])
@pytest.mark.parametrize("tile_bits", [10, 11, 12])
@pytest.mark.parametrize("n", [10, 11, 12, 13, 15, 17, 21])
def test_fast_random_circuits_within_tolerance(n, tile_bits, reorder):
    rng = np.random.default_rng(4000 + n)
    for trial in range(3 if n < 20 else 1):
        script = _random_script(rng, n, 60 + 40 * trial)
        init = rng.normal(size=2 ** n) + 1j * rng.normal(size=2 ** n)
        orc = po.Oracle(n, "corrected"); c = Circuit(n, tile_bits=tile_bits, reorder=reorder)
        orc.load_state(init); c.load_state(init)
        po.replay(orc, script); po.replay(c, script)
        _check(c.state(), orc.state(), f"n={n} trial={trial}\n{c.describe_plan()[:1500]}")
        st = c.stats()
        assert 0 < st["passes"] < len(script), st
        orc.close(); c.close()


@pytest.mark.parametrize("reorder", ["on", "off"])
@pytest.mark.parametrize("tile_bits", [10, 11, 12])
@pytest.mark.parametrize("n", [12, 14, 16, 20])
def test_fast_qft_and_fans(n, tile_bits, reorder):
    """Dense start state so every controlled phase acts; fans of all lengths, consecutive and scattered controls."""
    rng = np.random.default_rng(n)
    init = rng.normal(size=2 ** n) + 1j * rng.normal(size=2 ** n)
    init /= np.linalg.norm(init)
    for name, script in (("qft", [("qft",)]), ("qft.qft", [("qft",), ("qft",)]), ("fans", _fan_script(n, n))):
        orc = po.Oracle(n, "corrected"); c = Circuit(n, tile_bits=tile_bits, reorder=reorder)
        orc.load_state(init); c.load_state(init)
        po.replay(orc, script); po.replay(c, script)
        _check(c.state(), orc.state(), f"{name} n={n}\n{c.describe_plan()[:1500]}")
        orc.close(); c.close()


@pytest.mark.parametrize("tile_bits", [10, 12])
def test_fast_every_target_control_pair(tile_bits):
    n = 14
    rng = np.random.default_rng(7)
    init = rng.normal(size=2 ** n) + 1j * rng.normal(size=2 ** n)
    h = 1.0 / np.sqrt(2.0)
    mats = {  # one unitary per arithmetic class of the interpreter (common.h GateKind)
        "generic": _unitary(rng),
        "real": [0.6, 0, -0.8, 0, 0.8, 0, 0.6, 0],
        "hsym": [h, 0, -h, 0, h, 0, h, 0],
        "swap": [0, 0, 1, 0, 1, 0, 0, 0],
        "diag": [np.cos(0.4), np.sin(0.4), 0, 0, 0, 0, np.cos(-1.3), np.sin(-1.3)],
        "phase": [1, 0, 0, 0, 0, 0, 0.28, 0.96],
    }
    for cls, m in mats.items():
        orc = po.Oracle(n, "corrected"); c = Circuit(n, tile_bits=tile_bits)
        orc.load_state(init); c.load_state(init)
        for t in range(n):
            orc.apply_1q(m, t); c.apply_1q(m, t)
            for ctl in range(n):
                if ctl != t and (ctl + t) % 3 == 0:
                    orc.apply_c1q(m, ctl, t); c.apply_c1q(m, ctl, t)
        _check(c.state(), orc.state(), cls)
        orc.close(); c.close()


@pytest.mark.parametrize("reorder", ["on", "off"])
@pytest.mark.parametrize("tile_bits", [10, 11, 12])
@pytest.mark.parametrize("n", [11, 12, 14, 17, 21])
def test_fast_16_amplitudes_per_thread(n, tile_bits, reorder):
    """tile_kernel=ldg under math=fast: 16 amplitudes per thread, four pairing positions per segment."""
    if n < tile_bits:
        pytest.skip("shard smaller than the tile")
    rng = np.random.default_rng(7000 + n)
    init = rng.normal(size=2 ** n) + 1j * rng.normal(size=2 ** n)
    init /= np.linalg.norm(init)
    for name, script in (("random", _random_script(rng, n, 120)), ("qft", [("qft",)]), ("fans", _fan_script(n, n)),
                         ("brickwork", po.random_circuit_script(n, 6))):
        orc = po.Oracle(n, "corrected"); c = Circuit(n, tile_kernel="ldg", tile_bits=tile_bits, reorder=reorder)
        orc.load_state(init); c.load_state(init)
        po.replay(orc, script); po.replay(c, script)
        _check(c.state(), orc.state(), f"{name} n={n}\n{c.describe_plan()[:1500]}")
        orc.close(); c.close()


def test_fast_brickwork_deep_passes():
    """BASELINE config 5's circuit shape at 24 qubits, depth 12: the reordered plan (a pass follows its
    tile through several layers) needs under half the passes of the in-order plan and computes the
    same state within tolerance."""
    n = 24
    script = po.random_circuit_script(n, 12)
    orc = po.Oracle(n, "corrected")
    po.replay(orc, script)
    want = orc.state()
    orc.close()
    passes = {}
    for reorder in ("on", "off"):
        c = Circuit(n, reorder=reorder)
        po.replay(c, script)
        _check(c.state(), want, f"reorder={reorder}")
        passes[reorder] = c.stats()["passes"]
        c.close()
    assert passes["on"] * 2 < passes["off"], passes


def test_fast_measurement_and_shots_follow_the_state():
    """Sampling and collapse are the exact-sum kernels in both modes: on the SAME device state they give
    the oracle's outcomes.  (Seeded outcomes can differ from the reference's only where a draw lands
    within ~1e-15 of a CDF boundary; on this circuit none does.)"""
    n = 14
    script = po.random_circuit_script(n, 4, seed=31) + [("qft",)]
    orc = po.Oracle(n, "corrected"); c = Circuit(n)
    po.replay(orc, script); po.replay(c, script)
    _check(c.state(), orc.state())
    po.srand(5); want = orc.run_shots(3000); want_m = orc.measure_all()
    po.srand(5); got = c.run_shots(3000); got_m = c.measure_all()
    assert np.array_equal(got, want) and got_m == want_m
    orc.close(); c.close()


def test_fast_qft30_uniform_and_roundtrip():
    """BASELINE config 3 at full size in math=fast: uniform amplitudes within 1e-12, QFT then inverse returns the input."""
    import ctypes
    import math
    n = 30
    c = Circuit(n)
    c.qft()
    a = 2.0 ** (-n / 2)
    for idx in (0, 1, 12345, 2 ** 29 + 7, 2 ** 30 - 1):
        out = (ctypes.c_double * 2)()
        assert c.C.qcs_cuda_get_amplitude(c.e, idx, out) == 0
        assert abs(out[0] - a) <= TOL * a and abs(out[1]) <= TOL * a, (idx, out[0], out[1])
    c.close()
    c = Circuit(n)
    x = 0x2F0F3A71 & ((1 << n) - 1)
    for q in range(n):
        if (x >> q) & 1:
            c.x(q)
    c.qft()
    for i in reversed(range(n)):
        for j in reversed(range(i + 1, n)):
            c.cphase(j, i, -math.pi / float(1 << (j - i)))
        c.h(i)
    assert abs(c.get_probability(x) - 1.0) < 1e-12
    assert c.find_most_likely_state() == x
    c.close()
