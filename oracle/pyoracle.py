"""oracle/pyoracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes front-ends for the two CPU checkers:

* ``Oracle``  -- our own C restatement (oracle/qcs_oracle.c -> libqcs_oracle.so)
* ``RefLib``  -- the REAL reference compiled from /root/reference into
                 oracle/_ref/libqcsref_<mode>.so by oracle/Makefile (present
                 only after ``make -C oracle ref``; the built .so travels to
                 the GPU box, the sources never enter this repo)

Both expose the same Python surface as ``qcs_b200.Circuit`` (the product), so a
gate script can be replayed on any of the three and the amplitudes compared.
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import
this module.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
RAND_MAX = 2147483647

_libc = ctypes.CDLL(None)
_libc.rand.restype = ctypes.c_int
_libc.srand.argtypes = [ctypes.c_uint]


def srand(seed: int) -> None:
    """Seeds glibc's global rand() stream (shared by oracle, reference, product)."""
    _libc.srand(seed)


def rand_unit() -> float:
    """rand() / (double)RAND_MAX, exactly as src/qcs.c:261 and :596 compute it."""
    return _libc.rand() / float(RAND_MAX)


def build_oracle(force: bool = False) -> str:
    path = os.path.join(HERE, "libqcs_oracle.so")
    src = os.path.join(HERE, "qcs_oracle.c")
    if force or not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    return path


def build_ref() -> None:
    """(Re)builds oracle/_ref from /root/reference when that tree is present."""
    subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


def ref_available(mode: str = "seq") -> bool:
    return os.path.exists(os.path.join(REF_DIR, f"libqcsref_{mode}.so"))


_D = ctypes.c_double
_I = ctypes.c_int
_L = ctypes.c_long
_P = ctypes.c_void_p


def _m8(m) -> ctypes.Array:
    arr = (ctypes.c_double * 8)(*[float(x) for x in m])
    return arr


class Oracle:
    """State-level restatement; semantics = 'reference' | 'corrected'."""

    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            L = ctypes.CDLL(build_oracle())
            L.orc_create.restype = _P
            L.orc_create.argtypes = [_I, _I]
            L.orc_destroy.argtypes = [_P]
            L.orc_size.restype = _L
            L.orc_size.argtypes = [_P]
            L.orc_live.restype = ctypes.POINTER(_D)
            L.orc_live.argtypes = [_P]
            L.orc_scratch.restype = ctypes.POINTER(_D)
            L.orc_scratch.argtypes = [_P]
            L.orc_apply_1q.argtypes = [_P, ctypes.POINTER(_D), _I]
            L.orc_apply_c1q.argtypes = [_P, ctypes.POINTER(_D), _I, _I]
            L.orc_phase_flip.argtypes = [_P, _L]
            L.orc_diffusion.argtypes = [_P]
            L.orc_normalize.argtypes = [_P]
            L.orc_prob0.restype = _D
            L.orc_prob0.argtypes = [_P, _I]
            L.orc_measure.argtypes = [_P, _I, _D]
            L.orc_probability.restype = _D
            L.orc_probability.argtypes = [_P, _L]
            L.orc_argmax.restype = _L
            L.orc_argmax.argtypes = [_P]
            L.orc_sample.argtypes = [_P, ctypes.POINTER(_D), _I, ctypes.POINTER(_I)]
            for g in ("x", "y", "z", "h"):
                getattr(L, f"orc_gate_{g}").argtypes = [ctypes.POINTER(_D)]
            for g in ("p", "rx", "ry", "rz"):
                getattr(L, f"orc_gate_{g}").argtypes = [ctypes.POINTER(_D), _D]
            L.orc_grover_iterations.argtypes = [_I]
            L.orc_grover.argtypes = [_P, _L]
            L.orc_qft.argtypes = [_P]
            L.orc_bv.argtypes = [_P, _I]
            L.orc_ghz.argtypes = [_P]
            cls._lib = L
        return cls._lib

    def __init__(self, n_qubits: int, semantics: str = "reference"):
        self.L = self.lib()
        self.n = n_qubits
        self.semantics = semantics
        self.h_ = self.L.orc_create(n_qubits, 0 if semantics == "reference" else 1)
        if not self.h_:
            raise MemoryError("orc_create failed")

    def close(self):
        if self.h_:
            self.L.orc_destroy(self.h_)
            self.h_ = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- raw access -------------------------------------------------------
    def _view(self, ptr) -> np.ndarray:
        size = self.L.orc_size(self.h_)
        return np.ctypeslib.as_array(ptr, shape=(2 * size,))

    def state(self) -> np.ndarray:
        return self._view(self.L.orc_live(self.h_)).copy().view(np.complex128)

    def scratch(self) -> np.ndarray:
        return self._view(self.L.orc_scratch(self.h_)).copy().view(np.complex128)

    def load_state(self, amps: np.ndarray) -> None:
        self._view(self.L.orc_live(self.h_))[:] = np.asarray(amps, np.complex128).view(np.float64)

    def load_scratch(self, amps: np.ndarray) -> None:
        self._view(self.L.orc_scratch(self.h_))[:] = np.asarray(amps, np.complex128).view(np.float64)

    # --- gates ------------------------------------------------------------
    def _const(self, name, *args):
        m = (ctypes.c_double * 8)()
        getattr(self.L, f"orc_gate_{name}")(m, *args)
        return m

    def apply_1q(self, m, target):
        self.L.orc_apply_1q(self.h_, _m8(m), target)

    def apply_c1q(self, m, control, target):
        self.L.orc_apply_c1q(self.h_, _m8(m), control, target)

    def h(self, q): self.L.orc_apply_1q(self.h_, self._const("h"), q)
    def x(self, q): self.L.orc_apply_1q(self.h_, self._const("x"), q)
    def y(self, q): self.L.orc_apply_1q(self.h_, self._const("y"), q)
    def z(self, q): self.L.orc_apply_1q(self.h_, self._const("z"), q)
    def phase(self, q, a): self.L.orc_apply_1q(self.h_, self._const("p", a), q)
    def rx(self, q, a): self.L.orc_apply_1q(self.h_, self._const("rx", a), q)
    def ry(self, q, a): self.L.orc_apply_1q(self.h_, self._const("ry", a), q)
    def rz(self, q, a): self.L.orc_apply_1q(self.h_, self._const("rz", a), q)
    def cnot(self, c, t): self.L.orc_apply_c1q(self.h_, self._const("x"), c, t)
    def cphase(self, c, t, a): self.L.orc_apply_c1q(self.h_, self._const("p", a), c, t)
    def barrier(self): pass

    def phase_flip(self, idx): self.L.orc_phase_flip(self.h_, idx)
    def diffusion(self): self.L.orc_diffusion(self.h_)
    def normalize(self): self.L.orc_normalize(self.h_)

    # --- measurement ---------------------------------------------------------
    def measure(self, q) -> int:
        if q < 0 or q >= self.n:
            return 0
        return self.L.orc_measure(self.h_, q, rand_unit())

    def measure_all(self):
        return [self.measure(q) for q in range(self.n)]

    def run(self):
        self.measure_all()

    def reset(self, q):
        if self.measure(q) == 1:
            self.x(q)

    def run_shots(self, shots: int) -> np.ndarray:
        size = self.L.orc_size(self.h_)
        res = np.zeros(size, dtype=np.int32)
        if shots <= 0:
            return res
        u = np.array([rand_unit() for _ in range(shots)], dtype=np.float64)
        self.L.orc_sample(self.h_, u.ctypes.data_as(ctypes.POINTER(_D)), shots,
                          res.ctypes.data_as(ctypes.POINTER(_I)))
        return res

    def get_probability(self, idx) -> float:
        return self.L.orc_probability(self.h_, idx)

    def find_most_likely_state(self) -> int:
        return int(self.L.orc_argmax(self.h_))

    # --- drivers ------------------------------------------------------------
    def grover_search(self, sol): self.L.orc_grover(self.h_, sol)
    def qft(self): self.L.orc_qft(self.h_)
    def bv(self, s): self.L.orc_bv(self.h_, s)
    def ghz(self): self.L.orc_ghz(self.h_)


class RefLib:
    """The real reference through its own public API (include/qcs.h:31-74)."""

    _libs: dict = {}

    @classmethod
    def lib(cls, mode: str):
        if mode not in cls._libs:
            path = os.path.join(REF_DIR, f"libqcsref_{mode}.so")
            if not os.path.exists(path):
                raise FileNotFoundError(path)
            L = ctypes.CDLL(path)
            L.qc_create.restype = _P
            L.qc_create.argtypes = [_I]
            L.qc_destroy.argtypes = [_P]
            L.qc_destroy.restype = None
            for f in ("qc_h", "qc_x", "qc_y", "qc_z", "qc_reset"):
                getattr(L, f).argtypes = [_P, _I]
                getattr(L, f).restype = None
            for f in ("qc_phase", "qc_rx", "qc_ry", "qc_rz"):
                getattr(L, f).argtypes = [_P, _I, _D]
                getattr(L, f).restype = None
            L.qc_cnot.argtypes = [_P, _I, _I]
            L.qc_cnot.restype = None
            L.qc_cphase.argtypes = [_P, _I, _I, _D]
            L.qc_cphase.restype = None
            for f in ("qc_barrier", "qc_run", "qc_quantum_fourier_transform",
                      "qc_ghz_state", "qc_optimize", "qcsref_diffusion", "qcsref_normalize"):
                getattr(L, f).argtypes = [_P]
                getattr(L, f).restype = None
            L.qc_measure.argtypes = [_P, _I]
            L.qc_measure.restype = _I
            L.qc_measure_all.argtypes = [_P, ctypes.POINTER(_I)]
            L.qc_measure_all.restype = None
            L.qc_run_shots.argtypes = [_P, _I, ctypes.POINTER(_I)]
            L.qc_run_shots.restype = None
            L.qc_find_most_likely_state.argtypes = [_P]
            L.qc_find_most_likely_state.restype = _I
            L.qc_get_probability.argtypes = [_P, _I]
            L.qc_get_probability.restype = _D
            L.qc_grover_search.argtypes = [_P, _I]
            L.qc_grover_search.restype = None
            L.qc_bernstein_vazirani.argtypes = [_P, _I]
            L.qc_bernstein_vazirani.restype = None
            L.qc_get_num_qubits.argtypes = [_P]
            L.qc_get_num_gates.argtypes = [_P]
            L.qcsref_state_size.argtypes = [_P]
            L.qcsref_state_size.restype = _L
            for f in ("qcsref_copy_state", "qcsref_copy_scratch",
                      "qcsref_load_state", "qcsref_load_scratch"):
                getattr(L, f).argtypes = [_P, ctypes.POINTER(_D)]
                getattr(L, f).restype = None
            L.qcsref_phase_flip.argtypes = [_P, _I]
            L.qcsref_phase_flip.restype = None
            L.qcsref_apply_1q.argtypes = [_P, ctypes.POINTER(_D), _I]
            L.qcsref_apply_1q.restype = None
            L.qcsref_apply_c1q.argtypes = [_P, ctypes.POINTER(_D), _I, _I]
            L.qcsref_apply_c1q.restype = None
            L.qcsref_history_size.argtypes = [_P]
            L.qcsref_history_name.argtypes = [_P, _I]
            L.qcsref_history_name.restype = ctypes.c_char_p
            L.qcsref_history_target.argtypes = [_P, _I]
            L.qcsref_history_control.argtypes = [_P, _I]
            L.qcsref_history_param.argtypes = [_P, _I]
            L.qcsref_history_param.restype = _D
            cls._libs[mode] = L
        return cls._libs[mode]

    def __init__(self, n_qubits: int, mode: str = "seq"):
        """mode: seq | corrected | simd | omp | mt  (oracle/Makefile targets)."""
        self.L = self.lib(mode)
        self.n = n_qubits
        self.c = self.L.qc_create(n_qubits)
        if not self.c:
            raise MemoryError("qc_create failed")

    def close(self):
        if self.c:
            self.L.qc_destroy(self.c)
            self.c = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _buf(self):
        size = self.L.qcsref_state_size(self.c)
        return np.zeros(2 * size, dtype=np.float64)

    def state(self) -> np.ndarray:
        b = self._buf()
        self.L.qcsref_copy_state(self.c, b.ctypes.data_as(ctypes.POINTER(_D)))
        return b.view(np.complex128)

    def scratch(self) -> np.ndarray:
        b = self._buf()
        self.L.qcsref_copy_scratch(self.c, b.ctypes.data_as(ctypes.POINTER(_D)))
        return b.view(np.complex128)

    def load_state(self, amps):
        b = np.ascontiguousarray(np.asarray(amps, np.complex128)).view(np.float64)
        self.L.qcsref_load_state(self.c, b.ctypes.data_as(ctypes.POINTER(_D)))

    def load_scratch(self, amps):
        b = np.ascontiguousarray(np.asarray(amps, np.complex128)).view(np.float64)
        self.L.qcsref_load_scratch(self.c, b.ctypes.data_as(ctypes.POINTER(_D)))

    def apply_1q(self, m, target): self.L.qcsref_apply_1q(self.c, _m8(m), target)
    def apply_c1q(self, m, control, target): self.L.qcsref_apply_c1q(self.c, _m8(m), control, target)

    def h(self, q): self.L.qc_h(self.c, q)
    def x(self, q): self.L.qc_x(self.c, q)
    def y(self, q): self.L.qc_y(self.c, q)
    def z(self, q): self.L.qc_z(self.c, q)
    def phase(self, q, a): self.L.qc_phase(self.c, q, a)
    def rx(self, q, a): self.L.qc_rx(self.c, q, a)
    def ry(self, q, a): self.L.qc_ry(self.c, q, a)
    def rz(self, q, a): self.L.qc_rz(self.c, q, a)
    def cnot(self, c, t): self.L.qc_cnot(self.c, c, t)
    def cphase(self, c, t, a): self.L.qc_cphase(self.c, c, t, a)
    def barrier(self): self.L.qc_barrier(self.c)
    def reset(self, q): self.L.qc_reset(self.c, q)
    def phase_flip(self, idx): self.L.qcsref_phase_flip(self.c, idx)
    def diffusion(self): self.L.qcsref_diffusion(self.c)
    def normalize(self): self.L.qcsref_normalize(self.c)

    def measure(self, q) -> int: return self.L.qc_measure(self.c, q)

    def measure_all(self):
        r = (ctypes.c_int * self.n)()
        self.L.qc_measure_all(self.c, r)
        return list(r)

    def run(self): self.L.qc_run(self.c)

    def run_shots(self, shots: int) -> np.ndarray:
        size = self.L.qcsref_state_size(self.c)
        res = np.zeros(size, dtype=np.int32)
        self.L.qc_run_shots(self.c, shots, res.ctypes.data_as(ctypes.POINTER(_I)))
        return res

    def get_probability(self, idx) -> float: return self.L.qc_get_probability(self.c, idx)
    def find_most_likely_state(self) -> int: return self.L.qc_find_most_likely_state(self.c)
    def grover_search(self, sol): self.L.qc_grover_search(self.c, sol)
    def qft(self): self.L.qc_quantum_fourier_transform(self.c)
    def bv(self, s): self.L.qc_bernstein_vazirani(self.c, s)
    def ghz(self): self.L.qc_ghz_state(self.c)
    def optimize(self): self.L.qc_optimize(self.c)

    @property
    def num_qubits(self): return self.L.qc_get_num_qubits(self.c)

    @property
    def num_gates(self): return self.L.qc_get_num_gates(self.c)

    def history(self):
        out = []
        for i in range(self.L.qcsref_history_size(self.c)):
            out.append((self.L.qcsref_history_name(self.c, i).decode(),
                        self.L.qcsref_history_target(self.c, i),
                        self.L.qcsref_history_control(self.c, i),
                        self.L.qcsref_history_param(self.c, i)))
        return out


# ---------------------------------------------------------------------------
# Gate scripts: a list of tuples replayable on Oracle / RefLib / product.
# ---------------------------------------------------------------------------

# The scripts and their driver are workload definitions, not oracle arithmetic: they live with the
# product (qcs_b200/workloads.py) and are re-exported here for the tests.
from qcs_b200.workloads import random_circuit_script, replay, splitmix64  # noqa: E402,F401
