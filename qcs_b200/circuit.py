"""Python mirror of the reference's circuit API (include/qcs.h), over libqcs.so.

Method names follow the C functions minus the `qc_` prefix, so a test written
against the reference's test-suite reads the same here:

    c = Circuit(2); c.x(0); c.cnot(0, 1); assert c.find_most_likely_state() == 3

Every method is one call through the C API; no arithmetic happens in Python.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _ffi


class QcsError(RuntimeError):
    pass


def set_default(key: str, value: str | None) -> None:
    """Engine option for circuits created afterwards (semantics, fusion, dryrun, pass_flops)."""
    _, cuda = _ffi.load()
    cuda.qcs_cuda_set_default(key.encode(), None if value is None else str(value).encode())


def get_default(key: str) -> str | None:
    """The process-wide default set by set_default(), or None."""
    _, cuda = _ffi.load()
    buf = ctypes.create_string_buffer(256)
    n = cuda.qcs_cuda_get_default(key.encode(), buf, 256)
    return None if n < 0 else buf.value.decode()


class Circuit:
    def __init__(self, num_qubits: int, *, semantics: str | None = None,
                 fusion: str | None = None, dryrun: bool | None = None,
                 pass_flops: float | None = None, tile_kernel: str | None = None,
                 exchange: str | None = None, fuse_swaps: str | None = None,
                 peephole: str | None = None, fixed_low: int | None = None,
                 tile_bits: int | None = None, math: str | None = None,
                 reorder: str | None = None, reorder_segments: int | None = None,
                 swap_store: str | None = None, lazy_init: str | None = None,
                 fuse_argmax: str | None = None, remap_max: int | None = None,
                 victim_policy: str | None = None, fan_tables: str | None = None,
                 remap_buffer: str | None = None, tile_search: str | None = None):
        self.H, self.C = _ffi.load()
        opts = {"semantics": semantics, "fusion": fusion, "tile_kernel": tile_kernel, "exchange": exchange, "fuse_swaps": fuse_swaps, "peephole": peephole, "math": math, "reorder": reorder, "swap_store": swap_store, "lazy_init": lazy_init, "fuse_argmax": fuse_argmax, "victim_policy": victim_policy, "fan_tables": fan_tables, "remap_buffer": remap_buffer, "tile_search": tile_search,
                "remap_max": None if remap_max is None else str(int(remap_max)),
                "reorder_segments": None if reorder_segments is None else str(int(reorder_segments)), "fixed_low": None if fixed_low is None else str(int(fixed_low)),
                "tile_bits": None if tile_bits is None else str(int(tile_bits)),
                "dryrun": None if dryrun is None else ("1" if dryrun else "0"),
                "pass_flops": None if pass_flops is None else repr(float(pass_flops))}
        # per-circuit options travel as process-wide defaults around qc_create (the C API has no
        # configuration object); whatever defaults the caller had set before are put back afterwards
        saved = {k: get_default(k) for k, v in opts.items() if v is not None}
        for k, v in opts.items():
            if v is not None:
                set_default(k, v)
        try:
            self.c = self.H.qc_create(num_qubits)
        finally:
            for k, v in saved.items():
                set_default(k, v)
        if not self.c:
            raise QcsError(f"qc_create({num_qubits}) failed: {_ffi.last_error()}")
        self.e = self.H.qc_cuda_engine(self.c)
        self.n = num_qubits

    # -- lifetime ---------------------------------------------------------------
    def close(self):
        if getattr(self, "c", None):
            self.H.qc_destroy(self.c)
            self.c = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _ck(self, rc):
        if rc != 0:
            raise QcsError(_ffi.last_error())

    # -- gates (include/qcs.h) -----------------------------------------------------
    def h(self, q): self.H.qc_h(self.c, q)
    def x(self, q): self.H.qc_x(self.c, q)
    def y(self, q): self.H.qc_y(self.c, q)
    def z(self, q): self.H.qc_z(self.c, q)
    def cnot(self, control, target): self.H.qc_cnot(self.c, control, target)
    def phase(self, q, angle): self.H.qc_phase(self.c, q, angle)
    def rx(self, q, angle): self.H.qc_rx(self.c, q, angle)
    def ry(self, q, angle): self.H.qc_ry(self.c, q, angle)
    def rz(self, q, angle): self.H.qc_rz(self.c, q, angle)
    def cphase(self, control, target, angle): self.H.qc_cphase(self.c, control, target, angle)
    def barrier(self): self.H.qc_barrier(self.c)
    def reset(self, q): self.H.qc_reset(self.c, q)

    # -- measurement / execution ---------------------------------------------------
    def measure(self, q) -> int: return self.H.qc_measure(self.c, q)

    def measure_all(self):
        r = (ctypes.c_int * self.n)()
        self.H.qc_measure_all(self.c, r)
        return list(r)

    def run(self): self.H.qc_run(self.c)

    def run_shots(self, shots: int) -> np.ndarray:
        res = np.zeros(1 << self.n, dtype=np.int32)
        self.H.qc_run_shots(self.c, shots, res.ctypes.data_as(ctypes.POINTER(ctypes.c_int)))
        return res

    def run_shots_sparse(self, shots: int) -> np.ndarray:
        """Outcome of every shot (-1: dropped by the reference's rule), no dense histogram."""
        idx = np.empty(shots, dtype=np.int64)
        self.H.qc_run_shots_sparse(self.c, shots, idx.ctypes.data_as(ctypes.POINTER(ctypes.c_long)))
        return idx

    # -- state access -----------------------------------------------------------------
    def find_most_likely_state(self) -> int: return self.H.qc_find_most_likely_state(self.c)
    def get_probability(self, state: int) -> float: return self.H.qc_get_probability(self.c, state)
    def print_state(self, solution_index: int = -1): self.H.qc_print_state(self.c, solution_index)
    def print_circuit(self): self.H.qc_print_circuit(self.c)

    # -- algorithm drivers ---------------------------------------------------------------
    def grover_search(self, solution: int): self.H.qc_grover_search(self.c, solution)
    def quantum_fourier_transform(self): self.H.qc_quantum_fourier_transform(self.c)
    def bernstein_vazirani(self, hidden: int): self.H.qc_bernstein_vazirani(self.c, hidden)
    def ghz_state(self): self.H.qc_ghz_state(self.c)
    qft = quantum_fourier_transform
    bv = bernstein_vazirani
    ghz = ghz_state

    # -- introspection ---------------------------------------------------------------------
    @property
    def num_qubits(self) -> int: return self.H.qc_get_num_qubits(self.c)

    @property
    def num_gates(self) -> int: return self.H.qc_get_num_gates(self.c)

    def optimize(self): self.H.qc_optimize(self.c)

    # -- engine-level entry points (include/qcs_cuda.h) -----------------------------------------
    def apply_1q(self, m, target):
        arr = (ctypes.c_double * 8)(*[float(v) for v in m])
        return self.C.qcs_cuda_apply_1q(self.e, arr, target)

    def apply_c1q(self, m, control, target):
        arr = (ctypes.c_double * 8)(*[float(v) for v in m])
        return self.C.qcs_cuda_apply_c1q(self.e, arr, control, target)

    def phase_flip(self, index): return self.C.qcs_cuda_phase_flip(self.e, index)
    def diffusion(self): self._ck(self.C.qcs_cuda_diffusion(self.e))
    def normalize(self): self._ck(self.C.qcs_cuda_normalize(self.e))
    def flush(self): self._ck(self.C.qcs_cuda_flush(self.e))

    def prob0(self, q) -> float:
        p = ctypes.c_double()
        self._ck(self.C.qcs_cuda_prob0(self.e, q, ctypes.byref(p)))
        return p.value

    def _shard(self):
        world = self.C.qcs_cuda_dist_world()
        rank = self.C.qcs_cuda_dist_rank()
        local = (1 << self.n) // world
        return rank * local, local

    def state(self, which: int = 0) -> np.ndarray:
        """This rank's shard of the live (which=0) or scratch (which=1) buffer, complex128."""
        first, count = self._shard()
        out = np.empty(2 * count, dtype=np.float64)
        self._ck(self.C.qcs_cuda_read_amplitudes(
            self.e, which, first, count, out.ctypes.data_as(ctypes.POINTER(ctypes.c_double))))
        return out.view(np.complex128)

    def scratch(self) -> np.ndarray:
        return self.state(1)

    def load_state(self, amps, which: int = 0):
        first, count = self._shard()
        buf = np.ascontiguousarray(np.asarray(amps, dtype=np.complex128)).view(np.float64)
        assert buf.size == 2 * count
        self._ck(self.C.qcs_cuda_write_amplitudes(
            self.e, which, first, count, buf.ctypes.data_as(ctypes.POINTER(ctypes.c_double))))

    def load_scratch(self, amps):
        self.load_state(amps, 1)

    def sample_indices(self, u) -> np.ndarray:
        u = np.ascontiguousarray(u, dtype=np.float64)
        idx = np.empty(u.size, dtype=np.int64)
        self._ck(self.C.qcs_cuda_sample(
            self.e, u.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), u.size,
            idx.ctypes.data_as(ctypes.POINTER(ctypes.c_long))))
        return idx

    def marker(self, slot: int): self._ck(self.C.qcs_cuda_marker_record(self.e, slot))

    def marker_elapsed_ms(self, a: int, b: int) -> float:
        ms = ctypes.c_double()
        self._ck(self.C.qcs_cuda_marker_elapsed_ms(self.e, a, b, ctypes.byref(ms)))
        return ms.value

    def set_timing(self, enabled: bool): self._ck(self.C.qcs_cuda_set_timing(self.e, int(enabled)))

    def stats(self) -> dict:
        st = _ffi.Stats()
        self._ck(self.C.qcs_cuda_get_stats(self.e, ctypes.byref(st)))
        return st.as_dict()

    def reset_stats(self): self._ck(self.C.qcs_cuda_reset_stats(self.e))

    def layout(self) -> list:
        """perm[q] = physical index position of logical qubit q."""
        perm = (ctypes.c_int * self.n)()
        self._ck(self.C.qcs_cuda_get_layout(self.e, perm))
        return list(perm)

    def trace(self) -> list:
        """Dry-run engines: the ordered list of physical gates / position swaps a real engine would run."""
        out = []
        buf = (ctypes.c_double * 12)()
        n = self.C.qcs_cuda_trace_read(self.e, -1, buf)
        for i in range(n):
            self.C.qcs_cuda_trace_read(self.e, i, buf)
            if buf[0] == 2.0:
                out.append(("swap", int(buf[1]), int(buf[2])))
            elif buf[0] == 3.0:
                out.append(("swap_local", int(buf[1]), int(buf[2])))
            else:
                kf = int(buf[1])
                out.append(("gate", kf & 0xff, kf >> 8, int(buf[2]), int(buf[3]), [buf[4 + k] for k in range(8)]))
        return out

    def pass_info(self) -> list:
        """Passes of the last flush: dicts with ms (device time, timing on), flops_per_amp, tile_bits, segments."""
        out = []
        n = self.C.qcs_cuda_last_plan_pass_info(self.e, -1, None, None, None, None)
        for k in range(n):
            ms, fl = ctypes.c_double(), ctypes.c_double()
            tb, sg = ctypes.c_int(), ctypes.c_int()
            self.C.qcs_cuda_last_plan_pass_info(self.e, k, ctypes.byref(ms), ctypes.byref(fl), ctypes.byref(tb), ctypes.byref(sg))
            out.append({"ms": ms.value, "flops_per_amp": fl.value, "tile_bits": tb.value, "segments": sg.value})
        return out

    def pass_times(self) -> list:
        return [p["ms"] for p in self.pass_info()]

    def describe_plan(self) -> str:
        need = self.C.qcs_cuda_describe_last_plan(self.e, None, 0)
        buf = ctypes.create_string_buffer(int(need) + 1)
        self.C.qcs_cuda_describe_last_plan(self.e, buf, need + 1)
        return buf.value.decode()
