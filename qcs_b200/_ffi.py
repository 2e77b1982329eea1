"""ctypes bindings for the two native libraries (no torch types anywhere).

libqcs.so       -> the reference's public C API (include/qcs.h)
libqcs_cuda.so  -> the device ABI (include/qcs_cuda.h)

Loading fails loudly when the libraries are missing: there is no Python or CPU
implementation of the path to fall back to.
"""
from __future__ import annotations

import ctypes
import os

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_DIR = os.path.join(PKG, "lib")

_D = ctypes.c_double
_I = ctypes.c_int
_L = ctypes.c_long
_P = ctypes.c_void_p
_DP = ctypes.POINTER(ctypes.c_double)
_IP = ctypes.POINTER(ctypes.c_int)
_LP = ctypes.POINTER(ctypes.c_long)


class Stats(ctypes.Structure):
    _fields_ = [
        ("gates_submitted", _L), ("gates_executed", _L), ("passes", _L),
        ("kernel_launches", _L), ("segments", _L), ("remaps", _L),
        ("algorithmic_bytes", _D), ("pass_bytes", _D), ("pass_ms", _D),
        ("exchange_bytes", _D), ("exchange_ms", _D),
        ("fused_remaps", ctypes.c_long), ("fused_remap_pass_ms", _D),
        ("pass_flops_per_amp", _D), ("gates_cancelled", ctypes.c_long),
        ("multi_remaps", ctypes.c_long), ("fused_remap_bytes", _D),
        ("out_of_place_remaps", ctypes.c_long),
    ]

    def as_dict(self) -> dict:
        return {k: getattr(self, k) for k, _ in self._fields_}


# name -> (restype, argtypes): every symbol include/qcs_cuda.h declares
CUDA_ABI = {
    "qcs_cuda_state_create": (_I, [ctypes.POINTER(_P), _I]),
    "qcs_cuda_state_destroy": (None, [_P]),
    "qcs_cuda_apply_1q": (_I, [_P, _DP, _I]),
    "qcs_cuda_apply_c1q": (_I, [_P, _DP, _I, _I]),
    "qcs_cuda_phase_flip": (_I, [_P, _L]),
    "qcs_cuda_diffusion": (_I, [_P]),
    "qcs_cuda_normalize": (_I, [_P]),
    "qcs_cuda_prob0": (_I, [_P, _I, _DP]),
    "qcs_cuda_collapse": (_I, [_P, _I, _I]),
    "qcs_cuda_get_amplitude": (_I, [_P, _L, _DP]),
    "qcs_cuda_probability": (_I, [_P, _L, _DP]),
    "qcs_cuda_argmax": (_I, [_P, _LP]),
    "qcs_cuda_sample": (_I, [_P, _DP, _I, _LP]),
    "qcs_cuda_flush": (_I, [_P]),
    "qcs_cuda_read_amplitudes": (_I, [_P, _I, _L, _L, _DP]),
    "qcs_cuda_write_amplitudes": (_I, [_P, _I, _L, _L, _DP]),
    "qcs_cuda_set_default": (_I, [ctypes.c_char_p, ctypes.c_char_p]),
    "qcs_cuda_get_default": (_L, [ctypes.c_char_p, ctypes.c_char_p, _L]),
    "qcs_cuda_trim_pool": (_L, []),
    "qcs_cuda_num_qubits": (_I, [_P]),
    "qcs_cuda_get_layout": (_I, [_P, _IP]),
    "qcs_cuda_get_stats": (_I, [_P, ctypes.POINTER(Stats)]),
    "qcs_cuda_reset_stats": (_I, [_P]),
    "qcs_cuda_set_timing": (_I, [_P, _I]),
    "qcs_cuda_marker_record": (_I, [_P, _I]),
    "qcs_cuda_marker_elapsed_ms": (_I, [_P, _I, _I, _DP]),
    "qcs_cuda_probe_fp64": (_I, [_DP]),
    "qcs_cuda_pass_descriptor_bytes": (_L, []),
    "qcs_cuda_describe_last_plan": (_L, [_P, ctypes.c_char_p, _L]),
    "qcs_cuda_last_plan_raw": (_L, [_P, _L, ctypes.c_void_p, _L]),
    "qcs_cuda_last_plan_tables": (_L, [_P, _L, _DP, _L]),
    "qcs_cuda_last_plan_swap": (_L, [_P, _L, _IP, _IP]),
    "qcs_cuda_last_plan_pass_info": (_L, [_P, _L, _DP, _DP, _IP, _IP]),
    "qcs_cuda_last_error": (ctypes.c_char_p, []),
    "qcs_cuda_dist_unique_id": (_I, [ctypes.c_char_p]),
    "qcs_cuda_dist_init": (_I, [_I, _I, ctypes.c_char_p, _I]),
    "qcs_cuda_dist_finalize": (_I, []),
    "qcs_cuda_dist_init_plan_only": (_I, [_I, _I]),
    "qcs_cuda_trace_read": (_L, [_P, _L, _DP]),
    "qcs_cuda_dist_rank": (_I, []),
    "qcs_cuda_dist_world": (_I, []),
}

# every function include/qcs.h declares (+ the engine accessor of qcs_cuda.h)
HOST_API = {
    "qc_create": (_P, [_I]),
    "qc_destroy": (None, [_P]),
    "qc_h": (None, [_P, _I]), "qc_x": (None, [_P, _I]), "qc_y": (None, [_P, _I]),
    "qc_z": (None, [_P, _I]), "qc_cnot": (None, [_P, _I, _I]),
    "qc_phase": (None, [_P, _I, _D]), "qc_rx": (None, [_P, _I, _D]),
    "qc_ry": (None, [_P, _I, _D]), "qc_rz": (None, [_P, _I, _D]),
    "qc_barrier": (None, [_P]), "qc_reset": (None, [_P, _I]),
    "qc_measure": (_I, [_P, _I]), "qc_measure_all": (None, [_P, _IP]),
    "qc_run": (None, [_P]), "qc_run_shots": (None, [_P, _I, _IP]),
    "qc_find_most_likely_state": (_I, [_P]), "qc_get_probability": (_D, [_P, _I]),
    "qc_print_state": (None, [_P, _I]), "qc_print_circuit": (None, [_P]),
    "qc_grover_search": (None, [_P, _I]), "qc_quantum_fourier_transform": (None, [_P]),
    "qc_bernstein_vazirani": (None, [_P, _I]), "qc_ghz_state": (None, [_P]),
    "qc_get_num_qubits": (_I, [_P]), "qc_get_num_gates": (_I, [_P]),
    "qc_optimize": (None, [_P]),
    "qc_cphase": (None, [_P, _I, _I, _D]),
    "qc_add_gate": (None, [_P, ctypes.c_char_p, _I, _I, _D]),
    "qc_cuda_engine": (_P, [_P]),
    "qc_run_shots_sparse": (None, [_P, _I, _LP]),
}

_cuda = None
_host = None


def _bind(lib, table):
    for name, (res, args) in table.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args


def load():
    """Returns (host_lib, cuda_lib); raises if the native libraries are not built."""
    global _cuda, _host
    if _host is None:
        cuda_path = os.path.join(LIB_DIR, "libqcs_cuda.so")
        host_path = os.path.join(LIB_DIR, "libqcs.so")
        for p in (cuda_path, host_path):
            if not os.path.exists(p):
                raise RuntimeError(
                    f"{p} is missing: build it with `python -m qcs_b200.build` "
                    "(QCS_GPU_CUDA has no CPU fallback)")
        _cuda = ctypes.CDLL(cuda_path)
        _host = ctypes.CDLL(host_path)
        _bind(_cuda, CUDA_ABI)
        _bind(_host, HOST_API)
    return _host, _cuda


def last_error() -> str:
    _, cuda = load()
    return cuda.qcs_cuda_last_error().decode()
