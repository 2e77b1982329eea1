"""Synthetic workloads of BASELINE.json's configs and a driver that replays a gate script on any
backend exposing the qcs.h vocabulary (qcs_b200.Circuit; the test oracle uses the same scripts).

A script is a list of tuples: ("h", q), ("rz", q, theta), ("cnot", c, t), ("cphase", c, t, a),
("qft",), ("measure", q), ("run_shots", n), ("srand", seed), ...  Pure Python, no arithmetic on
amplitudes: product-side code (bench.py, scripts/) takes its workloads from here, never from oracle/.
"""
from __future__ import annotations

import ctypes
import math

_libc = ctypes.CDLL(None)
_libc.srand.argtypes = [ctypes.c_uint]
_libc.srand.restype = None


def srand(seed: int) -> None:
    """Seeds glibc's global rand() stream: the host layer draws from it exactly where the reference
    does (src/qcs.c:261, :596), so a user program's srand() behaves the same."""
    _libc.srand(seed)


def replay(backend, script) -> list:
    """Runs `script` on `backend`; returns the values the ops produced (measurement outcomes, shot
    histograms, argmax, probabilities) as (kind, value) pairs."""
    out = []
    for op in script:
        name, args = op[0], op[1:]
        if name == "srand":
            srand(args[0])
        elif name == "measure":
            out.append(("measure", backend.measure(*args)))
        elif name == "measure_all":
            out.append(("measure_all", tuple(backend.measure_all())))
        elif name == "run_shots":
            out.append(("run_shots", backend.run_shots(*args)))
        elif name == "argmax":
            out.append(("argmax", backend.find_most_likely_state()))
        elif name == "prob":
            out.append(("prob", backend.get_probability(*args)))
        elif name == "grover":
            backend.grover_search(*args)
        else:
            getattr(backend, name)(*args)
    return out


def splitmix64(seed: int):
    """Deterministic generator shared by every harness: no libc RNG, so CPU and
    GPU harnesses build the same circuit."""
    state = seed & 0xFFFFFFFFFFFFFFFF
    while True:
        state = (state + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        z = state
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        yield z ^ (z >> 31)


def random_circuit_script(n: int, depth: int, seed: int = 0x51C50034) -> list:
    """BASELINE config 5: per layer an H or an RZ(theta) on every qubit, then a
    brickwork of CNOT(q, q+1) on alternating pairs."""
    g = splitmix64(seed)
    script = []
    for layer in range(depth):
        for q in range(n):
            r = next(g)
            if r & 1:
                script.append(("h", q))
            else:
                theta = 2.0 * math.pi * ((next(g) >> 11) * 2.0 ** -53)
                script.append(("rz", q, theta))
        for q in range(layer % 2, n - 1, 2):
            script.append(("cnot", q, q + 1))
    return script
