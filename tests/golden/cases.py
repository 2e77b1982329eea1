"""Gate scripts shared by the golden-vector generator and the parity tests.

Each case is (n_qubits, script); a script is a list of tuples whose first item
is a method name common to oracle.pyoracle.Oracle / RefLib and
qcs_b200.Circuit (see oracle.pyoracle.replay).  The first block reproduces the
known-answer anchors listed in SURVEY.md section 8(c).
"""
import math

from qcs_b200.workloads import random_circuit_script  # workload definition, not oracle arithmetic


def _mixed8():
    s = []
    for q in range(8):
        s.append(("h", q))
    s += [("rx", 0, 0.3), ("ry", 1, 1.1), ("rz", 2, -0.7), ("phase", 3, 2.2),
          ("y", 4), ("z", 5), ("x", 6), ("cnot", 0, 7), ("cnot", 7, 3),
          ("cphase", 2, 5, 0.9), ("cphase", 6, 1, -1.3), ("ry", 7, 0.4),
          ("cnot", 3, 4), ("rx", 5, 2.5), ("barrier",), ("h", 2), ("cnot", 5, 0),
          ("rz", 7, 1.9), ("cphase", 0, 7, 0.25), ("y", 1), ("h", 6)]
    return s


CASES = {
    # --- SURVEY 8(c) anchors -------------------------------------------------
    "kat_h2": (2, [("h", 0), ("h", 1)]),
    "kat_cnot": (2, [("x", 0), ("cnot", 0, 1), ("argmax",)]),
    "kat_ghz2": (2, [("ghz",)]),
    "kat_qft3_x0": (3, [("x", 0), ("qft",)]),
    "kat_grover_3_5": (3, [("grover", 5), ("prob", 5)]),
    "kat_grover_4_5": (4, [("grover", 5), ("prob", 5)]),
    "kat_grover_5_5": (5, [("grover", 5), ("prob", 5)]),
    "kat_grover_6_42": (6, [("grover", 42), ("prob", 42)]),
    "kat_grover_8_5": (8, [("grover", 5), ("prob", 5)]),
    "kat_grover_10_5": (10, [("grover", 5), ("prob", 5), ("argmax",)]),
    "kat1_shots_measure": (3, [("srand", 42), ("h", 0), ("h", 1), ("ry", 2, 0.9),
                               ("run_shots", 1000), ("measure", 0), ("measure", 2),
                               ("measure", 1)]),
    "kat3_ghz_shots": (2, [("srand", 7), ("h", 0), ("cnot", 0, 1), ("run_shots", 1000)]),
    "kat4_run": (3, [("srand", 123), ("h", 0), ("h", 1), ("h", 2), ("run",),
                     ("argmax",), ("prob", 2)]),
    # --- wider coverage ---------------------------------------------------------
    "ghz5": (5, [("ghz",), ("argmax",)]),
    "bv6": (6, [("bv", 0b10110), ("argmax",)]),
    "bv9": (9, [("bv", 0b10110101), ("argmax",), ("srand", 9), ("run_shots", 500)]),
    "mixed8": (8, _mixed8()),
    "mixed8_measured": (8, [("srand", 2024)] + _mixed8() +
                        [("measure", 3), ("measure", 0), ("reset", 5), ("rx", 5, 0.7),
                         ("run_shots", 300), ("measure_all",)]),
    "qft10_prepared": (10, [("x", 1), ("ry", 3, 0.8), ("h", 7), ("cnot", 7, 2),
                            ("rz", 9, 0.6), ("qft",)]),
    "random10_d6": (10, random_circuit_script(10, 6)),
    "random12_d4_shots": (12, [("srand", 77)] + random_circuit_script(12, 4, seed=1234) +
                          [("run_shots", 2000), ("argmax",)]),
    "grover7_after_gates": (7, [("x", 2), ("h", 4), ("cnot", 4, 1), ("grover", 99),
                                ("prob", 99)]),
    "rotations": (1, [("rx", 0, math.pi), ("ry", 0, math.pi), ("h", 0),
                      ("rz", 0, math.pi), ("h", 0), ("phase", 0, math.pi)]),
    "invalid_qubits": (3, [("h", 0), ("h", 7), ("cnot", 1, 1), ("cnot", 0, 5), ("x", -1),
                           ("rz", 3, 0.5), ("h", 2)]),
}
