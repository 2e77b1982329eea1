"""CPU-side checks (no GPU): the C-ABI library loads and exports every symbol
include/*.h declares; the host layer's bookkeeping (history, num_gates,
qc_optimize, qc_print_circuit) matches the real reference; the product refuses
to run without a device instead of falling back."""
import ctypes
import os
import re
import subprocess

import pytest

from oracle import pyoracle as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def libs():
    from qcs_b200 import build
    return build.build_all()


def _declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qcs?_[a-z0-9_]+)\s*\(", text)))


def _exported(path):
    out = subprocess.check_output(["nm", "-D", "--defined-only", path], text=True)
    return {line.split()[-1] for line in out.splitlines() if " T " in line}


def test_cuda_abi_exports_every_declared_symbol(libs):
    declared = [s for s in _declared("qcs_cuda.h") if s.startswith("qcs_cuda_")]
    assert len(declared) >= 25
    missing = [s for s in declared if s not in _exported(libs["cuda"])]
    assert not missing, missing


def test_host_api_exports_every_declared_symbol(libs):
    declared = _declared("qcs.h")
    assert len(declared) == 30, declared   # the reference's 28 + qc_cphase + qc_add_gate
    missing = [s for s in declared if s not in _exported(libs["host"])]
    assert not missing, missing
    assert "qc_cuda_engine" in _exported(libs["host"])


def test_header_matches_reference_prototypes(libs):
    """Same function names as the reference header."""
    ref = ["qc_create", "qc_destroy", "qc_h", "qc_x", "qc_y", "qc_z", "qc_cnot", "qc_phase",
           "qc_rx", "qc_ry", "qc_rz", "qc_barrier", "qc_reset", "qc_measure", "qc_measure_all",
           "qc_run", "qc_run_shots", "qc_find_most_likely_state", "qc_get_probability",
           "qc_print_state", "qc_print_circuit", "qc_grover_search",
           "qc_quantum_fourier_transform", "qc_bernstein_vazirani", "qc_ghz_state",
           "qc_get_num_qubits", "qc_get_num_gates", "qc_optimize"]
    assert len(ref) == 28
    declared = _declared("qcs.h")
    for name in ref:
        assert name in declared, name


def test_bindings_cover_the_abi(libs):
    from qcs_b200 import _ffi
    declared = [s for s in _declared("qcs_cuda.h") if s.startswith("qcs_cuda_")]
    assert sorted(_ffi.CUDA_ABI) == sorted(declared)
    _ffi.load()


def test_no_cpu_fallback_without_device(libs):
    """On a box without a GPU qc_create must fail loudly (NULL + message), never compute on the CPU."""
    if os.path.exists("/dev/nvidiactl") or os.path.exists("/dev/nvidia0"):
        pytest.skip("a GPU is present")
    from qcs_b200 import Circuit, QcsError
    with pytest.raises(QcsError):
        Circuit(3)


def _history_script():
    return [("h", 0), ("h", 0), ("x", 1), ("cnot", 0, 1), ("cnot", 0, 1), ("y", 2), ("y", 2),
            ("rz", 1, 0.5), ("z", 1), ("z", 1), ("barrier",), ("h", 9), ("cnot", 1, 1),
            ("cphase", 0, 2, 0.3), ("x", 2), ("x", 2), ("phase", 0, 1.0)]


def test_history_and_optimize_match_reference(libs):
    if not po.ref_available("seq"):
        pytest.skip("oracle/_ref not built")
    from qcs_b200 import Circuit
    ref = po.RefLib(3, "seq")
    c = Circuit(3, dryrun=True)
    po.replay(ref, _history_script()); po.replay(c, _history_script())
    assert c.num_gates == ref.num_gates == len(_history_script())
    ref.optimize(); c.optimize()
    assert c.num_gates == ref.num_gates
    # H,H / CNOT,CNOT / Y,Y / Z,Z / X,X cancel; cascades rescan from the start
    assert c.num_gates == len(_history_script()) - 10


def _capture_stdout(fn):
    import tempfile
    libc = ctypes.CDLL(None)
    libc.fflush(None)
    saved = os.dup(1)
    with tempfile.TemporaryFile() as tmp:
        os.dup2(tmp.fileno(), 1)
        try:
            fn()
            libc.fflush(None)
        finally:
            os.dup2(saved, 1)
            os.close(saved)
        tmp.seek(0)
        return tmp.read()


def test_print_circuit_is_byte_identical(libs):
    if not po.ref_available("seq"):
        pytest.skip("oracle/_ref not built")
    from qcs_b200 import Circuit
    script = [("h", 0), ("x", 1), ("cnot", 0, 2), ("rz", 1, 0.2), ("cnot", 2, 1), ("barrier",), ("y", 0)]
    ref = po.RefLib(3, "seq"); c = Circuit(3, dryrun=True)
    po.replay(ref, script); po.replay(c, script)
    ref.L.qc_print_circuit.argtypes = [ctypes.c_void_p]
    want = _capture_stdout(lambda: ref.L.qc_print_circuit(ref.c))
    got = _capture_stdout(lambda: c.print_circuit())
    assert got == want and b"QUANTUM CIRCUIT" in got


def test_invalid_gates_are_recorded_but_not_queued(libs):
    """Reference quirk D13: a rejected gate still lands in the history (src/qcs.c:161-163)."""
    from qcs_b200 import Circuit
    c = Circuit(3, dryrun=True)
    c.h(7); c.cnot(1, 1); c.x(-1); c.h(0)
    assert c.num_gates == 4
    c.flush()
    assert c.stats()["gates_submitted"] == 1


def test_single_header_bundle_compiles_as_c89(tmp_path):
    """scripts/bundle.py emits one qcs.h; a C89 program that defines
    QCS_IMPLEMENTATION builds against it and links only libqcs_cuda.so.  Without a GPU qc_create
    must fail loudly (NULL + message), never fall back to the CPU."""
    import subprocess, sys, shutil
    from qcs_b200 import build
    build.build_all()
    hdr = tmp_path / "qcs.h"
    subprocess.check_call([sys.executable, os.path.join(ROOT, "scripts", "bundle.py"), str(hdr)])
    prog = tmp_path / "prog.c"
    prog.write_text('#define QCS_IMPLEMENTATION\n#include "qcs.h"\n#include <stdio.h>\n'
                    'int main(void) { t_q_circuit *c = qc_create(3);\n'
                    '  if (!c) { printf("no device\\n"); return 0; }\n'
                    '  qc_h(c, 0); qc_cnot(c, 0, 1); printf("%d %d\\n", qc_get_num_gates(c), qc_find_most_likely_state(c));\n'
                    '  qc_destroy(c); return 0; }\n')
    exe = tmp_path / "prog"
    lib = os.path.join(ROOT, "qcs_b200", "lib")
    subprocess.check_call(["gcc", "-std=c89", "-pedantic", "-Wall", "-ffp-contract=off", "-I", str(tmp_path),
                           str(prog), "-L" + lib, "-lqcs_cuda", "-lm", "-Wl,-rpath," + lib, "-o", str(exe)])
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    import torch
    if not torch.cuda.is_available():
        assert "no device" in r.stdout and "Error" in r.stderr
    else:
        assert r.stdout.split()[0] == "2"


def test_bench_reference_arm_contract_line():
    """bench.py --impl reference runs on the host cores (no GPU): one JSON line on stdout with the
    contract keys, the unmodified reference (oracle/_ref) or the C port as the timed thing."""
    import json, subprocess, sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                        "--qubits", "16", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step",
                "higher_is_better", "scaling", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "gates/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_circuit_options_do_not_clobber_process_defaults(libs):
    """Circuit(..., semantics=...) passes options as process-wide defaults around qc_create; a default the
    caller had set before must still be there afterwards (round-1 advice: it used to be wiped)."""
    from qcs_b200 import Circuit
    from qcs_b200.circuit import get_default, set_default
    try:
        set_default("semantics", "corrected")
        set_default("tile_bits", "12")
        c = Circuit(3, dryrun=True, semantics="reference", tile_bits=10)
        c.close()
        assert get_default("semantics") == "corrected" and get_default("tile_bits") == "12"
        assert get_default("dryrun") is None              # was not set before: not left behind either
        c = Circuit(3, dryrun=True)                       # no explicit option: the caller's default applies
        c.h(0); c.cnot(0, 1); c.flush()
        kinds = [t[1] for t in c.trace() if t[0] == "gate"]
        assert len(kinds) == 2
        c.close()
    finally:
        set_default("semantics", None)
        set_default("tile_bits", None)
    from qcs_b200 import _ffi
    assert _ffi.load()[1].qcs_cuda_trim_pool() >= 0
