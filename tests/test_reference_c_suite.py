"""The reference's OWN C test programs (reference test/test_main.c:26-59 and test/test_qc_*.c),
compiled unchanged from where they lie -- against the reference's own include/qcs.h -- by
`make -C oracle reftests`, with only the library swapped:

  oracle/_ref/ref_suite_cuda       test_main.c + the 20 test files, linked against this repository's
                                   libqcs.so (mode QCS_GPU_CUDA): the drop-in proof
  oracle/_ref/ref_tests_cuda       the same 20 test files behind oracle/ref_test_one.c (one test per run)
  oracle/_ref/ref_tests_seq        ... linked against the unmodified reference (sequential mode)
  oracle/_ref/ref_tests_corrected  ... linked against the reference with defects D1-D3 repaired

The programs are built in the CPU container (the sources never enter this repository) and travel
to the GPU box as binaries.
"""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
TESTS = ["test_qc_create_destroy", "test_qc_getters", "test_qc_h", "test_qc_x", "test_qc_y", "test_qc_z",
         "test_qc_cnot", "test_qc_phase", "test_qc_rotations", "test_qc_barrier", "test_qc_reset",
         "test_qc_measure", "test_qc_run", "test_qc_run_shots", "test_qc_state_access", "test_qc_print",
         "test_qc_grover_search", "test_qc_qft", "test_qc_bv", "test_qc_optimize"]
# what the UNMODIFIED reference fails in its sequential mode (defect D1)
REFERENCE_FAILS = {"test_qc_cnot", "test_qc_bv"}


def _need(name):
    path = os.path.join(REF, name)
    if not os.path.exists(path):
        pytest.skip(f"{path} not built (make -C oracle reftests needs /root/reference)")
    return path


def _run(exe, *args, sem=None):
    env = dict(os.environ)
    if sem:
        env["QCS_CUDA_SEMANTICS"] = sem
    return subprocess.run([exe, *args], capture_output=True, text=True, timeout=300, env=env)


def test_driver_lists_the_reference_suite():
    r = _run(_need("ref_tests_seq"), "list")
    assert r.returncode == 0 and r.stdout.split() == TESTS


@pytest.mark.parametrize("name", TESTS)
def test_unmodified_and_corrected_reference_on_cpu(name):
    """Pins the survey's claim on the CPU: the unmodified reference fails exactly test_qc_cnot and
    test_qc_bv, the corrected one passes all 20."""
    seq = _run(_need("ref_tests_seq"), name)
    cor = _run(_need("ref_tests_corrected"), name)
    assert (seq.returncode != 0) == (name in REFERENCE_FAILS), seq.stderr
    assert cor.returncode == 0 and "[PASSED]" in cor.stdout, cor.stderr


@pytest.mark.gpu
def test_reference_driver_unchanged_against_libqcs_corrected():
    """test/test_main.c + its 20 tests, unchanged, on the GPU through libqcs.so: 20/20."""
    r = _run(_need("ref_suite_cuda"), sem="corrected")
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("[PASSED]") == 20 and "ALL TESTS PASSED SUCCESSFULLY" in r.stdout


@pytest.mark.gpu
def test_reference_driver_unchanged_against_libqcs_reference_semantics():
    """Bug-compatible semantics: the suite stops where the unmodified reference's own run stops --
    at test_qc_cnot's assert (test/test_qc_cnot.c:10), after the same six passes."""
    r = _run(_need("ref_suite_cuda"), sem="reference")
    assert r.returncode != 0
    # (the driver never flushes stdout, and abort() drops what a pipe has buffered: the site of the
    # failing assert on stderr is the evidence -- the six tests in front of it passed or it would
    # name one of them)
    assert "test_qc_cnot.c:10: test_qc_cnot: Assertion `qc_find_most_likely_state(c) == 3' failed" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("name", TESTS)
def test_each_reference_test_same_verdict_as_the_reference(name):
    """Per test: libqcs.so under `reference` semantics passes / fails exactly like the unmodified
    reference, with the same failing assertion; under `corrected` like the corrected reference."""
    cuda = _need("ref_tests_cuda")
    want = _run(_need("ref_tests_seq"), name)
    got = _run(cuda, name, sem="reference")
    assert (got.returncode == 0) == (want.returncode == 0), got.stderr
    if want.returncode != 0:
        # "<prog>: <file>:<line>: <function>: Assertion `...' failed." -- same site
        assert got.stderr.split(": ", 1)[1].strip() == want.stderr.split(": ", 1)[1].strip()
    elif name != "test_qc_print":
        assert got.stdout == want.stdout
    got = _run(cuda, name, sem="corrected")
    assert got.returncode == 0 and "[PASSED]" in got.stdout, got.stderr
