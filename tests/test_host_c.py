"""The C test program of libcpecan_host.so (tests/host/host_tests.c): the reference's own CuTest cases for the path,
re-written against include/cpecan_host.h.  'cpu' needs no device (geometry, containers, file formats); 'gpu' runs the
fixture alignments through getAlignedPairsUsingAnchors / getExpectationsUsingAnchors."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "cpecan-signal_b200")
EXE = os.path.join(ROOT, "tests", "host", "host_tests")


def build():
    subprocess.check_call(["make", "-s", "-C", PKG])
    src = os.path.join(ROOT, "tests", "host", "host_tests.c")
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < max(os.path.getmtime(src), os.path.getmtime(os.path.join(PKG, "libcpecan_host.so"))):
        subprocess.check_call(["gcc", "-std=gnu99", "-O1", "-g", "-Wall", "-I" + os.path.join(ROOT, "include"), "-o", EXE, src,
                               "-L" + PKG, "-lcpecan_host", "-lcpecan_cuda", "-Wl,-rpath," + PKG, "-lm"])
    return EXE


def run(mode):
    exe = build()
    r = subprocess.run([exe, mode, os.path.join(ROOT, "tests", "golden"), os.path.join(PKG, "models")],
                       capture_output=True, text=True, timeout=600)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0, r.stderr[-2000:]
    assert " 0 failures" in r.stdout


def test_host_library_exports_header_symbols():
    import ctypes
    import re
    build()
    lib = ctypes.CDLL(os.path.join(PKG, "libcpecan_host.so"))
    txt = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "cpecan_host.h")).read(), flags=re.S)
    names = set(re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\((?!\*)", txt)) - {"defined", "void", "sizeof"}
    names = {n for n in names if re.search(r"^(st|sequence_|pairwise|diagonal|band|logAdd|get|sort|filter|emissions_|stateMachine|hmm|vanillaHmm|"
                                           r"continuousPairHmm|nanopore_|cpecan_host|deserialize_|destroy_|cigarRead|destructPairwise|checkPairwise|convertPairwise|cell_|dpDiagonal_|dpMatrix_)", n)}
    assert len(names) > 120
    for n in sorted(names):
        assert hasattr(lib, n), "libcpecan_host.so does not export %s" % n


def test_host_c_cpu():
    run("cpu")


@pytest.mark.gpu
def test_host_c_gpu():
    run("gpu")


def test_no_cpu_path_aborts():
    """Without a CUDA device the alignment entry point aborts with the reference's st_errAbort behaviour."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    exe = build()
    r = subprocess.run([exe, "gpu", os.path.join(ROOT, "tests", "golden"), os.path.join(PKG, "models")],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr
