"""The drop-in claim of include/cpecan_host.h as a binary: the reference's UNMODIFIED vanillaAlign.c, compiled against
the reference's own headers (oracle/Makefile target `dropin`, built where /root/reference exists) and linked against
libcpecan_host.so instead of the reference's impl/*.c, must write what the reference binary writes
(tests/golden/vanillaAlign/*, produced by oracle/_ref/vanillaAlign = the same source linked against the reference's own
library).  All four machines the CLI offers besides the HDP one: -s (threeState), default (vanilla), -f (fourState),
-e (echelon); the two strands run as two OpenMP sections through the host library at once."""
import gzip
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "cpecan-signal_b200")
GOLD = os.path.join(ROOT, "tests", "golden")
VA = os.path.join(GOLD, "vanillaAlign")
EXE = os.path.join(ROOT, "oracle", "_ref", "vanillaAlign_dropin")
P_TOL = 1e-4


def test_drop_in_binary_links_only_against_this_repo():
    """CPU: every undefined symbol of the binary that is not libc / libgomp comes from libcpecan_host.so."""
    if not os.path.exists(EXE):
        pytest.skip("oracle/_ref/vanillaAlign_dropin is built only where /root/reference exists")
    und = {l.split()[-1].split("@")[0] for l in subprocess.run(["nm", "-D", "--undefined-only", EXE], capture_output=True, text=True,
                                                             check=True).stdout.splitlines() if l.strip()}
    host = {l.split()[-1] for l in subprocess.run(["nm", "-D", "--defined-only", os.path.join(PKG, "libcpecan_host.so")],
                                                  capture_output=True, text=True, check=True).stdout.splitlines() if l.strip()}
    from_host = und & host
    for need in ("getAlignedPairsUsingAnchors", "getExpectationsUsingAnchors", "getStrawManStateMachine3", "getStateMachine4",
                 "getStateMachineEchelon", "getSignalStateMachine3Vanilla", "cigarRead", "nanopore_loadNanoporeReadFromFile",
                 "convertPairwiseForwardStrandAlignmentToAnchorPairs", "hmmContinuous_writeToFile", "sequence_padSequence",
                 "diagonalCalculationMultiPosteriorMatchProbs", "stList_sort"):
        assert need in from_host, need
    ldd = subprocess.run(["ldd", EXE], capture_output=True, text=True).stdout
    assert "libcpecan_host.so" in ldd and "libcpecan_ref" not in ldd


def _rows(path, multi=False):
    rows = {}
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "rt") as fh:
        for line in fh:
            f = line.rstrip("\n").split("\t")
            assert len(f) == 15
            key = (f[0], int(f[1]), f[2], f[3], f[4], int(f[5]))
            if multi:
                rows.setdefault(key, []).append(f)
            else:
                assert key not in rows
                rows[key] = [f]
    return rows


def _run(flags, extra, tmp_path, env=None):
    args = flags + ["-T", os.path.join(PKG, "models", "template_median68pA.model"),
                    "-C", os.path.join(PKG, "models", "complement_median68pA_pop2.model"), "-L", "readA",
                    "-q", os.path.join(GOLD, "ZymoC_ch_1_file1.npRead"), "-r", os.path.join(GOLD, "ZymoRef.txt")] + extra
    with open(os.path.join(VA, "guide.cigar")) as fin:
        r = subprocess.run([EXE] + args, stdin=fin, capture_output=True, text=True, timeout=600,
                           env=dict(os.environ, OMP_NUM_THREADS="2", **(env or {})))
    assert r.returncode == 0, r.stderr[-3000:]
    return r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("flag,tag", [("-s", "s"), ("", "v"), ("-f", "f"), ("-e", "e")])
def test_unmodified_vanilla_align_posteriors(tmp_path, flag, tag):
    if not os.path.exists(EXE):
        pytest.skip("oracle/_ref/vanillaAlign_dropin is built only where /root/reference exists")
    out = str(tmp_path / "post.tsv")
    stdout = _run([flag] if flag else [], ["-u", out], tmp_path)
    gold = os.path.join(VA, "out_%s.tsv" % tag)
    multi = tag == "e"                       # echelon lists a pair once per k-mer the event covers: rows repeat
    got, want = _rows(out, multi), _rows(gold + (".gz" if multi else ""), multi)
    tol = 1e-6 if tag in ("f", "e") else P_TOL + 1e-6      # the FP64 machines reproduce the printed posteriors
    for key, ws in want.items():
        if key not in got:
            assert all(float(w[12]) <= 0.01 + 2 * P_TOL for w in ws), "row %r missing" % (key,)
            continue
        gs = got[key]
        assert len(gs) == len(ws), key
        for g, w in zip(sorted(gs, key=lambda f: float(f[12])), sorted(ws, key=lambda f: float(f[12]))):
            assert g[:12] == w[:12] and g[13:] == w[13:], (g, w)        # every column but the posterior is text-identical
            assert abs(float(g[12]) - float(w[12])) <= tol
    for key, gs in got.items():
        if key not in want:
            assert all(float(g[12]) <= 0.01 + 2 * P_TOL for g in gs), "extra row %r" % (key,)
    want_line = open(os.path.join(VA, "stdout_%s.txt" % tag)).read().split()
    got_line = stdout.split()
    print(stdout.strip(), len(got), len(want))
    assert got_line[:2] == want_line[:2]                                              # label, number of guide anchors
    for g, w in zip(got_line[2:], want_line[2:]):                                      # "pairs(score)" per strand
        assert int(g.split("(")[0]) == int(w.split("(")[0])
        gs, ws = g.split("(")[1].rstrip(")"), w.split("(")[1].rstrip(")")
        assert ("nan" in gs) if "nan" in ws else abs(float(gs) - float(ws)) < 0.01


@pytest.mark.gpu
@pytest.mark.parametrize("flag,tag", [("-s", "s"), ("", "v")])
def test_unmodified_vanilla_align_exact_arithmetic(tmp_path, flag, tag):
    """CPECAN_EXACT=1 (cpecan_host_set_exact_arithmetic): the threeState / vanilla posteriors in FP64 and the reference's
    operation order -- the posterior file row for row what the reference binary writes, no row missing or extra, the
    printed posteriors (%f: six decimals) equal to 1e-6."""
    if not os.path.exists(EXE):
        pytest.skip("oracle/_ref/vanillaAlign_dropin is built only where /root/reference exists")
    out = str(tmp_path / "post.tsv")
    _run([flag] if flag else [], ["-u", out], tmp_path, env={"CPECAN_EXACT": "1"})
    got, want = _rows(out), _rows(os.path.join(VA, "out_%s.tsv" % tag))
    assert got.keys() == want.keys()
    worst = 0.0
    for key, ws in want.items():
        g, w = got[key][0], ws[0]
        assert g[:12] == w[:12] and g[13:] == w[13:], (g, w)
        worst = max(worst, abs(float(g[12]) - float(w[12])))
    print(tag, len(got), "rows, worst posterior difference", worst)
    assert worst <= 1.000001e-6


@pytest.mark.gpu
def test_unmodified_vanilla_align_expectations(tmp_path):
    """-t / -c: the expectation files of the template strand against the reference binary's."""
    if not os.path.exists(EXE):
        pytest.skip("oracle/_ref/vanillaAlign_dropin is built only where /root/reference exists")
    t, c = str(tmp_path / "t.exp"), str(tmp_path / "c.exp")
    _run(["-s"], ["-t", t, "-c", c], tmp_path)

    def read_exp(path):
        with open(path) as fh:
            return fh.readline().split(), np.array(fh.readline().split(), dtype=np.float64), np.array(fh.readline().split(), dtype=np.float64)
    gh, g1, g2 = read_exp(t)
    wh, w1, w2 = read_exp(os.path.join(VA, "t_s.exp"))
    assert gh == wh and g1.shape == w1.shape == (10,) and g2.shape == w2.shape == (4096,)
    np.testing.assert_allclose(g1[:9], w1[:9], rtol=2e-4, atol=2e-6)
    assert abs(g1[9] - w1[9]) <= 1e-4 * abs(w1[9])
    np.testing.assert_allclose(g2, w2, rtol=2e-4, atol=1e-4)
