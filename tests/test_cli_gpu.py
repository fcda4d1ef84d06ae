"""GPU test of cpecanAlign (the batched sibling of the reference's vanillaAlign CLI) against files written by the
UNMODIFIED reference binary (oracle/_ref/vanillaAlign, built by oracle/Makefile) on the fixture 2D read with a lastz
guide cigar: tests/golden/vanillaAlign/*.  Generator commands: oracle/make_golden.py (section 'vanillaAlign')."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "cpecan-signal_b200")
GOLD = os.path.join(ROOT, "tests", "golden")
VA = os.path.join(GOLD, "vanillaAlign")
EXE = os.path.join(PKG, "cpecanAlign")
P_TOL = 1e-4


def _run(args, stdin_path=None):
    subprocess.check_call(["make", "-s", "-C", PKG])
    with open(stdin_path) if stdin_path else open(os.devnull) as fin:
        r = subprocess.run([EXE] + args, stdin=fin, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


def _common(flag):
    return ([flag] if flag else []) + ["-T", os.path.join(PKG, "models", "template_median68pA.model"),
                                       "-C", os.path.join(PKG, "models", "complement_median68pA_pop2.model"),
                                       "-q", os.path.join(GOLD, "ZymoC_ch_1_file1.npRead"), "-r", os.path.join(GOLD, "ZymoRef.txt")]


def _rows(path):
    rows = {}
    for line in open(path):
        f = line.rstrip("\n").split("\t")
        assert len(f) == 15
        key = (f[0], int(f[1]), f[2], f[3], f[4], int(f[5]))
        assert key not in rows
        rows[key] = f
    return rows


def _compare_tsv(got_path, want_path, threshold=0.01):
    got, want = _rows(got_path), _rows(want_path)
    for key, w in want.items():
        if key not in got:
            assert float(w[12]) <= threshold + 2 * P_TOL, "row %r missing" % (key,)
            continue
        g = got[key]
        assert g[:12] == w[:12] and g[13:] == w[13:], (g, w)        # every column but the posterior is text-identical
        assert abs(float(g[12]) - float(w[12])) <= P_TOL + 1e-6
    for key, g in got.items():
        if key not in want:
            assert float(g[12]) <= threshold + 2 * P_TOL, "extra row %r" % (key,)
    return len(got), len(want)


@pytest.mark.parametrize("flag,tag", [("-s", "s"), ("", "v")])
def test_posteriors_tsv_and_stdout(tmp_path, flag, tag):
    out = str(tmp_path / "post.tsv")
    stdout = _run(_common(flag) + ["-L", "readA", "-u", out], os.path.join(VA, "guide.cigar"))
    n_got, n_want = _compare_tsv(out, os.path.join(VA, "out_%s.tsv" % tag))
    want_line = open(os.path.join(VA, "stdout_%s.txt" % tag)).read().split()
    got_line = stdout.split()
    print(stdout, n_got, n_want)
    assert got_line[0] == want_line[0] and got_line[1] == want_line[1]            # label, number of guide anchors
    for g, w in zip(got_line[2:], want_line[2:]):                                  # "pairs(score)" per strand
        gn, wn = int(g.split("(")[0]), int(w.split("(")[0])
        assert abs(gn - wn) <= 2
        gs, ws = g.split("(")[1].rstrip(")"), w.split("(")[1].rstrip(")")
        if "nan" in ws:
            assert "nan" in gs
        else:
            assert abs(float(gs) - float(ws)) < 0.05


@pytest.mark.parametrize("flag,tag", [("-f", "f"), ("-e", "e"), ("-d", "d")])
def test_posteriors_fp64_machines(tmp_path, hdp_fixture, flag, tag):
    """fourState, echelon and the HDP machine through cpecanAlign: the reference binary's posterior file row for row
    (echelon lists a pair once per k-mer the event covers: rows repeat), posteriors equal at the six printed decimals."""
    import gzip
    from collections import Counter
    out = str(tmp_path / "post.tsv")
    extra = ["-v", hdp_fixture["path"], "-w", hdp_fixture["path"]] if tag == "d" else []
    stdout = _run(_common(flag) + extra + ["-L", "readA", "-u", out], os.path.join(VA, "guide.cigar"))
    gold = os.path.join(VA, "out_%s.tsv" % tag)
    op = (lambda p: gzip.open(p + ".gz", "rt")) if not os.path.exists(gold) else open

    def rows(fh):
        c = Counter()
        for line in fh:
            f = line.rstrip("\n").split("\t")
            assert len(f) == 15
            c[tuple(f[:12]) + ("%.5f" % float(f[12]),) + tuple(f[13:])] += 1
        return c
    with open(out) as fh:
        got = rows(fh)
    with op(gold) as fh:
        want = rows(fh)
    # rows whose posterior rounds differently at the fifth decimal would show up on both sides
    only_got, only_want = got - want, want - got
    assert sum(only_got.values()) == sum(only_want.values()) <= 4, (list(only_got)[:3], list(only_want)[:3])
    for a, b in zip(sorted(only_got), sorted(only_want)):
        assert a[:12] == b[:12] and abs(float(a[12]) - float(b[12])) <= 2e-5
    want_line = open(os.path.join(VA, "stdout_%s.txt" % tag)).read().split()
    got_line = stdout.split()
    print(stdout.strip(), sum(got.values()), sum(want.values()))
    assert got_line[:2] == want_line[:2]
    for g, w in zip(got_line[2:], want_line[2:]):
        assert g.split("(")[0] == w.split("(")[0]


def test_hdp_expectation_file(tmp_path, hdp_fixture):
    """cpecanAlign -d -t / -c: the HdpHmm file of the template strand equals the reference binary's."""
    import gzip
    t, c = str(tmp_path / "t.exp"), str(tmp_path / "c.exp")
    _run(_common("-d") + ["-v", hdp_fixture["path"], "-w", hdp_fixture["path"], "-L", "readA", "-t", t, "-c", c],
         os.path.join(VA, "guide.cigar"))
    got = open(t).read().split("\n")
    with gzip.open(os.path.join(VA, "t_d.exp.gz"), "rt") as fh:
        want = fh.read().split("\n")
    assert got[0] == want[0] and len(got) == len(want)
    np.testing.assert_allclose(np.array(got[1].split(), dtype=np.float64), np.array(want[1].split(), dtype=np.float64), rtol=0, atol=2e-6)
    assert got[2] == want[2] and got[3] == want[3]


def _read_exp(path):
    with open(path) as fh:
        head = fh.readline().split()
        l1 = np.array(fh.readline().split(), dtype=np.float64)
        l2 = np.array(fh.readline().split(), dtype=np.float64)
    return head, l1, l2


def test_expectation_files(tmp_path):
    t, c = str(tmp_path / "t.exp"), str(tmp_path / "c.exp")
    _run(_common("-s") + ["-L", "readA", "-t", t, "-c", c], os.path.join(VA, "guide.cigar"))
    # complement: the fixture's complement event map is decreasing, the slice has a negative length; the reference
    # walks lX + lY = 244 one-cell diagonals along y = 0 (244 X->X steps, likelihood -234346.938040): reproduced
    for got_path, want_path in ((t, os.path.join(VA, "t_s.exp")), (c, os.path.join(VA, "c_s.exp"))):
        gh, g1, g2 = _read_exp(got_path)
        wh, w1, w2 = _read_exp(want_path)
        assert gh == wh and g1.shape == w1.shape == (10,) and g2.shape == w2.shape == (4096,)
        np.testing.assert_allclose(g1[:9], w1[:9], rtol=2e-4, atol=2e-6)
        assert abs(g1[9] - w1[9]) <= 1e-4 * abs(w1[9])
        np.testing.assert_allclose(g2, w2, rtol=2e-4, atol=1e-4)


def test_batch_manifest(tmp_path):
    """Two reads (the same fixture twice) in one GPU batch == two single runs."""
    cigar = open(os.path.join(VA, "guide.cigar")).readline().strip()
    man = tmp_path / "manifest.tsv"
    outs = [str(tmp_path / ("r%d.tsv" % i)) for i in range(2)]
    man.write_text("".join("read%d\t%s\t%s\t%s\t%s\n" % (i, os.path.join(GOLD, "ZymoC_ch_1_file1.npRead"),
                                                        os.path.join(GOLD, "ZymoRef.txt"), outs[i], cigar) for i in range(2)))
    stdout = _run(_common("-s")[:5] + ["--batch", str(man)])
    lines = stdout.strip().split("\n")
    assert len(lines) == 2 and lines[0].split()[1:] == lines[1].split()[1:]
    a, b = open(outs[0]).read().replace("read0", "readA"), open(outs[1]).read().replace("read1", "readA")
    assert a == b and len(a.splitlines()) > 900
