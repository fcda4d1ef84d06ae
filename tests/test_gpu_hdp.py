"""GPU parity of the threeStateHdp machine (FP64 kernel of cpecan_generic.cuh with the HDP's density
tables): the reference's goldens on its own serialised HDP fixture and fixture read, random small problems against the
oracle, and the unmodified vanillaAlign.c `-d` through libcpecan_host.so."""
import gzip
import os
import subprocess

import numpy as np
import pytest

import parity

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "cpecan-signal_b200")
GOLD = os.path.join(ROOT, "tests", "golden")
EXE = os.path.join(ROOT, "oracle", "_ref", "vanillaAlign_dropin")


def _same(got, want, score_tol=2):
    got, want = np.asarray(got, dtype=np.int64).reshape(-1, 3), np.asarray(want, dtype=np.int64).reshape(-1, 3)
    assert got.shape == want.shape, (got.shape, want.shape)
    assert np.array_equal(got[:, 1:], want[:, 1:])
    worst = int(np.abs(got[:, 0] - want[:, 0]).max()) if len(got) else 0
    assert worst <= score_tol, worst
    return worst


@pytest.mark.parametrize("tag,e,ragged,thr", [("hdp_e20_r00", 20, (0, 0), 0.01), ("hdp_e50_r11", 50, (1, 1), 0.01),
                                               ("hdp_e20_r10_t30", 20, (1, 0), 0.3)])
def test_fixture_hdp(engine, zymo, hdp_fixture, tag, e, ragged, thr):
    from cpecan_signal import HostBatch, default_params, hdp_hmm
    from cpecan_signal.engine import item_pairs
    g = hdp_fixture["golden"]
    mid = engine.upload_hdp(hdp_fixture["hdp"])
    batch = HostBatch([zymo["ref"]], [hdp_fixture["events"]], [zymo["anchors_template"]], model_ids=[mid], ragged=[ragged])
    res, pairs, totals = engine.align_batch(batch, hmm=hdp_hmm(), params=default_params(diagonalExpansion=e, threshold=thr),
                                            want_totals=True, pair_cap=200000)
    engine.release_model(mid)
    assert res[0]["status"] == 0
    worst = _same(parity.reverse_regions(item_pairs(res, pairs, 0)), g[tag + "_pairs"])
    wt = g[tag + "_totals"]
    mask = ~np.isnan(wt)
    assert np.array_equal(mask, ~np.isnan(totals[0]))
    dt = float(np.abs(totals[0][mask] - wt[mask]).max())
    print(tag, res[0]["n_pairs"], "worst score diff", worst, "worst |total diff|", dt)
    assert dt <= 1e-9 * float(np.abs(wt[mask]).max())


def test_random_small_problems_vs_oracle(engine, template_tables, hdp_fixture):
    """Random reads (events drawn around the pore model's levels, unscaled -- the HDP machine takes descaled events), even
    and odd expansions, traceback-heavy parameter sets, every ragged combination: each item against the oracle."""
    import oracleshim as O
    from cpecan_signal import HostBatch, default_params, hdp_hmm, synth
    from cpecan_signal.engine import item_pairs
    l1 = template_tables[0]
    h = hdp_fixture["hdp"]
    m = O.Model(O.THREE_STATE_HDP, hdp=h)
    mid = engine.upload_hdp(h)
    rng = np.random.default_rng(171)
    for mind, tbd, e, thr in [(5, 2, 4, 0.01), (20, 8, 11, 0.2), (60, 40, 20, 0.05), (1000, 40, 31, 0.01)]:
        reads, anchors, ragged, events = [], [], [], []
        for _ in range(8):
            r = synth.make_read(l1, int(rng.integers(1, 1 << 30)), lX=int(rng.integers(8, 200)), anchor_every=int(rng.integers(5, 60)))
            ev = np.array(r.events, dtype=np.float64).reshape(-1, 3).copy()
            ev[:, 0] = (ev[:, 0] - r.scale5[1]) / r.scale5[0]
            keep = rng.random(len(r.anchors)) < rng.choice([0.0, 0.3, 1.0])
            reads.append(r); events.append(ev); anchors.append(r.anchors[keep]); ragged.append((int(rng.integers(0, 2)), int(rng.integers(0, 2))))
        batch = HostBatch([r.ref for r in reads], events, anchors, model_ids=[mid] * len(reads), ragged=ragged)
        kw = dict(diagonalExpansion=e, minDiagsBetweenTraceBack=mind, traceBackDiagonals=tbd, threshold=thr)
        res, pairs, totals = engine.align_batch(batch, hmm=hdp_hmm(), params=default_params(**kw), want_totals=True, pair_cap=400000)
        worst = 0
        for i, r in enumerate(reads):
            want, wtot = O.align_banded(m, r.ref, events[i], anchors[i], params=O.default_params(**kw), ragged=ragged[i], want_totals=True)
            assert res[i]["status"] == 0, (i, kw)
            worst = max(worst, _same(parity.reverse_regions(item_pairs(res, pairs, i)), want))
            mask = ~np.isnan(wtot)
            assert np.array_equal(mask, ~np.isnan(totals[i]))
            if mask.any():
                assert float(np.abs(totals[i][mask] - wtot[mask]).max()) <= 1e-9 * max(1.0, float(np.abs(wtot[mask]).max()))
        print(kw, "worst score diff", worst)
    engine.release_model(mid)


def test_model_kinds_do_not_mix(engine, template_tables, hdp_fixture):
    from cpecan_signal import EngineError, HostBatch, hdp_hmm, synth, three_state_hmm
    l1, _, l3 = template_tables
    r = synth.make_read(l1, 5, lX=100)
    pore = engine.upload_model(l1, l3, np.full(4096, -2.3025850929940455))
    hid = engine.upload_hdp(hdp_fixture["hdp"])
    with pytest.raises(EngineError, match="upload_hdp"):
        engine.align_batch(HostBatch([r.ref], [r.events], [r.anchors], model_ids=[pore], ragged=[(1, 1)]), hmm=hdp_hmm())
    with pytest.raises(EngineError, match="upload_hdp"):
        engine.align_batch(HostBatch([r.ref], [r.events], [r.anchors], model_ids=[hid], ragged=[(1, 1)]))
    with pytest.raises(EngineError, match="not threeStateHdp"):
        engine.hdp_expectations_batch(HostBatch([r.ref], [r.events], [r.anchors], model_ids=[pore], ragged=[(1, 1)]), hmm=three_state_hmm())
    # a k-mer the HDP's alphabet cannot spell: reported per item (the reference ends the program; so does the host library)
    bad = r.ref[:40] + "N" + r.ref[41:]
    res, _, _ = engine.align_batch(HostBatch([bad, r.ref], [r.events, r.events], [r.anchors, r.anchors], model_ids=[hid, hid],
                                             ragged=[(1, 1), (1, 1)]), hmm=hdp_hmm(), pair_cap=100000)
    assert res[0]["status"] & 8 and res[1]["status"] == 0
    engine.release_model(pore)
    engine.release_model(hid)


def test_unmodified_vanilla_align_hdp(tmp_path, hdp_fixture):
    """`vanillaAlign -d -v X.nhdp -w X.nhdp`: the reference's CLI source against libcpecan_host.so writes the posterior
    file the reference binary (linked with the reference's own HDP sources) writes."""
    if not os.path.exists(EXE):
        pytest.skip("oracle/_ref/vanillaAlign_dropin is built only where /root/reference exists")
    out = str(tmp_path / "post.tsv")
    va = os.path.join(GOLD, "vanillaAlign")
    args = ["-d", "-v", hdp_fixture["path"], "-w", hdp_fixture["path"],
            "-T", os.path.join(PKG, "models", "template_median68pA.model"), "-C", os.path.join(PKG, "models", "complement_median68pA_pop2.model"),
            "-L", "readA", "-q", os.path.join(GOLD, "ZymoC_ch_1_file1.npRead"), "-r", os.path.join(GOLD, "ZymoRef.txt"), "-u", out]
    with open(os.path.join(va, "guide.cigar")) as fin:
        r = subprocess.run([EXE] + args, stdin=fin, capture_output=True, text=True, timeout=600, env=dict(os.environ, OMP_NUM_THREADS="2"))
    assert r.returncode == 0, r.stderr[-3000:]

    def rows(fh):
        d = {}
        for line in fh:
            f = line.rstrip("\n").split("\t")
            key = (f[0], int(f[1]), f[2], f[3], f[4], int(f[5]))
            assert key not in d
            d[key] = f
        return d
    with open(out) as fh:
        got = rows(fh)
    with gzip.open(os.path.join(va, "out_d.tsv.gz"), "rt") as fh:
        want = rows(fh)
    assert got.keys() == want.keys() and len(got) == 11665
    worst = 0.0
    for key, w in want.items():
        g = got[key]
        assert g[:12] == w[:12] and g[13:] == w[13:], (g, w)
        worst = max(worst, abs(float(g[12]) - float(w[12])))
    print(r.stdout.strip(), "worst posterior difference", worst)
    assert worst <= 1.000001e-6
    want_line = open(os.path.join(va, "stdout_d.txt")).read().split()
    got_line = r.stdout.split()
    assert got_line[:2] == want_line[:2]                                               # label, number of guide anchors
    for g, w in zip(got_line[2:], want_line[2:]):                                       # "pairs(score)" per strand
        assert g.split("(")[0] == w.split("(")[0]


@pytest.mark.parametrize("tag,e,ragged,thr", [("hdpexp_e50_r11", 50, (1, 1), 0.01), ("hdpexp_e20_r00_t30", 20, (0, 0), 0.3)])
def test_fixture_hdp_expectations(engine, zymo, hdp_fixture, tag, e, ragged, thr):
    """getExpectationsUsingAnchors with an HdpHmm: transition sums and likelihood to 1e-9 of the reference's, the
    event-to-k-mer assignments the reference's lists in the reference's order."""
    from cpecan_signal import HostBatch, default_params, hdp_hmm
    from cpecan_signal.engine import item_pairs
    g = hdp_fixture["golden"]
    mid = engine.upload_hdp(hdp_fixture["hdp"])
    batch = HostBatch([zymo["ref"]], [hdp_fixture["events"]], [zymo["anchors_template"]], model_ids=[mid], ragged=[ragged])
    vec, res, asg = engine.hdp_expectations_batch(batch, hmm=hdp_hmm(), params=default_params(diagonalExpansion=e, threshold=thr),
                                                  pseudocount=1e-4)
    engine.release_model(mid)
    assert res[0]["status"] == 0
    want = g[tag + "_vec"]
    np.testing.assert_allclose(vec[:9], want[:9], rtol=1e-9, atol=1e-12)
    assert abs(vec[-1] - want[9]) <= 1e-9 * abs(want[9])
    got = np.asarray(item_pairs(res, asg, 0), dtype=np.int64).reshape(-1, 3)
    assert np.array_equal(got[:, 1:], g[tag + "_assignments"])
    print(tag, len(got), "assignments; transitions", vec[:9])


@pytest.mark.parametrize("machine", ["three", "vanilla", "hdp"])
def test_fp64_expectations_vs_oracle(engine, template_tables, hdp_fixture, machine):
    """The expectation mode of the FP64 kernel -- threeState and vanilla with cpecan_cuda_set_exact_arithmetic or an odd
    expansion, threeStateHdp always -- on random reads: the oracle's sums to 1e-9 (k-mer skip counts / skip bins included),
    and for the HDP machine the oracle's assignment lists."""
    import oracleshim as O
    from cpecan_signal import HostBatch, default_params, hdp_hmm, synth, three_state_hmm, vanilla_gapx, vanilla_hmm
    from cpecan_signal.engine import item_pairs
    l1, l2, l3 = template_tables
    rng = np.random.default_rng({"three": 181, "vanilla": 182, "hdp": 183}[machine])
    if machine == "hdp":
        mid, hmm = engine.upload_hdp(hdp_fixture["hdp"]), hdp_hmm()
    elif machine == "vanilla":
        mid, hmm = engine.upload_model(l1, l3, vanilla_gapx(l2)), vanilla_hmm("template")
    else:
        mid, hmm = engine.upload_model(l1, l3, np.full(4096, -2.3025850929940455)), three_state_hmm()
    engine.set_exact_arithmetic(True)
    try:
        for mind, tbd, e in [(20, 8, 10), (60, 40, 21), (1000, 40, 30)]:
            reads, anchors, ragged, events = [], [], [], []
            for _ in range(6):
                r = synth.make_read(l1, int(rng.integers(1, 1 << 30)), lX=int(rng.integers(20, 250)), anchor_every=int(rng.integers(5, 60)),
                                    noise_dist="wald" if machine == "vanilla" else "gauss")
                ev = np.array(r.events, dtype=np.float64).reshape(-1, 3).copy()
                if machine == "hdp":
                    ev[:, 0] = (ev[:, 0] - r.scale5[1]) / r.scale5[0]
                keep = rng.random(len(r.anchors)) < rng.choice([0.3, 1.0])
                reads.append(r); events.append(ev); anchors.append(r.anchors[keep]); ragged.append((int(rng.integers(0, 2)), int(rng.integers(0, 2))))
            kw = dict(diagonalExpansion=e, minDiagsBetweenTraceBack=mind, traceBackDiagonals=tbd, threshold=0.05)
            batch = HostBatch([r.ref for r in reads], events, anchors, model_ids=[mid] * len(reads),
                              scales=None if machine == "hdp" else [r.scale5 for r in reads], ragged=ragged)
            if machine == "hdp":
                got, res, asg = engine.hdp_expectations_batch(batch, hmm=hmm, params=default_params(**kw))
                want, wasg = np.zeros(10), []
                m = O.Model(O.THREE_STATE_HDP, hdp=hdp_fixture["hdp"])
                for i, r in enumerate(reads):
                    v, a = O.hdp_expectations(m, r.ref, events[i], anchors[i], params=O.default_params(**kw), ragged=ragged[i], pseudocount=0.0)
                    want += v
                    assert np.array_equal(np.asarray(item_pairs(res, asg, i), dtype=np.int64).reshape(-1, 3), a), (i, kw)
                np.testing.assert_allclose(got[:9], want[:9], rtol=1e-9, atol=1e-12)
                assert abs(got[-1] - want[9]) <= 1e-9 * abs(want[9])
            else:
                got, res = engine.expectations_batch(batch, hmm=hmm, params=default_params(**kw))
                want = np.zeros_like(got)
                for i, r in enumerate(reads):
                    m = (O.Model(O.VANILLA, tables=(l1, l2, l3), scale5=r.scale5, strand=0) if machine == "vanilla"
                         else O.Model(O.THREE_STATE, tables=(l1, l2, l3), scale5=r.scale5))
                    want += O.expectations(m, r.ref, events[i], anchors[i], params=O.default_params(**kw), ragged=ragged[i], pseudocount=0.0)
                np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-12)
            assert (res["status"] == 0).all()
            print(machine, kw, "ok")
    finally:
        engine.set_exact_arithmetic(False)
        engine.release_model(mid)


def test_unmodified_vanilla_align_hdp_expectations(tmp_path, hdp_fixture):
    """`vanillaAlign -d -t`: the HdpHmm expectation file (impl/continuousHmm.c:704-749) of the template strand against the
    reference binary's -- header, transition sums and likelihood (six printed decimals), every assigned event mean and
    k-mer in order."""
    if not os.path.exists(EXE):
        pytest.skip("oracle/_ref/vanillaAlign_dropin is built only where /root/reference exists")
    t, c = str(tmp_path / "t.exp"), str(tmp_path / "c.exp")
    va = os.path.join(GOLD, "vanillaAlign")
    args = ["-d", "-v", hdp_fixture["path"], "-w", hdp_fixture["path"],
            "-T", os.path.join(PKG, "models", "template_median68pA.model"), "-C", os.path.join(PKG, "models", "complement_median68pA_pop2.model"),
            "-L", "readA", "-q", os.path.join(GOLD, "ZymoC_ch_1_file1.npRead"), "-r", os.path.join(GOLD, "ZymoRef.txt"), "-t", t, "-c", c]
    with open(os.path.join(va, "guide.cigar")) as fin:
        r = subprocess.run([EXE] + args, stdin=fin, capture_output=True, text=True, timeout=600, env=dict(os.environ, OMP_NUM_THREADS="2"))
    assert r.returncode == 0, r.stderr[-3000:]
    assert "got 13289 HDP assignments" in r.stderr
    got = open(t).read().split("\n")
    with gzip.open(os.path.join(va, "t_d.exp.gz"), "rt") as fh:
        want = fh.read().split("\n")
    assert got[0] == want[0] and len(got) == len(want)              # type, states, threshold, number of assignments
    np.testing.assert_allclose(np.array(got[1].split(), dtype=np.float64), np.array(want[1].split(), dtype=np.float64), rtol=0, atol=2e-6)
    assert got[2] == want[2] and got[3] == want[3]                  # event means, k-mers


def test_em_driver_hdp_estep(engine, zymo, hdp_fixture, tmp_path):
    """cpecan_signal.em.gpu_hdp_estep (single rank): the HdpHmm container -- transition sums, likelihood, the assigned
    event means and k-mers in order -- against the reference's own lists, and its text file in the reference's format."""
    from cpecan_signal import HostBatch, default_params, em, hdp_hmm
    g = hdp_fixture["golden"]
    mid = engine.upload_hdp(hdp_fixture["hdp"])
    batch = HostBatch([zymo["ref"]], [hdp_fixture["events"]], [zymo["anchors_template"]], model_ids=[mid], ragged=[(1, 1)])
    h = em.gpu_hdp_estep(engine, batch, hdp_hmm(), default_params(diagonalExpansion=50, threshold=0.01), distributed=False,
                         container=em.HdpHmm(1e-4, 0.01))
    engine.release_model(mid)
    want = g["hdpexp_e50_r11_vec"]
    np.testing.assert_allclose(h.transitions, want[:9], rtol=1e-9, atol=1e-12)
    assert abs(h.likelihood - want[9]) <= 1e-9 * abs(want[9])
    asg = g["hdpexp_e50_r11_assignments"]
    assert np.array_equal(h.means, hdp_fixture["events"][asg[:, 1], 0])
    assert [em.HdpHmm.kmer_string(k) for k in h.kmers] == [zymo["ref"][x:x + 6] for x in asg[:, 0]]
    out = str(tmp_path / "t.exp")
    h.write(out)
    back = em.HdpHmm.load(out)
    assert len(back.means) == len(asg) and np.array_equal(back.kmers, h.kmers)
