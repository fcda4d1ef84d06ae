"""CPU tests: the oracle restatement (oracle/cpecan_oracle.c) against
  * the golden vectors the UNMODIFIED reference produced (tests/golden/*.npz, generator oracle/make_golden.py), bit-exact;
  * the known-answer vectors the reference's own CuTest suites hold for this path:
      tests/pairwiseAlignerTest.c:74-137 (band walk), :596-665 (split points), :139-149 (logAdd property),
      tests/signalPairwiseTest.c:580-685 (8 pairs), :795-897 (5 pairs), :1163,1173,1293,1303 (987/986/999/953)."""
import os

import numpy as np
import pytest

import oracleshim as O


def _model(smt, tables, scale5=None, strand=None):
    # the goldens were made with stateMachine3Vanilla_setStrandTransitionsToDefaults(template) (oracle/ref_shim.c),
    # except the tiny case, which uses the constructor's values (strand=None)
    return O.Model(smt, tables=tables, scale5=scale5, strand=strand)


# ------------------------------------------------------------------------------------------ geometry known answers
def test_band_walk_known_answer():
    """tests/pairwiseAlignerTest.c:74-137: anchors (1,0),(2,1),(3,3), lX=6, lY=5, expansion 2."""
    want = [(0, 0, 0), (1, -1, 1), (2, -2, 2), (3, -1, 3), (4, -2, 4), (5, -1, 3), (6, -2, 4), (7, -3, 3), (8, -2, 2),
            (9, -1, 3), (10, 0, 2), (11, 1, 1)]
    got = O.band([(1, 0), (2, 1), (3, 3)], 6, 5, 2)
    assert [tuple(int(v) for v in r) for r in got] == want


def test_split_points_known_answer():
    """tests/pairwiseAlignerTest.c:596-665."""
    ms = 2000 * 2000
    none = np.zeros((0, 2), dtype=np.int64)
    assert O.split_points(none, 3000, 1000, ms, 0, 0).tolist() == [[0, 0, 3000, 1000]]
    lX, lY = 20000, 25000
    assert O.split_points(none, lX, lY, ms, 1, 1).tolist() == []
    assert O.split_points(none, lX, lY, ms, 1, 0).tolist() == [[18000, 23000, lX, lY]]
    assert O.split_points(none, lX, lY, ms, 0, 1).tolist() == [[0, 0, 2000, 2000]]
    assert O.split_points(none, lX, lY, ms, 0, 0).tolist() == [[0, 0, 2000, 2000], [18000, 23000, lX, lY]]
    anchors = [(2000, 2000), (4002, 4001), (5000, 5000), (8000, 6000), (9000, 9000), (10000, 14000), (15000, 15000),
               (16000, 16000)]
    assert O.split_points(anchors, lX, lY, ms, 0, 0).tolist() == [
        [0, 0, 3001, 3001], [3002, 3001, 9500, 11001], [9501, 12000, 12001, 14500], [13000, 14501, 18000, 18001],
        [18001, 23000, 20000, 25000]]


def test_log_add_property():
    """tests/pairwiseAlignerTest.c:139-149: |exp(logAdd(log i, log j)) - (i + j)| < 1e-3 over random pairs."""
    rng = np.random.default_rng(7)
    for i, j in rng.random((20000, 2)):
        k = i + j
        assert abs(np.exp(O.log_add(np.log(i), np.log(j))) - k) < 1e-3
    assert O.log_add(-np.inf, -3.0) == -3.0 and O.log_add(-3.0, -np.inf) == -3.0
    assert O.log_add(0.0, -7.5) == 0.0           # hard cut-off (impl/pairwiseAligner.c:251-255)
    assert O.log_add(-1.0, -2.0) == O.log_add(-2.0, -1.0)


def test_filter_overlap_monotone():
    rng = np.random.default_rng(3)
    pts = np.cumsum(rng.integers(-2, 6, size=(200, 2)), axis=0)
    out = O.filter_overlap(pts)
    assert len(out) > 0 and (np.diff(out[:, 0]) > 0).all() and (np.diff(out[:, 1]) > 0).all()


# ------------------------------------------------------------------------------------------ reference goldens
def test_tiny_known_answers(zymo, template_tables):
    """tests/signalPairwiseTest.c:580-685 (strawMan, 8 pairs at 0.2) and :795-897 (vanilla, 5 pairs at 0.5)."""
    ev = np.array([58.743435, 0.887833, 0.0571, 53.604965, 0.816836, 0.0571, 58.432015, 0.735143, 0.0571,
                   63.684352, 0.795437, 0.0571, 58.921430, 0.812959, 0.0571, 59.895882, 0.740952, 0.0571,
                   61.684303, 0.722332, 0.0571]).reshape(-1, 3)
    pairs, total = O.align_unbanded(_model(O.THREE_STATE, template_tables), "ACGATACGGACAT", ev,
                                    params=O.default_params(threshold=0.2))
    assert {(int(x), int(y)) for _, x, y in pairs} == {(0, 0), (1, 1), (2, 2), (3, 3), (4, 3), (5, 4), (6, 5), (7, 6)}
    assert np.array_equal(pairs, zymo["tiny_three_pairs"])
    assert total == float(zymo["tiny_three_total"])
    pairs, _ = O.align_unbanded(_model(O.VANILLA, template_tables), "ACGATACGGACAT", ev,
                                params=O.default_params(threshold=0.5))
    assert {(int(x), int(y)) for _, x, y in pairs} == {(2, 0), (3, 3), (5, 4), (6, 5), (7, 6)}
    assert np.array_equal(pairs, zymo["tiny_vanilla_pairs"])


@pytest.mark.parametrize("tag,smt,e,ragged,count", [
    ("three_e20_r00", O.THREE_STATE, 20, (0, 0), 987), ("vanilla_e20_r00", O.VANILLA, 20, (0, 0), 999),
    ("three_e50_r11", O.THREE_STATE, 50, (1, 1), None), ("vanilla_e50_r11", O.VANILLA, 50, (1, 1), None)])
def test_fixture_banded_bit_exact(zymo, template_tables, tag, smt, e, ragged, count):
    """BASELINE config 1 (tests/signalPairwiseTest.c:1116-1183, :1250-1310): identical (score, x, y) lists and identical
    per-diagonal totalProbability values as the unmodified reference."""
    rd = zymo["read"]
    m = _model(smt, template_tables, scale5=rd["template_params"], strand=0)
    pairs, totals = O.align_banded(m, zymo["ref"], rd["template_events"], zymo["anchors_template"],
                                   params=O.default_params(diagonalExpansion=e), ragged=ragged, want_totals=True)
    assert np.array_equal(pairs, zymo[tag + "_pairs"])
    assert np.array_equal(totals, zymo[tag + "_totals"], equal_nan=True)
    if count is not None:
        assert len(pairs) == count
    assert len({(int(x), int(y)) for _, x, y in pairs}) == len(pairs)
    assert pairs[:, 0].min() >= int(0.01 * 1e7) and pairs[:, 0].max() <= 10_000_000


@pytest.mark.parametrize("tag,e,ragged,count", [("four_e20_r11", 20, (1, 1), 988), ("four_e20_r00", 20, (0, 0), 988),
                                                  ("four_e50_r10", 50, (1, 0), None)])
def test_fixture_four_state_bit_exact(zymo, template_tables, tag, e, ragged, count):
    """The fourState machine (tests/signalPairwiseTest.c:1199-1236: 988 pairs with and without banding): the oracle's
    four-state cell, start / end vectors and S-state dot products against goldens of the unmodified reference.  (The
    device path does not build this machine yet -- DESIGN.md section 9; the oracle is pinned ahead of it.)"""
    g = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "zymo_four_state_golden.npz")))
    rd = zymo["read"]
    m = O.Model(O.FOUR_STATE, tables=template_tables, scale5=rd["template_params"])
    pairs, totals = O.align_banded(m, zymo["ref"], rd["template_events"], zymo["anchors_template"],
                                   params=O.default_params(diagonalExpansion=e), ragged=ragged, want_totals=True)
    assert np.array_equal(pairs, g[tag + "_pairs"])
    assert np.array_equal(totals, g[tag + "_totals"], equal_nan=True)
    if count is not None:
        assert len(pairs) == count
    if tag == "four_e20_r11":
        up, ut = O.align_unbanded(m, zymo["ref"], rd["template_events"], ragged=(1, 1))
        assert np.array_equal(up, g["four_unbanded_r11_pairs"]) and ut == float(g["four_unbanded_r11_total"])
        assert len(up) == 988


@pytest.mark.parametrize("tag,e,ragged,count", [("echelon_e20_r00_t15", 20, (0, 0), 857), ("echelon_e50_r10_t15", 50, (1, 0), None)])
def test_fixture_echelon_bit_exact(zymo, template_tables, tag, e, ragged, count):
    """The echelon machine (tests/signalPairwiseTest.c:1388-1449: 857 pairs banded, 1000 without banding at threshold
    0.15): seven states, events covering 1 .. 5 k-mers read from the padded sequence, Poisson duration term, the
    multi-state posterior.  Oracle pinned ahead of the device path, as for fourState."""
    g = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "zymo_echelon_golden.npz")))
    rd = zymo["read"]
    m = O.Model(O.ECHELON, tables=template_tables, scale5=rd["template_params"])
    prm = O.default_params(diagonalExpansion=e, threshold=0.15)
    pairs, totals = O.align_banded(m, zymo["ref"], rd["template_events"], zymo["anchors_template"], params=prm,
                                   ragged=ragged, want_totals=True)
    assert np.array_equal(pairs, g[tag + "_pairs"])
    assert np.array_equal(totals, g[tag + "_totals"], equal_nan=True)
    if count is not None:
        assert len(pairs) == count
        up, ut = O.align_unbanded(m, zymo["ref"], rd["template_events"], params=prm, ragged=(0, 0))
        assert np.array_equal(up, g["echelon_unbanded_r00_t15_pairs"]) and ut == float(g["echelon_unbanded_r00_t15_total"])
        assert len(up) == 1000


def test_fixture_band_size(zymo):
    """39 filtered anchors, 1692 diagonals, 140 469 band cells at e=20."""
    rd = zymo["read"]
    lX, lY = len(zymo["ref"]) - 5, len(rd["template_events"])
    assert (lX, lY) == (892, 799) and len(zymo["anchors_template"]) == 39
    b = O.band(zymo["anchors_template"], lX, lY, 20)
    assert len(b) == 1692 and int(((b[:, 2] - b[:, 1]) // 2 + 1).sum()) == 140469
    assert O.band_cells(zymo["anchors_template"], lX, lY, O.default_params(diagonalExpansion=20)) == 140469


@pytest.mark.parametrize("tag,smt,count", [("three", O.THREE_STATE, 986), ("vanilla", O.VANILLA, 953)])
def test_fixture_unbanded_bit_exact(zymo, template_tables, tag, smt, count):
    rd = zymo["read"]
    m = _model(smt, template_tables, scale5=rd["template_params"], strand=0)
    pairs, total = O.align_unbanded(m, zymo["ref"], rd["template_events"])
    assert len(pairs) == count
    assert np.array_equal(pairs, zymo[tag + "_unbanded_pairs"])
    assert total == float(zymo[tag + "_unbanded_total"])


@pytest.mark.parametrize("tag,smt", [("three", O.THREE_STATE), ("vanilla", O.VANILLA)])
@pytest.mark.parametrize("cfg,e,ragged", [("e20_r00", 20, (0, 0)), ("e50_r11", 50, (1, 1))])
def test_fixture_expectations(zymo, template_tables, tag, smt, cfg, e, ragged):
    """E-step (getExpectationsUsingAnchors, impl/pairwiseAligner.c:1571-1591) on the fixture.  Sums of ~1e5 doubles in
    the same order: the restatement reproduces the reference to rounding (<= 1e-12 relative)."""
    rd = zymo["read"]
    m = _model(smt, template_tables, scale5=rd["template_params"], strand=0)
    got = O.expectations(m, zymo["ref"], rd["template_events"], zymo["anchors_template"],
                         params=O.default_params(diagonalExpansion=e), ragged=ragged)
    want = zymo["%s_expectations_%s" % (tag, cfg)]
    assert got.shape == want.shape
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-15)


def test_fixture_complement(zymo):
    """Complement strand of the 2D fixture read against the reverse-complemented reference (vanillaAlign.c:635,772)."""
    from cpecan_signal import synth
    rd = zymo["read"]
    ref_rc = zymo["ref"][::-1].translate(str.maketrans("ACGT", "TGCA"))
    m = O.Model(O.THREE_STATE, model_file=synth.COMPLEMENT_MODEL, scale5=rd["complement_params"])
    pairs, totals = O.align_banded(m, ref_rc, rd["complement_events"], zymo["anchors_complement"],
                                   params=O.default_params(diagonalExpansion=50), ragged=(1, 1), want_totals=True)
    assert np.array_equal(pairs, zymo["three_e50_r11_complement_pairs"])
    assert np.array_equal(totals, zymo["three_e50_r11_complement_totals"], equal_nan=True)


@pytest.mark.parametrize("tag", ["s0", "s1", "s2", "s3", "s4"])
def test_synthetic_goldens_bit_exact(syn_golden, template_tables, tag):
    from cpecan_signal import synth
    idx, lX, e, r0, r1, every, mind = (int(v) for v in syn_golden[tag + "_meta"])
    r = synth.make_read(template_tables[0], idx, lX=lX, anchor_every=every)
    m = _model(O.THREE_STATE, template_tables, scale5=r.scale5)
    p = O.default_params(diagonalExpansion=e, minDiagsBetweenTraceBack=mind)
    pairs, totals = O.align_banded(m, r.ref, r.events, r.anchors, params=p, ragged=(r0, r1), want_totals=True)
    assert np.array_equal(pairs, syn_golden[tag + "_pairs"])
    assert np.array_equal(totals, syn_golden[tag + "_totals"], equal_nan=True)
    got = O.expectations(m, r.ref, r.events, r.anchors, params=p, ragged=(r0, r1))
    np.testing.assert_allclose(got, syn_golden[tag + "_expect"], rtol=1e-12, atol=1e-15)


def test_synthetic_split_regions(syn_golden, template_tables):
    """An anchor gap above splitMatrixBiggerThanThis: getSplitPoints + ragged sub-alignments (impl/pairwiseAligner.c:1356-1422)."""
    from cpecan_signal import synth
    idx, lX, e, r0, r1, every, mind, split = (int(v) for v in syn_golden["split_meta"])
    lo, hi = (int(v) for v in syn_golden["split_keep_lo_hi"])
    r = synth.make_read(template_tables[0], idx, lX=lX, anchor_every=every)
    keep = (r.anchors[:, 0] < lo) | (r.anchors[:, 0] > hi)
    m = _model(O.THREE_STATE, template_tables, scale5=r.scale5)
    p = O.default_params(diagonalExpansion=e, splitMatrixBiggerThanThis=split)
    assert np.array_equal(O.split_points(r.anchors[keep], r.lX, r.lY, split, r0, r1), syn_golden["split_points"])
    pairs, _ = O.align_banded(m, r.ref, r.events, r.anchors[keep], params=p, ragged=(r0, r1))
    assert np.array_equal(pairs, syn_golden["split_pairs"])
