"""CPU fuzz of the oracle against the live reference (oracle/_ref): small random problems that exercise what the fixed
cases cannot all reach at once -- several tracebacks per read (small minDiagsBetweenTraceBack), split regions (small
splitMatrixBiggerThanThis), sparse and dense anchors, tiny expansions, every ragged combination, every machine."""
import numpy as np
import pytest

import oracleshim as O
import refshim as R

pytestmark = pytest.mark.skipif(not R.available(), reason="oracle/_ref not built")


def _random_case(rng, template_tables):
    from cpecan_signal import synth
    lX = int(rng.integers(8, 90))
    r = synth.make_read(template_tables[0], int(rng.integers(1, 1 << 30)), lX=lX, anchor_every=int(rng.integers(5, 60)),
                        noise_dist="wald")
    keep = rng.random(len(r.anchors)) < rng.choice([0.0, 0.3, 1.0])
    anchors = r.anchors[keep]
    # the reference itself needs minDiagsBetweenTraceBack > traceBackDiagonals + 1 (it walks back past tracedBackTo otherwise)
    mind, tbd = [(5, 2), (20, 2), (20, 8), (60, 8), (60, 40), (1000, 40)][int(rng.integers(0, 6))]
    prm = dict(diagonalExpansion=int(rng.choice([2, 4, 10, 20])), minDiagsBetweenTraceBack=mind,
               traceBackDiagonals=tbd, threshold=float(rng.choice([0.01, 0.2])),
               splitMatrixBiggerThanThis=int(rng.choice([50, 400, 3000 * 3000])))
    ragged = (int(rng.integers(0, 2)), int(rng.integers(0, 2)))
    return r, anchors, prm, ragged


@pytest.mark.parametrize("smt", [O.THREE_STATE, O.VANILLA, O.FOUR_STATE, O.ECHELON])
def test_random_small_problems_bit_exact(template_tables, smt):
    from cpecan_signal import synth
    rng = np.random.default_rng(1000 + smt)
    for _ in range(12 if smt == O.ECHELON else 25):
        r, anchors, prm, ragged = _random_case(rng, template_tables)
        m = O.Model(smt, tables=template_tables, scale5=r.scale5, strand=0)
        got, gt = O.align_banded(m, r.ref, r.events, anchors, params=O.default_params(**prm), ragged=ragged, want_totals=True)
        want, wt = R.align_banded(smt, synth.TEMPLATE_MODEL, r.ref, r.events, anchors, params=R.default_params(**prm),
                                  scale5=r.scale5, strand=0, ragged=ragged, want_totals=True)
        assert np.array_equal(got, want), (smt, r.lX, r.lY, prm, ragged)
        if prm["splitMatrixBiggerThanThis"] > 1000:          # totals are indexed per region; compare them unsplit only
            assert np.array_equal(gt, wt, equal_nan=True), (smt, r.lX, r.lY, prm, ragged)
        if smt in (O.THREE_STATE, O.VANILLA):
            ge = O.expectations(m, r.ref, r.events, anchors, params=O.default_params(**prm), ragged=ragged)
            we = R.expectations(smt, synth.TEMPLATE_MODEL, r.ref, r.events, anchors, params=R.default_params(**prm),
                                scale5=r.scale5, strand=0, ragged=ragged)
            np.testing.assert_allclose(ge, we, rtol=1e-11, atol=1e-14)


@pytest.mark.skipif(not R.hdp_available(), reason="oracle/_ref/libcpecan_ref_hdp.so not built")
def test_random_small_problems_hdp_bit_exact(template_tables, hdp_fixture):
    """threeStateHdp: the oracle against the reference built with its own HDP sources -- pair lists, scores and totals
    bit-identical, expectation sums to 1e-11, the event-to-k-mer assignment lists identical."""
    rng = np.random.default_rng(1007)
    m = O.Model(O.THREE_STATE_HDP, hdp=hdp_fixture["hdp"])
    for _ in range(25):
        r, anchors, prm, ragged = _random_case(rng, template_tables)
        ev = np.array(r.events, dtype=np.float64).reshape(-1, 3).copy()
        ev[:, 0] = (ev[:, 0] - r.scale5[1]) / r.scale5[0]               # the HDP machine takes descaled events
        got, gt = O.align_banded(m, r.ref, ev, anchors, params=O.default_params(**prm), ragged=ragged, want_totals=True)
        want, wt = R.hdp_align_banded(hdp_fixture["path"], r.ref, ev, anchors, params=R.default_params(**prm), ragged=ragged,
                                      want_totals=True)
        assert np.array_equal(got, want), (r.lX, r.lY, prm, ragged)
        if prm["splitMatrixBiggerThanThis"] > 1000:
            assert np.array_equal(gt, wt, equal_nan=True), (r.lX, r.lY, prm, ragged)
        gv, ga = O.hdp_expectations(m, r.ref, ev, anchors, params=O.default_params(**prm), ragged=ragged, pseudocount=1e-4)
        wv, wa = R.hdp_expectations(hdp_fixture["path"], r.ref, ev, anchors, params=R.default_params(**prm), ragged=ragged,
                                    pseudocount=1e-4, threshold=prm["threshold"])
        np.testing.assert_allclose(gv, wv, rtol=1e-11, atol=1e-14)
        assert np.array_equal(ga[:, 1:], wa), (r.lX, r.lY, prm, ragged)
