"""CPU test: pins the oracle restatement bit-for-bit against the UNMODIFIED reference sources compiled into
oracle/_ref/libcpecan_ref.so, on seeded random inputs beyond the committed goldens.  Skipped where oracle/_ref was
never built (a machine without /root/reference and without the travelled .so)."""
import numpy as np
import pytest

import oracleshim as O
import refshim as R

pytestmark = pytest.mark.skipif(not R.available(), reason="oracle/_ref/libcpecan_ref.so not built (no /root/reference)")


def test_log_add_bit_exact():
    rng = np.random.default_rng(11)
    xs = np.concatenate([rng.uniform(-40, 5, 4000), [-np.inf, 0.0, -7.5, -1.0, -2.5, -4.5]])
    ys = np.concatenate([rng.uniform(-40, 5, 4000), [-3.0, -7.5, 0.0, 0.0, 0.0, 0.0]])
    for x, y in zip(xs, ys):
        assert O.log_add(x, y) == R.log_add(x, y)


def test_band_and_split_bit_exact():
    rng = np.random.default_rng(5)
    for _ in range(40):
        lX, lY = int(rng.integers(1, 400)), int(rng.integers(1, 400))
        n = int(rng.integers(0, 12))
        pts = np.stack([np.sort(rng.choice(lX, size=min(n, lX), replace=False)),
                        np.sort(rng.choice(lY, size=min(n, lX), replace=True))], axis=1) if n else np.zeros((0, 2), np.int64)
        pts = O.filter_overlap(pts)
        assert np.array_equal(pts, R.filter_overlap(pts))
        e = int(rng.choice([2, 4, 10, 20, 64]))
        assert np.array_equal(O.band(pts, lX, lY, e), R.band(pts, lX, lY, e))
        ms = int(rng.choice([50 * 50, 200 * 200, 3000 * 3000]))
        for rl, rr in ((0, 0), (1, 0), (0, 1), (1, 1)):
            assert np.array_equal(O.split_points(pts, lX, lY, ms, rl, rr), R.split_points(pts, lX, lY, ms, rl, rr))


@pytest.mark.parametrize("smt", [O.THREE_STATE, O.VANILLA, O.FOUR_STATE, O.ECHELON])
@pytest.mark.parametrize("idx,lX,e,ragged,every,mind", [(40, 260, 20, (1, 1), 50, 1000), (41, 700, 64, (0, 1), 50, 300),
                                                         (42, 420, 10, (1, 0), 200, 1000)])
def test_random_reads_bit_exact(template_tables, smt, idx, lX, e, ragged, every, mind):
    from cpecan_signal import synth
    r = synth.make_read(template_tables[0], idx, lX=lX, anchor_every=every,
                        noise_dist="wald" if smt in (O.VANILLA, O.ECHELON) else "gauss")
    m = O.Model(smt, tables=template_tables, scale5=r.scale5, strand=0)
    po = O.default_params(diagonalExpansion=e, minDiagsBetweenTraceBack=mind)
    pr = R.default_params(diagonalExpansion=e, minDiagsBetweenTraceBack=mind)
    got, gt = O.align_banded(m, r.ref, r.events, r.anchors, params=po, ragged=ragged, want_totals=True)
    want, wt = R.align_banded(smt, synth.TEMPLATE_MODEL, r.ref, r.events, r.anchors, params=pr, scale5=r.scale5,
                              strand=0, ragged=ragged, want_totals=True)
    assert np.array_equal(got, want)
    assert np.array_equal(gt, wt, equal_nan=True)
    if smt in (O.FOUR_STATE, O.ECHELON):
        return                     # the reference has no expectation container for this machine
    ge = O.expectations(m, r.ref, r.events, r.anchors, params=po, ragged=ragged)
    we = R.expectations(smt, synth.TEMPLATE_MODEL, r.ref, r.events, r.anchors, params=pr, scale5=r.scale5, strand=0,
                        ragged=ragged)
    np.testing.assert_allclose(ge, we, rtol=1e-12, atol=1e-15)
