"""Shared comparison helpers for the parity tests.

Tolerances are the ones BASELINE.json's north_star states: total log-probability within 1e-4 relative, posterior match
probabilities within 1e-4 absolute (scores are floor(p * 1e7), so 1000 score units), identical aligned-pair sets except
for pairs whose posterior lies within tolerance of the threshold."""
import numpy as np

P_TOL = 1e-4
SCORE_TOL = int(P_TOL * 1e7) + 2     # +2: floor() on both sides
TOTAL_RTOL = 1e-4


def pair_dict(pairs):
    pairs = np.asarray(pairs).reshape(-1, 3)
    d = {}
    for s, x, y in pairs:
        key = (int(x), int(y))
        assert key not in d, "duplicate aligned pair %r" % (key,)
        d[key] = int(s)
    return d


def compare_pairs(got, want, threshold=0.01):
    """Returns a dict of statistics; raises AssertionError on a violation."""
    g, w = pair_dict(got), pair_dict(want)
    thr_hi = int((threshold + P_TOL) * 1e7) + 2
    worst = 0
    for key, s in w.items():
        if key in g:
            worst = max(worst, abs(g[key] - s))
            assert abs(g[key] - s) <= SCORE_TOL, "pair %r: score %d vs reference %d" % (key, g[key], s)
        else:
            assert s <= thr_hi, "pair %r (score %d) missing and not within tolerance of the threshold" % (key, s)
    for key, s in g.items():
        if key not in w:
            assert s <= thr_hi, "extra pair %r (score %d) not within tolerance of the threshold" % (key, s)
    return dict(n_got=len(g), n_want=len(w), common=len(set(g) & set(w)), worst_score_diff=worst)


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
