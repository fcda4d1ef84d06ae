"""GPU parity of the reference's other two signal state machines -- fourState and echelon -- which
run through the FP64 kernel of cpecan_generic.cuh: the reference's fixture goldens (988 / 988 and 857 / 1000 pairs,
tests/signalPairwiseTest.c:1199-1236, 1388-1449; generated from the unmodified reference by oracle/make_golden.py) and
random small problems against the oracle."""
import os

import numpy as np
import pytest

import parity

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _batch(engine, tables, machine, refs, events, anchors, scales, ragged):
    from cpecan_signal import HostBatch, vanilla_gapx
    l1, l2, l3 = tables
    # getStateMachine4: emissions_signal_initEmissionsToZero (4096 zeros); echelon: the model's skip bins (60)
    gapx = np.zeros(4096) if machine == "four" else vanilla_gapx(l2)
    mid = engine.upload_model(l1, l3, gapx)
    return HostBatch(refs, events, anchors, model_ids=[mid] * len(refs), scales=scales, ragged=ragged), mid


def _same_pairs(got, want, score_tol=2):
    """FP64 on both sides: the same pairs in the same order (echelon lists a pair once per k-mer an event covers, so
    coordinates repeat), scores equal up to the last digit of floor(p * 1e7)."""
    got, want = np.asarray(got, dtype=np.int64).reshape(-1, 3), np.asarray(want, dtype=np.int64).reshape(-1, 3)
    assert got.shape == want.shape, (got.shape, want.shape)
    assert np.array_equal(got[:, 1:], want[:, 1:])
    worst = int(np.abs(got[:, 0] - want[:, 0]).max()) if len(got) else 0
    assert worst <= score_tol, worst
    return worst


@pytest.mark.parametrize("tag,e,ragged,count", [("four_e20_r11", 20, (1, 1), 988), ("four_e20_r00", 20, (0, 0), 988),
                                                 ("four_e50_r10", 50, (1, 0), None)])
def test_fixture_four_state(engine, zymo, template_tables, tag, e, ragged, count):
    from cpecan_signal import default_params, four_state_hmm
    from cpecan_signal.engine import item_pairs
    g = dict(np.load(os.path.join(GOLDEN, "zymo_four_state_golden.npz")))
    rd = zymo["read"]
    batch, mid = _batch(engine, template_tables, "four", [zymo["ref"]], [rd["template_events"]], [zymo["anchors_template"]],
                        [rd["template_params"]], [ragged])
    res, pairs, totals = engine.align_batch(batch, hmm=four_state_hmm(), params=default_params(diagonalExpansion=e),
                                            want_totals=True)
    engine.release_model(mid)
    assert res[0]["status"] == 0
    got = parity.reverse_regions(item_pairs(res, pairs, 0))
    worst = _same_pairs(got, g[tag + "_pairs"])
    want_t = g[tag + "_totals"]
    mask = ~np.isnan(want_t)
    assert np.array_equal(mask, ~np.isnan(totals[0]))
    dt = float(np.abs(totals[0][mask] - want_t[mask]).max())
    print(tag, len(got), "worst score diff", worst, "worst |total diff|", dt)
    assert dt <= 1e-9 * float(np.abs(want_t[mask]).max())
    if count is not None:
        assert len(got) == count


def test_fixture_four_state_unbanded(engine, zymo, template_tables):
    from cpecan_signal import four_state_hmm
    from cpecan_signal.engine import MODE_UNBANDED, item_pairs
    g = dict(np.load(os.path.join(GOLDEN, "zymo_four_state_golden.npz")))
    rd = zymo["read"]
    batch, mid = _batch(engine, template_tables, "four", [zymo["ref"]], [rd["template_events"]], [np.zeros((0, 2))],
                        [rd["template_params"]], [(1, 1)])
    res, pairs, _ = engine.align_batch(batch, hmm=four_state_hmm(), mode=MODE_UNBANDED)
    engine.release_model(mid)
    assert res[0]["status"] == 0
    got = item_pairs(res, pairs, 0)
    want = g["four_unbanded_r11_pairs"]
    st = parity.compare_pairs(got, want)
    assert st["n_got"] == st["n_want"] == 988 and st["worst_score_diff"] <= 2
    assert abs(res[0]["total_logprob"] - float(g["four_unbanded_r11_total"])) <= 1e-9 * abs(float(g["four_unbanded_r11_total"]))


@pytest.mark.parametrize("tag,e,ragged,count", [("echelon_e20_r00_t15", 20, (0, 0), 857), ("echelon_e50_r10_t15", 50, (1, 0), None)])
def test_fixture_echelon(engine, zymo, template_tables, tag, e, ragged, count):
    from cpecan_signal import default_params, echelon_hmm
    from cpecan_signal.engine import item_pairs
    g = dict(np.load(os.path.join(GOLDEN, "zymo_echelon_golden.npz")))
    rd = zymo["read"]
    batch, mid = _batch(engine, template_tables, "echelon", [zymo["ref"]], [rd["template_events"]], [zymo["anchors_template"]],
                        [rd["template_params"]], [ragged])
    res, pairs, totals = engine.align_batch(batch, hmm=echelon_hmm(), params=default_params(diagonalExpansion=e, threshold=0.15),
                                            want_totals=True, pair_cap=40000)
    engine.release_model(mid)
    assert res[0]["status"] == 0
    got = parity.reverse_regions(item_pairs(res, pairs, 0))
    worst = _same_pairs(got, g[tag + "_pairs"])
    want_t = g[tag + "_totals"]
    mask = ~np.isnan(want_t)
    assert np.array_equal(mask, ~np.isnan(totals[0]))
    dt = float(np.abs(totals[0][mask] - want_t[mask]).max())
    print(tag, len(got), "worst score diff", worst, "worst |total diff|", dt)
    assert dt <= 1e-9 * float(np.abs(want_t[mask]).max())
    if count is not None:
        assert len(got) == count


def test_fixture_echelon_unbanded(engine, zymo, template_tables):
    """1000 pairs without banding at threshold 0.15 (tests/signalPairwiseTest.c:1440-1447)."""
    from cpecan_signal import default_params, echelon_hmm
    from cpecan_signal.engine import MODE_UNBANDED, item_pairs
    g = dict(np.load(os.path.join(GOLDEN, "zymo_echelon_golden.npz")))
    rd = zymo["read"]
    batch, mid = _batch(engine, template_tables, "echelon", [zymo["ref"]], [rd["template_events"]], [np.zeros((0, 2))],
                        [rd["template_params"]], [(0, 0)])
    res, pairs, _ = engine.align_batch(batch, hmm=echelon_hmm(), params=default_params(threshold=0.15), mode=MODE_UNBANDED,
                                       pair_cap=40000)
    engine.release_model(mid)
    assert res[0]["status"] == 0
    got = np.asarray(item_pairs(res, pairs, 0), dtype=np.int64)
    want = np.asarray(g["echelon_unbanded_r00_t15_pairs"], dtype=np.int64)
    assert len(got) == len(want) == 1000
    # the reference lists ascending diagonals here, the device descending ones: compare as sorted multisets
    key = lambda p: np.lexsort((p[:, 0], p[:, 2], p[:, 1]))
    gs, ws = got[key(got)], want[key(want)]
    assert np.array_equal(gs[:, 1:], ws[:, 1:]) and int(np.abs(gs[:, 0] - ws[:, 0]).max()) <= 2
    assert abs(res[0]["total_logprob"] - float(g["echelon_unbanded_r00_t15_total"])) <= 1e-9 * abs(float(g["echelon_unbanded_r00_t15_total"]))


@pytest.mark.parametrize("machine", ["four", "echelon"])
def test_random_small_problems_vs_oracle(engine, template_tables, machine):
    """Fuzz as for the three-state / vanilla kernels: random small reads, traceback-heavy parameter sets, sparse or no
    anchors, every ragged combination -- each item against the oracle (same pair lists, totals to 1e-9)."""
    import oracleshim as O
    from cpecan_signal import default_params, echelon_hmm, four_state_hmm, synth
    from cpecan_signal.engine import item_pairs
    l1, l2, l3 = template_tables
    rng = np.random.default_rng(91 if machine == "four" else 92)
    smt = O.FOUR_STATE if machine == "four" else O.ECHELON
    hmm = four_state_hmm() if machine == "four" else echelon_hmm()
    for mind, tbd, e, thr in [(5, 2, 4, 0.01), (20, 8, 10, 0.2), (60, 40, 20, 0.05), (1000, 40, 30, 0.01)]:
        reads, anchors, ragged = [], [], []
        for _ in range(10):
            r = synth.make_read(l1, int(rng.integers(1, 1 << 30)), lX=int(rng.integers(8, 150)),
                                anchor_every=int(rng.integers(5, 60)), noise_dist="gauss" if machine == "four" else "wald")
            keep = rng.random(len(r.anchors)) < rng.choice([0.0, 0.3, 1.0])
            reads.append(r); anchors.append(r.anchors[keep]); ragged.append((int(rng.integers(0, 2)), int(rng.integers(0, 2))))
        batch, mid = _batch(engine, template_tables, machine, [r.ref for r in reads], [r.events for r in reads], anchors,
                            [r.scale5 for r in reads], ragged)
        kw = dict(diagonalExpansion=e, minDiagsBetweenTraceBack=mind, traceBackDiagonals=tbd, threshold=thr)
        res, pairs, totals = engine.align_batch(batch, hmm=hmm, params=default_params(**kw), want_totals=True,
                                                pair_cap=200000)
        engine.release_model(mid)
        worst = 0
        for i, r in enumerate(reads):
            m = O.Model(smt, tables=(l1, l2, l3), scale5=r.scale5)
            want, wtot = O.align_banded(m, r.ref, r.events, anchors[i], params=O.default_params(**kw), ragged=ragged[i],
                                        want_totals=True)
            assert res[i]["status"] == 0, (i, kw)
            worst = max(worst, _same_pairs(parity.reverse_regions(item_pairs(res, pairs, i)), want))
            mask = ~np.isnan(wtot)
            assert np.array_equal(mask, ~np.isnan(totals[i]))
            if mask.any():
                assert float(np.abs(totals[i][mask] - wtot[mask]).max()) <= 1e-9 * max(1.0, float(np.abs(wtot[mask]).max()))
        print(machine, kw, "worst score diff", worst)


def test_expectations_refused(engine, template_tables):
    from cpecan_signal import EngineError, four_state_hmm, synth
    r = synth.make_read(template_tables[0], 5, lX=100)
    batch, mid = _batch(engine, template_tables, "four", [r.ref], [r.events], [r.anchors], [r.scale5], [(1, 1)])
    with pytest.raises(EngineError, match="not defined"):
        engine.expectations_batch(batch, hmm=four_state_hmm())
    engine.release_model(mid)
