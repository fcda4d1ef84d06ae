"""GPU parity: the CUDA path (through the C-ABI of include/cpecan_cuda.h) against the golden vectors produced by the
unmodified reference (tests/golden/*.npz, generator oracle/make_golden.py) and against the oracle restatement on
seeded synthetic reads."""
import os

import numpy as np
import pytest

import parity

pytestmark = pytest.mark.gpu


def _three_state_batch(eng, tables, refs, events, anchors, scales, ragged):
    from cpecan_signal import HostBatch
    l1, _, l3 = tables
    mid = eng.upload_model(l1, l3, np.full(4096, -2.3025850929940455))
    return HostBatch(refs, events, anchors, model_ids=[mid] * len(refs), scales=scales, ragged=ragged)


@pytest.mark.parametrize("tag,e,ragged", [("three_e20_r00", 20, (0, 0)), ("three_e50_r11", 50, (1, 1))])
def test_fixture_banded_three_state(engine, zymo, template_tables, tag, e, ragged):
    """BASELINE config 1: reference tests/signalPairwiseTest.c:1116-1183 (987 aligned pairs)."""
    from cpecan_signal import default_params
    from cpecan_signal.engine import item_pairs
    rd = zymo["read"]
    batch = _three_state_batch(engine, template_tables, [zymo["ref"]], [rd["template_events"]],
                               [zymo["anchors_template"]], [rd["template_params"]], [ragged])
    res, pairs, totals = engine.align_batch(batch, params=default_params(diagonalExpansion=e), want_totals=True)
    assert res[0]["status"] == 0
    got = item_pairs(res, pairs, 0)
    want = zymo[tag + "_pairs"]
    stats = parity.compare_pairs(got, want)
    worst_total = parity.compare_totals(totals[0], zymo[tag + "_totals"])
    print(tag, stats, "worst |total diff|", worst_total, "cells", res[0]["band_cells"])
    assert stats["n_got"] == stats["n_want"] == 987
    # emission order: the device list reversed is the reference's list
    if stats["n_got"] == stats["n_want"]:
        assert np.array_equal(parity.reverse_regions(got)[:, 1:], want[:, 1:])
    if e == 20:
        assert res[0]["band_cells"] == 140469


def test_fixture_unbanded_three_state(engine, zymo, template_tables):
    """getAlignedPairsWithoutBanding on the fixture: 986 pairs (tests/signalPairwiseTest.c:1166-1173)."""
    from cpecan_signal.engine import MODE_UNBANDED, item_pairs
    rd = zymo["read"]
    batch = _three_state_batch(engine, template_tables, [zymo["ref"]], [rd["template_events"]], [np.zeros((0, 2))],
                               [rd["template_params"]], [(0, 0)])
    res, pairs, _ = engine.align_batch(batch, mode=MODE_UNBANDED)
    assert res[0]["status"] == 0
    got = item_pairs(res, pairs, 0)
    stats = parity.compare_pairs(got, zymo["three_unbanded_pairs"])
    print(stats, res[0]["total_logprob"], float(zymo["three_unbanded_total"]))
    assert abs(res[0]["total_logprob"] - float(zymo["three_unbanded_total"])) <= 1e-4 * abs(float(zymo["three_unbanded_total"]))
    assert stats["n_got"] == stats["n_want"] == 986


def test_tiny_known_answer(engine, zymo, template_tables):
    """tests/signalPairwiseTest.c:580-685: exactly 8 pairs {(0,0),(1,1),(2,2),(3,3),(4,3),(5,4),(6,5),(7,6)} at 0.2."""
    from cpecan_signal import default_params
    from cpecan_signal.engine import MODE_UNBANDED, item_pairs
    ev = np.array([58.743435, 0.887833, 0.0571, 53.604965, 0.816836, 0.0571, 58.432015, 0.735143, 0.0571,
                   63.684352, 0.795437, 0.0571, 58.921430, 0.812959, 0.0571, 59.895882, 0.740952, 0.0571,
                   61.684303, 0.722332, 0.0571]).reshape(-1, 3)
    batch = _three_state_batch(engine, template_tables, ["ACGATACGGACAT"], [ev], [np.zeros((0, 2))], None, [(0, 0)])
    res, pairs, _ = engine.align_batch(batch, params=default_params(threshold=0.2), mode=MODE_UNBANDED)
    got = item_pairs(res, pairs, 0)
    assert {(int(x), int(y)) for _, x, y in got} == {(0, 0), (1, 1), (2, 2), (3, 3), (4, 3), (5, 4), (6, 5), (7, 6)}
    parity.compare_pairs(got, zymo["tiny_three_pairs"], threshold=0.2)


@pytest.mark.parametrize("tag", ["s0", "s1", "s2", "s3", "s4"])
def test_synthetic_golden(engine, syn_golden, template_tables, tag):
    from cpecan_signal import default_params, synth
    from cpecan_signal.engine import item_pairs
    idx, lX, e, r0, r1, every, mind = (int(v) for v in syn_golden[tag + "_meta"])
    r = synth.make_read(template_tables[0], idx, lX=lX, anchor_every=every)
    batch = _three_state_batch(engine, template_tables, [r.ref], [r.events], [r.anchors], [r.scale5], [(r0, r1)])
    res, pairs, totals = engine.align_batch(batch, params=default_params(diagonalExpansion=e, minDiagsBetweenTraceBack=mind),
                                            want_totals=True)
    assert res[0]["status"] == 0
    got = item_pairs(res, pairs, 0)
    stats = parity.compare_pairs(got, syn_golden[tag + "_pairs"])
    worst_total = parity.compare_totals(totals[0], syn_golden[tag + "_totals"])
    print(tag, stats, worst_total, res[0]["n_tracebacks"])


def test_batch_mixed_widths_vs_oracle(engine, template_tables):
    """A batch mixing band widths (several kernel instantiations in one call), checked against the oracle."""
    import oracleshim as O
    from cpecan_signal import default_params, synth
    from cpecan_signal.engine import item_pairs
    l1, l2, l3 = template_tables
    reads = [synth.make_read(l1, 100 + i, lX=lx, anchor_every=ev)
             for i, (lx, ev) in enumerate([(400, 50), (1300, 50), (350, 400), (800, 50), (250, 50), (600, 300)])]
    e = 40
    batch = _three_state_batch(engine, template_tables, [r.ref for r in reads], [r.events for r in reads],
                               [r.anchors for r in reads], [r.scale5 for r in reads], [(1, 1)] * len(reads))
    res, pairs, totals = engine.align_batch(batch, params=default_params(diagonalExpansion=e), want_totals=True)
    op = O.default_params(diagonalExpansion=e)
    for i, r in enumerate(reads):
        m = O.Model(O.THREE_STATE, tables=(l1, l2, l3), scale5=r.scale5)
        want, wtot = O.align_banded(m, r.ref, r.events, r.anchors, params=op, ragged=(1, 1), want_totals=True)
        assert res[i]["status"] == 0
        assert res[i]["band_cells"] == O.band_cells(r.anchors, r.lX, r.lY, op, (1, 1))
        stats = parity.compare_pairs(item_pairs(res, pairs, i), want)
        parity.compare_totals(totals[i], wtot)
        print(i, stats)


def test_edge_shapes_vs_oracle(engine, template_tables):
    """Degenerate shapes in one batch: no anchors at all (the band is the whole matrix), a single anchor, anchors in the
    corners, one k-mer against one event, more events than a band of the given expansion can follow, every ragged
    combination."""
    import oracleshim as O
    from cpecan_signal import default_params, synth
    from cpecan_signal.engine import item_pairs
    l1, l2, l3 = template_tables
    base = [synth.make_read(l1, 600 + i, lX=lx) for i, lx in enumerate([120, 200, 90, 150, 60, 260])]
    refs = [r.ref for r in base]
    events = [r.events for r in base]
    anchors = [np.zeros((0, 2), np.int64),                       # none
               base[1].anchors[:1],                              # one
               np.array([[0, 0], [base[2].lX - 1, base[2].lY - 1]], dtype=np.int64),   # the corners
               base[3].anchors,
               base[4].anchors,
               base[5].anchors[::2]]
    refs.append(base[0].ref[:6]); events.append(base[0].events[:1]); anchors.append(np.zeros((0, 2), np.int64))   # 1 x 1
    refs.append(base[1].ref[:40]); events.append(base[1].events[:150]); anchors.append(np.zeros((0, 2), np.int64))  # lY >> lX
    scales = [r.scale5 for r in base] + [base[0].scale5, base[1].scale5]
    ragged = [(0, 0), (1, 0), (0, 1), (1, 1), (0, 0), (1, 1), (0, 0), (1, 1)]
    e = 20
    batch = _three_state_batch(engine, template_tables, refs, events, anchors, scales, ragged)
    res, pairs, totals = engine.align_batch(batch, params=default_params(diagonalExpansion=e), want_totals=True)
    for i in range(len(refs)):
        m = O.Model(O.THREE_STATE, tables=(l1, l2, l3), scale5=scales[i])
        want, wtot = O.align_banded(m, refs[i], events[i], anchors[i], params=O.default_params(diagonalExpansion=e),
                                    ragged=ragged[i], want_totals=True)
        assert res[i]["status"] == 0, i
        stats = parity.compare_pairs(item_pairs(res, pairs, i), want)
        parity.compare_totals(totals[i], wtot)
        print(i, stats)


@pytest.mark.parametrize("machine", ["three", "vanilla"])
def test_random_small_problems_vs_oracle(engine, template_tables, machine):
    """Fuzz: batches of random small reads under parameter sets that force many tracebacks per read (traceback every
    5 / 20 / 60 diagonals, 2 / 8 / 40 overlap diagonals), tiny expansions, sparse or no anchors and every ragged
    combination, both machines -- each item against the oracle (pairs, scores, per-diagonal totals)."""
    import oracleshim as O
    from cpecan_signal import default_params, synth, vanilla_hmm
    from cpecan_signal.engine import item_pairs
    l1, l2, l3 = template_tables
    rng = np.random.default_rng(77 if machine == "three" else 78)
    smt = O.THREE_STATE if machine == "three" else O.VANILLA
    for mind, tbd, e, thr in [(5, 2, 4, 0.01), (20, 8, 10, 0.2), (60, 40, 20, 0.01), (20, 2, 2, 0.01)]:
        reads, anchors, ragged = [], [], []
        for _ in range(14):
            r = synth.make_read(l1, int(rng.integers(1, 1 << 30)), lX=int(rng.integers(8, 120)),
                                anchor_every=int(rng.integers(5, 60)), noise_dist="gauss" if machine == "three" else "wald")
            keep = rng.random(len(r.anchors)) < rng.choice([0.0, 0.3, 1.0])
            reads.append(r); anchors.append(r.anchors[keep]); ragged.append((int(rng.integers(0, 2)), int(rng.integers(0, 2))))
        mk = _three_state_batch if machine == "three" else _vanilla_batch
        batch = mk(engine, template_tables, [r.ref for r in reads], [r.events for r in reads], anchors,
                   [r.scale5 for r in reads], ragged)
        kw = dict(diagonalExpansion=e, minDiagsBetweenTraceBack=mind, traceBackDiagonals=tbd, threshold=thr)
        res, pairs, totals = engine.align_batch(batch, hmm=None if machine == "three" else vanilla_hmm("template"),
                                                params=default_params(**kw), want_totals=True)
        worst = 0
        for i, r in enumerate(reads):
            m = O.Model(smt, tables=(l1, l2, l3), scale5=r.scale5, strand=0)
            want, wtot = O.align_banded(m, r.ref, r.events, anchors[i], params=O.default_params(**kw), ragged=ragged[i],
                                        want_totals=True)
            assert res[i]["status"] == 0, (i, kw)
            worst = max(worst, parity.compare_pairs(item_pairs(res, pairs, i), want, threshold=thr)["worst_score_diff"])
            parity.compare_totals(totals[i], wtot)
        # the E-step over the same batch: batch sums against the oracle's (transition / k-mer-skip / skip-bin counts)
        got, eres = engine.expectations_batch(batch, hmm=None if machine == "three" else vanilla_hmm("template"),
                                              params=default_params(**kw))
        want = sum(O.expectations(O.Model(smt, tables=(l1, l2, l3), scale5=r.scale5, strand=0), r.ref, r.events, anchors[i],
                                  params=O.default_params(**kw), ragged=ragged[i], pseudocount=0.0)
                   for i, r in enumerate(reads))
        assert (eres["status"] == 0).all()
        np.testing.assert_allclose(got[:-1], want[:-1], rtol=EXP_RTOL, atol=2e-4)
        assert abs(got[-1] - want[-1]) <= 1e-4 * abs(want[-1])
        print(machine, kw, "worst score diff", worst)


def test_c3_shape_batch_properties(engine, template_tables):
    """BASELINE config 3's shape (lX ~ 6700, ~8000 events, expansions 64 / 128 / 256, ~14 tracebacks per read) at a
    batch of several hundred reads, through properties that do not need the oracle -- plus the oracle on one read per
    expansion.  Every read appears twice, in different places of the batch: the copies must agree bit for bit, and a
    second run of the whole batch must reproduce the first (the result does not depend on which resident warp, which
    shared-memory ring size bucket or which neighbours a read gets).  The posteriors of one event over all reference
    positions sum to at most 1 (+ tolerance).  (The same over the events of one reference position does NOT hold in the
    reference either: an intermediate traceback is seeded with the end vector on its last diagonal, and the 41
    overlap diagonals do not always absorb that, so a position can be matched with probability ~1 on two diagonals
    45 apart -- 277 positions of the first read here.)"""
    import oracleshim as O
    from cpecan_signal import default_params, synth
    from cpecan_signal.engine import item_pairs
    l1, l2, l3 = template_tables
    uniq = [synth.make_read(l1, 4000 + i, lX=6700) for i in range(48)]
    for e in (64, 128, 256):
        reads = (uniq + uniq[::-1]) * 2                                  # 192 items, every read 4 times
        batch = _three_state_batch(engine, template_tables, [r.ref for r in reads], [r.events for r in reads],
                                   [r.anchors for r in reads], [r.scale5 for r in reads], [(1, 1)] * len(reads))
        prm = default_params(diagonalExpansion=e)
        res, pairs, _ = engine.align_batch(batch, params=prm)
        res2, pairs2, _ = engine.align_batch(batch, params=prm)
        assert (res["status"] == 0).all()
        assert np.array_equal(res["n_pairs"], res2["n_pairs"]) and np.array_equal(pairs, pairs2)
        assert np.array_equal(res["total_logprob"], res2["total_logprob"])
        n = len(uniq)
        for i in range(n):
            a = item_pairs(res, pairs, i)
            for j in (2 * n - 1 - i, 2 * n + i, 4 * n - 1 - i):
                assert np.array_equal(a, item_pairs(res, pairs, j)), "copies of read %d differ (items %d, %d)" % (i, i, j)
                assert res["total_logprob"][i] == res["total_logprob"][j]
            r = uniq[i]
            assert res["n_tracebacks"][i] >= 10 and res["band_cells"][i] > 6700 * e // 2
            assert a[:, 0].min() >= 100000 and a[:, 0].max() <= 10000000
            assert a[:, 1].min() >= 0 and a[:, 1].max() < r.lX and a[:, 2].min() >= 0 and a[:, 2].max() < r.lY
            key = a[:, 1].astype(np.int64) * (r.lY + 1) + a[:, 2]
            assert len(np.unique(key)) == len(key)                       # one posterior per (x, y)
            # traceback order: inside one traceback diagonals descending and x ascending inside a diagonal; the
            # tracebacks follow each other towards the end of the matrix
            dgl = (a[:, 1] + a[:, 2]).astype(np.int64)
            dd = np.diff(dgl)
            ups = np.flatnonzero(dd > 0)
            assert len(ups) <= res["n_tracebacks"][i] - 1
            assert (np.diff(a[:, 1])[dd == 0] > 0).all()
            seg_max = np.maximum.accumulate(dgl)
            assert (dgl[ups + 1] > seg_max[ups]).all()
            py = np.bincount(a[:, 2], weights=a[:, 0] / 1e7, minlength=r.lY)
            assert py.max() <= 1.0 + 1e-2
            assert np.isfinite(res["total_logprob"][i]) and -80000 < res["total_logprob"][i] < -5000
        k = {64: 0, 128: 1, 256: 2}[e]
        r = uniq[k]
        m = O.Model(O.THREE_STATE, tables=(l1, l2, l3), scale5=r.scale5)
        want, _ = O.align_banded(m, r.ref, r.events, r.anchors, params=O.default_params(diagonalExpansion=e), ragged=(1, 1))
        print(e, parity.compare_pairs(item_pairs(res, pairs, k), want))


# ------------------------------------------------------------------------------------------ E-step (expectations)
EXP_RTOL = 2e-4      # FP32 device path vs the reference's doubles; sums of ~1e5 terms


def _check_expectations(got, want):
    assert got.shape == want.shape == (9 + 4096 + 1,)
    np.testing.assert_allclose(got[:9], want[:9], rtol=EXP_RTOL, atol=2e-6)
    # a k-mer's skip count is a sum of a few posteriors, each within the 1e-4 posterior tolerance of north_star
    np.testing.assert_allclose(got[9:-1], want[9:-1], rtol=EXP_RTOL, atol=1e-4)
    assert abs(got[-1] - want[-1]) <= 1e-4 * abs(want[-1])


@pytest.mark.parametrize("cfg,e,ragged", [("e20_r00", 20, (0, 0)), ("e50_r11", 50, (1, 1))])
def test_fixture_expectations(engine, zymo, template_tables, cfg, e, ragged):
    """getExpectationsUsingAnchors on the reference's fixture read (golden from the unmodified reference, pseudocount
    1e-4 as vanillaAlign.c:675-676)."""
    from cpecan_signal import default_params
    rd = zymo["read"]
    batch = _three_state_batch(engine, template_tables, [zymo["ref"]], [rd["template_events"]],
                               [zymo["anchors_template"]], [rd["template_params"]], [ragged])
    got, res = engine.expectations_batch(batch, params=default_params(diagonalExpansion=e), pseudocount=1e-4)
    assert res[0]["status"] == 0
    _check_expectations(got, zymo["three_expectations_" + cfg])


def test_synthetic_expectations_batch_sum(engine, syn_golden, template_tables):
    """A batch's expectations are the sum of its reads' (the device accumulates over the batch)."""
    from cpecan_signal import default_params, synth
    tags = ["s0", "s1", "s3"]                       # same expansion groups are not required: one call per tag, then a joint call
    want_sum = None
    for tag in tags:
        idx, lX, e, r0, r1, every, mind = (int(v) for v in syn_golden[tag + "_meta"])
        r = synth.make_read(template_tables[0], idx, lX=lX, anchor_every=every)
        batch = _three_state_batch(engine, template_tables, [r.ref], [r.events], [r.anchors], [r.scale5], [(r0, r1)])
        got, res = engine.expectations_batch(batch, params=default_params(diagonalExpansion=e, minDiagsBetweenTraceBack=mind),
                                             pseudocount=1e-4)
        assert res[0]["status"] == 0
        _check_expectations(got, syn_golden[tag + "_expect"])
    # two reads with identical parameters in ONE call
    import oracleshim as O
    l1, l2, l3 = template_tables
    reads = [synth.make_read(l1, 300 + i, lX=500) for i in range(3)]
    batch = _three_state_batch(engine, template_tables, [r.ref for r in reads], [r.events for r in reads],
                               [r.anchors for r in reads], [r.scale5 for r in reads], [(1, 1)] * 3)
    got, res = engine.expectations_batch(batch, params=default_params(diagonalExpansion=30), pseudocount=0.0)
    want = np.zeros(9 + 4096 + 1)
    for r in reads:
        m = O.Model(O.THREE_STATE, tables=(l1, l2, l3), scale5=r.scale5)
        want += O.expectations(m, r.ref, r.events, r.anchors, params=O.default_params(diagonalExpansion=30), ragged=(1, 1),
                               pseudocount=0.0)
    _check_expectations(got, want)


def test_em_iterations_vs_oracle(engine, template_tables, tmp_path):
    """Two Baum-Welch iterations on the GPU (E-step on device, M-step + .hmm round trip on the host) against the same
    loop with the oracle as the E-step: trained transitions, k-mer skip probabilities and likelihoods agree."""
    import oracleshim as O
    from cpecan_signal import HostBatch, default_params, em, synth, three_state_hmm
    l1, l2, l3 = template_tables
    reads = [synth.make_read(l1, 800 + i, lX=300 + 50 * (i % 3)) for i in range(6)]
    e = 30
    mid = engine.upload_model(l1, l3, np.full(4096, -2.3025850929940455))
    batch = HostBatch([r.ref for r in reads], [r.events for r in reads], [r.anchors for r in reads],
                      model_ids=[mid] * len(reads), scales=[r.scale5 for r in reads], ragged=[(1, 1)] * len(reads))
    gm, om = em.ContinuousPairHmm(), em.ContinuousPairHmm()
    g_trans = o_trans = None
    g_gapx = o_gapx = None
    for it in range(2):
        hmm = three_state_hmm(g_trans)
        if g_gapx is not None:
            engine.update_model(mid, gapx=g_gapx)
        gvec = em.gpu_estep(engine, batch, hmm, default_params(diagonalExpansion=e), distributed=False)
        ovec = np.zeros(em.N_EXPECT)
        for r in reads:
            m = O.Model(O.THREE_STATE, tables=(l1, l2, l3), scale5=r.scale5, transitions=o_trans, gap_x=o_gapx)
            ovec += O.expectations(m, r.ref, r.events, r.anchors, params=O.default_params(diagonalExpansion=e),
                                   ragged=(1, 1), pseudocount=0.0)
        _check_expectations(gvec, ovec)
        gl = em.em_iteration(gm, gvec, len(reads), str(tmp_path / "g.hmm"))
        ol = em.em_iteration(om, ovec, len(reads), str(tmp_path / "o.hmm"))
        np.testing.assert_allclose(gl.transitions, ol.transitions, atol=3e-6)
        np.testing.assert_allclose(gl.kmer_skip_probs, ol.kmer_skip_probs, atol=3e-6)
        g_trans, g_gapx = gl.state_machine_params()
        o_trans, o_gapx = ol.state_machine_params()
    assert abs(gm.running_likelihoods[-1] - om.running_likelihoods[-1]) <= 1e-4 * abs(om.running_likelihoods[-1])


# ------------------------------------------------------------------------------------------ vanilla state machine
def _vanilla_batch(eng, tables, refs, events, anchors, scales, ragged):
    from cpecan_signal import HostBatch, vanilla_gapx
    l1, l2, l3 = tables
    mid = eng.upload_model(l1, l3, vanilla_gapx(l2))
    return HostBatch(refs, events, anchors, model_ids=[mid] * len(refs), scales=scales, ragged=ragged)


@pytest.mark.parametrize("tag,e,ragged,count", [("vanilla_e20_r00", 20, (0, 0), 999), ("vanilla_e50_r11", 50, (1, 1), None)])
def test_fixture_banded_vanilla(engine, zymo, template_tables, tag, e, ragged, count):
    """stateMachine3Vanilla on the fixture read: 999 aligned pairs banded (tests/signalPairwiseTest.c:1250-1310)."""
    from cpecan_signal import default_params, vanilla_hmm
    from cpecan_signal.engine import item_pairs
    rd = zymo["read"]
    batch = _vanilla_batch(engine, template_tables, [zymo["ref"]], [rd["template_events"]],
                           [zymo["anchors_template"]], [rd["template_params"]], [ragged])
    res, pairs, totals = engine.align_batch(batch, hmm=vanilla_hmm("template"), params=default_params(diagonalExpansion=e),
                                            want_totals=True)
    assert res[0]["status"] == 0
    got = item_pairs(res, pairs, 0)
    stats = parity.compare_pairs(got, zymo[tag + "_pairs"])
    worst_total = parity.compare_totals(totals[0], zymo[tag + "_totals"])
    print(tag, stats, worst_total)
    if count is not None:
        assert stats["n_got"] == stats["n_want"] == count


def test_fixture_unbanded_vanilla(engine, zymo, template_tables):
    """getAlignedPairsWithoutBanding, vanilla: 953 pairs (tests/signalPairwiseTest.c:1296-1303)."""
    from cpecan_signal import vanilla_hmm
    from cpecan_signal.engine import MODE_UNBANDED, item_pairs
    rd = zymo["read"]
    batch = _vanilla_batch(engine, template_tables, [zymo["ref"]], [rd["template_events"]], [np.zeros((0, 2))],
                           [rd["template_params"]], [(0, 0)])
    res, pairs, _ = engine.align_batch(batch, hmm=vanilla_hmm("template"), mode=MODE_UNBANDED)
    assert res[0]["status"] == 0
    stats = parity.compare_pairs(item_pairs(res, pairs, 0), zymo["vanilla_unbanded_pairs"])
    assert abs(res[0]["total_logprob"] - float(zymo["vanilla_unbanded_total"])) <= 1e-4 * abs(float(zymo["vanilla_unbanded_total"]))
    assert stats["n_got"] == stats["n_want"] == 953


def test_tiny_known_answer_vanilla(engine, zymo, template_tables):
    """tests/signalPairwiseTest.c:795-897: exactly 5 pairs {(2,0),(3,3),(5,4),(6,5),(7,6)} at threshold 0.5."""
    from cpecan_signal import default_params, vanilla_hmm
    from cpecan_signal.engine import MODE_UNBANDED, item_pairs
    ev = np.array([58.743435, 0.887833, 0.0571, 53.604965, 0.816836, 0.0571, 58.432015, 0.735143, 0.0571,
                   63.684352, 0.795437, 0.0571, 58.921430, 0.812959, 0.0571, 59.895882, 0.740952, 0.0571,
                   61.684303, 0.722332, 0.0571]).reshape(-1, 3)
    batch = _vanilla_batch(engine, template_tables, ["ACGATACGGACAT"], [ev], [np.zeros((0, 2))], None, [(0, 0)])
    res, pairs, _ = engine.align_batch(batch, hmm=vanilla_hmm(None), params=default_params(threshold=0.5), mode=MODE_UNBANDED)
    got = item_pairs(res, pairs, 0)
    assert {(int(x), int(y)) for _, x, y in got} == {(2, 0), (3, 3), (5, 4), (6, 5), (7, 6)}
    parity.compare_pairs(got, zymo["tiny_vanilla_pairs"], threshold=0.5)


@pytest.mark.parametrize("cfg,e,ragged", [("e20_r00", 20, (0, 0)), ("e50_r11", 50, (1, 1))])
def test_fixture_expectations_vanilla(engine, zymo, template_tables, cfg, e, ragged):
    """Vanilla E-step: 30 beta + 30 alpha skip-bin counts and the likelihood (impl/pairwiseAligner.c:478-498)."""
    from cpecan_signal import default_params, vanilla_hmm
    rd = zymo["read"]
    batch = _vanilla_batch(engine, template_tables, [zymo["ref"]], [rd["template_events"]],
                           [zymo["anchors_template"]], [rd["template_params"]], [ragged])
    got, res = engine.expectations_batch(batch, hmm=vanilla_hmm("template"), params=default_params(diagonalExpansion=e),
                                         pseudocount=1e-4)
    want = zymo["vanilla_expectations_" + cfg]
    assert res[0]["status"] == 0 and got.shape == want.shape == (61,)
    np.testing.assert_allclose(got[:60], want[:60], rtol=EXP_RTOL, atol=1e-4)
    assert abs(got[-1] - want[-1]) <= 1e-4 * abs(want[-1])


def test_vanilla_synthetic_vs_oracle(engine, template_tables):
    """Seeded synthetic reads with inverse-Gaussian noise through the vanilla machine, batch of mixed sizes."""
    import oracleshim as O
    from cpecan_signal import default_params, synth, vanilla_hmm
    from cpecan_signal.engine import item_pairs
    l1, l2, l3 = template_tables
    reads = [synth.make_read(l1, 900 + i, lX=lx, noise_dist="wald") for i, lx in enumerate([300, 1200, 500])]
    e = 40
    batch = _vanilla_batch(engine, template_tables, [r.ref for r in reads], [r.events for r in reads],
                           [r.anchors for r in reads], [r.scale5 for r in reads], [(1, 1)] * len(reads))
    res, pairs, totals = engine.align_batch(batch, hmm=vanilla_hmm("complement"), params=default_params(diagonalExpansion=e),
                                            want_totals=True)
    for i, r in enumerate(reads):
        m = O.Model(O.VANILLA, tables=(l1, l2, l3), scale5=r.scale5, strand=1)
        want, wtot = O.align_banded(m, r.ref, r.events, r.anchors, params=O.default_params(diagonalExpansion=e),
                                    ragged=(1, 1), want_totals=True)
        assert res[i]["status"] == 0
        print(i, parity.compare_pairs(item_pairs(res, pairs, i), want), parity.compare_totals(totals[i], wtot))


def test_em_iterations_vanilla_vs_oracle(engine, template_tables, tmp_path):
    """Two training iterations of the vanilla machine (trainModels.py with stateMachineType vanilla): E-step on the GPU
    (60 skip-bin sums + likelihood), ConditionalSignalHmm M-step with alpha / beta normalised separately, trained file
    round trip, the new bins uploaded as EMISSION_GAP_X_PROBS -- against the same loop with the oracle as the E-step."""
    import oracleshim as O
    from cpecan_signal import default_params, em, synth, vanilla_gapx, vanilla_hmm
    l1, l2, l3 = template_tables
    reads = [synth.make_read(l1, 950 + i, lX=350 + 60 * (i % 3), noise_dist="wald") for i in range(5)]
    e = 30
    batch = _vanilla_batch(engine, template_tables, [r.ref for r in reads], [r.events for r in reads],
                           [r.anchors for r in reads], [r.scale5 for r in reads], [(1, 1)] * len(reads))
    mid = int(batch.model_id[0])
    gm = em.ConditionalSignalHmm(match_model=l1, scaled_match_model=l3)
    om = em.ConditionalSignalHmm(match_model=l1, scaled_match_model=l3)
    g_bins = o_bins = None
    for it in range(2):
        if g_bins is not None:
            engine.update_model(mid, gapx=g_bins)
        gvec = em.gpu_estep(engine, batch, vanilla_hmm("template"), default_params(diagonalExpansion=e), distributed=False)
        assert gvec.shape == (em.N_EXPECT_VANILLA,)
        ovec = np.zeros(em.N_EXPECT_VANILLA)
        for r in reads:
            m = O.Model(O.VANILLA, tables=(l1, l2, l3), scale5=r.scale5, strand=0, gap_x=o_bins)
            ovec += O.expectations(m, r.ref, r.events, r.anchors, params=O.default_params(diagonalExpansion=e),
                                   ragged=(1, 1), pseudocount=0.0)
        np.testing.assert_allclose(gvec[:60], ovec[:60], rtol=EXP_RTOL, atol=1e-4)
        assert abs(gvec[-1] - ovec[-1]) <= 1e-4 * abs(ovec[-1])
        gl = em.em_iteration(gm, gvec, len(reads), str(tmp_path / "gv.hmm"))
        ol = em.em_iteration(om, ovec, len(reads), str(tmp_path / "ov.hmm"))
        assert abs(gl.kmer_skip_bins[:30].sum() - 1) < 1e-9 and abs(gl.kmer_skip_bins[30:].sum() - 1) < 1e-9
        np.testing.assert_allclose(gl.kmer_skip_bins, ol.kmer_skip_bins, atol=2e-5)
        g_bins, o_bins = gl.state_machine_params()[1], ol.state_machine_params()[1]
    assert abs(gm.running_likelihoods[-1] - om.running_likelihoods[-1]) <= 1e-4 * abs(om.running_likelihoods[-1])


def test_split_regions_as_items(engine, syn_golden, template_tables):
    """An anchor gap above splitMatrixBiggerThanThis: the regions of getSplitPoints become work items with ragged
    ends on the cut sides, and their pairs, shifted back and reversed per region, are what
    getAlignedPairsUsingAnchors returns (impl/pairwiseAligner.c:1356-1422, 1447-1454)."""
    import oracleshim as O
    from cpecan_signal import default_params, synth
    from cpecan_signal.engine import item_pairs
    idx, lX, e, r0, r1, every, mind, split = (int(v) for v in syn_golden["split_meta"])
    lo, hi = (int(v) for v in syn_golden["split_keep_lo_hi"])
    r = synth.make_read(template_tables[0], idx, lX=lX, anchor_every=every)
    anchors = r.anchors[(r.anchors[:, 0] < lo) | (r.anchors[:, 0] > hi)]
    regions = O.split_points(anchors, r.lX, r.lY, split, r0, r1)
    assert len(regions) == 2
    refs, evs, ans, rag = [], [], [], []
    j = 0
    for i, (x1, y1, x2, y2) in enumerate(regions):
        refs.append(r.ref[x1:x2 + 5])
        evs.append(r.events[y1:y2])
        sub = []
        while j < len(anchors) and anchors[j, 0] + anchors[j, 1] < x2 + y2:
            sub.append((anchors[j, 0] - x1, anchors[j, 1] - y1)); j += 1
        ans.append(np.array(sub, dtype=np.int64).reshape(-1, 2))
        rag.append((r0 or i > 0, r1 or i < len(regions) - 1))
    batch = _three_state_batch(engine, template_tables, refs, evs, ans, [r.scale5] * len(regions), rag)
    res, pairs, _ = engine.align_batch(batch, params=default_params(diagonalExpansion=e))
    got = []
    for i, (x1, y1, x2, y2) in enumerate(regions):
        assert res[i]["status"] == 0
        p = parity.reverse_regions(item_pairs(res, pairs, i)).astype(np.int64)
        p[:, 1] += x1; p[:, 2] += y1
        got.append(p)
    got = np.concatenate(got)
    stats = parity.compare_pairs(got, syn_golden["split_pairs"])
    print(stats)
    if stats["n_got"] == stats["n_want"]:
        assert np.array_equal(got[:, 1:], syn_golden["split_pairs"][:, 1:])


def test_argument_errors_are_reported(engine, template_tables):
    """Integer status + message, never an abort: banding parameters the reference rejects (traceBackDiagonals + 1 >=
    minDiagsBetweenTraceBack, impl/pairwiseAligner.c:880-884), a model with too few gap-X entries for the machine, an unknown
    state-machine type.  (An odd expansion is legal: it runs on the FP64 kernel, INTEGRATION.md.)"""
    from cpecan_signal import EngineError, HostBatch, default_params, synth, vanilla_hmm
    from cpecan_signal.engine import Hmm
    l1, l2, l3 = template_tables
    r = synth.make_read(l1, 7, lX=200)
    mid = engine.upload_model(l1, l3, np.full(4096, -2.3025850929940455))
    hb = HostBatch([r.ref], [r.events], [r.anchors], model_ids=[mid], scales=[r.scale5], ragged=[(1, 1)])
    with pytest.raises(EngineError, match="banding parameters"):
        engine.align_batch(hb, params=default_params(traceBackDiagonals=40, minDiagsBetweenTraceBack=41))
    engine.expectations_batch(hb, params=default_params(diagonalExpansion=21))     # odd expansions run on the FP64 kernel
    mid60 = engine.upload_model(l1, l3, np.full(60, 0.1))
    hb60 = HostBatch([r.ref], [r.events], [r.anchors], model_ids=[mid60], scales=[r.scale5], ragged=[(1, 1)])
    with pytest.raises(EngineError, match="gap-X"):
        engine.align_batch(hb60)                                   # three-state needs 4096 entries
    engine.align_batch(hb60, hmm=vanilla_hmm("template"))          # vanilla is fine with 60
    bad = Hmm()
    bad.sm_type = 3                                                # threeStateAsymmetric: not a signal machine
    with pytest.raises(EngineError, match="not implemented"):
        engine.align_batch(hb, hmm=bad)
    # a band wider than the widest shared-memory ring (a long read without anchors): refused, not mis-computed
    big = synth.make_read(l1, 8, lX=6000)
    hbw = HostBatch([big.ref], [big.events], [np.zeros((0, 2), np.int64)], model_ids=[mid], scales=[big.scale5], ragged=[(1, 1)])
    with pytest.raises(EngineError, match="band wider"):
        engine.align_batch(hbw)
    # empty batch and an item without events are legal
    res, pairs, _ = engine.align_batch(HostBatch([], [], []))
    assert len(res) == 0
    hb0 = HostBatch([r.ref], [np.zeros((0, 3))], [np.zeros((0, 2))], model_ids=[mid], scales=[r.scale5], ragged=[(1, 1)])
    res, pairs, _ = engine.align_batch(hb0)
    assert res[0]["status"] == 0 and res[0]["n_pairs"] == 0


@pytest.mark.parametrize("machine", ["three", "vanilla"])
def test_odd_expansions_vs_oracle(engine, template_tables, machine):
    """The reference takes odd diagonalExpansion values (its band edges then move backwards and by two cells between
    diagonals, impl/pairwiseAligner.c:98-170).  Such batches run on the FP64 kernel with explicit band edges: random
    reads, odd expansions, several traceback parameter sets, every ragged combination -- same pair lists as the oracle,
    scores to the last digit, totals to 1e-9, band cells exact."""
    import oracleshim as O
    from cpecan_signal import HostBatch, default_params, synth, three_state_hmm, vanilla_gapx, vanilla_hmm
    from cpecan_signal.engine import item_pairs
    l1, l2, l3 = template_tables
    rng = np.random.default_rng(131 if machine == "three" else 132)
    van = machine == "vanilla"
    hmm = vanilla_hmm("template") if van else three_state_hmm()
    mid = engine.upload_model(l1, l3, vanilla_gapx(l2) if van else np.full(4096, -2.3025850929940455))
    for mind, tbd, e, thr, lxs in [(5, 2, 3, 0.01, (8, 150)), (20, 8, 11, 0.2, (8, 150)), (60, 40, 21, 0.05, (20, 300)),
                                   (1000, 40, 51, 0.01, (700, 1500)), (1000, 40, 1, 0.01, (8, 300))]:
        reads, anchors, ragged = [], [], []
        for _ in range(8):
            r = synth.make_read(l1, int(rng.integers(1, 1 << 30)), lX=int(rng.integers(*lxs)),
                                anchor_every=int(rng.integers(5, 60)), noise_dist="wald" if van else "gauss")
            keep = rng.random(len(r.anchors)) < rng.choice([0.3, 1.0])
            reads.append(r); anchors.append(r.anchors[keep]); ragged.append((int(rng.integers(0, 2)), int(rng.integers(0, 2))))
        batch = HostBatch([r.ref for r in reads], [r.events for r in reads], anchors, model_ids=[mid] * len(reads),
                          scales=[r.scale5 for r in reads], ragged=ragged)
        kw = dict(diagonalExpansion=e, minDiagsBetweenTraceBack=mind, traceBackDiagonals=tbd, threshold=thr)
        res, pairs, totals = engine.align_batch(batch, hmm=hmm, params=default_params(**kw), want_totals=True, pair_cap=400000)
        worst = 0
        for i, r in enumerate(reads):
            m = (O.Model(O.VANILLA, tables=(l1, l2, l3), scale5=r.scale5, strand=0) if van
                 else O.Model(O.THREE_STATE, tables=(l1, l2, l3), scale5=r.scale5))
            want, wtot = O.align_banded(m, r.ref, r.events, anchors[i], params=O.default_params(**kw), ragged=ragged[i],
                                        want_totals=True)
            assert res[i]["status"] == 0, (i, kw, res[i]["status"])
            lX, lY = len(r.ref) - 5, len(r.events)
            assert res[i]["band_cells"] == O.band_cells(anchors[i], lX, lY, O.default_params(**kw))
            got = np.asarray(parity.reverse_regions(item_pairs(res, pairs, i)), dtype=np.int64).reshape(-1, 3)
            want = np.asarray(want, dtype=np.int64).reshape(-1, 3)
            assert got.shape == want.shape and np.array_equal(got[:, 1:], want[:, 1:]), (i, kw)
            if len(got):
                worst = max(worst, int(np.abs(got[:, 0] - want[:, 0]).max()))
            mask = ~np.isnan(wtot)
            assert np.array_equal(mask, ~np.isnan(totals[i]))
            if mask.any():
                assert float(np.abs(totals[i][mask] - wtot[mask]).max()) <= 1e-9 * max(1.0, float(np.abs(wtot[mask]).max()))
        print(machine, kw, "worst score diff", worst)
        assert worst <= 2
    engine.release_model(mid)
