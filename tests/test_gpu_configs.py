"""GPU parity at the sizes BASELINE.json's configurations name, against the oracle run on the box's host cores:
config 3 (>= 256 seeded reads per band expansion) and config 2 (a 2D read: template strand with the
template model, complement strand against the reverse-complemented reference with complement_median68pA_pop2)."""
import json
import multiprocessing as mp
import os

import numpy as np
import pytest

import parity

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _oracle_one(args):
    """(machine, model path, strand, ref, events, anchors, scale5, e) -> (pairs, totals) through the oracle."""
    import oracleshim as O
    smt, model_path, strand, ref, events, anchors, scale5, e = args
    m = O.Model(smt, model_file=model_path, scale5=scale5, strand=strand)
    return O.align_banded(m, ref, events, anchors, params=O.default_params(diagonalExpansion=e), ragged=(1, 1),
                          want_totals=True)


def _pool_map(fn, jobs):
    procs = min(len(jobs), os.cpu_count() or 1)
    with mp.get_context("fork").Pool(procs) as pool:
        return pool.map(fn, jobs, chunksize=1)


@pytest.mark.parametrize("e", [64, 128, 256])
def test_c3_sample_vs_oracle(engine, template_tables, e):
    """BASELINE config 3: 256 seeded reads (lX = 6700, ~8000 events, anchors every 50 k-mers, ragged (1,1)) per band
    expansion -- every aligned pair, every score and the total of every traceback segment against the oracle."""
    import oracleshim as O
    from cpecan_signal import HostBatch, default_params, synth
    from cpecan_signal.engine import item_pairs
    n = int(os.environ.get("CPECAN_C3_SAMPLE", "256"))
    l1, l2, l3 = template_tables
    first = {64: 70000, 128: 71000, 256: 72000}[e]
    reads = _pool_map(_make_read, [(first + i, 6700) for i in range(n)])
    mid = engine.upload_model(l1, l3, np.full(4096, -2.3025850929940455))
    batch = HostBatch([r.ref for r in reads], [r.events for r in reads], [r.anchors for r in reads],
                      model_ids=[mid] * n, scales=[r.scale5 for r in reads], ragged=[(1, 1)] * n)
    res, pairs, totals = engine.align_batch(batch, params=default_params(diagonalExpansion=e), want_totals=True)
    engine.release_model(mid)
    want = _pool_map(_oracle_one, [(O.THREE_STATE, synth.TEMPLATE_MODEL, None, r.ref, r.events, r.anchors, r.scale5, e)
                                   for r in reads])
    worst_score, worst_total, n_pairs, boundary, flips = 0, 0.0, 0, 0, []
    for i in range(n):
        assert res[i]["status"] == 0
        # logAdd segment flips (parity.FLIP_SCORE_TOL): one event moves the pairs of a few neighbouring diagonals; at most
        # 4 pairs per read and (below) 8 per 256 reads = ~2.2 million pairs.  Measured: 3 / 0 / 2 pairs at e = 64 / 128 / 256
        st = parity.compare_pairs(item_pairs(res, pairs, i), want[i][0], max_flips=4)
        flips += [(first + i,) + f for f in st["flips"]]
        worst_score = max(worst_score, st["worst_score_diff"])
        worst_total = max(worst_total, parity.compare_totals(totals[i], want[i][1]))
        n_pairs += st["n_want"]
        boundary += st["n_want"] + st["n_got"] - 2 * st["common"]
    assert len(flips) <= max(2, 8 * n // 256), flips
    rec = dict(config="C3", expansion=e, reads=n, aligned_pairs=n_pairs, worst_score_diff=worst_score,
               score_tolerance=parity.SCORE_TOL, worst_abs_total_diff=worst_total,
               logadd_segment_flips=[dict(read=f[0], pair=list(f[1]), score=f[2], reference=f[3]) for f in flips],
               pairs_only_on_one_side_within_tolerance_of_threshold=boundary)
    print(json.dumps(rec))
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "parity_c3_e%d.json" % e), "w") as fh:
            json.dump(rec, fh)


def _make_read(args):
    from cpecan_signal import synth
    idx, lX = args
    return synth.make_read(synth.load_model_file(synth.TEMPLATE_MODEL)[0], idx, lX=lX)


def test_c2_fixture_complement_strand(engine, zymo):
    """Config 2 on the reference's own 2D fixture read: the complement events against the reverse-complemented
    reference with complement_median68pA_pop2.model, e = 50, ragged (1,1) (vanillaAlign.c:635, 772-774); golden from the
    unmodified reference (oracle/make_golden.py)."""
    from cpecan_signal import HostBatch, default_params, synth
    from cpecan_signal.engine import item_pairs
    rd = zymo["read"]
    c1, c2, c3 = synth.load_model_file(synth.COMPLEMENT_MODEL)
    mid = engine.upload_model(c1, c3, np.full(4096, -2.3025850929940455))
    batch = HostBatch([synth.reverse_complement(zymo["ref"])], [rd["complement_events"]], [zymo["anchors_complement"]],
                      model_ids=[mid], scales=[rd["complement_params"]], ragged=[(1, 1)])
    res, pairs, totals = engine.align_batch(batch, params=default_params(diagonalExpansion=50), want_totals=True)
    engine.release_model(mid)
    assert res[0]["status"] == 0
    want = zymo["three_e50_r11_complement_pairs"]
    got = item_pairs(res, pairs, 0)
    st = parity.compare_pairs(got, want)
    wt = parity.compare_totals(totals[0], zymo["three_e50_r11_complement_totals"])
    print(st, wt)
    assert st["n_got"] == st["n_want"] == len(want)
    assert np.array_equal(parity.reverse_regions(got)[:, 1:], want[:, 1:])


@pytest.mark.parametrize("machine", ["three", "vanilla"])
def test_c2_synthetic_2d_read(engine, machine):
    """Config 2 at its named size: a synthetic 2D read, ~5000 k-mers / ~6000 events per strand, e = 50 -- the template
    strand with the template model, the complement strand (events drawn from the complement model along the reverse
    complement of the same reference) with complement_median68pA_pop2, in ONE batch with two models; both machines."""
    import oracleshim as O
    from cpecan_signal import HostBatch, default_params, synth, vanilla_gapx, vanilla_hmm
    from cpecan_signal.engine import item_pairs
    t1, t2, t3 = synth.load_model_file(synth.TEMPLATE_MODEL)
    c1, c2, c3 = synth.load_model_file(synth.COMPLEMENT_MODEL)
    dist = "gauss" if machine == "three" else "wald"
    rt = synth.make_read(t1, 52000, lX=5000, noise_dist=dist)
    rc = synth.make_read(c1, 52001, ref=synth.reverse_complement(rt.ref), noise_dist=dist)
    e = 50
    if machine == "three":
        gx = np.full(4096, -2.3025850929940455)
        mt, mc = engine.upload_model(t1, t3, gx), engine.upload_model(c1, c3, gx)
        hmm, smt, strands = None, O.THREE_STATE, (None, None)
    else:
        mt, mc = engine.upload_model(t1, t3, vanilla_gapx(t2)), engine.upload_model(c1, c3, vanilla_gapx(c2))
        smt, strands = O.VANILLA, (0, 1)
    want = _pool_map(_oracle_one, [(smt, synth.TEMPLATE_MODEL, strands[0], rt.ref, rt.events, rt.anchors, rt.scale5, e),
                                   (smt, synth.COMPLEMENT_MODEL, strands[1], rc.ref, rc.events, rc.anchors, rc.scale5, e)])
    for k, (r, mid, strand) in enumerate([(rt, mt, "template"), (rc, mc, "complement")]):
        batch = HostBatch([r.ref], [r.events], [r.anchors], model_ids=[mid], scales=[r.scale5], ragged=[(1, 1)])
        res, pairs, totals = engine.align_batch(batch, hmm=None if machine == "three" else vanilla_hmm(strand),
                                                params=default_params(diagonalExpansion=e), want_totals=True)
        assert res[0]["status"] == 0 and res[0]["n_tracebacks"] >= 8
        st = parity.compare_pairs(item_pairs(res, pairs, 0), want[k][0])
        wt = parity.compare_totals(totals[0], want[k][1])
        print(machine, strand, st, wt)
        assert st["n_want"] > 3000
    engine.release_model(mt); engine.release_model(mc)


@pytest.mark.parametrize("machine,e,n_reads", [("three", 64, 16), ("vanilla", 64, 16), ("three", 256, 6), ("vanilla", 128, 6), ("three", 65, 6)])
def test_config3_exact_arithmetic(engine, template_tables, machine, e, n_reads):
    """cpecan_cuda_set_exact_arithmetic: config 3 reads (lX = 6700; e = 64 / 128 / 256, and an odd one) on the FP64 kernel --
    the oracle's pair lists exactly (no flip budget), scores to the last digit."""
    import oracleshim as O
    from cpecan_signal import HostBatch, default_params, synth, three_state_hmm, vanilla_gapx, vanilla_hmm
    from cpecan_signal.engine import item_pairs
    l1, l2, l3 = template_tables
    van = machine == "vanilla"
    reads = [synth.make_read(l1, (7_100_000 if van else 7_000_000) + i, lX=6700, noise_dist="wald" if van else "gauss")
             for i in range(n_reads)]
    mid = engine.upload_model(l1, l3, vanilla_gapx(l2) if van else np.full(4096, -2.3025850929940455))
    batch = HostBatch([r.ref for r in reads], [r.events for r in reads], [r.anchors for r in reads],
                      model_ids=[mid] * len(reads), scales=[r.scale5 for r in reads], ragged=[(1, 1)] * len(reads))
    smt = O.VANILLA if van else O.THREE_STATE
    want = _pool_map(_oracle_one, [(smt, synth.TEMPLATE_MODEL, 0 if van else None, r.ref, r.events, r.anchors, r.scale5, e)
                                   for r in reads])
    engine.set_exact_arithmetic(True)
    try:
        res, pairs, _ = engine.align_batch(batch, hmm=vanilla_hmm("template") if van else three_state_hmm(),
                                           params=default_params(diagonalExpansion=e), pair_cap=engine.default_pair_capacity(batch, 3))
    finally:
        engine.set_exact_arithmetic(False)
        engine.release_model(mid)
    worst = 0
    for i in range(len(reads)):
        assert res[i]["status"] == 0
        got = np.asarray(parity.reverse_regions(item_pairs(res, pairs, i)), dtype=np.int64).reshape(-1, 3)
        w = np.asarray(want[i][0], dtype=np.int64).reshape(-1, 3)
        assert got.shape == w.shape and np.array_equal(got[:, 1:], w[:, 1:]), i
        worst = max(worst, int(np.abs(got[:, 0] - w[:, 0]).max()))
    print(machine, "e =", e, "exact arithmetic: worst score difference", worst, "of 1e7")
    assert worst <= 2
