"""Developer diagnostic (not a test): prints GPU-vs-oracle deviations per case.  Run on a GPU box:
   python tests/dev_gpu_diag.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cpecan-signal_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracleshim as O  # noqa: E402
from cpecan_signal import Engine, HostBatch, default_params, synth  # noqa: E402
from cpecan_signal.engine import MODE_UNBANDED, item_pairs  # noqa: E402


def report(tag, got, want, gtot, wtot):
    g = {(int(x), int(y)): int(s) for s, x, y in got}
    w = {(int(x), int(y)): int(s) for s, x, y in want}
    common = set(g) & set(w)
    diffs = sorted(((abs(g[k] - w[k]), k, g[k], w[k]) for k in common), reverse=True)
    only_g = sorted((g[k], k) for k in set(g) - set(w))[-3:]
    only_w = sorted((w[k], k) for k in set(w) - set(g))[-3:]
    line = "%-28s n=%d/%d worst=%s only_gpu=%s only_ref=%s" % (tag, len(g), len(w), diffs[:3], only_g, only_w)
    if gtot is not None and wtot is not None:
        mask = ~np.isnan(wtot)
        same = np.array_equal(mask, ~np.isnan(gtot))
        dd = np.abs(gtot[mask & ~np.isnan(gtot)] - wtot[mask & ~np.isnan(gtot)])
        line += " totals: mask_same=%s max|d|=%.3g at %s" % (same, dd.max() if dd.size else 0, int(np.argmax(dd)) if dd.size else -1)
    n_bad = sum(1 for d in diffs if d[0] > 1002)
    print(line, "n_bad=%d" % n_bad, flush=True)


def main():
    eng = Engine(0)
    l1, l2, l3 = synth.load_model_file(synth.TEMPLATE_MODEL)
    mid = eng.upload_model(l1, l3, np.full(4096, -2.3025850929940455))
    cases = []
    for i, (lx, every, e, rag, mind) in enumerate([(400, 50, 40, (1, 1), 1000), (1300, 50, 40, (1, 1), 1000),
                                                   (350, 400, 40, (1, 1), 1000), (900, 50, 40, (1, 1), 1000),
                                                   (3000, 50, 64, (1, 1), 1000), (6700, 50, 64, (1, 1), 1000),
                                                   (6700, 50, 128, (1, 1), 1000), (700, 50, 20, (1, 1), 200)]):
        r = synth.make_read(l1, 100 + i if i != 3 else 1, lX=lx, anchor_every=every)
        cases.append((r, e, rag, mind))
    for ci, (r, e, rag, mind) in enumerate(cases):
        batch = HostBatch([r.ref], [r.events], [r.anchors], model_ids=[mid], scales=[r.scale5], ragged=[rag])
        res, pairs, totals = eng.align_batch(batch, params=default_params(diagonalExpansion=e, minDiagsBetweenTraceBack=mind),
                                             want_totals=True)
        m = O.Model(O.THREE_STATE, tables=(l1, l2, l3), scale5=r.scale5)
        want, wtot = O.align_banded(m, r.ref, r.events, r.anchors,
                                    params=O.default_params(diagonalExpansion=e, minDiagsBetweenTraceBack=mind),
                                    ragged=rag, want_totals=True)
        report("case%d lX=%d e=%d st=%d tb=%d" % (ci, r.lX, e, res[0]["status"], res[0]["n_tracebacks"]),
               item_pairs(res, pairs, 0), want, totals[0], wtot)
    # unbanded
    r = synth.make_read(l1, 7, lX=500)
    batch = HostBatch([r.ref], [r.events], [np.zeros((0, 2))], model_ids=[mid], scales=[r.scale5], ragged=[(0, 0)])
    res, pairs, _ = eng.align_batch(batch, mode=MODE_UNBANDED)
    m = O.Model(O.THREE_STATE, tables=(l1, l2, l3), scale5=r.scale5)
    want, wtot = O.align_unbanded(m, r.ref, r.events)
    report("unbanded lX=500 tot=%.4f/%.4f" % (res[0]["total_logprob"], wtot), item_pairs(res, pairs, 0), want, None, None)
    print(eng.timing())


if __name__ == "__main__":
    main()
