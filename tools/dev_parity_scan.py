"""Developer tool (not a test): per-read worst score difference of a library build against the oracle on C3-shaped reads.
   CPECAN_LIB=... python tools/dev_parity_scan.py --e 64 --first 70000 --n 64 [--dump IDX]"""
import argparse
import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("cpecan-signal_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))


def make(args):
    from cpecan_signal import synth
    return synth.make_read(synth.load_model_file(synth.TEMPLATE_MODEL)[0], args[0], lX=args[1])


def oracle(args):
    import oracleshim as O
    from cpecan_signal import synth
    r, e = args
    m = O.Model(O.THREE_STATE, model_file=synth.TEMPLATE_MODEL, scale5=r.scale5)
    return O.align_banded(m, r.ref, r.events, r.anchors, params=O.default_params(diagonalExpansion=e), ragged=(1, 1), want_totals=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--e", type=int, default=64)
    ap.add_argument("--first", type=int, default=70000)
    ap.add_argument("--n", type=int, default=64)
    ap.add_argument("--lx", type=int, default=6700)
    ap.add_argument("--top", type=int, default=6)
    a = ap.parse_args()
    import parity
    from cpecan_signal import Engine, HostBatch, default_params, synth
    from cpecan_signal.engine import item_pairs
    with mp.get_context("fork").Pool(min(a.n, os.cpu_count())) as pool:
        reads = pool.map(make, [(a.first + i, a.lx) for i in range(a.n)])
        want = pool.map(oracle, [(r, a.e) for r in reads], chunksize=1)
    l1, l2, l3 = synth.load_model_file(synth.TEMPLATE_MODEL)
    eng = Engine(0)
    mid = eng.upload_model(l1, l3, np.full(4096, -2.3025850929940455))
    hb = HostBatch([r.ref for r in reads], [r.events for r in reads], [r.anchors for r in reads], model_ids=[mid] * a.n,
                   scales=[r.scale5 for r in reads], ragged=[(1, 1)] * a.n)
    res, pairs, totals = eng.align_batch(hb, params=default_params(diagonalExpansion=a.e), want_totals="terms")
    rows = []
    for i in range(a.n):
        g = parity.pair_dict(item_pairs(res, pairs, i)); w = parity.pair_dict(want[i][0])
        diffs = sorted(((abs(g[k] - w[k]), k, g[k], w[k]) for k in w if k in g), reverse=True)
        missing = [(k, w[k]) for k in w if k not in g and w[k] > 101002] + [(k, g[k]) for k in g if k not in w and g[k] > 101002]
        mask = ~np.isnan(want[i][1])
        dt = np.abs(totals[i][0][mask] - want[i][1][mask])
        rows.append((diffs[0][0] if diffs else 0, i, diffs[:a.top], missing[:4], float(dt.max()), int(np.flatnonzero(mask)[dt.argmax()])))
    rows.sort(reverse=True)
    print("lib", os.environ.get("CPECAN_LIB", "default"), "e", a.e)
    for worst, i, diffs, missing, dtmax, dtat in rows[:8]:
        print("read %d (idx %d): worst %d  max|dtotal| %.3g at d=%d  missing %s" % (i, a.first + i, worst, dtmax, dtat, missing))
        for df, k, gv, wv in diffs:
            print("      %6d  pair %s d=%d  got %d want %d" % (df, k, k[0] + k[1] + 2, gv, wv))
    print("over tolerance:", sum(1 for r in rows if r[0] > parity.SCORE_TOL), "of", a.n)
    allover = [(df, a.first + i, k, gv, wv) for worst, i, diffs, missing, dtmax, dtat in rows for df, k, gv, wv in diffs if df > parity.SCORE_TOL]
    print("pairs over tolerance:", sorted(allover, reverse=True))
    print("n pairs:", sum(len(w[0]) for w in want))


if __name__ == "__main__":
    main()
