"""Developer tool (not a test): latency of ONE alignment through the C-ABI with host buffers (cpecan_cuda_align_batch:
H2D, preparation, plan, kernel, D2H) -- the shape of the reference's per-read call getAlignedPairsUsingAnchors.
   python tools/dev_latency.py"""
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cpecan-signal_b200"))
from cpecan_signal import Engine, HostBatch, default_params, synth  # noqa: E402


def timed(eng, hb, params, n=15):
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        res, pairs, _ = eng.align_batch(hb, params=params)
        ts.append((time.perf_counter() - t0) * 1e3)
    tm = eng.timing()
    return float(np.median(ts[3:])), tm, int(res[0]["n_pairs"]), int(res[0]["band_cells"])


def main():
    l1, l2, l3 = synth.load_model_file(synth.TEMPLATE_MODEL)
    gold = os.path.join(ROOT, "tests", "golden")
    z = dict(np.load(os.path.join(gold, "zymo_golden.npz")))
    rd = synth.load_npread(os.path.join(gold, "ZymoC_ch_1_file1.npRead"))
    ref = open(os.path.join(gold, "ZymoRef.txt")).readline().strip()
    eng = Engine(0)
    mid = eng.upload_model(l1, l3, np.full(4096, -2.3025850929940455))
    out = {}
    cases = [("C1 fixture read (892 x 799, e=20)", HostBatch([ref], [rd["template_events"]], [z["anchors_template"]], model_ids=[mid],
                                                              scales=[rd["template_params"]], ragged=[(0, 0)]), default_params())]
    r2 = synth.make_read(l1, 52000, lX=5000)
    cases.append(("C2-sized strand (5000 x ~6000, e=50)", HostBatch([r2.ref], [r2.events], [r2.anchors], model_ids=[mid],
                                                                     scales=[r2.scale5], ragged=[(1, 1)]), default_params(diagonalExpansion=50)))
    r3 = synth.make_read(l1, 70000, lX=6700)
    for e in (64, 256):
        cases.append(("C3 read (6700 x ~8000, e=%d)" % e, HostBatch([r3.ref], [r3.events], [r3.anchors], model_ids=[mid],
                                                                    scales=[r3.scale5], ragged=[(1, 1)]), default_params(diagonalExpansion=e)))
    for name, hb, prm in cases:
        ms, tm, npairs, cells = timed(eng, hb, prm)
        out[name] = dict(ms_per_call=round(ms, 2), kernel_ms=round(tm["align_ms"], 2), band_cells=cells, pairs=npairs,
                         band_cells_per_s=round(cells / ms * 1e3))
        print(name, out[name], flush=True)
    # the two strands of vanillaAlign.c:737-790 as two host threads, one context each
    eng2 = Engine(0)
    mid2 = eng2.upload_model(l1, l3, np.full(4096, -2.3025850929940455))
    hb_a, prm = cases[0][1], cases[0][2]
    hb_b = HostBatch([ref], [rd["template_events"]], [z["anchors_template"]], model_ids=[mid2], scales=[rd["template_params"]], ragged=[(0, 0)])
    def both():
        th = [threading.Thread(target=lambda: eng.align_batch(hb_a, params=prm)), threading.Thread(target=lambda: eng2.align_batch(hb_b, params=prm))]
        [t.start() for t in th]; [t.join() for t in th]
    ts = []
    for _ in range(12):
        t0 = time.perf_counter(); both(); ts.append((time.perf_counter() - t0) * 1e3)
    out["two strands at once (two contexts, two host threads), C1"] = dict(ms_for_both=round(float(np.median(ts[3:])), 2))
    print(out["two strands at once (two contexts, two host threads), C1"])
    print(json.dumps(out))


if __name__ == "__main__":
    main()
