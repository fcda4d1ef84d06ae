"""Developer perf probe (not a test): kernel-only timing of one expansion; the command ncu captures.
   python tools/dev_perf.py --e 64 --n 1776 --reps 3 [--lx 6700]"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "cpecan-signal_b200"))
from bench import generate_reads  # noqa: E402
from cpecan_signal import Engine, HostBatch, default_params, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--e", type=int, default=64)
    ap.add_argument("--n", type=int, default=1776)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--lx", type=int, default=6700)
    ap.add_argument("--unique", type=int, default=0, help="generate only this many distinct reads and tile them")
    ap.add_argument("--machine", default="three", choices=["three", "vanilla", "four", "echelon", "hdp"],
                    help="four / echelon / hdp always run on the FP64 kernel")
    ap.add_argument("--mode", default="posterior", choices=["posterior", "em"])
    ap.add_argument("--thr", type=float, default=0.01, help="posterior threshold (1.1: no pair is ever reported)")
    ap.add_argument("--exact", action="store_true", help="cpecan_cuda_set_exact_arithmetic: the FP64 kernel")
    a = ap.parse_args()
    nu = a.unique or a.n
    reads = generate_reads(nu, 5_000_000, lX=a.lx)
    reads = [reads[i % nu] for i in range(a.n)]
    l1, l2, l3 = synth.load_model_file(synth.TEMPLATE_MODEL)
    eng = Engine(0)
    if a.exact:
        eng.set_exact_arithmetic(True)
    from cpecan_signal import echelon_hmm, four_state_hmm, hdp, hdp_hmm, vanilla_gapx, vanilla_hmm
    hmm = {"vanilla": lambda: vanilla_hmm("template"), "four": four_state_hmm, "echelon": echelon_hmm, "hdp": hdp_hmm,
           "three": lambda: None}[a.machine]()
    scales = [r.scale5 for r in reads]
    events = [r.events for r in reads]
    if a.machine == "hdp":
        mid = eng.upload_hdp(hdp.load_nhdp(os.path.join(ROOT, "tests", "golden", "hdp", "testTemplate.nhdp.gz")))
        events = [np.column_stack([(np.asarray(r.events).reshape(-1, 3)[:, 0] - r.scale5[1]) / r.scale5[0],
                                   np.asarray(r.events).reshape(-1, 3)[:, 1:]]) for r in reads]
        scales = None
    else:
        gapx = vanilla_gapx(l2) if a.machine in ("vanilla", "echelon") else (np.zeros(4096) if a.machine == "four" else np.full(4096, -2.3025850929940455))
        mid = eng.upload_model(l1, l3, gapx)
    hb = HostBatch([r.ref for r in reads], events, [r.anchors for r in reads],
                   model_ids=[mid] * len(reads), scales=scales, ragged=[(1, 1)] * len(reads))
    if a.mode == "em":
        eng.stage(hb, hmm=hmm, params=default_params(diagonalExpansion=a.e), mode=1, pair_cap=1)
    else:
        eng.stage(hb, hmm=hmm, params=default_params(diagonalExpansion=a.e, threshold=0.15 if a.machine == "echelon" and a.thr == 0.01 else a.thr),
                  pair_cap=eng.default_pair_capacity(hb, 12 if a.machine in ("echelon", "hdp") else 3))
    cells = eng.timing()["band_cells"]
    for i in range(a.reps):
        eng.run_staged()
        t = eng.timing()
        print("e=%d n=%d cells=%.3e align_ms=%.2f  %.2f Gcells/s  %.2f GCUPS  G=%d ctas=%d" % (
            a.e, a.n, cells, t["align_ms"], cells / t["align_ms"] / 1e6, 2 * cells / t["align_ms"] / 1e6,
            t["warps_per_item"], t["ctas"]), flush=True)
    if a.mode == "em":
        vec = np.zeros(eng.N_EXPECT_VANILLA if a.machine == "vanilla" else eng.N_EXPECT)
        eng.fetch_expectations(vec)
        print("expectations:", vec[:9], vec[-1])
        return
    res, _ = eng.fetch_staged()
    print("status!=0:", int((res["status"] != 0).sum()), "pairs:", int(res["n_pairs"].sum()))


if __name__ == "__main__":
    main()
