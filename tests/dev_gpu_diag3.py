"""Developer diagnostic: un-clamped log-posteriors GPU vs oracle."""
import os, sys
os.environ["CPECAN_DEBUG_LOGP"] = "1"
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cpecan-signal_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracleshim as O
from cpecan_signal import Engine, HostBatch, default_params, synth
from cpecan_signal.engine import item_pairs
eng = Engine(0)
l1, l2, l3 = synth.load_model_file(synth.TEMPLATE_MODEL)
mid = eng.upload_model(l1, l3, np.full(4096, -2.3025850929940455))
O.lib().oracle_debug_logp(1)
for idx, lx, e in ((100, 400, 40), (1, 900, 40), (106, 6700, 128), (105, 6700, 64)):
    r = synth.make_read(l1, idx, lX=lx)
    batch = HostBatch([r.ref], [r.events], [r.anchors], model_ids=[mid], scales=[r.scale5], ragged=[(1, 1)])
    res, pairs, _ = eng.align_batch(batch, params=default_params(diagonalExpansion=e))
    got = item_pairs(res, pairs, 0)
    m = O.Model(O.THREE_STATE, tables=(l1, l2, l3), scale5=r.scale5)
    want, _ = O.align_banded(m, r.ref, r.events, r.anchors, params=O.default_params(diagonalExpansion=e), ragged=(1, 1))
    g = {(int(x), int(y)): float(np.int32(s).view(np.float32)) for s, x, y in got}
    w = {(int(x), int(y)): s * 1e-9 for s, x, y in want}
    common = sorted(set(g) & set(w), key=lambda k: k[0] + k[1])
    dl = np.array([g[k] - w[k] for k in common])
    dd = np.array([k[0] + k[1] + 2 for k in common])
    print("case", idx, "n", len(g), len(w), "max|dlogp|", np.abs(dl).max(), "mean", dl.mean(), "rms", dl.std())
    bad = np.where(np.abs(dl) > 5e-5)[0]
    print("   n>5e-5:", len(bad), " diagonals:", sorted(set((dd[bad] // 1).tolist()))[:40])
    for i in bad[:12]:
        k = common[i]
        print("     d=%d x=%d y=%d gpu=%.6f ref=%.6f diff=%.2e" % (dd[i], k[0], k[1], g[k], w[k], dl[i]))
