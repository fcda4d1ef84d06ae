"""The E-step all-reduce of the C-ABI (cpecan_cuda_nccl_init / cpecan_cuda_allreduce_expectations): one context per
GPU, one host thread per context (the shape of a C driver), sums checked against the per-rank sums.  Needs NCCL; with
one GPU the communicator has one rank (the collective still runs), with more every visible GPU joins."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_cabi_allreduce_expectations(template_tables):
    import torch
    from cpecan_signal import Engine, HostBatch, default_params, synth, three_state_hmm
    n_dev = min(torch.cuda.device_count(), 8)
    assert n_dev >= 1
    l1, l2, l3 = template_tables
    engines = [Engine(d) for d in range(n_dev)]
    uid = engines[0].nccl_unique_id()
    per_rank, after, errs = [None] * n_dev, [None] * n_dev, []

    def work(r):
        try:
            eng = engines[r]
            eng.nccl_init(n_dev, r, uid)
            reads = [synth.make_read(l1, 3000 + 10 * r + i, lX=300 + 40 * i) for i in range(3)]
            mid = eng.upload_model(l1, l3, np.full(4096, -2.3025850929940455))
            hb = HostBatch([x.ref for x in reads], [x.events for x in reads], [x.anchors for x in reads],
                           model_ids=[mid] * 3, scales=[x.scale5 for x in reads], ragged=[(1, 1)] * 3)
            eng.stage(hb, hmm=three_state_hmm(), params=default_params(diagonalExpansion=30), mode=1, pair_cap=1)
            eng.run_staged()
            mine = np.zeros(Engine.N_EXPECT)
            eng.fetch_expectations(mine)
            per_rank[r] = mine
            eng.allreduce_expectations()
            tot = np.zeros(Engine.N_EXPECT)
            eng.fetch_expectations(tot)
            after[r] = tot
        except Exception as ex:       # noqa: BLE001
            errs.append(ex)

    ths = [threading.Thread(target=work, args=(r,)) for r in range(n_dev)]
    [t.start() for t in ths]
    [t.join(timeout=300) for t in ths]
    assert not errs, errs
    want = np.sum(per_rank, axis=0)
    for r in range(n_dev):
        np.testing.assert_allclose(after[r], want, rtol=1e-12, atol=1e-12)
        assert after[r][0] > 1.0 and np.isfinite(after[r][-1])
    for e in engines:
        e.close()
    print("ranks", n_dev, "sum of match->match expectations", want[0])
