"""CPU tests of the EM host logic: expectation container format / normalisation / M-step load against the reference's
own code (oracle/_ref where present, committed golden text otherwise), the read sharding, and a world_size-2 gloo run
of the all-reduced EM iteration with the oracle as the E-step."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

import oracleshim as O
import refshim as R
from cpecan_signal import em, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def _model_from(vec):
    m = em.ContinuousPairHmm()
    m.add_expectations(vec)
    return m


def test_write_matches_committed_reference_file(zymo):
    """tests/golden/zymo_three_e20.expectations was written by the reference's continuousPairHmm_writeToFile."""
    vec = zymo["three_expectations_e20_r00"]
    path = os.path.join(GOLDEN, "zymo_three_e20.expectations")
    want = open(path).read()
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        out = os.path.join(td, "x.hmm")
        _model_from(vec).write(out)
        assert open(out).read() == want
        back = em.ContinuousPairHmm.load(out)
        np.testing.assert_allclose(back.vector(), vec, atol=5.1e-7)      # "%f": 6 decimals


@pytest.mark.skipif(not R.available(), reason="oracle/_ref not built")
def test_container_against_live_reference(zymo, tmp_path):
    """write / normalise / loadTransitionsAndKmerGapProbs, byte for byte and bit for bit against impl/continuousHmm.c."""
    for key in ("three_expectations_e20_r00", "three_expectations_e50_r11"):
        vec = zymo[key]
        ref_raw, ref_norm, mine = str(tmp_path / "r.hmm"), str(tmp_path / "rn.hmm"), str(tmp_path / "m.hmm")
        R.write_pair_hmm(vec, ref_raw)
        R.write_pair_hmm(vec, ref_norm, normalize=True)
        m = _model_from(vec)
        m.write(mine)
        assert open(mine).read() == open(ref_raw).read()
        m.normalize()
        m.write(mine)
        assert open(mine).read() == open(ref_norm).read()
        t_ref, g_ref = R.load_pair_hmm(ref_norm, synth.TEMPLATE_MODEL)
        t, g = em.ContinuousPairHmm.load(mine).state_machine_params()
        assert np.array_equal(t, t_ref) and np.array_equal(g, g_ref)


def test_load_errors(tmp_path):
    p = tmp_path / "bad.hmm"
    p.write_text("2\t3\t4096\t\n0.1\t0.2\n\n")
    with pytest.raises(ValueError, match="Incorrect number of transitions"):
        em.ContinuousPairHmm.load(str(p))
    p.write_text("4\t3\t4096\t\n")
    with pytest.raises(ValueError, match="not a three-state"):
        em.ContinuousPairHmm.load(str(p))


def test_nan_expectations_are_skipped():
    m = em.ContinuousPairHmm()
    v = np.ones(em.N_EXPECT)
    v[3] = np.nan
    assert m.add_expectations(v) is False and m.transitions.sum() == 0


def test_shard_by_cells_balanced():
    rng = np.random.default_rng(0)
    cells = rng.integers(1_000_000, 4_000_000, size=1000)
    parts = em.shard_by_cells(cells, 8)
    allidx = np.sort(np.concatenate(parts))
    assert np.array_equal(allidx, np.arange(1000))
    loads = np.array([cells[p].sum() for p in parts])
    assert loads.max() / loads.mean() < 1.01
    assert em.shard_by_cells([5, 1], 4)[0].tolist() == [0]      # more ranks than reads: empty shards are fine


WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np
    import torch.distributed as dist
    sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "cpecan-signal_b200"))
    sys.path.insert(0, os.path.join({root!r}, "oracle"))
    import oracleshim as O
    from cpecan_signal import em, synth
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    l1, l2, l3 = synth.load_model_file(synth.TEMPLATE_MODEL)
    reads = [synth.make_read(l1, 700 + i, lX=150 + 40 * (i % 3)) for i in range(6)]
    p = O.default_params(diagonalExpansion=20)
    cells = [O.band_cells(r.anchors, r.lX, r.lY, p, (1, 1)) for r in reads]
    mine = em.shard_by_cells(cells, world)[rank]
    model = em.ContinuousPairHmm()
    hmm_path = sys.argv[1]
    trans, gapx = None, None
    liks = []
    for it in range(3):
        vec = np.zeros(em.N_EXPECT)
        for i in mine:
            r = reads[i]
            m = O.Model(O.THREE_STATE, tables=(l1, l2, l3), scale5=r.scale5, transitions=trans, gap_x=gapx)
            vec += O.expectations(m, r.ref, r.events, r.anchors, params=p, ragged=(1, 1), pseudocount=0.0)
        em.allreduce_sum(vec, world > 1)
        loaded = em.em_iteration(model, vec, len(reads), hmm_path, rank=rank, barrier=dist.barrier)
        trans, gapx = loaded.state_machine_params()
        liks.append(model.running_likelihoods[-1])
    if rank == 0:
        np.save(sys.argv[2], np.array(liks))
    dist.barrier()
    dist.destroy_process_group()
""")


def _run_workers(world, tmp_path, tag):
    script = tmp_path / ("w%s.py" % tag)
    script.write_text(WORKER.format(root=ROOT))
    hmm, out = str(tmp_path / ("t%s.hmm" % tag)), str(tmp_path / ("l%s.npy" % tag))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29611 + world), str(script), hmm, out]
    subprocess.run(cmd, check=True, timeout=600, capture_output=True)
    return np.load(out), open(hmm).read()


def test_em_world2_gloo_matches_single_rank(tmp_path):
    """Sharded E-step + all-reduce + identical M-step on every rank == the single-rank run; the likelihood of the
    trained model does not get worse (the reference's own EM test asserts the same with 5 % slack,
    tests/signalPairwiseTest.c:1604-1714)."""
    l1, h1 = _run_workers(1, tmp_path, "a")
    l2, h2 = _run_workers(2, tmp_path, "b")
    np.testing.assert_allclose(l2, l1, rtol=1e-12)
    assert h1 == h2
    assert l1[-1] >= l1[0] - 0.05 * abs(l1[0])
