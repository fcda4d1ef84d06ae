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


def test_py2_str_known_values():
    """str(float) of Python 2 (12 significant digits, '.0' for integers): the number format of the trained-model files
    scripts/nanoporeLib.py writes."""
    for x, want in [(0.1, "0.1"), (1.0 / 3.0, "0.333333333333"), (1.0, "1.0"), (1e-05, "1e-05"), (0.000244140625, "0.000244140625"),
                    (123456789012345.0, "1.23456789012e+14"), (-2.5, "-2.5"), (1e16, "1e+16"), (0.0, "0.0"),
                    (2.0 / 3.0, "0.666666666667"), (100.0, "100.0"), (float("-inf"), "-inf")]:
        assert em.py2_str(x) == want


def test_trained_model_file_round_trip(zymo, tmp_path):
    """The file the M-step writes (trainModels.py format) keeps 12 significant digits and is read back by the same
    loader as the C format."""
    m = _model_from(zymo["three_expectations_e20_r00"])
    m.normalize()
    path = str(tmp_path / "t.hmm")
    m.write_trained(path)
    lines = open(path).read().split("\n")
    assert lines[0] == "2\t3\t4096" and len(lines[1].split("\t")) == 10 and len(lines[2].split("\t")) == 4097
    back = em.ContinuousPairHmm.load(path)
    np.testing.assert_allclose(back.transitions, m.transitions, rtol=1e-11)
    np.testing.assert_allclose(back.kmer_skip_probs, m.kmer_skip_probs, rtol=1e-11)
    if R.available():      # and by the reference's own loader, into the state machine
        t_ref, g_ref = R.load_pair_hmm(path, synth.TEMPLATE_MODEL)
        t, g = back.state_machine_params()
        assert np.array_equal(t, t_ref) and np.array_equal(g, g_ref)


def test_vanilla_container(tmp_path):
    """ConditionalSignalHmm: alpha / beta normalised separately (nanoporeLib.py:1188-1197), files in both formats."""
    rng = np.random.default_rng(5)
    vec = np.concatenate([rng.uniform(0.5, 20.0, 60), [-12345.678]])
    m = em.ConditionalSignalHmm(pseudocount=1e-4)
    assert m.add_expectations(vec)
    np.testing.assert_allclose(m.kmer_skip_bins, vec[:60] + 1e-4)
    raw = m.kmer_skip_bins.copy()
    m.normalize()
    np.testing.assert_allclose(m.kmer_skip_bins[:30], raw[:30] / raw[:30].sum(), rtol=1e-15)
    np.testing.assert_allclose(m.kmer_skip_bins[30:], raw[30:] / raw[30:].sum(), rtol=1e-15)
    assert abs(m.kmer_skip_bins[:30].sum() - 1) < 1e-12 and abs(m.kmer_skip_bins[30:].sum() - 1) < 1e-12
    p = str(tmp_path / "v.hmm")
    m.write_trained(p)
    back = em.ConditionalSignalHmm.load(p)
    np.testing.assert_allclose(back.kmer_skip_bins, m.kmer_skip_bins, rtol=1e-11)
    assert back.likelihood == pytest.approx(-12345.678)
    assert not em.ConditionalSignalHmm().add_expectations(np.full(61, np.nan))
    with pytest.raises(ValueError):
        em.ContinuousPairHmm.load(p)                       # a vanilla file is not a three-state one
    with pytest.raises(ValueError):
        em.ConditionalSignalHmm().add_expectations(np.zeros(4106))


@pytest.mark.skipif(not R.available(), reason="oracle/_ref not built")
def test_vanilla_container_against_live_reference(tmp_path):
    """C-format writer byte for byte against vanillaHmm_writeToFile (raw and after the C container's joint
    normalisation); both of our file formats through the reference's loader into the vanilla state machine."""
    rng = np.random.default_rng(6)
    vec = np.concatenate([rng.uniform(0.01, 5.0, 60), [-4321.5]])
    l1, l2, l3 = synth.load_model_file(synth.TEMPLATE_MODEL)
    ref_raw, ref_norm, mine = str(tmp_path / "r.hmm"), str(tmp_path / "rn.hmm"), str(tmp_path / "m.hmm")
    R.write_vanilla_hmm(vec, synth.TEMPLATE_MODEL, ref_raw)
    R.write_vanilla_hmm(vec, synth.TEMPLATE_MODEL, ref_norm, normalize=True)
    m = em.ConditionalSignalHmm(match_model=l1, scaled_match_model=l3)     # implantMatchModels: MATCH and GAP_Y tables
    m.add_expectations(vec)
    m.write(mine)
    assert open(mine).read() == open(ref_raw).read()
    j = em.ConditionalSignalHmm(match_model=l1, scaled_match_model=l3)
    j.add_expectations(vec)
    j.normalize_joint()
    j.write(mine)
    assert open(mine).read() == open(ref_norm).read()
    m.normalize()
    for writer in (m.write, m.write_trained):
        writer(mine)
        bins = R.load_vanilla_hmm(mine, synth.TEMPLATE_MODEL)
        want = em.ConditionalSignalHmm.load(mine).state_machine_params()[1]
        assert np.array_equal(bins, want)


def test_cull_training_reads():
    """trainModels.py's culling rule: shuffled, cumulative length up to and including the read that crosses the mark."""
    lengths = np.array([500, 1200, 800, 300, 2500, 700])
    got = em.cull_training_reads(lengths, 2000, rng=np.random.default_rng(3))
    again = em.cull_training_reads(lengths, 2000, rng=np.random.default_rng(3))
    assert np.array_equal(got, again) and len(set(got.tolist())) == len(got)
    csum = np.cumsum(lengths[got])
    assert csum[-1] >= 2000 and (len(got) == 1 or csum[-2] < 2000)
    assert len(em.cull_training_reads(lengths, 10 ** 9, rng=np.random.default_rng(1))) == len(lengths)    # not enough data: all


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


def test_hdp_container_round_trip(tmp_path):
    """HdpHmm.write reproduces the reference CLI's `-d -t` file byte for byte from its parsed content."""
    import gzip
    gold = os.path.join(ROOT, "tests", "golden", "vanillaAlign", "t_d.exp.gz")
    h = em.HdpHmm.load(gold)
    assert len(h.means) == 13289 and h.threshold == 0.01 and em.HdpHmm.kmer_string(h.kmers[0]) == gzip.open(gold, "rt").read().split("\n")[3].split("\t")[0]
    out = str(tmp_path / "t.exp")
    h.write(out)
    assert open(out, "rb").read() == gzip.open(gold, "rb").read()


HDP_WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np
    import torch.distributed as dist
    sys.path.insert(0, os.path.join({root!r}, "cpecan-signal_b200"))
    from cpecan_signal import em
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    rng = np.random.default_rng(10 + rank)
    n = [5, 0, 12, 3][rank]                                   # ragged, one rank with nothing
    means, kmers = rng.normal(60, 5, n), rng.integers(0, 4096, n).astype(np.int32)
    v = np.arange(10, dtype=np.float64) * (rank + 1)
    em.allreduce_sum(v, True)
    m, k = em.gather_assignments(means, kmers, True)
    h = em.HdpHmm(1e-4, 0.01)
    h.add(v, m, k)
    h.write(sys.argv[1] + ".%d" % rank)
    dist.barrier()
    dist.destroy_process_group()
""")


def test_hdp_assignments_gather_world3_gloo(tmp_path):
    """Variable-length all-gather of the assignment lists (rank order) and all-reduce of the sums: every rank writes the
    same file, equal to the one built from the per-rank pieces by hand."""
    world = 3
    script = tmp_path / "hw.py"
    script.write_text(HDP_WORKER.format(root=ROOT))
    base = str(tmp_path / "h.exp")
    subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
                    "--master-port", "29641", str(script), base], check=True, timeout=600, capture_output=True)
    want = em.HdpHmm(1e-4, 0.01)
    ms, ks = [], []
    for r in range(world):
        rng = np.random.default_rng(10 + r)
        n = [5, 0, 12, 3][r]
        ms.append(rng.normal(60, 5, n)); ks.append(rng.integers(0, 4096, n).astype(np.int32))
    want.add(np.arange(10, dtype=np.float64) * sum(range(1, world + 1)), np.concatenate(ms), np.concatenate(ks))
    want.write(base + ".want")
    files = [open(base + ".%d" % r).read() for r in range(world)]
    assert files[0] == files[1] == files[2] == open(base + ".want").read()
    assert files[0].split("\n")[0].split("\t")[3] == "17"
