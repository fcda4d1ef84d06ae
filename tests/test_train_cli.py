"""cpecan_signal.train -- the command-line Baum-Welch driver (sibling of the reference's scripts/trainModels.py): its
caller-side glue (cigar -> guide anchors -> event-space anchors, strand slices) against the reference CLI's own numbers,
and one training run end to end."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
VA = os.path.join(GOLD, "vanillaAlign")


def _fixture_read():
    from cpecan_signal import train
    cigar = open(os.path.join(VA, "guide.cigar")).readline().strip()
    return train.prepare_read("readA", os.path.join(GOLD, "ZymoC_ch_1_file1.npRead"), os.path.join(GOLD, "ZymoRef.txt"), cigar)


def test_filter_to_remove_overlap_matches_the_oracle():
    import oracleshim as O
    from cpecan_signal import train
    rng = np.random.default_rng(5)
    for _ in range(50):
        n = int(rng.integers(0, 60))
        p = np.stack([rng.integers(0, 40, n), rng.integers(0, 40, n)], axis=1)
        p = p[np.lexsort((p[:, 1], p[:, 0]))] if n else p.reshape(0, 2)
        assert np.array_equal(train.filter_to_remove_overlap(p), O.filter_overlap(p))


def test_guide_anchors_of_the_fixture():
    """The reference CLI reports 175 guide anchors for the fixture read and its lastz cigar (stdout_*.txt); the template
    strand spans the events the posterior file of the reference covers."""
    from cpecan_signal import train
    cigar = train.parse_cigar(open(os.path.join(VA, "guide.cigar")).readline().strip())
    assert len(train.guide_anchors(cigar, 14)) == int(open(os.path.join(VA, "stdout_s.txt")).read().split()[1]) == 175
    jobs, length = _fixture_read()
    assert [j.strand for j in jobs] == [0]                     # the fixture's complement event map runs backwards
    t = jobs[0]
    assert (np.diff(t.anchors[:, 0]) > 0).all() and (np.diff(t.anchors[:, 1]) > 0).all()
    rows = [l.split("\t") for l in open(os.path.join(VA, "out_s.tsv"))]
    ys = np.array([int(r[5]) for r in rows])
    from cpecan_signal import synth
    emap = synth.load_npread(os.path.join(GOLD, "ZymoC_ch_1_file1.npRead"))["template_map"]
    y0 = int(emap[cigar["start2"]])
    assert ys.min() >= y0 and ys.max() < y0 + len(t.events)


@pytest.mark.gpu
def test_first_estep_equals_the_reference_cli_expectation_file():
    """One E-step of the fixture read's template strand through the trainer's glue, in the reference's own arithmetic:
    the expectation file `vanillaAlign -s -t` writes for it (pseudocount 1e-4 in every slot, six decimals)."""
    from cpecan_signal import Engine, default_params, synth, three_state_hmm, train
    jobs, _ = _fixture_read()
    eng = Engine(0)
    eng.set_exact_arithmetic(True)
    tables = synth.load_model_file(synth.TEMPLATE_MODEL)
    vec = train.estep_strand(eng, jobs, tables, "three", dict(hmm=three_state_hmm(), gapx=np.full(4096, -2.3025850929940455)),
                             default_params(diagonalExpansion=50), False)
    eng.close()
    with open(os.path.join(VA, "t_s.exp")) as fh:
        fh.readline()
        l1 = np.array(fh.readline().split(), dtype=np.float64)
        l2 = np.array(fh.readline().split(), dtype=np.float64)
    np.testing.assert_allclose(vec[:9] + 1e-4, l1[:9], rtol=0, atol=1.5e-6)
    assert abs(vec[-1] - l1[9]) <= 1.5e-6 * max(1.0, abs(l1[9]))
    np.testing.assert_allclose(vec[9:-1] + 1e-4, l2, rtol=0, atol=1.5e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("machine", ["three", "vanilla"])
def test_training_run(tmp_path, machine):
    """Three iterations on a two-read manifest: the template .hmm file is written in the trained format and reloads; from
    the first M-step on the likelihood does not get worse (the first M-step replaces the threeState machine's unnormalised
    0.1 gap-emission prior by k-mer skip probabilities that sum to one, which lowers the likelihood once; the reference's
    EM test allows 5 % slack, tests/signalPairwiseTest.c:1604-1714)."""
    from cpecan_signal import em, train
    cigar = open(os.path.join(VA, "guide.cigar")).readline().strip()
    man = tmp_path / "reads.tsv"
    row = "\t".join(["%s", os.path.join(GOLD, "ZymoC_ch_1_file1.npRead"), os.path.join(GOLD, "ZymoRef.txt"), "-", cigar])
    man.write_text(row % "readA" + "\n" + row % "readB" + "\n")
    t, c = str(tmp_path / "t.hmm"), str(tmp_path / "c.hmm")
    log = train.main(["--manifest", str(man), "--machine", machine, "--iterations", "3", "--out-template-hmm", t,
                      "--out-complement-hmm", c])
    liks = [l for name, it, l in log if name == "template"]
    print(machine, liks)
    assert len(liks) == 3 and liks[2] >= liks[1] - 0.05 * abs(liks[1])
    cls = em.ConditionalSignalHmm if machine == "vanilla" else em.ContinuousPairHmm
    trained = cls.load(t)
    if machine == "three":
        np.testing.assert_allclose(trained.transitions.reshape(3, 3).sum(axis=1), 1.0, rtol=1e-9)
    assert not os.path.exists(c)                               # the fixture has no complement stretch to train on


@pytest.mark.gpu
def test_hdp_pass_writes_the_reference_cli_file(tmp_path, hdp_fixture):
    """--machine hdp on the fixture read: the HdpHmm file of the template strand equals what `vanillaAlign -d -t` writes
    (13 289 assignments: every event mean and k-mer in order; sums at the six printed decimals)."""
    import gzip
    from cpecan_signal import train
    cigar = open(os.path.join(VA, "guide.cigar")).readline().strip()
    man = tmp_path / "reads.tsv"
    man.write_text("\t".join(["readA", os.path.join(GOLD, "ZymoC_ch_1_file1.npRead"), os.path.join(GOLD, "ZymoRef.txt"), cigar]) + "\n")
    t, c = str(tmp_path / "t.exp"), str(tmp_path / "c.exp")
    train.main(["--manifest", str(man), "--machine", "hdp", "--template-hdp", hdp_fixture["path"], "--complement-hdp", hdp_fixture["path"],
                "--out-template-hmm", t, "--out-complement-hmm", c])
    got = open(t).read().split("\n")
    with gzip.open(os.path.join(VA, "t_d.exp.gz"), "rt") as fh:
        want = fh.read().split("\n")
    assert got[0] == want[0] and len(got) == len(want)
    np.testing.assert_allclose(np.array(got[1].split(), dtype=np.float64), np.array(want[1].split(), dtype=np.float64), rtol=0, atol=2e-6)
    assert got[2] == want[2] and got[3] == want[3]
    assert open(c).read().split("\n")[0].split("\t")[3] == "0"          # the complement strand of the fixture has no stretch
