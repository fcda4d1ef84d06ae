"""threeStateHdp, CPU side: the .nhdp readers (Python and the host library's deserialize_nhdp), the
density query and the oracle's restatement of the machine, against goldens produced by the UNMODIFIED reference with its
HDP sources on the reference's own serialised fixture (oracle/make_golden.py hdp)."""
import ctypes as C
import os

import numpy as np

import oracleshim as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "cpecan-signal_b200")


def _kmer(k):
    return "".join("ACGT"[(k >> (2 * (5 - j))) & 3] for j in range(6))


def test_python_reader_sees_the_fixture(hdp_fixture):
    h = hdp_fixture["hdp"]
    assert h.alphabet == "ACEGOT" and h.kmer_length == 6 and h.grid_length == 100 and (h.grid_start, h.grid_stop) == (0.0, 100.0)
    assert h.density.shape == h.slopes.shape == (718, 100)        # 717 observed leaves and the base process
    assert (h.kmer_distr >= 0).all() and h.row[-1] == 717 and h.parent[-1] == -1
    # a k-mer without data of its own reads its nearest observed ancestor: the base process
    assert (h.kmer_distr == 717).sum() == 4096 - len(set(h.kmer_distr.tolist()) - {717})


def test_oracle_density_equals_the_reference(hdp_fixture):
    g, h = hdp_fixture["golden"], hdp_fixture["hdp"]
    m = O.Model(O.THREE_STATE_HDP, hdp=h)
    got = np.array([[O.hdp_density(m, k, x) for x in g["density_x"]] for k in g["density_kmers"]])
    assert np.array_equal(got, g["density"])                      # the same expressions in the same order: bit-identical
    assert (got >= 0).all() and got.max() > 0.02


def test_host_library_reads_and_queries_the_fixture(hdp_fixture):
    """deserialize_nhdp / get_nanopore_kmer_density of libcpecan_host.so (no GPU involved)."""
    lib = C.CDLL(os.path.join(PKG, "libcpecan_host.so"))
    lib.deserialize_nhdp.restype = C.c_void_p
    lib.get_nanopore_kmer_density.restype = C.c_double
    lib.get_nanopore_kmer_density.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_double)]
    lib.get_nanopore_hdp_kmer_length.restype = C.c_int64
    lib.get_nanopore_hdp_kmer_length.argtypes = [C.c_void_p]
    lib.get_nanopore_hdp_alphabet_size.restype = C.c_int64
    lib.get_nanopore_hdp_alphabet_size.argtypes = [C.c_void_p]
    lib.destroy_nanopore_hdp.argtypes = [C.c_void_p]
    nh = lib.deserialize_nhdp(hdp_fixture["path"].encode())
    assert nh and lib.get_nanopore_hdp_kmer_length(nh) == 6 and lib.get_nanopore_hdp_alphabet_size(nh) == 6
    g = hdp_fixture["golden"]
    got = np.array([[lib.get_nanopore_kmer_density(nh, _kmer(int(k)).encode(), C.byref(C.c_double(float(x))))
                     for x in g["density_x"]] for k in g["density_kmers"]])
    lib.destroy_nanopore_hdp(nh)
    assert np.array_equal(got, g["density"])


def test_oracle_alignment_equals_the_reference(hdp_fixture, zymo):
    """getAlignedPairsUsingAnchors with getHdpStateMachine3 / sequence_getKmer3 on the fixture read: the oracle's pair
    lists, scores and per-diagonal totals are the reference's."""
    g = hdp_fixture["golden"]
    m = O.Model(O.THREE_STATE_HDP, hdp=hdp_fixture["hdp"])
    for tag, e, ragged, thr in (("hdp_e20_r00", 20, (0, 0), 0.01), ("hdp_e50_r11", 50, (1, 1), 0.01), ("hdp_e20_r10_t30", 20, (1, 0), 0.3)):
        pairs, totals = O.align_banded(m, zymo["ref"], hdp_fixture["events"], zymo["anchors_template"],
                                       params=O.default_params(diagonalExpansion=e, threshold=thr), ragged=ragged, want_totals=True)
        want = g[tag + "_pairs"]
        assert np.array_equal(np.asarray(pairs, dtype=np.int64).reshape(-1, 3), want), tag
        wt = g[tag + "_totals"]
        mask = ~np.isnan(wt)
        assert np.array_equal(mask, ~np.isnan(totals)) and np.array_equal(totals[mask], wt[mask]), tag


def test_oracle_hdp_expectations_equal_the_reference(hdp_fixture, zymo):
    """getExpectationsUsingAnchors with an HdpHmm: the nine transition sums and the likelihood to 1e-12, the
    event-to-k-mer assignments (k-mer position, event index) in the reference's list order."""
    g = hdp_fixture["golden"]
    m = O.Model(O.THREE_STATE_HDP, hdp=hdp_fixture["hdp"])
    for tag, e, ragged, thr in (("hdpexp_e50_r11", 50, (1, 1), 0.01), ("hdpexp_e20_r00_t30", 20, (0, 0), 0.3)):
        vec, asg = O.hdp_expectations(m, zymo["ref"], hdp_fixture["events"], zymo["anchors_template"],
                                      params=O.default_params(diagonalExpansion=e, threshold=thr), ragged=ragged)
        np.testing.assert_allclose(vec, g[tag + "_vec"], rtol=1e-12, atol=1e-15)
        assert np.array_equal(asg[:, 1:], g[tag + "_assignments"]), tag
        assert set(asg[:, 0].tolist()) <= {0, 1, 2}


def test_reader_rejects_what_the_machine_cannot_use(tmp_path, hdp_fixture):
    """An HDP without finalized distributions (no data / splines not finalized) cannot be queried (the reference exits in
    dir_proc_density, impl/hdp.c:2578-2581): the Python reader raises, and a truncated file is reported, not mis-read."""
    import pytest
    from cpecan_signal import hdp
    lines = open(hdp_fixture["path"]).read().split("\n")
    unfinalized = lines[:3] + ["0"] + lines[4:]
    p = tmp_path / "u.nhdp"
    p.write_text("\n".join(unfinalized))
    with pytest.raises(ValueError, match="finalized"):
        hdp.load_nhdp(str(p))
    p2 = tmp_path / "t.nhdp"
    p2.write_text("\n".join(lines[:40000]))
    with pytest.raises(ValueError):
        hdp.load_nhdp(str(p2))
    # gzip-compressed files are read directly
    h = hdp.load_nhdp(os.path.join(ROOT, "tests", "golden", "hdp", "testTemplate.nhdp.gz"))
    assert h.density.shape == hdp_fixture["hdp"].density.shape and np.array_equal(h.kmer_distr, hdp_fixture["hdp"].kmer_distr)
