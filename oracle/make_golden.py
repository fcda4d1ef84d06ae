#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY.  Regenerates tests/golden/zymo_golden.npz and tests/golden/synthetic_golden.npz by
running the UNMODIFIED reference (oracle/_ref/libcpecan_ref.so + its vendored lastz, both built by oracle/Makefile
from /root/reference) on

  * the reference's own fixture read (tests/test_npReads/ZymoC_ch_1_file1.npRead vs ZymoRef.txt; the configuration of
    tests/signalPairwiseTest.c:1116-1183 and :1250-1310) -- anchors come from the reference's lastz pipeline;
  * a few small seeded synthetic reads (cpecan_signal.synth), including multi-traceback and split-region cases.

Only runs where /root/reference exists (the build container).  The outputs are committed; tests never need the
reference at run time.   Usage:  make -C oracle ref lastz && python oracle/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "cpecan-signal_b200"))

import refshim as R  # noqa: E402
from cpecan_signal import synth  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = os.environ.get("CPECAN_REFERENCE", "/root/reference")
T_MODEL = os.path.join(REFERENCE, "models", "template_median68pA.model")
C_MODEL = os.path.join(REFERENCE, "models", "complement_median68pA_pop2.model")


def revcomp(s):
    return s[::-1].translate(str.maketrans("ACGTacgt", "TGCAtgca"))


def main():
    os.chdir(os.path.join(HERE, "_ref"))  # the reference shells out to ./cPecanLastz
    ref = open(os.path.join(GOLDEN, "ZymoRef.txt")).readline().strip()
    npfile = os.path.join(GOLDEN, "ZymoC_ch_1_file1.npRead")
    rd = R.load_npread(npfile)
    out = {}
    anch_t, raw = R.fixture_anchors(ref, npfile, 0)
    anch_c, _ = R.fixture_anchors(ref, npfile, 1)
    out["anchors_raw_count"] = np.int64(raw)
    out["anchors_template"] = anch_t
    out["anchors_complement"] = anch_c
    tp, cp = rd["template_params"], rd["complement_params"]
    tev, cev = rd["template_events"], rd["complement_events"]

    def banded(tag, smt, model, refseq, ev, anchors, scale, strand, e, ragged):
        p = R.default_params(diagonalExpansion=e)
        pairs, totals = R.align_banded(smt, model, refseq, ev, anchors, params=p, scale5=scale, strand=strand,
                                       ragged=ragged, want_totals=True)
        out[tag + "_pairs"] = pairs
        out[tag + "_totals"] = totals
        print(tag, len(pairs), int(pairs[:, 0].sum()))

    # signalPairwiseTest.c:1154 (987 pairs) and the vanilla twin :1287 (999 pairs); ragged (0,0), e = 20
    banded("three_e20_r00", R.THREE_STATE, T_MODEL, ref, tev, anch_t, tp, 0, 20, (0, 0))
    banded("vanilla_e20_r00", R.VANILLA, T_MODEL, ref, tev, anch_t, tp, 0, 20, (0, 0))
    # vanillaAlign's configuration (vanillaAlign.c:202,371-373): e = 50, ragged (1,1); template and complement
    banded("three_e50_r11", R.THREE_STATE, T_MODEL, ref, tev, anch_t, tp, 0, 50, (1, 1))
    banded("three_e50_r11_complement", R.THREE_STATE, C_MODEL, revcomp(ref), cev, anch_c, cp, 1, 50, (1, 1))
    banded("vanilla_e50_r11", R.VANILLA, T_MODEL, ref, tev, anch_t, tp, 0, 50, (1, 1))
    for tag, smt in (("three", R.THREE_STATE), ("vanilla", R.VANILLA)):
        pairs, total = R.align_unbanded(smt, T_MODEL, ref, tev, scale5=tp)
        out[tag + "_unbanded_pairs"] = pairs
        out[tag + "_unbanded_total"] = np.float64(total)
        print(tag, "unbanded", len(pairs), total)
        out[tag + "_expectations_e20_r00"] = R.expectations(smt, T_MODEL, ref, tev, anch_t, scale5=tp)
        out[tag + "_expectations_e50_r11"] = R.expectations(smt, T_MODEL, ref, tev, anch_t, scale5=tp, ragged=(1, 1),
                                                            params=R.default_params(diagonalExpansion=50))
    # the tiny full-matrix known-answer case of signalPairwiseTest.c:580-685 (8 pairs at threshold 0.2)
    tiny_ref = "ACGATACGGACAT"
    tiny_ev = np.array([58.743435, 0.887833, 0.0571, 53.604965, 0.816836, 0.0571, 58.432015, 0.735143, 0.0571,
                        63.684352, 0.795437, 0.0571, 58.921430, 0.812959, 0.0571, 59.895882, 0.740952, 0.0571,
                        61.684303, 0.722332, 0.0571]).reshape(-1, 3)
    pairs, total = R.align_unbanded(R.THREE_STATE, T_MODEL, tiny_ref, tiny_ev, params=R.default_params(threshold=0.2))
    out["tiny_three_pairs"] = pairs
    out["tiny_three_total"] = np.float64(total)
    pairs, total = R.align_unbanded(R.VANILLA, T_MODEL, tiny_ref, tiny_ev, params=R.default_params(threshold=0.5))
    out["tiny_vanilla_pairs"] = pairs
    print("tiny", len(out["tiny_three_pairs"]), len(pairs))
    np.savez_compressed(os.path.join(GOLDEN, "zymo_golden.npz"), **out)
    # the expectation FILE as the reference's own writer lays it out (continuousPairHmm_writeToFile)
    R.write_pair_hmm(out["three_expectations_e20_r00"], os.path.join(GOLDEN, "zymo_three_e20.expectations"))

    # ---- seeded synthetic reads (small, so the committed file stays small) -------------------------------------
    match = synth.load_model_file(T_MODEL)[0]
    syn = {}
    cases = [  # (tag, read index, lX, expansion, ragged, anchor_every, minDiags)
        ("s0", 0, 300, 20, (1, 1), 50, 1000),
        ("s1", 1, 900, 40, (1, 1), 50, 1000),       # one intermediate traceback
        ("s2", 2, 1500, 64, (0, 0), 50, 1000),      # two intermediate tracebacks
        ("s3", 3, 700, 20, (1, 1), 50, 200),        # frequent tracebacks
        ("s4", 4, 500, 30, (0, 1), 500, 1000),      # sparse anchors -> wide band
    ]
    for tag, idx, lX, e, ragged, every, mind in cases:
        r = synth.make_read(match, idx, lX=lX, anchor_every=every)
        p = R.default_params(diagonalExpansion=e, minDiagsBetweenTraceBack=mind)
        pairs, totals = R.align_banded(R.THREE_STATE, T_MODEL, r.ref, r.events, r.anchors, params=p, scale5=r.scale5,
                                       ragged=ragged, want_totals=True)
        syn[tag + "_meta"] = np.array([idx, lX, e, ragged[0], ragged[1], every, mind], dtype=np.int64)
        syn[tag + "_pairs"] = pairs
        syn[tag + "_totals"] = totals
        syn[tag + "_expect"] = R.expectations(R.THREE_STATE, T_MODEL, r.ref, r.events, r.anchors, params=p,
                                              scale5=r.scale5, ragged=ragged)
        print(tag, r.lX, r.lY, len(r.anchors), len(pairs))
    # a split-region case: one anchor gap bigger than splitMatrixBiggerThanThis
    r = synth.make_read(match, 5, lX=1200, anchor_every=50)
    keep = (r.anchors[:, 0] < 300) | (r.anchors[:, 0] > 900)
    p = R.default_params(diagonalExpansion=20, splitMatrixBiggerThanThis=300 * 300)
    pairs, _ = R.align_banded(R.THREE_STATE, T_MODEL, r.ref, r.events, r.anchors[keep], params=p, scale5=r.scale5,
                              ragged=(1, 1))
    syn["split_meta"] = np.array([5, 1200, 20, 1, 1, 50, 1000, 300 * 300], dtype=np.int64)
    syn["split_keep_lo_hi"] = np.array([300, 900], dtype=np.int64)
    syn["split_pairs"] = pairs
    syn["split_points"] = R.split_points(r.anchors[keep], r.lX, r.lY, 300 * 300, 1, 1)
    print("split", len(pairs), syn["split_points"].tolist())
    np.savez_compressed(os.path.join(GOLDEN, "synthetic_golden.npz"), **syn)


def four_state_goldens():
    """The fourState machine (getStateMachine4) on the fixture read: the reference's own known answers are
    988 aligned pairs banded and 988 without banding at ragged (1,1), e = 20 (tests/signalPairwiseTest.c:1199-1236).
    Pins the oracle's four-state restatement ahead of the device path."""
    os.chdir(os.path.join(HERE, "_ref"))
    g = dict(np.load(os.path.join(GOLDEN, "zymo_golden.npz")))
    ref = open(os.path.join(GOLDEN, "ZymoRef.txt")).readline().strip()
    rd = R.load_npread(os.path.join(GOLDEN, "ZymoC_ch_1_file1.npRead"))
    tp, tev, anch = rd["template_params"], rd["template_events"], g["anchors_template"]
    out = {}
    for tag, e, ragged in (("four_e20_r11", 20, (1, 1)), ("four_e20_r00", 20, (0, 0)), ("four_e50_r10", 50, (1, 0))):
        pairs, totals = R.align_banded(R.FOUR_STATE, T_MODEL, ref, tev, anch, params=R.default_params(diagonalExpansion=e),
                                       scale5=tp, ragged=ragged, want_totals=True)
        out[tag + "_pairs"], out[tag + "_totals"] = pairs, totals
        print(tag, len(pairs), int(pairs[:, 0].sum()))
    pairs, total = R.align_unbanded(R.FOUR_STATE, T_MODEL, ref, tev, scale5=tp, ragged=(1, 1))
    out["four_unbanded_r11_pairs"], out["four_unbanded_r11_total"] = pairs, np.float64(total)
    print("four unbanded", len(pairs), total)
    np.savez_compressed(os.path.join(GOLDEN, "zymo_four_state_golden.npz"), **out)

    # The echelon machine (getStateMachineEchelon, 7 states, events covering 1 .. 5 k-mers, Poisson duration term,
    # diagonalCalculationMultiPosteriorMatchProbs): the reference's known answers are 857 pairs banded and 1000 without
    # banding at threshold 0.15, ragged (0,0) (tests/signalPairwiseTest.c:1388-1449).
    out = {}
    for tag, e, ragged, thr in (("echelon_e20_r00_t15", 20, (0, 0), 0.15), ("echelon_e50_r10_t15", 50, (1, 0), 0.15)):
        pairs, totals = R.align_banded(R.ECHELON, T_MODEL, ref, tev, anch, scale5=tp, ragged=ragged, want_totals=True,
                                       params=R.default_params(diagonalExpansion=e, threshold=thr))
        out[tag + "_pairs"], out[tag + "_totals"] = pairs, totals
        print(tag, len(pairs), int(pairs[:, 0].sum()))
    pairs, total = R.align_unbanded(R.ECHELON, T_MODEL, ref, tev, scale5=tp, ragged=(0, 0),
                                    params=R.default_params(threshold=0.15))
    out["echelon_unbanded_r00_t15_pairs"], out["echelon_unbanded_r00_t15_total"] = pairs, np.float64(total)
    print("echelon unbanded", len(pairs), total)
    np.savez_compressed(os.path.join(GOLDEN, "zymo_echelon_golden.npz"), **out)


def vanilla_align_goldens():
    """tests/golden/vanillaAlign/*: outputs of the UNMODIFIED reference CLI (oracle/_ref/vanillaAlign, `make -C oracle
    vanillaAlign lastz`) on the fixture 2D read.  The guide cigar comes from the reference's vendored lastz."""
    import shutil
    import subprocess
    import tempfile
    ref_dir = os.path.join(HERE, "_ref")
    out_dir = os.path.join(GOLDEN, "vanillaAlign")
    os.makedirs(out_dir, exist_ok=True)
    ref = open(os.path.join(GOLDEN, "ZymoRef.txt")).readline().strip()
    read = open(os.path.join(GOLDEN, "ZymoC_ch_1_file1.npRead")).read().split("\n")[1].strip()
    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, "ref.fa"), "w").write(">ZymoRef\n%s\n" % ref)
        open(os.path.join(td, "read.fa"), "w").write(">read\n%s\n" % read)
        cigar = subprocess.run([os.path.join(ref_dir, "cPecanLastz"), "--format=cigar", "ref.fa", "read.fa"], cwd=td,
                               capture_output=True, text=True, check=True).stdout.split("\n")[0] + "\n"
        open(os.path.join(out_dir, "guide.cigar"), "w").write(cigar)
        base = ["-T", T_MODEL, "-C", C_MODEL, "-L", "readA", "-q", os.path.join(GOLDEN, "ZymoC_ch_1_file1.npRead"),
                "-r", os.path.join(GOLDEN, "ZymoRef.txt")]
        for tag, flag in (("s", ["-s"]), ("v", []), ("f", ["-f"]), ("e", ["-e"])):
            tsv = os.path.join(td, "out_%s.tsv" % tag)
            r = subprocess.run([os.path.join(ref_dir, "vanillaAlign")] + flag + base + ["-u", tsv], input=cigar,
                               capture_output=True, text=True, check=True)
            open(os.path.join(out_dir, "stdout_%s.txt" % tag), "w").write(r.stdout)
            shutil.copy(tsv, os.path.join(out_dir, "out_%s.tsv" % tag))
        subprocess.run([os.path.join(ref_dir, "vanillaAlign"), "-s"] + base +
                       ["-t", os.path.join(out_dir, "t_s.exp"), "-c", os.path.join(out_dir, "c_s.exp")], input=cigar,
                       capture_output=True, text=True, check=True)


def descale_first_third(events, scale, shift):
    """nanopore_descaleNanoporeRead as the reference has it (impl/nanopore.c:34-38, 228-236): the loop runs over the flat
    array index in steps of NB_EVENT_PARAMS but stops at the NUMBER of events, so only the means of the first third of
    the events are descaled."""
    ev = np.array(events, dtype=np.float64).reshape(-1, 3).copy()
    k = (len(ev) + 2) // 3
    ev[:k, 0] = (ev[:k, 0] - shift) / scale
    return ev


def hdp_goldens():
    """threeStateHdp from the UNMODIFIED reference with its HDP sources (`make -C oracle refHdp
    vanillaAlignHdp`): the reference's serialised fixture tests/test_hdp/testTemplate.nhdp is committed (gzip) as
    tests/golden/hdp/testTemplate.nhdp.gz; goldens are get_nanopore_kmer_density on it, getAlignedPairsUsingAnchors with
    getHdpStateMachine3 on the fixture read (descaled as vanillaAlign does), and the CLI's own `-d` posterior file."""
    import ctypes as C
    import gzip
    import shutil
    import subprocess
    import tempfile
    from cpecan_signal import hdp as H
    ref_dir = os.path.join(HERE, "_ref")
    out_dir = os.path.join(GOLDEN, "hdp")
    os.makedirs(out_dir, exist_ok=True)
    src = os.path.join(REFERENCE, "tests", "test_hdp", "testTemplate.nhdp")
    with open(src, "rb") as fi, gzip.GzipFile(os.path.join(out_dir, "testTemplate.nhdp.gz"), "wb", mtime=0) as fo:
        shutil.copyfileobj(fi, fo)
    lib = C.CDLL(os.path.join(ref_dir, "libcpecan_ref_hdp.so"))
    lib.ref_hdp_density.restype = C.c_int64
    lib.ref_hdp_align_banded.restype = C.c_int64
    h = H.load_nhdp(src)
    # k-mers: every one with a Dirichlet process of its own that was observed, and every 37th of the rest
    own = [k for k in range(4096) if h.kmer_distr[k] != h.row[-1]]
    ks = np.array(sorted(set(own) | set(range(0, 4096, 37))), dtype=np.int64)
    kmers = "".join("".join("ACGT"[(k >> (2 * (5 - j))) & 3] for j in range(6)) for k in ks)
    xs = np.array([-5.0, 0.0, 0.3, 30.5, 45.123, 50.0, 50.50505050505051, 55.55, 60.606, 65.4321, 70.0, 80.8, 99.5, 100.0,
                   101.010101, 140.0])
    dens = np.zeros((len(ks), len(xs)))
    lib.ref_hdp_density(src.encode(), kmers.encode(), C.c_int64(len(ks)), xs.ctypes.data_as(C.c_void_p), C.c_int64(len(xs)),
                        dens.ctypes.data_as(C.c_void_p))
    out = {"density_kmers": ks, "density_x": xs, "density": dens}
    print("hdp density", dens.shape, float(dens.max()), int((dens == 0).sum()))
    g = dict(np.load(os.path.join(GOLDEN, "zymo_golden.npz")))
    ref = open(os.path.join(GOLDEN, "ZymoRef.txt")).readline().strip()
    rd = R.load_npread(os.path.join(GOLDEN, "ZymoC_ch_1_file1.npRead"))
    tp = rd["template_params"]
    ev = descale_first_third(rd["template_events"], tp[0], tp[1])
    anch = np.ascontiguousarray(g["anchors_template"], dtype=np.int64)
    lX, lY = len(ref) - 5, len(ev)
    for tag, e, ragged, thr in (("hdp_e20_r00", 20, (0, 0), 0.01), ("hdp_e50_r11", 50, (1, 1), 0.01), ("hdp_e20_r10_t30", 20, (1, 0), 0.3)):
        prm = R.default_params(diagonalExpansion=e, threshold=thr)
        cap = 64 * (lX + lY) + 1024
        pairs = np.zeros((cap, 3), dtype=np.int64)
        totals = np.zeros(lX + lY + 1)
        n = lib.ref_hdp_align_banded(src.encode(), ref.encode(), ev.ctypes.data_as(C.c_void_p), C.c_int64(lY),
                                     anch.ctypes.data_as(C.c_void_p), C.c_int64(len(anch)), C.byref(prm), ragged[0], ragged[1],
                                     pairs.ctypes.data_as(C.c_void_p), C.c_int64(cap), totals.ctypes.data_as(C.c_void_p),
                                     C.c_int64(len(totals)))
        assert 0 <= n <= cap
        out[tag + "_pairs"] = pairs[:n].copy()
        out[tag + "_totals"] = totals
        print(tag, n, int(pairs[:n, 0].sum()))
    # getExpectationsUsingAnchors with an HdpHmm: transition sums, likelihood, the event-to-k-mer assignments
    lib.ref_hdp_expectations.restype = C.c_int64
    for tag, e, ragged, thr in (("hdpexp_e50_r11", 50, (1, 1), 0.01), ("hdpexp_e20_r00_t30", 20, (0, 0), 0.3)):
        prm = R.default_params(diagonalExpansion=e, threshold=thr)
        cap = 64 * (lX + lY)
        asg = np.zeros((cap, 2), dtype=np.int64)
        vec = np.zeros(10)
        n = lib.ref_hdp_expectations(src.encode(), ref.encode(), ev.ctypes.data_as(C.c_void_p), C.c_int64(lY),
                                     anch.ctypes.data_as(C.c_void_p), C.c_int64(len(anch)), C.byref(prm), ragged[0], ragged[1],
                                     C.c_double(1e-4), C.c_double(thr), vec.ctypes.data_as(C.c_void_p),
                                     asg.ctypes.data_as(C.c_void_p), C.c_int64(cap))
        assert 0 <= n <= cap
        out[tag + "_vec"] = vec
        out[tag + "_assignments"] = asg[:n].copy()
        print(tag, n, vec)
    np.savez_compressed(os.path.join(out_dir, "zymo_hdp_golden.npz"), **out)
    # the CLI: vanillaAlign -d with the same HDP for both strands
    cigar = open(os.path.join(GOLDEN, "vanillaAlign", "guide.cigar")).read()
    with tempfile.TemporaryDirectory() as td:
        tsv = os.path.join(td, "out_d.tsv")
        r = subprocess.run([os.path.join(ref_dir, "vanillaAlign_hdp"), "-d", "-v", src, "-w", src, "-T", T_MODEL, "-C", C_MODEL,
                            "-L", "readA", "-q", os.path.join(GOLDEN, "ZymoC_ch_1_file1.npRead"),
                            "-r", os.path.join(GOLDEN, "ZymoRef.txt"), "-u", tsv], input=cigar, capture_output=True, text=True,
                           check=True)
        open(os.path.join(GOLDEN, "vanillaAlign", "stdout_d.txt"), "w").write(r.stdout)
        with open(tsv, "rb") as fi, gzip.GzipFile(os.path.join(GOLDEN, "vanillaAlign", "out_d.tsv.gz"), "wb", mtime=0) as fo:
            shutil.copyfileobj(fi, fo)
        print("vanillaAlign -d:", r.stdout.strip())
        # ... and its expectation files (-t / -c): the HdpHmm text format (impl/continuousHmm.c:690-749)
        t_exp, c_exp = os.path.join(GOLDEN, "vanillaAlign", "t_d.exp"), os.path.join(td, "c_d.exp")
        r = subprocess.run([os.path.join(ref_dir, "vanillaAlign_hdp"), "-d", "-v", src, "-w", src, "-T", T_MODEL, "-C", C_MODEL,
                            "-L", "readA", "-q", os.path.join(GOLDEN, "ZymoC_ch_1_file1.npRead"),
                            "-r", os.path.join(GOLDEN, "ZymoRef.txt"), "-t", t_exp, "-c", c_exp], input=cigar,
                           capture_output=True, text=True)
        print("vanillaAlign -d -t -c: rc", r.returncode, r.stderr[-300:])


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "hdp":
        hdp_goldens()
        sys.exit(0)
    main()
    four_state_goldens()
    vanilla_align_goldens()
