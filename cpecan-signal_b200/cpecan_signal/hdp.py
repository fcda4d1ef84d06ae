"""Reader of a serialised NanoporeHDP (the reference's .nhdp text format: serialize_nhdp, impl/nanopore_hdp.c:834-843, and
serialize_hdp, impl/hdp.c:2876-3007) -- the part the threeStateHdp machine reads: the sampling grid and, per observed
Dirichlet process, the posterior predictive density with its spline slopes (what get_nanopore_kmer_density evaluates,
impl/nanopore_hdp.c:390-392).  The data, the Gibbs sampler's state and the factor tree are read past: sampling and
building HDPs are out of scope."""
import gzip
from dataclasses import dataclass

import numpy as np

KMER_LENGTH = 6


@dataclass
class NanoporeHdp:
    alphabet: str                # sorted, as package_nanopore_hdp keeps it (impl/nanopore_hdp.c:38-55)
    kmer_length: int
    grid_start: float
    grid_stop: float
    grid_length: int
    density: np.ndarray          # [n_distr, grid_length]: posterior predictive of each observed process
    slopes: np.ndarray           # [n_distr, grid_length]: its spline slopes
    kmer_distr: np.ndarray       # [4096] int32: row read by each ACGT 6-mer (nearest observed ancestor), -1 = none
    parent: np.ndarray           # [num_dps] int64, -1 for the base process
    row: np.ndarray              # [num_dps] int32: row of each process or -1 (not observed)


def load_nhdp(path):
    op = gzip.open if str(path).endswith(".gz") else open
    with op(path, "rt") as fh:
        nxt = lambda: fh.readline().rstrip("\n")
        alphabet_size = int(nxt())
        alphabet = "".join(sorted(nxt().split()[0]))
        if len(alphabet) != alphabet_size:
            raise ValueError("alphabet %r does not have %d characters" % (alphabet, alphabet_size))
        kmer_length = int(nxt())
        splines_finalized, has_data, sample_gamma, num_dps = (int(nxt()) for _ in range(4))
        if not (splines_finalized and has_data):
            raise ValueError("the HDP holds no finalized distributions (has_data=%d, splines_finalized=%d)" % (has_data, splines_finalized))
        nxt()                                                    # data
        data_dp = np.array(nxt().split(), dtype=np.int64)
        nxt()                                                    # base parameters
        g = nxt().split()
        grid_start, grid_stop, grid_length = float(g[0]), float(g[1]), int(g[2])
        nxt()                                                    # gamma
        if sample_gamma:
            for _ in range(4):
                nxt()
        parent = np.full(num_dps, -1, dtype=np.int64)
        for i in range(num_dps):
            f = nxt().split("\t")
            if f[0] != "-":
                parent[i] = int(f[0])
        # mark_observed_dps (impl/hdp.c:1124-1152): processes holding data and all their ancestors
        observed = np.zeros(num_dps, dtype=bool)
        for i in np.unique(data_dp):
            while i >= 0 and not observed[i]:
                observed[i] = True
                i = parent[i]
        row = np.full(num_dps, -1, dtype=np.int32)
        dens = []
        for i in range(num_dps):
            line = nxt()
            if observed[i]:
                v = np.array(line.split(), dtype=np.float64)
                if v.size != grid_length:
                    raise ValueError("density of process %d has %d of %d grid points" % (i, v.size, grid_length))
                row[i] = len(dens)
                dens.append(v)
        slopes = np.zeros((len(dens), grid_length))
        for i in range(num_dps):
            line = nxt()
            if row[i] >= 0:
                v = np.array(line.split(), dtype=np.float64)
                if v.size != grid_length:
                    raise ValueError("slopes of process %d have %d of %d grid points" % (i, v.size, grid_length))
                slopes[row[i]] = v
    # kmer_id (impl/nanopore_hdp.c:348-380): big-endian word over the sorted alphabet; leaf process id = k-mer id
    kmer_distr = np.full(4 ** KMER_LENGTH, -1, dtype=np.int32)
    pos = {c: alphabet.find(c) for c in "ACGT"}
    if kmer_length == KMER_LENGTH and all(p >= 0 for p in pos.values()):
        digit = np.array([pos[c] for c in "ACGT"], dtype=np.int64)
        k = np.arange(4 ** KMER_LENGTH)
        ids = np.zeros_like(k)
        for j in range(KMER_LENGTH):
            ids = ids * alphabet_size + digit[(k >> (2 * (KMER_LENGTH - 1 - j))) & 3]
        for kk, i in zip(k, ids):
            while i >= 0 and row[i] < 0:
                i = parent[i]
            kmer_distr[kk] = row[i] if i >= 0 else -1
    return NanoporeHdp(alphabet, kmer_length, grid_start, grid_stop, grid_length, np.array(dens), slopes, kmer_distr, parent, row)
