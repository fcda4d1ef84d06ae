"""Seeded synthetic nanopore reads of the shape SURVEY.md 8(d) fixes for the C3 / C2 configurations, and the
text-format loaders for the two on-disk inputs of the path (.model, .npRead).

Reference formats restated here: pore model file (impl/stateMachine.c:242-320), .npRead (impl/nanopore.c:40-200).
"""
import os

import numpy as np

PKG_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TEMPLATE_MODEL = os.path.join(PKG_ROOT, "models", "template_median68pA.model")
COMPLEMENT_MODEL = os.path.join(PKG_ROOT, "models", "complement_median68pA_pop2.model")
N_KMERS = 4096
SEED = 20261018
BASES = np.frombuffer(b"ACGT", dtype=np.uint8)


def load_model_file(path):
    """Three whitespace-separated lines: 1 + 4096*5 match parameters (correlation, then level_mean, level_sd,
    noise_mean, noise_sd, noise_lambda per k-mer), 30 skip bins, 1 + 4096*5 extra-event ("gap Y") parameters."""
    with open(path) as fh:
        l1 = np.array(fh.readline().split(), dtype=np.float64)
        l2 = np.array(fh.readline().split(), dtype=np.float64)
        l3 = np.array(fh.readline().split(), dtype=np.float64)
    if l1.size != 1 + N_KMERS * 5 or l2.size != 30 or l3.size != 1 + N_KMERS * 5:
        raise ValueError("%s: not a 6-mer pore model (got %d / %d / %d tokens)" % (path, l1.size, l2.size, l3.size))
    return l1, l2, l3


def load_npread(path):
    """.npRead: line 1 = read length, #template events, #complement events, 5 template + 5 complement scaling
    parameters (13 tokens); line 2 the 2D read; lines 3/4 template event map and events (mean, noise, duration);
    lines 5/6 the same for the complement strand."""
    with open(path) as fh:
        head = fh.readline().split()
        L, nT, nC = int(head[0]), int(head[1]), int(head[2])
        tparams = np.array(head[3:8], dtype=np.float64)
        cparams = np.array(head[8:13], dtype=np.float64)
        twoD = fh.readline().split()[0]
        tmap = np.array(fh.readline().split(), dtype=np.int64)
        tev = np.array(fh.readline().split(), dtype=np.float64).reshape(-1, 3)
        cmap = np.array(fh.readline().split(), dtype=np.int64)
        cev = np.array(fh.readline().split(), dtype=np.float64).reshape(-1, 3)
    if len(tmap) != L or len(cmap) != L or len(tev) != nT or len(cev) != nC:
        raise ValueError("%s: inconsistent npRead" % path)
    return dict(read_length=L, twoD=twoD, template_params=tparams, complement_params=cparams,
                template_map=tmap, template_events=tev, complement_map=cmap, complement_events=cev)


def scale_match_table(match, scale5):
    """emissions_signal_scaleModel (impl/stateMachine.c:631-651): the MATCH table only."""
    scale, shift, var, scale_sd, var_sd = (float(v) for v in scale5)
    t = np.array(match, dtype=np.float64, copy=True)
    m = t[1:].reshape(N_KMERS, 5)
    m[:, 0] = m[:, 0] * scale + shift
    m[:, 1] = m[:, 1] * var
    m[:, 2] = m[:, 2] * scale_sd
    m[:, 4] = m[:, 4] * var_sd
    m[:, 3] = np.sqrt(np.power(m[:, 2], 3.0) / m[:, 4])
    return t


def filter_to_remove_overlap(pairs):
    """filterToRemoveOverlap (impl/pairwiseAligner.c:1160-1200) for lexicographically sorted, duplicate-free pairs:
    keep the pairs that are strictly increasing in x and y both from the back and from the front."""
    pairs = np.asarray(pairs, dtype=np.int64).reshape(-1, 2)
    n = len(pairs)
    keep = np.zeros(n, dtype=bool)
    px = py = np.iinfo(np.int64).max
    for i in range(n - 1, -1, -1):
        x, y = pairs[i]
        if x < px and y < py:
            keep[i] = True
        px, py = min(px, x), min(py, y)
    out = []
    px = py = np.iinfo(np.int64).min
    for i in range(n):
        x, y = pairs[i]
        if x > px and y > py and keep[i]:
            out.append((x, y))
        px, py = max(px, x), max(py, y)
    return np.asarray(out, dtype=np.int64).reshape(-1, 2)


class SyntheticRead:
    __slots__ = ("ref", "events", "anchors", "scale5", "lX", "lY")


def make_read(model_match, read_index, lX=6700, anchor_every=50, seed=SEED, skip_prob=0.10, stay_prob=0.25,
              noise_dist="gauss", ref=None):
    """One synthetic read (counter-based seeding: the read is a function of (seed, read_index) only).

    reference = uniform ACGT of lX+5 nt; walking its k-mers, each is skipped w.p. 0.10, otherwise emits one event
    plus Geometric(0.25) extra 'stay' events; event mean ~ N(mu_k, sd_k), noise ~ N(nu_k, tau_k) truncated > 0,
    duration ~ Exp(0.01), all under per-read scaling (scale, shift, var, scale_sd, var_sd) drawn as SURVEY 8(d) says;
    anchors = the true path sampled every `anchor_every` k-mers, run through filterToRemoveOverlap.
    ref: use this nucleotide string (ACGT only, lX = len - 5) instead of drawing one."""
    rng = np.random.default_rng([seed, read_index])
    codes = rng.integers(0, 4, size=lX + 5)
    if ref is not None:                       # a given reference (e.g. the reverse complement of another read's)
        lX = len(ref) - 5
        codes = np.searchsorted(BASES, np.frombuffer(ref.encode(), dtype=np.uint8))
    kidx = np.zeros(lX, dtype=np.int64)
    for j in range(6):
        kidx = kidx * 4 + codes[j:j + lX]
    scale5 = np.array([rng.uniform(0.95, 1.05), rng.uniform(-5.0, 20.0), rng.uniform(0.9, 1.3),
                       rng.uniform(0.9, 1.3), rng.uniform(0.9, 1.3)])
    scaled = scale_match_table(model_match, scale5)[1:].reshape(N_KMERS, 5)
    skipped = rng.random(lX) < skip_prob
    n_ev = np.where(skipped, 0, 1 + rng.geometric(1.0 - stay_prob, size=lX) - 1)
    owner = np.repeat(np.arange(lX), n_ev)
    lY = len(owner)
    k = kidx[owner]
    mean = rng.normal(scaled[k, 0], scaled[k, 1])
    if noise_dist == "gauss":
        noise = rng.normal(scaled[k, 2], scaled[k, 3])
    else:
        noise = rng.wald(scaled[k, 2], scaled[k, 4])
    noise = np.where(noise > 0.05, noise, 0.05)
    dur = rng.exponential(0.01, size=lY)
    first_event = np.cumsum(n_ev) - n_ev
    xs = np.arange(anchor_every // 2, lX, anchor_every)
    xs = xs[~skipped[xs]]
    anchors = filter_to_remove_overlap(np.stack([xs, first_event[xs]], axis=1)) if len(xs) else np.zeros((0, 2), np.int64)
    r = SyntheticRead()
    r.ref = BASES[codes].tobytes().decode()
    r.events = np.ascontiguousarray(np.stack([mean, noise, dur], axis=1))
    r.anchors = anchors
    r.scale5 = scale5
    r.lX, r.lY = lX, lY
    return r


def make_reads(n, lX=6700, first_index=0, model_path=TEMPLATE_MODEL, **kw):
    match = load_model_file(model_path)[0]
    return [make_read(match, first_index + i, lX=lX, **kw) for i in range(n)]


def reverse_complement(seq):
    return seq[::-1].translate(str.maketrans("ACGT", "TGCA"))
