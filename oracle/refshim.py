"""TEST INFRASTRUCTURE ONLY: ctypes binding of oracle/_ref/libcpecan_ref.so (the unmodified reference sources
compiled by oracle/Makefile) -- used by tests/, oracle/make_golden.py and bench.py's reference/cpu_baseline legs.
Never imported by the product package."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libcpecan_ref.so")

THREE_STATE = 2
VANILLA = 4
FOUR_STATE = 6
ECHELON = 5


class RefParams(C.Structure):
    _fields_ = [("threshold", C.c_double), ("minDiagsBetweenTraceBack", C.c_int64),
                ("traceBackDiagonals", C.c_int64), ("diagonalExpansion", C.c_int64),
                ("constraintDiagonalTrim", C.c_int64), ("splitMatrixBiggerThanThis", C.c_int64)]


def default_params(**kw):
    """pairwiseAlignmentBandingParameters_construct defaults (reference impl/pairwiseAligner.c:1428-1441)."""
    p = RefParams(0.01, 1000, 40, 20, 14, 3000 * 3000)
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def available():
    return os.path.exists(REF_SO)


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(REF_SO)
        _lib.ref_logAdd.restype = C.c_double
        _lib.ref_logAdd.argtypes = [C.c_double, C.c_double]
        for name in ("ref_align_banded", "ref_align_unbanded", "ref_expectations", "ref_split_points",
                     "ref_filter_overlap", "ref_fixture_anchors", "ref_load_npread"):
            getattr(_lib, name).restype = C.c_int64
    return _lib


def _dptr(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64).ctypes.data_as(C.c_void_p)


def _iptr(a):
    return a.ctypes.data_as(C.c_void_p)


def log_add(x, y):
    return lib().ref_logAdd(float(x), float(y))


def band(anchors, lX, lY, expansion):
    anchors = np.ascontiguousarray(np.asarray(anchors, dtype=np.int64).reshape(-1, 2))
    out = np.zeros((lX + lY + 1, 3), dtype=np.int64)
    lib().ref_band(_iptr(anchors), C.c_int64(len(anchors)), C.c_int64(lX), C.c_int64(lY), C.c_int64(expansion),
                   _iptr(out))
    return out


def split_points(anchors, lX, lY, max_matrix, ragged_left, ragged_right):
    anchors = np.ascontiguousarray(np.asarray(anchors, dtype=np.int64).reshape(-1, 2))
    out = np.zeros((len(anchors) + 2, 4), dtype=np.int64)
    n = lib().ref_split_points(_iptr(anchors), C.c_int64(len(anchors)), C.c_int64(lX), C.c_int64(lY),
                               C.c_int64(max_matrix), int(ragged_left), int(ragged_right), _iptr(out),
                               C.c_int64(len(out)))
    return out[:n]


def filter_overlap(pairs):
    pairs = np.ascontiguousarray(np.asarray(pairs, dtype=np.int64).reshape(-1, 2))
    out = np.zeros_like(pairs)
    n = lib().ref_filter_overlap(_iptr(pairs), C.c_int64(len(pairs)), _iptr(out))
    return out[:n]


def align_banded(sm_type, model_file, ref_seq, events, anchors, params=None, scale5=None, strand=0,
                 transitions=None, gap_x=None, ragged=(0, 0), want_totals=False):
    """Returns (pairs[n,3] int64 in the reference's emission order, totals[lX+lY+1] or None)."""
    events = np.ascontiguousarray(events, dtype=np.float64).reshape(-1, 3)
    anchors = np.ascontiguousarray(np.asarray(anchors, dtype=np.int64).reshape(-1, 2))
    params = params or default_params()
    lY = len(events)
    lX = max(len(ref_seq) - 5, 0)
    cap = 64 * (lX + lY) + 1024
    out = np.zeros((cap, 3), dtype=np.int64)
    totals = np.zeros(lX + lY + 1, dtype=np.float64) if want_totals else None
    n = lib().ref_align_banded(int(sm_type), model_file.encode(), _dptr(scale5), int(strand), _dptr(transitions),
                               _dptr(gap_x), ref_seq.encode(), _dptr(events), C.c_int64(lY), _iptr(anchors),
                               C.c_int64(len(anchors)), C.byref(params), int(ragged[0]), int(ragged[1]),
                               _iptr(out), C.c_int64(cap),
                               None if totals is None else _iptr(totals), C.c_int64(0 if totals is None else len(totals)))
    assert 0 <= n <= cap
    return out[:n].copy(), totals


def time_align_banded(sm_type, model_file, ref_seq, events, anchors, params=None, scale5=None, strand=0, ragged=(0, 0)):
    """Seconds spent inside the reference's getAlignedPairsUsingAnchors for one read (CPU baseline)."""
    events = np.ascontiguousarray(events, dtype=np.float64).reshape(-1, 3)
    anchors = np.ascontiguousarray(np.asarray(anchors, dtype=np.int64).reshape(-1, 2))
    params = params or default_params()
    sec = C.c_double(0.0)
    f = lib().ref_time_align_banded
    f.restype = C.c_int64
    n = f(int(sm_type), model_file.encode(), _dptr(scale5), int(strand), ref_seq.encode(), _dptr(events),
          C.c_int64(len(events)), _iptr(anchors), C.c_int64(len(anchors)), C.byref(params), int(ragged[0]),
          int(ragged[1]), C.byref(sec))
    return sec.value, n


def align_unbanded(sm_type, model_file, ref_seq, events, params=None, scale5=None, strand=0, transitions=None,
                   gap_x=None, ragged=(0, 0)):
    events = np.ascontiguousarray(events, dtype=np.float64).reshape(-1, 3)
    params = params or default_params()
    lY = len(events)
    lX = max(len(ref_seq) - 5, 0)
    cap = 64 * (lX + lY) + 1024
    out = np.zeros((cap, 3), dtype=np.int64)
    total = C.c_double(0.0)
    n = lib().ref_align_unbanded(int(sm_type), model_file.encode(), _dptr(scale5), int(strand), _dptr(transitions),
                                 _dptr(gap_x), ref_seq.encode(), _dptr(events), C.c_int64(lY), C.byref(params),
                                 int(ragged[0]), int(ragged[1]), _iptr(out), C.c_int64(cap), C.byref(total))
    assert 0 <= n <= cap
    return out[:n].copy(), total.value


def expectations(sm_type, model_file, ref_seq, events, anchors, params=None, scale5=None, strand=0,
                 transitions=None, gap_x=None, ragged=(0, 0), pseudocount=1e-4):
    events = np.ascontiguousarray(events, dtype=np.float64).reshape(-1, 3)
    anchors = np.ascontiguousarray(np.asarray(anchors, dtype=np.int64).reshape(-1, 2))
    params = params or default_params()
    out = np.zeros(4106 if sm_type == THREE_STATE else 61, dtype=np.float64)
    rc = lib().ref_expectations(int(sm_type), model_file.encode(), _dptr(scale5), int(strand), _dptr(transitions),
                                _dptr(gap_x), ref_seq.encode(), _dptr(events), C.c_int64(len(events)),
                                _iptr(anchors), C.c_int64(len(anchors)), C.byref(params), int(ragged[0]),
                                int(ragged[1]), C.c_double(pseudocount), _iptr(out))
    assert rc == 0
    return out


def load_npread(path):
    dims = np.zeros(3, dtype=np.int64)
    lib().ref_load_npread(path.encode(), _iptr(dims), None, None, None, None, None, None)
    L, nT, nC = (int(v) for v in dims)
    params = np.zeros(10)
    twoD = C.create_string_buffer(L + 8)
    tmap = np.zeros(L, dtype=np.int64)
    cmap = np.zeros(L, dtype=np.int64)
    tev = np.zeros((nT, 3))
    cev = np.zeros((nC, 3))
    lib().ref_load_npread(path.encode(), _iptr(dims), _iptr(params), twoD, _iptr(tmap), _iptr(tev), _iptr(cmap),
                          _iptr(cev))
    return dict(read_length=L, twoD=twoD.raw[:L].decode(), template_params=params[:5].copy(),
                complement_params=params[5:].copy(), template_map=tmap, template_events=tev,
                complement_map=cmap, complement_events=cev)


def fixture_anchors(ref_seq, npread_file, strand=0):
    """Runs the reference's lastz anchoring; must be called with cwd containing ./cPecanLastz."""
    out = np.zeros((1 << 16, 2), dtype=np.int64)
    raw = C.c_int64(0)
    n = lib().ref_fixture_anchors(ref_seq.encode(), npread_file.encode(), int(strand), _iptr(out),
                                  C.c_int64(len(out)), C.byref(raw))
    return out[:n].copy(), raw.value


def write_pair_hmm(exp, path, normalize=False):
    """continuousPairHmm_writeToFile (after continuousPairHmm_normalize when asked) on a 4106-vector."""
    exp = np.ascontiguousarray(exp, dtype=np.float64)
    assert exp.size == 9 + 4096 + 1
    lib().ref_write_pair_hmm(_dptr(exp), int(bool(normalize)), path.encode())


def load_pair_hmm(hmm_path, model_file):
    """hmmContinuous_loadSignalHmm into a strawMan machine: (transitions[9] in StateMachine3 order, gapX[4096])."""
    t = np.zeros(9)
    g = np.zeros(4096)
    lib().ref_load_pair_hmm(hmm_path.encode(), model_file.encode(), _dptr(t), _dptr(g))
    return t, g


def write_vanilla_hmm(bins61, model_file, path, normalize=False):
    """vanillaHmm_writeToFile (after the reference's joint vanillaHmm_normalizeKmerSkipBins when asked) on 60 bins +
    likelihood; the two model lines are those of the vanilla machine built from model_file."""
    bins61 = np.ascontiguousarray(bins61, dtype=np.float64)
    assert bins61.size == 61
    lib().ref_write_vanilla_hmm(_dptr(bins61), int(bool(normalize)), model_file.encode(), path.encode())


def load_vanilla_hmm(hmm_path, model_file):
    """hmmContinuous_loadSignalHmm into a vanilla machine: the 60 skip-bin probabilities the DP will see."""
    b = np.zeros(60)
    lib().ref_load_vanilla_hmm(hmm_path.encode(), model_file.encode(), _dptr(b))
    return b


# ---- threeStateHdp: the reference built WITH its HDP sources (oracle/_ref/libcpecan_ref_hdp.so, `make -C oracle refHdp`)
REF_HDP_SO = os.path.join(HERE, "_ref", "libcpecan_ref_hdp.so")
_hdp_lib = None


def hdp_available():
    return os.path.exists(REF_HDP_SO)


def hdp_lib():
    global _hdp_lib
    if _hdp_lib is None:
        _hdp_lib = C.CDLL(REF_HDP_SO)
        for name in ("ref_hdp_density", "ref_hdp_align_banded", "ref_hdp_expectations"):
            getattr(_hdp_lib, name).restype = C.c_int64
    return _hdp_lib


def hdp_align_banded(nhdp_file, ref_seq, events, anchors, params=None, ragged=(0, 0), want_totals=False):
    """getAlignedPairsUsingAnchors with getHdpStateMachine3(deserialize_nhdp(nhdp_file)) and sequence_getKmer3."""
    events = np.ascontiguousarray(events, dtype=np.float64).reshape(-1, 3)
    anchors = np.ascontiguousarray(np.asarray(anchors, dtype=np.int64).reshape(-1, 2))
    params = params or default_params()
    lY, lX = len(events), max(len(ref_seq) - 5, 0)
    cap = 64 * (lX + lY) + 1024
    out = np.zeros((cap, 3), dtype=np.int64)
    totals = np.zeros(lX + lY + 1, dtype=np.float64) if want_totals else None
    n = hdp_lib().ref_hdp_align_banded(nhdp_file.encode(), ref_seq.encode(), _dptr(events), C.c_int64(lY), _iptr(anchors),
                                       C.c_int64(len(anchors)), C.byref(params), int(ragged[0]), int(ragged[1]), _iptr(out),
                                       C.c_int64(cap), None if totals is None else _iptr(totals),
                                       C.c_int64(0 if totals is None else len(totals)))
    assert 0 <= n <= cap
    return out[:n].copy(), totals


def hdp_expectations(nhdp_file, ref_seq, events, anchors, params=None, ragged=(0, 0), pseudocount=1e-4, threshold=0.01):
    """(9 transition sums + likelihood, assignments[n, 2] = (k-mer position, event index)) of an HdpHmm."""
    events = np.ascontiguousarray(events, dtype=np.float64).reshape(-1, 3)
    anchors = np.ascontiguousarray(np.asarray(anchors, dtype=np.int64).reshape(-1, 2))
    params = params or default_params()
    cap = 64 * (max(len(ref_seq) - 5, 0) + len(events)) + 64
    asg = np.zeros((cap, 2), dtype=np.int64)
    vec = np.zeros(10)
    n = hdp_lib().ref_hdp_expectations(nhdp_file.encode(), ref_seq.encode(), _dptr(events), C.c_int64(len(events)), _iptr(anchors),
                                       C.c_int64(len(anchors)), C.byref(params), int(ragged[0]), int(ragged[1]),
                                       C.c_double(pseudocount), C.c_double(threshold), _iptr(vec), _iptr(asg), C.c_int64(cap))
    assert 0 <= n <= cap
    return vec, asg[:n].copy()
