"""TEST INFRASTRUCTURE ONLY: ctypes binding of oracle/libcpecan_oracle.so (this repo's plain-C restatement of the
reference path, oracle/cpecan_oracle.c).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs import this; the product package never does."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libcpecan_oracle.so")

THREE_STATE = 2
VANILLA = 4
FOUR_STATE = 6
ECHELON = 5
THREE_STATE_HDP = 7
N_KMERS = 4096
MODEL_STRIDE = 5

# stateMachine3_setTransitionsToNanoporeDefaults (reference impl/stateMachine.c:1278-1289), StateMachine3 field order
NANOPORE_TRANSITIONS = np.array([
    -0.23552123624314988,   # MATCH_CONTINUE
    -0.21880828092192281,   # MATCH_FROM_GAP_X
    -0.013406326748077823,  # MATCH_FROM_GAP_Y
    -1.6269694202638481,    # GAP_OPEN_X
    -4.3187242127300092,    # GAP_OPEN_Y
    -1.6269694202638481,    # GAP_EXTEND_X
    -4.3187242127239411,    # GAP_EXTEND_Y
    -np.inf,                # GAP_SWITCH_TO_X
    -np.inf,                # GAP_SWITCH_TO_Y
])
# stateMachine3Vanilla_construct (impl/stateMachine.c:1575-1579) + strand defaults (:1291-1303; note the float literals)
VANILLA_END = (-0.23552123624314988, -1.6269694202638481, -4.3187242127300092)
VANILLA_STRAND = {None: (0.17, float(np.float32(0.55))),
                  0: (float(np.float32(0.17)), float(np.float32(0.55))),
                  1: (float(np.float32(0.14)), float(np.float32(0.49)))}


# stateMachine4_construct's defaults (impl/stateMachine.c:993-1011) in the oracle's T4_* order: MATCH_CONTINUE,
# MATCH_FROM_SHORT_GAP_X, MATCH_FROM_SHORT_GAP_Y, MATCH_FROM_LONG_GAP_X, GAP_SHORT_OPEN_X, GAP_SHORT_EXTEND_X,
# GAP_SHORT_OPEN_Y, GAP_SHORT_EXTEND_Y, GAP_LONG_OPEN_X, GAP_LONG_EXTEND_X, GAP_LONG_SWITCH_TO_X
FOUR_STATE_TRANSITIONS = np.array([
    -0.23552123624314988, -0.21880828092192281, -0.013406326748077823, -5.6732801731704612,
    -1.6269694202638481, -1.6269694202638481, -4.7241893208381773, -4.724189320832104,
    -5.4173365013981227, -0.003442492794189331, -5.4173365013920494])


class OracleModel(C.Structure):
    _fields_ = [("sm_type", C.c_int32), ("strand", C.c_int32), ("match", C.c_void_p), ("gapy", C.c_void_p),
                ("gapx", C.c_void_p), ("trans", C.c_double * 9), ("vanilla", C.c_double * 5),
                ("trans4", C.c_double * 11), ("hdp_density", C.c_void_p), ("hdp_slopes", C.c_void_p),
                ("hdp_kmer_row", C.c_void_p), ("hdp_start", C.c_double), ("hdp_stop", C.c_double), ("hdp_len", C.c_int64)]


class OracleParams(C.Structure):
    _fields_ = [("threshold", C.c_double), ("minDiagsBetweenTraceBack", C.c_int64),
                ("traceBackDiagonals", C.c_int64), ("diagonalExpansion", C.c_int64),
                ("constraintDiagonalTrim", C.c_int64), ("splitMatrixBiggerThanThis", C.c_int64)]


def default_params(**kw):
    p = OracleParams(0.01, 1000, 40, 20, 14, 3000 * 3000)
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def build(force=False):
    if force or not os.path.exists(ORACLE_SO) or \
            os.path.getmtime(ORACLE_SO) < os.path.getmtime(os.path.join(HERE, "cpecan_oracle.c")):
        subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    return ORACLE_SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(ORACLE_SO)
        _lib.oracle_log_add.restype = C.c_double
        _lib.oracle_log_add.argtypes = [C.c_double, C.c_double]
        _lib.oracle_kmer_index.restype = C.c_int32
        for n in ("oracle_align_banded", "oracle_align_unbanded", "oracle_filter_overlap", "oracle_split_points",
                  "oracle_band_cells"):
            getattr(_lib, n).restype = C.c_int64
    return _lib


def load_model_file(path):
    """The .model text format (reference impl/stateMachine.c:242-320): line 1 = 1+4096*5 match params,
    line 2 = 30 skip bins, line 3 = 1+4096*5 gap-Y ('extra event') params."""
    with open(path) as fh:
        l1 = np.array(fh.readline().split(), dtype=np.float64)
        l2 = np.array(fh.readline().split(), dtype=np.float64)
        l3 = np.array(fh.readline().split(), dtype=np.float64)
    assert l1.size == 1 + N_KMERS * 5 and l2.size == 30 and l3.size == 1 + N_KMERS * 5
    return l1, l2, l3


class Model:
    """Host-side bundle of the tables a StateMachine3 / StateMachine3Vanilla carries."""

    def __init__(self, sm_type, model_file=None, tables=None, scale5=None, strand=None, transitions=None, gap_x=None,
                 hdp=None):
        self.sm_type = sm_type
        self.hdp = hdp
        if sm_type == THREE_STATE_HDP:
            # getHdpStateMachine3: the three-state transitions; emissions from the HDP (an object with grid_start,
            # grid_stop, grid_length, density[n, L], slopes[n, L], kmer_distr[4096]: cpecan_signal.hdp.load_nhdp's)
            self.match = self.gapy = np.zeros(1 + N_KMERS * MODEL_STRIDE)
            self.gapx = np.full(N_KMERS, -2.3025850929940455)
            self.trans = np.array(NANOPORE_TRANSITIONS if transitions is None else transitions, dtype=np.float64)
            self.vanilla = np.zeros(5)
            self._hd = np.ascontiguousarray(hdp.density, dtype=np.float64)
            self._hs = np.ascontiguousarray(hdp.slopes, dtype=np.float64)
            self._hk = np.ascontiguousarray(hdp.kmer_distr, dtype=np.int32)
            return
        l1, l2, l3 = tables if tables is not None else load_model_file(model_file)
        self.match = np.ascontiguousarray(l1, dtype=np.float64).copy()
        self.gapy = np.ascontiguousarray(l3, dtype=np.float64).copy()
        if sm_type == THREE_STATE:
            # stateMachine3_construct fills EMISSION_GAP_X_PROBS with log(0.1) (impl/stateMachine.c:1506-1508)
            self.gapx = np.full(N_KMERS, -2.3025850929940455) if gap_x is None else np.array(gap_x, dtype=np.float64)
        elif sm_type == FOUR_STATE:
            # getStateMachine4: emissions_signal_initEmissionsToZero, and the skip-bin line of the model file is not
            # loaded for this type (impl/stateMachine.c:374-386, :282-294)
            self.gapx = np.zeros(N_KMERS) if gap_x is None else np.array(gap_x, dtype=np.float64)
        else:
            # vanilla: 30 bins duplicated into [0..29] (beta) and [30..59] (alpha) (impl/stateMachine.c:282-294)
            self.gapx = np.concatenate([l2, l2]) if gap_x is None else np.array(gap_x, dtype=np.float64)
        self.trans = np.array(NANOPORE_TRANSITIONS if transitions is None else transitions, dtype=np.float64)
        mty, ete = VANILLA_STRAND[strand]
        self.vanilla = np.array([mty, ete, *VANILLA_END], dtype=np.float64)
        if scale5 is not None:
            lib().oracle_scale_model(self.match.ctypes.data_as(C.c_void_p), *[C.c_double(float(v)) for v in scale5])

    def cstruct(self):
        m = OracleModel()
        m.sm_type = self.sm_type
        m.strand = 0
        m.match = self.match.ctypes.data
        m.gapy = self.gapy.ctypes.data
        m.gapx = self.gapx.ctypes.data
        for i in range(9):
            m.trans[i] = self.trans[i]
        for i in range(5):
            m.vanilla[i] = self.vanilla[i]
        for i in range(11):
            m.trans4[i] = FOUR_STATE_TRANSITIONS[i]
        if self.sm_type == THREE_STATE_HDP:
            m.hdp_density = self._hd.ctypes.data
            m.hdp_slopes = self._hs.ctypes.data
            m.hdp_kmer_row = self._hk.ctypes.data
            m.hdp_start, m.hdp_stop, m.hdp_len = float(self.hdp.grid_start), float(self.hdp.grid_stop), int(self.hdp.grid_length)
        return m


def hdp_expectations(model, ref_seq, events, anchors, params=None, ragged=(0, 0), pseudocount=1e-4, threshold=None):
    """threeStateHdp: (vector of 9 transition sums + likelihood, assignments[n, 3] = (from state, k-mer position, event))."""
    events = np.ascontiguousarray(events, dtype=np.float64).reshape(-1, 3)
    anchors = np.ascontiguousarray(np.asarray(anchors, dtype=np.int64).reshape(-1, 2))
    params = params or default_params()
    lX = max(len(ref_seq) - 5, 0)
    expT = np.full(9, pseudocount, dtype=np.float64)
    lik = C.c_double(0.0)
    cap = 64 * (lX + len(events)) + 64
    asg = np.zeros((cap, 3), dtype=np.int64)
    cm = model.cstruct()
    f = lib().oracle_hdp_expectations
    f.restype = C.c_int64
    n = f(C.byref(cm), ref_seq.encode(), C.c_int64(lX), _iptr(events), C.c_int64(len(events)), _iptr(anchors),
          C.c_int64(len(anchors)), C.byref(params), int(ragged[0]), int(ragged[1]),
          C.c_double(params.threshold if threshold is None else threshold), _iptr(expT), C.byref(lik), _iptr(asg), C.c_int64(cap))
    assert 0 <= n <= cap
    return np.concatenate([expT, [lik.value]]), asg[:n].copy()


def hdp_density(model, kmer_index, x):
    """oracle_hdp_density: get_nanopore_kmer_density of an HDP Model for one ACGT k-mer index at x."""
    f = lib().oracle_hdp_density
    f.restype = C.c_double
    cs = model.cstruct()
    return f(C.byref(cs), C.c_int32(int(kmer_index)), C.c_double(float(x)))


def _iptr(a):
    return a.ctypes.data_as(C.c_void_p)


def log_add(x, y):
    return lib().oracle_log_add(float(x), float(y))


def band(anchors, lX, lY, expansion):
    anchors = np.ascontiguousarray(np.asarray(anchors, dtype=np.int64).reshape(-1, 2))
    xl = np.zeros(lX + lY + 1, dtype=np.int64)
    xr = np.zeros(lX + lY + 1, dtype=np.int64)
    lib().oracle_band(_iptr(anchors), C.c_int64(len(anchors)), C.c_int64(lX), C.c_int64(lY), C.c_int64(expansion),
                      _iptr(xl), _iptr(xr))
    return np.stack([np.arange(lX + lY + 1, dtype=np.int64), xl, xr], axis=1)


def filter_overlap(pairs):
    pairs = np.ascontiguousarray(np.asarray(pairs, dtype=np.int64).reshape(-1, 2))
    out = np.zeros_like(pairs)
    n = lib().oracle_filter_overlap(_iptr(pairs), C.c_int64(len(pairs)), _iptr(out))
    return out[:n]


def split_points(anchors, lX, lY, max_matrix, ragged_left, ragged_right):
    anchors = np.ascontiguousarray(np.asarray(anchors, dtype=np.int64).reshape(-1, 2))
    out = np.zeros((len(anchors) + 2, 4), dtype=np.int64)
    n = lib().oracle_split_points(_iptr(anchors), C.c_int64(len(anchors)), C.c_int64(lX), C.c_int64(lY),
                                  C.c_int64(max_matrix), int(ragged_left), int(ragged_right), _iptr(out))
    return out[:n]


def band_cells(anchors, lX, lY, params=None, ragged=(0, 0)):
    anchors = np.ascontiguousarray(np.asarray(anchors, dtype=np.int64).reshape(-1, 2))
    params = params or default_params()
    return lib().oracle_band_cells(_iptr(anchors), C.c_int64(len(anchors)), C.c_int64(lX), C.c_int64(lY),
                                   C.byref(params), int(ragged[0]), int(ragged[1]))


def debug_terms(n):
    """Arms capture of the two terms of diagonalCalculationTotalProbability; returns the (t1, t2) arrays."""
    t1 = np.full(n, np.nan)
    t2 = np.full(n, np.nan)
    lib().oracle_debug_terms(_iptr(t1), _iptr(t2))
    return t1, t2


def debug_terms_off():
    lib().oracle_debug_terms(None, None)


def align_banded(model, ref_seq, events, anchors, params=None, ragged=(0, 0), want_totals=False):
    events = np.ascontiguousarray(events, dtype=np.float64).reshape(-1, 3)
    anchors = np.ascontiguousarray(np.asarray(anchors, dtype=np.int64).reshape(-1, 2))
    params = params or default_params()
    lY = len(events)
    lX = max(len(ref_seq) - 5, 0)
    cap = 64 * (lX + lY) + 1024
    out = np.zeros((cap, 3), dtype=np.int64)
    totals = np.zeros(lX + lY + 1, dtype=np.float64) if want_totals else None
    cm = model.cstruct()
    n = lib().oracle_align_banded(C.byref(cm), ref_seq.encode(), C.c_int64(lX), _iptr(events), C.c_int64(lY),
                                  _iptr(anchors), C.c_int64(len(anchors)), C.byref(params), int(ragged[0]),
                                  int(ragged[1]), _iptr(out), C.c_int64(cap),
                                  None if totals is None else _iptr(totals))
    assert 0 <= n <= cap
    return out[:n].copy(), totals


def align_unbanded(model, ref_seq, events, params=None, ragged=(0, 0)):
    events = np.ascontiguousarray(events, dtype=np.float64).reshape(-1, 3)
    params = params or default_params()
    lY = len(events)
    lX = max(len(ref_seq) - 5, 0)
    cap = 64 * (lX + lY) + 1024
    out = np.zeros((cap, 3), dtype=np.int64)
    total = C.c_double(0.0)
    cm = model.cstruct()
    n = lib().oracle_align_unbanded(C.byref(cm), ref_seq.encode(), C.c_int64(lX), _iptr(events), C.c_int64(lY),
                                    C.byref(params), int(ragged[0]), int(ragged[1]), _iptr(out), C.c_int64(cap),
                                    C.byref(total))
    assert 0 <= n <= cap
    return out[:n].copy(), total.value


def expectations(model, ref_seq, events, anchors, params=None, ragged=(0, 0), pseudocount=1e-4):
    """Returns the reference's expectation vector layout: threeState 9 + 4096 + likelihood; vanilla 60 + likelihood."""
    events = np.ascontiguousarray(events, dtype=np.float64).reshape(-1, 3)
    anchors = np.ascontiguousarray(np.asarray(anchors, dtype=np.int64).reshape(-1, 2))
    params = params or default_params()
    lX = max(len(ref_seq) - 5, 0)
    nT = 9 if model.sm_type == THREE_STATE else 60
    expT = np.full(nT, pseudocount, dtype=np.float64)
    expS = np.full(N_KMERS, pseudocount, dtype=np.float64)
    lik = C.c_double(0.0)
    cm = model.cstruct()
    lib().oracle_expectations(C.byref(cm), ref_seq.encode(), C.c_int64(lX), _iptr(events), C.c_int64(len(events)),
                              _iptr(anchors), C.c_int64(len(anchors)), C.byref(params), int(ragged[0]),
                              int(ragged[1]), _iptr(expT), _iptr(expS), C.byref(lik))
    if model.sm_type == THREE_STATE:
        return np.concatenate([expT, expS, [lik.value]])
    return np.concatenate([expT, [lik.value]])
