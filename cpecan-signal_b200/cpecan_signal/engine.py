"""ctypes binding of the C-ABI in include/cpecan_cuda.h (libcpecan_cuda.so, built by __graft_entry__.build()).

There is no CPU path here: if the shared library or a CUDA device is missing, construction raises."""
import ctypes as C
import os

import numpy as np

PKG_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.environ.get("CPECAN_LIB") or os.path.join(PKG_ROOT, "libcpecan_cuda.so")   # CPECAN_LIB: developer A/B builds

SM_THREE_STATE = 2
SM_VANILLA = 4
SM_ECHELON = 5
SM_FOUR_STATE = 6
SM_THREE_STATE_HDP = 7
MODE_POSTERIOR = 0
MODE_EXPECTATION = 1
MODE_UNBANDED = 2
N_KMERS = 4096

ITEM_PAIR_OVERFLOW = 1
ITEM_NONFINITE = 2
ITEM_BAND_STEP = 4

# stateMachine3_setTransitionsToNanoporeDefaults (reference impl/stateMachine.c:1278-1289), StateMachine3 field order
NANOPORE_TRANSITIONS = (-0.23552123624314988, -0.21880828092192281, -0.013406326748077823, -1.6269694202638481,
                        -4.3187242127300092, -1.6269694202638481, -4.3187242127239411, -np.inf, -np.inf)


class Params(C.Structure):
    """PairwiseAlignmentParameters (reference inc/pairwiseAligner.h:80-91)."""
    _fields_ = [("threshold", C.c_double), ("minDiagsBetweenTraceBack", C.c_int64),
                ("traceBackDiagonals", C.c_int64), ("diagonalExpansion", C.c_int64),
                ("constraintDiagonalTrim", C.c_int64), ("splitMatrixBiggerThanThis", C.c_int64)]


def default_params(**kw):
    """pairwiseAlignmentBandingParameters_construct (reference impl/pairwiseAligner.c:1428-1441)."""
    p = Params(0.01, 1000, 40, 20, 14, 3000 * 3000)
    for k, v in kw.items():
        setattr(p, k, v)
    return p


class Hmm(C.Structure):
    _fields_ = [("sm_type", C.c_int32), ("reserved", C.c_int32), ("transitions", C.c_double * 9),
                ("vanilla", C.c_double * 5), ("four_state", C.c_double * 11)]


# stateMachine4_construct's defaults (reference impl/stateMachine.c:993-1011) in cpecan_hmm.four_state order
FOUR_STATE_TRANSITIONS = (-0.23552123624314988, -0.21880828092192281, -0.013406326748077823, -5.6732801731704612,
                          -1.6269694202638481, -1.6269694202638481, -4.7241893208381773, -4.724189320832104,
                          -5.4173365013981227, -0.003442492794189331, -5.4173365013920494)


def four_state_hmm(transitions=None):
    """StateMachine4 (getStateMachine4): model with 4096 gap-X log-probabilities (zeros by default)."""
    h = Hmm()
    h.sm_type = SM_FOUR_STATE
    t = FOUR_STATE_TRANSITIONS if transitions is None else transitions
    for i in range(11):
        h.four_state[i] = float(t[i])
    return h


def echelon_hmm():
    """StateMachineEchelon (getStateMachineEchelon): model with the 60 skip bins of the vanilla machine."""
    h = Hmm()
    h.sm_type = SM_ECHELON
    return h


def three_state_hmm(transitions=None):
    h = Hmm()
    h.sm_type = SM_THREE_STATE
    t = NANOPORE_TRANSITIONS if transitions is None else transitions
    for i in range(9):
        h.transitions[i] = float(t[i])
    return h


def hdp_hmm(transitions=None):
    """StateMachine3_HDP (getHdpStateMachine3): the three-state transitions; the items' models come from upload_hdp."""
    h = three_state_hmm(transitions)
    h.sm_type = SM_THREE_STATE_HDP
    return h


# stateMachine3Vanilla_construct (reference impl/stateMachine.c:1575-1579) and the strand defaults (:1291-1303, float
# literals there): TRANSITION_M_TO_Y_NOT_X, TRANSITION_E_TO_E, then the three DEFAULT_END_* log-probabilities
VANILLA_END = (-0.23552123624314988, -1.6269694202638481, -4.3187242127300092)
VANILLA_STRAND = {None: (0.17, float(np.float32(0.55))),
                  "template": (float(np.float32(0.17)), float(np.float32(0.55))),
                  "complement": (float(np.float32(0.14)), float(np.float32(0.49)))}


def vanilla_hmm(strand=None, m_to_y_not_x=None, e_to_e=None):
    """StateMachine3Vanilla: strand = None (constructor values), "template" or "complement"
    (stateMachine3Vanilla_setStrandTransitionsToDefaults)."""
    h = Hmm()
    h.sm_type = SM_VANILLA
    mty, ete = VANILLA_STRAND[strand]
    h.vanilla[0] = mty if m_to_y_not_x is None else float(m_to_y_not_x)
    h.vanilla[1] = ete if e_to_e is None else float(e_to_e)
    for i in range(3):
        h.vanilla[2 + i] = VANILLA_END[i]
    return h


def vanilla_gapx(skip_bins):
    """The 30 skip bins of a .model file as EMISSION_GAP_X_PROBS holds them for the vanilla machine: beta in [0, 30),
    alpha in [30, 60) (emissions_signal_loadPoreModel, impl/stateMachine.c:282-294)."""
    b = np.ascontiguousarray(skip_bins, dtype=np.float64)
    assert b.size == 30
    return np.concatenate([b, b])


class Batch(C.Structure):
    _fields_ = [("n_items", C.c_int64), ("ref", C.c_void_p), ("ref_off", C.c_void_p), ("events", C.c_void_p),
                ("ev_off", C.c_void_p), ("anchors", C.c_void_p), ("anchor_off", C.c_void_p),
                ("model_id", C.c_void_p), ("scale", C.c_void_p), ("ragged", C.c_void_p)]


class Result(C.Structure):
    _fields_ = [("n_pairs", C.c_int64), ("pair_off", C.c_int64), ("band_cells", C.c_int64),
                ("total_logprob", C.c_double), ("status", C.c_int32), ("n_tracebacks", C.c_int32)]


RESULT_DTYPE = np.dtype([("n_pairs", np.int64), ("pair_off", np.int64), ("band_cells", np.int64),
                         ("total_logprob", np.float64), ("status", np.int32), ("n_tracebacks", np.int32)])


class Timing(C.Structure):
    _fields_ = [("h2d_ms", C.c_double), ("prep_ms", C.c_double), ("plan_ms", C.c_double), ("align_ms", C.c_double),
                ("d2h_ms", C.c_double), ("total_ms", C.c_double), ("kernel_launches", C.c_int64),
                ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64), ("band_cells", C.c_int64),
                ("warps_per_item", C.c_int32), ("ctas", C.c_int32)]


class EngineError(RuntimeError):
    pass


def load_library():
    if not os.path.exists(LIB_PATH):
        raise EngineError("%s is missing: run __graft_entry__.build() (there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    lib.cpecan_cuda_last_error.restype = C.c_char_p
    lib.cpecan_cuda_last_error.argtypes = [C.c_void_p]
    lib.cpecan_cuda_destroy.argtypes = [C.c_void_p]
    lib.cpecan_cuda_destroy.restype = None
    return lib


class HostBatch:
    """Flat (structure-of-arrays) host batch in the layout cpecan_batch describes."""

    def __init__(self, refs, events, anchors, model_ids=None, scales=None, ragged=None):
        n = len(refs)
        self.n = n
        ref_len = np.fromiter((len(r) for r in refs), dtype=np.int64, count=n)
        self.ref_off = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(ref_len, out=self.ref_off[1:])
        self.ref = np.frombuffer("".join(refs).encode(), dtype=np.uint8).copy() if n else np.zeros(1, np.uint8)
        evs = [np.ascontiguousarray(e, dtype=np.float64).reshape(-1, 3) for e in events]
        self.ev_off = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(np.fromiter((len(e) for e in evs), dtype=np.int64, count=n), out=self.ev_off[1:])
        self.events = np.ascontiguousarray(np.concatenate(evs, axis=0)) if n and self.ev_off[-1] else np.zeros((1, 3))
        ans = [np.asarray(a, dtype=np.int64).reshape(-1, 2) for a in anchors]
        self.anchor_off = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(np.fromiter((len(a) for a in ans), dtype=np.int64, count=n), out=self.anchor_off[1:])
        self.anchors = np.ascontiguousarray(np.concatenate(ans, axis=0)) if n and self.anchor_off[-1] else np.zeros((1, 2), np.int64)
        self.model_id = np.zeros(n, dtype=np.int32) if model_ids is None else np.ascontiguousarray(model_ids, dtype=np.int32)
        self.scale = None if scales is None else np.ascontiguousarray(scales, dtype=np.float64).reshape(n, 5)
        self.ragged = np.zeros(n, dtype=np.uint8) if ragged is None else np.ascontiguousarray(
            [(int(a) & 1) | ((int(b) & 1) << 1) for a, b in ragged], dtype=np.uint8)

    def view(self, i0, i1):
        """Items [i0, i1) as a batch of their own WITHOUT copying the big arrays (views keep page-locked memory
        page-locked); only the small offset arrays are rebased."""
        v = HostBatch.__new__(HostBatch)
        v.n = int(i1 - i0)
        v.ref_off = np.ascontiguousarray(self.ref_off[i0:i1 + 1] - self.ref_off[i0])
        v.ev_off = np.ascontiguousarray(self.ev_off[i0:i1 + 1] - self.ev_off[i0])
        v.anchor_off = np.ascontiguousarray(self.anchor_off[i0:i1 + 1] - self.anchor_off[i0])
        v.ref = self.ref[int(self.ref_off[i0]):max(int(self.ref_off[i1]), int(self.ref_off[i0]) + 1)]
        v.events = self.events[int(self.ev_off[i0]):max(int(self.ev_off[i1]), int(self.ev_off[i0]) + 1)]
        v.anchors = self.anchors[int(self.anchor_off[i0]):max(int(self.anchor_off[i1]), int(self.anchor_off[i0]) + 1)]
        v.model_id = self.model_id[i0:i1]
        v.scale = None if self.scale is None else self.scale[i0:i1]
        v.ragged = self.ragged[i0:i1]
        return v

    @property
    def lX(self):
        return np.maximum(np.diff(self.ref_off) - 5, 0)

    @property
    def lY(self):
        return np.diff(self.ev_off)

    def cstruct(self):
        b = Batch()
        b.n_items = self.n
        b.ref = self.ref.ctypes.data
        b.ref_off = self.ref_off.ctypes.data
        b.events = self.events.ctypes.data
        b.ev_off = self.ev_off.ctypes.data
        b.anchors = self.anchors.ctypes.data
        b.anchor_off = self.anchor_off.ctypes.data
        b.model_id = self.model_id.ctypes.data
        b.scale = None if self.scale is None else self.scale.ctypes.data
        b.ragged = self.ragged.ctypes.data
        return b


class Engine:
    def __init__(self, device=0):
        self.lib = load_library()
        self.ctx = C.c_void_p()
        rc = self.lib.cpecan_cuda_init(int(device), C.byref(self.ctx))
        if rc != 0:
            raise EngineError("cpecan_cuda_init failed (rc=%d): no usable CUDA device; there is no CPU fallback" % rc)

    def close(self):
        if self.ctx:
            self.lib.cpecan_cuda_host_free.argtypes = [C.c_void_p, C.c_void_p]
            for p in getattr(self, "_pinned", []):
                self.lib.cpecan_cuda_host_free(self.ctx, p)
            self._pinned = []
            self.lib.cpecan_cuda_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise EngineError("%s failed (rc=%d): %s" % (what, rc, self.lib.cpecan_cuda_last_error(self.ctx).decode()))

    def upload_model(self, match, gapy, gapx):
        match = np.ascontiguousarray(match, dtype=np.float64)
        gapy = np.ascontiguousarray(gapy, dtype=np.float64)
        gapx = np.ascontiguousarray(gapx, dtype=np.float64)
        assert match.size == 1 + N_KMERS * 5 and gapy.size == 1 + N_KMERS * 5
        mid = C.c_int32(-1)
        self._check(self.lib.cpecan_cuda_upload_model(self.ctx, match.ctypes.data_as(C.c_void_p),
                                                      gapy.ctypes.data_as(C.c_void_p), gapx.ctypes.data_as(C.c_void_p),
                                                      C.c_int32(gapx.size), C.byref(mid)), "upload_model")
        return mid.value

    def upload_hdp(self, hdp):
        """A NanoporeHdp (cpecan_signal.hdp.load_nhdp) for the threeStateHdp machine; the id is released with release_model."""
        dens = np.ascontiguousarray(hdp.density, dtype=np.float64)
        slopes = np.ascontiguousarray(hdp.slopes, dtype=np.float64)
        kd = np.ascontiguousarray(hdp.kmer_distr, dtype=np.int32)
        assert dens.shape == slopes.shape == (dens.shape[0], hdp.grid_length) and kd.size == N_KMERS
        if hdp.kmer_length != 6:
            raise EngineError("the HDP is over %d-mers, the alignment over 6-mers" % hdp.kmer_length)
        mid = C.c_int32(-1)
        self._check(self.lib.cpecan_cuda_upload_hdp(self.ctx, C.c_double(hdp.grid_start), C.c_double(hdp.grid_stop),
                                                    C.c_int64(hdp.grid_length), C.c_int32(dens.shape[0]),
                                                    dens.ctypes.data_as(C.c_void_p), slopes.ctypes.data_as(C.c_void_p),
                                                    kd.ctypes.data_as(C.c_void_p), C.byref(mid)), "upload_hdp")
        return mid.value

    def update_model(self, model_id, match=None, gapy=None, gapx=None):
        arrs = [None if a is None else np.ascontiguousarray(a, dtype=np.float64) for a in (match, gapy, gapx)]
        ptrs = [None if a is None else a.ctypes.data_as(C.c_void_p) for a in arrs]
        self._check(self.lib.cpecan_cuda_update_model(self.ctx, C.c_int32(model_id), *ptrs), "update_model")

    def release_model(self, model_id):
        """Frees the device tables of an uploaded model (its id may be handed out again)."""
        self._check(self.lib.cpecan_cuda_release_model(self.ctx, C.c_int32(model_id)), "release_model")

    def set_resident_warps(self, warps_per_sm):
        """At most this many resident alignment warps per SM for the batches staged from now on (0 = all that fit)."""
        self._check(self.lib.cpecan_cuda_set_resident_warps(self.ctx, C.c_int32(int(warps_per_sm))), "set_resident_warps")

    def set_exact_arithmetic(self, on=True):
        """threeState / vanilla posterior batches staged from now on run on the FP64 kernel in the reference's own operation
        order (scores equal to the last digit) instead of the FP32 one."""
        self._check(self.lib.cpecan_cuda_set_exact_arithmetic(self.ctx, C.c_int32(1 if on else 0)), "set_exact_arithmetic")

    def device_info(self):
        sm, clk, mem = C.c_int32(), C.c_int32(), C.c_int64()
        self._check(self.lib.cpecan_cuda_device_info(self.ctx, C.byref(sm), C.byref(clk), C.byref(mem)), "device_info")
        return dict(sm_count=sm.value, clock_khz=clk.value, hbm_bytes=mem.value)

    def timing(self):
        t = Timing()
        self._check(self.lib.cpecan_cuda_get_timing(self.ctx, C.byref(t)), "get_timing")
        return {k: getattr(t, k) for k, _ in Timing._fields_}

    @staticmethod
    def default_pair_capacity(batch, per_event=4):
        return int(per_event * int(batch.ev_off[-1]) + 64 * batch.n + 1024)

    def pinned_empty(self, shape, dtype):
        """numpy array over page-locked host memory (cpecan_cuda_host_alloc); freed with the engine."""
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) if np.ndim(shape) else int(shape)
        nbytes = max(1, n * dtype.itemsize)
        self.lib.cpecan_cuda_host_alloc.restype = C.c_void_p
        p = self.lib.cpecan_cuda_host_alloc(self.ctx, C.c_int64(nbytes))
        if not p:
            raise EngineError("cpecan_cuda_host_alloc(%d) failed" % nbytes)
        self._pinned = getattr(self, "_pinned", [])
        self._pinned.append(p)
        buf = (C.c_char * nbytes).from_address(p)
        return np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)

    def pin_batch(self, batch):
        """Moves a HostBatch's arrays into page-locked memory (in place)."""
        for name in ("ref", "ref_off", "events", "ev_off", "anchors", "anchor_off", "model_id", "scale", "ragged"):
            a = getattr(batch, name)
            if a is None:
                continue
            b = self.pinned_empty(a.shape, a.dtype)
            b[...] = a
            setattr(batch, name, b)
        return batch

    def align_batch(self, batch, hmm=None, params=None, mode=MODE_POSTERIOR, pair_cap=None, want_totals=False,
                    out=None):
        """Returns (results structured array, pairs int32[n,3], totals list or None).  Pairs of item i are
        pairs[r['pair_off'] : r['pair_off'] + min(r['n_pairs'], cap_i)] in the reference's emission order.
        out = (results, pairs) lets the caller supply (e.g. page-locked) output buffers."""
        hmm = hmm or three_state_hmm()
        params = params or default_params()
        if out is not None:
            results, pairs = out
            pair_cap = len(pairs)
        else:
            pair_cap = pair_cap or self.default_pair_capacity(batch)
            pairs = np.zeros((pair_cap, 3), dtype=np.int32)
            results = np.zeros(batch.n, dtype=RESULT_DTYPE)
        totals = tot_off = None
        if want_totals:
            tot_off = np.zeros(batch.n + 1, dtype=np.int64)
            np.cumsum(3 * (batch.lX + batch.lY + 1), out=tot_off[1:])
            totals = np.full(int(tot_off[-1]), np.nan)
        cb = batch.cstruct()
        rc = self.lib.cpecan_cuda_align_batch(
            self.ctx, C.byref(hmm), C.byref(params), C.c_int32(mode), C.byref(cb), pairs.ctypes.data_as(C.c_void_p),
            C.c_int64(pair_cap), results.ctypes.data_as(C.c_void_p),
            None if totals is None else totals.ctypes.data_as(C.c_void_p),
            None if tot_off is None else tot_off.ctypes.data_as(C.c_void_p))
        self._check(rc, "align_batch")
        tl = None
        if want_totals:
            # per item: rows = (total handed to the diagonal, term 1, term 2); want_totals="terms" keeps all three
            tl = [totals[tot_off[i]:tot_off[i + 1]].reshape(3, -1) for i in range(batch.n)]
            if want_totals != "terms":
                tl = [t[0] for t in tl]
        return results, pairs, tl

    N_EXPECT = 9 + N_KMERS + 1
    N_EXPECT_VANILLA = 60 + 1

    def expectations_batch(self, batch, hmm=None, params=None, out=None, pseudocount=0.0):
        """getExpectationsUsingAnchors over a batch (reference impl/pairwiseAligner.c:1571-1591), summed on device.
        Returns (expectations[9 + 4096 + 1], results): transitions row-major from*3+to, k-mer skip counts, likelihood;
        the sums are ADDED to `out` (or to a fresh vector pre-filled with `pseudocount` in the count slots)."""
        hmm = hmm or three_state_hmm()
        params = params or default_params()
        if out is None:
            out = np.full(self.N_EXPECT_VANILLA if hmm.sm_type == SM_VANILLA else self.N_EXPECT, float(pseudocount),
                          dtype=np.float64)
            out[-1] = 0.0
        results = np.zeros(batch.n, dtype=RESULT_DTYPE)
        cb = batch.cstruct()
        self._check(self.lib.cpecan_cuda_expectations_batch(self.ctx, C.byref(hmm), C.byref(params), C.byref(cb),
                                                            out.ctypes.data_as(C.c_void_p),
                                                            results.ctypes.data_as(C.c_void_p)), "expectations_batch")
        return out, results

    def hdp_expectations_batch(self, batch, hmm=None, params=None, out=None, pseudocount=0.0, assignment_cap=None):
        """getExpectationsUsingAnchors with an HdpHmm over a batch: (expectations[9 + 4096 + 1] with the transition sums in
        [0, 9) and the likelihood last, results, assignments int32 [cap, 3] = (from state, k-mer position, event index);
        item i owns assignments[results[i].pair_off : + n_pairs] (item_pairs())."""
        hmm = hmm or hdp_hmm()
        params = params or default_params()
        if out is None:
            out = np.zeros(self.N_EXPECT, dtype=np.float64)
            out[:9] = float(pseudocount)
        cap = int(assignment_cap if assignment_cap is not None else 48 * (int(batch.ev_off[-1]) + 32 * batch.n) + 1024)
        asg = np.zeros((cap, 3), dtype=np.int32)
        results = np.zeros(batch.n, dtype=RESULT_DTYPE)
        cb = batch.cstruct()
        self._check(self.lib.cpecan_cuda_hdp_expectations_batch(self.ctx, C.byref(hmm), C.byref(params), C.byref(cb),
                                                                out.ctypes.data_as(C.c_void_p), asg.ctypes.data_as(C.c_void_p),
                                                                C.c_int64(cap), results.ctypes.data_as(C.c_void_p)),
                    "hdp_expectations_batch")
        return out, results, asg

    def expectations_device_ptr(self):
        p = C.c_void_p()
        self._check(self.lib.cpecan_cuda_expectations_device_ptr(self.ctx, C.byref(p)), "expectations_device_ptr")
        return p.value

    # multi-GPU training through the C-ABI's own NCCL communicator (include/cpecan_cuda.h)
    NCCL_ID_BYTES = 128

    def nccl_unique_id(self):
        buf = C.create_string_buffer(self.NCCL_ID_BYTES)
        self._check(self.lib.cpecan_cuda_nccl_unique_id(buf), "nccl_unique_id")
        return bytes(buf.raw)

    def nccl_init(self, n_ranks, rank, unique_id):
        assert len(unique_id) == self.NCCL_ID_BYTES
        self._check(self.lib.cpecan_cuda_nccl_init(self.ctx, C.c_int32(n_ranks), C.c_int32(rank), C.c_char_p(unique_id)), "nccl_init")

    def allreduce_expectations(self):
        """In-place ncclAllReduce(sum, fp64) of the device accumulator over the communicator of nccl_init()."""
        self._check(self.lib.cpecan_cuda_allreduce_expectations(self.ctx), "allreduce_expectations")

    def fetch_expectations(self, out):
        self._check(self.lib.cpecan_cuda_fetch_expectations(self.ctx, out.ctypes.data_as(C.c_void_p)), "fetch_expectations")
        return out

    # device-resident variant (kernel-only timing)
    def stage(self, batch, hmm=None, params=None, mode=MODE_POSTERIOR, pair_cap=None):
        hmm = hmm or three_state_hmm()
        params = params or default_params()
        pair_cap = pair_cap or self.default_pair_capacity(batch)
        cb = batch.cstruct()
        self._staged_n = batch.n
        self._staged_cap = pair_cap
        self._check(self.lib.cpecan_cuda_stage(self.ctx, C.byref(hmm), C.byref(params), C.c_int32(mode), C.byref(cb),
                                               C.c_int64(pair_cap)), "stage")

    def run_staged(self):
        self._check(self.lib.cpecan_cuda_run_staged(self.ctx), "run_staged")

    def restage_model(self, hmm):
        self._check(self.lib.cpecan_cuda_restage_model(self.ctx, C.byref(hmm)), "restage_model")

    def run_staged_async(self):
        self._check(self.lib.cpecan_cuda_run_staged_async(self.ctx), "run_staged_async")

    def wait(self):
        self._check(self.lib.cpecan_cuda_wait(self.ctx), "wait")

    def fetch_staged(self):
        pairs = np.zeros((self._staged_cap, 3), dtype=np.int32)
        results = np.zeros(self._staged_n, dtype=RESULT_DTYPE)
        self._check(self.lib.cpecan_cuda_fetch_staged(self.ctx, pairs.ctypes.data_as(C.c_void_p),
                                                      results.ctypes.data_as(C.c_void_p)), "fetch_staged")
        return results, pairs


def item_pairs(results, pairs, i, cap=None):
    """The pairs of item i.  An item that overflowed its share of the pair buffer (status bit ITEM_PAIR_OVERFLOW) has
    only the pairs that fitted in the buffer: its slice ends where the next item's begins."""
    r = results[i]
    n = int(r["n_pairs"])
    if int(r["status"]) & ITEM_PAIR_OVERFLOW:
        end = int(results[i + 1]["pair_off"]) if i + 1 < len(results) else None
        if end is None:
            raise EngineError("item %d overflowed its pair buffer (%d pairs): re-run with a larger pair_cap" % (i, n))
        n = end - int(r["pair_off"])
    return pairs[int(r["pair_off"]): int(r["pair_off"]) + n]
