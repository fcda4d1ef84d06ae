"""Baum-Welch training driver for the signal pair-HMM: the batched, multi-GPU sibling of the reference's
scripts/trainModels.py loop (:244-330) and of its expectation containers (impl/continuousHmm.c; scripts/nanoporeLib.py
:985-1054 ContinuousPairHmm, :1159-1225 ConditionalSignalHmm).

One EM iteration = E-step over this rank's shard of the reads on its GPU (Engine.expectations, summed on device) ->
ONE all-reduce of the expectation vector (4106 doubles three-state, 61 vanilla) across ranks (NCCL over NVLink on GPUs;
gloo in the CPU tests) -> the same normalisation on every rank -> rank 0 writes the .hmm file the way trainModels.py does
(nanoporeLib's writers: Python-2 str() of every float, 12 significant digits) -> every rank reloads it, as the next
round of vanillaAlign does with -y / -z, so the text round trip is part of the M-step.

Where the batch differs from the reference's loop: the reference sums per-read expectation FILES, which vanillaAlign
writes with "%f" (6 decimals, impl/continuousHmm.c:234-271); here the per-read sums never leave the device and are
added in FP64.  `write()` still produces that C format byte for byte (it is what `cpecanAlign -t/-c` emits)."""
import math
import os

import numpy as np

N_KMERS = 4096
N_EXPECT = 9 + N_KMERS + 1
N_BINS = 60
N_EXPECT_VANILLA = N_BINS + 1
MODEL_PARAMS = 5
THREE_STATE = 2
VANILLA = 4
M, X, Y = 0, 1, 2


def py2_str(x):
    """str(float) of the Python 2 the reference's scripts run under: '%.12g', with '.0' appended when the result
    looks like an integer (PyFloat_STR_PRECISION = 12, format_float_short's Py_DTSF_ADD_DOT_0)."""
    x = float(x)
    if math.isnan(x):
        return "nan"
    if math.isinf(x):
        return "inf" if x > 0 else "-inf"
    s = "%.12g" % x
    if "." not in s and "e" not in s:
        s += ".0"
    return s


class ContinuousPairHmm:
    """Expectations of the three-state machine: 3x3 transitions (row-major from*3+to), 4096 k-mer skip counts and the
    likelihood (reference inc/continuousHmm.h:7-25)."""

    def __init__(self, pseudocount=0.0):
        # hmmContinuous_getEmptyHmm(type, pseudocount, ...) puts the pseudocount in every slot (vanillaAlign.c:675-676)
        self.transitions = np.full(9, float(pseudocount))
        self.kmer_skip_probs = np.full(N_KMERS, float(pseudocount))
        self.likelihood = 0.0
        self.running_likelihoods = []

    # -- accumulation (continuousPairHmm_addTo*; nanoporeLib.py add_expectations_file) ---------------------------
    def add_expectations(self, vec):
        vec = np.asarray(vec, dtype=np.float64)
        if vec.shape != (N_EXPECT,):
            raise ValueError("expectation vector must have %d entries" % N_EXPECT)
        if np.isnan(vec[:9]).any():            # the reference writes header-only files for these and skips them
            return False
        self.transitions += vec[:9]
        self.kmer_skip_probs += vec[9:9 + N_KMERS]
        self.likelihood += float(vec[-1])
        return True

    def vector(self):
        return np.concatenate([self.transitions, self.kmer_skip_probs, [self.likelihood]])

    # -- continuousPairHmm_normalize (impl/continuousHmm.c:173-190, hmmDiscrete_normalize2 impl/discreteHmm.c:125-137)
    def normalize(self):
        t = self.transitions.reshape(3, 3)
        self.transitions = (t / t.sum(axis=1, keepdims=True)).reshape(9)
        self.kmer_skip_probs = self.kmer_skip_probs / self.kmer_skip_probs.sum()

    # -- text format of continuousPairHmm_writeToFile (impl/continuousHmm.c:234-271) -------------------------------
    def write(self, path):
        with open(path, "w") as fh:
            fh.write("%d\t%d\t%d\t\n" % (THREE_STATE, 3, N_KMERS))
            if not np.isnan(self.transitions).any():           # hmmContinuous_checkTransitions: header only on NaN
                fh.write("".join("%f\t" % v for v in self.transitions))
                fh.write("%f\n" % self.likelihood)
                fh.write("".join("%f\t" % v for v in self.kmer_skip_probs))
                fh.write("\n")

    # -- text format of nanoporeLib.ContinuousPairHmm.write (scripts/nanoporeLib.py:1027-1054): the trained model -----
    def write_trained(self, path):
        with open(path, "w") as fh:
            fh.write("2\t%d\t%d\n" % (3, N_KMERS))
            fh.write("".join(py2_str(v) + "\t" for v in self.transitions))
            fh.write(py2_str(self.likelihood) + "\n")
            fh.write("".join(py2_str(v) + "\t" for v in self.kmer_skip_probs))
            fh.write("\n")

    @classmethod
    def load(cls, path):
        """continuousPairHmm_loadFromFile (impl/continuousHmm.c:273-370): same checks, same errors (as exceptions)."""
        with open(path) as fh:
            head = fh.readline().split()
            if len(head) < 3:
                raise ValueError("%s: bad header" % path)
            typ, n_states, n_sym = int(head[0]), int(head[1]), int(head[2])
            if typ != THREE_STATE or n_states != 3 or n_sym != N_KMERS:
                raise ValueError("%s: not a three-state 6-mer HMM (type %d, %d states, %d symbols)" % (path, typ, n_states, n_sym))
            line1 = fh.readline().split()
            if len(line1) != 10:
                raise ValueError("Incorrect number of transitions in the input HMM file %s, got %d instead of 10" % (path, len(line1)))
            line2 = fh.readline().split()
            if len(line2) != N_KMERS:
                raise ValueError("Incorrect number of emissions in the input HMM file %s, got %d instead of %d" % (path, len(line2), N_KMERS))
        h = cls()
        h.transitions = np.array(line1[:9], dtype=np.float64)
        h.likelihood = float(line1[9])
        h.kmer_skip_probs = np.array(line2, dtype=np.float64)
        return h

    # -- continuousPairHmm_loadTransitionsAndKmerGapProbs (impl/continuousHmm.c:206-232) ----------------------------
    def state_machine_params(self):
        """(transitions[9] in StateMachine3 field order, gapX[4096] log-probabilities) as the DP will use them."""
        t = self.transitions.reshape(3, 3)
        with np.errstate(divide="ignore"):
            trans = np.array([
                math.log(t[M, M]) if t[M, M] > 0 else -np.inf,      # MATCH_CONTINUE
                math.log(t[X, M]) if t[X, M] > 0 else -np.inf,      # MATCH_FROM_GAP_X
                math.log(t[Y, M]) if t[Y, M] > 0 else -np.inf,      # MATCH_FROM_GAP_Y
                math.log(t[M, X]) if t[M, X] > 0 else -np.inf,      # GAP_OPEN_X
                math.log(t[M, Y]) if t[M, Y] > 0 else -np.inf,      # GAP_OPEN_Y
                math.log(1.0 - t[X, M]) if t[X, M] < 1 else -np.inf,  # GAP_EXTEND_X = log(1 - P(X->M))
                math.log(t[Y, Y]) if t[Y, Y] > 0 else -np.inf,      # GAP_EXTEND_Y
                math.log(t[Y, X]) if t[Y, X] > 0 else -np.inf,      # GAP_SWITCH_TO_X
                -np.inf,                                            # GAP_SWITCH_TO_Y = LOG_ZERO
            ])
            gapx = np.log(self.kmer_skip_probs)
        return trans, gapx


class ConditionalSignalHmm:
    """Expectations of the vanilla machine: 60 skip bins -- beta (M->X) in 0..29, alpha (X->X) in 30..59 -- and the
    likelihood; the two match-model lines ride along unused (reference inc/continuousHmm.h VanillaHmm;
    scripts/nanoporeLib.py:1159-1225)."""

    def __init__(self, pseudocount=0.0, match_model=None, scaled_match_model=None):
        self.kmer_skip_bins = np.full(N_BINS, float(pseudocount))
        n = 1 + N_KMERS * MODEL_PARAMS
        self.match_model = np.zeros(n) if match_model is None else np.asarray(match_model, dtype=np.float64)
        self.scaled_match_model = np.zeros(n) if scaled_match_model is None else np.asarray(scaled_match_model, dtype=np.float64)
        self.likelihood = 0.0
        self.running_likelihoods = []

    def add_expectations(self, vec):
        vec = np.asarray(vec, dtype=np.float64)
        if vec.shape != (N_EXPECT_VANILLA,):
            raise ValueError("expectation vector must have %d entries" % N_EXPECT_VANILLA)
        if np.isnan(vec[:N_BINS]).any():
            return False
        self.kmer_skip_bins += vec[:N_BINS]
        self.likelihood += float(vec[-1])      # nanoporeLib assigns the LAST file's value (:1183); one sum per batch here
        return True

    def vector(self):
        return np.concatenate([self.kmer_skip_bins, [self.likelihood]])

    def normalize(self):
        """nanoporeLib.ConditionalSignalHmm.normalize (:1188-1197): alpha and beta separately -- the C container
        normalises all 60 jointly and says so itself ("this is wrong", impl/continuousHmm.c:424-433); the training
        loop uses the Python one."""
        self.kmer_skip_bins[:30] = self.kmer_skip_bins[:30] / self.kmer_skip_bins[:30].sum()
        self.kmer_skip_bins[30:] = self.kmer_skip_bins[30:] / self.kmer_skip_bins[30:].sum()

    def normalize_joint(self):
        """vanillaHmm_normalizeKmerSkipBins (impl/continuousHmm.c:424-433)."""
        self.kmer_skip_bins = self.kmer_skip_bins / self.kmer_skip_bins.sum()

    # -- vanillaHmm_writeToFile (impl/continuousHmm.c:477-517): what vanillaAlign / cpecanAlign -t/-c emit ---------------
    def write(self, path):
        with open(path, "w") as fh:
            fh.write("%d\t%d\t%d\t\n" % (VANILLA, 3, N_KMERS))
            if not np.isnan(self.kmer_skip_bins).any():
                fh.write("".join("%f\t" % v for v in self.kmer_skip_bins))
                fh.write("%f\n" % self.likelihood)
                fh.write("".join("%f\t" % v for v in self.match_model) + "\n")
                fh.write("".join("%f\t" % v for v in self.scaled_match_model) + "\n")

    # -- nanoporeLib.ConditionalSignalHmm.write (:1199-1225): the trained model ----------------------------------------
    def write_trained(self, path):
        with open(path, "w") as fh:
            fh.write("4\t%d\t%d\n" % (3, N_KMERS))
            fh.write("".join(py2_str(v) + "\t" for v in self.kmer_skip_bins))
            fh.write(py2_str(self.likelihood) + "\n")
            fh.write("".join(py2_str(v) + "\t" for v in self.match_model) + "\n")
            fh.write("".join(py2_str(v) + "\t" for v in self.scaled_match_model) + "\n")

    @classmethod
    def load(cls, path):
        """vanillaHmm_loadFromFile (impl/continuousHmm.c:519-628): type 4, 3 states, 61 numbers on line 1, the two
        model lines of 1 + 4096 * 5 numbers."""
        with open(path) as fh:
            head = fh.readline().split()
            if len(head) < 3:
                raise ValueError("%s: bad header" % path)
            typ, n_states, n_sym = int(float(head[0])), int(float(head[1])), int(float(head[2]))
            if typ != VANILLA or n_states != 3 or n_sym != N_KMERS:
                raise ValueError("%s: not a vanilla 6-mer HMM (type %d, %d states, %d symbols)" % (path, typ, n_states, n_sym))
            line1 = fh.readline().split()
            if len(line1) != N_EXPECT_VANILLA:
                raise ValueError("Incorrect number of skip bins in the input HMM file %s, got %d instead of %d" % (path, len(line1), N_EXPECT_VANILLA))
            n = 1 + N_KMERS * MODEL_PARAMS
            line2, line3 = fh.readline().split(), fh.readline().split()
            if len(line2) != n or len(line3) != n:
                raise ValueError("Incorrect number of match-model parameters in the input HMM file %s" % path)
        h = cls(match_model=np.array(line2, dtype=np.float64), scaled_match_model=np.array(line3, dtype=np.float64))
        h.kmer_skip_bins = np.array(line1[:N_BINS], dtype=np.float64)
        h.likelihood = float(line1[N_BINS])
        return h

    def state_machine_params(self):
        """vanillaHmm_loadKmerSkipBinExpectations (impl/continuousHmm.c:457-466): the 60 bins go into
        EMISSION_GAP_X_PROBS as probabilities (the vanilla cell takes their logs itself); no transitions change."""
        return None, self.kmer_skip_bins.copy()


def cull_training_reads(read_lengths, training_amount, rng=None):
    """cull_training_files (scripts/trainModels.py:78-104) over reads already in memory: shuffle, then take reads until
    their cumulative 2D-read length reaches training_amount (the read that crosses the mark is included).  Returns
    the chosen indices in draw order; every rank that passes the same seeded rng draws the same set."""
    n = len(read_lengths)
    order = np.arange(n)
    (rng or np.random.default_rng()).shuffle(order)
    total, out = 0, []
    for i in order:
        out.append(int(i))
        total += int(read_lengths[i])
        if total >= training_amount:
            break
    return np.array(out, dtype=np.int64)


def shard_by_cells(cells, world):
    """Partition of reads over ranks balanced by band cells: longest-processing-time greedy.
    Returns a list of index arrays (sorted), one per rank; every read appears exactly once."""
    cells = np.asarray(cells, dtype=np.int64)
    order = np.argsort(-cells, kind="stable")
    load = np.zeros(world, dtype=np.int64)
    parts = [[] for _ in range(world)]
    for i in order:
        r = int(np.argmin(load))
        parts[r].append(int(i))
        load[r] += cells[i]
    return [np.array(sorted(p), dtype=np.int64) for p in parts]


def allreduce_sum(vec, group_ok):
    """Sum of a float64 vector over all ranks (in place for torch tensors).  vec: numpy array (CPU, gloo) or a CUDA
    tensor viewing the engine's device accumulator (NCCL)."""
    if not group_ok:
        return vec
    import torch
    import torch.distributed as dist
    if isinstance(vec, np.ndarray):
        if dist.get_backend() == "nccl":                     # NCCL reduces device tensors only: through the GPU and back
            t = torch.from_numpy(vec).cuda()
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            vec[...] = t.cpu().numpy()
            return vec
        t = torch.from_numpy(vec)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return vec
    dist.all_reduce(vec, op=dist.ReduceOp.SUM)
    return vec


class _DevView:
    """__cuda_array_interface__ wrapper so torch can all-reduce the engine's device buffer in place."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


def join_engine_communicator(engine):
    """One NCCL communicator over the engines of all torch.distributed ranks, owned by the C-ABI
    (cpecan_cuda_nccl_init): rank 0 draws the unique id, torch.distributed only carries its 128 bytes."""
    import torch.distributed as dist
    ids = [engine.nccl_unique_id() if dist.get_rank() == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    engine.nccl_init(dist.get_world_size(), dist.get_rank(), ids[0])


def gpu_estep(engine, batch, hmm, params, distributed):
    """E-step of this rank's shard on its GPU; the batch sums are all-reduced IN PLACE in the device accumulator
    (the kernel's atomics and the NCCL collective share one buffer), then fetched once.  The vector has 4106 entries
    for the three-state machine and 61 for the vanilla one (hmm.sm_type).  distributed: True = all-reduce through
    torch.distributed on a view of the device buffer; "cabi" = through the engine's own communicator
    (join_engine_communicator), the path a C driver takes."""
    import torch
    n = N_EXPECT_VANILLA if hmm is not None and int(hmm.sm_type) == VANILLA else N_EXPECT
    engine.stage(batch, hmm=hmm, params=params, mode=1, pair_cap=1)
    engine.run_staged()
    if distributed == "cabi":
        engine.allreduce_expectations()
    elif distributed:
        view = torch.as_tensor(_DevView(engine.expectations_device_ptr(), n), device="cuda")
        allreduce_sum(view, True)
        torch.cuda.synchronize()
    out = np.zeros(n)
    engine.fetch_expectations(out)
    return out


# ------------------------------------------------------------------------------------------------ threeStateHdp
THREE_STATE_HDP = 7


class HdpHmm:
    """Expectations of the HDP machine (reference inc/continuousHmm.h:27-38, impl/continuousHmm.c:630-749): the 3 x 3
    transition sums, the likelihood, and the event-to-k-mer ASSIGNMENTS (event mean, 6-mer) that feed the next round of
    Gibbs sampling.  Across ranks the transition sums are all-reduced like the other machines'; the assignment lists are
    of different lengths per rank and are all-GATHERED in rank order (gather_assignments)."""

    def __init__(self, pseudocount=0.0, threshold=0.01):
        self.transitions = np.full(9, float(pseudocount))
        self.likelihood = 0.0
        self.threshold = float(threshold)
        self.means = np.zeros(0)
        self.kmers = np.zeros(0, dtype=np.int32)             # ACGT 6-mer indices (A0 C1 G2 T3, first base most significant)

    def add(self, vec10, means, kmers):
        """vec10: 9 transition sums + likelihood; means / kmers: this batch's assignments in order."""
        vec10 = np.asarray(vec10, dtype=np.float64)
        if np.isnan(vec10[:9]).any():
            return False
        self.transitions += vec10[:9]
        self.likelihood += float(vec10[9])
        self.means = np.concatenate([self.means, np.asarray(means, dtype=np.float64)])
        self.kmers = np.concatenate([self.kmers, np.asarray(kmers, dtype=np.int32)])
        return True

    @staticmethod
    def kmer_string(k):
        return "".join("ACGT"[(int(k) >> (2 * (5 - j))) & 3] for j in range(6))

    def write(self, path):
        """hdpHmm_writeToFile (impl/continuousHmm.c:704-749), byte for byte: what `vanillaAlign -d -t` emits and
        updateHdpFromAssignments (vanillaAlign.c:142-154) reads back."""
        with open(path, "w") as fh:
            fh.write("%d\t%d\t%f\t%d\t\n" % (THREE_STATE_HDP, 3, self.threshold, len(self.means)))
            if not np.isnan(self.transitions).any():
                fh.write("".join("%f\t" % t for t in self.transitions) + "%f\n" % self.likelihood)
                fh.write("".join("%f\t" % m for m in self.means) + "\n")
                fh.write("".join(self.kmer_string(k) + "\t" for k in self.kmers) + "\n")

    @classmethod
    def load(cls, path):
        import gzip
        op = gzip.open if str(path).endswith(".gz") else open
        with op(path, "rt") as fh:
            head = fh.readline().split()
            if len(head) != 4 or int(head[0]) != THREE_STATE_HDP:
                raise ValueError("not an HdpHmm file: %r" % (head,))
            h = cls(0.0, float(head[2]))
            l1 = fh.readline().split()
            if len(l1) != 10:
                raise ValueError("expected 9 transitions and the likelihood, got %d numbers" % len(l1))
            h.transitions = np.array(l1[:9], dtype=np.float64)
            h.likelihood = float(l1[9])
            h.means = np.array(fh.readline().split(), dtype=np.float64)
            code = {c: i for i, c in enumerate("ACGT")}
            h.kmers = np.array([sum(code[c] << (2 * (5 - j)) for j, c in enumerate(k)) for k in fh.readline().split()], dtype=np.int32)
            if len(h.means) != int(head[3]) or len(h.kmers) != int(head[3]):
                raise ValueError("the file announces %s assignments and holds %d / %d" % (head[3], len(h.means), len(h.kmers)))
        return h


def gather_assignments(means, kmers, distributed):
    """All ranks' assignment lists concatenated in rank order on every rank: the lengths are exchanged first, the lists
    padded to the longest (torch.distributed.all_gather takes equal shapes; gloo on CPU tensors, NCCL on CUDA ones)."""
    means = np.ascontiguousarray(means, dtype=np.float64)
    kmers = np.ascontiguousarray(kmers, dtype=np.int32)
    if not distributed:
        return means, kmers
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    n = torch.tensor([len(means)], dtype=torch.int64, device=dev)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(t.item()) for t in sizes]
    longest = max(1, max(sizes))
    buf = torch.zeros(longest, 2, dtype=torch.float64, device=dev)       # (mean, k-mer index): 4095 is exact in a double
    if len(means):
        buf[:len(means), 0] = torch.from_numpy(means).to(dev)
        buf[:len(means), 1] = torch.from_numpy(kmers.astype(np.float64)).to(dev)
    parts = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    allm = np.concatenate([p[:k, 0].cpu().numpy() for p, k in zip(parts, sizes)])
    allk = np.concatenate([p[:k, 1].cpu().numpy() for p, k in zip(parts, sizes)]).astype(np.int32)
    return allm, allk


def batch_assignments(batch, results, assignments):
    """(event means, k-mer indices) of a batch's assignment triples (Engine.hdp_expectations_batch), item by item."""
    code = np.full(256, -1, dtype=np.int64)
    for i, c in enumerate(b"ACGT"):
        code[c] = i
    means, kmers = [], []
    for i in range(batch.n):
        n, off = int(results[i]["n_pairs"]), int(results[i]["pair_off"])
        if n == 0:
            continue
        t = assignments[off:off + n]
        x = t[:, 1].astype(np.int64) + int(batch.ref_off[i])
        y = t[:, 2].astype(np.int64) + int(batch.ev_off[i])
        k = np.zeros(n, dtype=np.int64)
        for j in range(6):
            k = k * 4 + code[batch.ref[x + j]]
        means.append(batch.events[y, 0])
        kmers.append(k.astype(np.int32))
    if not means:
        return np.zeros(0), np.zeros(0, dtype=np.int32)
    return np.concatenate(means), np.concatenate(kmers)


def gpu_hdp_estep(engine, batch, hmm, params, distributed, container=None):
    """E-step of this rank's shard with the HDP machine: transition sums + likelihood all-reduced (10 doubles), the
    assignments all-gathered; returns the HdpHmm container every rank ends up with."""
    vec, res, asg = engine.hdp_expectations_batch(batch, hmm=hmm, params=params)
    if int((res["status"] & 1).sum()):
        raise ValueError("assignment buffer too small")
    v10 = np.concatenate([vec[:9], vec[-1:]])
    allreduce_sum(v10, bool(distributed))
    means, kmers = gather_assignments(*batch_assignments(batch, res, asg), bool(distributed))
    out = container if container is not None else HdpHmm(0.0, params.threshold)
    out.add(v10, means, kmers)
    return out


def em_iteration(model, estep_sum, n_reads_total, hmm_path, rank=0, pseudocount=1e-4, barrier=None):
    """M-step shared by every rank.  estep_sum: the all-reduced expectation vector of the iteration.  Mirrors
    add_and_norm_expectations (scripts/trainModels.py:126-135): the model object carries its (normalised) values into
    the next iteration's sum, every read contributes the pseudocount of its own expectation file, likelihood restarts
    at zero, then normalise, write, remember the likelihood."""
    model.likelihood = 0.0
    vec = np.array(estep_sum, dtype=np.float64, copy=True)
    vec[:-1] += pseudocount * n_reads_total
    if not model.add_expectations(vec):
        # The device drops a read whose totals are non-finite from the batch sums (status CPECAN_ITEM_NONFINITE), so a
        # NaN here means the sums themselves are broken: normalising "the previous model plus pseudocounts" silently would
        # be a wrong M-step.  The reference skips such expectation files one by one (nanoporeLib.py:1006-1011).
        raise ValueError("E-step sums are not finite: no M-step was done")
    model.normalize()
    if rank == 0:
        tmp = hmm_path + ".tmp"
        model.write_trained(tmp)
        os.replace(tmp, hmm_path)
    if barrier is not None:
        barrier()
    model.running_likelihoods.append(model.likelihood)
    return type(model).load(hmm_path)        # what the next E-step sees: the text file, as with -y / -z
