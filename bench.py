#!/usr/bin/env python
"""bench.py -- banded signal pair-HMM forward/backward/posterior throughput (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--reads-per-gpu B] [--impl reference]

Workload (configuration C3): seeded synthetic reads, lX ~ 6700 reference k-mers, lY ~ 8000 events,
6-mer template model, anchors every 50 k-mers, three-state machine, threshold 0.01, ragged ends (1,1); the batch is
split evenly over the three band expansions e = 64 / 128 / 256 ("anchored band width 64-256").  One STEP = one pass of
the hot path over the whole batch of B reads per GPU (three kernel launches, one per expansion).  Reads are sharded
across ranks with no data-path collective (weak scaling: B reads per GPU).

  value      GCUPS = 2 * band cells / time / 1e9 with the batch resident in HBM (kernels only)
  e2e        the same metric through the C-ABI call with HOST buffers: H2D of the batch, preparation, plan, kernels
             and D2H of the aligned pairs all inside the timed region; the batch is streamed in sub-batches through
             three contexts per expansion so that one context's copies run under the others' kernels
  roofline   the binding bound: the FP32 issue-rate roofline SURVEY.md 8(d) defines (165 issue-ops per band cell against
             SMs x 128 lanes x clock); roofline_hbm: 25 algorithmic bytes per band cell / kernel time against
             MEASURED_PEAKS.json, with the DRAM traffic ncu measured per expansion (profiles/r2_dram_bytes_per_cell.json)
  cpu_baseline / --impl reference
             the reference's own C code (oracle/_ref, unmodified sources; else the oracle port) on all host cores,
             one process per read as the reference's drivers do, on a bounded sample of the same reads.
"""
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "cpecan-signal_b200"))

EXPANSIONS = (64, 128, 256)
LX = 6700
HBM_BYTES_PER_CELL = 25.0       # forward cell written + read once (24 B) + inputs/outputs (<1 B)
ISSUE_OPS_PER_CELL = 165.0
METRIC = "banded_fwd_bwd_posterior_gcups"


# ----------------------------------------------------------------------------------------------- synthetic input
def _gen_chunk(args):
    first, count, lX = args
    from cpecan_signal import synth
    match = synth.load_model_file(synth.TEMPLATE_MODEL)[0]
    return [synth.make_read(match, first + i, lX=lX) for i in range(count)]


def generate_reads(n, first_index, lX=LX, procs=None):
    procs = procs or min(32, os.cpu_count() or 1)
    chunk = max(1, (n + procs * 4 - 1) // (procs * 4))
    jobs = [(first_index + s, min(chunk, n - s), lX) for s in range(0, n, chunk)]
    if procs == 1 or n <= 8:
        parts = [_gen_chunk(j) for j in jobs]
    else:
        with mp.get_context("fork").Pool(procs) as pool:
            parts = pool.map(_gen_chunk, jobs)
    return [r for p in parts for r in p]


# ----------------------------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.samples = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); smax.append(float(f[1])); power.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(smax)) if smax else None,
                "power_w_max": float(max(power)) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------- CPU reference arm
def _mapped_cpu_library(kind):
    """The library that ran the timed call, as this worker process has it mapped (/proc/self/maps): makes "kind" provable
    from the output.  (libcpecan_oracle.so may be mapped as well: the parent counted the band cells with it before the
    fork; the timed call of kind "reference" goes through oracle/refshim.py into libcpecan_ref.so only.)"""
    want = "libcpecan_ref" if kind == "reference" else "libcpecan_oracle"
    try:
        with open("/proc/self/maps") as fh:
            libs = sorted({line.split()[-1] for line in fh if want in line})
        return [os.path.relpath(p, ROOT) for p in libs]
    except OSError:
        return []


def _cpu_one(args):
    """One read through the reference's getAlignedPairsUsingAnchors (process-per-read, as scripts/signalAlign.py)."""
    kind, read, e = args
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    if kind == "reference":
        import refshim as R
        from cpecan_signal import synth
        sec, n = R.time_align_banded(R.THREE_STATE, synth.TEMPLATE_MODEL, read.ref, read.events, read.anchors,
                                     params=R.default_params(diagonalExpansion=e), scale5=read.scale5, ragged=(1, 1))
    else:
        import oracleshim as O
        from cpecan_signal import synth
        m = O.Model(O.THREE_STATE, model_file=synth.TEMPLATE_MODEL, scale5=read.scale5)
        t0 = time.perf_counter()
        pairs, _ = O.align_banded(m, read.ref, read.events, read.anchors,
                                  params=O.default_params(diagonalExpansion=e), ragged=(1, 1))
        sec, n = time.perf_counter() - t0, len(pairs)
    return sec, n, _mapped_cpu_library(kind)


def cpu_kind():
    return "reference" if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libcpecan_ref.so")) else "port"


def cpu_cells(reads, exps):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracleshim as O
    return sum(O.band_cells(r.anchors, r.lX, r.lY, O.default_params(diagonalExpansion=e), (1, 1))
               for r, e in zip(reads, exps))


def run_cpu_sample(reads, exps, cores):
    """All `cores` host cores, one process per read.  Returns (wall seconds, band cells, summed core-seconds)."""
    kind = cpu_kind()
    cells = cpu_cells(reads, exps)
    jobs = [(kind, r, e) for r, e in zip(reads, exps)]
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(cores) as pool:
        res = pool.map(_cpu_one, jobs, chunksize=1)
    wall = time.perf_counter() - t0
    run_cpu_sample.libraries = sorted({lib for r in res for lib in r[2]})       # what the worker processes had mapped
    return wall, cells, sum(r[0] for r in res)


def reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path on the host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_step = max(3, min(3 * cores, 192))
    per_step -= per_step % 3
    reads = generate_reads(per_step, 10_000_000, procs=min(cores, 32))
    exps = [EXPANSIONS[i % 3] for i in range(per_step)]
    for _ in range(args.warmup if args.warmup < 2 else 1):       # one warm pass is enough for a CPU code path
        run_cpu_sample(reads[:min(len(reads), cores)], exps[:min(len(reads), cores)], cores)
    walls, cells = [], 0
    for _ in range(args.steps):
        wall, c, _ = run_cpu_sample(reads, exps, cores)
        walls.append(wall); cells = c
    t = float(np.mean(walls))
    gcups = 2.0 * cells / t / 1e9
    line = {"metric": METRIC, "value": gcups, "unit": "GCUPS", "impl": "reference", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "reads_per_s": len(reads) / t, "band_cells_per_s": cells / t,
            "config": {"workload": "C3 synthetic reads lX~6700 lY~8000, expansions 64/128/256 evenly, three-state, "
                                   "bounded sample of %d reads per step" % len(reads), "reads_per_step": len(reads)},
            "cpu_baseline": {"value": gcups, "unit": "GCUPS", "cores": cores, "kind": cpu_kind(),
                             "library_timed": getattr(run_cpu_sample, "libraries", []),
                             "sample": "%d reads per step, one process per read on %d cores" % (len(reads), cores)},
            "e2e": {"value": gcups, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- EM (config 5)
def em_arm(args):
    """BASELINE config 5: Baum-Welch iterations over a resident batch.  One step = E-step of this rank's reads
    (expectation kernels, sums in a device buffer) + ONE NCCL all-reduce of the 4106 doubles in place + the M-step
    (normalise, rank 0 writes the reference-format .hmm, every rank reloads it) + re-derivation of the per-column
    parameters on device.  Reads are sharded over ranks by band cells (weak scaling: B reads per GPU)."""
    import tempfile
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from cpecan_signal import Engine, HostBatch, default_params, em, synth, three_state_hmm
    B = min(args.reads_per_gpu, 125000)             # config 5: 1M reads over 8 GPUs = 125k per GPU (81 GB resident at e = 64)
    e = 64
    n_unique = min(B, args.unique_reads)
    reads = generate_reads(n_unique, 50_000_000 + rank * 1_000_000, procs=max(1, min(32, (os.cpu_count() or 1) // max(1, world))))
    reads = [reads[i % n_unique] for i in range(B)]
    l1, _, l3 = synth.load_model_file(synth.TEMPLATE_MODEL)
    eng = Engine(local)
    mid = eng.upload_model(l1, l3, np.full(4096, -2.3025850929940455))
    hb = HostBatch([r.ref for r in reads], [r.events for r in reads], [r.anchors for r in reads],
                   model_ids=[mid] * B, scales=[r.scale5 for r in reads], ragged=[(1, 1)] * B)
    params = default_params(diagonalExpansion=e)
    hmm = three_state_hmm()
    eng.stage(hb, hmm=hmm, params=params, mode=1, pair_cap=1)
    cells_rank = eng.timing()["band_cells"]
    if world > 1:
        em.join_engine_communicator(eng)          # the C-ABI's own NCCL communicator; torch carries only the id
    model = em.ContinuousPairHmm()
    td = tempfile.mkdtemp()
    path = [os.path.join(td, "template_trained.hmm")]
    if world > 1:
        dist.broadcast_object_list(path, src=0)
    hmm_path = path[0]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    estep_ms, allreduce_ms = [], []

    def step():
        nonlocal hmm
        eng.run_staged()
        estep_ms.append(eng.timing()["align_ms"])
        if world > 1:
            t_ar = time.perf_counter()
            eng.allreduce_expectations()              # ncclAllReduce(sum, fp64) of the 4106 doubles, in place, blocking
            allreduce_ms.append((time.perf_counter() - t_ar) * 1e3)
        vec = np.zeros(em.N_EXPECT)
        eng.fetch_expectations(vec)
        loaded = em.em_iteration(model, vec, B * world, hmm_path, rank=rank, barrier=barrier if world > 1 else None)
        trans, gapx = loaded.state_machine_params()
        eng.update_model(mid, gapx=gapx)
        hmm = three_state_hmm(trans)
        eng.restage_model(hmm)

    for _ in range(args.warmup):
        step()
    barrier()
    estep_ms.clear(); allreduce_ms.clear()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    barrier()
    wall = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([wall, float(cells_rank)], dtype=torch.float64, device="cuda")
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        wall, cells_total = float(tmax[0]), float(t[1])
    else:
        cells_total = float(cells_rank)
    if rank == 0:
        print(json.dumps({
            "metric": "em_iteration_gcups", "value": 2.0 * cells_total * args.steps / wall / 1e9, "unit": "GCUPS",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "reads_per_s": B * world * args.steps / wall,
            "config": {"workload": "C5: Baum-Welch iteration (E-step kernels + all-reduce of 4106 doubles + M-step), "
                                   "synthetic reads lX~6700 x lY~8000, e=64, three-state", "reads_per_gpu": B,
                       "distinct_reads_per_gpu": n_unique, "reads_total": B * world,
                       "allreduce_bytes": em.N_EXPECT * 8,
                       "collective": "ncclAllReduce(sum, fp64) in place on the device accumulator through the C-ABI "
                                     "(cpecan_cuda_allreduce_expectations)" if world > 1 else "none (1 rank)"},
            "estep_kernel_ms_rank0": float(np.mean(estep_ms)), "estep_gcups_rank0": 2.0 * cells_rank / float(np.mean(estep_ms)) / 1e6,
            "allreduce_ms_rank0": float(np.mean(allreduce_ms)) if allreduce_ms else None,
            "likelihoods": [float(v) for v in model.running_likelihoods[-3:]]}), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------- GPU arm
def main():
    # NCCL writes "NCCL version ..." to STDOUT at NCCL_DEBUG=VERSION (some pods export that): stdout is for the one JSON line
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--reads-per-gpu", type=int, default=99999,
                    help="reads per GPU and step; the default is BASELINE config 3's 100k-read batch (divisible by the 3 expansions)")
    ap.add_argument("--unique-reads", type=int, default=12288,
                    help="distinct synthetic reads generated per rank (2.7 ms of host time each); the batch tiles them, every "
                         "copy with its own host and device buffers")
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--e2e-subbatches", type=int, default=9,
                    help="end-to-end leg: the reads of one expansion are streamed in this many sub-batches through "
                         "--e2e-contexts contexts, so that one's copies run under the others' kernels")
    ap.add_argument("--e2e-contexts", type=int, default=3, help="contexts per expansion in the end-to-end leg (sub-batches alternate between them)")
    ap.add_argument("--e2e-resident-warps", type=int, default=8, help="resident alignment warps per SM and context in the end-to-end leg")
    ap.add_argument("--pairs-per-event", type=float, default=2.0, help="capacity of the aligned-pair buffers (the batch yields ~1.1)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (default, the driver's contract): --reads-per-gpu reads on EVERY GPU; strong: ONE batch of "
                         "--reads-per-gpu reads (BASELINE config 4: 'the same batch sharded across 2/4/8') dealt to the ranks "
                         "by band cells, longest first (em.shard_by_cells)")
    ap.add_argument("--machine", default="three", choices=["three", "vanilla"],
                    help="state machine of the posterior workload (the headline is the three-state machine)")
    ap.add_argument("--arithmetic", default="fp32", choices=["fp32", "exact"],
                    help="exact: cpecan_cuda_set_exact_arithmetic -- the FP64 kernel in the reference's own operation order "
                         "(use a smaller --reads-per-gpu: it runs at 5-8 %% of the FP32 rate)")
    ap.add_argument("--workload", default="posterior", choices=["posterior", "em"],
                    help="posterior = BASELINE config 3/4 (default, the headline); em = config 5, one Baum-Welch "
                         "iteration per step (E-step kernels + NCCL all-reduce + M-step)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    if args.impl == "reference":
        reference_arm(args)
        return
    if args.workload == "em":
        em_arm(args)
        return

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from cpecan_signal import Engine, HostBatch, default_params
    from cpecan_signal.engine import RESULT_DTYPE

    from cpecan_signal import em as _em
    B = args.reads_per_gpu - args.reads_per_gpu % 3
    n_unique = min(B, max(3, args.unique_reads - args.unique_reads % 3))
    if args.scaling == "strong" and world > 1:
        # ONE batch for the whole job: every rank generates the same distinct reads (seed 0), read i of the batch is
        # distinct read i % n_unique with expansion EXPANSIONS[i % 3]; the ranks take the shards of a longest-first
        # deal by band cells ((lX + lY) * (e + 40) is proportional to them for these reads)
        reads = generate_reads(n_unique, 0, procs=max(1, min(32, (os.cpu_count() or 1) // max(1, world))))
        weight = np.array([(reads[i % n_unique].lX + reads[i % n_unique].lY) * (EXPANSIONS[i % 3] + 40) for i in range(B)])
        mine = _em.shard_by_cells(weight, world)[rank]
        by_exp = [[reads[i % n_unique] for i in mine if i % 3 == j] for j in range(3)]
        B = len(mine)
    else:
        reads = generate_reads(n_unique, rank * 1_000_000, procs=max(1, min(32, (os.cpu_count() or 1) // max(1, world))))
        reads = [reads[i % n_unique] for i in range(B)]          # n_unique is a multiple of 3: a read keeps its expansion
        by_exp = [reads[j::3] for j in range(3)]
    engines, batches, outs = [], [], []
    l1 = l3 = None
    from cpecan_signal import synth
    l1, l2, l3 = synth.load_model_file(synth.TEMPLATE_MODEL)
    tbl_match, tbl_gapy = l1, l3
    from cpecan_signal import vanilla_gapx, vanilla_hmm
    vanilla = args.machine == "vanilla"
    gapx_tbl = vanilla_gapx(l2) if vanilla else np.full(4096, -2.3025850929940455)
    bench_hmm = vanilla_hmm("template") if vanilla else None
    if vanilla:
        # the vanilla machine reports up to 1.95 aligned pairs per event on these reads (three-state: 1.1)
        args.pairs_per_event = max(args.pairs_per_event, 3.0)
    order = sorted(range(3), key=lambda j: -EXPANSIONS[j])      # widest band first: its tail is filled by the others
    SUB = max(1, args.e2e_subbatches)
    host = Engine(local)                    # owns the page-locked host memory of both legs (stages nothing itself)
    for j in order:
        e = EXPANSIONS[j]
        sub = by_exp[j]
        eng = Engine(local)
        eng.set_exact_arithmetic(args.arithmetic == "exact")
        mid = eng.upload_model(l1, l3, gapx_tbl)
        hb = HostBatch([r.ref for r in sub], [r.events for r in sub], [r.anchors for r in sub],
                       model_ids=[mid] * len(sub), scales=[r.scale5 for r in sub], ragged=[(1, 1)] * len(sub))
        host.pin_batch(hb)
        # ~1.1 aligned pairs per event at threshold 0.01; room for the per-sub-batch rounding of the end-to-end leg
        cap = eng.default_pair_capacity(hb, per_event=args.pairs_per_event) + 2048 * SUB
        outs.append((host.pinned_empty(hb.n, RESULT_DTYPE), host.pinned_empty((cap, 3), np.int32)))
        engines.append(eng); batches.append(hb)
    params = [default_params(diagonalExpansion=EXPANSIONS[j]) for j in order]
    _free, _tot = torch.cuda.mem_get_info()
    print("rank %d: %d reads (%d distinct) pinned; device memory before staging: %.1f of %.1f GB free" % (rank, B, n_unique, _free / 1e9, _tot / 1e9), file=sys.stderr, flush=True)
    # host memory: at 8 ranks per box keep only the pinned batches and the small sample the CPU baseline needs
    n_keep = max(3, min(3 * (os.cpu_count() or 1), 192, B))
    n_keep -= n_keep % 3
    reads = [by_exp[i % 3][i // 3] for i in range(min(n_keep, 3 * min(len(b) for b in by_exp)))]
    del by_exp
    import gc
    gc.collect()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- resident (kernel-only) measurement -----------------------------------------------------------------
    for eng, hb, p, o in zip(engines, batches, params, outs):
        eng.stage(hb, hmm=bench_hmm, params=p, pair_cap=len(o[1]))
    cells_by_engine = [eng.timing()["band_cells"] for eng in engines]
    cells_rank = sum(cells_by_engine)
    _free, _tot = torch.cuda.mem_get_info()
    print("rank %d: batch staged; device memory: %.1f of %.1f GB free" % (rank, _free / 1e9, _tot / 1e9), file=sys.stderr, flush=True)
    def resident_step():
        """One pass of the hot path over the whole resident batch: the three expansions' kernels are enqueued
        together (each context has its own streams) so that the GPU stays full through their tails."""
        for eng in engines:
            eng.run_staged_async()
        for eng in engines:
            eng.wait()
        return max(eng.timing()["align_ms"] for eng in engines), sum(eng.timing()["kernel_launches"] for eng in engines)

    for _ in range(args.warmup):
        resident_step()
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    t0 = time.perf_counter()
    kern_ms = 0.0
    launches = 0
    l0 = sum(eng.timing()["kernel_launches"] for eng in engines)
    for _ in range(args.steps):
        span_ms, l1 = resident_step()
        kern_ms += span_ms                # CUDA-event span of the step's k_align2 launches (they start together)
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    launches = sum_over_ranks(float(l1 - l0))            # whole job, like value
    wall = max_over_ranks(wall)
    kern_ms_max = max_over_ranks(kern_ms)
    cells_total = sum_over_ranks(float(cells_rank))
    reads_total = sum_over_ranks(float(B))
    gcups = 2.0 * cells_total * args.steps / wall / 1e9

    # sanity: results of the last resident pass
    n_pairs = 0
    for eng in engines:
        res, _ = eng.fetch_staged()
        assert int((res["status"] != 0).sum()) == 0, "non-zero item status"
        n_pairs += int(res["n_pairs"].sum())

    info = engines[0].device_info()

    # ---- end-to-end through the C-ABI with host buffers ---------------------------------------------------
    # The same batch, streamed: the reads of an expansion go through the C-ABI in SUB sub-batches, in turn through
    # --e2e-contexts contexts, one blocking align_batch call each (H2D of the pinned host sub-batch, staging kernels, plan,
    # alignment kernels, D2H of the aligned pairs) issued from one host thread per context.  ctypes drops the GIL, so a
    # context's copies run under the other contexts' kernels; each context keeps half the resident warps
    # (cpecan_cuda_set_resident_warps) so that the forward-row rings of the six fit where the three resident ones did.
    from concurrent.futures import ThreadPoolExecutor
    for eng in engines:
        eng.close()                        # the resident batch leaves the device; the pinned host copies stay (host)
    lanes = []                             # per context: (engine, [(sub-batch, params, (results, pairs))])
    for hb, p, o in zip(batches, params, outs):
        bounds = np.linspace(0, hb.n, SUB + 1).astype(np.int64)
        pair = []
        for _ in range(min(max(1, args.e2e_contexts), SUB)):
            eng = Engine(local)
            eng.set_resident_warps(args.e2e_resident_warps)
            eng.set_exact_arithmetic(args.arithmetic == "exact")
            eng.upload_model(tbl_match, tbl_gapy, gapx_tbl)
            pair.append((eng, []))
        off = 0
        for k in range(SUB):
            sb = hb.view(int(bounds[k]), int(bounds[k + 1]))
            cap_k = Engine.default_pair_capacity(sb, per_event=args.pairs_per_event)
            assert off + cap_k <= len(o[1])
            pair[k % len(pair)][1].append((sb, p, (o[0][int(bounds[k]):int(bounds[k + 1])], o[1][off:off + cap_k])))
            off += cap_k
        lanes += pair
    pool = ThreadPoolExecutor(max_workers=len(lanes))
    counters = [[0, 0, 0] for _ in lanes]
    phase_ms = [dict(h2d_ms=0.0, prep_ms=0.0, plan_ms=0.0, align_ms=0.0, d2h_ms=0.0, call_ms=0.0) for _ in lanes]

    def run_lane(i):
        eng, tasks = lanes[i]
        for sb, p, o in tasks:
            t_call = time.perf_counter()
            eng.align_batch(sb, bench_hmm, p, 0, None, False, o)
            tm = eng.timing()
            counters[i][0] += tm["h2d_bytes"]; counters[i][1] += tm["d2h_bytes"]; counters[i][2] += tm["kernel_launches"]
            for k in ("h2d_ms", "prep_ms", "plan_ms", "align_ms", "d2h_ms"):
                phase_ms[i][k] += tm[k]
            phase_ms[i]["call_ms"] += (time.perf_counter() - t_call) * 1e3

    def e2e_step():
        for f in [pool.submit(run_lane, i) for i in range(len(lanes))]:
            f.result()

    e2e_step()
    for eng, tasks in lanes:               # the streamed results are the resident ones
        for sb, p, o in tasks:
            assert int((o[0]["status"] != 0).sum()) == 0, "non-zero item status (end-to-end leg)"
    assert sum(int(o[0]["n_pairs"].sum()) for eng, tasks in lanes for sb, p, o in tasks) == n_pairs
    barrier()
    counters = [[0, 0, 0] for _ in lanes]
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_wall = max_over_ranks(time.perf_counter() - t0)
    e2e_gcups = 2.0 * cells_total * args.steps / e2e_wall / 1e9
    h2d, d2h, e2e_launches = (sum_over_ranks(float(sum(c[i] for c in counters))) for i in range(3))   # whole job
    if os.environ.get("CPECAN_BENCH_PHASES"):
        for i, ph in enumerate(phase_ms):
            print("rank %d lane %d per step (ms): %s" % (rank, i, {k: round(v / (args.steps + 1), 1) for k, v in ph.items()}), file=sys.stderr, flush=True)
    pool.shutdown()
    for eng, _ in lanes:
        eng.close()
    host.close()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (k_align), from the live event timings of the resident steps ------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured" if "hbm_gbs" in peaks else "fallback"
    kern_s = kern_ms / 1e3                                   # this rank's kernels
    cells_steps = cells_rank * args.steps
    achieved_gbs = HBM_BYTES_PER_CELL * cells_steps / kern_s / 1e9
    sm_clock_hz = (clocks.get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)) * 1e6 if clocks else 1965e6
    issue_peak = info["sm_count"] * 128 * float(peaks.get("sm_max_mhz", 1965.0)) * 1e6
    issue_ach = ISSUE_OPS_PER_CELL * cells_steps / kern_s
    # DRAM bytes per band cell of k_align3, measured with ncu per expansion (tools: profiles/r2_dram_bytes_per_cell.json)
    dram_traffic_per_launch = None
    try:
        per_cell = json.load(open(os.path.join(ROOT, "profiles", "r2_dram_bytes_per_cell.json")))
        dram_traffic_per_launch = sum(float(per_cell[str(EXPANSIONS[j])]) * cells_by_engine[k] for k, j in enumerate(order)) / len(engines)
    except Exception:
        pass
    line = {
        "metric": METRIC, "value": gcups, "unit": "GCUPS", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True, "scaling": args.scaling if world > 1 else "weak",
        "vs_baseline": None, "dtype": "f64" if args.arithmetic == "exact" else "f32", "data": "synthetic",
        "reads_per_s": reads_total * args.steps / wall, "band_cells_per_s": cells_total * args.steps / wall,
        "config": {"workload": "C3: batched synthetic reads lX~6700 x lY~8000 events, 6-mer template model, anchors "
                               "every 50 k-mers, expansions 64/128/256 evenly, %s, threshold 0.01, ragged (1,1)" % (
                                   "vanilla machine" if vanilla else "three-state"),
                   "distinct_reads": "the batch tiles %d distinct reads; every copy has its own host and device buffers, so a "
                                     "copy shares no cache line with another" % n_unique,
                   "reads_per_gpu": B, "distinct_reads_per_gpu": n_unique, "reads_total": int(reads_total), "band_cells_per_step": int(cells_total),
                   "aligned_pairs_rank0": n_pairs,
                   "sharding": ("ONE batch of %d reads dealt to the ranks by band cells, longest first; no collective" % int(reads_total))
                   if (args.scaling == "strong" and world > 1) else "reads partitioned over ranks (the same number on every GPU), no collective",
                   "l2": "inputs + forward spill per step >> 126 MB L2 (no flush needed)"},
        "e2e": {"value": e2e_gcups, "unit": "GCUPS", "h2d_bytes_per_step": int(h2d / args.steps),
                "d2h_bytes_per_step": int(d2h / args.steps), "ms_per_step": e2e_wall / args.steps * 1e3,
                "reads_per_s": reads_total * args.steps / e2e_wall,
                "streaming": "%d sub-batches per expansion through %d contexts per GPU" % (SUB, len(lanes))},
        "gpu_launches": int(launches),
        "gpu_launches_e2e": int(e2e_launches),
        "kernel_ms_per_step": kern_ms_max / args.steps,
        # the bound that binds this kernel (north_star: FP32 / SFU issue rate against the per-cell op count of SURVEY.md
        # 8(d)); the HBM roofline of the forward-row spill is reported beside it
        "roofline": {"bound": "fp32_issue", "achieved": issue_ach / 1e12, "peak": issue_peak / 1e12,
                     "unit": "Tissue-op/s", "frac": issue_ach / issue_peak,
                     "ops_per_band_cell": ISSUE_OPS_PER_CELL, "sm_count": info["sm_count"],
                     "clock_mhz_for_peak": float(peaks.get("sm_max_mhz", 1965.0)),
                     "frac_at_observed_clock": issue_ach / (info["sm_count"] * 128 * sm_clock_hz),
                     "traffic": dram_traffic_per_launch, "traffic_unit": "DRAM bytes per launch (ncu dram__bytes_read.sum + "
                     "dram__bytes_write.sum per band cell and expansion, profiles/r2_dram_bytes_per_cell.json, x this run's cells)",
                     "peak_source": "148 SMs x 128 FP32 lanes x sm_max_mhz of MEASURED_PEAKS.json" if "sm_max_mhz" in peaks else "fallback 1965 MHz"},
        "roofline_hbm": {"bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved_gbs / hbm_peak, "traffic": dram_traffic_per_launch,
                         "measured_dram_gbs": dram_traffic_per_launch * len(engines) * args.steps / kern_s / 1e9 if dram_traffic_per_launch else None,
                         "algorithmic_bytes_per_launch": HBM_BYTES_PER_CELL * cells_rank / len(engines),
                         "peak_source": peak_src, "algorithmic_bytes_per_band_cell": HBM_BYTES_PER_CELL},
        "clocks": clocks,
    }
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n_s = len(reads)
        sample = reads[:n_s]
        exps = [EXPANSIONS[i % 3] for i in range(n_s)]    # reads[j::3] went to expansion j
        wall_c, cells_c, core_s = run_cpu_sample(sample, exps, cores)
        line["cpu_baseline"] = {"value": 2.0 * cells_c / wall_c / 1e9, "unit": "GCUPS", "cores": cores,
                                "kind": cpu_kind(),
                                "library_timed": getattr(run_cpu_sample, "libraries", []),
                                "sample": "%d reads of the same batch, one process per read on %d cores, %.1f s wall, "
                                          "%.1f core-s" % (n_s, cores, wall_c, core_s),
                                "band_cells_per_core_s": cells_c / core_s}
    if args.arithmetic == "exact":
        # the FP64 kernel (k_align_generic): the FP32 issue roofline and the ncu traffic figures above are k_align3's
        line["config"]["arithmetic"] = "exact: FP64 kernel k_align_generic in the reference's own operation order (cpecan_cuda_set_exact_arithmetic)"
        line["roofline"] = {"bound": "fp64_latency", "achieved": None, "peak": None, "unit": None, "frac": None, "traffic": None,
                            "note": "issue slots 24 % busy, long-scoreboard bound (profiles/r2_ncu_k_align_generic_e64_summary.txt)"}
        line.pop("roofline_hbm", None)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
