"""Baum-Welch training from the command line: the batched, multi-GPU sibling of the reference's scripts/trainModels.py
(:244-330), fed like `cpecanAlign --batch` instead of with fast5 directories and bwa (those stay out of scope).

    python -m cpecan_signal.train --manifest reads.tsv --machine three --iterations 5 --amount 100000 \
        --template-model T.model --complement-model C.model --out-template-hmm t.hmm --out-complement-hmm c.hmm
    torchrun --nproc-per-node 8 -m cpecan_signal.train ...          # one rank per GPU

manifest: one read per line, tab separated: label, .npRead file, reference file (one line of nucleotides), an optional
column that is ignored (cpecanAlign's posteriors file), the guide alignment as an exonerate cigar line ("cigar: ...").

Per iteration and strand: every rank runs the E-step of its share of the reads on its GPU (shared by band cells), the
expectation vectors are all-reduced, every rank does the same M-step, rank 0 writes the .hmm file and all ranks reload it
-- the text round trip the reference's loop has between two rounds of vanillaAlign (-y / -z).  What one strand of one
read contributes is what `vanillaAlign -t / -c` writes for it (getSignalExpectations, vanillaAlign.c:318-360): the guide
anchors trimmed, rebased and filtered (:278-299), the events of the aligned stretch (:301-316), the anchors re-mapped to
events (:98-102), ragged ends (1, 1)."""
import argparse
import os
import sys

import numpy as np

from . import em, synth
from .engine import Engine, HostBatch, default_params, three_state_hmm, vanilla_gapx, vanilla_hmm

_COMP = str.maketrans("ACGTacgt", "TGCAtgca")


def reverse_complement(s):
    return s[::-1].translate(_COMP)


def parse_cigar(line):
    """exonerate cigar as sonLib's cigarRead takes it: query (contig2) start end strand, target (contig1) start end
    strand, score, then (op, length) pairs: M match, D target only, I query only."""
    if not line.startswith("cigar:"):
        raise ValueError("not a cigar line: %r" % line[:40])
    f = line[6:].split()
    c = dict(contig2=f[0], start2=int(f[1]), end2=int(f[2]), strand2=f[3] == "+", contig1=f[4], start1=int(f[5]), end1=int(f[6]),
             strand1=f[7] == "+", score=float(f[8]))
    c["ops"] = [(f[i][0], int(f[i + 1])) for i in range(9, len(f) - 1, 2)]
    return c


def filter_to_remove_overlap(pairs):
    """filterToRemoveOverlap (impl/pairwiseAligner.c:1160-1200) on pairs sorted by (x, y): keep those that are strictly
    increasing in x and y scanning backwards AND forwards."""
    pairs = np.asarray(pairs, dtype=np.int64).reshape(-1, 2)
    n = len(pairs)
    keep = np.zeros(n, dtype=bool)
    px = py = np.iinfo(np.int64).max
    for i in range(n - 1, -1, -1):
        x, y = pairs[i]
        if x < px and y < py:
            keep[i] = True
        px, py = min(px, x), min(py, y)
    kept = set(map(tuple, pairs[keep].tolist()))               # the reference looks membership up by VALUE
    out = []
    px = py = np.iinfo(np.int64).min
    for i in range(n):
        x, y = int(pairs[i, 0]), int(pairs[i, 1])
        if x > px and y > py and (x, y) in kept:
            out.append((x, y))
        px, py = max(px, x), max(py, y)
    return np.array(out, dtype=np.int64).reshape(-1, 2)


def guide_anchors(c, trim):
    """guideAlignmentToRebasedAnchorPairs (vanillaAlign.c:278-299) + convertPairwiseForwardStrandAlignmentToAnchorPairs
    (impl/pairwiseAligner.c:1039-1063): reference interval rebased to 0 (a reverse-strand hit flipped), the inner part of
    every match run, sorted, filtered."""
    start1, end1 = c["start1"], c["end1"]
    shift = start1 if c["strand1"] else end1
    start1, end1 = start1 - shift, end1 - shift
    if not c["strand1"]:
        start1, end1 = end1, start1
    raw = []
    j, k = start1, c["start2"]
    for op, n in c["ops"]:
        if op == "M":
            raw += [(j + l, k + l) for l in range(trim, n - trim)]
        if op != "I":
            j += n
        if op != "D":
            k += n
    raw.sort()
    return filter_to_remove_overlap(raw)


class StrandJob:
    """What one strand of one read contributes to a batch: target nucleotides, events, event-space anchors, scaling."""

    def __init__(self, label, strand, ref, events, anchors, scale5):
        self.label, self.strand, self.ref, self.events, self.anchors, self.scale5 = label, strand, ref, events, anchors, scale5


def descale_events(events, scale5):
    """nanopore_descaleNanoporeRead AS THE REFERENCE HAS IT (impl/nanopore.c:34-38, 228-236): the loop steps through the
    flat array by three but stops at the NUMBER of events, so only the means of the first third of a strand's events are
    descaled.  The HDP machine was trained and is run on exactly that."""
    ev = np.array(events, dtype=np.float64).reshape(-1, 3).copy()
    k = (len(ev) + 2) // 3
    ev[:k, 0] = (ev[:k, 0] - scale5[1]) / scale5[0]
    return ev


def prepare_read(label, np_path, ref_path, cigar_line, trim=14, descale=False):
    """The two StrandJobs of a read (template, complement); a strand whose event slice is empty or runs backwards
    (a decreasing event map, vanillaAlign.c:645-651) is skipped for training.  Returns (jobs, 2D read length).
    descale: the HDP machine's events (vanillaAlign.c:609-612)."""
    ref = open(ref_path).readline().strip()
    rd = synth.load_npread(np_path)
    if descale:
        for name in ("template", "complement"):
            rd[name + "_events"] = descale_events(rd[name + "_events"], rd[name + "_params"])
    c = parse_cigar(cigar_line)
    if c["strand1"]:
        trimmed = ref[c["start1"]:c["end1"]]
    else:
        trimmed = reverse_complement(ref[c["end1"]:c["start1"]])
    anchors = guide_anchors(c, trim)
    jobs = []
    for s, name in enumerate(("template", "complement")):
        emap, events, scale5 = rd[name + "_map"], rd[name + "_events"], rd[name + "_params"]
        y0, y1 = int(emap[c["start2"]]), int(emap[c["end2"]])
        if y1 <= y0:
            continue
        target = reverse_complement(trimmed) if s else trimmed
        # getRemappedAnchorPairs (vanillaAlign.c:98-102): (x, map[y] - map[start2]), filtered again
        rm = np.stack([anchors[:, 0], emap[anchors[:, 1]] - emap[c["start2"]]], axis=1) if len(anchors) else anchors
        jobs.append(StrandJob(label, s, target, events[y0:y1], filter_to_remove_overlap(rm), scale5))
    return jobs, rd["read_length"]


def read_manifest(path):
    out = []
    for line in open(path):
        line = line.rstrip("\n")
        if not line or line.startswith("#"):
            continue
        f = line.split("\t")
        if len(f) not in (4, 5):
            raise ValueError("manifest line needs label, npRead, reference, [posteriors,] cigar: %r" % line[:80])
        out.append((f[0], f[1], f[2], f[-1]))
    return out


def estep_strand(engine, jobs, tables, machine, hmm, params, distributed):
    """All-reduced expectation vector of one strand over this rank's jobs."""
    l1, _, l3 = tables
    mid = engine.upload_model(l1, l3, hmm["gapx"])
    try:
        if jobs:
            batch = HostBatch([j.ref for j in jobs], [j.events for j in jobs], [j.anchors for j in jobs],
                              model_ids=[mid] * len(jobs), scales=[j.scale5 for j in jobs], ragged=[(1, 1)] * len(jobs))
        else:                                          # a rank without reads still takes part in the all-reduce
            batch = HostBatch(["ACGTAC"], [np.zeros((0, 3))], [np.zeros((0, 2), np.int64)], model_ids=[mid], ragged=[(1, 1)])
        return em.gpu_estep(engine, batch, hmm["hmm"], params, distributed)
    finally:
        engine.release_model(mid)


def hdp_pass(a, engine, s, name, jobs_all, params, world, rank, distributed):
    """One strand with the HDP machine: E-step of this rank's share, ten sums all-reduced, the assignment lists gathered in
    rank order, rank 0 writes the HdpHmm file (pseudocount 1e-4 per transition slot, as vanillaAlign.c:675-676)."""
    from . import hdp as hdp_mod
    from .engine import hdp_hmm
    path = a.template_hdp if s == 0 else a.complement_hdp
    if not path:
        raise SystemExit("--machine hdp needs --template-hdp and --complement-hdp")
    out = a.out_template_hmm if s == 0 else a.out_complement_hmm
    container = em.HdpHmm(1e-4, params.threshold)
    if jobs_all:
        cells = [len(j.ref) * (2 * a.diagonal_expansion + 1) for j in jobs_all]
        # contiguous shares in manifest order, so that the gathered lists come out in manifest order
        bounds = np.searchsorted(np.cumsum(cells), np.linspace(0, sum(cells), world + 1)[1:-1], side="left")
        mine = np.split(np.arange(len(jobs_all)), bounds)[rank]
        mid = engine.upload_hdp(hdp_mod.load_nhdp(path))
        jobs = [jobs_all[int(i)] for i in mine]
        if jobs:
            batch = HostBatch([j.ref for j in jobs], [j.events for j in jobs], [j.anchors for j in jobs], model_ids=[mid] * len(jobs),
                              ragged=[(1, 1)] * len(jobs))
        else:
            batch = HostBatch(["ACGTAC"], [np.zeros((0, 3))], [np.zeros((0, 2), np.int64)], model_ids=[mid], ragged=[(1, 1)])
        em.gpu_hdp_estep(engine, batch, hdp_hmm(), params, distributed, container=container)
        engine.release_model(mid)
    if rank == 0:
        container.write(out)
        print("%s: %d strands, %d HDP assignments, likelihood %.6f" % (name, len(jobs_all), len(container.means), container.likelihood), flush=True)
    return [(name, 0, container.likelihood)]


def main(argv=None):
    ap = argparse.ArgumentParser(prog="python -m cpecan_signal.train", description=__doc__.split("\n\n")[0])
    ap.add_argument("--manifest", required=True)
    ap.add_argument("--machine", default="three", choices=["three", "vanilla", "hdp"],
                    help="hdp: ONE pass that collects the event-to-k-mer assignments and transition sums of every strand into "
                         "the HdpHmm files the reference's Gibbs step reads (vanillaAlign.c:142-154); needs --template-hdp / "
                         "--complement-hdp (serialised NanoporeHDPs)")
    ap.add_argument("--template-hdp"); ap.add_argument("--complement-hdp")
    ap.add_argument("--iterations", type=int, default=10)
    ap.add_argument("--amount", type=int, default=0, help="train on reads until their 2D lengths add up to this (0: all reads)")
    ap.add_argument("--seed", type=int, default=0, help="of the shuffle that picks the training reads")
    ap.add_argument("--template-model", default=synth.TEMPLATE_MODEL)
    ap.add_argument("--complement-model", default=synth.COMPLEMENT_MODEL)
    ap.add_argument("--in-template-hmm"); ap.add_argument("--in-complement-hmm")
    ap.add_argument("--out-template-hmm", required=True); ap.add_argument("--out-complement-hmm", required=True)
    ap.add_argument("-x", "--diagonal-expansion", type=int, default=50)
    ap.add_argument("-D", "--threshold", type=float, default=0.01)
    ap.add_argument("-m", "--constraint-trim", type=int, default=14)
    ap.add_argument("--exact", action="store_true", help="FP64 E-step in the reference's own operation order")
    a = ap.parse_args(argv)

    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    distributed = world > 1
    if distributed:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    reads = read_manifest(a.manifest)
    prepared = [prepare_read(*r, trim=a.constraint_trim, descale=a.machine == "hdp") for r in reads]
    lengths = [p[1] for p in prepared]
    chosen = em.cull_training_reads(lengths, a.amount, np.random.default_rng(a.seed)) if a.amount > 0 else np.arange(len(reads))
    params = default_params(diagonalExpansion=a.diagonal_expansion, threshold=a.threshold, constraintDiagonalTrim=a.constraint_trim)
    engine = Engine(local)
    engine.set_exact_arithmetic(a.exact)
    vanilla = a.machine == "vanilla"
    log = []
    for s, name in enumerate(("template", "complement")):
        jobs_all = [j for i in chosen for j in prepared[int(i)][0] if j.strand == s]
        if a.machine == "hdp":
            log += hdp_pass(a, engine, s, name, jobs_all, params, world, rank, distributed)
            continue
        tables = synth.load_model_file(a.template_model if s == 0 else a.complement_model)
        if not jobs_all:
            if rank == 0:
                print("%s: no read has an aligned stretch of events on this strand; nothing to train" % name, flush=True)
            continue
        cells = [len(j.ref) * (2 * a.diagonal_expansion + 1) for j in jobs_all]        # a proxy is enough to balance the shards
        mine = [jobs_all[int(i)] for i in em.shard_by_cells(cells, world)[rank]]
        model = em.ConditionalSignalHmm(match_model=tables[0], scaled_match_model=tables[2]) if vanilla else em.ContinuousPairHmm()
        in_hmm = a.in_template_hmm if s == 0 else a.in_complement_hmm
        trans, gapx = None, None
        if in_hmm:
            trans, gapx = type(model).load(in_hmm).state_machine_params()
        out_hmm = a.out_template_hmm if s == 0 else a.out_complement_hmm
        for it in range(a.iterations):
            if vanilla:
                hmm = dict(hmm=vanilla_hmm(name), gapx=vanilla_gapx(tables[1]) if gapx is None else gapx)
            else:
                hmm = dict(hmm=three_state_hmm(trans), gapx=np.full(4096, -2.3025850929940455) if gapx is None else gapx)
            vec = estep_strand(engine, mine, tables, a.machine, hmm, params, distributed)
            barrier = None
            if distributed:
                import torch.distributed as dist
                barrier = dist.barrier
            loaded = em.em_iteration(model, vec, len(jobs_all), out_hmm, rank=rank, barrier=barrier)
            trans, gapx = loaded.state_machine_params()
            log.append((name, it, model.running_likelihoods[-1]))
            if rank == 0:
                print("%s iteration %d: %d strands, likelihood %.6f" % (name, it, len(jobs_all), model.running_likelihoods[-1]), flush=True)
    engine.close()
    if distributed:
        import torch.distributed as dist
        dist.destroy_process_group()
    return log


if __name__ == "__main__":
    main()
