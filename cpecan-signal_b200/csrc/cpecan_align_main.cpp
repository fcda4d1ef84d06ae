// cpecanAlign -- the batched, GPU-backed sibling of the reference's signal CLI (vanillaAlign.c).
//
// Single-read mode takes the reference's flags and stdin (an exonerate cigar of the 2D read against the reference)
// and produces the reference's outputs: the 15-column posteriors TSV (writePosteriorProbs, vanillaAlign.c:26-96,
// appended), the per-read stdout line (:791-793) or, with -t and -c, the two expectation files (:668-731).
// Batch mode (--batch manifest) does the same for many reads with ALL strands of ALL reads in one GPU batch
// (getAlignedPairsUsingAnchorsBatch); manifest lines are  label <TAB> npRead <TAB> reference <TAB> posteriors-out <TAB> cigar.
// Host work only: file formats, cigar -> anchors, coordinate bookkeeping.  The DP runs in libcpecan_cuda.so.
#include <getopt.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "cpecan_host.h"

namespace {

struct Cigar {                              // sonLib cigarRead: the QUERY lands in contig2 / start2 / end2
    std::string contig1, contig2;
    int64_t start1 = 0, end1 = 0, strand1 = 1, start2 = 0, end2 = 0, strand2 = 1;
    std::vector<std::pair<char, int64_t>> ops;   // 'M' match, 'D' target only (INDEL_X), 'I' query only (INDEL_Y)
};

bool parseCigar(const std::string &line, Cigar &c) {
    if (line.compare(0, 6, "cigar:") != 0) return false;
    std::istringstream in(line.substr(6));
    std::string s1, s2;
    double score;
    if (!(in >> c.contig2 >> c.start2 >> c.end2 >> s2 >> c.contig1 >> c.start1 >> c.end1 >> s1 >> score)) return false;
    c.strand1 = s1 == "+"; c.strand2 = s2 == "+";
    std::string op;
    int64_t len;
    while (in >> op >> len) c.ops.emplace_back(op[0], len);
    return true;
}

std::string revComp(const std::string &s) {
    std::string r(s.rbegin(), s.rend());
    for (char &ch : r) {
        switch (ch) { case 'A': ch = 'T'; break; case 'C': ch = 'G'; break; case 'G': ch = 'C'; break; case 'T': ch = 'A'; break;
                      case 'a': ch = 't'; break; case 'c': ch = 'g'; break; case 'g': ch = 'c'; break; case 't': ch = 'a'; break; default: break; }
    }
    return r;
}

// guideAlignmentToRebasedAnchorPairs (vanillaAlign.c:278-299): rebase the reference interval to 0, flip a reverse-strand
// hit, convertPairwiseForwardStrandAlignmentToAnchorPairs (impl/pairwiseAligner.c:1039-1063) with the trim, sort, filter
stList *guideAnchors(Cigar c, int64_t trim) {
    const int64_t shift = c.strand1 ? c.start1 : c.end1;
    c.start1 -= shift; c.end1 -= shift;
    if (!c.strand1) { c.strand1 = 1; std::swap(c.start1, c.end1); }
    stList *raw = stList_construct3(0, (void (*)(void *)) stIntTuple_destruct);
    int64_t j = c.start1, k = c.start2;
    for (auto &op : c.ops) {
        if (op.first == 'M') for (int64_t l = trim; l < op.second - trim; l++) stList_append(raw, stIntTuple_construct2(j + l, k + l));
        if (op.first != 'I') j += op.second;
        if (op.first != 'D') k += op.second;
    }
    stList_sort(raw, stIntTuple_cmpFn);
    stList *out = filterToRemoveOverlap(raw);
    stList_destruct(raw);
    return out;
}

struct Options {
    StateMachineType type = vanilla;
    int64_t diagExpansion = 50, constraintTrim = 14;
    double threshold = 0.01;
    std::string templateModel = "../../cPecan/models/template_median68pA.model";
    std::string complementModel = "../../cPecan/models/complement_median68pA_pop2.model";
    std::string label, npRead, target, posteriors, tHmm, cHmm, tExp, cExp, batch, templateHdp, complementHdp;
    NanoporeHDP *nHdp[2] = { nullptr, nullptr };           // threeStateHdp: one per strand, shared by all reads
};

struct Job {                                // one read: both strands
    std::string label, posteriors;
    Cigar cig;
    NanoporeRead *np = nullptr;
    std::string ref, trimmed, rcTrimmed;
    stList *anchors = nullptr;
    StateMachine *sM[2] = { nullptr, nullptr };
    Sequence *sX[2] = { nullptr, nullptr }, *sY[2] = { nullptr, nullptr };
    stList *remapped[2] = { nullptr, nullptr }, *pairs[2] = { nullptr, nullptr };
};

StateMachine *buildStateMachine(const std::string &model, const NanoporeReadAdjustmentParameters &npp, StateMachineType type,
                                int strand, const std::string &hmmFile, NanoporeHDP *nHdp) {   // vanillaAlign.c:104-140, 236-240
    StateMachine *sM = type == vanilla ? getSignalStateMachine3Vanilla(model.c_str())
                     : type == fourState ? getStateMachine4(model.c_str())
                     : type == echelon ? getStateMachineEchelon(model.c_str())
                     : type == threeStateHdp ? getHdpStateMachine3(nHdp) : getStrawManStateMachine3(model.c_str());
    if (type != threeStateHdp) emissions_signal_scaleModel(sM, npp.scale, npp.shift, npp.var, npp.scale_sd, npp.var_sd);
    if (type == vanilla) stateMachine3Vanilla_setStrandTransitionsToDefaults(sM, strand ? complement : template_);
    if (!hmmFile.empty()) { fprintf(stderr, "loading HMM from file, %s\n", hmmFile.c_str()); hmmContinuous_loadSignalHmm(hmmFile.c_str(), sM, type); }
    return sM;
}

void prepare(Job &job, const Options &o, const std::string &npFile, const std::string &targetFile) {
    std::ifstream rf(targetFile);
    if (!rf || !std::getline(rf, job.ref)) st_errAbort("cpecanAlign: cannot read the reference %s", targetFile.c_str());
    job.np = nanopore_loadNanoporeReadFromFile(npFile.c_str());
    if (o.type == threeStateHdp) nanopore_descaleNanoporeRead(job.np);               // vanillaAlign.c:609-612
    const Cigar &c = job.cig;
    // getSubSequence + reverse complement for a reverse-strand hit (vanillaAlign.c:626-635)
    job.trimmed = c.strand1 ? job.ref.substr((size_t) c.start1, (size_t) (c.end1 - c.start1))
                            : revComp(job.ref.substr((size_t) c.end1, (size_t) (c.start1 - c.end1)));
    job.rcTrimmed = revComp(job.trimmed);
    job.anchors = guideAnchors(c, o.constraintTrim);
    for (int s = 0; s < 2; s++) {
        const int64_t *map = s ? job.np->complementEventMap : job.np->templateEventMap;
        double *events = s ? job.np->complementEvents : job.np->templateEvents;
        const NanoporeReadAdjustmentParameters &npp = s ? job.np->complementParams : job.np->templateParams;
        // makeEventSequenceFromPairwiseAlignment (vanillaAlign.c:301-316)
        const int64_t y0 = map[c.start2], y1 = map[c.end2];
        job.sY[s] = sequence_construct2(y1 - y0, events + y0 * NB_EVENT_PARAMS, sequence_getEvent, sequence_sliceEventSequence2);
        std::string &tgt = s ? job.rcTrimmed : job.trimmed;
        job.sX[s] = sequence_construct2(sequence_correctSeqLength((int64_t) tgt.size(), event), (void *) tgt.c_str(),
                                        (o.type == vanilla || o.type == echelon) ? sequence_getKmer2
                                        : o.type == threeStateHdp ? sequence_getKmer3 : sequence_getKmer,     // vanillaAlign.c:224-250
                                        sequence_sliceNucleotideSequence2);
        if (o.type == echelon) sequence_padSequence(job.sX[s]);                      // vanillaAlign.c:196-198
        job.sM[s] = buildStateMachine(s ? o.complementModel : o.templateModel, npp, o.type, s, s ? o.cHmm : o.tHmm, o.nHdp[s]);
        // getRemappedAnchorPairs (vanillaAlign.c:98-102)
        stList *rm = nanopore_remapAnchorPairsWithOffset(job.anchors, const_cast<int64_t *>(map), c.start2);
        job.remapped[s] = filterToRemoveOverlap(rm);
        stList_destruct(rm);
    }
}

// writePosteriorProbs (vanillaAlign.c:26-96): appended, 15 tab-separated columns
void writePosteriors(const Job &job, int s) {
    FILE *fh = fopen(job.posteriors.c_str(), "a");
    if (!fh) st_errAbort("cpecanAlign: cannot open %s", job.posteriors.c_str());
    const bool forward = job.cig.strand1 != 0;
    const std::string &target = s ? job.rcTrimmed : job.trimmed;
    const double *events = s ? job.np->complementEvents : job.np->templateEvents;
    const NanoporeReadAdjustmentParameters &npp = s ? job.np->complementParams : job.np->templateParams;
    const int64_t evOff = (s ? job.np->complementEventMap : job.np->templateEventMap)[job.cig.start2];
    const int64_t refOff = s ? job.cig.end1 : job.cig.start1;
    const double *matchModel = job.sM[s]->EMISSION_MATCH_PROBS;
    for (int64_t i = 0; i < stList_length(job.pairs[s]); i++) {
        stIntTuple *t = (stIntTuple *) stList_get(job.pairs[s], i);
        const int64_t x = stIntTuple_get(t, 1);
        int64_t xAdj;
        if ((s == 0) == forward) xAdj = x + refOff;
        else { const int64_t refLen = (int64_t) target.size(); xAdj = (refLen - KMER_LENGTH) - (x + (refLen - refOff)); }
        const int64_t y = stIntTuple_get(t, 2) + evOff;
        const double p = (double) stIntTuple_get(t, 0) / PAIR_ALIGNMENT_PROB_1;
        const double mean = events[y * NB_EVENT_PARAMS], noise = events[y * NB_EVENT_PARAMS + 1], dur = events[y * NB_EVENT_PARAMS + 2];
        std::string k = target.substr((size_t) x, KMER_LENGTH);
        k.resize(KMER_LENGTH, '\0');
        const int64_t ki = emissions_discrete_getKmerIndex((void *) k.c_str());
        const double eLevel = matchModel[1 + ki * MODEL_PARAMS], eNoise = matchModel[1 + ki * MODEL_PARAMS + 2];
        const std::string refKmer = ((s == 0) == forward) ? k : revComp(k);
        fprintf(fh, "%s\t%lld\t%s\t%s\t%s\t%lld\t%f\t%f\t%f\t%s\t%f\t%f\t%f\t%f\t%f\n", job.cig.contig1.c_str(), (long long) xAdj,
                refKmer.c_str(), job.label.c_str(), s ? "c" : "t", (long long) y, mean, noise, dur, k.c_str(), eLevel, eNoise, p,
                (mean - npp.shift) / npp.scale, (eLevel - npp.shift) / npp.scale);
    }
    fclose(fh);
}

double posteriorScore(stList *pairs) {      // scoreByPosteriorProbabilityIgnoringGaps (vanillaAlign.c:163-177)
    double tot = 0.0;
    for (int64_t i = 0; i < stList_length(pairs); i++) tot += (double) stIntTuple_get((stIntTuple *) stList_get(pairs, i), 0);
    return 100.0 * tot / ((double) stList_length(pairs) * PAIR_ALIGNMENT_PROB_1);
}

void release(Job &job) {
    for (int s = 0; s < 2; s++) {
        if (job.pairs[s]) stList_destruct(job.pairs[s]);
        stList_destruct(job.remapped[s]);
        sequence_sequenceDestroy(job.sX[s]); sequence_sequenceDestroy(job.sY[s]);
        stateMachine_destruct(job.sM[s]);
    }
    stList_destruct(job.anchors);
    nanopore_nanoporeReadDestruct(job.np);
}

void usage() {
    fprintf(stderr, "cpecanAlign: GPU sibling of vanillaAlign.  Flags as vanillaAlign (-s strawMan, -f fourState, -e echelon,\n"
                    "-d sm3Hdp with -v/-w the template / complement .nhdp files, -T/-C models, -L label, -q npRead,\n"
                    "-r reference, -u posteriors TSV, -t/-c expectation files, -y/-z input HMMs, -x expansion, -D threshold, -m trim);\n"
                    "cigar on stdin.  --batch <manifest>: label, npRead, reference, posteriors file, cigar per line (tab separated).\n");
}

}  // namespace

int main(int argc, char **argv) {
    Options o;
    static struct option longOpts[] = {
        { "help", no_argument, 0, 'h' }, { "strawMan", no_argument, 0, 's' }, { "templateModel", required_argument, 0, 'T' },
        { "complementModel", required_argument, 0, 'C' }, { "readLabel", required_argument, 0, 'L' }, { "npRead", required_argument, 0, 'q' },
        { "reference", required_argument, 0, 'r' }, { "posteriors", required_argument, 0, 'u' }, { "templateHmm", required_argument, 0, 'y' },
        { "complementHmm", required_argument, 0, 'z' }, { "templateExpectations", required_argument, 0, 't' },
        { "complementExpectations", required_argument, 0, 'c' }, { "diagonalExpansion", required_argument, 0, 'x' },
        { "threshold", required_argument, 0, 'D' }, { "constraintTrim", required_argument, 0, 'm' }, { "batch", required_argument, 0, 'B' },
        { "fourState", no_argument, 0, 'f' }, { "echelon", no_argument, 0, 'e' }, { "sm3Hdp", no_argument, 0, 'd' },
        { "templateHdp", required_argument, 0, 'v' }, { "complementHdp", required_argument, 0, 'w' },
        { 0, 0, 0, 0 } };
    int key;
    while ((key = getopt_long(argc, argv, "hsfedv:w:T:C:L:q:r:u:y:z:t:c:x:D:m:B:", longOpts, nullptr)) != -1) {
        switch (key) {
            case 'h': usage(); return 0;
            case 's': o.type = threeState; break;
            case 'f': o.type = fourState; break;
            case 'e': o.type = echelon; break;
            case 'd': o.type = threeStateHdp; break;
            case 'v': o.templateHdp = optarg; break;
            case 'w': o.complementHdp = optarg; break;
            case 'T': o.templateModel = optarg; break;
            case 'C': o.complementModel = optarg; break;
            case 'L': o.label = optarg; break;
            case 'q': o.npRead = optarg; break;
            case 'r': o.target = optarg; break;
            case 'u': o.posteriors = optarg; break;
            case 'y': o.tHmm = optarg; break;
            case 'z': o.cHmm = optarg; break;
            case 't': o.tExp = optarg; break;
            case 'c': o.cExp = optarg; break;
            case 'x': o.diagExpansion = atoll(optarg); break;
            case 'D': o.threshold = atof(optarg); break;
            case 'm': o.constraintTrim = atoll(optarg); break;
            case 'B': o.batch = optarg; break;
            default: usage(); return 1;
        }
    }
    fprintf(stderr, "cpecanAlign - using %s model\n", o.type == vanilla ? "vanilla" : o.type == fourState ? "fourState"
            : o.type == echelon ? "echelon" : o.type == threeStateHdp ? "strawMan-HDP" : "strawMan");
    if (o.type == threeStateHdp) {                                                   // vanillaAlign.c:574-600
        if (o.templateHdp.empty() || o.complementHdp.empty()) st_errAbort("Need to have template and complement HDPs");
        o.nHdp[0] = deserialize_nhdp(o.templateHdp.c_str());
        o.nHdp[1] = deserialize_nhdp(o.complementHdp.c_str());
    }
    PairwiseAlignmentParameters *p = pairwiseAlignmentBandingParameters_construct();
    p->threshold = o.threshold; p->constraintDiagonalTrim = o.constraintTrim; p->diagonalExpansion = o.diagExpansion;

    std::deque<Job> jobs;                   // deque: Sequence objects keep pointers into a job's strings
    if (!o.batch.empty()) {
        std::ifstream mf(o.batch);
        if (!mf) st_errAbort("cpecanAlign: cannot open the manifest %s", o.batch.c_str());
        std::string line;
        while (std::getline(mf, line)) {
            if (line.empty() || line[0] == '#') continue;
            std::vector<std::string> f;
            size_t a = 0, b;
            while (f.size() < 4 && (b = line.find('\t', a)) != std::string::npos) { f.push_back(line.substr(a, b - a)); a = b + 1; }
            f.push_back(line.substr(a));
            if (f.size() != 5) st_errAbort("cpecanAlign: manifest line needs 5 tab-separated fields: %s", line.c_str());
            Job job;
            job.label = f[0]; job.posteriors = f[3];
            if (!parseCigar(f[4], job.cig)) st_errAbort("cpecanAlign: bad cigar for %s", f[0].c_str());
            jobs.push_back(job);
            prepare(jobs.back(), o, f[1], f[2]);
        }
    } else {
        Job job;
        job.label = o.label; job.posteriors = o.posteriors;
        std::string line;
        bool ok = false;
        while (std::getline(std::cin, line)) if ((ok = parseCigar(line, job.cig))) break;
        if (!ok) st_errAbort("cpecanAlign: no cigar on stdin");
        jobs.push_back(job);
        prepare(jobs.back(), o, o.npRead, o.target);
    }

    if (!o.tExp.empty() && !o.cExp.empty()) {
        // expectation routine (vanillaAlign.c:668-731): one read, pseudocount 1e-4, ragged ends (1,1)
        if (jobs.size() != 1) st_errAbort("cpecanAlign: expectation files are written per read (use the EM driver for batches)");
        if (o.type == fourState || o.type == echelon) st_errAbort("vanillaAlign - getting expectations not allowed for this HMM type, yet");   // vanillaAlign.c:669-672
        Job &job = jobs[0];
        for (int s = 0; s < 2; s++) {
            fprintf(stderr, "cpecanAlign - getting expectations for %s\n", s ? "complement" : "template");
            Hmm *hmm = hmmContinuous_getEmptyHmm(o.type, 0.0001, p->threshold);
            if (o.type == vanilla) vanillaHmm_implantMatchModelsintoHmm(job.sM[s], hmm);
            getExpectationsUsingAnchors(job.sM[s], hmm, job.sX[s], job.sY[s], job.remapped[s], p, diagonalCalculation_Expectations, 1, 1);
            if (o.type == threeStateHdp) fprintf(stderr, "cpecanAlign - got %lld HDP assignments\n", (long long) hmmContinuous_howManyAssignments(hmm));
            hmmContinuous_writeToFile((s ? o.cExp : o.tExp).c_str(), hmm, o.type);
            hmmContinuous_destruct(hmm, o.type);
        }
    } else {
        // alignment routine (vanillaAlign.c:733-797): every strand of every read in ONE GPU batch, ragged ends (1,1)
        const int64_t n = 2 * (int64_t) jobs.size();
        std::vector<StateMachine *> sMs; std::vector<Sequence *> xs, ys; std::vector<stList *> as, res((size_t) n, nullptr);
        for (Job &job : jobs) for (int s = 0; s < 2; s++) { sMs.push_back(job.sM[s]); xs.push_back(job.sX[s]); ys.push_back(job.sY[s]); as.push_back(job.remapped[s]); }
        getAlignedPairsUsingAnchorsBatch(n, sMs.data(), xs.data(), ys.data(), as.data(), p, true, true, res.data());
        for (size_t j = 0; j < jobs.size(); j++) {
            Job &job = jobs[j];
            double score[2];
            for (int s = 0; s < 2; s++) {
                job.pairs[s] = res[2 * j + (size_t) s];
                score[s] = posteriorScore(job.pairs[s]);
                stList_sort(job.pairs[s], sortByXPlusYCoordinate2);
                if (!job.posteriors.empty()) writePosteriors(job, s);
            }
            fprintf(stdout, "%s %lld\t%lld(%f)\t", job.label.c_str(), (long long) stList_length(job.anchors),
                    (long long) stList_length(job.pairs[0]), score[0]);
            fprintf(stdout, "%lld(%f)\n", (long long) stList_length(job.pairs[1]), score[1]);
        }
    }
    for (Job &job : jobs) release(job);
    for (NanoporeHDP *h : o.nHdp) if (h) destroy_nanopore_hdp(h);
    pairwiseAlignmentBandingParameters_destruct(p);
    fprintf(stderr, "cpecanAlign - SUCCESS: finished %zu read(s)\n", jobs.size());
    return 0;
}
