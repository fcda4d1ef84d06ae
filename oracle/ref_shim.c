/*
 * TEST INFRASTRUCTURE ONLY -- never linked into, imported by or executed from the product path.
 *
 * Flat C-ABI wrapper around the UNMODIFIED reference sources (compiled where they lie under
 * /root/reference by oracle/Makefile into oracle/_ref/libcpecan_ref.so).  It lets the Python
 * tests and bench.py's cpu_baseline leg call the reference's own
 *   getAlignedPairsUsingAnchors        (impl/pairwiseAligner.c:1456)
 *   getAlignedPairsWithoutBanding      (impl/pairwiseAligner.c:1512)
 *   getExpectationsUsingAnchors        (impl/pairwiseAligner.c:1571)
 *   band_construct / logAdd / getSplitPoints / filterToRemoveOverlap
 * with plain pointers, so the oracle restatement (oracle/cpecan_oracle.c) can be pinned
 * against the real thing on arbitrary inputs.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "sonLib.h"
#include "pairwiseAligner.h"
#include "stateMachine.h"
#include "continuousHmm.h"
#include "nanopore.h"
#include "emissionMatrix.h"

typedef struct {
    double threshold;
    int64_t minDiagsBetweenTraceBack;
    int64_t traceBackDiagonals;
    int64_t diagonalExpansion;
    int64_t constraintDiagonalTrim;
    int64_t splitMatrixBiggerThanThis;
} RefParams;

static PairwiseAlignmentParameters *makeParams(const RefParams *rp) {
    PairwiseAlignmentParameters *p = pairwiseAlignmentBandingParameters_construct();
    if (rp) {
        p->threshold = rp->threshold;
        p->minDiagsBetweenTraceBack = rp->minDiagsBetweenTraceBack;
        p->traceBackDiagonals = rp->traceBackDiagonals;
        p->diagonalExpansion = rp->diagonalExpansion;
        p->constraintDiagonalTrim = rp->constraintDiagonalTrim;
        p->splitMatrixBiggerThanThis = rp->splitMatrixBiggerThanThis;
    }
    return p;
}

/* smType: 2 = threeState (strawMan), 4 = vanilla.  scale5 = {scale, shift, var, scale_sd, var_sd} or NULL.
 * strand: 0 template, 1 complement (vanilla only).
 * transitions: NULL or 9 doubles in StateMachine3 field order (threeState only):
 *   MATCH_CONTINUE, MATCH_FROM_GAP_X, MATCH_FROM_GAP_Y, GAP_OPEN_X, GAP_OPEN_Y,
 *   GAP_EXTEND_X, GAP_EXTEND_Y, GAP_SWITCH_TO_X, GAP_SWITCH_TO_Y
 * gapX: NULL or 4096 log-probs for EMISSION_GAP_X_PROBS (threeState only). */
static StateMachine *makeStateMachine(int smType, const char *modelFile, const double *scale5, int strand,
                                      const double *transitions, const double *gapX) {
    StateMachine *sM;
    if (smType == threeState) {
        sM = getStrawManStateMachine3(modelFile);
    } else if (smType == vanilla) {
        sM = getSignalStateMachine3Vanilla(modelFile);
    } else if (smType == fourState) {
        sM = getStateMachine4(modelFile);
    } else if (smType == echelon) {
        sM = getStateMachineEchelon(modelFile);
    } else {
        fprintf(stderr, "ref_shim: unsupported state machine type %d\n", smType);
        return NULL;
    }
    if (scale5) emissions_signal_scaleModel(sM, scale5[0], scale5[1], scale5[2], scale5[3], scale5[4]);
    if (smType == vanilla) stateMachine3Vanilla_setStrandTransitionsToDefaults(sM, strand ? complement : template);
    if (smType == threeState && transitions) {
        StateMachine3 *s3 = (StateMachine3 *) sM;
        s3->TRANSITION_MATCH_CONTINUE = transitions[0];
        s3->TRANSITION_MATCH_FROM_GAP_X = transitions[1];
        s3->TRANSITION_MATCH_FROM_GAP_Y = transitions[2];
        s3->TRANSITION_GAP_OPEN_X = transitions[3];
        s3->TRANSITION_GAP_OPEN_Y = transitions[4];
        s3->TRANSITION_GAP_EXTEND_X = transitions[5];
        s3->TRANSITION_GAP_EXTEND_Y = transitions[6];
        s3->TRANSITION_GAP_SWITCH_TO_X = transitions[7];
        s3->TRANSITION_GAP_SWITCH_TO_Y = transitions[8];
    }
    if (smType == threeState && gapX) {
        for (int64_t i = 0; i < NUM_OF_KMERS; i++) sM->EMISSION_GAP_X_PROBS[i] = gapX[i];
    }
    return sM;
}

static void freeStateMachine(StateMachine *sM) {
    free(sM->EMISSION_MATCH_PROBS);
    free(sM->EMISSION_GAP_X_PROBS);
    free(sM->EMISSION_GAP_Y_PROBS);
    stateMachine_destruct(sM);
}

static stList *makeAnchors(const int64_t *anchors, int64_t nAnchors) {
    stList *l = stList_construct3(0, (void (*)(void *)) stIntTuple_destruct);
    for (int64_t i = 0; i < nAnchors; i++) stList_append(l, stIntTuple_construct2(anchors[2 * i], anchors[2 * i + 1]));
    return l;
}

static int64_t drainPairs(stList *pairs, int64_t *out, int64_t cap) {
    int64_t n = stList_length(pairs);
    for (int64_t i = 0; i < n && i < cap; i++) {
        stIntTuple *t = stList_get(pairs, i);
        out[3 * i] = stIntTuple_get(t, 0);
        out[3 * i + 1] = stIntTuple_get(t, 1);
        out[3 * i + 2] = stIntTuple_get(t, 2);
    }
    stList_destruct(pairs);
    return n;
}

/* ---- per-diagonal recorder: wraps the reference posterior callback so that the
 * totalProbability handed to every diagonal is captured as well. */
static double *g_totals = NULL;      /* indexed by xay, caller-sized lX+lY+1; NaN = not visited */
static int64_t g_totalsLen = 0;
static int64_t g_xOffset = 0, g_yOffset = 0; /* unused for totals (region-local xay) */

static void recordingPosteriorFn(StateMachine *sM, int64_t xay, DpMatrix *f, DpMatrix *b, Sequence *sX, Sequence *sY,
                                 double totalProbability, PairwiseAlignmentParameters *p, void *extraArgs) {
    if (g_totals && xay >= 0 && xay < g_totalsLen) g_totals[xay] = totalProbability;
    if (sM->type == echelon) diagonalCalculationMultiPosteriorMatchProbs(sM, xay, f, b, sX, sY, totalProbability, p, extraArgs);
    else diagonalCalculationPosteriorMatchProbs(sM, xay, f, b, sX, sY, totalProbability, p, extraArgs);
}

int64_t ref_align_banded(int smType, const char *modelFile, const double *scale5, int strand,
                         const double *transitions, const double *gapX,
                         const char *refSeq, const double *events, int64_t lY,
                         const int64_t *anchors, int64_t nAnchors, const RefParams *rp,
                         int raggedLeft, int raggedRight, int64_t *out, int64_t cap,
                         double *totalsOut, int64_t totalsLen) {
    StateMachine *sM = makeStateMachine(smType, modelFile, scale5, strand, transitions, gapX);
    if (!sM) return -1;
    PairwiseAlignmentParameters *p = makeParams(rp);
    int64_t lX = sequence_correctSeqLength(strlen(refSeq), event);
    Sequence *sX = sequence_construct2(lX, (void *) refSeq,
                                       (smType == vanilla || smType == echelon) ? sequence_getKmer2 : sequence_getKmer,
                                       sequence_sliceNucleotideSequence2);
    if (smType == echelon) sequence_padSequence(sX);     /* tests/signalPairwiseTest.c:1422-1423 */
    Sequence *sY = sequence_construct2(lY, (void *) events, sequence_getEvent, sequence_sliceEventSequence2);
    stList *anchorList = makeAnchors(anchors, nAnchors);
    g_totals = totalsOut; g_totalsLen = totalsLen; (void) g_xOffset; (void) g_yOffset;
    if (totalsOut) for (int64_t i = 0; i < totalsLen; i++) totalsOut[i] = NAN;
    stList *pairs = getAlignedPairsUsingAnchors(sM, sX, sY, anchorList, p,
                                                (totalsOut || smType == echelon) ? recordingPosteriorFn
                                                                                 : diagonalCalculationPosteriorMatchProbs,
                                                raggedLeft, raggedRight);
    g_totals = NULL; g_totalsLen = 0;
    int64_t n = drainPairs(pairs, out, cap);
    stList_destruct(anchorList);
    sequence_sequenceDestroy(sX);
    sequence_sequenceDestroy(sY);
    pairwiseAlignmentBandingParameters_destruct(p);
    freeStateMachine(sM);
    return n;
}

int64_t ref_align_unbanded(int smType, const char *modelFile, const double *scale5, int strand,
                           const double *transitions, const double *gapX,
                           const char *refSeq, const double *events, int64_t lY, const RefParams *rp,
                           int raggedLeft, int raggedRight, int64_t *out, int64_t cap, double *totalOut) {
    StateMachine *sM = makeStateMachine(smType, modelFile, scale5, strand, transitions, gapX);
    if (!sM) return -1;
    PairwiseAlignmentParameters *p = makeParams(rp);
    int64_t lX = sequence_correctSeqLength(strlen(refSeq), event);
    double tot = NAN;
    g_totals = &tot; g_totalsLen = 1; /* the same total is handed to every diagonal; record xay == 0 */
    stList *pairs = getAlignedPairsWithoutBanding(sM, (void *) refSeq, (void *) events, lX, lY, p,
                                                  (smType == vanilla || smType == echelon) ? sequence_getKmer2 : sequence_getKmer,
                                                  sequence_getEvent, recordingPosteriorFn, raggedLeft, raggedRight);
    g_totals = NULL; g_totalsLen = 0;
    if (totalOut) *totalOut = tot;
    int64_t n = drainPairs(pairs, out, cap);
    pairwiseAlignmentBandingParameters_destruct(p);
    freeStateMachine(sM);
    return n;
}

/* CPU-baseline helper: the same call as ref_align_banded, but only getAlignedPairsUsingAnchors itself is timed
 * (state-machine construction, i.e. parsing the .model text file, is outside the clock).  Returns the pair count. */
#include <time.h>
int64_t ref_time_align_banded(int smType, const char *modelFile, const double *scale5, int strand,
                              const char *refSeq, const double *events, int64_t lY,
                              const int64_t *anchors, int64_t nAnchors, const RefParams *rp,
                              int raggedLeft, int raggedRight, double *secondsOut) {
    StateMachine *sM = makeStateMachine(smType, modelFile, scale5, strand, NULL, NULL);
    if (!sM) return -1;
    PairwiseAlignmentParameters *p = makeParams(rp);
    int64_t lX = sequence_correctSeqLength(strlen(refSeq), event);
    Sequence *sX = sequence_construct2(lX, (void *) refSeq, smType == vanilla ? sequence_getKmer2 : sequence_getKmer,
                                       sequence_sliceNucleotideSequence2);
    Sequence *sY = sequence_construct2(lY, (void *) events, sequence_getEvent, sequence_sliceEventSequence2);
    stList *anchorList = makeAnchors(anchors, nAnchors);
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    stList *pairs = getAlignedPairsUsingAnchors(sM, sX, sY, anchorList, p, diagonalCalculationPosteriorMatchProbs,
                                                raggedLeft, raggedRight);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (secondsOut) *secondsOut = (double) (t1.tv_sec - t0.tv_sec) + 1e-9 * (double) (t1.tv_nsec - t0.tv_nsec);
    int64_t n = stList_length(pairs);
    stList_destruct(pairs);
    stList_destruct(anchorList);
    sequence_sequenceDestroy(sX);
    sequence_sequenceDestroy(sY);
    pairwiseAlignmentBandingParameters_destruct(p);
    freeStateMachine(sM);
    return n;
}

/* threeState: expOut = 9 transitions (row major from*3+to) + 4096 kmer-skip counts + 1 likelihood = 4106 doubles.
 * vanilla   : expOut = 60 skip-bin counts + 1 likelihood = 61 doubles. */
int64_t ref_expectations(int smType, const char *modelFile, const double *scale5, int strand,
                         const double *transitions, const double *gapX,
                         const char *refSeq, const double *events, int64_t lY,
                         const int64_t *anchors, int64_t nAnchors, const RefParams *rp,
                         int raggedLeft, int raggedRight, double pseudocount, double *expOut) {
    StateMachine *sM = makeStateMachine(smType, modelFile, scale5, strand, transitions, gapX);
    if (!sM) return -1;
    PairwiseAlignmentParameters *p = makeParams(rp);
    int64_t lX = sequence_correctSeqLength(strlen(refSeq), event);
    Sequence *sX = sequence_construct2(lX, (void *) refSeq, smType == vanilla ? sequence_getKmer2 : sequence_getKmer,
                                       sequence_sliceNucleotideSequence2);
    Sequence *sY = sequence_construct2(lY, (void *) events, sequence_getEvent, sequence_sliceEventSequence2);
    stList *anchorList = makeAnchors(anchors, nAnchors);
    Hmm *hmm = hmmContinuous_getEmptyHmm(smType, pseudocount, 0.0);
    if (smType == vanilla) vanillaHmm_implantMatchModelsintoHmm(sM, hmm);
    getExpectationsUsingAnchors(sM, hmm, sX, sY, anchorList, p, diagonalCalculation_Expectations, raggedLeft, raggedRight);
    if (smType == threeState) {
        ContinuousPairHmm *cp = (ContinuousPairHmm *) hmm;
        for (int i = 0; i < 9; i++) expOut[i] = cp->transitions[i];
        for (int i = 0; i < NUM_OF_KMERS; i++) expOut[9 + i] = cp->individualKmerGapProbs[i];
        expOut[9 + NUM_OF_KMERS] = hmm->likelihood;
    } else {
        VanillaHmm *vh = (VanillaHmm *) hmm;
        for (int i = 0; i < 60; i++) expOut[i] = vh->kmerSkipBins[i];
        expOut[60] = hmm->likelihood;
    }
    hmmContinuous_destruct(hmm, smType);
    stList_destruct(anchorList);
    sequence_sequenceDestroy(sX);
    sequence_sequenceDestroy(sY);
    pairwiseAlignmentBandingParameters_destruct(p);
    freeStateMachine(sM);
    return 0;
}

/* Band geometry: writes (xay, xmyL, xmyR) for xay = 0..lX+lY. */
void ref_band(const int64_t *anchors, int64_t nAnchors, int64_t lX, int64_t lY, int64_t expansion, int64_t *out) {
    stList *anchorList = makeAnchors(anchors, nAnchors);
    Band *band = band_construct(anchorList, lX, lY, expansion);
    BandIterator *it = bandIterator_construct(band);
    for (int64_t i = 0; i <= lX + lY; i++) {
        Diagonal d = bandIterator_getNext(it);
        out[3 * i] = diagonal_getXay(d);
        out[3 * i + 1] = diagonal_getMinXmy(d);
        out[3 * i + 2] = diagonal_getMaxXmy(d);
    }
    bandIterator_destruct(it);
    band_destruct(band);
    stList_destruct(anchorList);
}

double ref_logAdd(double x, double y) { return logAdd(x, y); }

/* getSplitPoints: writes 4-tuples (x1,y1,x2,y2); returns their number. */
int64_t ref_split_points(const int64_t *anchors, int64_t nAnchors, int64_t lX, int64_t lY, int64_t maxMatrix,
                         int raggedLeft, int raggedRight, int64_t *out, int64_t cap) {
    stList *anchorList = makeAnchors(anchors, nAnchors);
    stList *sp = getSplitPoints(anchorList, lX, lY, maxMatrix, raggedLeft, raggedRight);
    int64_t n = stList_length(sp);
    for (int64_t i = 0; i < n && i < cap; i++)
        for (int j = 0; j < 4; j++) out[4 * i + j] = stIntTuple_get(stList_get(sp, i), j);
    stList_destruct(sp);
    stList_destruct(anchorList);
    return n;
}

/* filterToRemoveOverlap on a sorted pair list; returns the surviving count. */
int64_t ref_filter_overlap(const int64_t *pairs, int64_t nPairs, int64_t *out) {
    stList *l = makeAnchors(pairs, nPairs);
    stList *f = filterToRemoveOverlap(l);
    int64_t n = stList_length(f);
    for (int64_t i = 0; i < n; i++) {
        out[2 * i] = stIntTuple_get(stList_get(f, i), 0);
        out[2 * i + 1] = stIntTuple_get(stList_get(f, i), 1);
    }
    stList_destruct(f);
    stList_destruct(l);
    return n;
}

/* Fixture support: the reference's own lastz anchoring + event-map remap + overlap filter, exactly as
 * tests/signalPairwiseTest.c:1141-1145 does it.  Needs ./cPecanLastz in the cwd (oracle/Makefile builds it).
 * strandSel: 0 template map, 1 complement map.  Returns the number of filtered anchors. */
int64_t ref_fixture_anchors(const char *refSeq, const char *npReadFile, int strandSel, int64_t *out, int64_t cap,
                            int64_t *rawCountOut) {
    NanoporeRead *np = nanopore_loadNanoporeReadFromFile(npReadFile);
    PairwiseAlignmentParameters *p = pairwiseAlignmentBandingParameters_construct();
    stList *anchorPairs = getBlastPairsForPairwiseAlignmentParameters((void *) refSeq, np->twoDread, p);
    if (rawCountOut) *rawCountOut = stList_length(anchorPairs);
    stList *remapped = nanopore_remapAnchorPairs(anchorPairs, strandSel ? np->complementEventMap : np->templateEventMap);
    stList *filtered = filterToRemoveOverlap(remapped);
    int64_t n = stList_length(filtered);
    for (int64_t i = 0; i < n && i < cap; i++) {
        out[2 * i] = stIntTuple_get(stList_get(filtered, i), 0);
        out[2 * i + 1] = stIntTuple_get(stList_get(filtered, i), 1);
    }
    stList_destruct(filtered);
    stList_destruct(remapped);
    stList_destruct(anchorPairs);
    pairwiseAlignmentBandingParameters_destruct(p);
    nanopore_nanoporeReadDestruct(np);
    return n;
}

/* npRead loader pass-through: sizes first (events may be NULL), then the data. */
int64_t ref_load_npread(const char *npReadFile, int64_t *dims3, double *params10, char *twoD,
                        int64_t *tMap, double *tEvents, int64_t *cMap, double *cEvents) {
    NanoporeRead *np = nanopore_loadNanoporeReadFromFile(npReadFile);
    dims3[0] = np->readLength; dims3[1] = np->nbTemplateEvents; dims3[2] = np->nbComplementEvents;
    if (params10) {
        params10[0] = np->templateParams.scale; params10[1] = np->templateParams.shift; params10[2] = np->templateParams.var;
        params10[3] = np->templateParams.scale_sd; params10[4] = np->templateParams.var_sd;
        params10[5] = np->complementParams.scale; params10[6] = np->complementParams.shift; params10[7] = np->complementParams.var;
        params10[8] = np->complementParams.scale_sd; params10[9] = np->complementParams.var_sd;
    }
    if (twoD) memcpy(twoD, np->twoDread, np->readLength);
    if (tMap) memcpy(tMap, np->templateEventMap, sizeof(int64_t) * np->readLength);
    if (cMap) memcpy(cMap, np->complementEventMap, sizeof(int64_t) * np->readLength);
    if (tEvents) memcpy(tEvents, np->templateEvents, sizeof(double) * 3 * np->nbTemplateEvents);
    if (cEvents) memcpy(cEvents, np->complementEvents, sizeof(double) * 3 * np->nbComplementEvents);
    nanopore_nanoporeReadDestruct(np);
    return 0;
}

/* ---- expectation container (impl/continuousHmm.c): file format, normalisation, M-step load --------------------- */

/* Fills a threeState ContinuousPairHmm from exp[4106] (9 transitions, 4096 k-mer skips, likelihood), optionally
 * normalises it (continuousPairHmm_normalize) and writes it with the reference's own writer. */
int64_t ref_write_pair_hmm(const double *exp, int normalize, const char *path) {
    Hmm *hmm = hmmContinuous_getEmptyHmm(threeState, 0.0, 0.0);
    ContinuousPairHmm *cp = (ContinuousPairHmm *) hmm;
    for (int i = 0; i < 9; i++) cp->transitions[i] = exp[i];
    for (int i = 0; i < NUM_OF_KMERS; i++) cp->individualKmerGapProbs[i] = exp[9 + i];
    hmm->likelihood = exp[9 + NUM_OF_KMERS];
    if (normalize) hmmContinuous_normalize(hmm, threeState);
    hmmContinuous_writeToFile(path, hmm, threeState);
    hmmContinuous_destruct(hmm, threeState);
    return 0;
}

/* Loads an .hmm file with the reference's loader into a strawMan state machine (hmmContinuous_loadSignalHmm ->
 * continuousPairHmm_loadTransitionsAndKmerGapProbs) and returns what the DP will see: the 9 transitions in
 * StateMachine3 field order and the 4096 gap-X emissions. */
int64_t ref_load_pair_hmm(const char *hmmPath, const char *modelFile, double *trans9, double *gapX4096) {
    StateMachine *sM = getStrawManStateMachine3(modelFile);
    hmmContinuous_loadSignalHmm(hmmPath, sM, threeState);
    StateMachine3 *s3 = (StateMachine3 *) sM;
    trans9[0] = s3->TRANSITION_MATCH_CONTINUE; trans9[1] = s3->TRANSITION_MATCH_FROM_GAP_X;
    trans9[2] = s3->TRANSITION_MATCH_FROM_GAP_Y; trans9[3] = s3->TRANSITION_GAP_OPEN_X;
    trans9[4] = s3->TRANSITION_GAP_OPEN_Y; trans9[5] = s3->TRANSITION_GAP_EXTEND_X;
    trans9[6] = s3->TRANSITION_GAP_EXTEND_Y; trans9[7] = s3->TRANSITION_GAP_SWITCH_TO_X;
    trans9[8] = s3->TRANSITION_GAP_SWITCH_TO_Y;
    for (int i = 0; i < NUM_OF_KMERS; i++) gapX4096[i] = sM->EMISSION_GAP_X_PROBS[i];
    freeStateMachine(sM);
    return 0;
}

/* The same for the vanilla container: bins[61] (60 skip bins + likelihood) and the two model lines taken from the
 * vanilla state machine built from modelFile (vanillaHmm_implantMatchModelsintoHmm); optionally normalised with the
 * reference's own (joint) vanillaHmm_normalizeKmerSkipBins; written with vanillaHmm_writeToFile. */
int64_t ref_write_vanilla_hmm(const double *bins, int normalize, const char *modelFile, const char *path) {
    StateMachine *sM = getSignalStateMachine3Vanilla(modelFile);
    Hmm *hmm = hmmContinuous_getEmptyHmm(vanilla, 0.0, 0.0);
    VanillaHmm *vh = (VanillaHmm *) hmm;
    for (int i = 0; i < 60; i++) vh->kmerSkipBins[i] = bins[i];
    hmm->likelihood = bins[60];
    vanillaHmm_implantMatchModelsintoHmm(sM, hmm);
    if (normalize) hmmContinuous_normalize(hmm, vanilla);
    hmmContinuous_writeToFile(path, hmm, vanilla);
    hmmContinuous_destruct(hmm, vanilla);
    freeStateMachine(sM);
    return 0;
}

/* Loads a vanilla .hmm file with the reference's loader (hmmContinuous_loadSignalHmm -> vanillaHmm_loadKmerSkipBin
 * Expectations) and returns the 60 skip-bin entries of EMISSION_GAP_X_PROBS the DP will see. */
int64_t ref_load_vanilla_hmm(const char *hmmPath, const char *modelFile, double *bins60) {
    StateMachine *sM = getSignalStateMachine3Vanilla(modelFile);
    hmmContinuous_loadSignalHmm(hmmPath, sM, vanilla);
    for (int i = 0; i < 60; i++) bins60[i] = sM->EMISSION_GAP_X_PROBS[i];
    freeStateMachine(sM);
    return 0;
}

#ifdef STANDIN_REAL_HDP
/* ---- threeStateHdp: only in oracle/_ref/libcpecan_ref_hdp.so, which also links the reference's HDP
 * sources (impl/hdp.c, impl/nanopore_hdp.c, impl/hdp_math_utils.c, impl/ranlib.c, impl/rnglib.c). */
#include "nanopore_hdp.h"

/* get_nanopore_kmer_density of a serialised NanoporeHDP: out[i * nx + j] for k-mer i (6 characters each) at x[j] */
int64_t ref_hdp_density(const char *nhdpFile, const char *kmers, int64_t nKmers, const double *x, int64_t nx, double *out) {
    NanoporeHDP *nhdp = deserialize_nhdp(nhdpFile);
    for (int64_t i = 0; i < nKmers; i++) {
        char km[7];
        memcpy(km, kmers + 6 * i, 6); km[6] = 0;
        for (int64_t j = 0; j < nx; j++) { double q = x[j]; out[i * nx + j] = get_nanopore_kmer_density(nhdp, km, &q); }
    }
    destroy_nanopore_hdp(nhdp);
    return nKmers * nx;
}

/* getAlignedPairsUsingAnchors with getHdpStateMachine3 and sequence_getKmer3 (vanillaAlign.c:132-135, 246-250) */
int64_t ref_hdp_align_banded(const char *nhdpFile, const char *refSeq, const double *events, int64_t lY,
                             const int64_t *anchors, int64_t nAnchors, const RefParams *rp,
                             int raggedLeft, int raggedRight, int64_t *out, int64_t cap,
                             double *totalsOut, int64_t totalsLen) {
    NanoporeHDP *nhdp = deserialize_nhdp(nhdpFile);
    StateMachine *sM = getHdpStateMachine3(nhdp);
    PairwiseAlignmentParameters *p = makeParams(rp);
    int64_t lX = sequence_correctSeqLength(strlen(refSeq), event);
    Sequence *sX = sequence_construct2(lX, (void *) refSeq, sequence_getKmer3, sequence_sliceNucleotideSequence2);
    Sequence *sY = sequence_construct2(lY, (void *) events, sequence_getEvent, sequence_sliceEventSequence2);
    stList *anchorList = makeAnchors(anchors, nAnchors);
    g_totals = totalsOut; g_totalsLen = totalsLen;
    if (totalsOut) for (int64_t i = 0; i < totalsLen; i++) totalsOut[i] = NAN;
    stList *pairs = getAlignedPairsUsingAnchors(sM, sX, sY, anchorList, p,
                                                totalsOut ? recordingPosteriorFn : diagonalCalculationPosteriorMatchProbs,
                                                raggedLeft, raggedRight);
    g_totals = NULL; g_totalsLen = 0;
    int64_t n = drainPairs(pairs, out, cap);
    stList_destruct(anchorList);
    sequence_sequenceDestroy(sX);
    sequence_sequenceDestroy(sY);
    pairwiseAlignmentBandingParameters_destruct(p);
    freeStateMachine(sM);
    destroy_nanopore_hdp(nhdp);
    return n;
}

/* getExpectationsUsingAnchors with an HdpHmm (vanillaAlign.c:318-360, 675): expOut = 9 transition sums (from * 3 + to,
 * pseudocount included) + likelihood; asgOut = (k-mer position, event index) of every assignment in list order */
int64_t ref_hdp_expectations(const char *nhdpFile, const char *refSeq, const double *events, int64_t lY,
                             const int64_t *anchors, int64_t nAnchors, const RefParams *rp,
                             int raggedLeft, int raggedRight, double pseudocount, double threshold,
                             double *expOut, int64_t *asgOut, int64_t cap) {
    NanoporeHDP *nhdp = deserialize_nhdp(nhdpFile);
    StateMachine *sM = getHdpStateMachine3(nhdp);
    PairwiseAlignmentParameters *p = makeParams(rp);
    int64_t lX = sequence_correctSeqLength(strlen(refSeq), event);
    Sequence *sX = sequence_construct2(lX, (void *) refSeq, sequence_getKmer3, sequence_sliceNucleotideSequence2);
    Sequence *sY = sequence_construct2(lY, (void *) events, sequence_getEvent, sequence_sliceEventSequence2);
    stList *anchorList = makeAnchors(anchors, nAnchors);
    Hmm *hmm = hmmContinuous_getEmptyHmm(threeStateHdp, pseudocount, threshold);
    getExpectationsUsingAnchors(sM, hmm, sX, sY, anchorList, p, diagonalCalculation_Expectations, raggedLeft, raggedRight);
    HdpHmm *hh = (HdpHmm *) hmm;
    for (int i = 0; i < 9; i++) expOut[i] = hh->transitions[i];
    expOut[9] = hmm->likelihood;
    int64_t n = hmmContinuous_howManyAssignments(hmm);
    for (int64_t i = 0; i < n && i < cap; i++) {
        asgOut[2 * i] = (const char *) stList_get(hh->kmerAssignments, i) - refSeq;
        asgOut[2 * i + 1] = ((const double *) stList_get(hh->eventAssignments, i) - events) / NB_EVENT_PARAMS;
    }
    stList_destruct(anchorList);
    sequence_sequenceDestroy(sX);
    sequence_sequenceDestroy(sY);
    pairwiseAlignmentBandingParameters_destruct(p);
    freeStateMachine(sM);
    destroy_nanopore_hdp(nhdp);
    return n;
}
#endif
