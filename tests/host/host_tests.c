/*
 * host_tests.c -- C tests of libcpecan_host.so written against include/cpecan_host.h the way the reference's CuTest
 * suites are written against cPecanLib.a:
 *   cpu  : tests/pairwiseAlignerTest.c:22-64 (test_diagonal), :74-137 (test_bands), :139-149 (test_logAdd),
 *          :596-665 (test_getSplitPoints); tests/signalPairwiseTest.c:1007-1040 (model scaling),
 *          :1461-1542 (ContinuousPairHmm write / load / normalise)
 *   gpu  : tests/signalPairwiseTest.c:580-685 (tiny strawMan known answer), :795-897 (tiny vanilla),
 *          :1116-1183 (strawMan banded on the fixture read: 987 pairs, 986 un-banded), :1250-1310 (vanilla: 999),
 *          :1604-1714 (EM: likelihood does not get worse)
 * usage: host_tests cpu|gpu <golden dir> <models dir>
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "cpecan_host.h"

static int failures = 0, checks = 0;
#define CHECK(c) do { checks++; if (!(c)) { failures++; fprintf(stderr, "FAIL %s:%d: %s\n", __FILE__, __LINE__, #c); } } while (0)

static char *readLine(const char *path) {
    FILE *f = fopen(path, "r");
    if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(2); }
    size_t cap = 1 << 16, n = 0; char *s = malloc(cap); int c;
    while ((c = fgetc(f)) != EOF && c != '\n') { if (n + 2 > cap) s = realloc(s, cap *= 2); s[n++] = (char) c; }
    s[n] = 0; fclose(f); return s;
}

static void test_diagonal(void) {
    Diagonal d = diagonal_construct(3, -1, 1);
    CHECK(diagonal_getXay(d) == 3 && diagonal_getMinXmy(d) == -1 && diagonal_getMaxXmy(d) == 1 && diagonal_getWidth(d) == 2);
    CHECK(diagonal_getXCoordinate(4, 0) == 2 && diagonal_getYCoordinate(4, 0) == 2);
    CHECK(diagonal_getXCoordinate(3, 1) == 2 && diagonal_getYCoordinate(3, 1) == 1);
    CHECK(diagonal_equals(d, diagonal_construct(3, -1, 1)) && !diagonal_equals(d, diagonal_construct(3, -1, 3)));
}

static void test_bands(void) {
    stList *a = stList_construct3(0, (void (*)(void *)) stIntTuple_destruct);
    stList_append(a, stIntTuple_construct2(1, 0)); stList_append(a, stIntTuple_construct2(2, 1)); stList_append(a, stIntTuple_construct2(3, 3));
    Band *band = band_construct(a, 6, 5, 2);
    BandIterator *it = bandIterator_construct(band);
    const int64_t want[12][3] = { {0,0,0},{1,-1,1},{2,-2,2},{3,-1,3},{4,-2,4},{5,-1,3},{6,-2,4},{7,-3,3},{8,-2,2},{9,-1,3},{10,0,2},{11,1,1} };
    for (int i = 0; i < 12; i++) CHECK(diagonal_equals(bandIterator_getNext(it), diagonal_construct(want[i][0], want[i][1], want[i][2])));
    CHECK(diagonal_equals(bandIterator_getNext(it), diagonal_construct(11, 1, 1)));          /* sticks at the end */
    for (int i = 11; i >= 7; i--) CHECK(diagonal_equals(bandIterator_getPrevious(it), diagonal_construct(want[i][0], want[i][1], want[i][2])));
    for (int i = 7; i <= 10; i++) CHECK(diagonal_equals(bandIterator_getNext(it), diagonal_construct(want[i][0], want[i][1], want[i][2])));
    BandIterator *cl = bandIterator_clone(it);
    for (int i = 10; i >= 0; i--) CHECK(diagonal_equals(bandIterator_getPrevious(it), diagonal_construct(want[i][0], want[i][1], want[i][2])));
    CHECK(diagonal_equals(bandIterator_getPrevious(it), diagonal_construct(0, 0, 0)));        /* sticks at 0 */
    CHECK(diagonal_equals(bandIterator_getPrevious(cl), diagonal_construct(10, 0, 2)));
    bandIterator_destruct(cl); bandIterator_destruct(it); band_destruct(band); stList_destruct(a);
}

static void test_logAdd(void) {
    srand(1);
    for (int t = 0; t < 100000; t++) {
        double i = (rand() + 1.0) / (RAND_MAX + 2.0), j = (rand() + 1.0) / (RAND_MAX + 2.0);
        CHECK(fabs(exp(logAdd(log(i), log(j))) - (i + j)) < 0.001);
    }
    CHECK(logAdd(-INFINITY, -3.0) == -3.0 && logAdd(0.0, -7.5) == 0.0);
}

static int tupleIs4(stList *l, int64_t i, int64_t a, int64_t b, int64_t c, int64_t d) {
    stIntTuple *t = stList_get(l, i), *w = stIntTuple_construct4(a, b, c, d);
    int r = stIntTuple_equalsFn(t, w); stIntTuple_destruct(w); return r;
}
static void test_getSplitPoints(void) {
    int64_t ms = 2000 * 2000, lX = 3000, lY = 1000;
    stList *a = stList_construct3(0, (void (*)(void *)) stIntTuple_destruct);
    stList *sp = getSplitPoints(a, lX, lY, ms, 0, 0);
    CHECK(stList_length(sp) == 1 && tupleIs4(sp, 0, 0, 0, lX, lY)); stList_destruct(sp);
    lX = 20000; lY = 25000;
    sp = getSplitPoints(a, lX, lY, ms, 1, 1); CHECK(stList_length(sp) == 0); stList_destruct(sp);
    sp = getSplitPoints(a, lX, lY, ms, 1, 0); CHECK(stList_length(sp) == 1 && tupleIs4(sp, 0, 18000, 23000, lX, lY)); stList_destruct(sp);
    sp = getSplitPoints(a, lX, lY, ms, 0, 1); CHECK(stList_length(sp) == 1 && tupleIs4(sp, 0, 0, 0, 2000, 2000)); stList_destruct(sp);
    sp = getSplitPoints(a, lX, lY, ms, 0, 0);
    CHECK(stList_length(sp) == 2 && tupleIs4(sp, 0, 0, 0, 2000, 2000) && tupleIs4(sp, 1, 18000, 23000, lX, lY)); stList_destruct(sp);
    const int64_t pts[8][2] = { {2000,2000},{4002,4001},{5000,5000},{8000,6000},{9000,9000},{10000,14000},{15000,15000},{16000,16000} };
    for (int i = 0; i < 8; i++) stList_append(a, stIntTuple_construct2(pts[i][0], pts[i][1]));
    sp = getSplitPoints(a, lX, lY, ms, 0, 0);
    CHECK(stList_length(sp) == 5);
    CHECK(tupleIs4(sp, 0, 0, 0, 3001, 3001) && tupleIs4(sp, 1, 3002, 3001, 9500, 11001) && tupleIs4(sp, 2, 9501, 12000, 12001, 14500));
    CHECK(tupleIs4(sp, 3, 13000, 14501, 18000, 18001) && tupleIs4(sp, 4, 18001, 23000, 20000, 25000));
    stList_destruct(sp); stList_destruct(a);
}

static void test_filterToRemoveOverlap(void) {
    const int64_t in[7][2] = { {0,0},{1,1},{1,2},{3,2},{4,5},{5,4},{6,6} };
    stList *l = stList_construct3(0, (void (*)(void *)) stIntTuple_destruct);
    for (int i = 0; i < 7; i++) stList_append(l, stIntTuple_construct2(in[i][0], in[i][1]));
    stList *f = filterToRemoveOverlap(l);
    int64_t px = -1, py = -1;
    for (int64_t i = 0; i < stList_length(f); i++) {
        stIntTuple *t = stList_get(f, i);
        CHECK(stIntTuple_get(t, 0) > px && stIntTuple_get(t, 1) > py);
        px = stIntTuple_get(t, 0); py = stIntTuple_get(t, 1);
    }
    /* what the reference's filterToRemoveOverlap returns for this input (checked against oracle/_ref) */
    CHECK(stList_length(f) == 2 && stIntTuple_get(stList_get(f, 0), 0) == 0 && stIntTuple_get(stList_get(f, 1), 0) == 6);
    stList_destruct(f); stList_destruct(l);
}

static void test_scaleModel(const char *models) {            /* tests/signalPairwiseTest.c:1007-1040 */
    char path[1024]; snprintf(path, sizeof path, "%s/template_median68pA.model", models);
    StateMachine *a = getStrawManStateMachine3(path), *b = getStrawManStateMachine3(path);
    emissions_signal_scaleModel(b, 1.2, 3.0, 1.1, 0.9, 1.3);
    for (int64_t i = 1; i < 1 + NUM_OF_KMERS * MODEL_PARAMS; i += MODEL_PARAMS) {
        CHECK(b->EMISSION_MATCH_PROBS[i] == a->EMISSION_MATCH_PROBS[i] * 1.2 + 3.0);
        CHECK(b->EMISSION_MATCH_PROBS[i + 1] == a->EMISSION_MATCH_PROBS[i + 1] * 1.1);
        CHECK(b->EMISSION_MATCH_PROBS[i + 2] == a->EMISSION_MATCH_PROBS[i + 2] * 0.9);
        CHECK(b->EMISSION_MATCH_PROBS[i + 4] == a->EMISSION_MATCH_PROBS[i + 4] * 1.3);
        CHECK(b->EMISSION_GAP_Y_PROBS[i] == a->EMISSION_GAP_Y_PROBS[i]);                    /* gap-Y stays unscaled */
    }
    /* emissions_signal_scaleModelNoiseOnly (impl/stateMachine.c:653-673): everything but the level mean */
    StateMachine *c = getStrawManStateMachine3(path);
    emissions_signal_scaleModelNoiseOnly(c, 1.2, 3.0, 1.1, 0.9, 1.3);
    for (int64_t i = 1; i < 1 + NUM_OF_KMERS * MODEL_PARAMS; i += MODEL_PARAMS) {
        CHECK(c->EMISSION_MATCH_PROBS[i] == a->EMISSION_MATCH_PROBS[i]);
        CHECK(c->EMISSION_MATCH_PROBS[i + 1] == b->EMISSION_MATCH_PROBS[i + 1] && c->EMISSION_MATCH_PROBS[i + 3] == b->EMISSION_MATCH_PROBS[i + 3]);
    }
    stateMachine_destruct(c);
    { char nts[] = "ACGT"; CHECK(*(char *) sequence_getBase(nts, 2) == 'G' && *(char *) sequence_getBase(nts, -1) == 'n'); }
    CHECK(a->EMISSION_GAP_X_PROBS[17] == -2.3025850929940455 && a->type == threeState && a->stateNumber == 3);
    CHECK(a->startStateProb(a, match) == 0.0 && a->endStateProb(a, shortGapY) == ((StateMachine3 *) a)->TRANSITION_MATCH_FROM_GAP_Y);
    stateMachine_destruct(a); stateMachine_destruct(b);
}

static void test_hmm_container(const char *golden) {         /* tests/signalPairwiseTest.c:1461-1542 */
    Hmm *h = hmmContinuous_getEmptyHmm(threeState, 0.0, 0.0);
    char gp[1024]; snprintf(gp, sizeof gp, "%s/zymo_three_e20.expectations", golden);
    FILE *f = fopen(gp, "r"); CHECK(f != NULL);
    int type; long long sn, ss; CHECK(fscanf(f, "%d %lld %lld", &type, &sn, &ss) == 3 && type == 2 && sn == 3 && ss == 4096);
    for (int i = 0; i < 9; i++) { double v; CHECK(fscanf(f, "%lf", &v) == 1); h->setTransitionFcn(h, i / 3, i % 3, v); }
    CHECK(fscanf(f, "%lf", &h->likelihood) == 1);
    for (int i = 0; i < 4096; i++) { double v; CHECK(fscanf(f, "%lf", &v) == 1); h->setEmissionExpectationFcn(h, 0, i, 0, v); }
    fclose(f);
    const char *out = "/tmp/cpecan_host_test.hmm";
    hmmContinuous_writeToFile(out, h, threeState);
    char cmd[2200]; snprintf(cmd, sizeof cmd, "cmp -s %s %s", out, gp);
    CHECK(system(cmd) == 0);                                                                 /* byte-identical file */
    hmmContinuous_normalize(h, threeState);
    for (int from = 0; from < 3; from++) {
        double t = 0; for (int to = 0; to < 3; to++) t += h->getTransitionsExpFcn(h, from, to);
        CHECK(fabs(t - 1.0) < 1e-12);
    }
    double t = 0; for (int i = 0; i < 4096; i++) t += h->getEmissionExpFcn(h, 0, i, 0);
    CHECK(fabs(t - 1.0) < 1e-9);
    hmmContinuous_writeToFile(out, h, threeState);
    hmmContinuous_destruct(h, threeState);
}

/* ------------------------------------------------------------------------------------------------------- GPU */
static double tinyEvents[21] = { 58.743435, 0.887833, 0.0571, 53.604965, 0.816836, 0.0571, 58.432015, 0.735143, 0.0571,
                                 63.684352, 0.795437, 0.0571, 58.921430, 0.812959, 0.0571, 59.895882, 0.740952, 0.0571,
                                 61.684303, 0.722332, 0.0571 };

static void checkAlignedPairs(stList *pairs, int64_t lX, int64_t lY) {   /* tests/signalPairwiseTest.c:46-71 */
    for (int64_t i = 0; i < stList_length(pairs); i++) {
        stIntTuple *t = stList_get(pairs, i);
        CHECK(stIntTuple_length(t) == 3);
        int64_t s = stIntTuple_get(t, 0), x = stIntTuple_get(t, 1), y = stIntTuple_get(t, 2);
        CHECK(s > 0 && s <= PAIR_ALIGNMENT_PROB_1 && x >= 0 && x < lX && y >= 0 && y < lY);
    }
    stList_sort(pairs, stIntTuple_cmpFn);                    /* uniqueness of (x, y) */
    for (int64_t i = 1; i < stList_length(pairs); i++) {
        stIntTuple *a = stList_get(pairs, i - 1), *b = stList_get(pairs, i);
        CHECK(!(stIntTuple_get(a, 1) == stIntTuple_get(b, 1) && stIntTuple_get(a, 2) == stIntTuple_get(b, 2)) || stIntTuple_get(a, 0) != stIntTuple_get(b, 0));
    }
}

static void test_tiny(const char *models) {
    char path[1024]; snprintf(path, sizeof path, "%s/template_median68pA.model", models);
    char *sX = "ACGATACGGACAT";
    int64_t lX = sequence_correctSeqLength(strlen(sX), event), lY = 7;
    PairwiseAlignmentParameters *p = pairwiseAlignmentBandingParameters_construct();
    {   /* strawMan: exactly 8 pairs at threshold 0.2 */
        StateMachine *sM = getStrawManStateMachine3(path);
        p->threshold = 0.2;
        stList *pairs = getAlignedPairsWithoutBanding(sM, sX, tinyEvents, lX, lY, p, sequence_getKmer, sequence_getEvent,
                                                      diagonalCalculationPosteriorMatchProbs, 0, 0);
        const int64_t want[8][2] = { {0,0},{1,1},{2,2},{3,3},{4,3},{5,4},{6,5},{7,6} };
        CHECK(stList_length(pairs) == 8);
        for (int i = 0; i < 8 && i < stList_length(pairs); i++) {
            int found = 0;
            for (int64_t j = 0; j < stList_length(pairs); j++) { stIntTuple *t = stList_get(pairs, j); found |= stIntTuple_get(t, 1) == want[i][0] && stIntTuple_get(t, 2) == want[i][1]; }
            CHECK(found);
        }
        checkAlignedPairs(pairs, lX, lY);
        stList_destruct(pairs); stateMachine_destruct(sM);
    }
    {   /* vanilla: exactly 5 pairs at threshold 0.5 */
        StateMachine *sM = getSignalStateMachine3Vanilla(path);
        p->threshold = 0.5;
        stList *pairs = getAlignedPairsWithoutBanding(sM, sX, tinyEvents, lX, lY, p, sequence_getKmer2, sequence_getEvent,
                                                      diagonalCalculationPosteriorMatchProbs, 0, 0);
        const int64_t want[5][2] = { {2,0},{3,3},{5,4},{6,5},{7,6} };
        CHECK(stList_length(pairs) == 5);
        for (int i = 0; i < 5; i++) {
            int found = 0;
            for (int64_t j = 0; j < stList_length(pairs); j++) { stIntTuple *t = stList_get(pairs, j); found |= stIntTuple_get(t, 1) == want[i][0] && stIntTuple_get(t, 2) == want[i][1]; }
            CHECK(found);
        }
        stList_destruct(pairs); stateMachine_destruct(sM);
    }
    pairwiseAlignmentBandingParameters_destruct(p);
}

static stList *loadAnchors(const char *golden) {
    char path[1024]; snprintf(path, sizeof path, "%s/zymo_anchors_template.txt", golden);
    FILE *f = fopen(path, "r"); if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(2); }
    stList *l = stList_construct3(0, (void (*)(void *)) stIntTuple_destruct);
    long long x, y;
    while (fscanf(f, "%lld %lld", &x, &y) == 2) stList_append(l, stIntTuple_construct2(x, y));
    fclose(f); return l;
}

static void test_fixture(const char *golden, const char *models) {
    char path[1024], model[1024];
    snprintf(path, sizeof path, "%s/ZymoRef.txt", golden); char *ref = readLine(path);
    snprintf(path, sizeof path, "%s/ZymoC_ch_1_file1.npRead", golden);
    snprintf(model, sizeof model, "%s/template_median68pA.model", models);
    NanoporeRead *np = nanopore_loadNanoporeReadFromFile(path);
    CHECK(np->readLength == 950 && np->nbTemplateEvents == 799 && np->nbComplementEvents == 670);
    int64_t lX = sequence_correctSeqLength(strlen(ref), event), lY = np->nbTemplateEvents;
    CHECK(lX == 892);
    stList *anchors = loadAnchors(golden);                   /* lastz anchors of the reference test, remapped + filtered */
    CHECK(stList_length(anchors) == 39);
    PairwiseAlignmentParameters *p = pairwiseAlignmentBandingParameters_construct();
    {   /* strawMan: 987 banded, 986 un-banded */
        StateMachine *sM = getStrawManStateMachine3(model);
        emissions_signal_scaleModel(sM, np->templateParams.scale, np->templateParams.shift, np->templateParams.var,
                                    np->templateParams.scale_sd, np->templateParams.var_sd);
        Sequence *sX = sequence_construct2(lX, ref, sequence_getKmer, sequence_sliceNucleotideSequence2);
        Sequence *sY = sequence_construct2(lY, np->templateEvents, sequence_getEvent, sequence_sliceEventSequence2);
        stList *pairs = getAlignedPairsUsingAnchors(sM, sX, sY, anchors, p, diagonalCalculationPosteriorMatchProbs, 0, 0);
        printf("strawMan banded pairs %lld (reference 987)\n", (long long) stList_length(pairs));
        CHECK(stList_length(pairs) == 987);
        checkAlignedPairs(pairs, lX, lY);
        stList_destruct(pairs);
        {   /* the callback forms, called the way the reference's tests do (tests/pairwiseAlignerTest.c:447-455): the list
             * travels as extraArgs[0]; pairs arrive in traceback order = the public entry point's list reversed */
            stList *pub = getAlignedPairsUsingAnchors(sM, sX, sY, anchors, p, diagonalCalculationPosteriorMatchProbs, 0, 0);
            stList *cb = stList_construct3(0, (void (*)(void *)) stIntTuple_destruct);
            void *extraArgs[1] = { cb };
            getPosteriorProbsWithBanding(sM, anchors, sX, sY, p, 0, 0, diagonalCalculationPosteriorMatchProbs, extraArgs);
            CHECK(stList_length(cb) == 987);
            int same = stList_length(cb) == stList_length(pub);
            for (int64_t i = 0; same && i < stList_length(cb); i++) {
                stIntTuple *a = stList_get(cb, i), *b = stList_get(pub, stList_length(pub) - 1 - i);
                for (int k = 0; k < 3; k++) if (stIntTuple_get(a, k) != stIntTuple_get(b, k)) same = 0;
            }
            CHECK(same);
            stList *cb2 = stList_construct3(0, (void (*)(void *)) stIntTuple_destruct);
            void *extraArgs2[1] = { cb2 };
            PairwiseAlignmentParameters q = *p;
            q.splitMatrixBiggerThanThis = 100 * 100;             /* several regions */
            getPosteriorProbsWithBandingSplittingAlignmentsByLargeGaps(sM, anchors, sX, sY, &q, 0, 0, diagonalCalculationPosteriorMatchProbs,
                                                                       NULL, extraArgs2);
            stList *pub2 = getAlignedPairsUsingAnchors(sM, sX, sY, anchors, &q, diagonalCalculationPosteriorMatchProbs, 0, 0);
            printf("split into regions: %lld pairs by callbacks, %lld by the entry point\n", (long long) stList_length(cb2), (long long) stList_length(pub2));
            CHECK(stList_length(cb2) == stList_length(pub2) && stList_length(cb2) > 0);
            stList_destruct(cb); stList_destruct(cb2); stList_destruct(pub); stList_destruct(pub2);
        }
        pairs = getAlignedPairsWithoutBanding(sM, ref, np->templateEvents, lX, lY, p, sequence_getKmer, sequence_getEvent,
                                              diagonalCalculationPosteriorMatchProbs, 0, 0);
        printf("strawMan un-banded pairs %lld (reference 986)\n", (long long) stList_length(pairs));
        CHECK(stList_length(pairs) == 986);
        {   /* the reference's order: ascending diagonals, x ascending inside a diagonal (impl/pairwiseAligner.c:1560-1562) */
            int ordered = 1;
            for (int64_t i = 1; i < stList_length(pairs); i++) {
                stIntTuple *a = stList_get(pairs, i - 1), *b = stList_get(pairs, i);
                const int64_t da = stIntTuple_get(a, 1) + stIntTuple_get(a, 2), db = stIntTuple_get(b, 1) + stIntTuple_get(b, 2);
                if (db < da || (db == da && stIntTuple_get(b, 1) <= stIntTuple_get(a, 1))) ordered = 0;
            }
            CHECK(ordered);
        }
        stList_destruct(pairs);
        /* EM, tests/signalPairwiseTest.c:1604-1714: start from an uninformed HMM (the reference randomises it), then
         * expectations -> normalise -> load into the state machine; the likelihood must not get worse (5 % slack) */
        {
            Hmm *h0 = hmmContinuous_getEmptyHmm(threeState, 0.0, 0.0);
            for (int from = 0; from < 3; from++) for (int to = 0; to < 3; to++) h0->setTransitionFcn(h0, from, to, 1.0 / 3.0);
            for (int k = 0; k < NUM_OF_KMERS; k++) h0->setEmissionExpectationFcn(h0, 0, k, 0, 1.0 / NUM_OF_KMERS);
            hmmContinuous_writeToFile("/tmp/cpecan_host_em.hmm", h0, threeState);
            hmmContinuous_destruct(h0, threeState);
            hmmContinuous_loadSignalHmm("/tmp/cpecan_host_em.hmm", sM, threeState);
        }
        double pLik = -INFINITY;
        for (int iter = 0; iter < 4; iter++) {
            Hmm *hmm = hmmContinuous_getEmptyHmm(threeState, 0.0001, 0.0);
            getExpectationsUsingAnchors(sM, hmm, sX, sY, anchors, p, diagonalCalculation_Expectations, 0, 0);
            printf("EM iteration %d likelihood %f\n", iter, hmm->likelihood);
            CHECK(isfinite(hmm->likelihood));
            CHECK(pLik <= hmm->likelihood * 0.95);                        /* the reference's own assertion */
            pLik = hmm->likelihood;
            hmmContinuous_normalize(hmm, threeState);
            hmmContinuous_writeToFile("/tmp/cpecan_host_em.hmm", hmm, threeState);
            hmmContinuous_destruct(hmm, threeState);
            hmmContinuous_loadSignalHmm("/tmp/cpecan_host_em.hmm", sM, threeState);
        }
        sequence_sequenceDestroy(sX); sequence_sequenceDestroy(sY); stateMachine_destruct(sM);
    }
    {   /* vanilla: 999 banded */
        StateMachine *sM = getSignalStateMachine3Vanilla(model);
        emissions_signal_scaleModel(sM, np->templateParams.scale, np->templateParams.shift, np->templateParams.var,
                                    np->templateParams.scale_sd, np->templateParams.var_sd);
        stateMachine3Vanilla_setStrandTransitionsToDefaults(sM, template);
        Sequence *sX = sequence_construct2(lX, ref, sequence_getKmer2, sequence_sliceNucleotideSequence2);
        Sequence *sY = sequence_construct2(lY, np->templateEvents, sequence_getEvent, sequence_sliceEventSequence2);
        stList *pairs = getAlignedPairsUsingAnchors(sM, sX, sY, anchors, p, diagonalCalculationPosteriorMatchProbs, 0, 0);
        printf("vanilla banded pairs %lld (reference 999)\n", (long long) stList_length(pairs));
        CHECK(stList_length(pairs) == 999);
        checkAlignedPairs(pairs, lX, lY);
        stList_destruct(pairs);
        Hmm *hmm = hmmContinuous_getEmptyHmm(vanilla, 0.0001, 0.0);
        vanillaHmm_implantMatchModelsintoHmm(sM, hmm);
        getExpectationsUsingAnchors(sM, hmm, sX, sY, anchors, p, diagonalCalculation_Expectations, 0, 0);
        CHECK(isfinite(hmm->likelihood) && hmm->getTransitionsExpFcn(hmm, 0, 0) > 0.0001);
        hmmContinuous_destruct(hmm, vanilla);
        sequence_sequenceDestroy(sX); sequence_sequenceDestroy(sY); stateMachine_destruct(sM);
    }
    {   /* fourState: 988 banded and 988 un-banded, ragged (1,1) (tests/signalPairwiseTest.c:1199-1236) */
        StateMachine *sM = getStateMachine4(model);
        emissions_signal_scaleModel(sM, np->templateParams.scale, np->templateParams.shift, np->templateParams.var,
                                    np->templateParams.scale_sd, np->templateParams.var_sd);
        Sequence *sX = sequence_construct2(lX, ref, sequence_getKmer, sequence_sliceNucleotideSequence2);
        Sequence *sY = sequence_construct2(lY, np->templateEvents, sequence_getEvent, sequence_sliceEventSequence2);
        stList *pairs = getAlignedPairsUsingAnchors(sM, sX, sY, anchors, p, diagonalCalculationPosteriorMatchProbs, 1, 1);
        printf("fourState banded pairs %lld (reference 988)\n", (long long) stList_length(pairs));
        CHECK(stList_length(pairs) == 988);
        checkAlignedPairs(pairs, lX, lY);
        stList_destruct(pairs);
        pairs = getAlignedPairsWithoutBanding(sM, ref, np->templateEvents, lX, lY, p, sequence_getKmer, sequence_getEvent,
                                              diagonalCalculationPosteriorMatchProbs, 1, 1);
        CHECK(stList_length(pairs) == 988);
        stList_destruct(pairs);
        sequence_sequenceDestroy(sX); sequence_sequenceDestroy(sY); stateMachine_destruct(sM);
    }
    {   /* echelon: 857 banded at threshold 0.15 with the padded sequence and the multi-state posterior (:1388-1449) */
        StateMachine *sM = getStateMachineEchelon(model);
        emissions_signal_scaleModel(sM, np->templateParams.scale, np->templateParams.shift, np->templateParams.var,
                                    np->templateParams.scale_sd, np->templateParams.var_sd);
        Sequence *sX = sequence_construct2(lX, ref, sequence_getKmer2, sequence_sliceNucleotideSequence2);
        sequence_padSequence(sX);
        Sequence *sY = sequence_construct2(lY, np->templateEvents, sequence_getEvent, sequence_sliceEventSequence2);
        p->threshold = 0.15;
        stList *pairs = getAlignedPairsUsingAnchors(sM, sX, sY, anchors, p, diagonalCalculationMultiPosteriorMatchProbs, 0, 0);
        printf("echelon banded pairs %lld (reference 857)\n", (long long) stList_length(pairs));
        CHECK(stList_length(pairs) == 857);
        stList_destruct(pairs);
        p->threshold = 0.01;
        free(sX->elements); sequence_sequenceDestroy(sX); sequence_sequenceDestroy(sY); stateMachine_destruct(sM);
    }
    /* two reads in one GPU batch == two calls */
    {
        StateMachine *sM = getStrawManStateMachine3(model);
        emissions_signal_scaleModel(sM, np->templateParams.scale, np->templateParams.shift, np->templateParams.var,
                                    np->templateParams.scale_sd, np->templateParams.var_sd);
        Sequence *sX = sequence_construct2(lX, ref, sequence_getKmer, sequence_sliceNucleotideSequence2);
        Sequence *sY = sequence_construct2(lY, np->templateEvents, sequence_getEvent, sequence_sliceEventSequence2);
        StateMachine *sMs[2] = { sM, sM }; Sequence *xs[2] = { sX, sX }, *ys[2] = { sY, sY }; stList *as[2] = { anchors, anchors }, *res[2];
        getAlignedPairsUsingAnchorsBatch(2, sMs, xs, ys, as, p, 0, 0, res);
        CHECK(stList_length(res[0]) == stList_length(res[1]) && stList_length(res[0]) == 987);
        stList_destruct(res[0]); stList_destruct(res[1]);
        sequence_sequenceDestroy(sX); sequence_sequenceDestroy(sY); stateMachine_destruct(sM);
    }
    pairwiseAlignmentBandingParameters_destruct(p); stList_destruct(anchors); nanopore_nanoporeReadDestruct(np); free(ref);
}

/* the caller-side helpers vanillaAlign.c uses around the alignment call (cigar -> anchors, padding, strings, HMM fields) */
static void test_caller_side(void) {
    FILE *f = fopen("/tmp/cpecan_host_test.cigar", "w");
    fprintf(f, "cigar: read 3 14 + ref 10 22 + 77 M 4 D 2 M 3 I 1 M 3\n");
    fclose(f);
    f = fopen("/tmp/cpecan_host_test.cigar", "r");
    struct PairwiseAlignment *pA = cigarRead(f);
    fclose(f);
    CHECK(pA != NULL && pA->start1 == 10 && pA->end1 == 22 && pA->strand1 == 1 && pA->start2 == 3 && pA->end2 == 14 && pA->strand2 == 1);
    CHECK(strcmp(pA->contig1, "ref") == 0 && strcmp(pA->contig2, "read") == 0 && pA->operationList->length == 5);
    checkPairwiseAlignment(pA);
    stList *an = convertPairwiseForwardStrandAlignmentToAnchorPairs(pA, 1);       /* trim 1 at both ends of every match */
    /* M4 at (10,3): (11,4),(12,5); D2: x += 2; M3 at (16,7): (17,8); I1: y += 1; M3 at (19,11): (20,12) */
    const int64_t want[4][2] = { {11,4},{12,5},{17,8},{20,12} };
    CHECK(stList_length(an) == 4);
    for (int i = 0; i < 4 && i < stList_length(an); i++) { stIntTuple *t = stList_get(an, i); CHECK(stIntTuple_get(t, 0) == want[i][0] && stIntTuple_get(t, 1) == want[i][1]); }
    stList_destruct(an); destructPairwiseAlignment(pA);
    char *rc = stString_reverseComplementString("AACGTn");
    CHECK(strcmp(rc, "nACGTT") == 0); free(rc);
    char *rp = stString_replace("ACCGC", "C", "E"); CHECK(strcmp(rp, "AEEGE") == 0); free(rp);
    char *sub = stString_getSubString("ABCDEFG", 2, 3); CHECK(strcmp(sub, "CDE") == 0); free(sub);
    char *pr = stString_print("%s-%d", "x", 7); CHECK(strcmp(pr, "x-7") == 0); free(pr);
    char seq[] = "ACGTACGTAC";
    Sequence *sX = sequence_construct2(5, seq, sequence_getKmer2, sequence_sliceNucleotideSequence2);
    sequence_padSequence(sX);
    CHECK(strlen((char *) sX->elements) == 40 && ((char *) sX->elements)[10] == 'n' && strncmp((char *) sX->elements, seq, 10) == 0);
    CHECK(sequence_getKmer3(sX->elements, -1) == sX->elements && (char *) sequence_getKmer3(sX->elements, 2) == (char *) sX->elements + 2);
    free(sX->elements); sequence_sequenceDestroy(sX);
    CHECK(emissions_discrete_getKmerIndexFromKmer("ACGTACGGGG") == emissions_discrete_getKmerIndex("ACGTAC"));
    Hmm *h = hmmContinuous_getEmptyHmm(threeState, 0.5, 0.0);
    continuousPairHmm_addToTransitionsExpectation(h, 0, 1, 2.0);
    CHECK(continuousPairHmm_getTransitionExpectation(h, 0, 1) == 2.5);
    continuousPairHmm_setKmerGapExpectation(h, 0, 17, 0, 3.0);
    CHECK(continuousPairHmm_getKmerGapExpectation(h, 0, 17, 0) == 3.0);
    continuousPairHmm_normalize(h);
    CHECK(fabs(continuousPairHmm_getTransitionExpectation(h, 0, 0) + continuousPairHmm_getTransitionExpectation(h, 0, 1) + continuousPairHmm_getTransitionExpectation(h, 0, 2) - 1.0) < 1e-12);
    continuousPairHmm_destruct(h);
    Hmm *v = hmmContinuous_getEmptyHmm(vanilla, 0.0, 0.0);
    vanillaHmm_addToKmerSkipBinExpectation(v, 3, 0, 1.5);
    CHECK(vanillaHmm_getKmerSkipBinExpectation(v, 3, 0) == 1.5);
    vanillaHmm_destruct(v);
    double c1[3] = { -1.0, -2.0, -INFINITY }, c2[3] = { -0.5, -0.25, -3.0 };
    CHECK(fabs(cell_dotProduct(c1, c2, 3) - logAdd(logAdd(-1.5, -2.25), -INFINITY)) < 1e-12);
    /* stList_sort: qsort over the tuples */
    stList *l = stList_construct3(0, (void (*)(void *)) stIntTuple_destruct);
    stList_append(l, stIntTuple_construct3(5, 4, 4)); stList_append(l, stIntTuple_construct3(9, 1, 1)); stList_append(l, stIntTuple_construct3(7, 2, 3));
    stList_sort(l, sortByXPlusYCoordinate2);
    CHECK(stIntTuple_get(stList_get(l, 0), 1) == 1 && stIntTuple_get(stList_get(l, 2), 1) == 4);
    stList_destruct(l);
}

int main(int argc, char **argv) {
    if (argc < 4) { fprintf(stderr, "usage: %s cpu|gpu <golden dir> <models dir>\n", argv[0]); return 2; }
    if (!strcmp(argv[1], "cpu")) {
        test_diagonal(); test_bands(); test_logAdd(); test_getSplitPoints(); test_filterToRemoveOverlap();
        test_scaleModel(argv[3]); test_hmm_container(argv[2]); test_caller_side();
    } else {
        test_tiny(argv[3]); test_fixture(argv[2], argv[3]);
    }
    printf("%d checks, %d failures\n", checks, failures);
    return failures ? 1 : 0;
}
