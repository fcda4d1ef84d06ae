// cpecan_host.cpp -- libcpecan_host.so: the reference's C API for the signal hot path (include/cpecan_host.h) on top
// of the CUDA C-ABI (include/cpecan_cuda.h).  Host work here is the integer / container / file-format code either
// side of the DP (anchors, band geometry, split points, state-machine and HMM containers, .npRead); every cell of the
// DP is computed by libcpecan_cuda.so.  There is no CPU fallback: without a usable CUDA device the alignment entry
// points abort exactly as the reference's st_errAbort does.
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <mutex>
#include <string>
#include <vector>

#include "cpecan_cuda.h"
#include "cpecan_host.h"

// ===================================================================================================== sonLib subset
struct _stList {
    std::vector<void *> v;
    void (*destruct)(void *) = nullptr;
};

extern "C" {

void st_errAbort(const char *format, ...) {
    va_list ap;
    va_start(ap, format);
    vfprintf(stderr, format, ap);
    va_end(ap);
    fputc('\n', stderr);
    abort();
}

stList *stList_construct(void) { return new _stList(); }
stList *stList_construct3(int64_t size, void (*destructElement)(void *)) {
    stList *l = new _stList();
    l->v.assign((size_t) size, nullptr);
    l->destruct = destructElement;
    return l;
}
void stList_destruct(stList *list) {
    if (!list) return;
    if (list->destruct) for (void *p : list->v) if (p) list->destruct(p);
    delete list;
}
int64_t stList_length(stList *list) { return list ? (int64_t) list->v.size() : 0; }
void *stList_get(stList *list, int64_t index) {
    if (index < 0 || index >= (int64_t) list->v.size()) st_errAbort("stList_get: index %lld out of bounds", (long long) index);
    return list->v[(size_t) index];
}
void stList_set(stList *list, int64_t index, void *item) {
    if (index < 0 || index >= (int64_t) list->v.size()) st_errAbort("stList_set: index %lld out of bounds", (long long) index);
    list->v[(size_t) index] = item;
}
void stList_append(stList *list, void *item) { list->v.push_back(item); }
void *stList_pop(stList *list) {
    if (list->v.empty()) st_errAbort("stList_pop: empty list");
    void *p = list->v.back();
    list->v.pop_back();
    return p;
}
// sonLib sorts the pointer array with qsort through a file-scope comparator; the same here (thread-local, so the two
// OpenMP sections of vanillaAlign.c may sort at once)
static thread_local int (*tlsCmp)(const void *, const void *) = nullptr;
static int sortTrampoline(const void *a, const void *b) { return tlsCmp(*(void *const *) a, *(void *const *) b); }
void stList_sort(stList *list, int (*cmpFn)(const void *a, const void *b)) {
    tlsCmp = cmpFn;
    if (!list->v.empty()) qsort(list->v.data(), list->v.size(), sizeof(void *), sortTrampoline);
}
void stList_setDestructor(stList *list, void (*destructElement)(void *)) { list->destruct = destructElement; }

static stIntTuple *tuple_n(int64_t n, const int64_t *vals) {
    int64_t *t = (int64_t *) malloc(sizeof(int64_t) * (size_t) (n + 1));
    t[0] = n;
    for (int64_t i = 0; i < n; i++) t[i + 1] = vals[i];
    return t;
}
stIntTuple *stIntTuple_construct2(int64_t a, int64_t b) { int64_t v[2] = { a, b }; return tuple_n(2, v); }
stIntTuple *stIntTuple_construct3(int64_t a, int64_t b, int64_t c) { int64_t v[3] = { a, b, c }; return tuple_n(3, v); }
stIntTuple *stIntTuple_construct4(int64_t a, int64_t b, int64_t c, int64_t d) { int64_t v[4] = { a, b, c, d }; return tuple_n(4, v); }
void stIntTuple_destruct(stIntTuple *t) { free(t); }
int64_t stIntTuple_length(stIntTuple *t) { return t[0]; }
int64_t stIntTuple_get(stIntTuple *t, int64_t index) {
    if (index < 0 || index >= t[0]) st_errAbort("stIntTuple_get: index %lld out of bounds", (long long) index);
    return t[index + 1];
}
int stIntTuple_cmpFn(const void *a, const void *b) {          // lexicographic, shorter first on a tie
    const int64_t *x = (const int64_t *) a, *y = (const int64_t *) b;
    const int64_t n = std::min(x[0], y[0]);
    for (int64_t i = 1; i <= n; i++) if (x[i] != y[i]) return x[i] < y[i] ? -1 : 1;
    return x[0] == y[0] ? 0 : (x[0] < y[0] ? -1 : 1);
}
int stIntTuple_equalsFn(const void *a, const void *b) { return stIntTuple_cmpFn(a, b) == 0; }

// ====================================================================================================== Sequence
// impl/pairwiseAligner.c:261-354
static double NULLEVENT_[2] = { -INFINITY, 0.0 };
static char NCHAR_[2] = "n";

Sequence *sequence_construct2(int64_t length, void *elements, void *(*getFcn)(void *, int64_t),
                              Sequence *(*sliceFcn)(Sequence *, int64_t, int64_t)) {
    Sequence *s = (Sequence *) malloc(sizeof(Sequence));
    s->length = length; s->elements = elements; s->get = getFcn; s->sliceFcn = sliceFcn;
    return s;
}
Sequence *sequence_construct(int64_t length, void *elements, void *(*getFcn)(void *, int64_t)) {
    return sequence_construct2(length, elements, getFcn, nullptr);
}
Sequence *sequence_sliceNucleotideSequence2(Sequence *in, int64_t start, int64_t sliceLength) {
    return sequence_construct2(sliceLength, (char *) in->elements + start, in->get, in->sliceFcn);
}
Sequence *sequence_sliceEventSequence2(Sequence *in, int64_t start, int64_t sliceLength) {
    return sequence_construct2(sliceLength, (double *) in->elements + start * NB_EVENT_PARAMS, in->get, in->sliceFcn);
}
void sequence_sequenceDestroy(Sequence *seq) { free(seq); }
void *sequence_getKmer(void *elements, int64_t index) { return index >= 0 ? (void *) ((char *) elements + index) : (void *) NCHAR_; }
void *sequence_getKmer2(void *elements, int64_t index) { return (char *) elements + (index > 0 ? index - 1 : 0); }
void *sequence_getBase(void *elements, int64_t index) {                                                     // impl/pairwiseAligner.c:308-312
    static char n[] = "n";
    return index >= 0 ? (void *) ((char *) elements + index) : (void *) n;
}
void *sequence_getKmer3(void *elements, int64_t index) { return (char *) elements + (index >= 0 ? index : 0); }   // :327-331
// impl/pairwiseAligner.c:282-285: 30 'n' behind the nucleotides (the echelon machine reads k-mers past the end); as in
// the reference the old buffer is not freed -- the caller still owns it
void sequence_padSequence(Sequence *sequence) {
    const char *e = (const char *) sequence->elements;
    const size_t n = strlen(e);
    char *padded = (char *) malloc(n + 31);
    memcpy(padded, e, n);
    memset(padded + n, 'n', 30);
    padded[n + 30] = 0;
    sequence->elements = padded;
}
void *sequence_getEvent(void *elements, int64_t index) {
    return index >= 0 ? (void *) ((double *) elements + index * NB_EVENT_PARAMS) : (void *) NULLEVENT_;
}
int64_t sequence_correctSeqLength(int64_t length, SequenceType type) {
    if (length <= 0) return 0;
    return type == nucleotide ? length : length - (KMER_LENGTH - 1);
}

// ============================================================================================ banding parameters
// defaults of impl/pairwiseAligner.c:1428-1441
PairwiseAlignmentParameters *pairwiseAlignmentBandingParameters_construct(void) {
    PairwiseAlignmentParameters *p = (PairwiseAlignmentParameters *) malloc(sizeof(PairwiseAlignmentParameters));
    p->threshold = 0.01;
    p->minDiagsBetweenTraceBack = 1000;
    p->traceBackDiagonals = 40;
    p->diagonalExpansion = 20;
    p->constraintDiagonalTrim = 14;
    p->anchorMatrixBiggerThanThis = 500 * 500;
    p->repeatMaskMatrixBiggerThanThis = 500 * 500;
    p->splitMatrixBiggerThanThis = 3000 * 3000;
    p->alignAmbiguityCharacters = false;
    p->gapGamma = 0.5f;
    return p;
}
void pairwiseAlignmentBandingParameters_destruct(PairwiseAlignmentParameters *p) { free(p); }

// ================================================================================================ Diagonal / Band
// impl/pairwiseAligner.c:36-227
Diagonal diagonal_construct(int64_t xay, int64_t xmyL, int64_t xmyR) {
    if ((xay + xmyL) % 2 != 0 || (xay + xmyR) % 2 != 0 || xmyL > xmyR)
        st_errAbort("Attempt to create diagonal with invalid coordinates: xay %lld xmyL %lld xmyR %lld",
                    (long long) xay, (long long) xmyL, (long long) xmyR);
    Diagonal d = { xay, xmyL, xmyR };
    return d;
}
int64_t diagonal_getXay(Diagonal d) { return d.xay; }
int64_t diagonal_getMinXmy(Diagonal d) { return d.xmyL; }
int64_t diagonal_getMaxXmy(Diagonal d) { return d.xmyR; }
int64_t diagonal_getWidth(Diagonal d) { return (d.xmyR - d.xmyL) / 2 + 1; }
int64_t diagonal_getXCoordinate(int64_t xay, int64_t xmy) { return (xay + xmy) / 2; }
int64_t diagonal_getYCoordinate(int64_t xay, int64_t xmy) { return (xay - xmy) / 2; }
int64_t diagonal_equals(Diagonal a, Diagonal b) { return a.xay == b.xay && a.xmyL == b.xmyL && a.xmyR == b.xmyR; }

}  // extern "C" (the opaque types below are C++ classes)

struct _band { std::vector<Diagonal> diagonals; };
struct _bandIterator { Band *band; int64_t index; };

extern "C" {

// Between consecutive anchors the band is a rectangle in matrix coordinates, grown by `expansion` in x-y and clamped to
// the matrix; every diagonal keeps the cells of that rectangle with the parity of the diagonal.
Band *band_construct(stList *anchorPairs, int64_t lX, int64_t lY, int64_t expansion) {
    Band *band = new _band();
    const int64_t nA = stList_length(anchorPairs);
    auto clampTo = [](int64_t z, int64_t l) { return z < 0 ? (int64_t) 0 : (z > l ? l : z); };
    int64_t ai = 0, nxay = 0, nxmy = 0, xL = 0, yL = 0, xU = 0, yU = 0;
    for (int64_t xay = 0; xay <= lX + lY; xay++) {
        auto parity = [xay](int64_t xmy) { return (xay + xmy) % 2 == 0 ? xmy : xmy + 1; };
        int64_t l = parity(xL - yL), r = parity(xU - yU), t;
        t = (xay + l) / 2; if (t < xL) l += 2 * (xL - t);
        t = (xay - l) / 2; if (yL < t) l += 2 * (t - yL);
        t = (xay + r) / 2; if (xU < t) r -= 2 * (t - xU);
        t = (xay - r) / 2; if (t < yU) r -= 2 * (yU - t);
        band->diagonals.push_back(diagonal_construct(xay, l, r));
        if (nxay == xay) {                                    // reached the next anchor: move on to the following box
            const int64_t pxay = nxay, pxmy = nxmy;
            int64_t x = lX, y = lY;
            if (ai < nA) {
                stIntTuple *a = (stIntTuple *) stList_get(anchorPairs, ai++);
                x = stIntTuple_get(a, 0) + 1; y = stIntTuple_get(a, 1) + 1;
            }
            nxay = x + y; nxmy = x - y;
            xL = clampTo((pxay + (pxmy - expansion)) / 2, lX);
            yL = clampTo((nxay - (nxmy - expansion)) / 2, lY);
            xU = clampTo((nxay + (nxmy + expansion)) / 2, lX);
            yU = clampTo((pxay - (pxmy + expansion)) / 2, lY);
        }
    }
    return band;
}
void band_destruct(Band *band) { delete band; }
BandIterator *bandIterator_construct(Band *band) { return new _bandIterator{ band, 0 }; }
void bandIterator_destruct(BandIterator *it) { delete it; }
BandIterator *bandIterator_clone(BandIterator *it) { return new _bandIterator{ it->band, it->index }; }
Diagonal bandIterator_getNext(BandIterator *it) {              // sticks at the last diagonal
    const int64_t n = (int64_t) it->band->diagonals.size();
    const Diagonal d = it->band->diagonals[(size_t) (it->index >= n ? n - 1 : it->index)];
    if (it->index < n) it->index++;
    return d;
}
Diagonal bandIterator_getPrevious(BandIterator *it) {          // sticks at diagonal 0
    if (it->index > 0) it->index--;
    return it->band->diagonals[(size_t) it->index];
}

// ========================================================================================================= logAdd
// impl/pairwiseAligner.c:235-255: four cubic segments (float-literal coefficients), hard cut at 7.5
static double la_segment(double x) {
    if (x <= 1.00f) return ((-0.009350833524763f * x + 0.130659527668286f) * x + 0.498799810682272f) * x + 0.693203116424741f;
    if (x <= 2.50f) return ((-0.014532321752540f * x + 0.139942324101744f) * x + 0.495635523139337f) * x + 0.692140569840976f;
    if (x <= 4.50f) return ((-0.004605031767994f * x + 0.063427417320019f) * x + 0.695956496475118f) * x + 0.514272634594009f;
    return ((-0.000458661602210f * x + 0.009695946122598f) * x + 0.930734667215156f) * x + 0.168037164329057f;
}
double logAdd(double x, double y) {
    if (x < y) return (x == -INFINITY || y - x >= 7.5) ? y : la_segment(y - x) + x;
    return (y == -INFINITY || x - y >= 7.5) ? x : la_segment(x - y) + y;
}

// ======================================================================================== anchors / split points
int sortByXPlusYCoordinate(const void *i, const void *j) {    // impl/pairwiseAligner.c: pairs (x, y)
    const int64_t a = stIntTuple_get((stIntTuple *) i, 0) + stIntTuple_get((stIntTuple *) i, 1);
    const int64_t b = stIntTuple_get((stIntTuple *) j, 0) + stIntTuple_get((stIntTuple *) j, 1);
    return a == b ? 0 : (a > b ? 1 : -1);
}
int sortByXPlusYCoordinate2(const void *i, const void *j) {   // aligned pairs (score, x, y)
    const int64_t a = stIntTuple_get((stIntTuple *) i, 1) + stIntTuple_get((stIntTuple *) i, 2);
    const int64_t b = stIntTuple_get((stIntTuple *) j, 1) + stIntTuple_get((stIntTuple *) j, 2);
    return a == b ? 0 : (a > b ? 1 : -1);
}

// impl/pairwiseAligner.c:1160-1200: keep the pairs strictly increasing in x and y scanning backwards AND forwards;
// membership of the backward pass is by VALUE (a sorted set in the reference)
stList *filterToRemoveOverlap(stList *pairs) {
    const int64_t n = stList_length(pairs);
    std::vector<std::pair<int64_t, int64_t>> kept;
    int64_t pX = INT64_MAX, pY = INT64_MAX;
    for (int64_t i = n - 1; i >= 0; i--) {
        stIntTuple *t = (stIntTuple *) stList_get(pairs, i);
        const int64_t x = stIntTuple_get(t, 0), y = stIntTuple_get(t, 1);
        if (x < pX && y < pY) kept.emplace_back(x, y);
        pX = std::min(pX, x); pY = std::min(pY, y);
    }
    std::sort(kept.begin(), kept.end());
    stList *out = stList_construct3(0, (void (*)(void *)) stIntTuple_destruct);
    pX = INT64_MIN; pY = INT64_MIN;
    for (int64_t i = 0; i < n; i++) {
        stIntTuple *t = (stIntTuple *) stList_get(pairs, i);
        const int64_t x = stIntTuple_get(t, 0), y = stIntTuple_get(t, 1);
        if (x > pX && y > pY && std::binary_search(kept.begin(), kept.end(), std::make_pair(x, y)))
            stList_append(out, stIntTuple_construct2(x, y));
        pX = std::max(pX, x); pY = std::max(pY, y);
    }
    return out;
}

// impl/pairwiseAligner.c:1289-1340: a gap between anchors whose rectangle exceeds maxMatrixSize cuts the alignment;
// each side keeps min(len / 2, sqrt(maxMatrixSize)) of the gap
stList *getSplitPoints(stList *anchorPairs, int64_t lX, int64_t lY, int64_t maxMatrixSize, bool raggedLeft, bool raggedRight) {
    stList *out = stList_construct3(0, (void (*)(void *)) stIntTuple_destruct);
    int64_t x1 = 0, y1 = 0, x2 = 0, y2 = 0;
    auto step = [&](int64_t x3, int64_t y3, bool skipBlock) {
        const int64_t gx = x3 - x2, gy = y3 - y2;
        if (gx * gy <= maxMatrixSize) return false;
        const int64_t maxLen = (int64_t) std::sqrt((double) maxMatrixSize);
        const int64_t hX = std::min(gx / 2, maxLen), hY = std::min(gy / 2, maxLen);
        if (!skipBlock) stList_append(out, stIntTuple_construct4(x1, y1, x2 + hX, y2 + hY));
        x1 = x3 - hX; y1 = y3 - hY;
        return true;
    };
    const int64_t nA = stList_length(anchorPairs);
    for (int64_t i = 0; i < nA; i++) {
        stIntTuple *a = (stIntTuple *) stList_get(anchorPairs, i);
        const int64_t x3 = stIntTuple_get(a, 0), y3 = stIntTuple_get(a, 1);
        step(x3, y3, raggedLeft && i == 0);
        x2 = x3 + 1; y2 = y3 + 1;
    }
    if (!step(lX, lY, raggedLeft && nA == 0) || !raggedRight) stList_append(out, stIntTuple_construct4(x1, y1, lX, lY));
    return out;
}

// ================================================================================================ state machines
int64_t emissions_discrete_getKmerIndex(void *kmerPtr) {      // impl/stateMachine.c:104-139; > NUM_OF_KMERS: invalid
    const char *s = (const char *) kmerPtr;
    int64_t v = 0;
    for (int j = 0; j < KMER_LENGTH; j++) {
        int b;
        switch (s[j]) { case 'A': b = 0; break; case 'C': b = 1; break; case 'G': b = 2; break; case 'T': b = 3; break; default: return NUM_OF_KMERS + 1; }
        v = v * 4 + b;
    }
    return v;
}

static void noCpuPath(const char *what) {
    st_errAbort("%s: this build evaluates the DP on the GPU only (libcpecan_cuda.so); there is no CPU path", what);
}
static void cellCalculateAbort(StateMachine *, double *, double *, double *, double *, void *, void *,
                               void (*)(double *, double *, int64_t, int64_t, double, double, void *), void *) {
    noCpuPath("StateMachine.cellCalculate");
}
static void updateExpAbort(double *, double *, int64_t, int64_t, double, double, void *) {
    noCpuPath("StateMachine.cellCalculateUpdateExpectations");
}
static double sm3_start(StateMachine *, int64_t s) { return s == match ? 0.0 : -INFINITY; }                 // :1168-1172
static double sm3_raggedStart(StateMachine *, int64_t s) { return s == match ? -INFINITY : 0.0; }           // :1174-1177
static double sm3_end(StateMachine *sM, int64_t s) {                                                        // :1179-1192
    StateMachine3 *m = (StateMachine3 *) sM;
    return s == match ? m->TRANSITION_MATCH_CONTINUE : (s == shortGapX ? m->TRANSITION_MATCH_FROM_GAP_X : m->TRANSITION_MATCH_FROM_GAP_Y);
}
static double sm3_raggedEnd(StateMachine *sM, int64_t s) {                                                  // :1194-1207
    StateMachine3 *m = (StateMachine3 *) sM;
    return s == match ? (m->TRANSITION_GAP_OPEN_X + m->TRANSITION_GAP_OPEN_Y) / 2.0
                      : (s == shortGapX ? m->TRANSITION_GAP_EXTEND_X : m->TRANSITION_GAP_EXTEND_Y);
}
static double sm3v_end(StateMachine *sM, int64_t s) {                                                       // :1223-1235
    StateMachine3Vanilla *m = (StateMachine3Vanilla *) sM;
    return s == match ? m->DEFAULT_END_MATCH_PROB : (s == shortGapX ? m->DEFAULT_END_FROM_X_PROB : m->DEFAULT_END_FROM_Y_PROB);
}
static double sm3v_raggedEnd(StateMachine *sM, int64_t s) {                                                 // :1209-1221
    StateMachine3Vanilla *m = (StateMachine3Vanilla *) sM;
    return s == match ? (m->DEFAULT_END_FROM_X_PROB + m->DEFAULT_END_FROM_Y_PROB) / 2.0
                      : (s == shortGapX ? m->DEFAULT_END_FROM_X_PROB : m->DEFAULT_END_FROM_Y_PROB);
}

static const int64_t TABLE_LEN = 1 + NUM_OF_KMERS * MODEL_PARAMS;

static std::vector<double> readDoublesLine(FILE *fh, const char *file) {
    std::vector<double> v;
    std::string line;
    int c;
    while ((c = fgetc(fh)) != EOF && c != '\n') line.push_back((char) c);
    if (c == EOF && line.empty()) st_errAbort("%s: unexpected end of file", file);
    const char *p = line.c_str();
    char *end;
    for (;;) {
        const double d = strtod(p, &end);
        if (end == p) break;
        v.push_back(d);
        p = end;
    }
    return v;
}

// emissions_signal_loadPoreModel (impl/stateMachine.c:242-320): 1 + 4096*5 match parameters, 30 skip bins,
// 1 + 4096*5 extra-event parameters; same aborts on malformed files
static void loadPoreModel(StateMachine *sM, const char *modelFile) {
    FILE *fh = fopen(modelFile, "r");
    if (!fh) st_errAbort("emissions_signal_loadPoreModel: cannot open %s", modelFile);
    std::vector<double> l1 = readDoublesLine(fh, modelFile);
    if ((int64_t) l1.size() != TABLE_LEN) st_errAbort("This stateMachine is not correct for signal model (match emissions)\n");
    std::vector<double> l2 = readDoublesLine(fh, modelFile);
    if (l2.size() != 30) st_errAbort("Did not get expected number of kmer skip bins, expected 30, got %lld\n", (long long) l2.size());
    std::vector<double> l3 = readDoublesLine(fh, modelFile);
    if ((int64_t) l3.size() != TABLE_LEN) st_errAbort("This stateMachine is not correct for signal model (dupeEvent - Y emissions)\n");
    fclose(fh);
    memcpy(sM->EMISSION_MATCH_PROBS, l1.data(), sizeof(double) * TABLE_LEN);
    memcpy(sM->EMISSION_GAP_Y_PROBS, l3.data(), sizeof(double) * TABLE_LEN);
    if (sM->type == vanilla || sM->type == echelon)             // :282-294
        for (int i = 0; i < 30; i++) { sM->EMISSION_GAP_X_PROBS[i] = l2[i]; sM->EMISSION_GAP_X_PROBS[i + 30] = l2[i]; }
}

static void initBase(StateMachine *sM, StateMachineType type, int64_t nGapX) {
    sM->type = type; sM->stateNumber = 3; sM->matchState = match; sM->parameterSetSize = NUM_OF_KMERS;
    sM->EMISSION_MATCH_PROBS = (double *) calloc((size_t) TABLE_LEN, sizeof(double));
    sM->EMISSION_GAP_Y_PROBS = (double *) calloc((size_t) TABLE_LEN, sizeof(double));
    sM->EMISSION_GAP_X_PROBS = (double *) calloc((size_t) nGapX, sizeof(double));
    sM->startStateProb = sm3_start; sM->raggedStartStateProb = sm3_raggedStart;
    sM->cellCalculate = cellCalculateAbort; sM->cellCalculateUpdateExpectations = updateExpAbort;
}

void stateMachine3_setTransitionsToNanoporeDefaults(StateMachine *sM) {                                     // :1278-1289
    StateMachine3 *m = (StateMachine3 *) sM;
    m->TRANSITION_MATCH_CONTINUE = -0.23552123624314988;
    m->TRANSITION_MATCH_FROM_GAP_X = -0.21880828092192281;
    m->TRANSITION_MATCH_FROM_GAP_Y = -0.013406326748077823;
    m->TRANSITION_GAP_OPEN_X = -1.6269694202638481;
    m->TRANSITION_GAP_OPEN_Y = -4.3187242127300092;
    m->TRANSITION_GAP_EXTEND_X = -1.6269694202638481;
    m->TRANSITION_GAP_EXTEND_Y = -4.3187242127239411;
    m->TRANSITION_GAP_SWITCH_TO_X = -INFINITY;
    m->TRANSITION_GAP_SWITCH_TO_Y = -INFINITY;
}

StateMachine *getStrawManStateMachine3(const char *modelFile) {                                             // :1725-1735
    StateMachine3 *m = (StateMachine3 *) calloc(1, sizeof(StateMachine3));
    initBase(&m->model, threeState, NUM_OF_KMERS);
    m->model.endStateProb = sm3_end; m->model.raggedEndStateProb = sm3_raggedEnd;
    stateMachine3_setTransitionsToNanoporeDefaults(&m->model);
    for (int64_t i = 0; i < NUM_OF_KMERS; i++) m->model.EMISSION_GAP_X_PROBS[i] = -2.3025850929940455;       // log(0.1), :1506-1508
    loadPoreModel(&m->model, modelFile);
    return &m->model;
}

void stateMachine3Vanilla_setStrandTransitionsToDefaults(StateMachine *sM, Strand strand) {                 // :1291-1303
    StateMachine3Vanilla *m = (StateMachine3Vanilla *) sM;
    if (strand == complement) { m->TRANSITION_M_TO_Y_NOT_X = 0.14f; m->TRANSITION_E_TO_E = 0.49f; }
    else { m->TRANSITION_M_TO_Y_NOT_X = 0.17f; m->TRANSITION_E_TO_E = 0.55f; }
}

StateMachine *getSignalStateMachine3Vanilla(const char *modelFile) {                                        // :1761-1771, 1560-1600
    StateMachine3Vanilla *m = (StateMachine3Vanilla *) calloc(1, sizeof(StateMachine3Vanilla));
    initBase(&m->model, vanilla, 60);
    m->model.endStateProb = sm3v_end; m->model.raggedEndStateProb = sm3v_raggedEnd;
    m->TRANSITION_M_TO_Y_NOT_X = 0.17;
    m->TRANSITION_E_TO_E = 0.55f;
    m->DEFAULT_END_MATCH_PROB = -0.23552123624314988;
    m->DEFAULT_END_FROM_X_PROB = -1.6269694202638481;
    m->DEFAULT_END_FROM_Y_PROB = -4.3187242127300092;
    loadPoreModel(&m->model, modelFile);
    return &m->model;
}

// getStateMachine4 (impl/stateMachine.c:1750-1759, constructor :960-1011): gap-X emissions all zero, the model's
// skip-bin line is not loaded for this type
static double sm4_start(StateMachine *, int64_t st) { return st == match ? 0 : LOG_ZERO; }                     // :775-779
static double sm4_raggedStart(StateMachine *, int64_t st) { return (st == shortGapY || st == longGapX) ? 0 : LOG_ZERO; }   // :791-794
static double sm4_end(StateMachine *sM, int64_t st) {                                                       // :796-810
    StateMachine4 *m = (StateMachine4 *) sM;
    switch (st) {
        case match: return m->TRANSITION_MATCH_CONTINUE;
        case shortGapX: return m->TRANSITION_MATCH_FROM_SHORT_GAP_X;
        case shortGapY: return m->TRANSITION_MATCH_FROM_SHORT_GAP_Y;
        case longGapX: return m->TRANSITION_MATCH_FROM_LONG_GAP_X;
    }
    return 0.0;
}
static double sm4_raggedEnd(StateMachine *sM, int64_t st) {                                                 // :812-826
    StateMachine4 *m = (StateMachine4 *) sM;
    return st == longGapX ? m->TRANSITION_GAP_LONG_EXTEND_X : m->TRANSITION_GAP_LONG_OPEN_X;
}
StateMachine *getStateMachine4(const char *modelFile) {
    StateMachine4 *m = (StateMachine4 *) calloc(1, sizeof(StateMachine4));
    initBase(&m->model, fourState, NUM_OF_KMERS);
    m->model.stateNumber = 4;
    m->model.startStateProb = sm4_start; m->model.raggedStartStateProb = sm4_raggedStart;
    m->model.endStateProb = sm4_end; m->model.raggedEndStateProb = sm4_raggedEnd;
    m->TRANSITION_MATCH_CONTINUE = -0.23552123624314988;
    m->TRANSITION_GAP_SHORT_OPEN_X = -1.6269694202638481;
    m->TRANSITION_GAP_SHORT_OPEN_Y = -4.7241893208381773;
    m->TRANSITION_GAP_LONG_OPEN_X = -5.4173365013981227;
    m->TRANSITION_GAP_SHORT_EXTEND_X = -1.6269694202638481;
    m->TRANSITION_MATCH_FROM_SHORT_GAP_X = -0.21880828092192281;
    m->TRANSITION_GAP_LONG_EXTEND_X = -0.003442492794189331;
    m->TRANSITION_MATCH_FROM_LONG_GAP_X = -5.6732801731704612;
    m->TRANSITION_MATCH_FROM_SHORT_GAP_Y = -0.013406326748077823;
    m->TRANSITION_GAP_SHORT_EXTEND_Y = -4.724189320832104;
    m->TRANSITION_GAP_LONG_SWITCH_TO_X = -5.4173365013920494;
    loadPoreModel(&m->model, modelFile);
    return &m->model;
}

// getStateMachineEchelon (impl/stateMachine.c:1773-1784, constructor :1602-1640): seven states, the model's 30 skip bins
// twice (as the vanilla machine), end "probabilities" that are not logs (:1617-1619, kept as they are)
static double sme_start(StateMachine *, int64_t st) { return st == 1 ? 0 : LOG_ZERO; }                        // :1237-1245
static double sme_raggedStart(StateMachine *, int64_t st) { return st == 6 ? 0 : LOG_ZERO; }                  // :1247-1253
static double sme_end(StateMachine *sM, int64_t st) {                                                       // :1255-1262
    StateMachineEchelon *m = (StateMachineEchelon *) sM;
    return st == 6 ? m->DEFAULT_END_FROM_X_PROB : m->DEFAULT_END_MATCH_PROB;
}
StateMachine *getStateMachineEchelon(const char *modelFile) {
    StateMachineEchelon *m = (StateMachineEchelon *) calloc(1, sizeof(StateMachineEchelon));
    initBase(&m->model, echelon, 60);
    m->model.stateNumber = 7;
    m->model.matchState = 1;
    m->model.startStateProb = sme_start; m->model.raggedStartStateProb = sme_raggedStart;
    m->model.endStateProb = sme_end; m->model.raggedEndStateProb = sme_end;
    m->DEFAULT_END_MATCH_PROB = 0.79015888282447311;
    m->DEFAULT_END_FROM_X_PROB = 0.19652425498269727;
    m->BACKGROUND_EVENT_PROB = -3.0;
    loadPoreModel(&m->model, modelFile);
    return &m->model;
}

// impl/stateMachine.c:631-651: the MATCH table only; the table's own noise_sd column is replaced
void emissions_signal_scaleModel(StateMachine *sM, double scale, double shift, double var, double scale_sd, double var_sd) {
    double *t = sM->EMISSION_MATCH_PROBS;
    for (int64_t i = 1; i < TABLE_LEN; i += MODEL_PARAMS) {
        t[i] = t[i] * scale + shift;
        t[i + 1] = t[i + 1] * var;
        t[i + 2] = t[i + 2] * scale_sd;
        t[i + 4] = t[i + 4] * var_sd;
        t[i + 3] = sqrt(pow(t[i + 2], 3.0) / t[i + 4]);
    }
}


// impl/stateMachine.c:653-673: the same without the level mean (deviation, noise mean, lambda, and the derived noise sd)
void emissions_signal_scaleModelNoiseOnly(StateMachine *sM, double scale, double shift, double var, double scale_sd, double var_sd) {
    (void) scale; (void) shift;
    double *t = sM->EMISSION_MATCH_PROBS;
    for (int64_t i = 1; i < TABLE_LEN; i += MODEL_PARAMS) {
        t[i + 1] = t[i + 1] * var;
        t[i + 2] = t[i + 2] * scale_sd;
        t[i + 4] = t[i + 4] * var_sd;
        t[i + 3] = sqrt(pow(t[i + 2], 3.0) / t[i + 4]);
    }
}

}  // extern "C"

// ============================================================================== NanoporeHDP (threeStateHdp)
// What get_nanopore_kmer_density needs of the reference's HierarchicalDirichletProcess: the sampling grid and, per
// OBSERVED Dirichlet process, the posterior predictive density on it with its spline slopes.  Everything else in a
// serialised HDP (data, Gibbs state, factor tree) is read past.
namespace {

struct HdpTables {
    double gridStart = 0, gridStop = 0;
    int64_t gridLength = 0, numDps = 0;
    std::vector<double> grid;               // linspace (impl/hdp_math_utils.c:497-510)
    std::vector<int64_t> parent;            // -1: the base process
    std::vector<int32_t> row;               // per process: its row in density / slopes, or -1 (not observed)
    std::vector<double> density, slopes;
    bool splinesFinalized = false;
    int32_t deviceModel = -1;               // cpecan_cuda_upload_hdp id, once an alignment has used it
};

std::vector<double> splitDoubles(const std::string &line) {
    std::vector<double> v;
    const char *p = line.c_str();
    char *end = nullptr;
    for (;;) {
        const double d = strtod(p, &end);
        if (end == p) break;
        v.push_back(d); p = end;
    }
    return v;
}

bool nextLine(std::ifstream &in, std::string &line, const char *what, const char *path) {
    if (!std::getline(in, line)) st_errAbort("deserialize_nhdp: %s ends before %s", path, what);
    return true;
}

// kmer_id (impl/nanopore_hdp.c:348-380): big-endian word over the HDP's sorted alphabet; the reference exits on a
// character outside it
int64_t hdpKmerId(const NanoporeHDP *nhdp, const char *kmer, bool fatal) {
    int64_t id = 0;
    for (int64_t i = 0; i < nhdp->kmer_length; i++) {
        const char *q = (const char *) memchr(nhdp->alphabet, kmer[i], (size_t) nhdp->alphabet_size);
        if (q == nullptr || kmer[i] == 0) {
            if (!fatal) return -1;
            fprintf(stderr, "vanillaAlign - ERROR: K-mer contains character outside alphabet. Got offending kmer is: %.*s. alphabet is %s\n",
                    (int) nhdp->kmer_length, kmer, nhdp->alphabet);
            exit(EXIT_FAILURE);
        }
        id = id * nhdp->alphabet_size + (q - nhdp->alphabet);
    }
    return id;
}

int32_t hdpRowOf(const HdpTables *t, int64_t dp) {           // impl/hdp.c:2588-2591: the nearest observed ancestor
    while (dp >= 0 && t->row[(size_t) dp] < 0) dp = t->parent[(size_t) dp];
    return dp < 0 ? -1 : t->row[(size_t) dp];
}

}  // namespace

extern "C" {

NanoporeHDP *deserialize_nhdp(const char *filepath) {
    std::ifstream in(filepath);
    if (!in) st_errAbort("deserialize_nhdp: cannot open %s", filepath);
    std::string line;
    NanoporeHDP *nhdp = (NanoporeHDP *) calloc(1, sizeof(NanoporeHDP));
    HdpTables *t = new HdpTables();
    nhdp->hdp = t;
    nextLine(in, line, "the alphabet size", filepath); nhdp->alphabet_size = atoll(line.c_str());
    nextLine(in, line, "the alphabet", filepath);
    { std::istringstream is(line); std::string a; is >> a;
      if ((int64_t) a.size() != nhdp->alphabet_size || a.empty()) st_errAbort("deserialize_nhdp: alphabet '%s' does not have %lld characters", a.c_str(), (long long) nhdp->alphabet_size);
      std::sort(a.begin(), a.end());                       // package_nanopore_hdp keeps it sorted (impl/nanopore_hdp.c:38-55)
      nhdp->alphabet = strdup(a.c_str()); }
    nextLine(in, line, "the k-mer length", filepath); nhdp->kmer_length = atoll(line.c_str());
    // ---- deserialize_hdp (impl/hdp.c:3009-3270)
    nextLine(in, line, "splines_finalized", filepath); t->splinesFinalized = atoll(line.c_str()) != 0;
    nextLine(in, line, "has_data", filepath); const bool hasData = atoll(line.c_str()) != 0;
    nextLine(in, line, "sample_gamma", filepath); const bool sampleGamma = atoll(line.c_str()) != 0;
    nextLine(in, line, "num_dps", filepath); t->numDps = atoll(line.c_str());
    if (t->numDps <= 0) st_errAbort("deserialize_nhdp: %s has no Dirichlet processes", filepath);
    std::vector<int64_t> dataDp;
    if (hasData) {
        nextLine(in, line, "the data", filepath);
        nextLine(in, line, "the data assignments", filepath);
        const char *p = line.c_str(); char *end = nullptr;
        for (;;) { const long long v = strtoll(p, &end, 10); if (end == p) break; dataDp.push_back(v); p = end; }
    }
    nextLine(in, line, "the base parameters", filepath);
    nextLine(in, line, "the sampling grid", filepath);
    { std::vector<double> g = splitDoubles(line);
      if (g.size() != 3 || g[2] < 2 || !(g[0] < g[1])) st_errAbort("deserialize_nhdp: bad sampling grid line in %s", filepath);
      t->gridStart = g[0]; t->gridStop = g[1]; t->gridLength = (int64_t) g[2]; }
    nextLine(in, line, "gamma", filepath);
    if (sampleGamma) for (int i = 0; i < 4; i++) nextLine(in, line, "the gamma distribution parameters", filepath);
    t->parent.assign((size_t) t->numDps, -1);
    for (int64_t id = 0; id < t->numDps; id++) {
        nextLine(in, line, "the parents", filepath);
        if (line.empty() || line[0] != '-') t->parent[(size_t) id] = atoll(line.c_str());
        if (t->parent[(size_t) id] >= t->numDps) st_errAbort("deserialize_nhdp: parent out of range in %s", filepath);
    }
    // mark_observed_dps (impl/hdp.c:1124-1152): the processes holding data, and all their ancestors
    std::vector<char> observed((size_t) t->numDps, 0);
    for (int64_t id : dataDp) {
        if (id < 0 || id >= t->numDps) st_errAbort("deserialize_nhdp: data point assigned to process %lld of %lld", (long long) id, (long long) t->numDps);
        for (int64_t dp = id; dp >= 0 && !observed[(size_t) dp]; dp = t->parent[(size_t) dp]) observed[(size_t) dp] = 1;
    }
    t->row.assign((size_t) t->numDps, -1);
    const int64_t n = t->gridLength;
    t->grid.resize((size_t) n);
    { const double dx = (t->gridStop - t->gridStart) / (double) (n - 1);
      for (int64_t i = 0; i < n - 1; i++) t->grid[(size_t) i] = t->gridStart + i * dx;
      t->grid[(size_t) n - 1] = t->gridStop; }
    if (hasData) {
        int32_t rows = 0;
        for (int64_t id = 0; id < t->numDps; id++) {
            nextLine(in, line, "the posterior predictive densities", filepath);
            std::vector<double> v = splitDoubles(line);
            if (v.empty()) {
                if (observed[(size_t) id]) st_errAbort("deserialize_nhdp: observed process %lld has no density in %s", (long long) id, filepath);
                continue;
            }
            if ((int64_t) v.size() != n) st_errAbort("deserialize_nhdp: density of process %lld has %lld of %lld grid points", (long long) id, (long long) v.size(), (long long) n);
            if (!observed[(size_t) id]) continue;             // never read by dir_proc_density
            t->row[(size_t) id] = rows++;
            t->density.insert(t->density.end(), v.begin(), v.end());
        }
        t->slopes.assign(t->density.size(), 0.0);
        if (t->splinesFinalized) {
            for (int64_t id = 0; id < t->numDps; id++) {
                nextLine(in, line, "the spline slopes", filepath);
                std::vector<double> v = splitDoubles(line);
                if (v.empty()) {
                    if (t->row[(size_t) id] >= 0) st_errAbort("deserialize_nhdp: observed process %lld has no spline slopes in %s", (long long) id, filepath);
                    continue;
                }
                if ((int64_t) v.size() != n) st_errAbort("deserialize_nhdp: slopes of process %lld have %lld of %lld grid points", (long long) id, (long long) v.size(), (long long) n);
                if (t->row[(size_t) id] >= 0) std::copy(v.begin(), v.end(), t->slopes.begin() + (size_t) t->row[(size_t) id] * (size_t) n);
            }
        }
    }
    return nhdp;                                              // the factor tree that follows is the Gibbs sampler's state
}

void cpecan_host_release_hdp_locked(NanoporeHDP *nhdp);

void destroy_nanopore_hdp(NanoporeHDP *nhdp) {
    if (!nhdp) return;
    cpecan_host_release_hdp_locked(nhdp);
    delete (HdpTables *) nhdp->hdp;
    free(nhdp->alphabet);
    free(nhdp);
}

int64_t get_nanopore_hdp_kmer_length(NanoporeHDP *nhdp) { return nhdp->kmer_length; }
int64_t get_nanopore_hdp_alphabet_size(NanoporeHDP *nhdp) { return nhdp->alphabet_size; }
char *get_nanopore_hdp_alphabet(NanoporeHDP *nhdp) { return strdup(nhdp->alphabet); }

// impl/nanopore_hdp.c:390-392 -> dir_proc_density (impl/hdp.c:2577-2599) -> grid_spline_interp
// (impl/hdp_math_utils.c:471-495): the same expressions in the same order
double get_nanopore_kmer_density(NanoporeHDP *nhdp, void *kmer, void *x) {
    const HdpTables *t = (const HdpTables *) nhdp->hdp;
    if (!t->splinesFinalized) { fprintf(stderr, "Must finalize distributions before querying densities.\n"); exit(EXIT_FAILURE); }
    const int64_t id = hdpKmerId(nhdp, (const char *) kmer, true);
    if (id >= t->numDps) { fprintf(stderr, "Hierarchical Dirichlet process has no Dirichlet process with this ID.\n"); exit(EXIT_FAILURE); }
    const int32_t row = hdpRowOf(t, id);
    if (row < 0) st_errAbort("get_nanopore_kmer_density: no observed Dirichlet process above this k-mer's (an HDP without data)");
    const int64_t n = t->gridLength;
    const double *xg = t->grid.data(), *y = t->density.data() + (size_t) row * (size_t) n, *slope = t->slopes.data() + (size_t) row * (size_t) n;
    const double q = *(const double *) x;
    double interp;
    if (q <= xg[0]) interp = y[0] - slope[0] * (xg[0] - q);
    else if (q >= xg[n - 1]) interp = y[n - 1] + slope[n - 1] * (q - xg[n - 1]);
    else {
        const double dx = xg[1] - xg[0];
        int64_t il = (int64_t) ((q - xg[0]) / dx);
        if (il > n - 2) il = n - 2;                          // the reference would read past the grid here
        const int64_t ir = il + 1;
        const double dy = y[ir] - y[il];
        const double a = slope[il] * dx - dy;
        const double b = dy - slope[ir] * dx;
        const double tl = (q - xg[il]) / dx;
        const double tr = 1.0 - tl;
        interp = tr * y[il] + tl * y[ir] + tl * tr * (a * tr + b * tl);
    }
    return interp > 0.0 ? interp : 0.0;
}

StateMachine *getHdpStateMachine3(NanoporeHDP *hdp) {                                                       // :1738-1749, 1513-1557
    StateMachine3_HDP *m = (StateMachine3_HDP *) calloc(1, sizeof(StateMachine3_HDP));
    initBase(&m->model, threeStateHdp, NUM_OF_KMERS);
    m->model.endStateProb = sm3_end; m->model.raggedEndStateProb = sm3_raggedEnd;
    stateMachine3_setTransitionsToNanoporeDefaults(&m->model);          // the nine transitions sit where StateMachine3's do
    for (int64_t i = 0; i < NUM_OF_KMERS; i++) m->model.EMISSION_GAP_X_PROBS[i] = -2.3025850929940455;
    m->hdpModel = hdp;
    m->getYGapProbFcn = get_nanopore_kmer_density;
    m->getMatchProbFcn = get_nanopore_kmer_density;
    return &m->model;
}

}  // extern "C"

// ===================================================================================================== GPU context
namespace {

std::mutex gMu;
cpecan_ctx *gCtx = nullptr;
int gDevice = -1;
int gExact = -1;                                // cpecan_host_set_exact_arithmetic / $CPECAN_EXACT
std::map<StateMachine *, int32_t> gModels;      // device copies of the tables, refreshed at every submit (the public
                                                // EMISSION_* / TRANSITION_* fields are mutable)
cpecan_ctx *gpuLocked() {
    if (!gCtx) {
        int dev = gDevice >= 0 ? gDevice : (getenv("CPECAN_DEVICE") ? atoi(getenv("CPECAN_DEVICE")) : 0);
        if (cpecan_cuda_init(dev, &gCtx) != CPECAN_OK)
            st_errAbort("cpecan: no usable CUDA device %d; the banded DP has no CPU fallback in this build", dev);
        if (gExact < 0) gExact = getenv("CPECAN_EXACT") ? atoi(getenv("CPECAN_EXACT")) : 0;
        cpecan_cuda_set_exact_arithmetic(gCtx, gExact);
    }
    return gCtx;
}

// the device tables of a state machine go with it (cpecan_cuda_release_model): a caller that builds one scaled state
// machine per read and strand, as vanillaAlign.c does, would otherwise grow the device model list without bound
void modelDropLocked(StateMachine *sM) {
    auto it = gModels.find(sM);
    if (it == gModels.end()) return;
    if (gCtx) cpecan_cuda_release_model(gCtx, it->second);
    gModels.erase(it);
}

// the device copy of an HDP belongs to the NanoporeHDP, not to the state machines that point at it
int32_t hdpModelIdLocked(cpecan_ctx *ctx, NanoporeHDP *nhdp) {
    HdpTables *t = (HdpTables *) nhdp->hdp;
    if (t->deviceModel >= 0) return t->deviceModel;
    if (nhdp->kmer_length != KMER_LENGTH) st_errAbort("cpecan: the HDP is over %lld-mers, the alignment over %d-mers", (long long) nhdp->kmer_length, KMER_LENGTH);
    if (!t->splinesFinalized || t->density.empty()) st_errAbort("cpecan: the HDP holds no finalized distributions");
    std::vector<int32_t> kmerRow((size_t) NUM_OF_KMERS, -1);
    for (int k = 0; k < NUM_OF_KMERS; k++) {
        char km[KMER_LENGTH + 1];
        for (int j = 0; j < KMER_LENGTH; j++) km[j] = "ACGT"[(k >> (2 * (KMER_LENGTH - 1 - j))) & 3];
        km[KMER_LENGTH] = 0;
        const int64_t id = hdpKmerId(nhdp, km, false);
        if (id >= 0 && id < t->numDps) kmerRow[(size_t) k] = hdpRowOf(t, id);
    }
    int32_t id = -1;
    if (cpecan_cuda_upload_hdp(ctx, t->gridStart, t->gridStop, t->gridLength, (int32_t) (t->density.size() / (size_t) t->gridLength),
                               t->density.data(), t->slopes.data(), kmerRow.data(), &id) != CPECAN_OK)
        st_errAbort("cpecan_cuda_upload_hdp: %s", cpecan_cuda_last_error(ctx));
    t->deviceModel = id;
    return id;
}

int32_t modelIdLocked(cpecan_ctx *ctx, StateMachine *sM) {
    if (sM->type == threeStateHdp) return hdpModelIdLocked(ctx, ((StateMachine3_HDP *) sM)->hdpModel);
    const int nGapX = (sM->type == vanilla || sM->type == echelon) ? 60 : NUM_OF_KMERS;
    auto it = gModels.find(sM);
    if (it != gModels.end()) {
        if (cpecan_cuda_update_model(ctx, it->second, sM->EMISSION_MATCH_PROBS, sM->EMISSION_GAP_Y_PROBS, sM->EMISSION_GAP_X_PROBS) != CPECAN_OK)
            st_errAbort("cpecan_cuda_update_model: %s", cpecan_cuda_last_error(ctx));
        return it->second;
    }
    int32_t id = -1;
    if (cpecan_cuda_upload_model(ctx, sM->EMISSION_MATCH_PROBS, sM->EMISSION_GAP_Y_PROBS, sM->EMISSION_GAP_X_PROBS, nGapX, &id) != CPECAN_OK)
        st_errAbort("cpecan_cuda_upload_model: %s", cpecan_cuda_last_error(ctx));
    gModels[sM] = id;
    return id;
}

cpecan_hmm hmmOf(StateMachine *sM) {
    cpecan_hmm h;
    memset(&h, 0, sizeof(h));
    if (sM->type == threeState) {
        StateMachine3 *m = (StateMachine3 *) sM;
        h.sm_type = CPECAN_SM_THREE_STATE;
        const double t[9] = { m->TRANSITION_MATCH_CONTINUE, m->TRANSITION_MATCH_FROM_GAP_X, m->TRANSITION_MATCH_FROM_GAP_Y,
                              m->TRANSITION_GAP_OPEN_X, m->TRANSITION_GAP_OPEN_Y, m->TRANSITION_GAP_EXTEND_X,
                              m->TRANSITION_GAP_EXTEND_Y, m->TRANSITION_GAP_SWITCH_TO_X, m->TRANSITION_GAP_SWITCH_TO_Y };
        memcpy(h.transitions, t, sizeof(t));
    } else if (sM->type == vanilla) {
        StateMachine3Vanilla *m = (StateMachine3Vanilla *) sM;
        h.sm_type = CPECAN_SM_VANILLA;
        const double v[5] = { m->TRANSITION_M_TO_Y_NOT_X, m->TRANSITION_E_TO_E, m->DEFAULT_END_MATCH_PROB,
                              m->DEFAULT_END_FROM_X_PROB, m->DEFAULT_END_FROM_Y_PROB };
        memcpy(h.vanilla, v, sizeof(v));
    } else if (sM->type == fourState) {
        StateMachine4 *m = (StateMachine4 *) sM;
        h.sm_type = CPECAN_SM_FOUR_STATE;
        const double t[11] = { m->TRANSITION_MATCH_CONTINUE, m->TRANSITION_MATCH_FROM_SHORT_GAP_X, m->TRANSITION_MATCH_FROM_SHORT_GAP_Y,
                               m->TRANSITION_MATCH_FROM_LONG_GAP_X, m->TRANSITION_GAP_SHORT_OPEN_X, m->TRANSITION_GAP_SHORT_EXTEND_X,
                               m->TRANSITION_GAP_SHORT_OPEN_Y, m->TRANSITION_GAP_SHORT_EXTEND_Y, m->TRANSITION_GAP_LONG_OPEN_X,
                               m->TRANSITION_GAP_LONG_EXTEND_X, m->TRANSITION_GAP_LONG_SWITCH_TO_X };
        memcpy(h.four_state, t, sizeof(t));
    } else if (sM->type == echelon) {
        h.sm_type = CPECAN_SM_ECHELON;
    } else if (sM->type == threeStateHdp) {
        StateMachine3_HDP *m = (StateMachine3_HDP *) sM;
        h.sm_type = CPECAN_SM_THREE_STATE_HDP;
        const double t[9] = { m->TRANSITION_MATCH_CONTINUE, m->TRANSITION_MATCH_FROM_GAP_X, m->TRANSITION_MATCH_FROM_GAP_Y,
                              m->TRANSITION_GAP_OPEN_X, m->TRANSITION_GAP_OPEN_Y, m->TRANSITION_GAP_EXTEND_X,
                              m->TRANSITION_GAP_EXTEND_Y, m->TRANSITION_GAP_SWITCH_TO_X, m->TRANSITION_GAP_SWITCH_TO_Y };
        memcpy(h.transitions, t, sizeof(t));
    } else {
        st_errAbort("cpecan: state machine type %d is not implemented on the GPU (threeState, vanilla, echelon, fourState, threeStateHdp are)", (int) sM->type);
    }
    return h;
}

bool sameHmm(const cpecan_hmm &a, const cpecan_hmm &b) {
    if (a.sm_type != b.sm_type) return false;
    for (int i = 0; i < 9; i++) if (!(a.transitions[i] == b.transitions[i] || (std::isinf(a.transitions[i]) && std::isinf(b.transitions[i])))) return false;
    for (int i = 0; i < 5; i++) if (a.vanilla[i] != b.vanilla[i]) return false;
    for (int i = 0; i < 11; i++) if (a.four_state[i] != b.four_state[i]) return false;
    return true;
}

cpecan_params paramsOf(const PairwiseAlignmentParameters *p) {
    cpecan_params q = { p->threshold, p->minDiagsBetweenTraceBack, p->traceBackDiagonals, p->diagonalExpansion,
                        p->constraintDiagonalTrim, p->splitMatrixBiggerThanThis };
    return q;
}

// flat batch under construction: one item per split region
struct Flat {
    std::string ref;
    std::vector<int64_t> refOff{ 0 }, evOff{ 0 }, anOff{ 0 }, anchors;
    std::vector<double> events;
    std::vector<int32_t> modelId;
    std::vector<uint8_t> ragged;
    std::vector<int64_t> offX, offY, owner;     // coordinate correction and owning read of every item
    int64_t n() const { return (int64_t) modelId.size(); }
};

void checkSequences(StateMachine *sM, Sequence *sX, Sequence *sY) {
    // the type-erased sequences are recognised by their accessors (SURVEY 7 "Function-pointer API"): anything else
    // cannot be read by the device path
    const bool k2 = sM->type == vanilla || sM->type == echelon;          // vanillaAlign.c:224-249
    const bool k3 = sM->type == threeStateHdp;                            // vanillaAlign.c:246-250
    void *(*wantX)(void *, int64_t) = k2 ? sequence_getKmer2 : (k3 ? sequence_getKmer3 : sequence_getKmer);
    if (sX->get != wantX) st_errAbort("cpecan: the reference sequence must use %s for this state machine", k2 ? "sequence_getKmer2" : (k3 ? "sequence_getKmer3" : "sequence_getKmer"));
    if (sY->get != sequence_getEvent) st_errAbort("cpecan: the read sequence must be an event sequence (sequence_getEvent)");
}

// the loop of getPosteriorProbsWithBandingSplittingAlignmentsByLargeGaps (impl/pairwiseAligner.c:1356-1422) as data
void addRead(Flat &f, int64_t owner, int32_t modelId, Sequence *sX, Sequence *sY, stList *anchorPairs,
             const PairwiseAlignmentParameters *p, bool raggedLeft, bool raggedRight) {
    const int64_t lX = sX->length, lY = sY->length;
    stList *split = getSplitPoints(anchorPairs, lX, lY, p->splitMatrixBiggerThanThis, raggedLeft, raggedRight);
    const int64_t nS = stList_length(split), nA = stList_length(anchorPairs);
    int64_t j = 0;
    for (int64_t i = 0; i < nS; i++) {
        stIntTuple *r = (stIntTuple *) stList_get(split, i);
        const int64_t x1 = stIntTuple_get(r, 0), y1 = stIntTuple_get(r, 1), x2 = stIntTuple_get(r, 2), y2 = stIntTuple_get(r, 3);
        // k-mers x1 .. x2-1 read nucleotides x1 .. x2+4 of the parent buffer (sequence_sliceNucleotideSequence2)
        if (x2 > x1) f.ref.append((const char *) sX->elements + x1, (size_t) (x2 - x1 + KMER_LENGTH - 1));
        f.refOff.push_back((int64_t) f.ref.size());
        const double *ev = (const double *) sY->elements + y1 * NB_EVENT_PARAMS;
        f.events.insert(f.events.end(), ev, ev + (y2 - y1) * NB_EVENT_PARAMS);
        f.evOff.push_back((int64_t) f.events.size() / NB_EVENT_PARAMS);
        while (j < nA) {
            stIntTuple *a = (stIntTuple *) stList_get(anchorPairs, j);
            const int64_t x = stIntTuple_get(a, 0), y = stIntTuple_get(a, 1);
            if (x + y >= x2 + y2) break;
            f.anchors.push_back(x - x1); f.anchors.push_back(y - y1);
            j++;
        }
        f.anOff.push_back((int64_t) f.anchors.size() / 2);
        f.modelId.push_back(modelId);
        f.ragged.push_back((uint8_t) (((raggedLeft || i > 0) ? 1 : 0) | ((raggedRight || i < nS - 1) ? 2 : 0)));
        f.offX.push_back(x1); f.offY.push_back(y1); f.owner.push_back(owner);
    }
    stList_destruct(split);
}

cpecan_batch batchOf(Flat &f) {
    if (f.events.empty()) f.events.push_back(0.0);
    if (f.anchors.empty()) f.anchors.push_back(0);
    cpecan_batch b;
    b.n_items = f.n();
    b.ref = f.ref.c_str(); b.ref_off = f.refOff.data();
    b.events = f.events.data(); b.ev_off = f.evOff.data();
    b.anchors = f.anchors.data(); b.anchor_off = f.anOff.data();
    b.model_id = f.modelId.data();
    b.scale = nullptr;                        // the StateMachine's tables are already scaled (emissions_signal_scaleModel)
    b.ragged = f.ragged.data();
    return b;
}

// aligned pairs of a flat batch -> one stList per read: every region shifted back and reversed
// (alignedPairCoordinateCorrectionFn pops, impl/pairwiseAligner.c:1447-1454)
void runPosterior(cpecan_ctx *ctx, const cpecan_hmm &hmm, const cpecan_params &prm, int mode, Flat &f, stList **results) {
    if (f.n() == 0) return;
    cpecan_batch b = batchOf(f);
    std::vector<cpecan_result> res((size_t) f.n());
    int64_t cap = 4 * (f.evOff.back() + 1) + 64 * f.n() + 1024;
    std::vector<int32_t> triples;
    for (int attempt = 0; attempt < 2; attempt++) {
        triples.assign((size_t) cap * 3, 0);
        if (cpecan_cuda_align_batch(ctx, &hmm, &prm, mode, &b, triples.data(), cap, res.data(), nullptr, nullptr) != CPECAN_OK)
            st_errAbort("cpecan_cuda_align_batch: %s", cpecan_cuda_last_error(ctx));
        bool overflow = false;
        for (auto &r : res) overflow |= (r.status & CPECAN_ITEM_PAIR_OVERFLOW) != 0;
        if (!overflow) break;
        int64_t need = 0;
        for (auto &r : res) need += r.n_pairs;
        cap = 8 * need + 1024;                // generous: slices are proportional to the event counts
        if (attempt == 1) st_errAbort("cpecan: aligned-pair buffer overflow");
    }
    for (int64_t i = 0; i < f.n(); i++) {
        if (res[(size_t) i].status & CPECAN_ITEM_BAND_STEP) st_errAbort("cpecan: anchors are not strictly increasing (run filterToRemoveOverlap)");
        if (hmm.sm_type == CPECAN_SM_THREE_STATE_HDP && (res[(size_t) i].status & CPECAN_ITEM_BAD_KMER)) {
            // kmer_to_word (impl/nanopore_hdp.c:358-373) ends the program on the first such k-mer
            fprintf(stderr, "vanillaAlign - ERROR: K-mer contains character outside alphabet.\n");
            exit(EXIT_FAILURE);
        }
        stList *out = results[f.owner[(size_t) i]];
        const int32_t *t = triples.data() + 3 * res[(size_t) i].pair_off;
        if (mode == CPECAN_MODE_UNBANDED) {
            // the reference lists ascending diagonals, x ascending inside a diagonal (impl/pairwiseAligner.c:1560-1562); the
            // device emits descending diagonals, x ascending: reverse the order of the diagonals, keep each one's run
            int64_t k = res[(size_t) i].n_pairs;
            while (k > 0) {
                int64_t k0 = k - 1;
                const int32_t dgl = t[3 * k0 + 1] + t[3 * k0 + 2];
                while (k0 > 0 && t[3 * (k0 - 1) + 1] + t[3 * (k0 - 1) + 2] == dgl) k0--;
                for (int64_t j = k0; j < k; j++)
                    stList_append(out, stIntTuple_construct3(t[3 * j], t[3 * j + 1] + f.offX[(size_t) i], t[3 * j + 2] + f.offY[(size_t) i]));
                k = k0;
            }
            continue;
        }
        for (int64_t k = res[(size_t) i].n_pairs - 1; k >= 0; k--)
            stList_append(out, stIntTuple_construct3(t[3 * k], t[3 * k + 1] + f.offX[(size_t) i], t[3 * k + 2] + f.offY[(size_t) i]));
    }
}

}  // namespace

extern "C" {

void cpecan_host_release_hdp_locked(NanoporeHDP *nhdp) {
    std::lock_guard<std::mutex> lk(gMu);
    HdpTables *t = (HdpTables *) nhdp->hdp;
    if (t && t->deviceModel >= 0 && gCtx) cpecan_cuda_release_model(gCtx, t->deviceModel);
    if (t) t->deviceModel = -1;
}

void cpecan_host_set_device(int device) {
    std::lock_guard<std::mutex> lk(gMu);
    if (gCtx && device != gDevice) st_errAbort("cpecan_host_set_device: the GPU context already exists");
    gDevice = device;
}

void cpecan_host_set_exact_arithmetic(int on) {
    std::lock_guard<std::mutex> lk(gMu);
    gExact = on ? 1 : 0;
    if (gCtx) cpecan_cuda_set_exact_arithmetic(gCtx, gExact);
}

void stateMachine_destruct(StateMachine *sM) {                 // impl/stateMachine.c:1786-1788 leaks the tables; freed here
    if (!sM) return;
    { std::lock_guard<std::mutex> lk(gMu); modelDropLocked(sM); }
    free(sM->EMISSION_MATCH_PROBS); free(sM->EMISSION_GAP_X_PROBS); free(sM->EMISSION_GAP_Y_PROBS);
    free(sM);
}

void diagonalCalculationPosteriorMatchProbs(StateMachine *, int64_t, DpMatrix *, DpMatrix *, Sequence *, Sequence *, double,
                                            PairwiseAlignmentParameters *, void *) {
    noCpuPath("diagonalCalculationPosteriorMatchProbs");
}
void diagonalCalculation_Expectations(StateMachine *, int64_t, DpMatrix *, DpMatrix *, Sequence *, Sequence *, double,
                                      PairwiseAlignmentParameters *, void *) {
    noCpuPath("diagonalCalculation_Expectations");
}
void diagonalCalculationMultiPosteriorMatchProbs(StateMachine *, int64_t, DpMatrix *, DpMatrix *, Sequence *, Sequence *, double,
                                                 PairwiseAlignmentParameters *, void *) {
    noCpuPath("diagonalCalculationMultiPosteriorMatchProbs");
}
// the posterior callback must be the one the machine's cells are made for (vanillaAlign.c:232-249)
static void checkPosteriorFn(StateMachine *sM, void (*fn)(StateMachine *, int64_t, DpMatrix *, DpMatrix *, Sequence *, Sequence *,
                                                          double, PairwiseAlignmentParameters *, void *), const char *who) {
    if (sM->type == echelon ? fn != diagonalCalculationMultiPosteriorMatchProbs : fn != diagonalCalculationPosteriorMatchProbs)
        st_errAbort("%s: the GPU path implements diagonalCalculationPosteriorMatchProbs (threeState, vanilla, fourState) and "
                    "diagonalCalculationMultiPosteriorMatchProbs (echelon)", who);
}

void getAlignedPairsUsingAnchorsBatch(int64_t n, StateMachine **sMs, Sequence **sXs, Sequence **sYs, stList **anchorPairs,
                                      PairwiseAlignmentParameters *p, bool raggedLeft, bool raggedRight, stList **results) {
    std::lock_guard<std::mutex> lk(gMu);            // callable from the two OpenMP sections of vanillaAlign.c:737-790
    cpecan_ctx *ctx = gpuLocked();
    const cpecan_params prm = paramsOf(p);
    for (int64_t i = 0; i < n; i++) results[i] = stList_construct3(0, (void (*)(void *)) stIntTuple_destruct);
    // one GPU batch per DISTINCT machine (type + transitions), whatever the order of the reads: a caller that submits
    // template, complement, template, complement, ... of the vanilla machine (whose strands differ in two transitions)
    // gets two batches, not one per strand of every read
    std::vector<cpecan_hmm> hmms;
    std::vector<std::vector<int64_t>> groups;
    for (int64_t i = 0; i < n; i++) {
        const cpecan_hmm h = hmmOf(sMs[i]);
        size_t g = 0;
        while (g < hmms.size() && !sameHmm(hmms[g], h)) g++;
        if (g == hmms.size()) { hmms.push_back(h); groups.emplace_back(); }
        groups[g].push_back(i);
    }
    for (size_t g = 0; g < hmms.size(); g++) {
        Flat f;
        for (int64_t j : groups[g]) {
            checkSequences(sMs[j], sXs[j], sYs[j]);
            // A negative length happens in the reference's own pipeline (vanillaAlign.c:645-651 slices the complement
            // events with a DECREASING event map); the reference then walks a degenerate band along y = 0 and reports no
            // pair (a posterior needs y > 0).
            if (sXs[j]->length < 0 || sYs[j]->length < 0) continue;
            addRead(f, j, modelIdLocked(ctx, sMs[j]), sXs[j], sYs[j], anchorPairs[j], p, raggedLeft, raggedRight);
        }
        runPosterior(ctx, hmms[g], prm, CPECAN_MODE_POSTERIOR, f, results);
    }
}

stList *getAlignedPairsUsingAnchors(StateMachine *sM, Sequence *SsX, Sequence *SsY, stList *anchorPairs,
                                    PairwiseAlignmentParameters *p,
                                    void (*diagonalPosteriorProbFn)(StateMachine *, int64_t, DpMatrix *, DpMatrix *, Sequence *,
                                                                    Sequence *, double, PairwiseAlignmentParameters *, void *),
                                    bool alignmentHasRaggedLeftEnd, bool alignmentHasRaggedRightEnd) {
    checkPosteriorFn(sM, diagonalPosteriorProbFn, "getAlignedPairsUsingAnchors");
    stList *result = nullptr;
    getAlignedPairsUsingAnchorsBatch(1, &sM, &SsX, &SsY, &anchorPairs, p, alignmentHasRaggedLeftEnd, alignmentHasRaggedRightEnd, &result);
    return result;
}

// impl/pairwiseAligner.c:1512-1569: full matrix, ONE totalProbability at the last diagonal.  The reference hands the
// pairs back in ascending-diagonal order, x ascending inside a diagonal; runPosterior re-orders the device's list so.
stList *getAlignedPairsWithoutBanding(StateMachine *sM, void *cX, void *cY, int64_t lX, int64_t lY, PairwiseAlignmentParameters *p,
                                      void *(*getXFcn)(void *, int64_t), void *(*getYFcn)(void *, int64_t),
                                      void (*diagonalPosteriorProbFn)(StateMachine *, int64_t, DpMatrix *, DpMatrix *, Sequence *,
                                                                      Sequence *, double, PairwiseAlignmentParameters *, void *),
                                      bool alignmentHasRaggedLeftEnd, bool alignmentHasRaggedRightEnd) {
    checkPosteriorFn(sM, diagonalPosteriorProbFn, "getAlignedPairsWithoutBanding");
    std::lock_guard<std::mutex> lk(gMu);
    cpecan_ctx *ctx = gpuLocked();
    Sequence sX = { lX, cX, getXFcn, nullptr }, sY = { lY, cY, getYFcn, nullptr };
    checkSequences(sM, &sX, &sY);
    Flat f;
    if (lX > 0) f.ref.append((const char *) cX, (size_t) (lX + KMER_LENGTH - 1));
    f.refOff.push_back((int64_t) f.ref.size());
    f.events.assign((const double *) cY, (const double *) cY + lY * NB_EVENT_PARAMS);
    f.evOff.push_back(lY);
    f.anOff.push_back(0);
    f.modelId.push_back(modelIdLocked(ctx, sM));
    f.ragged.push_back((uint8_t) ((alignmentHasRaggedLeftEnd ? 1 : 0) | (alignmentHasRaggedRightEnd ? 2 : 0)));
    f.offX.push_back(0); f.offY.push_back(0); f.owner.push_back(0);
    stList *result = stList_construct3(0, (void (*)(void *)) stIntTuple_destruct);
    cpecan_params prm = paramsOf(p);
    runPosterior(ctx, hmmOf(sM), prm, CPECAN_MODE_UNBANDED, f, &result);
    return result;
}

}  // extern "C"
namespace {
// HdpHmm (inc/continuousHmm.h:27-38): the assigned k-mers and event means are COPIED here (the reference keeps pointers
// into the caller's sequences)
struct HdpH {
    Hmm base; double transitions[9]; double threshold;
    void (*addToAssignments)(Hmm *, void *, void *);
    std::vector<std::array<char, KMER_LENGTH>> *kmers; std::vector<double> *means;
};
}  // namespace
extern "C" {

void getExpectationsUsingAnchors(StateMachine *sM, Hmm *hmmExpectations, Sequence *SsX, Sequence *SsY, stList *anchorPairs,
                                 PairwiseAlignmentParameters *p,
                                 void (*diagonalCalcExpectationFcn)(StateMachine *, int64_t, DpMatrix *, DpMatrix *, Sequence *,
                                                                    Sequence *, double, PairwiseAlignmentParameters *, void *),
                                 bool alignmentHasRaggedLeftEnd, bool alignmentHasRaggedRightEnd) {
    if (diagonalCalcExpectationFcn != diagonalCalculation_Expectations)
        st_errAbort("getExpectationsUsingAnchors: only diagonalCalculation_Expectations is implemented on the GPU");
    std::lock_guard<std::mutex> lk(gMu);
    cpecan_ctx *ctx = gpuLocked();
    checkSequences(sM, SsX, SsY);
    if (SsX->length < 0) return;
    Flat f;
    if (SsY->length < 0) {
        // vanillaAlign.c:645-651 slices the complement events with a DECREASING event map: a negative length.  The
        // reference's band code then walks lX + lY diagonals of one cell each along y = 0 -- the first lX + lY k-mers
        // against no event at all (on the fixture: 244 gap-X steps, likelihood -234346.938040 in its expectation file).
        // The same item is submitted here: lX + lY k-mers, zero events, no anchors.
        const int64_t lXd = SsX->length + SsY->length;
        if (lXd <= 0) return;
        Sequence sX = *SsX, sY = *SsY;
        sX.length = lXd; sY.length = 0;
        stList *none = stList_construct();
        addRead(f, 0, modelIdLocked(ctx, sM), &sX, &sY, none, p, alignmentHasRaggedLeftEnd, alignmentHasRaggedRightEnd);
        stList_destruct(none);
    } else
    addRead(f, 0, modelIdLocked(ctx, sM), SsX, SsY, anchorPairs, p, alignmentHasRaggedLeftEnd, alignmentHasRaggedRightEnd);
    if (f.n() == 0) return;
    cpecan_batch b = batchOf(f);
    const cpecan_hmm hmm = hmmOf(sM);
    const cpecan_params prm = paramsOf(p);
    std::vector<cpecan_result> res((size_t) f.n());
    std::vector<double> e(CPECAN_N_EXPECT, 0.0);
    if (sM->type == threeStateHdp) {
        // cell_signal_updateTransAndKmerSkipExpectations2 (impl/pairwiseAligner.c:445-476): transition sums, and every
        // transition into the match state with posterior >= the container's threshold appends (k-mer, event) to its lists
        HdpH *h = (HdpH *) hmmExpectations;
        cpecan_params q = prm;
        q.threshold = h->threshold;                             // only the assignments read it in this mode
        int64_t cap = 12 * (f.evOff.back() + 32 * f.n()) + 1024;
        std::vector<int32_t> asg;
        for (int attempt = 0;; attempt++) {
            asg.assign((size_t) cap * 3, 0);
            std::fill(e.begin(), e.end(), 0.0);
            if (cpecan_cuda_hdp_expectations_batch(ctx, &hmm, &q, &b, e.data(), asg.data(), cap, res.data()) != CPECAN_OK)
                st_errAbort("cpecan_cuda_hdp_expectations_batch: %s", cpecan_cuda_last_error(ctx));
            bool overflow = false;
            int64_t need = 0;
            for (auto &r : res) {
                overflow |= (r.status & CPECAN_ITEM_PAIR_OVERFLOW) != 0; need += r.n_pairs;
                if (r.status & CPECAN_ITEM_BAD_KMER) { fprintf(stderr, "vanillaAlign - ERROR: K-mer contains character outside alphabet.\n"); exit(EXIT_FAILURE); }
            }
            if (!overflow) break;
            if (attempt == 1) st_errAbort("cpecan: assignment buffer overflow");
            cap = 8 * need + 1024;
        }
        for (int from = 0; from < 3; from++)
            for (int to = 0; to < 3; to++) hmmExpectations->addToTransitionExpectationFcn(hmmExpectations, from, to, e[(size_t) (from * 3 + to)]);
        hmmExpectations->likelihood += e[9 + NUM_OF_KMERS];
        for (int64_t i = 0; i < f.n(); i++) {
            const int32_t *t = asg.data() + 3 * res[(size_t) i].pair_off;
            for (int64_t k = 0; k < res[(size_t) i].n_pairs; k++) {
                const char *km = (const char *) SsX->elements + (t[3 * k + 1] + f.offX[(size_t) i]);
                const double *ev = (const double *) SsY->elements + (t[3 * k + 2] + f.offY[(size_t) i]) * NB_EVENT_PARAMS;
                h->addToAssignments(hmmExpectations, (void *) km, (void *) ev);
            }
        }
        return;
    }
    if (cpecan_cuda_expectations_batch(ctx, &hmm, &prm, &b, e.data(), res.data()) != CPECAN_OK)
        st_errAbort("cpecan_cuda_expectations_batch: %s", cpecan_cuda_last_error(ctx));
    // accumulate through the container's own function pointers, as cell_signal_update* do (impl/pairwiseAligner.c:426-498)
    if (sM->type == threeState) {
        for (int from = 0; from < 3; from++)
            for (int to = 0; to < 3; to++) hmmExpectations->addToTransitionExpectationFcn(hmmExpectations, from, to, e[(size_t) (from * 3 + to)]);
        for (int64_t k = 0; k < NUM_OF_KMERS; k++)
            if (e[(size_t) (9 + k)] != 0.0) hmmExpectations->addToEmissionExpectationFcn(hmmExpectations, 0, k, 0, e[(size_t) (9 + k)]);
        hmmExpectations->likelihood += e[9 + NUM_OF_KMERS];
    } else {
        for (int bin = 0; bin < 60; bin++) hmmExpectations->addToTransitionExpectationFcn(hmmExpectations, bin, 0, e[(size_t) bin]);
        hmmExpectations->likelihood += e[60];
    }
}

// ================================================================================== expectation containers (HMMs)
}  // extern "C"

namespace {
struct PairHmm { Hmm base; double transitions[9]; double kmerGap[NUM_OF_KMERS]; };                          // inc/continuousHmm.h:7-25
struct VanHmm { Hmm base; double bins[60]; double *matchModel; double *scaledMatchModel; };
void hh_addT(Hmm *h, int64_t from, int64_t to, double p) { ((HdpH *) h)->transitions[from * 3 + to] += p; }
void hh_setT(Hmm *h, int64_t from, int64_t to, double p) { ((HdpH *) h)->transitions[from * 3 + to] = p; }
double hh_getT(Hmm *h, int64_t from, int64_t to) { return ((HdpH *) h)->transitions[from * 3 + to]; }
void hh_addA(Hmm *hm, void *kmer, void *event) {                                                            // :631-636
    HdpH *h = (HdpH *) hm;
    std::array<char, KMER_LENGTH> k;
    memcpy(k.data(), kmer, KMER_LENGTH);
    h->kmers->push_back(k); h->means->push_back(*(const double *) event);
}

void ph_addT(Hmm *h, int64_t from, int64_t to, double p) { ((PairHmm *) h)->transitions[from * 3 + to] += p; }
void ph_setT(Hmm *h, int64_t from, int64_t to, double p) { ((PairHmm *) h)->transitions[from * 3 + to] = p; }
double ph_getT(Hmm *h, int64_t from, int64_t to) { return ((PairHmm *) h)->transitions[from * 3 + to]; }
void ph_addE(Hmm *h, int64_t, int64_t k, int64_t, double p) { ((PairHmm *) h)->kmerGap[k] += p; }
void ph_setE(Hmm *h, int64_t, int64_t k, int64_t, double p) { ((PairHmm *) h)->kmerGap[k] = p; }
double ph_getE(Hmm *h, int64_t, int64_t k, int64_t) { return ((PairHmm *) h)->kmerGap[k]; }
void vh_addT(Hmm *h, int64_t bin, int64_t, double p) { ((VanHmm *) h)->bins[bin] += p; }
void vh_setT(Hmm *h, int64_t bin, int64_t, double p) { ((VanHmm *) h)->bins[bin] = p; }
double vh_getT(Hmm *h, int64_t bin, int64_t) { return ((VanHmm *) h)->bins[bin]; }

bool anyNaN(const double *v, int n) { for (int i = 0; i < n; i++) if (std::isnan(v[i])) { fprintf(stdout, "GOT NaN TRANS\n"); return true; } return false; }
}  // namespace

extern "C" {

Hmm *hmmContinuous_getEmptyHmm(StateMachineType type, double pseudocount, double threshold) {               // impl/continuousHmm.c:908-944
    if (type == threeStateHdp) {                               // hdpHmm_constructEmpty :644-681
        HdpH *h = (HdpH *) calloc(1, sizeof(HdpH));
        h->base.type = threeStateHdp; h->base.stateNumber = 3; h->base.symbolSetSize = 0; h->base.matrixSize = MODEL_PARAMS;
        h->base.addToTransitionExpectationFcn = hh_addT; h->base.setTransitionFcn = hh_setT; h->base.getTransitionsExpFcn = hh_getT;
        for (double &t : h->transitions) t = pseudocount;
        h->threshold = threshold; h->addToAssignments = hh_addA;
        h->kmers = new std::vector<std::array<char, KMER_LENGTH>>(); h->means = new std::vector<double>();
        return &h->base;
    }
    if (type == threeState) {
        PairHmm *h = (PairHmm *) calloc(1, sizeof(PairHmm));
        h->base.type = threeState; h->base.stateNumber = 3; h->base.symbolSetSize = NUM_OF_KMERS; h->base.matrixSize = MODEL_PARAMS;
        h->base.addToTransitionExpectationFcn = ph_addT; h->base.setTransitionFcn = ph_setT; h->base.getTransitionsExpFcn = ph_getT;
        h->base.addToEmissionExpectationFcn = ph_addE; h->base.setEmissionExpectationFcn = ph_setE; h->base.getEmissionExpFcn = ph_getE;
        h->base.getElementIndexFcn = emissions_discrete_getKmerIndex;
        for (double &t : h->transitions) t = pseudocount;
        for (double &k : h->kmerGap) k = pseudocount;
        return &h->base;
    }
    if (type == vanilla) {
        VanHmm *h = (VanHmm *) calloc(1, sizeof(VanHmm));
        h->base.type = vanilla; h->base.stateNumber = 3; h->base.symbolSetSize = NUM_OF_KMERS; h->base.matrixSize = MODEL_PARAMS;
        h->base.addToTransitionExpectationFcn = vh_addT; h->base.setTransitionFcn = vh_setT; h->base.getTransitionsExpFcn = vh_getT;
        for (double &b : h->bins) b = pseudocount;
        h->matchModel = (double *) calloc((size_t) TABLE_LEN, sizeof(double));
        h->scaledMatchModel = (double *) calloc((size_t) TABLE_LEN, sizeof(double));
        return &h->base;
    }
    st_errAbort("hmmContinuous_getEmptyHmm - ERROR: got unsupported HMM type %i\n", (int) type);
    return nullptr;
}

void vanillaHmm_implantMatchModelsintoHmm(StateMachine *sM, Hmm *hmm) {                                     // :443-453
    VanHmm *h = (VanHmm *) hmm;
    memcpy(h->matchModel, sM->EMISSION_MATCH_PROBS, sizeof(double) * TABLE_LEN);
    memcpy(h->scaledMatchModel, sM->EMISSION_GAP_Y_PROBS, sizeof(double) * TABLE_LEN);
}

void hmmContinuous_normalize(Hmm *hmm, StateMachineType type) {                                             // :946-954
    if (type == threeState) {                                  // continuousPairHmm_normalize :173-190
        PairHmm *h = (PairHmm *) hmm;
        for (int from = 0; from < 3; from++) {
            double total = 0.0;
            for (int to = 0; to < 3; to++) total += h->transitions[from * 3 + to];
            for (int to = 0; to < 3; to++) h->transitions[from * 3 + to] /= total;
        }
        double total = 0.0;
        for (double k : h->kmerGap) total += k;
        for (double &k : h->kmerGap) k /= total;
    } else if (type == vanilla) {                              // vanillaHmm_normalizeKmerSkipBins :424-433 (joint, as there)
        VanHmm *h = (VanHmm *) hmm;
        double total = 0.0;
        for (double b : h->bins) total += b;
        for (double &b : h->bins) b /= total;
    } else {
        st_errAbort("hmmContinuous_normalize: unsupported HMM type %i", (int) type);
    }
}

int64_t hmmContinuous_howManyAssignments(Hmm *hmm) {                                                        // :944-950
    if (hmm->type != threeStateHdp) st_errAbort("hmmContinuous: this type of Hmm doesn't have assignments got type: %lld", (long long) hmm->type);
    return (int64_t) ((HdpH *) hmm)->means->size();
}

void hmmContinuous_writeToFile(const char *outFile, Hmm *hmm, StateMachineType type) {                      // :956-970, 234-271, 477-517
    FILE *fh = fopen(outFile, "w");
    if (!fh) st_errAbort("hmmContinuous_writeToFile: cannot open %s", outFile);
    if (type == threeStateHdp) {                               // hdpHmm_writeToFile :704-749
        HdpH *h = (HdpH *) hmm;
        fprintf(fh, "%i\t%lld\t%lf\t%lld\t\n", (int) hmm->type, (long long) hmm->stateNumber, h->threshold, (long long) h->means->size());
        if (!anyNaN(h->transitions, 9)) {
            for (double t : h->transitions) fprintf(fh, "%f\t", t);
            fprintf(fh, "%f\n", hmm->likelihood);
            for (double m : *h->means) fprintf(fh, "%lf\t", m);
            fprintf(fh, "\n");
            for (auto &k : *h->kmers) { fwrite(k.data(), 1, KMER_LENGTH, fh); fputc('\t', fh); }
            fprintf(fh, "\n");
        }
        fclose(fh);
        return;
    }
    fprintf(fh, "%i\t%lld\t%lld\t\n", (int) hmm->type, (long long) hmm->stateNumber, (long long) hmm->symbolSetSize);
    if (type == threeState) {
        PairHmm *h = (PairHmm *) hmm;
        if (!anyNaN(h->transitions, 9)) {
            for (double t : h->transitions) fprintf(fh, "%f\t", t);
            fprintf(fh, "%f\n", hmm->likelihood);
            for (double k : h->kmerGap) fprintf(fh, "%f\t", k);
            fprintf(fh, "\n");
        }
    } else if (type == vanilla) {
        VanHmm *h = (VanHmm *) hmm;
        if (!anyNaN(h->bins, 60)) {
            for (double b : h->bins) fprintf(fh, "%f\t", b);
            fprintf(fh, "%f\n", hmm->likelihood);
            for (int64_t i = 0; i < TABLE_LEN; i++) fprintf(fh, "%f\t", h->matchModel[i]);
            fprintf(fh, "\n");
            for (int64_t i = 0; i < TABLE_LEN; i++) fprintf(fh, "%f\t", h->scaledMatchModel[i]);
            fprintf(fh, "\n");
        }
    } else {
        st_errAbort("hmmContinuous_writeToFile - ERROR: got unsupported HMM type %i\n", (int) type);
    }
    fclose(fh);
}

// impl/continuousHmm.c:895-906: load the last M-step into a live state machine
void hmmContinuous_loadSignalHmm(const char *hmmFile, StateMachine *sM, StateMachineType type) {
    FILE *fh = fopen(hmmFile, "r");
    if (!fh) st_errAbort("hmmContinuous_loadSignalHmm: cannot open %s", hmmFile);
    std::vector<double> head = readDoublesLine(fh, hmmFile);
    if (head.size() < 3) st_errAbort("Failed to parse the header of %s", hmmFile);
    if ((int) head[0] != (int) type) st_errAbort("hmmContinuous_loadSignalHmm: %s holds an HMM of type %d, not %d", hmmFile, (int) head[0], (int) type);
    std::vector<double> l1 = readDoublesLine(fh, hmmFile);
    if (type == threeState) {
        if (l1.size() != 10) st_errAbort("Incorrect number of transitions in the input HMM file %s, got %lld instead of 10\n", hmmFile, (long long) l1.size());
        std::vector<double> l2 = readDoublesLine(fh, hmmFile);
        if (l2.size() != NUM_OF_KMERS) st_errAbort("Incorrect number of emissions in the input HMM file %s, got %lld instead of %d\n", hmmFile, (long long) l2.size(), NUM_OF_KMERS);
        // continuousPairHmm_loadTransitionsAndKmerGapProbs (:206-232)
        StateMachine3 *m = (StateMachine3 *) sM;
        auto T = [&](int from, int to) { return l1[(size_t) (from * 3 + to)]; };
        m->TRANSITION_MATCH_CONTINUE = log(T(match, match));
        m->TRANSITION_GAP_OPEN_X = log(T(match, shortGapX));
        m->TRANSITION_GAP_OPEN_Y = log(T(match, shortGapY));
        m->TRANSITION_MATCH_FROM_GAP_X = log(T(shortGapX, match));
        m->TRANSITION_GAP_EXTEND_X = log(1 - T(shortGapX, match));
        m->TRANSITION_GAP_SWITCH_TO_Y = -INFINITY;
        m->TRANSITION_MATCH_FROM_GAP_Y = log(T(shortGapY, match));
        m->TRANSITION_GAP_EXTEND_Y = log(T(shortGapY, shortGapY));
        m->TRANSITION_GAP_SWITCH_TO_X = log(T(shortGapY, shortGapX));
        for (int64_t i = 0; i < NUM_OF_KMERS; i++) sM->EMISSION_GAP_X_PROBS[i] = log(l2[(size_t) i]);
    } else if (type == threeStateHdp) {                        // hdpHmm_loadTransitions (:683-702); the assignment lines are the sampler's
        if (l1.size() != 10) st_errAbort("Incorrect number of transitions in the input HMM file %s, got %lld instead of 10\n", hmmFile, (long long) l1.size());
        StateMachine3_HDP *m = (StateMachine3_HDP *) sM;
        auto T = [&](int from, int to) { return l1[(size_t) (from * 3 + to)]; };
        m->TRANSITION_MATCH_CONTINUE = log(T(match, match));
        m->TRANSITION_GAP_OPEN_X = log(T(match, shortGapX));
        m->TRANSITION_GAP_OPEN_Y = log(T(match, shortGapY));
        m->TRANSITION_MATCH_FROM_GAP_X = log(T(shortGapX, match));
        m->TRANSITION_GAP_EXTEND_X = log(1 - T(shortGapX, match));
        m->TRANSITION_GAP_SWITCH_TO_Y = -INFINITY;
        m->TRANSITION_MATCH_FROM_GAP_Y = log(T(shortGapY, match));
        m->TRANSITION_GAP_EXTEND_Y = log(T(shortGapY, shortGapY));
        m->TRANSITION_GAP_SWITCH_TO_X = log(T(shortGapY, shortGapX));
    } else if (type == vanilla) {                              // vanillaHmm_loadKmerSkipBinExpectations (:457-466)
        if (l1.size() != 61) st_errAbort("Incorrect number of skip bins in the input HMM file %s, got %lld instead of 61\n", hmmFile, (long long) l1.size());
        for (int i = 0; i < 60; i++) sM->EMISSION_GAP_X_PROBS[i] = l1[(size_t) i];
    } else {
        st_errAbort("hmmContinuous_loadSignalHmm - ERROR: got unsupported HMM type %i\n", (int) type);
    }
    fclose(fh);
}

void hmmContinuous_destruct(Hmm *hmm, StateMachineType type) {
    if (!hmm) return;
    if (type == threeStateHdp) { delete ((HdpH *) hmm)->kmers; delete ((HdpH *) hmm)->means; }
    if (type == vanilla) { free(((VanHmm *) hmm)->matchModel); free(((VanHmm *) hmm)->scaledMatchModel); }
    free(hmm);
}

// ========================================================================================================= .npRead
// impl/nanopore.c:40-200: 13 numbers (the reference asserts 12 and reads the 13th), the 2D read, then
// (event map, events) for the template and the complement strand
NanoporeRead *nanopore_loadNanoporeReadFromFile(const char *file) {
    FILE *fh = fopen(file, "r");
    if (!fh) st_errAbort("nanopore_loadNanoporeReadFromFile: cannot open %s", file);
    std::vector<double> head = readDoublesLine(fh, file);
    if (head.size() < 13) st_errAbort("error parsing nanopore read header of %s (%lld tokens)\n", file, (long long) head.size());
    NanoporeRead *np = (NanoporeRead *) calloc(1, sizeof(NanoporeRead));
    np->readLength = (int64_t) head[0]; np->nbTemplateEvents = (int64_t) head[1]; np->nbComplementEvents = (int64_t) head[2];
    np->templateParams = { head[3], head[4], head[5], head[6], head[7] };
    np->complementParams = { head[8], head[9], head[10], head[11], head[12] };
    std::string line;
    int c;
    while ((c = fgetc(fh)) != EOF && c != '\n') if (c != ' ' && c != '\t' && c != '\r') line.push_back((char) c);
    if ((int64_t) line.size() != np->readLength) st_errAbort("2D read length %lld does not match the header (%lld)\n", (long long) line.size(), (long long) np->readLength);
    np->twoDread = strdup(line.c_str());
    auto loadMap = [&](int64_t **dst) {
        std::vector<double> v = readDoublesLine(fh, file);
        if ((int64_t) v.size() != np->readLength) st_errAbort("event map has %lld entries, expected %lld\n", (long long) v.size(), (long long) np->readLength);
        *dst = (int64_t *) malloc(sizeof(int64_t) * v.size());
        for (size_t i = 0; i < v.size(); i++) (*dst)[i] = (int64_t) v[i];
    };
    auto loadEvents = [&](double **dst, int64_t n) {
        std::vector<double> v = readDoublesLine(fh, file);
        if ((int64_t) v.size() != n * NB_EVENT_PARAMS) st_errAbort("got %lld event values, expected %lld\n", (long long) v.size(), (long long) (n * NB_EVENT_PARAMS));
        *dst = (double *) malloc(sizeof(double) * std::max<size_t>(1, v.size()));
        memcpy(*dst, v.data(), sizeof(double) * v.size());
    };
    loadMap(&np->templateEventMap); loadEvents(&np->templateEvents, np->nbTemplateEvents);
    loadMap(&np->complementEventMap); loadEvents(&np->complementEvents, np->nbComplementEvents);
    np->scaled = true;
    fclose(fh);
    return np;
}
stList *nanopore_remapAnchorPairs(stList *anchorPairs, int64_t *eventMap) {                                 // :202-212
    return nanopore_remapAnchorPairsWithOffset(anchorPairs, eventMap, -1);
}
stList *nanopore_remapAnchorPairsWithOffset(stList *pairs, int64_t *eventMap, int64_t mapOffset) {          // :214-226
    stList *out = stList_construct3(0, (void (*)(void *)) stIntTuple_destruct);
    const int64_t base = mapOffset >= 0 ? eventMap[mapOffset] : 0;
    for (int64_t i = 0; i < stList_length(pairs); i++) {
        stIntTuple *t = (stIntTuple *) stList_get(pairs, i);
        stList_append(out, stIntTuple_construct2(stIntTuple_get(t, 0), eventMap[stIntTuple_get(t, 1)] - base));
    }
    return out;
}
void nanopore_nanoporeReadDestruct(NanoporeRead *np) {
    if (!np) return;
    free(np->twoDread); free(np->templateEventMap); free(np->templateEvents); free(np->complementEventMap); free(np->complementEvents);
    free(np);
}

}  // extern "C"

// ===================================================================== the rest of the caller-side surface
// What the reference's own callers of this path (vanillaAlign.c, tests/signalPairwiseTest.c) reach besides the entry
// points above: the per-field HMM accessors of inc/continuousHmm.h:39-100, the anchor conversion of a guide cigar,
// read descaling, and the few sonLib string / cigar helpers they call directly (sonLib is un-vendored, include.mk:2).
extern "C" {

void continuousPairHmm_addToTransitionsExpectation(Hmm *hmm, int64_t from, int64_t to, double p) { ph_addT(hmm, from, to, p); }
void continuousPairHmm_setTransitionExpectation(Hmm *hmm, int64_t from, int64_t to, double p) { ph_setT(hmm, from, to, p); }
double continuousPairHmm_getTransitionExpectation(Hmm *hmm, int64_t from, int64_t to) { return ph_getT(hmm, from, to); }
void continuousPairHmm_addToKmerGapExpectation(Hmm *hmm, int64_t state, int64_t k, int64_t ignore, double p) { ph_addE(hmm, state, k, ignore, p); }
void continuousPairHmm_setKmerGapExpectation(Hmm *hmm, int64_t state, int64_t k, int64_t ignore, double p) { ph_setE(hmm, state, k, ignore, p); }
double continuousPairHmm_getKmerGapExpectation(Hmm *hmm, int64_t state, int64_t k, int64_t ignore) { return ph_getE(hmm, state, k, ignore); }
void continuousPairHmm_normalize(Hmm *hmm) { hmmContinuous_normalize(hmm, threeState); }                    // :173-190
void continuousPairHmm_destruct(Hmm *hmm) { hmmContinuous_destruct(hmm, threeState); }
// impl/continuousHmm.c:206-232: the trained transitions and k-mer skip probabilities into a live state machine
void continuousPairHmm_loadTransitionsAndKmerGapProbs(StateMachine *sM, Hmm *hmm) {
    StateMachine3 *m = (StateMachine3 *) sM;
    m->TRANSITION_MATCH_CONTINUE = log(ph_getT(hmm, match, match));
    m->TRANSITION_GAP_OPEN_X = log(ph_getT(hmm, match, shortGapX));
    m->TRANSITION_GAP_OPEN_Y = log(ph_getT(hmm, match, shortGapY));
    m->TRANSITION_MATCH_FROM_GAP_X = log(ph_getT(hmm, shortGapX, match));
    m->TRANSITION_GAP_EXTEND_X = log(1 - ph_getT(hmm, shortGapX, match));
    m->TRANSITION_GAP_SWITCH_TO_Y = -INFINITY;
    m->TRANSITION_MATCH_FROM_GAP_Y = log(ph_getT(hmm, shortGapY, match));
    m->TRANSITION_GAP_EXTEND_Y = log(ph_getT(hmm, shortGapY, shortGapY));
    m->TRANSITION_GAP_SWITCH_TO_X = log(ph_getT(hmm, shortGapY, shortGapX));
    for (int64_t i = 0; i < NUM_OF_KMERS; i++) sM->EMISSION_GAP_X_PROBS[i] = log(ph_getE(hmm, 0, i, 0));
}
void continuousPairHmm_writeToFile(Hmm *hmm, FILE *fh) {                                                    // :234-271 (no header line)
    PairHmm *h = (PairHmm *) hmm;
    if (anyNaN(h->transitions, 9)) return;
    for (double t : h->transitions) fprintf(fh, "%f\t", t);
    fprintf(fh, "%f\n", hmm->likelihood);
    for (double k : h->kmerGap) fprintf(fh, "%f\t", k);
    fprintf(fh, "\n");
}
Hmm *continuousPairHmm_loadFromFile(const char *fileName) {                                                 // :273-340
    FILE *fh = fopen(fileName, "r");
    if (!fh) st_errAbort("continuousPairHmm_loadFromFile: cannot open %s", fileName);
    std::vector<double> head = readDoublesLine(fh, fileName);
    if (head.size() < 3) st_errAbort("Failed to parse the header of %s", fileName);
    Hmm *hmm = hmmContinuous_getEmptyHmm(threeState, 0.0, 0.0);
    std::vector<double> l1 = readDoublesLine(fh, fileName), l2 = readDoublesLine(fh, fileName);
    if (l1.size() != 10) st_errAbort("Incorrect number of transitions in the input HMM file %s, got %lld instead of 10\n", fileName, (long long) l1.size());
    if (l2.size() != NUM_OF_KMERS) st_errAbort("Incorrect number of emissions in the input HMM file %s, got %lld instead of %d\n", fileName, (long long) l2.size(), NUM_OF_KMERS);
    for (int i = 0; i < 9; i++) ((PairHmm *) hmm)->transitions[i] = l1[(size_t) i];
    hmm->likelihood = l1[9];
    for (int64_t i = 0; i < NUM_OF_KMERS; i++) ((PairHmm *) hmm)->kmerGap[i] = l2[(size_t) i];
    fclose(fh);
    return hmm;
}
void vanillaHmm_addToKmerSkipBinExpectation(Hmm *hmm, int64_t bin, int64_t ignore, double p) { vh_addT(hmm, bin, ignore, p); }
void vanillaHmm_setKmerSkipBinExpectation(Hmm *hmm, int64_t bin, int64_t ignore, double p) { vh_setT(hmm, bin, ignore, p); }
double vanillaHmm_getKmerSkipBinExpectation(Hmm *hmm, int64_t bin, int64_t ignore) { return vh_getT(hmm, bin, ignore); }
void vanillaHmm_normalizeKmerSkipBins(Hmm *hmm) { hmmContinuous_normalize(hmm, vanilla); }                  // :424-433
void vanillaHmm_loadKmerSkipBinExpectations(StateMachine *sM, Hmm *hmm) {                                   // :457-466
    for (int i = 0; i < 60; i++) sM->EMISSION_GAP_X_PROBS[i] = vh_getT(hmm, i, 0);
}
void vanillaHmm_destruct(Hmm *hmm) { hmmContinuous_destruct(hmm, vanilla); }

// impl/stateMachine.c:141-153
int64_t emissions_discrete_getKmerIndexFromKmer(void *kmer) {
    char k[KMER_LENGTH + 1];
    memcpy(k, kmer, KMER_LENGTH);
    k[KMER_LENGTH] = 0;
    return emissions_discrete_getKmerIndex(k);
}

// impl/nanopore.c:34-38, 228-236 -- the loop steps by NB_EVENT_PARAMS over nb_events INDICES, as the reference's does
// (only the first third of the events is descaled there; kept: this is what a caller of the reference gets)
void nanopore_descaleNanoporeRead(NanoporeRead *np) {
    for (int64_t i = 0; i < np->nbTemplateEvents; i += NB_EVENT_PARAMS) np->templateEvents[i] = (np->templateEvents[i] - np->templateParams.shift) / np->templateParams.scale;
    for (int64_t i = 0; i < np->nbComplementEvents; i += NB_EVENT_PARAMS) np->complementEvents[i] = (np->complementEvents[i] - np->complementParams.shift) / np->complementParams.scale;
    np->scaled = false;
}

// ---- sonLib helpers the callers use directly
void *st_malloc(size_t size) { void *p = malloc(size ? size : 1); if (!p) st_errAbort("st_malloc: out of memory"); return p; }
void *st_calloc(int64_t n, size_t size) { void *p = calloc((size_t) (n > 0 ? n : 1), size ? size : 1); if (!p) st_errAbort("st_calloc: out of memory"); return p; }
void st_uglyf(const char *format, ...) { va_list ap; va_start(ap, format); vfprintf(stderr, format, ap); va_end(ap); }
void st_logDebug(const char *, ...) {}
char *stString_copy(const char *s) { return s ? strdup(s) : nullptr; }
char *stString_print(const char *format, ...) {
    va_list ap, ap2;
    va_start(ap, format);
    va_copy(ap2, ap);
    const int n = vsnprintf(nullptr, 0, format, ap);
    va_end(ap);
    char *out = (char *) st_malloc((size_t) n + 1);
    vsnprintf(out, (size_t) n + 1, format, ap2);
    va_end(ap2);
    return out;
}
char *stString_getSubString(const char *s, int64_t start, int64_t length) {
    char *out = (char *) st_malloc((size_t) length + 1);
    memcpy(out, s + start, (size_t) length);
    out[length] = 0;
    return out;
}
char *stString_replace(const char *s, const char *what, const char *with) {
    std::string out;
    const size_t nw = strlen(what);
    while (*s) {
        if (nw && strncmp(s, what, nw) == 0) { out += with; s += nw; }
        else out.push_back(*s++);
    }
    return strdup(out.c_str());
}
char *stString_reverseComplementString(const char *s) {
    const size_t n = strlen(s);
    char *r = (char *) st_malloc(n + 1);
    for (size_t i = 0; i < n; i++) {
        const char c = s[n - 1 - i];
        char o = c;
        switch (c) {
            case 'A': o = 'T'; break; case 'C': o = 'G'; break; case 'G': o = 'C'; break; case 'T': o = 'A'; break;
            case 'a': o = 't'; break; case 'c': o = 'g'; break; case 'g': o = 'c'; break; case 't': o = 'a'; break;
        }
        r[i] = o;
    }
    r[n] = 0;
    return r;
}
char *stFile_getLineFromFile(FILE *fh) {                       // the next line without its newline; NULL at end of file
    std::string line;
    int c = fgetc(fh);
    if (c == EOF) return nullptr;
    for (; c != EOF && c != '\n'; c = fgetc(fh)) line.push_back((char) c);
    return strdup(line.c_str());
}

// ---- the guide alignment (exonerate cigar: "cigar: query qStart qEnd qStrand target tStart tEnd tStrand score (op len)*";
// sonLib's cigarRead puts the QUERY into contig2 / start2 / end2 and the target into contig1, M -> MATCH, D -> INDEL_X,
// I -> INDEL_Y, '+' -> strand 1)
static void destroyOp(void *op) { free(op); }
struct PairwiseAlignment *cigarRead(FILE *fh) {
    char *line = stFile_getLineFromFile(fh);
    if (!line) return nullptr;
    std::vector<std::string> tok;
    for (char *t = strtok(line, " \t\r"); t; t = strtok(nullptr, " \t\r")) tok.push_back(t);
    free(line);
    if (tok.size() < 10 || tok[0] != "cigar:") return nullptr;
    struct PairwiseAlignment *pA = (struct PairwiseAlignment *) st_malloc(sizeof(struct PairwiseAlignment));
    pA->contig2 = strdup(tok[1].c_str()); pA->start2 = atoll(tok[2].c_str()); pA->end2 = atoll(tok[3].c_str()); pA->strand2 = tok[4] == "+";
    pA->contig1 = strdup(tok[5].c_str()); pA->start1 = atoll(tok[6].c_str()); pA->end1 = atoll(tok[7].c_str()); pA->strand1 = tok[8] == "+";
    pA->score = (float) atof(tok[9].c_str());
    struct List *ops = (struct List *) st_malloc(sizeof(struct List));
    const size_t nOps = (tok.size() - 10) / 2;
    ops->list = (void **) st_malloc(sizeof(void *) * (nOps ? nOps : 1));
    ops->length = 0; ops->maxLength = (int64_t) nOps; ops->destroyElement = destroyOp;
    for (size_t i = 10; i + 1 < tok.size(); i += 2) {
        struct AlignmentOperation *op = (struct AlignmentOperation *) st_malloc(sizeof(struct AlignmentOperation));
        op->opType = tok[i] == "M" ? PAIRWISE_MATCH : (tok[i] == "D" ? PAIRWISE_INDEL_X : PAIRWISE_INDEL_Y);
        op->length = atoll(tok[i + 1].c_str());
        op->score = 0.f;
        ops->list[ops->length++] = op;
    }
    pA->operationList = ops;
    return pA;
}
void destructPairwiseAlignment(struct PairwiseAlignment *pA) {
    if (!pA) return;
    for (int64_t i = 0; i < pA->operationList->length; i++) free(pA->operationList->list[i]);
    free(pA->operationList->list); free(pA->operationList);
    free(pA->contig1); free(pA->contig2); free(pA);
}
void checkPairwiseAlignment(struct PairwiseAlignment *pA) {    // the operations must add up to the two intervals
    int64_t lx = 0, ly = 0;
    for (int64_t i = 0; i < pA->operationList->length; i++) {
        struct AlignmentOperation *op = (struct AlignmentOperation *) pA->operationList->list[i];
        if (op->opType != PAIRWISE_INDEL_Y) lx += op->length;
        if (op->opType != PAIRWISE_INDEL_X) ly += op->length;
    }
    if (llabs(pA->end1 - pA->start1) != lx || llabs(pA->end2 - pA->start2) != ly)
        st_errAbort("checkPairwiseAlignment: cigar operations do not match the intervals");
}
// impl/pairwiseAligner.c:1039-1063: the match columns of the guide alignment, `trim` dropped at both ends of every match
stList *convertPairwiseForwardStrandAlignmentToAnchorPairs(struct PairwiseAlignment *pA, int64_t trim) {
    stList *pairs = stList_construct3(0, (void (*)(void *)) stIntTuple_destruct);
    int64_t j = pA->start1, k = pA->start2;
    for (int64_t i = 0; i < pA->operationList->length; i++) {
        struct AlignmentOperation *op = (struct AlignmentOperation *) pA->operationList->list[i];
        if (op->opType == PAIRWISE_MATCH) for (int64_t l = trim; l < op->length - trim; l++) stList_append(pairs, stIntTuple_construct2(j + l, k + l));
        if (op->opType != PAIRWISE_INDEL_Y) j += op->length;
        if (op->opType != PAIRWISE_INDEL_X) k += op->length;
    }
    return pairs;
}

// ---- "methods tested and possibly useful elsewhere" (inc/pairwiseAligner.h:193-277): the per-cell / per-diagonal
// pieces of the CPU DP.  Here the diagonals live on the GPU and are never materialised on the host; the symbols exist
// so that code written against the header links, and abort (no CPU path) if anything calls them.
#define CPECAN_NO_HOST_DP(name, ret, args) ret name args { noCpuPath(#name); abort(); }
typedef struct _dpDiagonal DpDiagonal;
CPECAN_NO_HOST_DP(cell_calculateForward, void, (StateMachine *, double *, double *, double *, double *, void *, void *, void *))
CPECAN_NO_HOST_DP(cell_calculateBackward, void, (StateMachine *, double *, double *, double *, double *, void *, void *, void *))
CPECAN_NO_HOST_DP(dpDiagonal_construct, DpDiagonal *, (Diagonal, int64_t))
CPECAN_NO_HOST_DP(dpDiagonal_clone, DpDiagonal *, (DpDiagonal *))
CPECAN_NO_HOST_DP(dpDiagonal_equals, bool, (DpDiagonal *, DpDiagonal *))
CPECAN_NO_HOST_DP(dpDiagonal_destruct, void, (DpDiagonal *))
CPECAN_NO_HOST_DP(dpDiagonal_getCell, double *, (DpDiagonal *, int64_t))
CPECAN_NO_HOST_DP(dpDiagonal_dotProduct, double, (DpDiagonal *, DpDiagonal *))
CPECAN_NO_HOST_DP(dpDiagonal_zeroValues, void, (DpDiagonal *))
CPECAN_NO_HOST_DP(dpDiagonal_initialiseValues, void, (DpDiagonal *, StateMachine *, double (*)(StateMachine *, int64_t)))
CPECAN_NO_HOST_DP(dpMatrix_construct, DpMatrix *, (int64_t, int64_t))
CPECAN_NO_HOST_DP(dpMatrix_destruct, void, (DpMatrix *))
CPECAN_NO_HOST_DP(dpMatrix_getDiagonal, DpDiagonal *, (DpMatrix *, int64_t))
CPECAN_NO_HOST_DP(dpMatrix_getActiveDiagonalNumber, int64_t, (DpMatrix *))
CPECAN_NO_HOST_DP(dpMatrix_createDiagonal, DpDiagonal *, (DpMatrix *, Diagonal))
CPECAN_NO_HOST_DP(dpMatrix_deleteDiagonal, void, (DpMatrix *, int64_t))
CPECAN_NO_HOST_DP(diagonalCalculationForward, void, (StateMachine *, int64_t, DpMatrix *, Sequence *, Sequence *))
CPECAN_NO_HOST_DP(diagonalCalculationBackward, void, (StateMachine *, int64_t, DpMatrix *, Sequence *, Sequence *))
CPECAN_NO_HOST_DP(diagonalCalculationTotalProbability, double, (StateMachine *, int64_t, DpMatrix *, DpMatrix *, Sequence *, Sequence *))
// impl/pairwiseAligner.c:385-397: these two are plain arithmetic and are provided
double cell_dotProduct(double *cell1, double *cell2, int64_t stateNumber) {
    double total = cell1[0] + cell2[0];
    for (int64_t i = 1; i < stateNumber; i++) total = logAdd(total, cell1[i] + cell2[i]);
    return total;
}
double cell_dotProduct2(double *cell, StateMachine *sM, double (*getStateValue)(StateMachine *, int64_t)) {
    double total = cell[0] + getStateValue(sM, 0);
    for (int64_t i = 1; i < sM->stateNumber; i++) total = logAdd(total, cell[i] + getStateValue(sM, i));
    return total;
}

// getPosteriorProbsWithBanding (impl/pairwiseAligner.c:870-1006) hands every diagonal to a CALLBACK on the host; here the
// two callbacks the reference ships select the device mode, and the results come back as lists.  A caller that used the
// callback form directly gets the aligned pairs through extraArgs exactly as diagonalCalculationPosteriorMatchProbs would
// have appended them (extraArgs = the stList of aligned pairs, :756-795).
void getPosteriorProbsWithBanding(StateMachine *sM, stList *anchorPairs, Sequence *sX, Sequence *sY, PairwiseAlignmentParameters *p,
                                  bool alignmentHasRaggedLeftEnd, bool alignmentHasRaggedRightEnd,
                                  void (*diagonalPosteriorProbFn)(StateMachine *, int64_t, DpMatrix *, DpMatrix *, Sequence *, Sequence *,
                                                                  double, PairwiseAlignmentParameters *, void *),
                                  void *extraArgs) {
    checkPosteriorFn(sM, diagonalPosteriorProbFn, "getPosteriorProbsWithBanding");
    // one region, no splitting: splitMatrixBiggerThanThis is lifted for this call
    PairwiseAlignmentParameters q = *p;
    q.splitMatrixBiggerThanThis = INT64_MAX;
    stList *res = getAlignedPairsUsingAnchors(sM, sX, sY, anchorPairs, &q, diagonalPosteriorProbFn, alignmentHasRaggedLeftEnd, alignmentHasRaggedRightEnd);
    // getAlignedPairsUsingAnchors reversed the region's list (the coordinate-correction callback pops); undo it: the
    // callback form appends in traceback order -- to the list the reference's callbacks take out of extraArgs,
    // ((void **) extraArgs)[0] (impl/pairwiseAligner.c:761; callers pass "void *extraArgs[] = { alignedPairs, ... }")
    stList *out = (stList *) ((void **) extraArgs)[0];
    for (int64_t i = stList_length(res) - 1; i >= 0; i--) { stList_append(out, stList_get(res, i)); stList_set(res, i, nullptr); }
    stList_setDestructor(res, nullptr);
    stList_destruct(res);
}

// impl/pairwiseAligner.c:1356-1422: the regions between large anchor gaps one after the other through the callback form,
// coordinateCorrectionFn(x1, y1, extraArgs) after each (getAlignedPairsUsingAnchors passes the function that shifts the
// region's pairs back and pops them onto the final list, :1447-1454).  One GPU call per region here; the batched entry
// points submit all regions at once.
void getPosteriorProbsWithBandingSplittingAlignmentsByLargeGaps(
        StateMachine *sM, stList *anchorPairs, Sequence *SsX, Sequence *SsY, PairwiseAlignmentParameters *p,
        bool alignmentHasRaggedLeftEnd, bool alignmentHasRaggedRightEnd,
        void (*diagonalPosteriorProbFn)(StateMachine *, int64_t, DpMatrix *, DpMatrix *, Sequence *, Sequence *, double,
                                        PairwiseAlignmentParameters *, void *),
        void (*coordinateCorrectionFn)(int64_t, int64_t, void *), void *extraArgs) {
    stList *split = getSplitPoints(anchorPairs, SsX->length, SsY->length, p->splitMatrixBiggerThanThis, alignmentHasRaggedLeftEnd,
                                   alignmentHasRaggedRightEnd);
    const int64_t nS = stList_length(split), nA = stList_length(anchorPairs);
    int64_t j = 0;
    for (int64_t i = 0; i < nS; i++) {
        stIntTuple *r = (stIntTuple *) stList_get(split, i);
        const int64_t x1 = stIntTuple_get(r, 0), y1 = stIntTuple_get(r, 1), x2 = stIntTuple_get(r, 2), y2 = stIntTuple_get(r, 3);
        Sequence *sX3 = SsX->sliceFcn(SsX, x1, x2 - x1), *sY3 = SsY->sliceFcn(SsY, y1, y2 - y1);
        stList *sub = stList_construct3(0, (void (*)(void *)) stIntTuple_destruct);
        while (j < nA) {
            stIntTuple *a = (stIntTuple *) stList_get(anchorPairs, j);
            const int64_t x = stIntTuple_get(a, 0), y = stIntTuple_get(a, 1);
            if (x + y >= x2 + y2) break;
            stList_append(sub, stIntTuple_construct2(x - x1, y - y1));
            j++;
        }
        getPosteriorProbsWithBanding(sM, sub, sX3, sY3, p, alignmentHasRaggedLeftEnd || i > 0, alignmentHasRaggedRightEnd || i < nS - 1,
                                     diagonalPosteriorProbFn, extraArgs);
        if (coordinateCorrectionFn != nullptr) coordinateCorrectionFn(x1, y1, extraArgs);
        stList_destruct(sub);
        sequence_sequenceDestroy(sX3); sequence_sequenceDestroy(sY3);
    }
    stList_destruct(split);
}

}  // extern "C"
