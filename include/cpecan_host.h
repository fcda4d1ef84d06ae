/*
 * cpecan_host.h -- the reference's C API for the signal hot path, re-declared for libcpecan_host.so.
 *
 * This is the drop-in boundary on the CALLER's side: same names, argument meaning, struct layouts
 * and error behaviour as the reference's cPecanLib.a for this path: the reference's UNMODIFIED vanillaAlign.c, compiled
 * against the reference's own headers, links against libcpecan_host.so (plus stubs for the out-of-scope HDP symbols) and
 * reproduces the reference binary's output -- tests/test_vanilla_align_drop_in.py does exactly that.  Every entry point cites the reference declaration it replaces.  The DP itself runs in
 * libcpecan_cuda.so (include/cpecan_cuda.h); there is NO CPU fallback: getAlignedPairsUsingAnchors and friends abort
 * (st_errAbort semantics: message + abort) when no CUDA device is usable.
 *
 * sonLib is an un-vendored dependency of the reference (include.mk:2); the few container types the API exposes
 * (stList, stIntTuple) are provided here with the sonLib function names.
 */
#ifndef CPECAN_HOST_H_
#define CPECAN_HOST_H_

#include <stdbool.h>
#include <stdint.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------- sonLib subset (stList / stIntTuple) */
typedef struct _stList stList;
typedef int64_t stIntTuple;                                  /* sonLib: an int64 array whose slot 0 is the length */
stList *stList_construct(void);
stList *stList_construct3(int64_t size, void (*destructElement)(void *));
void stList_destruct(stList *list);
int64_t stList_length(stList *list);
void *stList_get(stList *list, int64_t index);
void stList_set(stList *list, int64_t index, void *item);
void stList_append(stList *list, void *item);
void *stList_pop(stList *list);
void stList_sort(stList *list, int (*cmpFn)(const void *a, const void *b));
void stList_setDestructor(stList *list, void (*destructElement)(void *));
stIntTuple *stIntTuple_construct2(int64_t a, int64_t b);
stIntTuple *stIntTuple_construct3(int64_t a, int64_t b, int64_t c);
stIntTuple *stIntTuple_construct4(int64_t a, int64_t b, int64_t c, int64_t d);
void stIntTuple_destruct(stIntTuple *t);
int64_t stIntTuple_get(stIntTuple *t, int64_t index);
int64_t stIntTuple_length(stIntTuple *t);
int stIntTuple_cmpFn(const void *a, const void *b);
int stIntTuple_equalsFn(const void *a, const void *b);
void st_errAbort(const char *format, ...);                   /* message on stderr + abort() */
void *st_malloc(size_t size);
void *st_calloc(int64_t n, size_t size);
void st_uglyf(const char *format, ...);                      /* printf to stderr */
void st_logDebug(const char *format, ...);
char *stString_copy(const char *s);
char *stString_print(const char *format, ...);
char *stString_getSubString(const char *s, int64_t start, int64_t length);
char *stString_replace(const char *s, const char *what, const char *with);
char *stString_reverseComplementString(const char *s);
char *stFile_getLineFromFile(FILE *fileHandle);              /* the next line without its newline, NULL at end of file */

/* sonLib's pairwiseAlignment.h: the guide alignment vanillaAlign.c reads from stdin (exonerate cigar) */
#define PAIRWISE_INDEL_X 0
#define PAIRWISE_INDEL_Y 1
#define PAIRWISE_MATCH 2
struct List { void **list; int64_t length; int64_t maxLength; void (*destroyElement)(void *); };
struct AlignmentOperation { int64_t opType; int64_t length; float score; };
struct PairwiseAlignment {
    char *contig1; int64_t start1; int64_t end1; int64_t strand1;
    char *contig2; int64_t start2; int64_t end2; int64_t strand2;
    float score;
    struct List *operationList;
};
struct PairwiseAlignment *cigarRead(FILE *fileHandle);
void destructPairwiseAlignment(struct PairwiseAlignment *pA);
void checkPairwiseAlignment(struct PairwiseAlignment *pA);

/* ---------------------------------------------------------------- inc/pairwiseAligner.h */
#define PAIR_ALIGNMENT_PROB_1 10000000                       /* inc/pairwiseAligner.h:26 */
#define LOG_ZERO (-INFINITY)

typedef enum { nucleotide = 0, kmer = 1, event = 2 } SequenceType;            /* :29-33 */

typedef struct _sequence Sequence;                                             /* :35-42 */
struct _sequence {
    int64_t length;
    void *elements;
    void *(*get)(void *elements, int64_t index);
    Sequence *(*sliceFcn)(Sequence *, int64_t, int64_t);
};
Sequence *sequence_construct(int64_t length, void *elements, void *(*getFcn)(void *, int64_t));             /* :50 */
Sequence *sequence_construct2(int64_t length, void *elements, void *(*getFcn)(void *, int64_t),
                              Sequence *(*sliceFcn)(Sequence *, int64_t, int64_t));                         /* :52-53 */
Sequence *sequence_sliceNucleotideSequence2(Sequence *inputSequence, int64_t start, int64_t sliceLength);   /* :59 */
Sequence *sequence_sliceEventSequence2(Sequence *inputSequence, int64_t start, int64_t sliceLength);        /* :61 */
void sequence_sequenceDestroy(Sequence *seq);                                                               /* :63 */
void *sequence_getKmer(void *elements, int64_t index);                                                      /* :67 */
void *sequence_getKmer2(void *elements, int64_t index);                                                     /* :70 */
void *sequence_getKmer3(void *elements, int64_t index);                                                     /* :72 */
void *sequence_getBase(void *elements, int64_t index);                                                      /* :66; nucleotide sequences (not a signal-path accessor) */
void sequence_padSequence(Sequence *sequence);                                                              /* :56 */
void *sequence_getEvent(void *elements, int64_t index);                                                     /* :75 */
int64_t sequence_correctSeqLength(int64_t length, SequenceType type);                                       /* :77 */

typedef struct _pairwiseAlignmentBandingParameters {                                                        /* :80-91 */
    double threshold;
    int64_t minDiagsBetweenTraceBack;
    int64_t traceBackDiagonals;
    int64_t diagonalExpansion;
    int64_t constraintDiagonalTrim;
    int64_t anchorMatrixBiggerThanThis;
    int64_t repeatMaskMatrixBiggerThanThis;
    int64_t splitMatrixBiggerThanThis;
    bool alignAmbiguityCharacters;
    float gapGamma;
} PairwiseAlignmentParameters;
PairwiseAlignmentParameters *pairwiseAlignmentBandingParameters_construct(void);                            /* :93 */
void pairwiseAlignmentBandingParameters_destruct(PairwiseAlignmentParameters *p);                           /* :95 */

typedef struct _diagonal { int64_t xay, xmyL, xmyR; } Diagonal;                                             /* :139-143 */
Diagonal diagonal_construct(int64_t xay, int64_t xmyL, int64_t xmyR);   /* aborts on odd parity / xmyL > xmyR */
int64_t diagonal_getXay(Diagonal diagonal);
int64_t diagonal_getMinXmy(Diagonal diagonal);
int64_t diagonal_getMaxXmy(Diagonal diagonal);
int64_t diagonal_getWidth(Diagonal diagonal);
int64_t diagonal_getXCoordinate(int64_t xay, int64_t xmy);
int64_t diagonal_getYCoordinate(int64_t xay, int64_t xmy);
int64_t diagonal_equals(Diagonal diagonal1, Diagonal diagonal2);

typedef struct _band Band;                                                                                  /* :165-171 */
Band *band_construct(stList *anchorPairs, int64_t lX, int64_t lY, int64_t expansion);
void band_destruct(Band *band);
typedef struct _bandIterator BandIterator;                                                                  /* :175-185 */
BandIterator *bandIterator_construct(Band *band);
void bandIterator_destruct(BandIterator *bandIterator);
BandIterator *bandIterator_clone(BandIterator *bandIterator);
Diagonal bandIterator_getNext(BandIterator *bandIterator);
Diagonal bandIterator_getPrevious(BandIterator *bandIterator);

double logAdd(double x, double y);                                                                          /* :191 */

typedef struct _stateMachine StateMachine;
typedef struct _hmm Hmm;
typedef struct _dpMatrix DpMatrix;                            /* opaque: the forward/backward matrices live on the GPU */

/* The two per-diagonal callbacks select the work mode by pointer identity; calling them directly
 * aborts -- the diagonals are never materialised on the host. */
void diagonalCalculationPosteriorMatchProbs(StateMachine *sM, int64_t xay, DpMatrix *forwardDpMatrix,       /* :246-249 */
                                            DpMatrix *backwardDpMatrix, Sequence *sX, Sequence *sY,
                                            double totalProbability, PairwiseAlignmentParameters *p, void *extraArgs);
void diagonalCalculation_Expectations(StateMachine *sM, int64_t xay, DpMatrix *forwardDpMatrix,             /* :260-264 */
                                      DpMatrix *backwardDpMatrix, Sequence *sX, Sequence *sY,
                                      double totalProbability, PairwiseAlignmentParameters *p, void *extraArgs);

stList *getAlignedPairsUsingAnchors(StateMachine *sM, Sequence *SsX, Sequence *SsY, stList *anchorPairs,    /* :288-296 */
                                    PairwiseAlignmentParameters *p,
                                    void (*diagonalPosteriorProbFn)(StateMachine *, int64_t, DpMatrix *, DpMatrix *,
                                                                    Sequence *, Sequence *, double,
                                                                    PairwiseAlignmentParameters *, void *),
                                    bool alignmentHasRaggedLeftEnd, bool alignmentHasRaggedRightEnd);
stList *getAlignedPairsWithoutBanding(StateMachine *sM, void *cX, void *cY, int64_t lX, int64_t lY,          /* :279-286 */
                                      PairwiseAlignmentParameters *p,
                                      void *(*getXFcn)(void *, int64_t), void *(*getYFcn)(void *, int64_t),
                                      void (*diagonalPosteriorProbFn)(StateMachine *, int64_t, DpMatrix *, DpMatrix *,
                                                                      Sequence *, Sequence *, double,
                                                                      PairwiseAlignmentParameters *, void *),
                                      bool alignmentHasRaggedLeftEnd, bool alignmentHasRaggedRightEnd);
void getExpectationsUsingAnchors(StateMachine *sM, Hmm *hmmExpectations, Sequence *SsX, Sequence *SsY,      /* :299-311 */
                                 stList *anchorPairs, PairwiseAlignmentParameters *p,
                                 void (*diagonalCalcExpectationFcn)(StateMachine *, int64_t, DpMatrix *, DpMatrix *,
                                                                    Sequence *, Sequence *, double,
                                                                    PairwiseAlignmentParameters *, void *),
                                 bool alignmentHasRaggedLeftEnd, bool alignmentHasRaggedRightEnd);

void diagonalCalculationMultiPosteriorMatchProbs(StateMachine *sM, int64_t xay, DpMatrix *forwardDpMatrix,  /* :251-254 */
                                                 DpMatrix *backwardDpMatrix, Sequence *sX, Sequence *sY,
                                                 double totalProbability, PairwiseAlignmentParameters *p, void *extraArgs);
/* The callback form (:266-277): the two callbacks above select the device mode; the aligned pairs are appended, in
 * traceback order, to the list the reference's callbacks take out of extraArgs -- ((void **) extraArgs)[0], i.e. callers
 * pass "void *extraArgs[] = { alignedPairs }" (impl/pairwiseAligner.c:761, tests/pairwiseAlignerTest.c:447-449). */
void getPosteriorProbsWithBanding(StateMachine *sM, stList *anchorPairs, Sequence *sX, Sequence *sY,
                                  PairwiseAlignmentParameters *p, bool alignmentHasRaggedLeftEnd, bool alignmentHasRaggedRightEnd,
                                  void (*diagonalPosteriorProbFn)(StateMachine *, int64_t, DpMatrix *, DpMatrix *,
                                                                  Sequence *, Sequence *, double,
                                                                  PairwiseAlignmentParameters *, void *),
                                  void *extraArgs);
/* :331-338: the same region by region between large anchor gaps (getSplitPoints), coordinateCorrectionFn(x1, y1, extraArgs)
 * after each region */
void getPosteriorProbsWithBandingSplittingAlignmentsByLargeGaps(
        StateMachine *sM, stList *anchorPairs, Sequence *SsX, Sequence *SsY, PairwiseAlignmentParameters *p,
        bool alignmentHasRaggedLeftEnd, bool alignmentHasRaggedRightEnd,
        void (*diagonalPosteriorProbFn)(StateMachine *, int64_t, DpMatrix *, DpMatrix *, Sequence *, Sequence *, double,
                                        PairwiseAlignmentParameters *, void *),
        void (*coordinateCorrectionFn)(int64_t, int64_t, void *), void *extraArgs);
stList *convertPairwiseForwardStrandAlignmentToAnchorPairs(struct PairwiseAlignment *pA, int64_t trim);     /* :109 */

/* "Methods tested and possibly useful elsewhere" (:193-245): the per-cell / per-diagonal pieces of the CPU DP.  The
 * diagonals live on the GPU here; the symbols exist so that code written against the reference header links, and all
 * but the two dot products (plain arithmetic) abort when called -- there is no CPU path. */
typedef struct _dpDiagonal DpDiagonal;
void cell_calculateForward(StateMachine *sM, double *current, double *lower, double *middle, double *upper, void *cX, void *cY, void *extraArgs);
void cell_calculateBackward(StateMachine *sM, double *current, double *lower, double *middle, double *upper, void *cX, void *cY, void *extraArgs);
double cell_dotProduct(double *cell1, double *cell2, int64_t stateNumber);
double cell_dotProduct2(double *cell1, StateMachine *sM, double (*getStateValue)(StateMachine *, int64_t));
DpDiagonal *dpDiagonal_construct(Diagonal diagonal, int64_t stateNumber);
DpDiagonal *dpDiagonal_clone(DpDiagonal *diagonal);
bool dpDiagonal_equals(DpDiagonal *diagonal1, DpDiagonal *diagonal2);
void dpDiagonal_destruct(DpDiagonal *dpDiagonal);
double *dpDiagonal_getCell(DpDiagonal *dpDiagonal, int64_t xmy);
double dpDiagonal_dotProduct(DpDiagonal *diagonal1, DpDiagonal *diagonal2);
void dpDiagonal_zeroValues(DpDiagonal *diagonal);
void dpDiagonal_initialiseValues(DpDiagonal *diagonal, StateMachine *sM, double (*getStateValue)(StateMachine *, int64_t));
DpMatrix *dpMatrix_construct(int64_t diagonalNumber, int64_t stateNumber);
void dpMatrix_destruct(DpMatrix *dpMatrix);
DpDiagonal *dpMatrix_getDiagonal(DpMatrix *dpMatrix, int64_t xay);
int64_t dpMatrix_getActiveDiagonalNumber(DpMatrix *dpMatrix);
DpDiagonal *dpMatrix_createDiagonal(DpMatrix *dpMatrix, Diagonal diagonal);
void dpMatrix_deleteDiagonal(DpMatrix *dpMatrix, int64_t xay);
void diagonalCalculationForward(StateMachine *sM, int64_t xay, DpMatrix *dpMatrix, Sequence *sX, Sequence *sY);
void diagonalCalculationBackward(StateMachine *sM, int64_t xay, DpMatrix *dpMatrix, Sequence *sX, Sequence *sY);
double diagonalCalculationTotalProbability(StateMachine *sM, int64_t xay, DpMatrix *forwardDpMatrix, DpMatrix *backwardDpMatrix,
                                           Sequence *sX, Sequence *sY);

int sortByXPlusYCoordinate(const void *i, const void *j);                                                   /* :316 */
int sortByXPlusYCoordinate2(const void *i, const void *j);                                                  /* :318 */
stList *filterToRemoveOverlap(stList *overlappingPairs);                                                    /* :324 */
stList *getSplitPoints(stList *anchorPairs, int64_t lX, int64_t lY, int64_t maxMatrixSize,                  /* :328-329 */
                       bool alignmentHasRaggedLeftEnd, bool alignmentHasRaggedRightEnd);

/* ---------------------------------------------------------------- inc/stateMachine.h */
#define KMER_LENGTH 6                                         /* inc/emissionMatrix.h:4-5 */
#define NUM_OF_KMERS 4096
#define MODEL_PARAMS 5

typedef enum { fiveState = 0, fiveStateAsymmetric = 1, threeState = 2, threeStateAsymmetric = 3, vanilla = 4,
               echelon = 5, fourState = 6, threeStateHdp = 7 } StateMachineType;                            /* :20-29 */
typedef enum { match = 0, shortGapX = 1, shortGapY = 2, longGapX = 3, longGapY = 4 } State;                 /* :31-33 */
#ifdef __cplusplus
typedef enum _strand { template_ = 0, complement = 1 } Strand;   /* `template` is reserved in C++ (:35-38) */
#else
typedef enum _strand { template = 0, complement = 1 } Strand;
#endif

struct _hmm {                                                                                               /* :47-74 */
    double likelihood;
    StateMachineType type;
    int64_t stateNumber;
    int64_t symbolSetSize;
    int64_t matrixSize;
    void (*addToTransitionExpectationFcn)(Hmm *hmm, int64_t from, int64_t to, double p);
    void (*setTransitionFcn)(Hmm *hmm, int64_t from, int64_t to, double p);
    double (*getTransitionsExpFcn)(Hmm *hmm, int64_t from, int64_t to);
    void (*addToEmissionExpectationFcn)(Hmm *hmm, int64_t state, int64_t x, int64_t y, double p);
    void (*setEmissionExpectationFcn)(Hmm *hmm, int64_t state, int64_t x, int64_t y, double p);
    double (*getEmissionExpFcn)(Hmm *hmm, int64_t state, int64_t x, int64_t y);
    int64_t (*getElementIndexFcn)(void *);
};

struct _stateMachine {                                                                                      /* :76-102 */
    StateMachineType type;
    int64_t stateNumber;
    int64_t matchState;
    int64_t parameterSetSize;
    double *EMISSION_MATCH_PROBS;
    double *EMISSION_GAP_X_PROBS;
    double *EMISSION_GAP_Y_PROBS;
    double (*startStateProb)(StateMachine *sM, int64_t state);
    double (*endStateProb)(StateMachine *sM, int64_t state);
    double (*raggedEndStateProb)(StateMachine *sM, int64_t state);
    double (*raggedStartStateProb)(StateMachine *sM, int64_t state);
    void (*cellCalculate)(StateMachine *sM, double *current, double *lower, double *middle, double *upper, void *cX,
                          void *cY, void (*doTransition)(double *, double *, int64_t, int64_t, double, double, void *),
                          void *extraArgs);                   /* aborts: the cells are computed on the GPU */
    void (*cellCalculateUpdateExpectations)(double *fromCells, double *toCells, int64_t from, int64_t to, double eP,
                                            double tP, void *extraArgs);
};

typedef struct _StateMachine3 {                                                                             /* :176-195 */
    StateMachine model;
    double TRANSITION_MATCH_CONTINUE;
    double TRANSITION_MATCH_FROM_GAP_X;
    double TRANSITION_MATCH_FROM_GAP_Y;
    double TRANSITION_GAP_OPEN_X;
    double TRANSITION_GAP_OPEN_Y;
    double TRANSITION_GAP_EXTEND_X;
    double TRANSITION_GAP_EXTEND_Y;
    double TRANSITION_GAP_SWITCH_TO_X;
    double TRANSITION_GAP_SWITCH_TO_Y;
    double (*getXGapProbFcn)(const double *emissionXGapProbs, void *i);
    double (*getYGapProbFcn)(const double *emissionYGapProbs, void *x, void *y);
    double (*getMatchProbFcn)(const double *emissionMatchProbs, void *x, void *y);
} StateMachine3;

typedef struct _StateMachine3vanilla {                                                                      /* :217-231 */
    StateMachine model;
    double TRANSITION_M_TO_Y_NOT_X;
    double TRANSITION_E_TO_E;
    double DEFAULT_END_MATCH_PROB;
    double DEFAULT_END_FROM_X_PROB;
    double DEFAULT_END_FROM_Y_PROB;
    double (*getKmerSkipProb)(StateMachine *sM, void *kmerList, bool getAlpha);
    double (*getScaledMatchProbFcn)(const double *scaledEventModel, void *kmer, void *event);
    double (*getMatchProbFcn)(const double *eventModel, void *kmer, void *event);
} StateMachine3Vanilla;

/* inc/nanopore_hdp.h:16-22.  Here `hdp` points at this library's own tables (what dir_proc_density reads: the sampling
 * grid, and per observed Dirichlet process the posterior predictive density and its spline slopes), not at a
 * HierarchicalDirichletProcess: Gibbs sampling and the construction of HDPs are out of scope. */
typedef struct _nanoporeHDP {
    void *hdp;
    char *alphabet;
    int64_t alphabet_size;
    int64_t kmer_length;
    void *distr_metric_memos;              /* unused */
} NanoporeHDP;

typedef struct _StateMachine3_HDP {                                                                         /* :197-215 */
    StateMachine model;
    double TRANSITION_MATCH_CONTINUE;
    double TRANSITION_MATCH_FROM_GAP_X;
    double TRANSITION_MATCH_FROM_GAP_Y;
    double TRANSITION_GAP_OPEN_X;
    double TRANSITION_GAP_OPEN_Y;
    double TRANSITION_GAP_EXTEND_X;
    double TRANSITION_GAP_EXTEND_Y;
    double TRANSITION_GAP_SWITCH_TO_X;
    double TRANSITION_GAP_SWITCH_TO_Y;
    double (*getXGapProbFcn)(const double *emissionXGapProbs, void *i);
    NanoporeHDP *hdpModel;
    double (*getYGapProbFcn)(NanoporeHDP *hdp, void *x, void *y);
    double (*getMatchProbFcn)(NanoporeHDP *hdp, void *x, void *y);
} StateMachine3_HDP;

typedef struct _StateMachine4 {                                                                             /* :132-170 */
    StateMachine model;
    double TRANSITION_MATCH_CONTINUE;
    double TRANSITION_MATCH_FROM_SHORT_GAP_X;
    double TRANSITION_MATCH_FROM_LONG_GAP_X;
    double TRANSITION_MATCH_FROM_SHORT_GAP_Y;
    double TRANSITION_GAP_SHORT_OPEN_X;
    double TRANSITION_GAP_SHORT_EXTEND_X;
    double TRANSITION_GAP_SHORT_OPEN_Y;
    double TRANSITION_GAP_SHORT_EXTEND_Y;
    double TRANSITION_GAP_LONG_OPEN_X;
    double TRANSITION_GAP_LONG_EXTEND_X;
    double TRANSITION_GAP_LONG_SWITCH_TO_X;
    double (*getXGapProbFcn)(const double *emissionXGapProbs, void *kmer);
    double (*getYGapProbFcn)(const double *scaledMatchModel, void *kmer, void *event);
    double (*getMatchProbFcn)(const double *matchModel, void *kmer, void *event);
} StateMachine4;

typedef struct _StateMachineEchelon {                                                                       /* :233-245 */
    StateMachine model;
    double BACKGROUND_EVENT_PROB;
    double DEFAULT_END_MATCH_PROB;
    double DEFAULT_END_FROM_X_PROB;
    double (*getKmerSkipProb)(StateMachine *sM, void *kmerList, bool getAlpha);
    double (*getDurationProb)(void *event, int64_t n);
    double (*getMatchProbFcn)(const double *eventModel, void *kmers, void *event, int64_t n);
    double (*getScaledMatchProbFcn)(const double *scaledEventModel, void *kmer, void *event);
} StateMachineEchelon;

StateMachine *getStrawManStateMachine3(const char *modelFile);                                              /* :372 */
StateMachine *getSignalStateMachine3Vanilla(const char *modelFile);                                         /* :380 */
StateMachine *getStateMachine4(const char *modelFile);                                                      /* :378 */
StateMachine *getStateMachineEchelon(const char *modelFile);                                                /* :382 */
/* threeStateHdp (impl/stateMachine.c:1738-1749): the three-state topology, match and gap-Y emissions =
 * get_nanopore_kmer_density of the HDP (the density itself, as the reference has it), gap-X emission log(0.1); the
 * reference sequence must use sequence_getKmer3 (vanillaAlign.c:246-250).  Its expectations go into the container of
 * hmmContinuous_getEmptyHmm(threeStateHdp, pseudocount, threshold): transition sums and the event-to-k-mer assignment
 * lists (impl/continuousHmm.c:630-749), written by hmmContinuous_writeToFile in the reference's format. */
StateMachine *getHdpStateMachine3(NanoporeHDP *hdp);                                                        /* :376 */
/* inc/nanopore_hdp.h: reading a serialised NanoporeHDP (impl/nanopore_hdp.c:845-873, impl/hdp.c:3009-3270) and querying it
 * (impl/nanopore_hdp.c:390-392 -> impl/hdp.c:2577-2599 -> impl/hdp_math_utils.c:471-495) */
NanoporeHDP *deserialize_nhdp(const char *filepath);
void destroy_nanopore_hdp(NanoporeHDP *nhdp);
double get_nanopore_kmer_density(NanoporeHDP *nhdp, void *kmer, void *x);   /* one scalar table query on the host, for callers and tests; the DP reads the same tables on the device and never calls it */
int64_t get_nanopore_hdp_kmer_length(NanoporeHDP *nhdp);
int64_t get_nanopore_hdp_alphabet_size(NanoporeHDP *nhdp);
char *get_nanopore_hdp_alphabet(NanoporeHDP *nhdp);
void emissions_signal_scaleModel(StateMachine *sM, double scale, double shift, double var, double scale_sd,
                                 double var_sd);                                                            /* :341-342 */
void emissions_signal_scaleModelNoiseOnly(StateMachine *sM, double scale, double shift, double var, double scale_sd,
                                          double var_sd);                                                   /* impl/stateMachine.c:653-673 */
void stateMachine3_setTransitionsToNanoporeDefaults(StateMachine *sM);
void stateMachine3Vanilla_setStrandTransitionsToDefaults(StateMachine *sM, Strand strand);                  /* :383 */
int64_t emissions_discrete_getKmerIndex(void *kmer);                                                        /* :305 */
int64_t emissions_discrete_getKmerIndexFromKmer(void *kmer);
void stateMachine_destruct(StateMachine *stateMachine);                                                     /* :387 */

/* ---------------------------------------------------------------- inc/continuousHmm.h */
Hmm *hmmContinuous_getEmptyHmm(StateMachineType type, double pseudocount, double threshold);                /* :107 */
void hmmContinuous_normalize(Hmm *hmm, StateMachineType type);                                              /* :113 */
void hmmContinuous_writeToFile(const char *outFile, Hmm *hmm, StateMachineType type);                       /* :116 */
void hmmContinuous_loadSignalHmm(const char *hmmFile, StateMachine *sM, StateMachineType type);             /* :104 */
void hmmContinuous_destruct(Hmm *hmm, StateMachineType type);                                               /* :119 */
void vanillaHmm_implantMatchModelsintoHmm(StateMachine *sM, Hmm *hmm);                                      /* :87 */
int64_t hmmContinuous_howManyAssignments(Hmm *hmm);                                                         /* :128: threeStateHdp containers only */
/* the per-field accessors (:39-100) over the containers hmmContinuous_getEmptyHmm returns */
void continuousPairHmm_addToTransitionsExpectation(Hmm *hmm, int64_t from, int64_t to, double p);
void continuousPairHmm_setTransitionExpectation(Hmm *hmm, int64_t from, int64_t to, double p);
double continuousPairHmm_getTransitionExpectation(Hmm *hmm, int64_t from, int64_t to);
void continuousPairHmm_addToKmerGapExpectation(Hmm *hmm, int64_t state, int64_t kmerIndex, int64_t ignore, double p);
void continuousPairHmm_setKmerGapExpectation(Hmm *hmm, int64_t state, int64_t kmerIndex, int64_t ignore, double p);
double continuousPairHmm_getKmerGapExpectation(Hmm *hmm, int64_t state, int64_t kmerIndex, int64_t ignore);
void continuousPairHmm_loadTransitionsAndKmerGapProbs(StateMachine *sM, Hmm *hmm);
void continuousPairHmm_normalize(Hmm *hmm);
void continuousPairHmm_destruct(Hmm *hmm);
void continuousPairHmm_writeToFile(Hmm *hmm, FILE *fileHandle);
Hmm *continuousPairHmm_loadFromFile(const char *fileName);
void vanillaHmm_addToKmerSkipBinExpectation(Hmm *hmm, int64_t bin, int64_t ignore, double p);
void vanillaHmm_setKmerSkipBinExpectation(Hmm *hmm, int64_t bin, int64_t ignore, double p);
double vanillaHmm_getKmerSkipBinExpectation(Hmm *hmm, int64_t bin, int64_t ignore);
void vanillaHmm_normalizeKmerSkipBins(Hmm *hmm);
void vanillaHmm_loadKmerSkipBinExpectations(StateMachine *sM, Hmm *hmm);
void vanillaHmm_destruct(Hmm *hmm);

/* ---------------------------------------------------------------- inc/nanopore.h */
#define NB_EVENT_PARAMS 3
typedef struct _nanoporeReadAdjustmentParameters { double scale, shift, var, scale_sd, var_sd; } NanoporeReadAdjustmentParameters;
typedef struct _nanoporeRead {                                                                              /* :14-29 */
    int64_t readLength;
    int64_t nbTemplateEvents;
    int64_t nbComplementEvents;
    NanoporeReadAdjustmentParameters templateParams;
    NanoporeReadAdjustmentParameters complementParams;
    char *twoDread;
    int64_t *templateEventMap;
    double *templateEvents;
    int64_t *complementEventMap;
    double *complementEvents;
    bool scaled;
} NanoporeRead;
NanoporeRead *nanopore_loadNanoporeReadFromFile(const char *nanoporeReadFile);                              /* :31 */
stList *nanopore_remapAnchorPairs(stList *anchorPairs, int64_t *eventMap);                                  /* :33 */
stList *nanopore_remapAnchorPairsWithOffset(stList *unmappedPairs, int64_t *eventMap, int64_t mapOffset);   /* :35 */
void nanopore_descaleNanoporeRead(NanoporeRead *npRead);                                                    /* :37 */
void nanopore_nanoporeReadDestruct(NanoporeRead *npRead);                                                   /* :39 */

/* ---------------------------------------------------------------- additive: batching and device selection */
/* Many (read, reference) pairs in one call: what N calls of getAlignedPairsUsingAnchors return, computed as one GPU
 * batch.  All arrays have n entries; result[i] is a new stList owned by the caller. */
void getAlignedPairsUsingAnchorsBatch(int64_t n, StateMachine **sMs, Sequence **sXs, Sequence **sYs, stList **anchorPairs,
                                      PairwiseAlignmentParameters *p, bool raggedLeft, bool raggedRight, stList **results);
/* Device for this process (default: $CPECAN_DEVICE or 0); must be called before the first alignment. */
void cpecan_host_set_device(int device);
/* Posteriors of the threeState / vanilla machines in FP64 and the reference's own operation order (scores equal to the
 * last digit; slower) instead of FP32 (default: $CPECAN_EXACT or 0); see cpecan_cuda_set_exact_arithmetic. */
void cpecan_host_set_exact_arithmetic(int on);

#ifdef __cplusplus
}
#endif
#endif
