/*
 * cpecan_cuda.h -- C-ABI of the B200 (sm_100a) banded signal pair-HMM engine.
 *
 * This is the additive device boundary SURVEY.md 8(b) calls for: plain pointers and sizes, integer status
 * codes, no torch / C++ types.  One call evaluates a BATCH of independent read-vs-reference alignments; each
 * batch entry ("work item") is what ONE call of the reference's
 *     getPosteriorProbsWithBanding           (impl/pairwiseAligner.c:870-1006)
 * computes, i.e. one banded region after getSplitPoints (impl/pairwiseAligner.c:1313-1340).  The host mirror of
 * the reference's public entry points (getAlignedPairsUsingAnchors, getExpectationsUsingAnchors,
 * getAlignedPairsWithoutBanding; include/cpecan_host.h) is a thin layer over these calls.
 *
 * There is no CPU fallback: every entry point returns CPECAN_ERR_CUDA when no usable device exists.
 */
#ifndef CPECAN_CUDA_H_
#define CPECAN_CUDA_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    CPECAN_OK = 0,
    CPECAN_ERR_CUDA = 1,        /* CUDA runtime error; see cpecan_cuda_last_error() */
    CPECAN_ERR_ARG = 2,         /* invalid argument */
    CPECAN_ERR_BAND_TOO_WIDE = 3,/* a band diagonal exceeds the widest kernel instantiation */
    CPECAN_ERR_NOMEM = 4
};

/* State-machine types, numbered as the reference's StateMachineType (inc/stateMachine.h:20-29). */
enum { CPECAN_SM_THREE_STATE = 2, CPECAN_SM_VANILLA = 4, CPECAN_SM_ECHELON = 5, CPECAN_SM_FOUR_STATE = 6, CPECAN_SM_THREE_STATE_HDP = 7 };

/* Work modes. */
enum {
    CPECAN_MODE_POSTERIOR = 0,  /* diagonalCalculationPosteriorMatchProbs   (impl/pairwiseAligner.c:756-795) */
    CPECAN_MODE_EXPECTATION = 1,/* diagonalCalculation_Expectations         (impl/pairwiseAligner.c:841-863) */
    CPECAN_MODE_UNBANDED = 2    /* getAlignedPairsWithoutBanding schedule   (impl/pairwiseAligner.c:1512-1569) */
};

/* per-item status bits returned in cpecan_result.status */
enum {
    CPECAN_ITEM_OK = 0,
    CPECAN_ITEM_PAIR_OVERFLOW = 1, /* more aligned pairs than pair_cap; n_pairs holds the true count */
    CPECAN_ITEM_NONFINITE = 2,     /* total probability was -inf / NaN (read skipped, reference would emit NaNs) */
    CPECAN_ITEM_BAND_STEP = 4,     /* band edge moved backwards or by more than one cell between diagonals (never for anchors that
                                      went through filterToRemoveOverlap): the item is not aligned */
    CPECAN_ITEM_BAD_KMER = 8       /* expectation mode: the reference holds a non-ACGT k-mer, every path is -inf; the read is
                                      dropped from the sums as the reference drops it (status also carries NONFINITE).
                                      threeStateHdp, any mode: a k-mer the HDP's alphabet cannot spell was met (its density
                                      read as 0); the reference ends the program there (impl/nanopore_hdp.c:358-373) */
};

typedef struct cpecan_ctx cpecan_ctx;

/* PairwiseAlignmentParameters (inc/pairwiseAligner.h:80-91; defaults impl/pairwiseAligner.c:1428-1441). */
typedef struct {
    double threshold;
    int64_t minDiagsBetweenTraceBack;
    int64_t traceBackDiagonals;
    int64_t diagonalExpansion;
    int64_t constraintDiagonalTrim;
    int64_t splitMatrixBiggerThanThis;
} cpecan_params;

/* What a StateMachine3 / StateMachine3Vanilla carries besides its tables (inc/stateMachine.h:176-231). */
typedef struct {
    int32_t sm_type;        /* CPECAN_SM_* */
    int32_t reserved;
    double transitions[9];  /* threeState, StateMachine3 field order: MATCH_CONTINUE, MATCH_FROM_GAP_X,
                               MATCH_FROM_GAP_Y, GAP_OPEN_X, GAP_OPEN_Y, GAP_EXTEND_X, GAP_EXTEND_Y,
                               GAP_SWITCH_TO_X, GAP_SWITCH_TO_Y (log space, -inf allowed) */
    double vanilla[5];      /* vanilla: TRANSITION_M_TO_Y_NOT_X, TRANSITION_E_TO_E, DEFAULT_END_MATCH_PROB,
                               DEFAULT_END_FROM_X_PROB, DEFAULT_END_FROM_Y_PROB */
    double four_state[11];  /* fourState (StateMachine4, inc/stateMachine.h:130-152; defaults impl/stateMachine.c:993-1011):
                               MATCH_CONTINUE, MATCH_FROM_SHORT_GAP_X, MATCH_FROM_SHORT_GAP_Y, MATCH_FROM_LONG_GAP_X,
                               GAP_SHORT_OPEN_X, GAP_SHORT_EXTEND_X, GAP_SHORT_OPEN_Y, GAP_SHORT_EXTEND_Y, GAP_LONG_OPEN_X,
                               GAP_LONG_EXTEND_X, GAP_LONG_SWITCH_TO_X (log space).  echelon carries no parameters of its
                               own: its skip bins are the model's 60 gap-X entries, its end vector the reference's constants. */
} cpecan_hmm;

/* A batch of work items in flat (structure-of-arrays) host buffers.  All *_off arrays have n_items+1 entries.
 *   ref        nucleotides of every item back to back; item i owns ref[ref_off[i] .. ref_off[i+1]) and has
 *              lX = max(len - 5, 0) k-mers (sequence_correctSeqLength, impl/pairwiseAligner.c:339-354)
 *   events     reference layout: 3 doubles (mean, noise, duration) per event (inc/nanopore.h NB_EVENT_PARAMS);
 *              item i owns events[3*ev_off[i] .. 3*ev_off[i+1])
 *   anchors    (x, y) int64 pairs in sequence coordinates, strictly increasing in both (filterToRemoveOverlap);
 *              item i owns pairs anchor_off[i] .. anchor_off[i+1]
 *   model_id   per item, from cpecan_cuda_upload_model
 *   scale      per item 5 doubles (scale, shift, var, scale_sd, var_sd) applied to the MATCH table exactly as
 *              emissions_signal_scaleModel does (impl/stateMachine.c:631-651), or NULL when the uploaded tables
 *              are already scaled
 *   ragged     per item bit0 = alignmentHasRaggedLeftEnd, bit1 = alignmentHasRaggedRightEnd
 */
typedef struct {
    int64_t n_items;
    const char *ref;        const int64_t *ref_off;
    const double *events;   const int64_t *ev_off;
    const int64_t *anchors; const int64_t *anchor_off;
    const int32_t *model_id;
    const double *scale;
    const uint8_t *ragged;
} cpecan_batch;

/* Per-item results (host buffers owned by the caller). */
typedef struct {
    int64_t n_pairs;        /* aligned pairs found (may exceed the capacity given; see status) */
    int64_t pair_off;       /* offset (in pairs) of this item's pairs inside the pairs buffer */
    int64_t band_cells;     /* C = sum over diagonals of the band width (SURVEY.md 8(d) work unit) */
    double total_logprob;   /* totalProbability handed to the LAST diagonal (xay = lX+lY) of the item */
    int32_t status;         /* CPECAN_ITEM_* bits */
    int32_t n_tracebacks;
} cpecan_result;

/* Timing of the last batch call, measured with CUDA events on the engine's own streams. */
typedef struct {
    double h2d_ms, prep_ms, plan_ms, align_ms, d2h_ms, total_ms;
    int64_t kernel_launches;    /* launches of this library's kernels during the call */
    int64_t h2d_bytes, d2h_bytes;
    int64_t band_cells;         /* sum over items */
    int32_t warps_per_item;     /* kernel instantiation used for the widest bucket */
    int32_t ctas;
} cpecan_timing;

int cpecan_cuda_init(int device, cpecan_ctx **ctx_out);
void cpecan_cuda_destroy(cpecan_ctx *ctx);
const char *cpecan_cuda_last_error(cpecan_ctx *ctx);

/* Upload one pore model: the three lines of a .model file as emissions_signal_loadPoreModel leaves them
 * (impl/stateMachine.c:242-320).  match/gapy: 1 + 4096*5 doubles; gapx: n_gapx doubles (threeState: 4096 log
 * probabilities = EMISSION_GAP_X_PROBS; vanilla: 60 skip-bin probabilities). */
int cpecan_cuda_upload_model(cpecan_ctx *ctx, const double *match, const double *gapy, const double *gapx,
                             int32_t n_gapx, int32_t *model_id_out);

/* Upload one NanoporeHDP for the threeStateHdp machine (getHdpStateMachine3, impl/stateMachine.c:1738-1749): what
 * get_nanopore_kmer_density (impl/nanopore_hdp.c:390-392) reads.  The sampling grid is linspace(grid_start, grid_stop,
 * grid_length) (impl/hdp_math_utils.c:497-510); `density` and `slopes` hold n_distr rows of grid_length doubles (the
 * posterior predictive of an observed Dirichlet process and its spline slopes, impl/hdp.c:2540-2575); kmer_distr[k]
 * is the row the k-th ACGT 6-mer reads (its own process, or the nearest observed ancestor: impl/hdp.c:2588-2591), or
 * -1 when the HDP's alphabet cannot spell it.  The id returned lives in the same space as pore-model ids and is
 * released the same way; items of a threeStateHdp batch must carry such an id, items of the other machines must not. */
int cpecan_cuda_upload_hdp(cpecan_ctx *ctx, double grid_start, double grid_stop, int64_t grid_length, int32_t n_distr,
                           const double *density, const double *slopes, const int32_t *kmer_distr, int32_t *model_id_out);

/* Overwrite tables of an uploaded model in place (NULL = keep): what the M-step loaders do to a live StateMachine
 * (continuousPairHmm_loadTransitionsAndKmerGapProbs rewrites EMISSION_GAP_X_PROBS, impl/continuousHmm.c:206-232). */
int cpecan_cuda_update_model(cpecan_ctx *ctx, int32_t model_id, const double *match, const double *gapy,
                             const double *gapx);
/* Free the device tables of an uploaded model; its id may be handed out again by a later upload.  What
 * stateMachine_destruct (impl/stateMachine.c:1786-1788) is to the host tables. */
int cpecan_cuda_release_model(cpecan_ctx *ctx, int32_t model_id);

/* Posterior match probabilities for a batch.
 *   pairs_out   int32 triples (score, x, y), score = floor(p * 1e7) (PAIR_ALIGNMENT_PROB_1), x / y sequence
 *               coordinates local to the item; pair_cap_total = capacity of pairs_out in triples.  Item i gets a
 *               slice proportional to its event count; results[i].pair_off/n_pairs locate it.  Pairs appear in
 *               the order the reference's traceback emits them (descending diagonals inside a traceback,
 *               ascending x inside a diagonal); getAlignedPairsUsingAnchors then reverses each region
 *               (alignedPairCoordinateCorrectionFn pops, impl/pairwiseAligner.c:1447-1454) -- the host mirror does that.
 *   totals_out  optional debug output, may be NULL: item i owns 3 * (lX + lY + 1) doubles starting at tot_off[i]
 *               (or packed back to back when tot_off is NULL): [xay] the totalProbability handed to diagonal xay
 *               (NaN where no posterior was computed), [(lX+lY+1) + xay] and [2*(lX+lY+1) + xay] the two terms of
 *               diagonalCalculationTotalProbability on the diagonals where it was recomputed.
 */
int cpecan_cuda_align_batch(cpecan_ctx *ctx, const cpecan_hmm *hmm, const cpecan_params *params, int32_t mode,
                            const cpecan_batch *batch, int32_t *pairs_out, int64_t pair_cap_total,
                            cpecan_result *results, double *totals_out, const int64_t *tot_off);

/* Baum-Welch expectations for a batch (getExpectationsUsingAnchors, impl/pairwiseAligner.c:1571-1591), summed over
 * the batch on device.  expectations_out: threeState 9 + 4096 + 1 doubles (transitions row-major from*3+to, k-mer
 * skip counts, likelihood); vanilla 60 + 1.  Values are ADDED to what the buffer already holds (pseudocounts). */
#define CPECAN_N_EXPECT (9 + 4096 + 1)   /* threeState: transitions, k-mer skip counts, likelihood */
#define CPECAN_N_EXPECT_VANILLA (60 + 1)  /* vanilla: 30 beta + 30 alpha skip-bin counts, likelihood */
int cpecan_cuda_expectations_batch(cpecan_ctx *ctx, const cpecan_hmm *hmm, const cpecan_params *params,
                                   const cpecan_batch *batch, double *expectations_out, cpecan_result *results);
/* threeStateHdp: what getExpectationsUsingAnchors leaves in an HdpHmm (cell_signal_updateTransAndKmerSkipExpectations2,
 * impl/pairwiseAligner.c:445-476).  expectations_out as above (9 transition sums, the k-mer part stays untouched, the
 * likelihood); assignments_out: int32 triples (from state, k-mer position, event index), one for every transition into
 * the match state whose posterior is >= params->threshold, in the order the reference appends them to its lists
 * (traceback order of the diagonals, ascending x, from match / gap X / gap Y); sliced per item like pairs_out
 * (results[i].pair_off / n_pairs; CPECAN_ITEM_PAIR_OVERFLOW when an item's slice was too small). */
int cpecan_cuda_hdp_expectations_batch(cpecan_ctx *ctx, const cpecan_hmm *hmm, const cpecan_params *params,
                                       const cpecan_batch *batch, double *expectations_out, int32_t *assignments_out,
                                       int64_t assignment_cap_total, cpecan_result *results);
/* The same in steps, for the EM driver: stage(mode = CPECAN_MODE_EXPECTATION) + run_staged() leave the batch sums in a
 * device buffer of CPECAN_N_EXPECT doubles; expectations_device_ptr() exposes it so that the caller can all-reduce it
 * in place across GPUs (NCCL over NVLink) before fetch_expectations() ADDS it to a host vector. */
int cpecan_cuda_fetch_expectations(cpecan_ctx *ctx, double *expectations_out);
int cpecan_cuda_expectations_device_ptr(cpecan_ctx *ctx, double **dev_ptr_out);
/* Multi-GPU training without Python: one context per GPU (one process or host thread each),
 * joined in ONE NCCL communicator.  Rank 0 calls nccl_unique_id() and hands the 128 bytes to the others by whatever
 * means the driver has (a file, a pipe, MPI); every rank then calls nccl_init().  allreduce_expectations() sums the
 * device accumulators of all ranks in place -- ncclAllReduce(ncclSum, ncclDouble) over NVLink / NVSwitch, 4106 doubles
 * (vanilla: 61), on the context's stream -- so that the following fetch_expectations() returns the whole job's sums on
 * every rank: what scripts/trainModels.py:244-330 gets by adding up one expectation file per read.  libnccl.so.2 is
 * opened at run time (dlopen); CPECAN_ERR_CUDA with a message if it is missing. */
#define CPECAN_NCCL_UNIQUE_ID_BYTES 128
int cpecan_cuda_nccl_unique_id(void *id_out);
int cpecan_cuda_nccl_init(cpecan_ctx *ctx, int32_t n_ranks, int32_t rank, const void *id);
int cpecan_cuda_allreduce_expectations(cpecan_ctx *ctx);

/* Device-resident variant used to measure kernel-only throughput: stage() copies and prepares a batch in HBM once,
 * run_staged() re-runs plan + align kernels on it (results stay on device), fetch_staged() copies results back. */
int cpecan_cuda_stage(cpecan_ctx *ctx, const cpecan_hmm *hmm, const cpecan_params *params, int32_t mode,
                      const cpecan_batch *batch, int64_t pair_cap_total);
int cpecan_cuda_run_staged(cpecan_ctx *ctx);
/* New transitions / updated model tables for the batch that is already staged (the M-step between two E-steps of a
 * resident batch): re-derives the per-column parameter records on device, no host<->device copy of the reads. */
int cpecan_cuda_restage_model(cpecan_ctx *ctx, const cpecan_hmm *hmm);
/* The same split in two, so that several contexts (e.g. one per band expansion) keep the GPU full together:
 * run_staged_async() only enqueues, wait() blocks until the context's kernels are done and fills the timing. */
int cpecan_cuda_run_staged_async(cpecan_ctx *ctx);
int cpecan_cuda_wait(cpecan_ctx *ctx);
int cpecan_cuda_fetch_staged(cpecan_ctx *ctx, int32_t *pairs_out, cpecan_result *results);

/* Page-locked host buffers for callers that want asynchronous, full-rate host<->device copies. */
void *cpecan_cuda_host_alloc(cpecan_ctx *ctx, int64_t bytes);
void cpecan_cuda_host_free(cpecan_ctx *ctx, void *p);

/* Share of the device one context takes: at most `warps_per_sm` resident alignment warps per SM for the batches staged
 * afterwards (0 = as many as fit, the default).  Callers that stream sub-batches through several contexts at once give
 * each a share, so that the forward-row rings (sized by the resident warps) of all of them fit in HBM. */
int cpecan_cuda_set_resident_warps(cpecan_ctx *ctx, int32_t warps_per_sm);

/* Arithmetic of the threeState / vanilla posterior batches staged afterwards.  0 (default): the FP32 kernel (k_align3),
 * whose posteriors agree with the reference's FP64 ones to ~1e-4 apart from rare logAdd segment flips.
 * 1: the FP64 kernel in the reference's own operation order -- the same pair lists, scores equal to the last digit of
 * floor(p * 1e7) -- at 8 - 10 % of the FP32 rate.  Batches with an odd diagonalExpansion always run that
 * way (their band is one the FP32 kernel does not walk).  Expectation batches follow the same switch (sums to 1e-9 of
 * the reference's instead of 2e-4). */
int cpecan_cuda_set_exact_arithmetic(cpecan_ctx *ctx, int32_t on);

int cpecan_cuda_get_timing(cpecan_ctx *ctx, cpecan_timing *out);
int cpecan_cuda_device_info(cpecan_ctx *ctx, int32_t *sm_count, int32_t *clock_khz, int64_t *hbm_bytes);

#ifdef __cplusplus
}
#endif
#endif
