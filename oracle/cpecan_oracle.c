/*
 * cpecan_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, double-precision CPU restatement of the reference's banded pair-HMM
 * forward / backward / posterior-match-probability path, written for this repo
 * (not copied) so that it can travel to machines where /root/reference does not
 * exist.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it; the product library never does.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py checks this file
 * bit-for-bit (identical (score,x,y) lists, identical per-diagonal totals)
 * against the unmodified reference sources compiled into oracle/_ref/ on the
 * reference's own fixture (987 / 986 / 999 / 953 aligned pairs,
 * tests/signalPairwiseTest.c:1163,1173,1293,1303; fourState 988 / 988, :1227,1236;
 * echelon 857 / 1000, :1434,1445) and on seeded synthetic reads; the resulting
 * golden vectors are committed under tests/golden/.  Machines: threeState
 * (strawMan), vanilla, fourState, echelon.
 *
 * Each function cites the reference file:line whose behaviour it restates.
 * "ref" below = impl/pairwiseAligner.c, "sm" = impl/stateMachine.c.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define NEG_INF (-INFINITY)
#define N_KMERS 4096
#define KMER_LEN 6
#define MODEL_STRIDE 5
#define SM_THREE_STATE 2
#define SM_VANILLA 4
#define SM_THREE_STATE_HDP 7
#define SM_FOUR_STATE 6       /* inc/stateMachine.h:20-29 */
#define SM_ECHELON 5
#define ECH_GAPX 6             /* match0 .. match5 = 0 .. 5, gapX = 6 (sm:1164-1166) */
enum { ST_M = 0, ST_X = 1, ST_Y = 2, ST_LX = 3 };       /* match, shortGapX, shortGapY, longGapX (inc/stateMachine.h:31-33) */
/* StateMachine4 transitions in the order this file keeps them (inc/stateMachine.h StateMachine4) */
enum { T4_MATCH_CONTINUE = 0, T4_MATCH_FROM_SHORT_GAP_X, T4_MATCH_FROM_SHORT_GAP_Y, T4_MATCH_FROM_LONG_GAP_X,
       T4_GAP_SHORT_OPEN_X, T4_GAP_SHORT_EXTEND_X, T4_GAP_SHORT_OPEN_Y, T4_GAP_SHORT_EXTEND_Y,
       T4_GAP_LONG_OPEN_X, T4_GAP_LONG_EXTEND_X, T4_GAP_LONG_SWITCH_TO_X, T4_COUNT };

typedef struct {
    int32_t sm_type;          /* 2 = threeState (strawMan), 4 = vanilla, 6 = fourState, 5 = echelon */
    int32_t strand;           /* unused by the DP; kept for the host mirror */
    const double *match;      /* EMISSION_MATCH_PROBS: 1 + 4096*5 (already scaled per read) */
    const double *gapy;       /* EMISSION_GAP_Y_PROBS: 1 + 4096*5 (never scaled, sm:631-651) */
    const double *gapx;       /* threeState: 4096 log-probs; vanilla, echelon: 60 skip-bin probabilities */
    double trans[9];          /* threeState, StateMachine3 field order (inc/stateMachine.h:179-187) */
    double vanilla[5];        /* M_TO_Y_NOT_X, E_TO_E, END_MATCH, END_FROM_X, END_FROM_Y */
    double trans4[T4_COUNT];  /* fourState, T4_* order; gapx then holds 4096 zeros (emissions_signal_initEmissionsToZero) */
    /* threeStateHdp (sm_type 7): what get_nanopore_kmer_density reads of a NanoporeHDP -- the sampling grid
     * linspace(start, stop, len), per distinct observed Dirichlet process a row of posterior predictive densities and
     * one of spline slopes, and for every ACGT 6-mer the row it reads (its nearest observed ancestor's) or -1 */
    const double *hdp_density, *hdp_slopes;
    const int32_t *hdp_kmer_row;
    double hdp_start, hdp_stop;
    int64_t hdp_len;
} OracleModel;

typedef struct {
    double threshold;
    int64_t minDiagsBetweenTraceBack;
    int64_t traceBackDiagonals;
    int64_t diagonalExpansion;
    int64_t constraintDiagonalTrim;
    int64_t splitMatrixBiggerThanThis;
} OracleParams;

/* ------------------------------------------------------------------ logAdd
 * ref:235-255.  Four cubic segments with float-literal coefficients, hard cut at 7.5. */
static double la_poly(double x) {
    if (x <= 1.00f)
        return ((-0.009350833524763f * x + 0.130659527668286f) * x + 0.498799810682272f) * x + 0.693203116424741f;
    if (x <= 2.50f)
        return ((-0.014532321752540f * x + 0.139942324101744f) * x + 0.495635523139337f) * x + 0.692140569840976f;
    if (x <= 4.50f)
        return ((-0.004605031767994f * x + 0.063427417320019f) * x + 0.695956496475118f) * x + 0.514272634594009f;
    return ((-0.000458661602210f * x + 0.009695946122598f) * x + 0.930734667215156f) * x + 0.168037164329057f;
}

double oracle_log_add(double x, double y) {
    if (x < y) return (x == NEG_INF || y - x >= 7.5) ? y : la_poly(y - x) + x;
    return (y == NEG_INF || x - y >= 7.5) ? x : la_poly(x - y) + y;
}
#define LA oracle_log_add

/* ------------------------------------------------------------- k-mer index
 * sm:104-139: A0 C1 G2 T3, anything else contributes NUM_OF_KMERS+1, so any
 * non-ACGT base pushes the index above 4096.  Returned as -1 here. */
static int base_index(char b) {
    switch (b) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; default: return -1; }
}
static int32_t kmer_index(const char *s) {
    int32_t v = 0;
    for (int j = 0; j < KMER_LEN; j++) {
        int b = base_index(s[j]);
        if (b < 0) return -1;
        v = v * 4 + b;
    }
    return v;
}
int32_t oracle_kmer_index(const char *s) { return kmer_index(s); }

/* model getters, sm:221-240: index > 4096 reads as 0.0 */
static double mdl(const double *tbl, int32_t k, int col) { return k < 0 ? 0.0 : tbl[1 + k * MODEL_STRIDE + col]; }

/* sm:333-343 */
static double log_gauss(double x, double mu, double sigma) {
    if (sigma == 0.0) return NEG_INF;
    double a = (x - mu) / sigma;
    return -0.91893853320467267 - log(sigma) + (-0.5 * a * a);
}
/* sm:322-331 */
static double log_inv_gauss(double x, double mu, double lambda) {
    double a = (x - mu) / mu;
    return (log(lambda) - 1.8378770664093453 - 3 * log(x) - lambda * a * a / x) / 2;
}
/* sm:595-629 (strawMan: two Gaussians) */
static double emit_two_gauss(const double *tbl, int32_t k, const double *ev) {
    return log_gauss(ev[0], mdl(tbl, k, 0), mdl(tbl, k, 1)) + log_gauss(ev[1], mdl(tbl, k, 2), mdl(tbl, k, 3));
}
/* sm:499-528 (vanilla: Gaussian level x inverse-Gaussian noise) */
static double emit_gauss_invgauss(const double *tbl, int32_t k, const double *ev) {
    return log_gauss(ev[0], mdl(tbl, k, 0), mdl(tbl, k, 1)) + log_inv_gauss(ev[1], mdl(tbl, k, 2), mdl(tbl, k, 4));
}

/* emissions_signal_scaleModel, sm:631-651: match table only */
void oracle_scale_model(double *match, double scale, double shift, double var, double scale_sd, double var_sd) {
    for (int64_t i = 1; i < N_KMERS * MODEL_STRIDE + 1; i += MODEL_STRIDE) {
        match[i] = match[i] * scale + shift;
        match[i + 1] = match[i + 1] * var;
        match[i + 2] = match[i + 2] * scale_sd;
        match[i + 4] = match[i + 4] * var_sd;
        match[i + 3] = sqrt(pow(match[i + 2], 3.0) / match[i + 4]);
    }
}

/* ------------------------------------------------------------------- band
 * ref:98-184.  anchors are sequence coordinates; matrix coordinates are +1. */
static int64_t fix_parity(int64_t xay, int64_t xmy) { return (xay + xmy) % 2 == 0 ? xmy : xmy + 1; }
static int64_t clampz(int64_t z, int64_t l) { return z < 0 ? 0 : (z > l ? l : z); }

void oracle_band(const int64_t *anchors, int64_t nA, int64_t lX, int64_t lY, int64_t e, int64_t *xmyL, int64_t *xmyR) {
    int64_t ai = 0, xay = 0, pxay = 0, pxmy = 0, nxay = 0, nxmy = 0, xL = 0, yL = 0, xU = 0, yU = 0;
    while (xay <= lX + lY) {
        int64_t l = fix_parity(xay, xL - yL), r = fix_parity(xay, xU - yU);
        int64_t t;
        t = (xay + l) / 2; if (t < xL) l += 2 * (xL - t);
        t = (xay - l) / 2; if (yL < t) l += 2 * (t - yL);
        t = (xay + r) / 2; if (xU < t) r -= 2 * (t - xU);
        t = (xay - r) / 2; if (t < yU) r -= 2 * (yU - t);
        xmyL[xay] = l; xmyR[xay] = r;
        if (nxay == xay++) {
            pxay = nxay; pxmy = nxmy;
            int64_t x = lX, y = lY;
            if (ai < nA) { x = anchors[2 * ai] + 1; y = anchors[2 * ai + 1] + 1; ai++; }
            nxay = x + y; nxmy = x - y;
            xL = clampz((pxay + (pxmy - e)) / 2, lX);
            yL = clampz((nxay - (nxmy - e)) / 2, lY);
            xU = clampz((nxay + (nxmy + e)) / 2, lX);
            yU = clampz((pxay - (pxmy + e)) / 2, lY);
        }
    }
}

/* ------------------------------------------------- filterToRemoveOverlap
 * ref:1160-1200.  Input sorted lexicographically; keeps pairs that are strictly
 * increasing in x and y both scanning backwards and forwards. */
int64_t oracle_filter_overlap(const int64_t *pairs, int64_t n, int64_t *out) {
    char *keep = calloc(n ? n : 1, 1);
    int64_t pX = INT64_MAX, pY = INT64_MAX;
    for (int64_t i = n - 1; i >= 0; i--) {
        int64_t x = pairs[2 * i], y = pairs[2 * i + 1];
        if (x < pX && y < pY) keep[i] = 1;
        pX = x < pX ? x : pX; pY = y < pY ? y : pY;
    }
    /* the reference looks membership up by VALUE (a sorted set): a duplicate of a kept pair is "in the set" too */
    pX = INT64_MIN; pY = INT64_MIN;
    int64_t m = 0;
    for (int64_t i = 0; i < n; i++) {
        int64_t x = pairs[2 * i], y = pairs[2 * i + 1];
        int inSet = keep[i];
        if (!inSet) {
            for (int64_t j = i + 1; j < n && pairs[2 * j] == x && pairs[2 * j + 1] == y; j++) if (keep[j]) { inSet = 1; break; }
            for (int64_t j = i - 1; !inSet && j >= 0 && pairs[2 * j] == x && pairs[2 * j + 1] == y; j--) if (keep[j]) { inSet = 1; break; }
        }
        if (x > pX && y > pY && inSet) { out[2 * m] = x; out[2 * m + 1] = y; m++; }
        pX = x > pX ? x : pX; pY = y > pY ? y : pY;
    }
    free(keep);
    return m;
}

/* ----------------------------------------------------------- split points
 * ref:1289-1340.  Returns 4-tuples (x1,y1,x2,y2). */
static int split_step(int64_t *x1, int64_t *y1, int64_t x2, int64_t y2, int64_t x3, int64_t y3, int64_t *out,
                      int64_t *n, int64_t maxMatrix, int skipBlock) {
    int64_t lX2 = x3 - x2, lY2 = y3 - y2;
    if (lX2 * lY2 > maxMatrix) {
        int64_t maxLen = (int64_t) sqrt((double) maxMatrix);
        int64_t hX = lX2 / 2 > maxLen ? maxLen : lX2 / 2;
        int64_t hY = lY2 / 2 > maxLen ? maxLen : lY2 / 2;
        if (!skipBlock) { out[4 * *n] = *x1; out[4 * *n + 1] = *y1; out[4 * *n + 2] = x2 + hX; out[4 * *n + 3] = y2 + hY; (*n)++; }
        *x1 = x3 - hX; *y1 = y3 - hY;
        return 1;
    }
    return 0;
}
int64_t oracle_split_points(const int64_t *anchors, int64_t nA, int64_t lX, int64_t lY, int64_t maxMatrix,
                            int raggedLeft, int raggedRight, int64_t *out) {
    int64_t x1 = 0, y1 = 0, x2 = 0, y2 = 0, n = 0;
    for (int64_t i = 0; i < nA; i++) {
        int64_t x3 = anchors[2 * i], y3 = anchors[2 * i + 1];
        split_step(&x1, &y1, x2, y2, x3, y3, out, &n, maxMatrix, raggedLeft && i == 0);
        x2 = x3 + 1; y2 = y3 + 1;
    }
    if (!split_step(&x1, &y1, x2, y2, lX, lY, out, &n, maxMatrix, raggedLeft && nA == 0) || !raggedRight) {
        out[4 * n] = x1; out[4 * n + 1] = y1; out[4 * n + 2] = lX; out[4 * n + 3] = lY; n++;
    }
    return n;
}

/* optional debug capture of the two terms of the total (set by oracle_debug_terms) */
static int g_dbg_logp = 0;   /* when set, the score slot carries round(log p * 1e9) of the UNCLAMPED posterior */
void oracle_debug_logp(int on) { g_dbg_logp = on; }
static int64_t g_dbg_row = -1; static double *g_dbg_F = NULL, *g_dbg_B = NULL;   /* capture one diagonal's cells */
void oracle_debug_row(int64_t d, double *F, double *B) { g_dbg_row = d; g_dbg_F = F; g_dbg_B = B; }
static double *g_dbg_t1 = NULL, *g_dbg_t2 = NULL;
void oracle_debug_terms(double *t1, double *t2) { g_dbg_t1 = t1; g_dbg_t2 = t2; }

/* ------------------------------------------------------------ DP plumbing */
typedef struct {
    const OracleModel *m;
    const char *ref;          /* nucleotides of this (sub-)region: ref[0 .. lX+4] */
    const double *ev;         /* events of this (sub-)region: 3 doubles each */
    int64_t lX, lY;
    const int64_t *xmyL, *xmyR;
    int S;                    /* states per cell: 3, fourState 4, echelon 7 */
    int64_t refLen;           /* characters from ref to the end of the sequence (echelon reads into the padding) */
    double **F, **B;          /* per-diagonal cell arrays (S doubles per cell) or NULL when not alive */
    /* expectation accumulators (mode 2) */
    double total;
    double *expT;             /* threeState: 9 transitions; vanilla: 60 bins */
    double *expSkip;          /* threeState: 4096 k-mer skip counts */
} Dp;

static double *cell_at(Dp *dp, double **mat, int64_t xay, int64_t xmy) {
    if (xay < 0 || xay > dp->lX + dp->lY || mat[xay] == NULL) return NULL;
    if (xmy < dp->xmyL[xay] || xmy > dp->xmyR[xay]) return NULL;
    return mat[xay] + ((xmy - dp->xmyL[xay]) / 2) * dp->S;
}
static int64_t width_of(Dp *dp, int64_t xay) { return (dp->xmyR[xay] - dp->xmyL[xay]) / 2 + 1; }
static double *new_diag(Dp *dp, double **mat, int64_t xay, double fill) {
    int64_t n = width_of(dp, xay) * dp->S;
    if (mat[xay] == NULL) mat[xay] = malloc(sizeof(double) * n);
    for (int64_t i = 0; i < n; i++) mat[xay][i] = fill;
    return mat[xay];
}
static void drop_diag(double **mat, int64_t xay) { if (mat[xay]) { free(mat[xay]); mat[xay] = NULL; } }

enum { MODE_FWD = 0, MODE_BWD = 1, MODE_EXP = 2 };

/* HDP expectations (ref:445-476): a transition into the match state with posterior >= threshold is an ASSIGNMENT of the
 * cell's event to its k-mer; recorded as (from state, k-mer position, event index) relative to the buffers handed to
 * oracle_hdp_expectations */
static int64_t *g_asg = NULL;
static int64_t g_asg_cap = 0, g_asg_n = 0;
static double g_asg_thr = 0.0;
static const char *g_asg_ref = NULL, *g_cell_kmer = NULL;
static const double *g_asg_ev = NULL, *g_cell_ev = NULL;

/* ref:365-383 (forward/backward transition), ref:426-443 and ref:478-498 (expectation updates) */
static void transition(Dp *dp, int mode, double *nb, double *cur, int from, int to, double eP, double tP, int32_t kx, int bin) {
    if (mode == MODE_FWD) {
        cur[to] = LA(cur[to], nb[from] + (eP + tP));
    } else if (mode == MODE_BWD) {
        nb[from] = LA(nb[from], cur[to] + (eP + tP));
    } else {
        double p = exp(nb[from] + cur[to] + (eP + tP) - dp->total);
        if (dp->m->sm_type == SM_THREE_STATE) {
            dp->expT[from * 3 + to] += p;
            if (to == ST_X && kx >= 0) dp->expSkip[kx] += p;
        } else if (dp->m->sm_type == SM_THREE_STATE_HDP) {
            dp->expT[from * 3 + to] += p;
            if (to == ST_M && p >= g_asg_thr && g_asg != NULL) {
                if (g_asg_n < g_asg_cap) {
                    g_asg[3 * g_asg_n] = from;
                    g_asg[3 * g_asg_n + 1] = g_cell_kmer - g_asg_ref;
                    g_asg[3 * g_asg_n + 2] = (g_cell_ev - g_asg_ev) / 3;
                }
                g_asg_n++;
            }
        } else {
            if (from == ST_M && to == ST_X) dp->expT[bin] += p;
            if (from == ST_X && to == ST_X) dp->expT[bin + 30] += p;
        }
    }
}

/* sm:388-419: skip bin from |mu(k_i) - mu(k_{i-1})| in 0.5 pA steps, clamped to 29 */
static int skip_bin(const double *match, const char *p) {
    double d = fabs(mdl(match, kmer_index(p + 1), 0) - mdl(match, kmer_index(p), 0));
    int64_t b = (int64_t) (d / 0.5);
    return b >= 30 ? 29 : (int) b;
}

/* ---- echelon (getStateMachineEchelon, sm:1773-1784): events may cover n = 1 .. 5 k-mers --------------------------
 * The reference pads the nucleotide sequence with 30 'n' (ref:282-285; tests/signalPairwiseTest.c:1422) and reads up
 * to 6 n + 6 characters past the getKmer2 pointer; rch() is that padded read. */
static char rch(const Dp *dp, int64_t i) { return (i >= 0 && i < dp->refLen) ? dp->ref[i] : 'n'; }
static int32_t kmer_index_at(const Dp *dp, int64_t i) {
    char b[KMER_LEN];
    for (int j = 0; j < KMER_LEN; j++) b[j] = rch(dp, i + j);
    return kmer_index(b);
}
/* sm:388-419 on the padded sequence */
static int skip_bin_at(const Dp *dp, int64_t i) {
    double d = fabs(mdl(dp->m->match, kmer_index_at(dp, i + 1), 0) - mdl(dp->m->match, kmer_index_at(dp, i), 0));
    int64_t b = (int64_t) (d / 0.5);
    return b >= 30 ? 29 : (int) b;
}
/* emissions_signal_multipleKmerMatchProb (sm:530-549), quirks included: the fold starts from 0.0 (= log 1), and the
 * run-off test looks at the single character 6 n past the pointer */
static double multi_kmer_match(const Dp *dp, const double *tbl, int64_t i, const double *ev, int n) {
    double p = 0.0;
    for (int j = 0; j < n; j++) {
        char last = rch(dp, i + KMER_LEN * n);
        if (last >= 'A' && last <= 'Z') p = LA(p, emit_gauss_invgauss(tbl, kmer_index_at(dp, i + j + 1), ev));
        else return NEG_INF;
    }
    return p - log((double) n);
}
/* emissions_signal_poissonPosteriorProb (sm:345-370) of the event's duration */
static double duration_prob(const double *ev, int n) {
    static const double lfact[6] = { 0.0, 0.0, 0.69314718056, 1.79175946923, 3.17805383035, 4.78749174278 };
    double lambda = ev[2] / 0.00332005312085;
    return (n + 1) * 0.1397619423751586 + n * log(lambda) - lfact[n] - 2 * lambda;
}

/* get_nanopore_kmer_density (impl/nanopore_hdp.c:390-392) -> dir_proc_density (impl/hdp.c:2577-2599) ->
 * grid_spline_interp (impl/hdp_math_utils.c:471-495) on linspace(start, stop, len) (:497-510).  k = ACGT k-mer index. */
double oracle_hdp_density(const OracleModel *m, int32_t k, double q) {
    const int64_t n = m->hdp_len;
    const int32_t row = k >= 0 ? m->hdp_kmer_row[k] : -1;
    if (row < 0) return NAN;                                  /* the reference exits: character outside the alphabet */
    const double *y = m->hdp_density + (int64_t) row * n, *slope = m->hdp_slopes + (int64_t) row * n;
    const double step = (m->hdp_stop - m->hdp_start) / ((double) (n - 1));
#define HDP_X(i) ((i) == n - 1 ? m->hdp_stop : m->hdp_start + (i) * step)
    double interp;
    if (q <= HDP_X(0)) interp = y[0] - slope[0] * (HDP_X(0) - q);
    else if (q >= HDP_X(n - 1)) interp = y[n - 1] + slope[n - 1] * (q - HDP_X(n - 1));
    else {
        double dx = HDP_X(1) - HDP_X(0);
        int64_t il = (int64_t) ((q - HDP_X(0)) / dx);
        if (il > n - 2) il = n - 2;
        int64_t ir = il + 1;
        double dy = y[ir] - y[il];
        double a = slope[il] * dx - dy;
        double b = dy - slope[ir] * dx;
        double tl = (q - HDP_X(il)) / dx;
        double tr = 1.0 - tl;
        interp = tr * y[il] + tl * y[ir] + tl * tr * (a * tr + b * tl);
    }
#undef HDP_X
    return interp > 0.0 ? interp : 0.0;
}

/* One cell: sm:1305-1334 (threeState) and sm:1368-1409 (vanilla).  ix / iy are SEQUENCE indices (matrix - 1);
 * the neighbour pointers may be NULL exactly as dpDiagonal_getCell returns NULL outside the band (ref:562-568). */
static void cell(Dp *dp, int mode, double *cur, double *lower, double *middle, double *upper, int64_t ix, int64_t iy) {
    const OracleModel *m = dp->m;
    static const double nullEvent[3] = { NEG_INF, 0.0, 0.0 };         /* ref:261-262 */
    const double *ev = iy >= 0 ? dp->ev + 3 * iy : nullEvent;
    if (m->sm_type == SM_THREE_STATE) {
        /* sequence_getKmer (ref:314-318): index < 0 reads the literal "n" => invalid k-mer */
        int32_t k = ix >= 0 ? kmer_index(dp->ref + ix) : -1;
        const double *t = m->trans;
        if (lower) {
            double eP = k < 0 ? NEG_INF : m->gapx[k];                  /* sm:175-187 */
            transition(dp, mode, lower, cur, ST_M, ST_X, eP, t[3], k, 0);
            transition(dp, mode, lower, cur, ST_X, ST_X, eP, t[5], k, 0);
            transition(dp, mode, lower, cur, ST_Y, ST_X, eP, t[7], k, 0);
        }
        if (middle) {
            double eP = emit_two_gauss(m->match, k, ev);
            transition(dp, mode, middle, cur, ST_M, ST_M, eP, t[0], k, 0);
            transition(dp, mode, middle, cur, ST_X, ST_M, eP, t[1], k, 0);
            transition(dp, mode, middle, cur, ST_Y, ST_M, eP, t[2], k, 0);
        }
        if (upper) {
            double eP = emit_two_gauss(m->gapy, k, ev);
            transition(dp, mode, upper, cur, ST_M, ST_Y, eP, t[4], k, 0);
            transition(dp, mode, upper, cur, ST_Y, ST_Y, eP, t[6], k, 0);
        }
    } else if (m->sm_type == SM_THREE_STATE_HDP) {
        /* stateMachine3HDP_cellCalculate (sm:1336-1366): gap X log(0.1); match and gap Y the DENSITY (not its log) of the
         * k-mer's distribution at the event mean; sequence_getKmer3 (ref:327-331): index < 0 reads k-mer 0 */
        int32_t k = kmer_index(dp->ref + (ix >= 0 ? ix : 0));
        const double *t = m->trans;
        g_cell_kmer = dp->ref + (ix >= 0 ? ix : 0); g_cell_ev = ev;
        if (lower) {
            double eP = -2.3025850929940455;
            transition(dp, mode, lower, cur, ST_M, ST_X, eP, t[3], k, 0);
            transition(dp, mode, lower, cur, ST_X, ST_X, eP, t[5], k, 0);
            transition(dp, mode, lower, cur, ST_Y, ST_X, eP, t[7], k, 0);
        }
        if (middle) {
            double eP = oracle_hdp_density(m, k, ev[0]);
            transition(dp, mode, middle, cur, ST_M, ST_M, eP, t[0], k, 0);
            transition(dp, mode, middle, cur, ST_X, ST_M, eP, t[1], k, 0);
            transition(dp, mode, middle, cur, ST_Y, ST_M, eP, t[2], k, 0);
        }
        if (upper) {
            double eP = oracle_hdp_density(m, k, ev[0]);
            transition(dp, mode, upper, cur, ST_M, ST_Y, eP, t[4], k, 0);
            transition(dp, mode, upper, cur, ST_Y, ST_Y, eP, t[6], k, 0);
        }
    } else if (m->sm_type == SM_ECHELON) {
        /* stateMachineEchelon_cellCalculate (sm:1411-1460); getKmer2 pointer clamped at 0 (ref:320-325) */
        const int64_t i = ix > 0 ? ix - 1 : 0;
        const int bin = skip_bin_at(dp, i);
        const double a_mx = m->gapx[bin], la_mx = log(a_mx), la_mh = log(1 - a_mx);
        const double a_xx = m->gapx[bin + 30], la_xx = log(a_xx), la_xh = log(1 - a_xx);
        if (lower) {
            for (int n = 1; n < 6; n++) transition(dp, mode, lower, cur, n, ECH_GAPX, 0, la_mx, -1, 0);
            transition(dp, mode, lower, cur, ECH_GAPX, ECH_GAPX, 0, la_xx, -1, 0);
        }
        if (middle) {
            for (int n = 1; n < 6; n++)
                for (int from = 0; from < 6; from++)
                    transition(dp, mode, middle, cur, from, n, multi_kmer_match(dp, m->match, i, ev, n),
                               la_mh + duration_prob(ev, n), -1, 0);
            for (int n = 1; n < 6; n++)
                transition(dp, mode, middle, cur, ECH_GAPX, n, multi_kmer_match(dp, m->match, i, ev, n),
                           la_xh + duration_prob(ev, n), -1, 0);
        }
        if (upper) {
            for (int n = 1; n < 6; n++)
                transition(dp, mode, upper, cur, n, 0, emit_gauss_invgauss(m->gapy, kmer_index_at(dp, i + 1), ev),
                           la_mh + duration_prob(ev, 0), -1, 0);
        }
    } else if (m->sm_type == SM_FOUR_STATE) {
        /* stateMachine4_cellCalculate (sm:867-897); emissions of getStateMachine4 (sm:1750-1759): gap X from the
         * all-zero k-mer table, match and short gap Y the two-Gaussian model on the MATCH / GAP_Y tables */
        int32_t k = ix >= 0 ? kmer_index(dp->ref + ix) : -1;
        const double *t = m->trans4;
        if (lower) {
            double eP = k < 0 ? NEG_INF : m->gapx[k];
            transition(dp, mode, lower, cur, ST_M, ST_X, eP, t[T4_GAP_SHORT_OPEN_X], k, 0);
            transition(dp, mode, lower, cur, ST_X, ST_X, eP, t[T4_GAP_SHORT_EXTEND_X], k, 0);
            transition(dp, mode, lower, cur, ST_M, ST_LX, eP, t[T4_GAP_LONG_OPEN_X], k, 0);
            transition(dp, mode, lower, cur, ST_LX, ST_LX, eP, t[T4_GAP_LONG_EXTEND_X], k, 0);
            transition(dp, mode, lower, cur, ST_Y, ST_LX, eP, t[T4_GAP_LONG_SWITCH_TO_X], k, 0);
        }
        if (middle) {
            double eP = emit_two_gauss(m->match, k, ev);
            transition(dp, mode, middle, cur, ST_M, ST_M, eP, t[T4_MATCH_CONTINUE], k, 0);
            transition(dp, mode, middle, cur, ST_X, ST_M, eP, t[T4_MATCH_FROM_SHORT_GAP_X], k, 0);
            transition(dp, mode, middle, cur, ST_Y, ST_M, eP, t[T4_MATCH_FROM_SHORT_GAP_Y], k, 0);
            transition(dp, mode, middle, cur, ST_LX, ST_M, eP, t[T4_MATCH_FROM_LONG_GAP_X], k, 0);
        }
        if (upper) {
            double eP = emit_two_gauss(m->gapy, k, ev);
            transition(dp, mode, upper, cur, ST_M, ST_Y, eP, t[T4_GAP_SHORT_OPEN_Y], k, 0);
            transition(dp, mode, upper, cur, ST_Y, ST_Y, eP, t[T4_GAP_SHORT_EXTEND_Y], k, 0);
        }
    } else {
        /* sequence_getKmer2 (ref:320-325): pointer to the PREVIOUS k-mer, clamped at 0 */
        const char *p = dp->ref + (ix > 0 ? ix - 1 : 0);
        int bin = skip_bin(m->match, p);
        int32_t k = kmer_index(p + 1);
        double a_mx = m->gapx[bin];
        double a_my = (1 - a_mx) * m->vanilla[0];
        double a_mm = 1.0f - a_my - a_mx;
        double a_yy = m->vanilla[1];
        double a_ym = 1.0f - a_yy;
        double a_xx = m->gapx[bin + 30];
        double a_xm = 1.0f - a_xx;
        if (lower) {
            transition(dp, mode, lower, cur, ST_M, ST_X, 0, log(a_mx), k, bin);
            transition(dp, mode, lower, cur, ST_X, ST_X, 0, log(a_xx), k, bin);
        }
        if (middle) {
            double eP = emit_gauss_invgauss(m->match, k, ev);
            transition(dp, mode, middle, cur, ST_M, ST_M, eP, log(a_mm), k, bin);
            transition(dp, mode, middle, cur, ST_X, ST_M, eP, log(a_xm), k, bin);
            transition(dp, mode, middle, cur, ST_Y, ST_M, eP, log(a_ym), k, bin);
        }
        if (upper) {
            double eP = emit_gauss_invgauss(m->gapy, k, ev);
            transition(dp, mode, upper, cur, ST_M, ST_Y, eP, log(a_my), k, bin);
            transition(dp, mode, upper, cur, ST_Y, ST_Y, eP, log(a_yy), k, bin);
        }
    }
}

/* ref:681-712: sweep one diagonal in ascending xmy.  cur from matrix `mc`, lower/upper from `m1` (xay-1), middle from `m2`. */
static void sweep(Dp *dp, int mode, double **mc, int64_t xay, double **m1, double **m2, int useLowerUpper) {
    for (int64_t xmy = dp->xmyL[xay]; xmy <= dp->xmyR[xay]; xmy += 2) {
        int64_t ix = (xay + xmy) / 2 - 1, iy = (xay - xmy) / 2 - 1;
        double *cur = cell_at(dp, mc, xay, xmy);
        double *lower = useLowerUpper ? cell_at(dp, m1, xay - 1, xmy - 1) : NULL;
        double *middle = cell_at(dp, m2, xay - 2, xmy);
        double *upper = useLowerUpper ? cell_at(dp, m1, xay - 1, xmy + 1) : NULL;
        cell(dp, mode, cur, lower, middle, upper, ix, iy);
    }
}

/* ref:391-397 and ref:587-597: left folds, state 0 first, cells in ascending xmy from -inf */
static double diag_dot(Dp *dp, double *a, double *b, int64_t xay) {
    double tot = NEG_INF;
    int64_t w = width_of(dp, xay);
    const int S = dp->S;
    for (int64_t i = 0; i < w; i++) {
        double c = a[S * i] + b[S * i];
        for (int st = 1; st < S; st++) c = LA(c, a[S * i + st] + b[S * i + st]);
        tot = LA(tot, c);
    }
    return tot;
}

/* ref:736-754 */
static double total_probability(Dp *dp, int64_t xay) {
    double tot = diag_dot(dp, dp->F[xay], dp->B[xay], xay);
    if (g_dbg_t1) { g_dbg_t1[xay] = tot; g_dbg_t2[xay] = NAN; }
    if (xay + 1 <= dp->lX + dp->lY && dp->B[xay + 1] != NULL && xay - 1 >= 0 && dp->F[xay - 1] != NULL) {
        /* match-only forward step from F[xay-1] into a -inf clone shaped like B[xay+1] */
        double **tmp = calloc(dp->lX + dp->lY + 2, sizeof(double *));
        new_diag(dp, tmp, xay + 1, NEG_INF);
        sweep(dp, MODE_FWD, tmp, xay + 1, NULL, dp->F, 0);
        double t2 = diag_dot(dp, tmp[xay + 1], dp->B[xay + 1], xay + 1);
        if (g_dbg_t2) g_dbg_t2[xay] = t2;
        tot = LA(tot, t2);
        free(tmp[xay + 1]);
        free(tmp);
    }
    return tot;
}

typedef struct { int64_t *out; int64_t cap, n; int64_t offX, offY; } PairSink;

/* ref:756-795 */
static void posterior_diag(Dp *dp, int64_t xay, double total, double threshold, PairSink *sink) {
    if (xay == g_dbg_row && g_dbg_F) {
        int64_t w = (dp->xmyR[xay] - dp->xmyL[xay]) / 2 + 1;
        for (int64_t i = 0; i < dp->S * w; i++) { g_dbg_F[i] = dp->F[xay][i]; g_dbg_B[i] = dp->B[xay][i]; }
    }
    for (int64_t xmy = dp->xmyL[xay]; xmy <= dp->xmyR[xay]; xmy += 2) {
        int64_t x = (xay + xmy) / 2, y = (xay - xmy) / 2;
        if (x > 0 && y > 0 && dp->m->sm_type == SM_ECHELON) {
            /* diagonalCalculationMultiPosteriorMatchProbs (ref:797-839): state s = 1 .. 5 pairs the event with s k-mers */
            const double *f = cell_at(dp, dp->F, xay, xmy), *b = cell_at(dp, dp->B, xay, xmy);
            for (int st = 1; st < 6; st++) {
                double p = exp((f[st] + b[st]) - total);
                if (p >= threshold) {
                    if (p > 1.0) p = 1.0;
                    p = floor(p * 10000000);
                    for (int n = 0; n < st; n++) {
                        if (sink->n < sink->cap) {
                            sink->out[3 * sink->n] = (int64_t) p;
                            sink->out[3 * sink->n + 1] = (x + n) - 1 + sink->offX;
                            sink->out[3 * sink->n + 2] = y - 1 + sink->offY;
                        }
                        sink->n++;
                    }
                }
            }
        } else if (x > 0 && y > 0) {
            double lp = cell_at(dp, dp->F, xay, xmy)[ST_M] + cell_at(dp, dp->B, xay, xmy)[ST_M] - total;
            double p = exp(lp);
            if (p >= threshold) {
                if (p > 1.0) p = 1.0;
                if (sink->n < sink->cap) {
                    sink->out[3 * sink->n] = g_dbg_logp ? (int64_t) llround(lp * 1e9) : (int64_t) floor(p * 10000000);
                    sink->out[3 * sink->n + 1] = x - 1 + sink->offX;
                    sink->out[3 * sink->n + 2] = y - 1 + sink->offY;
                }
                sink->n++;
            }
        }
    }
}

static void state_vector(const OracleModel *m, int which /*0 start,1 raggedStart,2 end,3 raggedEnd*/, double *v) {
    if (m->sm_type == SM_ECHELON) {
        /* sm:1237-1262; the end "probabilities" are NOT logs in the reference (sm:1617-1619) and are used as they are */
        for (int st = 0; st < 7; st++) v[st] = NEG_INF;
        if (which == 0) v[1] = 0;
        else if (which == 1) v[ECH_GAPX] = 0;
        else { for (int st = 0; st < 6; st++) v[st] = 0.79015888282447311; v[ECH_GAPX] = 0.19652425498269727; }
        return;
    }
    if (m->sm_type == SM_FOUR_STATE) {
        const double *t = m->trans4;
        if (which == 0) { v[0] = 0; v[1] = v[2] = v[3] = NEG_INF; }                          /* sm:775-779 (shared with 5) */
        else if (which == 1) { v[0] = v[1] = NEG_INF; v[2] = 0; v[3] = 0; }                  /* sm:791-794 */
        else if (which == 2) { v[0] = t[T4_MATCH_CONTINUE]; v[1] = t[T4_MATCH_FROM_SHORT_GAP_X];      /* sm:796-810 */
                               v[2] = t[T4_MATCH_FROM_SHORT_GAP_Y]; v[3] = t[T4_MATCH_FROM_LONG_GAP_X]; }
        else { v[0] = v[1] = v[2] = t[T4_GAP_LONG_OPEN_X]; v[3] = t[T4_GAP_LONG_EXTEND_X]; }  /* sm:812-826 */
        return;
    }
    if (which == 0) { v[0] = 0; v[1] = NEG_INF; v[2] = NEG_INF; return; }                 /* sm:1168-1172 */
    if (which == 1) { v[0] = NEG_INF; v[1] = 0; v[2] = 0; return; }                       /* sm:1174-1177 */
    if (m->sm_type == SM_THREE_STATE || m->sm_type == SM_THREE_STATE_HDP) {
        const double *t = m->trans;
        if (which == 2) { v[0] = t[0]; v[1] = t[1]; v[2] = t[2]; }                        /* sm:1179-1192 */
        else { v[0] = (t[3] + t[4]) / 2.0; v[1] = t[5]; v[2] = t[6]; }                    /* sm:1194-1207 */
    } else {
        const double *e = m->vanilla;
        if (which == 2) { v[0] = e[2]; v[1] = e[3]; v[2] = e[4]; }                        /* sm:1223-1235 */
        else { v[0] = (e[3] + e[4]) / 2.0; v[1] = e[3]; v[2] = e[4]; }                    /* sm:1209-1221 */
    }
}
static void fill_diag(Dp *dp, double **mat, int64_t xay, const double *v) {
    double *c = new_diag(dp, mat, xay, 0.0);
    for (int64_t i = 0; i < width_of(dp, xay); i++) for (int st = 0; st < dp->S; st++) c[dp->S * i + st] = v[st];
}

/* ref:870-1006: one banded region with periodic traceback.
 * mode_exp = 0: posterior pairs into sink; 1: expectations into dp->expT/expSkip, likelihood += total per diagonal.
 * totals (optional, region-local xay index): the totalProbability handed to each diagonal. */
static void banded_region(const OracleModel *m, const char *ref, int64_t lX, const double *ev, int64_t lY,
                          const int64_t *anchors, int64_t nA, const OracleParams *p, int raggedLeft, int raggedRight,
                          int mode_exp, PairSink *sink, double *expT, double *expSkip, double *likelihood,
                          double *totals) {
    int64_t D = lX + lY;
    if (D == 0) return;
    Dp dp; memset(&dp, 0, sizeof(dp));
    dp.m = m; dp.S = m->sm_type == SM_FOUR_STATE ? 4 : (m->sm_type == SM_ECHELON ? 7 : 3); dp.refLen = (int64_t) strlen(ref); dp.ref = ref; dp.ev = ev; dp.lX = lX; dp.lY = lY; dp.expT = expT; dp.expSkip = expSkip;
    int64_t *xl = malloc(sizeof(int64_t) * (D + 1)), *xr = malloc(sizeof(int64_t) * (D + 1));
    oracle_band(anchors, nA, lX, lY, p->diagonalExpansion, xl, xr);
    dp.xmyL = xl; dp.xmyR = xr;
    dp.F = calloc(D + 2, sizeof(double *)); dp.B = calloc(D + 2, sizeof(double *));
    double v[7];
    state_vector(m, raggedLeft ? 1 : 0, v);
    fill_diag(&dp, dp.F, 0, v);
    int64_t tracedBackTo = 0;
    for (int64_t xay = 1; xay <= D; xay++) {
        new_diag(&dp, dp.F, xay, NEG_INF);
        sweep(&dp, MODE_FWD, dp.F, xay, dp.F, dp.F, 1);
        int atEnd = xay == D;
        int tracebackPoint = xay >= tracedBackTo + p->minDiagsBetweenTraceBack
                             && width_of(&dp, xay) <= p->diagonalExpansion * 2 + 1;
        if (!(atEnd || tracebackPoint)) continue;
        state_vector(m, (atEnd && raggedRight) ? 3 : 2, v);
        fill_diag(&dp, dp.B, xay, v);
        if (xay > tracedBackTo + 1) new_diag(&dp, dp.B, xay - 1, NEG_INF);
        int64_t tracedBackFrom = xay - (atEnd ? 0 : p->traceBackDiagonals + 1);
        double total = NEG_INF;
        int64_t count = 0;
        for (int64_t d2 = xay; d2 > tracedBackTo; d2--) {
            if (d2 > tracedBackTo + 2) new_diag(&dp, dp.B, d2 - 2, NEG_INF);
            if (d2 > tracedBackTo + 1) sweep(&dp, MODE_BWD, dp.B, d2, dp.B, dp.B, 1);
            if (d2 <= tracedBackFrom) {
                if (count++ % 10 == 0) total = total_probability(&dp, d2);
                if (totals) totals[d2] = total;
                if (!mode_exp) {
                    posterior_diag(&dp, d2, total, p->threshold, sink);
                } else {
                    /* ref:841-863: likelihood += total once per diagonal; cur = B[d2], neighbours from F */
                    *likelihood += total;
                    dp.total = total;
                    sweep(&dp, MODE_EXP, dp.B, d2, dp.F, dp.F, 1);
                }
                if (d2 < tracedBackFrom || atEnd) drop_diag(dp.F, d2);
            }
            if (d2 + 1 <= D) drop_diag(dp.B, d2 + 1);
        }
        drop_diag(dp.B, tracedBackTo + 1);
        drop_diag(dp.F, tracedBackTo);
        tracedBackTo = tracedBackFrom;
    }
    for (int64_t i = 0; i <= D; i++) { drop_diag(dp.F, i); drop_diag(dp.B, i); }
    free(dp.F); free(dp.B); free(xl); free(xr);
}

/* ref:1356-1422 + ref:1456-1484 / 1571-1591: split at large anchor gaps, run each region, shift coordinates back. */
static void run_split(const OracleModel *m, const char *ref, int64_t lX, const double *ev, int64_t lY,
                      const int64_t *anchors, int64_t nA, const OracleParams *p, int raggedLeft, int raggedRight,
                      int mode_exp, PairSink *sink, double *expT, double *expSkip, double *likelihood, double *totals) {
    int64_t *sp = malloc(sizeof(int64_t) * 4 * (nA + 2));
    int64_t nS = oracle_split_points(anchors, nA, lX, lY, p->splitMatrixBiggerThanThis, raggedLeft, raggedRight, sp);
    int64_t j = 0;
    int64_t *sub = malloc(sizeof(int64_t) * 2 * (nA + 1));
    for (int64_t i = 0; i < nS; i++) {
        int64_t x1 = sp[4 * i], y1 = sp[4 * i + 1], x2 = sp[4 * i + 2], y2 = sp[4 * i + 3];
        int64_t k = 0;
        while (j < nA) {
            int64_t x = anchors[2 * j], y = anchors[2 * j + 1];
            if (x + y >= x2 + y2) break;
            sub[2 * k] = x - x1; sub[2 * k + 1] = y - y1; k++; j++;
        }
        int64_t first = 0;
        if (sink) { sink->offX = x1; sink->offY = y1; first = sink->n; }
        /* totals are only meaningful for a single region (region-local diagonal index) */
        banded_region(m, ref + x1, x2 - x1, ev + 3 * y1, y2 - y1, sub, k, p, raggedLeft || i > 0,
                      raggedRight || i < nS - 1, mode_exp, sink, expT, expSkip, likelihood, nS == 1 ? totals : NULL);
        /* ref:1447-1454: each region's pairs are moved to the result list by popping, i.e. in reverse order */
        if (sink && sink->n <= sink->cap) {
            for (int64_t a = first, b = sink->n - 1; a < b; a++, b--)
                for (int c = 0; c < 3; c++) {
                    int64_t t = sink->out[3 * a + c]; sink->out[3 * a + c] = sink->out[3 * b + c]; sink->out[3 * b + c] = t;
                }
        }
    }
    free(sub); free(sp);
}

int64_t oracle_align_banded(const OracleModel *m, const char *ref, int64_t lX, const double *events, int64_t lY,
                            const int64_t *anchors, int64_t nA, const OracleParams *p, int raggedLeft, int raggedRight,
                            int64_t *out, int64_t cap, double *totals) {
    PairSink sink = { out, cap, 0, 0, 0 };
    if (totals) for (int64_t i = 0; i <= lX + lY; i++) totals[i] = NAN;
    run_split(m, ref, lX, events, lY, anchors, nA, p, raggedLeft, raggedRight, 0, &sink, NULL, NULL, NULL, totals);
    return sink.n;
}

/* threeState: expT[9], expSkip[4096]; vanilla: expT[60], expSkip unused.  Accumulates (caller pre-fills pseudocounts). */
void oracle_expectations(const OracleModel *m, const char *ref, int64_t lX, const double *events, int64_t lY,
                         const int64_t *anchors, int64_t nA, const OracleParams *p, int raggedLeft, int raggedRight,
                         double *expT, double *expSkip, double *likelihood) {
    run_split(m, ref, lX, events, lY, anchors, nA, p, raggedLeft, raggedRight, 1, NULL, expT, expSkip, likelihood, NULL);
}

/* threeStateHdp: expT[9] transition sums, the likelihood, and the assignments (from, k-mer position, event index) in the
 * order the reference's lists hold them; returns their number (which may exceed cap). */
int64_t oracle_hdp_expectations(const OracleModel *m, const char *ref, int64_t lX, const double *events, int64_t lY,
                                const int64_t *anchors, int64_t nA, const OracleParams *p, int raggedLeft, int raggedRight,
                                double threshold, double *expT, double *likelihood, int64_t *asgOut, int64_t cap) {
    double skip[1] = { 0 };
    g_asg = asgOut; g_asg_cap = cap; g_asg_n = 0; g_asg_thr = threshold; g_asg_ref = ref; g_asg_ev = events;
    run_split(m, ref, lX, events, lY, anchors, nA, p, raggedLeft, raggedRight, 1, NULL, expT, skip, likelihood, NULL);
    g_asg = NULL;
    return g_asg_n;
}

/* ref:1512-1569: full matrix (band with no anchors, expansion 2), ONE total at the last diagonal, posteriors in
 * ascending diagonal order. */
int64_t oracle_align_unbanded(const OracleModel *m, const char *ref, int64_t lX, const double *events, int64_t lY,
                              const OracleParams *p, int raggedLeft, int raggedRight, int64_t *out, int64_t cap,
                              double *totalOut) {
    int64_t D = lX + lY;
    Dp dp; memset(&dp, 0, sizeof(dp));
    dp.m = m; dp.S = m->sm_type == SM_FOUR_STATE ? 4 : (m->sm_type == SM_ECHELON ? 7 : 3); dp.refLen = (int64_t) strlen(ref); dp.ref = ref; dp.ev = events; dp.lX = lX; dp.lY = lY;
    int64_t *xl = malloc(sizeof(int64_t) * (D + 1)), *xr = malloc(sizeof(int64_t) * (D + 1));
    oracle_band(NULL, 0, lX, lY, 2, xl, xr);
    dp.xmyL = xl; dp.xmyR = xr;
    dp.F = calloc(D + 2, sizeof(double *)); dp.B = calloc(D + 2, sizeof(double *));
    for (int64_t i = 0; i <= D; i++) { new_diag(&dp, dp.F, i, NEG_INF); new_diag(&dp, dp.B, i, NEG_INF); }
    double v[7];
    state_vector(m, raggedLeft ? 1 : 0, v); fill_diag(&dp, dp.F, 0, v);
    state_vector(m, raggedRight ? 3 : 2, v); fill_diag(&dp, dp.B, D, v);
    for (int64_t i = 0; i <= D; i++) sweep(&dp, MODE_FWD, dp.F, i, dp.F, dp.F, 1);
    for (int64_t i = D; i > 0; i--) sweep(&dp, MODE_BWD, dp.B, i, dp.B, dp.B, 1);
    double total = total_probability(&dp, D);
    if (totalOut) *totalOut = total;
    PairSink sink = { out, cap, 0, 0, 0 };
    for (int64_t i = 0; i <= D; i++) posterior_diag(&dp, i, total, p->threshold, &sink);
    for (int64_t i = 0; i <= D; i++) { drop_diag(dp.F, i); drop_diag(dp.B, i); }
    free(dp.F); free(dp.B); free(xl); free(xr);
    return sink.n;
}

/* Band cells C = sum_d width(d) over all split regions (SURVEY.md 8(d) definition of the work unit). */
int64_t oracle_band_cells(const int64_t *anchors, int64_t nA, int64_t lX, int64_t lY, const OracleParams *p,
                          int raggedLeft, int raggedRight) {
    int64_t *sp = malloc(sizeof(int64_t) * 4 * (nA + 2));
    int64_t nS = oracle_split_points(anchors, nA, lX, lY, p->splitMatrixBiggerThanThis, raggedLeft, raggedRight, sp);
    int64_t j = 0, cells = 0;
    int64_t *sub = malloc(sizeof(int64_t) * 2 * (nA + 1));
    for (int64_t i = 0; i < nS; i++) {
        int64_t x1 = sp[4 * i], y1 = sp[4 * i + 1], x2 = sp[4 * i + 2], y2 = sp[4 * i + 3], k = 0;
        while (j < nA) {
            int64_t x = anchors[2 * j], y = anchors[2 * j + 1];
            if (x + y >= x2 + y2) break;
            sub[2 * k] = x - x1; sub[2 * k + 1] = y - y1; k++; j++;
        }
        int64_t D = (x2 - x1) + (y2 - y1);
        int64_t *xl = malloc(sizeof(int64_t) * (D + 1)), *xr = malloc(sizeof(int64_t) * (D + 1));
        oracle_band(sub, k, x2 - x1, y2 - y1, p->diagonalExpansion, xl, xr);
        for (int64_t d = 0; d <= D; d++) cells += (xr[d] - xl[d]) / 2 + 1;
        free(xl); free(xr);
    }
    free(sub); free(sp);
    return cells;
}
