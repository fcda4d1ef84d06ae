// cpecan_generic.cuh -- the reference's other two signal state machines on device: fourState (match, short gap X,
// short gap Y, long gap X; impl/stateMachine.c:867-897, factory :1750-1759) and echelon (match0 .. match5, gap X: an
// event may cover 1 .. 5 k-mers; :1411-1460, emissions :530-549, Poisson duration :345-370, factory :1773-1784), with
// the banded schedule of getPosteriorProbsWithBanding (impl/pairwiseAligner.c:870-1006) and, for echelon,
// diagonalCalculationMultiPosteriorMatchProbs (:797-839).
//
// Also here: threeStateHdp (impl/stateMachine.c:1336-1366, factory :1738-1749) -- the three-state
// topology with the match and gap-Y emissions read from a NanoporeHDP: the spline-interpolated posterior predictive
// density of the k-mer's Dirichlet process at the event mean (impl/nanopore_hdp.c:390-392, impl/hdp.c:2577-2599,
// impl/hdp_math_utils.c:471-495).  The reference adds that DENSITY (not its logarithm) to the log-transition; so does
// this kernel.
//
// These are the reference's experimental machines (its own test accepts 857 of 1000 pairs for echelon); here they run
// in FP64 -- B200 issues FP64 at half the FP32 rate -- with the reference's arithmetic as it stands: absolute
// log-probabilities, logAdd as lo + cubic(hi - lo), transitions pulled on the way forward and PUSHED on the way back
// in the reference's order (a cell's backward value is folded from diagonal d+2 first, then from the cell above, then
// from the cell to the right), so the results agree with the reference to the last digits and no offsets are needed.
// One warp per alignment; the last diagonals live in shared memory as S planes of N doubles (3 buffers forward, 4 on
// the way back: B of d+1 is kept for the second term of the total probability); forward cells spill to HBM as S doubles.
// A cell reads only its column's record and its event: k-mer lookups, model scaling and every logarithm that does not
// depend on the cell happen once per column / per event in k_prep_generic.
// Band and traceback points come from k_plan3.  Expectations are not defined for these machines in the reference
// (cellCalculateUpdateExpectations is NULL for echelon) and are refused.
#pragma once
#include <math_constants.h>
#include "cpecan_kernels.cuh"

namespace cpecan {

#define CPG_NI (-CUDART_INF)
#ifndef CPG_MINB
#define CPG_MINB 1
#endif

struct GenParams {
    int sm;                 // 6 = fourState, 5 = echelon, 7 = threeStateHdp, and -- for band shapes k_align3 does not take (odd expansions) --
                            // 2 = threeState, 4 = vanilla (StateMachineType, inc/stateMachine.h:20-29)
    double t3[9];           // threeState transitions, StateMachine3 field order
    double van[5];          // vanilla: M_TO_Y_NOT_X, E_TO_E, END_MATCH, END_FROM_X, END_FROM_Y
    double t4[11];          // fourState transitions: MATCH_CONTINUE, MATCH_FROM_SHORT_GAP_X, MATCH_FROM_SHORT_GAP_Y,
                            // MATCH_FROM_LONG_GAP_X, GAP_SHORT_OPEN_X, GAP_SHORT_EXTEND_X, GAP_SHORT_OPEN_Y, GAP_SHORT_EXTEND_Y,
                            // GAP_LONG_OPEN_X, GAP_LONG_EXTEND_X, GAP_LONG_SWITCH_TO_X
    double threshold;
};

struct KernelArgsG {
    const Item *items;
    const int *order;
    int n_items;
    int *queue;
    const char *ref;
    const long long *ref_off;
    const double *events;         // reference layout: (mean, noise, duration) per event
    const long long *ev_src_off;
    const ModelTables *models;
    const double *scale;          // 5 per item or null
    const int2 *bands;            // plan: (lo, hi) of every diagonal, item i at band_off[i]
    const long long *band_off;
    const int *tbs;
    const int *flags;
    double *scratch;              // per warp: ring_rows rows of N cells of S doubles
    long long scratch_stride;     // doubles per warp
    int ring_rows, ringN;
    int *pairs;
    ItemOut *out;
    double *totals;
    double *expect;               // EXPECTATION mode: the batch sums (layout of CPECAN_N_EXPECT / CPECAN_N_EXPECT_VANILLA)
    const double *colp;           // k_prep_generic: per-column records, gen_ncol(sm) planes of lX + 2 doubles per item at
                                  // gen_ncol(sm) * xp_off; null for echelon
    const double *rowp;           // vanilla, echelon: log(noise) of every event (item at ev_off, entry y = matrix row);
    long long row2;               // echelon: rowp[row2 + ...] = log(duration / 0.00332005312085) of every event
    DevParams P;
    GenParams G;
};

__host__ __device__ inline size_t generic_smem_bytes(int ringN, int S) { return (size_t) 4 * S * ringN * sizeof(double); }

// impl/pairwiseAligner.c:235-255 with the reference's float-literal coefficients.  The four segments and the two
// early returns are SELECTED, not branched on: neighbouring lanes sit in different segments all the time, and a divergent
// warp would evaluate every segment one after the other.  Same coefficients, same Horner form, same result.
__device__ __forceinline__ double g_poly(double x) {
    const bool s0 = x <= 1.00f, s1 = x <= 2.50f, s2 = x <= 4.50f;
    const double c3 = s0 ? -0.009350833524763f : (s1 ? -0.014532321752540f : (s2 ? -0.004605031767994f : -0.000458661602210f));
    const double c2 = s0 ? 0.130659527668286f : (s1 ? 0.139942324101744f : (s2 ? 0.063427417320019f : 0.009695946122598f));
    const double c1 = s0 ? 0.498799810682272f : (s1 ? 0.495635523139337f : (s2 ? 0.695956496475118f : 0.930734667215156f));
    const double c0 = s0 ? 0.693203116424741f : (s1 ? 0.692140569840976f : (s2 ? 0.514272634594009f : 0.168037164329057f));
    return ((c3 * x + c2) * x + c1) * x + c0;
}
__device__ __forceinline__ double g_la(double x, double y) {
    const bool xl = x < y;
    const double lo = xl ? x : y, hi = xl ? y : x;
    const double d = hi - lo;
    const double r = g_poly(d) + lo;
    return (lo == CPG_NI || d >= 7.5) ? hi : r;
}
// impl/stateMachine.c:333-343 and :322-331
__device__ __forceinline__ double g_log_gauss(double x, double mu, double sigma) {
    if (sigma == 0.0) return CPG_NI;
    const double a = (x - mu) / sigma;
    return -0.91893853320467267 - log(sigma) + (-0.5 * a * a);
}
__device__ __forceinline__ double g_log_inv_gauss(double x, double mu, double lambda) {
    const double a = (x - mu) / mu;
    return (log(lambda) - 1.8378770664093453 - 3 * log(x) - lambda * a * a / x) / 2;
}

// the same two densities with the logarithms that depend only on the k-mer (log sigma, log lambda) or only on the event
// (log x) handed in: they are computed once per column / per event by k_prep_generic instead of once per cell
__device__ __forceinline__ double g_log_gauss_c(double x, double mu, double sigma, double logSigma) {
    if (sigma == 0.0) return CPG_NI;
    const double a = (x - mu) / sigma;
    return -0.91893853320467267 - logSigma + (-0.5 * a * a);
}
__device__ __forceinline__ double g_log_inv_gauss_c(double x, double mu, double lambda, double logLambda, double logX) {
    const double a = (x - mu) / mu;
    return (logLambda - 1.8378770664093453 - 3 * logX - lambda * a * a / x) / 2;
}

// Per-column records of the FP64 kernel (everything of a cell's emissions and transitions that depends only on its
// column: k-mer index, the logarithms of the scaled model's deviations, the vanilla machine's seven log transitions).
//   threeState / fourState (14): k | gap-X emission | match: mu sd log(sd) nu tau log(tau) | gap Y: the same six
//   vanilla (21): k | skip bin | log a_mx | log a_xx | log a_mm | log a_xm | log a_ym | log a_my | log a_yy |
//                 match: mu sd log(sd) nu lambda log(lambda) | gap Y: the same six
// (planes, so that neighbouring lanes read neighbouring doubles; a cell never touches the 4096-row model tables)
//   threeStateHdp (1): the row of the HDP's density table the column's k-mer reads
//   echelon (41): log a_mx | log(1 - a_mx) | log a_xx | log(1 - a_xx) | run-off mask (bit n: the character 6 n past the
//                 pointer is a nucleotide) | gap Y of k-mer i+1: mu sd log(sd) nu lambda log(lambda) | the same six of the
//                 scaled match rows of k-mers i+1 .. i+5; per event: log noise and log(duration / 0.00332)
// The values are the reference's own expressions (impl/stateMachine.c:322-343, 388-427, 631-651), evaluated once.
__host__ __device__ inline int gen_ncol(int sm) { return sm == 4 ? 21 : (sm == 7 ? 1 : (sm == 5 ? 41 : 14)); }

__global__ void k_prep_generic(const Item *items, const long long *ref_off, const char *ref, const ModelTables *models,
                               const double *scale, GenParams G, double *colp, double *rowp, long long row2, const double *events,
                               const long long *ev_src_off) {
    const int i = blockIdx.x;
    const Item it = items[i];
    const ModelTables mt = models[it.model_id];
    const char *r = ref + ref_off[i];
    const int refLen = (int) (ref_off[i + 1] - ref_off[i]);
    const int sm = G.sm, NC = gen_ncol(sm);
    const long long CS = it.lX + 2;
    double *cq = colp + (long long) NC * it.xp_off;
    const bool scaled = scale != nullptr;
    double sc = 1, sh = 0, var = 1, scsd = 1, varsd = 1;
    if (scaled) { const double *s5 = scale + 5 * i; sc = s5[0]; sh = s5[1]; var = s5[2]; scsd = s5[3]; varsd = s5[4]; }
    auto kmerAt = [&](int p) -> int {
        int v = 0;
        for (int j = 0; j < 6; j++) { const int q = p + j; const int b = base_code((q >= 0 && q < refLen) ? r[q] : 'n'); if (b < 0) return -1; v = v * 4 + b; }
        return v;
    };
    // scaled MATCH row as emissions_signal_scaleModel leaves it; the gap-Y row is never scaled
    auto matchRow = [&](int k, double &mu, double &sd, double &nu, double &tau, double &lam) {
        if (k < 0) { mu = sd = nu = tau = lam = 0.0; return; }
        const double *m = mt.match + 1 + 5 * k;
        mu = m[0]; sd = m[1]; nu = m[2]; tau = m[3]; lam = m[4];
        if (scaled) { mu = mu * sc + sh; sd = sd * var; nu = nu * scsd; lam = lam * varsd; tau = sqrt(pow(nu, 3.0) / lam); }
    };
    for (int x = blockIdx.y * blockDim.x + threadIdx.x; x <= it.lX; x += gridDim.y * blockDim.x) {
        if (sm == 7) {
            const int k = kmerAt(x >= 1 ? x - 1 : 0);
            cq[x] = (double) (k >= 0 ? mt.hdp_kmer[k] : -1);
        } else if (sm == 5) {
            const int p = x >= 2 ? x - 2 : 0;
            double mu0, mu1, sd, nu, tau, lam;
            matchRow(kmerAt(p), mu0, sd, nu, tau, lam);
            matchRow(kmerAt(p + 1), mu1, sd, nu, tau, lam);
            long long bin = (long long) (fabs(mu1 - mu0) / 0.5);
            bin = bin >= 30 ? 29 : bin;
            const double a_mx = mt.gapx[bin], a_xx = mt.gapx[bin + 30];
            cq[x] = log(a_mx); cq[CS + x] = log(1 - a_mx); cq[2 * CS + x] = log(a_xx); cq[3 * CS + x] = log(1 - a_xx);
            int mask = 0;
            for (int n = 1; n < 6; n++) {
                const int q = p + 6 * n;
                const char last = (q >= 0 && q < refLen) ? r[q] : 'n';
                if (last >= 'A' && last <= 'Z') mask |= 1 << n;
            }
            cq[4 * CS + x] = (double) mask;
            const int ky = kmerAt(p + 1);
            double y[5] = { 0.0, 0.0, 0.0, 0.0, 0.0 };
            if (ky >= 0) for (int j = 0; j < 5; j++) y[j] = mt.gapy[1 + 5 * ky + j];
            cq[5 * CS + x] = y[0]; cq[6 * CS + x] = y[1]; cq[7 * CS + x] = log(y[1]); cq[8 * CS + x] = y[2]; cq[9 * CS + x] = y[4];
            cq[10 * CS + x] = log(y[4]);
            for (int n = 1; n < 6; n++) {
                double mu, s2, n2, t2, l2;
                matchRow(kmerAt(p + n), mu, s2, n2, t2, l2);
                double *o = cq + (long long) (11 + 6 * (n - 1)) * CS + x;
                o[0] = mu; o[CS] = s2; o[2 * CS] = log(s2); o[3 * CS] = n2; o[4 * CS] = l2; o[5 * CS] = log(l2);
            }
        } else if (sm == 4) {
            const int p = x >= 2 ? x - 2 : 0;
            double mu0, mu1, sd, nu, tau, lam;
            matchRow(kmerAt(p), mu0, sd, nu, tau, lam);
            const int k = kmerAt(p + 1);
            matchRow(k, mu1, sd, nu, tau, lam);
            long long bin = (long long) (fabs(mu1 - mu0) / 0.5);
            bin = bin >= 30 ? 29 : bin;
            const double a_mx = mt.gapx[bin];
            const double a_my = (1 - a_mx) * G.van[0];
            const double a_mm = 1.0f - a_my - a_mx;
            const double a_yy = G.van[1];
            const double a_ym = 1.0f - a_yy;
            const double a_xx = mt.gapx[bin + 30];
            const double a_xm = 1.0f - a_xx;
            cq[x] = (double) k; cq[CS + x] = (double) bin;
            cq[2 * CS + x] = log(a_mx); cq[3 * CS + x] = log(a_xx); cq[4 * CS + x] = log(a_mm); cq[5 * CS + x] = log(a_xm);
            cq[6 * CS + x] = log(a_ym); cq[7 * CS + x] = log(a_my); cq[8 * CS + x] = log(a_yy);
            cq[9 * CS + x] = mu1; cq[10 * CS + x] = sd; cq[11 * CS + x] = log(sd); cq[12 * CS + x] = nu; cq[13 * CS + x] = lam;
            cq[14 * CS + x] = log(lam);
            double y[5] = { 0.0, 0.0, 0.0, 0.0, 0.0 };
            if (k >= 0) for (int j = 0; j < 5; j++) y[j] = mt.gapy[1 + 5 * k + j];
            cq[15 * CS + x] = y[0]; cq[16 * CS + x] = y[1]; cq[17 * CS + x] = log(y[1]); cq[18 * CS + x] = y[2]; cq[19 * CS + x] = y[4];
            cq[20 * CS + x] = log(y[4]);
        } else {
            const int k = x >= 1 ? kmerAt(x - 1) : -1;
            double mu, sd, nu, tau, lam;
            matchRow(k, mu, sd, nu, tau, lam);
            cq[x] = (double) k; cq[CS + x] = k < 0 ? CPG_NI : mt.gapx[k];
            cq[2 * CS + x] = mu; cq[3 * CS + x] = sd; cq[4 * CS + x] = log(sd); cq[5 * CS + x] = nu; cq[6 * CS + x] = tau; cq[7 * CS + x] = log(tau);
            double y[5] = { 0.0, 0.0, 0.0, 0.0, 0.0 };
            if (k >= 0) for (int j = 0; j < 5; j++) y[j] = mt.gapy[1 + 5 * k + j];
            cq[8 * CS + x] = y[0]; cq[9 * CS + x] = y[1]; cq[10 * CS + x] = log(y[1]); cq[11 * CS + x] = y[2]; cq[12 * CS + x] = y[3];
            cq[13 * CS + x] = log(y[3]);
        }
    }
    if ((sm == 4 || sm == 5) && rowp != nullptr) {
        const double *evs = events + 3 * ev_src_off[i];
        double *rq = rowp + it.ev_off;
        for (int y = blockIdx.y * blockDim.x + threadIdx.x; y <= it.lY; y += gridDim.y * blockDim.x) {
            rq[y] = y >= 1 ? log(evs[3 * (y - 1) + 1]) : 0.0;
            if (sm == 5) rq[row2 + y] = y >= 1 ? log(evs[3 * (y - 1) + 2] / 0.00332005312085) : 0.0;
        }
    }
}

template <int SM> struct GenTraits;
template <> struct GenTraits<6> { static constexpr int S = 4, NC = 14; };
template <> struct GenTraits<5> { static constexpr int S = 7, NC = 0; };
template <> struct GenTraits<2> { static constexpr int S = 3, NC = 14; };
template <> struct GenTraits<7> { static constexpr int S = 3, NC = 1; };
template <> struct GenTraits<4> { static constexpr int S = 3, NC = 21; };

template <int SM>
__global__ void __launch_bounds__(32, CPG_MINB) k_align_generic(const KernelArgsG A) {
    constexpr int S = GenTraits<SM>::S, NC = GenTraits<SM>::NC;
    extern __shared__ __align__(16) unsigned char smraw[];
    double *ring = reinterpret_cast<double *>(smraw);           // 4 buffers x S planes x N
    const int N = A.ringN, NM = N - 1;
    const int lane = threadIdx.x;
    const DevParams &P = A.P;
    const double NI = CPG_NI;
    double *rows = A.scratch + (long long) blockIdx.x * A.scratch_stride;
    const int R = A.ring_rows;
    auto buf = [&](int b) -> double * { return ring + (size_t) b * S * N; };

    for (;;) {
        int qi = 0;
        if (lane == 0) qi = atomicAdd(A.queue, 1);
        qi = __shfl_sync(CP_FULL, qi, 0);
        if (qi >= A.n_items) break;
        const int itemIdx = A.order[qi];
        const Item it = A.items[itemIdx];
        const int lX = it.lX, lY = it.lY, D = lX + lY;
        const int planFlags = A.flags[itemIdx];
        if (D == 0 || (planFlags & 16) != 0) {                  // nothing to align, or a diagonal of the band is empty
            if (lane == 0) { ItemOut &o = A.out[itemIdx]; o.n_pairs = 0; o.status = D == 0 ? 0 : 6; o.total_logprob = 0.0; o.n_tracebacks = 0; }
            continue;
        }
        const double *evs = A.events + 3 * A.ev_src_off[itemIdx];
        const ModelTables mt = A.models[it.model_id];
        const int2 *bandp = A.bands + A.band_off[itemIdx];
        const long long CS = lX + 2, CSg = CS;                 // CSg: the stride under a name cellGen does not shadow
        const double *colq = A.colp + (long long) gen_ncol(SM) * it.xp_off;
        const double *rowq = (SM == 4 || SM == 5) ? A.rowp + it.ev_off : nullptr;
        const int *tbp = A.tbs + it.pad1;
        int *pairs = A.pairs + 3 * it.pair_off;
        double *dbgTot = (A.totals != nullptr && it.tot_off >= 0) ? A.totals + it.tot_off : nullptr;
        int nPairs = 0, status = 0, nTb = 0;
        double lastTotal = 0.0, lik = 0.0;

        auto bandOf = [&](int d, int &l, int &h) { if (d >= 0 && d <= D) { const int2 b = bandp[d]; l = b.x; h = b.y; } else { l = 0; h = -1; } };
        // dir_proc_density of the k-mer's distribution (impl/hdp.c:2577-2599) through grid_spline_interp
        // (impl/hdp_math_utils.c:471-495); the grid is linspace(start, stop, n) (:497-510)
        auto hdpDensity = [&](int t, double q) -> double {
            if (t < 0) { status |= 8; return 0.0; }
            const double *yv = mt.hdp_y + (long long) t * mt.hdp_len, *sl = mt.hdp_slope + (long long) t * mt.hdp_len;
            const int n = mt.hdp_len - 1;
            double r;
            if (q <= mt.hdp_x0) r = yv[0] - sl[0] * (mt.hdp_x0 - q);
            else if (q >= mt.hdp_xn) r = yv[n] + sl[n] * (q - mt.hdp_xn);
            else {
                const double dx = mt.hdp_x1 - mt.hdp_x0;
                int il = (int) ((q - mt.hdp_x0) / dx);
                il = il > n - 1 ? n - 1 : il;
                const double xl = __dadd_rn(mt.hdp_x0, __dmul_rn((double) il, mt.hdp_step));     // linspace's own rounding, no FMA
                const double dy = yv[il + 1] - yv[il];
                const double a = sl[il] * dx - dy, b = dy - sl[il + 1] * dx;
                const double tl = (q - xl) / dx, tr = 1.0 - tl;
                r = tr * yv[il] + tl * yv[il + 1] + tl * tr * (a * tr + b * tl);
            }
            return r > 0.0 ? r : 0.0;
        };

        // Everything a cell reads from global memory -- its column's record, its event, (vanilla) log noise -- as one
        // value: the sweeps load it ONCE per cell and one chunk AHEAD of its use (the first uses of these loads were
        // half of all stall samples when they sat inside the cell body)
        struct CellIn { double c[NC > 0 ? NC : 1]; double em, en, edur, lx; };
        auto loadIn = [&](int x, int y, bool ok) -> CellIn {
            CellIn in;
#pragma unroll
            for (int j = 0; j < (NC > 0 ? NC : 1); j++) in.c[j] = 0.0;
            in.em = NI; in.en = 0.0; in.edur = 0.0; in.lx = 0.0;
            if (ok) {
                if (NC > 0) {
                    const double *cq = colq + x;
#pragma unroll
                    for (int j = 0; j < NC; j++) in.c[j] = __ldg(cq + j * CS);
                }
                if (y >= 1) { const double *e = evs + 3 * (y - 1); in.em = __ldg(e); in.en = __ldg(e + 1); in.edur = __ldg(e + 2); }
                if (SM == 4) in.lx = __ldg(rowq + (y >= 0 ? y : 0));
            }
            return in;
        };

        // One cell in the reference's transition order.  FWD: cur[to] (+)= nb[from] + (eP + tP) (pull); BWD: the same
        // statement with the roles swapped (push), impl/pairwiseAligner.c:365-383.  lo / mi / up: the neighbours
        // (x-1, y), (x-1, y-1), (x, y-1), has*: whether they exist (inside the band of a live diagonal).
        // Expectation mode (diagonalCalculation_Expectations, impl/pairwiseAligner.c:841-863): cur = the cell's BACKWARD values,
        // the neighbours' FORWARD values; every transition's posterior p = exp(F + B + (eP + tP) - total) goes to
        //   threeState: the 3 x 3 transition sums and, into gap X, the k-mer's skip count (:425-443)
        //   vanilla:    the skip bin's beta (match -> gap X) or alpha (gap X -> gap X) count (:478-498)
        //   HDP:        the transition sums, and a match with p >= threshold is an event-to-k-mer ASSIGNMENT (:445-476)
        double accT[9] = { 0, 0, 0, 0, 0, 0, 0, 0, 0 };
        double expTotal = 0.0;
        int expK = -1, expBin = 0, asgMask = 0;
        auto expUp = [&](int from, int to, double p) {
            if (SM == 2) { accT[from * 3 + to] += p; if (to == 1 && expK >= 0) atomicAdd(A.expect + 9 + expK, p); }
            else if (SM == 7) { accT[from * 3 + to] += p; if (to == 0 && p >= A.G.threshold) asgMask |= 1 << from; }
            else if (SM == 4) {
                if (from == 0 && to == 1) atomicAdd(A.expect + expBin, p);
                if (from == 1 && to == 1) atomicAdd(A.expect + expBin + 30, p);
            }
        };
        auto cellGen = [&](int fwd, int x, int y, const CellIn &in, double *cur, double *lo_, bool hasLo, double *mi_, bool hasMi,
                           double *up_, bool hasUp) {
#define TRG(nb, from, to, eptp) do { if (fwd == 1) cur[to] = g_la(cur[to], nb[from] + (eptp)); \
                                     else if (fwd == 0) nb[from] = g_la(nb[from], cur[to] + (eptp)); \
                                     else expUp(from, to, exp(nb[from] + cur[to] + (eptp) - expTotal)); } while (0)
            const double em = in.em, en = in.en, edur = in.edur;
            const double *cq = in.c;
            constexpr int CS = 1;                                 // the record's planes are consecutive in CellIn
            if (SM == 2) {
                // stateMachine3_cellCalculate (impl/stateMachine.c:1305-1334); sequence_getKmer: index < 0 reads "n"
                const int k = (int) cq[0];
                const double *t = A.G.t3;
                expK = k;
                if (hasLo) {
                    const double eP = cq[CS];
                    TRG(lo_, 0, 1, eP + t[3]); TRG(lo_, 1, 1, eP + t[5]); TRG(lo_, 2, 1, eP + t[7]);
                }
                if (hasMi) {
                    const double eP = g_log_gauss_c(em, cq[2 * CS], cq[3 * CS], cq[4 * CS]) + g_log_gauss_c(en, cq[5 * CS], cq[6 * CS], cq[7 * CS]);
                    TRG(mi_, 0, 0, eP + t[0]); TRG(mi_, 1, 0, eP + t[1]); TRG(mi_, 2, 0, eP + t[2]);
                }
                if (hasUp) {
                    const double eP = g_log_gauss_c(em, cq[8 * CS], cq[9 * CS], cq[10 * CS]) + g_log_gauss_c(en, cq[11 * CS], cq[12 * CS], cq[13 * CS]);
                    TRG(up_, 0, 2, eP + t[4]); TRG(up_, 2, 2, eP + t[6]);
                }
            } else if (SM == 7) {
                // stateMachine3HDP_cellCalculate (impl/stateMachine.c:1336-1366); sequence_getKmer3: index < 0 reads k-mer 0
                const int k = (int) cq[0];                    // the density-table row of k-mer max(x - 1, 0)
                const double *t = A.G.t3;
                if (hasLo) {
                    const double eP = -2.3025850929940455;
                    TRG(lo_, 0, 1, eP + t[3]); TRG(lo_, 1, 1, eP + t[5]); TRG(lo_, 2, 1, eP + t[7]);
                }
                if (hasMi) {
                    const double eP = hdpDensity(k, em);
                    TRG(mi_, 0, 0, eP + t[0]); TRG(mi_, 1, 0, eP + t[1]); TRG(mi_, 2, 0, eP + t[2]);
                }
                if (hasUp) {
                    const double eP = hdpDensity(k, em);
                    TRG(up_, 0, 2, eP + t[4]); TRG(up_, 2, 2, eP + t[6]);
                }
            } else if (SM == 4) {
                // stateMachine3Vanilla_cellCalculate (impl/stateMachine.c:1368-1409); sequence_getKmer2: the pointer to the
                // PREVIOUS k-mer, clamped at 0; the skip bin of (k-mer i, k-mer i+1) on the scaled match table -- all of
                // it per column (k_prep_generic)
                expBin = (int) cq[CS];
                const double lx = in.lx;
                if (hasLo) { TRG(lo_, 0, 1, 0 + cq[2 * CS]); TRG(lo_, 1, 1, 0 + cq[3 * CS]); }
                if (hasMi) {
                    const double eP = g_log_gauss_c(em, cq[9 * CS], cq[10 * CS], cq[11 * CS]) + g_log_inv_gauss_c(en, cq[12 * CS], cq[13 * CS], cq[14 * CS], lx);
                    TRG(mi_, 0, 0, eP + cq[4 * CS]); TRG(mi_, 1, 0, eP + cq[5 * CS]); TRG(mi_, 2, 0, eP + cq[6 * CS]);
                }
                if (hasUp) {
                    const double eP = g_log_gauss_c(em, cq[15 * CS], cq[16 * CS], cq[17 * CS]) + g_log_inv_gauss_c(en, cq[18 * CS], cq[19 * CS], cq[20 * CS], lx);
                    TRG(up_, 0, 2, eP + cq[7 * CS]); TRG(up_, 2, 2, eP + cq[8 * CS]);
                }
            } else if (SM == 6) {
                const double *t = A.G.t4;
                if (hasLo) {
                    const double eP = cq[CS];
                    TRG(lo_, 0, 1, eP + t[4]); TRG(lo_, 1, 1, eP + t[5]); TRG(lo_, 0, 3, eP + t[8]); TRG(lo_, 3, 3, eP + t[9]); TRG(lo_, 2, 3, eP + t[10]);
                }
                if (hasMi) {
                    const double eP = g_log_gauss_c(em, cq[2 * CS], cq[3 * CS], cq[4 * CS]) + g_log_gauss_c(en, cq[5 * CS], cq[6 * CS], cq[7 * CS]);
                    TRG(mi_, 0, 0, eP + t[0]); TRG(mi_, 1, 0, eP + t[1]); TRG(mi_, 2, 0, eP + t[2]); TRG(mi_, 3, 0, eP + t[3]);
                }
                if (hasUp) {
                    const double eP = g_log_gauss_c(em, cq[8 * CS], cq[9 * CS], cq[10 * CS]) + g_log_gauss_c(en, cq[11 * CS], cq[12 * CS], cq[13 * CS]);
                    TRG(up_, 0, 2, eP + t[6]); TRG(up_, 2, 2, eP + t[7]);
                }
            } else {
                // echelon: sequence_getKmer2 pointer i (clamped at 0), skip bin of (k-mer i, k-mer i+1) on the scaled match
                // table, n = 1 .. 5 k-mers per event (impl/stateMachine.c:1411-1460); everything that depends on the column
                // or on the event alone comes from k_prep_generic
                const double *eq = colq + x;
                const double la_mx = eq[0], la_mh = eq[CSg], la_xx = eq[2 * CSg], la_xh = eq[3 * CSg];
                if (hasLo) {
#pragma unroll
                    for (int n = 1; n < 6; n++) TRG(lo_, n, 6, 0 + la_mx);
                    TRG(lo_, 6, 6, 0 + la_xx);
                }
                const double lfact[6] = { 0.0, 0.0, 0.69314718056, 1.79175946923, 3.17805383035, 4.78749174278 };
                const double lambda = edur / 0.00332005312085;
                const double llam = rowq[A.row2 + (y >= 0 ? y : 0)];
                const double lx = rowq[y >= 0 ? y : 0];
                if (hasMi) {
                    // emissions_signal_multipleKmerMatchProb (:530-549): the fold starts from 0.0 and the run-off test
                    // looks at the single character 6 n past the pointer
                    const int mask = (int) eq[4 * CSg];
                    const double logn[6] = { 0.0, 0.0, 0.69314718055994529, 1.0986122886681098, 1.3862943611198906, 1.6094379124341003 };
                    double mk[6];
                    double fold = 0.0;
#pragma unroll
                    for (int n = 1; n < 6; n++) {
                        const double *m = eq + (long long) (11 + 6 * (n - 1)) * CSg;
                        fold = g_la(fold, g_log_gauss_c(em, m[0], m[CSg], m[2 * CSg]) + g_log_inv_gauss_c(en, m[3 * CSg], m[4 * CSg], m[5 * CSg], lx));
                        mk[n] = (mask >> n) & 1 ? fold - logn[n] : NI;
                    }
#pragma unroll
                    for (int n = 1; n < 6; n++) {
                        const double dp = (n + 1) * 0.1397619423751586 + n * llam - lfact[n] - 2 * lambda;
#pragma unroll
                        for (int from = 0; from < 6; from++) TRG(mi_, from, n, mk[n] + (la_mh + dp));
                    }
#pragma unroll
                    for (int n = 1; n < 6; n++) {
                        const double dp = (n + 1) * 0.1397619423751586 + n * llam - lfact[n] - 2 * lambda;
                        TRG(mi_, 6, n, mk[n] + (la_xh + dp));
                    }
                }
                if (hasUp) {
                    const double eP = g_log_gauss_c(em, eq[5 * CSg], eq[6 * CSg], eq[7 * CSg]) + g_log_inv_gauss_c(en, eq[8 * CSg], eq[9 * CSg], eq[10 * CSg], lx);
                    const double dp0 = 1 * 0.1397619423751586 + 0 * llam - lfact[0] - 2 * lambda;
#pragma unroll
                    for (int n = 1; n < 6; n++) TRG(up_, n, 0, eP + (la_mh + dp0));
                }
            }
#undef TRG
        };
        // note: multipleKmerMatchProb returns -inf from inside its loop as soon as the run-off test fails, and the test
        // does not depend on the loop variable: for a given n the value is either the full fold or -inf (as above)

        auto stateVector = [&](int which, double *v) {       // 0 start, 1 ragged start, 2 end, 3 ragged end
            if (SM == 5) {
                for (int st = 0; st < 7; st++) v[st] = NI;
                if (which == 0) v[1] = 0; else if (which == 1) v[6] = 0;
                else { for (int st = 0; st < 6; st++) v[st] = 0.79015888282447311; v[6] = 0.19652425498269727; }   // not logs (:1617-1619)
            } else if (SM == 2 || SM == 7 || SM == 4) {
                // impl/stateMachine.c:1168-1235
                if (which == 0) { v[0] = 0; v[1] = v[2] = NI; }
                else if (which == 1) { v[0] = NI; v[1] = 0; v[2] = 0; }
                else if (SM != 4) {
                    const double *t = A.G.t3;
                    if (which == 2) { v[0] = t[0]; v[1] = t[1]; v[2] = t[2]; } else { v[0] = (t[3] + t[4]) / 2.0; v[1] = t[5]; v[2] = t[6]; }
                } else {
                    const double *e = A.G.van;
                    if (which == 2) { v[0] = e[2]; v[1] = e[3]; v[2] = e[4]; } else { v[0] = (e[3] + e[4]) / 2.0; v[1] = e[3]; v[2] = e[4]; }
                }
            } else {
                const double *t = A.G.t4;
                if (which == 0) { v[0] = 0; v[1] = v[2] = v[3] = NI; }
                else if (which == 1) { v[0] = v[1] = NI; v[2] = 0; v[3] = 0; }
                else if (which == 2) { v[0] = t[0]; v[1] = t[1]; v[2] = t[2]; v[3] = t[3]; }
                else { v[0] = v[1] = v[2] = t[8]; v[3] = t[9]; }
            }
        };
        auto clearBuf = [&](int b, int l, int h) {           // -inf over positions l .. h of every plane
            double *p = buf(b);
            for (int x = l + lane; x <= h; x += 32)
#pragma unroll
                for (int st = 0; st < S; st++) p[st * N + (x & NM)] = NI;
        };
        auto rowPtr = [&](int d, int x) -> double * { return rows + ((long long) (d % R) * N + (x & NM)) * S; };

        // ---- diagonal 0 -------------------------------------------------------------------------------------------
        int f0 = 0, f1 = 1, f2 = 2;                            // ring buffers of diagonals d, d-1, d-2 (forward)
        int lo = 0, hi = 0, lo1 = 0, hi1 = -1, lo2 = 0, hi2 = -1;   // band of d, d-1, d-2
        if (lane == 0) {
            double v[S];
            stateVector((it.flags & 1) ? 1 : 0, v);
            for (int st = 0; st < S; st++) { buf(f0)[st * N] = v[st]; rowPtr(0, 0)[st] = v[st]; }
        }
        __syncwarp();
        int dcur = 0, tracedBackTo = 0;

        while (tracedBackTo < D) {
            const int Dt = tbp[nTb];
            const bool atEnd = Dt == D;
            const int tbf = Dt - (atEnd ? 0 : P.tbDiags + 1);
            // =============================== forward ========================================================
            int2 bN = make_int2(0, -1), bNN = make_int2(0, -1);    // bands of the next two diagonals, fetched two diagonals ahead
            if (dcur + 1 <= Dt) bN = bandp[dcur + 1];
            if (dcur + 2 <= Dt) bNN = bandp[dcur + 2];
            CellIn nxt = loadIn(bN.x + lane, dcur + 1 - (bN.x + lane), dcur + 1 <= Dt && bN.x + lane <= bN.y);
            for (int d = dcur + 1; d <= Dt; d++) {
                { const int t = f2; f2 = f1; f1 = f0; f0 = t; }
                lo2 = lo1; hi2 = hi1; lo1 = lo; hi1 = hi;
                lo = bN.x; hi = bN.y;
                bN = bNN;
                if (d + 2 <= Dt) bNN = bandp[d + 2];
                double *F0 = buf(f0), *F1 = buf(f1), *F2 = buf(f2);
                for (int xb = lo; xb <= hi; xb += 32) {
                    const int x = xb + lane;
                    const CellIn in = nxt;
                    // the next chunk of this diagonal, or the first one of the next diagonal
                    if (xb + 32 <= hi) nxt = loadIn(x + 32, d - (x + 32), x + 32 <= hi);
                    else nxt = loadIn(bN.x + lane, d + 1 - (bN.x + lane), d + 1 <= Dt && bN.x + lane <= bN.y);
                    if (x <= hi) {
                        double cur[S], nl[S], nm[S], nu_[S];
                        const bool hasLo = x - 1 >= lo1 && x - 1 <= hi1, hasMi = x - 1 >= lo2 && x - 1 <= hi2, hasUp = x >= lo1 && x <= hi1;
#pragma unroll
                        for (int st = 0; st < S; st++) {
                            cur[st] = NI;
                            nl[st] = hasLo ? F1[st * N + ((x - 1) & NM)] : NI;
                            nm[st] = hasMi ? F2[st * N + ((x - 1) & NM)] : NI;
                            nu_[st] = hasUp ? F1[st * N + (x & NM)] : NI;
                        }
                        cellGen(1, x, d - x, in, cur, nl, hasLo, nm, hasMi, nu_, hasUp);
                        double *rp = rowPtr(d, x);
#pragma unroll
                        for (int st = 0; st < S; st++) { F0[st * N + (x & NM)] = cur[st]; __stcs(rp + st, cur[st]); }
                    }
                }
                __syncwarp();
            }
            dcur = Dt;

            // =============================== traceback ======================================================
            nTb++;
            {
                int b0 = 0, b1 = 1, b2 = 2, bp = 3;            // ring buffers of B of d, d-1, d-2 and d+1
                int blo = lo, bhi = hi;                        // band of d
                double endv[S];
                stateVector((atEnd && (it.flags & 2)) ? 3 : 2, endv);
                for (int x = blo + lane; x <= bhi; x += 32)
#pragma unroll
                    for (int st = 0; st < S; st++) buf(b0)[st * N + (x & NM)] = endv[st];
                int l1 = 0, h1 = -1, l2 = 0, h2 = -1, lp = 0, hp = -1;   // band of d-1, d-2, d+1
                bandOf(Dt - 1, l1, h1);
                if (Dt > tracedBackTo + 1) clearBuf(b1, l1 - 1, h1 + 1);
                __syncwarp();
                double total = NI;
                long long count = 0;
                CellIn nxtB = loadIn(blo + lane, Dt - (blo + lane), blo + lane <= bhi);
                for (int d = Dt; d > tracedBackTo; d--) {
                    bandOf(d - 2, l2, h2);
                    const bool liveMi = d > tracedBackTo + 2, sweepB = d > tracedBackTo + 1;
                    if (liveMi) clearBuf(b2, l2 - 1, h2 + 1);
                    __syncwarp();
                    double *B0 = buf(b0), *B1 = buf(b1), *B2 = buf(b2), *Bp = buf(bp);
                    // the posterior of this diagonal reads the forward match value of its cells: the first chunk's is
                    // requested here, before the sweep, the following ones a chunk ahead
                    double fPre = NI;
                    if (SM != 5 && P.mode != 1 && d <= tbf && blo + lane <= bhi) fPre = __ldcs(rowPtr(d, blo + lane));
                    if (sweepB) {
                        for (int xb = blo; xb <= bhi; xb += 32) {
                            const int x = xb + lane;
                            const bool act = x <= bhi;
                            const CellIn in = nxtB;
                            // the next chunk of this diagonal, or the first one of diagonal d - 1 (whose band is l1 .. h1)
                            if (xb + 32 <= bhi) nxtB = loadIn(x + 32, d - (x + 32), x + 32 <= bhi);
                            else nxtB = loadIn(l1 + lane, d - 1 - (l1 + lane), d - 1 > tracedBackTo && l1 + lane <= h1);
                            double cur[S], nb[S];
                            if (act)
#pragma unroll
                                for (int st = 0; st < S; st++) cur[st] = B0[st * N + (x & NM)];
                            const bool hasLo = act && x - 1 >= l1 && x - 1 <= h1, hasMi = act && liveMi && x - 1 >= l2 && x - 1 <= h2,
                                       hasUp = act && x >= l1 && x <= h1;
                            // the reference's order inside a cell is lower, middle, upper; the three go to different cells, so
                            // only the order in which ONE cell receives matters: from d+2 (earlier), then as "upper" of the
                            // cell above it, then as "lower" of the cell to its right (ascending x - y)
                            if (hasMi) {
#pragma unroll
                                for (int st = 0; st < S; st++) nb[st] = B2[st * N + ((x - 1) & NM)];
                                cellGen(0, x, d - x, in, cur, nullptr, false, nb, true, nullptr, false);
#pragma unroll
                                for (int st = 0; st < S; st++) B2[st * N + ((x - 1) & NM)] = nb[st];
                            }
                            if (hasUp) {
#pragma unroll
                                for (int st = 0; st < S; st++) nb[st] = B1[st * N + (x & NM)];
                                cellGen(0, x, d - x, in, cur, nullptr, false, nullptr, false, nb, true);
#pragma unroll
                                for (int st = 0; st < S; st++) B1[st * N + (x & NM)] = nb[st];
                            }
                            __syncwarp();
                            if (hasLo) {
#pragma unroll
                                for (int st = 0; st < S; st++) nb[st] = B1[st * N + ((x - 1) & NM)];
                                cellGen(0, x, d - x, in, cur, nb, true, nullptr, false, nullptr, false);
#pragma unroll
                                for (int st = 0; st < S; st++) B1[st * N + ((x - 1) & NM)] = nb[st];
                            }
                            __syncwarp();
                        }
                    }
                    if (d <= tbf) {
                        const bool doTotal = P.mode == 2 ? (d == Dt) : (count % P.totalEvery == 0);
                        count++;
                        if (doTotal) {
                            // diagonalCalculationTotalProbability (impl/pairwiseAligner.c:736-754): serial left folds
                            auto diagDot = [&](auto getA, const double *Bb, int l, int h) -> double {
                                double tot = NI;
                                for (int xb = l; xb <= h; xb += 32) {
                                    const int x = xb + lane;
                                    double c = NI;
                                    if (x <= h) {
                                        double a[S];
                                        getA(x, a);
                                        c = a[0] + Bb[(x & NM)];
#pragma unroll
                                        for (int st = 1; st < S; st++) c = g_la(c, a[st] + Bb[st * N + (x & NM)]);
                                    }
                                    const int cnt = min(32, h - xb + 1);
                                    for (int j = 0; j < cnt; j++) tot = g_la(tot, __shfl_sync(CP_FULL, c, j));
                                }
                                return tot;
                            };
                            double tot = diagDot([&](int x, double *a) { const double *rp = rowPtr(d, x);
#pragma unroll
                                                     for (int st = 0; st < S; st++) a[st] = rp[st]; }, B0, blo, bhi);
                            double t2 = CUDART_NAN;
                            if (d < Dt && d - 1 >= 0) {
                                // match-only forward step from F[d-1] into a -inf clone shaped like B[d+1], dotted with B[d+1]
                                int lm, hm;
                                bandOf(d - 1, lm, hm);
                                t2 = diagDot([&](int x, double *a) {
                                    double nm[S];
                                    const bool hasMi = x - 1 >= lm && x - 1 <= hm;
#pragma unroll
                                    for (int st = 0; st < S; st++) { a[st] = NI; nm[st] = NI; }
                                    if (hasMi) {
                                        const double *rp = rowPtr(d - 1, x - 1);
#pragma unroll
                                        for (int st = 0; st < S; st++) nm[st] = rp[st];
                                    }
                                    cellGen(1, x, d + 1 - x, loadIn(x, d + 1 - x, true), a, nullptr, false, nm, hasMi, nullptr, false);
                                }, Bp, lp, hp);
                                tot = g_la(tot, t2);
                            }
                            total = tot;
                            if (!(tot > -1e300)) status |= 2;
                            if (dbgTot != nullptr && lane == 0) { dbgTot[(D + 1) + d] = tot; dbgTot[2 * (D + 1) + d] = t2; }
                        }
                        if (d == D) lastTotal = total;
                        if (dbgTot != nullptr && lane == 0) dbgTot[d] = total;
                        if (P.mode == 1) {
                            // expectations of this diagonal; a diagonal whose total is not finite contributes nothing
                            if (S == 3 && total > -1e300 && total < 1e300) {
                                lik += total;
                                expTotal = total;
                                for (int xb = blo; xb <= bhi; xb += 32) {
                                    const int x = xb + lane, y = d - x;
                                    asgMask = 0;
                                    if (x <= bhi) {
                                        double cur[S], nl[S], nm[S], nu_[S];
                                        // the previous traceback freed the forward diagonals below the one it stopped at
                                        // (impl/pairwiseAligner.c:983-990): on the first diagonal above it, "middle" is NULL
                                        const bool hasLo = x - 1 >= l1 && x - 1 <= h1, hasUp = x >= l1 && x <= h1,
                                                   hasMi = d - 2 >= tracedBackTo && x - 1 >= l2 && x - 1 <= h2;
#pragma unroll
                                        for (int st = 0; st < S; st++) {
                                            cur[st] = B0[st * N + (x & NM)];
                                            nl[st] = hasLo ? __ldcs(rowPtr(d - 1, x - 1) + st) : NI;
                                            nm[st] = hasMi ? __ldcs(rowPtr(d - 2, x - 1) + st) : NI;
                                            nu_[st] = hasUp ? __ldcs(rowPtr(d - 1, x) + st) : NI;
                                        }
                                        cellGen(2, x, y, loadIn(x, y, true), cur, nl, hasLo, nm, hasMi, nu_, hasUp);
                                    }
                                    if (SM == 7) {
                                        // assignments in the reference's order: ascending x, from match / gap X / gap Y
                                        const int cnt = __popc(asgMask);
                                        int incl = cnt;
#pragma unroll
                                        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(CP_FULL, incl, o); if (lane >= o) incl += t; }
                                        int pos = nPairs + incl - cnt;
                                        for (int from = 0; from < 3; from++)
                                            if (asgMask & (1 << from)) {
                                                if (pos < it.pair_cap) { pairs[3 * pos] = from; pairs[3 * pos + 1] = x >= 1 ? x - 1 : 0; pairs[3 * pos + 2] = y - 1; }
                                                pos++;
                                            }
                                        nPairs += __shfl_sync(CP_FULL, incl, 31);
                                    }
                                }
                            } else status |= 2;
                        } else
                        // posteriors: impl/pairwiseAligner.c:756-795 (match state) and :797-839 (echelon: states 1 .. 5,
                        // an event matched to st k-mers gives st pairs)
                        for (int xb = blo; xb <= bhi; xb += 32) {
                            const int x = xb + lane, y = d - x;
                            const double fCur = fPre;
                            if (SM != 5 && x + 32 <= bhi) fPre = __ldcs(rowPtr(d, x + 32));
                            int cnt = 0;
                            int sc_[6] = { 0, 0, 0, 0, 0, 0 };
                            if (x <= bhi && x > 0 && y > 0) {
                                const double *rp = rowPtr(d, x);
                                if (SM == 5) {
#pragma unroll
                                    for (int st = 1; st < 6; st++) {
                                        double p = exp((rp[st] + B0[st * N + (x & NM)]) - total);
                                        if (p >= A.G.threshold) { if (p > 1.0) p = 1.0; sc_[st] = (int) floor(p * 10000000); cnt += st; } else sc_[st] = -1;
                                    }
                                } else {
                                    double p = exp(fCur + B0[(x & NM)] - total);
                                    if (p >= A.G.threshold) { if (p > 1.0) p = 1.0; sc_[1] = (int) floor(p * 10000000); cnt = 1; } else sc_[1] = -1;
                                }
                            }
                            int incl = cnt;                     // inclusive prefix sum over the lanes
#pragma unroll
                            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(CP_FULL, incl, o); if (lane >= o) incl += t; }
                            int pos = nPairs + incl - cnt;
                            if (cnt > 0) {
                                if (SM == 5) {
                                    for (int st = 1; st < 6; st++)
                                        if (sc_[st] >= 0)
                                            for (int n = 0; n < st; n++) {
                                                if (pos < it.pair_cap) { pairs[3 * pos] = sc_[st]; pairs[3 * pos + 1] = (x + n) - 1; pairs[3 * pos + 2] = y - 1; }
                                                pos++;
                                            }
                                } else if (pos < it.pair_cap) { pairs[3 * pos] = sc_[1]; pairs[3 * pos + 1] = x - 1; pairs[3 * pos + 2] = y - 1; }
                            }
                            nPairs += __shfl_sync(CP_FULL, incl, 31);
                        }
                    }
                    // next diagonal: B(d) becomes "d+1", d-1 -> d, d-2 -> d-1, the old d+1 buffer is the new d-2
                    { const int t = bp; bp = b0; b0 = b1; b1 = b2; b2 = t; }
                    lp = blo; hp = bhi; blo = l1; bhi = h1; l1 = l2; h1 = h2;
                    __syncwarp();
                }
            }
            tracedBackTo = tbf;

            // =============================== restore the forward state at Dt ==============================
            if (tracedBackTo < D) {
                f0 = 0; f1 = 1; f2 = 2;
                bandOf(Dt - 1, lo1, hi1);
                lo2 = 0; hi2 = -1;                                // re-derived when the sweep advances
                for (int x = lo + lane; x <= hi; x += 32) { const double *rp = rowPtr(Dt, x);
#pragma unroll
                    for (int st = 0; st < S; st++) buf(f0)[st * N + (x & NM)] = rp[st]; }
                for (int x = lo1 + lane; x <= hi1; x += 32) { const double *rp = rowPtr(Dt - 1, x);
#pragma unroll
                    for (int st = 0; st < S; st++) buf(f1)[st * N + (x & NM)] = rp[st]; }
                __syncwarp();
            }
        }
        status = __reduce_or_sync(CP_FULL, status);          // bit 8 is raised by single lanes
        if (P.mode == 1 && S == 3) {
            if (SM != 4) {
#pragma unroll
                for (int i = 0; i < 9; i++) {
                    double v = accT[i];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(CP_FULL, v, o);
                    if (lane == 0 && v != 0.0) atomicAdd(A.expect + i, v);
                }
            }
            if (lane == 0) atomicAdd(A.expect + (SM == 4 ? 60 : 9 + 4096), lik);
        }
        if (lane == 0) {
            ItemOut &o = A.out[itemIdx];
            o.n_pairs = nPairs;
            o.status = status | (nPairs > it.pair_cap ? 1 : 0);
            o.total_logprob = lastTotal;
            o.n_tracebacks = nTb;
        }
    }
}

}  // namespace cpecan
