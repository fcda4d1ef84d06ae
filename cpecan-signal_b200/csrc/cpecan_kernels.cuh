// cpecan_kernels.cuh -- shared device code of the banded signal pair-HMM engine (sm_100a): parameter / item records,
// the plan kernel (the band of band_construct and the traceback schedule, walked once per item) and the staging kernels
// that turn the reference's FP64 inputs into the FP32 records the alignment kernel (cpecan_align3.cuh) streams.
//
// Reference semantics restated here (nothing is translated from the C sources):
//   band_construct                    impl/pairwiseAligner.c:98-184
//   k-mer indexing / model tables     impl/stateMachine.c:104-153, 221-240
//   emissions_signal_scaleModel       impl/stateMachine.c:631-651
//   strawMan / vanilla emissions      impl/stateMachine.c:322-343, 499-528, 595-629
//   vanilla skip bins / transitions   impl/stateMachine.c:388-427, 1368-1409
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

namespace cpecan {

#define CP_NEG_INF (__int_as_float(0xff800000))
#define CP_FULL 0xffffffffu

struct DevParams {
    float tMC, tMX, tMY, tOX, tOY, tEX, tEY, tSX, tSY;  // StateMachine3 transitions (log)
    float startv[3], rstartv[3], endv[3], rendv[3];      // state vectors M, X, Y
    float threshold;
    int minDiags, tbDiags, expansion;
    int totalEvery;    // 10 (impl/pairwiseAligner.c:956)
    int rebaseEvery;   // diagonals between integer re-basing of the FP32 rows
    int mode;          // CPECAN_MODE_*
    int machine;       // 0 = three-state (strawMan), 1 = vanilla
    float vYM, vYY;    // vanilla: log(1 - E_TO_E), log(E_TO_E)
    int hasSX;
    int dbgLogP;       // debug: emit the raw float bits of log p (un-clamped) instead of the integer score
};

struct __align__(16) Item {
    long long xp_off;     // first x-parameter record (record i <-> matrix x = i, 0..lX)
    long long ev_off;     // first float4 event (entry i <-> matrix y = i, 0..lY; entry 0 is a dummy)
    long long an_off;     // first anchor pair (int64 x, y sequence coordinates)
    long long pair_off;   // first output triple
    long long tot_off;    // first debug total (or -1)
    int lX, lY, nA, flags;
    int pair_cap, model_id, pad0, pad1;
};

struct __align__(8) ItemOut {
    long long band_cells;
    double total_logprob;
    int n_pairs, status, n_tracebacks, max_width, max_rows, pad;
};

// ------------------------------------------------------------------------------------------------ plan kernel (k_align3)
// One thread per item walks the band exactly as band_construct does (impl/pairwiseAligner.c:98-184: x-y limits with
// parity fixing and C integer division, so odd expansions come out as in the reference) and leaves what k_align3 needs:
//   * 2 bits per diagonal d = 0 .. D at word d / 16: bit 0 = lo(d) - lo(d-1), bit 1 = hi(d) - hi(d-1) (x limits);
//   * the list of diagonals at which getPosteriorProbsWithBanding traces back (impl/pairwiseAligner.c:903-918);
//   * band cells, widest diagonal, longest run of forward rows alive at a traceback;
//   * flag bit 4 (CPECAN_ITEM_BAND_STEP) when a band edge moves backwards or by more than one cell (never for anchors
//     that went through filterToRemoveOverlap and an even expansion): k_align3 skips such an item; flag bit 16 when a
//     diagonal is empty (no kernel aligns that);
__device__ __forceinline__ long long cp_clampz(long long z, long long l) { return z < 0 ? 0 : (z > l ? l : z); }
__global__ void k_plan3(const Item *items, int n, const long long *anchors, DevParams P, ItemOut *out, unsigned *bits,
                        int *tbs, int *flags, int2 *bands /* or null */, const long long *band_off) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Item it = items[i];
    const long long *an = anchors + 2 * it.an_off;
    const long long lX = it.lX, lY = it.lY, D = lX + lY, e = P.expansion;
    unsigned *bw = bits + it.pad0;
    int *tb = tbs + it.pad1;
    // the FP64 kernel (cpecan_generic.cuh) takes the band as explicit (lo, hi) per diagonal: it also serves the bands of odd
    // expansions, whose edges move backwards and by two cells
    int2 *bandOut = bands != nullptr ? bands + band_off[i] : nullptr;
    long long ai = 0, pxay = 0, pxmy = 0, nxay = 0, nxmy = 0, xL = 0, yL = 0, xU = 0, yU = 0;
    long long cells = 0;
    int maxw = 0, maxrows = 1, ntb = 0, tracedBackTo = 0, flag = 0;
    long long plo = 0, phi = 0;
    unsigned word = 0;
    for (long long xay = 0; xay <= D; xay++) {
        long long l = xL - yL, r = xU - yU;
        if ((xay + l) % 2 != 0) l++;
        if ((xay + r) % 2 != 0) r++;
        long long t;
        t = (xay + l) / 2; if (t < xL) l += 2 * (xL - t);
        t = (xay - l) / 2; if (yL < t) l += 2 * (t - yL);
        t = (xay + r) / 2; if (xU < t) r -= 2 * (t - xU);
        t = (xay - r) / 2; if (t < yU) r -= 2 * (yU - t);
        const long long lo = (xay + l) / 2, hi = (xay + r) / 2;
        if (bandOut != nullptr) bandOut[xay] = make_int2((int) lo, (int) hi);
        if (xay > 0) {
            const long long dl = lo - plo, dh = hi - phi;
            if (dl < 0 || dl > 1 || dh < 0 || dh > 1) flag |= 4;
            word |= (unsigned) ((dl & 1) | ((dh & 1) << 1)) << ((xay & 15) * 2);
        } else if (lo != 0 || hi != 0) flag |= 4;
        if ((xay & 15) == 15 || xay == D) { bw[xay >> 4] = word; word = 0; }
        plo = lo; phi = hi;
        const int w = (int) (hi - lo + 1);
        if (w < 1) flag |= 4 | 16;
        cells += w;
        maxw = max(maxw, w);
        if (xay > 0) {
            const bool atEnd = xay == D;
            const bool tbp = P.mode == 2 ? false : (xay >= tracedBackTo + P.minDiags && w <= 2 * e + 1);
            if (atEnd || tbp) {
                maxrows = max(maxrows, (int) xay - tracedBackTo + 1);
                tracedBackTo = atEnd ? (int) xay : (int) xay - (P.tbDiags + 1);
                tb[ntb++] = (int) xay;
            }
        }
        if (nxay == xay) {
            pxay = nxay; pxmy = nxmy;
            long long x = lX, y = lY;
            if (ai < it.nA) { x = an[2 * ai] + 1; y = an[2 * ai + 1] + 1; ai++; }
            nxay = x + y; nxmy = x - y;
            xL = cp_clampz((pxay + (pxmy - e)) / 2, lX);
            yL = cp_clampz((nxay - (nxmy - e)) / 2, lY);
            xU = cp_clampz((nxay + (nxmy + e)) / 2, lX);
            yU = cp_clampz((pxay - (pxmy + e)) / 2, lY);
        }
    }
    ItemOut o;
    o.band_cells = cells; o.total_logprob = 0.0; o.n_pairs = 0; o.status = 0; o.n_tracebacks = ntb;
    o.max_width = maxw; o.max_rows = maxrows + 2; o.pad = 0;
    out[i] = o;
    atomicOr(flags + i, flag);
}

// ------------------------------------------------------------------------------------------------ preparation
// Events: reference layout (mean, noise, duration) doubles -> float4 (mean - centre, noise, 1 / noise,
// -1.5 log(noise)); the last two serve the inverse-Gaussian noise term of the vanilla machine.  Entry 0 of each item
// is the finite dummy standing in for "no event" (y = 0).
__global__ void k_prep_events(const Item *items, const long long *ev_src_off, const double *events,
                              const double *centre, float4 *out) {
    const int i = blockIdx.x;
    const Item it = items[i];
    const double c0 = centre[i];
    const double *src = events + 3 * ev_src_off[i];
    float4 *dst = out + it.ev_off;
    for (int y = blockIdx.y * blockDim.x + threadIdx.x; y <= it.lY; y += gridDim.y * blockDim.x) {
        float4 v = make_float4(0.f, 1.f, 1.f, 0.f);
        if (y > 0) {
            const double m = src[3 * (y - 1)], n = src[3 * (y - 1) + 1];
            v.x = (float) (m - c0); v.y = (float) n; v.z = (float) (1.0 / n); v.w = (float) (-1.5 * log(n));
        }
        dst[y] = v;
    }
}

struct ModelTables {      // device pointers of one uploaded pore model
    const double *match, *gapy, *gapx;
    int n_gapx;
    // an HDP model (cpecan_cuda_upload_hdp) instead: per distinct distribution the posterior predictive density on the
    // sampling grid and its spline slopes, and for every ACGT 6-mer the distribution it reads (-1: none)
    const double *hdp_y, *hdp_slope;
    const int *hdp_kmer;
    double hdp_x0, hdp_x1, hdp_xn, hdp_step;   // grid[0], grid[1], grid[n-1], (stop - start) / (n - 1)
    int hdp_len;
};

__device__ __forceinline__ int base_code(char b) {
    return b == 'A' ? 0 : (b == 'C' ? 1 : (b == 'G' ? 2 : (b == 'T' ? 3 : -1)));
}
__device__ __forceinline__ int kmer_code(const char *s) {   // impl/stateMachine.c:104-139; -1 <=> index > 4096
    int v = 0;
#pragma unroll
    for (int j = 0; j < 6; j++) { int b = base_code(s[j]); if (b < 0) return -1; v = v * 4 + b; }
    return v;
}

// x-parameter record: 4 float4 per matrix column x (sequence index x-1), stored as four PLANES of lX + 2 float4 per
// item (a[0..lX+1], b[..], c[..], d[..]); record lX+1 is the all -inf dummy read for columns beyond the matrix.
//   a = (mu_m - c0, -1/(2 sd_m^2), nu_m, q_m)      b = (K_m, mu_y - c0, -1/(2 sd_y^2), nu_y)
//   c = (q_y, K_y, gap-X emission | log a_my, k-mer index | skip bin as int bits)
//   d = vanilla only: (log a_mx, log a_xx, log a_mm, log a_xm)
// three-state (two log-Gaussians, impl/stateMachine.c:333-343, 595-629):  q = -1/(2 tau^2),
//   emission = K + a.y (m - mu)^2 + q (n - nu)^2,  K = -2 * 0.9189 - log sd - log tau
// vanilla (Gaussian level x inverse-Gaussian noise, :322-331, 499-528):  q = -lambda / (2 nu^2),
//   emission = K + a.y (m - mu)^2 + q (n - nu)^2 / n - 1.5 log n,  K = -0.9189 - log sd + (log lambda - log 2 pi) / 2
// The match table is scaled per read exactly as emissions_signal_scaleModel does (:636-650); the gap-Y ("extra
// event") table never is.  x = 0 reads the literal "n" in the three-state machine => every emission is LOG_ZERO; the
// vanilla machine reads through sequence_getKmer2 (impl/pairwiseAligner.c:320-325): columns 0 and 1 both use the
// k-mer pair starting at nucleotide 0.
__global__ void k_prep_xparams(const Item *items, const long long *ref_off, const char *ref,
                               const ModelTables *models, const double *scale /*5 per item or null*/,
                               const double *centre, float4 *out, int machine, double mToYNotX, int *flags) {
    const int i = blockIdx.x;
    const Item it = items[i];
    const ModelTables mt = models[it.model_id];
    const char *r = ref + ref_off[i];
    const double c0 = centre[i];
    double sc = 1, sh = 0, var = 1, scsd = 1, varsd = 1;
    const bool scaled = scale != nullptr;
    if (scaled) { sc = scale[5 * i]; sh = scale[5 * i + 1]; var = scale[5 * i + 2]; scsd = scale[5 * i + 3]; varsd = scale[5 * i + 4]; }
    float4 *dst = out + (machine ? 4 : 3) * it.xp_off;      // the fourth plane (per-column transitions) is the vanilla machine's
    const float ninf = CP_NEG_INF;
    const double HL2PI = 0.91893853320467267;
    for (int x = blockIdx.y * blockDim.x + threadIdx.x; x <= it.lX + 1; x += gridDim.y * blockDim.x) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = make_float4(ninf, 0.f, 0.f, 0.f);
        float4 c = make_float4(0.f, ninf, machine ? 0.f : ninf, __int_as_float(-1)), d = make_float4(ninf, ninf, ninf, ninf);
        int k = -1, kprev = -1;
        if (x <= it.lX) {
            if (machine == 0) { if (x > 0) k = kmer_code(r + x - 1); }
            else { const int p = x >= 2 ? x - 2 : 0; kprev = kmer_code(r + p); k = kmer_code(r + p + 1); }
        }
        // a reference k-mer that is not ACGT makes every path -inf: the E-step drops such a read (flag bit 8)
        if (flags != nullptr && k < 0 && x >= (machine ? 0 : 1) && x <= it.lX) atomicOr(flags + i, 8);
        auto scaledMean = [&](int kk) { return kk < 0 ? 0.0 : (scaled ? mt.match[1 + 5 * kk] * sc + sh : mt.match[1 + 5 * kk]); };
        if (k >= 0) {
            const double *m = mt.match + 1 + 5 * k;
            double mu = m[0], sd = m[1], nu = m[2], tau = m[3], lam = m[4];
            if (scaled) { mu = mu * sc + sh; sd = sd * var; nu = nu * scsd; lam = lam * varsd; tau = sqrt(pow(nu, 3.0) / lam); }
            const double *g = mt.gapy + 1 + 5 * k;
            const double muy = g[0], sdy = g[1], nuy = g[2], tauy = g[3], lamy = g[4];
            if (machine == 0) {
                if (sd != 0.0 && tau != 0.0) {
                    a = make_float4((float) (mu - c0), (float) (-0.5 / (sd * sd)), (float) nu, (float) (-0.5 / (tau * tau)));
                    b.x = (float) (-2.0 * HL2PI - log(sd) - log(tau));
                }
                if (sdy != 0.0 && tauy != 0.0) {
                    b.y = (float) (muy - c0); b.z = (float) (-0.5 / (sdy * sdy)); b.w = (float) nuy;
                    c.x = (float) (-0.5 / (tauy * tauy));
                    c.y = (float) (-2.0 * HL2PI - log(sdy) - log(tauy));
                }
                c.z = (float) mt.gapx[k];
                c.w = __int_as_float(k);
            } else {
                if (sd != 0.0) {
                    a = make_float4((float) (mu - c0), (float) (-0.5 / (sd * sd)), (float) nu, (float) (-lam / (2.0 * nu * nu)));
                    b.x = (float) (-HL2PI - log(sd) + 0.5 * (log(lam) - 2.0 * HL2PI));
                }
                if (sdy != 0.0) {
                    b.y = (float) (muy - c0); b.z = (float) (-0.5 / (sdy * sdy)); b.w = (float) nuy;
                    c.x = (float) (-lamy / (2.0 * nuy * nuy));
                    c.y = (float) (-HL2PI - log(sdy) + 0.5 * (log(lamy) - 2.0 * HL2PI));
                }
            }
        }
        if (machine == 1 && x <= it.lX) {
            // emissions_signal_getKmerSkipBin (:388-419) on the SCALED match table, then the per-cell transition
            // probabilities of stateMachine3Vanilla_cellCalculate (:1378-1390), logs taken once per column
            const double dmu = fabs(scaledMean(k) - scaledMean(kprev));
            long long bin = (long long) (dmu / 0.5);
            bin = bin >= 30 ? 29 : bin;
            const double a_mx = mt.gapx[bin], a_xx = mt.gapx[bin + 30];
            const double a_my = (1 - a_mx) * mToYNotX, a_mm = 1.0 - a_my - a_mx, a_xm = 1.0 - a_xx;
            d = make_float4((float) log(a_mx), (float) log(a_xx), (float) log(a_mm), (float) log(a_xm));
            c.z = (float) log(a_my);
            c.w = __int_as_float((int) bin);
        }
        const int plane = it.lX + 2;          // planes per item: lanes of a warp read neighbouring records of one plane
        dst[x] = a; dst[plane + x] = b; dst[2 * plane + x] = c;
        if (machine) dst[3 * plane + x] = d;
    }
}

#define CP_INT_MIN (-2147483647 - 1)

}  // namespace cpecan
