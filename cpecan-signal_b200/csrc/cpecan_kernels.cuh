// cpecan_kernels.cuh -- sm_100a device code for the banded signal pair-HMM (forward / backward / posterior).
//
// What is computed (reference semantics, restated for the GPU -- nothing here is translated from the C sources):
//   * the banded forward sweep, the periodic traceback with a backward sweep re-started from the end-state vector,
//     the local totalProbability recomputed every 10th diagonal and the thresholded posterior match
//     probabilities of impl/pairwiseAligner.c:870-1006 (getPosteriorProbsWithBanding) and :714-795;
//   * over the three-state signal machine of impl/stateMachine.c:1305-1334 with the two-Gaussian k-mer/event
//     emissions of :595-629 and the 4-segment cubic logAdd of impl/pairwiseAligner.c:235-255.
//
// Mapping (B200-first):
//   * one CTA of G warps per alignment ("work item"); persistent CTAs pull items from a global queue;
//   * cells are owned by x: ring slot s = x mod N, N = 32*G*K, thread t owns slots t*K .. t*K+K-1.  With that
//     ownership the "upper" neighbour (x, y-1) is the thread's own previous value, "lower" (x-1, y) and "middle"
//     (x-1, y-1) are the previous slot's last two values: registers, plus ONE warp shuffle per diagonal for the
//     slot that crosses a lane (and a 5-float shared-memory hand-off across warps when G > 1);
//   * the k-mer side of the emissions (11 coefficients per reference position, computed once per read by
//     k_prep_xparams in FP64) stays in registers while x is inside the band; events stream through the slots
//     systolically (one shuffle per diagonal), so the hot loop touches memory only to spill the forward row;
//   * all DP values are FP32 held relative to a PER-THREAD integer offset that follows the thread's own cells
//     (logAdd is translation invariant and shifting by an integer is exact).  A per-row offset is not enough:
//     inside one diagonal the reference's values span hundreds of nats (the un-scaled gap-Y table makes whole
//     regions of the band astronomically unlikely), and the cells that matter are not the row maximum;
//   * logAdd reproduces the reference's cubic segments and its 7.5 cut-off, and every fold (per-cell transition
//     order, per-diagonal dot products) runs in the reference's order because the approximation is not associative.
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
    int hasSX;
    int dbgLogP;       // debug: emit the raw float bits of log p (un-clamped) instead of the integer score
};

struct __align__(16) Item {
    long long xp_off;     // first x-parameter record (record i <-> matrix x = i, 0..lX)
    long long ev_off;     // first float2 event (entry i <-> matrix y = i, 0..lY; entry 0 is a dummy)
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

// ------------------------------------------------------------------------------------------------ logAdd
// impl/pairwiseAligner.c:238-255.  Coefficients are the reference's float literals (exact in FP32).
__device__ __forceinline__ float la_poly(float x) {
    const bool s1 = x <= 1.0f, s2 = x <= 2.5f, s3 = x <= 4.5f;
    const float a = s1 ? -0.009350833524763f : (s2 ? -0.014532321752540f : (s3 ? -0.004605031767994f : -0.000458661602210f));
    const float b = s1 ? 0.130659527668286f : (s2 ? 0.139942324101744f : (s3 ? 0.063427417320019f : 0.009695946122598f));
    const float c = s1 ? 0.498799810682272f : (s2 ? 0.495635523139337f : (s3 ? 0.695956496475118f : 0.930734667215156f));
    const float d = s1 ? 0.693203116424741f : (s2 ? 0.692140569840976f : (s3 ? 0.514272634594009f : 0.168037164329057f));
    return fmaf(fmaf(fmaf(a, x, b), x, c), x, d);
}

__device__ __forceinline__ float logadd(float x, float y) {
    const float hi = fmaxf(x, y), lo = fminf(x, y);
    const float d = hi - lo;              // NaN when both are -inf, +inf when only lo is
    const float r = lo + la_poly(d);
    return (d < 7.5f) ? r : hi;           // NaN / inf / >= 7.5 all fall through to hi
}

// ------------------------------------------------------------------------------------------------ band geometry
// impl/pairwiseAligner.c:98-184 in closed form: between the previous anchor p and the next anchor n (matrix
// coordinates = sequence + 1) the band is the rectangle [p.x - e/2, n.x + e/2] x [p.y - e/2, n.y + e/2] clamped to
// the matrix; diagonal d (p.x+p.y < d <= n.x+n.y) keeps the cells x in [max(xL, d - yL), min(xU, d - yU)].
struct BandWalker {
    const long long *anchors;  // pairs
    int nA, lX, lY, h;
    int ai;                    // index of the "next" anchor of the current box
    int pd, nd;                // x+y of previous / next point
    int xL, yL, xU, yU;

    __device__ __forceinline__ void setBox() {
        int px = 0, py = 0, nx = lX, ny = lY;
        if (ai > 0) { px = (int) anchors[2 * (ai - 1)] + 1; py = (int) anchors[2 * (ai - 1) + 1] + 1; }
        if (ai < nA) { nx = (int) anchors[2 * ai] + 1; ny = (int) anchors[2 * ai + 1] + 1; }
        pd = px + py; nd = nx + ny;
        xL = min(max(px - h, 0), lX);
        yL = min(max(ny + h, 0), lY);
        xU = min(max(nx + h, 0), lX);
        yU = min(max(py - h, 0), lY);
    }
    __device__ __forceinline__ void init(const long long *a, int nA_, int lX_, int lY_, int expansion) {
        anchors = a; nA = nA_; lX = lX_; lY = lY_; h = expansion / 2; ai = 0; setBox();
    }
    // x range of diagonal d; moves the box forwards or backwards as needed
    __device__ __forceinline__ void range(int d, int &lo, int &hi) {
        while (d > nd && ai < nA) { ai++; setBox(); }
        while (d <= pd && ai > 0) { ai--; setBox(); }
        lo = max(xL, d - yL);
        hi = min(xU, d - yU);
    }
};

// ------------------------------------------------------------------------------------------------ plan kernel
// One thread per item: walks the band once and reports the work (band cells), the widest diagonal and the longest
// run of forward rows that must be alive at a traceback (sizes the per-CTA forward spill ring).
__global__ void k_plan(const Item *items, int n, const long long *anchors, DevParams P, ItemOut *out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Item it = items[i];
    BandWalker bw;
    bw.init(anchors + 2 * it.an_off, it.nA, it.lX, it.lY, P.expansion);
    const int D = it.lX + it.lY;
    long long cells = 0;
    int maxw = 0, maxrows = 1, ntb = 0, tracedBackTo = 0;
    for (int d = 0; d <= D; d++) {
        int lo, hi;
        bw.range(d, lo, hi);
        int w = hi - lo + 1;
        cells += w;
        maxw = max(maxw, w);
        if (d == 0) continue;
        bool atEnd = d == D;
        bool tb = P.mode == 2 ? false : (d >= tracedBackTo + P.minDiags && w <= 2 * P.expansion + 1);
        if (atEnd || tb) {
            maxrows = max(maxrows, d - tracedBackTo + 1);
            tracedBackTo = atEnd ? d : d - (P.tbDiags + 1);
            ntb++;
        }
    }
    ItemOut o;
    o.band_cells = cells; o.total_logprob = 0.0; o.n_pairs = 0; o.status = 0; o.n_tracebacks = ntb;
    o.max_width = maxw; o.max_rows = maxrows + 2; o.pad = 0;
    out[i] = o;
}

// ------------------------------------------------------------------------------------------------ preparation
// Events: reference layout (mean, noise, duration) doubles -> float2 (mean - centre, noise), entry 0 of each item is
// the finite dummy standing in for "no event" (y = 0).
__global__ void k_prep_events(const Item *items, const long long *ev_src_off, const double *events,
                              const double *centre, float2 *out) {
    const int i = blockIdx.x;
    const Item it = items[i];
    const double c0 = centre[i];
    const double *src = events + 3 * ev_src_off[i];
    float2 *dst = out + it.ev_off;
    for (int y = blockIdx.y * blockDim.x + threadIdx.x; y <= it.lY; y += gridDim.y * blockDim.x) {
        float2 v;
        if (y == 0) { v.x = 0.f; v.y = 1.f; }
        else { v.x = (float) (src[3 * (y - 1)] - c0); v.y = (float) src[3 * (y - 1) + 1]; }
        dst[y] = v;
    }
}

struct ModelTables {      // device pointers of one uploaded pore model
    const double *match, *gapy, *gapx;
    int n_gapx;
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

// x-parameter record, 12 floats per matrix column x (sequence index x-1), three-state machine:
//   [0] mu_m - c0   [1] -1/(2 sd_m^2)   [2] nu_m   [3] -1/(2 tau_m^2)
//   [4] K_m         [5] mu_y - c0       [6] -1/(2 sd_y^2)   [7] nu_y
//   [8] -1/(2 tau_y^2)   [9] K_y        [10] gap-X emission (log)   [11] k-mer index as int bits (-1: none)
// so that  match(x, ev) = K_m + c1 (m - mu)^2 + c2 (n - nu)^2  (two log-Gaussians, impl/stateMachine.c:333-343,
// 595-629) and likewise for the gap-Y ("extra event") table, which the reference never scales (:631-651).
// x = 0 (sequence index -1) reads the literal "n" in the reference => every emission is LOG_ZERO.
__global__ void k_prep_xparams3(const Item *items, const long long *ref_off, const char *ref,
                                const ModelTables *models, const double *scale /*5 per item or null*/,
                                const double *centre, float4 *out) {
    const int i = blockIdx.x;
    const Item it = items[i];
    const ModelTables mt = models[it.model_id];
    const char *r = ref + ref_off[i];
    const double c0 = centre[i];
    double sc = 1, sh = 0, var = 1, scsd = 1, varsd = 1;
    const bool scaled = scale != nullptr;
    if (scaled) { sc = scale[5 * i]; sh = scale[5 * i + 1]; var = scale[5 * i + 2]; scsd = scale[5 * i + 3]; varsd = scale[5 * i + 4]; }
    float4 *dst = out + 3 * it.xp_off;
    const float ninf = CP_NEG_INF;
    // record lX + 1 is the all -inf dummy the second-generation kernel reads for columns beyond the matrix
    for (int x = blockIdx.y * blockDim.x + threadIdx.x; x <= it.lX + 1; x += gridDim.y * blockDim.x) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = make_float4(ninf, 0.f, 0.f, 0.f), c = make_float4(0.f, ninf, ninf, __int_as_float(-1));
        const int k = (x > 0 && x <= it.lX) ? kmer_code(r + x - 1) : -1;
        if (k >= 0) {
            const double *m = mt.match + 1 + 5 * k;
            double mu = m[0], sd = m[1], nu = m[2], tau = m[3], lam = m[4];
            if (scaled) {   // emissions_signal_scaleModel, impl/stateMachine.c:636-650
                mu = mu * sc + sh; sd = sd * var; nu = nu * scsd; lam = lam * varsd; tau = sqrt(pow(nu, 3.0) / lam);
            }
            if (sd != 0.0 && tau != 0.0) {
                a.x = (float) (mu - c0); a.y = (float) (-0.5 / (sd * sd)); a.z = (float) nu; a.w = (float) (-0.5 / (tau * tau));
                b.x = (float) (-2.0 * 0.91893853320467267 - log(sd) - log(tau));
            }
            const double *g = mt.gapy + 1 + 5 * k;
            double muy = g[0], sdy = g[1], nuy = g[2], tauy = g[3];
            if (sdy != 0.0 && tauy != 0.0) {
                b.y = (float) (muy - c0); b.z = (float) (-0.5 / (sdy * sdy)); b.w = (float) nuy;
                c.x = (float) (-0.5 / (tauy * tauy));
                c.y = (float) (-2.0 * 0.91893853320467267 - log(sdy) - log(tauy));
            }
            c.z = (float) mt.gapx[k];
            c.w = __int_as_float(k);
        }
        dst[3 * x] = a; dst[3 * x + 1] = b; dst[3 * x + 2] = c;
    }
}

// ------------------------------------------------------------------------------------------------ helpers
template <int G> __device__ __forceinline__ void cta_sync() {
    if (G > 1) __syncthreads(); else __syncwarp();
}

// Hand NV floats to the NEXT thread of the CTA ring (thread t receives what thread t-1 passed in).
template <int G, int NV>
__device__ __forceinline__ void pass_to_next(float (&v)[NV], float *sm, int parity) {
    const int lane = threadIdx.x & 31;
    if (G > 1) {
        const int warp = threadIdx.x >> 5;
        float *buf = sm + parity * (G * NV);
        if (lane == 31) {
#pragma unroll
            for (int i = 0; i < NV; i++) buf[warp * NV + i] = v[i];
        }
        __syncthreads();
        const int pw = (warp + G - 1) % G;
#pragma unroll
        for (int i = 0; i < NV; i++) {
            float r = __shfl_up_sync(CP_FULL, v[i], 1);
            v[i] = lane == 0 ? buf[pw * NV + i] : r;
        }
    } else {
#pragma unroll
        for (int i = 0; i < NV; i++) v[i] = __shfl_sync(CP_FULL, v[i], (lane + 31) & 31);
    }
}
// Hand NV floats to the PREVIOUS thread (thread t receives what thread t+1 passed in).
template <int G, int NV>
__device__ __forceinline__ void pass_to_prev(float (&v)[NV], float *sm, int parity) {
    const int lane = threadIdx.x & 31;
    if (G > 1) {
        const int warp = threadIdx.x >> 5;
        float *buf = sm + parity * (G * NV);
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < NV; i++) buf[warp * NV + i] = v[i];
        }
        __syncthreads();
        const int nw = (warp + 1) % G;
#pragma unroll
        for (int i = 0; i < NV; i++) {
            float r = __shfl_down_sync(CP_FULL, v[i], 1);
            v[i] = lane == 31 ? buf[nw * NV + i] : r;
        }
    } else {
#pragma unroll
        for (int i = 0; i < NV; i++) v[i] = __shfl_sync(CP_FULL, v[i], (lane + 1) & 31);
    }
}

template <int G> __device__ __forceinline__ int cta_max_int(int v, int *red) {
    v = __reduce_max_sync(CP_FULL, v);
    if (G > 1) {
        __syncthreads();                 // protect red[] from the previous use
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
        __syncthreads();
        int m = red[0];
#pragma unroll
        for (int w = 1; w < G; w++) m = max(m, red[w]);
        v = m;
    }
    return v;
}

template <int G> __device__ __forceinline__ int cta_any(int pred) {
    if (G > 1) return __syncthreads_or(pred);
    return __any_sync(CP_FULL, pred);
}

// Left fold  acc = logadd(acc, v[0]), logadd(acc, v[1]), ...  over buf[0..w) (the order of
// dpDiagonal_dotProduct, impl/pairwiseAligner.c:587-597) executed by warp 0.  Exactly equal to the serial fold:
// an element can change acc only if it is within 7.5 of the running prefix maximum (acc >= prefix max and
// logadd returns its larger argument unchanged beyond 7.5), so only those are folded -- serially, in order.
__device__ __forceinline__ float warp_ordered_fold(const float *buf, int w) {
    const int lane = threadIdx.x & 31;
    float acc = CP_NEG_INF, runmax = CP_NEG_INF;
    for (int base = 0; base < w; base += 32) {
        const int idx = base + lane;
        const float v = idx < w ? buf[idx] : CP_NEG_INF;
        float m = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { float t = __shfl_up_sync(CP_FULL, m, o); if (lane >= o) m = fmaxf(m, t); }
        float excl = __shfl_up_sync(CP_FULL, m, 1);
        excl = lane == 0 ? runmax : fmaxf(excl, runmax);
        const bool live = v > excl - 7.6f;     // margin: a spurious "live" element is folded exactly anyway
        unsigned mask = __ballot_sync(CP_FULL, live);
        while (mask) {
            const int b = __ffs(mask) - 1;
            mask &= mask - 1;
            acc = logadd(acc, __shfl_sync(CP_FULL, v, b));
        }
        runmax = fmaxf(runmax, __shfl_sync(CP_FULL, m, 31));
    }
    return acc;
}

// ------------------------------------------------------------------------------------------------ the kernel
struct KernelArgs {
    const Item *items;
    const int *order;             // processing order (largest first)
    int n_items;
    int *queue;                   // global work counter
    const long long *anchors;
    const float4 *xparams;
    const float2 *events;
    float *scratch;               // per CTA: ring of forward rows, 3*N floats each
    int *scratch_off;             // per CTA: ring of per-thread row offsets, T ints each
    long long scratch_stride;     // floats per CTA
    int ring_rows;
    int *pairs;                   // triples
    ItemOut *out;
    double *totals;               // debug, may be null
    DevParams P;
};

#define CP_INT_MIN (-2147483647 - 1)
#define CP_SHIFT_T 8              // re-centre a thread's registers when they drift this far from zero

template <int G, int K, bool HAS_SX>
__global__ void __launch_bounds__(32 * G) k_align(const KernelArgs A) {
    constexpr int T = 32 * G;        // threads per CTA
    constexpr int N = T * K;         // ring slots
    constexpr int NMASK = N - 1;
    __shared__ float sm_x[2 * G * 6];
    __shared__ int sm_red[G > 1 ? G : 1];
    __shared__ float sm_fold[N];
    __shared__ int sm_score[N];
    __shared__ float sm_bcast[2];
    __shared__ int sm_item;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int sbase = tid * K;
    const DevParams &P = A.P;
    float *rows = A.scratch + (long long) blockIdx.x * A.scratch_stride;
    int *rowoff = A.scratch_off + (long long) blockIdx.x * A.ring_rows * T;
    const int R = A.ring_rows;
    const float NI = CP_NEG_INF;

    for (;;) {
        cta_sync<G>();
        if (tid == 0) sm_item = atomicAdd(A.queue, 1);
        cta_sync<G>();
        const int qi = sm_item;
        if (qi >= A.n_items) break;
        const int itemIdx = A.order[qi];
        const Item it = A.items[itemIdx];
        const int lX = it.lX, lY = it.lY, D = lX + lY;
        const float4 *xp = A.xparams + 3 * it.xp_off;
        const float2 *evp = A.events + it.ev_off;
        int *pairs = A.pairs + 3 * it.pair_off;
        double *dbgTot = (A.totals != nullptr && it.tot_off >= 0) ? A.totals + it.tot_off : nullptr;
        int nPairs = 0, status = 0, nTb = 0;
        double lastTotal = 0.0;
        const bool unbanded = P.mode == 2;

        if (D == 0) {
            if (tid == 0) { ItemOut &o = A.out[itemIdx]; o.n_pairs = 0; o.status = 0; o.total_logprob = 0.0; o.n_tracebacks = 0; }
            continue;
        }

        BandWalker bw;
        bw.init(A.anchors + 2 * it.an_off, it.nA, lX, lY, P.expansion);

        // per-slot state ------------------------------------------------------------------------------
        int xs[K];
        float4 pa[K], pb[K], pc[K];      // x-parameter records
        float2 ev[K];                    // event of the slot's current cell
        float vM[K], vX[K], vY[K];       // row d-1 (forward) / G of row d+1 (backward)
        float wM[K], wX[K], wY[K];       // row d-2 (forward) / G of row d+2 (backward)
        int to = 0;                      // this thread's integer offset: true value = register + to
        float mcur = NI, mprev = NI;     // maxima of the v and w registers (for the offset choice)

        auto loadParams = [&](int k, int x) {
            if (x <= lX) { pa[k] = xp[3 * x]; pb[k] = xp[3 * x + 1]; pc[k] = xp[3 * x + 2]; }
            else { pa[k] = make_float4(0.f, 0.f, 0.f, 0.f); pb[k] = make_float4(NI, 0.f, 0.f, 0.f); pc[k] = make_float4(0.f, NI, NI, 0.f); }
        };
        auto loadEvent = [&](int y) -> float2 {
            return (y >= 1 && y <= lY) ? evp[y] : make_float2(0.f, 1.f);
        };
        auto emitM = [&](int k) -> float {
            const float dm = ev[k].x - pa[k].x, dn = ev[k].y - pa[k].z;
            return fmaf(pa[k].y, dm * dm, fmaf(pa[k].w, dn * dn, pb[k].x));
        };
        auto emitY = [&](int k) -> float {
            const float dm = ev[k].x - pb[k].y, dn = ev[k].y - pb[k].w;
            return fmaf(pb[k].z, dm * dm, fmaf(pc[k].x, dn * dn, pc[k].y));
        };
        // (re)assign every slot for base lo at diagonal d and fetch its parameters and event from memory
        auto assignAll = [&](int lo, int d) {
#pragma unroll
            for (int k = 0; k < K; k++) {
                const int x = lo + ((sbase + k - lo) & NMASK);
                xs[k] = x;
                loadParams(k, x);
                ev[k] = loadEvent(d - x);
            }
        };
        auto rowPtr = [&](int row) -> float4 * { return reinterpret_cast<float4 *>(rows + (long long) row * (3 * N)); };
        auto storeRow = [&](int row) {
            float4 *rp = rowPtr(row);
            if (K == 4) {
                rp[tid] = make_float4(vM[0], vM[1], vM[2], vM[3]);
                rp[T + tid] = make_float4(vX[0], vX[1], vX[2], vX[3]);
                rp[2 * T + tid] = make_float4(vY[0], vY[1], vY[2], vY[3]);
            } else {
                float *r = reinterpret_cast<float *>(rp);
#pragma unroll
                for (int k = 0; k < K; k++) { r[sbase + k] = vM[k]; r[N + sbase + k] = vX[k]; r[2 * N + sbase + k] = vY[k]; }
            }
            rowoff[row * T + tid] = to;
        };
        auto loadRow = [&](int row, float (&fM)[K], float (&fX)[K], float (&fY)[K]) -> int {
            const float4 *rp = rowPtr(row);
            if (K == 4) {
                float4 a = rp[tid], b = rp[T + tid], c = rp[2 * T + tid];
                fM[0] = a.x; fM[1] = a.y; fM[2] = a.z; fM[3] = a.w;
                fX[0] = b.x; fX[1] = b.y; fX[2] = b.z; fX[3] = b.w;
                fY[0] = c.x; fY[1] = c.y; fY[2] = c.z; fY[3] = c.w;
            } else {
                const float *r = reinterpret_cast<const float *>(rp);
#pragma unroll
                for (int k = 0; k < K; k++) { fM[k] = r[sbase + k]; fX[k] = r[N + sbase + k]; fY[k] = r[2 * N + sbase + k]; }
            }
            return rowoff[row * T + tid];
        };
        auto max3K = [&](const float (&a)[K], const float (&b)[K], const float (&c)[K]) -> float {
            float m = NI;
#pragma unroll
            for (int k = 0; k < K; k++) m = fmaxf(m, fmaxf(a[k], fmaxf(b[k], c[k])));
            return m;
        };
        // Offset choice for the next row: follow the largest of what this thread holds and what flows in from
        // the neighbour (absolute integer units).  All lanes of a warp re-centre together so the branch is uniform.
        auto recentre = [&](float mine, int nto, float incoming, float (&nb2)[3], float &nb2b) {
            const int aM = mine > -1e30f ? to + (int) floorf(mine) : CP_INT_MIN;
            const int aI = incoming > -1e30f ? nto + (int) floorf(incoming) : CP_INT_MIN;
            const int tgt = max(aM, aI);
            const int c = tgt == CP_INT_MIN ? 0 : tgt - to;
            if (__any_sync(CP_FULL, c >= CP_SHIFT_T || c <= -CP_SHIFT_T)) {
                const float cf = (float) c;
#pragma unroll
                for (int k = 0; k < K; k++) { vM[k] -= cf; vX[k] -= cf; vY[k] -= cf; wM[k] -= cf; wX[k] -= cf; wY[k] -= cf; }
                nb2[0] -= cf; nb2[1] -= cf; nb2[2] -= cf; nb2b -= cf;
                mcur -= cf; mprev -= cf;
                to += c;
            }
        };

        // ---- diagonal 0: the single cell (0,0) holds the start vector (impl/pairwiseAligner.c:897-898) ----
        int rowF = 0;            // ring row of the newest forward diagonal
        int dcur = 0;            // newest forward diagonal
        int tracedBackTo = 0;
        {
            const float *sv = (it.flags & 1) ? P.rstartv : P.startv;
#pragma unroll
            for (int k = 0; k < K; k++) {
                const bool z = (sbase + k) == 0;
                vM[k] = z ? sv[0] : NI; vX[k] = z ? sv[1] : NI; vY[k] = z ? sv[2] : NI;
            }
            to = 0;
            storeRow(0);
        }
        int iter = 0;            // parity for the shared-memory hand-off buffers

        while (tracedBackTo < D) {
            // =============================== forward phase ===============================================
            // (re)build the forward registers from the two newest stored rows
            int lo, hi;
            bw.range(dcur, lo, hi);
            assignAll(lo, dcur);
            to = loadRow(rowF, vM, vX, vY);
            if (dcur > 0) {
                const int rp = rowF == 0 ? R - 1 : rowF - 1;
                const int tw = loadRow(rp, wM, wX, wY);
                const float adj = (float) (tw - to);
#pragma unroll
                for (int k = 0; k < K; k++) { wM[k] += adj; wX[k] += adj; wY[k] += adj; }
            } else {
#pragma unroll
                for (int k = 0; k < K; k++) { wM[k] = NI; wX[k] = NI; wY[k] = NI; }
            }
            mcur = max3K(vM, vX, vY);
            mprev = max3K(wM, wX, wY);
            // neighbour values of row d-2 (the previous slot's row dcur-1), converted to this thread's units
            float n2[3];
            float dummy = 0.f;
            {
                float h[4] = { wM[K - 1], wX[K - 1], wY[K - 1], __int_as_float(to) };
                pass_to_next<G, 4>(h, sm_x, iter & 1); iter++;
                const float dl = (float) (__float_as_int(h[3]) - to);
                n2[0] = h[0] + dl; n2[1] = h[1] + dl; n2[2] = h[2] + dl;
            }
            int Dt = -1;
            bool atEnd = false;
            while (true) {
                const int d = dcur + 1;
                int plo = lo;
                bw.range(d, lo, hi);
                if (lo - plo > 1 || lo < plo) status |= 4;
                // hand-off: previous slot's row d-1 values, its event and its offset
                float n1[6] = { vM[K - 1], vX[K - 1], vY[K - 1], ev[K - 1].x, ev[K - 1].y, __int_as_float(to) };
                pass_to_next<G, 6>(n1, sm_x, iter & 1); iter++;
                const int nto = __float_as_int(n1[5]);
                recentre(fmaxf(fmaxf(mcur, mprev), fmaxf(n2[0], fmaxf(n2[1], n2[2]))), nto,
                         fmaxf(n1[0], fmaxf(n1[1], n1[2])), n2, dummy);
                {
                    const float dl = (float) (nto - to);
                    n1[0] += dl; n1[1] += dl; n1[2] += dl;
                }
                float cM[K], cX[K], cY[K];
#pragma unroll
                for (int k = K - 1; k >= 0; k--) {
                    const int x = lo + ((sbase + k - lo) & NMASK);
                    if (x != xs[k]) { xs[k] = x; loadParams(k, x); }
                    // events move with the anti-diagonal: cell (x, d-x) takes the event of (x-1, d-1-(x-1))
                    ev[k] = k > 0 ? ev[k - 1] : make_float2(n1[3], n1[4]);
                    if (x == lo) ev[k] = loadEvent(d - lo);
                    const bool inb = x <= hi;
                    const float lM = k > 0 ? vM[k - 1] : n1[0], lXv = k > 0 ? vX[k - 1] : n1[1], lYv = k > 0 ? vY[k - 1] : n1[2];
                    const float mM = k > 0 ? wM[k - 1] : n2[0], mX = k > 0 ? wX[k - 1] : n2[1], mY = k > 0 ? wY[k - 1] : n2[2];
                    const float uM = vM[k], uY = vY[k];
                    // impl/stateMachine.c:1314-1333, transitions folded in code order, emission added once
                    float tX = logadd(lM + P.tOX, lXv + P.tEX);
                    if (HAS_SX) tX = logadd(tX, lYv + P.tSX);
                    float tM = logadd(logadd(mM + P.tMC, mX + P.tMX), mY + P.tMY);
                    float tY = logadd(uM + P.tOY, uY + P.tEY);
                    tX += pc[k].z; tM += emitM(k); tY += emitY(k);
                    cM[k] = inb ? tM : NI; cX[k] = inb ? tX : NI; cY[k] = inb ? tY : NI;
                }
                n2[0] = n1[0]; n2[1] = n1[1]; n2[2] = n1[2];
#pragma unroll
                for (int k = 0; k < K; k++) { wM[k] = vM[k]; wX[k] = vX[k]; wY[k] = vY[k]; vM[k] = cM[k]; vX[k] = cX[k]; vY[k] = cY[k]; }
                mprev = mcur;
                mcur = max3K(vM, vX, vY);
                dcur = d;
                rowF = rowF + 1 == R ? 0 : rowF + 1;
                storeRow(rowF);
                atEnd = d == D;
                const bool tbPoint = !unbanded && d >= tracedBackTo + P.minDiags && (hi - lo + 1) <= 2 * P.expansion + 1;
                if (atEnd || tbPoint) { Dt = d; break; }
            }

            // =============================== traceback ===================================================
            // impl/pairwiseAligner.c:920-992.  Backward values live in registers as G = B + emission of the cell.
            nTb++;
            cta_sync<G>();
            const int tracedBackFrom = Dt - (atEnd ? 0 : P.tbDiags + 1);
            BandWalker bb = bw;
            int blo, bhi;
            bb.range(Dt, blo, bhi);
            assignAll(blo, Dt);
            int rowB = rowF;
            float totSt = NI;           // total in units of totBase
            int totBase = 0;
            int count = 0;
            float m2next = NI;          // G_M of row d+2 at the next slot (this thread's units)
            float n2dummy[3] = { NI, NI, NI };
            to = 0; mcur = NI; mprev = NI;
#pragma unroll
            for (int k = 0; k < K; k++) { vM[k] = NI; vX[k] = NI; vY[k] = NI; wM[k] = NI; wX[k] = NI; wY[k] = NI; }
            const float *endv = (atEnd && (it.flags & 2)) ? P.rendv : P.endv;

            for (int d = Dt; d > tracedBackTo; d--) {
                float bM[K], bX[K], bY[K];
                bool inb[K];
                if (d == Dt) {
#pragma unroll
                    for (int k = 0; k < K; k++) {
                        inb[k] = xs[k] <= bhi;
                        bM[k] = inb[k] ? endv[0] : NI; bX[k] = inb[k] ? endv[1] : NI; bY[k] = inb[k] ? endv[2] : NI;
                    }
                } else {
                    const int plo = blo;
                    bb.range(d, blo, bhi);
                    if (plo - blo > 1 || blo > plo) status |= 4;
                    // hand-off from the next slot: G_X and G_M of row d+1, its event and its offset
                    float n1[5] = { vX[0], vM[0], ev[0].x, ev[0].y, __int_as_float(to) };
                    pass_to_prev<G, 5>(n1, sm_x, iter & 1); iter++;
                    const int nto = __float_as_int(n1[4]);
                    recentre(fmaxf(fmaxf(mcur, mprev), m2next), nto, fmaxf(n1[0], n1[1]), n2dummy, m2next);
                    {
                        const float dl = (float) (nto - to);
                        n1[0] += dl; n1[1] += dl;
                    }
#pragma unroll
                    for (int k = 0; k < K; k++) {
                        const int x = blo + ((sbase + k - blo) & NMASK);
                        if (x != xs[k]) { xs[k] = x; loadParams(k, x); }
                        ev[k] = k < K - 1 ? ev[k + 1] : make_float2(n1[2], n1[3]);
                        if (x == blo + N - 1) ev[k] = loadEvent(d - x);
                        inb[k] = x <= bhi;
                        const float gx1 = k < K - 1 ? vX[k + 1] : n1[0];
                        const float gm2 = k < K - 1 ? wM[k + 1] : m2next;
                        const float gy1 = vY[k];
                        // predecessor P accumulates (impl/pairwiseAligner.c:378-383 in pull form): first from row d+2
                        // (P as "middle"), then from row d+1 in ascending x-y: P as "upper", then P as "lower".
                        float tM = logadd(logadd(gm2 + P.tMC, gy1 + P.tOY), gx1 + P.tOX);
                        float tX = logadd(gm2 + P.tMX, gx1 + P.tEX);
                        float tY = logadd(gm2 + P.tMY, gy1 + P.tEY);
                        if (HAS_SX) tY = logadd(tY, gx1 + P.tSX);
                        bM[k] = inb[k] ? tM : NI; bX[k] = inb[k] ? tX : NI; bY[k] = inb[k] ? tY : NI;
                    }
                    m2next = n1[1];
                    rowB = rowB == 0 ? R - 1 : rowB - 1;
                }

                if (d <= tracedBackFrom) {
                    float fM[K], fX[K], fY[K];
                    const int toF = loadRow(rowB, fM, fX, fY);
                    const int u = toF + to;                     // units of F + B in this thread
                    const bool doTotal = unbanded ? (d == Dt) : (count % P.totalEvery == 0);
                    count++;
                    if (doTotal) {
                        // term 1: dot(F[d], B[d])                           (impl/pairwiseAligner.c:741)
                        float c1[K];
                        float lm = NI;
#pragma unroll
                        for (int k = 0; k < K; k++) {
                            c1[k] = inb[k] ? logadd(logadd(fM[k] + bM[k], fX[k] + bX[k]), fY[k] + bY[k]) : NI;
                            lm = fmaxf(lm, c1[k]);
                        }
                        const int base = cta_max_int<G>(lm > -1e30f ? u + (int) floorf(lm) : CP_INT_MIN, sm_red);
                        const float ub = (float) (u - base);
#pragma unroll
                        for (int k = 0; k < K; k++) sm_fold[(xs[k] - blo) & NMASK] = c1[k] + ub;
                        cta_sync<G>();
                        if (tid < 32) { float t = warp_ordered_fold(sm_fold, N); if (tid == 0) sm_bcast[0] = t; }
                        cta_sync<G>();
                        const float t1 = sm_bcast[0];
                        float tot = t1;
                        if (d < Dt && base != CP_INT_MIN) {
                            // term 2: matches that jump over diagonal d     (impl/pairwiseAligner.c:744-752):
                            // a match-only forward step from F[d-1] into the cells of row d+1, dotted with B[d+1].
                            const int rp = rowB == 0 ? R - 1 : rowB - 1;
                            float gM[K], gX[K], gY[K];
                            const int toF1 = loadRow(rp, gM, gX, gY);
                            float nb[4] = { gM[K - 1], gX[K - 1], gY[K - 1], __int_as_float(toF1) };
                            pass_to_next<G, 4>(nb, sm_x, iter & 1); iter++;
                            const float dl = (float) (__float_as_int(nb[3]) - toF1);
                            const float ub2 = (float) (toF1 + to - base);
#pragma unroll
                            for (int k = 0; k < K; k++) {
                                const float mM = k > 0 ? gM[k - 1] : nb[0] + dl, mX = k > 0 ? gX[k - 1] : nb[1] + dl,
                                            mY = k > 0 ? gY[k - 1] : nb[2] + dl;
                                const float md = logadd(logadd(mM + P.tMC, mX + P.tMX), mY + P.tMY);
                                // vM[k] still holds G_M = B_M + match emission of the row d+1 cell of this slot
                                sm_fold[(xs[k] - blo) & NMASK] = (md + vM[k]) + ub2;
                            }
                            cta_sync<G>();
                            if (tid < 32) { float t = warp_ordered_fold(sm_fold, N); if (tid == 0) sm_bcast[1] = t; }
                            cta_sync<G>();
                            tot = logadd(t1, sm_bcast[1]);
                        }
                        totSt = tot;
                        totBase = base == CP_INT_MIN ? 0 : base;
                        if (!(tot > -1e30f)) status |= 2;
                        if (dbgTot != nullptr && tid == 0) {   // debug: the two terms, absolute
                            dbgTot[(D + 1) + d] = (double) t1 + (double) totBase;
                            dbgTot[2 * (D + 1) + d] = d < Dt ? (double) sm_bcast[1] + (double) totBase : (double) NAN;
                        }
                    }
                    const double totAbs = (double) totSt + (double) totBase;
                    if (d == D) lastTotal = totAbs;
                    if (dbgTot != nullptr && tid == 0) dbgTot[d] = totAbs;
                    // posteriors (impl/pairwiseAligner.c:768-793)
                    const float corr = (float) (u - totBase) - totSt;
                    int sc[K];
                    int any = 0;
#pragma unroll
                    for (int k = 0; k < K; k++) {
                        const int x = xs[k], y = d - x;
                        const float lp = (fM[k] + bM[k]) + corr;
                        float p = __expf(lp);
                        const bool ok = inb[k] && x > 0 && y > 0 && p >= P.threshold;
                        p = fminf(p, 1.0f);
                        sc[k] = ok ? (P.dbgLogP ? __float_as_int(lp) : (int) floorf(p * 10000000.0f)) : 0x7fffffff;
                        any |= ok;
                    }
                    if (cta_any<G>(any)) {
#pragma unroll
                        for (int k = 0; k < K; k++) sm_score[(xs[k] - blo) & NMASK] = sc[k];
                        cta_sync<G>();
                        if (tid < 32) {
                            for (int base = 0; base < N; base += 32) {
                                const int s = sm_score[base + lane];
                                const unsigned mask = __ballot_sync(CP_FULL, s != 0x7fffffff);
                                if (s != 0x7fffffff) {
                                    const int pos = nPairs + __popc(mask & ((1u << lane) - 1u));
                                    if (pos < it.pair_cap) {
                                        const int x = blo + base + lane;
                                        pairs[3 * pos] = s; pairs[3 * pos + 1] = x - 1; pairs[3 * pos + 2] = d - x - 1;
                                    }
                                }
                                nPairs += __popc(mask);
                            }
                        }
                        cta_sync<G>();
                    }
                }

                // G of this row becomes "row d+1" for the next step
#pragma unroll
                for (int k = 0; k < K; k++) {
                    wM[k] = vM[k]; wX[k] = vX[k]; wY[k] = vY[k];
                    vM[k] = inb[k] ? bM[k] + emitM(k) : NI;
                    vX[k] = inb[k] ? bX[k] + pc[k].z : NI;
                    vY[k] = inb[k] ? bY[k] + emitY(k) : NI;
                }
                mprev = mcur;
                mcur = max3K(bM, bX, bY);      // follow B itself: the emissions of unlikely cells are hugely negative
            }
            tracedBackTo = tracedBackFrom;
        }

        if (tid == 0) {
            ItemOut &o = A.out[itemIdx];
            o.n_pairs = nPairs;
            o.status = status | (nPairs > it.pair_cap ? 1 : 0);
            o.total_logprob = lastTotal;
            o.n_tracebacks = nTb;
        }
    }
}

}  // namespace cpecan
