// cpecan_align2.cuh -- the sm_100a kernel for the banded signal pair-HMM (three-state and vanilla machines):
// forward sweep, periodic traceback (backward values, local totals, posteriors or E-step sums), one launch per batch
// and ring-size bucket.  Reference: impl/pairwiseAligner.c:870-1006 (schedule), :681-795 and :841-863 (diagonal
// calculations), impl/stateMachine.c:1305-1334 and :1368-1409 (cell bodies).
//
// The mapping comes out of the ncu profiles kept under profiles/ (a CTA per alignment with x-owned register rows lost
// 60 % of its lanes outside the ragged band and sat on block barriers; a register-resident variant of this kernel was
// instruction-cache bound):
//
//   * ONE WARP per alignment, no block barriers.  The reference positions x are cut into CHUNKS of 32 consecutive x
//     placed relative to the band of the diagonal.  Per diagonal the warp loops (a run-time loop, one copy of the
//     code) over just the chunks that intersect the band, so the cost of a diagonal follows its band width.
//   * All per-cell state lives in a shared-memory ring of N positions (x mod N), two float4 (M, X, Y, offset) per
//     position: the values of diagonals d-1 and d-2.  A cell reads its own and its neighbour's entries (LDS.128) and
//     writes the new value over its own d-2 entry; walking the chunks in DESCENDING x on the way forward (ASCENDING
//     on the way back) makes that in-place update safe, so two buffers suffice.
//   * The k-mer side of the emissions (3 or 4 float4 column records) and the event are streamed from L1/L2 through a
//     software pipeline over the (diagonal, chunk) tasks: requested at the top of task i, reduced at its end to the
//     3 - 8 values task i+1 needs (see the forward sweep for why the loads stay a whole task ahead).
//   * Every cell carries its own offset (an integer stored as float): values are FP32 relative to it.  logAdd is
//     translation invariant, so a cell is computed in the units U = max(own, neighbours' offsets) and re-based to
//     its own maximum afterwards; nothing is lost at log-probabilities of -40 000 and nothing depends on where in
//     the diagonal the probability mass sits.
//   * logAdd: the reference's 4-segment cubic and 7.5 cut-off (impl/pairwiseAligner.c:235-255) as max + q(|x-y|),
//     q = cubic - identity; on the dependent chain the segment is chosen by predicated immediate-operand FFMAs, off
//     the chain its coefficients come from a shared-memory table with one LDS.128 (bit-identical results).
//   * Aligned pairs leave in the reference's order for free: the backward walk visits x ascending inside a diagonal.
#pragma once
#include "cpecan_kernels.cuh"

namespace cpecan {

#define CP_BIG 1.0e30f
#define CP_POS_INF (__int_as_float(0x7f800000))

// Multiplier coefficients of the logAdd segments, kept in registers for the whole kernel (an FFMA takes only one
// immediate; without this ptxas re-materialises the eight constants in every loop iteration).
struct LaCoef { float a4, c4, a3, c3, a2, c2, a1, c1; };
__device__ __forceinline__ LaCoef la_coef() {
    LaCoef k;
    asm volatile("mov.f32 %0, 0fB9F07885;" : "=f"(k.a4));
    asm volatile("mov.f32 %0, 0fBD8DDAF8;" : "=f"(k.c4));
    asm volatile("mov.f32 %0, 0fBB96E5CE;" : "=f"(k.a3));
    asm volatile("mov.f32 %0, 0fBE9BAB98;" : "=f"(k.c3));
    asm volatile("mov.f32 %0, 0fBC6E18FA;" : "=f"(k.a2));
    asm volatile("mov.f32 %0, 0fBF011E08;" : "=f"(k.c2));
    asm volatile("mov.f32 %0, 0fBC19343D;" : "=f"(k.a1));
    asm volatile("mov.f32 %0, 0fBF004EA8;" : "=f"(k.c1));
    return k;
}

// max(x, y) + q(|x - y|) for |x - y| < 7.5, else max(x, y); q = the reference's cubic segment minus the identity,
// evaluated as (a t + b) t^2 + ((c - 1) t + d).  NaN (both -inf) and +inf differences fall through to max.
__device__ __forceinline__ float logadd2p(float x, float y, const LaCoef &k) {
    float r;
    asm("{\n\t"
        ".reg .pred p1, p2, p3, p5;\n\t"
        ".reg .f32 d, a, a2, u, v;\n\t"
        "sub.f32 d, %1, %2;\n\t"
        "abs.f32 a, d;\n\t"
        "max.f32 %0, %1, %2;\n\t"
        "setp.le.f32 p3, a, 0f40900000;\n\t"            // 4.5
        "setp.le.f32 p2, a, 0f40200000;\n\t"            // 2.5
        "setp.le.f32 p1, a, 0f3F800000;\n\t"            // 1.0
        "setp.lt.f32 p5, a, 0f40F00000;\n\t"            // 7.5
        "mul.f32 a2, a, a;\n\t"
        "fma.rn.f32 u, a, %3, 0f3C1EDBBF;\n\t"          // (4.5, 7.5)
        "fma.rn.f32 v, a, %4, 0f3E2C11EF;\n\t"
        "@p3 fma.rn.f32 u, a, %5, 0f3D81E63C;\n\t"      // (2.5, 4.5]
        "@p3 fma.rn.f32 v, a, %6, 0f3F03A75F;\n\t"
        "@p2 fma.rn.f32 u, a, %7, 0f3E0F4D0A;\n\t"      // (1, 2.5]
        "@p2 fma.rn.f32 v, a, %8, 0f3F313020;\n\t"
        "@p1 fma.rn.f32 u, a, %9, 0f3E05CB9C;\n\t"      // [0, 1]
        "@p1 fma.rn.f32 v, a, %10, 0f3F3175C2;\n\t"
        "fma.rn.f32 u, u, a2, v;\n\t"
        "@p5 add.f32 %0, %0, u;\n\t"
        "}"
        : "=f"(r) : "f"(x), "f"(y), "f"(k.a4), "f"(k.c4), "f"(k.a3), "f"(k.c3), "f"(k.a2), "f"(k.c2), "f"(k.a1), "f"(k.c1));
    return r;
}

// logAdd coefficient table in shared memory: entry i serves |x - y| in ((i - 1) / 2, i / 2] -- the reference's segment
// bounds 1, 2.5, 4.5 and the cut-off 7.5 are all multiples of 1/2 and closed on the right, as ceil(2 |x - y|) is.
// An entry holds (a, b, c - 1, d) of the segment's cubic a t^3 + b t^2 + c t + d; entry 16 (beyond the cut-off) is zero.
#define CP_LAT_ENTRIES 17
#define CP_LAT_BYTES 288
__device__ __forceinline__ void la_table_init(float4 *tbl, int lane) {
    if (lane < CP_LAT_ENTRIES) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (lane <= 2) v = make_float4(__int_as_float(0xBC19343D), __int_as_float(0x3E05CB9C), __int_as_float(0xBF004EA8), __int_as_float(0x3F3175C2));
        else if (lane <= 5) v = make_float4(__int_as_float(0xBC6E18FA), __int_as_float(0x3E0F4D0A), __int_as_float(0xBF011E08), __int_as_float(0x3F313020));
        else if (lane <= 9) v = make_float4(__int_as_float(0xBB96E5CE), __int_as_float(0x3D81E63C), __int_as_float(0xBE9BAB98), __int_as_float(0x3F03A75F));
        else if (lane <= 15) v = make_float4(__int_as_float(0xB9F07885), __int_as_float(0x3C1EDBBF), __int_as_float(0xBD8DDAF8), __int_as_float(0x3E2C11EF));
        tbl[lane] = v;
    }
}

// max(x, y) + q(|x - y|) for |x - y| < 7.5, else max(x, y); q = the reference's cubic segment (impl/pairwiseAligner.c:
// 235-255) minus the identity, evaluated as (a t + b) t^2 + ((c - 1) t + d) with the segment's coefficients fetched by
// one LDS.128 (13 instructions; selecting them with predicated immediate FFMAs took 19, with a tree of selects more).
// NaN (both -inf) and +inf differences fall through to max: cvt gives 0 / INT_MAX for them and the add is predicated.
__device__ __forceinline__ float logadd2(float x, float y, unsigned tbl) {
    float r;
    asm("{\n\t"
        ".reg .pred p5;\n\t"
        ".reg .f32 d, a, a2, t, c3, c2, c1, c0, u, v;\n\t"
        ".reg .s32 i;\n\t"
        ".reg .u32 ad;\n\t"
        "sub.f32 d, %1, %2;\n\t"
        "abs.f32 a, d;\n\t"
        "max.f32 %0, %1, %2;\n\t"
        "setp.lt.f32 p5, a, 0f40F00000;\n\t"            // 7.5
        "add.f32 t, a, a;\n\t"
        "cvt.rpi.s32.f32 i, t;\n\t"
        "min.s32 i, i, 16;\n\t"
        "shl.b32 ad, i, 4;\n\t"
        "add.u32 ad, ad, %3;\n\t"
        "ld.shared.v4.f32 {c3, c2, c1, c0}, [ad];\n\t"
        "mul.f32 a2, a, a;\n\t"
        "fma.rn.f32 u, a, c3, c2;\n\t"
        "fma.rn.f32 v, a, c1, c0;\n\t"
        "fma.rn.f32 u, u, a2, v;\n\t"
        "@p5 add.f32 %0, %0, u;\n\t"
        "}"
        : "=f"(r) : "f"(x), "f"(y), "r"(tbl));
    return r;
}

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }

// column records + event of one (diagonal, chunk) task, as loaded for this lane
struct FwdRec { float4 a, b, c, d, ev; };
struct BwdRec { float4 a, b, c, dR, ev, F; };

struct KernelArgs2 {
    const Item *items;
    const int *order;
    int n_items;
    int *queue;
    const long long *anchors;
    const float4 *xparams;        // 4 float4 per matrix column x = 0 .. lX+1 (the last record is the all -inf dummy)
    const float4 *events;
    float4 *scratch;              // per warp: ring of forward rows, N float4 (M, X, Y, offset) each
    long long scratch_stride;     // float4 per warp
    int ring_rows;
    int ringN;                    // ring positions (power of two)
    int zero;                     // always 0, but only the host knows (see the register prefetch in k_align2)
    int *pairs;
    ItemOut *out;
    double *totals;
    double *expect;               // EXPECT: 9 transition sums, 4096 k-mer skip sums, 1 likelihood (batch totals)
    DevParams P;
};

__host__ __device__ inline size_t align2_smem_bytes(int ringN) { return (size_t) ringN * (2 * 16) + CP_LAT_BYTES; }

// Left fold  acc = logadd(acc, v[0]), logadd(acc, v[1]), ...  in ascending x (the order of dpDiagonal_dotProduct,
// impl/pairwiseAligner.c:587-597), fed 32 consecutive elements (one per lane, -inf where there is none) at a time,
// straight from the registers of the pass that produces them.
// Exactly equal to the serial fold, at the cost of the few elements near the ridge only:
//   * an element more than 7.5 below the running prefix maximum cannot change acc (acc >= prefix maximum);
//   * an element at least 7.5 + 12 above the prefix maximum RESETS the fold: acc <= prefix maximum + ln(count) +
//     count * 6e-4 (the cubic over-estimates by at most 5.5e-4 per step) < prefix maximum + 12 for count <= 4096, so
//     logadd returns the element itself and everything before it is forgotten.
struct OrderedFold {
    float acc, runmax;
    int base;                              // integer units of acc (block_units)
    __device__ __forceinline__ void reset() { acc = CP_NEG_INF; runmax = CP_NEG_INF; base = CP_INT_MIN; }

    __device__ __forceinline__ void block(float v, unsigned k) {
        const int lane = threadIdx.x & 31;
        float m = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { float t = __shfl_up_sync(CP_FULL, m, o); if (lane >= o) m = fmaxf(m, t); }
        float excl = __shfl_up_sync(CP_FULL, m, 1);
        excl = lane == 0 ? runmax : fmaxf(excl, runmax);
        unsigned mask = __ballot_sync(CP_FULL, v > excl - 7.6f);
        const unsigned resets = __ballot_sync(CP_FULL, v >= excl + 19.5f);
        if (resets) {
            const int j = 31 - __clz(resets);
            acc = __shfl_sync(CP_FULL, v, j);
            mask &= j == 31 ? 0u : ~((2u << j) - 1u);
        }
        while (mask) {
            const int b = __ffs(mask) - 1;
            mask &= mask - 1;
            acc = logadd2(acc, __shfl_sync(CP_FULL, v, b), k);
        }
        runmax = fmaxf(runmax, __shfl_sync(CP_FULL, m, 31));
    }

    // Elements c1 in their own integer units us (floats holding integers).  The fold keeps its accumulator in the
    // units of the largest element seen so far -- an exact integer shift when a block raises it -- so the values near
    // the ridge stay small whatever the offsets of the cells are, and nothing has to be parked until the diagonal's
    // maximum is known.
    __device__ __forceinline__ void block_units(float c1, float us, bool valid, unsigned k) {
        valid = valid && c1 > -1e30f;
        const int cm = __reduce_max_sync(CP_FULL, valid ? (int) us + (int) floorf(c1) : CP_INT_MIN);
        if (cm > base) {
            if (base != CP_INT_MIN) { const float sh = (float) (base - cm); acc += sh; runmax += sh; }
            base = cm;
        }
        block(valid ? c1 + (us - (float) base) : CP_NEG_INF, k);
    }
};

// re-base a cell to its own maximum (integer shift, exact); a cell that is all -inf drifts down by 64 per step
__device__ __forceinline__ void rebase(float &a, float &b, float &c, float &off) {
    const float m = fmaxf(fmaxf(a, b), fmaxf(c, -64.0f));
    const float s = (m + 12582912.0f) - 12582912.0f;
    a -= s; b -= s; c -= s; off += s;
}

#ifndef CP_MINB
#define CP_MINB 16
#endif
// MACH: 0 = three-state (strawMan) machine, 1 = vanilla machine (per-column transitions, inverse-Gaussian noise term)
template <int MACH, bool HAS_SX, bool EXPECT>
__global__ void __launch_bounds__(32, CP_MINB) k_align2(const KernelArgs2 A) {
    extern __shared__ __align__(16) unsigned char smraw[];
    const int N = A.ringN, NM = N - 1;
    float4 *ring = reinterpret_cast<float4 *>(smraw);             // 2 * N entries

    const int lane = threadIdx.x;
    const DevParams &P = A.P;
    const float NI = CP_NEG_INF;
    float4 *rows = A.scratch + (long long) blockIdx.x * A.scratch_stride;
    const int R = A.ring_rows;
    const float4 NIENT = make_float4(NI, NI, NI, -CP_BIG);
    // logAdd coefficient table behind the ring
    la_table_init(ring + 2 * N, threadIdx.x);
    __syncwarp();
    const unsigned K = (unsigned) __cvta_generic_to_shared(ring + 2 * N);
#define LA(a, b) logadd2((a), (b), K)
    const LaCoef KP = la_coef();
#define LAP(a, b) logadd2p((a), (b), KP)
    // three-state: global transitions; vanilla: M->X, X->X, M->M, X->M, M->Y come from the column records
    const float gMC = P.tMC, gMX = P.tMX, gOX = P.tOX, gOY = P.tOY, gEX = P.tEX, tSX = P.tSX;
    const float tMY = MACH ? P.vYM : P.tMY, tEY = MACH ? P.vYY : P.tEY;
    // emission of the match (Y = false) or extra-event (Y = true) state from a column record and an event record
    auto emit = [](const float4 pa, const float4 pb, const float4 pc, const float4 ev, bool Y) -> float {
        const float dm = ev.x - (Y ? pb.y : pa.x), dn = ev.y - (Y ? pb.w : pa.z);
        const float c1 = Y ? pb.z : pa.y, q = Y ? pc.x : pa.w, k0 = Y ? pc.y : pb.x;
        if (MACH) return fmaf(c1, dm * dm, fmaf(q * ev.z, dn * dn, k0)) + ev.w;
        return fmaf(c1, dm * dm, fmaf(q, dn * dn, k0));
    };

    for (;;) {
        int qi = 0;
        if (lane == 0) qi = atomicAdd(A.queue, 1);
        qi = __shfl_sync(CP_FULL, qi, 0);
        if (qi >= A.n_items) break;
        const int itemIdx = A.order[qi];
        const Item it = A.items[itemIdx];
        const int lX = it.lX, lY = it.lY, D = lX + lY;
        // the planes (a, b, c and, vanilla, d) of the column records, each lX + 2 float4 long: neighbouring lanes read
        // neighbouring 16-byte records, so a warp's LDG.128 touches 4-5 cache lines instead of 16
        const float4 *xpA = A.xparams + (MACH ? 4 : 3) * it.xp_off, *xpB = xpA + (lX + 2), *xpC = xpB + (lX + 2), *xpD = xpC + (lX + 2);
        const float4 *evp = A.events + it.ev_off;
        int *pairs = A.pairs + 3 * it.pair_off;
        double *dbgTot = (A.totals != nullptr && it.tot_off >= 0) ? A.totals + it.tot_off : nullptr;
        int nPairs = 0, status = 0, nTb = 0;
        double lastTotal = 0.0;
        const bool unbanded = P.mode == 2;
        const float logThrLo = __logf(P.threshold) - 1e-3f;
        double eT[9] = { 0, 0, 0, 0, 0, 0, 0, 0, 0 };            // EXPECT: this lane's share of the transition sums
        double eLik = 0.0;      // pre-filter; the exact test is p >= threshold

        if (D == 0) {
            if (lane == 0) { ItemOut &o = A.out[itemIdx]; o.n_pairs = 0; o.status = 0; o.total_logprob = 0.0; o.n_tracebacks = 0; }
            continue;
        }
        BandWalker bw;
        bw.init(A.anchors + 2 * it.an_off, it.nA, lX, lY, P.expansion);

        auto resetRing = [&]() {
            __syncwarp();
            for (int i = lane; i < 2 * N; i += 32) ring[i] = NIENT;
            __syncwarp();
        };

        // ---- diagonal 0: the single cell (0,0) holds the start vector (impl/pairwiseAligner.c:897-898) ----------
        resetRing();
        float4 *A1 = ring, *A2 = ring + N;       // A1: newest diagonal, A2: the one before (forward); mirrored backward
        if (lane == 0) {
            const float *sv = (it.flags & 1) ? P.rstartv : P.startv;
            const float4 e = make_float4(sv[0], sv[1], sv[2], 0.f);
            A1[0] = e; rows[0] = e;
        }
        __syncwarp();

        int dcur = 0, tracedBackTo = 0;
        int lo = 0, hi = 0;

        while (tracedBackTo < D) {
            // =============================== forward sweep ===============================================
            int Dt = -1;
            bool atEnd = false;
            {
                // The (diagonal, chunk) tasks are walked as ONE loop with the inputs of the next task (k-mer record,
                // event) requested before the current one is computed, so global-load latency hides behind a full
                // chunk of arithmetic, also across the step to the next diagonal.
                int rowF = dcur % R;
                int d = dcur + 1;
                {
                    const int plo = lo, phi = hi;
                    bw.range(d, lo, hi);
                    if (lo < plo || lo > plo + 1 || hi < phi || hi > phi + 1) status |= 4;
                }
                // lanes are placed relative to the window [lo-1, hi+1] of the diagonal (the cells one outside the band
                // are written as -inf for the neighbours that will read them): chunk j holds x = wlo + 32 j + lane
                int wlo = max(lo - 1, 0), c = (min(hi + 1, lX) - wlo) >> 5;     // any start: 8 consecutive float4 are 32 distinct banks
                rowF = rowF + 1 == R ? 0 : rowF + 1;
                float4 *frow = rows + (long long) rowF * N;
                // Software pipeline over the tasks: the column records and the event of task i+1 are requested at the
                // top of task i and consumed at its END, where they are reduced to the emissions (and, vanilla, the
                // transitions) of task i+1 -- 3 to 8 values carried over instead of 14 to 22 record registers, and no
                // register copies.  What keeps the loads a whole task ahead of their use:
                //   * they are issued before the warp barrier -- ptxas does not move loads across it, and left to itself
                //     it sinks them to the end of the iteration (shorter live ranges under the 128-register cap);
                //   * nothing is in flight when they are issued (ptxas gives all of them the same scoreboard, a counter,
                //     so a wait for an older load would wait for these too);
                //   * every component of a record is read (a dead one -- the k-mer index, an E-step input -- is OR-ed
                //     in under a mask only the host knows to be zero), else ptxas re-uses its register while the
                //     LDG.128 is in flight and the write-after-write hazard waits out the load.
                // (Tried and measured worse: register double buffering, 14 MOVs per task and ptxas places the copy of an
                // early-dying component right behind the load; the same unrolled by two, whose body falls out of the
                // instruction caches, no_instruction 0.24 -> 0.81 per issue; prefetch.global.L1 + late loads, -4 %.)
                FwdRec r;
                auto load = [&](int dd, int xbase) {
                    const int x = xbase + lane;
                    const int xx = min(x, lX + 1);
                    r.a = xpA[xx]; r.b = xpB[xx]; r.c = xpC[xx];
                    if (MACH) r.d = xpD[xx];
                    r.ev = evp[min(max(dd - x, 0), lY)];
                };
                struct { float eM, eY, eX, tOX, tEX, tMC, tMX, tOY; } E;
                auto reduce = [&]() {
                    E.eM = emit(r.a, r.b, r.c, r.ev, false); E.eY = emit(r.a, r.b, r.c, r.ev, true);
                    const float cz = __int_as_float(__float_as_int(r.c.z) | (__float_as_int(r.c.w) & A.zero));
                    E.eX = MACH ? 0.f : cz;
                    // impl/stateMachine.c:1368-1409: the vanilla transitions are those of THIS column
                    E.tOX = MACH ? r.d.x : gOX; E.tEX = MACH ? r.d.y : gEX; E.tMC = MACH ? r.d.z : gMC;
                    E.tMX = MACH ? r.d.w : gMX; E.tOY = MACH ? cz : gOY;
                };
                // one (diagonal, chunk) task; true when the sweep ends
                auto fstep = [&]() -> bool {
                    const bool last = c == 0;
                    int nd = d, nc = c - 1, nlo = lo, nhi = hi, nwlo = wlo;
                    bool stop = false;
                    if (last) {
                        const bool tbPoint = !unbanded && d >= tracedBackTo + P.minDiags && (hi - lo + 1) <= 2 * P.expansion + 1;
                        stop = d == D || tbPoint;
                        if (!stop) {
                            nd = d + 1;
                            bw.range(nd, nlo, nhi);
                            if (nlo < lo || nlo > lo + 1 || nhi < hi || nhi > hi + 1) status |= 4;
                            nwlo = max(nlo - 1, 0);
                            nc = (min(nhi + 1, lX) - nwlo) >> 5;
                        }
                    }
                    {
                        const int x = wlo + (c << 5) + lane;
                        const int s = x & NM, sl = (x - 1) & NM;
                        load(nd, nwlo + (max(nc, 0) << 5));          // next task (harmless clamped addresses at the stop)
                        const float4 own = A1[s], L = A1[sl], Mi = A2[sl];
                        const bool inb = x >= lo && x <= hi;
                        __syncwarp();
                        const float U = fmaxf(own.w, fmaxf(L.w, Mi.w));
                        // impl/stateMachine.c:1314-1333: transitions folded in code order, emission added once
                        float tX = LA(L.x + E.tOX, L.y + E.tEX);
                        if (HAS_SX) tX = LA(tX, L.z + tSX);
                        float tM = LAP(LAP(Mi.x + E.tMC, Mi.y + E.tMX), Mi.z + tMY);
                        float tY = LA(own.x + E.tOY, own.z + tEY);
                        const float Um = inb ? U : CP_POS_INF;       // a cell outside the band comes out as -inf
                        float cM = tM + (E.eM + (Mi.w - Um)), cX = tX + (E.eX + (L.w - Um)), cY = tY + (E.eY + (own.w - Um));
                        float co = inb ? U : -CP_BIG;
                        rebase(cM, cX, cY, co);
                        const float4 e = make_float4(cM, cX, cY, co);
                        A2[s] = e;                                   // descending x: in place over the d-2 entry
                        if (inb) frow[s] = e;
                        __syncwarp();
                        reduce();
                    }
                    if (last) {
                        { float4 *t = A1; A1 = A2; A2 = t; }
                        dcur = d;
                        if (stop) { Dt = d; atEnd = d == D; return true; }
                        d = nd; lo = nlo; hi = nhi; wlo = nwlo;
                        if ((d & 15) == 0) {
                            // columns / events that enter the band during the next diagonals: first touch comes from
                            // DRAM, so pull them into L2 well ahead (one lane stalling stalls the warp)
                            const int px = hi + 32 + lane, py = (d - lo) + 32 + lane;
                            if (px <= lX + 1) { prefetch_l2(xpA + px); prefetch_l2(xpB + px); prefetch_l2(xpC + px); if (MACH) prefetch_l2(xpD + px); }
                            if (py <= lY) prefetch_l2(evp + py);
                        }
                        rowF = rowF + 1 == R ? 0 : rowF + 1;
                        frow = rows + (long long) rowF * N;
                    }
                    c = nc;
                    return false;
                };
                load(d, wlo + (c << 5));
                reduce();
                while (!fstep()) { }
            }

            // =============================== traceback ===================================================
            // impl/pairwiseAligner.c:920-992.  The ring now holds G = B + emission of diagonals d+1 (A1) and d+2 (A2).
            nTb++;
            const int tracedBackFrom = Dt - (atEnd ? 0 : P.tbDiags + 1);
            BandWalker bb = bw;
            int blo = lo, bhi = hi;
            resetRing();
            const float *endv = (atEnd && (it.flags & 2)) ? P.rendv : P.endv;
            const float endM = endv[0], endX = endv[1], endY = endv[2];
            float totSt = NI, totBase = 0.f;
            int tillTotal = 0;
            {
                int rowB = Dt % R;
                for (int d = Dt; d > tracedBackTo; d--) {
                    if (d < Dt) {
                        const int pl = blo, ph = bhi;
                        bb.range(d, blo, bhi);
                        if (blo > pl || blo < pl - 1 || bhi > ph || bhi < ph - 1) status |= 4;
                        rowB = rowB == 0 ? R - 1 : rowB - 1;
                    }
                    const float4 *frow = rows + (long long) rowB * N;
                    const bool post = d <= tracedBackFrom;
                    const int wlo = max(blo - 1, 0), nch = ((min(bhi + 1, lX) - wlo) >> 5) + 1;
                    const int plo = max(blo, 1), phi = min(bhi, d - 1);   // cells with x > 0 and y > 0 report posteriors
                    bool doTotal = false;
                    if (post) { doTotal = unbanded ? (d == Dt) : (tillTotal == 0); tillTotal = tillTotal == 0 ? P.totalEvery - 1 : tillTotal - 1; }   // every totalEvery-th posterior diagonal, from the first
                    // The same software pipeline as in the forward sweep: the records of the next chunk are requested at
                    // the top of a chunk and reduced at its end to what the cell needs (emissions, the forward cell's match
                    // value and units, the k-mer index for the E-step, the vanilla transitions).
                    BwdRec q;
                    struct { float eM, eY, eX, Fx, Fw, tOX, tEX, tMC, tMX, myLog; int kw; } G;
                    auto loadB = [&](int cc) {
                        const int x = wlo + (cc << 5) + lane;
                        const int xx = min(x, lX + 1);
                        q.a = xpA[xx]; q.b = xpB[xx]; q.c = xpC[xx];
                        if (MACH) q.dR = xpD[min(x + 1, lX + 1)];
                        q.ev = evp[min(max(d - x, 0), lY)];
                        q.F = NIENT;
                        if (post && x >= blo && x <= bhi) q.F = frow[x & NM];
                    };
                    auto reduceB = [&]() {
                        G.eM = emit(q.a, q.b, q.c, q.ev, false); G.eY = emit(q.a, q.b, q.c, q.ev, true);
                        G.kw = __float_as_int(q.c.w);
                        const float cz = EXPECT ? q.c.z : __int_as_float(__float_as_int(q.c.z) | (G.kw & A.zero));   // see the forward sweep
                        G.eX = MACH ? 0.f : cz; G.myLog = cz;
                        G.tOX = MACH ? q.dR.x : 0.f; G.tEX = MACH ? q.dR.y : 0.f; G.tMC = MACH ? q.dR.z : 0.f; G.tMX = MACH ? q.dR.w : 0.f;
                        G.Fx = q.F.x; G.Fw = q.F.w;
                    };
                    // the first chunk's records are requested here, ahead of the per-diagonal bookkeeping below
                    if (!doTotal) loadB(0);
                    if (d - 2 > tracedBackTo && d - 2 <= tracedBackFrom) {
                        // the forward cells of diagonal d-2 were written >= 1000 diagonals ago: DRAM -> L2 now
                        const int rowP = rowB >= 2 ? rowB - 2 : rowB - 2 + R;
                        const float4 *fp = rows + (long long) rowP * N;
                        // one lane per 128-byte line (8 cells): a single instruction covers 256 cells of the window
                        const int p0 = max(wlo - 32, 0), pw = (nch + 1) << 5;
                        for (int o = lane << 3; o < pw; o += 256) prefetch_l2(fp + ((p0 + o) & NM));
                    }
                    if ((d & 15) == 0) {
                        const int px = blo - 32 - lane, py = (d - bhi) - 32 - lane;
                        if (px >= 0) { prefetch_l2(xpA + px); prefetch_l2(xpB + px); prefetch_l2(xpC + px); if (MACH) prefetch_l2(xpD + px); }
                        if (py >= 1) prefetch_l2(evp + py);
                    }

                    // E-step inputs: forward cells of the three predecessors live on diagonals d-1 and d-2
                    int el1 = 0, eh1 = -1, el2 = 0, eh2 = -1;
                    const float4 *frow1 = frow, *frow2 = frow;
                    float aT[9] = { 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f };
                    if (EXPECT && post) {
                        BandWalker be = bb;
                        if (d >= 1) be.range(d - 1, el1, eh1);
                        // the reference has freed the forward diagonals below tracedBackTo by now (impl/pairwiseAligner.c:
                        // 971-985), so on the first diagonal of a later traceback the match predecessors are absent
                        if (d >= 2 && d - 2 >= tracedBackTo) be.range(d - 2, el2, eh2);
                        const int r1 = rowB == 0 ? R - 1 : rowB - 1, r2 = r1 == 0 ? R - 1 : r1 - 1;
                        frow1 = rows + (long long) r1 * N; frow2 = rows + (long long) r2 * N;
                    }

                    // B of one cell from the ring (pull form of impl/pairwiseAligner.c:378-383: first from diagonal
                    // d+2 as "middle", then d+1 in ascending x-y as "upper", then "lower"), in units U
                    // a cell outside the band comes out as -inf (its inputs are converted to units +inf)
                    // (vanilla: the transitions INTO a successor cell are those of the successor's column: pdR = record d
                    // of column x+1; M->Y uses this column's log a_my)
                    auto cellB = [&](int x, int s, bool inb, const float4 pdR, float myLog, float &bM, float &bX, float &bY, float &U) {
                        if (d == Dt) { bM = inb ? endM : NI; bX = inb ? endX : NI; bY = inb ? endY : NI; U = 0.f; return; }
                        const float tOX = MACH ? pdR.x : gOX, tEX = MACH ? pdR.y : gEX, tMC = MACH ? pdR.z : gMC,
                                    tMX = MACH ? pdR.w : gMX, tOY = MACH ? myLog : gOY;
                        const int sr = (x + 1) & NM;
                        const float4 own = A1[s], R1 = A1[sr], R2 = A2[sr];
                        U = fmaxf(own.w, fmaxf(R1.w, R2.w));
                        const float Um = inb ? U : CP_POS_INF;
                        const float gm2 = R2.x + (R2.w - Um), gx1 = R1.y + (R1.w - Um), gy1 = own.z + (own.w - Um);
                        bM = LAP(LAP(gm2 + tMC, gy1 + tOY), gx1 + tOX);
                        bX = LA(gm2 + tMX, gx1 + tEX);
                        bY = LA(gm2 + tMY, gy1 + tEY);
                        if (HAS_SX) bY = LA(bY, gx1 + tSX);
                    };
                    // posterior of one cell + G = B + emission, re-based, back into the ring
                    auto cellPost = [&](int x, int s, bool inb, float bM, float bX, float bY, float U, float eM, float eY,
                                        float eX, float Fx, float Fw, float myLog, int kw) {
                        if (EXPECT) {
                            if (post) {
                                // diagonalCalculation_Expectations (impl/pairwiseAligner.c:841-863): for every transition
                                // into this cell p = exp(F_pred[from] + B[to] + eP + tP - total) (:426-443)
                                float4 FL = NIENT, FM = NIENT, FU = NIENT;
                                if (inb) {
                                    if (x - 1 >= el1 && x - 1 <= eh1) FL = frow1[(x - 1) & NM];
                                    if (x - 1 >= el2 && x - 1 <= eh2) FM = frow2[(x - 1) & NM];
                                    if (x >= el1 && x <= eh1) FU = frow1[s];
                                }
                                float4 pdo = NIENT;
                                if (MACH) pdo = xpD[min(x, lX + 1)];
                                const float tOX = MACH ? pdo.x : gOX, tEX = MACH ? pdo.y : gEX, tMC = MACH ? pdo.z : gMC,
                                            tMX = MACH ? pdo.w : gMX, tOY = MACH ? myLog : gOY;
                                const float kX = (bX + eX) + (((FL.w + U) - totBase) - totSt);
                                const float pMX = __expf(FL.x + tOX + kX), pXX = __expf(FL.y + tEX + kX);
                                if (MACH) {
                                    // cell_signal_updateBetaAndAlphaProb (impl/pairwiseAligner.c:478-498): skip bins only
                                    const int bin = kw;
                                    if (bin >= 0) {
                                        if (pMX > 1e-13f) atomicAdd(A.expect + bin, (double) pMX);
                                        if (pXX > 1e-13f) atomicAdd(A.expect + 30 + bin, (double) pXX);
                                    }
                                } else {
                                    const float kM = (bM + eM) + (((FM.w + U) - totBase) - totSt);
                                    const float kY = (bY + eY) + (((FU.w + U) - totBase) - totSt);
                                    const float pYX = HAS_SX ? __expf(FL.z + tSX + kX) : 0.f;
                                    aT[1] += pMX; aT[4] += pXX; aT[7] += pYX;             // from * 3 + to, to = X (1)
                                    aT[0] += __expf(FM.x + tMC + kM); aT[3] += __expf(FM.y + tMX + kM); aT[6] += __expf(FM.z + tMY + kM);
                                    aT[2] += __expf(FU.x + tOY + kY); aT[8] += __expf(FU.z + tEY + kY);
                                    const float pk = pMX + pXX + pYX;
                                    const int kmer = kw;
                                    if (pk > 1e-13f && kmer >= 0) atomicAdd(A.expect + 9 + kmer, (double) pk);
                                }
                            }
                        } else
                        if (post) {
                            // impl/pairwiseAligner.c:768-793; exp only for the cells that can reach the threshold
                            const float lp = (Fx + bM) + (((Fw + U) - totBase) - totSt);
                            bool ok = x >= plo && x <= phi && lp >= logThrLo;
                            float p = 0.f;
                            if (ok) { p = __expf(lp); ok = p >= P.threshold; }
                            const unsigned mask = __ballot_sync(CP_FULL, ok);
                            if (ok) {
                                const int pos = nPairs + __popc(mask & ((1u << lane) - 1u));
                                if (pos < it.pair_cap) {
                                    p = fminf(p, 1.0f);
                                    pairs[3 * pos] = P.dbgLogP ? __float_as_int(lp) : (int) floorf(p * 10000000.0f);
                                    pairs[3 * pos + 1] = x - 1; pairs[3 * pos + 2] = d - x - 1;
                                }
                            }
                            nPairs += __popc(mask);
                        }
                        float gM = bM + eM, gX = bX + eX, gY = bY + eY, go = inb ? U : -CP_BIG;   // b is -inf outside the band
                        rebase(gM, gX, gY, go);
                        A2[s] = make_float4(gM, gX, gY, go);
                    };

                    if (!doTotal) {
                        reduceB();
                        for (int c = 0; c < nch; c++) {                // ascending x: in-place update of the d+2 entries
                            const int x = wlo + (c << 5) + lane, s = x & NM;
                            const bool inb = x >= blo && x <= bhi;
                            const auto cur = G;
                            loadB(min(c + 1, nch - 1));                // the next chunk (after the last one: again this one)
                            float bM, bX, bY, U;
                            cellB(x, s, inb, make_float4(cur.tOX, cur.tEX, cur.tMC, cur.tMX), cur.myLog, bM, bX, bY, U);
                            __syncwarp();
                            cellPost(x, s, inb, bM, bX, bY, U, cur.eM, cur.eY, cur.eX, cur.Fx, cur.Fw, cur.myLog, cur.kw);
                            __syncwarp();
                            reduceB();
                        }
                    } else {
                        // ---- totalProbability (impl/pairwiseAligner.c:736-754), recomputed every 10th posterior diagonal
                        // pass 1: B into the ring (units in .w), dot-product terms and their units aside
                        OrderedFold fold1;
                        fold1.reset();
                        float4 Fn = frow[(wlo + lane) & NM];           // forward cell of chunk 0 (used inside the band only)
                        for (int c = 0; c < nch; c++) {
                            const int x = wlo + (c << 5) + lane, s = x & NM;
                            const bool inb = x >= blo && x <= bhi;
                            const float4 F = Fn;
                            Fn = frow[(x + (c + 1 < nch ? 32 : 0)) & NM];                  // next chunk's, a chunk ahead of its use
                            float bM, bX, bY, U;
                            float4 pdR = NIENT;
                            float myLog = 0.f;
                            if (MACH) { pdR = xpD[min(x + 1, lX + 1)]; myLog = xpC[min(x, lX + 1)].z; }
                            cellB(x, s, inb, pdR, myLog, bM, bX, bY, U);
                            __syncwarp();
                            A2[s] = make_float4(bM, bX, bY, inb ? U : -CP_BIG);
                            // dot product of F and B over the diagonal, folded as the cells come (ascending x)
                            float c1 = NI, us = 0.f;
                            if (inb) { c1 = LA(LA(F.x + bM, F.y + bX), F.z + bY); us = F.w + U; }
                            fold1.block_units(c1, us, inb, K);
                            __syncwarp();
                        }
                        loadB(0);                                      // pass 2's first chunk: in flight during the second term
                        const int base = fold1.base;
                        const float fbase = base == CP_INT_MIN ? 0.f : (float) base;
                        const float t1 = fold1.acc;
                        float tot = t1, t2v = NI;
                        if (d < Dt && base != CP_INT_MIN) {
                            // term 2: matches jumping over diagonal d = a match-only forward step from F[d-1] into the
                            // cells of diagonal d+1, dotted with B[d+1] (G_M of d+1 is still in A1)
                            BandWalker b2w = bb;
                            int l1, h1, lm1, hm1;
                            b2w.range(d + 1, l1, h1);
                            b2w.range(d - 1, lm1, hm1);
                            const int rowM = rowB == 0 ? R - 1 : rowB - 1;
                            const float4 *fprev = rows + (long long) rowM * N;
                            OrderedFold fold2;
                            fold2.reset();
                            float4 Fp = fprev[(l1 + lane - 1) & NM];
                            for (int xb = l1; xb <= h1; xb += 32) {
                                const int x = xb + lane;
                                float val = NI;
                                const float4 F = Fp;
                                Fp = fprev[(x + 31) & NM];             // the next round's, a round ahead of its use
                                if (x <= h1 && x - 1 >= lm1 && x - 1 <= hm1) {
                                    const float4 Gn = A1[x & NM];
                                    float4 pdx = NIENT;
                                    if (MACH) pdx = xpD[min(x, lX + 1)];
                                    const float md = LA(LA(F.x + (MACH ? pdx.z : gMC), F.y + (MACH ? pdx.w : gMX)), F.z + tMY);
                                    val = (md + Gn.x) + ((F.w + Gn.w) - fbase);
                                }
                                fold2.block(val, K);
                            }
                            t2v = fold2.acc;
                            tot = LA(t1, t2v);
                        }
                        totSt = tot;
                        totBase = fbase;
                        if (!(tot > -1e30f)) status |= 2;
                        if (dbgTot != nullptr && lane == 0) {
                            dbgTot[(D + 1) + d] = (double) t1 + (double) totBase;
                            dbgTot[2 * (D + 1) + d] = d < Dt ? (double) t2v + (double) totBase : (double) NAN;
                        }
                        __syncwarp();
                        // pass 2: posteriors and G from the parked B
                        reduceB();
                        for (int c = 0; c < nch; c++) {
                            const int x = wlo + (c << 5) + lane, s = x & NM;
                            const bool inb = x >= blo && x <= bhi;
                            const float4 b = A2[s];
                            const auto cur = G;
                            loadB(min(c + 1, nch - 1));
                            __syncwarp();
                            cellPost(x, s, inb, b.x, b.y, b.z, b.w, cur.eM, cur.eY, cur.eX, cur.Fx, cur.Fw, cur.myLog, cur.kw);
                            __syncwarp();
                            reduceB();
                        }
                    }
                    if (post && (EXPECT || d == D || dbgTot != nullptr)) {     // FP64 only where somebody reads it
                        const double totAbs = (double) totSt + (double) totBase;
                        if (EXPECT) {
#pragma unroll
                            for (int i = 0; i < 9; i++) eT[i] += (double) aT[i];
                            eLik += totAbs;                       // "a hack": once per diagonal (:852-857)
                        }
                        if (d == D) lastTotal = totAbs;
                        if (dbgTot != nullptr && lane == 0) dbgTot[d] = totAbs;
                    }
                    { float4 *t = A1; A1 = A2; A2 = t; }
                }
            }
            tracedBackTo = tracedBackFrom;

            // =============================== restore the forward state at Dt ==============================
            if (tracedBackTo < D) {
                resetRing();
                A1 = ring; A2 = ring + N;
                int l1, h1;
                BandWalker br = bw;
                br.range(Dt - 1, l1, h1);
                br.range(Dt, lo, hi);
                const float4 *f0 = rows + (long long) (Dt % R) * N;
                const float4 *f1 = rows + (long long) ((Dt - 1) % R) * N;
                for (int x = lo + lane; x <= hi; x += 32) A1[x & NM] = f0[x & NM];
                for (int x = l1 + lane; x <= h1; x += 32) A2[x & NM] = f1[x & NM];
                __syncwarp();
            }
        }

        if (EXPECT) {
#pragma unroll
            for (int i = 0; i < (MACH ? 0 : 9); i++) {
                double v = eT[i];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(CP_FULL, v, o);
                if (lane == 0) atomicAdd(A.expect + i, v);
            }
            if (lane == 0) atomicAdd(A.expect + (MACH ? 60 : 9 + 4096), eLik);
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

#undef LA
}  // namespace cpecan
