// cpecan_align3.cuh -- the sm_100a kernel for the banded signal pair-HMM (three-state and vanilla machines), second
// generation: forward sweep, periodic traceback (backward values, local totals, posteriors or E-step sums), one launch
// per batch and ring-size bucket.  Reference: impl/pairwiseAligner.c:870-1006 (schedule), :681-795 and :841-863
// (diagonal calculations), impl/stateMachine.c:1305-1334 and :1368-1409 (cell bodies).
//
// Mapping (kept from k_align2, whose ncu profiles are under profiles/r1_*): ONE WARP per alignment, no block barriers;
// per diagonal a run-time loop over the chunks of 32 consecutive x that intersect the band; the cells of the last two
// diagonals in a shared-memory ring of float4 (M, X, Y, offset) -- every cell carries its own integer offset, values
// are FP32 relative to it; logAdd = the reference's 4-segment cubic, 7.5 cut-off.
//
// What is new, each item out of the round-1 profile (16.4 warp-instructions per band cell, 14 % of them per-diagonal
// bookkeeping, the backward pass re-deriving every emission):
//   * the band comes from the plan kernel as 2 bits per diagonal (lo / hi advance) and the traceback points as a
//     list: no anchor arithmetic, no box walking, no sanity checks inside the sweeps (k_plan3 does them once);
//   * the forward sweep stores, per band cell, (F_match, offset, e_match, e_extra-event): the backward sweep reads
//     ONE float4 per cell and evaluates no Gaussian and loads no column record except the gap emission;
//     the other two forward states are stored only on the diagonals that need them (every 10th posterior diagonal and
//     its predecessor for the total probability, the 41 + 2 diagonals around a traceback point); the E-step keeps
//     all three states and stores the two emissions in a second plane;
//   * cells outside the band are produced by indexing the all -inf dummy column record, not by per-cell selects;
//   * the logAdd table index is a round-up FFMA with a magic constant (no F2I, no clamp, no shift);
//   * E-step: k-mer skip sums accumulate per COLUMN in shared memory and leave with one atomic when the column leaves
//     the band (was: one FP64 global atomic per band cell).
#pragma once
#include "cpecan_kernels.cuh"
#include "cpecan_logadd.cuh"

namespace cpecan {

// max(x, y) + q(min(|x - y|, 8)): the table of logadd2 (17 entries, entry i serves |x - y| in ((i-1)/2, i/2], entry 16
// all zero) indexed by ceil(2 |x - y|) taken from the mantissa of a round-up FFMA; kadj = table address - 16 * 0x4B000000.
// NaN (both -inf) and +inf differences clamp to 8 (fminf returns the number), entry 16 adds exactly 0.
__device__ __forceinline__ float la3(float x, float y, unsigned kadj) {
    const float m = fmaxf(x, y);
    const float a = fminf(fabsf(x - y), 8.0f);
    const float u = __fmaf_ru(a, 2.0f, 8388608.0f);
    const unsigned ad = (unsigned) __float_as_int(u) * 16u + kadj;
    float c3, c2, c1, c0;
    asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(c3), "=f"(c2), "=f"(c1), "=f"(c0) : "r"(ad));
    const float p = fmaf(a, c3, c2), q = fmaf(a, c1, c0);
    return m + fmaf(p, a * a, q);
}

// a whole ring entry with ONE LDS.128 (left to itself ptxas loads the two or three components that are used with scalar
// LDS at a 16-byte stride: 4-way bank conflicts, 20 wavefronts instead of 12 per backward chunk)
__device__ __forceinline__ float4 lds4(const float4 *p) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "r"((unsigned) __cvta_generic_to_shared(p)));
    return v;
}

struct KernelArgs3 {
    const Item *items;
    const int *order;
    int n_items;
    int *queue;
    const float4 *xparams;        // planes of lX + 2 float4 per item (record lX+1 is the all -inf dummy)
    const float4 *events;
    const unsigned *bits;         // plan: 2 bits per diagonal (bit 0: lo advances, bit 1: hi advances), item at Item.pad0
    const int *tbs;               // plan: the diagonals at which the reference traces back, item at Item.pad1
    const int *flags;             // plan / preparation flags per item (CPECAN_ITEM_* bits): flagged items are skipped
    float4 *scratch;              // per warp: ring_rows rows of N float4, then spec_rows rows of N float2
    long long scratch_stride;     // float4 per warp
    int ring_rows, spec_rows;
    int all_spec;                 // POSTERIOR: every forward row keeps its X, Y states (traceback points closer together than a zone)
    int ringN;                    // ring positions (power of two)
    int zero;                     // always 0, but only the host knows (keeps a dead record component "read")
    int *pairs;
    ItemOut *out;
    double *totals;
    double *expect;               // EXPECT: 9 transition sums, 4096 k-mer skip sums, 1 likelihood (batch totals)
    DevParams P;
};

// CP_TMA_ROWS (an A/B build, off by default -- profiles/r2_tma_rows_ab.txt): the traceback fetches the forward row of the
// NEXT diagonal with cp.async.bulk (TMA, one instruction from one lane, completion on an mbarrier) into a shared-memory
// stage instead of one LDG.128 per cell; costs 2 N float4 of shared memory per warp.
#ifdef CP_TMA_ROWS
#define CP_TMA_STAGE_BYTES(ringN) ((size_t) (ringN) * 32 + 16)
#else
#define CP_TMA_STAGE_BYTES(ringN) ((size_t) 0)
#endif
__host__ __device__ inline size_t align3_smem_bytes(int ringN, bool colAcc) {
    return (size_t) ringN * (2 * 16) + CP_LAT_BYTES + (colAcc ? (size_t) ringN * 4 : 0) + CP_TMA_STAGE_BYTES(ringN);
}

#ifdef CP_TMA_ROWS
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned phase) {
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra WAIT_%=;\n\t}"
                 :: "r"(bar), "r"(phase) : "memory");
}
#endif

// forward rows are written once and read once, ~1000 diagonals later: keep them out of L1 (evict-first), which the
// column records and events re-read on every diagonal need
#ifdef CP_ROWS_PLAIN
#define ROWST4(p, v) (*(p) = (v))
#define ROWST2(p, v) (*(p) = (v))
#define ROWLD4(p) (*(p))
#define ROWLD2(p) (*(p))
#else
#define ROWST4(p, v) __stcs((p), (v))
#define ROWST2(p, v) __stcs((p), (v))
#define ROWLD4(p) __ldcs(p)
#define ROWLD2(p) __ldcs(p)
#endif

#ifndef CP_MINB3
#define CP_MINB3 16
#endif

// MACH: 0 = three-state (strawMan) machine, 1 = vanilla machine (per-column transitions, inverse-Gaussian noise term)
template <int MACH, bool HAS_SX, bool EXPECT>
__global__ void __launch_bounds__(32, CP_MINB3) k_align3(const KernelArgs3 A) {
    extern __shared__ __align__(16) unsigned char smraw[];
    const int N = A.ringN, NM = N - 1;
    float4 *ring = reinterpret_cast<float4 *>(smraw);             // 2 * N entries
    float *colAcc = reinterpret_cast<float *>(smraw + (size_t) N * 32 + CP_LAT_BYTES);   // EXPECT: N floats (three-state: per column; vanilla: the 60 skip bins)
#ifdef CP_TMA_ROWS
    float4 *stage = reinterpret_cast<float4 *>(smraw + (size_t) N * 32 + CP_LAT_BYTES + (EXPECT ? (size_t) N * 4 : 0));   // 2 x N
    const unsigned stageAddr = (unsigned) __cvta_generic_to_shared(stage);
    const unsigned barAddr = stageAddr + (unsigned) N * 32;
    if (threadIdx.x == 0) { mbar_init(barAddr, 1); mbar_init(barAddr + 8, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    unsigned tmaPhase0 = 0, tmaPhase1 = 0;
#endif

    const int lane = threadIdx.x;
    const DevParams &P = A.P;
    const float NI = CP_NEG_INF;
    float4 *rows = A.scratch + (long long) blockIdx.x * A.scratch_stride;
    const int R = A.ring_rows;
    float2 *plane2 = reinterpret_cast<float2 *>(rows + (long long) R * N);     // POSTERIOR: special rows; EXPECT: (eM, eY) of every row
    const float4 NIENT = make_float4(NI, NI, NI, -CP_BIG);
    la_table_init(ring + 2 * N, threadIdx.x);
    __syncwarp();
    const unsigned K = (unsigned) __cvta_generic_to_shared(ring + 2 * N);
    const unsigned KADJ = K - 0xB0000000u;
#define LA(a, b) la3((a), (b), KADJ)
#ifdef CP_ALLTABLE
#define LAP(a, b) la3((a), (b), KADJ)
#else
    const LaCoef KP = la_coef();
#define LAP(a, b) logadd2p((a), (b), KP)
#endif
    const float gMC = P.tMC, gMX = P.tMX, gOX = P.tOX, gOY = P.tOY, gEX = P.tEX, tSX = P.tSX;
    const float tMY = MACH ? P.vYM : P.tMY, tEY = MACH ? P.vYY : P.tEY;
#ifndef CP_NJ
#define CP_NJ 1
#endif
    // chunks per task of the posterior sweeps.  2 gives every lane two independent dependency chains (165 registers, 12
    // resident warps): measured equal at e = 128 / 256 and 11 % slower at e = 64, where a diagonal has 3 - 4 chunks and the odd
    // one wastes half a task (profiles/r2_tma_rows_ab.txt) -- so 1 it is; the E-step's body is too wide for two anyway
    constexpr int NJ = EXPECT ? 1 : CP_NJ;
    const int ZONE = P.tbDiags + 3;           // forward diagonals tbf-1 .. Dt keep all three states
    const bool allSpec = EXPECT || A.all_spec != 0;
    const bool unbanded = P.mode == 2;
    const float logThrLo = __logf(P.threshold) - 1e-3f;     // pre-filter; the exact test is p >= threshold
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
        const int planFlags = A.flags[itemIdx];
        const float4 *xpA = A.xparams + (MACH ? 4 : 3) * it.xp_off, *xpB = xpA + (lX + 2), *xpC = xpB + (lX + 2), *xpD = xpC + (lX + 2);
        const float4 *evp = A.events + it.ev_off;
        const unsigned *bitsp = A.bits + it.pad0;
        const int *tbp = A.tbs + it.pad1;
        int *pairs = A.pairs + 3 * it.pair_off;
        double *dbgTot = (A.totals != nullptr && it.tot_off >= 0) ? A.totals + it.tot_off : nullptr;
        int nPairs = 0, status = 0, nTb = 0;
        double lastTotal = 0.0;
        double eT[9] = { 0, 0, 0, 0, 0, 0, 0, 0, 0 };            // EXPECT: this lane's share of the transition sums
        double eLik = 0.0;
        double eBin0 = 0.0, eBin1 = 0.0;                             // EXPECT, vanilla: this lane's two skip bins (lane, lane + 32)

        if (D == 0 || (planFlags & (EXPECT ? 12 : 4)) != 0) {
            // nothing to align, or the plan refused the item (band edge jumps; E-step: a non-ACGT reference k-mer makes
            // the total -inf and the reference drops the read)
            if (lane == 0) { ItemOut &o = A.out[itemIdx]; o.n_pairs = 0; o.status = D == 0 ? 0 : (planFlags & 12) | 2; o.total_logprob = 0.0; o.n_tracebacks = 0; }
            continue;
        }
        auto bandBits = [&](int d) -> unsigned { return (bitsp[d >> 4] >> ((d & 15) << 1)) & 3u; };
        auto rowOf = [&](int d) -> int { return d % R; };

        auto resetRing = [&]() {
            __syncwarp();
            for (int i = lane; i < 2 * N; i += 32) ring[i] = NIENT;
            __syncwarp();
        };
        if (EXPECT) { for (int i = lane; i < N; i += 32) colAcc[i] = 0.f; }

        // ---- diagonal 0: the single cell (0,0) holds the start vector (impl/pairwiseAligner.c:897-898) ----------
        resetRing();
        float4 *A1 = ring, *A2 = ring + N;       // A1: newest diagonal, A2: the one before (forward); mirrored backward
        if (lane == 0) {
            const float *sv = (it.flags & 1) ? P.rstartv : P.startv;
            A1[0] = make_float4(sv[0], sv[1], sv[2], 0.f);
            if (EXPECT) { rows[0] = make_float4(sv[0], sv[1], sv[2], 0.f); plane2[0] = make_float2(NI, NI); }
            else { rows[0] = make_float4(sv[0], 0.f, NI, NI); plane2[allSpec ? 0 : (long long) (ZONE + 1) * N] = make_float2(sv[1], sv[2]); }   // zone 1, slot 1
        }
        __syncwarp();

        int dcur = 0, tracedBackTo = 0;
        int lo = 0, hi = 0;

        while (tracedBackTo < D) {
            // the traceback point of this segment comes from the plan (impl/pairwiseAligner.c:903-918)
            const int Dt = tbp[nTb];
            const bool atEnd = Dt == D;
            const int tbf = Dt - (atEnd ? 0 : P.tbDiags + 1);           // tracedBackFrom
            const int zonePar = nTb & 1;
            const int segStart = dcur, prevZoneBase = tracedBackTo - 1;
            // forward states X, Y of diagonal dd (POSTERIOR: kept only where a total or a restart needs them)
            auto specRowOf = [&](int dd) -> float2 * {
                if (allSpec) return plane2 + (long long) rowOf(dd) * N;
                int slot;
                if (dd <= segStart) slot = (zonePar ^ 1) * ZONE + (dd - prevZoneBase);
                else { const int kk = tbf - dd; slot = kk <= 1 ? zonePar * ZONE + (1 - kk) : 2 * ZONE + 2 * (kk / 10 - 1) + kk % 10; }
                return plane2 + (long long) slot * N;
            };

            // =============================== forward sweep ===============================================
            if (dcur < Dt) {
                int rowF = rowOf(dcur);
                int d = dcur + 1;
                unsigned fw = bitsp[d >> 4];
                { const unsigned b = (fw >> ((d & 15) << 1)) & 3u; lo += b & 1; hi += b >> 1; }
                int wlo = max(lo - 1, 0), c = (min(hi + 1, lX) - wlo) >> 5;
                rowF = rowF + 1 == R ? 0 : rowF + 1;
                float4 *frow = rows + (long long) rowF * N;
                int k = tbf - d;                                   // special-row bookkeeping (see specRowOf)
                int kq = k > 0 ? k / 10 : 0, k10 = k > 0 ? k % 10 : 0;
                bool spec = allSpec || k <= 1 || k10 <= 1;
                float2 *srow = allSpec ? plane2 + (long long) rowF * N
                                      : plane2 + (long long) (k <= 1 ? zonePar * ZONE + (1 - k) : 2 * ZONE + 2 * (kq - 1) + k10) * N;
                // Software pipeline over the tasks: a task is NJ (= 1, see above) neighbouring chunks of one diagonal; the
                // column records and the events of task i+1 are requested at
                // the top of task i (before the warp barrier, which ptxas does not move loads across) and reduced at its END
                // to the emissions (vanilla: + transitions) task i+1 needs.  c is the TOP chunk of the task: it covers chunks
                // c, c-1, .. c-NJ+1 as far as they exist.
                struct { float4 a, b, c, d, ev; } r[NJ];
                bool inb[NJ], ninb[NJ];
                auto load = [&](int j, int dd, int x, bool in) {
                    const int xx = in ? x : lX + 1;                   // outside the band: the all -inf dummy record
                    r[j].a = xpA[xx]; r[j].b = xpB[xx]; r[j].c = xpC[xx];
                    if (MACH) r[j].d = xpD[xx];
                    r[j].ev = evp[in ? dd - x : 0];
                };
                struct { float eM, eY, eX, tOX, tEX, tMC, tMX, tOY; } E[NJ];
                auto reduce = [&](int j) {
                    E[j].eM = emit(r[j].a, r[j].b, r[j].c, r[j].ev, false); E[j].eY = emit(r[j].a, r[j].b, r[j].c, r[j].ev, true);
                    const float cz = __int_as_float(__float_as_int(r[j].c.z) | (__float_as_int(r[j].c.w) & A.zero));
                    E[j].eX = MACH ? 0.f : cz;           // vanilla: the dummy record's transitions are -inf
                    // impl/stateMachine.c:1368-1409: the vanilla transitions are those of THIS column
                    E[j].tOX = MACH ? r[j].d.x : gOX; E[j].tEX = MACH ? r[j].d.y : gEX; E[j].tMC = MACH ? r[j].d.z : gMC;
                    E[j].tMX = MACH ? r[j].d.w : gMX; E[j].tOY = MACH ? cz : gOY;
                };
#pragma unroll
                for (int j = 0; j < NJ; j++) {
                    const int x0 = wlo + ((c - j) << 5) + lane;
                    ninb[j] = c - j >= 0 && x0 >= lo && x0 <= hi;
                    load(j, d, x0, ninb[j]);
                    inb[j] = ninb[j];
                    reduce(j);
                }
                for (;;) {
                    const bool last = c < NJ;                        // the task reaches chunk 0
                    int nd = d, nc = c - NJ, nlo = lo, nhi = hi, nwlo = wlo;
                    const bool stop = last && d >= Dt;
                    if (last && !stop) {
                        nd = d + 1;
                        if ((nd & 15) == 0) fw = bitsp[nd >> 4];
                        const unsigned b = (fw >> ((nd & 15) << 1)) & 3u;
                        nlo = lo + (b & 1); nhi = hi + (b >> 1);
                        nwlo = max(nlo - 1, 0);
                        nc = (min(nhi + 1, lX) - nwlo) >> 5;
                    }
                    {
                        float4 own[NJ], L[NJ], Mi[NJ];
                        int sj[NJ];
#pragma unroll
                        for (int j = 0; j < NJ; j++) {
                            const int nx = nwlo + ((nc - j) << 5) + lane;
                            ninb[j] = nc - j >= 0 && nx >= nlo && nx <= nhi && !stop;
                            load(j, nd, nx, ninb[j]);                // next task (the dummy record at the stop)
                        }
#pragma unroll
                        for (int j = 0; j < NJ; j++) {
                            const int x = wlo + ((c - j) << 5) + lane;
                            const int s = x & NM, sl = (x - 1) & NM;
                            sj[j] = s;
                            own[j] = A1[s]; L[j] = A1[sl]; Mi[j] = A2[sl];
                        }
                        __syncwarp();
                        float4 e4[NJ];
#pragma unroll
                        for (int j = 0; j < NJ; j++) {
                            const float U = fmaxf(own[j].w, fmaxf(L[j].w, Mi[j].w));
                            // impl/stateMachine.c:1314-1333: transitions folded in code order, emission added once
                            float tX = LA(L[j].x + E[j].tOX, L[j].y + E[j].tEX);
                            if (HAS_SX) tX = LA(tX, L[j].z + tSX);
                            float tM = LAP(LAP(Mi[j].x + E[j].tMC, Mi[j].y + E[j].tMX), Mi[j].z + tMY);
                            float tY = LA(own[j].x + E[j].tOY, own[j].z + tEY);
                            float cM = tM + (E[j].eM + (Mi[j].w - U)), cX = tX + (E[j].eX + (L[j].w - U)), cY = tY + (E[j].eY + (own[j].w - U));
                            float co = U;
                            rebase(cM, cX, cY, co);
                            if (!inb[j]) co = -CP_BIG;               // the emissions of the dummy record made the cell -inf
                            e4[j] = make_float4(cM, cX, cY, co);
                        }
#pragma unroll
                        for (int j = 0; j < NJ; j++) {
                            if (c - j >= 0) {                        // (a chunk that does not exist stores nothing)
                                const int s = sj[j];
                                A2[s] = e4[j];                       // descending x: in place over the d-2 entry
                                if (inb[j]) {
                                    if (EXPECT) { ROWST4(frow + s, e4[j]); ROWST2(srow + s, make_float2(E[j].eM, E[j].eY)); }
                                    else { ROWST4(frow + s, make_float4(e4[j].x, e4[j].w, E[j].eM, E[j].eY)); if (spec) ROWST2(srow + s, make_float2(e4[j].y, e4[j].z)); }
                                }
                            }
                        }
                        __syncwarp();
#pragma unroll
                        for (int j = 0; j < NJ; j++) { inb[j] = ninb[j]; reduce(j); }
                    }
                    if (last) {
                        { float4 *t = A1; A1 = A2; A2 = t; }
                        if (stop) break;
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
                        if (allSpec) srow = plane2 + (long long) rowF * N;
                        else {
                            k--;
                            if (k10 == 0) { k10 = 9; kq--; } else k10--;
                            spec = k <= 1 || k10 <= 1;
                            if (spec) srow = plane2 + (long long) (k <= 1 ? zonePar * ZONE + (1 - k) : 2 * ZONE + 2 * (kq - 1) + k10) * N;
                        }
                    }
                    c = nc;
                }
                dcur = Dt;
            }

            // =============================== traceback ===================================================
            // impl/pairwiseAligner.c:920-992.  The ring holds G = B + emission of diagonals d+1 (A1) and d+2 (A2).
            nTb++;
            int blo = lo, bhi = hi;
            resetRing();
            const float *endv = (atEnd && (it.flags & 2)) ? P.rendv : P.endv;
            const float endM = endv[0], endX = endv[1], endY = endv[2];
            float totSt = NI, totBase = 0.f;
            int tillTotal = 0;
            {
                int rowB = rowOf(Dt);
                // band bits through a one-word cache: the per-diagonal step reads bits(d + 1), a load every 16 diagonals
                int bwIdx = -1;
                unsigned bwWord = 0;
                auto bandBitsC = [&](int dd) -> unsigned {
                    if ((dd >> 4) != bwIdx) { bwIdx = dd >> 4; bwWord = bitsp[bwIdx]; }
                    return (bwWord >> ((dd & 15) << 1)) & 3u;
                };
#ifdef CP_TMA_ROWS
                // one lane asks the TMA unit for the band cells of a forward row: positions (x & NM), x = l .. h, of the row in
                // HBM into the same positions of stage buffer `sb` (two pieces when the ring wraps)
                auto tmaFetchRow = [&](int row, int l, int h, int sb) {
                    if (lane == 0 && !EXPECT) {
                        const float4 *src = rows + (long long) row * N;
                        const int p0 = l & NM, p1 = h & NM;
                        const unsigned bar = barAddr + 8u * sb, dst = stageAddr + (unsigned) sb * N * 16u;
                        if (p0 <= p1) {
                            const unsigned bytes = (unsigned) (p1 - p0 + 1) * 16u;
                            mbar_expect_tx(bar, bytes);
                            bulk_g2s(dst + p0 * 16u, src + p0, bytes, bar);
                        } else {
                            const unsigned b0 = (unsigned) (N - p0) * 16u, b1 = (unsigned) (p1 + 1) * 16u;
                            mbar_expect_tx(bar, b0 + b1);
                            bulk_g2s(dst + p0 * 16u, src + p0, b0, bar);
                            bulk_g2s(dst, src, b1, bar);
                        }
                    }
                };
                if (!EXPECT) tmaFetchRow(rowB, blo, bhi, 0);
#endif
                for (int d = Dt; d > tracedBackTo; d--) {
                    if (d < Dt) {
                        const unsigned b = bandBitsC(d + 1);
                        blo -= b & 1; bhi -= b >> 1;
                        rowB = rowB == 0 ? R - 1 : rowB - 1;
                    }
                    const float4 *frow = rows + (long long) rowB * N;
                    const float2 *erow = plane2 + (long long) rowB * N;          // EXPECT
                    const bool post = d <= tbf;
#ifdef CP_TMA_ROWS
                    const int sb = (Dt - d) & 1;
                    if (!EXPECT) {
                        // the row of the next diagonal goes into the other stage buffer (every lane is done with it: the warp
                        // barriers at the end of the previous diagonal), then wait for this diagonal's row
                        if (d - 1 > tracedBackTo) {
                            const unsigned bn = bandBitsC(d);
                            tmaFetchRow(rowB == 0 ? R - 1 : rowB - 1, blo - (int) (bn & 1), bhi - (int) (bn >> 1), sb ^ 1);
                        }
                        mbar_wait(barAddr + 8u * sb, sb ? tmaPhase1 : tmaPhase0);
                        if (sb) tmaPhase1 ^= 1; else tmaPhase0 ^= 1;
                    }
                    const float4 *srow_stage = stage + (size_t) sb * N;
#endif
                    const int wlo = max(blo - 1, 0), nch = ((min(bhi + 1, lX) - wlo) >> 5) + 1;
                    const int plo = max(blo, 1), phi = min(bhi, d - 1);   // cells with x > 0 and y > 0 report posteriors
                    bool doTotal = false;
                    if (post) { doTotal = unbanded ? (d == Dt) : (tillTotal == 0); tillTotal = tillTotal == 0 ? P.totalEvery - 1 : tillTotal - 1; }   // every totalEvery-th posterior diagonal, from the first
                    // What one cell of the backward sweep needs besides the ring: the forward sweep's record (match value,
                    // offset, the two emissions), the gap emission of the column (three-state) or the transitions of the
                    // successors' columns (vanilla), the k-mer index / skip bin for the E-step.  Requested a chunk ahead.
                    struct QRec { float4 F; float4 dR; float2 dO; float cz, ex; int kw; };
                    struct GRec { float eM, eY, eX, Fx, Fw, tOX, tEX, tMC, tMX, myLog, oOX, oEX; int kw; };
                    QRec qq[NJ];
                    GRec GG[NJ];
                    auto loadBj = [&](int j, int cc) {                 // chunk cc (one that does not exist reads as outside the band)
                        QRec &q = qq[j];
                        const int x = wlo + (cc << 5) + lane;
                        const bool in = cc < nch && x >= blo && x <= bhi;
                        const int xx = in ? x : lX + 1;
                        const float4 *cp = xpC + xx;
                        q.cz = reinterpret_cast<const float *>(cp)[2];
                        q.ex = MACH ? (in ? 0.f : NI) : q.cz;
                        q.kw = EXPECT ? reinterpret_cast<const int *>(cp)[3] : 0;
                        if (MACH) q.dR = xpD[min(x + 1, lX + 1)];
                        // vanilla E-step: the transitions into THIS column's gap-X state (log a_mx, log a_xx), a chunk ahead too
                        q.dO = make_float2(0.f, 0.f);
                        if (EXPECT && MACH) q.dO = *reinterpret_cast<const float2 *>(xpD + min(x, lX + 1));
                        q.F = make_float4(NI, NI, NI, NI);
                        if (EXPECT) { float2 e2 = make_float2(NI, NI); if (in) e2 = ROWLD2(erow + (x & NM)); q.F.z = e2.x; q.F.w = e2.y; }
#ifdef CP_TMA_ROWS
                        else if (in) q.F = srow_stage[x & NM];
#else
                        else if (in) q.F = ROWLD4(frow + (x & NM));
#endif
                    };
                    auto reduceBj = [&](int j) {
                        const QRec &q = qq[j];
                        GRec &G = GG[j];
                        // POSTERIOR record: (F_match, offset, eM, eY); EXPECT: (-, -, eM, eY)
                        G.eM = q.F.z; G.eY = q.F.w; G.Fx = q.F.x; G.Fw = EXPECT ? 0.f : q.F.y;
                        G.eX = q.ex; G.myLog = q.cz; G.kw = q.kw;
                        G.tOX = MACH ? q.dR.x : 0.f; G.tEX = MACH ? q.dR.y : 0.f; G.tMC = MACH ? q.dR.z : 0.f; G.tMX = MACH ? q.dR.w : 0.f;
                        G.oOX = q.dO.x; G.oEX = q.dO.y;
                    };
                    auto loadB = [&](int cc) { loadBj(0, cc); };
                    auto reduceB = [&]() { reduceBj(0); };
                    GRec &G = GG[0];
                    if (!doTotal) {
#pragma unroll
                        for (int j = 0; j < NJ; j++) loadBj(j, j);
                    }
                    if (d - 2 > tracedBackTo) {
                        // the forward records of diagonal d-2 were written >= 1000 diagonals ago: DRAM -> L2 now
                        const int rowP = rowB >= 2 ? rowB - 2 : rowB - 2 + R;
                        const float4 *fp = rows + (long long) rowP * N;
                        const int p0 = max(wlo - 32, 0), pw = (nch + 1) << 5;
#pragma unroll 1
                        for (int o = lane << 3; o < pw; o += 256) prefetch_l2(fp + ((p0 + o) & NM));
                        if (EXPECT) { const float2 *ep = plane2 + (long long) rowP * N; for (int o = lane << 4; o < pw; o += 512) prefetch_l2(ep + ((p0 + o) & NM)); }
                    }
                    if ((d & 15) == 0) {
                        const int px = blo - 32 - lane;
                        if (px >= 0) { prefetch_l2(xpC + px); if (MACH) prefetch_l2(xpD + px); }
                    }

                    // E-step inputs: forward cells of the three predecessors live on diagonals d-1 and d-2
                    int el1 = 0, eh1 = -1, el2 = 0, eh2 = -1;
                    const float4 *frow1 = frow, *frow2 = frow;
                    float aT[9] = { 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f };
                    if (EXPECT && post) {
                        const unsigned b0 = bandBits(d);
                        el1 = blo - (int) (b0 & 1); eh1 = bhi - (int) (b0 >> 1);
                        // the reference has freed the forward diagonals below tracedBackTo by now (impl/pairwiseAligner.c:
                        // 971-985), so on the first diagonal of a later traceback the match predecessors are absent
                        if (d >= 2 && d - 2 >= tracedBackTo) { const unsigned b1 = bandBits(d - 1); el2 = el1 - (int) (b1 & 1); eh2 = eh1 - (int) (b1 >> 1); }
                        const int r1 = rowB == 0 ? R - 1 : rowB - 1, r2 = r1 == 0 ? R - 1 : r1 - 1;
                        frow1 = rows + (long long) r1 * N; frow2 = rows + (long long) r2 * N;
                    }

                    // B of one cell from the ring (pull form of impl/pairwiseAligner.c:378-383: first from diagonal
                    // d+2 as "middle", then d+1 in ascending x-y as "upper", then "lower"), in units U.  Outside the band
                    // the value is garbage that the -inf emissions of the dummy record turn into G = -inf.
                    // (vanilla: the transitions INTO a successor cell are those of the successor's column: pdR = record d
                    // of column x+1; M->Y uses this column's log a_my)
                    auto cellB = [&](int x, int s, const float4 pdR, float myLog, float &bM, float &bX, float &bY, float &U) {
                        if (d == Dt) { bM = endM; bX = endX; bY = endY; U = 0.f; return; }
                        const float tOX = MACH ? pdR.x : gOX, tEX = MACH ? pdR.y : gEX, tMC = MACH ? pdR.z : gMC,
                                    tMX = MACH ? pdR.w : gMX, tOY = MACH ? myLog : gOY;
                        const int sr = (x + 1) & NM;
                        const float4 own = lds4(A1 + s), R1 = lds4(A1 + sr), R2 = lds4(A2 + sr);
                        U = fmaxf(own.w, fmaxf(R1.w, R2.w));
                        const float gm2 = R2.x + (R2.w - U), gx1 = R1.y + (R1.w - U), gy1 = own.z + (own.w - U);
                        bM = LAP(LAP(gm2 + tMC, gy1 + tOY), gx1 + tOX);
                        bX = LA(gm2 + tMX, gx1 + tEX);
                        bY = LA(gm2 + tMY, gy1 + tEY);
                        if (HAS_SX) bY = LA(bY, gx1 + tSX);
                    };
                    // posterior of one cell + G = B + emission, re-based, back into the ring
                    // E-step: the forward cells of the three predecessors of cell x (lower, middle, upper), requested at the top
                    // of a chunk so that the ~100 instructions of cellB hide their latency (they were the kernel's hottest
                    // stall, 14 % of all samples, when loaded at their first use)
                    struct Pred { float4 L, M, U; };
                    auto loadPred = [&](int x, bool inb) -> Pred {
                        Pred P3 = { NIENT, NIENT, NIENT };
                        if (EXPECT && post && inb) {
                            if (x - 1 >= el1 && x - 1 <= eh1) P3.L = frow1[(x - 1) & NM];
                            if (x - 1 >= el2 && x - 1 <= eh2) P3.M = frow2[(x - 1) & NM];
                            if (x >= el1 && x <= eh1) P3.U = frow1[x & NM];
                        }
                        return P3;
                    };
                    auto cellPost = [&](int x, int s, bool inb, float bM, float bX, float bY, float U, float eM, float eY,
                                        float eX, float Fx, float Fw, float myLog, int kw, const Pred &P3, float oOX, float oEX) {
                        if (EXPECT) {
                            if (post) {
                                // diagonalCalculation_Expectations (impl/pairwiseAligner.c:841-863): for every transition
                                // into this cell p = exp(F_pred[from] + B[to] + eP + tP - total) (:426-443)
                                const float4 FL = P3.L, FM = P3.M, FU = P3.U;
                                const float tOX = MACH ? oOX : gOX, tEX = MACH ? oEX : gEX, tMC = gMC, tMX = gMX, tOY = MACH ? myLog : gOY;
                                const float kX = (bX + eX) + (((FL.w + U) - totBase) - totSt);
                                float pMX = __expf(FL.x + tOX + kX), pXX = __expf(FL.y + tEX + kX);
                                if (MACH) {
                                    // cell_signal_updateBetaAndAlphaProb (impl/pairwiseAligner.c:478-498): skip bins only;
                                    // 60 bins: warp-aggregated per bin would need a sort, the sums go out per lane
                                    // 60 bins in shared memory (global FP64 atomics from every lane of every warp met on 60
                                    // addresses); FP32 over a few diagonals, then into this item's FP64 registers
                                    const int bin = kw;
                                    if (bin >= 0) {
                                        if (pMX > 1e-13f) atomicAdd(colAcc + bin, pMX);
                                        if (pXX > 1e-13f) atomicAdd(colAcc + 30 + bin, pXX);
                                    }
                                } else {
                                    const float kM = (bM + eM) + (((FM.w + U) - totBase) - totSt);
                                    const float kY = (bY + eY) + (((FU.w + U) - totBase) - totSt);
                                    const float pYX = HAS_SX ? __expf(FL.z + tSX + kX) : 0.f;
                                    aT[1] += pMX; aT[4] += pXX; aT[7] += pYX;             // from * 3 + to, to = X (1)
                                    aT[0] += __expf(FM.x + tMC + kM); aT[3] += __expf(FM.y + tMX + kM); aT[6] += __expf(FM.z + tMY + kM);
                                    aT[2] += __expf(FU.x + tOY + kY); aT[8] += __expf(FU.z + tEY + kY);
                                    // k-mer skip sums: per column in shared memory, out when the column leaves the band
                                    if (inb) colAcc[s] += pMX + pXX + pYX;
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
                        float gM = bM + eM, gX = bX + eX, gY = bY + eY, go = U;   // the emissions are -inf outside the band
                        rebase(gM, gX, gY, go);
                        if (!inb) go = -CP_BIG;
                        A2[s] = make_float4(gM, gX, gY, go);
                    };

                    if (!doTotal) {
#pragma unroll
                        for (int j = 0; j < NJ; j++) reduceBj(j);
                        for (int c = 0; c < nch; c += NJ) {            // ascending x: in-place update of the d+2 entries; NJ chunks per task
                            GRec cur[NJ];
                            Pred P3[NJ];
                            float bM[NJ], bX[NJ], bY[NJ], U[NJ];
#pragma unroll
                            for (int j = 0; j < NJ; j++) cur[j] = GG[j];
#pragma unroll
                            for (int j = 0; j < NJ; j++) loadBj(j, c + NJ + j);      // the next task's chunks
#pragma unroll
                            for (int j = 0; j < NJ; j++) {
                                const int x = wlo + ((c + j) << 5) + lane;
                                P3[j] = loadPred(x, c + j < nch && x >= blo && x <= bhi);
                                cellB(x, x & NM, make_float4(cur[j].tOX, cur[j].tEX, cur[j].tMC, cur[j].tMX), cur[j].myLog, bM[j], bX[j], bY[j], U[j]);
                            }
                            __syncwarp();
#pragma unroll
                            for (int j = 0; j < NJ; j++) {
                                if (c + j < nch) {                     // (a chunk that does not exist stores nothing and reports nothing)
                                    const int x = wlo + ((c + j) << 5) + lane;
                                    cellPost(x, x & NM, x >= blo && x <= bhi, bM[j], bX[j], bY[j], U[j], cur[j].eM, cur[j].eY, cur[j].eX,
                                             cur[j].Fx, cur[j].Fw, cur[j].myLog, cur[j].kw, P3[j], cur[j].oOX, cur[j].oEX);
                                }
                            }
                            __syncwarp();
#pragma unroll
                            for (int j = 0; j < NJ; j++) reduceBj(j);
                        }
                    } else {
                        // ---- totalProbability (impl/pairwiseAligner.c:736-754), recomputed every 10th posterior diagonal
                        // pass 1: B into the ring (units in .w), dot-product terms and their units aside
                        OrderedFold fold1;
                        fold1.reset();
                        const float2 *srow = specRowOf(d);
                        // the forward cell (M, X, Y, offset) of a chunk is requested a chunk ahead of its use
                        auto loadF = [&](int cc) -> float4 {
                            const int x = wlo + (cc << 5) + lane, s = x & NM;
                            float4 F = NIENT;
                            if (cc < nch && x >= blo && x <= bhi) {
                                if (EXPECT) F = frow[s];
                                else { const float4 f = frow[s]; const float2 xy = srow[s]; F = make_float4(f.x, xy.x, xy.y, f.y); }
                            }
                            return F;
                        };
                        float4 Fn = loadF(0);
                        for (int c = 0; c < nch; c++) {
                            const int x = wlo + (c << 5) + lane, s = x & NM;
                            const bool inb = x >= blo && x <= bhi;
                            const float4 F = Fn;
                            Fn = loadF(c + 1);
                            float bM, bX, bY, U;
                            float4 pdR = NIENT;
                            float myLog = 0.f;
                            if (MACH) { pdR = xpD[min(x + 1, lX + 1)]; myLog = xpC[min(x, lX + 1)].z; }
                            cellB(x, s, pdR, myLog, bM, bX, bY, U);
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
                            const unsigned bu = bandBits(d + 1), bd = bandBits(d);
                            const int l1 = blo + (int) (bu & 1), h1 = bhi + (int) (bu >> 1);
                            const int lm1 = blo - (int) (bd & 1), hm1 = bhi - (int) (bd >> 1);
                            const int rowM = rowB == 0 ? R - 1 : rowB - 1;
                            const float4 *fprev = rows + (long long) rowM * N;
                            const float2 *sprev = specRowOf(d - 1);
                            OrderedFold fold2;
                            fold2.reset();
                            auto loadFp = [&](int xb) -> float4 {        // forward cell of x - 1 on diagonal d - 1, a round ahead
                                const int x = xb + lane;
                                float4 F = NIENT;
                                if (x <= h1 && x - 1 >= lm1 && x - 1 <= hm1) {
                                    if (EXPECT) F = fprev[(x - 1) & NM];
                                    else { const float4 f = fprev[(x - 1) & NM]; const float2 xy = sprev[(x - 1) & NM]; F = make_float4(f.x, xy.x, xy.y, f.y); }
                                }
                                return F;
                            };
                            float4 Fpn = loadFp(l1);
                            for (int xb = l1; xb <= h1; xb += 32) {
                                const int x = xb + lane;
                                float val = NI;
                                const float4 F = Fpn;
                                Fpn = loadFp(xb + 32);
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
                            tot = logadd2(t1, t2v, K);
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
                            const Pred P3 = loadPred(x, inb);
                            __syncwarp();
                            cellPost(x, s, inb, b.x, b.y, b.z, b.w, cur.eM, cur.eY, cur.eX, cur.Fx, cur.Fw, cur.myLog, cur.kw, P3, cur.oOX, cur.oEX);
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
                    if (EXPECT && !MACH) {
                        // the column that leaves the band with this diagonal (the band of d-1 ends one column earlier), or,
                        // on the last diagonal of the traceback, every column still inside: one atomic per column
                        __syncwarp();
                        const bool lastDiag = d == tracedBackTo + 1;
                        const unsigned b0 = lastDiag ? 0u : bandBits(d);
                        const int xFrom = lastDiag ? blo : (b0 >> 1 ? bhi : bhi + 1);
                        for (int x = xFrom + lane; x <= bhi; x += 32) {
                            const float v = colAcc[x & NM];
                            colAcc[x & NM] = 0.f;
                            const int kmer = __float_as_int(xpC[x].w);
                            if (v > 1e-13f && kmer >= 0) atomicAdd(A.expect + 9 + kmer, (double) v);
                        }
                        __syncwarp();
                    }
                    if (EXPECT && MACH && post && ((d & 7) == 0 || d == tracedBackTo + 1)) {
                        __syncwarp();
                        eBin0 += (double) colAcc[lane]; colAcc[lane] = 0.f;
                        if (lane < 28) { eBin1 += (double) colAcc[lane + 32]; colAcc[lane + 32] = 0.f; }
                        __syncwarp();
                    }
                    { float4 *t = A1; A1 = A2; A2 = t; }
                }
            }
            tracedBackTo = tbf;

            // =============================== restore the forward state at Dt ==============================
            if (tracedBackTo < D) {
                resetRing();
                A1 = ring; A2 = ring + N;
                const unsigned b = bandBits(Dt);
                const int l1 = lo - (int) (b & 1), h1 = hi - (int) (b >> 1);
                const float4 *f0 = rows + (long long) rowOf(Dt) * N;
                const float4 *f1 = rows + (long long) rowOf(Dt - 1) * N;
                if (EXPECT) {
                    for (int x = lo + lane; x <= hi; x += 32) A1[x & NM] = f0[x & NM];
                    for (int x = l1 + lane; x <= h1; x += 32) A2[x & NM] = f1[x & NM];
                } else {
                    const float2 *s0 = specRowOf(Dt), *s1 = specRowOf(Dt - 1);
                    for (int x = lo + lane; x <= hi; x += 32) { const float4 f = f0[x & NM]; const float2 xy = s0[x & NM]; A1[x & NM] = make_float4(f.x, xy.x, xy.y, f.y); }
                    for (int x = l1 + lane; x <= h1; x += 32) { const float4 f = f1[x & NM]; const float2 xy = s1[x & NM]; A2[x & NM] = make_float4(f.x, xy.x, xy.y, f.y); }
                }
                __syncwarp();
            }
        }

        if (EXPECT) {
            // an item whose totals went non-finite contributes nothing to the transition sums and the likelihood (the
            // reference drops such a read); its k-mer / skip-bin sums were NaN-free by construction (p > 1e-13 tests)
            const bool bad = (status & 2) != 0;
#pragma unroll
            for (int i = 0; i < (MACH ? 0 : 9); i++) {
                double v = eT[i];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(CP_FULL, v, o);
                if (lane == 0 && !bad && v == v) atomicAdd(A.expect + i, v);
            }
            if (lane == 0 && !bad && eLik == eLik) atomicAdd(A.expect + (MACH ? 60 : 9 + 4096), eLik);
            if (MACH && !bad) {
                if (eBin0 > 0.0) atomicAdd(A.expect + lane, eBin0);
                if (lane < 28 && eBin1 > 0.0) atomicAdd(A.expect + 32 + lane, eBin1);
            }
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
#undef LAP
}  // namespace cpecan
