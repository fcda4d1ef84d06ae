// cpecan_logadd.cuh -- pieces shared by the alignment kernels (cpecan_align3.cuh): the two bit-identical FP32 forms of
// the reference's logAdd (impl/pairwiseAligner.c:235-255: 4-segment cubic, 7.5 cut-off) as max + q(|x - y|), q = cubic -
// identity -- predicated immediate-operand FFMAs for the dependent chain, a 17-entry shared-memory coefficient table
// off the chain --, the exact ordered fold of dpDiagonal_dotProduct (:587-597) that touches only the elements that can
// matter, and the integer re-basing of a cell.  (The round-1 kernel k_align2 that lived here was replaced by k_align3;
// its profiles stay under profiles/r1_*.)
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

}  // namespace cpecan
