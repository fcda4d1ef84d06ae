// cpecan_cuda.cu -- host side of the C-ABI declared in include/cpecan_cuda.h: context, staging, kernel launches.
// All device work runs on the context's own streams; timings come from CUDA events on those streams.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <numeric>
#include <string>
#include <vector>

#include "cpecan_cuda.h"
#include "cpecan_kernels.cuh"
#include "cpecan_logadd.cuh"
#include "cpecan_align3.cuh"
#include "cpecan_generic.cuh"

using namespace cpecan;

namespace {

// one warp per alignment; ring of N = 128 << bucket positions (power of two) in shared memory
constexpr int NCFG2 = 6;
constexpr int CFG2_MARGIN = 40;                   // ring positions beyond the widest diagonal (window + the rest of its last chunk)
inline int cfg2N(int b) { return 128 << b; }

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + std::min<size_t>(bytes / 8, (size_t) 64 << 20) + 256;   // growth slack, bounded: the big buffers are tens of GB
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

struct Model {
    double *match = nullptr, *gapy = nullptr, *gapx = nullptr; int n_gapx = 0; bool live = false;
    // cpecan_cuda_upload_hdp
    bool hdp = false; double *hdpY = nullptr, *hdpSlope = nullptr; int *hdpKmer = nullptr; int hdpLen = 0;
    double x0 = 0, x1 = 0, xn = 0, step = 0;
};

struct Bucket {
    std::vector<int> order;      // item indices, largest first
    int nCta = 0, ringRows = 0, specRows = 0;
    long long stride = 0;        // float4 per CTA
    size_t scratchOff = 0;       // float4 into the scratch buffer
    size_t orderOff = 0;         // ints into the order buffer
};

}  // namespace

// ---- NCCL through dlopen: the library is needed only by multi-GPU training
namespace {
struct Nccl {
    void *h = nullptr;
    int (*getUniqueId)(void *) = nullptr;
    int (*commInitRank)(void **, int, char[128], int) = nullptr;      // ncclUniqueId is passed by value: 128 bytes
    int (*allReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*commDestroy)(void *) = nullptr;
    const char *(*getErrorString)(int) = nullptr;
};
struct NcclId { char b[CPECAN_NCCL_UNIQUE_ID_BYTES]; };
Nccl *ncclLib(std::string &err) {
    static Nccl L;
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    if (L.h) return &L;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { err = std::string("libnccl.so.2 not found: ") + dlerror(); return nullptr; }
    L.getUniqueId = (int (*)(void *)) dlsym(h, "ncclGetUniqueId");
    L.commInitRank = (int (*)(void **, int, char[128], int)) dlsym(h, "ncclCommInitRank");
    L.allReduce = (int (*)(const void *, void *, size_t, int, int, void *, cudaStream_t)) dlsym(h, "ncclAllReduce");
    L.commDestroy = (int (*)(void *)) dlsym(h, "ncclCommDestroy");
    L.getErrorString = (const char *(*)(int)) dlsym(h, "ncclGetErrorString");
    if (!L.getUniqueId || !L.commInitRank || !L.allReduce) { err = "libnccl.so.2 lacks ncclGetUniqueId / ncclCommInitRank / ncclAllReduce"; return nullptr; }
    L.h = h;
    return &L;
}
}  // namespace


struct cpecan_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[8] = {};
    cudaStream_t bstream[8] = {};     // one stream per ring-size bucket: their kernels overlap, so one bucket's tail
    cudaEvent_t bev[8] = {};          // of long alignments is filled by the next bucket's CTAs
    bool running = false;
    cudaDeviceProp prop{};
    std::string err;
    std::mutex mu;
    std::vector<Model> models;
    DevBuf dModels;
    bool modelsDirty = true;

    // staged batch
    int64_t n = 0;
    int mode = 0;
    int machine = 0;             // 0 three-state, 1 vanilla
    int generic = 0;             // 0, or the StateMachineType run by k_align_generic: 6 fourState, 5 echelon; and 2 threeState /
                                 // 4 vanilla when the band is one k_align3 does not take (odd diagonalExpansion)
    GenParams G{};
    bool exact = false;          // cpecan_cuda_set_exact_arithmetic: threeState / vanilla posteriors on the FP64 kernel
    double mToYNotX = 0.0;
    DevParams P{};
    bool hasSX = false;
    std::vector<Item> hItems;
    std::vector<ItemOut> hOut;
    DevBuf dItems, dOut, dRef, dRefOff, dEvSrc, dEvSrcOff, dAnchors, dScale, dCentre, dXp, dEv, dPairs, dOrder,
           dQueue, dScratch, dTotals, dCompact, dCompactOff, dExpect, dBits, dTbs, dFlags, dBands, dBandOff, dColp, dRowp;
    int64_t pairCapTotal = 0, totalsLen = 0;
    Bucket buckets[NCFG2];
    int stagedMaxLX = 0, stagedMaxLY = 0;
    long long stagedEvTot = 0;
    bool stagedScaled = false;
    int occ2[NCFG2][2][2][2] = {};  // [bucket][machine][hasSX][expect]
    int occCap = 0;              // resident warps per SM this context may take (0 = all that fit)
    cudaEvent_t evBlock = nullptr;   // cudaEventBlockingSync: host threads sleep while they wait (several contexts per process)
    bool wantTotals = false;
    std::vector<int64_t> hTotOff;
    cpecan_timing timing{};
    void *ncclComm = nullptr;    // ncclComm_t (cpecan_cuda_nccl_init)
    int ncclRanks = 0;
};

namespace {

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                         \
            return CPECAN_ERR_CUDA;                                                                \
        }                                                                                          \
    } while (0)

// Waits for a stream without spinning: a caller that streams sub-batches through several contexts has one host thread
// per context, and most of them are waiting at any time.
static cudaError_t waitStream(cpecan_ctx *ctx, cudaStream_t s) {
    cudaError_t e = cudaEventRecord(ctx->evBlock, s);
    return e != cudaSuccess ? e : cudaEventSynchronize(ctx->evBlock);
}

// k_align3<MACH, HAS_SX, EXPECT>: dispatch over the instantiations (the vanilla machine has no Y->X transition)
template <typename F> auto dispatchK2(int mach, bool sx, bool ex, F f) {
    if (mach) return ex ? f(k_align3<1, false, true>) : f(k_align3<1, false, false>);
    if (sx) return ex ? f(k_align3<0, true, true>) : f(k_align3<0, true, false>);
    return ex ? f(k_align3<0, false, true>) : f(k_align3<0, false, false>);
}
inline size_t smemCfg2(int cfg, int mach, bool ex) { (void) mach; return align3_smem_bytes(cfg2N(cfg), ex); }   // E-step: per-column sums (three-state) or the 60 skip bins (vanilla)
int occCfg2(int cfg, int mach, bool sx, bool ex) {
    const size_t bytes = smemCfg2(cfg, mach, ex);
    return dispatchK2(mach, sx, ex, [&](auto k) {
        int nb = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k, 32, bytes);
        return nb; });
}
void launchCfg2(int cfg, int mach, bool sx, bool expect, const KernelArgs3 &a, int nCta, cudaStream_t s) {
    const size_t bytes = smemCfg2(cfg, mach, expect);
    dispatchK2(mach, sx, expect, [&](auto k) { k<<<nCta, 32, bytes, s>>>(a); return 0; });
}

// fourState / echelon: the FP64 kernel; ring size per bucket as for k_align3, occupancy from its shared-memory need
template <typename F> auto dispatchGen(int sm, F f) {
    switch (sm) {
        case CPECAN_SM_ECHELON: return f(k_align_generic<5>, 7);
        case CPECAN_SM_THREE_STATE: return f(k_align_generic<2>, 3);
        case CPECAN_SM_THREE_STATE_HDP: return f(k_align_generic<7>, 3);
        case CPECAN_SM_VANILLA: return f(k_align_generic<4>, 3);
        default: return f(k_align_generic<6>, 4);
    }
}
int genStates(int sm) { return sm == CPECAN_SM_ECHELON ? 7 : (sm == CPECAN_SM_FOUR_STATE ? 4 : 3); }
// Dynamic shared memory: every alignment kernel gets the device's opt-in maximum ONCE per device and process.  (Setting the
// attribute per ring size, as a context was created or a batch staged, lowered it for a moment under the kernels other
// host threads were launching through their own contexts: "invalid argument".)
void setMaxSmemOnce(int device, int optin) {
    static std::once_flag flags[64];
    std::call_once(flags[device & 63], [&]() {
        for (int mach = 0; mach < 2; mach++)
            for (int sx = 0; sx < 2; sx++)
                for (int ex = 0; ex < 2; ex++)
                    dispatchK2(mach, sx != 0, ex != 0, [&](auto k) { return cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, optin); });
        const int sms[5] = { CPECAN_SM_THREE_STATE, CPECAN_SM_VANILLA, CPECAN_SM_ECHELON, CPECAN_SM_FOUR_STATE, CPECAN_SM_THREE_STATE_HDP };
        for (int sm : sms) dispatchGen(sm, [&](auto k, int) { return cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, optin); });
        cudaGetLastError();
    });
}

int occGen(int cfg, int sm, size_t optin) {
    return dispatchGen(sm, [&](auto k, int S) {
        const size_t bytes = generic_smem_bytes(cfg2N(cfg), S);
        if (bytes > optin) return 0;
        int nb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k, 32, bytes) != cudaSuccess) { cudaGetLastError(); return 0; }
        return nb; });
}

__global__ void k_compact(const Item *items, const ItemOut *out, const long long *dstOff, int n, const int *src, int *dst) {
    const int i = blockIdx.x;
    if (i >= n) return;
    const Item it = items[i];
    const int np = min(out[i].n_pairs, it.pair_cap);
    const int *s = src + 3 * it.pair_off;
    int *d = dst + 3 * dstOff[i];
    for (int j = threadIdx.x; j < 3 * np; j += blockDim.x) d[j] = s[j];
}

// transitions and state vectors of the machine (everything but the banding parameters)
int fillMachine(cpecan_ctx *ctx, const cpecan_hmm *hmm) {
    DevParams &P = ctx->P;
    const float NI = -INFINITY;
    P.startv[0] = 0.f; P.startv[1] = NI; P.startv[2] = NI;                 // impl/stateMachine.c:1168-1172
    P.rstartv[0] = NI; P.rstartv[1] = 0.f; P.rstartv[2] = 0.f;             // :1174-1177
    if (hmm->sm_type == CPECAN_SM_THREE_STATE) {
        const double *t = hmm->transitions;
        P.tMC = (float) t[0]; P.tMX = (float) t[1]; P.tMY = (float) t[2]; P.tOX = (float) t[3]; P.tOY = (float) t[4];
        P.tEX = (float) t[5]; P.tEY = (float) t[6]; P.tSX = (float) t[7]; P.tSY = (float) t[8];
        P.endv[0] = (float) t[0]; P.endv[1] = (float) t[1]; P.endv[2] = (float) t[2];           // :1179-1192
        P.rendv[0] = (float) ((t[3] + t[4]) / 2.0); P.rendv[1] = (float) t[5]; P.rendv[2] = (float) t[6];  // :1194-1207
        ctx->hasSX = !(std::isinf(t[7]) && t[7] < 0);
        ctx->machine = 0;
        P.vYM = P.vYY = 0.f;
    } else if (hmm->sm_type == CPECAN_SM_VANILLA) {
        // stateMachine3Vanilla: per-column transitions come from the skip bins (impl/stateMachine.c:1368-1409); here
        // only the two global ones and the end vectors (:1209-1235)
        const double *v = hmm->vanilla;
        const double a_yy = v[1];
        P.vYY = (float) std::log(a_yy);
        P.vYM = (float) std::log(1.0 - a_yy);
        P.tMC = P.tMX = P.tMY = P.tOX = P.tOY = P.tEX = P.tEY = 0.f; P.tSX = P.tSY = NI;
        P.endv[0] = (float) v[2]; P.endv[1] = (float) v[3]; P.endv[2] = (float) v[4];
        P.rendv[0] = (float) ((v[3] + v[4]) / 2.0); P.rendv[1] = (float) v[3]; P.rendv[2] = (float) v[4];
        ctx->hasSX = false;
        ctx->machine = 1;
        ctx->mToYNotX = v[0];
    } else if (hmm->sm_type == CPECAN_SM_THREE_STATE_HDP) {
        // the FP64 kernel with the three-state transitions; emissions from the item's HDP model
        ctx->generic = hmm->sm_type;
        ctx->G.sm = hmm->sm_type;
        for (int i = 0; i < 9; i++) ctx->G.t3[i] = hmm->transitions[i];
        P.tMC = P.tMX = P.tMY = P.tOX = P.tOY = P.tEX = P.tEY = 0.f; P.tSX = P.tSY = NI;
        ctx->hasSX = false;
        ctx->machine = 0;
        P.vYM = P.vYY = 0.f;
    } else if (hmm->sm_type == CPECAN_SM_FOUR_STATE || hmm->sm_type == CPECAN_SM_ECHELON) {
        // the FP64 kernel of cpecan_generic.cuh: everything it needs is the type and, fourState, the 11 transitions
        ctx->generic = hmm->sm_type;
        ctx->G.sm = hmm->sm_type;
        for (int i = 0; i < 11; i++) ctx->G.t4[i] = hmm->four_state[i];
        P.tMC = P.tMX = P.tMY = P.tOX = P.tOY = P.tEX = P.tEY = 0.f; P.tSX = P.tSY = NI;
        ctx->hasSX = false;
        ctx->machine = 0;
        P.vYM = P.vYY = 0.f;
    } else {
        ctx->err = "state machine type not implemented on device (threeState = 2, vanilla = 4, echelon = 5, fourState = 6, threeStateHdp = 7 are)";
        return CPECAN_ERR_ARG;
    }
    if (hmm->sm_type == CPECAN_SM_THREE_STATE || hmm->sm_type == CPECAN_SM_VANILLA) ctx->generic = 0;
    P.machine = ctx->machine;
    P.hasSX = ctx->hasSX;
    return CPECAN_OK;
}

// threeState / vanilla on the FP64 kernel
void routeGeneric(cpecan_ctx *ctx, const cpecan_hmm *hmm) {
    ctx->generic = hmm->sm_type;
    ctx->G.sm = hmm->sm_type;
    for (int i = 0; i < 9; i++) ctx->G.t3[i] = hmm->transitions[i];
    for (int i = 0; i < 5; i++) ctx->G.van[i] = hmm->vanilla[i];
}

int fillDevParams(cpecan_ctx *ctx, const cpecan_hmm *hmm, const cpecan_params *p, int mode) {
    if (p->diagonalExpansion < 0 || p->traceBackDiagonals < 1 ||
        p->minDiagsBetweenTraceBack < 2 || p->traceBackDiagonals + 1 >= p->minDiagsBetweenTraceBack) {
        ctx->err = "invalid banding parameters (impl/pairwiseAligner.c:880-884 of the reference)";
        return CPECAN_ERR_ARG;
    }
    int rc = fillMachine(ctx, hmm);
    if (rc != CPECAN_OK) return rc;
    if ((ctx->generic == CPECAN_SM_FOUR_STATE || ctx->generic == CPECAN_SM_ECHELON) && mode == CPECAN_MODE_EXPECTATION) {
        ctx->err = "expectations are not defined for the fourState / echelon machines (the reference has no update function for them)";
        return CPECAN_ERR_ARG;
    }
    // with an odd expansion band_construct's band edges move backwards and by two cells from one diagonal to the next
    // (impl/pairwiseAligner.c:98-170): k_align3's one-step band walk does not take that; the FP64 kernel, which reads
    // explicit band edges, does (and it is the one cpecan_cuda_set_exact_arithmetic asks for)
    if (!ctx->generic && (p->diagonalExpansion % 2 != 0 || ctx->exact)) routeGeneric(ctx, hmm);
    ctx->G.threshold = p->threshold;
    DevParams &P = ctx->P;
    P.threshold = (float) p->threshold;
    P.minDiags = (int) p->minDiagsBetweenTraceBack;
    P.tbDiags = (int) p->traceBackDiagonals;
    P.expansion = (int) p->diagonalExpansion;
    P.totalEvery = 10;
    P.rebaseEvery = 1;
    P.mode = mode;
    P.dbgLogP = getenv("CPECAN_DEBUG_LOGP") ? atoi(getenv("CPECAN_DEBUG_LOGP")) : 0;
    return CPECAN_OK;
}

// traceback points closer together than the zone of forward rows that keep all three states: keep them on every row
bool allSpec(const cpecan_ctx *ctx) { return ctx->P.minDiags - ctx->P.tbDiags - 1 < ctx->P.tbDiags + 3; }

void launchPrepX(cpecan_ctx *ctx, cudaStream_t s) {
    dim3 gx((unsigned) ctx->n, (unsigned) std::min(64, (ctx->stagedMaxLX + 256) / 256 + 1));
    k_prep_xparams<<<gx, 256, 0, s>>>(ctx->dItems.as<Item>(), ctx->dRefOff.as<long long>(), ctx->dRef.as<char>(),
                                     ctx->dModels.as<ModelTables>(), ctx->stagedScaled ? ctx->dScale.as<double>() : nullptr,
                                     ctx->dCentre.as<double>(), ctx->dXp.as<float4>(), ctx->machine, ctx->mToYNotX,
                                     ctx->dFlags.as<int>());
    ctx->timing.kernel_launches += 1;
}

// per-column records of the FP64 kernel (cpecan_generic.cuh); echelon has none
void launchPrepGeneric(cpecan_ctx *ctx, cudaStream_t s) {
    if (!ctx->generic) return;
    dim3 g((unsigned) ctx->n, (unsigned) std::min(64, (std::max(ctx->stagedMaxLX, ctx->stagedMaxLY) + 256) / 256 + 1));
    k_prep_generic<<<g, 256, 0, s>>>(ctx->dItems.as<Item>(), ctx->dRefOff.as<long long>(), ctx->dRef.as<char>(),
                                     ctx->dModels.as<ModelTables>(), ctx->stagedScaled ? ctx->dScale.as<double>() : nullptr,
                                     ctx->G, ctx->dColp.as<double>(),
                                     (ctx->generic == CPECAN_SM_VANILLA || ctx->generic == CPECAN_SM_ECHELON) ? ctx->dRowp.as<double>() : nullptr,
                                     ctx->stagedEvTot, ctx->dEvSrc.as<double>(), ctx->dEvSrcOff.as<long long>());
    ctx->timing.kernel_launches += 1;
}

}  // namespace

extern "C" {

int cpecan_cuda_init(int device, cpecan_ctx **ctx_out) {
    if (!ctx_out) return CPECAN_ERR_ARG;
    *ctx_out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0 || device < 0 || device >= count) return CPECAN_ERR_CUDA;
    cpecan_ctx *ctx = new cpecan_ctx();
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&ctx->prop, device) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return CPECAN_ERR_CUDA;
    }
    for (auto &e : ctx->ev) cudaEventCreate(&e);
    for (auto &st : ctx->bstream) cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    for (auto &e : ctx->bev) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->evBlock, cudaEventBlockingSync | cudaEventDisableTiming);
    setMaxSmemOnce(device, (int) ctx->prop.sharedMemPerBlockOptin);
    for (int c = 0; c < NCFG2; c++)
        for (int mach = 0; mach < 2; mach++)
            for (int sx = 0; sx < 2; sx++)
                for (int ex = 0; ex < 2; ex++)        // 0: ring too large for this device's shared memory
                    ctx->occ2[c][mach][sx][ex] = smemCfg2(c, mach, ex != 0) > ctx->prop.sharedMemPerBlockOptin ? 0 : occCfg2(c, mach, sx != 0, ex != 0);
    *ctx_out = ctx;
    return CPECAN_OK;
}

void cpecan_cuda_destroy(cpecan_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto &m : ctx->models) if (m.live) { cudaFree(m.match); cudaFree(m.gapy); cudaFree(m.gapx); cudaFree(m.hdpY); cudaFree(m.hdpSlope); cudaFree(m.hdpKmer); }
    DevBuf *bufs[] = { &ctx->dModels, &ctx->dItems, &ctx->dOut, &ctx->dRef, &ctx->dRefOff, &ctx->dEvSrc, &ctx->dEvSrcOff,
                       &ctx->dAnchors, &ctx->dScale, &ctx->dCentre, &ctx->dXp, &ctx->dEv, &ctx->dPairs, &ctx->dOrder,
                       &ctx->dQueue, &ctx->dScratch, &ctx->dTotals, &ctx->dCompact, &ctx->dCompactOff, &ctx->dExpect,
                       &ctx->dBits, &ctx->dTbs, &ctx->dFlags, &ctx->dBands, &ctx->dBandOff, &ctx->dColp, &ctx->dRowp };
    for (auto *b : bufs) b->release();
    for (auto &e : ctx->ev) cudaEventDestroy(e);
    for (auto &e : ctx->bev) cudaEventDestroy(e);
    if (ctx->evBlock) cudaEventDestroy(ctx->evBlock);
    if (ctx->ncclComm) { std::string e; Nccl *L = ncclLib(e); if (L && L->commDestroy) L->commDestroy(ctx->ncclComm); }
    for (auto &st : ctx->bstream) cudaStreamDestroy(st);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char *cpecan_cuda_last_error(cpecan_ctx *ctx) { return ctx ? ctx->err.c_str() : "no context (no usable CUDA device?)"; }

int cpecan_cuda_upload_model(cpecan_ctx *ctx, const double *match, const double *gapy, const double *gapx,
                             int32_t n_gapx, int32_t *model_id_out) {
    if (!ctx || !match || !gapy || !gapx || n_gapx <= 0 || !model_id_out) return CPECAN_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    const size_t tbl = (1 + 4096 * 5) * sizeof(double);
    Model m;
    m.n_gapx = n_gapx;
    CK(cudaMalloc(&m.match, tbl));
    CK(cudaMalloc(&m.gapy, tbl));
    CK(cudaMalloc(&m.gapx, n_gapx * sizeof(double)));
    CK(cudaMemcpy(m.match, match, tbl, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(m.gapy, gapy, tbl, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(m.gapx, gapx, n_gapx * sizeof(double), cudaMemcpyHostToDevice));
    m.live = true;
    size_t slot = 0;
    while (slot < ctx->models.size() && ctx->models[slot].live) slot++;      // reuse a released slot
    if (slot == ctx->models.size()) ctx->models.push_back(m); else ctx->models[slot] = m;
    ctx->modelsDirty = true;
    *model_id_out = (int32_t) slot;
    return CPECAN_OK;
}

int cpecan_cuda_upload_hdp(cpecan_ctx *ctx, double grid_start, double grid_stop, int64_t grid_length, int32_t n_distr,
                           const double *density, const double *slopes, const int32_t *kmer_distr, int32_t *model_id_out) {
    if (!ctx || !density || !slopes || !kmer_distr || !model_id_out) return CPECAN_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (grid_length < 2 || grid_length > 0x7fffffff || n_distr < 1 || !(grid_start < grid_stop)) { ctx->err = "upload_hdp: bad grid or no distribution"; return CPECAN_ERR_ARG; }
    for (int k = 0; k < 4096; k++) if (kmer_distr[k] < -1 || kmer_distr[k] >= n_distr) { ctx->err = "upload_hdp: k-mer points outside the distributions"; return CPECAN_ERR_ARG; }
    CK(cudaSetDevice(ctx->device));
    Model m;
    m.hdp = true;
    m.hdpLen = (int) grid_length;
    // linspace (impl/hdp_math_utils.c:497-510): grid[i] = start + i * dx, the last point is `stop` itself
    m.step = (grid_stop - grid_start) / (double) (grid_length - 1);
    m.x0 = grid_start + 0 * m.step;
    m.x1 = grid_length == 2 ? grid_stop : grid_start + 1 * m.step;
    m.xn = grid_stop;
    const size_t bytes = (size_t) n_distr * (size_t) grid_length * sizeof(double);
    CK(cudaMalloc(&m.hdpY, bytes));
    CK(cudaMalloc(&m.hdpSlope, bytes));
    CK(cudaMalloc(&m.hdpKmer, 4096 * sizeof(int)));
    CK(cudaMemcpy(m.hdpY, density, bytes, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(m.hdpSlope, slopes, bytes, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(m.hdpKmer, kmer_distr, 4096 * sizeof(int), cudaMemcpyHostToDevice));
    m.live = true;
    size_t slot = 0;
    while (slot < ctx->models.size() && ctx->models[slot].live) slot++;
    if (slot == ctx->models.size()) ctx->models.push_back(m); else ctx->models[slot] = m;
    ctx->modelsDirty = true;
    *model_id_out = (int32_t) slot;
    return CPECAN_OK;
}

int cpecan_cuda_release_model(cpecan_ctx *ctx, int32_t model_id) {
    if (!ctx) return CPECAN_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (model_id < 0 || model_id >= (int32_t) ctx->models.size() || !ctx->models[model_id].live) { ctx->err = "release_model: bad model id"; return CPECAN_ERR_ARG; }
    CK(cudaSetDevice(ctx->device));
    if (ctx->running) { ctx->err = "release_model: a run is in flight"; return CPECAN_ERR_ARG; }
    CK(waitStream(ctx, ctx->stream));
    Model &m = ctx->models[model_id];
    cudaFree(m.match); cudaFree(m.gapy); cudaFree(m.gapx); cudaFree(m.hdpY); cudaFree(m.hdpSlope); cudaFree(m.hdpKmer);
    m = Model();
    ctx->modelsDirty = true;
    return CPECAN_OK;
}

int cpecan_cuda_update_model(cpecan_ctx *ctx, int32_t model_id, const double *match, const double *gapy,
                             const double *gapx) {
    if (!ctx) return CPECAN_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (model_id < 0 || model_id >= (int32_t) ctx->models.size() || !ctx->models[model_id].live) { ctx->err = "update_model: bad model id"; return CPECAN_ERR_ARG; }
    CK(cudaSetDevice(ctx->device));
    CK(waitStream(ctx, ctx->stream));
    const size_t tbl = (1 + 4096 * 5) * sizeof(double);
    Model &m = ctx->models[model_id];
    if (m.hdp) { ctx->err = "update_model: an HDP model has no pore-model tables"; return CPECAN_ERR_ARG; }
    if (match) CK(cudaMemcpy(m.match, match, tbl, cudaMemcpyHostToDevice));
    if (gapy) CK(cudaMemcpy(m.gapy, gapy, tbl, cudaMemcpyHostToDevice));
    if (gapx) CK(cudaMemcpy(m.gapx, gapx, m.n_gapx * sizeof(double), cudaMemcpyHostToDevice));
    return CPECAN_OK;
}

}  // extern "C"

namespace {

// On every error exit of stage() the context holds NO staged batch (run_staged / fetch_staged then do nothing).
struct StageGuard {
    cpecan_ctx *c;
    bool ok = false;
    ~StageGuard() { if (!ok) { c->n = 0; for (auto &b : c->buckets) { b.order.clear(); b.nCta = 0; } } }
};

int stageL(cpecan_ctx *ctx, const cpecan_hmm *hmm, const cpecan_params *params, int32_t mode, const cpecan_batch *B,
           int64_t pair_cap_total, bool wantTotals) {
    CK(cudaSetDevice(ctx->device));
    if (ctx->running) { ctx->err = "stage: a run is in flight"; return CPECAN_ERR_ARG; }
    StageGuard guard{ctx};
    ctx->wantTotals = wantTotals;
    int rc = fillDevParams(ctx, hmm, params, mode);
    if (rc != CPECAN_OK) return rc;
    const int64_t n = B->n_items;
    ctx->n = n;
    ctx->mode = mode;
    ctx->timing = cpecan_timing{};
    if (n == 0) { guard.ok = true; return CPECAN_OK; }
    cudaStream_t s = ctx->stream;
    const bool wantHdp = ctx->generic == CPECAN_SM_THREE_STATE_HDP;
    const int needGapx = wantHdp ? 0 : (ctx->generic == CPECAN_SM_ECHELON ? 60 : (ctx->machine ? 60 : 4096));

    if (ctx->modelsDirty) {
        std::vector<ModelTables> mt(ctx->models.size());
        for (size_t i = 0; i < mt.size(); i++) {
            const Model &m = ctx->models[i];
            mt[i].match = m.match; mt[i].gapy = m.gapy; mt[i].gapx = m.gapx; mt[i].n_gapx = m.n_gapx;
            mt[i].hdp_y = m.hdpY; mt[i].hdp_slope = m.hdpSlope; mt[i].hdp_kmer = m.hdpKmer; mt[i].hdp_len = m.hdpLen;
            mt[i].hdp_x0 = m.x0; mt[i].hdp_x1 = m.x1; mt[i].hdp_xn = m.xn; mt[i].hdp_step = m.step;
        }
        CK(ctx->dModels.ensure(std::max<size_t>(1, mt.size()) * sizeof(ModelTables)));
        if (!mt.empty()) CK(cudaMemcpy(ctx->dModels.p, mt.data(), mt.size() * sizeof(ModelTables), cudaMemcpyHostToDevice));
        ctx->modelsDirty = false;
    }

    // ---- host-side layout: prefix sums, per-item records -------------------------------------------------
    ctx->hItems.resize(n);
    std::vector<double> centre(n);
    long long xpTot = 0, evTot = 0, totTot = 0, bitsTot = 0, tbTot = 0, bandTot = 0;
    std::vector<long long> bandOff(ctx->generic ? n : 0);
    const long long tbSpacing = std::max<long long>(1, ctx->P.minDiags - ctx->P.tbDiags - 1);
    int64_t sumLY = 0;
    for (int64_t i = 0; i < n; i++) sumLY += (B->ev_off[i + 1] - B->ev_off[i]) + 32;
    long long pairTot = 0;
    ctx->hTotOff.assign(n + 1, 0);
    int maxLX = 0, maxLY = 0;
    for (int64_t i = 0; i < n; i++) {
        Item &it = ctx->hItems[i];
        const int64_t refLen = B->ref_off[i + 1] - B->ref_off[i];
        const int64_t lX = refLen >= 5 ? refLen - 5 : 0, lY = B->ev_off[i + 1] - B->ev_off[i];
        if (B->model_id[i] < 0 || B->model_id[i] >= (int) ctx->models.size() || !ctx->models[B->model_id[i]].live) { ctx->err = "bad model id"; return CPECAN_ERR_ARG; }
        if (ctx->models[B->model_id[i]].hdp != wantHdp) {
            ctx->err = "the threeStateHdp machine takes models from cpecan_cuda_upload_hdp, the other machines from cpecan_cuda_upload_model";
            return CPECAN_ERR_ARG;
        }
        if (ctx->models[B->model_id[i]].n_gapx < needGapx) {
            ctx->err = "model has too few gap-X entries for this state machine (threeState: 4096 log-probabilities, vanilla: 60 skip bins)";
            return CPECAN_ERR_ARG;
        }
        it.xp_off = xpTot; it.ev_off = evTot; it.an_off = B->anchor_off[i];
        it.lX = (int) lX; it.lY = (int) lY; it.nA = (int) (B->anchor_off[i + 1] - B->anchor_off[i]);
        it.flags = B->ragged ? B->ragged[i] : 0;
        it.model_id = B->model_id[i];
        // plan outputs: 2 bits per diagonal, and the traceback points (at least minDiags - tbDiags - 1 diagonals apart)
        it.pad0 = (int) bitsTot; it.pad1 = (int) tbTot;
        bitsTot += ((lX + lY) >> 4) + 1; tbTot += (lX + lY) / tbSpacing + 2;
        if (ctx->generic) { bandOff[i] = bandTot; bandTot += lX + lY + 1; }
        if (bitsTot > 0x7fffffffLL || tbTot > 0x7fffffffLL) { ctx->err = "batch too large for the plan buffers"; return CPECAN_ERR_ARG; }
        // pair capacity: proportional share of the caller's buffer
        long long cap = (long long) ((long double) pair_cap_total * (long double) (lY + 32) / (long double) sumLY);
        it.pair_off = pairTot; it.pair_cap = (int) std::min<long long>(cap, 0x7fffffff);
        pairTot += it.pair_cap;
        it.tot_off = ctx->wantTotals ? totTot : -1;
        ctx->hTotOff[i] = totTot;
        totTot += 3 * (lX + lY + 1);   // total, term 1, term 2 per diagonal
        xpTot += lX + 2; evTot += lY + 1;
        maxLX = std::max(maxLX, (int) lX); maxLY = std::max(maxLY, (int) lY);
        const double sc = B->scale ? B->scale[5 * i] : 1.0, sh = B->scale ? B->scale[5 * i + 1] : 0.0;
        centre[i] = 68.0 * sc + sh;
    }
    ctx->hTotOff[n] = totTot;
    ctx->pairCapTotal = pairTot;
    ctx->totalsLen = totTot;
    const int64_t refBytes = B->ref_off[n], nEv = B->ev_off[n], nAn = B->anchor_off[n];

    // ---- H2D ---------------------------------------------------------------------------------------------
    CK(cudaEventRecord(ctx->ev[0], s));
    CK(ctx->dItems.ensure(n * sizeof(Item)));
    CK(ctx->dOut.ensure(n * sizeof(ItemOut)));
    CK(ctx->dRef.ensure(refBytes + 16));
    CK(ctx->dRefOff.ensure((n + 1) * sizeof(long long)));
    CK(ctx->dEvSrc.ensure(std::max<int64_t>(1, nEv) * 3 * sizeof(double)));
    CK(ctx->dEvSrcOff.ensure((n + 1) * sizeof(long long)));
    CK(ctx->dAnchors.ensure(std::max<int64_t>(1, nAn) * 2 * sizeof(long long)));
    CK(ctx->dCentre.ensure(n * sizeof(double)));
    if (!ctx->generic) {
        CK(ctx->dXp.ensure(xpTot * (ctx->machine ? 4 : 3) * sizeof(float4)));
        CK(ctx->dEv.ensure(evTot * sizeof(float4)));
    } else {
        CK(ctx->dColp.ensure(std::max<long long>(1, xpTot * gen_ncol(ctx->generic)) * sizeof(double)));
        if (ctx->generic == CPECAN_SM_VANILLA) CK(ctx->dRowp.ensure(evTot * sizeof(double)));
        if (ctx->generic == CPECAN_SM_ECHELON) CK(ctx->dRowp.ensure(2 * evTot * sizeof(double)));
        ctx->stagedEvTot = evTot;
    }
    CK(ctx->dPairs.ensure(std::max<long long>(1, pairTot) * 3 * sizeof(int)));
    CK(ctx->dBits.ensure(std::max<long long>(1, bitsTot) * sizeof(unsigned)));
    CK(ctx->dTbs.ensure(std::max<long long>(1, tbTot) * sizeof(int)));
    CK(ctx->dFlags.ensure(n * sizeof(int)));
    CK(cudaMemsetAsync(ctx->dFlags.p, 0, n * sizeof(int), s));
    if (ctx->generic) {
        CK(ctx->dBands.ensure(std::max<long long>(1, bandTot) * sizeof(int2)));
        CK(ctx->dBandOff.ensure(n * sizeof(long long)));
        CK(cudaMemcpyAsync(ctx->dBandOff.p, bandOff.data(), n * sizeof(long long), cudaMemcpyHostToDevice, s));
    }
    CK(cudaMemcpyAsync(ctx->dItems.p, ctx->hItems.data(), n * sizeof(Item), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->dRef.p, B->ref, refBytes, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->dRefOff.p, B->ref_off, (n + 1) * sizeof(long long), cudaMemcpyHostToDevice, s));
    if (nEv) CK(cudaMemcpyAsync(ctx->dEvSrc.p, B->events, nEv * 3 * sizeof(double), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->dEvSrcOff.p, B->ev_off, (n + 1) * sizeof(long long), cudaMemcpyHostToDevice, s));
    if (nAn) CK(cudaMemcpyAsync(ctx->dAnchors.p, B->anchors, nAn * 2 * sizeof(long long), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->dCentre.p, centre.data(), n * sizeof(double), cudaMemcpyHostToDevice, s));
    ctx->stagedScaled = false;
    if (B->scale) {
        CK(ctx->dScale.ensure(n * 5 * sizeof(double)));
        CK(cudaMemcpyAsync(ctx->dScale.p, B->scale, n * 5 * sizeof(double), cudaMemcpyHostToDevice, s));
        ctx->stagedScaled = true;
    }
    ctx->stagedMaxLX = maxLX; ctx->stagedMaxLY = maxLY;
    ctx->timing.h2d_bytes = n * (int64_t) sizeof(Item) + refBytes + 2 * (n + 1) * 8 + nEv * 24 + nAn * 16 + n * 8 + (B->scale ? n * 40 : 0);
    CK(cudaEventRecord(ctx->ev[1], s));

    // ---- preparation kernels -----------------------------------------------------------------------------
    if (!ctx->generic) {
        dim3 ge((unsigned) n, (unsigned) std::min(64, (maxLY + 256) / 256 + 1));
        k_prep_events<<<ge, 256, 0, s>>>(ctx->dItems.as<Item>(), ctx->dEvSrcOff.as<long long>(), ctx->dEvSrc.as<double>(),
                                        ctx->dCentre.as<double>(), ctx->dEv.as<float4>());
        ctx->timing.kernel_launches += 1;
        launchPrepX(ctx, s);
    } else launchPrepGeneric(ctx, s);
    CK(cudaEventRecord(ctx->ev[2], s));

    // ---- plan: band cells, widest diagonal, longest run of live forward rows -------------------------------
    k_plan3<<<(unsigned) ((n + 63) / 64), 64, 0, s>>>(ctx->dItems.as<Item>(), (int) n, ctx->dAnchors.as<long long>(), ctx->P, ctx->dOut.as<ItemOut>(),
                                                    ctx->dBits.as<unsigned>(), ctx->dTbs.as<int>(), ctx->dFlags.as<int>(),
                                                    ctx->generic ? ctx->dBands.as<int2>() : nullptr, ctx->dBandOff.as<long long>());
    ctx->timing.kernel_launches += 1;
    ctx->hOut.resize(n);
    CK(cudaMemcpyAsync(ctx->hOut.data(), ctx->dOut.p, n * sizeof(ItemOut), cudaMemcpyDeviceToHost, s));
    CK(cudaEventRecord(ctx->ev[3], s));
    CK(waitStream(ctx, s));
    {
        float ms = 0;
        cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]); ctx->timing.h2d_ms = ms;
        cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[2]); ctx->timing.prep_ms = ms;
        cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]); ctx->timing.plan_ms = ms;
    }

    // ---- bucket by the ring size an alignment needs, order largest first ---------------------------------------
    for (auto &b : ctx->buckets) { b.order.clear(); b.nCta = 0; b.ringRows = 0; }
    const int sx = ctx->hasSX ? 1 : 0, ex = mode == CPECAN_MODE_EXPECTATION ? 1 : 0;
    int occGenCache[NCFG2] = {};
    if (ctx->generic) for (int b = 0; b < NCFG2; b++) occGenCache[b] = occGen(b, ctx->generic, ctx->prop.sharedMemPerBlockOptin);
    const int genS = genStates(ctx->generic);
    int64_t cells = 0;
    for (int64_t i = 0; i < n; i++) {
        const ItemOut &o = ctx->hOut[i];
        cells += o.band_cells;
        int b = 0;
        // the FP64 kernel keeps every live diagonal in a buffer of its own: the ring only has to hold one diagonal and the
        // two positions either side that a sweep clears
        const int margin = ctx->generic ? 8 : CFG2_MARGIN;
        while (b < NCFG2 && (cfg2N(b) < o.max_width + margin ||
                             (ctx->generic ? occGenCache[b] : ctx->occ2[b][ctx->machine][sx][ex]) == 0)) b++;
        if (b == NCFG2) { ctx->err = "band wider than the widest ring this device's shared memory holds"; return CPECAN_ERR_BAND_TOO_WIDE; }
        ctx->buckets[b].order.push_back((int) i);
        ctx->buckets[b].ringRows = std::max(ctx->buckets[b].ringRows, o.max_rows);
    }
    ctx->timing.band_cells = cells;
    size_t scratch4 = 0, orderInts = 0;
    std::vector<int> orderAll;
    orderAll.reserve(n);
    for (int b = 0; b < NCFG2; b++) {
        Bucket &bk = ctx->buckets[b];
        if (bk.order.empty()) continue;
        std::stable_sort(bk.order.begin(), bk.order.end(), [&](int a, int c) { return ctx->hOut[a].band_cells > ctx->hOut[c].band_cells; });
        int occ = std::max(1, ctx->generic ? occGenCache[b] : ctx->occ2[b][ctx->machine][sx][ex]);
        if (ctx->occCap > 0) occ = std::min(occ, ctx->occCap);       // cpecan_cuda_set_resident_warps
        bk.nCta = (int) std::min<int64_t>((int64_t) bk.order.size(), (int64_t) ctx->prop.multiProcessorCount * occ);
        // forward rows: one float4 record per ring position and row; then the second plane (float2): posteriors keep
        // the other two forward states of 2 * (tbDiags + 3) zone rows and 2 rows per 10 diagonals, the E-step the two
        // emissions of every row
        bk.specRows = (ex || allSpec(ctx)) ? bk.ringRows : 2 * (ctx->P.tbDiags + 3) + 2 * (bk.ringRows / 10 + 2);
        bk.stride = (long long) bk.ringRows * cfg2N(b) + ((long long) bk.specRows * cfg2N(b) + 1) / 2;
        if (ctx->generic) bk.stride = ((long long) bk.ringRows * cfg2N(b) * genS + 1) / 2;     // S doubles per cell, in float4 units
        bk.scratchOff = scratch4; scratch4 += (size_t) bk.stride * bk.nCta;
        bk.orderOff = orderInts; orderInts += bk.order.size();
        orderAll.insert(orderAll.end(), bk.order.begin(), bk.order.end());
        ctx->timing.warps_per_item = 1;
        ctx->timing.ctas = bk.nCta;
    }
    CK(ctx->dScratch.ensure(std::max<size_t>(1, scratch4) * sizeof(float4)));
    CK(ctx->dOrder.ensure(std::max<size_t>(1, orderInts) * sizeof(int)));
    CK(ctx->dQueue.ensure(16 * sizeof(int)));
    CK(ctx->dExpect.ensure(CPECAN_N_EXPECT * sizeof(double)));
    CK(cudaMemcpyAsync(ctx->dOrder.p, orderAll.data(), orderAll.size() * sizeof(int), cudaMemcpyHostToDevice, s));
    if (ctx->wantTotals) CK(ctx->dTotals.ensure(std::max<int64_t>(1, totTot) * sizeof(double)));
    CK(waitStream(ctx, s));
    guard.ok = true;
    return CPECAN_OK;
}

}  // namespace

extern "C" {

int cpecan_cuda_stage(cpecan_ctx *ctx, const cpecan_hmm *hmm, const cpecan_params *params, int32_t mode,
                      const cpecan_batch *B, int64_t pair_cap_total) {
    if (!ctx || !hmm || !params || !B || B->n_items < 0) return CPECAN_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    return stageL(ctx, hmm, params, mode, B, pair_cap_total, false);
}

int cpecan_cuda_restage_model(cpecan_ctx *ctx, const cpecan_hmm *hmm) {
    if (!ctx || !hmm) return CPECAN_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    if (ctx->running) { ctx->err = "restage_model: a run is in flight"; return CPECAN_ERR_ARG; }
    const int machineBefore = ctx->machine, genericBefore = ctx->generic;
    int rc = fillMachine(ctx, hmm);
    if (rc != CPECAN_OK) return rc;
    if (genericBefore && !ctx->generic) routeGeneric(ctx, hmm);      // the staged batch runs on the FP64 kernel (odd expansion / exact)
    if (ctx->n == 0) return CPECAN_OK;
    if (ctx->machine != machineBefore || ctx->generic != genericBefore) { ctx->err = "restage_model: the staged batch was prepared for the other state machine"; return CPECAN_ERR_ARG; }
    if (ctx->generic) {                                               // transitions travel as kernel arguments; the vanilla machine's
        launchPrepGeneric(ctx, ctx->stream);                          // per-column log transitions and all cached logarithms are re-derived
        CK(cudaGetLastError());
        CK(waitStream(ctx, ctx->stream));
        return CPECAN_OK;
    }
    launchPrepX(ctx, ctx->stream);
    CK(cudaGetLastError());
    CK(waitStream(ctx, ctx->stream));
    return CPECAN_OK;
}

}  // extern "C"

namespace {

int runAsyncL(cpecan_ctx *ctx) {
    CK(cudaSetDevice(ctx->device));
    if (ctx->n == 0) return CPECAN_OK;
    if (ctx->running) { ctx->err = "run_staged_async: the previous run was not waited for"; return CPECAN_ERR_ARG; }
    cudaStream_t s = ctx->stream;
    CK(cudaMemsetAsync(ctx->dQueue.p, 0, 16 * sizeof(int), s));
    if (ctx->mode == CPECAN_MODE_EXPECTATION) CK(cudaMemsetAsync(ctx->dExpect.p, 0, CPECAN_N_EXPECT * sizeof(double), s));
    if (ctx->wantTotals) CK(cudaMemsetAsync(ctx->dTotals.p, 0xff, ctx->totalsLen * sizeof(double), s));   // all-ones = NaN
    CK(cudaEventRecord(ctx->ev[4], s));
    int launches = 0;
    for (int b = 0; b < NCFG2; b++) {
        Bucket &bk = ctx->buckets[b];
        if (bk.order.empty()) continue;
        KernelArgs3 a;
        a.items = ctx->dItems.as<Item>();
        a.order = ctx->dOrder.as<int>() + bk.orderOff;
        a.n_items = (int) bk.order.size();
        a.queue = ctx->dQueue.as<int>() + b;
        a.xparams = ctx->dXp.as<float4>();
        a.events = ctx->dEv.as<float4>();
        a.bits = ctx->dBits.as<unsigned>();
        a.tbs = ctx->dTbs.as<int>();
        a.flags = ctx->dFlags.as<int>();
        a.scratch = ctx->dScratch.as<float4>() + bk.scratchOff;
        a.scratch_stride = bk.stride;
        a.ring_rows = bk.ringRows;
        a.spec_rows = bk.specRows;
        a.all_spec = allSpec(ctx) ? 1 : 0;
        a.ringN = cfg2N(b);
        a.zero = 0;
        a.pairs = ctx->dPairs.as<int>();
        a.out = ctx->dOut.as<ItemOut>();
        a.totals = ctx->wantTotals ? ctx->dTotals.as<double>() : nullptr;
        a.expect = ctx->dExpect.as<double>();
        a.P = ctx->P;
        CK(cudaStreamWaitEvent(ctx->bstream[b], ctx->ev[4], 0));
        if (ctx->generic) {
            KernelArgsG g;
            g.items = a.items; g.order = a.order; g.n_items = a.n_items; g.queue = a.queue;
            g.ref = ctx->dRef.as<char>(); g.ref_off = ctx->dRefOff.as<long long>();
            g.events = ctx->dEvSrc.as<double>(); g.ev_src_off = ctx->dEvSrcOff.as<long long>();
            g.models = ctx->dModels.as<ModelTables>(); g.scale = ctx->stagedScaled ? ctx->dScale.as<double>() : nullptr;
            g.bands = ctx->dBands.as<int2>(); g.band_off = ctx->dBandOff.as<long long>(); g.tbs = a.tbs; g.flags = a.flags;
            g.scratch = reinterpret_cast<double *>(a.scratch); g.scratch_stride = bk.stride * 2;
            g.ring_rows = bk.ringRows; g.ringN = cfg2N(b);
            g.pairs = a.pairs; g.out = a.out; g.totals = a.totals; g.expect = a.expect; g.P = ctx->P; g.G = ctx->G;
            g.colp = ctx->dColp.as<double>(); g.rowp = ctx->dRowp.as<double>(); g.row2 = ctx->stagedEvTot;
            dispatchGen(ctx->generic, [&](auto k, int S) {
                k<<<bk.nCta, 32, generic_smem_bytes(cfg2N(b), S), ctx->bstream[b]>>>(g); return 0; });
        } else
        launchCfg2(b, ctx->machine, ctx->hasSX, ctx->mode == CPECAN_MODE_EXPECTATION, a, bk.nCta, ctx->bstream[b]);
        CK(cudaEventRecord(ctx->bev[b], ctx->bstream[b]));
        CK(cudaStreamWaitEvent(s, ctx->bev[b], 0));
        launches++;
    }
    CK(cudaEventRecord(ctx->ev[5], s));
    CK(cudaGetLastError());
    ctx->timing.kernel_launches += launches;
    ctx->running = true;
    return CPECAN_OK;
}

int waitL(cpecan_ctx *ctx) {
    if (!ctx->running) return CPECAN_OK;
    CK(cudaSetDevice(ctx->device));
    ctx->running = false;
    CK(waitStream(ctx, ctx->stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev[4], ctx->ev[5]);
    ctx->timing.align_ms = ms;
    return CPECAN_OK;
}

int fetchL(cpecan_ctx *ctx, int32_t *pairs_out, cpecan_result *results) {
    CK(cudaSetDevice(ctx->device));
    const int64_t n = ctx->n;
    if (n == 0) return CPECAN_OK;
    cudaStream_t s = ctx->stream;
    CK(cudaEventRecord(ctx->ev[6], s));
    // the plan wrote band_cells / max_width into dOut, the align kernel the rest
    std::vector<ItemOut> out(n);
    CK(cudaMemcpyAsync(out.data(), ctx->dOut.p, n * sizeof(ItemOut), cudaMemcpyDeviceToHost, s));
    CK(waitStream(ctx, s));
    std::vector<long long> dst(n + 1, 0);
    for (int64_t i = 0; i < n; i++) {
        const int np = std::min(out[i].n_pairs, ctx->hItems[i].pair_cap);
        dst[i + 1] = dst[i] + np;
        results[i].n_pairs = out[i].n_pairs;
        results[i].pair_off = dst[i];
        results[i].band_cells = out[i].band_cells;
        results[i].total_logprob = out[i].total_logprob;
        results[i].status = out[i].status;
        results[i].n_tracebacks = out[i].n_tracebacks;
    }
    ctx->timing.d2h_bytes = n * (int64_t) sizeof(ItemOut);
    if (pairs_out && dst[n] > 0) {
        CK(ctx->dCompactOff.ensure((n + 1) * sizeof(long long)));
        // the packed copy goes into the raw-event staging buffer when it fits: that one is dead once the batch is
        // prepared (24 B per event against ~13 B of aligned pairs), and a 100k-read batch has no 10 GB to spare
        const size_t packedBytes = (size_t) dst[n] * 3 * sizeof(int);
        int *packed = ctx->dEvSrc.as<int>();
        if (packedBytes > ctx->dEvSrc.cap || ctx->generic) { CK(ctx->dCompact.ensure(packedBytes)); packed = ctx->dCompact.as<int>(); }
        CK(cudaMemcpyAsync(ctx->dCompactOff.p, dst.data(), (n + 1) * sizeof(long long), cudaMemcpyHostToDevice, s));
        k_compact<<<(unsigned) n, 128, 0, s>>>(ctx->dItems.as<Item>(), ctx->dOut.as<ItemOut>(), ctx->dCompactOff.as<long long>(), (int) n,
                                              ctx->dPairs.as<int>(), packed);
        ctx->timing.kernel_launches += 1;
        CK(cudaMemcpyAsync(pairs_out, packed, packedBytes, cudaMemcpyDeviceToHost, s));
        ctx->timing.d2h_bytes += dst[n] * 12;
    }
    CK(cudaEventRecord(ctx->ev[7], s));
    CK(waitStream(ctx, s));
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev[6], ctx->ev[7]);
    ctx->timing.d2h_ms = ms;
    ctx->timing.total_ms = ctx->timing.h2d_ms + ctx->timing.prep_ms + ctx->timing.plan_ms + ctx->timing.align_ms + ctx->timing.d2h_ms;
    return CPECAN_OK;
}

int fetchExpectL(cpecan_ctx *ctx, double *expectations_out) {
    CK(cudaSetDevice(ctx->device));
    if (ctx->mode != CPECAN_MODE_EXPECTATION) { ctx->err = "fetch_expectations: the staged batch is not in expectation mode"; return CPECAN_ERR_ARG; }
    const int len = ctx->machine ? CPECAN_N_EXPECT_VANILLA : CPECAN_N_EXPECT;
    std::vector<double> tmp(len);
    CK(cudaMemcpyAsync(tmp.data(), ctx->dExpect.p, len * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(waitStream(ctx, ctx->stream));
    for (int i = 0; i < len; i++) expectations_out[i] += tmp[i];
    ctx->timing.d2h_bytes += len * 8;
    return CPECAN_OK;
}

}  // namespace

extern "C" {

int cpecan_cuda_run_staged_async(cpecan_ctx *ctx) {
    if (!ctx) return CPECAN_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    return runAsyncL(ctx);
}

int cpecan_cuda_wait(cpecan_ctx *ctx) {
    if (!ctx) return CPECAN_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    return waitL(ctx);
}

int cpecan_cuda_run_staged(cpecan_ctx *ctx) {
    if (!ctx) return CPECAN_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    int rc = runAsyncL(ctx);
    return rc == CPECAN_OK ? waitL(ctx) : rc;
}

int cpecan_cuda_fetch_staged(cpecan_ctx *ctx, int32_t *pairs_out, cpecan_result *results) {
    if (!ctx || !results) return CPECAN_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    return fetchL(ctx, pairs_out, results);
}

// One lock for the whole call: two host threads sharing a context (the two strands of vanillaAlign.c:737-790) serialise
// here instead of interleaving their stage / run / fetch steps.
int cpecan_cuda_align_batch(cpecan_ctx *ctx, const cpecan_hmm *hmm, const cpecan_params *params, int32_t mode,
                            const cpecan_batch *batch, int32_t *pairs_out, int64_t pair_cap_total,
                            cpecan_result *results, double *totals_out, const int64_t *tot_off) {
    if (!ctx) return CPECAN_ERR_CUDA;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!hmm || !params || !batch || batch->n_items < 0 || !results) { ctx->err = "align_batch: null argument"; return CPECAN_ERR_ARG; }
    if (mode != CPECAN_MODE_POSTERIOR && mode != CPECAN_MODE_UNBANDED) { ctx->err = "align_batch: mode must be POSTERIOR or UNBANDED"; return CPECAN_ERR_ARG; }
    int rc = stageL(ctx, hmm, params, mode, batch, pair_cap_total, totals_out != nullptr);
    if (rc == CPECAN_OK) rc = runAsyncL(ctx);
    if (rc == CPECAN_OK) rc = waitL(ctx);
    if (rc == CPECAN_OK) rc = fetchL(ctx, pairs_out, results);
    if (rc == CPECAN_OK && totals_out && ctx->n > 0) {
        // debug totals: copy each item's slice to the caller's layout
        std::vector<double> tmp(ctx->totalsLen);
        CK(cudaMemcpy(tmp.data(), ctx->dTotals.p, ctx->totalsLen * sizeof(double), cudaMemcpyDeviceToHost));
        for (int64_t i = 0; i < ctx->n; i++) {
            const int64_t len = ctx->hTotOff[i + 1] - ctx->hTotOff[i];
            const int64_t dstOff = tot_off ? tot_off[i] : ctx->hTotOff[i];
            memcpy(totals_out + dstOff, tmp.data() + ctx->hTotOff[i], len * sizeof(double));
        }
    }
    ctx->wantTotals = false;
    return rc;
}

int cpecan_cuda_expectations_batch(cpecan_ctx *ctx, const cpecan_hmm *hmm, const cpecan_params *params,
                                   const cpecan_batch *batch, double *expectations_out, cpecan_result *results) {
    if (!ctx) return CPECAN_ERR_CUDA;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!hmm || !params || !batch || batch->n_items < 0) { ctx->err = "expectations_batch: null argument"; return CPECAN_ERR_ARG; }
    if (!expectations_out || !results) { ctx->err = "expectations_batch: null output"; return CPECAN_ERR_ARG; }
    int rc = stageL(ctx, hmm, params, CPECAN_MODE_EXPECTATION, batch, 0, false);
    if (rc == CPECAN_OK) rc = runAsyncL(ctx);
    if (rc == CPECAN_OK) rc = waitL(ctx);
    if (rc == CPECAN_OK) rc = fetchL(ctx, nullptr, results);
    if (rc == CPECAN_OK && ctx->n > 0) rc = fetchExpectL(ctx, expectations_out);
    return rc;
}

int cpecan_cuda_hdp_expectations_batch(cpecan_ctx *ctx, const cpecan_hmm *hmm, const cpecan_params *params,
                                       const cpecan_batch *batch, double *expectations_out, int32_t *assignments_out,
                                       int64_t assignment_cap_total, cpecan_result *results) {
    if (!ctx) return CPECAN_ERR_CUDA;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!hmm || !params || !batch || batch->n_items < 0) { ctx->err = "hdp_expectations_batch: null argument"; return CPECAN_ERR_ARG; }
    if (!expectations_out || !results || (!assignments_out && assignment_cap_total > 0)) { ctx->err = "hdp_expectations_batch: null output"; return CPECAN_ERR_ARG; }
    if (hmm->sm_type != CPECAN_SM_THREE_STATE_HDP) { ctx->err = "hdp_expectations_batch: the state machine is not threeStateHdp"; return CPECAN_ERR_ARG; }
    int rc = stageL(ctx, hmm, params, CPECAN_MODE_EXPECTATION, batch, assignment_cap_total, false);
    if (rc == CPECAN_OK) rc = runAsyncL(ctx);
    if (rc == CPECAN_OK) rc = waitL(ctx);
    if (rc == CPECAN_OK) rc = fetchL(ctx, assignments_out, results);
    if (rc == CPECAN_OK && ctx->n > 0) rc = fetchExpectL(ctx, expectations_out);
    return rc;
}

int cpecan_cuda_fetch_expectations(cpecan_ctx *ctx, double *expectations_out) {
    if (!ctx || !expectations_out) return CPECAN_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    return fetchExpectL(ctx, expectations_out);
}

int cpecan_cuda_nccl_unique_id(void *id_out) {
    if (!id_out) return CPECAN_ERR_ARG;
    std::string err;
    Nccl *L = ncclLib(err);
    if (!L) return CPECAN_ERR_CUDA;
    return L->getUniqueId(id_out) == 0 ? CPECAN_OK : CPECAN_ERR_CUDA;
}

int cpecan_cuda_nccl_init(cpecan_ctx *ctx, int32_t n_ranks, int32_t rank, const void *id) {
    if (!ctx || !id || n_ranks < 1 || rank < 0 || rank >= n_ranks) return CPECAN_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    Nccl *L = ncclLib(ctx->err);
    if (!L) return CPECAN_ERR_CUDA;
    if (ctx->ncclComm) { if (L->commDestroy) L->commDestroy(ctx->ncclComm); ctx->ncclComm = nullptr; }
    // ncclResult_t ncclCommInitRank(ncclComm_t *comm, int nranks, ncclUniqueId commId, int rank): the id goes by value
    typedef int (*InitFn)(void **, int, NcclId, int);
    NcclId nid;
    memcpy(nid.b, id, sizeof(nid.b));
    const int rc = ((InitFn) L->commInitRank)(&ctx->ncclComm, n_ranks, nid, rank);
    if (rc != 0) { ctx->err = std::string("ncclCommInitRank: ") + (L->getErrorString ? L->getErrorString(rc) : "error"); ctx->ncclComm = nullptr; return CPECAN_ERR_CUDA; }
    ctx->ncclRanks = n_ranks;
    return CPECAN_OK;
}

int cpecan_cuda_allreduce_expectations(cpecan_ctx *ctx) {
    if (!ctx) return CPECAN_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    if (!ctx->ncclComm) { ctx->err = "allreduce_expectations: call cpecan_cuda_nccl_init first"; return CPECAN_ERR_ARG; }
    if (ctx->running) { ctx->err = "allreduce_expectations: a run is in flight (wait for it)"; return CPECAN_ERR_ARG; }
    Nccl *L = ncclLib(ctx->err);
    if (!L) return CPECAN_ERR_CUDA;
    CK(ctx->dExpect.ensure(CPECAN_N_EXPECT * sizeof(double)));
    const size_t len = ctx->machine ? CPECAN_N_EXPECT_VANILLA : CPECAN_N_EXPECT;
    const int rc = L->allReduce(ctx->dExpect.p, ctx->dExpect.p, len, /* ncclDouble */ 8, /* ncclSum */ 0, ctx->ncclComm, ctx->stream);
    if (rc != 0) { ctx->err = std::string("ncclAllReduce: ") + (L->getErrorString ? L->getErrorString(rc) : "error"); return CPECAN_ERR_CUDA; }
    CK(waitStream(ctx, ctx->stream));
    return CPECAN_OK;
}

int cpecan_cuda_expectations_device_ptr(cpecan_ctx *ctx, double **dev_ptr_out) {
    if (!ctx || !dev_ptr_out) return CPECAN_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    CK(ctx->dExpect.ensure(CPECAN_N_EXPECT * sizeof(double)));
    *dev_ptr_out = ctx->dExpect.as<double>();
    return CPECAN_OK;
}

void *cpecan_cuda_host_alloc(cpecan_ctx *ctx, int64_t bytes) {
    if (!ctx || bytes <= 0) return nullptr;
    void *p = nullptr;
    cudaSetDevice(ctx->device);
    if (cudaHostAlloc(&p, (size_t) bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}

void cpecan_cuda_host_free(cpecan_ctx *ctx, void *p) {
    if (ctx && p) { cudaSetDevice(ctx->device); cudaFreeHost(p); }
}

int cpecan_cuda_set_resident_warps(cpecan_ctx *ctx, int32_t warps_per_sm) {
    if (!ctx) return CPECAN_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (warps_per_sm < 0) { ctx->err = "set_resident_warps: negative"; return CPECAN_ERR_ARG; }
    ctx->occCap = warps_per_sm;
    return CPECAN_OK;
}

int cpecan_cuda_set_exact_arithmetic(cpecan_ctx *ctx, int32_t on) {
    if (!ctx) return CPECAN_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->exact = on != 0;
    return CPECAN_OK;
}

int cpecan_cuda_get_timing(cpecan_ctx *ctx, cpecan_timing *out) {
    if (!ctx || !out) return CPECAN_ERR_ARG;
    *out = ctx->timing;
    return CPECAN_OK;
}

int cpecan_cuda_device_info(cpecan_ctx *ctx, int32_t *sm_count, int32_t *clock_khz, int64_t *hbm_bytes) {
    if (!ctx) return CPECAN_ERR_ARG;
    if (sm_count) *sm_count = ctx->prop.multiProcessorCount;
    if (clock_khz) *clock_khz = ctx->prop.clockRate;
    if (hbm_bytes) *hbm_bytes = (int64_t) ctx->prop.totalGlobalMem;
    return CPECAN_OK;
}

}  // extern "C"
