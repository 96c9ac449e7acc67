// pgi_api.cu — C-ABI (include/pgi.h) over the sm_100a kernels of pgi_kernels.cuh.
// Host-side plumbing only: device memory, one stream per context, pinned staging, CUDA-event timing.
// No CPU compute path exists here: without a usable device every entry point returns PGI_ERR_CUDA.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/pgi.h"
#include "pgi_kernels.cuh"
#include "pgi_astar.cuh"
#include "pgi_matcher.cuh"
#include "pgi_features.cuh"
#include "pgi_nvtx.h"

using namespace pgi;

namespace {

struct Registration {
    double *d_corr = nullptr;
    uint64_t *d_offset = nullptr;
    double *d_thr = nullptr;
    uint32_t *d_pairTable = nullptr;
    uint32_t *d_sampler = nullptr;
    uint16_t *d_itersTab = nullptr;
    uint64_t *d_itersOff = nullptr;
    uint64_t nPairs = 0, nRows = 0;
    uint32_t maxN = 0;
    std::vector<uint64_t> h_offset;
    bool borrowed = false;  // device buffers belong to another context (pgi_share_pairs)
    size_t capRows = 0, capOffset = 0, capThr = 0, capPairTable = 0, capSampler = 0, capIters = 0, capTables = 0;
};

constexpr int kMaxChunks = 64;
constexpr uint32_t kK5MaxSmemPts = 11264;  // 176 KB of float4 + 22 KB of queues; also keeps point slots per thread <= 64 (mask bits)

}  // namespace

struct pgi_ctx {
    pgi_config cfg;
    cudaStream_t stream = nullptr;
    std::string err;
    Registration reg, tmp;
    // wave buffers
    uint32_t waveCap = 0, hypCap = 0, bitsStride = 0;
    bool fbScratch = false;
    uint64_t maskCap = 0;
    uint32_t *d_pairId = nullptr, *d_hypOffset = nullptr, *d_bits = nullptr;
    double *d_hyp = nullptr;
    SlotState *d_state = nullptr;
    uint8_t *d_masks = nullptr, *d_fbCounts = nullptr;
    // staging of pgi_register_scene's compact inputs; kept between registrations (cudaMalloc / cudaFree of GB-sized
    // buffers costs up to hundreds of milliseconds and synchronises the device)
    double *d_stFocal = nullptr, *d_stSize = nullptr;
    uint64_t *d_stKpOff = nullptr;
    float2 *d_stKp = nullptr;
    uint2 *d_stPv = nullptr, *d_stM = nullptr;
    size_t capStFocal = 0, capStSize = 0, capStKpOff = 0, capStKp = 0, capStPv = 0, capStM = 0;
    uint32_t *d_dkList = nullptr, *d_dkCtl = nullptr;  // work list of the split K4 (K4a -> K4b -> K4c)
    uint32_t k4aSmem = 0;  // PGI_K4A_SMEM: dynamic shared memory per K4a CTA, an occupancy throttle (L2 footprint of the 5 KB stacks)
    uint32_t k4bCtas = 2;  // K4b CTAs per SM (PGI_K4B_CTAS): fewer lanes = more polynomials per lane = better refill balance
    bool k4Split = true;   // PGI_K4_SPLIT=0 selects the one-kernel k4_fallback_solve
    bool k1Tma = true;     // PGI_K1_TMA=0: always the direct-load K1
    uint32_t numSms = 148;
    uint32_t *d_k3Scratch = nullptr;  // per wave slot: K3 vote totals + arrival ticket (zero between launches)
    uint64_t *d_maskOffset = nullptr;
    pgi_verdict *d_verdicts = nullptr;
    double *d_fbSols = nullptr;
    float4 *d_fbSolsF = nullptr;
    unsigned long long *d_counters = nullptr;
    // pinned staging
    uint32_t *h_pairId = nullptr, *h_hypOffset = nullptr;
    double *h_hyp = nullptr;
    uint64_t *h_maskOffset = nullptr;
    pgi_verdict *h_verdicts = nullptr;
    unsigned long long *h_counters = nullptr;
    uint32_t hWaveCap = 0, hHypCap = 0;
    // in flight
    bool inFlight = false;
    uint32_t waveN = 0, waveFlags = 0;
    uint64_t waveMaskRows = 0;
    int nChunks = 0;
    cudaEvent_t evBegin = nullptr, evStart = nullptr, evK1 = nullptr, evK2 = nullptr, evK3 = nullptr, evFbEnd = nullptr;
    cudaEvent_t evChunk[kMaxChunks][2];
    bool fbLaunched = false;
    pgi_stats stats;
    // ---- K6: device mirror of the pose graph + per-search slabs (pgi_graph_*) ----
    AdjDev *d_adj = nullptr;
    uint32_t gV = 0, gCap = 0, gWords = 0;
    uint32_t *d_gCounts = nullptr;   // [0, V): committed entries per vertex, [V, 2V): committed + predicted
    double *d_simT = nullptr;
    HeapItemDev *d_heaps = nullptr;
    ArenaNodeDev *d_arenas = nullptr;
    uint32_t heapCap = 0, arenaCap = 0, searchSlots = 0;
    pgi_query *d_queries = nullptr;
    pgi_search_result *d_results = nullptr;
    uint32_t *d_expBits = nullptr, *d_nextQuery = nullptr;
    uint32_t queryCap = 0;
    ApplyEntry *d_apply = nullptr;
    uint32_t applyCap = 0;
    ApplyEntry *h_apply = nullptr;
    uint32_t *h_gCounts = nullptr;
    pgi_query *h_queries = nullptr;
    pgi_search_result *h_results = nullptr;
    uint32_t *h_expBits = nullptr;
    uint32_t hApplyCap = 0, hQueryCap = 0;
    int popLookahead = 1, stagedPush = 1;
    cudaEvent_t evS0 = nullptr, evS1 = nullptr;
    pgi_search_stats sstats;
};

namespace {

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            char buf__[256];                                                                       \
            snprintf(buf__, sizeof buf__, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                     __FILE__, __LINE__);                                                          \
            ctx->err = buf__;                                                                      \
            return e__ == cudaErrorMemoryAllocation ? PGI_ERR_NOMEM : PGI_ERR_CUDA;                \
        }                                                                                          \
    } while (0)

template <typename T>
pgi_status growDevice(pgi_ctx *ctx, T **p, size_t &cap, size_t need)
{
    if (need <= cap && *p) return PGI_OK;
    if (*p) CK(cudaFree(*p));
    *p = nullptr;
    cap = 0;
    CK(cudaMalloc((void **)p, std::max<size_t>(need, 1) * sizeof(T)));
    cap = need;
    return PGI_OK;
}

void freeReg(Registration &r)
{
    if (r.borrowed) { r = Registration(); return; }
    cudaFree(r.d_corr); cudaFree(r.d_offset); cudaFree(r.d_thr); cudaFree(r.d_pairTable);
    cudaFree(r.d_sampler); cudaFree(r.d_itersTab); cudaFree(r.d_itersOff);
    r = Registration();
}

// cv::RNG (multiply-with-carry) on the host, for the sampler tables.
struct HostRng {
    uint64_t state;
    explicit HostRng(uint64_t s) : state(s ? s : 0xffffffffULL) {}
    uint32_t next()
    {
        state = (uint64_t)(uint32_t)state * 4164903690U + (uint32_t)(state >> 32);
        return (uint32_t)state;
    }
    int uniform(int a, int b) { return a == b ? a : (int)(next() % (uint32_t)(b - a) + a); }
};

// Per distinct N: the uniform sampler's (iteration -> 5 indices) table (persistent-pool partial
// Fisher-Yates driven by cv::RNG(0)) and the termination table
// iterations(inlier count) = ceil(log(1-0.99)/log(1-(k/N)^5)) clamped to [1, maxIters].
pgi_status buildTables(pgi_ctx *ctx, Registration &r)
{
    const uint32_t maxIters = ctx->cfg.fallback_max_iters;
    std::map<uint32_t, uint32_t> tableOfN;
    std::vector<uint32_t> pairTable(r.nPairs);
    for (uint64_t p = 0; p < r.nPairs; p++) {
        const uint32_t N = (uint32_t)(r.h_offset[p + 1] - r.h_offset[p]);
        auto it = tableOfN.find(N);
        if (it == tableOfN.end()) it = tableOfN.emplace(N, (uint32_t)tableOfN.size()).first;
        pairTable[p] = it->second;
    }
    const size_t T = tableOfN.size();
    std::vector<uint32_t> sampler(T * maxIters * 5, 0);
    std::vector<uint64_t> itersOff(T, 0);
    std::vector<uint16_t> itersTab;
    std::vector<std::pair<uint32_t, uint32_t>> order(tableOfN.begin(), tableOfN.end());
    std::sort(order.begin(), order.end(), [](auto &a, auto &b) { return a.second < b.second; });
    std::vector<int> pool;
    for (auto &kv : order) {
        const int N = (int)kv.first;
        const uint32_t t = kv.second;
        itersOff[t] = itersTab.size();
        for (int k = 0; k <= N; k++) {
            int it;
            const double w = N > 0 ? (double)k / (double)N : 0.0;
            const double p5 = w * w * w * w * w;
            if (p5 <= 0.0)
                it = (int)maxIters;
            else if (p5 >= 1.0)
                it = 1;
            else {
                const double v = std::log(1.0 - 0.99) / std::log(1.0 - p5);
                it = v >= (double)maxIters ? (int)maxIters : (int)std::ceil(v);
                if (it < 1) it = 1;
            }
            itersTab.push_back((uint16_t)it);
        }
        if (N >= 5) {
            pool.resize(N);
            for (int i = 0; i < N; i++) pool[i] = i;
            HostRng rng(0);
            uint32_t *out = sampler.data() + (size_t)t * maxIters * 5;
            for (uint32_t itn = 0; itn < maxIters; itn++) {
                int size = N;
                for (int i = 0; i < 5; i++) {
                    const int j = rng.uniform(0, size);
                    out[(size_t)itn * 5 + i] = (uint32_t)pool[j];
                    std::swap(pool[j], pool[--size]);
                }
            }
        }
    }
    pgi_status st;
    if ((st = growDevice(ctx, &r.d_pairTable, r.capPairTable, (size_t)r.nPairs)) != PGI_OK) return st;
    if ((st = growDevice(ctx, &r.d_sampler, r.capSampler, sampler.size())) != PGI_OK) return st;
    if ((st = growDevice(ctx, &r.d_itersTab, r.capIters, itersTab.size())) != PGI_OK) return st;
    if ((st = growDevice(ctx, &r.d_itersOff, r.capTables, T)) != PGI_OK) return st;
    CK(cudaMemcpyAsync(r.d_pairTable, pairTable.data(), pairTable.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(r.d_sampler, sampler.data(), sampler.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(r.d_itersTab, itersTab.data(), itersTab.size() * 2, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(r.d_itersOff, itersOff.data(), itersOff.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->stats.h2d_bytes += pairTable.size() * 4 + sampler.size() * 4 + itersTab.size() * 2 + itersOff.size() * 8;
    return PGI_OK;
}

pgi_status registerDense(pgi_ctx *ctx, Registration &r, uint64_t nPairs, const uint64_t *offset, const double *corr,
                         const double *thr)
{
    if (!offset || (!corr && offset[nPairs] > 0) || !thr) { ctx->err = "null argument"; return PGI_ERR_INVALID; }
    for (uint64_t p = 0; p < nPairs; p++)
        if (offset[p + 1] < offset[p] || offset[p + 1] - offset[p] > 0x7fffffffULL) { ctx->err = "bad corr_offset"; return PGI_ERR_INVALID; }
    r.nPairs = nPairs;
    r.nRows = offset[nPairs];
    r.h_offset.assign(offset, offset + nPairs + 1);
    r.maxN = 0;
    for (uint64_t p = 0; p < nPairs; p++) r.maxN = std::max<uint32_t>(r.maxN, (uint32_t)(offset[p + 1] - offset[p]));
    pgi_status st;
    if ((st = growDevice(ctx, &r.d_corr, r.capRows, (size_t)r.nRows * 4)) != PGI_OK) return st;
    if ((st = growDevice(ctx, &r.d_offset, r.capOffset, (size_t)nPairs + 1)) != PGI_OK) return st;
    if ((st = growDevice(ctx, &r.d_thr, r.capThr, (size_t)nPairs)) != PGI_OK) return st;
    if (r.nRows) CK(cudaMemcpyAsync(r.d_corr, corr, r.nRows * 32, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(r.d_offset, offset, (nPairs + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (nPairs) CK(cudaMemcpyAsync(r.d_thr, thr, nPairs * 8, cudaMemcpyHostToDevice, ctx->stream));
    ctx->stats.h2d_bytes += r.nRows * 32 + (nPairs + 1) * 8 + nPairs * 8;
    return buildTables(ctx, r);
}

pgi_status ensureWave(pgi_ctx *ctx, const Registration &r, uint32_t n, uint32_t nHyp, uint32_t flags, uint64_t maskRows)
{
    const uint32_t stride = (r.maxN + 31) / 32 + 1;
    if (n > ctx->waveCap || stride > ctx->bitsStride) {
        const uint32_t cap = std::max<uint32_t>(std::max(n, ctx->waveCap), 64);
        cudaFree(ctx->d_pairId); cudaFree(ctx->d_hypOffset); cudaFree(ctx->d_state); cudaFree(ctx->d_bits);
        cudaFree(ctx->d_maskOffset); cudaFree(ctx->d_verdicts); cudaFree(ctx->d_fbSols); cudaFree(ctx->d_fbCounts); cudaFree(ctx->d_fbSolsF);
        cudaFree(ctx->d_k3Scratch); cudaFree(ctx->d_dkList);
        ctx->d_fbSolsF = nullptr; ctx->d_k3Scratch = nullptr; ctx->d_dkList = nullptr;
        ctx->d_pairId = ctx->d_hypOffset = ctx->d_bits = nullptr; ctx->d_state = nullptr; ctx->d_maskOffset = nullptr;
        ctx->d_verdicts = nullptr; ctx->d_fbSols = nullptr; ctx->d_fbCounts = nullptr;
        ctx->fbScratch = false;
        ctx->waveCap = 0;
        ctx->bitsStride = std::max(stride, ctx->bitsStride);
        CK(cudaMalloc((void **)&ctx->d_pairId, (size_t)cap * 4));
        CK(cudaMalloc((void **)&ctx->d_hypOffset, ((size_t)cap + 1) * 4));
        CK(cudaMalloc((void **)&ctx->d_state, (size_t)cap * sizeof(SlotState)));
        CK(cudaMalloc((void **)&ctx->d_bits, (size_t)cap * 2 * ctx->bitsStride * 4));
        CK(cudaMalloc((void **)&ctx->d_maskOffset, ((size_t)cap + 1) * 8));
        CK(cudaMalloc((void **)&ctx->d_verdicts, (size_t)cap * sizeof(pgi_verdict)));
        CK(cudaMalloc((void **)&ctx->d_k3Scratch, (size_t)cap * 8 * 4));
        CK(cudaMemsetAsync(ctx->d_k3Scratch, 0, (size_t)cap * 8 * 4, ctx->stream));
        ctx->waveCap = cap;
    }
    if ((flags & PGI_WAVE_FALLBACK) && !ctx->fbScratch) {
        // (after a failed attempt some of these may be allocated: release them first, cudaFree(nullptr) is a no-op)
        cudaFree(ctx->d_fbSols); cudaFree(ctx->d_fbCounts); cudaFree(ctx->d_fbSolsF); cudaFree(ctx->d_dkList);
        ctx->d_fbSols = nullptr; ctx->d_fbCounts = nullptr; ctx->d_fbSolsF = nullptr; ctx->d_dkList = nullptr;
        CK(cudaMalloc((void **)&ctx->d_fbSols, (size_t)ctx->waveCap * kFbChunk * 90 * 8));
        CK(cudaMalloc((void **)&ctx->d_fbCounts, (size_t)ctx->waveCap * kFbChunk));
        CK(cudaMalloc((void **)&ctx->d_fbSolsF, (size_t)ctx->waveCap * kFbChunk * 30 * sizeof(float4)));
        CK(cudaMalloc((void **)&ctx->d_dkList, (size_t)ctx->waveCap * kFbChunk * 4));
        if (!ctx->d_dkCtl) CK(cudaMalloc((void **)&ctx->d_dkCtl, 2 * 4));
        ctx->fbScratch = true;
    }
    if (nHyp > ctx->hypCap) {
        cudaFree(ctx->d_hyp);
        ctx->d_hyp = nullptr;
        ctx->hypCap = 0;  // (a failed allocation below must not leave a stale capacity behind a null pointer)
        const uint32_t cap = std::max<uint32_t>(nHyp, std::max<uint32_t>(ctx->waveCap, 64));
        CK(cudaMalloc((void **)&ctx->d_hyp, (size_t)cap * 7 * 8));
        ctx->hypCap = cap;
    }
    if ((flags & PGI_WAVE_MASKS) && maskRows > ctx->maskCap) {
        cudaFree(ctx->d_masks);
        ctx->d_masks = nullptr;
        ctx->maskCap = 0;
        CK(cudaMalloc((void **)&ctx->d_masks, (size_t)maskRows));
        ctx->maskCap = maskRows;
    }
    if (n > ctx->hWaveCap) {
        const uint32_t cap = std::max<uint32_t>(std::max(n, ctx->hWaveCap), 64);
        cudaFreeHost(ctx->h_pairId); cudaFreeHost(ctx->h_hypOffset); cudaFreeHost(ctx->h_maskOffset); cudaFreeHost(ctx->h_verdicts);
        ctx->h_pairId = ctx->h_hypOffset = nullptr; ctx->h_maskOffset = nullptr; ctx->h_verdicts = nullptr;
        ctx->hWaveCap = 0;
        CK(cudaMallocHost((void **)&ctx->h_pairId, (size_t)cap * 4));
        CK(cudaMallocHost((void **)&ctx->h_hypOffset, ((size_t)cap + 1) * 4));
        CK(cudaMallocHost((void **)&ctx->h_maskOffset, ((size_t)cap + 1) * 8));
        CK(cudaMallocHost((void **)&ctx->h_verdicts, (size_t)cap * sizeof(pgi_verdict)));
        ctx->hWaveCap = cap;
    }
    if (nHyp > ctx->hHypCap) {
        cudaFreeHost(ctx->h_hyp);
        ctx->h_hyp = nullptr;
        const uint32_t cap = std::max<uint32_t>(nHyp, std::max<uint32_t>(ctx->hWaveCap, 64));
        CK(cudaMallocHost((void **)&ctx->h_hyp, (size_t)cap * 7 * 8));
        ctx->hHypCap = cap;
    }
    return PGI_OK;
}

pgi_status submitWave(pgi_ctx *ctx, const Registration &r, uint32_t n, const uint32_t *pairId, const uint32_t *hypOffset,
                      const double *hyp, uint32_t flags, double thrOverride, uint32_t testMinOverride, bool scoreOnly)
{
    if (ctx->inFlight) { ctx->err = "a wave is already in flight"; return PGI_ERR_STATE; }
    if (n == 0) { ctx->waveN = 0; ctx->inFlight = true; ctx->fbLaunched = false; ctx->nChunks = 0; ctx->waveFlags = flags; return PGI_OK; }
    if (!pairId || !hypOffset) { ctx->err = "null argument"; return PGI_ERR_INVALID; }
    if (!r.d_offset) { ctx->err = "no pairs registered"; return PGI_ERR_STATE; }
    const uint32_t nHyp = hypOffset[n];
    if (nHyp && !hyp) { ctx->err = "hypotheses missing"; return PGI_ERR_INVALID; }
    uint64_t maskRows = 0;
    for (uint32_t i = 0; i < n; i++) {
        if (pairId[i] >= r.nPairs) { ctx->err = "pair_id out of range"; return PGI_ERR_INVALID; }
        if (hypOffset[i + 1] < hypOffset[i]) { ctx->err = "hyp_offset not monotone"; return PGI_ERR_INVALID; }
    }
    pgi_status st = ensureWave(ctx, r, n, nHyp, flags, 0);
    if (st != PGI_OK) return st;
    for (uint32_t i = 0; i < n; i++) {
        ctx->h_maskOffset[i] = maskRows;
        maskRows += r.h_offset[pairId[i] + 1] - r.h_offset[pairId[i]];
    }
    ctx->h_maskOffset[n] = maskRows;
    if ((st = ensureWave(ctx, r, n, nHyp, flags, maskRows)) != PGI_OK) return st;
    memcpy(ctx->h_pairId, pairId, (size_t)n * 4);
    memcpy(ctx->h_hypOffset, hypOffset, ((size_t)n + 1) * 4);
    if (nHyp) memcpy(ctx->h_hyp, hyp, (size_t)nHyp * 56);
    cudaStream_t s = ctx->stream;
    CK(cudaEventRecord(ctx->evBegin, s));
    CK(cudaMemcpyAsync(ctx->d_pairId, ctx->h_pairId, (size_t)n * 4, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->d_hypOffset, ctx->h_hypOffset, ((size_t)n + 1) * 4, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->d_maskOffset, ctx->h_maskOffset, ((size_t)n + 1) * 8, cudaMemcpyHostToDevice, s));
    if (nHyp) CK(cudaMemcpyAsync(ctx->d_hyp, ctx->h_hyp, (size_t)nHyp * 56, cudaMemcpyHostToDevice, s));
    CK(cudaMemsetAsync(ctx->d_counters, 0, 3 * sizeof(unsigned long long), s));
    ctx->stats.h2d_bytes += (size_t)n * 4 + ((size_t)n + 1) * 12 + (size_t)nHyp * 56;

    WaveArgs a;
    a.corr = r.d_corr; a.offset = r.d_offset; a.thr = r.d_thr; a.pairTable = r.d_pairTable;
    a.samplerTab = r.d_sampler; a.itersTab = r.d_itersTab; a.itersTabOff = r.d_itersOff;
    a.n = n; a.pairId = ctx->d_pairId; a.hypOffset = ctx->d_hypOffset; a.hyp = ctx->d_hyp; a.state = ctx->d_state;
    a.bits = ctx->d_bits; a.bitsStride = ctx->bitsStride; a.masks = ctx->d_masks; a.maskOffset = ctx->d_maskOffset;
    a.verdicts = ctx->d_verdicts; a.flags = flags; a.minInliers = ctx->cfg.min_inliers;
    a.testMinInliers = testMinOverride ? testMinOverride : ctx->cfg.test_min_inliers;
    a.fbMaxIters = ctx->cfg.fallback_max_iters; a.thrMultiplier = ctx->cfg.threshold_multiplier;
    a.thrOverride = thrOverride;
    a.fbSols = ctx->d_fbSols; a.fbSolsF = ctx->d_fbSolsF; a.fbCounts = ctx->d_fbCounts; a.counters = ctx->d_counters;
    a.k3Scratch = ctx->d_k3Scratch; a.dkList = ctx->d_dkList; a.dkCtl = ctx->d_dkCtl;

    CK(cudaEventRecord(ctx->evStart, s));  // kernel-only timing: the wave's small H2D copies are before this event
    // full waves stream the correspondences through shared memory with bulk async copies (persistent CTAs, one per SM);
    // small waves and waves that want byte masks take the one-CTA-per-pair kernel
    if (ctx->k1Tma && n >= 2u * ctx->numSms && !(flags & PGI_WAVE_MASKS)) {
        const int k1Smem = kK1Stages * kK1TileRows * 32 + (int)sizeof(K1Meta);
        CK(cudaFuncSetAttribute(k1_score_hypotheses_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, k1Smem));
        k1_score_hypotheses_tma<<<std::min<uint32_t>(n, ctx->numSms), kK1TmaThreads, k1Smem, s>>>(a);
    } else
        k1_score_hypotheses<<<n, kCtaThreads, 0, s>>>(a);
    CK(cudaEventRecord(ctx->evK1, s));
    ctx->stats.launches += 1;
    ctx->fbLaunched = false;
    ctx->nChunks = 0;
    if (!scoreOnly) {
        // the solve is one long serial chain per pair: on small waves give every pair a warp of its own (no divergence
        // serialisation between pairs, all SMs used); large waves pack 64 pairs per CTA for throughput
        if (n <= 4096u)
            k2_fivept_first_solution<<<n, 32, 0, s>>>(a, 1);
        else
            k2_fivept_first_solution<<<(n + 63) / 64, 64, 0, s>>>(a, 0);
        CK(cudaEventRecord(ctx->evK2, s));
        ctx->stats.launches += 1;
        if (flags & PGI_WAVE_FALLBACK) {
            const int chunks = (int)((ctx->cfg.fallback_max_iters + kFbChunk - 1) / kFbChunk);
            // FP32 staging area of K5: the largest pair if it fits, else 0 (those pairs take the FP64-only path)
            uint32_t smemPts = r.maxN <= kK5MaxSmemPts ? std::max<uint32_t>(r.maxN, 1) : kK5MaxSmemPts;
            // + per-warp compaction queues: 8 warps x min(smemPts, 2048) uint16
            const size_t k5Smem = (size_t)smemPts * 16 + (size_t)8 * (smemPts < kK5QueuePts ? ((smemPts + 31u) & ~31u) : kK5QueuePts) * 2 +
                                  (size_t)((smemPts + 31) / 32) * 4;  // + inlier bit mask of the LO refit
            if (k5Smem > 40 * 1024)  // static shared memory (~4 KB) counts towards the 48 KB default limit
                CK(cudaFuncSetAttribute(k5_fallback_score, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k5Smem));
            if (ctx->k4aSmem > 40 * 1024)
                CK(cudaFuncSetAttribute(k4a_polynomial, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->k4aSmem));
            ctx->nChunks = std::min(chunks, kMaxChunks);
            for (int c = 0; c < ctx->nChunks; c++) {
                const uint32_t threads = n * kFbChunk;
                if (ctx->k4Split) {
                    CK(cudaMemsetAsync(ctx->d_dkCtl, 0, 8, s));
                    k4a_polynomial<<<(threads + 127) / 128, 128, ctx->k4aSmem, s>>>(a, c);
                    k4b_roots<<<std::min<uint32_t>((threads + 127) / 128, ctx->numSms * ctx->k4bCtas), 128, 0, s>>>(a);
                    k4c_solutions<<<(threads + 127) / 128, 128, 0, s>>>(a, c);
                    ctx->stats.launches += 2;
                } else
                    k4_fallback_solve<<<(threads + 127) / 128, 128, 0, s>>>(a, c);
                CK(cudaEventRecord(ctx->evChunk[c][0], s));
                k5_fallback_score<<<n, kCtaThreads, k5Smem, s>>>(a, c, c == ctx->nChunks - 1 ? 1 : 0, smemPts);
                CK(cudaEventRecord(ctx->evChunk[c][1], s));
                ctx->stats.launches += 2;
            }
            ctx->fbLaunched = true;
        }
        CK(cudaEventRecord(ctx->evFbEnd, s));
        // small waves are latency bound: cover each pair by several point-range CTAs until the grid fills the GPU
        const uint32_t split = n >= 592u ? 1u : std::min(8u, (592u + n - 1) / n);
        k3_decompose_vote<<<n * split, kK3Threads, 0, s>>>(a, split);
        ctx->stats.launches += 1;
    } else {
        CK(cudaEventRecord(ctx->evK2, s));
        CK(cudaEventRecord(ctx->evFbEnd, s));
    }
    CK(cudaEventRecord(ctx->evK3, s));
    CK(cudaGetLastError());
    ctx->inFlight = true;
    ctx->waveN = n;
    ctx->waveFlags = flags;
    ctx->waveMaskRows = maskRows;
    ctx->stats.pairs += n;
    return PGI_OK;
}

// A failed CUDA call inside a wait must not leave the context "in flight" for ever (every later submit would answer
// PGI_ERR_STATE): whatever the exit path, the wave is over once the wait returns.
struct WaveEndGuard {
    pgi_ctx *ctx;
    ~WaveEndGuard()
    {
        if (ctx->inFlight) {
            cudaStreamSynchronize(ctx->stream);  // best effort: nothing of the failed wave may still be running
            ctx->inFlight = false;
        }
    }
};

pgi_status finishWave(pgi_ctx *ctx)
{
    // caller has synchronised the stream
    if (ctx->waveN == 0) { ctx->inFlight = false; return PGI_OK; }
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, ctx->evStart, ctx->evK1)); ctx->stats.ms_score += ms;
    CK(cudaEventElapsedTime(&ms, ctx->evK1, ctx->evK2)); ctx->stats.ms_fivept += ms;
    if (ctx->fbLaunched) {
        for (int c = 0; c < ctx->nChunks; c++) {
            CK(cudaEventElapsedTime(&ms, c == 0 ? ctx->evK2 : ctx->evChunk[c - 1][1], ctx->evChunk[c][0]));
            ctx->stats.ms_fallback_solve += ms;
            CK(cudaEventElapsedTime(&ms, ctx->evChunk[c][0], ctx->evChunk[c][1]));
            ctx->stats.ms_fallback_score += ms;
        }
    }
    CK(cudaEventElapsedTime(&ms, ctx->evFbEnd, ctx->evK3)); ctx->stats.ms_decompose += ms;
    CK(cudaEventElapsedTime(&ms, ctx->evBegin, ctx->evK3)); ctx->stats.ms_total += ms;
    ctx->stats.corr_evals += ctx->h_counters[0];
    ctx->stats.fallback_pairs += ctx->h_counters[1];
    ctx->stats.fallback_models += ctx->h_counters[2];
    ctx->inFlight = false;
    return PGI_OK;
}

}  // namespace

extern "C" {

static void freeGraph(pgi_ctx *ctx);

const char *pgi_version(void) { return "pgi 0.1 (sm_100a)"; }

int32_t pgi_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    int ok = 0;
    for (int d = 0; d < n; d++) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, d) == cudaSuccess && p.major == 10) ok++;
    }
    return ok;
}

pgi_status pgi_create(const pgi_config *cfg, pgi_ctx **out)
{
    if (!cfg || !out) return PGI_ERR_INVALID;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= cfg->device || cfg->device < 0) { cudaGetLastError(); return PGI_ERR_CUDA; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess) return PGI_ERR_CUDA;
    if (prop.major != 10) return PGI_ERR_CUDA;  // kernels are sm_100a-only; there is no other path
    pgi_ctx *ctx = new pgi_ctx();
    ctx->cfg = *cfg;
    if (const char *e = getenv("PGI_K4_SPLIT")) ctx->k4Split = atoi(e) != 0;
    if (const char *e = getenv("PGI_K1_TMA")) ctx->k1Tma = atoi(e) != 0;
    ctx->numSms = (uint32_t)std::max(1, prop.multiProcessorCount);
    if (const char *e = getenv("PGI_K4A_SMEM")) ctx->k4aSmem = (uint32_t)std::max(0, std::min(200 * 1024, atoi(e)));
    if (const char *e = getenv("PGI_K4B_CTAS")) ctx->k4bCtas = (uint32_t)std::max(1, std::min(4, atoi(e)));
    if (ctx->cfg.min_inliers == 0) ctx->cfg.min_inliers = 20;
    if (ctx->cfg.test_min_inliers == 0) ctx->cfg.test_min_inliers = 5;
    if (ctx->cfg.fallback_max_iters == 0) ctx->cfg.fallback_max_iters = 1000;
    if (ctx->cfg.fallback_max_iters > (uint32_t)kMaxChunks * kFbChunk) ctx->cfg.fallback_max_iters = kMaxChunks * kFbChunk;
    if (ctx->cfg.threshold_multiplier == 0.0) ctx->cfg.threshold_multiplier = 3.0 / 2.0;
    memset(&ctx->stats, 0, sizeof ctx->stats);
    memset(&ctx->sstats, 0, sizeof ctx->sstats);
    auto fail = [&](pgi_status s) { pgi_destroy(ctx); return s; };
    if (cudaSetDevice(cfg->device) != cudaSuccess) return fail(PGI_ERR_CUDA);
    {
        // flags bit 0: background context (lowest stream priority) — used for the hypothesis-independent fallback
        // prefetch so that the latency-critical wave kernels of the foreground context get SM slots first
        int prioLow = 0, prioHigh = 0;
        cudaDeviceGetStreamPriorityRange(&prioLow, &prioHigh);
        const int prio = (ctx->cfg.flags & 1u) ? prioLow : prioHigh;
        if (cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, prio) != cudaSuccess) return fail(PGI_ERR_CUDA);
    }
    cudaEvent_t *evs[] = {&ctx->evBegin, &ctx->evStart, &ctx->evK1, &ctx->evK2, &ctx->evK3, &ctx->evFbEnd};
    for (auto e : evs)
        if (cudaEventCreate(e) != cudaSuccess) return fail(PGI_ERR_CUDA);
    for (int c = 0; c < kMaxChunks; c++)
        for (int k = 0; k < 2; k++) {
            ctx->evChunk[c][k] = nullptr;
            if (cudaEventCreate(&ctx->evChunk[c][k]) != cudaSuccess) return fail(PGI_ERR_CUDA);
        }
    if (cudaMalloc((void **)&ctx->d_counters, 3 * sizeof(unsigned long long)) != cudaSuccess) return fail(PGI_ERR_NOMEM);
    if (cudaMallocHost((void **)&ctx->h_counters, 3 * sizeof(unsigned long long)) != cudaSuccess) return fail(PGI_ERR_NOMEM);
    memset(ctx->h_counters, 0, 3 * sizeof(unsigned long long));
    *out = ctx;
    return PGI_OK;
}

pgi_status pgi_destroy(pgi_ctx *ctx)
{
    if (!ctx) return PGI_OK;
    cudaSetDevice(ctx->cfg.device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    freeGraph(ctx);
    if (ctx->evS0) cudaEventDestroy(ctx->evS0);
    if (ctx->evS1) cudaEventDestroy(ctx->evS1);
    freeReg(ctx->reg);
    freeReg(ctx->tmp);
    cudaFree(ctx->d_pairId); cudaFree(ctx->d_hypOffset); cudaFree(ctx->d_bits); cudaFree(ctx->d_hyp);
    cudaFree(ctx->d_state); cudaFree(ctx->d_masks); cudaFree(ctx->d_fbCounts); cudaFree(ctx->d_maskOffset);
    cudaFree(ctx->d_verdicts); cudaFree(ctx->d_fbSols); cudaFree(ctx->d_fbSolsF); cudaFree(ctx->d_counters);
    cudaFree(ctx->d_k3Scratch); cudaFree(ctx->d_dkList); cudaFree(ctx->d_dkCtl);
    cudaFree(ctx->d_stFocal); cudaFree(ctx->d_stSize); cudaFree(ctx->d_stKpOff); cudaFree(ctx->d_stKp); cudaFree(ctx->d_stPv); cudaFree(ctx->d_stM);
    cudaFreeHost(ctx->h_pairId); cudaFreeHost(ctx->h_hypOffset); cudaFreeHost(ctx->h_hyp);
    cudaFreeHost(ctx->h_maskOffset); cudaFreeHost(ctx->h_verdicts); cudaFreeHost(ctx->h_counters);
    cudaEvent_t evs[] = {ctx->evBegin, ctx->evStart, ctx->evK1, ctx->evK2, ctx->evK3, ctx->evFbEnd};
    for (auto e : evs)
        if (e) cudaEventDestroy(e);
    for (int c = 0; c < kMaxChunks; c++)
        for (int k = 0; k < 2; k++)
            if (ctx->evChunk[c][k]) cudaEventDestroy(ctx->evChunk[c][k]);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    cudaGetLastError();
    delete ctx;
    return PGI_OK;
}

const char *pgi_last_error(pgi_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

pgi_status pgi_register_pairs(pgi_ctx *ctx, uint64_t n_pairs, const uint64_t *corr_offset, const double *corr_xy4,
                              const double *thr_norm)
{
    PgiNvtxRange nvtxRange("pgi:register_pairs");
    if (!ctx) return PGI_ERR_INVALID;
    if (ctx->inFlight) { ctx->err = "a wave is in flight"; return PGI_ERR_STATE; }
    CK(cudaSetDevice(ctx->cfg.device));
    if (ctx->reg.borrowed) ctx->reg = Registration();
    return registerDense(ctx, ctx->reg, n_pairs, corr_offset, corr_xy4, thr_norm);
}

pgi_status pgi_register_scene(pgi_ctx *ctx, uint64_t n_views, const double *focal, const double *size_wh,
                              const uint64_t *kp_offset, const float *kp_xy, uint64_t n_pairs,
                              const uint32_t *pair_views, const uint64_t *m_offset, const uint32_t *matches,
                              double thr_px)
{
    PgiNvtxRange nvtxRange("pgi:register_scene (H2D + K0)");
    if (!ctx) return PGI_ERR_INVALID;
    if (ctx->inFlight) { ctx->err = "a wave is in flight"; return PGI_ERR_STATE; }
    if (!focal || !size_wh || !kp_offset || !kp_xy || !pair_views || !m_offset || (!matches && m_offset[n_pairs])) {
        ctx->err = "null argument";
        return PGI_ERR_INVALID;
    }
    CK(cudaSetDevice(ctx->cfg.device));
    if (ctx->reg.borrowed) ctx->reg = Registration();
    Registration &r = ctx->reg;
    const uint64_t nKp = kp_offset[n_views], nRows = m_offset[n_pairs];
    for (uint64_t p = 0; p < n_pairs; p++) {
        if (pair_views[2 * p] >= n_views || pair_views[2 * p + 1] >= n_views) { ctx->err = "view id out of range"; return PGI_ERR_INVALID; }
        if (m_offset[p + 1] < m_offset[p] || m_offset[p + 1] - m_offset[p] > 0x7fffffffULL) { ctx->err = "bad m_offset"; return PGI_ERR_INVALID; }
    }
    r.nPairs = n_pairs;
    r.nRows = nRows;
    r.h_offset.assign(m_offset, m_offset + n_pairs + 1);
    r.maxN = 0;
    for (uint64_t p = 0; p < n_pairs; p++) r.maxN = std::max<uint32_t>(r.maxN, (uint32_t)(m_offset[p + 1] - m_offset[p]));
    pgi_status st;
    if ((st = growDevice(ctx, &r.d_corr, r.capRows, (size_t)nRows * 4)) != PGI_OK) return st;
    if ((st = growDevice(ctx, &r.d_offset, r.capOffset, (size_t)n_pairs + 1)) != PGI_OK) return st;
    if ((st = growDevice(ctx, &r.d_thr, r.capThr, (size_t)n_pairs)) != PGI_OK) return st;
    // staging buffers for the compact inputs (owned by the context, grown on demand)
    if ((st = growDevice(ctx, &ctx->d_stFocal, ctx->capStFocal, (size_t)n_views)) != PGI_OK) return st;
    if ((st = growDevice(ctx, &ctx->d_stSize, ctx->capStSize, (size_t)n_views * 2)) != PGI_OK) return st;
    if ((st = growDevice(ctx, &ctx->d_stKpOff, ctx->capStKpOff, (size_t)n_views + 1)) != PGI_OK) return st;
    if ((st = growDevice(ctx, &ctx->d_stKp, ctx->capStKp, (size_t)nKp)) != PGI_OK) return st;
    if ((st = growDevice(ctx, &ctx->d_stPv, ctx->capStPv, (size_t)n_pairs)) != PGI_OK) return st;
    if ((st = growDevice(ctx, &ctx->d_stM, ctx->capStM, (size_t)nRows)) != PGI_OK) return st;
    double *d_focal = ctx->d_stFocal, *d_size = ctx->d_stSize;
    uint64_t *d_kpOff = ctx->d_stKpOff;
    float2 *d_kp = ctx->d_stKp;
    uint2 *d_pv = ctx->d_stPv, *d_m = ctx->d_stM;
    cudaStream_t s = ctx->stream;
    auto cleanup = [&]() {};
#define CKC(call) do { cudaError_t e2__ = (call); if (e2__ != cudaSuccess) { cleanup(); ctx->err = std::string(#call) + ": " + cudaGetErrorString(e2__); return e2__ == cudaErrorMemoryAllocation ? PGI_ERR_NOMEM : PGI_ERR_CUDA; } } while (0)
    CKC(cudaMemcpyAsync(d_focal, focal, n_views * 8, cudaMemcpyHostToDevice, s));
    CKC(cudaMemcpyAsync(d_size, size_wh, n_views * 16, cudaMemcpyHostToDevice, s));
    CKC(cudaMemcpyAsync(d_kpOff, kp_offset, (n_views + 1) * 8, cudaMemcpyHostToDevice, s));
    CKC(cudaMemcpyAsync(d_kp, kp_xy, nKp * 8, cudaMemcpyHostToDevice, s));
    CKC(cudaMemcpyAsync(d_pv, pair_views, n_pairs * 8, cudaMemcpyHostToDevice, s));
    CKC(cudaMemcpyAsync(d_m, matches, nRows * 8, cudaMemcpyHostToDevice, s));
    CKC(cudaMemcpyAsync(r.d_offset, m_offset, (n_pairs + 1) * 8, cudaMemcpyHostToDevice, s));
    ctx->stats.h2d_bytes += n_views * 24 + (n_views + 1) * 8 + nKp * 8 + n_pairs * 8 + nRows * 8 + (n_pairs + 1) * 8;
    if (n_pairs) {
        CKC(cudaEventRecord(ctx->evStart, s));
        const uint32_t gx = std::max<uint32_t>(1, std::min<uint32_t>((r.maxN + 255) / 256, 64));
        const uint32_t gy = (uint32_t)std::min<uint64_t>(n_pairs, 65535), gz = (uint32_t)((n_pairs + gy - 1) / gy);
        k0_build_correspondences<<<dim3(gx, gy, gz), 256, 0, s>>>(d_focal, d_size, d_kpOff, d_kp, n_pairs, d_pv, r.d_offset,
                                                                   d_m, thr_px, reinterpret_cast<double4 *>(r.d_corr), r.d_thr);
        CKC(cudaEventRecord(ctx->evK1, s));
        CKC(cudaGetLastError());
        ctx->stats.launches += 1;
    }
    CKC(cudaStreamSynchronize(s));
    if (n_pairs) {
        float ms = 0;
        CKC(cudaEventElapsedTime(&ms, ctx->evStart, ctx->evK1));
        ctx->stats.ms_correspondences += ms;
    }
    cleanup();
#undef CKC
    return buildTables(ctx, r);
}

pgi_status pgi_share_pairs(pgi_ctx *ctx, pgi_ctx *owner)
{
    if (!ctx || !owner || ctx == owner) return PGI_ERR_INVALID;
    if (ctx->inFlight) { ctx->err = "a wave is in flight"; return PGI_ERR_STATE; }
    if (ctx->cfg.device != owner->cfg.device) { ctx->err = "contexts are on different devices"; return PGI_ERR_INVALID; }
    if (ctx->cfg.fallback_max_iters != owner->cfg.fallback_max_iters) { ctx->err = "fallback_max_iters differ"; return PGI_ERR_INVALID; }
    freeReg(ctx->reg);
    ctx->reg = owner->reg;  // shallow copy of the device pointers and host offsets
    ctx->reg.borrowed = true;
    return PGI_OK;
}

pgi_status pgi_read_pair(pgi_ctx *ctx, uint32_t pair_id, double *corr_xy4, uint64_t capacity_rows, uint64_t *n_rows,
                         double *thr_norm)
{
    if (!ctx) return PGI_ERR_INVALID;
    const Registration &r = ctx->reg;
    if (pair_id >= r.nPairs) { ctx->err = "pair_id out of range"; return PGI_ERR_INVALID; }
    CK(cudaSetDevice(ctx->cfg.device));
    const uint64_t n = r.h_offset[pair_id + 1] - r.h_offset[pair_id];
    if (n_rows) *n_rows = n;
    if (corr_xy4) {
        if (capacity_rows < n) { ctx->err = "capacity too small"; return PGI_ERR_INVALID; }
        CK(cudaMemcpyAsync(corr_xy4, r.d_corr + 4 * r.h_offset[pair_id], n * 32, cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (thr_norm) CK(cudaMemcpyAsync(thr_norm, r.d_thr + pair_id, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return PGI_OK;
}

pgi_status pgi_submit_wave(pgi_ctx *ctx, uint32_t n, const uint32_t *pair_id, const uint32_t *hyp_offset,
                           const double *hyp_q_t, uint32_t flags)
{
    PgiNvtxRange nvtxRange("pgi:submit_wave (H2D hypotheses + K1/K2/K4/K5/K3 launches)");
    if (!ctx) return PGI_ERR_INVALID;
    CK(cudaSetDevice(ctx->cfg.device));
    return submitWave(ctx, ctx->reg, n, pair_id, hyp_offset, hyp_q_t, flags, 0.0, 0, false);
}

pgi_status pgi_wait_wave(pgi_ctx *ctx, pgi_verdict *out, uint8_t *masks_or_null)
{
    PgiNvtxRange nvtxRange("pgi:wait_wave (sync + D2H verdicts)");
    if (!ctx) return PGI_ERR_INVALID;
    if (!ctx->inFlight) { ctx->err = "no wave in flight"; return PGI_ERR_STATE; }
    WaveEndGuard guard{ctx};
    CK(cudaSetDevice(ctx->cfg.device));
    const uint32_t n = ctx->waveN;
    if (n) {
        if (!out) { ctx->err = "null output"; return PGI_ERR_INVALID; }
        CK(cudaMemcpyAsync(ctx->h_verdicts, ctx->d_verdicts, (size_t)n * sizeof(pgi_verdict), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(ctx->h_counters, ctx->d_counters, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
        if (masks_or_null && (ctx->waveFlags & PGI_WAVE_MASKS) && ctx->waveMaskRows)
            CK(cudaMemcpyAsync(masks_or_null, ctx->d_masks, ctx->waveMaskRows, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        memcpy(out, ctx->h_verdicts, (size_t)n * sizeof(pgi_verdict));
        ctx->stats.d2h_bytes += (size_t)n * sizeof(pgi_verdict) + 24 + ((masks_or_null && (ctx->waveFlags & PGI_WAVE_MASKS)) ? ctx->waveMaskRows : 0);
    }
    return finishWave(ctx);
}

pgi_status pgi_wait_wave_device(pgi_ctx *ctx, void *verdicts_device)
{
    PgiNvtxRange nvtxRange("pgi:wait_wave_device");
    if (!ctx) return PGI_ERR_INVALID;
    if (!ctx->inFlight) { ctx->err = "no wave in flight"; return PGI_ERR_STATE; }
    WaveEndGuard guard{ctx};
    CK(cudaSetDevice(ctx->cfg.device));
    const uint32_t n = ctx->waveN;
    if (n) {
        if (!verdicts_device) { ctx->err = "null output"; return PGI_ERR_INVALID; }
        CK(cudaMemcpyAsync(verdicts_device, ctx->d_verdicts, (size_t)n * sizeof(pgi_verdict), cudaMemcpyDeviceToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->h_counters, ctx->d_counters, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        ctx->stats.d2h_bytes += 24;
    }
    return finishWave(ctx);
}

int32_t pgi_estimate_pose(pgi_ctx *ctx, const double *corr_xy4, uint64_t n, double thr_norm, const double *guesses_q_t,
                          uint32_t n_guesses, double *pose_q_t_out, uint8_t *mask_out, uint64_t *inlier_number_out,
                          pgi_verdict *verdict_or_null)
{
    if (!ctx) return PGI_ERR_INVALID;
    if (ctx->inFlight) { ctx->err = "a wave is in flight"; return PGI_ERR_STATE; }
    if (n > 0x7fffffffULL) { ctx->err = "too many correspondences"; return PGI_ERR_INVALID; }
    if (cudaSetDevice(ctx->cfg.device) != cudaSuccess) return PGI_ERR_CUDA;
    const uint64_t off[2] = {0, n};
    pgi_status st = registerDense(ctx, ctx->tmp, 1, off, corr_xy4, &thr_norm);
    if (st != PGI_OK) return st;
    const uint32_t pid = 0, hoff[2] = {0, n_guesses};
    const uint32_t flags = PGI_WAVE_PATH | PGI_WAVE_FALLBACK | PGI_WAVE_NO_TEST | (mask_out ? PGI_WAVE_MASKS : 0u);
    st = submitWave(ctx, ctx->tmp, 1, &pid, hoff, guesses_q_t, flags, 0.0, 0, false);
    if (st != PGI_OK) return st;
    pgi_verdict v;
    st = pgi_wait_wave(ctx, &v, mask_out);
    if (st != PGI_OK) return st;
    if (pose_q_t_out) {
        for (int k = 0; k < 4; k++) pose_q_t_out[k] = v.q[k];
        for (int k = 0; k < 3; k++) pose_q_t_out[4 + k] = v.t[k];
    }
    if (inlier_number_out) *inlier_number_out = v.inlier_count;
    if (verdict_or_null) *verdict_or_null = v;
    return v.accepted ? 1 : 0;
}

int32_t pgi_test_pose(pgi_ctx *ctx, const double *corr_xy4, uint64_t n, double thr, uint64_t min_inliers,
                      const double *pose_q_t, uint64_t *inlier_number_out)
{
    if (!ctx) return PGI_ERR_INVALID;
    if (ctx->inFlight) { ctx->err = "a wave is in flight"; return PGI_ERR_STATE; }
    if (!pose_q_t || n > 0x7fffffffULL || min_inliers == 0 || min_inliers > 0xffffffffULL) { ctx->err = "bad argument"; return PGI_ERR_INVALID; }
    if (cudaSetDevice(ctx->cfg.device) != cudaSuccess) return PGI_ERR_CUDA;
    const uint64_t off[2] = {0, n};
    const double thrReg = thr;
    pgi_status st = registerDense(ctx, ctx->tmp, 1, off, corr_xy4, &thrReg);
    if (st != PGI_OK) return st;
    const uint32_t pid = 0, hoff[2] = {0, 1};
    st = submitWave(ctx, ctx->tmp, 1, &pid, hoff, pose_q_t, PGI_WAVE_PATH, thr, (uint32_t)min_inliers, true);
    if (st != PGI_OK) return st;
    // score-only wave: read the slot state instead of a verdict
    SlotState s;
    CK(cudaMemcpyAsync(&s, ctx->d_state, sizeof s, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(ctx->h_counters, ctx->d_counters, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    st = finishWave(ctx);
    if (st != PGI_OK) return st;
    if (inlier_number_out) *inlier_number_out = s.testCount;
    return (s.flags & ST_TEST_PASSED) ? 1 : 0;
}

// ---- K6: device pose graph + batched A* -------------------------------------------------------------------
static void freeGraph(pgi_ctx *ctx)
{
    cudaFree(ctx->d_adj); cudaFree(ctx->d_gCounts); cudaFree(ctx->d_simT); cudaFree(ctx->d_heaps); cudaFree(ctx->d_arenas);
    cudaFree(ctx->d_queries); cudaFree(ctx->d_results); cudaFree(ctx->d_expBits); cudaFree(ctx->d_nextQuery); cudaFree(ctx->d_apply);
    cudaFreeHost(ctx->h_apply); cudaFreeHost(ctx->h_gCounts); cudaFreeHost(ctx->h_queries); cudaFreeHost(ctx->h_results);
    cudaFreeHost(ctx->h_expBits);
    ctx->d_adj = nullptr; ctx->d_gCounts = nullptr; ctx->d_simT = nullptr; ctx->d_heaps = nullptr; ctx->d_arenas = nullptr;
    ctx->d_queries = nullptr; ctx->d_results = nullptr; ctx->d_expBits = nullptr; ctx->d_nextQuery = nullptr; ctx->d_apply = nullptr;
    ctx->h_apply = nullptr; ctx->h_gCounts = nullptr; ctx->h_queries = nullptr; ctx->h_results = nullptr; ctx->h_expBits = nullptr;
    ctx->gV = ctx->gCap = ctx->gWords = 0;
    ctx->queryCap = ctx->applyCap = ctx->hApplyCap = ctx->hQueryCap = 0;
    ctx->heapCap = ctx->arenaCap = ctx->searchSlots = 0;
}

pgi_status pgi_graph_init(pgi_ctx *ctx, uint32_t n_views, const double *sim_to_next)
{
    if (!ctx) return PGI_ERR_INVALID;
    if (!sim_to_next || n_views == 0 || n_views > 65535u) { ctx->err = "bad graph size (1..65535 views)"; return PGI_ERR_INVALID; }
    CK(cudaSetDevice(ctx->cfg.device));
    CK(cudaStreamSynchronize(ctx->stream));
    const uint32_t V = n_views;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, ctx->cfg.device));
    uint32_t slotsPerSm = 16, heapCap = 0;
    if (const char *e = getenv("PGI_ASTAR_SLOTS_PER_SM")) slotsPerSm = (uint32_t)std::max(4, std::min(48, atoi(e)));
    slotsPerSm = (slotsPerSm / kAstarWarps) * kAstarWarps;
    if (const char *e = getenv("PGI_ASTAR_HEAP")) heapCap = (uint32_t)std::max(1024, atoi(e));
    if (const char *e = getenv("PGI_ASTAR_POP")) ctx->popLookahead = atoi(e) != 0;
    if (const char *e = getenv("PGI_ASTAR_STAGED")) ctx->stagedPush = atoi(e) != 0;
    if (!heapCap) {
        const uint64_t want = (uint64_t)V * V / 2;
        heapCap = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(want, 65536), 1u << 20);
    }
    const uint32_t slots = (uint32_t)prop.multiProcessorCount * slotsPerSm;
    const uint32_t arenaCap = 16384;
    if (ctx->gV != V || ctx->heapCap != heapCap || ctx->searchSlots != slots) {
        freeGraph(ctx);
        ctx->gV = V; ctx->gCap = V; ctx->gWords = (V + 31) / 32;
        CK(cudaMalloc((void **)&ctx->d_adj, (size_t)V * ctx->gCap * sizeof(AdjDev)));
        CK(cudaMalloc((void **)&ctx->d_gCounts, (size_t)2 * V * 4));
        CK(cudaMalloc((void **)&ctx->d_simT, (size_t)V * V * 8));
        CK(cudaMalloc((void **)&ctx->d_heaps, (size_t)slots * heapCap * sizeof(HeapItemDev)));
        CK(cudaMalloc((void **)&ctx->d_arenas, (size_t)slots * arenaCap * sizeof(ArenaNodeDev)));
        CK(cudaMalloc((void **)&ctx->d_nextQuery, 4));
        CK(cudaMallocHost((void **)&ctx->h_gCounts, (size_t)2 * V * 4));
        ctx->heapCap = heapCap; ctx->arenaCap = arenaCap; ctx->searchSlots = slots;
    }
    CK(cudaMemsetAsync(ctx->d_gCounts, 0, (size_t)2 * V * 4, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_simT, sim_to_next, (size_t)V * V * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->sstats.h2d_bytes += (uint64_t)V * V * 8;
    if (!ctx->evS0) { CK(cudaEventCreate(&ctx->evS0)); CK(cudaEventCreate(&ctx->evS1)); }
    return PGI_OK;
}

pgi_status pgi_graph_apply(pgi_ctx *ctx, uint32_t n_entries, const pgi_adj_entry *entries, const uint32_t *committed_count,
                           const uint32_t *total_count)
{
    if (!ctx) return PGI_ERR_INVALID;
    if (!ctx->d_adj) { ctx->err = "pgi_graph_init has not been called"; return PGI_ERR_STATE; }
    if ((n_entries && !entries) || !committed_count || !total_count) { ctx->err = "null argument"; return PGI_ERR_INVALID; }
    CK(cudaSetDevice(ctx->cfg.device));
    const uint32_t V = ctx->gV;
    for (uint32_t v = 0; v < V; v++)
        if (total_count[v] < committed_count[v] || total_count[v] > ctx->gCap) { ctx->err = "bad edge-list count"; return PGI_ERR_INVALID; }
    for (uint32_t i = 0; i < n_entries; i++)
        if (entries[i].vertex >= V || entries[i].next >= V || entries[i].index >= ctx->gCap) { ctx->err = "edge-list entry out of range"; return PGI_ERR_INVALID; }
    if (n_entries > ctx->hApplyCap) {
        const uint32_t cap = std::max<uint32_t>(n_entries * 2, 4096);
        cudaFreeHost(ctx->h_apply); ctx->h_apply = nullptr; ctx->hApplyCap = 0;
        cudaFree(ctx->d_apply); ctx->d_apply = nullptr; ctx->applyCap = 0;
        CK(cudaMallocHost((void **)&ctx->h_apply, (size_t)cap * sizeof(ApplyEntry)));
        CK(cudaMalloc((void **)&ctx->d_apply, (size_t)cap * sizeof(ApplyEntry)));
        ctx->hApplyCap = ctx->applyCap = cap;
    }
    static_assert(sizeof(ApplyEntry) == sizeof(pgi_adj_entry), "apply entry layout");
    cudaStream_t s = ctx->stream;
    memcpy(ctx->h_gCounts, committed_count, (size_t)V * 4);
    memcpy(ctx->h_gCounts + V, total_count, (size_t)V * 4);
    CK(cudaMemcpyAsync(ctx->d_gCounts, ctx->h_gCounts, (size_t)2 * V * 4, cudaMemcpyHostToDevice, s));
    if (n_entries) {
        memcpy(ctx->h_apply, entries, (size_t)n_entries * sizeof(ApplyEntry));
        CK(cudaMemcpyAsync(ctx->d_apply, ctx->h_apply, (size_t)n_entries * sizeof(ApplyEntry), cudaMemcpyHostToDevice, s));
        k6_graph_apply<<<(n_entries + 255) / 256, 256, 0, s>>>(ctx->d_adj, ctx->gCap, ctx->d_apply, n_entries);
        CK(cudaGetLastError());
        ctx->sstats.launches += 1;
    }
    CK(cudaStreamSynchronize(s));  // the pinned staging buffers are reused by the next call
    ctx->sstats.h2d_bytes += (uint64_t)2 * V * 4 + (uint64_t)n_entries * sizeof(ApplyEntry);
    return PGI_OK;
}

pgi_status pgi_graph_search(pgi_ctx *ctx, uint32_t n, const pgi_query *queries, uint32_t max_depth, double weight,
                            pgi_search_result *results, uint32_t *expanded_bits)
{
    PgiNvtxRange nvtxRange("pgi:graph_search (K6)");
    if (!ctx) return PGI_ERR_INVALID;
    if (!ctx->d_adj) { ctx->err = "pgi_graph_init has not been called"; return PGI_ERR_STATE; }
    if (n == 0) return PGI_OK;
    if (!queries || !results || !expanded_bits || max_depth > 7) { ctx->err = "bad argument"; return PGI_ERR_INVALID; }
    CK(cudaSetDevice(ctx->cfg.device));
    const uint32_t V = ctx->gV, words = ctx->gWords;
    for (uint32_t i = 0; i < n; i++)
        if (queries[i].src >= V || queries[i].dst >= V) { ctx->err = "query vertex out of range"; return PGI_ERR_INVALID; }
    if (n > ctx->queryCap) {
        const uint32_t cap = std::max<uint32_t>(n, 4096);
        cudaFree(ctx->d_queries); cudaFree(ctx->d_results); cudaFree(ctx->d_expBits);
        cudaFreeHost(ctx->h_queries); cudaFreeHost(ctx->h_results); cudaFreeHost(ctx->h_expBits);
        ctx->d_queries = nullptr; ctx->d_results = nullptr; ctx->d_expBits = nullptr;
        ctx->h_queries = nullptr; ctx->h_results = nullptr; ctx->h_expBits = nullptr;
        ctx->queryCap = ctx->hQueryCap = 0;
        CK(cudaMalloc((void **)&ctx->d_queries, (size_t)cap * sizeof(pgi_query)));
        CK(cudaMalloc((void **)&ctx->d_results, (size_t)cap * sizeof(pgi_search_result)));
        CK(cudaMalloc((void **)&ctx->d_expBits, (size_t)cap * words * 4));
        CK(cudaMallocHost((void **)&ctx->h_queries, (size_t)cap * sizeof(pgi_query)));
        CK(cudaMallocHost((void **)&ctx->h_results, (size_t)cap * sizeof(pgi_search_result)));
        CK(cudaMallocHost((void **)&ctx->h_expBits, (size_t)cap * words * 4));
        ctx->queryCap = ctx->hQueryCap = cap;
    }
    cudaStream_t s = ctx->stream;
    memcpy(ctx->h_queries, queries, (size_t)n * sizeof(pgi_query));
    CK(cudaMemcpyAsync(ctx->d_queries, ctx->h_queries, (size_t)n * sizeof(pgi_query), cudaMemcpyHostToDevice, s));
    CK(cudaMemsetAsync(ctx->d_nextQuery, 0, 4, s));
    SearchArgs a;
    a.adj = ctx->d_adj; a.cap = ctx->gCap; a.ccnt = ctx->d_gCounts; a.cnt = ctx->d_gCounts + V; a.simT = ctx->d_simT;
    a.V = V; a.words = words; a.n = n; a.queries = ctx->d_queries; a.maxDepth = max_depth;
    a.weight = weight; a.oneMinusWeight = 1.0 - weight;
    a.heaps = ctx->d_heaps; a.heapCap = ctx->heapCap; a.arenas = ctx->d_arenas; a.arenaCap = ctx->arenaCap;
    a.results = ctx->d_results; a.expandedBits = ctx->d_expBits; a.nextQuery = ctx->d_nextQuery;
    a.popLookahead = ctx->popLookahead;
    a.staged = ctx->stagedPush;
    const uint32_t ctas = std::min<uint32_t>((n + kAstarWarps - 1) / kAstarWarps, ctx->searchSlots / kAstarWarps);
    const size_t smem = (size_t)kAstarWarps * ((size_t)kStageItems * sizeof(HeapItemDev) + (size_t)kChunk * 2 + (size_t)2 * words * 4);
    CK(cudaFuncSetAttribute(k6_astar_search, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 48 * 1024)));
    CK(cudaEventRecord(ctx->evS0, s));
    k6_astar_search<<<ctas, kAstarWarps * 32, smem, s>>>(a);
    CK(cudaEventRecord(ctx->evS1, s));
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(ctx->h_results, ctx->d_results, (size_t)n * sizeof(pgi_search_result), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(ctx->h_expBits, ctx->d_expBits, (size_t)n * words * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    memcpy(results, ctx->h_results, (size_t)n * sizeof(pgi_search_result));
    memcpy(expanded_bits, ctx->h_expBits, (size_t)n * words * 4);
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, ctx->evS0, ctx->evS1));
    pgi_search_stats &st = ctx->sstats;
    st.ms_search += ms;
    st.launches += 1;
    st.queries += n;
    uint32_t longest = 0;
    for (uint32_t i = 0; i < n; i++) {
        st.pops += results[i].touched;
        st.pushes += results[i].pushes;
        st.overflows += results[i].status != 0;
        st.kcycles_sum += results[i].kcycles;
        longest = std::max(longest, results[i].kcycles);
    }
    st.kcycles_longest += longest;
    st.h2d_bytes += (uint64_t)n * sizeof(pgi_query);
    st.d2h_bytes += (uint64_t)n * (sizeof(pgi_search_result) + (uint64_t)words * 4);
    return PGI_OK;
}

pgi_status pgi_graph_stats(pgi_ctx *ctx, pgi_search_stats *out, int32_t reset)
{
    if (!ctx) return PGI_ERR_INVALID;
    if (out) *out = ctx->sstats;
    if (reset) memset(&ctx->sstats, 0, sizeof ctx->sstats);
    return PGI_OK;
}

pgi_status pgi_guided_match(pgi_ctx *ctx, uint32_t n_src, const float *kp_src, const float *desc_src, uint32_t n_dst,
                            const float *kp_dst, const float *desc_dst, uint32_t dim, const double *pose_q_t,
                            const double *K_src, const double *K_dst, const int32_t *size_src, const int32_t *size_dst,
                            int32_t bin_number, uint32_t *matches_out, double *ratios_out, uint32_t *n_out,
                            double *prepared_or_null)
{
    PgiNvtxRange nvtxRange("pgi:guided_match (K7)");
    if (!ctx) return PGI_ERR_INVALID;
    if (!pose_q_t || !K_src || !K_dst || !size_src || !size_dst || !n_out || (n_src && (!kp_src || !desc_src || !matches_out || !ratios_out)) ||
        (n_dst && (!kp_dst || !desc_dst)) || dim == 0 || bin_number > kMatchMaxBins) {
        ctx->err = "bad argument";
        return PGI_ERR_INVALID;
    }
    CK(cudaSetDevice(ctx->cfg.device));
    *n_out = 0;
    if (n_src == 0 || n_dst == 0) return PGI_OK;
    cudaStream_t s = ctx->stream;
    const size_t bS = (size_t)n_src, bD = (size_t)n_dst;
    // one scratch allocation per call: this entry point is the reference-shaped single-pair call
    unsigned char *d = nullptr;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t oKpS = 0, oDS = oKpS + al(bS * 8), oKpD = oDS + al(bS * dim * 4), oDD = oKpD + al(bD * 8), oBin = oDD + al(bD * dim * 4),
                 oList = oBin + al(bD), oCand = oList + al(bD * 4), oRatio = oCand + al(bS * 4), oM = oRatio + al(bS * 8),
                 oR = oM + al(bS * 8), oN = oR + al(bS * 8), oPrep = oN + 256, total = oPrep + 256;
    CK(cudaMalloc((void **)&d, total));
    auto fail = [&](cudaError_t e) { cudaFree(d); ctx->err = cudaGetErrorString(e); return PGI_ERR_CUDA; };
    cudaError_t e;
    if ((e = cudaMemcpyAsync(d + oKpS, kp_src, bS * 8, cudaMemcpyHostToDevice, s)) != cudaSuccess) return fail(e);
    if ((e = cudaMemcpyAsync(d + oDS, desc_src, bS * dim * 4, cudaMemcpyHostToDevice, s)) != cudaSuccess) return fail(e);
    if ((e = cudaMemcpyAsync(d + oKpD, kp_dst, bD * 8, cudaMemcpyHostToDevice, s)) != cudaSuccess) return fail(e);
    if ((e = cudaMemcpyAsync(d + oDD, desc_dst, bD * dim * 4, cudaMemcpyHostToDevice, s)) != cudaSuccess) return fail(e);
    MatchArgs a;
    a.kpS = reinterpret_cast<const float2 *>(d + oKpS); a.dS = reinterpret_cast<const float *>(d + oDS);
    a.kpD = reinterpret_cast<const float2 *>(d + oKpD); a.dD = reinterpret_cast<const float *>(d + oDD);
    a.nS = n_src; a.nD = n_dst; a.dim = dim;
    {
        double E[9];
        // E = [t]x R of the pose (pose.h:46-55), as the matcher reads it from Pose::getEssentialMatrix (matcher.h:218)
        double R[9];
        const double *q = pose_q_t;
        const double tx = 2.0 * q[0], ty = 2.0 * q[1], tz = 2.0 * q[2];
        const double twx = tx * q[3], twy = ty * q[3], twz = tz * q[3], txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
        const double tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
        R[0] = 1.0 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
        R[3] = txy + twz; R[4] = 1.0 - (txx + tzz); R[5] = tyz - twx;
        R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1.0 - (txx + tyy);
        const double Cx[9] = {0.0, -q[6], q[5], q[6], 0.0, -q[4], -q[5], q[4], 0.0};
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) E[i * 3 + j] = Cx[i * 3 + 0] * R[j] + (Cx[i * 3 + 1] * R[3 + j] + Cx[i * 3 + 2] * R[6 + j]);
        for (int k = 0; k < 9; k++) { a.E[k] = E[k]; a.Ks[k] = K_src[k]; a.Kd[k] = K_dst[k]; }
    }
    a.wS = size_src[0]; a.hS = size_src[1]; a.wD = size_dst[0]; a.hD = size_dst[1]; a.binNumber = bin_number;
    a.binOfD = d + oBin; a.binList = reinterpret_cast<uint32_t *>(d + oList); a.cand = reinterpret_cast<uint32_t *>(d + oCand);
    a.candRatio = reinterpret_cast<double *>(d + oRatio); a.matches = reinterpret_cast<uint32_t *>(d + oM);
    a.ratios = reinterpret_cast<double *>(d + oR); a.nOut = reinterpret_cast<uint32_t *>(d + oN);
    a.prepOut = reinterpret_cast<double *>(d + oPrep);
    k7_guided_match<<<1, kMatchThreads, 0, s>>>(a);
    if ((e = cudaGetLastError()) != cudaSuccess) return fail(e);
    uint32_t n = 0;
    if ((e = cudaMemcpyAsync(&n, d + oN, 4, cudaMemcpyDeviceToHost, s)) != cudaSuccess) return fail(e);
    if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return fail(e);
    if (n > n_src) n = n_src;
    if (n) {
        if ((e = cudaMemcpy(matches_out, d + oM, (size_t)n * 8, cudaMemcpyDeviceToHost)) != cudaSuccess) return fail(e);
        if ((e = cudaMemcpy(ratios_out, d + oR, (size_t)n * 8, cudaMemcpyDeviceToHost)) != cudaSuccess) return fail(e);
    }
    if (prepared_or_null && (e = cudaMemcpy(prepared_or_null, d + oPrep, 14 * 8, cudaMemcpyDeviceToHost)) != cudaSuccess) return fail(e);
    cudaFree(d);
    *n_out = n;
    ctx->stats.launches += 1;
    ctx->stats.h2d_bytes += bS * (8 + dim * 4) + bD * (8 + dim * 4);
    ctx->stats.d2h_bytes += (size_t)n * 16 + 4;
    return PGI_OK;
}

pgi_status pgi_match_features(pgi_ctx *ctx, uint32_t n_src, const float *desc_src, uint32_t n_dst, const float *desc_dst,
                              uint32_t dim, uint32_t *matches_out, double *ratios_out, uint32_t *n_out)
{
    PgiNvtxRange nvtxRange("pgi:match_features (K8)");
    if (!ctx) return PGI_ERR_INVALID;
    if (!n_out || (n_src && (!desc_src || !matches_out || !ratios_out)) || (n_dst && !desc_dst) || (dim != 128 && dim != 64)) {
        ctx->err = "bad argument (descriptors must have 64 or 128 dimensions)";
        return PGI_ERR_INVALID;
    }
    CK(cudaSetDevice(ctx->cfg.device));
    *n_out = 0;
    if (n_src == 0 || n_dst == 0) return PGI_OK;
    cudaStream_t s = ctx->stream;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t bS = n_src, bD = n_dst;
    const size_t oA = 0, oB = oA + al(bS * dim * 4), oF = oB + al(bD * dim * 4), oG = oF + al(bS * sizeof(Knn2)),
                 oM = oG + al(bD * sizeof(Knn2)), oR = oM + al(bS * 8), oN = oR + al(bS * 8), total = oN + 256;
    unsigned char *d = nullptr;
    CK(cudaMalloc((void **)&d, total));
    auto fail = [&](cudaError_t e) { cudaFree(d); ctx->err = cudaGetErrorString(e); return PGI_ERR_CUDA; };
    cudaError_t e;
    if ((e = cudaMemcpyAsync(d + oA, desc_src, bS * dim * 4, cudaMemcpyHostToDevice, s)) != cudaSuccess) return fail(e);
    if ((e = cudaMemcpyAsync(d + oB, desc_dst, bD * dim * 4, cudaMemcpyHostToDevice, s)) != cudaSuccess) return fail(e);
    const float *A = reinterpret_cast<const float *>(d + oA), *B = reinterpret_cast<const float *>(d + oB);
    Knn2 *F = reinterpret_cast<Knn2 *>(d + oF), *G = reinterpret_cast<Knn2 *>(d + oG);
    if (dim == 128) {
        k8_knn2<128><<<(n_src + kFeatThreads - 1) / kFeatThreads, kFeatThreads, 0, s>>>(A, n_src, B, n_dst, F);  // feature_utils.h:158-159
        k8_knn2<128><<<(n_dst + kFeatThreads - 1) / kFeatThreads, kFeatThreads, 0, s>>>(B, n_dst, A, n_src, G);  // :162-163
    } else {
        k8_knn2<64><<<(n_src + kFeatThreads - 1) / kFeatThreads, kFeatThreads, 0, s>>>(A, n_src, B, n_dst, F);
        k8_knn2<64><<<(n_dst + kFeatThreads - 1) / kFeatThreads, kFeatThreads, 0, s>>>(B, n_dst, A, n_src, G);
    }
    k8_mutual<<<1, 1024, 0, s>>>(F, n_src, G, n_dst, reinterpret_cast<uint32_t *>(d + oM), reinterpret_cast<double *>(d + oR),
                                 reinterpret_cast<uint32_t *>(d + oN));
    if ((e = cudaGetLastError()) != cudaSuccess) return fail(e);
    uint32_t n = 0;
    if ((e = cudaMemcpyAsync(&n, d + oN, 4, cudaMemcpyDeviceToHost, s)) != cudaSuccess) return fail(e);
    if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return fail(e);
    if (n > n_src) n = n_src;
    std::vector<uint32_t> m((size_t)2 * n);
    std::vector<double> r(n);
    if (n) {
        if ((e = cudaMemcpy(m.data(), d + oM, (size_t)n * 8, cudaMemcpyDeviceToHost)) != cudaSuccess) return fail(e);
        if ((e = cudaMemcpy(r.data(), d + oR, (size_t)n * 8, cudaMemcpyDeviceToHost)) != cudaSuccess) return fail(e);
    }
    cudaFree(d);
    // std::sort of (ratio, &matches[i]) pairs (feature_utils.h:188): by ratio, then by the address of the query's
    // match vector, i.e. by query index
    std::vector<uint32_t> order(n);
    for (uint32_t i = 0; i < n; i++) order[i] = i;
    std::sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return r[x] != r[y] ? r[x] < r[y] : m[2 * x] < m[2 * y]; });
    for (uint32_t i = 0; i < n; i++) {
        matches_out[2 * i] = m[2 * order[i]];
        matches_out[2 * i + 1] = m[2 * order[i] + 1];
        ratios_out[i] = r[order[i]];
    }
    *n_out = n;
    ctx->stats.launches += 3;
    ctx->stats.h2d_bytes += (bS + bD) * dim * 4;
    ctx->stats.d2h_bytes += (size_t)n * 16 + 4;
    return PGI_OK;
}

pgi_status pgi_get_stats(pgi_ctx *ctx, pgi_stats *out)
{
    if (!ctx || !out) return PGI_ERR_INVALID;
    *out = ctx->stats;
    return PGI_OK;
}

pgi_status pgi_reset_stats(pgi_ctx *ctx)
{
    if (!ctx) return PGI_ERR_INVALID;
    memset(&ctx->stats, 0, sizeof ctx->stats);
    return PGI_OK;
}

// ---- unit-level device entry points for the parity tests -----------------------------------------
namespace {
__global__ void kdbg_sampson(const double4 *rows, uint64_t n, const double *E9, double *out)
{
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double E[9];
    for (int k = 0; k < 9; k++) E[k] = E9[k];
    const double4 c = rows[i];
    out[i] = sampsonSq(c.x, c.y, c.z, c.w, E);
}
__global__ void kdbg_five_point(const double *x1, const double *x2, uint32_t n, int dkMaxIters, double dkTolSq, double *Eout,
                                int *count)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double a[10], b[10];
    for (int k = 0; k < 10; k++) { a[k] = x1[10 * (size_t)i + k]; b[k] = x2[10 * (size_t)i + k]; }
    // 1000 sweeps (with or without a tolerance) = K2's configuration: the exact cycle jump; fewer sweeps = K4's plain loop
    count[i] = dkMaxIters >= 1000 ? fivePoint<true>(a, b, Eout + 90 * (size_t)i, 10, dkMaxIters, dkTolSq)
                                  : fivePoint<false>(a, b, Eout + 90 * (size_t)i, 10, dkMaxIters, dkTolSq);
}
}  // namespace

pgi_status pgi_dbg_sampson(pgi_ctx *ctx, const double *corr_xy4, uint64_t n, const double *E, double *out)
{
    if (!ctx || !corr_xy4 || !E || !out) return PGI_ERR_INVALID;
    CK(cudaSetDevice(ctx->cfg.device));
    double *d_c = nullptr, *d_E = nullptr, *d_o = nullptr;
    CK(cudaMalloc((void **)&d_c, std::max<uint64_t>(n, 1) * 32));
    CK(cudaMalloc((void **)&d_E, 72));
    CK(cudaMalloc((void **)&d_o, std::max<uint64_t>(n, 1) * 8));
    CK(cudaMemcpy(d_c, corr_xy4, n * 32, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_E, E, 72, cudaMemcpyHostToDevice));
    if (n) kdbg_sampson<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(reinterpret_cast<double4 *>(d_c), n, d_E, d_o);
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaMemcpy(out, d_o, n * 8, cudaMemcpyDeviceToHost));
    cudaFree(d_c); cudaFree(d_E); cudaFree(d_o);
    ctx->stats.launches += 1;
    return PGI_OK;
}

pgi_status pgi_dbg_five_point(pgi_ctx *ctx, const double *x1, const double *x2, uint32_t n_problems, int32_t dk_max_iters,
                              double dk_tol_sq, double *E_out, int32_t *count_out)
{
    if (!ctx || !x1 || !x2 || !E_out || !count_out) return PGI_ERR_INVALID;
    CK(cudaSetDevice(ctx->cfg.device));
    double *d_a = nullptr, *d_b = nullptr, *d_E = nullptr;
    int *d_c = nullptr;
    const size_t n = std::max<uint32_t>(n_problems, 1);
    CK(cudaMalloc((void **)&d_a, n * 80));
    CK(cudaMalloc((void **)&d_b, n * 80));
    CK(cudaMalloc((void **)&d_E, n * 720));
    CK(cudaMalloc((void **)&d_c, n * 4));
    CK(cudaMemset(d_E, 0, n * 720));
    CK(cudaMemcpy(d_a, x1, (size_t)n_problems * 80, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_b, x2, (size_t)n_problems * 80, cudaMemcpyHostToDevice));
    if (n_problems) kdbg_five_point<<<(n_problems + 63) / 64, 64, 0, ctx->stream>>>(d_a, d_b, n_problems, dk_max_iters, dk_tol_sq, d_E, d_c);
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    CK(cudaMemcpy(E_out, d_E, (size_t)n_problems * 720, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(count_out, d_c, (size_t)n_problems * 4, cudaMemcpyDeviceToHost));
    cudaFree(d_a); cudaFree(d_b); cudaFree(d_E); cudaFree(d_c);
    ctx->stats.launches += 1;
    return PGI_OK;
}

pgi_status pgi_dbg_pose_from_essential(pgi_ctx *ctx, const double *E, const double *corr_xy4, uint64_t n, double *R,
                                       double *t, uint64_t *votes)
{
    if (!ctx || !E || !corr_xy4 || !R || !t || !votes || n > 0x7fffffffULL) return PGI_ERR_INVALID;
    CK(cudaSetDevice(ctx->cfg.device));
    double *d_c = nullptr, *d_E = nullptr, *d_R = nullptr, *d_t = nullptr;
    unsigned long long *d_v = nullptr;
    CK(cudaMalloc((void **)&d_c, std::max<uint64_t>(n, 1) * 32));
    CK(cudaMalloc((void **)&d_E, 72));
    CK(cudaMalloc((void **)&d_R, 72));
    CK(cudaMalloc((void **)&d_t, 24));
    CK(cudaMalloc((void **)&d_v, 32));
    CK(cudaMemcpy(d_c, corr_xy4, n * 32, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_E, E, 72, cudaMemcpyHostToDevice));
    kdbg_pose_from_essential<<<1, kCtaThreads, 0, ctx->stream>>>(d_E, reinterpret_cast<double4 *>(d_c), (uint32_t)n, d_R, d_t, d_v);
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    unsigned long long hv[4];
    CK(cudaMemcpy(R, d_R, 72, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(t, d_t, 24, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hv, d_v, 32, cudaMemcpyDeviceToHost));
    for (int k = 0; k < 4; k++) votes[k] = hv[k];
    cudaFree(d_c); cudaFree(d_E); cudaFree(d_R); cudaFree(d_t); cudaFree(d_v);
    ctx->stats.launches += 1;
    return PGI_OK;
}

pgi_status pgi_dbg_fp64_peak(pgi_ctx *ctx, int32_t fused, double *tflops_out)
{
    if (!ctx || !tflops_out) return PGI_ERR_INVALID;
    CK(cudaSetDevice(ctx->cfg.device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, ctx->cfg.device));
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 14;
    double *d = nullptr;
    CK(cudaMalloc((void **)&d, (size_t)blocks * threads * 8));
    k_fp64_peak<<<blocks, threads, 0, ctx->stream>>>(d, 64, fused);  // warm-up
    double best = 0.0;
    for (int rep = 0; rep < 5; rep++) {
        CK(cudaEventRecord(ctx->evStart, ctx->stream));
        k_fp64_peak<<<blocks, threads, 0, ctx->stream>>>(d, iters, fused);
        CK(cudaEventRecord(ctx->evK1, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, ctx->evStart, ctx->evK1));
        const double flops = (double)blocks * threads * iters * 8.0 * 2.0;  // 8 chains x (mul + add)
        best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    cudaFree(d);
    ctx->stats.launches += 6;
    *tflops_out = best;
    return PGI_OK;
}

}  // extern "C"
