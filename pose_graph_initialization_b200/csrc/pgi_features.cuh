// pgi_features.cuh — K8: brute-force descriptor matching, the reference's matchFeatures (feature_utils.h:103-210):
//   BRUTEFORCE_SL2 knnMatch(k = 2) source->destination and destination->source (:151-163), a match survives if its
//   squared-L2 ratio best/second is below 0.90 AND it is the mutual nearest neighbour (:172-184), survivors are sorted by
//   that ratio (:188, PROSAC order) and written as (queryIdx, trainIdx, ratio) (:196-204).
//
// k8_knn2: one thread per query descriptor (its 128 floats live in registers), the train descriptors stream through shared
// memory in tiles every thread of the CTA reads at the same address (a broadcast: no bank conflicts); best and second-best
// squared distance with the index of the best, scanning train indices in ascending order with strict `<` (the first of
// equal distances stays first, as cv::batchDistance keeps it).  The distance is accumulated in FP32 in dimension order,
// multiply then add (-fmad=false: never fused) — the definition the oracle states; OpenCV's own SIMD accumulation order is
// not reproducible across builds, agreement with cv2 is checked at 1e-6 and on the surviving matches.
// k8_mutual: ratio test in double (`d1 < 0.90 * d2`: float < double), mutual check, ratio = d1 / d2 in FP32 widened,
// compaction in query order.  The sort by (ratio, query index) is host work on the short survivor list.
#pragma once
#include <cfloat>
#include <cstdint>
#include <cuda_runtime.h>

namespace pgi {

constexpr int kFeatThreads = 128;
constexpr int kFeatTileRows = 32;
constexpr int kFeatMaxDim = 128;

struct Knn2 {
    float d1, d2;
    uint32_t i1;
    uint32_t have;  // how many neighbours exist (0, 1 or 2)
};

template <int DIM>
__global__ void __launch_bounds__(kFeatThreads) k8_knn2(const float *__restrict__ query, uint32_t nQ, const float *__restrict__ train,
                                                        uint32_t nT, Knn2 *__restrict__ out)
{
    __shared__ __align__(16) float sT[kFeatTileRows * DIM];
    const uint32_t q = blockIdx.x * kFeatThreads + threadIdx.x;
    float a[DIM];
    if (q < nQ) {
#pragma unroll
        for (int k = 0; k < DIM; k += 4) {
            const float4 v = *reinterpret_cast<const float4 *>(query + (size_t)q * DIM + k);
            a[k] = v.x; a[k + 1] = v.y; a[k + 2] = v.z; a[k + 3] = v.w;
        }
    }
    float d1 = FLT_MAX, d2 = FLT_MAX;
    uint32_t i1 = 0xffffffffu;
    for (uint32_t t0 = 0; t0 < nT; t0 += kFeatTileRows) {
        const uint32_t rows = nT - t0 < (uint32_t)kFeatTileRows ? nT - t0 : (uint32_t)kFeatTileRows;
        __syncthreads();
        for (uint32_t e = threadIdx.x * 4; e < rows * DIM; e += kFeatThreads * 4)
            *reinterpret_cast<float4 *>(sT + e) = *reinterpret_cast<const float4 *>(train + (size_t)t0 * DIM + e);
        __syncthreads();
        if (q < nQ) {
            for (uint32_t r = 0; r < rows; r++) {
                const float *b = sT + r * DIM;
                float acc = 0.0f;
#pragma unroll
                for (int k = 0; k < DIM; k += 4) {
                    const float4 v = *reinterpret_cast<const float4 *>(b + k);
                    float d = a[k] - v.x; acc = acc + d * d;
                    d = a[k + 1] - v.y; acc = acc + d * d;
                    d = a[k + 2] - v.z; acc = acc + d * d;
                    d = a[k + 3] - v.w; acc = acc + d * d;
                }
                if (acc < d1) { d2 = d1; d1 = acc; i1 = t0 + r; }
                else if (acc < d2) d2 = acc;
            }
        }
    }
    if (q < nQ) out[q] = Knn2{d1, d2, i1, nT >= 2 ? 2u : nT};
}

// feature_utils.h:172-184 for every query i, survivors compacted in query order by one CTA-wide scan per 1 024 queries
__global__ void __launch_bounds__(1024) k8_mutual(const Knn2 *__restrict__ fwd, uint32_t nQ, const Knn2 *__restrict__ bwd, uint32_t nT,
                                                  uint32_t *__restrict__ matches /*2 per survivor*/, double *__restrict__ ratios,
                                                  uint32_t *__restrict__ nOut)
{
    __shared__ uint32_t sWarp[32], sBase;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) sBase = 0;
    __syncthreads();
    for (uint32_t i0 = 0; i0 < nQ; i0 += 1024) {
        const uint32_t i = i0 + threadIdx.x;
        bool ok = false;
        Knn2 f{};
        if (i < nQ) {
            f = fwd[i];
            if (f.have >= 2 && bwd[f.i1].have >= 2)  // :174-176
                ok = ((double)f.d1 < 0.90 * (double)f.d2) && (bwd[f.i1].i1 == i);  // :178-179
        }
        const uint32_t m = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) sWarp[warp] = __popc(m);
        __syncthreads();
        uint32_t off = sBase;
        for (int w = 0; w < warp; w++) off += sWarp[w];
        if (ok) {
            const uint32_t o = off + __popc(m & ((1u << lane) - 1u));
            matches[2 * o] = i;
            matches[2 * o + 1] = f.i1;
            ratios[o] = (double)(f.d1 / f.d2);  // :182 float / float
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t t = 0;
            for (int w = 0; w < 32; w++) t += sWarp[w];
            sBase += t;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *nOut = sBase;
}

}  // namespace pgi
