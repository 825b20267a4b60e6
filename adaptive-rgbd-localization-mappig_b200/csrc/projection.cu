// csrc/projection.cu — Matcher::ProjectionMatch (reference Features/matcher.cpp:90-143): landmarks already
// projected into a frame are matched to the frame's features inside a square window (Frame::GetFeaturesInArea,
// Core/frame.cpp:258-274: a linear scan in feature order, |dx| < r && |dy| < r in float).
//
// The reference walks the landmarks in order and a feature taken by an earlier landmark (one with Observations() > 0) is skipped
// by later ones, so the result depends on the order.  Two kernels keep that exact:
//   1. proj_candidates_kernel (warp per landmark, all landmarks in parallel): window test + 256-bit Hamming distance for every
//      feature, candidates compacted in feature order as (distance << 16 | feature) keys — the order-independent part, and
//      all of the arithmetic.
//   2. proj_resolve_kernel (one warp): landmarks in order; the lanes take a landmark's candidates that are still free, the two
//      smallest keys are exactly the reference's (best, second best) under its strict '<' updates in arrival order; octave /
//      ratio test (double); the accepted feature is marked taken.  A window of 8 px holds ~1 candidate, so this pass is a few
//      tens of instructions per landmark.
#include "orbf_internal.h"

namespace {

constexpr uint32_t PJ_NONE = 0xFFFFFFFFu;

__global__ void __launch_bounds__(128) proj_candidates_kernel(const float* __restrict__ kpx, const float* __restrict__ kpy, const uint8_t* __restrict__ desc, int nFeat,
    const uint8_t* __restrict__ lmDesc, const float* __restrict__ projX, const float* __restrict__ projY, const uint8_t* __restrict__ lmFlags, int nLm, float radius,
    uint32_t* __restrict__ cand, int* __restrict__ candCount)
{
    const int lane = threadIdx.x & 31, i = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (i >= nLm) return;
    int n = 0;
    if (lmFlags[i] & 1) {
        const uint4 la = __ldg(reinterpret_cast<const uint4*>(lmDesc + (size_t)i * 32)), lb = __ldg(reinterpret_cast<const uint4*>(lmDesc + (size_t)i * 32) + 1);
        const float px = projX[i], py = projY[i];
        uint32_t* out = cand + (size_t)i * nFeat;
        for (int j0 = 0; j0 < nFeat; j0 += 32) {
            const int j = j0 + lane;
            bool in = false;
            if (j < nFeat) {
                const float dx = __fsub_rn(kpx[j], px), dy = __fsub_rn(kpy[j], py);
                in = fabsf(dx) < radius && fabsf(dy) < radius;
            }
            const unsigned m = __ballot_sync(0xffffffffu, in);
            if (in) {
                const uint4 fa = __ldg(reinterpret_cast<const uint4*>(desc + (size_t)j * 32)), fb = __ldg(reinterpret_cast<const uint4*>(desc + (size_t)j * 32) + 1);
                const int d = __popc(la.x ^ fa.x) + __popc(la.y ^ fa.y) + __popc(la.z ^ fa.z) + __popc(la.w ^ fa.w)
                            + __popc(lb.x ^ fb.x) + __popc(lb.y ^ fb.y) + __popc(lb.z ^ fb.z) + __popc(lb.w ^ fb.w);
                out[n + __popc(m & ((1u << lane) - 1))] = ((uint32_t)d << 16) | (uint32_t)j;
            }
            n += __popc(m);
        }
    }
    if (lane == 0) candCount[i] = n;
}

__device__ __forceinline__ void pj_insert(uint32_t& m1, uint32_t& m2, uint32_t key)
{
    m2 = min(m2, max(m1, key));
    m1 = min(m1, key);
}

__global__ void __launch_bounds__(32) proj_resolve_kernel(const uint32_t* __restrict__ cand, const int* __restrict__ candCount, const int* __restrict__ kpOct, int nFeat,
    const uint8_t* __restrict__ lmFlags, int nLm, const uint8_t* __restrict__ featTaken, float nnRatio, int thHigh, uint8_t* taken /* [nFeat] scratch */,
    int* __restrict__ bestIdx, int* __restrict__ nMatches)
{
    const int lane = threadIdx.x;
    for (int j = lane; j < nFeat; j += 32) taken[j] = featTaken ? featTaken[j] : 0;
    __syncwarp();
    int nm = 0;
    for (int i = 0; i < nLm; ++i) {
        int best = -1;
        const int n = (lmFlags[i] & 1) ? candCount[i] : 0;
        if (n > 0) {
            uint32_t k1 = PJ_NONE, k2 = PJ_NONE;
            for (int c = lane; c < n; c += 32) {
                const uint32_t key = cand[(size_t)i * nFeat + c];
                if (!taken[key & 0xFFFFu]) pj_insert(k1, k2, key);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const uint32_t o1 = __shfl_xor_sync(0xffffffffu, k1, o), o2 = __shfl_xor_sync(0xffffffffu, k2, o);
                pj_insert(k1, k2, o1);
                pj_insert(k1, k2, o2);
            }
            if (k1 != PJ_NONE && (int)(k1 >> 16) <= thHigh) {
                const int j1 = (int)(k1 & 0xFFFFu);
                bool ok = true;
                if (k2 != PJ_NONE) {
                    const int j2 = (int)(k2 & 0xFFFFu);
                    // bestLevel == bestLevel2 && bestDist1 > mfNNratio * bestDist2 (float ratio promoted to double)
                    if (kpOct[j1] == kpOct[j2] && (double)(k1 >> 16) > __dmul_rn((double)nnRatio, (double)(k2 >> 16))) ok = false;
                }
                if (ok) {
                    best = j1;
                    ++nm;
                    if (lane == 0 && (lmFlags[i] & 2)) taken[j1] = 1;
                }
            }
            __syncwarp();
        }
        if (lane == 0) bestIdx[i] = best;
    }
    if (lane == 0) *nMatches = nm;
}

}  // namespace

int orbf_launch_projection_match(orbf_context* c, const float* d_kpx, const float* d_kpy, const int* d_kpoct, const uint8_t* d_desc, int nFeat, const uint8_t* d_lmDesc,
    const float* d_projX, const float* d_projY, const uint8_t* d_lmFlags, int nLm, const uint8_t* d_featTaken, float radius, float nnRatio, int thHigh,
    uint32_t* d_cand, int* d_candCount, uint8_t* d_taken, int* d_bestIdx, int* d_nMatches)
{
    if (nLm <= 0) return ORBF_OK;
    proj_candidates_kernel<<<(nLm + 3) / 4, 128, 0, c->stream>>>(d_kpx, d_kpy, d_desc, nFeat, d_lmDesc, d_projX, d_projY, d_lmFlags, nLm, radius, d_cand, d_candCount);
    ORBF_LAUNCH_CHECK(c);
    proj_resolve_kernel<<<1, 32, 0, c->stream>>>(d_cand, d_candCount, d_kpoct, nFeat, d_lmFlags, nLm, d_featTaken, nnRatio, thHigh, d_taken, d_bestIdx, d_nMatches);
    ORBF_LAUNCH_CHECK(c);
    return ORBF_OK;
}
