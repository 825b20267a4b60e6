// csrc/match.cu — brute-force Hamming kNN-2 + Lowe ratio (+ optional mutual-NN cross-check).
// Replaces cv::BFMatcher(NORM_HAMMING)->knnMatch(q, t, k=2) and the ratio test of Matcher::KnnMatch
// (reference Features/matcher.cpp:55-66, :23-35); result order = (distance asc, trainIdx asc), which the
// packed key (dist << 16 | trainIdx) reproduces under unsigned min.
//
// Register-tiled POPC/LOP3 kernel, integer-ALU bound (8 POPC32 per descriptor pair), no tensor cores:
// each thread keeps R query descriptors (R x 8 words) in registers; a warp walks its share of the train
// descriptors staged in shared memory with broadcast 128-bit loads, so one LDS pair feeds 32*R pairs.
// Cross-check (north-star extension, quirk Q10) reuses the same distances: per train row the warp takes
// REDUX.MIN over (dist << 16 | queryIdx) and folds it into a shared / global atomicMin.
#include "orbf_internal.h"

namespace {

constexpr int KN_R = 4, KN_WARPS = 4, KN_THREADS = KN_WARPS * 32, KN_QT = 32 * KN_R, KN_CHUNK = 1024;
constexpr uint32_t KEY_NONE = 0xFFFFFFFFu;

__device__ __forceinline__ void top2_insert(uint32_t& m1, uint32_t& m2, uint32_t key)
{
    m2 = min(m2, max(m1, key));
    m1 = min(m1, key);
}

template <bool CROSS>
__global__ void __launch_bounds__(KN_THREADS) knn2_kernel(MatchSet ms, int K)
{
    uint32_t* __restrict__ knn = ms.knn;
    uint32_t* __restrict__ rev = ms.rev;
    __shared__ __align__(16) uint32_t sT[KN_CHUNK * 8];
    __shared__ uint32_t sRev[CROSS ? KN_CHUNK : 1];
    __shared__ uint32_t sMerge[KN_WARPS][KN_R][2][32];
    const int pair = ms.pair0 + blockIdx.y;
    int qs = 0, ts = 0;
    if (ms.pairs) { qs = ms.pairs[2 * pair]; ts = ms.pairs[2 * pair + 1]; }
    const int nq = ms.qCounts ? ms.qCounts[qs] : ms.nq, nt = ms.tCounts ? ms.tCounts[ts] : ms.nt;
    const int qBase = blockIdx.x * KN_QT;
    if (qBase >= nq) return;
    const uint8_t* Q = ms.qdesc + (long long)qs * ms.qStride;
    const uint8_t* T = ms.tdesc + (long long)ts * ms.tStride;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    uint32_t q[KN_R][8], m1[KN_R], m2[KN_R];
#pragma unroll
    for (int r = 0; r < KN_R; ++r) {
        const int qi = qBase + r * 32 + lane;
        m1[r] = m2[r] = KEY_NONE;
        if (qi < nq) {
            const uint4 a = __ldg(reinterpret_cast<const uint4*>(Q + (long long)qi * 32));
            const uint4 b = __ldg(reinterpret_cast<const uint4*>(Q + (long long)qi * 32 + 16));
            q[r][0] = a.x; q[r][1] = a.y; q[r][2] = a.z; q[r][3] = a.w; q[r][4] = b.x; q[r][5] = b.y; q[r][6] = b.z; q[r][7] = b.w;
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) q[r][i] = 0;
        }
    }
    for (int c0 = 0; c0 < nt; c0 += KN_CHUNK) {
        const int cn = min(KN_CHUNK, nt - c0);
        __syncthreads();
        for (int i = threadIdx.x; i < cn * 2; i += KN_THREADS)
            reinterpret_cast<uint4*>(sT)[i] = __ldg(reinterpret_cast<const uint4*>(T + (long long)c0 * 32) + i);
        if (CROSS) for (int i = threadIdx.x; i < cn; i += KN_THREADS) sRev[i] = KEY_NONE;
        __syncthreads();
        for (int j = warp; j < cn; j += KN_WARPS) {
            const uint4 ta = reinterpret_cast<const uint4*>(sT)[2 * j], tb = reinterpret_cast<const uint4*>(sT)[2 * j + 1];
            uint32_t kmin = KEY_NONE;
#pragma unroll
            for (int r = 0; r < KN_R; ++r) {
                const int d = __popc(q[r][0] ^ ta.x) + __popc(q[r][1] ^ ta.y) + __popc(q[r][2] ^ ta.z) + __popc(q[r][3] ^ ta.w)
                    + __popc(q[r][4] ^ tb.x) + __popc(q[r][5] ^ tb.y) + __popc(q[r][6] ^ tb.z) + __popc(q[r][7] ^ tb.w);
                top2_insert(m1[r], m2[r], ((uint32_t)d << 16) | (uint32_t)(c0 + j));
                if (CROSS) {
                    const int qi = qBase + r * 32 + lane;
                    if (qi < nq) kmin = min(kmin, ((uint32_t)d << 16) | (uint32_t)qi);
                }
            }
            if (CROSS) {
                const uint32_t wmin = __reduce_min_sync(0xffffffffu, kmin);
                if (lane == 0 && wmin != KEY_NONE) atomicMin(&sRev[j], wmin);
            }
        }
        if (CROSS) {
            __syncthreads();
            for (int i = threadIdx.x; i < cn; i += KN_THREADS)
                if (sRev[i] != KEY_NONE) atomicMin(&rev[(long long)pair * K + c0 + i], sRev[i]);
        }
    }
    // merge the warps' partial top-2 lists
#pragma unroll
    for (int r = 0; r < KN_R; ++r) { sMerge[warp][r][0][lane] = m1[r]; sMerge[warp][r][1][lane] = m2[r]; }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int r = 0; r < KN_R; ++r) {
            uint32_t a = KEY_NONE, b = KEY_NONE;
#pragma unroll
            for (int w = 0; w < KN_WARPS; ++w) { top2_insert(a, b, sMerge[w][r][0][lane]); top2_insert(a, b, sMerge[w][r][1][lane]); }
            const int qi = qBase + r * 32 + lane;
            if (qi < nq) {
                uint2* o = reinterpret_cast<uint2*>(knn + ((long long)pair * K + qi) * 2);
                *o = make_uint2(a, b);
            }
        }
    }
}

constexpr int MS_THREADS = 256;

__global__ void __launch_bounds__(MS_THREADS) match_select_kernel(MatchSet ms, int K, float ratio, int cross)
{
    const uint32_t* __restrict__ knn = ms.knn;
    const uint32_t* __restrict__ rev = ms.rev;
    orbf_dmatch* __restrict__ matches = ms.matches;
    int* __restrict__ matchCount = ms.matchCount;
    __shared__ int sWarp[MS_THREADS / 32];
    __shared__ int sBase;
    const int pair = ms.pair0 + blockIdx.x;
    int qs = 0, ts = 0;
    if (ms.pairs) { qs = ms.pairs[2 * pair]; ts = ms.pairs[2 * pair + 1]; }
    const int nq = ms.qCounts ? ms.qCounts[qs] : ms.nq, nt = ms.tCounts ? ms.tCounts[ts] : ms.nt;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) sBase = 0;
    __syncthreads();
    orbf_dmatch* out = matches + (long long)pair * K;
    for (int base = 0; base < nq; base += MS_THREADS) {
        const int i = base + threadIdx.x;
        bool keep = false;
        uint32_t k1 = 0;
        if (i < nq && nt >= 2) {
            const uint2 kk = *reinterpret_cast<const uint2*>(knn + ((long long)pair * K + i) * 2);
            k1 = kk.x;
            if (kk.y != KEY_NONE) {
                const float d1 = (float)(kk.x >> 16), d2 = (float)(kk.y >> 16);
                keep = d1 < __fmul_rn(ratio, d2);                       // m1.distance < mfNNratio * m2.distance (float)
                if (keep && cross) keep = (rev[(long long)pair * K + (kk.x & 0xFFFFu)] & 0xFFFFu) == (uint32_t)i;
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) sWarp[warp] = __popc(m);
        __syncthreads();
        int off = sBase;
        for (int w = 0; w < warp; ++w) off += sWarp[w];
        if (keep && matches) {
            orbf_dmatch dm;
            dm.queryIdx = i; dm.trainIdx = (int)(k1 & 0xFFFFu); dm.imgIdx = 0; dm.distance = (float)(k1 >> 16);
            out[off + __popc(m & ((1u << lane) - 1))] = dm;
        }
        __syncthreads();
        if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < MS_THREADS / 32; ++w) t += sWarp[w]; sBase += t; }
        __syncthreads();
    }
    if (threadIdx.x == 0) matchCount[pair] = sBase;
}

}  // namespace

int orbf_launch_knn2(orbf_context* c, const MatchSet& ms, int npairs, bool cross)
{
    const int maxNq = ms.qCounts ? c->K : ms.nq;
    if (maxNq <= 0 || npairs <= 0) return ORBF_OK;
    dim3 grid((maxNq + KN_QT - 1) / KN_QT, npairs);
    if (cross) {
        ORBF_CUDA(c, cudaMemsetAsync(ms.rev + (size_t)ms.pair0 * c->K, 0xFF, (size_t)npairs * c->K * sizeof(uint32_t), c->stream));
        knn2_kernel<true><<<grid, KN_THREADS, 0, c->stream>>>(ms, c->K);
    } else knn2_kernel<false><<<grid, KN_THREADS, 0, c->stream>>>(ms, c->K);
    ORBF_LAUNCH_CHECK(c);
    return ORBF_OK;
}

int orbf_launch_match_select(orbf_context* c, const MatchSet& ms, int npairs, float ratio, bool cross)
{
    if (npairs <= 0) return ORBF_OK;
    match_select_kernel<<<npairs, MS_THREADS, 0, c->stream>>>(ms, c->K, ratio, cross ? 1 : 0);
    ORBF_LAUNCH_CHECK(c);
    return ORBF_OK;
}
