// csrc/match.cu — brute-force Hamming kNN-2 + Lowe ratio (+ optional mutual-NN cross-check).
// Replaces cv::BFMatcher(NORM_HAMMING)->knnMatch(q, t, k=2) and the ratio test of Matcher::KnnMatch
// (reference Features/matcher.cpp:55-66, :23-35); result order = (distance asc, trainIdx asc), which the
// packed key (dist << 16 | trainIdx) reproduces under unsigned min.
//
// Register-tiled POPC/LOP3 kernel, integer-ALU bound, no tensor cores: each thread keeps R query descriptors
// (R x 8 words) in registers; a warp walks its share of the train descriptors staged in shared memory with
// broadcast 128-bit loads, so one LDS pair feeds 32*R pairs.  The 256-bit popcount runs through a carry-save
// adder tree (4 CSAs = 8 LOP3 turn the 8 XOR words into words of weight 1,1,2,4), so a pair costs 4 POPC
// instead of 8: B200's POPC pipe issues 16 lanes/clk/SM against 64 for LOP3 (profiles/int_pipe_peaks.json), which
// moves the bound from the POPC pipe (0.5 clk/pair/SM) to the ALU pipe (~0.27).  The running top-2 is only
// touched when a pair beats the current second best (one ISETP per pair on the common path).
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

__device__ __forceinline__ uint32_t lop3_xor3(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t lop3_maj(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// Hamming distance of two 256-bit rows, returned as (popc(w1a) + popc(w1b)) and the weight-2 / weight-4 words' counts
// folded by the caller with IMADs (FMA pipe): d = p1a + p1b + 2*p2 + 4*p4.
__device__ __forceinline__ uint32_t hamming256_csa(const uint32_t q[8], const uint4& ta, const uint4& tb)
{
    const uint32_t x0 = q[0] ^ ta.x, x1 = q[1] ^ ta.y, x2 = q[2] ^ ta.z, x3 = q[3] ^ ta.w;
    const uint32_t x4 = q[4] ^ tb.x, x5 = q[5] ^ tb.y, x6 = q[6] ^ tb.z, x7 = q[7] ^ tb.w;
    const uint32_t s1 = lop3_xor3(x0, x1, x2), c1 = lop3_maj(x0, x1, x2);
    const uint32_t s2 = lop3_xor3(x3, x4, x5), c2 = lop3_maj(x3, x4, x5);
    const uint32_t s3 = lop3_xor3(s1, s2, x6), c3 = lop3_maj(s1, s2, x6);
    const uint32_t s4 = lop3_xor3(c1, c2, c3), c4 = lop3_maj(c1, c2, c3);
    return (uint32_t)__popc(c4) * 4u + ((uint32_t)__popc(s4) * 2u + (uint32_t)(__popc(s3) + __popc(x7)));
}

template <bool CROSS>
__global__ void __launch_bounds__(KN_THREADS, 6) knn2_kernel(MatchSet ms, int K)
{
    uint32_t* __restrict__ knn = ms.knn;
    uint32_t* __restrict__ rev = ms.rev;
    __shared__ __align__(16) uint32_t sT[KN_CHUNK * 8];
    __shared__ uint32_t sRev[CROSS ? KN_CHUNK : 1];
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
    static_assert(KN_R == 4, "cross-check reduction is written for 4 queries per thread");
    uint32_t qkey[KN_R];                    // query index = low half of the reverse key; rows past nq get a key above every valid one
#pragma unroll
    for (int r = 0; r < KN_R; ++r) { const int qi = qBase + r * 32 + lane; qkey[r] = qi < nq ? (uint32_t)qi : 0xF0000000u; }   // d << 16 <= 0x01000000: no wrap
    for (int c0 = 0; c0 < nt; c0 += KN_CHUNK) {
        const int cn = min(KN_CHUNK, nt - c0);
        __syncthreads();
        for (int i = threadIdx.x; i < cn * 2; i += KN_THREADS)
            reinterpret_cast<uint4*>(sT)[i] = __ldg(reinterpret_cast<const uint4*>(T + (long long)c0 * 32) + i);
        if (CROSS) for (int i = threadIdx.x; i < cn; i += KN_THREADS) sRev[i] = KEY_NONE;
        __syncthreads();
        for (int j = warp; j < cn; j += KN_WARPS) {
            const uint4 ta = reinterpret_cast<const uint4*>(sT)[2 * j], tb = reinterpret_cast<const uint4*>(sT)[2 * j + 1];
            const uint32_t idx = (uint32_t)(c0 + j);
            uint32_t key[KN_R];
#pragma unroll
            for (int r = 0; r < KN_R; ++r) key[r] = hamming256_csa(q[r], ta, tb) * 65536u + idx;
            // train rows arrive in ascending index order, so a pair enters the top-2 iff its key is below the second best
            bool any = false;
#pragma unroll
            for (int r = 0; r < KN_R; ++r) any |= key[r] < m2[r];
            if (any) {
#pragma unroll
                for (int r = 0; r < KN_R; ++r) top2_insert(m1[r], m2[r], key[r]);
            }
            if (CROSS) {
                // per train row: best query of this warp's 128 (rows past nq carry an all-ones query key)
                const uint32_t a = min(key[0] - idx + qkey[0], key[1] - idx + qkey[1]), b = min(key[2] - idx + qkey[2], key[3] - idx + qkey[3]);
                const uint32_t wmin = __reduce_min_sync(0xffffffffu, min(a, b));
                if (lane == 0 && wmin < 0xF0000000u) atomicMin(&sRev[j], wmin);
            }
        }
        if (CROSS) {
            __syncthreads();
            for (int i = threadIdx.x; i < cn; i += KN_THREADS)
                if (sRev[i] != KEY_NONE) atomicMin(&rev[(long long)pair * K + c0 + i], sRev[i]);
        }
    }
    // merge the warps' partial top-2 lists (the staging area aliases the train rows, which are no longer needed)
    __syncthreads();
    uint32_t (*sMerge)[KN_R][2][32] = reinterpret_cast<uint32_t (*)[KN_R][2][32]>(sT);
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

// ---- Landmark::ComputeDistinctiveDescriptors (Core/landmark.cpp:219-273): one warp per landmark ------------------------------
// Lane j holds observation j's descriptor (chunks of 32 observations); for every row i the distances to all observations are
// formed with the same XOR/POPC arithmetic, the row median sorted[(size_t)(0.5 * (N - 1))] is found by rank counting over
// shuffles, and the first row with the strictly smallest median wins.
constexpr int DD_MAX_OBS = 128;

__global__ void __launch_bounds__(128) distinctive_kernel(const uint8_t* __restrict__ desc, const int* __restrict__ offsets, int nLandmarks,
    int* __restrict__ best, int* __restrict__ bestMedian)
{
    __shared__ uint16_t sDist[4][DD_MAX_OBS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int l = blockIdx.x * 4 + warp;
    if (l >= nLandmarks) return;
    const int a = offsets[l], N = min(offsets[l + 1] - a, DD_MAX_OBS);
    if (N <= 0) { if (lane == 0) { best[l] = -1; if (bestMedian) bestMedian[l] = -1; } return; }
    const int m = (int)(0.5 * (double)(N - 1));
    int bestIdx = 0, bestMed = 0x7fffffff;
    uint16_t* dist = sDist[warp];
    for (int i = 0; i < N; ++i) {
        const uint4 ia = __ldg(reinterpret_cast<const uint4*>(desc + (size_t)(a + i) * 32)), ib = __ldg(reinterpret_cast<const uint4*>(desc + (size_t)(a + i) * 32) + 1);
        for (int j = lane; j < N; j += 32) {
            const uint4 ja = __ldg(reinterpret_cast<const uint4*>(desc + (size_t)(a + j) * 32)), jb = __ldg(reinterpret_cast<const uint4*>(desc + (size_t)(a + j) * 32) + 1);
            dist[j] = (uint16_t)(__popc(ia.x ^ ja.x) + __popc(ia.y ^ ja.y) + __popc(ia.z ^ ja.z) + __popc(ia.w ^ ja.w)
                + __popc(ib.x ^ jb.x) + __popc(ib.y ^ jb.y) + __popc(ib.z ^ jb.z) + __popc(ib.w ^ jb.w));
        }
        __syncwarp();
        // the m-th smallest: the value v with #{< v} <= m < #{<= v}
        int med = -1;
        for (int j = lane; j < N; j += 32) {
            const int v = dist[j];
            int less = 0, leq = 0;
            for (int k = 0; k < N; ++k) { const int u = dist[k]; less += u < v; leq += u <= v; }
            if (less <= m && m < leq) med = v;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) med = max(med, __shfl_xor_sync(0xffffffffu, med, o));
        if (med < bestMed) { bestMed = med; bestIdx = i; }
        __syncwarp();
    }
    if (lane == 0) { best[l] = bestIdx; if (bestMedian) bestMedian[l] = bestMed; }
}

}  // namespace

int orbf_launch_distinctive(orbf_context* c, const uint8_t* d_desc, const int* d_offsets, int nLandmarks, int* d_best, int* d_median)
{
    if (nLandmarks <= 0) return ORBF_OK;
    distinctive_kernel<<<(nLandmarks + 3) / 4, 128, 0, c->stream>>>(d_desc, d_offsets, nLandmarks, d_best, d_median);
    ORBF_LAUNCH_CHECK(c);
    return ORBF_OK;
}

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
