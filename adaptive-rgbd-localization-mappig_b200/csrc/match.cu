// csrc/match.cu — brute-force Hamming kNN-2 + Lowe ratio (+ optional mutual-NN cross-check).
// Replaces cv::BFMatcher(NORM_HAMMING)->knnMatch(q, t, k=2) and the ratio test of Matcher::KnnMatch
// (reference Features/matcher.cpp:55-66, :23-35); result order = (distance asc, trainIdx asc), which the
// packed key (dist << 16 | trainIdx) reproduces under unsigned min.
//
// The O(nq * nt * 256) distance matrix is the one compute-bound piece of the path (64 KB of descriptors per pair, 2.6e8 bit
// operations), so it runs on the tensor cores as an exact integer GEMM: every descriptor bit becomes a signed byte (1 -> +1,
// 0 -> -1), dot(a, b) over the 256 bytes = 256 - 2 * hamming(a, b), and IMMA (mma.sync m16n8k32 s8*s8 -> s32) produces
// 16 x 8 distances per 8 instructions.  The previous POPC/LOP3 kernel needed ~100 warp instructions for the same 128
// distances and was bound by the integer ALU pipe (0.63e12 distances/s); B200's s8 mma.sync path sustains 0.48 IMMA/clk/SM
// = 2.2e12 distances/s (tools/bmma_probe.cu).  The bit -> byte expansion is one PRMT per 4 bytes (the selector nibbles are the
// bits, the pool bytes are 0xFF / 0x01); the order of the 256 dimensions is irrelevant as long as both operands use the
// same one, so the expansion is laid out to make the fragment loads contiguous.
//
// CTA = 8 warps = 256 query rows (two m16 tiles per warp, A fragments resident in 64 registers); train rows are expanded in
// chunks of 256 into shared memory as [k-step pair][column][lane-in-group][4 words], which makes every B-fragment load a
// conflict-free LDS.128.  The epilogue is branch-free and packed two columns per register (VIMNMX.U16x2): per n8 tile and
// row slot a running top-2 of 15-bit (256 - distance, tile) codes, decoded into (distance << 16 | trainIdx) keys once per
// chunk; for the cross-check, the maximum over the thread's rows and 3 shuffles over the 8 row groups give the warp's best
// (256 - distance, row) per column, stored per warp and folded over the 8 warps at the end of the chunk.
#include "orbf_internal.h"

namespace {

constexpr int KM_WARPS = 8, KM_THREADS = KM_WARPS * 32, KM_ROWS = KM_WARPS * 32, KM_CHUNK = 256;
constexpr size_t KM_SMEM = (size_t)4 * KM_CHUNK * 16 * sizeof(uint32_t) + KM_WARPS * (KM_CHUNK / 2) * sizeof(uint32_t);
constexpr uint32_t KEY_NONE = 0xFFFFFFFFu;
// Accumulator encoding: train bytes are +-32 and query bytes +-1, so an accumulator that starts at 32 * 256 + rc holds
// v = 64 * h + rc with h = 256 - hamming (0..256) and rc = 32 - (row within the warp's 32) in the low 6 bits: one 15-bit
// value that orders a column's candidates by (distance asc, query asc) and still tells which row it came from.  Rows past nq
// have all-zero bytes and start at 0 (v = 0, rc = 0 marks them).
constexpr int V_SHIFT = 6, V_START = 32 * 256;

__device__ __forceinline__ void top2_insert(uint32_t& m1, uint32_t& m2, uint32_t key)
{
    m2 = min(m2, max(m1, key));
    m1 = min(m1, key);
}

// 4 descriptor bits (bits s, s+4, s+8, s+12 of `half`) -> 4 signed bytes (+MAG / -MAG)
template <int MAG>
__device__ __forceinline__ uint32_t expand4(uint32_t half, int s)
{
    constexpr uint32_t pool = ((uint32_t)MAG << 8) | (uint32_t)((256 - MAG) & 0xFF);          // byte 0 = -MAG (bit clear), byte 1 = +MAG (bit set)
    return __byte_perm(pool, 0u, (half >> s) & 0x1111u);
}

__device__ __forceinline__ uint32_t pack16(int lo, int hi)           // lo | hi << 16 on the FMA pipe (the ALU pipe is the busy one)
{
    uint32_t r;
    asm("mad.lo.u32 %0, %1, 65536, %2;" : "=r"(r) : "r"(hi), "r"(lo));
    return r;
}

__device__ __forceinline__ void imma_first(int (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1, int c01, int c23)
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%11,%11};\n"
                 : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "r"(c01), "r"(c23));
}
__device__ __forceinline__ void imma_acc(int (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <bool CROSS>
__global__ void __launch_bounds__(KM_THREADS, 2) knn2_kernel(MatchSet ms, int K)
{
    extern __shared__ __align__(16) uint32_t kmSmem[];
    uint4* sB4 = reinterpret_cast<uint4*>(kmSmem);                 // [4][KM_CHUNK][4] uint4
    uint32_t* sCol = kmSmem + 4 * KM_CHUNK * 16;                   // [KM_WARPS][KM_CHUNK / 2] per-warp column maxima, two columns per word
    uint32_t* __restrict__ knn = ms.knn;
    uint32_t* __restrict__ rev = ms.rev;
    const int pair = ms.pair0 + blockIdx.y;
    int qs = 0, ts = 0;
    if (ms.pairs) { qs = ms.pairs[2 * pair]; ts = ms.pairs[2 * pair + 1]; }
    const int nq = ms.qCounts ? ms.qCounts[qs] : ms.nq, nt = ms.tCounts ? ms.tCounts[ts] : ms.nt;
    const int qBase = blockIdx.x * KM_ROWS;
    if (qBase >= nq) return;
    const uint32_t* Q = reinterpret_cast<const uint32_t*>(ms.qdesc + (long long)qs * ms.qStride);
    const uint32_t* T = reinterpret_cast<const uint32_t*>(ms.tdesc + (long long)ts * ms.tStride);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;

    // row slot r = 2 * mtile + half: row-in-warp (r >> 1) * 16 + (r & 1) * 8 + g
    uint32_t A[2][8][4];                     // [mtile][k-step][a0..a3]
    int cinit[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int inWarp = (r >> 1) * 16 + (r & 1) * 8 + g, row = qBase + warp * 32 + inWarp;
        const bool valid = row < nq;
        cinit[r] = valid ? V_START + 32 - inWarp : 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t w = valid ? __ldg(Q + (long long)row * 8 + 2 * q + (t >> 1)) : 0u;
            const uint32_t half = (t & 1) ? (w >> 16) : (w & 0xFFFFu);
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                // k-step 2q + e: a0/a1 (rows g / g+8) pair with b0, a2/a3 with b1
                A[r >> 1][2 * q + e][(r & 1)] = valid ? expand4<1>(half, 2 * e) : 0u;
                A[r >> 1][2 * q + e][2 + (r & 1)] = valid ? expand4<1>(half, 2 * e + 1) : 0u;
            }
        }
    }
    uint32_t k1[4], k2[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) k1[r] = k2[r] = KEY_NONE;

    for (int c0 = 0; c0 < nt; c0 += KM_CHUNK) {
        const int cn = min(KM_CHUNK, nt - c0);
        __syncthreads();
        // expand the chunk: item e = ((q * KM_CHUNK + col) * 4 + tt) -> one 16-byte store, consecutive threads consecutive addresses
        for (int e = threadIdx.x; e < 4 * KM_CHUNK * 4; e += KM_THREADS) {
            const int q = e >> 10, col = (e >> 2) & (KM_CHUNK - 1), tt = e & 3;
            uint4 o = make_uint4(0u, 0u, 0u, 0u);
            if (col < cn) {
                const uint32_t w = __ldg(T + (long long)(c0 + col) * 8 + 2 * q + (tt >> 1));
                const uint32_t half = (tt & 1) ? (w >> 16) : (w & 0xFFFFu);
                o = make_uint4(expand4<32>(half, 0), expand4<32>(half, 1), expand4<32>(half, 2), expand4<32>(half, 3));
            }
            sB4[e] = o;
        }
        __syncthreads();

        // per chunk and row slot: packed running top-2 of each of the thread's two column streams (low half = column 2t of every
        // tile, high half = 2t + 1); the row code in the low 6 bits is replaced by 32 - tile, so ties keep the earlier column
        uint32_t b1[4] = {0u, 0u, 0u, 0u}, b2[4] = {0u, 0u, 0u, 0u};
        uint32_t tcode = 32u * 0x10001u;
        const int ntiles = (cn + 7) >> 3;
        for (int tile = 0; tile < ntiles; ++tile, tcode -= 0x10001u) {
            const int colB = tile * 8;
            uint4 b[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) b[q] = sB4[(q * KM_CHUNK + colB + g) * 4 + t];
            int acc[2][4];
            imma_first(acc[0], A[0][0], b[0].x, b[0].y, cinit[0], cinit[1]);
            imma_first(acc[1], A[1][0], b[0].x, b[0].y, cinit[2], cinit[3]);
            imma_acc(acc[0], A[0][1], b[0].z, b[0].w);
            imma_acc(acc[1], A[1][1], b[0].z, b[0].w);
#pragma unroll
            for (int q = 1; q < 4; ++q) {
                imma_acc(acc[0], A[0][2 * q], b[q].x, b[q].y);
                imma_acc(acc[1], A[1][2 * q], b[q].x, b[q].y);
                imma_acc(acc[0], A[0][2 * q + 1], b[q].z, b[q].w);
                imma_acc(acc[1], A[1][2 * q + 1], b[q].z, b[q].w);
            }
            if (colB + 8 > cn) {
                // ragged last tile: columns past nt score hamming 256 and carry an index >= nt, so every valid column beats them
                const bool v0 = colB + 2 * t < cn, v1 = colB + 2 * t + 1 < cn;
#pragma unroll
                for (int m = 0; m < 2; ++m) {
                    if (!v0) acc[m][0] = acc[m][2] = 0;
                    if (!v1) acc[m][1] = acc[m][3] = 0;
                }
            }
            uint32_t P[4];                    // row slot r: columns 2t (low half) and 2t + 1 (high half)
#pragma unroll
            for (int r = 0; r < 4; ++r) P[r] = pack16(acc[r >> 1][2 * (r & 1)], acc[r >> 1][2 * (r & 1) + 1]);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const uint32_t e = (P[r] & 0xFFC0FFC0u) | tcode;
                b2[r] = __vmaxu2(b2[r], __vminu2(b1[r], e));
                b1[r] = __vmaxu2(b1[r], e);
            }
            if (CROSS) {
                // columns: best (h, lowest query) of the warp's 32 rows; every (warp, column) is visited once, so a plain store
                uint32_t m = __vmaxu2(__vimax3_u16x2(P[0], P[1], P[2]), P[3]);
                m = __vmaxu2(m, __shfl_xor_sync(0xffffffffu, m, 4));
                m = __vmaxu2(m, __shfl_xor_sync(0xffffffffu, m, 8));
                m = __vmaxu2(m, __shfl_xor_sync(0xffffffffu, m, 16));
                if (g == 0) sCol[warp * (KM_CHUNK / 2) + tile * 4 + t] = m;
            }
        }
        // fold the chunk's candidates into the rows' (distance << 16 | trainIdx) top-2
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t e = ((k & 2) ? b2[r] : b1[r]) >> ((k & 1) * 16) & 0xFFFFu;
                if (e) {
                    const uint32_t col = (uint32_t)c0 + (32u - (e & 63u)) * 8u + 2u * t + (k & 1);
                    top2_insert(k1[r], k2[r], ((256u - (e >> V_SHIFT)) << 16) | col);
                }
            }
        }
        if (CROSS) {
            __syncthreads();
            for (int i = threadIdx.x; i < cn; i += KM_THREADS) {
                uint32_t best = 0u;           // h << 9 | (7 - warp) << 6 | rc: highest h, then lowest row
#pragma unroll
                for (int w = 0; w < KM_WARPS; ++w) {
                    const uint32_t v = (sCol[w * (KM_CHUNK / 2) + (i >> 1)] >> ((i & 1) * 16)) & 0xFFFFu;
                    if (v & 63u) best = max(best, ((v >> V_SHIFT) << 9) | ((uint32_t)(KM_WARPS - 1 - w) << V_SHIFT) | (v & 63u));
                }
                if (best) {
                    const uint32_t row = (uint32_t)qBase + (KM_WARPS - 1 - ((best >> V_SHIFT) & 7u)) * 32u + 32u - (best & 63u);
                    atomicMin(&rev[(long long)pair * K + c0 + i], (256u - (best >> 9)) * 65536u + row);
                }
            }
        }
    }
    // merge the 4 lanes of each row group; lane t == 0 writes
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
        for (int o = 1; o <= 2; o <<= 1) {
            const uint32_t o1 = __shfl_xor_sync(0xffffffffu, k1[r], o), o2 = __shfl_xor_sync(0xffffffffu, k2[r], o);
            top2_insert(k1[r], k2[r], o1);
            top2_insert(k1[r], k2[r], o2);
        }
        const int row = qBase + warp * 32 + (r >> 1) * 16 + (r & 1) * 8 + g;
        if (t == 0 && row < nq) {
            const uint32_t a = (int)(k1[r] & 0xFFFFu) < nt ? k1[r] : KEY_NONE, b = (int)(k2[r] & 0xFFFFu) < nt ? k2[r] : KEY_NONE;
            *reinterpret_cast<uint2*>(knn + ((long long)pair * K + row) * 2) = make_uint2(a, b);
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
    dim3 grid((maxNq + KM_ROWS - 1) / KM_ROWS, npairs);
    if (cross) ORBF_CUDA(c, cudaFuncSetAttribute(knn2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KM_SMEM));
    else ORBF_CUDA(c, cudaFuncSetAttribute(knn2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KM_SMEM));
    if (cross) {
        ORBF_CUDA(c, cudaMemsetAsync(ms.rev + (size_t)ms.pair0 * c->K, 0xFF, (size_t)npairs * c->K * sizeof(uint32_t), c->stream));
        knn2_kernel<true><<<grid, KM_THREADS, KM_SMEM, c->stream>>>(ms, c->K);
    } else knn2_kernel<false><<<grid, KM_THREADS, KM_SMEM, c->stream>>>(ms, c->K);
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
