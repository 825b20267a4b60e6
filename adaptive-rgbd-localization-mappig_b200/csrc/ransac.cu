// csrc/ransac.cu — RANSAC-Kabsch 3D-3D (reference Ransac::Iterate and helpers, Odometry/ransac.cpp:155-431;
// Kabsch::Compute, Odometry/kabsch.cpp:14-57).
//
// Three kernels per batch of frame pairs:
//   prepare : depth filter (ransac.cpp:175-189), std::sort replay (:199, quirk Q6), gather of the sorted 3D-3D
//             pairs into a packed array, glibc-rand sample table (SampleMatches :269-293, quirk Q5);
//   hyp     : ONE WARP PER HYPOTHESIS.  Each executed iteration of the reference depends only on its own sample,
//             so all `iterations` hypotheses run concurrently: up to 19 refit rounds of
//               T = weighted Kabsch(inliers)  — pcl::TransformationFromCorrespondences restated: the f32
//                   incremental mean/covariance recurrence is replayed in match order (lane-uniform), then a
//                   3x3 two-sided Jacobi SVD in registers, R = U diag(1,1,sign(det U det V)) V^T;
//               (err, inliers) = Mahalanobis scoring of ALL good pairs (ErrorFunction2 :350-414, f64, closed
//                   form 3x3 Cholesky), lanes striding over pairs, inlier mask by __ballot_sync, error summed in
//                   match order;
//   select  : the sequential accept / skip-ahead / early-exit rule (:233-249) replayed over the hypothesis
//             results in sample order, identity fallback (:252-264), inlier list of the winner re-scored.
// All float/double arithmetic is written in the oracle's operation order and compiled with -fmad=false, so
// inlier sets, errors and poses are bit-identical to the oracle's.
#include <cfloat>
#include <cmath>

#include "orbf_internal.h"
#include "replay.h"

namespace {

struct Pt6 { float sx, sy, sz, tx, ty, tz; float pad0, pad1; };   // 32 B: one sorted good correspondence

struct RState { float rmse; int win, nBest, n, realIters, validIters, done; };   // sequential accept-rule state of one pair

struct RansacParams {
    RansacSet rs;
    RState* state; int hypLo, hypHi, pair0;
    orbf_ransac_config cfg;
    orbf_dmatch* good; int* goodCount; Pt6* pts;
    int* samples; const int* userSamples;
    orbf_hyp_trace* hyp; orbf_ransac_result* res; orbf_dmatch* inliers;
    double* depthCov;
    int K, iters, S;
    int covReset;             // standalone call: ignore (and do not touch) the covariance latched on the context
    const float* composeIn; float* composeOut;   // one-pair Odometry::Compute: pose2 = T12 * pose1 written by the select kernel (may be NULL)
    int tabRows;              // sample-table rows ransac_prepare draws itself; ransac_table_kernel completes the table for the pairs that go on
    double covX, covY;
};


// ------------------------------------------------------------------------------------------------------
constexpr int PR_THREADS = 128;

// ---- std::sort replay, parallel over the independent sub-ranges of the introsort recursion ---------------------------
// libstdc++'s __introsort_loop partitions [first, last) around a median-of-3 pivot and then treats [first, cut) and
// [cut, last) independently, so the ranges alive at one recursion depth can be partitioned concurrently (one thread each)
// without changing a single comparison or swap inside any range.  The __final_insertion_sort that follows is a stable
// insertion sort over the whole array; since every element of an earlier leaf is <= every element of a later one, it never
// moves an element across a leaf boundary, i.e. it stably sorts each leaf (<= 16 elements) in place — done here by rank
// counting, one thread per element.  Ranges that exhaust the depth limit are heap-sorted by their thread exactly as the
// library does (already sorted, so the final pass leaves them alone).  Keys: ordered float bits << 32 | position.
struct HiLess { ORBF_HD bool operator()(unsigned long long a, unsigned long long b) const { return (a >> 32) < (b >> 32); } };
constexpr int SORT_QCAP = 128;

constexpr int SORT_THREADS = PR_THREADS - 32;     // warp 1 draws the sample table meanwhile
__device__ __forceinline__ void sort_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(SORT_THREADS) : "memory"); }

// One range of the introsort recursion partitioned by a WARP, with the result of libstdc++'s sequential __unguarded_partition: the k-th
// element (from the left) that stops the left scan — not less than the pivot — is exchanged with the k-th element (from the right) that
// stops the right scan — not greater than the pivot — for as long as the two positions have not crossed; neither scan ever looks at an
// element the other side has already moved before they cross.  So the stop positions follow from two ballot prefix sums over the
// untouched range, K = the number of pairs before the crossing, and the cut is min(left stop K, right stop K - 1) (left stop 0 when
// nothing is exchanged).  Exhaustively compared with the sequential loop on random ranges with heavy ties (authoring container).  The
// u16 position lists live in the range's own part of keys2, which is only written after the recursion.
__device__ int warp_unguarded_partition(unsigned long long* keys, unsigned long long* keys2, int first, int last, int lane)
{
    const uint32_t pv = (uint32_t)(keys[first] >> 32);
    const int lo0 = first + 1, n = last - lo0;
    uint16_t* Lpos = reinterpret_cast<uint16_t*>(keys2 + first);
    uint16_t* Rpos = Lpos + n;
    const unsigned below = (1u << lane) - 1;
    int nL = 0, nR = 0;
    for (int b = 0; b < n; b += 32) {
        const int i = lo0 + b + lane, j = last - 1 - (b + lane);
        const bool in = b + lane < n;
        const bool fl = in && !((uint32_t)(keys[i] >> 32) < pv), fr = in && !(pv < (uint32_t)(keys[j] >> 32));
        const unsigned ml = __ballot_sync(0xffffffffu, fl), mr = __ballot_sync(0xffffffffu, fr);
        if (fl) Lpos[nL + __popc(ml & below)] = (uint16_t)i;
        if (fr) Rpos[nR + __popc(mr & below)] = (uint16_t)j;
        nL += __popc(ml); nR += __popc(mr);
    }
    __syncwarp();
    const int nmin = min(nL, nR);
    int K = 0;
    for (int b = 0; b < nmin; b += 32) {
        const int k = b + lane;
        const unsigned m = __ballot_sync(0xffffffffu, k < nmin && Lpos[k] < Rpos[k]);
        K += __popc(m);
        if (m != 0xffffffffu) break;
    }
    for (int k = lane; k < K; k += 32) {
        const int p = Lpos[k], q = Rpos[k];
        const unsigned long long a = keys[p];
        keys[p] = keys[q]; keys[q] = a;
    }
    int cut;
    if (K == 0) cut = Lpos[0];
    else { cut = Rpos[K - 1]; if (K < nL && (int)Lpos[K] < cut) cut = Lpos[K]; }
    __syncwarp();
    return cut;
}

__device__ void parallel_std_sort(unsigned long long* keys, unsigned long long* keys2, uint16_t* leafStart, uint16_t* leafEnd, int M, int tid)
{
    __shared__ int qF[2][SORT_QCAP], qL[2][SORT_QCAP], qD[2][SORT_QCAP];
    __shared__ int qN[2];
    replay::IntroSort<unsigned long long, HiLess> S{ keys, HiLess() };
    for (int i = tid; i < M; i += SORT_THREADS) { leafStart[i] = (uint16_t)i; leafEnd[i] = (uint16_t)(i + 1); }
    if (tid == 0) {
        int lg = 0;
        for (int t = M; t > 1; t >>= 1) ++lg;
        qN[0] = 0; qN[1] = 0;
        if (M > 16) { qF[0][0] = 0; qL[0][0] = M; qD[0][0] = 2 * lg; qN[0] = 1; }
    }
    sort_barrier();
    if (M <= 16) {
        for (int i = tid; i < M; i += SORT_THREADS) { leafStart[i] = 0; leafEnd[i] = (uint16_t)M; }
    }
    int cur = 0;
    while (true) {
        const int n = qN[cur];
        if (n == 0) break;
        // the first levels of the recursion hold one, two, ... long ranges: a warp each (the partition of a 400-element range by one
        // thread is the kernel's critical path); later levels hold many short ranges: a thread each
        const bool byWarp = n <= SORT_THREADS / 32;
        for (int t = byWarp ? (tid >> 5) : tid; t < n; t += byWarp ? n : SORT_THREADS) {
            const int first = qF[cur][t], last = qL[cur][t];
            int depth = qD[cur][t];
            if (depth == 0) { if (!byWarp || (tid & 31) == 0) S.heap_sort(first, last); continue; }      // leaf of singletons: already in final order
            --depth;
            const int mid = first + (last - first) / 2;
            int cut;
            if (byWarp) {
                if ((tid & 31) == 0) S.move_median_to_first(first, first + 1, mid, last - 1);
                __syncwarp();
                cut = warp_unguarded_partition(keys, keys2, first, last, tid & 31);
                if ((tid & 31) != 0) continue;                       // lane 0 files the two parts
            } else {
                S.move_median_to_first(first, first + 1, mid, last - 1);
                cut = S.unguarded_partition(first + 1, last, first);
            }
            const int lo[2] = { cut, first }, hi[2] = { last, cut };
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                if (hi[c] - lo[c] > 16) {
                    const int slot = atomicAdd(&qN[cur ^ 1], 1);
                    qF[cur ^ 1][slot] = lo[c]; qL[cur ^ 1][slot] = hi[c]; qD[cur ^ 1][slot] = depth;
                } else
                    for (int i = lo[c]; i < hi[c]; ++i) { leafStart[i] = (uint16_t)lo[c]; leafEnd[i] = (uint16_t)hi[c]; }
            }
        }
        sort_barrier();
        if (tid == 0) qN[cur] = 0;
        cur ^= 1;
        sort_barrier();
    }
    sort_barrier();
    for (int i = tid; i < M; i += SORT_THREADS) {
        const int s0 = leafStart[i], e0 = leafEnd[i];
        const unsigned long long k = keys[i];
        const uint32_t ki = (uint32_t)(k >> 32);
        int rank = s0;
        for (int j = s0; j < e0; ++j) { const uint32_t kj = (uint32_t)(keys[j] >> 32); rank += (kj < ki) || (kj == ki && j < i); }
        keys2[rank] = k;
    }
    sort_barrier();
}

// ---- glibc rand() replay by one warp ---------------------------------------------------------------------------------
// glibc's TYPE_3 generator is x[i] = x[i-31] + x[i-3] (mod 2^32) with x[0..30] from the 16807 LCG, x[31..33] = x[0..2], 310
// discarded values and output x[i] >> 1 (csrc/replay.h holds the circular-buffer form the library uses).  A block of 31
// consecutive values follows from the previous block P as N[j] = sum_k P[j-3k] + P[28 + j % 3]: a stride-3 inclusive scan,
// i.e. four shuffle steps for 31 values instead of 31 dependent steps.  The warp fills a buffer of `rand() % M` values;
// the duplicate-rejecting loop of SampleMatches (ransac.cpp:269-293) then streams over it (rows_from_stream) without any
// generator or modulo arithmetic.  Returns false when the buffer was too short (caller falls back to the scalar replay).
constexpr int RAND_BUF = 3072;     // rand() values buffered per pair: 1536 draws of SampleMatches (a 200 x 4 table needs ~860)

__device__ __forceinline__ void cas(int& a, int& b) { const int lo = min(a, b), hi = max(a, b); a = lo; b = hi; }

// SampleMatches' duplicate-rejecting loop (ransac.cpp:269-293) as a streaming state machine over the id stream: the ids arrive
// by shuffle, 32 at a time, and every lane runs the same register-resident state (no divergence, no local memory, no
// dependent shared-memory loads on the critical path).  A row closes when it holds S distinct ids; they are emitted ascending.
template <int SMAX>
__device__ bool rows_from_stream(const uint16_t* pid, int np, int S, int iters, int* tab, int lane)
{
    int sel[SMAX];
#pragma unroll
    for (int q = 0; q < SMAX; ++q) sel[q] = 0x7fffffff;
    int cnt = 0, row = 0;
    for (int base = 0; base < np && row < iters; base += 32) {
        const int mine = (base + lane < np) ? (int)pid[base + lane] : 0;
        const int lim = min(32, np - base);
        for (int j = 0; j < lim && row < iters; ++j) {
            const int id = __shfl_sync(0xffffffffu, mine, j);
            bool dup = false;
#pragma unroll
            for (int q = 0; q < SMAX; ++q) dup |= sel[q] == id;          // unused slots hold INT_MAX
            if (dup) continue;
#pragma unroll
            for (int q = 0; q < SMAX; ++q) if (q == cnt) sel[q] = id;
            if (++cnt < S) continue;
            if (SMAX == 4) { cas(sel[0], sel[1]); cas(sel[2], sel[3]); cas(sel[0], sel[2]); cas(sel[1], sel[3]); cas(sel[1], sel[2]); }
            else {
#pragma unroll
                for (int pass = 0; pass < SMAX; ++pass)
#pragma unroll
                    for (int q = pass & 1; q + 1 < SMAX; q += 2) cas(sel[q], sel[q + 1]);
            }
            int v = sel[0];
#pragma unroll
            for (int q = 1; q < SMAX; ++q) if (lane == q) v = sel[q];
            if (lane < S) tab[(long long)row * S + lane] = v;
            ++row; cnt = 0;
#pragma unroll
            for (int q = 0; q < SMAX; ++q) sel[q] = 0x7fffffff;
        }
    }
    return row == iters;
}

// nRand = how many rand() values to buffer (a multiple of 2, <= RAND_BUF): a full 200 x 4 table needs ~860, its first rows a few dozen
__device__ bool warp_sample_table(uint32_t seed, int M, int S, int iters, int* tab, uint16_t* ids, int lane, int nRand = RAND_BUF)
{
    uint16_t* pid = ids + RAND_BUF;                 // [RAND_BUF / 2] min of each rand pair = the id SampleMatches draws
    uint32_t P = 0;
    {
        if (seed == 0) seed = 1;
        int32_t word = (int32_t)seed;
        for (int i = 0; i < 31; ++i) {
            if (i > 0) {
                const int32_t hi = word / 127773, lo = word % 127773;
                word = 16807 * lo - 2836 * hi;
                if (word < 0) word += 2147483647;
            }
            if (lane == (i >= 3 ? i - 3 : 28 + i)) P = (uint32_t)word;    // lane j holds x[3 + j]; x[31..33] = x[0..2]
        }
    }
    const int base = 28 + lane % 3;
    for (int blk = 0; (blk - 10) * 31 < nRand; ++blk) {
        uint32_t v = P;
#pragma unroll
        for (int off = 3; off < 32; off <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, v, off);
            if (lane >= off) v += t;
        }
        v += __shfl_sync(0xffffffffu, P, base);
        P = v;
        if (blk >= 10) {
            const int n = (blk - 10) * 31 + lane;
            if (lane < 31 && n < RAND_BUF) ids[n] = (uint16_t)((int)(v >> 1) % M);
        }
    }
    __syncwarp();
    for (int k = lane; k < nRand / 2; k += 32) pid[k] = (uint16_t)min((int)ids[2 * k], (int)ids[2 * k + 1]);
    __syncwarp();
    return S <= 4 ? rows_from_stream<4>(pid, nRand / 2, S, iters, tab, lane) : rows_from_stream<8>(pid, nRand / 2, S, iters, tab, lane);
}

__global__ void __launch_bounds__(PR_THREADS) ransac_prepare_kernel(RansacParams P)
{
    extern __shared__ __align__(16) uint8_t smem[];
    orbf_dmatch* sm = reinterpret_cast<orbf_dmatch*>(smem);     // K entries
    __shared__ int sWarp[PR_THREADS / 32];
    __shared__ int sBase;
    const int pair = P.pair0 + blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int qs = 0, ts = 0;
    if (P.rs.pairs) { qs = P.rs.pairs[2 * pair]; ts = P.rs.pairs[2 * pair + 1]; }
    const float* sz = P.rs.sz + (long long)qs * P.rs.slotStride;
    const float* tz = P.rs.tz + (long long)ts * P.rs.slotStride;
    const orbf_dmatch* m12 = P.rs.matches + (long long)pair * P.K;
    const int nm = P.rs.matchCount[pair];
    if (tid == 0) sBase = 0;
    __syncthreads();
    if ((unsigned)nm >= P.cfg.min_inlier_th) {
        for (int base = 0; base < nm; base += PR_THREADS) {
            const int i = base + tid;
            bool keep = false;
            orbf_dmatch m;
            if (i < nm) {
                m = m12[i];
                keep = true;
                if (P.cfg.check_depth) {
                    const float a = sz[m.queryIdx], b = tz[m.trainIdx];
                    if (isnan(a) || isnan(b) || a <= 0 || b <= 0) keep = false;
                }
            }
            const unsigned mk = __ballot_sync(0xffffffffu, keep);
            if (lane == 0) sWarp[warp] = __popc(mk);
            __syncthreads();
            int off = sBase;
            for (int w = 0; w < warp; ++w) off += sWarp[w];
            if (keep) sm[off + __popc(mk & ((1u << lane) - 1))] = m;
            __syncthreads();
            if (tid == 0) { int t = 0; for (int w = 0; w < PR_THREADS / 32; ++w) t += sWarp[w]; sBase += t; }
            __syncthreads();
        }
    }
    const int M = sBase;
    orbf_dmatch* good = P.good + (long long)pair * P.K;
    // sample table (needs only M): warp 1 draws it while the other three warps replay std::sort
    int* tab = P.samples + (long long)pair * P.iters * P.S;
    if (P.userSamples)
        for (int i = tid; i < P.iters * P.S; i += PR_THREADS) tab[i] = P.userSamples[i];
    __syncthreads();
    if (warp == 1) {
        if (!P.userSamples) {
            uint16_t* ids = reinterpret_cast<uint16_t*>(smem + (size_t)P.K * (sizeof(orbf_dmatch) + 2 * sizeof(unsigned long long) + 2 * sizeof(uint16_t)));
            // the first tabRows rows only (the loop rarely gets further); ransac_table_kernel draws the rest for pairs that go on
            const int rows = P.tabRows, nRand = rows < P.iters ? min(RAND_BUF, 64 * rows) : RAND_BUF;
            bool done = false;
            if (M >= P.S) done = warp_sample_table(P.cfg.seed + (uint32_t)pair, M, P.S, rows, tab, ids, lane, nRand);
            if (!done && lane == 0) {
                if (M >= P.S) {     // buffer exhausted (tiny M, many duplicate draws): scalar replay from the start
                    replay::GlibcRand g;
                    g.seed(P.cfg.seed + (uint32_t)pair);
                    for (int k = 0; k < rows; ++k) replay::sample_row(g, M, P.S, tab + (long long)k * P.S);
                } else for (int i = 0; i < P.iters * P.S; ++i) tab[i] = -1;
            }
        }
    } else {
        const int st = tid < 32 ? tid : tid - 32;                              // 0..SORT_THREADS-1
        if (P.cfg.sort_mode == 0) {
            unsigned long long* keys = reinterpret_cast<unsigned long long*>(sm + P.K);
            unsigned long long* keys2 = keys + P.K;
            uint16_t* leafStart = reinterpret_cast<uint16_t*>(keys2 + P.K);
            uint16_t* leafEnd = leafStart + P.K;
            for (int i = st; i < M; i += SORT_THREADS) {
                uint32_t b = __float_as_uint(sm[i].distance);
                b ^= (b >> 31) ? 0xFFFFFFFFu : 0x80000000u;                 // float order -> unsigned order
                keys[i] = ((unsigned long long)b << 32) | (unsigned)i;
            }
            sort_barrier();
            parallel_std_sort(keys, keys2, leafStart, leafEnd, M, st);
            for (int i = st; i < M; i += SORT_THREADS) good[i] = sm[(uint32_t)keys2[i]];
        } else if (P.cfg.sort_mode == 2) {   // stable: rank = #{j : d_j < d_i or (d_j == d_i and j < i)}
            for (int i = st; i < M; i += SORT_THREADS) {
                const float d = sm[i].distance;
                int r = 0;
                for (int j = 0; j < M; ++j) { const float e = sm[j].distance; r += (e < d) || (e == d && j < i); }
                good[r] = sm[i];
            }
        } else for (int i = st; i < M; i += SORT_THREADS) good[i] = sm[i];
    }
    __syncthreads();
    // packed, sorted 3D-3D pairs
    const float* sx = P.rs.sx + (long long)qs * P.rs.slotStride; const float* sy = P.rs.sy + (long long)qs * P.rs.slotStride;
    const float* tx = P.rs.tx + (long long)ts * P.rs.slotStride; const float* ty = P.rs.ty + (long long)ts * P.rs.slotStride;
    Pt6* pts = P.pts + (long long)pair * P.K;
    for (int i = tid; i < M; i += PR_THREADS) {
        const orbf_dmatch m = good[i];
        Pt6 p;
        p.sx = sx[m.queryIdx]; p.sy = sy[m.queryIdx]; p.sz = sz[m.queryIdx];
        p.tx = tx[m.trainIdx]; p.ty = ty[m.trainIdx]; p.tz = tz[m.trainIdx];
        p.pad0 = p.pad1 = 0.f;
        pts[i] = p;
    }
    orbf_hyp_trace* tr = P.hyp + (long long)pair * P.iters;
    for (int k = tid; k < P.iters; k += PR_THREADS) { tr[k].n_refined = 0; tr[k].rounds = -1; tr[k].refined_error = 1e6; }
    if (tid == 0) {
        P.goodCount[pair] = M;
        RState st;
        st.rmse = 1e6f; st.win = -1; st.nBest = 0; st.n = 0; st.realIters = 0; st.validIters = 0;
        st.done = ((unsigned)nm >= P.cfg.min_inlier_th && (unsigned)M >= P.cfg.min_inlier_th && M >= P.S && M <= 2048) ? 0 : 1;
        P.state[pair] = st;
    }
}

// depth covariance latch (quirk Q7: static local initialised by the first DepthCovariance() call of the process)
// The whole sample table for the pairs whose loop is still running after the first tabRows hypotheses (one warp per pair; the
// rows prepare already drew come out identical: same seed, same stream).
__global__ void __launch_bounds__(32) ransac_table_kernel(RansacParams P)
{
    __shared__ uint16_t ids[RAND_BUF + RAND_BUF / 2];
    const int pair = P.pair0 + blockIdx.x, lane = threadIdx.x;
    if (P.state[pair].done) return;
    const int M = P.goodCount[pair];
    int* tab = P.samples + (long long)pair * P.iters * P.S;
    bool done = false;
    if (M >= P.S) done = warp_sample_table(P.cfg.seed + (uint32_t)pair, M, P.S, P.iters, tab, ids, lane);
    if (!done && lane == 0 && M >= P.S) {
        replay::GlibcRand g;
        g.seed(P.cfg.seed + (uint32_t)pair);
        for (int k = 0; k < P.iters; ++k) replay::sample_row(g, M, P.S, tab + (long long)k * P.S);
    }
}

__global__ void ransac_latch_kernel(RansacParams P, int npairs)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double cz = P.cfg.depth_cov;
    if (!(cz >= 0.0)) {
        cz = P.covReset ? -1.0 : *P.depthCov;   // value latched by an earlier call on this context (negative if none)
        // The reference latches on the first ErrorFunction2 call of the process, i.e. in the first pair (in call order) that
        // gets as far as scoring — Iterate returns early below minInlierTh matches / good matches (ransac.cpp:165,191) — at
        // the first correspondence ComputeInliersAndError does not skip (:326) and ErrorFunction2 does not reject as NaN (:362).
        for (int q = 0; q < npairs && !(cz >= 0.0); ++q) {
            const int pair = P.pair0 + q;
            const int M = P.goodCount[pair];
            if ((unsigned)P.rs.matchCount[pair] < P.cfg.min_inlier_th || (unsigned)M < P.cfg.min_inlier_th || M > 2048) continue;
            const Pt6* pts = P.pts + (long long)pair * P.K;
            for (int i = 0; i < M; ++i) {
                const Pt6 p = pts[i];
                if (p.sz == 0.0f || p.tx == 0.0f) continue;
                if (isnan(p.sz) || isnan(p.tz)) continue;
                const double sd = 0.01 * (double)p.sz * (double)p.sz;
                cz = sd * sd;
                break;
            }
        }
    }
    *P.depthCov = cz;
}

// ---- 3x3 SVD (two-sided Jacobi), same operation order as the oracle ---------------------------------------
struct M3 { float m[3][3]; };

__device__ __forceinline__ void rot_rows(M3& a, int p, int q, float c, float s)
{
#pragma unroll
    for (int j = 0; j < 3; ++j) { const float x = a.m[p][j], y = a.m[q][j]; a.m[p][j] = c * x + s * y; a.m[q][j] = c * y - s * x; }
}
__device__ __forceinline__ void rot_cols(M3& a, int p, int q, float c, float s)
{
#pragma unroll
    for (int i = 0; i < 3; ++i) { const float x = a.m[i][p], y = a.m[i][q]; a.m[i][p] = c * x - s * y; a.m[i][q] = s * x + c * y; }
}

__device__ void svd3(const float* A, M3& U, float S[3], M3& V)
{
    M3 M;
    float scale = 0.f;
#pragma unroll
    for (int i = 0; i < 9; ++i) scale = fmaxf(scale, fabsf(A[i]));
    if (!(scale > 0.f)) scale = 1.f;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) { M.m[i][j] = A[3 * i + j] / scale; U.m[i][j] = V.m[i][j] = (i == j) ? 1.f : 0.f; }
    const float precision = 2.f * FLT_EPSILON, tiny = FLT_MIN;
    float maxDiag = fmaxf(fabsf(M.m[0][0]), fmaxf(fabsf(M.m[1][1]), fabsf(M.m[2][2])));
    for (int sweep = 0; sweep < 30; ++sweep) {
        bool finished = true;
#pragma unroll
        for (int p = 1; p < 3; ++p)
#pragma unroll
            for (int q = 0; q < p; ++q) {
                const float thr = fmaxf(tiny, precision * maxDiag);
                if (!(fabsf(M.m[p][q]) > thr || fabsf(M.m[q][p]) > thr)) continue;
                finished = false;
                const float m00 = M.m[p][p], m01 = M.m[p][q], m10 = M.m[q][p], m11 = M.m[q][q];
                float c1 = 1.f, s1 = 0.f;
                const float t = m00 + m11, d = m10 - m01;
                if (fabsf(d) >= tiny) {
                    const float u = t / d;
                    const float tmp = sqrtf(1.f + u * u);
                    s1 = 1.f / tmp;
                    c1 = u / tmp;
                }
                const float x = c1 * m00 + s1 * m10;
                const float y = c1 * m01 + s1 * m11;
                const float z = c1 * m11 - s1 * m01;
                float c2 = 1.f, s2 = 0.f;
                if (fabsf(y) >= tiny) {
                    const float tau = (x - z) / (2.f * y);
                    const float w = sqrtf(tau * tau + 1.f);
                    const float tt = (tau > 0.f) ? -1.f / (tau + w) : -1.f / (tau - w);
                    c2 = 1.f / sqrtf(tt * tt + 1.f);
                    s2 = tt * c2;
                }
                const float cL = c1 * c2 + s1 * s2, sL = s1 * c2 - c1 * s2;
                rot_rows(M, p, q, cL, sL);
                rot_cols(M, p, q, c2, s2);
                rot_cols(U, p, q, cL, -sL);
                rot_cols(V, p, q, c2, s2);
                maxDiag = fmaxf(maxDiag, fmaxf(fabsf(M.m[p][p]), fabsf(M.m[q][q])));
            }
        if (finished) break;
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float a = M.m[i][i];
        S[i] = fabsf(a);
        if (a < 0.f) for (int r = 0; r < 3; ++r) U.m[r][i] = -U.m[r][i];
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        int k = i;
        for (int j = i + 1; j < 3; ++j) if (S[j] > S[k]) k = j;
        if (k != i) {
            float t = S[i]; S[i] = S[k]; S[k] = t;
            for (int r = 0; r < 3; ++r) {
                t = U.m[r][i]; U.m[r][i] = U.m[r][k]; U.m[r][k] = t;
                t = V.m[r][i]; V.m[r][i] = V.m[r][k]; V.m[r][k] = t;
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) S[i] *= scale;
}

__device__ __forceinline__ float det3(const M3& a)
{
    return a.m[0][0] * (a.m[1][1] * a.m[2][2] - a.m[1][2] * a.m[2][1]) - a.m[0][1] * (a.m[1][0] * a.m[2][2] - a.m[1][2] * a.m[2][0])
        + a.m[0][2] * (a.m[1][0] * a.m[2][1] - a.m[1][1] * a.m[2][0]);
}

struct Tfc {   // pcl::TransformationFromCorrespondences
    float accW, m1[3], m2[3], C[3][3];
    // The recurrence as PCL writes it (one correspondence at a time).  tfc_from_mask runs exactly these operations, split into
    // nine per-lane chains; the scalar form is kept (not compiled) as the statement of what they compute.
#if 0
    __device__ void add(const float p[3], const float q[3], float w)
    {
        if (w == 0.0f) return;
        accW += w;
        const float alpha = w / accW;
        float d1[3], d2[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) { d1[i] = p[i] - m1[i]; d2[i] = q[i] - m2[i]; }
        const float oma = 1.0f - alpha;
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) { const float outer = d2[r] * d1[c]; C[r][c] = oma * (C[r][c] + alpha * outer); }
#pragma unroll
        for (int i = 0; i < 3; ++i) { m1[i] += alpha * d1[i]; m2[i] += alpha * d2[i]; }
    }
#endif
    __device__ void transform(float* T) const
    {
        M3 U, V; float S[3];
        svd3(&C[0][0], U, S, V);
        const float s22 = (det3(U) * det3(V) < 0.0f) ? -1.0f : 1.0f;
        float R[3][3];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                const float us = U.m[i][2] * s22;
                R[i][j] = (U.m[i][0] * V.m[j][0] + U.m[i][1] * V.m[j][1]) + us * V.m[j][2];
            }
        for (int i = 0; i < 3; ++i) {
            const float rm = (R[i][0] * m1[0] + R[i][1] * m1[1]) + R[i][2] * m1[2];
            T[4 * i + 0] = R[i][0]; T[4 * i + 1] = R[i][1]; T[4 * i + 2] = R[i][2];
            T[4 * i + 3] = m2[i] - rm;
        }
        T[12] = 0; T[13] = 0; T[14] = 0; T[15] = 1;
    }
};

// GetTransformFromMatches (ransac.cpp:295-313) for the inlier set in mask[], by one warp, bit-identical to feeding the
// points to Tfc::add in match order:
//   (a) the inliers are compacted in order (ballot prefix) together with their weights w = 1 / (z_from * z_to);
//   (b) accW is the sequential float prefix sum of w (one dependent FADD per point — the only truly serial part), and
//       alpha_k = w_k / accW_k is then formed for 32 points at a time;
//   (c) the recurrence C = (1-alpha)(C + alpha d2 d1^T), m += alpha d splits into independent scalar chains: lane (r, c)
//       owns C[r][c] plus private copies of m1[c] and m2[r], so a point costs ~10 FP ops per lane instead of ~60 per warp.
// idx / al: per-warp shared scratch of at least M entries.
__device__ void tfc_from_mask(const Pt6* pts, const uint32_t* mask, int words, uint16_t* idx, float* al, Tfc& out, int lane)
{
    const uint32_t lt = (1u << lane) - 1;
    int n = 0;
    for (int w = 0; w < words; ++w) {
        const uint32_t mk = mask[w];
        const int i = (w << 5) + lane;
        bool use = (mk >> lane) & 1u;
        float wt = 0.f;
        if (use) {
            const Pt6 p = pts[i];
            if (isnan(p.sz) || isnan(p.tz)) use = false;
            else { wt = 1.0f / (p.sz * p.tz); if (wt == 0.0f) use = false; }
        }
        const uint32_t um = __ballot_sync(0xffffffffu, use);
        if (use) { const int pos = n + __popc(um & lt); idx[pos] = (uint16_t)i; al[pos] = wt; }
        n += __popc(um);
    }
    __syncwarp();
    float acc = 0.f;
    for (int base = 0; base < n; base += 32) {
        const int cnt = min(32, n - base);
        float myAcc = 1.f;
        for (int j = 0; j < cnt; ++j) { acc += al[base + j]; if (lane == j) myAcc = acc; }
        __syncwarp();
        if (lane < cnt) al[base + lane] = al[base + lane] / myAcc;
        __syncwarp();
    }
    const int r = (lane % 9) / 3, c = lane % 3;              // lanes 9..31 repeat the nine chains
    float Cv = 0.f, m1c = 0.f, m2r = 0.f;
    for (int k = 0; k < n; ++k) {
        const float alpha = al[k], oma = 1.0f - alpha;
        const float* pf = reinterpret_cast<const float*>(pts + idx[k]);
        const float d1 = pf[c] - m1c, d2 = pf[3 + r] - m2r;
        const float outer = d2 * d1;
        Cv = oma * (Cv + alpha * outer);
        m1c += alpha * d1;
        m2r += alpha * d2;
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        out.m1[i] = __shfl_sync(0xffffffffu, m1c, i);
        out.m2[i] = __shfl_sync(0xffffffffu, m2r, 3 * i);
#pragma unroll
        for (int j = 0; j < 3; ++j) out.C[i][j] = __shfl_sync(0xffffffffu, Cv, 3 * i + j);
    }
    out.accW = acc;
    __syncwarp();
}

// ErrorFunction2 (ransac.cpp:350-414); cz = depth covariance (explicit, quirk Q7)
__device__ double mahal2(const Pt6& p, const double* T, double cz, double covX, double covY)
{
    const double dmax = DBL_MAX;
    if (isnan(p.sz) || isnan(p.tz)) return dmax;
    const double a[3] = { (double)p.sx, (double)p.sy, (double)p.sz }, b[3] = { (double)p.tx, (double)p.ty, (double)p.tz };
    double dl[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double mu = ((T[4 * i] * a[0] + T[4 * i + 1] * a[1]) + T[4 * i + 2] * a[2]) + T[4 * i + 3];
        dl[i] = mu - b[i];
    }
    {
        const double sq = (dl[0] * dl[0] + dl[1] * dl[1]) + dl[2] * dl[2];
        const double s1 = fmax(covX, cz), s2 = fmax(covX, cz);
        if (sq > 2.0 * (s1 + s2)) return dmax;
    }
    const double c1[3] = { covX * a[2], covY * a[2], cz };
    const double c2[3] = { covX * b[2], covY * b[2], cz };
    double S[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const double v = ((T[i] * c1[0]) * T[j] + (T[4 + i] * c1[1]) * T[4 + j]) + (T[8 + i] * c1[2]) * T[8 + j];
            S[i][j] = v + ((i == j) ? c2[i] : 0.0);
        }
    if (isnan(dl[2])) return dmax;
    double L[3][3] = { { 0, 0, 0 }, { 0, 0, 0 }, { 0, 0, 0 } };
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        double x = S[k][k];
        for (int j = 0; j < k; ++j) x -= L[k][j] * L[k][j];
        if (!(x > 0.0)) return dmax;
        const double lkk = sqrt(x);
        L[k][k] = lkk;
        for (int i = k + 1; i < 3; ++i) {
            double v = S[i][k];
            for (int j = 0; j < k; ++j) v -= L[i][j] * L[k][j];
            L[i][k] = v / lkk;
        }
    }
    double y[3], xs[3];
    y[0] = dl[0] / L[0][0];
    y[1] = (dl[1] - L[1][0] * y[0]) / L[1][1];
    y[2] = ((dl[2] - L[2][0] * y[0]) - L[2][1] * y[1]) / L[2][2];
    xs[2] = y[2] / L[2][2];
    xs[1] = (y[1] - L[2][1] * xs[2]) / L[1][1];
    xs[0] = ((y[0] - L[1][0] * xs[1]) - L[2][0] * xs[2]) / L[0][0];
    const double d2 = (dl[0] * xs[0] + dl[1] * xs[1]) + dl[2] * xs[2];
    if (!(d2 >= 0.0)) return dmax;
    return d2;
}

constexpr int HY_WARPS = 4;
constexpr int MAX_WORDS = 64;   // supports up to 2048 good matches per pair

// ComputeInliersAndError (ransac.cpp:315-348) by one warp: fills mask[] (bit i of word i/32), returns error
__device__ double score_all(const Pt6* pts, int M, const float* T4f, double cz, const RansacParams& P, uint32_t* mask, int& nInl)
{
    const int lane = threadIdx.x & 31;
    double T[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) T[i] = (double)T4f[i];
    const double thr = (double)(P.cfg.max_mahal * P.cfg.max_mahal);
    double mean = 0.0;
    int cnt = 0;
    for (int b = 0; b < M; b += 32) {
        const int i = b + lane;
        bool in = false;
        double d = 0.0;
        if (i < M) {
            const Pt6 p = pts[i];
            if (!(p.sz == 0.0f || p.tx == 0.0f)) {            // sic: target.x (quirk Q8)
                d = mahal2(p, T, cz, P.covX, P.covY);
                in = !(d > thr) && (d >= 0.0);
            }
        }
        unsigned mk = __ballot_sync(0xffffffffu, in);
        if (lane == 0) mask[b >> 5] = mk;
        cnt += __popc(mk);
        while (mk) {                                           // meanError += mahalDist, in match order
            const int src = __ffs(mk) - 1;
            mk &= mk - 1;
            mean += __shfl_sync(0xffffffffu, d, src);
        }
    }
    __syncwarp();
    nInl = cnt;
    if (cnt < 3) return 1e9;
    mean /= (double)cnt;
    return sqrt(mean);
}

__global__ void __launch_bounds__(HY_WARPS * 32) ransac_hyp_kernel(RansacParams P)
{
    __shared__ uint32_t sMask[HY_WARPS][MAX_WORDS];
    extern __shared__ __align__(16) uint8_t hsmem[];             // per warp: K alpha floats + K inlier indices (tfc_from_mask)
    float* sAl = reinterpret_cast<float*>(hsmem);
    uint16_t* sIdx = reinterpret_cast<uint16_t*>(sAl + HY_WARPS * P.K);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int pair = P.pair0 + blockIdx.y;
    const int k = P.hypLo + blockIdx.x * HY_WARPS + warp;
    if (k >= P.hypHi) return;
    if (P.state[pair].done) return;          // the sequential loop already ended before this wave (early exit / skip-ahead)
    const int M = P.goodCount[pair];
    orbf_hyp_trace* tr = P.hyp + (long long)pair * P.iters + k;
    const float I16[16] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1 };
    float refT[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) refT[i] = I16[i];
    int nRefined = 0, rounds = -1;
    double refErr = 1e6;
    if ((unsigned)M >= P.cfg.min_inlier_th && M >= P.S && M <= MAX_WORDS * 32) {
        const Pt6* pts = P.pts + (long long)pair * P.K;
        const int* row = P.samples + ((long long)pair * P.iters + k) * P.S;
        const double cz = *P.depthCov;
        uint32_t* mask = sMask[warp];
        const int words = (M + 31) >> 5;
        for (int w = lane; w < words; w += 32) mask[w] = 0;
        __syncwarp();
        if (lane == 0) for (int s = 0; s < P.S; ++s) { const int id = row[s]; if (id >= 0 && id < M) mask[id >> 5] |= 1u << (id & 31); }
        __syncwarp();
        rounds = 0;
        for (int refinements = 1; refinements < 20; ++refinements) {
            Tfc tfc;
            tfc_from_mask(pts, mask, words, sIdx + warp * P.K, sAl + warp * P.K, tfc, lane);
            float T[16];
            tfc.transform(T);
            ++rounds;
            __syncwarp();
            int nInl;
            const double err = score_all(pts, M, T, cz, P, mask, nInl);
            if ((unsigned)nInl < P.cfg.min_inlier_th || err > (double)P.cfg.max_mahal) break;
            if (nInl >= nRefined && err <= refErr) {
                const int prev = nRefined;
#pragma unroll
                for (int i = 0; i < 16; ++i) refT[i] = T[i];
                nRefined = nInl;
                refErr = err;
                if (nInl == prev) break;
            } else break;
        }
    }
    if (lane == 0) {
        tr->n_refined = nRefined; tr->rounds = rounds; tr->refined_error = refErr;
        for (int i = 0; i < 16; ++i) tr->T[i] = refT[i];
    }
}

// The same hypothesis by a whole CTA (4 warps), for the first waves: they hold only a few hypotheses per pair (2, then 6), so a warp per
// hypothesis leaves the GPU idle and the refit loop's latency is what the stage costs (on the bench sequence 507 of 511 pairs end in
// the first wave).  Warp 0 fits the transform (the PCL recurrence and its accW prefix are sequential by definition); ALL warps then
// score the pairs (f64 Mahalanobis, independent per pair); the inliers' distances are compacted in match order and summed by one
// thread from shared memory — the reference's sequential double accumulation, 8 loads in flight per add instead of two shuffles per
// add.  Bit-identical to ransac_hyp_kernel (same operations in the same order; tests/test_gpu_ransac.py compares the traces).
constexpr int HC_WARPS = 4, HC_THREADS = HC_WARPS * 32;

__global__ void __launch_bounds__(HC_THREADS) ransac_hyp_coop_kernel(RansacParams P)
{
    __shared__ uint32_t sMask[MAX_WORDS];
    __shared__ int sBase[MAX_WORDS + 1];
    __shared__ float sT[16];
    __shared__ double sErr;
    extern __shared__ __align__(16) uint8_t hsmem[];             // K doubles (distances by match index), K doubles (compacted), K floats, K u16
    double* sD = reinterpret_cast<double*>(hsmem);
    double* sDense = sD + P.K;
    float* sAl = reinterpret_cast<float*>(sDense + P.K);
    uint16_t* sIdx = reinterpret_cast<uint16_t*>(sAl + P.K);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int pair = P.pair0 + blockIdx.y;
    const int k = P.hypLo + blockIdx.x;
    if (k >= P.hypHi) return;
    if (P.state[pair].done) return;          // the sequential loop already ended before this wave (early exit / skip-ahead)
    const int M = P.goodCount[pair];
    orbf_hyp_trace* tr = P.hyp + (long long)pair * P.iters + k;
    float refT[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) refT[i] = (i % 5 == 0) ? 1.f : 0.f;
    int nRefined = 0, rounds = -1;
    double refErr = 1e6;
    if ((unsigned)M >= P.cfg.min_inlier_th && M >= P.S && M <= MAX_WORDS * 32) {
        const Pt6* pts = P.pts + (long long)pair * P.K;
        const int* row = P.samples + ((long long)pair * P.iters + k) * P.S;
        const double cz = *P.depthCov;
        const double thr = (double)(P.cfg.max_mahal * P.cfg.max_mahal);
        const int words = (M + 31) >> 5;
        for (int w = tid; w < words; w += HC_THREADS) sMask[w] = 0;
        __syncthreads();
        if (tid == 0) for (int s = 0; s < P.S; ++s) { const int id = row[s]; if (id >= 0 && id < M) sMask[id >> 5] |= 1u << (id & 31); }
        __syncthreads();
        rounds = 0;
        for (int refinements = 1; refinements < 20; ++refinements) {
            if (warp == 0) {
                Tfc tfc;
                tfc_from_mask(pts, sMask, words, sIdx, sAl, tfc, lane);
                float T[16];
                tfc.transform(T);
                if (lane == 0) for (int i = 0; i < 16; ++i) sT[i] = T[i];
            }
            ++rounds;
            __syncthreads();
            float T4[16];
            double T[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) { T4[i] = sT[i]; T[i] = (double)T4[i]; }
            // ComputeInliersAndError (ransac.cpp:315-348): chunks of 32 matches per warp, so a warp's ballot is one word of the mask
            for (int b = warp * 32; b < M; b += HC_THREADS) {
                const int i = b + lane;
                bool in = false;
                double d = 0.0;
                if (i < M) {
                    const Pt6 p = pts[i];
                    if (!(p.sz == 0.0f || p.tx == 0.0f)) {            // sic: target.x (quirk Q8)
                        d = mahal2(p, T, cz, P.covX, P.covY);
                        in = !(d > thr) && (d >= 0.0);
                    }
                    sD[i] = d;
                }
                const unsigned mk = __ballot_sync(0xffffffffu, in);
                if (lane == 0) sMask[b >> 5] = mk;
            }
            __syncthreads();
            if (warp == 0) {                                       // exclusive prefix of the words' inlier counts
                int carry = 0;
                for (int w0 = 0; w0 < words; w0 += 32) {
                    const int w = w0 + lane;
                    const int c = w < words ? __popc(sMask[w]) : 0;
                    int incl = c;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
                    if (w < words) sBase[w] = carry + incl - c;
                    carry += __shfl_sync(0xffffffffu, incl, 31);
                }
                if (lane == 0) sBase[words] = carry;
            }
            __syncthreads();
            for (int b = warp * 32; b < M; b += HC_THREADS) {      // inlier distances, compacted in match order
                const uint32_t mk = sMask[b >> 5];
                if ((mk >> lane) & 1u) sDense[sBase[b >> 5] + __popc(mk & ((1u << lane) - 1u))] = sD[b + lane];
            }
            __syncthreads();
            const int nInl = sBase[words];
            if (tid == 0) {                                        // meanError += mahalDist, in match order (sequential double adds)
                double mean = 0.0;
                int j = 0;
                for (; j + 8 <= nInl; j += 8) {
                    const double v0 = sDense[j], v1 = sDense[j + 1], v2 = sDense[j + 2], v3 = sDense[j + 3], v4 = sDense[j + 4], v5 = sDense[j + 5],
                        v6 = sDense[j + 6], v7 = sDense[j + 7];
                    mean += v0; mean += v1; mean += v2; mean += v3; mean += v4; mean += v5; mean += v6; mean += v7;
                }
                for (; j < nInl; ++j) mean += sDense[j];
                double e = 1e9;
                if (nInl >= 3) { mean /= (double)nInl; e = sqrt(mean); }
                sErr = e;
            }
            __syncthreads();
            const double err = sErr;
            if ((unsigned)nInl < P.cfg.min_inlier_th || err > (double)P.cfg.max_mahal) break;
            if (nInl >= nRefined && err <= refErr) {
                const int prev = nRefined;
#pragma unroll
                for (int i = 0; i < 16; ++i) refT[i] = T4[i];
                nRefined = nInl;
                refErr = err;
                if (nInl == prev) break;
            } else break;
        }
    }
    if (tid == 0) {
        tr->n_refined = nRefined; tr->rounds = rounds; tr->refined_error = refErr;
        for (int i = 0; i < 16; ++i) tr->T[i] = refT[i];
    }
}

// The reference's accept / skip-ahead / early-exit rule (ransac.cpp:233-249), continued over the hypotheses of the
// wave that just finished: one thread per pair, hypotheses consumed strictly in sample order.
__global__ void ransac_replay_kernel(RansacParams P, int npairs)
{
    if (blockIdx.x * blockDim.x + threadIdx.x >= npairs) return;
    const int pair = P.pair0 + blockIdx.x * blockDim.x + threadIdx.x;
    RState st = P.state[pair];
    if (st.done) return;
    const int M = P.goodCount[pair];
    const unsigned minInl = P.cfg.min_inlier_th;
    const orbf_hyp_trace* hyp = P.hyp + (long long)pair * P.iters;
    while (st.n < P.iters && st.realIters < P.hypHi) {
        const orbf_hyp_trace h = hyp[st.realIters];
        st.realIters++;
        bool brk = false;
        if (h.n_refined > 0) {
            st.validIters++;
            if (h.refined_error <= (double)st.rmse && h.n_refined >= st.nBest && (unsigned)h.n_refined >= minInl) {
                st.rmse = (float)h.refined_error;
                st.win = st.realIters - 1;
                st.nBest = h.n_refined;
                if ((double)h.n_refined > (double)M * 0.5) st.n += 10;
                if ((double)h.n_refined > (double)M * 0.75) st.n += 10;
                if ((double)h.n_refined > (double)M * 0.8) brk = true;
            }
        }
        if (brk) { st.done = 1; break; }
        st.n++;
    }
    if (st.n >= P.iters) st.done = 1;
    P.state[pair] = st;
}

// One CTA (4 warps) per pair.  The winner's inlier set only needs the pass / fail flag of every pair (its error is known from the
// hypothesis trace), so all four warps score; the identity fallback (no valid hypothesis at all, ransac.cpp:252-262) needs the ordered
// error sum as well and is left to warp 0.
constexpr int SEL_THREADS = 128;
__global__ void __launch_bounds__(SEL_THREADS) ransac_select_kernel(RansacParams P)
{
    __shared__ uint32_t sMask[MAX_WORDS];
    __shared__ int sIdentity;
    __shared__ double sErr;
    const int pair = P.pair0 + blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int M = P.goodCount[pair];
    const int nm = P.rs.matchCount[pair];
    orbf_ransac_result* res = P.res + pair;
    const orbf_hyp_trace* hyp = P.hyp + (long long)pair * P.iters;
    const orbf_dmatch* good = P.good + (long long)pair * P.K;
    const Pt6* pts = P.pts + (long long)pair * P.K;
    orbf_dmatch* inl = P.inliers + (long long)pair * P.K;
    const double cz = *P.depthCov;
    const unsigned minInl = P.cfg.min_inlier_th;
    const RState st = P.state[pair];
    float rmse = st.rmse;
    const int win = st.win, realIters = st.realIters;
    int validIters = st.validIters, usedIdentity = 0;
    const bool runnable = (unsigned)nm >= minInl && (unsigned)M >= minInl && M <= MAX_WORDS * 32;
    float T[16] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1 };
    bool haveMask = false;
    if (runnable && win >= 0) {                                // vRefinedMatches of the winner, recomputed (flags only)
        double Td[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) { T[i] = hyp[win].T[i]; Td[i] = (double)T[i]; }
        const double thr = (double)(P.cfg.max_mahal * P.cfg.max_mahal);
        for (int b = warp * 32; b < M; b += SEL_THREADS) {
            const int i = b + lane;
            bool in = false;
            if (i < M) {
                const Pt6 p = pts[i];
                if (!(p.sz == 0.0f || p.tx == 0.0f)) {          // sic: target.x (quirk Q8)
                    const double d = mahal2(p, Td, cz, P.covX, P.covY);
                    in = !(d > thr) && (d >= 0.0);
                }
            }
            const unsigned mk = __ballot_sync(0xffffffffu, in);
            if (lane == 0) sMask[b >> 5] = mk;
        }
        haveMask = true;
    } else if (runnable && validIters == 0) {
        if (warp == 0) {
            int nInl = 0;
            const double err = score_all(pts, M, T, cz, P, sMask, nInl);
            if (lane == 0) { sIdentity = ((unsigned)nInl > minInl && err < (double)P.cfg.max_mahal) ? 1 : 0; sErr = err; }
        }
        __syncthreads();
        if (sIdentity) {
            haveMask = true; usedIdentity = 1;
            rmse = (float)((double)rmse + sErr);
            validIters = 1;
        }
    }
    __syncthreads();
    if (warp != 0) return;
    int outCount = 0;
    if (haveMask) {
        for (int b = 0; b < M; b += 32) {
            const uint32_t mk = sMask[b >> 5];
            const int i = b + lane;
            if (mk & (1u << lane)) inl[outCount + __popc(mk & ((1u << lane) - 1))] = good[i];
            outCount += __popc(mk);
        }
    }
    if (lane == 0) {
        orbf_ransac_result r;
        r.ok = ((unsigned)outCount >= minInl) ? 1 : 0;
        r.rmse = rmse;
        for (int i = 0; i < 16; ++i) r.T12[i] = T[i];
        r.n_inliers = outCount; r.n_good = ((unsigned)nm >= minInl) ? M : 0;
        r.real_iters = realIters; r.valid_iters = validIters; r.used_identity = usedIdentity;
        r.depth_cov_used = cz;
        *res = r;
    }
    // Odometry::Compute's composition rule for a one-pair call (odometry.cpp:82-84): cv::Mat's 4x4 float product, every element the four
    // products summed left to right — the arithmetic of compose_kernel, without its launch
    if (P.composeOut && blockIdx.x == 0 && lane < 16) {
        const int rr = lane >> 2, cc = lane & 3;
        const float* B = P.composeIn;
        float a0 = T[0], a1 = T[1], a2 = T[2], a3 = T[3];
        if (rr == 1) { a0 = T[4]; a1 = T[5]; a2 = T[6]; a3 = T[7]; }
        if (rr == 2) { a0 = T[8]; a1 = T[9]; a2 = T[10]; a3 = T[11]; }
        if (rr == 3) { a0 = T[12]; a1 = T[13]; a2 = T[14]; a3 = T[15]; }
        float t = __fmul_rn(a0, B[cc]);
        t = __fadd_rn(t, __fmul_rn(a1, B[4 + cc]));
        t = __fadd_rn(t, __fmul_rn(a2, B[8 + cc]));
        t = __fadd_rn(t, __fmul_rn(a3, B[12 + cc]));
        P.composeOut[lane] = B[lane];
        P.composeOut[16 + lane] = t;
    }
}

// Kabsch::Compute (kabsch.cpp:14-57), single thread (the reference never calls it on the hot path)
__global__ void kabsch_kernel(const float* A, const float* B, int n, float* T)
{
    if (threadIdx.x != 0) return;
    const float I16[16] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1 };
    for (int i = 0; i < 16; ++i) T[i] = I16[i];
    if (n == 0) return;
    float ca[3] = { 0, 0, 0 }, cb[3] = { 0, 0, 0 };
    for (int i = 0; i < n; ++i) for (int k = 0; k < 3; ++k) { ca[k] += A[3 * i + k]; cb[k] += B[3 * i + k]; }
    for (int k = 0; k < 3; ++k) { ca[k] /= (float)n; cb[k] /= (float)n; }
    float H[9] = { 0, 0, 0, 0, 0, 0, 0, 0, 0 };
    for (int i = 0; i < n; ++i)
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) H[3 * r + c] += (A[3 * i + r] - ca[r]) * (B[3 * i + c] - cb[c]);
    M3 V, W; float S[3];
    svd3(H, V, S, W);
    M3 Hm;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Hm.m[i][j] = H[3 * i + j];
    const float det = det3(Hm);
    const float d = (det != 0.f) ? (float)((det > 0.f) - (det < 0.f)) : 1.f;
    float R[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R[i][j] = (W.m[i][0] * V.m[j][0] + W.m[i][1] * V.m[j][1]) + (W.m[i][2] * d) * V.m[j][2];
    for (int i = 0; i < 3; ++i) {
        const float rc = (R[i][0] * ca[0] + R[i][1] * ca[1]) + R[i][2] * ca[2];
        T[4 * i] = R[i][0]; T[4 * i + 1] = R[i][1]; T[4 * i + 2] = R[i][2];
        T[4 * i + 3] = cb[i] - rc;
    }
}

}  // namespace

int orbf_launch_kabsch(orbf_context* c, const float* dA, const float* dB, int n, float* dT)
{
    kabsch_kernel<<<1, 32, 0, c->stream>>>(dA, dB, n, dT);
    ORBF_LAUNCH_CHECK(c);
    return ORBF_OK;
}

int orbf_ransac_reserve(orbf_context* c, const orbf_ransac_config& cfg)
{
    if (cfg.iterations < 1 || cfg.sample_size < 1 || cfg.sample_size > ORBF_MAX_SAMPLE || cfg.iterations > 100000) return ORBF_ERR_ARG;
    if (c->K > MAX_WORDS * 32) return ORBF_ERR_ARG;
    const size_t needHyp = (size_t)c->P * cfg.iterations;
    if ((size_t)c->hypCap < needHyp || !c->d_hyp) {
        ORBF_CUDA(c, cudaDeviceSynchronize());
        if (c->d_hyp) cudaFree(c->d_hyp);
        if (c->d_samples) cudaFree(c->d_samples);
        c->d_hyp = nullptr; c->d_samples = nullptr;
        ORBF_CUDA(c, cudaMalloc((void**)&c->d_hyp, needHyp * sizeof(orbf_hyp_trace)));
        ORBF_CUDA(c, cudaMalloc((void**)&c->d_samples, needHyp * ORBF_MAX_SAMPLE * sizeof(int)));
        c->hypCap = (int)needHyp;
    }
    const size_t needPts = (size_t)c->P * c->K;
    if (c->ptsCap < needPts) {
        ORBF_CUDA(c, cudaDeviceSynchronize());
        if (c->d_pts) cudaFree(c->d_pts);
        c->d_pts = nullptr; c->ptsCap = 0;
        ORBF_CUDA(c, cudaMalloc(&c->d_pts, needPts * sizeof(Pt6)));
        c->ptsCap = needPts;
    }
    const size_t smem = (size_t)c->K * (sizeof(orbf_dmatch) + 2 * sizeof(unsigned long long) + 2 * sizeof(uint16_t)) + (RAND_BUF + RAND_BUF / 2) * sizeof(uint16_t);
    {   // static (sort queues) + dynamic shared memory exceed the 48 KB default: always opt in
        cudaError_t e = cudaFuncSetAttribute(ransac_prepare_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return orbf_cuda_fail(c, e, "ransac smem attr", __FILE__, __LINE__);
    }
    return ORBF_OK;
}

// firstWave / lastWave: the hypothesis waves [firstWave, lastWave) of this launch (5 waves in all).  firstWave == 0 prepares the pairs;
// a later launch with firstWave > 0 on the same set continues where the earlier one stopped (the one-pair call checks the pair's
// done flag after two waves and only then queues the rest).  Every launch ends with the select kernel.
int orbf_launch_ransac(orbf_context* c, const RansacSet& rs, int pair0, int npairs, const orbf_ransac_config& cfg,
    const int* d_userSamples, bool standalone, bool fullTable, bool probeOnly, int firstWave, int lastWave, const float* d_composeIn, float* d_composeOut)
{
    if (npairs <= 0) return ORBF_OK;
    {
        const int r = orbf_ransac_reserve(c, cfg);
        if (r != ORBF_OK) return r;
    }
    const int iters = cfg.iterations, S = (int)cfg.sample_size;
    if (!probeOnly) { c->lastRs = rs; c->lastRansacCfg = cfg; }
    RansacParams P;
    P.rs = rs; P.cfg = cfg; P.good = c->d_good; P.goodCount = c->d_goodCount; P.pts = reinterpret_cast<Pt6*>(c->d_pts);
    P.samples = c->d_samples; P.userSamples = d_userSamples; P.hyp = c->d_hyp; P.res = c->d_rres; P.inliers = c->d_inliers;
    // slot 0 = the covariance latched on the context (quirk Q7), written only while it is still negative; slot 1 = per-call value of
    // standalone calls and of calls with an explicit covariance, which must neither see nor replace the latched one
    P.covReset = standalone ? 1 : 0;
    P.depthCov = c->d_depthCov + ((standalone || cfg.depth_cov >= 0.0) ? 1 : 0); P.K = c->K; P.iters = iters; P.S = S;
    P.state = reinterpret_cast<RState*>(c->d_rstate); P.hypLo = 0; P.hypHi = iters; P.pair0 = pair0;
    P.composeIn = d_composeIn; P.composeOut = d_composeIn ? d_composeOut : nullptr;
    constexpr int LAZY_ROWS = 8;                     // = the end of the second hypothesis wave
    P.tabRows = (fullTable || d_userSamples) ? iters : std::min(iters, LAZY_ROWS);
    {   // raster covariances of ErrorFunction2 (ransac.cpp:352-359), host libm like the reference
        const double ax = 58.0 / 180.0 * M_PI, ay = 45.0 / 180.0 * M_PI;
        const double sx = 3 * tan(ax / 640), sy = 3 * tan(ay / 480);
        P.covX = sx * sx; P.covY = sy * sy;
    }
    const size_t smem = (size_t)c->K * (sizeof(orbf_dmatch) + 2 * sizeof(unsigned long long) + 2 * sizeof(uint16_t)) + (RAND_BUF + RAND_BUF / 2) * sizeof(uint16_t);
    orbf_prof_begin(c, ST_RANSAC_PREPARE);
    if (firstWave == 0) {
    ransac_prepare_kernel<<<npairs, PR_THREADS, smem, c->stream>>>(P);
    ORBF_LAUNCH_CHECK(c);
    // every group runs the latch (a no-op once a value >= 0 is there): groups of one sequence are enqueued in pair order on one stream,
    // so a group whose pairs never reach scoring leaves the latch to the next one
    ransac_latch_kernel<<<1, 32, 0, c->stream>>>(P, npairs);
    ORBF_LAUNCH_CHECK(c);
    }
    orbf_prof_end(c, ST_RANSAC_PREPARE);
    if (probeOnly) return ORBF_OK;
    orbf_prof_begin(c, ST_RANSAC_HYP);
    // waves of hypotheses: the reference usually stops after a handful of iterations (> 80 % inliers ends the loop,
    // accepted hypotheses skip 10-20 iterations ahead), so later waves find their pair already done and exit at once
    const size_t hypSmem = (size_t)HY_WARPS * c->K * (sizeof(float) + sizeof(uint16_t));
    {   // static (masks) + dynamic shared memory can exceed the 48 KB default while the dynamic part alone does not: always opt in
        cudaError_t e = cudaFuncSetAttribute(ransac_hyp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hypSmem);
        if (e != cudaSuccess) return orbf_cuda_fail(c, e, "ransac hyp smem attr", __FILE__, __LINE__);
    }
    const size_t coopSmem = (size_t)c->K * (2 * sizeof(double) + sizeof(float) + sizeof(uint16_t));
    {
        cudaError_t e = cudaFuncSetAttribute(ransac_hyp_coop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)coopSmem);
        if (e != cudaSuccess) return orbf_cuda_fail(c, e, "ransac coop smem attr", __FILE__, __LINE__);
    }
    // (a first wave of 2 covers the common case — on well-matched consecutive frames the loop ends after its first accepted
    // hypothesis — and keeps that wave inside one residency of the GPU; every later wave is a few microseconds when nothing is left)
    const int waveEnd[5] = { 2, 8, 32, 96, iters };
    int lo = firstWave > 0 ? std::min(waveEnd[std::min(firstWave, 5) - 1], iters) : 0;
    for (int w = firstWave; w < std::min(lastWave, 5) && lo < iters; ++w) {
        const int hi = std::min(waveEnd[w], iters);
        if (hi <= lo) continue;
        if (hi > P.tabRows && lo <= P.tabRows && P.tabRows < iters) {          // this wave reads rows prepare did not draw
            ransac_table_kernel<<<npairs, 32, 0, c->stream>>>(P);
            ORBF_LAUNCH_CHECK(c);
            P.tabRows = iters;
        }
        P.hypLo = lo; P.hypHi = hi;
        if (hi - lo <= 8) {                   // few hypotheses per pair: a CTA per hypothesis (parallel scoring) instead of a warp
            ransac_hyp_coop_kernel<<<dim3(hi - lo, npairs), HC_THREADS, coopSmem, c->stream>>>(P);
        } else {
            dim3 grid((hi - lo + HY_WARPS - 1) / HY_WARPS, npairs);
            ransac_hyp_kernel<<<grid, HY_WARPS * 32, hypSmem, c->stream>>>(P);
        }
        ORBF_LAUNCH_CHECK(c);
        ransac_replay_kernel<<<(npairs + 127) / 128, 128, 0, c->stream>>>(P, npairs);
        ORBF_LAUNCH_CHECK(c);
        lo = hi;
    }
    orbf_prof_end(c, ST_RANSAC_HYP);
    orbf_prof_begin(c, ST_RANSAC_SELECT);
    ransac_select_kernel<<<npairs, SEL_THREADS, 0, c->stream>>>(P);
    ORBF_LAUNCH_CHECK(c);
    orbf_prof_end(c, ST_RANSAC_SELECT);
    return ORBF_OK;
}

const int* orbf_ransac_done_flag(orbf_context* c, int pair) { return &reinterpret_cast<const RState*>(c->d_rstate)[pair].done; }

// ---- Odometry::Compute, RANSAC strategy (Odometry/odometry.cpp:78-90) along a device-resident sequence --------------------------
// Composition rule pose[k + 1] = T12[k] * pose[k]: cv::Mat's 4x4 float product (every element = four products summed left to right
// in float), inherently sequential in k; 16 lanes own one element each and pass the pose along through shared memory.
// Inlier flags: Frame::mvbOutlier starts all-true, SetInlier(m.trainIdx) clears it for the winner's inliers.
namespace {
__global__ void __launch_bounds__(32) compose_kernel(const orbf_ransac_result* __restrict__ res, int npairs, const float* __restrict__ pose0, float* __restrict__ poses)
{
    __shared__ float sP[16];
    const int lane = threadIdx.x, r = (lane >> 2) & 3, c = lane & 3;
    if (lane < 16) { sP[lane] = pose0[lane]; poses[lane] = pose0[lane]; }
    __syncwarp();
    for (int k = 0; k < npairs; ++k) {
        const float* A = res[k].T12;
        float t = 0.f;
        if (lane < 16) {
            t = __fmul_rn(A[4 * r], sP[c]);
            t = __fadd_rn(t, __fmul_rn(A[4 * r + 1], sP[4 + c]));
            t = __fadd_rn(t, __fmul_rn(A[4 * r + 2], sP[8 + c]));
            t = __fadd_rn(t, __fmul_rn(A[4 * r + 3], sP[12 + c]));
        }
        __syncwarp();
        if (lane < 16) { sP[lane] = t; poses[(size_t)(k + 1) * 16 + lane] = t; }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(256) inlier_flag_kernel(const orbf_ransac_result* __restrict__ res, const orbf_dmatch* __restrict__ inliers, int K, int npairs,
    uint8_t* __restrict__ outlier /* [npairs + 1][K], preset to 1 */)
{
    const int p = blockIdx.x;
    const int n = res[p].n_inliers;
    for (int i = threadIdx.x; i < n; i += 256) outlier[(size_t)(p + 1) * K + inliers[(size_t)p * K + i].trainIdx] = 0;
}

// Ransac::mpSourceCloud / mpTargetCloud (Odometry/ransac.cpp:163-189): the 3D points of the depth-valid matches in m12 order (before the
// sort), one pcl::PointXYZ (16 bytes: x, y, z, 1.0f) per point; both stay empty when m12 has fewer than min_inlier_th entries.  CTA per pair.
__global__ void __launch_bounds__(256) ransac_clouds_kernel(RansacSet rs, int K, int pair0, unsigned minInl, int checkDepth, float4* __restrict__ src,
    float4* __restrict__ tgt, int* __restrict__ counts)
{
    __shared__ int sWarp[8];
    __shared__ int sBase;
    const int pair = pair0 + blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int qs = 0, ts = 0;
    if (rs.pairs) { qs = rs.pairs[2 * pair]; ts = rs.pairs[2 * pair + 1]; }
    const long long so = (long long)qs * rs.slotStride, to = (long long)ts * rs.slotStride;
    const orbf_dmatch* m12 = rs.matches + (long long)pair * K;
    const int nm = rs.matchCount[pair];
    float4* os = src + (long long)pair * K; float4* ot = tgt + (long long)pair * K;
    if (tid == 0) sBase = 0;
    __syncthreads();
    if ((unsigned)nm >= minInl) {
        for (int base = 0; base < nm; base += 256) {
            const int i = base + tid;
            bool keep = false;
            float4 a = make_float4(0.f, 0.f, 0.f, 1.f), b = a;
            if (i < nm) {
                const orbf_dmatch m = m12[i];
                a.x = rs.sx[so + m.queryIdx]; a.y = rs.sy[so + m.queryIdx]; a.z = rs.sz[so + m.queryIdx];
                b.x = rs.tx[to + m.trainIdx]; b.y = rs.ty[to + m.trainIdx]; b.z = rs.tz[to + m.trainIdx];
                keep = !(checkDepth && (isnan(a.z) || isnan(b.z) || a.z <= 0 || b.z <= 0));
            }
            const unsigned mk = __ballot_sync(0xffffffffu, keep);
            if (lane == 0) sWarp[warp] = __popc(mk);
            __syncthreads();
            int off = sBase;
            for (int w = 0; w < warp; ++w) off += sWarp[w];
            if (keep) { const int o = off + __popc(mk & ((1u << lane) - 1)); os[o] = a; ot[o] = b; }
            __syncthreads();
            if (tid == 0) { int t = 0; for (int w = 0; w < 8; ++w) t += sWarp[w]; sBase += t; }
            __syncthreads();
        }
    }
    if (tid == 0) counts[pair] = sBase;
}
}  // namespace

int orbf_launch_ransac_clouds(orbf_context* c, int pair0, int npairs)
{
    if (npairs <= 0) return ORBF_OK;
    if (!c->lastRs.matches) return ORBF_ERR_STATE;
    const size_t need = (size_t)c->P * c->K;
    if (c->cloudCap < need) {
        if (c->d_cloudSrc) cudaFree(c->d_cloudSrc);
        if (c->d_cloudTgt) cudaFree(c->d_cloudTgt);
        if (c->d_cloudCount) cudaFree(c->d_cloudCount);
        c->d_cloudSrc = c->d_cloudTgt = nullptr; c->d_cloudCount = nullptr; c->cloudCap = 0;
        ORBF_CUDA(c, cudaMalloc(&c->d_cloudSrc, need * sizeof(float4)));
        ORBF_CUDA(c, cudaMalloc(&c->d_cloudTgt, need * sizeof(float4)));
        ORBF_CUDA(c, cudaMalloc(&c->d_cloudCount, (size_t)c->P * sizeof(int)));
        c->cloudCap = need;
    }
    ransac_clouds_kernel<<<npairs, 256, 0, c->stream>>>(c->lastRs, c->K, pair0, c->lastRansacCfg.min_inlier_th, c->lastRansacCfg.check_depth, c->d_cloudSrc,
        c->d_cloudTgt, c->d_cloudCount);
    ORBF_LAUNCH_CHECK(c);
    return ORBF_OK;
}

int orbf_launch_compose(orbf_context* c, int npairs, const float* d_pose0, float* d_poses, uint8_t* d_outlier)
{
    compose_kernel<<<1, 32, 0, c->stream>>>(c->d_rres, npairs, d_pose0, d_poses);
    ORBF_LAUNCH_CHECK(c);
    if (d_outlier) {
        ORBF_CUDA(c, cudaMemsetAsync(d_outlier, 1, (size_t)(npairs + 1) * c->K, c->stream));
        if (npairs > 0) { inlier_flag_kernel<<<npairs, 256, 0, c->stream>>>(c->d_rres, c->d_inliers, c->K, npairs, d_outlier); ORBF_LAUNCH_CHECK(c); }
    }
    return ORBF_OK;
}
