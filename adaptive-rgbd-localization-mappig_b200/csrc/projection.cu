// csrc/projection.cu — the windowed / word-bucketed Hamming searches of the reference matcher:
// Matcher::ProjectionMatch (Features/matcher.cpp:90-143), the search of Matcher::Fuse (:212-296) and Matcher::BoWMatch (:145-209).
//
// ProjectionMatch: landmarks already
// projected into a frame are matched to the frame's features inside a square window (Frame::GetFeaturesInArea,
// Core/frame.cpp:258-274: a linear scan in feature order, |dx| < r && |dy| < r in float).
//
// The reference walks the landmarks in order and a feature taken by an earlier landmark (one with Observations() > 0) is skipped
// by later ones, so the result depends on the order.  Two kernels keep that exact:
//   1. proj_candidates_kernel (warp per landmark, all landmarks in parallel): window test + 256-bit Hamming distance for every
//      feature, candidates compacted in feature order as (distance << 16 | feature) keys — the order-independent part, and
//      all of the arithmetic.
//   2. proj_resolve_kernel (one warp): landmarks in order; among a landmark's candidates that are still free, the two smallest
//      keys are exactly the reference's (best, second best) under its strict '<' updates in arrival order; octave / ratio test
//      (double); the accepted feature is marked taken.  32 landmarks are fetched into registers at a time and resolved one after
//      the other by their owning lanes against taken flags in shared memory (0.1 us per landmark instead of 2.7 us when every
//      landmark's turn waited on global memory).
#include "orbf_internal.h"
#include "undistort_device.h"

namespace {

constexpr uint32_t PJ_NONE = 0xFFFFFFFFu;

constexpr int PJ_REG = 8;          // candidates per landmark the resolve pass keeps in registers

__global__ void __launch_bounds__(128) proj_candidates_kernel(const float* __restrict__ kpx, const float* __restrict__ kpy, const int* __restrict__ kpOct,
    const uint8_t* __restrict__ desc, int nFeat, const uint8_t* __restrict__ lmDesc, const float* __restrict__ projX, const float* __restrict__ projY,
    const uint8_t* __restrict__ lmFlags, int nLm, float radius, uint32_t* __restrict__ cand, int* __restrict__ candOct /* [nLm][PJ_REG] */, int* __restrict__ candCount)
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
                const int pos = n + __popc(m & ((1u << lane) - 1));
                out[pos] = ((uint32_t)d << 16) | (uint32_t)j;
                if (pos < PJ_REG) candOct[(size_t)i * PJ_REG + pos] = kpOct[j];
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

// One warp, landmarks in order.  The pass is latency-, not work-bound (a window holds a handful of candidates), so nothing of a
// landmark's turn may wait on global memory and as few turns as possible may wait on each other:
//   * the warp takes 32 landmarks at a time — lane L holds landmark L's count, its first PJ_REG candidate keys and their octaves in
//     registers, fetched while the previous batch was resolved;
//   * every lane decides its landmark at once against the taken flags (shared memory) as they stand; a lane whose candidates include
//     a feature that an EARLIER lane of the batch has just accepted (and will mark taken) is in conflict: everything before the first
//     such lane is final and is committed, the rest decides again.  Without contention a batch takes one round;
//   * a landmark with more than PJ_REG candidates (crowded windows) is resolved in its turn by the whole warp from the candidate list.
// Among a landmark's free candidates the two smallest (distance << 16 | feature) keys are exactly the reference's (best, second best)
// under its strict '<' updates in feature order.
struct PjBatch { int n, fl; uint32_t k[PJ_REG]; int oc[PJ_REG]; };

__device__ __forceinline__ void pj_load(PjBatch& b, int i, int nLm, int nFeat, const uint32_t* __restrict__ cand, const int* __restrict__ candOct,
    const int* __restrict__ candCount, const uint8_t* __restrict__ lmFlags)
{
    b.fl = i < nLm ? lmFlags[i] : 0;
    b.n = (b.fl & 1) ? candCount[i] : 0;
#pragma unroll
    for (int c = 0; c < PJ_REG; ++c) {
        const bool on = c < b.n && b.n <= PJ_REG;
        b.k[c] = on ? cand[(size_t)i * nFeat + c] : PJ_NONE;
        b.oc[c] = on ? candOct[(size_t)i * PJ_REG + c] : -1;
    }
}

__global__ void __launch_bounds__(32) proj_resolve_kernel(const uint32_t* __restrict__ cand, const int* __restrict__ candOct, const int* __restrict__ candCount,
    const int* __restrict__ kpOct, int nFeat, const uint8_t* __restrict__ lmFlags, int nLm, const uint8_t* __restrict__ featTaken, float nnRatio, int thHigh,
    int* __restrict__ bestIdx, int* __restrict__ nMatches)
{
    extern __shared__ uint8_t taken[];                        // [nFeat]
    const int lane = threadIdx.x;
    PjBatch cur, nxt;
    pj_load(cur, lane, nLm, nFeat, cand, candOct, candCount, lmFlags);
    for (int j = lane; j < nFeat; j += 32) taken[j] = featTaken ? featTaken[j] : 0;
    __syncwarp();
    int nm = 0;
    for (int i0 = 0; i0 < nLm; i0 += 32) {
        pj_load(nxt, i0 + 32 + lane, nLm, nFeat, cand, candOct, candCount, lmFlags);          // in flight while this batch is resolved
        int best = -1;
        const unsigned big = __ballot_sync(0xffffffffu, cur.n > PJ_REG);
        unsigned pending = __ballot_sync(0xffffffffu, cur.n > 0 && cur.n <= PJ_REG);
        int pos = 0;
        while (pos < 32) {
            const unsigned bigRem = big & (0xffffffffu << pos);
            const int stop = bigRem ? __ffs(bigRem) - 1 : 32;
            const unsigned range = (stop >= 32 ? 0xffffffffu : ((1u << stop) - 1)) & (0xffffffffu << pos);
            unsigned grp = pending & range;
            while (grp) {
                const bool mine = (grp >> lane) & 1;
                int acc = -1;
                if (mine) {
                    uint32_t k1 = PJ_NONE, k2 = PJ_NONE; int o1 = -1, o2 = -1;
#pragma unroll
                    for (int c = 0; c < PJ_REG; ++c) {
                        const uint32_t key = cur.k[c];
                        if (key != PJ_NONE && !taken[key & 0xFFFFu]) {
                            if (key < k1) { k2 = k1; o2 = o1; k1 = key; o1 = cur.oc[c]; }
                            else if (key < k2) { k2 = key; o2 = cur.oc[c]; }
                        }
                    }
                    // bestDist1 <= TH_HIGH, and not (bestLevel == bestLevel2 && bestDist1 > mfNNratio * bestDist2) — float ratio promoted to double
                    if (k1 != PJ_NONE && (int)(k1 >> 16) <= thHigh
                        && !(k2 != PJ_NONE && o1 == o2 && (double)(k1 >> 16) > __dmul_rn((double)nnRatio, (double)(k2 >> 16))))
                        acc = (int)(k1 & 0xFFFFu);
                }
                // features this round would mark taken, and the lanes after their owner that hold them as a candidate
                unsigned takers = __ballot_sync(0xffffffffu, mine && acc >= 0 && (cur.fl & 2));
                bool conflict = false;
                while (takers) {
                    const int a = __ffs(takers) - 1;
                    takers &= takers - 1;
                    const uint32_t tj = (uint32_t)__shfl_sync(0xffffffffu, acc, a);
                    if (mine && lane > a) {
#pragma unroll
                        for (int c = 0; c < PJ_REG; ++c) conflict |= cur.k[c] != PJ_NONE && (cur.k[c] & 0xFFFFu) == tj;
                    }
                }
                const unsigned cm = __ballot_sync(0xffffffffu, conflict);
                const unsigned fin = cm ? ((1u << (__ffs(cm) - 1)) - 1) : 0xffffffffu;         // lanes before the first conflict are final
                if (mine && ((fin >> lane) & 1)) {
                    best = acc;
                    if (acc >= 0) { ++nm; if (cur.fl & 2) taken[acc] = 1; }
                }
                __syncwarp();
                grp &= ~fin;
            }
            if (stop < 32) {        // a crowded window: the whole warp on landmark i0 + stop
                const int ib = i0 + stop, nb = __shfl_sync(0xffffffffu, cur.n, stop), flb = __shfl_sync(0xffffffffu, cur.fl, stop);
                uint32_t k1 = PJ_NONE, k2 = PJ_NONE;
                for (int c = lane; c < nb; c += 32) {
                    const uint32_t key = cand[(size_t)ib * nFeat + c];
                    if (!taken[key & 0xFFFFu]) pj_insert(k1, k2, key);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const uint32_t a1 = __shfl_xor_sync(0xffffffffu, k1, o), a2 = __shfl_xor_sync(0xffffffffu, k2, o);
                    pj_insert(k1, k2, a1);
                    pj_insert(k1, k2, a2);
                }
                if (k1 != PJ_NONE && (int)(k1 >> 16) <= thHigh) {
                    const int j1 = (int)(k1 & 0xFFFFu);
                    bool ok = true;
                    if (k2 != PJ_NONE) {
                        const int j2 = (int)(k2 & 0xFFFFu);
                        if (kpOct[j1] == kpOct[j2] && (double)(k1 >> 16) > __dmul_rn((double)nnRatio, (double)(k2 >> 16))) ok = false;
                    }
                    if (ok && lane == stop) {
                        best = j1;
                        ++nm;
                        if (flb & 2) taken[j1] = 1;
                    }
                }
                __syncwarp();
            }
            pos = stop + 1;
        }
        if (i0 + lane < nLm) bestIdx[i0 + lane] = best;
        cur = nxt;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nm += __shfl_xor_sync(0xffffffffu, nm, o);
    if (lane == 0) *nMatches = nm;
}


// ---- Matcher::Fuse, search part (Features/matcher.cpp:212-296): warp per landmark, landmarks independent ---------------------------
// p3Dc = Rcw * p3Dw + tcw exactly as cv::gemm evaluates a 3x3 * 3x1 product (float, left to right) followed by the double
// alpha / beta combination; projection with separate multiply and add (the library is built with -fmad=false, quirk Q4).
struct FuseCamera { float R[9], t[3], fx, fy, cx, cy, mbf, minX, maxX, minY, maxY; };

__global__ void __launch_bounds__(128) fuse_search_kernel(FuseCamera cam, const float* __restrict__ kpx, const float* __restrict__ kpy, const float* __restrict__ uRight,
    const uint8_t* __restrict__ desc, int nFeat, const float* __restrict__ lmPos, const uint8_t* __restrict__ lmDesc, const uint8_t* __restrict__ lmValid, int nLm,
    float radius, int thLow, int* __restrict__ bestIdx, int* __restrict__ bestDist)
{
    const int lane = threadIdx.x & 31, i = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (i >= nLm) return;
    uint32_t best = PJ_NONE;
    if (lmValid[i]) {
        float pc[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            float t = __fmul_rn(cam.R[3 * r], lmPos[3 * i]);
            t = __fadd_rn(t, __fmul_rn(cam.R[3 * r + 1], lmPos[3 * i + 1]));
            t = __fadd_rn(t, __fmul_rn(cam.R[3 * r + 2], lmPos[3 * i + 2]));
            pc[r] = __double2float_rn(__dadd_rn((double)t, (double)cam.t[r]));
        }
        if (!(pc[2] < 0.0f)) {
            const float invz = __fdiv_rn(1.0f, pc[2]);
            const float u = __fadd_rn(__fmul_rn(cam.fx, __fmul_rn(pc[0], invz)), cam.cx), v = __fadd_rn(__fmul_rn(cam.fy, __fmul_rn(pc[1], invz)), cam.cy);
            if (u >= cam.minX && u < cam.maxX && v >= cam.minY && v < cam.maxY) {
                const float ur = __fsub_rn(u, __fmul_rn(cam.mbf, invz));
                const uint4 la = __ldg(reinterpret_cast<const uint4*>(lmDesc + (size_t)i * 32)), lb = __ldg(reinterpret_cast<const uint4*>(lmDesc + (size_t)i * 32) + 1);
                for (int j = lane; j < nFeat; j += 32) {
                    const float x = kpx[j], y = kpy[j];
                    if (!(fabsf(__fsub_rn(x, u)) < radius && fabsf(__fsub_rn(y, v)) < radius)) continue;
                    const float ex = __fsub_rn(u, x), ey = __fsub_rn(v, y), r = uRight[j];
                    float e2 = __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
                    if (r >= 0) {
                        const float er = __fsub_rn(ur, r);
                        e2 = __fadd_rn(e2, __fmul_rn(er, er));
                        if (e2 > 7.8f) continue;
                    } else if (e2 > 5.99f) continue;
                    const uint4 fa = __ldg(reinterpret_cast<const uint4*>(desc + (size_t)j * 32)), fb = __ldg(reinterpret_cast<const uint4*>(desc + (size_t)j * 32) + 1);
                    const int d = __popc(la.x ^ fa.x) + __popc(la.y ^ fa.y) + __popc(la.z ^ fa.z) + __popc(la.w ^ fa.w)
                                + __popc(lb.x ^ fb.x) + __popc(lb.y ^ fb.y) + __popc(lb.z ^ fb.z) + __popc(lb.w ^ fb.w);
                    best = min(best, ((uint32_t)d << 16) | (uint32_t)j);          // strict '<' in feature order = smallest (distance, index)
                }
            }
        }
    }
    best = __reduce_min_sync(0xffffffffu, best);
    if (lane == 0) {
        const bool ok = best != PJ_NONE && (int)(best >> 16) <= thLow;
        bestIdx[i] = ok ? (int)(best & 0xFFFFu) : -1;
        bestDist[i] = ok ? (int)(best >> 16) : -1;
    }
}

// ---- Matcher::BoWMatch (Features/matcher.cpp:145-209) --------------------------------------------------------------------------
// Entry e of keyframe 1's flattened feature vector (= the reference's processing order: words ascending, bucket order) finds
// its word in keyframe 2 by binary search, then its best / second best in that bucket; the std::set of used train indices is
// "the first entry in processing order wins", i.e. an atomicMin of the entry number per train feature.
__global__ void __launch_bounds__(128) bow_best_kernel(const int* __restrict__ words1, const int* __restrict__ off1, const int* __restrict__ idx1, int nw1,
    const uint8_t* __restrict__ desc1, const int* __restrict__ words2, const int* __restrict__ off2, const int* __restrict__ idx2, int nw2,
    const uint8_t* __restrict__ desc2, float nnRatio, int thLow, int nEntries, int* __restrict__ entryTrain, int* __restrict__ entryDist, int* __restrict__ firstUser)
{
    const int lane = threadIdx.x & 31, e = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (e >= nEntries) return;
    // word of entry e: last a with off1[a] <= e
    int lo = 0, hi = nw1 - 1;
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (off1[mid] <= e) lo = mid; else hi = mid - 1; }
    const int w = words1[lo];
    int l2 = 0, h2 = nw2;                                        // lower_bound of w in words2
    while (l2 < h2) { const int mid = (l2 + h2) >> 1; if (words2[mid] < w) l2 = mid + 1; else h2 = mid; }
    int train = -1, dist = -1;
    if (l2 < nw2 && words2[l2] == w) {
        const int q = idx1[e], b0 = off2[l2], nb = off2[l2 + 1] - b0;
        const uint4 qa = __ldg(reinterpret_cast<const uint4*>(desc1 + (size_t)q * 32)), qb = __ldg(reinterpret_cast<const uint4*>(desc1 + (size_t)q * 32) + 1);
        uint32_t k1 = PJ_NONE, k2 = PJ_NONE;
        for (int p = lane; p < nb; p += 32) {
            const int t = idx2[b0 + p];
            const uint4 ta = __ldg(reinterpret_cast<const uint4*>(desc2 + (size_t)t * 32)), tb = __ldg(reinterpret_cast<const uint4*>(desc2 + (size_t)t * 32) + 1);
            const int d = __popc(qa.x ^ ta.x) + __popc(qa.y ^ ta.y) + __popc(qa.z ^ ta.z) + __popc(qa.w ^ ta.w)
                        + __popc(qb.x ^ tb.x) + __popc(qb.y ^ tb.y) + __popc(qb.z ^ tb.z) + __popc(qb.w ^ tb.w);
            pj_insert(k1, k2, ((uint32_t)d << 16) | (uint32_t)p);          // position in the bucket: strict '<' keeps the earlier one
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const uint32_t o1 = __shfl_xor_sync(0xffffffffu, k1, o), o2 = __shfl_xor_sync(0xffffffffu, k2, o);
            pj_insert(k1, k2, o1);
            pj_insert(k1, k2, o2);
        }
        if (k1 != PJ_NONE && (int)(k1 >> 16) <= thLow) {
            const float d1 = (float)(k1 >> 16), d2 = k2 != PJ_NONE ? (float)(k2 >> 16) : __int_as_float(0x7f800000);   // (float)DBL_MAX = +inf
            if (d1 < __fmul_rn(nnRatio, d2)) { train = idx2[b0 + (int)(k1 & 0xFFFFu)]; dist = (int)(k1 >> 16); }
        }
    }
    if (lane == 0) {
        entryTrain[e] = train; entryDist[e] = dist;
        if (train >= 0) atomicMin(&firstUser[train], e);
    }
}

// entries that own their train feature, compacted in processing order (single CTA: the lists are a few thousand entries)
__global__ void __launch_bounds__(256) bow_emit_kernel(const int* __restrict__ idx1, const int* __restrict__ entryTrain, const int* __restrict__ entryDist,
    const int* __restrict__ firstUser, int nEntries, orbf_dmatch* __restrict__ out, int* __restrict__ nOut)
{
    __shared__ int sWarp[8];
    __shared__ int sBase;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) sBase = 0;
    __syncthreads();
    for (int base = 0; base < nEntries; base += 256) {
        const int e = base + threadIdx.x;
        const int t = e < nEntries ? entryTrain[e] : -1;
        const bool keep = t >= 0 && firstUser[t] == e;
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) sWarp[warp] = __popc(m);
        __syncthreads();
        int off = sBase;
        for (int w = 0; w < warp; ++w) off += sWarp[w];
        if (keep) {
            orbf_dmatch dm; dm.queryIdx = idx1[e]; dm.trainIdx = t; dm.imgIdx = -1; dm.distance = (float)entryDist[e];
            out[off + __popc(m & ((1u << lane) - 1))] = dm;
        }
        __syncthreads();
        if (threadIdx.x == 0) { int s = 0; for (int w = 0; w < 8; ++w) s += sWarp[w]; sBase += s; }
        __syncthreads();
    }
    if (threadIdx.x == 0) *nOut = sBase;
}


// ---- Frame::UndistortKeyPoints (Core/frame.cpp:286-313): cv::undistortPoints(pts, pts, K, dist, Mat(), K) --------------------------
// cvUndistortPointsInternal with the default criteria (5 fixed iterations), every operation in double in OpenCV's order (the
// library is compiled with -fmad=false, so nothing contracts); thread per point.
__global__ void __launch_bounds__(128) undistort_kernel(UndistortParams U, const float* __restrict__ xy, int n, float* __restrict__ out)
{
    const int i = blockIdx.x * 128 + threadIdx.x;
    if (i >= n) return;
    undistort_point(U, xy[2 * i], xy[2 * i + 1], &out[2 * i], &out[2 * i + 1]);
}

// ---- Frame::ExtractFeatures' per-keypoint tail (Core/frame.cpp:148-164) for keypoints that did not come out of the ORB extractor
// (the adaptive-FAST route): mvKeysUn, mvuRight and mvKeys3Dc from the keypoint position and its raw depth sample (0 = outside the
// image / no sample).  Same operation order as describe.cu; thread per keypoint.
__global__ void __launch_bounds__(128) unproject_kernel(UndistortParams U, int distorted, const float* __restrict__ xy, const uint16_t* __restrict__ raw, int n,
    float cx, float cy, float invfx, float invfy, float mbf, float depthFactor, float* __restrict__ xyz, float* __restrict__ uright, float* __restrict__ xyUn)
{
    const int i = blockIdx.x * 128 + threadIdx.x;
    if (i >= n) return;
    const float x = xy[2 * i], y = xy[2 * i + 1];
    float uX = x, uY = y;
    if (distorted) undistort_point(U, x, y, &uX, &uY);
    float X = 0.f, Y = 0.f, Z = 0.f, ur = -1.f;
    const float z = __fmul_rn((float)raw[i], depthFactor);
    if (z > 0) {
        ur = __fsub_rn(uX, __fdiv_rn(mbf, z));
        X = __fmul_rn(__fmul_rn(__fsub_rn(uX, cx), z), invfx);
        Y = __fmul_rn(__fmul_rn(__fsub_rn(uY, cy), z), invfy);
        Z = z;
    }
    xyz[3 * i] = X; xyz[3 * i + 1] = Y; xyz[3 * i + 2] = Z; uright[i] = ur; xyUn[2 * i] = uX; xyUn[2 * i + 1] = uY;
}

}  // namespace

int orbf_launch_unproject(orbf_context* c, const float* d_xy, const uint16_t* d_raw, int n, float* d_xyz, float* d_uright, float* d_xyUn)
{
    if (n <= 0) return ORBF_OK;
    const orbf_config& g = c->cfg;
    UndistortParams U{(double)g.fx, (double)g.fy, (double)g.cx, (double)g.cy, (double)g.k1, (double)g.k2, (double)g.p1, (double)g.p2, (double)g.k3};
    unproject_kernel<<<(n + 127) / 128, 128, 0, c->stream>>>(U, g.k1 != 0.f ? 1 : 0, d_xy, d_raw, n, g.cx, g.cy, 1.0f / g.fx, 1.0f / g.fy, g.mbf, g.depth_factor,
        d_xyz, d_uright, d_xyUn);
    ORBF_LAUNCH_CHECK(c);
    return ORBF_OK;
}

int orbf_launch_projection_match(orbf_context* c, const float* d_kpx, const float* d_kpy, const int* d_kpoct, const uint8_t* d_desc, int nFeat, const uint8_t* d_lmDesc,
    const float* d_projX, const float* d_projY, const uint8_t* d_lmFlags, int nLm, const uint8_t* d_featTaken, float radius, float nnRatio, int thHigh,
    uint32_t* d_cand, int* d_candOct, int* d_candCount, int* d_bestIdx, int* d_nMatches)
{
    if (nLm <= 0) return ORBF_OK;
    proj_candidates_kernel<<<(nLm + 3) / 4, 128, 0, c->stream>>>(d_kpx, d_kpy, d_kpoct, d_desc, nFeat, d_lmDesc, d_projX, d_projY, d_lmFlags, nLm, radius, d_cand, d_candOct,
        d_candCount);
    ORBF_LAUNCH_CHECK(c);
    const size_t smem = (size_t)std::max(nFeat, 1);           // nFeat <= 65535 (16-bit feature index in the candidate keys)
    ORBF_CUDA(c, cudaFuncSetAttribute(proj_resolve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    proj_resolve_kernel<<<1, 32, smem, c->stream>>>(d_cand, d_candOct, d_candCount, d_kpoct, nFeat, d_lmFlags, nLm, d_featTaken, nnRatio, thHigh, d_bestIdx, d_nMatches);
    ORBF_LAUNCH_CHECK(c);
    return ORBF_OK;
}

int orbf_launch_fuse_search(orbf_context* c, const float* Rcw, const float* tcw, const float* camera /* fx fy cx cy mbf minX maxX minY maxY */, const float* d_kpx,
    const float* d_kpy, const float* d_uright, const uint8_t* d_desc, int nFeat, const float* d_lmPos, const uint8_t* d_lmDesc, const uint8_t* d_lmValid, int nLm,
    float radius, int thLow, int* d_bestIdx, int* d_bestDist)
{
    if (nLm <= 0) return ORBF_OK;
    FuseCamera cam;
    for (int i = 0; i < 9; ++i) cam.R[i] = Rcw[i];
    for (int i = 0; i < 3; ++i) cam.t[i] = tcw[i];
    cam.fx = camera[0]; cam.fy = camera[1]; cam.cx = camera[2]; cam.cy = camera[3]; cam.mbf = camera[4];
    cam.minX = camera[5]; cam.maxX = camera[6]; cam.minY = camera[7]; cam.maxY = camera[8];
    fuse_search_kernel<<<(nLm + 3) / 4, 128, 0, c->stream>>>(cam, d_kpx, d_kpy, d_uright, d_desc, nFeat, d_lmPos, d_lmDesc, d_lmValid, nLm, radius, thLow, d_bestIdx, d_bestDist);
    ORBF_LAUNCH_CHECK(c);
    return ORBF_OK;
}

int orbf_launch_bow_match(orbf_context* c, const int* d_words1, const int* d_off1, const int* d_idx1, int nw1, const uint8_t* d_desc1, const int* d_words2,
    const int* d_off2, const int* d_idx2, int nw2, const uint8_t* d_desc2, int nTrain, float nnRatio, int thLow, int nEntries, int* d_entryTrain, int* d_entryDist,
    int* d_firstUser, orbf_dmatch* d_out, int* d_nOut)
{
    ORBF_CUDA(c, cudaMemsetAsync(d_nOut, 0, sizeof(int), c->stream));
    if (nEntries <= 0 || nw1 <= 0 || nw2 <= 0) return ORBF_OK;
    ORBF_CUDA(c, cudaMemsetAsync(d_firstUser, 0x7F, (size_t)std::max(nTrain, 1) * sizeof(int), c->stream));
    bow_best_kernel<<<(nEntries + 3) / 4, 128, 0, c->stream>>>(d_words1, d_off1, d_idx1, nw1, d_desc1, d_words2, d_off2, d_idx2, nw2, d_desc2, nnRatio, thLow, nEntries,
        d_entryTrain, d_entryDist, d_firstUser);
    ORBF_LAUNCH_CHECK(c);
    bow_emit_kernel<<<1, 256, 0, c->stream>>>(d_idx1, d_entryTrain, d_entryDist, d_firstUser, nEntries, d_out, d_nOut);
    ORBF_LAUNCH_CHECK(c);
    return ORBF_OK;
}

int orbf_launch_undistort(orbf_context* c, const float* d_xy, int n, float fx, float fy, float cx, float cy, const float* dist5, float* d_out)
{
    if (n <= 0) return ORBF_OK;
    UndistortParams U{(double)fx, (double)fy, (double)cx, (double)cy, (double)dist5[0], (double)dist5[1], (double)dist5[2], (double)dist5[3], (double)dist5[4]};
    undistort_kernel<<<(n + 127) / 128, 128, 0, c->stream>>>(U, d_xy, n, d_out);
    ORBF_LAUNCH_CHECK(c);
    return ORBF_OK;
}
