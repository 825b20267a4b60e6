// csrc/quadtree.cu — quadtree keypoint distribution (reference DistributeOctTree + DivideNode,
// Features/orbextractor.cpp:412-663) for one (frame, level) per CTA.
//
// The reference is a sequential std::list algorithm whose OUTPUT ORDER defines descriptor row order and
// therefore match indices.  It is restated here as level-synchronous parallel rounds that reproduce the
// list order exactly:
//   * keys never move; each key carries the list position of its node (u16, shared memory);
//   * a phase-1 pass splits every multi-key node: children are created in (parent list order, n1..n4)
//     order and pushed to the FRONT, so the new list is reverse(creation order) ++ kept single-key nodes;
//   * phase 2 (entered when size + 3*nToExpand > N) expands the nodes created by the previous round,
//     largest count first; equal counts are ordered by creation sequence, newest first (the oracle's
//     definition of quirk Q3 — the reference compares heap pointers there).  The sequential "stop as soon
//     as size >= N" rule becomes a prefix sum over the sorted gains (nonempty children - 1);
//   * per leaf the max-response key wins, first in candidate order on ties (strict '>' in the reference).
// Before that, the per-cell candidate slots written by fast.cu are gathered into one contiguous list in
// the reference's push_back order (cell row-major, pixel row-major inside a cell).
#include "orbf_internal.h"

namespace {

constexpr int QT_THREADS = 256;
constexpr int QT_SMEM_KEYS = 8192;

struct QtParams {
    const LevelGeom* lg; const CellDesc* cells;
    const uint32_t* cellCand; const int* cellCount;
    uint32_t* cand; int* candCount; uint16_t* nodeScratch;
    uint32_t* lkp; int* lkpCount;
    int cellSlotTotal, nCellsTotal, candTotal, kpStageTotal;
    int slot0, NM, maxCells;
};

__device__ __forceinline__ int warp_excl_scan(int v, int lane, int& total)
{
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    total = __shfl_sync(0xffffffffu, x, 31);
    return x - v;
}

// exclusive scan of src[0..n) into dst (may alias), executed by warp 0 only; returns the total to warp 0's lanes
__device__ int scan_by_warp0(const int* src, int* dst, int n, int lane)
{
    int carry = 0;
    for (int b = 0; b < n; b += 32) {
        const int i = b + lane;
        const int v = (i < n) ? src[i] : 0;
        int tot;
        const int e = warp_excl_scan(v, lane, tot);
        if (i < n) dst[i] = carry + e;
        carry += tot;
    }
    return carry;
}

__device__ __forceinline__ int quadrant(uint32_t key, short4 r)
{
    const int x = key & 0x7FF, y = (key >> 11) & 0x7FF;
    const int mx = r.x + ((r.z - r.x + 1) >> 1), my = r.y + ((r.w - r.y + 1) >> 1);   // ceil(half) (DivideNode :414-415)
    return (x < mx) ? ((y < my) ? 0 : 2) : ((y < my) ? 1 : 3);
}

__global__ void __launch_bounds__(QT_THREADS) quadtree_kernel(QtParams P)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const int NM = P.NM;
    short4* rect0 = reinterpret_cast<short4*>(smem);
    short4* rect1 = rect0 + NM;
    int* cnt0 = reinterpret_cast<int*>(rect1 + NM);
    int* cnt1 = cnt0 + NM;
    int4* cc = reinterpret_cast<int4*>(cnt1 + NM);          // child key counts (also reused as 'best' at the end)
    ushort4* childPos = reinterpret_cast<ushort4*>(cc + NM);
    int* nchByRank = reinterpret_cast<int*>(childPos + NM);  // number of non-empty children, indexed by processing rank
    int* baseByRank = nchByRank + NM;
    int* keptTmp = baseByRank + NM;
    uint16_t* keptPos = reinterpret_cast<uint16_t*>(keptTmp + NM);
    uint16_t* rankOf = keptPos + NM;
    uint8_t* expd = reinterpret_cast<uint8_t*>(rankOf + NM);
    int* cellTmp = reinterpret_cast<int*>(smem + align_up((int)((uint8_t*)(expd + NM) - smem), 16));
    uint16_t* nodeSm = reinterpret_cast<uint16_t*>(cellTmp + 2 * P.maxCells);
    __shared__ int sM, sE, sFinish, sPhase2, sNToExpand, sN, sCut;

    const int level = blockIdx.x, slot = P.slot0 + blockIdx.y;
    const LevelGeom g = P.lg[level];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = g.nfeat;
    uint32_t* cand = P.cand + (long long)slot * P.candTotal + g.candOff;
    uint32_t* lkp = P.lkp + (long long)slot * P.kpStageTotal + g.kpOff;

    // ---- gather the cell slots into the contiguous, reference-ordered candidate list ----------------
    int* cellCnt = cellTmp; int* cellOff = cellTmp + P.maxCells;
    for (int c = tid; c < g.nCells; c += QT_THREADS) cellCnt[c] = P.cellCount[(long long)slot * P.nCellsTotal + g.cell0 + c];
    __syncthreads();
    if (warp == 0) { const int t = scan_by_warp0(cellCnt, cellOff, g.nCells, lane); if (lane == 0) sN = t; }
    __syncthreads();
    const int n = sN;
    // one cell per thread: a warp per cell walked ~40 cells one after the other, each a chain of dependent global loads
    // (descriptor -> candidates -> store); here every thread has at most two such chains and its loads are independent
    for (int c = tid; c < g.nCells; c += QT_THREADS) {
        const int k = cellCnt[c], o = cellOff[c];
        const uint32_t* src = P.cellCand + (long long)slot * P.cellSlotTotal + P.cells[g.cell0 + c].slotOff;
        int i = 0;
        for (; i + 4 <= k; i += 4) {
            const uint32_t a0 = src[i], a1 = src[i + 1], a2 = src[i + 2], a3 = src[i + 3];
            cand[o + i] = a0; cand[o + i + 1] = a1; cand[o + i + 2] = a2; cand[o + i + 3] = a3;
        }
        for (; i < k; ++i) cand[o + i] = src[i];
    }
    if (tid == 0) P.candCount[slot * ORBF_MAX_LEVELS + level] = n;
    if (n == 0) { if (tid == 0) P.lkpCount[slot * ORBF_MAX_LEVELS + level] = 0; return; }
    uint16_t* nodeOf = (n <= QT_SMEM_KEYS) ? nodeSm : (P.nodeScratch + (long long)slot * P.candTotal + g.candOff);
    __syncthreads();

    // ---- roots (orbextractor.cpp:470-506) -------------------------------------------------------------
    const int maxYrel = g.h - 2 * ORBF_MINB;          // maxBorderY - minBorderY
    for (int r = tid; r < g.nIni; r += QT_THREADS) keptTmp[r] = 0;
    __syncthreads();
    for (int k = tid; k < n; k += QT_THREADS) {
        int r = (int)__fdiv_rn((float)(cand[k] & 0x7FF), g.hX);
        r = min(r, g.nIni - 1);
        nodeOf[k] = (uint16_t)r;
        atomicAdd(&keptTmp[r], 1);
    }
    __syncthreads();
    if (tid == 0) {
        int m = 0;
        for (int r = 0; r < g.nIni; ++r) {
            keptPos[r] = (uint16_t)m;
            if (keptTmp[r] > 0) {
                rect0[m] = make_short4((short)(int)(g.hX * (float)r), 0, (short)(int)(g.hX * (float)(r + 1)), (short)maxYrel);
                cnt0[m] = keptTmp[r];
                ++m;
            }
        }
        sM = m; sE = 0; sFinish = 0; sPhase2 = 0;
    }
    __syncthreads();
    for (int k = tid; k < n; k += QT_THREADS) nodeOf[k] = keptPos[nodeOf[k]];
    __syncthreads();

    short4* rc = rect0; short4* rn = rect1; int* cc_ = cnt0; int* cn = cnt1;
    // ---- rounds -----------------------------------------------------------------------------------------
    while (true) {
        const int m = sM, Efresh = sE, phase2 = sPhase2;
        // 1. child counts of every node this round may split
        for (int i = tid; i < m; i += QT_THREADS) cc[i] = make_int4(0, 0, 0, 0);
        if (tid == 0) sNToExpand = 0;
        __syncthreads();
        for (int k = tid; k < n; k += QT_THREADS) {
            const int nd = nodeOf[k];
            if (cc_[nd] > 1 && (!phase2 || nd < Efresh)) atomicAdd(reinterpret_cast<int*>(&cc[nd]) + quadrant(cand[k], rc[nd]), 1);
        }
        __syncthreads();
        // 2. which nodes are split, and in which order (processing rank)
        int nCandNodes = 0;
        if (!phase2) {
            for (int i = tid; i < m; i += QT_THREADS) keptTmp[i] = cc_[i] > 1 ? 1 : 0;
            __syncthreads();
            if (warp == 0) { const int t = scan_by_warp0(keptTmp, baseByRank, m, lane); if (lane == 0) sCut = t; }
            __syncthreads();
            nCandNodes = sCut;
            for (int i = tid; i < m; i += QT_THREADS) {
                const bool e = cc_[i] > 1;
                expd[i] = e;
                if (e) rankOf[i] = (uint16_t)baseByRank[i];
            }
            __syncthreads();
        } else {
            // candidates: fresh multi-key nodes at positions [0, Efresh); order: count desc, then position asc
            for (int i = tid; i < m; i += QT_THREADS) {
                expd[i] = 0;
                if (i < Efresh && cc_[i] > 1) {
                    const int ci = cc_[i];
                    int r = 0;
                    for (int j = 0; j < Efresh; ++j) { const int cj = cc_[j]; r += (cj > 1) && (cj > ci || (cj == ci && j < i)); }
                    rankOf[i] = (uint16_t)r;
                }
            }
            for (int i = tid; i < m; i += QT_THREADS) keptTmp[i] = (i < Efresh && cc_[i] > 1) ? 1 : 0;
            __syncthreads();
            if (warp == 0) { const int t = scan_by_warp0(keptTmp, baseByRank, m, lane); if (lane == 0) sCut = t; }
            __syncthreads();
            nCandNodes = sCut;
        }
        // 3. non-empty children per rank, prefix sums (creation index base), phase-2 cut
        for (int i = tid; i < m; i += QT_THREADS) {
            if (cc_[i] > 1 && (!phase2 || i < Efresh)) {
                const int4 c4 = cc[i];
                nchByRank[rankOf[i]] = (c4.x > 0) + (c4.y > 0) + (c4.z > 0) + (c4.w > 0);
            }
        }
        __syncthreads();
        if (phase2) {
            if (warp == 0) {
                // size after processing ranks 0..r = m + sum(nch - 1); first r reaching N ends the round
                int carry = m, cut = nCandNodes - 1;
                bool found = false;
                for (int b = 0; b < nCandNodes && !found; b += 32) {
                    const int i = b + lane;
                    const int v = (i < nCandNodes) ? nchByRank[i] - 1 : 0;
                    int tot;
                    const int e = warp_excl_scan(v, lane, tot);
                    const bool hit = (i < nCandNodes) && (carry + e + v >= N);
                    const unsigned mk = __ballot_sync(0xffffffffu, hit);
                    if (mk) { cut = b + __ffs(mk) - 1; found = true; }
                    carry += tot;
                }
                if (lane == 0) sCut = cut;
            }
            __syncthreads();
            const int cut = sCut;
            for (int i = tid; i < m; i += QT_THREADS) expd[i] = (i < Efresh && cc_[i] > 1 && rankOf[i] <= cut) ? 1 : 0;
            nCandNodes = (nCandNodes > 0) ? cut + 1 : 0;
            __syncthreads();
        }
        if (warp == 0) { const int t = scan_by_warp0(nchByRank, baseByRank, nCandNodes, lane); if (lane == 0) sE = t; }
        for (int i = tid; i < m; i += QT_THREADS) keptTmp[i] = expd[i] ? 0 : 1;
        __syncthreads();
        const int E = sE;
        if (warp == 0) { const int t = scan_by_warp0(keptTmp, keptTmp, m, lane); if (lane == 0) sCut = t; }
        __syncthreads();
        const int newM = E + sCut;
        // 4. write the next node table
        int localExpand = 0;
        for (int i = tid; i < m; i += QT_THREADS) {
            if (expd[i]) {
                const short4 r = rc[i];
                const int4 c4 = cc[i];
                const int hx = (r.z - r.x + 1) >> 1, hy = (r.w - r.y + 1) >> 1;
                int c = baseByRank[rankOf[i]];
                ushort4 cp = make_ushort4(0, 0, 0, 0);
                if (c4.x > 0) { const int p = E - 1 - c++; cp.x = (unsigned short)p; rn[p] = make_short4(r.x, r.y, (short)(r.x + hx), (short)(r.y + hy)); cn[p] = c4.x; localExpand += c4.x > 1; }
                if (c4.y > 0) { const int p = E - 1 - c++; cp.y = (unsigned short)p; rn[p] = make_short4((short)(r.x + hx), r.y, r.z, (short)(r.y + hy)); cn[p] = c4.y; localExpand += c4.y > 1; }
                if (c4.z > 0) { const int p = E - 1 - c++; cp.z = (unsigned short)p; rn[p] = make_short4(r.x, (short)(r.y + hy), (short)(r.x + hx), r.w); cn[p] = c4.z; localExpand += c4.z > 1; }
                if (c4.w > 0) { const int p = E - 1 - c++; cp.w = (unsigned short)p; rn[p] = make_short4((short)(r.x + hx), (short)(r.y + hy), r.z, r.w); cn[p] = c4.w; localExpand += c4.w > 1; }
                childPos[i] = cp;
            } else {
                const int p = E + keptTmp[i];
                keptPos[i] = (uint16_t)p;
                rn[p] = rc[i]; cn[p] = cc_[i];
            }
        }
        if (localExpand) atomicAdd(&sNToExpand, localExpand);
        __syncthreads();
        // 5. move the keys
        for (int k = tid; k < n; k += QT_THREADS) {
            const int nd = nodeOf[k];
            if (expd[nd]) {
                const ushort4 cp = childPos[nd];
                const int q = quadrant(cand[k], rc[nd]);
                nodeOf[k] = (q == 0) ? cp.x : (q == 1) ? cp.y : (q == 2) ? cp.z : cp.w;
            } else nodeOf[k] = keptPos[nd];
        }
        __syncthreads();
        // 6. termination rules (orbextractor.cpp:581-639)
        if (tid == 0) {
            const int prevSize = m;
            sM = newM;
            if (newM >= N || newM == prevSize) sFinish = 1;
            else if (!phase2 && newM + sNToExpand * 3 > N) sPhase2 = 1;
        }
        { short4* t = rc; rc = rn; rn = t; int* u = cc_; cc_ = cn; cn = u; }
        __syncthreads();
        if (sFinish) break;
    }
    // ---- best key per leaf, emitted in list order ----------------------------------------------------
    const int m = sM;
    uint32_t* best = reinterpret_cast<uint32_t*>(cc);
    for (int i = tid; i < m; i += QT_THREADS) best[i] = 0;
    __syncthreads();
    for (int k = tid; k < n; k += QT_THREADS) atomicMax(&best[nodeOf[k]], ((cand[k] >> 22) << 24) | (0xFFFFFFu - (uint32_t)k));
    __syncthreads();
    for (int i = tid; i < m; i += QT_THREADS) lkp[i] = cand[0xFFFFFFu - (best[i] & 0xFFFFFFu)];
    if (tid == 0) P.lkpCount[slot * ORBF_MAX_LEVELS + level] = m;
}

}  // namespace

int orbf_launch_quadtree(orbf_context* c, int slot0, int n)
{
    int NM = 0, maxCells = 0;
    for (int l = 0; l < c->L; ++l) { NM = std::max(NM, c->lg[l].kpCap + 8); maxCells = std::max(maxCells, c->lg[l].nCells); }
    NM = align_up(NM, 8);
    size_t smem = (size_t)NM * (8 + 8 + 4 + 4 + 16 + 8 + 4 + 4 + 4 + 2 + 2 + 1) + 32;
    smem += (size_t)2 * maxCells * sizeof(int) + (size_t)QT_SMEM_KEYS * sizeof(uint16_t);
    if (smem > 200 * 1024) return ORBF_ERR_GEOMETRY;
    cudaError_t e = cudaFuncSetAttribute(quadtree_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return orbf_cuda_fail(c, e, "quadtree smem attr", __FILE__, __LINE__);
    QtParams P;
    P.lg = c->d_lg; P.cells = c->d_cells; P.cellCand = c->d_cellCand; P.cellCount = c->d_cellCount;
    P.cand = c->d_cand; P.candCount = c->d_candCount; P.nodeScratch = c->d_nodeScratch; P.lkp = c->d_lkp; P.lkpCount = c->d_lkpCount;
    P.cellSlotTotal = c->cellSlotTotal; P.nCellsTotal = c->nCellsTotal; P.candTotal = c->candTotal; P.kpStageTotal = c->kpStageTotal;
    P.slot0 = slot0; P.NM = NM; P.maxCells = maxCells;
    dim3 grid(c->L, n);
    quadtree_kernel<<<grid, QT_THREADS, smem, c->stream>>>(P);
    ORBF_LAUNCH_CHECK(c);
    return ORBF_OK;
}
