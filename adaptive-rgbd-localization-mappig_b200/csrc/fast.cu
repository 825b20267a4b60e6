// csrc/fast.cu — per-cell FAST-9/16 with non-max suppression and the iniTh -> minTh fallback
// (reference ComputeKeyPointsOctTree, Features/orbextractor.cpp:665-723, calling cv::FAST(roi, th, true)).
//
// Work unit = a STRIP of up to 4 adjacent cells of one cell row (their scored interiors tile the row without gaps),
// one CTA of 4 warps per (strip, frame), all levels in one launch:
//   0. one thread fetches the strip + 3-px ring halo with a single TMA box load (cp.async.bulk.tensor.3d -> UTMALDG);
//      a u8 box must start on a 16-byte boundary of the level row, so the interior begins at tile column ax in [3, 18];
//   1. pretest, 4 pixels per thread, byte-parallel: a 9-arc contains one pixel of each opposite ring pair, so for the four
//      pairs (0,8) (4,12) (2,10) (6,14) at least one member must differ from the centre by more than th.  |r - v| for 4
//      pixels is one VABSDIFF4, "> th" a carry trick into bit 7 of each byte; 8 % of the pixels of the bench texture
//      survive (true corners: 3.5 %) and are appended to a CTA-wide list (shared-memory atomics);
//   2. corner strength of two survivors at a time on packed u16x2 lanes with 3-input min/max (VIMNMX3.U16x2):
//      S = max(v - A, B - v), A = min over the 16 arcs of the arc maximum, B = max over the arcs of the arc minimum
//      (cv::FAST response = S - 1, corner iff S > th); 80 packed ops per pixel pair;
//   3. NMS over the survivor list: strict '>' against the 8 neighbours *inside the same cell interior* (quirk Q1: each
//      cell is its own cv::FAST call, so neighbours in an adjacent cell count as 0); kept corners go to their cell's list;
//   4. one warp per cell ranks its kept corners by tile offset (= row-major order inside the cell) and writes them to
//      the cell's private slot range in exactly the reference's push_back order;
//   5. a cell with no survivor at iniTh is redone by its warp at minTh (orbextractor.cpp:709-712; rare, scalar path).
// Arithmetic is all integer min/max/compare: bit-exact by construction.  Bound by the integer ALU pipe, not by HBM.
#include "orbf_internal.h"
#include "fast_device.h"

namespace {

constexpr int FS_WARPS = ORBF_STRIP_CELLS, FS_THREADS = FS_WARPS * 32;

struct FastParams {
    CUtensorMap maps[ORBF_MAX_LEVELS];
    const StripDesc* strips; const CellDesc* cells;
    uint32_t* cellCand; int* cellCount;
    int cellSlotTotal, nCellsTotal, iniTh, minTh, slot0, z0;   // z coordinate of a slot: slot - z0 on level 0, slot elsewhere
    int keptCap;                                                // capacity of a cell's kept-corner list (max CellDesc::cap)
    short BW[ORBF_MAX_LEVELS], BH[ORBF_MAX_LEVELS];
};

// strict 8-neighbour maximum test of score byte *s (value v > 0); l / r: a left / right neighbour column exists in this cell
__device__ __forceinline__ bool nms_keep(const uint8_t* s, int v, int BW, bool l, bool r)
{
    bool k = v > s[-BW] && v > s[BW];
    if (l) k = k && v > s[-1] && v > s[-BW - 1] && v > s[BW - 1];
    if (r) k = k && v > s[1] && v > s[-BW + 1] && v > s[BW + 1];
    return k;
}

// One warp orders a cell's kept corners (by tile offset; offsets are distinct) and writes them out with their response.
__device__ __forceinline__ void emit_cell(const uint16_t* kept, const uint8_t* sorg, int n, int BW, int c0, uint32_t* out, int outX0, int outY0, int lane)
{
    for (int i = lane; i < n; i += 32) {
        const uint32_t e = kept[i];
        int rank = 0;
        for (int j = 0; j < n; ++j) rank += kept[j] < e;
        const int row = (int)e / BW, col = (int)e - row * BW;
        out[rank] = (uint32_t)(outX0 + col - c0) | ((uint32_t)(outY0 + row) << 11) | ((uint32_t)sorg[e] << 22);
    }
}

__global__ void __launch_bounds__(FS_THREADS) fast_strip_kernel(const __grid_constant__ FastParams P)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ int sCount, sCorner;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const StripDesc sd = P.strips[blockIdx.x];
    const int slot = P.slot0 + blockIdx.y;
    constexpr int BW = ORBF_FAST_BW;                                       // compile-time pitch: every tile offset is an immediate
    const int level = sd.level, BH = P.BH[level];
    const int W = sd.w, h = sd.h;
    const int xs = (sd.x0 - 3) & ~15, ax = sd.x0 - xs;                      // TMA boxes of bytes start on 16-byte boundaries
    uint8_t* tile = smem;                                                   // BH x BW level pixels, interior (0,0) at [3][ax]
    uint8_t* score = smem + align_up(BW * BH, 128);                         // (h + 2) x BW responses, interior (0,0) at [1][ax], zero elsewhere
    uint16_t* list = reinterpret_cast<uint16_t*>(score + align_up(BW * (BH - 4), 16));   // pretest survivors: y * BW + tile column
    uint16_t* kept = list + align_up(ORBF_STRIP_MAX_W * (BH - 6), 8);                    // [cell][keptCap] NMS survivors (tile offsets)
    __shared__ int sKept[FS_WARPS];
    __shared__ int sC0[FS_WARPS + 1];                                       // first tile column of each cell, then the strip's end

    if (tid == 0) { mbar_init(&bar, 1); sCount = 0; sCorner = 0; }
    if (tid < FS_WARPS) sKept[tid] = 0;
    if (tid <= FS_WARPS) sC0[tid] = tid < sd.nCells ? ax + P.cells[sd.firstCell + tid].x0 - sd.x0 : ax + W + (tid > sd.nCells ? 4096 : 0);
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&bar, (uint32_t)(BW * BH));
        tma_load_3d(tile, &P.maps[level], xs, sd.y0 - 3, level == 0 ? slot - P.z0 : slot, &bar);
    }
    for (int i = tid; i < (BW * (h + 2)) >> 2; i += FS_THREADS) reinterpret_cast<uint32_t*>(score)[i] = 0;
    mbar_wait(&bar, 0);
    __syncthreads();

    const int th = P.iniTh;
    const int keptCap = P.keptCap;
    // ---- 1. pretest: 4 pixels (one aligned word of the centre row) per thread -------------------------------------
    {
        const int wFirst = ax >> 2, wpr = ((ax + W + 3) >> 2) - wFirst, nTasks = h * wpr;
        const uint32_t rcp = ((1u << 20) + wpr - 1) / wpr;
        // per byte x = |r - v|: bit 7 of ((x & 0x7f) + K) | x (th < 128, K = 127 - th) or of ((x & 0x7f) + K) & x (th >= 128,
        // K = 255 - th) is set iff x > th
        const bool hiTh = th >= 128;
        const uint32_t K = (uint32_t)(hiTh ? 255 - th : 127 - th) * 0x01010101u, M7 = 0x7F7F7F7Fu;
        for (int t = tid; t < nTasks; t += FS_THREADS) {
            const int row = (int)(((uint32_t)t * rcp) >> 20), wi = wFirst + (t - row * wpr);
            const uint32_t* c = reinterpret_cast<const uint32_t*>(tile + (row + 3) * BW) + wi;
            const uint32_t C = c[0];
            const uint32_t* u2 = reinterpret_cast<const uint32_t*>(tile + (row + 1) * BW) + wi;     // image row y - 2
            const uint32_t* d2 = reinterpret_cast<const uint32_t*>(tile + (row + 5) * BW) + wi;     // image row y + 2
            uint32_t ad[8];
            ad[0] = __vabsdiffu4(*reinterpret_cast<const uint32_t*>(tile + (row + 6) * BW + 4 * wi), C);     // ring 0  ( 0, +3)
            ad[1] = __vabsdiffu4(*reinterpret_cast<const uint32_t*>(tile + row * BW + 4 * wi), C);           // ring 8  ( 0, -3)
            ad[2] = __vabsdiffu4(__funnelshift_r(C, c[1], 24), C);                                           // ring 4  (+3,  0)
            ad[3] = __vabsdiffu4(__funnelshift_r(c[-1], C, 8), C);                                           // ring 12 (-3,  0); word -1 of tile row >= 3 is inside the tile
            ad[4] = __vabsdiffu4(__funnelshift_r(d2[0], d2[1], 16), C);                                      // ring 2  (+2, +2)
            ad[5] = __vabsdiffu4(__funnelshift_r(u2[-1], u2[0], 16), C);                                     // ring 10 (-2, -2)
            ad[6] = __vabsdiffu4(__funnelshift_r(u2[0], u2[1], 16), C);                                      // ring 6  (+2, -2)
            ad[7] = __vabsdiffu4(__funnelshift_r(d2[-1], d2[0], 16), C);                                     // ring 14 (-2, +2)
            uint32_t all = 0x80808080u;
#pragma unroll
            for (int k = 0; k < 8; k += 2) {
                const uint32_t ta = (ad[k] & M7) + K, tb = (ad[k + 1] & M7) + K;
                all &= hiTh ? ((ta & ad[k]) | (tb & ad[k + 1])) : (ta | tb | ad[k] | ad[k + 1]);
            }
            const int xi = 4 * wi - ax;                                     // interior x of byte 0: mask pixels outside [0, W)
            if (xi < 0) all &= 0xFFFFFFFFu << (8 * (-xi));
            if (W - xi < 4) all &= 0xFFFFFFFFu >> (8 * (4 - (W - xi)));
            if (all) {
                int pos = atomicAdd(&sCount, __popc(all));
                const int e = row * BW + 4 * wi;
                if (all & 0x80u) list[pos++] = (uint16_t)e;
                if (all & 0x8000u) list[pos++] = (uint16_t)(e + 1);
                if (all & 0x800000u) list[pos++] = (uint16_t)(e + 2);
                if (all & 0x80000000u) list[pos] = (uint16_t)(e + 3);
            }
        }
    }
    __syncthreads();
    const int n = sCount;
    const uint8_t* org = tile + 3 * BW;
    uint8_t* sorg = score + BW;
    // ---- 2. corner strength, two survivors per thread; the corners (44 % of the survivors) are compacted in place at the front
    // of the list so that the NMS pass runs with full warps.  A round reads 2 * FS_THREADS entries before anything is appended,
    // and the appends of round k stay below the entries consumed so far, so the only hazard is inside a round (the barrier).
    for (int base = 0; base < n; base += 2 * FS_THREADS) {
        const int ia = base + 2 * tid;
        int ea = 0, eb = 0, sa = 0, sb = 0;
        if (ia < n) {
            ea = list[ia]; eb = list[min(ia + 1, n - 1)];
            const uint32_t s = ring_strength_x2(org + ea, org + eb, BW);
            sa = (int)(s & 0xFFFFu); sb = ia + 1 < n ? (int)(s >> 16) : 0;
            if (sa > th) sorg[ea] = (uint8_t)(sa - 1);
            if (sb > th) sorg[eb] = (uint8_t)(sb - 1);
        }
        __syncthreads();
        const int na = sa > th, nb = sb > th;
        if (na + nb) {
            int pos = atomicAdd(&sCorner, na + nb);
            if (na) list[pos++] = (uint16_t)ea;
            if (nb) list[pos] = (uint16_t)eb;
        }
    }
    __syncthreads();
    // ---- 3. NMS over the corner list, kept corners appended to their cell's list -------------------------------------
    {
        const int nc = sCorner;
        const int b1 = sC0[1], b2 = sC0[2], b3 = sC0[3];
        for (int i = tid; i < nc; i += FS_THREADS) {
            const int e = list[i];
            const int v = sorg[e];
            const int row = e / BW, col = e - row * BW;
            const int k = (col >= b1) + (col >= b2) + (col >= b3);
            if (nms_keep(sorg + e, v, BW, col > sC0[k], col < sC0[k + 1] - 1)) {
                const int pos = atomicAdd(&sKept[k], 1);
                if (pos < keptCap) kept[k * keptCap + pos] = (uint16_t)e;
            }
        }
    }
    __syncthreads();
    // ---- 4./5. one warp per cell: ordered output, minTh fallback ------------------------------------------------------
    if (warp < sd.nCells) {
        const int cellIdx = sd.firstCell + warp;
        const CellDesc cd = P.cells[cellIdx];
        const int c0 = sC0[warp], cw = cd.w;
        uint32_t* out = P.cellCand + (long long)slot * P.cellSlotTotal + cd.slotOff;
        const int outX0 = cd.x0 + cd.relx, outY0 = cd.y0 + cd.rely;
        uint16_t* myKept = kept + warp * keptCap;
        int total = min(sKept[warp], keptCap);
        if (total == 0 && P.minTh < th) {
            // no corner at iniTh in this cell, so its score columns are still all zero: rescore the cell at minTh
            const int t2 = P.minTh, npix = cw * h;
            const uint32_t rcpW = ((1u << 20) + cw - 1) / cw;
            for (int p = lane; p < npix; p += 32) {
                const int y = (int)(((uint32_t)p * rcpW) >> 20), x = p - y * cw;
                const uint8_t* c = org + y * BW + c0 + x;
                const int v = c[0], hiT = v + t2, loT = v - t2;
                const int p0 = c[3 * BW], p8 = c[-3 * BW], p4 = c[3], p12 = c[-3];
                const bool pass = (((p0 > hiT) | (p8 > hiT)) & ((p4 > hiT) | (p12 > hiT))) | (((p0 < loT) | (p8 < loT)) & ((p4 < loT) | (p12 < loT)));
                if (pass) {
                    const int s = ring_strength(c, BW);
                    if (s > t2) sorg[y * BW + c0 + x] = (uint8_t)(s - 1);
                }
            }
            __syncwarp();
            for (int p = lane; p < npix; p += 32) {
                const int y = (int)(((uint32_t)p * rcpW) >> 20), x = p - y * cw;
                const int e = y * BW + c0 + x;
                const int v = sorg[e];
                if (v && nms_keep(sorg + e, v, BW, x > 0, x < cw - 1)) {
                    const int pos = atomicAdd(&sKept[warp], 1);
                    if (pos < keptCap) myKept[pos] = (uint16_t)e;
                }
            }
            __syncwarp();
            total = min(sKept[warp], keptCap);
        }
        emit_cell(myKept, sorg, total, BW, c0, out, outX0, outY0, lane);
        if (lane == 0) P.cellCount[(long long)slot * P.nCellsTotal + cellIdx] = total;
    }
}

}  // namespace

int orbf_launch_fast(orbf_context* c, int slot0, int n)
{
    {
        const int r = orbf_refresh_maps(c);
        if (r != ORBF_OK) return r;
    }
    FastParams P;
    size_t smem = 0;
    const int keptCap = ((c->maxCellW + 1) / 2) * ((c->maxCellH + 1) / 2);     // strict 8-neighbour maxima: <= 1 per 2x2 block
    P.keptCap = keptCap;
    for (int l = 0; l < c->L; ++l) {
        P.maps[l] = c->tmFast[l];
        P.BW[l] = (short)c->fastBW[l]; P.BH[l] = (short)c->fastBH[l];
        const int BW = c->fastBW[l], BH = c->fastBH[l];
        // 27.2 KB at 640x480: 8 CTAs per SM (the kernel gains ~6 % per extra resident CTA at this point)
        const size_t need = (size_t)align_up(BW * BH, 128) + align_up(BW * (BH - 4), 16) + (size_t)align_up(ORBF_STRIP_MAX_W * (BH - 6), 8) * sizeof(uint16_t)
            + (size_t)FS_WARPS * keptCap * sizeof(uint16_t) + 16;
        smem = std::max(smem, need);
    }
    P.strips = c->d_strips; P.cells = c->d_cells; P.cellCand = c->d_cellCand; P.cellCount = c->d_cellCount;
    P.cellSlotTotal = c->cellSlotTotal; P.nCellsTotal = c->nCellsTotal; P.iniTh = c->cfg.ini_th_fast; P.minTh = c->cfg.min_th_fast;
    P.slot0 = slot0; P.z0 = c->cur_slot0;
    if (smem > 200 * 1024) return ORBF_ERR_GEOMETRY;
    {   // static + dynamic shared memory can exceed the 48 KB default while the dynamic part alone does not: always opt in
        cudaError_t e = cudaFuncSetAttribute(fast_strip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return orbf_cuda_fail(c, e, "fast smem attr", __FILE__, __LINE__);
    }
    dim3 grid(c->nStrips, n);
    fast_strip_kernel<<<grid, FS_THREADS, smem, c->stream>>>(P);
    ORBF_LAUNCH_CHECK(c);
    return ORBF_OK;
}
