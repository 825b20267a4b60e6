// csrc/fast.cu — per-cell FAST-9/16 with non-max suppression and the iniTh -> minTh fallback
// (reference ComputeKeyPointsOctTree, Features/orbextractor.cpp:665-723, calling cv::FAST(roi, th, true)).
//
// Work unit = a STRIP of up to 4 adjacent cells of one cell row (their scored interiors tile the row without gaps),
// one CTA of 4 warps per (strip, frame), all levels in one launch:
//   0. one thread fetches the strip + 3-px ring halo with a single TMA box load (cp.async.bulk.tensor.3d -> UTMALDG);
//      a u8 box must start on a 16-byte boundary of the level row, so the interior begins at tile column ax in [3, 18];
//   1. pretest, 8 pixels per thread, byte-parallel: a 9-arc contains one pixel of each opposite ring pair, so for the four
//      pairs (0,8) (4,12) (2,10) (6,14) at least one member must differ from the centre by more than th.  |r - v| for 4
//      pixels is one VABSDIFF4, "> th" a carry trick into bit 7 of each byte; 8 % of the pixels of the bench texture
//      survive (true corners: 3.5 %) and are appended to a CTA-wide list of bounded size;
//   2. corner strength of two survivors at a time on packed u16x2 lanes with 3-input min/max (VIMNMX3.U16x2):
//      S = max(v - A, B - v), A = min over the 16 arcs of the arc maximum, B = max over the arcs of the arc minimum
//      (cv::FAST response = S - 1, corner iff S > th); 80 packed ops per pixel pair.  Responses go to a score plane in which
//      the cells of the strip are one zero column apart, so that
//   3. NMS over the corner list is eight unconditional strict compares (quirk Q1: each cell is its own cv::FAST call, so
//      neighbours in an adjacent cell count as 0); a kept corner sets its bit in a per-row bitmap;
//   4. one warp per cell turns the cell's bits into the reference's push_back order (row-major inside the cell) with a
//      popcount prefix over the rows — no sorting — and writes the cell's private slot range;
//   5. cells left empty after NMS are redone at minTh (orbextractor.cpp:709-712) through the same vector path, restricted to
//      those cells.  A strip whose survivors exceed the list (noise-like content) takes the dense scalar path instead.
// Arithmetic is all integer min/max/compare: bit-exact by construction.  Not HBM-bound: the binding resource is the shared-memory
// data pipe (ncu: 81 % of the peak wavefront rate, 46 % of the wavefronts are the ring gathers of step 2 at ~2.2 wavefronts per
// warp-wide byte gather), with instruction issue close behind (71 %) and the CTA barriers between the phases as the top stall reason —
// profiles/r2i_fast_kernel.txt, r2i_fast_phases.txt.  Measured and not kept: 16-pixel pretest tasks with LDS.128 rows (12 instead of 24
// load instructions per 16 pixels, ~28 % fewer pretest wavefronts): 0.9038 vs 0.9019 ms — the pretest is not what the pipe waits for.
#include "orbf_internal.h"
#include "fast_device.h"

// tuning switches (profiles/r2*_fast_variants.txt has the measurements behind the defaults)
#ifndef FV_APPEND_COOP
#define FV_APPEND_COOP 0        // pretest append: 1 = warp prefix + one atomic per warp, 0 = one atomic per task with survivors
#endif
#ifndef FV_CORNER_COOP
#define FV_CORNER_COOP 1        // corner compaction in the strength rounds: 1 = ballot-aggregated, 0 = one atomic per thread
#endif
#ifndef FV_CELLTAB
#define FV_CELLTAB 0            // cell index of a column: 1 = shared-memory table, 0 = three compares
#endif
#ifndef FV_LB
#define FV_LB 10                // __launch_bounds__ minimum CTAs per SM (0 = unconstrained)
#endif

namespace {

constexpr int FS_WARPS = ORBF_STRIP_CELLS, FS_THREADS = FS_WARPS * 32;
constexpr int BW = ORBF_FAST_BW;                 // tile pitch (= TMA box width): every tile offset is an immediate
constexpr int SP = BW + FS_WARPS;                // score-plane pitch: cell k of the strip is shifted right by k columns
constexpr int BITW = (SP + 31) / 32;             // words of a row of the kept-corner bitmap
constexpr int LIST_CAP = 2048;                   // pretest survivors a strip may hold (u16 tile offsets)

struct FastParams {
    CUtensorMap maps[ORBF_MAX_LEVELS];
    const StripDesc* strips; const CellDesc* cells;
    uint32_t* cellCand; int* cellCount;
    int cellSlotTotal, nCellsTotal, iniTh, minTh, slot0, z0;   // z coordinate of a slot: slot - z0 on level 0, slot elsewhere
    int scoreOff, listOff, bitsOff, sbmOff, sbmWords;           // byte offsets of the shared-memory regions; words of the survivor bitmap
    // region-adapted variant (orbf_extract_adapted): iniTh of a cell = regionTh[cellRegion[cell]]; NULL => iniTh everywhere
    const uint8_t* cellRegion; const int* regionTh; int regionThStride;   // stride: entries between the tables of consecutive frames (videos) of the launch, or 0
    short BH[ORBF_MAX_LEVELS];
};

struct Strip { int W, h, ax; };      // scored width / rows of the strip, tile column of its first scored pixel

// ---- 1. pretest over the strip interior; survivors (tile offsets row * BW + column) are appended to list -----------------------
// sMask[w] = 0x80 in every byte of tile word w (columns 4w .. 4w + 3) that lies inside the scored interior — the start value of
// the flag chain, so that pixels outside [ax, ax + W) never survive (no per-task edge arithmetic).
// onlyEmpty: keep a survivor only if its cell has no kept corner yet (minTh pass)
// cell index of tile column col
__device__ __forceinline__ int cell_of(const uint8_t* sCellOf, const int* sC0, int col)
{
#if FV_CELLTAB
    return sCellOf[col];
#else
    return (col >= sC0[1]) + (col >= sC0[2]) + (col >= sC0[3]);
#endif
}

__device__ __forceinline__ void pretest(const uint8_t* tile, uint8_t* sbm, const Strip& S, int th, bool onlyEmpty, const uint32_t* sMask,
    const uint8_t* sCellOf, const int* sC0, const int* sHas, int tid)
{
    const int wFirst = S.ax >> 3, wpr = ((S.ax + S.W + 7) >> 3) - wFirst, nTasks = S.h * wpr;
    const uint32_t rcp = ((1u << 20) + wpr - 1) / wpr;
    // per byte x = |r - v| (th < 128, K = 127 - th in every byte): bit 7 of (x + K) | x is set if x > th.  The add runs over the whole
    // word: a byte with x >= 129 + th carries into its left neighbour, which can only turn a neighbour with x == th into a false
    // positive (the exact strength decides later), never drop a corner: x + K overflows the byte only when x >= 129, and then bit 7 of
    // x itself is set.  th >= 128 never reaches this function (the kernel takes the dense path).
    const uint32_t K = (uint32_t)(127 - th) * 0x01010101u;
    for (int t = tid; t < nTasks; t += FS_THREADS) {
        uint32_t all[2];
        int e;
        {
            const int row = (int)(((uint32_t)t * rcp) >> 20), x8 = 8 * (wFirst + (t - row * wpr));
            e = row * BW + x8;
            const uint8_t* p = tile + e;                                                 // image row y - 3 of the task's 8 columns
            const uint2 r0 = *reinterpret_cast<const uint2*>(p + 6 * BW);                // ring 0  ( 0, +3)
            const uint2 r8 = *reinterpret_cast<const uint2*>(p);                         // ring 8  ( 0, -3)
            const uint2 c = *reinterpret_cast<const uint2*>(p + 3 * BW);
            const uint32_t cl = *reinterpret_cast<const uint32_t*>(p + 3 * BW - 4), cr = *reinterpret_cast<const uint32_t*>(p + 3 * BW + 8);
            const uint2 u = *reinterpret_cast<const uint2*>(p + BW);                      // image row y - 2
            const uint32_t ul = *reinterpret_cast<const uint32_t*>(p + BW - 4), ur = *reinterpret_cast<const uint32_t*>(p + BW + 8);
            const uint2 d = *reinterpret_cast<const uint2*>(p + 5 * BW);                  // image row y + 2
            const uint32_t dl = *reinterpret_cast<const uint32_t*>(p + 5 * BW - 4), dr = *reinterpret_cast<const uint32_t*>(p + 5 * BW + 8);
            const uint2 m0 = *reinterpret_cast<const uint2*>(sMask + (x8 >> 2));
#pragma unroll
            for (int w = 0; w < 2; ++w) {
                // the word's own 4 bytes, the word left of it (.L) and right of it (.R), on the three rows that need shifted neighbours
                const uint32_t C = w ? c.y : c.x, CL = w ? c.x : cl, CR = w ? cr : c.y;
                const uint32_t U = w ? u.y : u.x, UL = w ? u.x : ul, UR = w ? ur : u.y;
                const uint32_t D = w ? d.y : d.x, DL = w ? d.x : dl, DR = w ? dr : d.y;
                uint32_t ad[8];
                ad[0] = __vabsdiffu4(w ? r0.y : r0.x, C);
                ad[1] = __vabsdiffu4(w ? r8.y : r8.x, C);
                ad[2] = __vabsdiffu4(__funnelshift_r(C, CR, 24), C);                                             // ring 4  (+3,  0)
                ad[3] = __vabsdiffu4(__funnelshift_r(CL, C, 8), C);                                              // ring 12 (-3,  0)
                ad[4] = __vabsdiffu4(__funnelshift_r(D, DR, 16), C);                                             // ring 2  (+2, +2)
                ad[5] = __vabsdiffu4(__funnelshift_r(UL, U, 16), C);                                             // ring 10 (-2, -2)
                ad[6] = __vabsdiffu4(__funnelshift_r(U, UR, 16), C);                                             // ring 6  (+2, -2)
                ad[7] = __vabsdiffu4(__funnelshift_r(DL, D, 16), C);                                             // ring 14 (-2, +2)
                uint32_t a = w ? m0.y : m0.x;
#pragma unroll
                for (int k = 0; k < 8; k += 2) a &= (ad[k] + K) | (ad[k + 1] + K) | ad[k] | ad[k + 1];
                all[w] = a;
            }
            if (onlyEmpty && (all[0] | all[1])) {            // rare pass: drop survivors of cells that already have a kept corner
#pragma unroll
                for (int b = 0; b < 8; ++b)
                    if (sHas[cell_of(sCellOf, sC0, x8 + b)]) all[b >> 2] &= ~(0x80u << (8 * (b & 3)));
            }
        }
        // flag bytes -> one bit per pixel: byte (e >> 3) of the survivor bitmap, whose bit index is the tile offset
        sbm[e >> 3] = (uint8_t)(((((all[0] >> 7) * 0x00204081u) >> 21) & 0xFu) | ((((all[1] >> 7) * 0x00204081u) >> 17) & 0xF0u));
    }
}

// ---- 1b. survivor bitmap -> list of tile offsets in raster order.  Lanes of a warp then work on neighbouring pixels in the strength and
// NMS steps: their byte gathers fall into distinct banks far more often than with an arbitrary order (the shared-memory data pipe is what
// binds this kernel).  Every thread takes 64 consecutive bits; returns the survivor count (uniform).  Nothing is written when the count
// exceeds LIST_CAP.
__device__ __forceinline__ int compact_survivors(const uint32_t* sbm32, uint16_t* list, int nWords, int* sTot, int tid)
{
    const int lane = tid & 31, warp = tid >> 5;
    int total = 0;
    for (int w0 = 0; w0 < nWords; w0 += 2 * FS_THREADS) {
        const int w = w0 + 2 * tid;
        uint2 v = make_uint2(0u, 0u);
        if (w < nWords) v = *reinterpret_cast<const uint2*>(sbm32 + w);          // the bitmap is padded to an even word count
        const int cnt = __popc(v.x) + __popc(v.y);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) sTot[warp] = incl;
        __syncthreads();
        int pos = total + incl - cnt;
#pragma unroll
        for (int k = 0; k < FS_WARPS; ++k) { const int t = sTot[k]; if (k < warp) pos += t; total += t; }
        __syncthreads();                    // sTot is rewritten by the next round
        if (total <= LIST_CAP) {
            const int e0 = 32 * w;
            while (v.x) { const int b = __ffs((int)v.x) - 1; v.x &= v.x - 1; list[pos++] = (uint16_t)(e0 + b); }
            while (v.y) { const int b = __ffs((int)v.y) - 1; v.y &= v.y - 1; list[pos++] = (uint16_t)(e0 + 32 + b); }
        }
    }
    return total;
}

// ---- 2. corner strength, two survivors per thread; the corners (44 % of the survivors) are compacted in place at the front of
// the list — as score-plane positions — so that the NMS pass runs with full warps.  A round reads 2 * FS_THREADS entries before
// anything is appended, and the appends of round k stay below the entries consumed so far, so the only hazard is inside a round
// (the barrier).  Ends with a barrier: *sCorner is the corner count.
__device__ __forceinline__ void strengths(const uint8_t* org, uint8_t* sorg, uint16_t* list, int n, int* sCorner, const int* sTh, const uint8_t* sCellOf, const int* sC0, int tid)
{
    const int lane = tid & 31;
#if !FV_CELLTAB
    const int b1 = sC0[1], b2 = sC0[2], b3 = sC0[3];
#endif
    for (int base = 0; base < n; base += 2 * FS_THREADS) {
        const int ia = base + 2 * tid;
        int pa = 0, pb = 0, sa = 0, sb = 0, tha = 255, thb = 255;
        if (ia < n) {
            const int ea = list[ia], eb = list[min(ia + 1, n - 1)];
            const uint32_t s = ring_strength_x2(org + ea, org + eb, BW);
            sa = (int)(s & 0xFFFFu); sb = ia + 1 < n ? (int)(s >> 16) : 0;
            const int ra = ea / BW, rb = eb / BW;                        // score-plane position = tile offset + (SP - BW) * row + cell index
#if FV_CELLTAB
            const int ka = sCellOf[ea - ra * BW], kb = sCellOf[eb - rb * BW];
            pa = ea + (SP - BW) * ra + ka;
            pb = eb + (SP - BW) * rb + kb;
            tha = sTh[ka]; thb = sTh[kb];
#else
            const int ca = ea - ra * BW, cb = eb - rb * BW;
            const int ka = (ca >= b1) + (ca >= b2) + (ca >= b3), kb = (cb >= b1) + (cb >= b2) + (cb >= b3);
            pa = ea + (SP - BW) * ra + ka;
            pb = eb + (SP - BW) * rb + kb;
            tha = sTh[ka]; thb = sTh[kb];                                 // a cell's own threshold decides (they differ in the region-adapted variant)
#endif
            if (sa > tha) sorg[pa] = (uint8_t)(sa - 1);
            if (sb > thb) sorg[pb] = (uint8_t)(sb - 1);
        }
        __syncthreads();
        const int na = sa > tha, nb = sb > thb;
#if FV_CORNER_COOP
        // ballot-aggregated compaction: one shared-memory atomic per warp and round
        const uint32_t ma = __ballot_sync(0xffffffffu, na), mb = __ballot_sync(0xffffffffu, nb);
        if (ma | mb) {
            int base = 0;
            if (lane == 0) base = atomicAdd(sCorner, __popc(ma) + __popc(mb));
            base = __shfl_sync(0xffffffffu, base, 0);
            const uint32_t lt = (1u << lane) - 1u;
            if (na) list[base + __popc(ma & lt)] = (uint16_t)pa;
            if (nb) list[base + __popc(ma) + __popc(mb & lt)] = (uint16_t)pb;
        }
#else
        if (na + nb) {
            int pos = atomicAdd(sCorner, na + nb);
            if (na) list[pos++] = (uint16_t)pa;
            if (nb) list[pos] = (uint16_t)pb;
        }
#endif
    }
    __syncthreads();
    (void)lane;
}

// ---- 3. NMS over the corner list: strict '>' against the 8 neighbours of the score plane (zero between cells and around the strip)
__device__ __forceinline__ bool is_local_max(const uint8_t* s, int v)
{
    return (v > s[-1]) & (v > s[1]) & (v > s[-SP - 1]) & (v > s[-SP]) & (v > s[-SP + 1]) & (v > s[SP - 1]) & (v > s[SP]) & (v > s[SP + 1]);
}

__device__ __forceinline__ void nms(const uint8_t* sorg, const uint16_t* list, int nc, uint32_t* bits, int* sHas, const uint8_t* sCellOfG, const int* sC0, int tid)
{
    for (int i = tid; i < nc; i += FS_THREADS) {
        const int p = list[i];
        if (is_local_max(sorg + p, sorg[p])) {
            const int row = p / SP, col = p - row * SP;                   // col = tile column + cell index
            atomicOr(&bits[row * BITW + (col >> 5)], 1u << (col & 31));
#if FV_CELLTAB
            sHas[sCellOfG[col]] = 1;
#else
            sHas[(col > sC0[1]) + (col > sC0[2] + 1) + (col > sC0[3] + 2)] = 1;
#endif
        }
    }
}

// dense scalar path of one cell (one warp): every pixel scored at threshold th, NMS by scanning the cell.  Used when the
// survivor list of the strip overflows (noise-like content); slow, exact, rare.  k = index of the cell in its strip, c0 = its
// first tile column.
__device__ void dense_cell(const uint8_t* org, uint8_t* sorg, uint32_t* bits, int* sHas, int k, int c0, int cw, int h, int th, int lane)
{
    const int npix = cw * h;
    const uint32_t rcpW = ((1u << 20) + cw - 1) / cw;
    for (int p = lane; p < npix; p += 32) {
        const int y = (int)(((uint32_t)p * rcpW) >> 20), x = p - y * cw;
        const uint8_t* c = org + y * BW + c0 + x;
        const int v = c[0], hiT = v + th, loT = v - th;
        const int p0 = c[3 * BW], p8 = c[-3 * BW], p4 = c[3], p12 = c[-3];
        const bool pass = (((p0 > hiT) | (p8 > hiT)) & ((p4 > hiT) | (p12 > hiT))) | (((p0 < loT) | (p8 < loT)) & ((p4 < loT) | (p12 < loT)));
        if (pass) {
            const int s = ring_strength(c, BW);
            if (s > th) sorg[y * SP + c0 + k + x] = (uint8_t)(s - 1);
        }
    }
    __syncwarp();
    for (int p = lane; p < npix; p += 32) {
        const int y = (int)(((uint32_t)p * rcpW) >> 20), x = p - y * cw;
        const int col = c0 + k + x;
        const uint8_t* s = sorg + y * SP + col;
        const int v = s[0];
        if (v != 0 && is_local_max(s, v)) {
            atomicOr(&bits[y * BITW + (col >> 5)], 1u << (col & 31));
            sHas[k] = 1;
        }
    }
    __syncwarp();
}

#if FV_LB
__global__ void __launch_bounds__(FS_THREADS, FV_LB) fast_strip_kernel
#else
__global__ void __launch_bounds__(FS_THREADS) fast_strip_kernel
#endif
    (const __grid_constant__ FastParams P)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ int sCorner, sDense;
    __shared__ int sTot[FS_WARPS];                                          // per-warp survivor counts of the ordered compaction
    __shared__ int sHas[FS_WARPS];                                          // cell has a kept corner (cells the strip does not have: 1)
    __shared__ int sTh[FS_WARPS];                                           // threshold of each cell in the current pass
    __shared__ int sC0[FS_WARPS + 1];                                       // first tile column of each cell, then the strip's end
    __shared__ __align__(8) uint32_t sMask[BW / 4 + 2];                     // per tile word: 0x80 in the bytes inside the scored interior
    __shared__ uint8_t sCellOf[BW], sCellOfG[SP];                           // cell index of a tile column / of a score-plane column
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const StripDesc sd = P.strips[blockIdx.x];
    const int slot = P.slot0 + blockIdx.y;
    const int level = sd.level, BH = P.BH[level];
    Strip S;
    S.W = sd.w; S.h = sd.h;
    const int xs = (sd.x0 - 3) & ~15;                                       // TMA boxes of bytes start on 16-byte boundaries
    S.ax = sd.x0 - xs;
    uint8_t* tile = smem;                                                   // BH x BW level pixels, interior (0,0) at [3][ax]
    uint8_t* score = smem + P.scoreOff;                                     // (h + 2) x SP responses, interior (0,0) of cell k at [1][ax + k]
    uint32_t* bits = reinterpret_cast<uint32_t*>(smem + P.bitsOff);         // [h][BITW] kept corners, score-plane columns
    uint16_t* list = reinterpret_cast<uint16_t*>(smem + P.listOff);         // pretest survivors (tile offsets), then corners (score-plane positions)
    uint8_t* sbm = smem + P.sbmOff;                                         // survivor bitmap of the pretest: bit index = tile offset (h x BW bits)

    if (tid == 0) { mbar_init(&bar, 1); sCorner = 0; sDense = 0; }
    if (tid < FS_WARPS) {
        sHas[tid] = tid < sd.nCells ? 0 : 1;
        int t = P.iniTh;
        if (P.regionTh && tid < sd.nCells) t = P.regionTh[blockIdx.y * P.regionThStride + P.cellRegion[sd.firstCell + tid]];
        sTh[tid] = tid < sd.nCells ? t : 255;
    }
    if (tid <= FS_WARPS) sC0[tid] = tid < sd.nCells ? S.ax + P.cells[sd.firstCell + tid].x0 - sd.x0 : S.ax + S.W + (tid > sd.nCells ? 4096 : 0);
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&bar, (uint32_t)(BW * BH));
        tma_load_3d(tile, &P.maps[level], xs, sd.y0 - 3, level == 0 ? slot - P.z0 : slot, &bar);
    }
    {   // score plane and bitmap (adjacent regions, multiples of 16 bytes) start at zero
        uint4* z = reinterpret_cast<uint4*>(score);
        const int n16 = (P.listOff - P.scoreOff) >> 4;
        for (int i = tid; i < n16; i += FS_THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    {
        const int b1 = sC0[1], b2 = sC0[2], b3 = sC0[3], lo = S.ax, hi = S.ax + S.W;
        if (tid < BW / 4 + 2) {
            uint32_t m = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) { const int col = 4 * tid + j; if (col >= lo && col < hi) m |= 0x80u << (8 * j); }
            sMask[tid] = m;
        }
#if FV_CELLTAB
        for (int col = tid; col < SP; col += FS_THREADS) {
            if (col < BW) sCellOf[col] = (uint8_t)((col >= b1) + (col >= b2) + (col >= b3));
            sCellOfG[col] = (uint8_t)((col > b1) + (col > b2 + 1) + (col > b3 + 2));
        }
#else
        (void)b1; (void)b2; (void)b3;
#endif
    }
    if (warp == 0) mbar_wait(&bar, 0);                                      // one warp polls the mbarrier, the others sleep in the barrier
    __syncthreads();

    const uint8_t* org = tile + 3 * BW;
    uint8_t* sorg = score + SP;
    int th = min(min(sTh[0], sTh[1]), min(sTh[2], sTh[3]));                 // pretest at the lowest cell threshold of the strip
    bool second = false;
    while (true) {
        int n = LIST_CAP + 1;                                               // the byte-parallel pretest is written for th < 128
        if (th < 128) {
            pretest(tile, sbm, S, th, second, sMask, sCellOf, sC0, sHas, tid);
            __syncthreads();
            n = compact_survivors(reinterpret_cast<const uint32_t*>(sbm), list, min(P.sbmWords, (S.h * BW + 31) >> 5), sTot, tid);
        }
        __syncthreads();
        if (n > LIST_CAP) {                // dense path below; nothing was scored in this pass
            if (tid == 0) sDense = 1;
            break;
        }
        strengths(org, sorg, list, n, &sCorner, sTh, sCellOf, sC0, tid);
        nms(sorg, list, sCorner, bits, sHas, sCellOfG, sC0, tid);
        __syncthreads();
        if (second || (sHas[0] & sHas[1] & sHas[2] & sHas[3])) break;
        {   // cells left empty at a threshold above minTh get the second pass
            bool any = false;
            for (int k = 0; k < FS_WARPS; ++k) any |= !sHas[k] && sTh[k] > P.minTh;
            if (!any) break;
        }
        __syncthreads();
        if (tid == 0) sCorner = 0;
        if (tid < FS_WARPS) { if (sTh[tid] <= P.minTh) sHas[tid] |= 2; sTh[tid] = P.minTh; }     // bit 1: already ran at (or below) minTh, nothing to redo
        th = P.minTh; second = true;       // cells without a kept corner at iniTh: the same pass at minTh, restricted to them
        __syncthreads();
    }
    __syncthreads();
    // ---- 4. one warp per cell: dense passes if flagged, then the ordered output -----------------------------------------------
    if (warp < sd.nCells) {
        const int cellIdx = sd.firstCell + warp;
        const CellDesc cd = P.cells[cellIdx];
        const int c0 = sC0[warp], cw = cd.w, h = S.h;
        if (sDense) {
            // overflow in the first pass: nothing is scored yet, every cell goes through iniTh and, if still empty, minTh; overflow in
            // the second pass: only the cells the first pass left empty are redone, at minTh.  Responses do not depend on the
            // threshold, so a dense pass over a cell that already holds some only adds to them.
            int cellTh0 = P.iniTh;
            if (P.regionTh) cellTh0 = P.regionTh[blockIdx.y * P.regionThStride + P.cellRegion[cellIdx]];
            if (!second) dense_cell(org, sorg, bits, sHas, warp, c0, cw, h, cellTh0, lane);
            if (!sHas[warp] && P.minTh < cellTh0) dense_cell(org, sorg, bits, sHas, warp, c0, cw, h, P.minTh, lane);
        }
        uint32_t* out = P.cellCand + (long long)slot * P.cellSlotTotal + cd.slotOff;
        const int outX0 = cd.x0 + cd.relx, outY0 = cd.y0 + cd.rely;
        const int f0 = c0 + warp;                                            // first score-plane column of the cell
        int total = 0;
        for (int r0 = 0; r0 < h; r0 += 32) {
            const int row = r0 + lane;
            unsigned long long m = 0;
            if (row < h) {
                const uint32_t* b = bits + row * BITW + (f0 >> 5);
                const int sh = f0 & 31;
                const uint32_t w0 = b[0], w1 = b[1], w2 = b[2];              // words past the row's end belong to columns past the cell
                m = ((unsigned long long)__funnelshift_r(w1, w2, sh) << 32) | __funnelshift_r(w0, w1, sh);
                if (cw < 64) m &= (1ull << cw) - 1ull;
            }
            const int cnt = __popcll(m);
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
            int pos = total + incl - cnt;
            const uint32_t rowKey = (uint32_t)outX0 | ((uint32_t)(outY0 + row) << 11);
            const uint8_t* srow = sorg + row * SP + f0;
            if (cw <= 32) {                     // the usual geometry: 32-bit bit loop
                uint32_t m32 = (uint32_t)m;
                while (m32) {
                    const int x = __ffs((int)m32) - 1;
                    m32 &= m32 - 1;
                    out[pos++] = (rowKey + (uint32_t)x) | ((uint32_t)srow[x] << 22);
                }
            } else {
                while (m) {
                    const int x = __ffsll((long long)m) - 1;
                    m &= m - 1;
                    out[pos++] = (rowKey + (uint32_t)x) | ((uint32_t)srow[x] << 22);
                }
            }
            total += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) P.cellCount[(long long)slot * P.nCellsTotal + cellIdx] = total;
    }
}

}  // namespace

int orbf_launch_fast(orbf_context* c, int slot0, int n, bool adapted, bool perVideo)
{
    {
        const int r = orbf_refresh_maps(c);
        if (r != ORBF_OK) return r;
    }
    FastParams P;
    int maxBH = 0;
    for (int l = 0; l < c->L; ++l) {
        P.maps[l] = c->tmFast[l];
        P.BH[l] = (short)c->fastBH[l];
        maxBH = std::max(maxBH, c->fastBH[l]);
    }
    if (c->maxCellW > 64) return ORBF_ERR_GEOMETRY;                 // a cell's row of the kept-corner bitmap is read as one 64-bit field
    // regions: [tile BH x BW] [score (h + 2) x SP | bitmap h x BITW words (+ 2 words the output pass may read past the last row) | survivor
    // bitmap h x BW bits] [list]; 18.2 KB at 640x480
    const int maxH = maxBH - 6;
    P.scoreOff = align_up(BW * maxBH, 128);
    P.bitsOff = P.scoreOff + align_up(SP * (maxH + 2), 16);
    P.sbmOff = P.bitsOff + align_up(4 * (BITW * maxH + 2), 16);
    P.sbmWords = align_up((maxH * BW + 31) / 32, 2);
    P.listOff = P.sbmOff + align_up(4 * P.sbmWords, 16);
    const size_t smem = (size_t)P.listOff + LIST_CAP * sizeof(uint16_t);
    P.strips = c->d_strips; P.cells = c->d_cells; P.cellCand = c->d_cellCand; P.cellCount = c->d_cellCount;
    P.cellSlotTotal = c->cellSlotTotal; P.nCellsTotal = c->nCellsTotal; P.iniTh = c->cfg.ini_th_fast; P.minTh = c->cfg.min_th_fast;
    P.slot0 = slot0; P.z0 = c->cur_slot0;
    P.cellRegion = adapted ? c->d_cellRegion : nullptr; P.regionTh = adapted ? c->d_regionTh : nullptr; P.regionThStride = perVideo ? 25 : 0;
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
