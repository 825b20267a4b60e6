// csrc/fast.cu — per-cell FAST-9/16 with non-max suppression and the iniTh -> minTh fallback
// (reference ComputeKeyPointsOctTree, Features/orbextractor.cpp:665-723, calling cv::FAST(roi, th, true)).
//
// One warp per (cell, frame).  The cell's scored interior plus a 3-px ring halo is staged in shared memory;
// corner strength S(p) = max over the 16 arcs of 9 contiguous ring pixels of min(centre - ring) for both
// polarities (cv::FAST response = S - 1, corner iff S > th); NMS is strict '>' against the 8 neighbours
// *inside the same cell interior* (quirk Q1: each cell is its own cv::FAST call, so neighbours in an
// adjacent cell count as 0); a cell with no survivor at iniTh is redone at minTh.  Survivors are written
// row-major (warp-ballot compaction) to the cell's private slot range, so the candidate list a later stage
// gathers cell by cell is in exactly the reference's push_back order.
#include "orbf_internal.h"

namespace {

constexpr int FC_WARPS = 4, FC_THREADS = FC_WARPS * 32;

// Corner strength without ever negating a min/max result:
//   bright arcs: min over 9 contiguous (v - ring) = v - max_arc(ring);  dark arcs: min (ring - v) = min_arc(ring) - v
//   S = max(v - A, B - v),  A = min over the 16 arcs of the arc maximum,  B = max over the 16 arcs of the arc minimum.
// (nvcc 12.9 for sm_100a miscompiles max(x, -max(...)) chains — the negation is lost when ptxas fuses them into
//  VIMNMX3; tools/nvcc_minmax_bug.cu reproduces it.  The raw-pixel formulation is also cheaper: no 16 subtractions.)
__device__ __forceinline__ int ring_strength(const uint8_t* p, int rp)
{
    // ring offsets in cv::FAST order: (0,3)(1,3)(2,2)(3,1)(3,0)(3,-1)(2,-2)(1,-3)(0,-3)(-1,-3)(-2,-2)(-3,-1)(-3,0)(-3,1)(-2,2)(-1,3)
    int r[25];
    r[0] = p[3 * rp];      r[1] = p[3 * rp + 1];  r[2] = p[2 * rp + 2];   r[3] = p[rp + 3];
    r[4] = p[3];           r[5] = p[-rp + 3];     r[6] = p[-2 * rp + 2];  r[7] = p[-3 * rp + 1];
    r[8] = p[-3 * rp];     r[9] = p[-3 * rp - 1]; r[10] = p[-2 * rp - 2]; r[11] = p[-rp - 3];
    r[12] = p[-3];         r[13] = p[rp - 3];     r[14] = p[2 * rp - 2];  r[15] = p[3 * rp - 1];
#pragma unroll
    for (int k = 16; k < 25; ++k) r[k] = r[k - 16];
    int A = 255, B = 0;
#pragma unroll
    for (int k = 0; k < 16; k += 2) {
        int lo = min(r[k + 1], r[k + 2]), hi = max(r[k + 1], r[k + 2]);
#pragma unroll
        for (int j = 3; j <= 8; ++j) { lo = min(lo, r[k + j]); hi = max(hi, r[k + j]); }
        A = min(A, min(max(hi, r[k]), max(hi, r[k + 9])));
        B = max(B, max(min(lo, r[k]), min(lo, r[k + 9])));
    }
    const int v = p[0];
    return max(v - A, B - v);
}

// One WARP per (cell, frame); FC_WARPS cells per CTA.  No block barriers: the three phases (quick test ->
// dense strength -> NMS + ordered write) only need __syncwarp.  Because a single warp walks the cell in
// row-major chunks of 32 pixels, one ballot per chunk yields the reference's output order directly.
__global__ void __launch_bounds__(FC_THREADS) fast_cell_kernel(PyrView pv, const CellDesc* __restrict__ cells,
    int nCellsTotal, uint32_t* __restrict__ cellCand, int* __restrict__ cellCount, int cellSlotTotal, int iniTh, int minTh,
    int slot0, int regPitch, int regBytes, int scoreBytes, int warpBytes)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cellIdx = blockIdx.x * FC_WARPS + warp;
    if (cellIdx >= nCellsTotal) return;
    uint8_t* reg = smem + warp * warpBytes;                       // (h+6) x regPitch : level pixels, word-aligned columns
    uint8_t* score = reg + regBytes;                              // (h+2) x (w+2)    : response, zero border
    uint16_t* list = reinterpret_cast<uint16_t*>(score + scoreBytes);   // quick-test survivors

    const CellDesc cd = cells[cellIdx];
    const int slot = slot0 + blockIdx.y;
    const LevelView lv = pv.lv[cd.level];
    const int w = cd.w, h = cd.h, sp = w + 2, npix = w * h;
    const uint32_t rcpW = ((1u << 20) + w - 1) / w;               // y = (p * rcpW) >> 20 is exact for p < 2^20 / w
    // stage (h+6) rows of (w+6) pixels with aligned 32-bit loads; ax = misalignment of the first column
    const int gx0 = cd.x0 - 3, ax = gx0 & 3, wordsPerRow = (ax + w + 6 + 3) >> 2;
    const uint8_t* img = lv.base + (long long)slot * lv.frameStride + (long long)(cd.y0 - 3) * lv.pitch + (gx0 - ax);
    {
        const uint32_t rcpR = ((1u << 20) + wordsPerRow - 1) / wordsPerRow;
        const int nWords = (h + 6) * wordsPerRow;
        for (int i = lane; i < nWords; i += 32) {
            const int r = (int)(((uint32_t)i * rcpR) >> 20), c = i - r * wordsPerRow;
            *reinterpret_cast<uint32_t*>(reg + r * regPitch + 4 * c) = __ldg(reinterpret_cast<const uint32_t*>(img + (long long)r * lv.pitch) + c);
        }
    }
    const uint8_t* org = reg + 3 * regPitch + 3 + ax;             // interior pixel (0,0)
    uint32_t* out = cellCand + (long long)slot * cellSlotTotal + cd.slotOff;
    int total = 0;
    int th = iniTh;
    for (int attempt = 0; attempt < 2; ++attempt) {
        for (int i = lane; i < (h + 2) * sp; i += 32) score[i] = 0;
        __syncwarp();
        // 1. cheap necessary condition (an arc of 9 contains one pixel of every opposite pair, all of one polarity), compacted
        int nList = 0;
        for (int base = 0; base < npix; base += 32) {
            const int p = base + lane;
            bool pass = false;
            if (p < npix) {
                const int y = (int)(((uint32_t)p * rcpW) >> 20), x = p - y * w;
                const uint8_t* c = org + y * regPitch + x;
                const int v = c[0], hiT = v + th, loT = v - th;
                const int p0 = c[3 * regPitch], p8 = c[-3 * regPitch], p4 = c[3], p12 = c[-3];
                pass = (((p0 > hiT) | (p8 > hiT)) & ((p4 > hiT) | (p12 > hiT))) | (((p0 < loT) | (p8 < loT)) & ((p4 < loT) | (p12 < loT)));
            }
            const unsigned m = __ballot_sync(0xffffffffu, pass);
            if (pass) list[nList + __popc(m & ((1u << lane) - 1))] = (uint16_t)p;
            nList += __popc(m);
        }
        __syncwarp();
        // 2. full strength on the dense survivor list
        for (int i = lane; i < nList; i += 32) {
            const int p = list[i];
            const int y = (int)(((uint32_t)p * rcpW) >> 20), x = p - y * w;
            const int s = ring_strength(org + y * regPitch + x, regPitch);
            if (s > th) score[(y + 1) * sp + x + 1] = (uint8_t)(s - 1);
        }
        __syncwarp();
        // 3. NMS + row-major ordered compaction
        total = 0;
        for (int base = 0; base < npix; base += 32) {
            const int p = base + lane;
            bool keep = false;
            int v = 0, x = 0, y = 0;
            if (p < npix) {
                y = (int)(((uint32_t)p * rcpW) >> 20); x = p - y * w;
                const uint8_t* s = score + (y + 1) * sp + x + 1;
                v = s[0];
                keep = v > 0 && v > s[-1] && v > s[1] && v > s[-sp - 1] && v > s[-sp] && v > s[-sp + 1] && v > s[sp - 1]
                    && v > s[sp] && v > s[sp + 1];
            }
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            if (keep) {
                const int xr = cd.x0 + x + cd.relx, yr = cd.y0 + y + cd.rely;
                out[total + __popc(m & ((1u << lane) - 1))] = (uint32_t)xr | ((uint32_t)yr << 11) | ((uint32_t)v << 22);
            }
            total += __popc(m);
        }
        if (total > 0 || minTh >= th) break;
        th = minTh;   // empty cell at iniTh: rerun at minTh (orbextractor.cpp:709-712)
        __syncwarp();
    }
    if (lane == 0) cellCount[(long long)slot * nCellsTotal + cellIdx] = total;
}

}  // namespace

int orbf_launch_fast(orbf_context* c, int slot0, int n)
{
    PyrView pv = orbf_pyr_view(c, false);
    const int regPitch = align_up(c->maxCellW + 6 + 3, 4);
    const int regBytes = (c->maxCellH + 6) * regPitch;
    const int scoreBytes = align_up((c->maxCellH + 2) * (c->maxCellW + 2), 4);
    const int warpBytes = align_up(regBytes + scoreBytes + c->maxCellW * c->maxCellH * (int)sizeof(uint16_t), 16);
    const size_t smem = (size_t)warpBytes * FC_WARPS;
    if (smem > 200 * 1024) return ORBF_ERR_GEOMETRY;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(fast_cell_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return orbf_cuda_fail(c, e, "fast smem attr", __FILE__, __LINE__);
    }
    dim3 grid((c->nCellsTotal + FC_WARPS - 1) / FC_WARPS, n);
    fast_cell_kernel<<<grid, FC_THREADS, smem, c->stream>>>(pv, c->d_cells, c->nCellsTotal, c->d_cellCand, c->d_cellCount,
        c->cellSlotTotal, c->cfg.ini_th_fast, c->cfg.min_th_fast, slot0, regPitch, regBytes, scoreBytes, warpBytes);
    ORBF_LAUNCH_CHECK(c);
    return ORBF_OK;
}
