// csrc/fast.cu — per-cell FAST-9/16 with non-max suppression and the iniTh -> minTh fallback
// (reference ComputeKeyPointsOctTree, Features/orbextractor.cpp:665-723, calling cv::FAST(roi, th, true)).
//
// One CTA per (cell, frame).  The cell's scored interior plus a 3-px ring halo is staged in shared memory;
// corner strength S(p) = max over the 16 arcs of 9 contiguous ring pixels of min(centre - ring) for both
// polarities (cv::FAST response = S - 1, corner iff S > th); NMS is strict '>' against the 8 neighbours
// *inside the same cell interior* (quirk Q1: each cell is its own cv::FAST call, so neighbours in an
// adjacent cell count as 0); a cell with no survivor at iniTh is redone at minTh.  Survivors are written
// row-major (warp-ballot compaction) to the cell's private slot range, so the candidate list a later stage
// gathers cell by cell is in exactly the reference's push_back order.
#include "orbf_internal.h"

namespace {

constexpr int FC_THREADS = 128;

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

__global__ void __launch_bounds__(FC_THREADS) fast_cell_kernel(PyrView pv, const CellDesc* __restrict__ cells,
    int nCellsTotal, uint32_t* __restrict__ cellCand, int* __restrict__ cellCount, int cellSlotTotal, int iniTh, int minTh,
    int slot0, int regPitch, int maxW, int maxH)
{
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t* reg = smem;                                        // (maxH+6) x regPitch : level pixels
    uint8_t* score = reg + (maxH + 6) * regPitch;               // (maxH+2) x (maxW+2)  : response, zero border
    uint16_t* list = reinterpret_cast<uint16_t*>(score + align_up((maxH + 2) * (maxW + 2), 16));   // quick-test survivors
    __shared__ int sListCount;
    __shared__ int sWarpCount[FC_THREADS / 32];

    const CellDesc cd = cells[blockIdx.x];
    const int slot = slot0 + blockIdx.y;
    const LevelView lv = pv.lv[cd.level];
    const uint8_t* img = lv.base + (long long)slot * lv.frameStride;
    const int w = cd.w, h = cd.h, sp = w + 2;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int i = tid; i < (h + 6) * (w + 6); i += FC_THREADS) {
        const int r = i / (w + 6), c = i - r * (w + 6);
        reg[r * regPitch + c] = __ldg(img + (long long)(cd.y0 - 3 + r) * lv.pitch + cd.x0 - 3 + c);
    }
    uint32_t* out = cellCand + (long long)slot * cellSlotTotal + cd.slotOff;
    int total = 0;
    int th = iniTh;
    for (int attempt = 0; attempt < 2; ++attempt) {
        for (int i = tid; i < (h + 2) * sp; i += FC_THREADS) score[i] = 0;
        if (tid == 0) sListCount = 0;
        __syncthreads();
        // 1. cheap necessary condition (an arc of 9 contains one pixel of every opposite pair), compacted
        for (int base = 0; base < w * h; base += FC_THREADS) {
            const int p = base + tid;
            bool pass = false;
            if (p < w * h) {
                const int y = p / w, x = p - y * w;
                const uint8_t* c = reg + (y + 3) * regPitch + x + 3;
                const int v = c[0], hiT = v + th, loT = v - th;
                const int p0 = c[3 * regPitch], p8 = c[-3 * regPitch], p4 = c[3], p12 = c[-3];
                pass = ((p0 > hiT) | (p8 > hiT) | (p0 < loT) | (p8 < loT)) & ((p4 > hiT) | (p12 > hiT) | (p4 < loT) | (p12 < loT));
            }
            const unsigned m = __ballot_sync(0xffffffffu, pass);
            int wbase = 0;
            if (lane == 0 && m) wbase = atomicAdd(&sListCount, __popc(m));
            wbase = __shfl_sync(0xffffffffu, wbase, 0);
            if (pass) list[wbase + __popc(m & ((1u << lane) - 1))] = (uint16_t)p;
        }
        __syncthreads();
        // 2. full strength on the dense survivor list
        const int nList = sListCount;
        for (int i = tid; i < nList; i += FC_THREADS) {
            const int p = list[i];
            const int y = p / w, x = p - y * w;
            const int s = ring_strength(reg + (y + 3) * regPitch + x + 3, regPitch);
            if (s > th) score[(y + 1) * sp + x + 1] = (uint8_t)(s - 1);
        }
        __syncthreads();
        // 3. NMS + row-major ordered compaction: each warp owns a contiguous quarter of the pixel range
        const int per = (w * h + FC_THREADS / 32 - 1) / (FC_THREADS / 32);
        const int pBeg = warp * per, pEnd = min(pBeg + per, w * h);
        int cnt = 0;
        for (int base = pBeg; base < pEnd; base += 32) {
            const int p = base + lane;
            bool keep = false;
            if (p < pEnd) {
                const int y = p / w, x = p - y * w;
                const uint8_t* s = score + (y + 1) * sp + x + 1;
                const int v = s[0];
                keep = v > 0 && v > s[-1] && v > s[1] && v > s[-sp - 1] && v > s[-sp] && v > s[-sp + 1] && v > s[sp - 1]
                    && v > s[sp] && v > s[sp + 1];
            }
            cnt += __popc(__ballot_sync(0xffffffffu, keep));
        }
        if (lane == 0) sWarpCount[warp] = cnt;
        __syncthreads();
        int off = 0;
        total = 0;
        for (int k = 0; k < FC_THREADS / 32; ++k) { if (k < warp) off += sWarpCount[k]; total += sWarpCount[k]; }
        if (total > 0) {
            for (int base = pBeg; base < pEnd; base += 32) {
                const int p = base + lane;
                bool keep = false;
                int v = 0, x = 0, y = 0;
                if (p < pEnd) {
                    y = p / w; x = p - y * w;
                    const uint8_t* s = score + (y + 1) * sp + x + 1;
                    v = s[0];
                    keep = v > 0 && v > s[-1] && v > s[1] && v > s[-sp - 1] && v > s[-sp] && v > s[-sp + 1] && v > s[sp - 1]
                        && v > s[sp] && v > s[sp + 1];
                }
                const unsigned m = __ballot_sync(0xffffffffu, keep);
                if (keep) {
                    const int xr = cd.x0 + x + cd.relx, yr = cd.y0 + y + cd.rely;
                    out[off + __popc(m & ((1u << lane) - 1))] = (uint32_t)xr | ((uint32_t)yr << 11) | ((uint32_t)v << 22);
                }
                off += __popc(m);
            }
            break;
        }
        if (minTh >= th) break;
        th = minTh;   // empty cell at iniTh: rerun at minTh (orbextractor.cpp:709-712)
        __syncthreads();
    }
    if (tid == 0) cellCount[(long long)slot * nCellsTotal + blockIdx.x] = total;
}

}  // namespace

int orbf_launch_fast(orbf_context* c, int slot0, int n)
{
    PyrView pv = orbf_pyr_view(c, false);
    const int regPitch = align_up(c->maxCellW + 6, 4) + 4;
    const size_t smem = (size_t)(c->maxCellH + 6) * regPitch + align_up((c->maxCellH + 2) * (c->maxCellW + 2), 16)
        + (size_t)c->maxCellW * c->maxCellH * sizeof(uint16_t) + 16;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(fast_cell_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return orbf_cuda_fail(c, e, "fast smem attr", __FILE__, __LINE__);
    }
    dim3 grid(c->nCellsTotal, n);
    fast_cell_kernel<<<grid, FC_THREADS, smem, c->stream>>>(pv, c->d_cells, c->nCellsTotal, c->d_cellCand, c->d_cellCount,
        c->cellSlotTotal, c->cfg.ini_th_fast, c->cfg.min_th_fast, slot0, regPitch, c->maxCellW, c->maxCellH);
    ORBF_LAUNCH_CHECK(c);
    return ORBF_OK;
}
