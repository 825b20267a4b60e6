// csrc/pyramid.cu — image pyramid (reference ComputePyramid, Features/orbextractor.cpp:833-857) and the
// 7x7 sigma-2 Gaussian blur (orbextractor.cpp:795-796) as register-blocked stencils (coalesced 32-bit row reads).
//
// Level l is cv::resize(level l-1, INTER_LINEAR) — OpenCV's fixed-point bilinear: horizontal
// S[sx]*a0 + S[sx+1]*a1 with 11-bit coefficients, vertical (((b0*(h0>>4))>>16) + ((b1*(h1>>4))>>16) + 2) >> 2.
// The 19-px copyMakeBorder frame of the reference is never read downstream and is not materialised.
// The blur is OpenCV's fixed-point Gaussian: taps [18,34,48,56,48,34,18]/256, horizontal Q8.8 (u16),
// vertical Q16.16, (v + 32768) >> 16, BORDER_REFLECT_101 at the true level edge.
// HBM-bound: algorithmic bytes per 640x480 frame = 926,546 rd + 643,332 wr (resize), 950,532 rd + wr (blur).
#include "orbf_internal.h"

namespace {

// ---- bilinear resize, register-blocked: a lane owns 4 adjacent dst columns x RS_ROWS dst rows ------------------------
// The x coefficients (source offset, a0, a1) of the 4 columns are loaded once and reused down the rows; source pixels
// are read straight through L1 (byte loads: neighbouring lanes hit the same 128-byte lines), the 4 results of a row
// leave as one aligned 32-bit store.  No shared memory, no barriers.
constexpr int RS_WARPS = 4, RS_THREADS = RS_WARPS * 32, RS_ROWS = 4, RS_COLS = 128;

__global__ void __launch_bounds__(RS_THREADS) resize_kernel(LevelView src, uint8_t* __restrict__ dstBase,
    long long dstFrameStride, int dstPitch, int dw, int dh, const ResizeCoef* __restrict__ tx,
    const ResizeCoef* __restrict__ ty, int slot0)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int slot = slot0 + blockIdx.z;
    const uint8_t* sImg = src.base + (long long)slot * src.frameStride;
    uint8_t* dImg = dstBase + (long long)slot * dstFrameStride;
    const int xb = blockIdx.x * RS_COLS + lane * 4;
    const int y0 = (blockIdx.y * RS_WARPS + warp) * RS_ROWS;
    if (xb >= dw || y0 >= dh) return;
    int sx0[4], sx1[4], a0[4], a1[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int x = min(xb + k, dw - 1);                      // columns past the edge recompute the last one (never stored as valid)
        const ResizeCoef cx = tx[x];
        sx0[k] = cx.ofs; sx1[k] = min(cx.ofs + 1, src.w - 1); a0[k] = cx.a0; a1[k] = cx.a1;
    }
#pragma unroll
    for (int r = 0; r < RS_ROWS; ++r) {
        const int y = y0 + r;
        if (y >= dh) break;
        const ResizeCoef cy = ty[y];
        const uint8_t* r0 = sImg + (long long)cy.ofs * src.pitch;
        const uint8_t* r1 = sImg + (long long)min(cy.ofs + 1, src.h - 1) * src.pitch;
        const int b0 = cy.a0, b1 = cy.a1;
        uint32_t packed = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int h0 = __ldg(r0 + sx0[k]) * a0[k] + __ldg(r0 + sx1[k]) * a1[k];
            const int h1 = __ldg(r1 + sx0[k]) * a0[k] + __ldg(r1 + sx1[k]) * a1[k];
            const int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
            packed |= (uint32_t)(v & 255) << (8 * k);
        }
        *reinterpret_cast<uint32_t*>(dImg + (long long)y * dstPitch + xb) = packed;   // pitch % 128 == 0: aligned, in-plane
    }
}

// ---- 7x7 Gaussian, register-blocked, no shared memory ---------------------------------------------------------
// The fixed-point result (sum_ij k_i k_j p_ij + 32768) >> 16 is exact integer arithmetic, so pass order is free.
// A lane owns 4 adjacent columns (one aligned 32-bit word per row; the neighbouring words come from __shfl, so a
// warp reads each row as one coalesced 128-byte segment).  Horizontal pass: two IDP4A per pixel on byte windows
// cut out with funnel shifts (taps 18,34,48,56 | 48,34,18,0).  Vertical pass: 7-row ring of the 32-bit row sums in
// registers, walked down a strip of BL_ROWS rows.  BORDER_REFLECT_101 at the true level edge (orbextractor.cpp:796).
constexpr int BL_WARPS = 4, BL_THREADS = BL_WARPS * 32, BL_ROWS = 16, BL_COLS = 128;

__device__ __forceinline__ int reflect101(int p, int n)
{
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = (p < 0) ? -p : 2 * (n - 1) - p;
    return p;
}

__device__ __forceinline__ uint32_t load_word_reflect(const uint8_t* __restrict__ row, int xw, int w)
{
    if (xw >= 0 && xw + 3 < w) return __ldg(reinterpret_cast<const uint32_t*>(row + xw));
    uint32_t v = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) v |= (uint32_t)__ldg(row + reflect101(xw + k, w)) << (8 * k);
    return v;
}

__device__ __forceinline__ void hrow7(uint32_t W1, uint32_t edge, int lane, uint32_t out[4])
{
    uint32_t W0 = __shfl_up_sync(0xffffffffu, W1, 1), W2 = __shfl_down_sync(0xffffffffu, W1, 1);
    if (lane == 0) W0 = edge;
    if (lane == 31) W2 = edge;
    const uint32_t kA = 18u | (34u << 8) | (48u << 16) | (56u << 24), kB = 48u | (34u << 8) | (18u << 16);
    // output i uses bytes [i-3, i+3]: window A = bytes i-3..i, window B = bytes i+1..i+4 (last tap weight 0)
    out[0] = __dp4a(__funnelshift_r(W0, W1, 8), kA, __dp4a(__funnelshift_r(W1, W2, 8), kB, 0u));
    out[1] = __dp4a(__funnelshift_r(W0, W1, 16), kA, __dp4a(__funnelshift_r(W1, W2, 16), kB, 0u));
    out[2] = __dp4a(__funnelshift_r(W0, W1, 24), kA, __dp4a(__funnelshift_r(W1, W2, 24), kB, 0u));
    out[3] = __dp4a(W1, kA, __dp4a(W2, kB, 0u));
}

template <bool INTERIOR>
__device__ __forceinline__ void blur_strip(const uint8_t* __restrict__ sImg, uint8_t* __restrict__ dImg, int pitch, int dstPitch,
    int w, int h, int x, int y0, int lane)
{
    // stage all BL_ROWS + 6 source rows of the strip first: every load of the strip is in flight at once
    // (lane 0 / lane 31 also fetch the word left / right of the warp's 128-byte segment)
    uint32_t w1[BL_ROWS + 6], ed[BL_ROWS + 6];
    const int xe = (lane == 0) ? x - 4 : x + 4;
    const bool edgeLane = lane == 0 || lane == 31;
    if (INTERIOR) {     // warp-uniform: all 32 words of every row of the strip lie inside the level, no row reflection
        const uint8_t* p = sImg + (long long)(y0 - 3) * pitch + x;
        // edge word: loaded when it is inside the level; at the level's left / right edge (w % 4 == 0 there) it is the
        // REFLECT_101 image of the lane's own word: b[-1..-3] = b[1..3] and b[w..w+2] = b[w-2..w-4]
        const bool loadEdge = edgeLane && xe >= 0 && xe + 3 < w;
        const uint32_t perm = (lane == 0) ? 0x1230u : 0x0012u;
#pragma unroll
        for (int r = 0; r < BL_ROWS + 6; ++r) {
            w1[r] = __ldg(reinterpret_cast<const uint32_t*>(p));
            ed[r] = loadEdge ? __ldg(reinterpret_cast<const uint32_t*>(p + (xe - x))) : __byte_perm(w1[r], 0u, perm);
            p += pitch;
        }
    } else {
#pragma unroll 1
        for (int r = 0; r < BL_ROWS + 6; ++r) {
            const uint8_t* row = sImg + (long long)reflect101(y0 - 3 + r, h) * pitch;
            w1[r] = load_word_reflect(row, x, w);
            ed[r] = edgeLane ? load_word_reflect(row, xe, w) : 0u;
        }
    }
    uint32_t ring[7][4];
#pragma unroll
    for (int r = 0; r < 6; ++r) hrow7(w1[r], ed[r], lane, ring[r]);
    uint8_t* q = dImg + (long long)y0 * dstPitch + x;
#pragma unroll
    for (int i = 0; i < BL_ROWS; ++i) {
        hrow7(w1[i + 6], ed[i + 6], lane, ring[(i + 6) % 7]);
        uint32_t packed = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t acc = 18u * (ring[i % 7][k] + ring[(i + 6) % 7][k]) + 34u * (ring[(i + 1) % 7][k] + ring[(i + 5) % 7][k])
                + 48u * (ring[(i + 2) % 7][k] + ring[(i + 4) % 7][k]) + 56u * ring[(i + 3) % 7][k];
            packed |= ((acc + 32768u) >> 16) << (8 * k);
        }
        if (INTERIOR || (x < w && y0 + i < h)) *reinterpret_cast<uint32_t*>(q) = packed;
        q += dstPitch;
    }
}

__global__ void __launch_bounds__(BL_THREADS) blur7_kernel(LevelView src, uint8_t* __restrict__ dstBase,
    long long dstFrameStride, int dstPitch, int slot0)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int slot = slot0 + blockIdx.z;
    const uint8_t* sImg = src.base + (long long)slot * src.frameStride;
    uint8_t* dImg = dstBase + (long long)slot * dstFrameStride;
    const int w = src.w, h = src.h;
    const int xb = blockIdx.x * BL_COLS, x = xb + lane * 4;
    const int y0 = (blockIdx.y * BL_WARPS + warp) * BL_ROWS;
    if (y0 >= h) return;
    const bool interior = (xb + BL_COLS + 4 <= w || xb + BL_COLS == w) && y0 >= 3 && y0 + BL_ROWS + 3 <= h;
    if (interior) blur_strip<true>(sImg, dImg, src.pitch, dstPitch, w, h, x, y0, lane);
    else blur_strip<false>(sImg, dImg, src.pitch, dstPitch, w, h, x, y0, lane);
}

}  // namespace

int orbf_launch_pyramid(orbf_context* c, int slot0, int n)
{
    PyrView pv = orbf_pyr_view(c, false);
    for (int l = 1; l < c->L; ++l) {
        const LevelGeom& g = c->lg[l];
        dim3 grid((g.w + RS_COLS - 1) / RS_COLS, (g.h + RS_WARPS * RS_ROWS - 1) / (RS_WARPS * RS_ROWS), n);
        resize_kernel<<<grid, RS_THREADS, 0, c->stream>>>(pv.lv[l - 1], c->d_pyr[l], (long long)g.plane, g.pitch, g.w, g.h,
            c->d_resizeTab + g.tabX, c->d_resizeTab + g.tabY, slot0);
        ORBF_LAUNCH_CHECK(c);
    }
    return ORBF_OK;
}

int orbf_launch_blur(orbf_context* c, int slot0, int n)
{
    PyrView pv = orbf_pyr_view(c, false);
    for (int l = 0; l < c->L; ++l) {
        const LevelGeom& g = c->lg[l];
        dim3 grid((g.w + BL_COLS - 1) / BL_COLS, (g.h + BL_WARPS * BL_ROWS - 1) / (BL_WARPS * BL_ROWS), n);
        blur7_kernel<<<grid, BL_THREADS, 0, c->stream>>>(pv.lv[l], c->d_blur[l], (long long)g.plane, g.pitch, slot0);
        ORBF_LAUNCH_CHECK(c);
    }
    return ORBF_OK;
}
