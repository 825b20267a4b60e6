// csrc/pyramid.cu — image pyramid (reference ComputePyramid, Features/orbextractor.cpp:833-857) and the 7x7 sigma-2
// Gaussian blur (orbextractor.cpp:795-796) as TMA-fed tile stencils: ONE WARP PER TILE, four tiles per CTA.
//
// Each warp fetches its source box (tile + halo) with one cp.async.bulk.tensor.3d (UTMALDG) into shared memory, waits
// on its own mbarrier, and walks down the tile with every lane owning 4 adjacent output columns, so each output row
// leaves as one coalesced 128-byte store.  Out-of-image box elements arrive as zeros; BORDER_REFLECT_101 is patched
// into the halo of edge tiles only.  All levels of all frames run in one launch per stage (tile tables).
//
// Level l is cv::resize(level l-1, INTER_LINEAR) — OpenCV's fixed-point bilinear: horizontal S[sx]*a0 + S[sx+1]*a1 with
// 11-bit coefficients (one IDP.2A per pixel and source row), vertical (((b0*(h0>>4))>>16) + ((b1*(h1>>4))>>16) + 2) >> 2;
// horizontally interpolated source rows are reused between consecutive output rows.  The 19-px copyMakeBorder frame of
// the reference is never read downstream and is not materialised.
// The blur is OpenCV's fixed-point Gaussian: taps [18,34,48,56,48,34,18]/256, horizontal Q8.8, vertical Q16.16,
// (v + 32768) >> 16 — exact integer arithmetic, so the pass order is free: horizontal pass = 10 IDP.4A per 4 pixels with
// the taps pre-shifted into byte-coefficient words (no data shifts), vertical pass on a 7-row register ring.
// HBM-bound by design: algorithmic bytes per 640x480 frame = 926,546 rd + 643,332 wr (resize), 950,532 rd + wr (blur).
#include "orbf_internal.h"

namespace {

constexpr int TL_W = 128, TL_WARPS = 4, TL_THREADS = TL_WARPS * 32;
constexpr int BL_H = 36, BL_BW = 160, BL_BH = BL_H + 6, BL_TILE_BYTES = ((BL_BW * BL_BH + 127) / 128) * 128;   // 6 turns of the 6-entry row-pair ring
constexpr int RS_H = 16;

struct StageParams {
    CUtensorMap maps[ORBF_MAX_LEVELS];          // source of level l's stage (blur: level l; resize: level l-1)
    const TileDesc* tiles; int nTiles;
    uint8_t* dst[ORBF_MAX_LEVELS]; long long dstFrameStride[ORBF_MAX_LEVELS];
    int dstPitch[ORBF_MAX_LEVELS];
    short w[ORBF_MAX_LEVELS], h[ORBF_MAX_LEVELS];      // destination size (blur: = source size)
    short BW[ORBF_MAX_LEVELS], BH[ORBF_MAX_LEVELS];    // source box (resize)
    int tabX[ORBF_MAX_LEVELS], tabY[ORBF_MAX_LEVELS];
    const ResizeCoef* tab;
    int slot0, z0, srcLevel0;                    // z of a slot in maps[l]: slot - z0 when the source is the caller's input plane
    int tileStride;                              // bytes of shared memory per warp (resize)
};

// ---- 7x7 Gaussian ------------------------------------------------------------------------------------------------------
// horizontal 7-tap sums of the 4 pixels whose first byte is byte 4 of the 12-byte window (W0, W1, W2)
__device__ __forceinline__ void hrow7(const uint32_t* p, uint32_t out[4])
{
    const uint32_t W0 = p[0], W1 = p[1], W2 = p[2];
    out[0] = __dp4a(W0, 0x30221200u, __dp4a(W1, 0x12223038u, 0u));                              // bytes 1..7
    out[1] = __dp4a(W0, 0x22120000u, __dp4a(W1, 0x22303830u, __dp4a(W2, 0x00000012u, 0u)));      // bytes 2..8
    out[2] = __dp4a(W0, 0x12000000u, __dp4a(W1, 0x30383022u, __dp4a(W2, 0x00001222u, 0u)));      // bytes 3..9
    out[3] = __dp4a(W1, 0x38302212u, __dp4a(W2, 0x00122230u, 0u));                              // bytes 4..10
}

__global__ void __launch_bounds__(TL_THREADS) blur_tile_kernel(const __grid_constant__ StageParams P)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bars[TL_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tileIdx = blockIdx.x * TL_WARPS + warp;
    if (tileIdx >= P.nTiles) return;
    const TileDesc td = P.tiles[tileIdx];
    const int level = td.level, x0 = td.x0, y0 = td.y0;
    const int slot = P.slot0 + blockIdx.y;
    const int w = P.w[level], h = P.h[level];
    uint8_t* tile = smem + warp * BL_TILE_BYTES;                    // BL_BH x BL_BW; image (x0, y0) at tile [3][16]
    if (lane == 0) {
        mbar_init(&bars[warp], 1);
        mbar_expect_tx(&bars[warp], BL_BW * BL_BH);
        tma_load_3d(tile, &P.maps[level], x0 - 16, y0 - 3, level == P.srcLevel0 ? slot - P.z0 : slot, &bars[warp]);
    }
    __syncwarp();
    mbar_wait(&bars[warp], 0);
    // BORDER_REFLECT_101 at the true level edges: rows first, then columns (corners inherit both)
    if (y0 == 0 || y0 + BL_H + 3 > h) {
        uint32_t* t32 = reinterpret_cast<uint32_t*>(tile);
#pragma unroll
        for (int k = 1; k <= 3; ++k) {
            if (y0 == 0)
                for (int i = lane; i < BL_BW / 4; i += 32) t32[(3 - k) * (BL_BW / 4) + i] = t32[(3 + k) * (BL_BW / 4) + i];
            const int tr = h - 1 + k - y0 + 3, sr = h - 1 - k - y0 + 3;     // image rows h-1+k <- h-1-k
            if (tr < BL_BH && sr >= 0)
                for (int i = lane; i < BL_BW / 4; i += 32) t32[tr * (BL_BW / 4) + i] = t32[sr * (BL_BW / 4) + i];
        }
        __syncwarp();
    }
    if (x0 == 0 || x0 + TL_W + 3 > w) {
        const int cR = 16 + (w - 1 - x0);
        for (int r = lane; r < BL_BH; r += 32) {
            uint8_t* t = tile + r * BL_BW;
#pragma unroll
            for (int k = 1; k <= 3; ++k) {
                if (x0 == 0) t[16 - k] = t[16 + k];
                if (cR + k < BL_BW && cR - k >= 0) t[cR + k] = t[cR - k];
            }
        }
        __syncwarp();
    }
    const uint32_t* rowp = reinterpret_cast<const uint32_t*>(tile) + 3 + lane;     // words of tile columns 12 + 4*lane ..
    const int x = x0 + 4 * lane;
    uint8_t* q = P.dst[level] + (long long)slot * P.dstFrameStride[level] + (long long)y0 * P.dstPitch[level] + x;
    const int pitch = P.dstPitch[level];
    const bool colOk = x < w;
    // Vertical pass on ROW PAIRS: ring entry j holds, per pixel, the horizontal sums (< 2^16) of tile rows r and r + 1 packed into one
    // word, so an output row is three IDP.2A (taps 18,34 | 48,56 | 48,34 against the pairs of rows y..y+5) on top of 18 * row y+6 +
    // rounding: 4 multiply-adds + 1 PRMT (the new pair) per pixel instead of 3 adds + 4 multiply-adds.
    uint32_t pr[6][4], prevH[4];
    {
        uint32_t h0[4], h1[4];
        hrow7(rowp, h0);
#pragma unroll
        for (int r = 0; r < 5; ++r) {
            hrow7(rowp + (r + 1) * (BL_BW / 4), h1);
#pragma unroll
            for (int k = 0; k < 4; ++k) { pr[r][k] = __byte_perm(h0[k], h1[k], 0x5410); h0[k] = h1[k]; }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) prevH[k] = h0[k];                 // tile row 5
    }
#pragma unroll 1
    for (int turn = 0; turn < BL_H / 6; ++turn) {
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            const int row = turn * 6 + i;                             // output row: tile rows row .. row + 6; pr[(row + j) % 6] = rows (row + j, row + j + 1)
            uint32_t hn[4];
            hrow7(rowp + (row + 6) * (BL_BW / 4), hn);
            uint32_t acc[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                acc[k] = __dp2a_lo(pr[i % 6][k], 0x2212u, __dp2a_lo(pr[(i + 2) % 6][k], 0x3830u, __dp2a_lo(pr[(i + 4) % 6][k], 0x2230u, 18u * hn[k] + 32768u)));
                pr[(i + 5) % 6][k] = __byte_perm(prevH[k], hn[k], 0x5410);          // rows (row + 5, row + 6): first used by the next output row
                prevH[k] = hn[k];
            }
            // acc < 2^24: the rounded result is byte 2 of each accumulator
            const uint32_t packed = __byte_perm(__byte_perm(acc[0], acc[1], 0x0062), __byte_perm(acc[2], acc[3], 0x0062), 0x5410);
            if (colOk && y0 + row < h) *reinterpret_cast<uint32_t*>(q) = packed;
            q += pitch;
        }
    }
}

// ---- bilinear resize -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TL_THREADS) resize_tile_kernel(const __grid_constant__ StageParams P)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bars[TL_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tileIdx = blockIdx.x * TL_WARPS + warp;
    if (tileIdx >= P.nTiles) return;
    const TileDesc td = P.tiles[tileIdx];
    const int level = td.level, x0 = td.x0, y0 = td.y0;
    const int slot = P.slot0 + blockIdx.y;
    const int w = P.w[level], h = P.h[level], BW = P.BW[level], BH = P.BH[level];
    const ResizeCoef* tx = P.tab + P.tabX[level];
    const ResizeCoef* ty = P.tab + P.tabY[level];
    const int xs = tx[x0].ofs & ~15, ys = ty[y0].ofs;               // source box origin (x on a 16-byte boundary)
    uint8_t* tile = smem + warp * P.tileStride;
    if (lane == 0) {
        mbar_init(&bars[warp], 1);
        mbar_expect_tx(&bars[warp], (uint32_t)(BW * BH));
        tma_load_3d(tile, &P.maps[level], xs, ys, level == P.srcLevel0 ? slot - P.z0 : slot, &bars[warp]);
    }
    // per-lane column coefficients: 4 adjacent destination columns
    const int x = x0 + 4 * lane;
    int o[4];
    uint32_t pk[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const ResizeCoef c = tx[min(x + k, w - 1)];                 // columns past the edge recompute the last one (never stored)
        o[k] = c.ofs - xs;
        pk[k] = (uint32_t)(uint16_t)c.a0 | ((uint32_t)(uint16_t)c.a1 << 16);
    }
    const int wi0 = o[0] >> 2, sh = 8 * (o[0] & 3);
    uint32_t sel[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { const int d = o[k] - o[0]; sel[k] = (uint32_t)d | ((uint32_t)(d + 1) << 4); }
    __syncwarp();
    mbar_wait(&bars[warp], 0);

    const uint8_t* tcol = tile + 4 * wi0;
#define RS_HROW(r, out)                                                                                             \
    do {   /* horizontally interpolated row r of the tile, already >> 4 (the vertical pass consumes h >> 4) */      \
        const uint32_t* p_ = reinterpret_cast<const uint32_t*>(tcol + (r) * BW);                                    \
        const uint32_t Wa_ = p_[0], Wb_ = p_[1], Wc_ = p_[2];                                                        \
        const uint32_t T0_ = __funnelshift_r(Wa_, Wb_, sh), T1_ = __funnelshift_r(Wb_, Wc_, sh);  /* 8 bytes from column o[0] */ \
        _Pragma("unroll") for (int k = 0; k < 4; ++k) out[k] = __dp2a_lo(pk[k], __byte_perm(T0_, T1_, sel[k]), 0u) >> 4; \
    } while (0)
    uint8_t* q = P.dst[level] + (long long)slot * P.dstFrameStride[level] + (long long)y0 * P.dstPitch[level] + x;
    const int pitch = P.dstPitch[level];
    const bool colOk = x < w;
    uint32_t hA[4], hB[4];
    int rowB = -1;                                                   // tile row held in hB
    const int rows = min(RS_H, h - y0);
    // row coefficients of the tile's RS_H output rows: lane i fetches row i once, the loop takes them by shuffle (no dependent global
    // load inside the loop)
    static_assert(RS_H <= 32, "one lane per output row");
    const ResizeCoef myRow = ty[min(y0 + lane, h - 1)];
    const int myOfs = myRow.ofs - ys;
    const uint32_t myAB = (uint32_t)(uint16_t)myRow.a0 | ((uint32_t)(uint16_t)myRow.a1 << 16);
#pragma unroll 1
    for (int i = 0; i < rows; ++i) {
        const int r = __shfl_sync(0xffffffffu, myOfs, i);
        const uint32_t ab = __shfl_sync(0xffffffffu, myAB, i);
        if (r == rowB) {
#pragma unroll
            for (int k = 0; k < 4; ++k) hA[k] = hB[k];
        } else RS_HROW(r, hA);
        RS_HROW(r + 1, hB);                                            // r + 1 is inside the box; past the source edge its weight is 0
        rowB = r + 1;
        const uint32_t b0 = ab << 16, b1 = ab & 0xFFFF0000u;
        uint32_t v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = (__umulhi(hA[k], b0) + __umulhi(hB[k], b1) + 2u) >> 2;     // (b*(h>>4))>>16 == umulhi(h>>4, b<<16)
        const uint32_t packed = __byte_perm(__byte_perm(v[0], v[1], 0x0040), __byte_perm(v[2], v[3], 0x0040), 0x5410);
        if (colOk) *reinterpret_cast<uint32_t*>(q) = packed;         // pitch % 128 == 0: aligned, in-plane
        q += pitch;
    }
#undef RS_HROW
}

// ---- BGR -> gray (Frame::Frame, Core/frame.cpp:23: cv::cvtColor(imColor, mImGray, CV_BGR2GRAY)) ------------------------------
// OpenCV's 8-bit path is fixed point with 15 fractional bits: (B*3735 + G*19235 + R*9798 + 16384) >> 15 (pinned against
// cv2 4.13.0 in tests/test_ingest.py).  A thread converts 4 pixels: three aligned word reads, one word write.
__global__ void __launch_bounds__(256) bgr2gray_kernel(const uint8_t* __restrict__ bgr, int bgrPitch, long long bgrFrameStride,
    uint8_t* __restrict__ gray, int grayPitch, long long grayFrameStride, int w, int h, int slot0)
{
    const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4, y = blockIdx.y, f = blockIdx.z;
    if (x4 >= w) return;
    const uint8_t* src = bgr + (long long)f * bgrFrameStride + (long long)y * bgrPitch + 3 * x4;
    uint8_t* dst = gray + (long long)(slot0 + f) * grayFrameStride + (long long)y * grayPitch + x4;
    uint8_t px[12];
    if (x4 + 3 < w && ((reinterpret_cast<uintptr_t>(src) & 3) == 0)) {
        const uint32_t* s32 = reinterpret_cast<const uint32_t*>(src);
        *reinterpret_cast<uint32_t*>(px) = __ldg(s32); *reinterpret_cast<uint32_t*>(px + 4) = __ldg(s32 + 1); *reinterpret_cast<uint32_t*>(px + 8) = __ldg(s32 + 2);
    } else {
#pragma unroll
        for (int i = 0; i < 12; ++i) px[i] = (x4 + i / 3 < w) ? __ldg(src + i) : 0;
    }
    uint32_t out = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t v = (px[3 * k] * 3735u + px[3 * k + 1] * 19235u + px[3 * k + 2] * 9798u + 16384u) >> 15;
        out |= v << (8 * k);
    }
    if (x4 + 3 < w) *reinterpret_cast<uint32_t*>(dst) = out;
    else for (int k = 0; x4 + k < w; ++k) dst[k] = (uint8_t)(out >> (8 * k));
}

void fill_common(orbf_context* c, StageParams& P, int slot0)
{
    for (int l = 0; l < c->L; ++l) { P.w[l] = (short)c->lg[l].w; P.h[l] = (short)c->lg[l].h; P.tabX[l] = c->lg[l].tabX; P.tabY[l] = c->lg[l].tabY; }
    P.tab = c->d_resizeTab; P.slot0 = slot0; P.z0 = c->cur_slot0;
}

}  // namespace

int orbf_launch_pyramid(orbf_context* c, int slot0, int n)
{
    if (c->L < 2) return ORBF_OK;
    int r = orbf_refresh_maps(c);
    if (r != ORBF_OK) return r;
    StageParams P;
    fill_common(c, P, slot0);
    P.srcLevel0 = 1;                                                 // level 1 reads the caller's input plane
    int maxBytes = 0;
    for (int l = 1; l < c->L; ++l) {
        P.maps[l] = c->tmResize[l];
        P.dst[l] = c->d_pyr[l]; P.dstFrameStride[l] = (long long)c->lg[l].plane; P.dstPitch[l] = c->lg[l].pitch;
        P.BW[l] = (short)c->rsBW[l]; P.BH[l] = (short)c->rsBH[l];
        maxBytes = std::max(maxBytes, align_up(c->rsBW[l] * c->rsBH[l], 128));
    }
    P.tileStride = maxBytes;
    const size_t smem = (size_t)maxBytes * TL_WARPS + 16;            // + slack: a lane reads up to 11 bytes past its last column
    if (smem > 200 * 1024) return ORBF_ERR_GEOMETRY;
    {
        cudaError_t e = cudaFuncSetAttribute(resize_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return orbf_cuda_fail(c, e, "resize smem attr", __FILE__, __LINE__);
    }
    // the chain level l-1 -> l is a true dependency: one launch per level, each covering every frame of the batch
    for (int l = 1; l < c->L; ++l) {
        P.tiles = c->d_rsTiles + c->rsTile0[l]; P.nTiles = c->rsTileN[l];
        dim3 grid((P.nTiles + TL_WARPS - 1) / TL_WARPS, n);
        resize_tile_kernel<<<grid, TL_THREADS, smem, c->stream>>>(P);
        ORBF_LAUNCH_CHECK(c);
    }
    return ORBF_OK;
}

int orbf_launch_blur(orbf_context* c, int slot0, int n)
{
    int r = orbf_refresh_maps(c);
    if (r != ORBF_OK) return r;
    StageParams P;
    fill_common(c, P, slot0);
    P.srcLevel0 = 0; P.tileStride = BL_TILE_BYTES;
    for (int l = 0; l < c->L; ++l) {
        P.maps[l] = c->tmBlur[l];
        P.dst[l] = c->d_blur[l]; P.dstFrameStride[l] = (long long)c->lg[l].plane; P.dstPitch[l] = c->lg[l].pitch;
        P.BW[l] = BL_BW; P.BH[l] = BL_BH;
    }
    P.tiles = c->d_blTiles; P.nTiles = c->nBlTiles;
    const size_t smem = (size_t)BL_TILE_BYTES * TL_WARPS;
    dim3 grid((P.nTiles + TL_WARPS - 1) / TL_WARPS, n);
    blur_tile_kernel<<<grid, TL_THREADS, smem, c->stream>>>(P);
    ORBF_LAUNCH_CHECK(c);
    return ORBF_OK;
}

int orbf_launch_bgr2gray(orbf_context* c, const uint8_t* d_bgr, int bgrPitch, long long bgrFrameStride, int slot0, int n)
{
    const int w = c->cfg.width, h = c->cfg.height;
    dim3 grid(((w + 3) / 4 + 255) / 256, h, n);
    bgr2gray_kernel<<<grid, 256, 0, c->stream>>>(d_bgr, bgrPitch, bgrFrameStride, c->d_in, c->inPitch, (long long)c->inPlane, w, h, slot0);
    ORBF_LAUNCH_CHECK(c);
    return ORBF_OK;
}

// geometry constants the context needs to build tile tables and TMA boxes
void orbf_stage_tile_geometry(int* tileW, int* blurH, int* blurBW, int* blurBH, int* resizeH)
{
    *tileW = TL_W; *blurH = BL_H; *blurBW = BL_BW; *blurBH = BL_BH; *resizeH = RS_H;
}
