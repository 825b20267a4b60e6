// csrc/pyramid.cu — image pyramid (reference ComputePyramid, Features/orbextractor.cpp:833-857) and the
// 7x7 sigma-2 Gaussian blur (orbextractor.cpp:795-796) as shared-memory-staged stencils.
//
// Level l is cv::resize(level l-1, INTER_LINEAR) — OpenCV's fixed-point bilinear: horizontal
// S[sx]*a0 + S[sx+1]*a1 with 11-bit coefficients, vertical (((b0*(h0>>4))>>16) + ((b1*(h1>>4))>>16) + 2) >> 2.
// The 19-px copyMakeBorder frame of the reference is never read downstream and is not materialised.
// The blur is OpenCV's fixed-point Gaussian: taps [18,34,48,56,48,34,18]/256, horizontal Q8.8 (u16),
// vertical Q16.16, (v + 32768) >> 16, BORDER_REFLECT_101 at the true level edge.
// HBM-bound: algorithmic bytes per 640x480 frame = 926,546 rd + 643,332 wr (resize), 950,532 rd + wr (blur).
#include "orbf_internal.h"

namespace {

constexpr int RS_TW = 64, RS_TH = 16, RS_THREADS = 256;

__global__ void __launch_bounds__(RS_THREADS) resize_kernel(LevelView src, uint8_t* __restrict__ dstBase,
    long long dstFrameStride, int dstPitch, int dw, int dh, const ResizeCoef* __restrict__ tx,
    const ResizeCoef* __restrict__ ty, int slot0, int regPitch, int regRows)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const int slot = slot0 + blockIdx.z;
    const uint8_t* sImg = src.base + (long long)slot * src.frameStride;
    uint8_t* dImg = dstBase + (long long)slot * dstFrameStride;
    const int x0 = blockIdx.x * RS_TW, y0 = blockIdx.y * RS_TH;
    const int x1 = min(x0 + RS_TW, dw) - 1, y1 = min(y0 + RS_TH, dh) - 1;
    const int sxFirst = tx[x0].ofs, sxLast = min(tx[x1].ofs + 1, src.w - 1);
    const int syFirst = ty[y0].ofs, syLast = min(ty[y1].ofs + 1, src.h - 1);
    const int sxA = sxFirst & ~15;                       // 16-byte aligned start column
    const int vecPerRow = (sxLast - sxA) / 16 + 1;
    const int rows = syLast - syFirst + 1;
    // stage the source footprint with 16-byte vector loads (rows are 16-byte aligned: pitch % 16 == 0)
    for (int i = threadIdx.x; i < rows * vecPerRow; i += RS_THREADS) {
        const int r = i / vecPerRow, v = i - r * vecPerRow;
        const int gx = sxA + v * 16;
        uint4 val = make_uint4(0, 0, 0, 0);
        if (gx + 16 <= src.pitch) val = __ldg(reinterpret_cast<const uint4*>(sImg + (long long)(syFirst + r) * src.pitch + gx));
        else for (int b = 0; b < 16 && gx + b < src.pitch; ++b)
            reinterpret_cast<uint8_t*>(&val)[b] = sImg[(long long)(syFirst + r) * src.pitch + gx + b];
        *reinterpret_cast<uint4*>(smem + r * regPitch + v * 16) = val;
    }
    __syncthreads();
    const int ly = threadIdx.x / (RS_TW / 4), lx = (threadIdx.x % (RS_TW / 4)) * 4;
    const int y = y0 + ly, xb = x0 + lx;
    if (y >= dh || xb >= dw) return;
    const ResizeCoef cy = ty[y];
    const uint8_t* r0 = smem + (cy.ofs - syFirst) * regPitch - sxA;
    const uint8_t* r1 = smem + (min(cy.ofs + 1, src.h - 1) - syFirst) * regPitch - sxA;
    const int b0 = cy.a0, b1 = cy.a1;
    uint32_t packed = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int x = xb + k;
        if (x < dw) {
            const ResizeCoef cx = tx[x];
            const int sx = cx.ofs, sx1 = min(sx + 1, src.w - 1);
            const int h0 = r0[sx] * cx.a0 + r0[sx1] * cx.a1;
            const int h1 = r1[sx] * cx.a0 + r1[sx1] * cx.a1;
            const int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
            packed |= (uint32_t)(v & 255) << (8 * k);
        }
    }
    *reinterpret_cast<uint32_t*>(dImg + (long long)y * dstPitch + xb) = packed;   // pitch % 128 == 0: aligned, in-plane
}

constexpr int BL_TW = 64, BL_TH = 32, BL_THREADS = 256;
constexpr int BL_RW = BL_TW + 6, BL_RH = BL_TH + 6, BL_RP = 72;

__device__ __forceinline__ int reflect101(int p, int n)
{
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = (p < 0) ? -p : 2 * (n - 1) - p;
    return p;
}

__global__ void __launch_bounds__(BL_THREADS) blur7_kernel(LevelView src, uint8_t* __restrict__ dstBase,
    long long dstFrameStride, int dstPitch, int slot0)
{
    __shared__ uint8_t sIn[BL_RH][BL_RP];
    __shared__ uint16_t sH[BL_RH][BL_TW];
    const int slot = slot0 + blockIdx.z;
    const uint8_t* sImg = src.base + (long long)slot * src.frameStride;
    uint8_t* dImg = dstBase + (long long)slot * dstFrameStride;
    const int w = src.w, h = src.h;
    const int x0 = blockIdx.x * BL_TW, y0 = blockIdx.y * BL_TH;
    for (int i = threadIdx.x; i < BL_RH * BL_RW; i += BL_THREADS) {
        const int r = i / BL_RW, c = i - r * BL_RW;
        const int gy = reflect101(y0 - 3 + r, h), gx = reflect101(x0 - 3 + c, w);
        sIn[r][c] = __ldg(sImg + (long long)gy * src.pitch + gx);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < BL_RH * BL_TW; i += BL_THREADS) {
        const int r = i / BL_TW, c = i - r * BL_TW;
        const uint8_t* p = &sIn[r][c];
        const int acc = 18 * (p[0] + p[6]) + 34 * (p[1] + p[5]) + 48 * (p[2] + p[4]) + 56 * p[3];
        sH[r][c] = (uint16_t)acc;
    }
    __syncthreads();
    // each thread: 4 adjacent columns x 2 rows -> two aligned 32-bit stores
    const int cx = (threadIdx.x % 16) * 4, ry = (threadIdx.x / 16) * 2;
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
        const int y = y0 + ry + rr;
        if (y >= h) break;
        uint32_t packed = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int c = cx + k;
            const uint32_t acc = 18u * (sH[ry + rr][c] + sH[ry + rr + 6][c]) + 34u * (sH[ry + rr + 1][c] + sH[ry + rr + 5][c])
                + 48u * (sH[ry + rr + 2][c] + sH[ry + rr + 4][c]) + 56u * sH[ry + rr + 3][c];
            packed |= ((acc + 32768u) >> 16) << (8 * k);
        }
        if (x0 + cx < dstPitch) *reinterpret_cast<uint32_t*>(dImg + (long long)y * dstPitch + x0 + cx) = packed;
    }
}

}  // namespace

int orbf_launch_pyramid(orbf_context* c, int slot0, int n)
{
    PyrView pv = orbf_pyr_view(c, false);
    for (int l = 1; l < c->L; ++l) {
        const LevelGeom& g = c->lg[l];
        const LevelGeom& s = c->lg[l - 1];
        const double sx = (double)s.w / g.w, sy = (double)s.h / g.h;
        const int regPitch = align_up((int)(RS_TW * sx) + 4 + 16, 16) + 16;
        const int regRows = (int)(RS_TH * sy) + 4;
        const size_t smem = (size_t)regPitch * regRows;
        dim3 grid((g.w + RS_TW - 1) / RS_TW, (g.h + RS_TH - 1) / RS_TH, n);
        resize_kernel<<<grid, RS_THREADS, smem, c->stream>>>(pv.lv[l - 1], c->d_pyr[l], (long long)g.plane, g.pitch, g.w, g.h,
            c->d_resizeTab + g.tabX, c->d_resizeTab + g.tabY, slot0, regPitch, regRows);
        ORBF_LAUNCH_CHECK(c);
    }
    return ORBF_OK;
}

int orbf_launch_blur(orbf_context* c, int slot0, int n)
{
    PyrView pv = orbf_pyr_view(c, false);
    for (int l = 0; l < c->L; ++l) {
        const LevelGeom& g = c->lg[l];
        dim3 grid((g.w + BL_TW - 1) / BL_TW, (g.h + BL_TH - 1) / BL_TH, n);
        blur7_kernel<<<grid, BL_THREADS, 0, c->stream>>>(pv.lv[l], c->d_blur[l], (long long)g.plane, g.pitch, slot0);
        ORBF_LAUNCH_CHECK(c);
    }
    return ORBF_OK;
}
