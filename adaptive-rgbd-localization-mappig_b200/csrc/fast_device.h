// csrc/fast_device.h — FAST-9/16 device helpers shared by fast.cu (per-cell ORB-SLAM2 route) and adaptive.cu
// (whole-image response plane of the adaptive-threshold route).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

// Corner strength without ever negating a min/max result (nvcc 12.9 for sm_100a loses the negation when ptxas fuses
// max(x, -max(...)) chains into VIMNMX3; tools/nvcc_minmax_bug.cu reproduces it).  Scalar form: fallback path only.
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

// Strength of two pixels at once: lane halves hold pixel a (low) and pixel b (high), every value <= 255.
__device__ __forceinline__ uint32_t ring_strength_x2(const uint8_t* pa, const uint8_t* pb, int rp)
{
    const int o[16] = { 3 * rp, 3 * rp + 1, 2 * rp + 2, rp + 3, 3, -rp + 3, -2 * rp + 2, -3 * rp + 1,
        -3 * rp, -3 * rp - 1, -2 * rp - 2, -rp - 3, -3, rp - 3, 2 * rp - 2, 3 * rp - 1 };
    uint32_t r[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) r[k] = (uint32_t)pa[o[k]] + ((uint32_t)pb[o[k]] << 16);
    uint32_t mx3[16], mn3[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        mx3[k] = __vimax3_u16x2(r[k], r[(k + 1) & 15], r[(k + 2) & 15]);
        mn3[k] = __vimin3_u16x2(r[k], r[(k + 1) & 15], r[(k + 2) & 15]);
    }
    uint32_t mx9[16], mn9[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {      // arc k = ring pixels k .. k+8
        mx9[k] = __vimax3_u16x2(mx3[k], mx3[(k + 3) & 15], mx3[(k + 6) & 15]);
        mn9[k] = __vimin3_u16x2(mn3[k], mn3[(k + 3) & 15], mn3[(k + 6) & 15]);
    }
    uint32_t A = __vimin3_u16x2(mx9[0], mx9[1], mx9[2]), B = __vimax3_u16x2(mn9[0], mn9[1], mn9[2]);
#pragma unroll
    for (int k = 3; k < 15; k += 2) { A = __vimin3_u16x2(A, mx9[k], mx9[k + 1]); B = __vimax3_u16x2(B, mn9[k], mn9[k + 1]); }
    A = __vminu2(A, mx9[15]); B = __vmaxu2(B, mn9[15]);
    const uint32_t v = (uint32_t)pa[0] + ((uint32_t)pb[0] << 16);
    // max(v - A, 0) and max(B - v, 0) per half: both differences are formed on values ordered first, so no borrow crosses halves
    return __vmaxu2(__vmaxu2(v, A) - A, __vmaxu2(B, v) - v);
}

// pass flags of one pixel pair (halves = two adjacent pixels): a half of the result is 0 iff that pixel passes
__device__ __forceinline__ uint32_t pretest_x2(uint32_t v, uint32_t p0, uint32_t p8, uint32_t p4, uint32_t p12, uint32_t th1)
{
    const uint32_t mb = __vminu2(__vmaxu2(p0, p8), __vmaxu2(p4, p12));     // bright: both opposite pairs hold a pixel > v + th
    const uint32_t md = __vmaxu2(__vminu2(p0, p8), __vminu2(p4, p12));     // dark:   both opposite pairs hold a pixel < v - th
    const uint32_t vh = v + th1, dh = md + th1;                           // th1 = (th + 1) in both halves; sums < 2^16
    const uint32_t fb = __vminu2(mb, vh) ^ vh;                            // 0 iff mb >= v + th + 1
    const uint32_t fd = __vmaxu2(dh, v) ^ v;                              // 0 iff md + th + 1 <= v
    return __vminu2(fb, fd);
}

}  // namespace
