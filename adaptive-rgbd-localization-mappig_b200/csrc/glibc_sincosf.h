// csrc/glibc_sincosf.h — sinf / cosf exactly as glibc (>= 2.28; pinned here against 2.39) evaluates them, for host and device.
//
// Why: computeOrbDescriptor (reference Features/orbextractor.cpp:45-46) writes  float a = (float)cos(angle), b = (float)sin(angle);
// with a FLOAT argument under `using namespace std`, so overload resolution picks std::cos(float) / std::sin(float) — libm's
// cosf / sinf, not the double functions.  (the reference-source build used by the tests compiles that very line; (float)cos((double)angle) differs from cosf(angle)
// in the last bit for ~2.6 % of the angles, which can move a rotated test point across a rounding boundary.)  Like std::sort and
// rand() (csrc/replay.h) this is library behaviour the reference inherits, so it is replayed, not approximated.
//
// Algorithm (glibc sysdeps/ieee754/flt-32/s_sinf.c, s_cosf.c, sincosf.h; from ARM's optimized-routines): everything in double —
// |x| < pi/4: polynomial directly; |x| < 120: n = round(x * 2/pi) through a scaled float-to-int conversion, r = x - n * pi/2,
// sine or cosine polynomial by quadrant parity, sign by quadrant; one rounding to float at the end.  Coefficients read out of the
// libm.so.6 of this image (__sincosf_table).  Valid for 0 <= x < 120 (the descriptor's angle is fastAtan2 degrees * pi/180 <= 2 pi);
// the host version is compared with libm over EVERY float in [0, 6.2832] (tests/test_replay_sincosf.py), the device version with
// the host version.  No operation here may be contracted into an FMA (the library is built with -fmad=false / -ffp-contract=off;
// the exhaustive test shows the FMA build of glibc returns the same floats on this domain).
#pragma once
#include <stdint.h>
#include <string.h>

#ifndef ORBF_HD
#ifdef __CUDACC__
#define ORBF_HD __host__ __device__
#else
#define ORBF_HD
#endif
#endif

namespace replay {

ORBF_HD inline uint32_t f32_bits(float f)
{
#ifdef __CUDA_ARCH__
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
ORBF_HD inline uint32_t abstop12(float x) { return (f32_bits(x) >> 20) & 0x7ffu; }

// quadrant parity even: sine polynomial of (x, x2 = x * x); odd: cosine polynomial; neg: the cosine polynomial negated (n & 2)
ORBF_HD inline float sincosf_poly(double x, double x2, int n, bool neg)
{
    const double c0 = 0x1p0, c1 = -0x1.ffffffd0c621cp-2, c2 = 0x1.55553e1068f19p-5, c3 = -0x1.6c087e89a359dp-10, c4 = 0x1.99343027bf8c3p-16;
    const double s1 = -0x1.555545995a603p-3, s2 = 0x1.1107605230bc4p-7, s3 = -0x1.994eb3774cf24p-13;
    if ((n & 1) == 0) {
        const double x3 = x * x2;
        const double t1 = s2 + x2 * s3;
        const double x5 = x3 * x2;
        const double s = x + x3 * s1;
        return (float)(s + x5 * t1);
    }
    const double sg = neg ? -1.0 : 1.0;                   // table[1] of glibc holds the negated cosine coefficients: same magnitudes
    const double x4 = x2 * x2;
    const double t2 = sg * c3 + x2 * (sg * c4);
    const double t1 = sg * c0 + x2 * (sg * c1);
    const double x6 = x4 * x2;
    const double c = t1 + x4 * (sg * c2);
    return (float)(c + x6 * t2);
}

ORBF_HD inline double sincosf_reduce(double x, int* np)
{
    const double hpi_inv = 0x1.45F306DC9C883p+23, hpi = 0x1.921FB54442D18p0;      // 2/pi * 2^24, pi/2
    const double r = x * hpi_inv;
    const int n = ((int32_t)r + 0x800000) >> 24;
    *np = n;
    return x - (double)n * hpi;
}

ORBF_HD inline float glibc_sinf(float y)              // 0 <= y < 120
{
    double x = (double)y;
    if (abstop12(y) < abstop12(0x1.921FB6p-1f)) {
        if (abstop12(y) < abstop12(0x1p-12f)) return y;
        return sincosf_poly(x, x * x, 0, false);
    }
    int n;
    x = sincosf_reduce(x, &n);
    const double s = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
    return sincosf_poly(x * s, x * x, n, (n & 2) != 0);
}

ORBF_HD inline float glibc_cosf(float y)              // 0 <= y < 120
{
    double x = (double)y;
    if (abstop12(y) < abstop12(0x1.921FB6p-1f)) {
        if (abstop12(y) < abstop12(0x1p-12f)) return 1.0f;
        return sincosf_poly(x, x * x, 1, false);
    }
    int n;
    x = sincosf_reduce(x, &n);
    const double s = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
    return sincosf_poly(x * s, x * x, n ^ 1, (n & 2) != 0);
}

}  // namespace replay
