// csrc/undistort_device.h — cv::undistortPoints(pts, pts, K, dist, Mat(), K) for one point, as Frame::UndistortKeyPoints calls it
// (Core/frame.cpp:286-313): cvUndistortPointsInternal with the default criteria (5 fixed iterations), every operation in double in
// OpenCV's order (the library is compiled with -fmad=false, so nothing contracts).  Shared by projection.cu (orbf_undistort_points)
// and describe.cu (mvKeysUn feeding mvuRight / mvKeys3Dc).
#pragma once
#include <cuda_runtime.h>

struct UndistortParams { double fx, fy, cx, cy, k1, k2, p1, p2, k3; };

__device__ __forceinline__ void undistort_point(const UndistortParams& U, float pxf, float pyf, float* ox, float* oy)
{
    const double px = pxf, py = pyf, ifx = 1.0 / U.fx, ify = 1.0 / U.fy;
    double x = (px - U.cx) * ifx, y = (py - U.cy) * ify;
    const double x0 = x, y0 = y;
    for (int j = 0; j < 5; ++j) {
        const double r2 = x * x + y * y;
        const double icdist = (1 + ((0.0 * r2 + 0.0) * r2 + 0.0) * r2) / (1 + ((U.k3 * r2 + U.k2) * r2 + U.k1) * r2);
        if (icdist < 0) { x = (px - U.cx) * ifx; y = (py - U.cy) * ify; break; }
        const double deltaX = 2 * U.p1 * x * y + U.p2 * (r2 + 2 * x * x) + 0.0 * r2 + 0.0 * r2 * r2;
        const double deltaY = U.p1 * (r2 + 2 * y * y) + 2 * U.p2 * x * y + 0.0 * r2 + 0.0 * r2 * r2;
        x = (x0 - deltaX) * icdist;
        y = (y0 - deltaY) * icdist;
    }
    const double xx = U.fx * x + 0.0 * y + U.cx, yy = 0.0 * x + U.fy * y + U.cy, ww = 1.0 / (0.0 * x + 0.0 * y + 1.0);
    *ox = (float)(xx * ww);
    *oy = (float)(yy * ww);
}
