// oracle/ref_utils/utils_stubs.h — force-included (-include) in front of the reference's Utils/utils.cpp so that the WHOLE file compiles
// verbatim although only LoadImages (utils.cpp:16-38) is called: declarations for the OpenCV / Eigen / PCL entry points its other
// functions (FindHomography, DistanceFiler, TestRecallPrecision, AddNormal — off the hot path, never executed here) mention.  Every
// stand-in that would have to compute something aborts.  Test infrastructure (oracle/_ref), not product code.
#pragma once
#include <cstdlib>
#include <opencv2/opencv.hpp>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <Eigen/Core>

#ifndef CV_64F
#define CV_64F 6
#endif
#ifndef CV_RANSAC
#define CV_RANSAC 8
#endif

namespace cv {
struct Exception {};
struct Vec2f {
    float x, y;
    Vec2f(float a, float b) : x(a), y(b) {}
    Vec2f operator-(const Vec2f& o) const { return Vec2f(x - o.x, y - o.y); }
};
inline double norm(const Vec2f&) { std::abort(); }
inline double norm(const Vec2f&, const Vec2f&) { std::abort(); }
template <typename... A> Mat findHomography(A&&...) { std::abort(); }
}  // namespace cv

namespace Eigen {
struct ArrayXd {
    static ArrayXd LinSpaced(int, double, double) { std::abort(); }
    int size() const { return 0; }
    double operator()(int, int) const { return 0.0; }
};
}  // namespace Eigen

namespace pcl {
struct Normal {};
namespace search {
template <typename P> struct KdTree {
    typedef KdTree* Ptr;
    template <typename C> void setInputCloud(const C&) {}
};
}  // namespace search
template <typename P, typename N> struct NormalEstimationOMP {
    template <typename C> void setInputCloud(const C&) {}
    template <typename T> void setSearchMethod(const T&) {}
    void setKSearch(int) {}
    template <typename C> void compute(C&) { std::abort(); }
};
template <typename A, typename B, typename C> void concatenateFields(const A&, const B&, C&) { std::abort(); }
}  // namespace pcl
