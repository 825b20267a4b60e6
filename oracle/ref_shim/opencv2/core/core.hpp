// oracle/ref_shim/opencv2/core/core.hpp — TEST INFRASTRUCTURE ONLY.
// A minimal stand-in for the parts of OpenCV that /root/reference/Features/orbextractor.cpp (and the adaptive-detector sources
// next to it) touch, so that those reference sources compile VERBATIM, from where they lie, with g++ alone (oracle/Makefile, target
// _ref).  The container types (Mat with ROI views and reference counting, KeyPoint, Point, Size, Rect, Input/OutputArray) are
// written here; the image-processing entry points (resize, FAST, GaussianBlur, fastAtan2) delegate to the routines of
// oracle/orb_oracle.cpp that tests/test_oracle_vs_cv2.py pins bit for bit against the real library (cv2 4.13).  What the _ref build
// therefore adds is the reference-AUTHORED logic run from the reference's own source: cell grid, FAST fallback, quadtree
// distribution, orientation, steering, descriptor assembly, output order.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <list>
#include <map>
#include <memory>
#include <set>
#include <string>
#include <vector>

#include "../../../oracle_api.h"

typedef unsigned char uchar;
#define CV_8U 0
#define CV_8UC1 0
#define CV_16U 2
#define CV_32F 5
#define CV_32FC1 5
#define CV_32FC2 13
#define CV_8UC3 16
#define CV_PI 3.1415926535897932384626433832795
#define CV_WRAP
#define CV_OUT
#define CV_IN_OUT
#define CV_EXPORTS
#define CV_Assert(x) assert(x)

inline int cvRound(double v) { return (int)std::nearbyint(v); }       // round half to even, like the SSE2 cvtsd2si OpenCV uses
inline int cvRound(float v) { return (int)std::nearbyintf(v); }
inline int cvRound(int v) { return v; }
inline int cvFloor(double v) { return (int)std::floor(v); }
inline int cvCeil(double v) { return (int)std::ceil(v); }

namespace cv {

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
    template <typename U> Point_(const Point_<U>& o) : x((T)o.x), y((T)o.y) {}
    Point_& operator*=(float s) { x = (T)(x * s); y = (T)(y * s); return *this; }
};
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;
template <typename T> struct Point3_ {
    T x, y, z;
    Point3_() : x(0), y(0), z(0) {}
    Point3_(T x_, T y_, T z_) : x(x_), y(y_), z(z_) {}
};
typedef Point3_<float> Point3f;

struct DMatch {
    int queryIdx, trainIdx, imgIdx; float distance;
    DMatch() : queryIdx(-1), trainIdx(-1), imgIdx(-1), distance(3.402823466e+38f) {}
    DMatch(int q, int t, float d) : queryIdx(q), trainIdx(t), imgIdx(-1), distance(d) {}
    DMatch(int q, int t, int i, float d) : queryIdx(q), trainIdx(t), imgIdx(i), distance(d) {}
    bool operator<(const DMatch& m) const { return distance < m.distance; }       // cv::DMatch::operator< (types.hpp)
};

struct Size { int width, height; Size() : width(0), height(0) {} Size(int w, int h) : width(w), height(h) {} };
struct Rect { int x, y, width, height; Rect() : x(0), y(0), width(0), height(0) {} Rect(int x_, int y_, int w, int h) : x(x_), y(y_), width(w), height(h) {} };
struct Range { int start, end; Range() : start(0), end(0) {} Range(int s, int e) : start(s), end(e) {} };
struct Scalar {
    double v[4];
    Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { v[0] = a; v[1] = b; v[2] = c; v[3] = d; }
    static Scalar all(double a) { return Scalar(a, a, a, a); }
};

struct KeyPoint {
    Point2f pt; float size, angle, response; int octave, class_id;
    KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(float x, float y, float size_, float angle_ = -1, float response_ = 0, int octave_ = 0, int class_id_ = -1)
        : pt(x, y), size(size_), angle(angle_), response(response_), octave(octave_), class_id(class_id_) {}
};

class Mat;
struct MatZeros { int rows, cols, type; };            // what Mat::zeros returns: assigned INTO an existing Mat of the same shape

class Mat {
public:
    int rows, cols; size_t step; uchar* data;
    Mat() : rows(0), cols(0), step(0), data(nullptr), type_(0), esz_(1) {}
    Mat(int r, int c, int type) : Mat() { create(r, c, type); }
    Mat(Size sz, int type) : Mat() { create(sz.height, sz.width, type); }
    Mat(int r, int c, int type, void* ext, size_t stp = 0) : rows(r), cols(c), step(stp ? stp : (size_t)c * esz(type)), data((uchar*)ext), type_(type), esz_(esz(type)) {}
    Mat(const MatZeros& z) : Mat() { *this = z; }
    Mat(const struct MatMul& m);                                       // A * B
    static int esz(int type) { return type == CV_32F ? 4 : type == CV_32FC2 ? 8 : type == CV_16U ? 2 : type == CV_8UC3 ? 3 : 1; }
    int channels() const { return type_ == CV_32FC2 ? 2 : type_ == CV_8UC3 ? 3 : 1; }
    void create(int r, int c, int type)
    {
        if (data && r == rows && c == cols && type == type_) return;       // cv::Mat::create: no reallocation for the same shape
        buf_ = std::shared_ptr<uchar>(new uchar[(size_t)std::max(r, 0) * std::max(c, 0) * esz(type) + 64], std::default_delete<uchar[]>());
        rows = r; cols = c; type_ = type; esz_ = esz(type); step = (size_t)c * esz_; data = buf_.get();
    }
    void release() { buf_.reset(); rows = cols = 0; step = 0; data = nullptr; }
    bool empty() const { return !data || rows == 0 || cols == 0; }
    int type() const { return type_; }
    size_t step1() const { return step / esz_; }
    Size size() const { return Size(cols, rows); }
    Mat operator()(const Rect& r) const { Mat m(*this); m.data = data + (size_t)r.y * step + (size_t)r.x * esz_; m.rows = r.height; m.cols = r.width; return m; }
    Mat operator()(const Range& rr, const Range& cr) const { return (*this)(Rect(cr.start, rr.start, cr.end - cr.start, rr.end - rr.start)); }
    Mat row(int r) const { return rowRange(r, r + 1); }
    Mat rowRange(int a, int b) const { Mat m(*this); m.data = data + (size_t)a * step; m.rows = b - a; return m; }
    Mat colRange(int a, int b) const { Mat m(*this); m.data = data + (size_t)a * esz_; m.cols = b - a; return m; }
    Mat clone() const
    {
        Mat m(rows, cols, type_);
        for (int r = 0; r < rows; ++r) std::memcpy(m.data + (size_t)r * m.step, data + (size_t)r * step, (size_t)cols * esz_);
        return m;
    }
    void copyTo(const class _OutputArray& dst) const;                  // defined below _OutputArray (a temporary view is a valid destination)
    Mat col(int c) const { return colRange(c, c + 1); }
    Mat t() const
    {
        assert(type_ == CV_32F);
        Mat m(cols, rows, CV_32F);
        for (int r = 0; r < rows; ++r) for (int c = 0; c < cols; ++c) m.at<float>(c, r) = at<float>(r, c);
        return m;
    }
    static Mat eye(int r, int c, int type)
    {
        assert(type == CV_32F);
        Mat m(r, c, type);
        for (int i = 0; i < r; ++i) for (int j = 0; j < c; ++j) m.at<float>(i, j) = i == j ? 1.f : 0.f;
        return m;
    }
    void resize(size_t nrows)                                          // cv::Mat::resize: the first rows are kept
    {
        Mat m((int)nrows, cols, type_);
        std::memset(m.data, 0, (size_t)m.rows * m.step);
        for (int r = 0; r < std::min(rows, m.rows); ++r) std::memcpy(m.data + (size_t)r * m.step, data + (size_t)r * step, (size_t)cols * esz_);
        *this = m;
    }
    Mat reshape(int cn) const                                          // N x 2 CV_32F <-> N x 1 CV_32FC2 on the same data (continuous only)
    {
        assert(step == (size_t)cols * esz_);
        Mat m(*this);
        if (cn == 2) { assert(type_ == CV_32F && cols == 2); m.type_ = CV_32FC2; m.esz_ = 8; m.cols = 1; }
        else { assert(cn == 1 && type_ == CV_32FC2); m.type_ = CV_32F; m.esz_ = 4; m.cols = cols * 2; }
        m.step = (size_t)m.cols * m.esz_;
        return m;
    }
    // cv::Mat::convertTo(dst, CV_32F, alpha) from 16-bit unsigned: float(src) * float(alpha), one rounding (what the oracle's
    // unprojection restates; pinned against cv2 in tests/test_ingest.py)
    void convertTo(Mat& dst, int rtype, double alpha = 1.0) const
    {
        assert(type_ == CV_16U && rtype == CV_32F);
        dst.create(rows, cols, CV_32F);
        const float a = (float)alpha;
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < cols; ++c) dst.at<float>(r, c) = (float)*reinterpret_cast<const uint16_t*>(data + (size_t)r * step + (size_t)c * 2) * a;
    }
    template <typename T> T* ptr(int r) { return reinterpret_cast<T*>(data + (size_t)r * step); }
    template <typename T> const T* ptr(int r) const { return reinterpret_cast<const T*>(data + (size_t)r * step); }
    template <typename T> T& at(int r, int c) { return *reinterpret_cast<T*>(data + (size_t)r * step + (size_t)c * sizeof(T)); }
    template <typename T> const T& at(int r, int c) const { return *reinterpret_cast<const T*>(data + (size_t)r * step + (size_t)c * sizeof(T)); }
    template <typename T> T& at(int i) { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }                 // vectors
    template <typename T> const T& at(int i) const { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }
    uchar* ptr(int r = 0) { return data + (size_t)r * step; }
    const uchar* ptr(int r = 0) const { return data + (size_t)r * step; }
    static MatZeros zeros(int r, int c, int type) { MatZeros z = { r, c, type }; return z; }
    // m = Mat::zeros(...): cv::MatExpr assignment creates (a no-op for an unchanged shape, so a rowRange view keeps pointing into its
    // parent) and then clears in place
    Mat& operator=(const MatZeros& z)
    {
        create(z.rows, z.cols, z.type);
        for (int r = 0; r < rows; ++r) std::memset(data + (size_t)r * step, 0, (size_t)cols * esz_);
        return *this;
    }
private:
    int type_, esz_;
    std::shared_ptr<uchar> buf_;
};

// cv::InputArray / cv::OutputArray: proxies around a Mat (all the reference passes through them)
class _InputArray {
public:
    _InputArray() : m_(nullptr) {}
    _InputArray(const Mat& m) : m_(const_cast<Mat*>(&m)) {}
    bool empty() const { return !m_ || m_->empty(); }
    Mat getMat() const { return m_ ? *m_ : Mat(); }
protected:
    Mat* m_;
};
class _OutputArray : public _InputArray {
public:
    _OutputArray() {}
    _OutputArray(Mat& m) : _InputArray(m) {}
    _OutputArray(const Mat& m) : _InputArray(m) {}                     // OpenCV allows a temporary header (a view) as destination
    void create(int r, int c, int type) const { if (m_) m_->create(r, c, type); }
    void release() const { if (m_) m_->release(); }
};
inline void Mat::copyTo(const _OutputArray& dst_) const
{
    dst_.create(rows, cols, type_);                                    // no reallocation for an unchanged shape: a view keeps pointing into its parent
    Mat dst = dst_.getMat();
    for (int r = 0; r < rows; ++r) std::memmove(dst.data + (size_t)r * dst.step, data + (size_t)r * step, (size_t)cols * esz_);
}
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;
inline const _OutputArray& noArray() { static _OutputArray none; return none; }

// `A * B + C` on float matrices = cv::gemm(A, B, 1, C, 1) through cv::MatExpr: every element is the products summed left to right in
// float, then (float)((double)sum * alpha + (double)c * beta) (OpenCV's small-matrix GEMM; the same statement as the oracle's, which
// tests/test_fuse_bow.py and tests/test_trajectory.py pin against cv2.gemm)
struct MatMul {
    Mat a, b;
    operator Mat() const { return eval(); }
    Mat eval() const                                                   // A * B alone: gemm with beta = 0
    {
        Mat d(a.rows, b.cols, CV_32F);
        for (int i = 0; i < d.rows; ++i)
            for (int j = 0; j < d.cols; ++j) {
                float t = a.at<float>(i, 0) * b.at<float>(0, j);
                for (int k = 1; k < a.cols; ++k) t = t + a.at<float>(i, k) * b.at<float>(k, j);
                d.at<float>(i, j) = (float)((double)t * 1.0);
            }
        return d;
    }
};
inline Mat::Mat(const MatMul& m) : Mat() { *this = m.eval(); }
inline MatMul operator*(const Mat& a, const Mat& b) { assert(a.type() == CV_32F && b.type() == CV_32F && a.cols == b.rows); MatMul m = { a, b }; return m; }
inline Mat operator-(const Mat& a)
{
    assert(a.type() == CV_32F);
    Mat d(a.rows, a.cols, CV_32F);
    for (int i = 0; i < a.rows; ++i) for (int j = 0; j < a.cols; ++j) d.at<float>(i, j) = -a.at<float>(i, j);
    return d;
}
// cv::Mat_<float>(r, c) << v0, v1, ...
template <typename T> class Mat_ : public Mat {
public:
    Mat_(int r, int c) : Mat(r, c, CV_32F) { static_assert(sizeof(T) == 4, "float only"); }
    struct Init {
        Mat m; int k;
        Init& operator,(T v) { m.at<T>(k / m.cols, k % m.cols) = v; ++k; return *this; }
        operator Mat() const { return m; }
    };
    Init operator<<(T v) { Init i = { *this, 0 }; i, v; return i; }
};
inline Mat operator+(const MatMul& m, const Mat& c)
{
    assert(c.type() == CV_32F && c.rows == m.a.rows && c.cols == m.b.cols);
    Mat d(m.a.rows, m.b.cols, CV_32F);
    for (int i = 0; i < d.rows; ++i)
        for (int j = 0; j < d.cols; ++j) {
            float t = m.a.at<float>(i, 0) * m.b.at<float>(0, j);
            for (int k = 1; k < m.a.cols; ++k) t = t + m.a.at<float>(i, k) * m.b.at<float>(k, j);
            d.at<float>(i, j) = (float)((double)t * 1.0 + (double)c.at<float>(i, j) * 1.0);
        }
    return d;
}

enum { NORM_L2 = 4, NORM_HAMMING = 6 };
// cv::norm(a, b, NORM_HAMMING) on two u8 rows: number of differing bits
inline double norm(InputArray a_, InputArray b_, int normType)
{
    const Mat a = a_.getMat(), b = b_.getMat();
    assert(normType == NORM_HAMMING && a.type() == CV_8U && a.rows == b.rows && a.cols == b.cols);
    int d = 0;
    for (int r = 0; r < a.rows; ++r) for (int c = 0; c < a.cols; ++c) d += __builtin_popcount((unsigned)(a.ptr(r)[c] ^ b.ptr(r)[c]));
    return (double)d;
}

enum { BORDER_CONSTANT = 0, BORDER_REFLECT_101 = 4, BORDER_DEFAULT = 4, BORDER_ISOLATED = 16 };
enum { INTER_LINEAR = 1 };

inline float fastAtan2(float y, float x) { return orc_fast_atan2(y, x); }

// cv::FAST(image, keypoints, threshold, nonmaxSuppression): FAST-9/16 on the matrix handed in (a view is its own image)
inline void FAST(InputArray image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression = true)
{
    const Mat m = image.getMat();
    assert(nonmaxSuppression && m.type() == CV_8UC1);
    keypoints.clear();
    if (m.rows < 7 || m.cols < 7) return;
    std::vector<orc_cand> out((size_t)m.rows * m.cols / 2 + 16);
    int n = 0;
    const int rc = orc_fast_roi(m.data, (int)m.step, m.cols, m.rows, threshold, out.data(), (int)out.size(), &n);
    assert(rc == ORC_OK); (void)rc;
    for (int i = 0; i < n; ++i) keypoints.push_back(KeyPoint((float)out[i].x, (float)out[i].y, 7.f, -1.f, (float)out[i].score));
}

inline void resize(InputArray src_, OutputArray dst_, Size dsize, double = 0, double = 0, int interpolation = INTER_LINEAR)
{
    assert(interpolation == INTER_LINEAR);
    const Mat src = src_.getMat();
    dst_.create(dsize.height, dsize.width, src.type());
    Mat dst = dst_.getMat();
    const int rc = orc_resize_linear(src.data, src.cols, src.rows, (int)src.step, dst.data, dst.cols, dst.rows, (int)dst.step);
    assert(rc == ORC_OK); (void)rc;
}

// cv::copyMakeBorder, BORDER_REFLECT_101 (gfedcb|abcdefgh|gfedcba).  src may be a view into dst (the reference builds its padded
// pyramid that way, orbextractor.cpp:841-851): the interior is copied only when it lives elsewhere.
inline void copyMakeBorder(InputArray src_, OutputArray dst_, int top, int bottom, int left, int right, int borderType)
{
    assert((borderType & ~BORDER_ISOLATED) == BORDER_REFLECT_101);
    const Mat src = src_.getMat();
    dst_.create(src.rows + top + bottom, src.cols + left + right, src.type());
    Mat dst = dst_.getMat();
    auto refl = [](int p, int n) { if (n == 1) return 0; while (p < 0 || p >= n) p = p < 0 ? -p : 2 * (n - 1) - p; return p; };
    for (int r = 0; r < src.rows; ++r) {
        uchar* d = dst.data + (size_t)(r + top) * dst.step;
        const uchar* s = src.data + (size_t)r * src.step;
        if (d + left != s) std::memmove(d + left, s, (size_t)src.cols);
        for (int c = 0; c < left; ++c) d[c] = s[refl(c - left, src.cols)];
        for (int c = 0; c < right; ++c) d[left + src.cols + c] = s[refl(src.cols + c, src.cols)];
    }
    for (int r = 0; r < top; ++r) std::memcpy(dst.data + (size_t)r * dst.step, dst.data + (size_t)(top + refl(r - top, src.rows)) * dst.step, (size_t)dst.cols);
    for (int r = 0; r < bottom; ++r)
        std::memcpy(dst.data + (size_t)(top + src.rows + r) * dst.step, dst.data + (size_t)(top + refl(src.rows + r, src.rows)) * dst.step, (size_t)dst.cols);
}

inline void GaussianBlur(InputArray src_, OutputArray dst_, Size ksize, double sigmaX, double sigmaY = 0, int borderType = BORDER_DEFAULT)
{
    assert(ksize.width == 7 && ksize.height == 7 && sigmaX == 2 && sigmaY == 2 && borderType == BORDER_REFLECT_101);
    const Mat src = src_.getMat();
    Mat tmp = src.clone();                                             // the reference blurs in place
    dst_.create(src.rows, src.cols, src.type());
    Mat dst = dst_.getMat();
    const int rc = orc_gaussian_blur7(tmp.data, tmp.cols, tmp.rows, (int)tmp.step, dst.data, (int)dst.step);
    assert(rc == ORC_OK); (void)rc;
}

// cv::Ptr (OpenCV 3): a shared pointer with implicit construction from a raw pointer and conversions between related types
template <typename T> class Ptr {
public:
    Ptr() {}
    Ptr(std::nullptr_t) {}
    template <typename Y> Ptr(Y* p) : p_(p) {}
    template <typename Y> Ptr(const Ptr<Y>& o) : p_(o.shared()) {}
    void reset() { p_.reset(); }
    template <typename Y> void reset(Y* p) { p_.reset(p); }
    T* operator->() const { return p_.get(); }
    T& operator*() const { return *p_; }
    T* get() const { return p_.get(); }
    bool empty() const { return !p_; }
    explicit operator bool() const { return (bool)p_; }
    const std::shared_ptr<T>& shared() const { return p_; }
private:
    std::shared_ptr<T> p_;
};

class Algorithm { public: virtual ~Algorithm() {} };
class Feature2D : public Algorithm {
public:
    virtual ~Feature2D() {}
    virtual void detect(InputArray, std::vector<KeyPoint>&, InputArray = noArray()) {}
    virtual void compute(InputArray, std::vector<KeyPoint>&, OutputArray) {}
    virtual void detectAndCompute(InputArray, InputArray, std::vector<KeyPoint>&, OutputArray, bool = false) {}
    virtual int defaultNorm() const { return NORM_L2; }
};
typedef Feature2D FeatureDetector;
typedef Feature2D DescriptorExtractor;

// cv::FastFeatureDetector: cv::FAST on the matrix handed in (threshold is an int: a double argument truncates, as in OpenCV)
class FastFeatureDetector : public Feature2D {
public:
    static Ptr<FastFeatureDetector> create(int threshold = 10, bool nonmaxSuppression = true, int = 2)
    {
        FastFeatureDetector* d = new FastFeatureDetector; d->th_ = threshold; d->nms_ = nonmaxSuppression;
        return Ptr<FastFeatureDetector>(d);
    }
    void detect(InputArray image, std::vector<KeyPoint>& keypoints, InputArray = noArray()) override { FAST(image, keypoints, th_, nms_); }
    int getThreshold() const { return th_; }
private:
    int th_ = 10; bool nms_ = true;
};

// detectors / descriptors the reference can be configured with but the hot path never runs: they exist so that the sources compile,
// and stop the process if a test ever reaches them
#define ORBF_SHIM_UNUSED_FEATURE2D(NAME, NORM)                                                                                  \
    class NAME : public Feature2D {                                                                                             \
    public:                                                                                                                     \
        template <typename... A> static Ptr<NAME> create(A...) { return Ptr<NAME>(new NAME); }                                  \
        void detect(InputArray, std::vector<KeyPoint>&, InputArray = noArray()) override { std::abort(); }                      \
        void compute(InputArray, std::vector<KeyPoint>&, OutputArray) override {}                                               \
        int defaultNorm() const override { return NORM; }                                                                       \
    };
ORBF_SHIM_UNUSED_FEATURE2D(ORB, NORM_HAMMING)
ORBF_SHIM_UNUSED_FEATURE2D(BRISK, NORM_HAMMING)
ORBF_SHIM_UNUSED_FEATURE2D(GFTTDetector, NORM_L2)
namespace xfeatures2d {
ORBF_SHIM_UNUSED_FEATURE2D(StarDetector, NORM_L2)
ORBF_SHIM_UNUSED_FEATURE2D(SURF, NORM_L2)
ORBF_SHIM_UNUSED_FEATURE2D(SIFT, NORM_L2)
ORBF_SHIM_UNUSED_FEATURE2D(BriefDescriptorExtractor, NORM_HAMMING)
ORBF_SHIM_UNUSED_FEATURE2D(FREAK, NORM_HAMMING)
ORBF_SHIM_UNUSED_FEATURE2D(LATCH, NORM_HAMMING)
typedef SURF SurfFeatureDetector;
typedef SIFT SiftFeatureDetector;
}  // namespace xfeatures2d

// cv::BFMatcher::knnMatch(query, train, matches, 2), NORM_HAMMING, no cross-check: per query row the two nearest train rows, ties to
// the lower train index — the routine tests/test_oracle_vs_cv2.py pins against cv2.BFMatcher.knnMatch (orc_knn2).  Fewer than two train
// rows give shorter lists, as in OpenCV.
class DescriptorMatcher : public Algorithm {
public:
    virtual void knnMatch(InputArray query, InputArray train, std::vector<std::vector<DMatch>>& matches, int k, InputArray = noArray(), bool = false) const = 0;
};
class BFMatcher : public DescriptorMatcher {
public:
    static Ptr<BFMatcher> create(int normType = NORM_L2, bool crossCheck = false)
    {
        assert(normType == NORM_HAMMING && !crossCheck); (void)normType; (void)crossCheck;
        return Ptr<BFMatcher>(new BFMatcher);
    }
    void knnMatch(InputArray query, InputArray train, std::vector<std::vector<DMatch>>& matches, int k, InputArray = noArray(), bool = false) const override
    {
        const Mat q = query.getMat().clone(), t = train.getMat().clone();                  // tight rows
        assert(k == 2 && (q.empty() || q.cols == 32) && (t.empty() || t.cols == 32)); (void)k;
        matches.assign((size_t)q.rows, std::vector<DMatch>());
        if (q.rows == 0 || t.rows == 0) return;
        std::vector<int> i1((size_t)q.rows), d1((size_t)q.rows), i2((size_t)q.rows), d2((size_t)q.rows);
        orc_knn2(q.data, q.rows, t.data, t.rows, i1.data(), d1.data(), i2.data(), d2.data());
        for (int i = 0; i < q.rows; ++i) {
            if (i1[(size_t)i] >= 0) matches[(size_t)i].push_back(DMatch(i, i1[(size_t)i], 0, (float)d1[(size_t)i]));
            if (i2[(size_t)i] >= 0) matches[(size_t)i].push_back(DMatch(i, i2[(size_t)i], 0, (float)d2[(size_t)i]));
        }
    }
};

// drawing / display entry points of Matcher::Draw* (never on the hot path): no-ops so that the source compiles
struct DrawMatchesFlags { enum { DEFAULT = 0, DRAW_OVER_OUTIMG = 1, NOT_DRAW_SINGLE_POINTS = 2, DRAW_RICH_KEYPOINTS = 4 }; };
enum { COLOR_BGR2GRAY = 6, COLOR_GRAY2BGR = 8 };
template <typename... A> inline void drawMatches(A&&...) {}
template <typename... A> inline void drawKeypoints(A&&...) {}
template <typename... A> inline void imshow(A&&...) {}
// cv::cvtColor(src, dst, COLOR_BGR2GRAY) on 8-bit images: OpenCV's 15-bit fixed point (pinned against cv2 in tests/test_ingest.py);
// COLOR_GRAY2BGR (Matcher::DrawInlierPoints only) is a no-op here
inline void cvtColor(InputArray src_, OutputArray dst_, int code)
{
    if (code != COLOR_BGR2GRAY) return;
    const Mat src = src_.getMat();
    assert(src.type() == CV_8UC3);
    Mat out(src.rows, src.cols, CV_8UC1);
    for (int r = 0; r < src.rows; ++r) {
        const uchar* s = src.ptr(r); uchar* d = out.ptr(r);
        for (int c = 0; c < src.cols; ++c) d[c] = (uchar)((s[3 * c] * 3735u + s[3 * c + 1] * 19235u + s[3 * c + 2] * 9798u + 16384u) >> 15);
    }
    dst_.create(out.rows, out.cols, CV_8UC1);
    out.copyTo(dst_);
}
// cv::undistortPoints(src, dst, K, dist, Mat(), K) on N x 1 CV_32FC2 points: the routine tests/test_undistort.py pins against cv2
inline void undistortPoints(InputArray src_, OutputArray dst_, InputArray K_, InputArray dist_, InputArray = noArray(), InputArray P_ = noArray())
{
    const Mat src = src_.getMat(), K = K_.getMat(), dist = dist_.getMat(), P = P_.getMat();
    assert(src.type() == CV_32FC2 && src.cols == 1 && K.type() == CV_32F && dist.type() == CV_32F && !P.empty());
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) assert(P.at<float>(i, j) == K.at<float>(i, j));
    float d5[5] = { 0, 0, 0, 0, 0 };
    for (int i = 0; i < std::min(dist.rows * dist.cols, 5); ++i) d5[i] = dist.at<float>(i);
    std::vector<float> out((size_t)src.rows * 2 + 2);
    const Mat tight = src.clone();
    const int rc = orc_undistort_points(reinterpret_cast<const float*>(tight.data), src.rows, K.at<float>(0, 0), K.at<float>(1, 1), K.at<float>(0, 2), K.at<float>(1, 2), d5, out.data());
    assert(rc == ORC_OK); (void)rc;
    dst_.create(src.rows, 1, CV_32FC2);
    Mat dst = dst_.getMat();
    for (int i = 0; i < src.rows; ++i) std::memcpy(dst.data + (size_t)i * dst.step, &out[2 * (size_t)i], 8);
}
template <typename... A> inline void rectangle(A&&...) {}
template <typename... A> inline void circle(A&&...) {}
inline int waitKey(int = 0) { return -1; }

// cv::KeyPointsFilter::retainBest (features2d/src/keypoint.cpp): nth_element on the response, then every keypoint whose response
// equals the one at the cut is kept as well
struct KeyPointsFilter {
    static void retainBest(std::vector<KeyPoint>& keypoints, int n_points)
    {
        if (n_points >= 0 && keypoints.size() > (size_t)n_points) {
            if (n_points == 0) { keypoints.clear(); return; }
            std::nth_element(keypoints.begin(), keypoints.begin() + n_points - 1, keypoints.end(),
                [](const KeyPoint& a, const KeyPoint& b) { return a.response > b.response; });
            const float ambiguous = keypoints[(size_t)n_points - 1].response;
            auto new_end = std::partition(keypoints.begin() + n_points, keypoints.end(), [ambiguous](const KeyPoint& k) { return k.response >= ambiguous; });
            keypoints.resize((size_t)(new_end - keypoints.begin()));
        }
    }
};

}  // namespace cv
