// tests/cpp/stub/opencv2/core.hpp — NOT OpenCV: a stand-in with the handful of cv:: declarations include/orbfront_host.hpp touches when
// ORBF_WITH_OPENCV is defined (this image has no OpenCV headers).  It exists so that the reference-signature overloads
// (ORBextractor::operator()(cv::InputArray, cv::InputArray, std::vector<cv::KeyPoint>&, cv::OutputArray), Extractor::Extract,
// Matcher::DescriptorDistance(const cv::Mat&, const cv::Mat&)) are compiled and run by the test suite; layouts follow OpenCV 4
// (cv::KeyPoint 28 bytes, cv::DMatch 16 bytes, cv::Point3f 12 bytes).
#pragma once
#include <cstddef>
#include <cstdint>
#include <memory>
#include <vector>

#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5

namespace cv {

template <typename T> struct Point_ { T x, y; Point_(T a = 0, T b = 0) : x(a), y(b) {} };
using Point2f = Point_<float>;
template <typename T> struct Point3_ { T x, y, z; Point3_(T a = 0, T b = 0, T c = 0) : x(a), y(b), z(c) {} };
using Point3f = Point3_<float>;

struct KeyPoint {
    Point2f pt; float size = 0, angle = -1, response = 0; int octave = 0, class_id = -1;
};
struct DMatch {
    int queryIdx = -1, trainIdx = -1, imgIdx = -1; float distance = 3.4e38f;
    bool operator<(const DMatch& m) const { return distance < m.distance; }
};

class Mat {
public:
    int rows = 0, cols = 0; uint8_t* data = nullptr; size_t step = 0;
    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(int r, int c, int type, void* ext, size_t stepBytes = 0) : rows(r), cols(c), data(static_cast<uint8_t*>(ext)), step(stepBytes ? stepBytes : (size_t)c), type_(type) {}
    void create(int r, int c, int type)
    {
        rows = r; cols = c; type_ = type; step = (size_t)c * (type == CV_32F ? 4 : 1);
        store = std::make_shared<std::vector<uint8_t>>((size_t)r * step); data = store->data();
    }
    void release() { rows = cols = 0; step = 0; data = nullptr; store.reset(); }
    bool empty() const { return rows == 0 || cols == 0 || !data; }
    int type() const { return type_; }
    uint8_t* ptr(int r = 0) { return data + (size_t)r * step; }
    const uint8_t* ptr(int r = 0) const { return data + (size_t)r * step; }
private:
    int type_ = CV_8U; std::shared_ptr<std::vector<uint8_t>> store;
};

class _InputArray {
public:
    _InputArray() {}
    _InputArray(const Mat& m) : m_(const_cast<Mat*>(&m)) {}
    Mat getMat() const { return m_ ? *m_ : Mat(); }
    bool empty() const { return !m_ || m_->empty(); }
protected:
    Mat* m_ = nullptr;
};
class _OutputArray : public _InputArray {
public:
    _OutputArray() {}
    _OutputArray(Mat& m) { m_ = &m; }
    void create(int r, int c, int type) const { if (m_) m_->create(r, c, type); }
    void release() const { if (m_) m_->release(); }
};
using InputArray = const _InputArray&;
using OutputArray = const _OutputArray&;
inline const _OutputArray& noArray() { static _OutputArray none; return none; }

}  // namespace cv
