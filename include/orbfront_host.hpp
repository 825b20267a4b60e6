// include/orbfront_host.hpp — C++ host mirror of the reference's operator interface for the hot path, header-only,
// on top of the C ABI in orbfront.h (liborbfront_b200.so).  Same class names, method names, argument meaning and
// error behaviour as the reference (INTEGRATION.md lists, per reference file, the lines a maintainer changes to bind it):
//
//   ORBextractor   Features/orbextractor.h:24-84     ctor, operator(), detect/compute/detectAndCompute, Get* getters, mvImagePyramid
//   Extractor      Features/extractor.h:6-48         Extract(), mNorm (ORB_SLAM2 route and the adaptive-FAST detector; the OpenCV-contrib
//                                                    descriptor routes are out of scope)
//   Frame          Core/frame.h:60-130, frame.cpp:135-170,276-343   mvKeys / mvKeysUn / mvKeys3Dc / mvuRight / mDescriptors / N / mvbOutlier,
//                                                    ExtractFeatures(), pose and outlier accessors, image bounds
//   KeyFrame       Core/keyframe.h:11, keyframe.cpp:30-52   the feature storage a KeyFrame copies out of its Frame
//   Matcher        Features/matcher.h:10-18          Matcher(nnratio), KnnMatch(Frame&, Frame&, .), KnnMatch(KeyFrame*, Frame&, .), DescriptorDistance
//   Ransac         Odometry/ransac.h:13-69           ctors, setters, Iterate(F1, F2, m12), Iterate(), rmse / mvInliers / mT12 / mpSourceCloud / mpTargetCloud
//   Kabsch         Odometry/kabsch.h:6-15            Compute(setA, setB)
//   Odometry       Odometry/odometry.h:24            void Compute(F1, F2, matches), RANSAC strategy: pose composition + inlier flags
//
// Types: with OpenCV present (define ORBF_WITH_OPENCV) cv::KeyPoint / cv::DMatch / cv::Point3f are used directly (orbf_keypoint and
// orbf_dmatch have their exact layout) and ORBextractor::operator() / Extractor::Extract / Matcher::DescriptorDistance gain the
// reference's cv::InputArray / cv::OutputArray / cv::Mat signatures (compile-checked against tests/cpp/stub/opencv2/core.hpp); without it
// the PODs below stand in.  With Eigen present (define ORBF_WITH_EIGEN) Kabsch::Compute takes Eigen::MatrixXf and returns
// Eigen::Matrix4f (compile-checked against tests/cpp/stub/Eigen/Core); otherwise Matrix4f is a 16-float row-major POD with operator()(r, c) and MatrixXf a minimal column-major stand-in.
// PCL is never required: PointXYZ / PointCloud below have pcl::PointXYZ's 16-byte record and a `points` vector.
// Every object shares one process-wide device context (Runtime), created on first use: Matcher is constructed per call on
// the reference's stack (System/tracking.cpp:197), so its constructor allocates nothing.
// No CPU fallback: without a CUDA device the first call throws orbf::Error(ORBF_ERR_CUDA).  All arithmetic of the path (extraction,
// undistortion, unprojection, matching, RANSAC, Kabsch, pose composition) runs on the device; what stays on the host is what the
// reference keeps in pointer graphs (Landmark* bookkeeping) and the copying of results into the caller's containers.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "orbfront.h"

#ifdef ORBF_WITH_OPENCV
#include <opencv2/core.hpp>
#endif
#ifdef ORBF_WITH_EIGEN
#include <Eigen/Core>
#endif

namespace orbf {

struct Error : std::runtime_error {
    int status;
    Error(int s, const std::string& what) : std::runtime_error(what + ": " + orbf_status_string(s)), status(s) {}
};

#ifdef ORBF_WITH_OPENCV
using KeyPoint = cv::KeyPoint;
using DMatch = cv::DMatch;
using Point3f = cv::Point3f;
#else
struct Point2f { float x, y; };
struct Point3f { float x, y, z; Point3f(float a = 0, float b = 0, float c = 0) : x(a), y(b), z(c) {} };
struct KeyPoint { Point2f pt; float size, angle, response; int octave, class_id; };             // cv::KeyPoint, 28 bytes
struct DMatch {                                                                              // cv::DMatch, 16 bytes
    int queryIdx, trainIdx, imgIdx; float distance;
    bool operator<(const DMatch& m) const { return distance < m.distance; }
};
#endif
static_assert(sizeof(KeyPoint) == sizeof(orbf_keypoint) && sizeof(DMatch) == sizeof(orbf_dmatch), "POD mirrors must match the C ABI");

// Row-major image / descriptor matrix view-or-owner (the subset of cv::Mat the path touches).
template <typename T>
struct Mat_ {
    int rows = 0, cols = 0; size_t step = 0;          // step in elements
    T* data = nullptr; std::vector<T> owned;
    Mat_() {}
    Mat_(const Mat_& o) { *this = o; }
    Mat_& operator=(const Mat_& o)
    {
        rows = o.rows; cols = o.cols; step = o.step; owned = o.owned;
        data = o.owned.empty() ? o.data : owned.data();      // owners deep-copy, views stay views
        return *this;
    }
    Mat_(int r, int c) { create(r, c); }
    Mat_(int r, int c, T* external, size_t stepElems = 0) : rows(r), cols(c), step(stepElems ? stepElems : (size_t)c), data(external) {}
    void create(int r, int c) { rows = r; cols = c; step = (size_t)c; owned.assign((size_t)r * c, T()); data = owned.data(); }
    void release() { rows = cols = 0; step = 0; owned.clear(); data = nullptr; }
    bool empty() const { return rows == 0 || cols == 0 || !data; }
    T* ptr(int r) { return data + (size_t)r * step; }
    const T* ptr(int r) const { return data + (size_t)r * step; }
};
using Mat8u = Mat_<uint8_t>;
using Mat16u = Mat_<uint16_t>;

struct Matrix4f {                                      // row-major; Eigen::Map<Eigen::Matrix<float,4,4,Eigen::RowMajor>>(m) on the Eigen side
    float m[16];
    Matrix4f() { setIdentity(); }
    void setIdentity() { for (int i = 0; i < 16; ++i) m[i] = (i % 5 == 0) ? 1.f : 0.f; }
    float& operator()(int r, int c) { return m[4 * r + c]; }
    float operator()(int r, int c) const { return m[4 * r + c]; }
};

#ifdef ORBF_WITH_EIGEN
using MatrixXf = Eigen::MatrixXf;
#else
struct MatrixXf {                                      // the subset of Eigen::MatrixXf Kabsch::Compute reads: column-major, rows = points
    MatrixXf() {}
    MatrixXf(int r, int c) : r_(r), c_(c), d((size_t)r * c, 0.f) {}
    int rows() const { return r_; }
    int cols() const { return c_; }
    float& operator()(int r, int c) { return d[(size_t)c * r_ + r]; }
    float operator()(int r, int c) const { return d[(size_t)c * r_ + r]; }
private:
    int r_ = 0, c_ = 0; std::vector<float> d;
};
#endif

struct PointXYZ {                                      // pcl::PointXYZ: float data[4] = { x, y, z, 1.0f }, 16 bytes
    float x, y, z, w;
    PointXYZ(float a = 0, float b = 0, float c = 0) : x(a), y(b), z(c), w(1.f) {}
};
static_assert(sizeof(PointXYZ) == 16, "pcl::PointXYZ record");
struct PointCloud { std::vector<PointXYZ> points; size_t size() const { return points.size(); } };     // pcl::PointCloud<pcl::PointXYZ>::points

// Utils/common.h:35-44,67,71: the FR1 intrinsics the reference compiles in.  Its distortion coefficients (k1 = 0.262383, ...) are NOT the
// default here: the synthetic benchmarks are undistorted, so k1 = 0 keeps mvKeysUn = mvKeys (frame.cpp:288-291);
// Runtime::SetCalibration(Calibration::FR1Distorted()) switches the device path to cv::undistortPoints' arithmetic.
struct Calibration {
    float fx = 517.3f, fy = 516.5f, cx = 318.6f, cy = 255.3f, mbf = 40.0f, depthFactor = 1.0f / 5000.0f;
    float k1 = 0.f, k2 = 0.f, p1 = 0.f, p2 = 0.f, k3 = 0.f;
    static Calibration FR1Distorted() { Calibration c; c.k1 = 0.262383f; c.k2 = -0.953104f; c.p1 = -0.005358f; c.p2 = 0.002628f; c.k3 = 1.163314f; return c; }
};

// ---- process-wide device context ---------------------------------------------------------------------------------------
class Runtime {
public:
    // (Re)creates the shared context when the extractor parameters or frame size change; cheap otherwise.
    static orbf_context* Get(int width, int height, int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh)
    {
        Runtime& r = inst();
        std::lock_guard<std::mutex> g(r.mu);
        orbf_config c;
        orbf_default_config(&c);
        c.width = width; c.height = height; c.nfeatures = nfeatures; c.scale_factor = scaleFactor; c.nlevels = nlevels;
        c.ini_th_fast = iniTh; c.min_th_fast = minTh; c.max_frames = 2; c.max_pairs = 2; c.device = r.device;
        c.fx = r.cal.fx; c.fy = r.cal.fy; c.cx = r.cal.cx; c.cy = r.cal.cy; c.mbf = r.cal.mbf; c.depth_factor = r.cal.depthFactor;
        c.k1 = r.cal.k1; c.k2 = r.cal.k2; c.p1 = r.cal.p1; c.p2 = r.cal.p2; c.k3 = r.cal.k3;
        if (r.ctx && std::memcmp(&c, &r.cfg, sizeof(c)) == 0) return r.ctx;
        if (r.ctx) { orbf_destroy(r.ctx); r.ctx = nullptr; }
        orbf_context* h = nullptr;
        const int rc = orbf_create(&c, &h);
        if (rc != ORBF_OK) {
            const std::string detail = h ? orbf_last_error(h) : "";
            if (h) orbf_destroy(h);
            throw Error(rc, "orbf_create " + detail);
        }
        r.ctx = h; r.cfg = c;
        return h;
    }
    // The context of the last extractor configuration (Matcher / Ransac / Kabsch do not depend on the image geometry).
    static orbf_context* Current()
    {
        Runtime& r = inst();
        { std::lock_guard<std::mutex> g(r.mu); if (r.ctx) return r.ctx; }
        return Get(640, 480, 1000, 1.2f, 8, 20, 7);      // Utils/common.h:77, Features/extractor.cpp:86
    }
    static void SetDevice(int dev) { inst().device = dev; }
    // Calibration:: of Utils/common.h as the device path uses it (takes effect with the next extractor call: the context is re-created)
    static void SetCalibration(const Calibration& k) { Runtime& r = inst(); std::lock_guard<std::mutex> g(r.mu); r.cal = k; }
    static Calibration GetCalibration() { Runtime& r = inst(); std::lock_guard<std::mutex> g(r.mu); return r.cal; }
    static int Width() { return inst().cfg.width; }
    static int Height() { return inst().cfg.height; }
    static void Shutdown() { Runtime& r = inst(); std::lock_guard<std::mutex> g(r.mu); if (r.ctx) { orbf_destroy(r.ctx); r.ctx = nullptr; } }
    static std::mutex& Lock() { return inst().call; }    // serialises calls: a context is thread-safe per handle, not per call
private:
    static Runtime& inst() { static Runtime r; return r; }
    ~Runtime() { if (ctx) orbf_destroy(ctx); }
    orbf_context* ctx = nullptr; orbf_config cfg; int device = 0; std::mutex mu, call; Calibration cal;
};

inline void check(int rc, const char* what) { if (rc != ORBF_OK) throw Error(rc, what); }

// ---- ORBextractor ------------------------------------------------------------------------------------------------------
class ORBextractor {
public:
    enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };

    ORBextractor(int nfeatures_, float scaleFactor_, int nlevels_, int iniThFAST_, int minThFAST_)
        : nfeatures(nfeatures_), scaleFactor(scaleFactor_), nlevels(nlevels_), iniThFAST(iniThFAST_), minThFAST(minThFAST_)
    {
        // scale tables as in orbextractor.cpp:346-381 (float * double(scaleFactor) -> float); device tables are identical
        mvScaleFactor.resize(nlevels); mvLevelSigma2.resize(nlevels); mvInvScaleFactor.resize(nlevels); mvInvLevelSigma2.resize(nlevels);
        mvScaleFactor[0] = 1.0f; mvLevelSigma2[0] = 1.0f;
        for (int i = 1; i < nlevels; i++) {
            mvScaleFactor[i] = (float)((double)mvScaleFactor[i - 1] * (double)scaleFactor_);
            mvLevelSigma2[i] = mvScaleFactor[i] * mvScaleFactor[i];
        }
        for (int i = 0; i < nlevels; i++) { mvInvScaleFactor[i] = 1.0f / mvScaleFactor[i]; mvInvLevelSigma2[i] = 1.0f / mvLevelSigma2[i]; }
    }

    // Compute the ORB features and descriptors on an image.  Mask is ignored, as in the reference (orbextractor.h:36).
    void operator()(const Mat8u& image, const Mat8u& /*mask*/, std::vector<KeyPoint>& keypoints, Mat8u& descriptors)
    {
        if (image.empty()) return;                                           // orbextractor.cpp:758-759: outputs untouched
        std::lock_guard<std::mutex> g(Runtime::Lock());
        orbf_context* ctx = Runtime::Get(image.cols, image.rows, nfeatures, (float)scaleFactor, nlevels, iniThFAST, minThFAST);
        int cap = 0, n = 0;
        check(orbf_keypoint_capacity(ctx, &cap), "orbf_keypoint_capacity");
        keypoints.resize((size_t)cap);
        std::vector<uint8_t> desc((size_t)cap * 32);
        check(orbf_extract(ctx, image.data, image.cols, image.rows, (int)image.step, reinterpret_cast<orbf_keypoint*>(keypoints.data()),
                  desc.data(), cap, &n), "orbf_extract");
        keypoints.resize((size_t)n);
        if (n == 0) { descriptors.release(); return; }                       // orbextractor.cpp:776-777
        descriptors.create(n, 32);
        std::memcpy(descriptors.data, desc.data(), (size_t)n * 32);
        lastW = image.cols; lastH = image.rows;
    }
#ifdef ORBF_WITH_OPENCV
    // The reference's own signatures (orbextractor.h:36-49): 8-bit single-channel image in, descriptors created in the callee as
    // n x 32 CV_8U (orbextractor.cpp:779) or released when nothing was found (:776-777).
    void operator()(cv::InputArray _image, cv::InputArray /*_mask*/, std::vector<cv::KeyPoint>& _keypoints, cv::OutputArray _descriptors)
    {
        if (_image.empty()) return;
        cv::Mat image = _image.getMat();
        if (image.type() != CV_8UC1) throw std::invalid_argument("orbf::ORBextractor: image must be CV_8UC1");      // assert at orbextractor.cpp:762
        Mat8u d;
        (*this)(Mat8u(image.rows, image.cols, image.data, (size_t)image.step), Mat8u(), _keypoints, d);
        if (d.empty()) { _descriptors.release(); return; }
        _descriptors.create(d.rows, 32, CV_8U);
        cv::Mat out = _descriptors.getMat();
        for (int r = 0; r < d.rows; ++r) std::memcpy(out.ptr(r), d.ptr(r), 32);
    }
    void detectAndCompute(cv::InputArray image, cv::InputArray mask, std::vector<cv::KeyPoint>& keypoints, cv::OutputArray descriptors,
        bool /*useProvidedKeypoints*/ = false) { (*this)(image, mask, keypoints, descriptors); }
    void detect(cv::InputArray image, std::vector<cv::KeyPoint>& keypoints, cv::InputArray mask = cv::noArray())
    {
        cv::Mat d; (*this)(image, mask, keypoints, d);
    }
    void compute(cv::InputArray image, std::vector<cv::KeyPoint>& keypoints, cv::OutputArray descriptors) { (*this)(image, cv::noArray(), keypoints, descriptors); }
#endif
    void detectAndCompute(const Mat8u& image, const Mat8u& mask, std::vector<KeyPoint>& keypoints, Mat8u& descriptors,
        bool /*useProvidedKeypoints*/ = false) { (*this)(image, mask, keypoints, descriptors); }      // orbextractor.cpp:828-831
    void detect(const Mat8u& image, std::vector<KeyPoint>& keypoints, const Mat8u& mask = Mat8u()) { Mat8u d; (*this)(image, mask, keypoints, d); }
    void compute(const Mat8u& image, std::vector<KeyPoint>& keypoints, Mat8u& descriptors) { (*this)(image, Mat8u(), keypoints, descriptors); }

    int GetLevels() { return nlevels; }
    float GetScaleFactor() { return (float)scaleFactor; }
    std::vector<float> GetScaleFactors() { return mvScaleFactor; }
    std::vector<float> GetInverseScaleFactors() { return mvInvScaleFactor; }
    std::vector<float> GetScaleSigmaSquares() { return mvLevelSigma2; }
    std::vector<float> GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }

    // mvImagePyramid (public member in the reference, orbextractor.h:57): `extractor.mvImagePyramid[level]` reads like the reference's
    // std::vector<cv::Mat>, but a level crosses the link only when it is asked for (the pyramid lives in HBM)
    struct Pyramid {
        ORBextractor* self;
        size_t size() const { return (size_t)self->nlevels; }
        Mat8u operator[](int level) const { return self->ImagePyramidLevel(level); }
    } mvImagePyramid{this};
    ORBextractor(const ORBextractor&) = delete;
    ORBextractor& operator=(const ORBextractor&) = delete;
    Mat8u ImagePyramidLevel(int level, bool blurred = false)
    {
        std::lock_guard<std::mutex> g(Runtime::Lock());
        orbf_context* ctx = Runtime::Get(lastW, lastH, nfeatures, (float)scaleFactor, nlevels, iniThFAST, minThFAST);
        std::vector<int32_t> w(nlevels), h(nlevels);
        check(orbf_get_tables(ctx, nullptr, nullptr, nullptr, nullptr, nullptr, w.data(), h.data()), "orbf_get_tables");
        Mat8u m(h[level], w[level]);
        check(orbf_pyramid_level(ctx, 0, level, blurred ? 1 : 0, m.data, w[level]), "orbf_pyramid_level");
        return m;
    }

protected:
    int nfeatures; double scaleFactor; int nlevels, iniThFAST, minThFAST;
    std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;
    int lastW = 640, lastH = 480;
};

// ---- Extractor facade ----------------------------------------------------------------------------------------------------
enum { NORM_HAMMING = 6 };                                // cv::NORM_HAMMING
class Extractor {
public:
    enum eAlgorithm { ORB = 0, ORB_SLAM2, FAST, GFTT, STAR, BRISK, FREAK, BRIEF, LATCH, SURF, SIFT };
    enum eMode { NORMAL = 0, ADAPTIVE };
    eAlgorithm mDetectorAlgorithm, mDescriptorAlgorithm; eMode mMode;

    Extractor(const eAlgorithm& detector = ORB_SLAM2, const eAlgorithm& descriptor = ORB_SLAM2, const eMode& mode = NORMAL, int nFeatures = 1000)
        : mDetectorAlgorithm(detector), mDescriptorAlgorithm(descriptor), mMode(mode), mnFeatures(nFeatures)
    {
        // the reference terminates on an unknown enum (extractor.cpp:27,108,132).  On the GPU: the ORB_SLAM2 route, and the
        // adaptive FAST detector of CreateAdaptiveDetector (extractor.cpp:52-77); the OpenCV-contrib routes are out of scope
        if (detector == ORB_SLAM2 && descriptor == ORB_SLAM2) {
            mpDetector.reset(new ORBextractor(nFeatures, 1.2f, 8, 20, 7));   // extractor.cpp:86
            mNorm() = NORM_HAMMING;                                          // defaultNorm() of ORB (extractor.cpp:35)
        } else if (detector == FAST && mode == ADAPTIVE) {
            orbf_default_adaptive_config(&mAdaptive);
            mAdaptive.retain_best = nFeatures;
            mAdaptiveState.assign((size_t)mAdaptive.grid * mAdaptive.grid, 0.0);    // <= 0: start from init_th (20)
        } else throw std::invalid_argument("orbf::Extractor: only the ORB_SLAM2 and adaptive-FAST routes run on the GPU");
    }
    void Extract(const Mat8u& image, const Mat8u& mask, std::vector<KeyPoint>& keypoints, Mat8u& descriptors)
    {
        if (mpDetector) { mpDetector->detectAndCompute(image, mask, keypoints, descriptors); return; }   // extractor.cpp:41-42
        // adaptive route: detect -> retainBest(nFeatures) (extractor.cpp:44-46); the descriptor extractors of this route are
        // OpenCV(-contrib) objects and stay with the caller
        std::lock_guard<std::mutex> g(Runtime::Lock());
        orbf_context* ctx = Runtime::Get(image.cols, image.rows, mnFeatures, 1.2f, 8, 20, 7);
        const int cap = mAdaptive.max_per_cell * mAdaptive.grid * mAdaptive.grid;
        keypoints.resize((size_t)cap);
        int n = 0;
        check(orbf_adaptive_detect(ctx, &mAdaptive, image.data, 1, (int)image.step, (int64_t)image.step * image.rows, mAdaptiveState.data(),
                  reinterpret_cast<orbf_keypoint*>(keypoints.data()), &n, cap, nullptr, nullptr), "orbf_adaptive_detect");
        keypoints.resize((size_t)n);
        descriptors.release();
    }
#ifdef ORBF_WITH_OPENCV
    void Extract(cv::InputArray _image, cv::InputArray /*mask*/, std::vector<cv::KeyPoint>& keypoints, cv::OutputArray descriptors)   // extractor.h:33
    {
        if (mpDetector) { (*mpDetector)(_image, cv::noArray(), keypoints, descriptors); return; }
        cv::Mat image = _image.getMat();
        Mat8u none;
        Extract(Mat8u(image.rows, image.cols, image.data, (size_t)image.step), Mat8u(), keypoints, none);
        descriptors.release();
    }
#endif
    const std::vector<double>& AdaptiveThresholds() const { return mAdaptiveState; }
    bool IsOrbSlam2() const { return (bool)mpDetector; }
    int Features() const { return mnFeatures; }   // DetectorAdjuster::mThresh of the 3x3 cells
    static int& mNorm() { static int n = NORM_HAMMING; return n; }
    ORBextractor* detector() { return mpDetector.get(); }
private:
    std::unique_ptr<ORBextractor> mpDetector;
    int mnFeatures;
    orbf_adaptive_config mAdaptive;
    std::vector<double> mAdaptiveState;
};

// ---- Frame: the storage the path reads and writes ----------------------------------------------------------------------
class Landmark;                                           // Core/landmark.h: pointer-graph state, opaque here (stays with the caller)

class Frame {
public:
    Frame() {}
    Frame(const Mat8u& imGray, const Mat16u& imDepthRaw, double timestamp = 0.0) : mImGray(imGray), mImDepthRaw(imDepthRaw), mTimestamp(timestamp) {}
    virtual ~Frame() {}

    // Frame::ExtractFeatures (frame.cpp:135-170): extraction, UndistortKeyPoints, the depth gather and the unprojection — all on the device.
    // ORB_SLAM2 route: one orbf_extract_batch call with the depth plane (the describe kernel samples it and unprojects), results copied
    // back once.  Adaptive-FAST route: the detector, then orbf_unproject_keypoints on its keypoints.
    void ExtractFeatures(Extractor* pExtractor)
    {
        if (mImGray.empty()) { N = 0; return; }
        const uint16_t* depth = mImDepthRaw.empty() ? nullptr : mImDepthRaw.data;
        if (pExtractor->IsOrbSlam2()) {
            std::lock_guard<std::mutex> g(Runtime::Lock());
            orbf_context* ctx = Runtime::Get(mImGray.cols, mImGray.rows, pExtractor->Features(), 1.2f, 8, 20, 7);        // extractor.cpp:86
            int cap = 0, n = 0, n2 = 0;
            check(orbf_keypoint_capacity(ctx, &cap), "orbf_keypoint_capacity");
            check(orbf_extract_batch(ctx, 0, 1, mImGray.data, (int64_t)mImGray.step, (int64_t)mImGray.step * mImGray.rows, depth,
                      (int64_t)mImDepthRaw.step, (int64_t)mImDepthRaw.step * mImDepthRaw.rows), "orbf_extract_batch");
            mvKeys.resize((size_t)cap); mvKeys3Dc.resize((size_t)cap);
            std::vector<uint8_t> desc((size_t)cap * 32);
            static_assert(sizeof(Point3f) == 12, "Point3f must be three packed floats");
            check(orbf_download_frame(ctx, 0, reinterpret_cast<orbf_keypoint*>(mvKeys.data()), desc.data(), reinterpret_cast<float*>(mvKeys3Dc.data()), cap, &n),
                "orbf_download_frame");
            std::vector<float> xyUn((size_t)cap * 2); mvuRight.resize((size_t)cap);
            check(orbf_download_keys_un(ctx, 0, xyUn.data(), mvuRight.data(), cap, &n2), "orbf_download_keys_un");
            N = (size_t)n;
            mvKeys.resize(N); mvKeys3Dc.resize(N); mvuRight.resize(N);
            if (n == 0) mDescriptors.release();
            else { mDescriptors.create(n, 32); std::memcpy(mDescriptors.data, desc.data(), (size_t)n * 32); }
            mvKeysUn = mvKeys;
            for (size_t i = 0; i < N; ++i) { mvKeysUn[i].pt.x = xyUn[2 * i]; mvKeysUn[i].pt.y = xyUn[2 * i + 1]; }
        } else {
            pExtractor->Extract(mImGray, Mat8u(), mvKeys, mDescriptors);
            N = mvKeys.size();
            mvKeys3Dc.assign(N, Point3f(0, 0, 0)); mvuRight.assign(N, -1.f); mvKeysUn = mvKeys;
            if (N) {
                std::lock_guard<std::mutex> g(Runtime::Lock());
                std::vector<float> xyUn(N * 2);
                check(orbf_unproject_keypoints(Runtime::Current(), reinterpret_cast<const orbf_keypoint*>(mvKeys.data()), (int)N, depth, mImGray.cols, mImGray.rows,
                          (int64_t)mImDepthRaw.step, reinterpret_cast<float*>(mvKeys3Dc.data()), mvuRight.data(), xyUn.data()), "orbf_unproject_keypoints");
                for (size_t i = 0; i < N; ++i) { mvKeysUn[i].pt.x = xyUn[2 * i]; mvKeysUn[i].pt.y = xyUn[2 * i + 1]; }
            }
        }
        mvbOutlier.assign(N, false);                                          // frame.cpp:145
        mvpLandmarks.assign(N, nullptr);
        ComputeImageBounds();
    }

    // Frame::ComputeImageBounds (frame.cpp:315-343): the undistorted image corners (cv::undistortPoints on the device when k1 != 0)
    void ComputeImageBounds()
    {
        const Calibration K = Runtime::GetCalibration();
        const float w = (float)mImGray.cols, h = (float)mImGray.rows;
        if (K.k1 != 0.f) {
            float xy[8] = { 0.f, 0.f, w, 0.f, 0.f, h, w, h };
            const float dist[5] = { K.k1, K.k2, K.p1, K.p2, K.k3 };
            std::lock_guard<std::mutex> g(Runtime::Lock());
            check(orbf_undistort_points(Runtime::Current(), xy, 4, K.fx, K.fy, K.cx, K.cy, dist, xy), "orbf_undistort_points");
            mnMinX = std::min(xy[0], xy[4]); mnMaxX = std::max(xy[2], xy[6]); mnMinY = std::min(xy[1], xy[3]); mnMaxY = std::max(xy[5], xy[7]);
        } else { mnMinX = 0.f; mnMaxX = w; mnMinY = 0.f; mnMaxY = h; }
    }

    // pose and flags as Odometry::Compute and Matcher::KnnMatch touch them (frame.cpp:172-210,276-284)
    virtual void SetPose(const Matrix4f& Tcw) { mTcw = Tcw; mbHasPose = true; }
    virtual Matrix4f GetPose() { return mTcw; }
    void SetOutlier(const size_t& idx) { mvbOutlier[idx] = true; }
    void SetInlier(const size_t& idx) { mvbOutlier[idx] = false; }
    bool IsOutlier(const size_t& idx) { return mvbOutlier[idx] == true; }
    bool IsInlier(const size_t& idx) { return mvbOutlier[idx] == false; }
    std::vector<bool> GetOutliers() { return mvbOutlier; }
    virtual void AddLandmark(Landmark* pLM, const size_t& idx) { mvpLandmarks[idx] = pLM; }
    virtual Landmark* GetLandmark(const size_t& idx) { return mvpLandmarks[idx]; }
    virtual std::vector<Landmark*> GetLandmarks() { return mvpLandmarks; }

    Mat8u mImGray; Mat16u mImDepthRaw; double mTimestamp = 0.0;
    std::vector<KeyPoint> mvKeys, mvKeysUn;
    std::vector<Point3f> mvKeys3Dc;
    std::vector<float> mvuRight;
    Mat8u mDescriptors;
    size_t N = 0;
    float mnMinX = 0.f, mnMinY = 0.f, mnMaxX = 0.f, mnMaxY = 0.f;
protected:
    std::vector<bool> mvbOutlier;
    std::vector<Landmark*> mvpLandmarks;
    Matrix4f mTcw; bool mbHasPose = false;
};

// KeyFrame (Core/keyframe.h:11): a Frame whose feature storage was copied out of the frame it was made from (keyframe.cpp:30-52); the
// covisibility graph / spanning tree / database hooks are control plane and not mirrored.
class KeyFrame : public Frame {
public:
    explicit KeyFrame(Frame& frame) : Frame(frame) {}
};

// ---- Matcher -----------------------------------------------------------------------------------------------------------
class Matcher {
public:
    Matcher(float nnratio = 0.6f) : mfNNratio(nnratio), TH_LOW(50), TH_HIGH(100) {}     // matcher.cpp:10-21; allocates nothing

    static double DescriptorDistance(const Mat8u& a, const Mat8u& b)                    // matcher.cpp:355-358
    {
        int d = 0;
        check(orbf_descriptor_distance(a.data, b.data, a.cols, &d), "orbf_descriptor_distance");
        return (double)d;
    }

#ifdef ORBF_WITH_OPENCV
    static double DescriptorDistance(const cv::Mat& a, const cv::Mat& b)                // matcher.h:18
    {
        int d = 0;
        check(orbf_descriptor_distance(a.data, b.data, a.cols, &d), "orbf_descriptor_distance");
        return (double)d;
    }
#endif

    // Matcher::KnnMatch(KeyFrame*, Frame&, .) (matcher.cpp:23-53): kNN-2 + ratio of the keyframe's descriptors against the frame's on the
    // device; the survivors then pass the reference's landmark conditions in query order, exactly as written there — pKF1 must hold a
    // landmark at queryIdx that `isBad` does not reject, F2's feature must still be free; an accepted match hands the landmark to F2 and
    // marks the feature an outlier until the pose optimisation says otherwise.  Landmark is opaque here: pass `isBad` (Landmark::isBad,
    // landmark.h) — by default no landmark is bad.  Like the reference, vMatches12 is appended to, not cleared.
    template <typename IsBad>
    size_t KnnMatch(KeyFrame* pKF1, Frame& F2, std::vector<DMatch>& vMatches12, IsBad isBad)
    {
        const int nq = pKF1->mDescriptors.rows, nt = F2.mDescriptors.rows;
        if (nq == 0 || nt < 2) return vMatches12.size();
        std::vector<DMatch> all((size_t)nq);
        int n = 0;
        {
            std::lock_guard<std::mutex> g(Runtime::Lock());
            check(orbf_knn_match(Runtime::Current(), pKF1->mDescriptors.data, nq, F2.mDescriptors.data, nt, mfNNratio, 0,
                      reinterpret_cast<orbf_dmatch*>(all.data()), nq, &n), "orbf_knn_match");
        }
        const std::vector<Landmark*> vpLandmarksKF1 = pKF1->GetLandmarks();
        for (int i = 0; i < n; ++i) {
            const size_t i1 = (size_t)all[i].queryIdx, i2 = (size_t)all[i].trainIdx;
            Landmark* pLM = vpLandmarksKF1[i1];
            if (!pLM) continue;
            if (isBad(pLM)) continue;
            if (F2.GetLandmark(i2)) continue;
            F2.AddLandmark(pLM, i2);
            F2.SetOutlier(i2);
            vMatches12.push_back(all[i]);
        }
        return vMatches12.size();
    }
    size_t KnnMatch(KeyFrame* pKF1, Frame& F2, std::vector<DMatch>& vMatches12) { return KnnMatch(pKF1, F2, vMatches12, [](Landmark*) { return false; }); }

    // kNN-2 + Lowe ratio on the frames' descriptors (matcher.cpp:55-66).  The reference then filters on Landmark* state
    // (matcher.cpp:70-83: F1 holds a live landmark at queryIdx, F2's slot is free) — pointer-graph bookkeeping that stays on
    // the host: pass it as `accept`; without it every ratio survivor is returned, in query order.
    template <typename Accept>
    size_t KnnMatch(Frame& pF1, Frame& pF2, std::vector<DMatch>& vMatches12, Accept accept, bool crossCheck = false)
    {
        vMatches12.clear();
        const int nq = pF1.mDescriptors.rows, nt = pF2.mDescriptors.rows;
        if (nq == 0 || nt < 2) return 0;                                     // matchesKnn[i][1] needs two train rows (matcher.cpp:65)
        std::lock_guard<std::mutex> g(Runtime::Lock());
        orbf_context* ctx = Runtime::Current();
        std::vector<DMatch> all((size_t)nq);
        int n = 0;
        check(orbf_knn_match(ctx, pF1.mDescriptors.data, nq, pF2.mDescriptors.data, nt, mfNNratio, crossCheck ? 1 : 0,
                  reinterpret_cast<orbf_dmatch*>(all.data()), nq, &n), "orbf_knn_match");
        for (int i = 0; i < n; ++i) if (accept(all[i])) vMatches12.push_back(all[i]);
        return vMatches12.size();
    }
    size_t KnnMatch(Frame& pF1, Frame& pF2, std::vector<DMatch>& vMatches12, bool crossCheck = false)
    {
        return KnnMatch(pF1, pF2, vMatches12, [](const DMatch&) { return true; }, crossCheck);
    }

    // One projected landmark as Matcher::ProjectionMatch reads it (matcher.cpp:94-103): the Landmark* graph stays with the caller.
    struct ProjectedLandmark {
        const uint8_t* descriptor;          // Landmark::GetDescriptor(), 32 bytes
        float projX, projY;                 // mTrackProjX / mTrackProjY
        bool inView;                        // mbTrackInView && !isBad()
        bool observed;                      // Observations() > 0
    };
    // Matcher::ProjectionMatch (matcher.cpp:90-143).  featureTaken[j] = pFrame->GetLandmark(j) exists and has Observations() > 0.
    // bestIdx[i] = feature the reference would pass to pFrame->AddLandmark(vpLandmarks[i], .) or -1; returns nmatches.
    size_t ProjectionMatch(Frame& frame, const std::vector<ProjectedLandmark>& landmarks, const std::vector<uint8_t>& featureTaken, float th,
        std::vector<int>& bestIdx)
    {
        const int L = (int)landmarks.size(), N = (int)frame.mvKeys.size();
        bestIdx.assign((size_t)L, -1);
        if (L == 0) return 0;
        std::vector<uint8_t> d((size_t)L * 32), flags((size_t)L);
        std::vector<float> px((size_t)L), py((size_t)L), kx((size_t)N), ky((size_t)N);
        std::vector<int32_t> oct((size_t)N);
        for (int i = 0; i < L; ++i) {
            std::copy(landmarks[i].descriptor, landmarks[i].descriptor + 32, d.begin() + (size_t)i * 32);
            px[i] = landmarks[i].projX; py[i] = landmarks[i].projY;
            flags[i] = (uint8_t)((landmarks[i].inView ? 1 : 0) | (landmarks[i].observed ? 2 : 0));
        }
        for (int j = 0; j < N; ++j) { kx[j] = frame.mvKeys[j].pt.x; ky[j] = frame.mvKeys[j].pt.y; oct[j] = frame.mvKeys[j].octave; }
        int n = 0;
        std::lock_guard<std::mutex> g(Runtime::Lock());
        check(orbf_projection_match(Runtime::Current(), -1, kx.data(), ky.data(), oct.data(), frame.mDescriptors.data, N, d.data(), px.data(), py.data(),
                  flags.data(), L, featureTaken.empty() ? nullptr : featureTaken.data(), th, mfNNratio, (int)TH_HIGH, bestIdx.data(), &n),
            "orbf_projection_match");
        return (size_t)n;
    }

    // Matcher::Fuse (matcher.cpp:212-296), the computing half: projection of every landmark with the keyframe pose and the windowed
    // search for its best feature.  lmValid[i] = pLM && !pLM->isBad() && !pLM->IsInKeyFrame(pKF); bestIdx[i] = feature or -1.  The
    // caller then runs the reference's Replace / AddObservation / AddLandmark branch (matcher.cpp:297-311) on the survivors.
    void FuseSearch(Frame& kf, const float Rcw[9], const float tcw[3], const float camera[9] /* fx fy cx cy mbf minX maxX minY maxY */,
        const std::vector<Point3f>& worldPos, const std::vector<const uint8_t*>& descriptors, const std::vector<uint8_t>& lmValid, float th,
        std::vector<int>& bestIdx)
    {
        const int L = (int)worldPos.size(), N = (int)kf.mvKeysUn.size();
        bestIdx.assign((size_t)L, -1);
        if (L == 0) return;
        std::vector<uint8_t> d((size_t)L * 32);
        std::vector<float> pos((size_t)L * 3), kx((size_t)N), ky((size_t)N);
        for (int i = 0; i < L; ++i) {
            std::copy(descriptors[i], descriptors[i] + 32, d.begin() + (size_t)i * 32);
            pos[3 * i] = worldPos[i].x; pos[3 * i + 1] = worldPos[i].y; pos[3 * i + 2] = worldPos[i].z;
        }
        for (int j = 0; j < N; ++j) { kx[j] = kf.mvKeysUn[j].pt.x; ky[j] = kf.mvKeysUn[j].pt.y; }
        std::lock_guard<std::mutex> g(Runtime::Lock());
        check(orbf_fuse_search(Runtime::Current(), -1, Rcw, tcw, camera, kx.data(), ky.data(), kf.mvuRight.data(), kf.mDescriptors.data, N, pos.data(), d.data(),
                  lmValid.data(), L, th, (int)TH_LOW, bestIdx.data(), nullptr), "orbf_fuse_search");
    }

    // Matcher::BoWMatch (matcher.cpp:145-209).  A DBoW3::FeatureVector is a std::map<NodeId, std::vector<unsigned>>: pass it flattened
    // (words ascending, bucket offsets, feature indices in bucket order).
    struct FlatFeatureVector { std::vector<int32_t> words, offsets, indices; };
    int BoWMatch(const FlatFeatureVector& fv1, const Mat8u& desc1, const FlatFeatureVector& fv2, const Mat8u& desc2, std::vector<DMatch>& vMatches12)
    {
        vMatches12.assign(fv1.indices.size(), DMatch());
        int n = 0;
        std::lock_guard<std::mutex> g(Runtime::Lock());
        check(orbf_bow_match(Runtime::Current(), fv1.words.data(), fv1.offsets.data(), fv1.indices.data(), (int)fv1.words.size(), desc1.data, desc1.rows,
                  fv2.words.data(), fv2.offsets.data(), fv2.indices.data(), (int)fv2.words.size(), desc2.data, desc2.rows, mfNNratio, (int)TH_LOW,
                  reinterpret_cast<orbf_dmatch*>(vMatches12.data()), (int)vMatches12.size(), &n), "orbf_bow_match");
        vMatches12.resize((size_t)n);
        return n;
    }

private:
    float mfNNratio; double TH_LOW, TH_HIGH;
};

// ---- Ransac ------------------------------------------------------------------------------------------------------------
class Ransac {
public:
    Ransac() : Ransac(200, 20, 3.0f, 4) {}                                               // ransac.cpp:8-17
    Ransac(int iters, unsigned minInlierTh, float maxMahalanobisDist, unsigned sampleSize)
        : mpSourceCloud(std::make_shared<PointCloud>()), mpTargetCloud(std::make_shared<PointCloud>())
    {
        SetParameters(iters, minInlierTh, maxMahalanobisDist, sampleSize);
    }
    // ransac.cpp:26-36: frames and matches bound at construction, solved by Iterate().  (The reference leaves the four parameters
    // uninitialised on this path; here they take the defaults of Ransac().)
    Ransac(KeyFrame* pKF1, KeyFrame* pKF2, const std::vector<DMatch>& vMatches12) : Ransac()
    {
        mpSourceFrame = pKF1; mpTargetFrame = pKF2; mvMatchesS2T = vMatches12;
    }
    void SetParameters(int iters, unsigned minInlierTh, float maxMahalanobisDist, unsigned sampleSize)
    {
        mIterations = iters; mMinInlierTh = minInlierTh; mMaxMahalanobisDistance = maxMahalanobisDist; mSampleSize = sampleSize;
    }
    void SetIterations(int iters) { mIterations = iters; }
    void SetMaxMahalanobisDistance(float dist) { mMaxMahalanobisDistance = dist; }
    void SetSampleSize(unsigned sampleSize) { mSampleSize = sampleSize; }
    void SetInlierThreshold(unsigned th) { mMinInlierTh = th; }
    void CheckDepth(bool check_) { mCheckDepth = check_; }
    // The reference draws from libc rand() seeded with srand(clock()) (main.cpp:27, quirk Q5): the seed is explicit here and
    // advances by one per call, so a run is reproducible.
    static unsigned& Seed() { static unsigned s = 42; return s; }
    // DepthCovariance's static local (ransac.cpp:416-421, quirk Q7): latched by the first scored pair of the process.
    static double& LatchedDepthCovariance() { static double c = -1.0; return c; }

    bool Iterate(Frame* pF1, Frame* pF2, const std::vector<DMatch>& m12)                  // ransac.cpp:155-267
    {
        mpSourceFrame = pF1; mpTargetFrame = pF2;
        return Solve(m12);
    }
    bool Iterate() { return Solve(mvMatchesS2T); }                                        // ransac.cpp:44-153: the same loop on the bound frames
    bool SolvedOnDevice() const { return mbSolved; }                                      // false: the early return at ransac.cpp:166-167

    float rmse = 1e6f;
    std::vector<DMatch> mvInliers;
    Matrix4f mT12;
    // ransac.cpp:163-189: the 3D points of the depth-valid matches in m12 order, for the GICP refinement that may follow; gathered on
    // the device (orbf_ransac_clouds keeps them there for a device-side consumer) and copied here
    std::shared_ptr<PointCloud> mpSourceCloud, mpTargetCloud;

private:
    bool Solve(const std::vector<DMatch>& m12) { return Solve(m12, nullptr, nullptr); }
    // pose1 / pose2 (Odometry::Compute): the composition rule T12 * pose1 rides on the same call and synchronisation
    bool Solve(const std::vector<DMatch>& m12, const Matrix4f* pose1, Matrix4f* pose2)
    {
        rmse = 1e6f; mvInliers.clear(); mT12.setIdentity(); mbSolved = false;
        mpSourceCloud->points.clear(); mpTargetCloud->points.clear();
        if (m12.size() < mMinInlierTh) return false;                                     // ransac.cpp:166-167: nothing touched
        std::lock_guard<std::mutex> g(Runtime::Lock());
        orbf_context* ctx = Runtime::Current();
        orbf_ransac_config cfg;
        orbf_default_ransac_config(&cfg);
        cfg.iterations = mIterations; cfg.min_inlier_th = mMinInlierTh; cfg.max_mahal = mMaxMahalanobisDistance; cfg.sample_size = mSampleSize;
        cfg.check_depth = mCheckDepth ? 1 : 0; cfg.depth_cov = LatchedDepthCovariance(); cfg.seed = Seed()++;
        orbf_ransac_result res;
        std::vector<DMatch> inl(m12.size() ? m12.size() : 1);
        static_assert(sizeof(Point3f) == 12, "Point3f must be three packed floats");
        const Frame *pF1 = mpSourceFrame, *pF2 = mpTargetFrame;
        int nc = 0;
        mpSourceCloud->points.resize(m12.size()); mpTargetCloud->points.resize(m12.size());
        // Iterate, the clouds it leaves for GICP (ransac.cpp:163-189) and, for Odometry::Compute, the composed pose: one call, one synchronisation
        check(orbf_odometry_compute(ctx, &cfg, reinterpret_cast<const float*>(pF1->mvKeys3Dc.data()), (int)pF1->mvKeys3Dc.size(),
                  reinterpret_cast<const float*>(pF2->mvKeys3Dc.data()), (int)pF2->mvKeys3Dc.size(),
                  reinterpret_cast<const orbf_dmatch*>(m12.data()), (int)m12.size(), reinterpret_cast<orbf_dmatch*>(inl.data()), (int)inl.size(), &res,
                  reinterpret_cast<float*>(mpSourceCloud->points.data()), reinterpret_cast<float*>(mpTargetCloud->points.data()), (int)m12.size(), &nc,
                  pose1 ? pose1->m : nullptr, pose2 ? pose2->m : nullptr), "orbf_odometry_compute");
        if (LatchedDepthCovariance() < 0.0 && res.depth_cov_used >= 0.0) LatchedDepthCovariance() = res.depth_cov_used;
        rmse = res.rmse;
        std::memcpy(mT12.m, res.T12, sizeof(res.T12));
        mvInliers.assign(inl.begin(), inl.begin() + res.n_inliers);
        mpSourceCloud->points.resize((size_t)nc); mpTargetCloud->points.resize((size_t)nc);
        mbSolved = true;
        return res.ok != 0;
    }
    friend class Odometry;
    int mIterations = 200; unsigned mMinInlierTh = 20; float mMaxMahalanobisDistance = 3.0f; unsigned mSampleSize = 4; bool mCheckDepth = true;
    Frame* mpSourceFrame = nullptr; Frame* mpTargetFrame = nullptr;
    std::vector<DMatch> mvMatchesS2T;
    bool mbSolved = false;
};

// ---- Kabsch ------------------------------------------------------------------------------------------------------------
class Kabsch {
public:
    // setA / setB: N x 3, rows = points (kabsch.cpp:14-57)
    Matrix4f Compute(const std::vector<Point3f>& setA, const std::vector<Point3f>& setB)
    {
        std::lock_guard<std::mutex> g(Runtime::Lock());
        Matrix4f T;
        const int n = (int)std::min(setA.size(), setB.size());
        check(orbf_kabsch(Runtime::Current(), reinterpret_cast<const float*>(setA.data()), reinterpret_cast<const float*>(setB.data()), n, T.m),
            "orbf_kabsch");
        return T;
    }
    // the reference's signature (kabsch.h:10): Eigen::MatrixXf in (or the stand-in above), 4x4 out
#ifdef ORBF_WITH_EIGEN
    Eigen::Matrix4f Compute(const Eigen::MatrixXf& setA, const Eigen::MatrixXf& setB)
    {
        const Matrix4f T = Compute(rowsOf(setA), rowsOf(setB));
        return Eigen::Map<const Eigen::Matrix<float, 4, 4, Eigen::RowMajor>>(T.m);
    }
#else
    Matrix4f Compute(const MatrixXf& setA, const MatrixXf& setB) { return Compute(rowsOf(setA), rowsOf(setB)); }
#endif
private:
    static std::vector<Point3f> rowsOf(const MatrixXf& M)
    {
        std::vector<Point3f> v((size_t)M.rows());
        for (int i = 0; i < (int)M.rows(); ++i) v[(size_t)i] = Point3f(M(i, 0), M(i, 1), M(i, 2));
        return v;
    }
};

// ---- Odometry (RANSAC strategy only; the ICP / bundle-adjustment strategies are out of scope) ------------------------------
class Odometry {
public:
    enum eAlgorithm { RANSAC = 0, ADAPTIVE_RICP, MOTION_ONLY_BA, ADAPTIVE_RBA };
    Odometry(const eAlgorithm& algorithm = RANSAC) : mOdometryAlgorithm(algorithm), mpRansac(new Ransac(200, 20, 3.0f, 4)) {}   // odometry.cpp:14
    // Odometry::Compute, RANSAC strategy (odometry.cpp:78-90): Ransac::Iterate, then the composition rule T12 * pF1->GetPose() into
    // pF2's pose (cv::Mat's float product, evaluated by the device) and SetInlier(m.trainIdx) for the inliers.  void like the
    // reference; mbConverged keeps what Iterate returned.
    void Compute(Frame* pF1, Frame* pF2, const std::vector<DMatch>& vMatches12)
    {
        if (mOdometryAlgorithm != RANSAC) throw std::invalid_argument("orbf::Odometry: only the RANSAC strategy is on the GPU path");
        const Matrix4f pose1 = pF1->GetPose();
        Matrix4f T = pose1;                                 // Iterate returns before anything runs on too few matches: T12 = I, and I * pose is the pose, exactly
        mpRansac->mpSourceFrame = pF1; mpRansac->mpTargetFrame = pF2;
        mbConverged = mpRansac->Solve(vMatches12, &pose1, &T);            // Iterate + the composition rule on the device, one call
        mT12 = mpRansac->mT12; mvInliers = mpRansac->mvInliers;
        pF2->SetPose(T);
        for (const auto& m : mvInliers) pF2->SetInlier((size_t)m.trainIdx);
    }
    // The rest of Odometry::Compute (odometry.cpp:82-90) for a device-resident sequence (after orbf_track_sequence /
    // orbf_ransac_pairs on `npairs` consecutive pairs): poses[k + 1] = T12[k] * poses[k] with cv::Mat's float product, and the
    // frames' mvbOutlier flags after SetInlier(m.trainIdx).  poses: (npairs + 1) row-major 4x4; outlier may be null.
    static void ComposeTrajectory(int npairs, const Matrix4f* pose0, std::vector<Matrix4f>& poses, std::vector<uint8_t>* outlier = nullptr)
    {
        std::lock_guard<std::mutex> g(Runtime::Lock());
        orbf_context* ctx = Runtime::Current();
        static_assert(sizeof(Matrix4f) == 16 * sizeof(float), "poses are written as one contiguous block");
        poses.assign((size_t)npairs + 1, Matrix4f());
        int cap = 0;
        check(orbf_keypoint_capacity(ctx, &cap), "orbf_keypoint_capacity");
        if (outlier) outlier->assign(((size_t)npairs + 1) * (size_t)cap, 1);
        check(orbf_compose_trajectory(ctx, npairs, pose0 ? pose0->m : nullptr, poses[0].m, outlier ? outlier->data() : nullptr),
            "orbf_compose_trajectory");
    }
    eAlgorithm mOdometryAlgorithm;
    Matrix4f mT12;
    std::vector<DMatch> mvInliers;
    bool mbConverged = false;
    Ransac* ransac() { return mpRansac.get(); }
private:
    std::unique_ptr<Ransac> mpRansac;
};

}  // namespace orbf
