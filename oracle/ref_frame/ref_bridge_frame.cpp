// oracle/ref_frame/ref_bridge_frame.cpp — TEST INFRASTRUCTURE ONLY.
// (Its own directory: a quoted #include looks next to the including file first, and next to the other bridges sit the Core/ stand-ins.)
// C entry points over the reference's own Frame and Landmark: Core/frame.cpp, Core/keyframe.cpp, Core/landmark.cpp, Core/map.cpp,
// Features/extractor.cpp (+ the detector sources it names), Features/matcher.cpp and Features/orbextractor.cpp, each handed to g++
// verbatim from /root/reference by oracle/Makefile (target _ref): oracle/_ref/libframe_ref.so.  Here the reference's REAL Core classes
// are used (the stand-ins of ref_shim/Core belong to the other two libraries): -I$(REF) comes before -Iref_shim.
// Run from the reference's source: Frame::Frame (BGR -> gray, depth scale, K and the distortion vector from Utils/common.h),
// Frame::ExtractFeatures (Extractor::Extract -> ORBextractor, UndistortKeyPoints, the depth gather at the distorted position, mvuRight,
// the unprojection of the undistorted point, ComputeImageBounds), Landmark::ComputeDistinctiveDescriptors.
// Stubbed (each stops the process if reached): Database::Erase (needs the DBoW3 vocabulary) and Converter::toDescriptorVector (g2o).
#include "../ref_shim/ref_bridge.cpp"        // Features/orbextractor.cpp + the bump arena that fixes quirk Q3's tie order (ArenaScope)

#include <map>
#include <mutex>
#include <set>
#include <opencv2/opencv.hpp>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include "DBoW3/DBoW3.h"
#include "DBoW3/QueryResults.h"
// protected members of the reference classes (KeyFrame::mbBad) are set by the bridge without touching the source
#define private public
#define protected public
#include "Core/frame.h"
#include "Core/keyframe.h"
#include "Core/keyframedatabase.h"
#include "Core/landmark.h"
#include "Core/map.h"
#include "Features/extractor.h"
#include "Utils/common.h"
#include "Utils/converter.h"
#undef private
#undef protected

void Database::Erase(KeyFrame*) { std::abort(); }
std::vector<cv::Mat> Converter::toDescriptorVector(const cv::Mat&) { std::abort(); }

extern "C" {

// Frame(imColor, imDepth, 0) + ExtractFeatures(Extractor(ORB_SLAM2, ORB_SLAM2, NORMAL)) (Core/frame.cpp:18-45, 135-170; the calibration
// is the one compiled into Utils/common.h: FR1 with its distortion).  Outputs (any may be NULL): mvKeys, mDescriptors, mvKeysUn (x, y),
// mvKeys3Dc, mvuRight, mImGray, and Frame::mnMinX / mnMaxX / mnMinY / mnMaxY.
int ref_frame_extract(const uint8_t* bgr, const uint16_t* depth, int w, int h, orc_keypoint* kps, uint8_t* desc, float* xy_un, float* xyz, float* uright,
    int cap, int* n_out, uint8_t* gray, float* bounds4)
{
    if (!bgr || !depth || !n_out) return ORC_ERR_ARG;
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    int rc = ORC_OK;
    ArenaScope scope;
    {
        static Extractor* ex = nullptr;             // built once (prints its banner), outside nothing: lives in the arena-free heap of the first call
        if (!ex) { g_arena.active = false; ex = new Extractor(Extractor::ORB_SLAM2, Extractor::ORB_SLAM2, Extractor::NORMAL); g_arena.active = g_arena.base != nullptr; }
        cv::Mat color(h, w, CV_8UC3, const_cast<uint8_t*>(bgr), (size_t)w * 3);
        cv::Mat d16(h, w, CV_16U, const_cast<uint16_t*>(depth), (size_t)w * 2);
        Frame::mbInitialComputations = true;        // every call computes the image bounds, like the first frame of a run
        Frame f(color, d16, 0.0);
        f.ExtractFeatures(ex);
        *n_out = (int)f.N;
        if ((int)f.N > cap) rc = ORC_ERR_CAPACITY;
        else
            for (size_t i = 0; i < f.N; ++i) {
                const cv::KeyPoint& k = f.mvKeys[i];
                if (kps) { orc_keypoint o = { k.pt.x, k.pt.y, k.size, k.angle, k.response, k.octave, k.class_id }; kps[i] = o; }
                if (desc) std::memcpy(desc + 32 * i, f.mDescriptors.ptr((int)i), 32);
                if (xy_un) { xy_un[2 * i] = f.mvKeysUn[i].pt.x; xy_un[2 * i + 1] = f.mvKeysUn[i].pt.y; }
                if (xyz) { xyz[3 * i] = f.mvKeys3Dc[i].x; xyz[3 * i + 1] = f.mvKeys3Dc[i].y; xyz[3 * i + 2] = f.mvKeys3Dc[i].z; }
                if (uright) uright[i] = f.mvuRight[i];
            }
        if (gray) for (int r = 0; r < h; ++r) std::memcpy(gray + (size_t)r * w, f.mImGray.ptr(r), (size_t)w);
        if (bounds4) { bounds4[0] = Frame::mnMinX; bounds4[1] = Frame::mnMaxX; bounds4[2] = Frame::mnMinY; bounds4[3] = Frame::mnMaxY; }
    }
    return rc;
}

// Landmark::ComputeDistinctiveDescriptors (Core/landmark.cpp:219-273) for a batch of landmarks: landmark l is observed by
// offsets[l + 1] - offsets[l] keyframes, observation k holding descriptor row offsets[l] + k (bad[row] != 0: that keyframe isBad()).
// The reference walks a std::map<KeyFrame*, size_t>, i.e. in ADDRESS order: the keyframes of a landmark are constructed in one block at
// ascending addresses, so that order is the observation order.  out_desc[l] = the landmark's descriptor afterwards (32 bytes; zeros
// when it stayed empty), has[l] = whether it was set.
int ref_distinctive_descriptors(const uint8_t* desc, const uint8_t* bad, const int* offsets, int n_landmarks, uint8_t* out_desc, uint8_t* has)
{
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    Extractor::mNorm = cv::NORM_HAMMING;
    Map map;
    cv::Mat pos = cv::Mat::eye(3, 1, CV_32F);
    for (int l = 0; l < n_landmarks; ++l) {
        const int a = offsets[l], n = offsets[l + 1] - a;
        std::memset(out_desc + (size_t)l * 32, 0, 32); has[l] = 0;
        if (n <= 0) continue;
        void* block = std::malloc(sizeof(KeyFrame) * (size_t)n);
        std::vector<KeyFrame*> kfs;
        for (int k = 0; k < n; ++k) {
            Frame f;
            f.mnId = (long unsigned)k; f.N = 1;
            f.mDescriptors = cv::Mat(1, 32, CV_8UC1);
            std::memcpy(f.mDescriptors.data, desc + (size_t)(a + k) * 32, 32);
            f.mvuRight.assign(1, -1.f);
            f.mvKeys.resize(1); f.mvKeysUn.resize(1); f.mvKeys3Dc.resize(1);
            kfs.push_back(new (static_cast<char*>(block) + sizeof(KeyFrame) * (size_t)k) KeyFrame(f, &map, nullptr));
        }
        {
            Landmark lm(pos, kfs[0], &map);
            for (int k = 0; k < n; ++k) lm.AddObservation(kfs[(size_t)k], 0);
            for (int k = 0; k < n; ++k) if (bad && bad[a + k]) kfs[(size_t)k]->mbBad = true;
            lm.ComputeDistinctiveDescriptors();
            const cv::Mat d = lm.GetDescriptor();
            if (!d.empty()) { std::memcpy(out_desc + (size_t)l * 32, d.ptr(0), 32); has[l] = 1; }
        }
        for (KeyFrame* kf : kfs) kf->~KeyFrame();
        std::free(block);
    }
    return ORC_OK;
}

}  // extern "C"
