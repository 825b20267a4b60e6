// oracle/ref_frame/ref_bridge_frame.cpp — TEST INFRASTRUCTURE ONLY.
// (Its own directory: a quoted #include looks next to the including file first, and next to the other bridges sit the Core/ stand-ins.)
// C entry points over the reference's own Frame and Landmark: Core/frame.cpp, Core/keyframe.cpp, Core/landmark.cpp, Core/map.cpp,
// Features/extractor.cpp (+ the detector sources it names), Features/matcher.cpp, Features/orbextractor.cpp, Odometry/odometry.cpp and
// Odometry/ransac.cpp, each handed to g++
// verbatim from /root/reference by oracle/Makefile (target _ref): oracle/_ref/libframe_ref.so.  Here the reference's REAL Core classes
// are used (the stand-ins of ref_shim/Core belong to the other two libraries): -I$(REF) comes before -Iref_shim.
// Run from the reference's source: Frame::Frame (BGR -> gray, depth scale, K and the distortion vector from Utils/common.h),
// Frame::ExtractFeatures (Extractor::Extract -> ORBextractor, UndistortKeyPoints, the depth gather at the distorted position, mvuRight,
// the unprojection of the undistorted point, ComputeImageBounds), Landmark::ComputeDistinctiveDescriptors, Odometry::Compute (RANSAC
// strategy: Ransac::Iterate on real Frame objects, the composition rule through cv::Mat, SetPose, SetInlier).
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
#include "Features/matcher.h"
#include "Utils/common.h"
#include "Utils/converter.h"
#include "Odometry/generalizedicp.h"
#include "Odometry/odometry.h"
#include "Odometry/pnpsolver.h"
#include "Odometry/ransac.h"
#undef private
#undef protected

// the two refinements behind the RANSAC strategy of Odometry::Compute are third-party iterative solvers (PCL GICP, g2o): never constructed here
GeneralizedICP::GeneralizedICP(int, double) { std::abort(); }
bool GeneralizedICP::Compute(const pcl::PointCloud<pcl::PointXYZ>::Ptr, const pcl::PointCloud<pcl::PointXYZ>::Ptr, const Eigen::Matrix4f&, const bool) { std::abort(); }
int PnPSolver::Compute(Frame*) { std::abort(); }
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

// Matcher(ratio).KnnMatch(Frame& F1, Frame& F2, matches) (Features/matcher.cpp:55-88) on the reference's REAL Frame / Landmark / Map
// objects.  lm_obs1[i]: -1 = feature i of F1 holds no landmark, else that landmark's Observations(); outlier1[i]: F1.IsOutlier(i);
// lm_obs2[j]: the same for F2's slots before the call.  slot2[j] (out): index of the F1 feature whose landmark sits in slot j of F2 after
// the call, -1 = none or the landmark that was there before; outlier2[j] (out): F2's mvbOutlier.
int ref_real_knn_match_frames(const uint8_t* q, int nq, const uint8_t* t, int nt, float ratio, const int* lm_obs1, const uint8_t* outlier1, const int* lm_obs2,
    orc_dmatch* out, int cap, int* n_out, int* slot2, uint8_t* outlier2)
{
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    Extractor::mNorm = cv::NORM_HAMMING;
    Map map;
    Frame f1, f2;
    auto fill = [](Frame& f, const uint8_t* d, int n) {
        f.N = (size_t)n; f.mvbOutlier.assign((size_t)n, false); f.mvpLandmarks.assign((size_t)n, nullptr); f.mvuRight.assign((size_t)n, -1.f);
        f.mDescriptors = cv::Mat(n, 32, CV_8UC1);
        if (n > 0) std::memcpy(f.mDescriptors.data, d, (size_t)n * 32);
    };
    fill(f1, q, nq); fill(f2, t, nt);
    cv::Mat pos = cv::Mat::eye(3, 1, CV_32F);
    std::vector<std::unique_ptr<Landmark>> lms1, lms2;
    std::map<Landmark*, int> owner;
    for (int i = 0; i < nq; ++i) {
        if (outlier1[i]) f1.mvbOutlier[(size_t)i] = true;
        if (lm_obs1[i] < 0) continue;
        lms1.emplace_back(new Landmark(pos, &map, &f1, (size_t)i));
        lms1.back()->nObs = lm_obs1[i];
        f1.mvpLandmarks[(size_t)i] = lms1.back().get(); owner[lms1.back().get()] = i;
    }
    for (int j = 0; j < nt; ++j) {
        if (lm_obs2[j] < 0) continue;
        lms2.emplace_back(new Landmark(pos, &map, &f2, (size_t)j));
        lms2.back()->nObs = lm_obs2[j];
        f2.mvpLandmarks[(size_t)j] = lms2.back().get();
    }
    Matcher m(ratio);
    std::vector<cv::DMatch> res;
    *n_out = (int)m.KnnMatch(f1, f2, res);
    for (int j = 0; j < nt; ++j) {
        Landmark* p = f2.mvpLandmarks[(size_t)j];
        slot2[j] = (p && owner.count(p)) ? owner[p] : -1;
        outlier2[j] = f2.mvbOutlier[(size_t)j] ? 1 : 0;
    }
    if (*n_out > cap) return ORC_ERR_CAPACITY;
    for (size_t i = 0; i < res.size(); ++i) { orc_dmatch o = { res[i].queryIdx, res[i].trainIdx, res[i].imgIdx, res[i].distance }; out[i] = o; }
    return ORC_OK;
}

// Ransac::DepthCovariance of THIS library's copy of Odometry/ransac.cpp (quirk Q7: the first call in the process fixes the value)
double ref_frame_depth_covariance(double depth)
{
    Ransac r;
    return r.DepthCovariance(depth);
}

// Odometry(RANSAC).Compute(pF1, pF2, m12) (Odometry/odometry.cpp:9-31, 44, 78-90) on the reference's real Frame objects after srand(seed):
// Ransac::Iterate, T12 * pF1->GetPose() as cv::Mat evaluates it, pF2->SetPose, pF2->SetInlier(m.trainIdx).  pose1: F1's Tcw (row-major 4x4);
// outlier2 [ndst]: F2's mvbOutlier, all true before the call (what the matcher's SetOutlier leaves for matched features).
int ref_odometry_compute(const float* src_xyz, int nsrc, const float* dst_xyz, int ndst, const orc_dmatch* m12, int nm, unsigned seed, const float* pose1,
    float* pose2, uint8_t* outlier2, orc_ransac_out* out, orc_dmatch* inliers_out, int cap)
{
    if (!out || !pose1 || !pose2) return ORC_ERR_ARG;
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    std::memset(out, 0, sizeof(*out));
    Frame f1, f2;
    auto fill = [](Frame& f, const float* xyz, int n) {
        f.N = (size_t)n; f.mvKeys3Dc.resize((size_t)n); f.mvbOutlier.assign((size_t)n, true); f.mvpLandmarks.assign((size_t)n, nullptr);
        for (int i = 0; i < n; ++i) f.mvKeys3Dc[(size_t)i] = cv::Point3f(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    };
    fill(f1, src_xyz, nsrc); fill(f2, dst_xyz, ndst);
    cv::Mat T1(4, 4, CV_32F);
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) T1.at<float>(i, j) = pose1[4 * i + j];
    f1.SetPose(T1);
    std::vector<cv::DMatch> ms((size_t)nm);
    for (int i = 0; i < nm; ++i) ms[(size_t)i] = cv::DMatch(m12[i].queryIdx, m12[i].trainIdx, m12[i].imgIdx, m12[i].distance);
    // the reference's constructor leaves mpBA unset for the RANSAC strategy and its destructor deletes it when non-null: built in zeroed storage
    void* mem = std::calloc(1, sizeof(Odometry));
    Odometry* od = new (mem) Odometry(Odometry::RANSAC);
    srand(seed);
    od->Compute(&f1, &f2, ms);
    const Ransac* r = od->mpRansac;
    out->rmse = r->rmse; out->n_inliers = (int)r->mvInliers.size(); out->ok = out->n_inliers >= 20 ? 1 : 0;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) out->T12[4 * i + j] = r->mT12(i, j);
    int rc = out->n_inliers > cap ? ORC_ERR_CAPACITY : ORC_OK;
    for (int i = 0; i < out->n_inliers && inliers_out && rc == ORC_OK; ++i) {
        const cv::DMatch& m = r->mvInliers[(size_t)i];
        orc_dmatch o = { m.queryIdx, m.trainIdx, m.imgIdx, m.distance };
        inliers_out[i] = o;
    }
    const cv::Mat T2 = f2.GetPose();
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) pose2[4 * i + j] = T2.at<float>(i, j);
    if (outlier2) for (int j = 0; j < ndst; ++j) outlier2[j] = f2.mvbOutlier[(size_t)j] ? 1 : 0;
    od->~Odometry();
    std::free(mem);
    return rc;
}

}  // extern "C"
