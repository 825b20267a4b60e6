// oracle/ref_shim/Core/frame.h — TEST INFRASTRUCTURE ONLY.
// Stand-in for the reference's Core/frame.h + frame.cpp (which pull in DBoW3, PCL filters and the whole map graph and cannot be compiled
// here): the members of Frame that Odometry/ransac.cpp and Features/matcher.cpp touch, with the reference's names and types
// (Core/frame.h:82-130).  The few methods are restated from Core/frame.cpp (cited per method); AddLandmark additionally appends to a log
// so that a bridge can report which slot each landmark took.
#pragma once
#include <cmath>
#include <utility>
#include <vector>
#include <opencv2/opencv.hpp>
#include "DBoW3/DBoW3.h"

class Landmark;
class KeyFrame;

class Frame {
public:
    virtual ~Frame() {}

    // frame.cpp:172-186, 276-284
    virtual void AddLandmark(Landmark* pLandmark, const size_t& idx) { mvpLandmarks[idx] = pLandmark; addLog.push_back(std::make_pair(pLandmark, idx)); }
    virtual std::vector<Landmark*> GetLandmarks() { return mvpLandmarks; }
    virtual Landmark* GetLandmark(const size_t& idx) { getLog.push_back(idx); return mvpLandmarks[idx]; }
    void SetOutlier(const size_t& idx) { mvbOutlier[idx] = true; }
    void SetInlier(const size_t& idx) { mvbOutlier[idx] = false; }
    bool IsOutlier(const size_t& idx) { return mvbOutlier[idx] == true; }
    bool IsInlier(const size_t& idx) { return mvbOutlier[idx] == false; }

    // frame.cpp:80-84
    virtual cv::Mat GetCameraCenter() { return mOw.clone(); }
    virtual cv::Mat GetRotation() { return mRcw.clone(); }
    virtual cv::Mat GetTranslation() { return mtcw.clone(); }

    // frame.cpp:258-274
    std::vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r) const
    {
        std::vector<size_t> vIndices;
        vIndices.reserve(N);
        for (size_t i = 0; i < N; ++i) {
            const cv::KeyPoint& kpU = mvKeysUn[i];
            const float distx = kpU.pt.x - x;
            const float disty = kpU.pt.y - y;
            if (std::fabs(distx) < r && std::fabs(disty) < r)
                vIndices.push_back(i);
        }
        return vIndices;
    }

    void Resize(size_t n)                     // stand-in only: size every per-feature container
    {
        N = n; mvKeys.resize(n); mvKeysUn.resize(n); mvKeys3Dc.resize(n); mvuRight.assign(n, -1.f);
        mvbOutlier.assign(n, false); mvpLandmarks.assign(n, nullptr);
    }

    cv::Mat mImColor, mImGray, mImDepth;
    std::vector<cv::KeyPoint> mvKeys;
    std::vector<cv::KeyPoint> mvKeysUn;
    std::vector<cv::Point3f> mvKeys3Dc;
    std::vector<float> mvuRight;
    cv::Mat mDescriptors;
    DBoW3::BowVector mBowVec;
    DBoW3::FeatureVector mFeatVec;
    size_t N = 0;

    cv::Mat mRcw, mtcw, mOw;
    std::vector<bool> mvbOutlier;
    std::vector<Landmark*> mvpLandmarks;
    std::vector<std::pair<Landmark*, size_t>> addLog;     // stand-in only
    std::vector<size_t> getLog;                           // stand-in only
};
