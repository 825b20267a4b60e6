// oracle/ref_shim/ref_bridge_adaptive.cpp — TEST INFRASTRUCTURE ONLY.
// C entry points over the reference's own adaptive detector chain (row a-17, BASELINE config 4): Features/extractor.cpp
// (Extractor(FAST, ., ADAPTIVE): CreateAdaptiveDetector + Extract), videogridadaptedfeaturedetector.cpp,
// videodynamicadaptedfeaturedetector.cpp, detectoradjuster.cpp and statefulfeaturedetector.cpp, each compiled verbatim from
// /root/reference by oracle/Makefile (target _ref) against the OpenCV stand-in of this directory (cv::FAST = the cv2-pinned routine
// of orb_oracle.cpp; cv::KeyPointsFilter::retainBest restated there).  Run from the reference's source: the 3 x 3 grid with its
// 31-pixel overlap, the per-cell controllers (tooFew / tooMany / good, five attempts, state carried from frame to frame),
// keepStrongest through the real std::nth_element, the aggregation order, the final retainBest(nFeatures).
#include <cstring>
#include <vector>

#include "../oracle_api.h"
#include <opencv2/features2d.hpp>
#include <opencv2/xfeatures2d.hpp>

// the controllers' state lives in protected / private members; the bridge reads it without touching the source
#define private public
#define protected public
#include "Features/extractor.h"
#include "Features/detectoradjuster.h"
#include "Features/statefulfeaturedetector.h"
#include "Features/videodynamicadaptedfeaturedetector.h"
#include "Features/videogridadaptedfeaturedetector.h"
#undef private
#undef protected

extern "C" {

// new Extractor(FAST, ORB, ADAPTIVE) (Features/extractor.cpp:15-37, 52-77)
void* ref_adaptive_create(void) { return new Extractor(Extractor::FAST, Extractor::ORB, Extractor::ADAPTIVE); }
void ref_adaptive_destroy(void* p) { delete static_cast<Extractor*>(p); }

// Extractor::Extract(image, noArray(), keypoints, descriptors) (extractor.cpp:39-50) on the next frame of the video; thresh (optional,
// 9 doubles): DetectorAdjuster::mThresh of every grid cell after the call, row-major.
int ref_adaptive_extract(void* p, const uint8_t* img, int w, int h, int stride, orc_keypoint* out, int cap, int* n_out, double* thresh)
{
    Extractor* ex = static_cast<Extractor*>(p);
    if (!ex || !img || !n_out) return ORC_ERR_ARG;
    cv::Mat image(h, w, CV_8UC1, const_cast<uint8_t*>(img), (size_t)stride);
    std::vector<cv::KeyPoint> keys;
    cv::Mat desc;
    ex->Extract(image, cv::noArray(), keys, desc);
    *n_out = (int)keys.size();
    if (thresh) {
        VideoGridAdaptedFeatureDetector* grid = dynamic_cast<VideoGridAdaptedFeatureDetector*>(ex->mpDetector.get());
        if (!grid) return ORC_ERR_ARG;
        for (size_t k = 0; k < grid->vpDetectors.size(); ++k) {
            VideoDynamicAdaptedFeatureDetector* dyn = dynamic_cast<VideoDynamicAdaptedFeatureDetector*>(grid->vpDetectors[k].get());
            if (!dyn) return ORC_ERR_ARG;
            thresh[k] = dyn->mpAdjuster->mThresh;
        }
    }
    if ((int)keys.size() > cap) return ORC_ERR_CAPACITY;
    for (size_t i = 0; i < keys.size() && out; ++i) {
        const cv::KeyPoint& k = keys[i];
        orc_keypoint o = { k.pt.x, k.pt.y, k.size, k.angle, k.response, k.octave, k.class_id };
        out[i] = o;
    }
    return ORC_OK;
}

}  // extern "C"
