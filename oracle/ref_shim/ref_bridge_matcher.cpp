// oracle/ref_shim/ref_bridge_matcher.cpp — TEST INFRASTRUCTURE ONLY.
// C entry points over the reference's own Matcher (Features/matcher.cpp, handed to g++ verbatim from /root/reference by oracle/Makefile,
// target _ref): both KnnMatch overloads (row a-16), ProjectionMatch, BoWMatch and Fuse.  Frame / KeyFrame / Landmark
// are the stand-ins of ref_shim/Core (the reference's own pull in DBoW3, PCL and the map graph); cv::BFMatcher::knnMatch, cv::norm and
// the cv::Mat product are the cv2-pinned routines of the oracle.  Run from the reference's source: the ratio rule and landmark filters of
// KnnMatch, the ordered best / second-best search with its octave rule, the feature-vector merge walk with its used-train set, the
// projection, gates and search of Fuse.
#include <cstring>
#include <memory>
#include <vector>

#include "../oracle_api.h"
#include <opencv2/opencv.hpp>
#include "Core/frame.h"
#include "Core/keyframe.h"
#include "Core/landmark.h"
#include "Utils/common.h"

#define private public
#include "Features/matcher.h"
#undef private
#include "Features/extractor.h"

namespace {
cv::Mat desc_mat(const uint8_t* d, int n)
{
    cv::Mat m(n, 32, CV_8UC1);
    if (n > 0) std::memcpy(m.data, d, (size_t)n * 32);
    return m;
}
void fill_features(Frame& f, const float* kp_x, const float* kp_y, const int* kp_octave, const float* u_right, const uint8_t* desc, int n)
{
    f.Resize((size_t)n);
    for (int j = 0; j < n; ++j) {
        f.mvKeysUn[(size_t)j] = cv::KeyPoint(kp_x[j], kp_y[j], 31.f, -1.f, 0.f, kp_octave ? kp_octave[j] : 0);
        f.mvKeys[(size_t)j] = f.mvKeysUn[(size_t)j];
        if (u_right) f.mvuRight[(size_t)j] = u_right[j];
    }
    f.mDescriptors = desc_mat(desc, n);
}
Matcher make_matcher(float ratio, double th_low, double th_high)
{
    Extractor::mNorm = cv::NORM_HAMMING;              // what Extractor's constructor stores for a binary descriptor (extractor.cpp:36)
    Matcher m(ratio);
    if (th_low >= 0) m.TH_LOW = th_low;
    if (th_high >= 0) m.TH_HIGH = th_high;
    return m;
}
}  // namespace

extern "C" {

// Matcher(ratio).KnnMatch(Frame& F1, Frame& F2, matches) (matcher.cpp:55-88) with every feature of F1 holding an inlier landmark and F2
// empty: the landmark filters pass everything, what is left is kNN-2 + the ratio rule in query order.
int ref_knn_match_frames(const uint8_t* q, int nq, const uint8_t* t, int nt, float ratio, orc_dmatch* out, int cap, int* n_out)
{
    Frame f1, f2;
    f1.Resize((size_t)nq); f2.Resize((size_t)nt);
    f1.mDescriptors = desc_mat(q, nq); f2.mDescriptors = desc_mat(t, nt);
    std::vector<Landmark> lms((size_t)nq);
    for (int i = 0; i < nq; ++i) f1.mvpLandmarks[(size_t)i] = &lms[(size_t)i];
    Matcher m = make_matcher(ratio, -1, -1);
    std::vector<cv::DMatch> res;
    *n_out = (int)m.KnnMatch(f1, f2, res);
    if (*n_out > cap) return ORC_ERR_CAPACITY;
    for (size_t i = 0; i < res.size(); ++i) { orc_dmatch o = { res[i].queryIdx, res[i].trainIdx, res[i].imgIdx, res[i].distance }; out[i] = o; }
    return ORC_OK;
}

// Matcher(ratio).KnnMatch(KeyFrame* KF1, Frame& F2, matches) (matcher.cpp:23-53).  kf_lm[i]: landmark id held by feature i of the
// keyframe (0 = none); lm_bad[id]: isBad(); f2_lm[j] in / out: landmark id in slot j of the frame (0 = free).
int ref_knn_match_keyframe(const uint8_t* q, int nq, const uint8_t* t, int nt, float ratio, const int* kf_lm, const uint8_t* lm_bad, int n_ids,
    int* f2_lm, orc_dmatch* out, int cap, int* n_out)
{
    KeyFrame kf; Frame f2;
    kf.Resize((size_t)nq); f2.Resize((size_t)nt);
    kf.mDescriptors = desc_mat(q, nq); f2.mDescriptors = desc_mat(t, nt);
    std::vector<Landmark> lms((size_t)n_ids);
    for (int id = 0; id < n_ids; ++id) lms[(size_t)id].mbBad = lm_bad[id] != 0;
    for (int i = 0; i < nq; ++i) kf.mvpLandmarks[(size_t)i] = kf_lm[i] > 0 ? &lms[(size_t)kf_lm[i]] : nullptr;
    for (int j = 0; j < nt; ++j) f2.mvpLandmarks[(size_t)j] = f2_lm[j] > 0 ? &lms[(size_t)f2_lm[j]] : nullptr;
    Matcher m = make_matcher(ratio, -1, -1);
    std::vector<cv::DMatch> res;
    *n_out = (int)m.KnnMatch(&kf, f2, res);
    for (int j = 0; j < nt; ++j) f2_lm[j] = f2.mvpLandmarks[(size_t)j] ? (int)(f2.mvpLandmarks[(size_t)j] - lms.data()) : 0;
    if (*n_out > cap) return ORC_ERR_CAPACITY;
    for (size_t i = 0; i < res.size(); ++i) { orc_dmatch o = { res[i].queryIdx, res[i].trainIdx, res[i].imgIdx, res[i].distance }; out[i] = o; }
    return ORC_OK;
}

// Matcher(ratio).ProjectionMatch(frame, landmarks, radius) (matcher.cpp:90-143).  lm_flags[i] bit 0: mbTrackInView && !isBad(), bit 1:
// Observations() > 0; feat_taken[j] (may be NULL): slot j already holds a landmark with Observations() > 0.  best_idx[i]: the slot
// Frame::AddLandmark was called with for landmark i, or -1.
int ref_projection_match(const float* kp_x, const float* kp_y, const int* kp_octave, const uint8_t* desc, int n_feat, const uint8_t* lm_desc,
    const float* proj_x, const float* proj_y, const uint8_t* lm_flags, int n_landmarks, const uint8_t* feat_taken, float radius, float nn_ratio,
    double th_high, int* best_idx, int* n_matches)
{
    Frame f;
    fill_features(f, kp_x, kp_y, kp_octave, nullptr, desc, n_feat);
    std::vector<Landmark> lms((size_t)n_landmarks);
    std::vector<Landmark*> ptrs((size_t)n_landmarks);
    for (int i = 0; i < n_landmarks; ++i) {
        Landmark& l = lms[(size_t)i];
        l.mbTrackInView = (lm_flags[i] & 1) != 0; l.nObs = (lm_flags[i] & 2) ? 1 : 0;
        l.mTrackProjX = proj_x[i]; l.mTrackProjY = proj_y[i];
        l.mDescriptor = desc_mat(lm_desc + (size_t)i * 32, 1);
        ptrs[(size_t)i] = &l; best_idx[i] = -1;
    }
    Landmark occupied; occupied.nObs = 1;
    for (int j = 0; j < n_feat && feat_taken; ++j) if (feat_taken[j]) f.mvpLandmarks[(size_t)j] = &occupied;
    Matcher m = make_matcher(nn_ratio, -1, th_high);
    *n_matches = (int)m.ProjectionMatch(&f, ptrs, radius);
    for (const auto& e : f.addLog) best_idx[e.first - lms.data()] = (int)e.second;
    return ORC_OK;
}

// Matcher(ratio).BoWMatch(KF1, KF2, matches) (matcher.cpp:145-209) on feature vectors given as CSR (ascending word ids)
int ref_bow_match(const int* words1, const int* off1, const int* idx1, int nw1, const uint8_t* desc1, int n1, const int* words2, const int* off2,
    const int* idx2, int nw2, const uint8_t* desc2, int n2, float nn_ratio, double th_low, orc_dmatch* out, int cap, int* n_out)
{
    KeyFrame k1, k2;
    k1.Resize((size_t)n1); k2.Resize((size_t)n2);
    k1.mDescriptors = desc_mat(desc1, n1); k2.mDescriptors = desc_mat(desc2, n2);
    for (int a = 0; a < nw1; ++a) k1.mFeatVec[(unsigned)words1[a]].assign(idx1 + off1[a], idx1 + off1[a + 1]);
    for (int b = 0; b < nw2; ++b) k2.mFeatVec[(unsigned)words2[b]].assign(idx2 + off2[b], idx2 + off2[b + 1]);
    Matcher m = make_matcher(nn_ratio, th_low, -1);
    std::vector<cv::DMatch> res;
    *n_out = m.BoWMatch(&k1, &k2, res);
    if (*n_out > cap) return ORC_ERR_CAPACITY;
    for (size_t i = 0; i < res.size(); ++i) { orc_dmatch o = { res[i].queryIdx, res[i].trainIdx, res[i].imgIdx, res[i].distance }; out[i] = o; }
    return ORC_OK;
}

// Matcher().Fuse(KF, landmarks, radius) (matcher.cpp:212-311) with the reference's compiled-in FR1 calibration (Utils/common.h:35-38,
// mbf 40).  lm_state[i]: 0 = null pointer, 1 = valid, 2 = isBad(), 3 = already observed in the keyframe.  best_idx[i]: the feature the
// search settled on when it passed TH_LOW (read from the graph edit the reference then makes), else -1.  Returns nFused in *n_fused.
int ref_fuse(const float* Rcw, const float* tcw, float min_x, float max_x, float min_y, float max_y, const float* kp_x, const float* kp_y,
    const float* u_right, const uint8_t* desc, int n_feat, const float* lm_pos, const uint8_t* lm_desc, const uint8_t* lm_state, int n_landmarks,
    float radius, double th_low, int* best_idx, int* n_fused)
{
    KeyFrame kf;
    fill_features(kf, kp_x, kp_y, nullptr, u_right, desc, n_feat);
    kf.mnMinX = min_x; kf.mnMaxX = max_x; kf.mnMinY = min_y; kf.mnMaxY = max_y;
    kf.mRcw = cv::Mat(3, 3, CV_32F); kf.mtcw = cv::Mat(3, 1, CV_32F); kf.mOw = cv::Mat(3, 1, CV_32F);
    for (int i = 0; i < 3; ++i) { kf.mtcw.at<float>(i) = tcw[i]; kf.mOw.at<float>(i) = 0.f; for (int j = 0; j < 3; ++j) kf.mRcw.at<float>(i, j) = Rcw[3 * i + j]; }
    std::vector<Landmark> lms((size_t)n_landmarks);
    std::vector<Landmark*> ptrs((size_t)n_landmarks);
    for (int i = 0; i < n_landmarks; ++i) {
        Landmark& l = lms[(size_t)i];
        l.mbBad = lm_state[i] == 2;
        if (lm_state[i] == 3) l.mObservedIn.insert(&kf);
        l.mWorldPos = cv::Mat(3, 1, CV_32F);
        for (int k = 0; k < 3; ++k) l.mWorldPos.at<float>(k) = lm_pos[3 * i + k];
        l.mDescriptor = desc_mat(lm_desc + (size_t)i * 32, 1);
        ptrs[(size_t)i] = lm_state[i] == 0 ? nullptr : &l;
        best_idx[i] = -1;
    }
    Landmark::Log().clear();
    Matcher m = make_matcher(0.6f, th_low, -1);
    *n_fused = m.Fuse(&kf, ptrs, radius);
    // every landmark that passed TH_LOW made exactly one KeyFrame::GetLandmark(bestIdx) call, in landmark order, followed by the edit
    // that names it: AddObservation on itself (free slot) or a Replace between itself and the slot's landmark
    size_t g = 0;
    for (const Landmark::Event& e : Landmark::Log()) {
        if (g >= kf.getLog.size()) return ORC_ERR_GEOMETRY;
        const size_t slot = kf.getLog[g++];
        Landmark* cur = e.a;
        if (e.kind == 1) cur = (best_idx[e.a - lms.data()] < 0) ? e.a : e.b;       // the one of the two that is not in a slot yet
        best_idx[cur - lms.data()] = (int)slot;
    }
    Landmark::Log().clear();
    return ORC_OK;
}

}  // extern "C"
