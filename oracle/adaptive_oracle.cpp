// oracle/adaptive_oracle.cpp — CPU restatement of the reference's adaptive FAST detector route (TEST INFRASTRUCTURE ONLY).
// Follows /root/reference:
//   Extractor::CreateAdaptiveDetector   Features/extractor.cpp:52-77     FAST adjuster (20, 2, 10000, 1.3, 0.7), 3x3 grid,
//                                                                         min 600/9 = 67, max 1020/9 = 113, 5 iterations, edge 31
//   VideoGridAdaptedFeatureDetector::detect   Features/videogridadaptedfeaturedetector.cpp:52-84  overlapping sub-images, keepStrongest
//   VideoDynamicAdaptedFeatureDetector::detect Features/videodynamicadaptedfeaturedetector.cpp:24-44 the per-cell threshold loop
//   DetectorAdjuster::detect/tooFew/tooMany/good  Features/detectoradjuster.cpp:22-59            cv::FastFeatureDetector::create(int(mThresh))
//   Extractor::Extract (non-ORB_SLAM2 route)  Features/extractor.cpp:44-46                      KeyPointsFilter::retainBest(nFeatures)
// cv::FAST on a sub-image view = orc_fast_roi (pinned against cv2 in tests/test_oracle_vs_cv2.py).
// Definitions where the reference is unspecified (std::nth_element leaves both the order and, among equal responses at the
// cut, the kept subset implementation-defined; quirk Q15): keepStrongest keeps the N largest responses, equal responses in
// detection (row-major) order, and the survivors stay in detection order.  retainBest keeps every keypoint whose response is
// >= the N-th largest (OpenCV keeps all ties), in aggregate order.
#include <algorithm>
#include <cmath>
#include <vector>

#include "oracle_api.h"

int orc_adaptive_default(orc_adaptive_cfg* c)
{
    if (!c) return ORC_ERR_ARG;
    c->grid = 3; c->edge = 31; c->max_iters = 5;
    const int minFeatures = 600, maxFeatures = (int)(minFeatures * 1.7);          // extractor.cpp:65-66
    c->min_features = (int)std::round(minFeatures / 9.0f);                        // :71
    c->max_features = (int)std::round(maxFeatures / 9.0f);                        // :72
    c->max_per_cell = maxFeatures / 9;                                            // videogridadaptedfeaturedetector.cpp:58
    c->init_th = 20; c->min_th = 2; c->max_th = 10000; c->inc = 1.3; c->dec = 0.7;   // extractor.cpp:56
    return ORC_OK;
}

static void keep_strongest(std::vector<orc_cand>& v, int N)
{
    if ((int)v.size() <= N) return;
    std::vector<int> idx(v.size());
    for (size_t i = 0; i < idx.size(); ++i) idx[i] = (int)i;
    std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return v[a].score > v[b].score; });
    std::vector<char> keep(v.size(), 0);
    for (int i = 0; i < N; ++i) keep[idx[i]] = 1;
    std::vector<orc_cand> out;
    for (size_t i = 0; i < v.size(); ++i) if (keep[i]) out.push_back(v[i]);
    v.swap(out);
}

int orc_adaptive_detect(const orc_adaptive_cfg* cfg, const uint8_t* img, int w, int h, int stride, double* thresh, int retain_best,
    orc_keypoint* out, int cap, int* n_out, int* cell_found, int* cell_thresh)
{
    if (!cfg || !img || !thresh || !n_out || cfg->grid < 1 || cfg->grid > 5) return ORC_ERR_ARG;
    const int g = cfg->grid;
    std::vector<orc_keypoint> all;
    std::vector<orc_cand> kps;
    for (int i = 0; i < g; ++i) {
        const int rowstart = std::max((i * h) / g - cfg->edge, 0), rowend = std::min(h, ((i + 1) * h) / g + cfg->edge);
        for (int j = 0; j < g; ++j) {
            const int colstart = std::max((j * w) / g - cfg->edge, 0), colend = std::min(w, ((j + 1) * w) / g + cfg->edge);
            double& th = thresh[i * g + j];
            int iterCount = cfg->max_iters, usedTh = 0;
            const int sw = colend - colstart, sh = rowend - rowstart;
            do {
                usedTh = (int)th;                                                  // FastFeatureDetector::create(int threshold)
                int n = 0;
                kps.assign((size_t)sw * sh / 4 + 16, orc_cand());
                const int rc = orc_fast_roi(img + (size_t)rowstart * stride + colstart, stride, sw, sh, usedTh, kps.data(), (int)kps.size(), &n);
                if (rc != ORC_OK) return rc;
                kps.resize(n);
                if (n < cfg->min_features) {                                       // tooFew
                    th *= cfg->dec;
                    if (th < cfg->min_th) th = cfg->min_th;
                } else if (n > cfg->max_features) {                                // tooMany
                    th *= cfg->inc;
                    if (th > cfg->max_th) th = cfg->max_th;
                    break;
                } else break;
                iterCount--;
            } while (iterCount > 0 && th > cfg->min_th && th < cfg->max_th);       // good()
            if (cell_found) cell_found[i * g + j] = (int)kps.size();
            if (cell_thresh) cell_thresh[i * g + j] = usedTh;
            keep_strongest(kps, cfg->max_per_cell);
            for (const orc_cand& c : kps) {
                orc_keypoint k;
                k.x = (float)(c.x + colstart); k.y = (float)(c.y + rowstart); k.size = 7.f; k.angle = -1.f;
                k.response = (float)c.score; k.octave = 0; k.class_id = -1;
                all.push_back(k);
            }
        }
    }
    if (retain_best > 0 && (int)all.size() > retain_best) {
        std::vector<float> r(all.size());
        for (size_t i = 0; i < all.size(); ++i) r[i] = all[i].response;
        std::nth_element(r.begin(), r.begin() + (retain_best - 1), r.end(), std::greater<float>());
        const float cut = r[retain_best - 1];
        std::vector<orc_keypoint> keep;
        for (const orc_keypoint& k : all) if (k.response >= cut) keep.push_back(k);
        all.swap(keep);
    }
    *n_out = (int)all.size();
    if ((int)all.size() > cap) return ORC_ERR_CAPACITY;
    if (out) std::copy(all.begin(), all.end(), out);
    return ORC_OK;
}

// ---- Landmark::ComputeDistinctiveDescriptors (Core/landmark.cpp:219-273), SURVEY.md §8f rank 1 ----------------------------
// For every landmark: all-pairs Hamming distances of its observed descriptors, per row the median sorted[(size_t)(0.5*(N-1))],
// the first row with the strictly smallest median is the landmark's descriptor.  offsets[l]..offsets[l+1] delimit landmark l's
// rows in `desc`; best[l] = chosen row index inside the landmark (-1 for a landmark without observations).
int orc_distinctive_descriptors(const uint8_t* desc, const int* offsets, int n_landmarks, int* best, int* best_median)
{
    if (!desc || !offsets || !best || n_landmarks < 0) return ORC_ERR_ARG;
    std::vector<double> row;
    for (int l = 0; l < n_landmarks; ++l) {
        const int a = offsets[l], N = offsets[l + 1] - a;
        best[l] = -1;
        if (best_median) best_median[l] = -1;
        if (N <= 0) continue;
        double bestMedian = 1e300;
        int bestIdx = 0;
        for (int i = 0; i < N; ++i) {
            row.assign(N, 0.0);
            for (int j = 0; j < N; ++j) row[j] = (i == j) ? 0.0 : (double)orc_hamming(desc + (size_t)(a + i) * 32, desc + (size_t)(a + j) * 32);
            std::sort(row.begin(), row.end());
            const double median = row[(size_t)(0.5 * (N - 1))];
            if (median < bestMedian) { bestMedian = median; bestIdx = i; }
        }
        best[l] = bestIdx;
        if (best_median) best_median[l] = (int)bestMedian;
    }
    return ORC_OK;
}

// ---- Matcher::ProjectionMatch (Features/matcher.cpp:90-143), SURVEY.md §8f rank 1 ---------------------------------------------
// Landmarks already projected into the frame are matched IN ORDER: Frame::GetFeaturesInArea (Core/frame.cpp:258-274) is a linear
// scan in feature order over the square window |dx| < r && |dy| < r (float); a feature whose slot holds a landmark with
// Observations() > 0 is skipped; best / second best by strict '<' on the double-valued Hamming distance; accepted when
// best <= TH_HIGH unless both come from the same octave and best > ratio * second (float ratio promoted to double);
// Frame::AddLandmark then puts the landmark into the slot, which later landmarks see.
//   lm_flags[i] bit 0: mbTrackInView && !isBad();  bit 1: Observations() > 0.
//   feat_taken[j] (may be NULL): slot j initially holds a landmark with Observations() > 0.
int orc_projection_match(const float* kp_x, const float* kp_y, const int* kp_octave, const uint8_t* desc, int n_feat, const uint8_t* lm_desc,
    const float* proj_x, const float* proj_y, const uint8_t* lm_flags, int n_landmarks, const uint8_t* feat_taken, float radius, float nn_ratio,
    double th_high, int* best_idx, int* n_matches)
{
    if (n_feat < 0 || n_landmarks < 0 || !best_idx || !n_matches) return ORC_ERR_ARG;
    if (n_landmarks > 0 && (!lm_desc || !proj_x || !proj_y || !lm_flags)) return ORC_ERR_ARG;
    if (n_feat > 0 && (!kp_x || !kp_y || !kp_octave || !desc)) return ORC_ERR_ARG;
    std::vector<uint8_t> taken(std::max(n_feat, 1), 0);
    if (feat_taken) std::copy(feat_taken, feat_taken + n_feat, taken.begin());
    int nm = 0;
    for (int i = 0; i < n_landmarks; ++i) {
        best_idx[i] = -1;
        if (!(lm_flags[i] & 1)) continue;
        double bestDist1 = 1.7976931348623157e308, bestDist2 = 1.7976931348623157e308;
        int bestLevel = -1, bestLevel2 = -1, bestIdx = -1;
        bool any = false;
        for (int j = 0; j < n_feat; ++j) {
            const float distx = kp_x[j] - proj_x[i], disty = kp_y[j] - proj_y[i];
            if (!(std::fabs(distx) < radius && std::fabs(disty) < radius)) continue;
            any = true;
            if (taken[j]) continue;
            const double dist = (double)orc_hamming(lm_desc + (size_t)i * 32, desc + (size_t)j * 32);
            if (dist < bestDist1) {
                bestDist2 = bestDist1; bestDist1 = dist; bestLevel2 = bestLevel; bestLevel = kp_octave[j]; bestIdx = j;
            } else if (dist < bestDist2) {
                bestLevel2 = kp_octave[j]; bestDist2 = dist;
            }
        }
        if (!any) continue;
        if (bestDist1 <= th_high) {
            if (bestLevel == bestLevel2 && bestDist1 > nn_ratio * bestDist2) continue;
            best_idx[i] = bestIdx;
            if (lm_flags[i] & 2) taken[bestIdx] = 1;
            ++nm;
        }
    }
    *n_matches = nm;
    return ORC_OK;
}

// ---- Matcher::Fuse, search part (Features/matcher.cpp:212-296), SURVEY.md §8f rank 1 ---------------------------------------------
// Per landmark (independent of the others): p3Dc = Rcw * p3Dw + tcw as cv::gemm evaluates it for 3x3 * 3x1 (float products summed
// left to right in float, then (float)((double)t * 1.0 + (double)tcw * 1.0); pinned against cv2.gemm in tests/test_fuse_bow.py),
// rejected for z < 0, pinhole projection with separate multiply and add, KeyFrame::IsInImage, ur = u - mbf / z;
// Frame::GetFeaturesInArea window in feature order; stereo (mvuRight >= 0) / mono reprojection gates 7.8 / 5.99; strict '<' on the
// Hamming distance; accepted when best <= TH_LOW.  What follows in the reference (Replace / AddObservation / AddLandmark,
// matcher.cpp:297-311) edits the map graph and stays with the caller.
//   lm_valid[i]: pLM && !pLM->isBad() && !pLM->IsInKeyFrame(pKF).   best_idx[i] = -1 when nothing is fused.
int orc_fuse_search(const float* Rcw /* 9, row-major */, const float* tcw /* 3 */, float fx, float fy, float cx, float cy, float mbf, float min_x, float max_x,
    float min_y, float max_y, const float* kp_x, const float* kp_y, const float* u_right, const uint8_t* desc, int n_feat, const float* lm_pos /* 3 per landmark */,
    const uint8_t* lm_desc, const uint8_t* lm_valid, int n_landmarks, float radius, double th_low, int* best_idx, int* best_dist)
{
    if (n_feat < 0 || n_landmarks < 0 || !best_idx || !Rcw || !tcw) return ORC_ERR_ARG;
    if (n_landmarks > 0 && (!lm_pos || !lm_desc || !lm_valid)) return ORC_ERR_ARG;
    if (n_feat > 0 && (!kp_x || !kp_y || !u_right || !desc)) return ORC_ERR_ARG;
    for (int i = 0; i < n_landmarks; ++i) {
        best_idx[i] = -1;
        if (best_dist) best_dist[i] = -1;
        if (!lm_valid[i]) continue;
        float pc[3];
        for (int r = 0; r < 3; ++r) {
            float t = Rcw[3 * r] * lm_pos[3 * i];
            t = t + Rcw[3 * r + 1] * lm_pos[3 * i + 1];
            t = t + Rcw[3 * r + 2] * lm_pos[3 * i + 2];
            pc[r] = (float)((double)t * 1.0 + (double)tcw[r] * 1.0);
        }
        if (pc[2] < 0.0f) continue;
        const float invz = 1 / pc[2];
        const float x = pc[0] * invz, y = pc[1] * invz;
        float u = fx * x; u = u + cx;
        float v = fy * y; v = v + cy;
        if (!(u >= min_x && u < max_x && v >= min_y && v < max_y)) continue;
        float ur = mbf * invz; ur = u - ur;
        double bestDist = 1.7976931348623157e308;
        int bestIdx = -1;
        for (int j = 0; j < n_feat; ++j) {
            const float distx = kp_x[j] - u, disty = kp_y[j] - v;
            if (!(std::fabs(distx) < radius && std::fabs(disty) < radius)) continue;
            const float ex = u - kp_x[j], ey = v - kp_y[j];
            if (u_right[j] >= 0) {
                const float er = ur - u_right[j];
                float e2 = ex * ex; e2 = e2 + ey * ey; e2 = e2 + er * er;
                if (e2 > 7.8f) continue;
            } else {
                float e2 = ex * ex; e2 = e2 + ey * ey;
                if (e2 > 5.99f) continue;
            }
            const double dist = (double)orc_hamming(lm_desc + (size_t)i * 32, desc + (size_t)j * 32);
            if (dist < bestDist) { bestDist = dist; bestIdx = j; }
        }
        if (bestDist <= th_low) {
            best_idx[i] = bestIdx;
            if (best_dist) best_dist[i] = (int)bestDist;
        }
    }
    return ORC_OK;
}

// ---- Matcher::BoWMatch (Features/matcher.cpp:145-209), SURVEY.md §8f rank 1 -------------------------------------------------------
// The DBoW3 feature vectors arrive as CSR: words1[nw1] ascending node ids, off1[nw1 + 1], idx1[...] feature indices in the
// std::vector order (same for keyframe 2).  For every word present in both, every feature of keyframe 1 in that word looks for its
// best / second best (strict '<', bucket order) among keyframe 2's features of the word; kept when best <= TH_LOW and
// (float)best < ratio * (float)second, unless the train feature was already used by an earlier query (std::set trainIdxs).
int orc_bow_match(const int* words1, const int* off1, const int* idx1, int nw1, const uint8_t* desc1, const int* words2, const int* off2, const int* idx2,
    int nw2, const uint8_t* desc2, float nn_ratio, double th_low, orc_dmatch* out, int cap, int* n_out)
{
    if (nw1 < 0 || nw2 < 0 || !n_out) return ORC_ERR_ARG;
    std::vector<orc_dmatch> res;
    std::vector<int> used;
    int a = 0, b = 0;
    while (a < nw1 && b < nw2) {
        if (words1[a] == words2[b]) {
            for (int e1 = off1[a]; e1 < off1[a + 1]; ++e1) {
                const int q = idx1[e1];
                double bestDist1 = 1.7976931348623157e308, bestDist2 = 1.7976931348623157e308;
                int bestTrain = -1;
                for (int e2 = off2[b]; e2 < off2[b + 1]; ++e2) {
                    const int t = idx2[e2];
                    const double dist = (double)orc_hamming(desc1 + (size_t)q * 32, desc2 + (size_t)t * 32);
                    if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestTrain = t; }
                    else if (dist < bestDist2) bestDist2 = dist;
                }
                if (bestDist1 <= th_low) {
                    if (static_cast<float>(bestDist1) < nn_ratio * static_cast<float>(bestDist2)) {
                        if (std::find(used.begin(), used.end(), bestTrain) != used.end()) continue;
                        orc_dmatch m; m.queryIdx = q; m.trainIdx = bestTrain; m.imgIdx = -1; m.distance = static_cast<float>(bestDist1);
                        res.push_back(m);
                        used.push_back(bestTrain);
                    }
                }
            }
            ++a; ++b;
        } else if (words1[a] < words2[b]) {
            while (a < nw1 && words1[a] < words2[b]) ++a;          // lower_bound on a sorted map
        } else {
            while (b < nw2 && words2[b] < words1[a]) ++b;
        }
    }
    *n_out = (int)res.size();
    if ((int)res.size() > cap) return ORC_ERR_CAPACITY;
    if (out) std::copy(res.begin(), res.end(), out);
    return ORC_OK;
}

// ---- Frame::UndistortKeyPoints (Core/frame.cpp:286-313) = cv::undistortPoints(pts, pts, K, dist, Mat(), K), SURVEY.md §8f rank 2 ---
// OpenCV's cvUndistortPointsInternal with its default criteria (5 fixed iterations, no epsilon test), all in double: normalise,
// iterate x = (x0 - deltaX) * icdist with the radial (k1 k2 k3) and tangential (p1 p2) terms, re-project with the same camera
// matrix, round to float.  dist = {k1, k2, p1, p2, k3}.  Pinned against cv2.undistortPoints in tests/test_undistort.py.
int orc_undistort_points(const float* xy /* n x 2 */, int n, float fx, float fy, float cx, float cy, const float* dist /* 5 */, float* out /* n x 2 */)
{
    if (n < 0 || !dist || (n > 0 && (!xy || !out))) return ORC_ERR_ARG;
    const double k0 = dist[0], k1 = dist[1], p1 = dist[2], p2 = dist[3], k4 = dist[4];
    const double dfx = fx, dfy = fy, dcx = cx, dcy = cy, ifx = 1.0 / dfx, ify = 1.0 / dfy;
    for (int i = 0; i < n; ++i) {
        const double px = xy[2 * i], py = xy[2 * i + 1];
        double x = (px - dcx) * ifx, y = (py - dcy) * ify;
        const double x0 = x, y0 = y;
        for (int j = 0; j < 5; ++j) {
            const double r2 = x * x + y * y;
            const double icdist = (1 + ((0.0 * r2 + 0.0) * r2 + 0.0) * r2) / (1 + ((k4 * r2 + k1) * r2 + k0) * r2);
            if (icdist < 0) { x = (px - dcx) * ifx; y = (py - dcy) * ify; break; }
            const double deltaX = 2 * p1 * x * y + p2 * (r2 + 2 * x * x) + 0.0 * r2 + 0.0 * r2 * r2;
            const double deltaY = p1 * (r2 + 2 * y * y) + 2 * p2 * x * y + 0.0 * r2 + 0.0 * r2 * r2;
            x = (x0 - deltaX) * icdist;
            y = (y0 - deltaY) * icdist;
        }
        const double xx = dfx * x + 0.0 * y + dcx, yy = 0.0 * x + dfy * y + dcy, ww = 1.0 / (0.0 * x + 0.0 * y + 1.0);
        out[2 * i] = (float)(xx * ww);
        out[2 * i + 1] = (float)(yy * ww);
    }
    return ORC_OK;
}
