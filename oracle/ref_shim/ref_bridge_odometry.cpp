// oracle/ref_shim/ref_bridge_odometry.cpp — TEST INFRASTRUCTURE ONLY.
// C entry points over the reference's own Ransac and Kabsch classes, compiled verbatim from /root/reference/Odometry/ransac.cpp and
// Odometry/kabsch.cpp by oracle/Makefile (target _ref) against the stand-ins for Eigen, PCL, boost and OpenCV in this directory:
// oracle/_ref/libodometry_ref.so.  Tests check oracle/match_ransac_oracle.cpp's restatement of rows a-11 ... a-15 against the
// reference SOURCE itself (tests/test_oracle_vs_ref_odometry.py).
//
// Shared with the oracle: the three third-party numerical routines (PCL's TransformationFromCorrespondences, Eigen's 3x3 LLT solve
// and Jacobi SVD — those libraries are not in this image, see ref_shim/Eigen/Core).  Run from the reference's source: everything else.
//
// Quirk Q7: Ransac::DepthCovariance keeps the covariance of the FIRST depth it is ever called with in a function-local
// static, i.e. once per process.  ref_depth_covariance() exposes that function, so a test can latch a chosen value first (and hand the
// same value to the oracle) or read back what a first Iterate() call latched.
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <limits>
#include <memory>
#include <mutex>
#include <set>
#include <vector>

#include "../oracle_api.h"
// every stand-in header first, so that the define below only reaches the reference's own ransac.h / kabsch.h
#include <Eigen/Core>
#include <Eigen/Eigen>
#include <boost/make_shared.hpp>
#include <opencv2/opencv.hpp>
#include <pcl/common/transformation_from_correspondences.h>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include "Core/frame.h"
#include "Core/keyframe.h"

// the reference keeps SampleMatches / DepthCovariance / the setters' fields private; the bridge reads them without touching the source
#define private public
#include "Odometry/ransac.cpp"
#include "Odometry/kabsch.cpp"
#undef private

namespace {
std::mutex g_mutex;          // libc rand() and the Q7 static are process-wide

void fill_frame(Frame& f, const float* xyz, int n)
{
    f.mvKeys3Dc.resize((size_t)n);
    for (int i = 0; i < n; ++i) f.mvKeys3Dc[(size_t)i] = cv::Point3f(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
}
}  // namespace

extern "C" {

// Ransac::DepthCovariance(depth) (ransac.cpp:416-421): the first call in the process fixes the value every later call returns
double ref_depth_covariance(double depth)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    Ransac r;
    return r.DepthCovariance(depth);
}

// form 0: Ransac(iters, minInlierTh, maxMahal, sampleSize).Iterate(pF1, pF2, m12)          (ransac.cpp:155-267)
// form 1: Ransac(pKF1, pKF2, m12) + SetParameters(...) + Iterate()                         (ransac.cpp:26-36, 44-153)
// srand(seed) first: the sample loop draws from libc rand() (ransac.cpp:269-292).  clouds (optional): mpSourceCloud / mpTargetCloud
// as x, y, z triples, n_cloud points each.
int ref_ransac_iterate(const orc_ransac_cfg* cfg, const float* src_xyz, int nsrc, const float* dst_xyz, int ndst, const orc_dmatch* m12, int nm,
    unsigned seed, int form, orc_dmatch* inliers_out, int cap, orc_ransac_out* out, float* src_cloud, float* dst_cloud, int* n_cloud)
{
    if (!cfg || !out) return ORC_ERR_ARG;
    std::lock_guard<std::mutex> lock(g_mutex);
    std::memset(out, 0, sizeof(*out));
    std::vector<cv::DMatch> matches((size_t)nm);
    for (int i = 0; i < nm; ++i) {
        if (m12[i].queryIdx < 0 || m12[i].queryIdx >= nsrc || m12[i].trainIdx < 0 || m12[i].trainIdx >= ndst) return ORC_ERR_ARG;
        matches[(size_t)i] = cv::DMatch(m12[i].queryIdx, m12[i].trainIdx, m12[i].imgIdx, m12[i].distance);
    }
    KeyFrame f1, f2;
    fill_frame(f1, src_xyz, nsrc);
    fill_frame(f2, dst_xyz, ndst);
    std::unique_ptr<Ransac> r;
    bool ok;
    srand(seed);
    if (form == 0) {
        r.reset(new Ransac(cfg->iterations, cfg->min_inlier_th, cfg->max_mahal, cfg->sample_size));
        r->CheckDepth(cfg->check_depth != 0);
        ok = r->Iterate(&f1, &f2, matches);
    } else {
        r.reset(new Ransac(&f1, &f2, matches));
        r->SetParameters(cfg->iterations, cfg->min_inlier_th, cfg->max_mahal, cfg->sample_size);
        r->CheckDepth(cfg->check_depth != 0);
        ok = r->Iterate();
    }
    out->ok = ok ? 1 : 0;
    out->rmse = r->rmse;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) out->T12[4 * i + j] = r->mT12(i, j);
    out->n_inliers = (int)r->mvInliers.size();
    out->n_good = (int)r->mpSourceCloud->points.size();
    if (n_cloud) *n_cloud = (int)r->mpSourceCloud->points.size();
    for (size_t i = 0; i < r->mpSourceCloud->points.size(); ++i) {
        const pcl::PointXYZ& a = r->mpSourceCloud->points[i]; const pcl::PointXYZ& b = r->mpTargetCloud->points[i];
        if (src_cloud) { src_cloud[3 * i] = a.x; src_cloud[3 * i + 1] = a.y; src_cloud[3 * i + 2] = a.z; }
        if (dst_cloud) { dst_cloud[3 * i] = b.x; dst_cloud[3 * i + 1] = b.y; dst_cloud[3 * i + 2] = b.z; }
    }
    if (out->n_inliers > cap) return ORC_ERR_CAPACITY;
    for (int i = 0; i < out->n_inliers && inliers_out; ++i) {
        const cv::DMatch& m = r->mvInliers[(size_t)i];
        orc_dmatch o = { m.queryIdx, m.trainIdx, m.imgIdx, m.distance };
        inliers_out[i] = o;
    }
    return ORC_OK;
}

// Ransac::SampleMatches (ransac.cpp:269-292) called `iterations` times after srand(seed) on a list of M matches whose queryIdx is
// their position: table[k][0 .. S) = the sampled positions of call k in the returned (ascending, std::set) order, -1 padded.
int ref_sample_table(unsigned seed, int M, int iterations, int sample_size, int* table)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    Ransac r(iterations, 20, 3.0f, (uint)sample_size);
    std::vector<cv::DMatch> ms((size_t)M);
    for (int i = 0; i < M; ++i) ms[(size_t)i] = cv::DMatch(i, i, 0, (float)i);
    srand(seed);
    for (int k = 0; k < iterations; ++k) {
        const std::vector<cv::DMatch> s = r.SampleMatches(ms);
        int j = 0;
        for (; j < (int)s.size() && j < sample_size; ++j) table[(size_t)k * sample_size + j] = s[(size_t)j].queryIdx;
        for (; j < sample_size; ++j) table[(size_t)k * sample_size + j] = -1;
    }
    return ORC_OK;
}

// Ransac::ComputeInliersAndError (ransac.cpp:313-348) for one transformation: returns the error, inlier positions into m12
double ref_inliers_and_error(const float* src_xyz, int nsrc, const float* dst_xyz, int ndst, const orc_dmatch* m12, int nm, const float* T16,
    float max_mahal, int* inlier_pos, int* n_inliers)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    KeyFrame f1, f2;
    fill_frame(f1, src_xyz, nsrc);
    fill_frame(f2, dst_xyz, ndst);
    Ransac r(200, 20, max_mahal, 4);
    r.mpSourceFrame = &f1; r.mpTargetFrame = &f2;
    std::vector<cv::DMatch> ms((size_t)nm), inl;
    for (int i = 0; i < nm; ++i) ms[(size_t)i] = cv::DMatch(m12[i].queryIdx, m12[i].trainIdx, i, m12[i].distance);   // imgIdx = position
    Eigen::Matrix4f T;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) T(i, j) = T16[4 * i + j];
    const double e = r.ComputeInliersAndError(ms, T, inl);
    *n_inliers = (int)inl.size();
    for (size_t i = 0; i < inl.size(); ++i) inlier_pos[i] = inl[i].imgIdx;
    return e;
}

// Kabsch::Compute(setA, setB) (kabsch.cpp:14-57): rows are points
int ref_kabsch(const float* A, const float* B, int n, float* T16)
{
    Eigen::MatrixXf a(n, 3), b(n, 3);
    for (int i = 0; i < n; ++i) for (int k = 0; k < 3; ++k) { a(i, k) = A[3 * i + k]; b(i, k) = B[3 * i + k]; }
    Kabsch kb;
    const Eigen::Matrix4f T = kb.Compute(a, b);
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) T16[4 * i + j] = T(i, j);
    return ORC_OK;
}

}  // extern "C"
