// tests/cpp/host_api_demo.cpp — drives the C++ host mirror (include/orbfront_host.hpp) exactly like the reference's
// tracking loop does (System/tracking.cpp:38-46, 193-208): Frame::ExtractFeatures -> Matcher(ratio).KnnMatch(last, cur)
// -> Ransac::Iterate(last, cur, m12).  Reads raw frames from argv[1], writes every result to argv[2] for
// tests/test_cpp_host.py to compare with the oracle.  Exit code 3 = no CUDA device (there is no CPU fallback).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "orbfront_host.hpp"

using namespace orbf;

static void put(FILE* f, const void* p, size_t n) { if (n && fwrite(p, 1, n, f) != n) { perror("write"); exit(2); } }
static void puti(FILE* f, int v) { put(f, &v, sizeof(v)); }

int main(int argc, char** argv)
{
    if (argc < 3) { fprintf(stderr, "usage: %s <in.raw> <out.raw>\n", argv[0]); return 2; }
    FILE* in = fopen(argv[1], "rb");
    if (!in) { perror(argv[1]); return 2; }
    int hdr[3];
    if (fread(hdr, sizeof(int), 3, in) != 3) return 2;
    const int n = hdr[0], w = hdr[1], h = hdr[2];
    std::vector<uint8_t> gray((size_t)n * w * h);
    std::vector<uint16_t> depth((size_t)n * w * h);
    if (fread(gray.data(), 1, gray.size(), in) != gray.size() || fread(depth.data(), 2, depth.size(), in) != depth.size()) return 2;
    fclose(in);
    FILE* out = fopen(argv[2], "wb");
    if (!out) { perror(argv[2]); return 2; }
    try {
        Extractor extractor(Extractor::ORB_SLAM2, Extractor::ORB_SLAM2, Extractor::NORMAL);
        Odometry odometry(Odometry::RANSAC);
        Ransac::Seed() = 42;
        std::vector<Frame> frames;
        for (int i = 0; i < n; ++i) {
            Frame f(Mat8u(h, w, gray.data() + (size_t)i * w * h), Mat16u(h, w, depth.data() + (size_t)i * w * h), (double)i);
            f.ExtractFeatures(&extractor);
            puti(out, (int)f.N);
            put(out, f.mvKeys.data(), f.N * sizeof(KeyPoint));
            put(out, f.mDescriptors.data, f.N * 32);
            put(out, f.mvKeys3Dc.data(), f.N * sizeof(Point3f));
            frames.push_back(f);
            if (i == 0) continue;
            std::vector<DMatch> m12;
            Matcher matcher(0.8f);
            matcher.KnnMatch(frames[i - 1], frames[i], m12, /*crossCheck=*/true);
            odometry.Compute(&frames[i - 1], &frames[i], m12);
            const bool ok = odometry.mbConverged;
            puti(out, (int)m12.size());
            put(out, m12.data(), m12.size() * sizeof(DMatch));
            puti(out, ok ? 1 : 0);
            put(out, &odometry.ransac()->rmse, sizeof(float));
            put(out, odometry.mT12.m, sizeof(float) * 16);
            puti(out, (int)odometry.mvInliers.size());
            put(out, odometry.mvInliers.data(), odometry.mvInliers.size() * sizeof(DMatch));
        }
        // Matcher::DescriptorDistance and Kabsch::Compute on the first frame's data
        const Mat8u a(1, 32, frames[0].mDescriptors.ptr(0)), b(1, 32, frames[0].mDescriptors.ptr(1));
        puti(out, (int)Matcher::DescriptorDistance(a, b));
        std::vector<Point3f> A(frames[0].mvKeys3Dc.begin(), frames[0].mvKeys3Dc.begin() + 50), B = A;
        for (auto& p : B) { const float x = p.x; p.x = -p.y + 0.1f; p.y = x - 0.2f; p.z += 0.3f; }      // 90 deg about z + translation
        const Matrix4f T = Kabsch().Compute(A, B);
        put(out, T.m, sizeof(float) * 16);
        {   // Kabsch::Compute(MatrixXf, MatrixXf): the reference's signature (kabsch.h:10), rows = points
            MatrixXf MA((int)A.size(), 3), MB((int)B.size(), 3);
            for (int i = 0; i < (int)A.size(); ++i) { MA(i, 0) = A[i].x; MA(i, 1) = A[i].y; MA(i, 2) = A[i].z; MB(i, 0) = B[i].x; MB(i, 1) = B[i].y; MB(i, 2) = B[i].z; }
            const Matrix4f T2 = Kabsch().Compute(MA, MB);
            put(out, T2.m, sizeof(float) * 16);
        }
        // poses after Odometry::Compute's composition rule (odometry.cpp:82-86); frame 0 keeps the identity
        for (int i = 0; i < n; ++i) { const Matrix4f P = frames[i].GetPose(); put(out, P.m, sizeof(float) * 16); }
        // ORBextractor::mvImagePyramid[2] of the last frame extracted (orbextractor.h:57)
        {
            const Mat8u lvl = extractor.detector()->mvImagePyramid[2];
            puti(out, (int)extractor.detector()->mvImagePyramid.size()); puti(out, lvl.cols); puti(out, lvl.rows);
            put(out, lvl.data, (size_t)lvl.cols * lvl.rows);
        }
        if (n >= 2) {
            // Matcher::KnnMatch(KeyFrame*, Frame&, .) (matcher.cpp:23-53): landmarks at the even features of the keyframe, every fifth of
            // them bad, one feature of the frame already taken
            KeyFrame kf0(frames[0]), kf1(frames[1]);
            for (size_t j = 0; j < kf0.N; j += 2) kf0.AddLandmark(reinterpret_cast<Landmark*>((uintptr_t)(j + 1)), j);
            Frame f2 = frames[1];
            if (f2.N > 7) f2.AddLandmark(reinterpret_cast<Landmark*>((uintptr_t)0x7fff), 7);
            std::vector<DMatch> mk;
            Matcher(0.8f).KnnMatch(&kf0, f2, mk, [](Landmark* p) { return ((uintptr_t)p - 1) % 10 == 0; });
            puti(out, (int)mk.size());
            put(out, mk.data(), mk.size() * sizeof(DMatch));
            int nOut = 0;
            for (size_t j = 0; j < f2.N; ++j) nOut += f2.IsOutlier(j) ? 1 : 0;
            puti(out, nOut);
            // Ransac(KeyFrame*, KeyFrame*, matches).Iterate() (ransac.cpp:26-36,44-153) and the clouds it leaves behind
            std::vector<DMatch> m01;
            Matcher(0.8f).KnnMatch(frames[0], frames[1], m01, /*crossCheck=*/true);
            Ransac::Seed() = 100;
            Ransac bound(&kf0, &kf1, m01);
            const bool ok2 = bound.Iterate();
            puti(out, ok2 ? 1 : 0);
            put(out, bound.mT12.m, sizeof(float) * 16);
            puti(out, (int)bound.mvInliers.size());
            put(out, bound.mvInliers.data(), bound.mvInliers.size() * sizeof(DMatch));
            puti(out, (int)bound.mpSourceCloud->points.size());
            put(out, bound.mpSourceCloud->points.data(), bound.mpSourceCloud->points.size() * sizeof(PointXYZ));
            put(out, bound.mpTargetCloud->points.data(), bound.mpTargetCloud->points.size() * sizeof(PointXYZ));
            // too few matches: the early return leaves everything cleared
            std::vector<DMatch> few(m01.begin(), m01.begin() + std::min<size_t>(5, m01.size()));
            Ransac r3;
            const bool ok3 = r3.Iterate(&frames[0], &frames[1], few);
            puti(out, ok3 ? 1 : 0); puti(out, (int)r3.mpSourceCloud->points.size()); puti(out, (int)r3.mvInliers.size());
        }
        // Extractor(FAST, ., ADAPTIVE): the stateful grid detector over the same frames (extractor.cpp:52-77)
        Extractor adaptive(Extractor::FAST, Extractor::BRIEF, Extractor::ADAPTIVE);
        for (int i = 0; i < n; ++i) {
            std::vector<KeyPoint> kps; Mat8u none;
            adaptive.Extract(Mat8u(h, w, gray.data() + (size_t)i * w * h), Mat8u(), kps, none);
            puti(out, (int)kps.size());
            put(out, kps.data(), kps.size() * sizeof(KeyPoint));
        }
        {   // Frame::ExtractFeatures on the adaptive route: the detector's keypoints unprojected on the device (frame.cpp:138-164)
            Extractor adaptive2(Extractor::FAST, Extractor::BRIEF, Extractor::ADAPTIVE);
            Frame fa(Mat8u(h, w, gray.data()), Mat16u(h, w, depth.data()), 0.0);
            fa.ExtractFeatures(&adaptive2);
            puti(out, (int)fa.N);
            put(out, fa.mvKeys.data(), fa.N * sizeof(KeyPoint));
            put(out, fa.mvKeys3Dc.data(), fa.N * sizeof(Point3f));
            put(out, fa.mvuRight.data(), fa.N * sizeof(float));
        }
        put(out, adaptive.AdaptiveThresholds().data(), 9 * sizeof(double));
        Runtime::Shutdown();
    } catch (const Error& e) {
        fprintf(stderr, "host_api_demo: %s\n", e.what());
        fclose(out);
        return e.status == ORBF_ERR_CUDA ? 3 : 4;
    }
    fclose(out);
    return 0;
}
