// tests/cpp/host_api_demo.cpp — drives the C++ host mirror (include/orbfront_host.hpp) exactly like the reference's
// tracking loop does (System/tracking.cpp:38-46, 193-208): Frame::ExtractFeatures -> Matcher(ratio).KnnMatch(last, cur)
// -> Ransac::Iterate(last, cur, m12).  Reads raw frames from argv[1], writes every result to argv[2] for
// tests/test_cpp_host.py to compare with the oracle.  Exit code 3 = no CUDA device (there is no CPU fallback).
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
            const bool ok = odometry.Compute(&frames[i - 1], &frames[i], m12);
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
        // Extractor(FAST, ., ADAPTIVE): the stateful grid detector over the same frames (extractor.cpp:52-77)
        Extractor adaptive(Extractor::FAST, Extractor::BRIEF, Extractor::ADAPTIVE);
        for (int i = 0; i < n; ++i) {
            std::vector<KeyPoint> kps; Mat8u none;
            adaptive.Extract(Mat8u(h, w, gray.data() + (size_t)i * w * h), Mat8u(), kps, none);
            puti(out, (int)kps.size());
            put(out, kps.data(), kps.size() * sizeof(KeyPoint));
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
