// tools/latency_cpp.cpp — the reference's per-frame tracking loop (System/tracking.cpp:38-46,193-208) through the C++ host mirror
// (include/orbfront_host.hpp), ONE frame at a time from pageable host memory, timed per call with std::chrono the way
// Tests/detector-descriptor-speed-test.cpp:53-70 times its extractor (TickMeter around the call, mean over the frames):
//   Frame f(gray, depth); f.ExtractFeatures(&extractor);  Matcher(0.8f).KnnMatch(last, f, m12, cross);  odometry.Compute(&last, &f, m12);
// argv[1] = raw frames (int n, w, h; n gray planes; n u16 depth planes), prints one JSON object.  Built and run by bench.py.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "orbfront_host.hpp"

using namespace orbf;
using Clock = std::chrono::steady_clock;

static double pct(std::vector<double> v, double p)
{
    if (v.empty()) return 0.0;
    std::sort(v.begin(), v.end());
    const double pos = p / 100.0 * (double)(v.size() - 1);
    const size_t i = (size_t)pos;
    return i + 1 < v.size() ? v[i] + (pos - (double)i) * (v[i + 1] - v[i]) : v[i];
}

int main(int argc, char** argv)
{
    if (argc < 2) return 2;
    FILE* in = fopen(argv[1], "rb");
    if (!in) return 2;
    int hdr[3];
    if (fread(hdr, sizeof(int), 3, in) != 3) return 2;
    const int n = hdr[0], w = hdr[1], h = hdr[2];
    std::vector<uint8_t> gray((size_t)n * w * h);
    std::vector<uint16_t> depth((size_t)n * w * h);
    if (fread(gray.data(), 1, gray.size(), in) != gray.size() || fread(depth.data(), 2, depth.size(), in) != depth.size()) return 2;
    fclose(in);
    try {
        Extractor extractor(Extractor::ORB_SLAM2, Extractor::ORB_SLAM2, Extractor::NORMAL);
        Odometry odometry(Odometry::RANSAC);
        Ransac::Seed() = 42;
        std::vector<double> tE, tM, tR, tT;
        Frame last;
        long inliers = 0, matches = 0;
        for (int i = 0; i < n; ++i) {
            const auto t0 = Clock::now();
            Frame f(Mat8u(h, w, gray.data() + (size_t)i * w * h), Mat16u(h, w, depth.data() + (size_t)i * w * h), (double)i);
            f.ExtractFeatures(&extractor);
            const auto t1 = Clock::now();
            if (i > 0) {
                std::vector<DMatch> m12;
                Matcher matcher(0.8f);
                matcher.KnnMatch(last, f, m12, /*crossCheck=*/true);
                const auto t2 = Clock::now();
                odometry.Compute(&last, &f, m12);
                const auto t3 = Clock::now();
                if (i >= 8) {                                // the first calls pay allocation / module load
                    const auto ms = [](Clock::time_point a, Clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
                    tE.push_back(ms(t0, t1)); tM.push_back(ms(t1, t2)); tR.push_back(ms(t2, t3)); tT.push_back(ms(t0, t3));
                    inliers += (long)odometry.mvInliers.size(); matches += (long)m12.size();
                }
            }
            last = f;
        }
        double mean = 0; for (double v : tT) mean += v;
        mean = tT.empty() ? 0.0 : mean / (double)tT.size();
        printf("{\"frames\": %d, \"p50\": %.6f, \"p99\": %.6f, \"mean\": %.6f, \"p50_by_call\": {\"extract\": %.6f, \"match\": %.6f, \"ransac\": %.6f}, "
               "\"mean_matches\": %.2f, \"mean_inliers\": %.2f}\n", (int)tT.size(), pct(tT, 50), pct(tT, 99), mean, pct(tE, 50), pct(tM, 50), pct(tR, 50),
            tT.empty() ? 0.0 : (double)matches / (double)tT.size(), tT.empty() ? 0.0 : (double)inliers / (double)tT.size());
        Runtime::Shutdown();
    } catch (const Error& e) {
        fprintf(stderr, "latency_cpp: %s\n", e.what());
        return e.status == ORBF_ERR_CUDA ? 3 : 4;
    }
    return 0;
}
