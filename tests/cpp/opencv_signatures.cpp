// tests/cpp/opencv_signatures.cpp — the reference's own call shapes against include/orbfront_host.hpp with ORBF_WITH_OPENCV
// (compiled against tests/cpp/stub/opencv2/core.hpp; with a real OpenCV the same lines compile unchanged):
//   (*mpORBextractor)(im, cv::Mat(), mvKeys, mDescriptors)            Features/orbextractor.h:36, Core/frame.cpp:137 via Extractor
//   extractor.Extract(im, cv::Mat(), keys, descriptors)               Features/extractor.h:33
//   Matcher::DescriptorDistance(d.row(i), d.row(j))                   Features/matcher.h:18
// argv[1] = raw frames (as host_api_demo), argv[2] = output.  Exit code 3 = no CUDA device.
#include <cstdio>
#include <cstdlib>
#include <vector>

#define ORBF_WITH_OPENCV
#include "orbfront_host.hpp"

int main(int argc, char** argv)
{
    if (argc < 3) return 2;
    FILE* in = fopen(argv[1], "rb");
    if (!in) return 2;
    int hdr[3];
    if (fread(hdr, sizeof(int), 3, in) != 3) return 2;
    const int w = hdr[1], h = hdr[2];
    std::vector<uint8_t> gray((size_t)w * h);
    if (fread(gray.data(), 1, gray.size(), in) != gray.size()) return 2;
    fclose(in);
    FILE* out = fopen(argv[2], "wb");
    if (!out) return 2;
    try {
        cv::Mat im(h, w, CV_8UC1, gray.data());
        orbf::ORBextractor orb(1000, 1.2f, 8, 20, 7);
        std::vector<cv::KeyPoint> keys; cv::Mat desc;
        orb(im, cv::Mat(), keys, desc);
        orbf::Extractor ex;
        std::vector<cv::KeyPoint> keys2; cv::Mat desc2;
        ex.Extract(im, cv::Mat(), keys2, desc2);
        std::vector<cv::KeyPoint> keys3;
        orb.detect(im, keys3);
        const int n = (int)keys.size();
        int same = (keys2.size() == keys.size() && keys3.size() == keys.size() && desc2.rows == desc.rows) ? 1 : 0;
        for (int i = 0; i < n && same; ++i) same = std::memcmp(desc.ptr(i), desc2.ptr(i), 32) == 0 && keys[i].pt.x == keys2[i].pt.x;
        fwrite(&n, sizeof(int), 1, out);
        fwrite(keys.data(), sizeof(cv::KeyPoint), (size_t)n, out);
        for (int i = 0; i < n; ++i) fwrite(desc.ptr(i), 1, 32, out);
        fwrite(&same, sizeof(int), 1, out);
        const cv::Mat a(1, 32, CV_8U, desc.ptr(0)), b(1, 32, CV_8U, desc.ptr(1));
        const int dist = (int)orbf::Matcher::DescriptorDistance(a, b);
        fwrite(&dist, sizeof(int), 1, out);
        cv::Mat empty; std::vector<cv::KeyPoint> untouched(3);
        orb(empty, cv::Mat(), untouched, desc2);                       // _image.empty(): outputs untouched (orbextractor.cpp:758-759)
        const int kept = (int)untouched.size();
        fwrite(&kept, sizeof(int), 1, out);
        orbf::Runtime::Shutdown();
    } catch (const orbf::Error& e) {
        fprintf(stderr, "opencv_signatures: %s\n", e.what());
        fclose(out);
        return e.status == ORBF_ERR_CUDA ? 3 : 4;
    }
    fclose(out);
    return 0;
}
