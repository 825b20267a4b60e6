// tests/cpp/tum_io_demo.cpp — include/orbfront_tum.hpp driven from a file of poses and an association file, for tests/test_tum_io.py:
//   tum_io_demo <associations.txt> <poses.bin> <trajectory_out.txt> <associations_out.txt>
// poses.bin = int32 n, n doubles (timestamps), n * 16 floats (row-major Tcw).
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <iomanip>

#include "orbfront_tum.hpp"

int main(int argc, char** argv)
{
    if (argc != 5) return 2;
    std::vector<std::string> rgb, depth;
    std::vector<double> ts;
    orbf::LoadImages(argv[1], rgb, depth, ts);
    {
        std::ofstream out(argv[4]);
        for (size_t i = 0; i < ts.size(); ++i) out << std::setprecision(17) << ts[i] << "|" << rgb[i] << "|" << depth[i] << "\n";
    }
    std::ifstream in(argv[2], std::ios::binary);
    int32_t n = 0;
    in.read(reinterpret_cast<char*>(&n), 4);
    std::vector<double> t(n);
    std::vector<float> poses((size_t)n * 16);
    in.read(reinterpret_cast<char*>(t.data()), (std::streamsize)n * 8);
    in.read(reinterpret_cast<char*>(poses.data()), (std::streamsize)n * 64);
    if (!in) return 3;
    orbf::SaveTrajectory(argv[3], t, poses.data(), (size_t)n);
    bool threw = false;
    try { orbf::LoadImages("/nonexistent/associations.txt", rgb, depth, ts); } catch (const std::runtime_error&) { threw = true; }
    return threw ? 0 : 4;
}
