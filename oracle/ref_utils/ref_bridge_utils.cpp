// oracle/ref_utils/ref_bridge_utils.cpp — drives the reference's own LoadImages (Utils/utils.cpp:16-38, compiled verbatim from where it
// lies, never copied) for tests/test_tum_io.py: ref_utils_demo <associations.txt> <out.txt> writes one `timestamp|rgb|depth` line per
// entry, the format tests/cpp/tum_io_demo.cpp writes for include/orbfront_tum.hpp.  Test infrastructure (oracle/_ref).
#include <fstream>
#include <iomanip>
#include <string>
#include <vector>

void LoadImages(const std::string& associationFilename, std::vector<std::string>& vImageFilenamesRGB, std::vector<std::string>& vImageFilenamesD,
    std::vector<double>& vTimestamps);

int main(int argc, char** argv)
{
    if (argc != 3) return 2;
    std::vector<std::string> rgb, depth;
    std::vector<double> ts;
    LoadImages(argv[1], rgb, depth, ts);
    std::ofstream out(argv[2]);
    for (size_t i = 0; i < ts.size(); ++i) out << std::setprecision(17) << ts[i] << "|" << rgb[i] << "|" << depth[i] << "\n";
    return 0;
}
