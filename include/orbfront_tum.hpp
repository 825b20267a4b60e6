// include/orbfront_tum.hpp — the TUM RGB-D on-disk formats either side of the path for a C++ host, header-only,
// no dependency: the association file the reference reads and the trajectory file it writes.  Host-side file I/O only; nothing of the
// hot path's arithmetic lives here.  adaptive-rgbd-localization-mappig_b200/tum.py is the same in Python (tests, tools); the two are
// compared byte for byte by tests/test_tum_io.py.
//
//   LoadImages        Utils/utils.cpp:16-38          same name, arguments and parsing rule
//   toQuaternion      Utils/converter.cpp:149-161    = Eigen::Quaterniond(Eigen::Matrix3d), restated (Eigen is not required)
//   SaveTrajectory    System/tracking.cpp:544-580    the line format (`timestamp tx ty tz qx qy qz qw`, fixed, 6 / 9 decimals) and the
//                                                    camera-centre arithmetic of :566-571; the keyframe-graph walk above it (:556-564)
//                                                    is map bookkeeping and stays with the caller, who passes world-to-camera poses
#pragma once
#include <cmath>
#include <fstream>
#include <iomanip>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace orbf {

// Every non-empty line of the association file is `t_rgb rgb_file t_depth depth_file`; the first timestamp is the frame's.  Like the
// reference, a line is parsed whenever it is not empty — names that are missing come out as "" (operator>> on an exhausted stream).
// Compared with the reference's own function (Utils/utils.cpp compiled verbatim, oracle/_ref/ref_utils_demo) by tests/test_tum_io.py.
// Two deviations, both where the reference has no defined result: its `while (!f.eof())` never ends on a file that cannot be opened —
// this throws; on a whitespace-only line it pushes an uninitialised double as the timestamp — this pushes 0.
inline void LoadImages(const std::string& associationFilename, std::vector<std::string>& vImageFilenamesRGB,
    std::vector<std::string>& vImageFilenamesD, std::vector<double>& vTimestamps)
{
    std::ifstream in(associationFilename.c_str());
    if (!in.is_open()) throw std::runtime_error("LoadImages: cannot open " + associationFilename);
    std::string line;
    while (std::getline(in, line)) {
        if (line.empty()) continue;
        std::istringstream fields(line);
        double tRgb = 0.0, tDepth = 0.0;
        std::string rgb, depth;
        fields >> tRgb >> rgb >> tDepth >> depth;
        vTimestamps.push_back(tRgb);
        vImageFilenamesRGB.push_back(rgb);
        vImageFilenamesD.push_back(depth);
    }
}

// (x, y, z, w) of the rotation R (row-major 3x3, float): the float entries widened to double (Converter::toMatrix3d), Eigen's
// Quaterniond(Matrix3d) branch structure (trace > 0, else the largest diagonal entry), the result narrowed to float.
inline void toQuaternion(const float* R, float q[4])
{
    double m[3][3];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) m[r][c] = (double)R[3 * r + c];
    double v[4] = { 0, 0, 0, 0 };
    double t = m[0][0] + m[1][1] + m[2][2];
    if (t > 0.0) {
        t = std::sqrt(t + 1.0); v[3] = 0.5 * t; t = 0.5 / t;
        v[0] = (m[2][1] - m[1][2]) * t; v[1] = (m[0][2] - m[2][0]) * t; v[2] = (m[1][0] - m[0][1]) * t;
    } else {
        int i = 0;
        if (m[1][1] > m[0][0]) i = 1;
        if (m[2][2] > m[i][i]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = std::sqrt(m[i][i] - m[j][j] - m[k][k] + 1.0); v[i] = 0.5 * t; t = 0.5 / t;
        v[3] = (m[k][j] - m[j][k]) * t; v[j] = (m[j][i] + m[i][j]) * t; v[k] = (m[k][i] + m[i][k]) * t;
    }
    for (int a = 0; a < 4; ++a) q[a] = (float)v[a];
}

// One line of the trajectory file from a world-to-camera pose Tcw (row-major 4x4): Rwc = Rcw^T, twc = -Rwc * tcw with cv::Mat's float
// product (three products summed left to right), q = toQuaternion(Rwc).
inline std::string TrajectoryLine(double timestamp, const float* Tcw)
{
    float Rwc[9], twc[3], q[4];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) Rwc[3 * r + c] = Tcw[4 * c + r];
    for (int r = 0; r < 3; ++r) {
        volatile float acc = Rwc[3 * r] * Tcw[3];                 // volatile: no contraction into FMA whatever the caller's flags
        volatile float p1 = Rwc[3 * r + 1] * Tcw[7];
        acc = acc + p1;
        volatile float p2 = Rwc[3 * r + 2] * Tcw[11];
        acc = acc + p2;
        twc[r] = -acc;
    }
    toQuaternion(Rwc, q);
    std::ostringstream out;
    out << std::fixed << std::setprecision(6) << timestamp << std::setprecision(9) << " " << twc[0] << " " << twc[1] << " " << twc[2] << " " << q[0] << " "
        << q[1] << " " << q[2] << " " << q[3];
    return out.str();
}

// The frame trajectory in the format evaluate_ate.py / evaluate_rpe.py read: one line per frame, poses = n row-major 4x4 Tcw back to back
// (what orbf_compose_trajectory / orbf::SequenceShard::compose return).
inline void SaveTrajectory(const std::string& filename, const std::vector<double>& vTimestamps, const float* poses, size_t n)
{
    if (vTimestamps.size() < n) throw std::invalid_argument("SaveTrajectory: fewer timestamps than poses");
    std::ofstream f(filename.c_str());
    if (!f.is_open()) throw std::runtime_error("SaveTrajectory: cannot open " + filename);
    for (size_t i = 0; i < n; ++i) f << TrajectoryLine(vTimestamps[i], poses + 16 * i) << "\n";
}

}  // namespace orbf
