// tests/cpp/eigen_signatures.cpp — Kabsch::Compute with the reference's own signature (Odometry/kabsch.h:10) under ORBF_WITH_EIGEN,
// compiled against the stand-in tests/cpp/stub/Eigen/Core (with a real Eigen the same lines compile unchanged).
// Prints the 16 floats of the result for a fixed point set; exit code 3 = no CUDA device.
#include <cstdio>

#define ORBF_WITH_EIGEN
#include "orbfront_host.hpp"

int main()
{
    try {
        const int n = 40;
        Eigen::MatrixXf A(n, 3), B(n, 3);
        for (int i = 0; i < n; ++i) {
            const float x = 0.1f * (float)i, y = 0.05f * (float)(i * i % 17), z = 1.0f + 0.02f * (float)i;
            A(i, 0) = x; A(i, 1) = y; A(i, 2) = z;
            B(i, 0) = -y + 0.1f; B(i, 1) = x - 0.2f; B(i, 2) = z + 0.3f;            // 90 degrees about z, then a translation
        }
        const Eigen::Matrix4f T = orbf::Kabsch().Compute(A, B);
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) printf("%.9g ", T(i, j));
        printf("\n");
        orbf::Runtime::Shutdown();
    } catch (const orbf::Error& e) {
        fprintf(stderr, "eigen_signatures: %s\n", e.what());
        return e.status == ORBF_ERR_CUDA ? 3 : 4;
    }
    return 0;
}
