// =====================================================================================
// oracle/match_ransac_oracle.cpp — TEST INFRASTRUCTURE ONLY (parity checker, never shipped)
//
// CPU restatement of
//   /root/reference/Features/matcher.cpp:10-88,355-358   (BFMatcher knnMatch k=2 + Lowe ratio)
//   /root/reference/Odometry/ransac.cpp:155-431          (Iterate, SampleMatches,
//        GetTransformFromMatches, ComputeInliersAndError, ErrorFunction2, DepthCovariance)
//   /root/reference/Odometry/kabsch.cpp:14-57            (unweighted Kabsch)
// Third-party arithmetic absent from /root/reference (unpinned versions):
//   cv::BFMatcher(NORM_HAMMING).knnMatch  — order (distance asc, trainIdx asc); pinned vs cv2 4.13.
//   pcl::TransformationFromCorrespondences (PCL common/impl/transformation_from_correspondences.hpp)
//        — restated from the published algorithm (incremental weighted mean/covariance, 3x3 SVD,
//        R = U diag(1,1,sign(det U det V)) V^T).  NOT probe-verified (PCL absent): pose parity
//        against a real reference build is tolerance-level (1e-5), "parity unpinned".
//   Eigen JacobiSVD<Matrix3f>, Matrix3d::llt().solve — restated as a two-sided Jacobi 3x3 SVD and
//        an unblocked 3x3 Cholesky; checked against numpy in tests/test_oracle_ransac.py.
// Float semantics: no FMA contraction (build with -ffp-contract=off), strictly left-to-right sums.
// Pinned (round 2): the reference-AUTHORED logic restated here is bit-identical to the reference's own ransac.cpp / kabsch.cpp /
// matcher.cpp compiled verbatim into oracle/_ref (tests/test_oracle_vs_ref_odometry.py, test_oracle_vs_ref_matcher.py); the PCL / Eigen
// arithmetic named above is shared with that build (orc_tfc_transform, orc_llt3_solve, orc_svd3) and stays "parity unpinned".
// =====================================================================================
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <mutex>
#include <set>
#include <vector>

#include "oracle_api.h"

namespace {

// Sensitivity probe (tests/test_oracle_ransac.py): Eigen evaluates small fixed-size inner products as a balanced tree
// (a0 + (a1 + a2), (a0 + a1) + (a2 + a3)) when it unrolls a reduction, and left to right otherwise; which one a given build of the
// reference uses is not knowable here (the library is absent, section 7 of DESIGN.md).  g_sumTree = 1 switches every 3- / 4-term sum of
// the PCL / Eigen restatements to the tree order so that the effect on inlier sets and poses can be measured.  Default 0: left to right,
// the order the CUDA kernels implement.
int g_sumTree = 0;
template <typename T> inline T sum3(T a, T b, T c) { return g_sumTree ? a + (b + c) : (a + b) + c; }
template <typename T> inline T sum4(T a, T b, T c, T d) { return g_sumTree ? (a + b) + (c + d) : ((a + b) + c) + d; }

inline int popc64(uint64_t v) { return __builtin_popcountll(v); }

inline int hamming256(const uint8_t* a, const uint8_t* b)
{
    uint64_t x[4], y[4];
    std::memcpy(x, a, 32); std::memcpy(y, b, 32);
    return popc64(x[0] ^ y[0]) + popc64(x[1] ^ y[1]) + popc64(x[2] ^ y[2]) + popc64(x[3] ^ y[3]);
}

// ---- 3x3 SVD: two-sided (Kogbetliantz) Jacobi in f32, A = U diag(S) V^T, S sorted descending ----
struct M3 { float m[3][3]; };

inline void rot_rows(M3& a, int p, int q, float c, float s)
{
    for (int j = 0; j < 3; ++j) {
        const float x = a.m[p][j], y = a.m[q][j];
        a.m[p][j] = c * x + s * y;
        a.m[q][j] = c * y - s * x;
    }
}
inline void rot_cols(M3& a, int p, int q, float c, float s)
{
    for (int i = 0; i < 3; ++i) {
        const float x = a.m[i][p], y = a.m[i][q];
        a.m[i][p] = c * x - s * y;
        a.m[i][q] = s * x + c * y;
    }
}

void svd3(const float* A, M3& U, float S[3], M3& V)
{
    M3 M;
    float scale = 0.f;
    for (int i = 0; i < 9; ++i) scale = std::max(scale, std::fabs(A[i]));
    if (!(scale > 0.f)) scale = 1.f;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
        M.m[i][j] = A[3 * i + j] / scale;
        U.m[i][j] = V.m[i][j] = (i == j) ? 1.f : 0.f;
    }
    const float precision = 2.f * FLT_EPSILON, tiny = FLT_MIN;
    float maxDiag = std::max(std::fabs(M.m[0][0]), std::max(std::fabs(M.m[1][1]), std::fabs(M.m[2][2])));
    for (int sweep = 0; sweep < 30; ++sweep) {
        bool finished = true;
        for (int p = 1; p < 3; ++p)
            for (int q = 0; q < p; ++q) {
                const float thr = std::max(tiny, precision * maxDiag);
                if (!(std::fabs(M.m[p][q]) > thr || std::fabs(M.m[q][p]) > thr)) continue;
                finished = false;
                // 2x2 block in (p,q) index order
                const float m00 = M.m[p][p], m01 = M.m[p][q], m10 = M.m[q][p], m11 = M.m[q][q];
                // step 1: row rotation (c1,s1) that symmetrises the block
                float c1 = 1.f, s1 = 0.f;
                const float t = m00 + m11, d = m10 - m01;
                if (std::fabs(d) >= tiny) {
                    const float u = t / d;
                    const float tmp = std::sqrt(1.f + u * u);
                    s1 = 1.f / tmp;
                    c1 = u / tmp;
                }
                const float x = c1 * m00 + s1 * m10;   // symmetric block [[x,y],[y,z]]
                const float y = c1 * m01 + s1 * m11;
                const float z = c1 * m11 - s1 * m01;
                // step 2: Jacobi rotation (c2,s2) diagonalising it
                float c2 = 1.f, s2 = 0.f;
                if (std::fabs(y) >= tiny) {
                    const float tau = (x - z) / (2.f * y);
                    const float w = std::sqrt(tau * tau + 1.f);
                    const float tt = (tau > 0.f) ? -1.f / (tau + w) : -1.f / (tau - w);
                    c2 = 1.f / std::sqrt(tt * tt + 1.f);
                    s2 = tt * c2;
                }
                const float cL = c1 * c2 + s1 * s2, sL = s1 * c2 - c1 * s2;
                rot_rows(M, p, q, cL, sL);
                rot_cols(M, p, q, c2, s2);
                rot_cols(U, p, q, cL, -sL);
                rot_cols(V, p, q, c2, s2);
                maxDiag = std::max(maxDiag, std::max(std::fabs(M.m[p][p]), std::fabs(M.m[q][q])));
            }
        if (finished) break;
    }
    for (int i = 0; i < 3; ++i) {
        const float a = M.m[i][i];
        S[i] = std::fabs(a);
        if (a < 0.f) for (int r = 0; r < 3; ++r) U.m[r][i] = -U.m[r][i];
    }
    for (int i = 0; i < 3; ++i) {          // selection sort, descending, first max wins
        int k = i;
        for (int j = i + 1; j < 3; ++j) if (S[j] > S[k]) k = j;
        if (k != i) {
            std::swap(S[i], S[k]);
            for (int r = 0; r < 3; ++r) { std::swap(U.m[r][i], U.m[r][k]); std::swap(V.m[r][i], V.m[r][k]); }
        }
    }
    for (int i = 0; i < 3; ++i) S[i] *= scale;
}

inline float det3(const M3& a)
{
    return a.m[0][0] * (a.m[1][1] * a.m[2][2] - a.m[1][2] * a.m[2][1])
        - a.m[0][1] * (a.m[1][0] * a.m[2][2] - a.m[1][2] * a.m[2][0])
        + a.m[0][2] * (a.m[1][0] * a.m[2][1] - a.m[1][1] * a.m[2][0]);
}

// pcl::TransformationFromCorrespondences restated: incremental weighted statistics + SVD.
struct Tfc {
    float accW = 0.f;
    float m1[3] = { 0, 0, 0 }, m2[3] = { 0, 0, 0 };
    float C[3][3] = { { 0, 0, 0 }, { 0, 0, 0 }, { 0, 0, 0 } };
    void add(const float* p, const float* q, float w)
    {
        if (w == 0.0f) return;
        accW += w;
        const float alpha = w / accW;
        float d1[3], d2[3];
        for (int i = 0; i < 3; ++i) { d1[i] = p[i] - m1[i]; d2[i] = q[i] - m2[i]; }
        const float oma = 1.0f - alpha;
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) {
                const float outer = d2[r] * d1[c];
                C[r][c] = oma * (C[r][c] + alpha * outer);
            }
        for (int i = 0; i < 3; ++i) { m1[i] += alpha * d1[i]; m2[i] += alpha * d2[i]; }
    }
    void transform(float* T) const
    {
        M3 U, V; float S[3];
        svd3(&C[0][0], U, S, V);
        const float s22 = (det3(U) * det3(V) < 0.0f) ? -1.0f : 1.0f;
        float R[3][3];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                const float us = U.m[i][2] * s22;
                R[i][j] = sum3(U.m[i][0] * V.m[j][0], U.m[i][1] * V.m[j][1], us * V.m[j][2]);
            }
        for (int i = 0; i < 3; ++i) {
            const float rm = sum3(R[i][0] * m1[0], R[i][1] * m1[1], R[i][2] * m1[2]);
            T[4 * i + 0] = R[i][0]; T[4 * i + 1] = R[i][1]; T[4 * i + 2] = R[i][2];
            T[4 * i + 3] = m2[i] - rm;
        }
        T[12] = 0; T[13] = 0; T[14] = 0; T[15] = 1;
    }
};

// ErrorFunction2 (ransac.cpp:350-414) with the static-local depth covariance made explicit (Q7).
struct MahalConst { double cov_x, cov_y; };
MahalConst mahal_const()
{
    const double ax = 58.0 / 180.0 * M_PI, ay = 45.0 / 180.0 * M_PI;
    const double sx = 3 * std::tan(ax / 640), sy = 3 * std::tan(ay / 480);
    return { sx * sx, sy * sy };
}

// Eigen's `S.llt().solve(b)` for a 3x3 double matrix, restated: unblocked Cholesky S = L L^T (lower), then L y = b, L^T x = y.
// false: S is not positive definite.
bool llt3_solve(const double S[3][3], const double b[3], double xs[3])
{
    double L[3][3] = { { 0 } };
    for (int k = 0; k < 3; ++k) {
        double x = S[k][k];
        for (int j = 0; j < k; ++j) x -= L[k][j] * L[k][j];
        if (!(x > 0.0)) return false;
        const double lkk = std::sqrt(x);
        L[k][k] = lkk;
        for (int i = k + 1; i < 3; ++i) {
            double v = S[i][k];
            for (int j = 0; j < k; ++j) v -= L[i][j] * L[k][j];
            L[i][k] = v / lkk;
        }
    }
    double y[3];
    y[0] = b[0] / L[0][0];
    y[1] = (b[1] - L[1][0] * y[0]) / L[1][1];
    y[2] = ((b[2] - L[2][0] * y[0]) - L[2][1] * y[1]) / L[2][2];
    xs[2] = y[2] / L[2][2];
    xs[1] = (y[1] - L[2][1] * xs[2]) / L[1][1];
    xs[0] = ((y[0] - L[1][0] * xs[1]) - L[2][0] * xs[2]) / L[0][0];
    return true;
}

double mahal2(const float* x1, const float* x2, const double T[16], double cz)
{
    static const MahalConst K = mahal_const();
    const double dmax = std::numeric_limits<double>::max();
    if (std::isnan(x1[2]) || std::isnan(x2[2])) return dmax;
    const double a[3] = { x1[0], x1[1], x1[2] }, b[3] = { x2[0], x2[1], x2[2] };
    double mu12[3], dl[3];
    for (int i = 0; i < 3; ++i) mu12[i] = sum4(T[4 * i] * a[0], T[4 * i + 1] * a[1], T[4 * i + 2] * a[2], T[4 * i + 3]);
    for (int i = 0; i < 3; ++i) dl[i] = mu12[i] - b[i];
    {
        const double sq = sum3(dl[0] * dl[0], dl[1] * dl[1], dl[2] * dl[2]);
        const double s1 = std::max(K.cov_x, cz), s2 = std::max(K.cov_x, cz);
        if (sq > 2.0 * (s1 + s2)) return dmax;
    }
    const double c1[3] = { K.cov_x * a[2], K.cov_y * a[2], cz };
    const double c2[3] = { K.cov_x * b[2], K.cov_y * b[2], cz };
    // S = R^T diag(c1) R + diag(c2)   (literal: rotation_mat.transpose() * cov1 * rotation_mat)
    double S[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            const double v = sum3((T[4 * 0 + i] * c1[0]) * T[4 * 0 + j], (T[4 * 1 + i] * c1[1]) * T[4 * 1 + j], (T[4 * 2 + i] * c1[2]) * T[4 * 2 + j]);
            S[i][j] = v + ((i == j) ? c2[i] : 0.0);
        }
    if (std::isnan(dl[2])) return dmax;
    double xs[3];
    if (!llt3_solve(S, dl, xs)) return dmax;                  // not positive definite (degenerate depth)
    const double d2 = sum3(dl[0] * xs[0], dl[1] * xs[1], dl[2] * xs[2]);
    if (!(d2 >= 0.0)) return dmax;
    return d2;
}

struct Ctx {
    const float* src; const float* dst;
    double cz; bool czLatched;
    float maxMahal;
};

void transform_from(const Ctx& c, const std::vector<orc_dmatch>& ms, float* T)
{
    Tfc tfc;
    for (const auto& m : ms) {
        const float* from = c.src + 3 * m.queryIdx;
        const float* to = c.dst + 3 * m.trainIdx;
        if (std::isnan(from[2]) || std::isnan(to[2])) continue;
        const float w = 1.0f / (from[2] * to[2]);
        tfc.add(from, to, w);
    }
    tfc.transform(T);
}

double inliers_and_error(Ctx& c, const std::vector<orc_dmatch>& all, const float* T4f, std::vector<orc_dmatch>& inl)
{
    inl.clear();
    double mean = 0.0;
    double T[16];
    for (int i = 0; i < 16; ++i) T[i] = (double)T4f[i];
    const double thr = (double)(c.maxMahal * c.maxMahal);
    for (const auto& m : all) {
        const float* o = c.src + 3 * m.queryIdx;
        const float* t = c.dst + 3 * m.trainIdx;
        if (o[2] == 0.0f || t[0] == 0.0f) continue;              // sic: target.x (quirk Q8)
        if (!c.czLatched && !std::isnan(o[2]) && !std::isnan(t[2])) {  // Q7: first call of DepthCovariance
            const double sd = 0.01 * (double)o[2] * (double)o[2];
            c.cz = sd * sd; c.czLatched = true;
        }
        const double d = mahal2(o, t, T, c.cz);
        if (d > thr) continue;
        if (!(d >= 0.0)) continue;
        mean += d;
        inl.push_back(m);
    }
    if (inl.size() < 3) mean = 1e9;
    else { mean /= inl.size(); mean = std::sqrt(mean); }
    return mean;
}

void sample_libc(int M, int S, int* row)
{
    std::set<size_t> ids;
    int safety = 0;
    while ((int)ids.size() < S && M >= S) {
        int id1 = rand() % M;
        int id2 = rand() % M;
        if (id1 > id2) id1 = id2;
        ids.insert(id1);
        if (++safety > 10000) break;
    }
    int k = 0;
    for (size_t id : ids) row[k++] = (int)id;
    for (; k < S; ++k) row[k] = -1;
}

}  // namespace

extern "C" {

int orc_hamming(const uint8_t* a, const uint8_t* b) { return hamming256(a, b); }

int orc_knn2(const uint8_t* q, int nq, const uint8_t* t, int nt, int* idx1, int* d1, int* idx2, int* d2)
{
    for (int i = 0; i < nq; ++i) {
        int b1 = -1, b2 = -1, e1 = 1 << 30, e2 = 1 << 30;
        const uint8_t* a = q + (size_t)i * 32;
        for (int j = 0; j < nt; ++j) {
            const int d = hamming256(a, t + (size_t)j * 32);
            if (d < e1) { e2 = e1; b2 = b1; e1 = d; b1 = j; }
            else if (d < e2) { e2 = d; b2 = j; }
        }
        idx1[i] = b1; d1[i] = (b1 < 0) ? -1 : e1;
        idx2[i] = b2; d2[i] = (b2 < 0) ? -1 : e2;
    }
    return ORC_OK;
}

// Matcher::KnnMatch's device-side part: kNN-2 + ratio (matcher.cpp:55-66), optional mutual-NN
// cross-check (north-star extension, quirk Q10; default off in the reference).
int orc_knn_match(const uint8_t* q, int nq, const uint8_t* t, int nt, float ratio, int cross_check, orc_dmatch* out,
    int cap, int* n)
{
    *n = 0;
    if (nq <= 0 || nt < 2) return ORC_OK;   // reference indexes matchesKnn[i][1]: needs >= 2 train rows
    std::vector<int> i1(nq), e1(nq), i2(nq), e2(nq);
    orc_knn2(q, nq, t, nt, i1.data(), e1.data(), i2.data(), e2.data());
    std::vector<int> r1, rd1, r2, rd2;
    if (cross_check) {
        r1.resize(nt); rd1.resize(nt); r2.resize(nt); rd2.resize(nt);
        orc_knn2(t, nt, q, nq, r1.data(), rd1.data(), r2.data(), rd2.data());
    }
    int k = 0;
    for (int i = 0; i < nq; ++i) {
        const float da = (float)e1[i], db = (float)e2[i];
        if (!(da < ratio * db)) continue;
        if (cross_check && r1[i1[i]] != i) continue;
        if (k >= cap) { *n = k; return ORC_ERR_CAPACITY; }
        out[k].queryIdx = i; out[k].trainIdx = i1[i]; out[k].imgIdx = 0; out[k].distance = da;
        ++k;
    }
    *n = k;
    return ORC_OK;
}

// srand() / rand() are the real libc calls (quirk Q5) and share one process-wide state: callers on several threads (bench.py's CPU
// legs run the oracle frame-parallel) are serialised here, so that every table is the uninterrupted stream of its own seed.
static std::mutex g_randMutex;

int orc_libc_rand_sequence(unsigned seed, int n, int* out)
{
    std::lock_guard<std::mutex> lock(g_randMutex);
    srand(seed);
    for (int i = 0; i < n; ++i) out[i] = rand();
    return ORC_OK;
}

int orc_sample_table_libc(unsigned seed, int M, int iterations, int sample_size, int* table)
{
    std::lock_guard<std::mutex> lock(g_randMutex);
    srand(seed);
    for (int k = 0; k < iterations; ++k) sample_libc(M, sample_size, table + (size_t)k * sample_size);
    return ORC_OK;
}

int orc_std_sort_dmatch(orc_dmatch* m, int n)
{
    std::sort(m, m + n, [](const orc_dmatch& a, const orc_dmatch& b) { return a.distance < b.distance; });
    return ORC_OK;
}

int orc_svd3(const float* A, float* U, float* S, float* V)
{
    M3 u, v;
    svd3(A, u, S, v);
    std::memcpy(U, u.m, sizeof(u.m)); std::memcpy(V, v.m, sizeof(v.m));
    return ORC_OK;
}

int orc_weighted_transform(const float* src_xyz, const float* dst_xyz, int n, float* T16)
{
    Tfc tfc;
    for (int i = 0; i < n; ++i) {
        const float* from = src_xyz + 3 * i; const float* to = dst_xyz + 3 * i;
        if (std::isnan(from[2]) || std::isnan(to[2])) continue;
        tfc.add(from, to, 1.0f / (from[2] * to[2]));
    }
    tfc.transform(T16);
    return ORC_OK;
}

// 0 = inner sums left to right (default, what the CUDA path implements), 1 = Eigen's balanced-tree order: a sensitivity probe only
int orc_set_sum_order(int tree) { const int old = g_sumTree; g_sumTree = tree ? 1 : 0; return old; }

// pcl::TransformationFromCorrespondences over explicit (point, corresponding point, weight) triples, in the order given
int orc_tfc_transform(const float* p_xyz, const float* q_xyz, const float* w, int n, float* T16)
{
    Tfc tfc;
    for (int i = 0; i < n; ++i) tfc.add(p_xyz + 3 * i, q_xyz + 3 * i, w[i]);
    tfc.transform(T16);
    return ORC_OK;
}

// Eigen `S.llt().solve(b)`, 3x3 double, S row-major; returns 1 when S is positive definite (x written), else 0
int orc_llt3_solve(const double* S9, const double* b3, double* x3)
{
    double S[3][3];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) S[i][j] = S9[3 * i + j];
    return llt3_solve(S, b3, x3) ? 1 : 0;
}

double orc_mahalanobis2(const float* p1, const float* p2, const float* T16, double depth_cov)
{
    double T[16];
    for (int i = 0; i < 16; ++i) T[i] = T16[i];
    return mahal2(p1, p2, T, depth_cov);
}

// Ransac::Iterate(Frame*, Frame*, m12)  (ransac.cpp:155-267)
int orc_ransac_iterate(const orc_ransac_cfg* cfg, const float* src_xyz, int nsrc, const float* dst_xyz, int ndst,
    const orc_dmatch* m12, int nm, int sort_mode, const int* sample_table, unsigned seed, orc_dmatch* inliers_out,
    int cap, orc_ransac_out* out, orc_hyp_debug* per_hyp, orc_dmatch* good_sorted_out, int* sample_table_out)
{
    if (!cfg || !out) return ORC_ERR_ARG;
    std::memset(out, 0, sizeof(*out));
    float rmse = 1e6;
    float T12[16] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1 };
    std::vector<orc_dmatch> inliersBest;
    auto finish = [&](bool ok) {
        out->ok = ok ? 1 : 0; out->rmse = rmse; std::memcpy(out->T12, T12, sizeof(T12));
        out->n_inliers = (int)inliersBest.size();
        if ((int)inliersBest.size() > cap) return (int)ORC_ERR_CAPACITY;
        if (inliers_out) std::copy(inliersBest.begin(), inliersBest.end(), inliers_out);
        return (int)ORC_OK;
    };
    const int iters = cfg->iterations;
    const unsigned S = cfg->sample_size, minInl = cfg->min_inlier_th;
    if (per_hyp) for (int k = 0; k < iters; ++k) { std::memset(&per_hyp[k], 0, sizeof(orc_hyp_debug)); per_hyp[k].rounds = -1; }
    out->depth_cov_used = cfg->depth_cov;
    if ((unsigned)nm < minInl) return finish(false);

    Ctx c; c.src = src_xyz; c.dst = dst_xyz; c.maxMahal = cfg->max_mahal;
    c.cz = cfg->depth_cov; c.czLatched = cfg->depth_cov >= 0.0;

    std::vector<orc_dmatch> good; good.reserve(nm);
    for (int i = 0; i < nm; ++i) {
        const orc_dmatch& m = m12[i];
        if (m.queryIdx < 0 || m.queryIdx >= nsrc || m.trainIdx < 0 || m.trainIdx >= ndst) return ORC_ERR_ARG;
        const float sz = src_xyz[3 * m.queryIdx + 2], tz = dst_xyz[3 * m.trainIdx + 2];
        if (cfg->check_depth) {
            if (std::isnan(sz) || std::isnan(tz)) continue;
            if (sz <= 0 || tz <= 0) continue;
        }
        good.push_back(m);
    }
    out->n_good = (int)good.size();
    if (good.size() < minInl) return finish(false);
    auto less = [](const orc_dmatch& a, const orc_dmatch& b) { return a.distance < b.distance; };
    if (sort_mode == 0) std::sort(good.begin(), good.end(), less);
    else if (sort_mode == 2) std::stable_sort(good.begin(), good.end(), less);
    if (good_sorted_out) std::copy(good.begin(), good.end(), good_sorted_out);

    std::vector<int> table((size_t)std::max(iters, 0) * S, -1);
    if (good.size() >= S) {
        if (sample_table) std::copy(sample_table, sample_table + table.size(), table.begin());
        else orc_sample_table_libc(seed, (int)good.size(), iters, (int)S, table.data());
    }
    if (sample_table_out) std::copy(table.begin(), table.end(), sample_table_out);

    int realIters = 0, validIters = 0;
    double inlierError;
    for (int n = 0; n < iters && good.size() >= S; ++n) {
        double refinedError = 1e6;
        std::vector<orc_dmatch> refined, inl;
        const int* row = &table[(size_t)realIters * S];
        for (unsigned k = 0; k < S; ++k) if (row[k] >= 0) inl.push_back(good[row[k]]);
        float refinedT[16] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1 };
        const int hyp = realIters;
        realIters++;
        int rounds = 0;
        for (int refinements = 1; refinements < 20; ++refinements) {
            float T[16];
            transform_from(c, inl, T);
            ++rounds;
            inlierError = inliers_and_error(c, good, T, inl);
            if (inl.size() < minInl || inlierError > (double)cfg->max_mahal) break;
            if (inl.size() >= refined.size() && inlierError <= refinedError) {
                const size_t prev = refined.size();
                std::memcpy(refinedT, T, sizeof(T));
                refined = inl;
                refinedError = inlierError;
                if (inl.size() == prev) break;
            } else break;
        }
        if (per_hyp) {
            per_hyp[hyp].n_refined = (int)refined.size(); per_hyp[hyp].rounds = rounds;
            per_hyp[hyp].refined_error = refinedError; std::memcpy(per_hyp[hyp].T, refinedT, sizeof(refinedT));
        }
        if (!refined.empty()) {
            validIters++;
            if (refinedError <= (double)rmse && refined.size() >= inliersBest.size() && refined.size() >= minInl) {
                rmse = (float)refinedError;
                std::memcpy(T12, refinedT, sizeof(T12));
                inliersBest = refined;
                if (refined.size() > good.size() * 0.5) n += 10;
                if (refined.size() > good.size() * 0.75) n += 10;
                if (refined.size() > good.size() * 0.8) break;
            }
        }
    }
    if (validIters == 0) {
        const float I[16] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1 };
        std::vector<orc_dmatch> inl;
        inlierError = inliers_and_error(c, good, I, inl);
        if (inl.size() > minInl && inlierError < (double)cfg->max_mahal) {
            std::memcpy(T12, I, sizeof(T12));
            inliersBest = inl;
            rmse = (float)((double)rmse + inlierError);   // float += double: summed in double, rounded to float
            validIters++;
            out->used_identity = 1;
        }
    }
    out->real_iters = realIters; out->valid_iters = validIters; out->depth_cov_used = c.cz;
    return finish(inliersBest.size() >= minInl);
}

// Kabsch::Compute (kabsch.cpp:14-57): rows of setA/setB are points; returns T mapping A -> B.
int orc_kabsch(const float* A, const float* B, int n, float* T)
{
    const float I[16] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1 };
    std::memcpy(T, I, sizeof(I));
    if (n == 0) return ORC_OK;
    float ca[3] = { 0, 0, 0 }, cb[3] = { 0, 0, 0 };
    for (int i = 0; i < n; ++i) for (int k = 0; k < 3; ++k) { ca[k] += A[3 * i + k]; cb[k] += B[3 * i + k]; }
    for (int k = 0; k < 3; ++k) { ca[k] /= (float)n; cb[k] /= (float)n; }
    float H[9] = { 0 };   // A'^T B'
    for (int i = 0; i < n; ++i)
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) H[3 * r + c] += (A[3 * i + r] - ca[r]) * (B[3 * i + c] - cb[c]);
    M3 V, W; float S[3];
    svd3(H, V, S, W);        // H = V diag(S) W^T   (reference names: V = matrixU, W = matrixV)
    M3 Hm; std::memcpy(Hm.m, H, sizeof(H));
    const float det = det3(Hm);
    const float d = (det != 0.f) ? (float)((det > 0.f) - (det < 0.f)) : 1.f;
    float R[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            R[i][j] = (W.m[i][0] * V.m[j][0] + W.m[i][1] * V.m[j][1]) + (W.m[i][2] * d) * V.m[j][2];
    for (int i = 0; i < 3; ++i) {
        const float rc = (R[i][0] * ca[0] + R[i][1] * ca[1]) + R[i][2] * ca[2];
        T[4 * i] = R[i][0]; T[4 * i + 1] = R[i][1]; T[4 * i + 2] = R[i][2];
        T[4 * i + 3] = cb[i] - rc;
    }
    return ORC_OK;
}

}  // extern "C"

// ---- Odometry::Compute, RANSAC strategy: composition rule + inlier flags (Odometry/odometry.cpp:78-90), SURVEY.md §8f rank 3 ----
// pose[k + 1] = T12[k] * pose[k] as cv::Mat evaluates a 4x4 * 4x4 float product (cv::gemm's small-matrix path: every element is
// the four products summed left to right in float; pinned against cv2.gemm in tests/test_trajectory.py), sequentially along
// the sequence.  outlier[k + 1][j] starts true (Frame::mvbOutlier) and is cleared for every inlier's trainIdx (SetInlier).
int orc_compose_trajectory(const float* T12 /* [npairs][16] */, int npairs, const float* pose0 /* 16 */, float* poses /* [npairs + 1][16] */)
{
    if (npairs < 0 || !pose0 || !poses || (npairs > 0 && !T12)) return ORC_ERR_ARG;
    for (int i = 0; i < 16; ++i) poses[i] = pose0[i];
    for (int k = 0; k < npairs; ++k) {
        const float* A = T12 + (size_t)k * 16; const float* B = poses + (size_t)k * 16; float* D = poses + (size_t)(k + 1) * 16;
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 4; ++c) {
                float t = A[4 * r] * B[c];
                t = t + A[4 * r + 1] * B[4 + c];
                t = t + A[4 * r + 2] * B[8 + c];
                t = t + A[4 * r + 3] * B[12 + c];
                D[4 * r + c] = t;
            }
    }
    return ORC_OK;
}
