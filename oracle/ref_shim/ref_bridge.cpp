// oracle/ref_shim/ref_bridge.cpp — TEST INFRASTRUCTURE ONLY.
// C entry points over the reference's own ORBextractor (compiled verbatim from /root/reference/Features/orbextractor.cpp by
// oracle/Makefile, target _ref, against the OpenCV stand-in in this directory), so that tests can check oracle/orb_oracle.cpp's
// restatement against the reference SOURCE itself: oracle/_ref/liborb_ref.so.
//
// Quirk Q3: DistributeOctTree sorts (count, ExtractorNode*) pairs, so ties between equally populated nodes are
// broken by heap address — the reference is only deterministic for a given allocator.  Inside a bridge call every allocation
// comes from a bump arena that never reuses memory, which makes "address order" equal "creation order": the one order the oracle
// (and the CUDA kernels) define for those ties.  The reference code itself is untouched.
#include <cstdio>
#include <cstdlib>
#include <new>

// The reference source is this translation unit (found through -I$(REF)): its file-static helpers (IC_Angle, computeOrbDescriptor)
// become reachable without touching a line of it.
#include "Features/orbextractor.cpp"

namespace {
struct Arena { char* base = nullptr; size_t cap = 0, used = 0; bool active = false; };
thread_local Arena g_arena;
constexpr size_t ARENA_BYTES = (size_t)1 << 30;

struct ArenaScope {
    ArenaScope()
    {
        if (!g_arena.base) { g_arena.base = (char*)std::malloc(ARENA_BYTES); g_arena.cap = g_arena.base ? ARENA_BYTES : 0; }   // pages are touched on use
        g_arena.used = 0; g_arena.active = g_arena.base != nullptr;
    }
    ~ArenaScope() { g_arena.active = false; }
};
inline bool in_arena(void* p) { return g_arena.base && (char*)p >= g_arena.base && (char*)p < g_arena.base + g_arena.cap; }
}  // namespace

void* operator new(size_t n)
{
    if (g_arena.active) {
        const size_t a = (g_arena.used + 15) & ~(size_t)15;
        if (a + n <= g_arena.cap) { g_arena.used = a + n; return g_arena.base + a; }
        std::fprintf(stderr, "ref_bridge: arena exhausted\n"); std::abort();
    }
    void* p = std::malloc(n ? n : 1);
    if (!p) throw std::bad_alloc();
    return p;
}
void* operator new[](size_t n) { return operator new(n); }
void operator delete(void* p) noexcept { if (p && !in_arena(p)) std::free(p); }
void operator delete[](void* p) noexcept { operator delete(p); }
void operator delete(void* p, size_t) noexcept { operator delete(p); }
void operator delete[](void* p, size_t) noexcept { operator delete(p); }

extern "C" {

// ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)(image, noArray(), keypoints, descriptors) — the call
// Extractor::Extract makes (Features/extractor.cpp:41).  pyramid (optional): the levels of mvImagePyramid back to back, tight rows.
int ref_orb_extract(const uint8_t* img, int w, int h, int stride, int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh,
    orc_keypoint* kps, uint8_t* desc, int cap, int* n_out, uint8_t* pyramid, float* tables /* 4 * nlevels, optional */)
{
    if (!img || !n_out) return ORC_ERR_ARG;
    int rc = ORC_OK;
    {
        ArenaScope scope;
        {
            ORBextractor ex(nfeatures, scaleFactor, nlevels, iniTh, minTh);
            cv::Mat image(h, w, CV_8UC1, const_cast<uint8_t*>(img), (size_t)stride);
            std::vector<cv::KeyPoint> keys;
            cv::Mat descriptors;
            ex(image, cv::noArray(), keys, descriptors);
            *n_out = (int)keys.size();
            if ((int)keys.size() > cap) rc = ORC_ERR_CAPACITY;
            else {
                for (size_t i = 0; i < keys.size(); ++i) {
                    const cv::KeyPoint& k = keys[i];
                    if (kps) { orc_keypoint o = { k.pt.x, k.pt.y, k.size, k.angle, k.response, k.octave, k.class_id }; kps[i] = o; }
                    if (desc) std::memcpy(desc + 32 * i, descriptors.ptr((int)i), 32);
                }
            }
            if (pyramid)
                for (int l = 0; l < nlevels; ++l) {
                    const cv::Mat& m = ex.mvImagePyramid[l];
                    for (int r = 0; r < m.rows; ++r) { std::memcpy(pyramid, m.ptr(r), (size_t)m.cols); pyramid += m.cols; }
                }
            if (tables) {
                const std::vector<float> a = ex.GetScaleFactors(), b = ex.GetInverseScaleFactors(), c = ex.GetScaleSigmaSquares(), d = ex.GetInverseScaleSigmaSquares();
                for (int l = 0; l < nlevels; ++l) { tables[l] = a[l]; tables[nlevels + l] = b[l]; tables[2 * nlevels + l] = c[l]; tables[3 * nlevels + l] = d[l]; }
            }
        }
    }
    return rc;
}

// ---- finer-grained pins: the reference's own helpers on caller-chosen inputs ------------------------------------------------
namespace {
struct Probe : ORBextractor {      // protected members / methods of the reference class
    Probe(int nf, float sf, int nl, int a, int b) : ORBextractor(nf, sf, nl, a, b) {}
    using ORBextractor::DistributeOctTree;
    using ORBextractor::pattern;
    using ORBextractor::umax;
    using ORBextractor::mnFeaturesPerLevel;
};
}  // namespace

// ORBextractor::DistributeOctTree (orbextractor.cpp:466-663) on caller-supplied candidates (x, y relative to minBorder, response):
// out_idx = index into cands of every returned keypoint, in the returned order.
int ref_distribute(const orc_cand* cands, int n, int minX, int maxX, int minY, int maxY, int N, int* out_idx, int cap, int* n_out)
{
    ArenaScope scope;
    int rc = ORC_OK;
    {
        Probe ex(1000, 1.2f, 8, 20, 7);
        std::vector<cv::KeyPoint> keys;
        for (int i = 0; i < n; ++i) keys.push_back(cv::KeyPoint((float)cands[i].x, (float)cands[i].y, 7.f, -1.f, (float)cands[i].score, 0, i));   // class_id = input index
        std::vector<cv::KeyPoint> res = n ? ex.DistributeOctTree(keys, minX, maxX, minY, maxY, N, 0) : std::vector<cv::KeyPoint>();
        *n_out = (int)res.size();
        if ((int)res.size() > cap) rc = ORC_ERR_CAPACITY;
        else for (size_t i = 0; i < res.size(); ++i) out_idx[i] = res[i].class_id;
    }
    return rc;
}

// IC_Angle (orbextractor.cpp:14-39) with the constructor's umax table (:386-403)
int ref_ic_angle(const uint8_t* img, int w, int h, int stride, const int* xs, const int* ys, int n, float* angles)
{
    ArenaScope scope;
    {
        Probe ex(1000, 1.2f, 8, 20, 7);
        cv::Mat image(h, w, CV_8UC1, const_cast<uint8_t*>(img), (size_t)stride);
        for (int i = 0; i < n; ++i) angles[i] = IC_Angle(image, cv::Point2f((float)xs[i], (float)ys[i]), ex.umax);
    }
    return ORC_OK;
}

// computeOrbDescriptor (orbextractor.cpp:43-85) on an already blurred image, keypoint angle in degrees
int ref_orb_descriptor(const uint8_t* blurred, int w, int h, int stride, const int* xs, const int* ys, const float* angles, int n, uint8_t* desc)
{
    ArenaScope scope;
    {
        Probe ex(1000, 1.2f, 8, 20, 7);
        cv::Mat image(h, w, CV_8UC1, const_cast<uint8_t*>(blurred), (size_t)stride);
        for (int i = 0; i < n; ++i) {
            cv::KeyPoint kp((float)xs[i], (float)ys[i], 31.f, angles[i]);
            computeOrbDescriptor(kp, image, &ex.pattern[0], desc + (size_t)32 * i);
        }
    }
    return ORC_OK;
}

// constructor tables (orbextractor.cpp:346-403): features per level and umax
int ref_tables(int nfeatures, float scaleFactor, int nlevels, int* nfeat_per_level, int* umax16)
{
    ArenaScope scope;
    {
        Probe ex(nfeatures, scaleFactor, nlevels, 20, 7);
        for (int l = 0; l < nlevels; ++l) nfeat_per_level[l] = ex.mnFeaturesPerLevel[l];
        for (int v = 0; v < 16; ++v) umax16[v] = ex.umax[v];
    }
    return ORC_OK;
}

}  // extern "C"
