// =====================================================================================
// oracle/orb_oracle.cpp  —  TEST INFRASTRUCTURE ONLY (the parity checker, never the product)
//
// Dependency-free CPU restatement of the reference's ORB extraction path
//   /root/reference/Features/orbextractor.cpp   (ctor :346-404, ComputePyramid :833-857,
//   ComputeKeyPointsOctTree :665-746, DistributeOctTree :466-663, DivideNode :412-464,
//   IC_Angle :14-39, computeOrbDescriptor :43-85, operator() :756-815)
//   /root/reference/Core/frame.cpp:135-170      (depth gather + unprojection)
// plus the OpenCV-internal arithmetic those call sites delegate to (cv::resize INTER_LINEAR 8U,
// cv::FAST 9/16 + NMS, cv::GaussianBlur 7x7 sigma 2 8U fixed point, cv::fastAtan2), restated
// from OpenCV's published algorithms and pinned bit-for-bit against cv2 4.13.0 by
// tests/test_oracle_vs_cv2.py.
//
// PARITY STATUS: the reference ships no golden vectors / known-answer tests for this path
//  and cannot be built as a whole here (needs OpenCV/PCL/Eigen, absent).
// Pinned (round 2) against the reference's own Features/orbextractor.cpp compiled verbatim into
// oracle/_ref over an OpenCV stand-in (tests/test_oracle_vs_ref.py: keypoints incl. order, angles,
// descriptors, pyramid byte-identical); against cv2 for every stage the reference delegates to
// OpenCV; against committed goldens (tests/golden/).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load this library.  Build: oracle/Makefile  (g++ -O2 -ffp-contract=off, no -march=native:
// no FMA contraction, quirk Q4).
// =====================================================================================
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <list>
#include <utility>
#include <functional>
#include <vector>

#include "oracle_api.h"

namespace {

// ---- cvRound / cvFloor / cvCeil semantics (OpenCV fast_math.hpp): round-half-to-even --------
inline int cv_round(double v) { return (int)std::nearbyint(v); }  // default FE_TONEAREST
inline int cv_round(float v) { return (int)std::nearbyintf(v); }
inline int cv_floor(double v) { return (int)std::floor(v); }
inline int cv_ceil(double v) { return (int)std::ceil(v); }

const int kEdgeThreshold = 19;   // orbextractor.cpp:12
const int kHalfPatch = 15;       // orbextractor.cpp:11
const int kPatchSize = 31;       // orbextractor.cpp:10

const int8_t kPattern[1024] = {
#include "rbrief_pattern.inc"
};

// ---------------------------------------------------------------------------------------------
// a-0: constructor tables (orbextractor.cpp:346-404)
// ---------------------------------------------------------------------------------------------
struct Tables {
    std::vector<float> scale, inv_scale, sigma2, inv_sigma2;
    std::vector<int> nfeat;
    int umax[16];
};

Tables make_tables(int nfeatures, float scaleFactorF, int nlevels)
{
    Tables t;
    const double scaleFactor = scaleFactorF;  // member is a double initialised from a float (orbextractor.h:69)
    t.scale.resize(nlevels); t.sigma2.resize(nlevels);
    t.inv_scale.resize(nlevels); t.inv_sigma2.resize(nlevels);
    t.scale[0] = 1.0f; t.sigma2[0] = 1.0f;
    for (int i = 1; i < nlevels; ++i) {
        t.scale[i] = (float)((double)t.scale[i - 1] * scaleFactor);   // float*double -> double -> float
        t.sigma2[i] = t.scale[i] * t.scale[i];
    }
    for (int i = 0; i < nlevels; ++i) {
        t.inv_scale[i] = 1.0f / t.scale[i];
        t.inv_sigma2[i] = 1.0f / t.sigma2[i];
    }
    t.nfeat.resize(nlevels);
    const float factor = (float)(1.0 / scaleFactor);                  // 1.0f / double -> double -> float
    float desired = (float)nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int l = 0; l < nlevels - 1; ++l) {
        t.nfeat[l] = cv_round(desired);
        sum += t.nfeat[l];
        desired *= factor;
    }
    t.nfeat[nlevels - 1] = std::max(nfeatures - sum, 0);

    // circular patch row half-widths (orbextractor.cpp:389-403)
    int v, v0;
    const int vmax = cv_floor(kHalfPatch * std::sqrt(2.f) / 2 + 1);
    const int vmin = cv_ceil(kHalfPatch * std::sqrt(2.f) / 2);
    const double hp2 = kHalfPatch * kHalfPatch;
    for (v = 0; v < 16; ++v) t.umax[v] = 0;
    for (v = 0; v <= vmax; ++v) t.umax[v] = cv_round(std::sqrt(hp2 - v * v));
    for (v = kHalfPatch, v0 = 0; v >= vmin; --v) {
        while (t.umax[v0] == t.umax[v0 + 1]) ++v0;
        t.umax[v] = v0;
        ++v0;
    }
    return t;
}

void level_size(const Tables& t, int level, int w, int h, int* lw, int* lh)
{
    const float s = t.inv_scale[level];                               // orbextractor.cpp:836-838
    *lw = cv_round((float)w * s);
    *lh = cv_round((float)h * s);
}

// ---------------------------------------------------------------------------------------------
// P2: cv::resize(INTER_LINEAR) on CV_8UC1 — OpenCV's fixed-point bilinear (11-bit coefficients).
// ---------------------------------------------------------------------------------------------
struct ResizeTab { std::vector<int> ofs; std::vector<short> a0, a1; };

ResizeTab resize_tab(int src, int dst)
{
    ResizeTab r; r.ofs.resize(dst); r.a0.resize(dst); r.a1.resize(dst);
    const double inv_scale = (double)dst / src;
    const double scale = 1. / inv_scale;
    for (int d = 0; d < dst; ++d) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = cv_floor(f);
        f -= s;
        if (s < 0) { f = 0; s = 0; }
        if (s >= src - 1) { f = 0; s = src - 1; }
        r.ofs[d] = s;
        r.a0[d] = (short)cv_round((1.f - f) * 2048.f);
        r.a1[d] = (short)cv_round(f * 2048.f);
    }
    return r;
}

void resize_linear_u8(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dw, int dh, int dstride)
{
    ResizeTab tx = resize_tab(sw, dw), ty = resize_tab(sh, dh);
    std::vector<int> row0(dw), row1(dw);
    int cached0 = -1, cached1 = -1;
    auto hrow = [&](int sy, std::vector<int>& out) {
        const uint8_t* S = src + (size_t)sy * sstride;
        for (int x = 0; x < dw; ++x) {
            const int sx = tx.ofs[x];
            const int sx1 = std::min(sx + 1, sw - 1);  // a1 == 0 whenever sx == sw-1
            out[x] = S[sx] * tx.a0[x] + S[sx1] * tx.a1[x];
        }
    };
    for (int y = 0; y < dh; ++y) {
        const int sy0 = ty.ofs[y];
        const int sy1 = std::min(sy0 + 1, sh - 1);
        if (cached1 == sy0) { std::swap(row0, row1); std::swap(cached0, cached1); }
        if (cached0 != sy0) { hrow(sy0, row0); cached0 = sy0; }
        if (cached1 != sy1) { if (sy1 == sy0) row1 = row0; else hrow(sy1, row1); cached1 = sy1; }
        const int b0 = ty.a0[y], b1 = ty.a1[y];
        uint8_t* D = dst + (size_t)y * dstride;
        for (int x = 0; x < dw; ++x)
            D[x] = (uint8_t)((((b0 * (row0[x] >> 4)) >> 16) + ((b1 * (row1[x] >> 4)) >> 16) + 2) >> 2);
    }
}

// ---------------------------------------------------------------------------------------------
// P1: cv::FAST TYPE_9_16 corner strength.  S(p) = max over the 16 arcs of 9 contiguous ring
// pixels of min(d) for both polarities (d = centre - ring, and ring - centre); p is a corner at
// threshold th iff S > th; OpenCV's response = S - 1.
// ---------------------------------------------------------------------------------------------
const int kRingDx[16] = { 0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1 };
const int kRingDy[16] = { 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3 };

inline int fast_strength(const uint8_t* p, const int* ofs)
{
    int d[25];
    const int v = p[0];
    for (int k = 0; k < 16; ++k) d[k] = v - p[ofs[k]];
    for (int k = 16; k < 25; ++k) d[k] = d[k - 16];
    int best = -255;
    // bright-centre arcs: min over 9 contiguous d ; dark-centre arcs: min over 9 contiguous (-d)
    for (int k = 0; k < 16; k += 2) {
        int lo = d[k + 1], hi = d[k + 1];
        for (int j = 2; j <= 8; ++j) { lo = std::min(lo, d[k + j]); hi = std::max(hi, d[k + j]); }
        best = std::max(best, std::min(lo, d[k]));
        best = std::max(best, std::min(lo, d[k + 9]));
        best = std::max(best, -std::max(hi, d[k]));
        best = std::max(best, -std::max(hi, d[k + 9]));
    }
    return best;
}

// cv::FAST(roi, kps, th, nonmaxSuppression=true) on a roiW x roiH view: emits (x, y, response)
// row-major; 3-px frame never scored; NMS strict '>' against the 8 neighbours' scores where
// non-corners / unscored pixels hold 0.
void fast_roi(const uint8_t* roi, int stride, int roiW, int roiH, int th, std::vector<int>& scoreBuf,
    std::vector<orc_cand>& out)
{
    out.clear();
    if (roiW < 7 || roiH < 7) return;
    int ofs[16];
    for (int k = 0; k < 16; ++k) ofs[k] = kRingDy[k] * stride + kRingDx[k];
    scoreBuf.assign((size_t)roiW * roiH, 0);
    bool any = false;
    for (int y = 3; y < roiH - 3; ++y) {
        const uint8_t* row = roi + (size_t)y * stride;
        for (int x = 3; x < roiW - 3; ++x) {
            // cheap necessary condition first: an arc of 9 must contain one of each opposite pair
            const int v = row[x];
            const int hiT = v + th, loT = v - th;
            const int p0 = row[x + ofs[0]], p8 = row[x + ofs[8]];
            if (!((p0 > hiT) | (p8 > hiT) | (p0 < loT) | (p8 < loT))) continue;
            const int p4 = row[x + ofs[4]], p12 = row[x + ofs[12]];
            if (!((p4 > hiT) | (p12 > hiT) | (p4 < loT) | (p12 < loT))) continue;
            const int s = fast_strength(row + x, ofs);
            if (s > th) { scoreBuf[(size_t)y * roiW + x] = s - 1; any = true; }
        }
    }
    if (!any) return;
    for (int y = 3; y < roiH - 3; ++y)
        for (int x = 3; x < roiW - 3; ++x) {
            const int s = scoreBuf[(size_t)y * roiW + x];
            const int* c = &scoreBuf[(size_t)y * roiW + x];
            if (s <= 0) continue;  // non-corner; a response-0 corner (th == 0) can never win the strict '>' either
            if (s > c[-1] && s > c[1] && s > c[-roiW - 1] && s > c[-roiW] && s > c[-roiW + 1]
                && s > c[roiW - 1] && s > c[roiW] && s > c[roiW + 1]) {
                orc_cand k; k.x = x; k.y = y; k.score = s;
                out.push_back(k);
            }
        }
}

// a-2: per-cell FAST with ini/min threshold fallback (orbextractor.cpp:665-723).  Output
// coordinates are relative to minBorder (16); order = cell row-major, row-major inside a cell.
// cellTh (optional): iniTh of the cell whose first scored pixel is (x, y) — the region-adapted variant (orc_extract_adapted)
void fast_cells(const uint8_t* img, int w, int h, int stride, int iniTh, int minTh, std::vector<orc_cand>& cands,
    std::vector<int>* cellFallback, const std::function<int(int, int)>* cellTh = nullptr)
{
    cands.clear();
    const int minBorderX = kEdgeThreshold - 3, minBorderY = minBorderX;
    const int maxBorderX = w - kEdgeThreshold + 3, maxBorderY = h - kEdgeThreshold + 3;
    const float W = 30;
    const float width = (float)(maxBorderX - minBorderX), height = (float)(maxBorderY - minBorderY);
    const int nCols = (int)(width / W), nRows = (int)(height / W);
    if (nCols <= 0 || nRows <= 0) return;
    const int wCell = (int)std::ceil(width / nCols), hCell = (int)std::ceil(height / nRows);
    std::vector<int> scoreBuf;
    std::vector<orc_cand> cell;
    if (cellFallback) cellFallback->assign((size_t)nRows * nCols, -1);
    for (int i = 0; i < nRows; ++i) {
        const float iniY = (float)(minBorderY + i * hCell);
        float maxY = iniY + hCell + 6;
        if (iniY >= maxBorderY - 3) continue;
        if (maxY > maxBorderY) maxY = (float)maxBorderY;
        for (int j = 0; j < nCols; ++j) {
            const float iniX = (float)(minBorderX + j * wCell);
            float maxX = iniX + wCell + 6;
            if (iniX >= maxBorderX - 6) continue;
            if (maxX > maxBorderX) maxX = (float)maxBorderX;
            const int x0 = (int)iniX, y0 = (int)iniY, rw = (int)maxX - x0, rh = (int)maxY - y0;
            const uint8_t* roi = img + (size_t)y0 * stride + x0;
            fast_roi(roi, stride, rw, rh, cellTh ? (*cellTh)(x0 + 3, y0 + 3) : iniTh, scoreBuf, cell);
            int fb = 0;
            if (cell.empty()) { fast_roi(roi, stride, rw, rh, minTh, scoreBuf, cell); fb = 1; }
            if (cellFallback) (*cellFallback)[(size_t)i * nCols + j] = cell.empty() ? (fb ? 2 : 0) : fb;
            for (orc_cand k : cell) {
                k.x += j * wCell; k.y += i * hCell;
                cands.push_back(k);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// a-3: DistributeOctTree (orbextractor.cpp:466-663) + DivideNode (:412-464), literal list
// algorithm.  The reference breaks equal-count ties in phase 2 by heap pointer value (quirk Q3,
// run-to-run non-deterministic); the oracle DEFINES the tie order as creation sequence number
// (later-created node expands first among equals), i.e. sort key (count, seq) ascending,
// iterated from the back.
// ---------------------------------------------------------------------------------------------
struct QNode {
    int ulx, uly, urx, ury, blx, bly, brx, bry;
    std::vector<int> keys;
    bool noMore = false;
    long seq = 0;
    std::list<QNode>::iterator lit;
};

void divide_node(const QNode& n, const std::vector<orc_cand>& c, QNode out[4])
{
    const int halfX = (int)std::ceil((float)(n.urx - n.ulx) / 2);
    const int halfY = (int)std::ceil((float)(n.bry - n.uly) / 2);
    QNode &n1 = out[0], &n2 = out[1], &n3 = out[2], &n4 = out[3];
    n1.ulx = n.ulx; n1.uly = n.uly; n1.urx = n.ulx + halfX; n1.ury = n.uly;
    n1.blx = n.ulx; n1.bly = n.uly + halfY; n1.brx = n.ulx + halfX; n1.bry = n.uly + halfY;
    n2.ulx = n1.urx; n2.uly = n1.ury; n2.urx = n.urx; n2.ury = n.ury;
    n2.blx = n1.brx; n2.bly = n1.bry; n2.brx = n.urx; n2.bry = n.uly + halfY;
    n3.ulx = n1.blx; n3.uly = n1.bly; n3.urx = n1.brx; n3.ury = n1.bry;
    n3.blx = n.blx; n3.bly = n.bly; n3.brx = n1.brx; n3.bry = n.bly;
    n4.ulx = n3.urx; n4.uly = n3.ury; n4.urx = n2.brx; n4.ury = n2.bry;
    n4.blx = n3.brx; n4.bly = n3.bry; n4.brx = n.brx; n4.bry = n.bry;
    for (int k : n.keys) {
        const float px = (float)c[k].x, py = (float)c[k].y;
        if (px < (float)n1.urx) {
            if (py < (float)n1.bry) n1.keys.push_back(k); else n3.keys.push_back(k);
        } else if (py < (float)n1.bry) n2.keys.push_back(k);
        else n4.keys.push_back(k);
    }
    for (int q = 0; q < 4; ++q) if (out[q].keys.size() == 1) out[q].noMore = true;
}

int distribute_octtree(const std::vector<orc_cand>& c, int minX, int maxX, int minY, int maxY, int N,
    std::vector<int>& result)
{
    result.clear();
    const int nIni = (int)std::round((float)(maxX - minX) / (maxY - minY));
    if (nIni < 1) return ORC_ERR_GEOMETRY;
    const float hX = (float)(maxX - minX) / nIni;
    std::list<QNode> nodes;
    std::vector<QNode*> ini(nIni);
    long seq = 0;
    for (int i = 0; i < nIni; ++i) {
        QNode n;
        n.ulx = (int)(hX * (float)i); n.uly = 0;
        n.urx = (int)(hX * (float)(i + 1)); n.ury = 0;
        n.blx = n.ulx; n.bly = maxY - minY;
        n.brx = n.urx; n.bry = maxY - minY;
        n.seq = seq++;
        nodes.push_back(n);
        ini[i] = &nodes.back();
    }
    for (size_t k = 0; k < c.size(); ++k) {
        int r = (int)((float)c[k].x / hX);
        if (r >= nIni) r = nIni - 1;  // cannot happen for x < maxX-minX; guards the oracle only
        ini[r]->keys.push_back((int)k);
    }
    for (auto it = nodes.begin(); it != nodes.end();) {
        if (it->keys.size() == 1) { it->noMore = true; ++it; }
        else if (it->keys.empty()) it = nodes.erase(it);
        else ++it;
    }
    bool finish = false;
    typedef std::pair<int, long> SizeSeq;
    std::vector<std::pair<SizeSeq, QNode*>> expandable;
    auto push_children = [&](QNode ch[4], int* nToExpand) {
        for (int q = 0; q < 4; ++q) {
            if (ch[q].keys.empty()) continue;
            ch[q].seq = seq++;
            nodes.push_front(ch[q]);
            if (nodes.front().keys.size() > 1) {
                if (nToExpand) ++*nToExpand;
                expandable.push_back({ { (int)nodes.front().keys.size(), nodes.front().seq }, &nodes.front() });
                nodes.front().lit = nodes.begin();
            }
        }
    };
    while (!finish) {
        int prevSize = (int)nodes.size();
        int nToExpand = 0;
        expandable.clear();
        for (auto it = nodes.begin(); it != nodes.end();) {
            if (it->noMore) { ++it; continue; }
            QNode ch[4];
            divide_node(*it, c, ch);
            push_children(ch, &nToExpand);
            it = nodes.erase(it);
        }
        if ((int)nodes.size() >= N || (int)nodes.size() == prevSize) finish = true;
        else if ((int)nodes.size() + nToExpand * 3 > N) {
            while (!finish) {
                prevSize = (int)nodes.size();
                std::vector<std::pair<SizeSeq, QNode*>> prev = expandable;
                expandable.clear();
                std::sort(prev.begin(), prev.end(),
                    [](const std::pair<SizeSeq, QNode*>& a, const std::pair<SizeSeq, QNode*>& b) { return a.first < b.first; });
                for (int j = (int)prev.size() - 1; j >= 0; --j) {
                    QNode ch[4];
                    divide_node(*prev[j].second, c, ch);
                    push_children(ch, nullptr);
                    nodes.erase(prev[j].second->lit);
                    if ((int)nodes.size() >= N) break;
                }
                if ((int)nodes.size() >= N || (int)nodes.size() == prevSize) finish = true;
            }
        }
    }
    result.reserve(nodes.size());
    for (auto& n : nodes) {
        int best = n.keys[0];
        int bestScore = c[best].score;
        for (size_t k = 1; k < n.keys.size(); ++k)
            if (c[n.keys[k]].score > bestScore) { best = n.keys[k]; bestScore = c[best].score; }
        result.push_back(best);
    }
    return ORC_OK;
}

// ---------------------------------------------------------------------------------------------
// P4: cv::fastAtan2 (degrees, [0,360)), scalar f32 polynomial, no FMA.
// ---------------------------------------------------------------------------------------------
float fast_atan2_deg(float y, float x)
{
    const float scale = (float)(180 / 3.1415926535897932384626433832795);
    const float p1 = 0.9997878412794807f * scale;
    const float p3 = -0.3258083974640975f * scale;
    const float p5 = 0.1555786518463281f * scale;
    const float p7 = -0.04432655554792128f * scale;
    const float eps = (float)2.2204460492503131e-16;
    const float ax = std::fabs(x), ay = std::fabs(y);
    float a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + eps);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + eps);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

// a-5: IC_Angle (orbextractor.cpp:14-39) on the un-blurred level.
float ic_angle(const uint8_t* img, int stride, int x, int y, const int* umax)
{
    int m01 = 0, m10 = 0;
    const uint8_t* c = img + (size_t)y * stride + x;
    for (int u = -kHalfPatch; u <= kHalfPatch; ++u) m10 += u * c[u];
    for (int v = 1; v <= kHalfPatch; ++v) {
        int vsum = 0;
        const int d = umax[v];
        for (int u = -d; u <= d; ++u) {
            const int p = c[u + v * stride], m = c[u - v * stride];
            vsum += (p - m);
            m10 += u * (p + m);
        }
        m01 += v * vsum;
    }
    return fast_atan2_deg((float)m01, (float)m10);
}

// ---------------------------------------------------------------------------------------------
// P3: cv::GaussianBlur(7x7, sigma 2, BORDER_REFLECT_101) on CV_8UC1 — OpenCV's fixed-point
// smoothing: kernel [18,34,48,56,48,34,18]/256, horizontal Q8.8, vertical Q16.16, round-half-up.
// ---------------------------------------------------------------------------------------------
inline int reflect101(int p, int n)
{
    if (n == 1) return 0;
    while (p < 0 || p >= n) { if (p < 0) p = -p; else p = 2 * (n - 1) - p; }
    return p;
}

void gaussian_blur7(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride)
{
    static const int K[7] = { 18, 34, 48, 56, 48, 34, 18 };
    std::vector<uint16_t> tmp((size_t)w * h);
    std::vector<int> xi((size_t)w * 7);
    for (int x = 0; x < w; ++x) for (int k = 0; k < 7; ++k) xi[(size_t)x * 7 + k] = reflect101(x + k - 3, w);
    for (int y = 0; y < h; ++y) {
        const uint8_t* S = src + (size_t)y * sstride;
        uint16_t* T = &tmp[(size_t)y * w];
        for (int x = 0; x < w; ++x) {
            int acc = 0;
            const int* ix = &xi[(size_t)x * 7];
            for (int k = 0; k < 7; ++k) acc += K[k] * S[ix[k]];
            T[x] = (uint16_t)acc;
        }
    }
    for (int y = 0; y < h; ++y) {
        const uint16_t* R[7];
        for (int k = 0; k < 7; ++k) R[k] = &tmp[(size_t)reflect101(y + k - 3, h) * w];
        uint8_t* D = dst + (size_t)y * dstride;
        for (int x = 0; x < w; ++x) {
            uint32_t acc = 0;
            for (int k = 0; k < 7; ++k) acc += (uint32_t)K[k] * R[k][x];
            D[x] = (uint8_t)((acc + 32768u) >> 16);
        }
    }
}

// a-7: computeOrbDescriptor (orbextractor.cpp:43-85): steered BRIEF on the blurred level.
void rbrief(const uint8_t* img, int stride, int x, int y, float angleDeg, uint8_t* desc)
{
    const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
    const float angle = angleDeg * factorPI;
    // `(float)cos(angle)` with a float argument under `using namespace std` (orbextractor.cpp:7-8,45-46) is std::cos(float): libm's
    // cosf / sinf, not the double functions (checked against the reference's own source in oracle/_ref)
    const float a = cosf(angle), b = sinf(angle);
    const uint8_t* c = img + (size_t)y * stride + x;
    const int8_t* pat = kPattern;
    for (int i = 0; i < 32; ++i, pat += 32) {
        int val = 0;
        for (int k = 0; k < 8; ++k) {
            const float x0 = (float)pat[4 * k + 0], y0 = (float)pat[4 * k + 1];
            const float x1 = (float)pat[4 * k + 2], y1 = (float)pat[4 * k + 3];
            const int t0 = c[cv_round(x0 * b + y0 * a) * stride + cv_round(x0 * a - y0 * b)];
            const int t1 = c[cv_round(x1 * b + y1 * a) * stride + cv_round(x1 * a - y1 * b)];
            val |= (t0 < t1) << k;
        }
        desc[i] = (uint8_t)val;
    }
}

}  // namespace

// =============================================================================================
// C API (ctypes-friendly)
// =============================================================================================
extern "C" {

int orc_tables(int nfeatures, float scaleFactor, int nlevels, float* scale, float* inv_scale, float* sigma2,
    float* inv_sigma2, int* nfeat_per_level, int* umax16)
{
    if (nlevels < 1 || nlevels > ORC_MAX_LEVELS) return ORC_ERR_ARG;
    Tables t = make_tables(nfeatures, scaleFactor, nlevels);
    for (int i = 0; i < nlevels; ++i) {
        if (scale) scale[i] = t.scale[i];
        if (inv_scale) inv_scale[i] = t.inv_scale[i];
        if (sigma2) sigma2[i] = t.sigma2[i];
        if (inv_sigma2) inv_sigma2[i] = t.inv_sigma2[i];
        if (nfeat_per_level) nfeat_per_level[i] = t.nfeat[i];
    }
    if (umax16) for (int i = 0; i < 16; ++i) umax16[i] = t.umax[i];
    return ORC_OK;
}

int orc_level_sizes(int w, int h, float scaleFactor, int nlevels, int* ws, int* hs)
{
    if (nlevels < 1 || nlevels > ORC_MAX_LEVELS) return ORC_ERR_ARG;
    Tables t = make_tables(1000, scaleFactor, nlevels);
    for (int l = 0; l < nlevels; ++l) level_size(t, l, w, h, &ws[l], &hs[l]);
    return ORC_OK;
}

int orc_resize_linear(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dw, int dh, int dstride)
{
    if (sw < 1 || sh < 1 || dw < 1 || dh < 1) return ORC_ERR_ARG;
    resize_linear_u8(src, sw, sh, sstride, dst, dw, dh, dstride);
    return ORC_OK;
}

// Pyramid: levels written back to back (tight rows, stride == level width) into `out`.
int orc_pyramid(const uint8_t* img, int w, int h, int stride, float scaleFactor, int nlevels, uint8_t* out)
{
    if (nlevels < 1 || nlevels > ORC_MAX_LEVELS) return ORC_ERR_ARG;
    Tables t = make_tables(1000, scaleFactor, nlevels);
    size_t off = 0, prevOff = 0;
    int pw = w, ph = h;
    for (int l = 0; l < nlevels; ++l) {
        int lw, lh;
        level_size(t, l, w, h, &lw, &lh);
        if (l == 0) for (int y = 0; y < h; ++y) std::memcpy(out + (size_t)y * w, img + (size_t)y * stride, w);
        else resize_linear_u8(out + prevOff, pw, ph, pw, out + off, lw, lh, lw);
        prevOff = off; pw = lw; ph = lh;
        off += (size_t)lw * lh;
    }
    return ORC_OK;
}

int orc_fast_roi(const uint8_t* roi, int stride, int w, int h, int th, orc_cand* out, int cap, int* n)
{
    std::vector<int> buf; std::vector<orc_cand> v;
    fast_roi(roi, stride, w, h, th, buf, v);
    *n = (int)v.size();
    if ((int)v.size() > cap) return ORC_ERR_CAPACITY;
    std::copy(v.begin(), v.end(), out);
    return ORC_OK;
}

int orc_fast_strength_map(const uint8_t* img, int w, int h, int stride, int16_t* out)
{
    int ofs[16];
    for (int k = 0; k < 16; ++k) ofs[k] = kRingDy[k] * stride + kRingDx[k];
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x)
            out[(size_t)y * w + x] = (x < 3 || y < 3 || x >= w - 3 || y >= h - 3)
                ? (int16_t)-256 : (int16_t)fast_strength(img + (size_t)y * stride + x, ofs);
    return ORC_OK;
}

int orc_fast_cells(const uint8_t* img, int w, int h, int stride, int iniTh, int minTh, orc_cand* out, int cap, int* n)
{
    std::vector<orc_cand> v;
    fast_cells(img, w, h, stride, iniTh, minTh, v, nullptr);
    *n = (int)v.size();
    if ((int)v.size() > cap) return ORC_ERR_CAPACITY;
    std::copy(v.begin(), v.end(), out);
    return ORC_OK;
}

int orc_distribute(const orc_cand* cands, int n, int minX, int maxX, int minY, int maxY, int N, int* out_idx, int cap,
    int* n_out)
{
    std::vector<orc_cand> c(cands, cands + n);
    std::vector<int> r;
    if (n == 0) { *n_out = 0; return ORC_OK; }
    int rc = distribute_octtree(c, minX, maxX, minY, maxY, N, r);
    if (rc != ORC_OK) return rc;
    *n_out = (int)r.size();
    if ((int)r.size() > cap) return ORC_ERR_CAPACITY;
    std::copy(r.begin(), r.end(), out_idx);
    return ORC_OK;
}

float orc_fast_atan2(float y, float x) { return fast_atan2_deg(y, x); }

int orc_ic_angle(const uint8_t* img, int w, int h, int stride, const int* xs, const int* ys, int n, float* angles)
{
    Tables t = make_tables(1000, 1.2f, 1);
    for (int i = 0; i < n; ++i) {
        if (xs[i] < kHalfPatch || ys[i] < kHalfPatch || xs[i] >= w - kHalfPatch || ys[i] >= h - kHalfPatch) return ORC_ERR_ARG;
        angles[i] = ic_angle(img, stride, xs[i], ys[i], t.umax);
    }
    return ORC_OK;
}

int orc_gaussian_blur7(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride)
{
    if (w < 1 || h < 1) return ORC_ERR_ARG;
    gaussian_blur7(src, w, h, sstride, dst, dstride);
    return ORC_OK;
}

int orc_rbrief(const uint8_t* blurred, int w, int h, int stride, const int* xs, const int* ys, const float* angles, int n,
    uint8_t* desc)
{
    for (int i = 0; i < n; ++i) {
        if (xs[i] < 19 || ys[i] < 19 || xs[i] >= w - 19 || ys[i] >= h - 19) return ORC_ERR_ARG;
        rbrief(blurred, stride, xs[i], ys[i], angles[i], desc + (size_t)i * 32);
    }
    return ORC_OK;
}

const int8_t* orc_pattern(void) { return kPattern; }

static int extract_impl(const orc_extract_cfg* cfg, const uint8_t* img, int stride, orc_keypoint* kps, uint8_t* desc, int cap,
    int* n_out, orc_extract_debug* dbg, int grid, const int* regionTh);

// a-8: ORBextractor::operator() (orbextractor.cpp:756-815).  Optional per-stage dumps in `dbg`.
int orc_extract(const orc_extract_cfg* cfg, const uint8_t* img, int stride, orc_keypoint* kps, uint8_t* desc, int cap,
    int* n_out, orc_extract_debug* dbg)
{
    return extract_impl(cfg, img, stride, kps, desc, cap, n_out, dbg, 0, nullptr);
}

// BASELINE config 4, 8-level variant (SURVEY.md quirk Q14 — a north-star extension, DEFINED here; the reference's adaptive routes are
// single-scale): the ORB-SLAM2 extractor with iniThFAST replaced, per region of a grid x grid partition of the image, by the state of a
// DetectorAdjuster-style controller (detectoradjuster.cpp:22-59) that is fed the number of keypoints the extractor returned in that
// region and carries its threshold from frame to frame like the stateful detectors (videodynamicadaptedfeaturedetector.cpp:24-44):
//   * threshold of a FAST cell = max(minThFAST, min(254, (int)state[region])) — int truncation as detectoradjuster.cpp:27 — where the
//     cell's region is that of its first scored pixel (x, y) of level l mapped to the image: ((int)((float)y * scale[l]) * grid / height,
//     (int)((float)x * scale[l]) * grid / width), clamped to grid - 1;  the minThFAST fallback of empty cells is unchanged;
//   * after the frame: found[region] = keypoints returned there (region of ((int)kp.y * grid / height, (int)kp.x * grid / width));
//     found < min_features => state *= dec (tooFew), found > max_features => state *= inc (tooMany), clamped to [min_th, max_th]:
//     one controller step per frame.
// thresh [grid * grid] in / out (<= 0: init_th); region_th / region_found (optional): thresholds used for this frame, keypoints found.
int orc_extract_adapted(const orc_extract_cfg* cfg, const orc_adaptive_cfg* acfg, const uint8_t* img, int stride, double* thresh, orc_keypoint* kps,
    uint8_t* desc, int cap, int* n_out, int* region_th, int* region_found)
{
    if (!cfg || !acfg || !thresh || acfg->grid < 1 || acfg->grid > 5) return ORC_ERR_ARG;
    const int g = acfg->grid;
    std::vector<int> th(g * g);
    for (int r = 0; r < g * g; ++r) {
        if (!(thresh[r] > 0)) thresh[r] = acfg->init_th;
        th[r] = std::max(cfg->min_th_fast, std::min(254, (int)thresh[r]));
        if (region_th) region_th[r] = th[r];
    }
    const int rc = extract_impl(cfg, img, stride, kps, desc, cap, n_out, nullptr, g, th.data());
    if (rc != ORC_OK) return rc;
    std::vector<int> found(g * g, 0);
    for (int i = 0; i < *n_out; ++i) {
        const int ry = std::min(g - 1, (int)kps[i].y * g / cfg->height), rx = std::min(g - 1, (int)kps[i].x * g / cfg->width);
        found[ry * g + rx]++;
    }
    for (int r = 0; r < g * g; ++r) {
        if (region_found) region_found[r] = found[r];
        if (found[r] < acfg->min_features) { thresh[r] *= acfg->dec; if (thresh[r] < acfg->min_th) thresh[r] = acfg->min_th; }
        else if (found[r] > acfg->max_features) { thresh[r] *= acfg->inc; if (thresh[r] > acfg->max_th) thresh[r] = acfg->max_th; }
    }
    return ORC_OK;
}

static int extract_impl(const orc_extract_cfg* cfg, const uint8_t* img, int stride, orc_keypoint* kps, uint8_t* desc, int cap,
    int* n_out, orc_extract_debug* dbg, int grid, const int* regionTh)
{
    if (!cfg || !n_out) return ORC_ERR_ARG;
    *n_out = 0;
    const int w = cfg->width, h = cfg->height, nl = cfg->nlevels;
    if (!img || w <= 0 || h <= 0) return ORC_OK;  // empty image: outputs untouched (:758-759)
    if (nl < 1 || nl > ORC_MAX_LEVELS) return ORC_ERR_ARG;
    Tables t = make_tables(cfg->nfeatures, cfg->scale_factor, nl);

    std::vector<int> lw(nl), lh(nl);
    std::vector<size_t> off(nl + 1, 0);
    for (int l = 0; l < nl; ++l) { level_size(t, l, w, h, &lw[l], &lh[l]); off[l + 1] = off[l] + (size_t)lw[l] * lh[l]; }
    for (int l = 0; l < nl; ++l) if (lw[l] < 2 * kEdgeThreshold + 1 || lh[l] < 2 * kEdgeThreshold + 1) return ORC_ERR_GEOMETRY;
    std::vector<uint8_t> pyr(off[nl]);
    int rc = orc_pyramid(img, w, h, stride, cfg->scale_factor, nl, pyr.data());
    if (rc != ORC_OK) return rc;
    if (dbg && dbg->pyramid) std::memcpy(dbg->pyramid, pyr.data(), pyr.size());

    std::vector<std::vector<orc_keypoint>> all(nl);
    std::vector<std::vector<std::pair<int, int>>> allXY(nl);
    int candTotal = 0;
    for (int l = 0; l < nl; ++l) {
        const uint8_t* L = pyr.data() + off[l];
        std::vector<orc_cand> cands;
        const float sc = t.scale[l];
        const std::function<int(int, int)> cellTh = [&](int x, int y) {
            const int ry = std::min(grid - 1, (int)((float)y * sc) * grid / h), rx = std::min(grid - 1, (int)((float)x * sc) * grid / w);
            return regionTh[ry * grid + rx];
        };
        fast_cells(L, lw[l], lh[l], lw[l], cfg->ini_th_fast, cfg->min_th_fast, cands, nullptr, regionTh ? &cellTh : nullptr);
        if (dbg && dbg->cands) {
            if (candTotal + (int)cands.size() > dbg->cand_cap) return ORC_ERR_CAPACITY;
            std::copy(cands.begin(), cands.end(), dbg->cands + candTotal);
        }
        if (dbg && dbg->n_cands) dbg->n_cands[l] = (int)cands.size();
        candTotal += (int)cands.size();
        const int minB = kEdgeThreshold - 3;
        const int maxBX = lw[l] - kEdgeThreshold + 3, maxBY = lh[l] - kEdgeThreshold + 3;
        std::vector<int> keep;
        if (!cands.empty()) {
            rc = distribute_octtree(cands, minB, maxBX, minB, maxBY, t.nfeat[l], keep);
            if (rc != ORC_OK) return rc;
        }
        const int scaledPatch = (int)(kPatchSize * t.scale[l]);
        for (int idx : keep) {
            orc_keypoint k;
            const int x = cands[idx].x + minB, y = cands[idx].y + minB;
            k.x = (float)x; k.y = (float)y;
            k.size = (float)scaledPatch;
            k.response = (float)cands[idx].score;
            k.octave = l; k.class_id = -1;
            k.angle = ic_angle(L, lw[l], x, y, t.umax);
            all[l].push_back(k);
            allXY[l].push_back({ x, y });
        }
        if (dbg && dbg->n_kps) dbg->n_kps[l] = (int)all[l].size();
    }
    int total = 0;
    for (int l = 0; l < nl; ++l) total += (int)all[l].size();
    *n_out = total;
    if (total > cap) return ORC_ERR_CAPACITY;
    int o = 0;
    std::vector<uint8_t> blurred;
    for (int l = 0; l < nl; ++l) {
        if (all[l].empty()) continue;                         // level skipped (:791-792)
        blurred.resize((size_t)lw[l] * lh[l]);
        gaussian_blur7(pyr.data() + off[l], lw[l], lh[l], lw[l], blurred.data(), lw[l]);
        if (dbg && dbg->blurred) std::memcpy(dbg->blurred + off[l], blurred.data(), blurred.size());
        for (size_t i = 0; i < all[l].size(); ++i) {
            orc_keypoint k = all[l][i];
            rbrief(blurred.data(), lw[l], allXY[l][i].first, allXY[l][i].second, k.angle, desc + (size_t)o * 32);
            if (dbg && dbg->level_xy) { dbg->level_xy[2 * o] = allXY[l][i].first; dbg->level_xy[2 * o + 1] = allXY[l][i].second; }
            if (l != 0) { k.x *= t.scale[l]; k.y *= t.scale[l]; }
            kps[o++] = k;
        }
    }
    return ORC_OK;
}

// a-10: Frame::ExtractFeatures depth part (frame.cpp:148-164), with k1 == 0 (mvKeysUn = mvKeys).
// depth_u16 * depth_factor reproduces Frame's convertTo(CV_32F, 1/5000) (frame.cpp:24).
int orc_unproject(const orc_keypoint* kps, int n, const uint16_t* depth_u16, const float* depth_f32, int w, int h,
    int dstride_elems, float depth_factor, float fx, float fy, float cx, float cy, float mbf, float* xyz, float* uright)
{
    const float invfx = 1.0f / fx, invfy = 1.0f / fy;
    for (int i = 0; i < n; ++i) {
        const int u = (int)kps[i].x, v = (int)kps[i].y;
        float X = 0, Y = 0, Z = 0, ur = -1;
        if (u >= 0 && v >= 0 && u < w && v < h) {
            const float z = depth_f32 ? depth_f32[(size_t)v * dstride_elems + u]
                                      : (float)depth_u16[(size_t)v * dstride_elems + u] * depth_factor;
            if (z > 0) {
                ur = kps[i].x - mbf / z;
                X = (kps[i].x - cx) * z * invfx;
                Y = (kps[i].y - cy) * z * invfy;
                Z = z;
            }
        }
        xyz[3 * i] = X; xyz[3 * i + 1] = Y; xyz[3 * i + 2] = Z;
        if (uright) uright[i] = ur;
    }
    return ORC_OK;
}

// The same with mvKeysUn (Core/frame.cpp:148-164): the depth is looked up at the distorted keypoint kp (:152-155), mvuRight and
// mvKeys3Dc are formed from the undistorted kpU = xy_un[i] (:157-162).  xy_un == NULL is the k1 == 0 shortcut (frame.cpp:288-291).
int orc_unproject_un(const orc_keypoint* kps, const float* xy_un, int n, const uint16_t* depth_u16, const float* depth_f32, int w, int h,
    int dstride_elems, float depth_factor, float fx, float fy, float cx, float cy, float mbf, float* xyz, float* uright)
{
    if (!xy_un) return orc_unproject(kps, n, depth_u16, depth_f32, w, h, dstride_elems, depth_factor, fx, fy, cx, cy, mbf, xyz, uright);
    const float invfx = 1.0f / fx, invfy = 1.0f / fy;
    for (int i = 0; i < n; ++i) {
        const int u = (int)kps[i].x, v = (int)kps[i].y;
        const float xu = xy_un[2 * i], yu = xy_un[2 * i + 1];
        float X = 0, Y = 0, Z = 0, ur = -1;
        if (u >= 0 && v >= 0 && u < w && v < h) {
            const float z = depth_f32 ? depth_f32[(size_t)v * dstride_elems + u]
                                      : (float)depth_u16[(size_t)v * dstride_elems + u] * depth_factor;
            if (z > 0) {
                ur = xu - mbf / z;
                X = (xu - cx) * z * invfx;
                Y = (yu - cy) * z * invfy;
                Z = z;
            }
        }
        xyz[3 * i] = X; xyz[3 * i + 1] = Y; xyz[3 * i + 2] = Z;
        if (uright) uright[i] = ur;
    }
    return ORC_OK;
}

}  // extern "C"
