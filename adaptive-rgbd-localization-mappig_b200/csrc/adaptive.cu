// csrc/adaptive.cu — the reference's adaptive FAST detector route (BASELINE config 4):
//   Extractor::CreateAdaptiveDetector            Features/extractor.cpp:52-77
//   VideoGridAdaptedFeatureDetector::detect      Features/videogridadaptedfeaturedetector.cpp:52-84
//   VideoDynamicAdaptedFeatureDetector::detect   Features/videodynamicadaptedfeaturedetector.cpp:24-44
//   DetectorAdjuster (FAST branch)               Features/detectoradjuster.cpp:22-59
//
// The reference runs cv::FAST up to 5 times per grid cell and frame, each time at a new threshold.  cv::FAST's response
// does not depend on its threshold, so the GPU computes ONE response plane per frame at a floor
// threshold and every controller iteration becomes a lookup:
//   1. score_plane_kernel   FAST response of every pixel at the floor threshold (TMA tile + packed u16x2 pretest/strength,
//                           the same device code as the per-cell extractor), written as a u8 plane;
//   2. cell_collect_kernel  per overlapping grid cell: 3-px unscored frame and NMS clipped to the cell exactly like
//                           cv::FAST on the sub-image view, row-major candidate list + histogram of responses;
//   3. control_kernel       the per-cell threshold loop replayed over the frames of the batch in order (the threshold state
//                           carries from frame to frame), count(th) = #{response >= th} from the histogram;
//   4. cell_select_kernel   keypoints of the final detection, keepStrongest(max_per_cell) in detection order.
// If the loop ever visits a threshold below the floor (a run of "too few"), the batch is redone at the minimum threshold.
#include <algorithm>
#include <vector>

#include "fast_device.h"
#include "orbf_internal.h"

#define CTX_ENTER(c)                                                                   \
    do {                                                                               \
        if (!(c)) return ORBF_ERR_ARG;                                                 \
        cudaError_t e_ = cudaSetDevice((c)->cfg.device);                               \
        if (e_ != cudaSuccess) return orbf_cuda_fail((c), e_, "cudaSetDevice", __FILE__, __LINE__); \
    } while (0)
#define TRY(x) do { int r__ = (x); if (r__ != ORBF_OK) return r__; } while (0)

namespace {

constexpr int SP_W = 128, SP_H = 32, SP_BW = 160, SP_BH = SP_H + 6, SP_THREADS = 128;
constexpr int AD_MAX_CELLS = 25, AD_BATCH = 16;

struct AdGeom { int x0, y0, x1, y1; };   // cell sub-image [x0, x1) x [y0, y1)

struct AdParams {
    CUtensorMap map;                      // input planes of the sub-batch
    uint8_t* score; long long scoreFrameStride; int scorePitch;
    int w, h, floorTh, nCells, candCap, maxPerCell;
    AdGeom cell[AD_MAX_CELLS];
    uint32_t* cand; int* candCount; int* ge;          // [frames][cells][candCap], [frames][cells], [frames][cells][257]
    int* finalTh; int* found; int* under;             // [frames][cells], [frames][cells], [1]
    double* thresh;                                   // [cells] controller state (in/out)
    orbf_keypoint* outKp; int* outCount;              // [frames][cells][maxPerCell], [frames][cells]
    int nFrames, minFeat, maxFeat, maxIters;
    double minTh, maxTh, inc, dec;
};

// ---- 1. response plane -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SP_THREADS) score_plane_kernel(const __grid_constant__ AdParams P)
{
    __shared__ __align__(128) uint8_t tile[SP_BW * SP_BH];        // image (x0, y0) at [3][16]
    __shared__ __align__(16) uint8_t score[SP_W * SP_H];
    __shared__ uint16_t list[SP_W * SP_H];
    __shared__ __align__(8) uint64_t bar;
    __shared__ int sCount;
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * SP_W, y0 = blockIdx.y * SP_H, f = blockIdx.z;
    if (tid == 0) { mbar_init(&bar, 1); sCount = 0; }
    __syncthreads();
    if (tid == 0) { mbar_expect_tx(&bar, SP_BW * SP_BH); tma_load_3d(tile, &P.map, x0 - 16, y0 - 3, f, &bar); }
    for (int i = tid; i < SP_W * SP_H / 4; i += SP_THREADS) reinterpret_cast<uint32_t*>(score)[i] = 0;
    mbar_wait(&bar, 0);
    __syncthreads();
    const int th = P.floorTh;
    const uint32_t th1 = (uint32_t)(th + 1) * 0x00010001u;
    // pixels closer than 3 px to the image border are never scored (cv::FAST)
    for (int t = tid; t < SP_H * (SP_W / 4); t += SP_THREADS) {
        const int row = t >> 5, wc = t & 31;
        const int y = y0 + row, x = x0 + 4 * wc;
        if (y < 3 || y >= P.h - 3 || x + 3 < 3 || x >= P.w - 3) continue;
        const uint32_t* c = reinterpret_cast<const uint32_t*>(tile + (row + 3) * SP_BW) + 4 + wc;
        const uint32_t Cp = c[-1], C = c[0], Cn = c[1];
        const uint32_t U = *reinterpret_cast<const uint32_t*>(tile + row * SP_BW + 16 + 4 * wc);
        const uint32_t D = *reinterpret_cast<const uint32_t*>(tile + (row + 6) * SP_BW + 16 + 4 * wc);
        const uint32_t L = __funnelshift_r(Cp, C, 8), R = __funnelshift_r(C, Cn, 24);
        const uint32_t f01 = pretest_x2(__byte_perm(C, 0, 0x4140), __byte_perm(D, 0, 0x4140), __byte_perm(U, 0, 0x4140),
            __byte_perm(R, 0, 0x4140), __byte_perm(L, 0, 0x4140), th1);
        const uint32_t f23 = pretest_x2(__byte_perm(C, 0, 0x4342), __byte_perm(D, 0, 0x4342), __byte_perm(U, 0, 0x4342),
            __byte_perm(R, 0, 0x4342), __byte_perm(L, 0, 0x4342), th1);
        uint32_t flags = ((f01 & 0xFFFFu) == 0 ? 1u : 0u) | ((f01 >> 16) == 0 ? 2u : 0u) | ((f23 & 0xFFFFu) == 0 ? 4u : 0u) | ((f23 >> 16) == 0 ? 8u : 0u);
#pragma unroll
        for (int b = 0; b < 4; ++b) if (x + b < 3 || x + b >= P.w - 3) flags &= ~(1u << b);
        if (flags) {
            int pos = atomicAdd(&sCount, __popc(flags));
            const int e = row * SP_W + 4 * wc;
#pragma unroll
            for (int b = 0; b < 4; ++b) if (flags & (1u << b)) list[pos++] = (uint16_t)(e + b);
        }
    }
    __syncthreads();
    const int n = sCount;
    for (int i = tid; 2 * i < n; i += SP_THREADS) {
        const int ea = list[2 * i], eb = list[min(2 * i + 1, n - 1)];
        const uint8_t* pa = tile + ((ea >> 7) + 3) * SP_BW + 16 + (ea & 127);
        const uint8_t* pb = tile + ((eb >> 7) + 3) * SP_BW + 16 + (eb & 127);
        const uint32_t s = ring_strength_x2(pa, pb, SP_BW);
        const int sa = (int)(s & 0xFFFFu), sb = (int)(s >> 16);
        if (sa > th) score[ea] = (uint8_t)(sa - 1);
        if (sb > th) score[eb] = (uint8_t)(sb - 1);
    }
    __syncthreads();
    uint8_t* dst = P.score + (long long)f * P.scoreFrameStride + (long long)y0 * P.scorePitch + x0;
    for (int t = tid; t < SP_H * (SP_W / 4); t += SP_THREADS) {
        const int row = t >> 5, wc = t & 31;
        if (y0 + row < P.h && x0 + 4 * wc < P.scorePitch)
            *reinterpret_cast<uint32_t*>(dst + (long long)row * P.scorePitch + 4 * wc) = reinterpret_cast<const uint32_t*>(score)[t];
    }
}

// ---- 2. per-cell NMS, ordered candidate list, response histogram -------------------------------------------------------
constexpr int CC_WARPS = 8, CC_THREADS = CC_WARPS * 32, CC_MAX_ROWS = 1024;

// kept-corner bits of one aligned score word (byte b = pixel x + b); scored area [xa, xb) x [ya, yb)
__device__ __forceinline__ uint32_t nms_word(const uint8_t* plane, int pitch, int x, int y, uint32_t word, int xa, int xb, int ya, int yb)
{
    uint32_t keep = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        const int v = (word >> (8 * b)) & 255, px = x + b;
        if (v == 0 || px < xa || px >= xb) continue;
        const uint8_t* s = plane + (long long)y * pitch + px;
        const bool l = px > xa, r = px < xb - 1, u = y > ya, d = y < yb - 1;     // neighbours outside the scored area count as 0
        bool k = true;
        if (l) k = k && v > s[-1];
        if (r) k = k && v > s[1];
        if (u) { k = k && v > s[-pitch]; if (l) k = k && v > s[-pitch - 1]; if (r) k = k && v > s[-pitch + 1]; }
        if (d) { k = k && v > s[pitch]; if (l) k = k && v > s[pitch - 1]; if (r) k = k && v > s[pitch + 1]; }
        if (k) keep |= 1u << b;
    }
    return keep;
}

__global__ void __launch_bounds__(CC_THREADS) cell_collect_kernel(const __grid_constant__ AdParams P)
{
    __shared__ int rowCount[CC_MAX_ROWS];
    __shared__ int hist[256];
    const int cellIdx = blockIdx.x, f = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const AdGeom g = P.cell[cellIdx];
    const int xa = g.x0 + 3, xb = g.x1 - 3, ya = g.y0 + 3, yb = g.y1 - 3;      // cv::FAST never scores the sub-image's 3-px frame
    const int rows = max(yb - ya, 0);
    const uint8_t* plane = P.score + (long long)f * P.scoreFrameStride;
    const int pitch = P.scorePitch;
    const int w0 = xa >> 2, w1 = (xb + 3) >> 2;                                 // aligned words covering [xa, xb)
    for (int i = tid; i < 256; i += CC_THREADS) hist[i] = 0;
    for (int i = tid; i < rows; i += CC_THREADS) rowCount[i] = 0;
    __syncthreads();
    const uint32_t lt = (1u << lane) - 1;
    uint32_t* out = P.cand + ((long long)f * P.nCells + cellIdx) * P.candCap;
    for (int pass = 0; pass < 2; ++pass) {
        for (int r = warp; r < rows; r += CC_WARPS) {
            const int y = ya + r;
            int base = pass ? rowCount[r] : 0, cnt = 0;
            for (int wb = w0; wb < w1; wb += 32) {
                const int wi = wb + lane;
                uint32_t word = 0, keep = 0;
                if (wi < w1) {
                    word = __ldg(reinterpret_cast<const uint32_t*>(plane + (long long)y * pitch) + wi);
                    if (word) keep = nms_word(plane, pitch, 4 * wi, y, word, xa, xb, ya, yb);
                }
                if (!__any_sync(0xffffffffu, keep != 0)) continue;
                const int c = __popc(keep);
                const uint32_t b0 = __ballot_sync(0xffffffffu, c & 1), b1 = __ballot_sync(0xffffffffu, c & 2), b2 = __ballot_sync(0xffffffffu, c & 4);
                if (pass == 0) {
#pragma unroll
                    for (int b = 0; b < 4; ++b) if (keep & (1u << b)) atomicAdd(&hist[(word >> (8 * b)) & 255], 1);
                } else {
                    int pos = base + cnt + __popc(b0 & lt) + 2 * __popc(b1 & lt) + 4 * __popc(b2 & lt);
#pragma unroll
                    for (int b = 0; b < 4; ++b)
                        if ((keep & (1u << b)) && pos < P.candCap)
                            out[pos++] = (uint32_t)(4 * wi + b) | ((uint32_t)y << 12) | (((word >> (8 * b)) & 255u) << 24);
                }
                cnt += __popc(b0) + 2 * __popc(b1) + 4 * __popc(b2);
            }
            if (pass == 0 && lane == 0) rowCount[r] = cnt;
        }
        __syncthreads();
        if (pass == 0) {
            if (tid == 0) {                                                 // exclusive prefix over rows (<= ~300 rows), suffix sums of the histogram
                int acc = 0;
                for (int r = 0; r < rows; ++r) { const int c = rowCount[r]; rowCount[r] = acc; acc += c; }
                int* ge = P.ge + ((long long)f * P.nCells + cellIdx) * 257;
                int suf = 0;
                ge[256] = 0;
                for (int v = 255; v >= 0; --v) { suf += hist[v]; ge[v] = suf; }
                P.candCount[f * P.nCells + cellIdx] = min(acc, P.candCap);
            }
            __syncthreads();
        }
    }
}

// ---- 3. threshold controllers, frame after frame -------------------------------------------------------------------------
__global__ void control_kernel(const __grid_constant__ AdParams P)
{
    const int cellIdx = threadIdx.x;
    if (cellIdx >= P.nCells) return;
    double th = P.thresh[cellIdx];
    for (int f = 0; f < P.nFrames; ++f) {
        const int* ge = P.ge + ((long long)f * P.nCells + cellIdx) * 257;
        int iterCount = P.maxIters, usedTh = 0, n = 0;
        do {
            usedTh = (int)th;
            if (usedTh < P.floorTh) atomicExch(P.under, 1);
            n = ge[min(max(usedTh, 0), 256)];
            if (n < P.minFeat) { th *= P.dec; if (th < P.minTh) th = P.minTh; }
            else if (n > P.maxFeat) { th *= P.inc; if (th > P.maxTh) th = P.maxTh; break; }
            else break;
            iterCount--;
        } while (iterCount > 0 && th > P.minTh && th < P.maxTh);
        P.finalTh[f * P.nCells + cellIdx] = usedTh;
        P.found[f * P.nCells + cellIdx] = n;
    }
    P.thresh[cellIdx] = th;
}

// ---- 4. final detection + keepStrongest, in detection order ----------------------------------------------------------------
constexpr int CS_THREADS = 256;

__global__ void __launch_bounds__(CS_THREADS) cell_select_kernel(const __grid_constant__ AdParams P)
{
    __shared__ int sWarp[2][CS_THREADS / 32];
    __shared__ int sBase[2];
    const int cellIdx = blockIdx.x, f = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int slot = f * P.nCells + cellIdx;
    const int* ge = P.ge + (long long)slot * 257;
    const int th = min(max(P.finalTh[slot], 0), 256), N = P.maxPerCell;
    const int n = ge[th];
    int cut = th, takeEq = 0x7fffffff;                       // keep response > cut, plus the first takeEq with response == cut
    if (n > N) {
        int r = th;
        while (r < 255 && ge[r + 1] >= N) ++r;               // ge[r] >= N > ge[r + 1]
        cut = r; takeEq = N - ge[r + 1];
    } else cut = th - 1;                                      // everything >= th
    const uint32_t* cand = P.cand + (long long)slot * P.candCap;
    const int nc = P.candCount[slot];
    orbf_keypoint* out = P.outKp + (long long)slot * N;
    if (tid == 0) { sBase[0] = 0; sBase[1] = 0; }
    __syncthreads();
    const uint32_t lt = (1u << lane) - 1;
    for (int base = 0; base < nc; base += CS_THREADS) {
        const int i = base + tid;
        uint32_t e = 0;
        int resp = -1;
        if (i < nc) { e = cand[i]; resp = (int)(e >> 24); }
        const bool isEq = (n > N) && resp == cut, gt = resp > cut;
        const uint32_t mEq = __ballot_sync(0xffffffffu, isEq);
        if (lane == 0) sWarp[0][warp] = __popc(mEq);
        __syncthreads();
        int eqBefore = sBase[0];
        for (int k = 0; k < warp; ++k) eqBefore += sWarp[0][k];
        const bool keep = gt || (isEq && eqBefore + __popc(mEq & lt) < takeEq);
        const uint32_t mK = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) sWarp[1][warp] = __popc(mK);
        __syncthreads();
        int pos = sBase[1];
        for (int k = 0; k < warp; ++k) pos += sWarp[1][k];
        pos += __popc(mK & lt);
        if (keep && pos < N) {
            orbf_keypoint k;
            k.x = (float)(e & 0xFFF); k.y = (float)((e >> 12) & 0xFFF); k.size = 7.f; k.angle = -1.f; k.response = (float)resp;
            k.octave = 0; k.class_id = -1;
            out[pos] = k;
        }
        __syncthreads();
        if (tid == 0) {
            int a = 0, b = 0;
            for (int k = 0; k < CS_THREADS / 32; ++k) { a += sWarp[0][k]; b += sWarp[1][k]; }
            sBase[0] += a; sBase[1] += b;
        }
        __syncthreads();
    }
    if (tid == 0) P.outCount[slot] = min(sBase[1], N);
}

// ---- BASELINE config 4, 8-level variant: per-region controllers around the ORB extractor ------------------------------------------
// One CTA after the quadtree of frame slot `slot`: found[region] = keypoints the extractor returned there (their image coordinates
// are level coordinates * scale[level], as the output assembly computes them), log of (threshold used, found) for the frame, then one
// DetectorAdjuster step per region (tooFew: *= dec, tooMany: *= inc, clamped) and the integer thresholds of the NEXT frame's cells.
struct RegionParams {
    const uint32_t* lkp; const int* lkpCount; int kpStageTotal, L, slot, frameIdx;
    int kpOff[ORBF_MAX_LEVELS]; float scale[ORBF_MAX_LEVELS];
    int grid, width, height, minFeat, maxFeat, minThFast;
    double minTh, maxTh, inc, dec;
    int* regionTh; double* state; int* log;         // log [frame][2][grid * grid]: thresholds used, keypoints found
    int nVideos;                                    // block v: slot + v, frame log row frameIdx + v, state / thresholds of video v (25 entries each)
};

__global__ void __launch_bounds__(256) region_control_kernel(const RegionParams P)
{
    __shared__ int sFound[AD_MAX_CELLS];
    const int tid = threadIdx.x, g2 = P.grid * P.grid;
    if (tid < AD_MAX_CELLS) sFound[tid] = 0;
    __syncthreads();
    const int v = blockIdx.x, slot = P.slot + v;
    for (int l = 0; l < P.L; ++l) {
        const int n = P.lkpCount[slot * ORBF_MAX_LEVELS + l];
        const uint32_t* kp = P.lkp + (long long)slot * P.kpStageTotal + P.kpOff[l];
        for (int i = tid; i < n; i += 256) {
            const uint32_t key = kp[i];
            float x = (float)((int)(key & 0x7FF) + ORBF_MINB), y = (float)((int)((key >> 11) & 0x7FF) + ORBF_MINB);
            if (l > 0) { x = __fmul_rn(x, P.scale[l]); y = __fmul_rn(y, P.scale[l]); }
            const int ry = min(P.grid - 1, (int)y * P.grid / P.height), rx = min(P.grid - 1, (int)x * P.grid / P.width);
            atomicAdd(&sFound[ry * P.grid + rx], 1);
        }
    }
    __syncthreads();
    if (tid < g2) {
        const int found = sFound[tid];
        int* log = P.log + (long long)(P.frameIdx + v) * 2 * g2;
        int* regionTh = P.regionTh + v * AD_MAX_CELLS; double* state = P.state + v * AD_MAX_CELLS;
        log[tid] = regionTh[tid]; log[g2 + tid] = found;
        double st = state[tid];
        if (found < P.minFeat) { st *= P.dec; if (st < P.minTh) st = P.minTh; }
        else if (found > P.maxFeat) { st *= P.inc; if (st > P.maxTh) st = P.maxTh; }
        state[tid] = st;
        regionTh[tid] = max(P.minThFast, min(254, (int)st));
    }
}

}  // namespace

int orbf_launch_region_control(orbf_context* c, int slot, int frameIdx, const orbf_adaptive_config& cfg, int nVideos)
{
    RegionParams P;
    P.lkp = c->d_lkp; P.lkpCount = c->d_lkpCount; P.kpStageTotal = c->kpStageTotal; P.L = c->L; P.slot = slot; P.frameIdx = frameIdx;
    for (int l = 0; l < c->L; ++l) { P.kpOff[l] = c->lg[l].kpOff; P.scale[l] = c->scale[l]; }
    P.grid = cfg.grid; P.width = c->cfg.width; P.height = c->cfg.height; P.minFeat = cfg.min_features; P.maxFeat = cfg.max_features;
    P.minThFast = c->cfg.min_th_fast; P.minTh = cfg.min_th; P.maxTh = cfg.max_th; P.inc = cfg.inc; P.dec = cfg.dec;
    P.regionTh = c->d_regionTh; P.state = c->d_regionState; P.log = c->d_regionLog;
    P.nVideos = nVideos;
    region_control_kernel<<<nVideos, 256, 0, c->stream>>>(P);
    ORBF_LAUNCH_CHECK(c);
    return ORBF_OK;
}

// region of every FAST cell for a grid x grid partition: that of the cell's first scored pixel mapped to the image (the same float
// product and integer division as the oracle's definition), and room for nFrames log rows
int orbf_region_tables(orbf_context* c, int grid, int nFrames, int nVideos)
{
    const int g2 = grid * grid;
    if (c->cellRegionGrid != grid) {
        std::vector<uint8_t> reg(c->h_cells.size());
        for (size_t i = 0; i < c->h_cells.size(); ++i) {
            const CellDesc& d = c->h_cells[i];
            const float sc = c->scale[d.level];
            const int ry = std::min(grid - 1, (int)((float)d.y0 * sc) * grid / c->cfg.height), rx = std::min(grid - 1, (int)((float)d.x0 * sc) * grid / c->cfg.width);
            reg[i] = (uint8_t)(ry * grid + rx);
        }
        if (!c->d_cellRegion) ORBF_CUDA(c, cudaMalloc((void**)&c->d_cellRegion, std::max<size_t>(reg.size(), 1)));

        ORBF_CUDA(c, cudaMemcpyAsync(c->d_cellRegion, reg.data(), reg.size(), cudaMemcpyHostToDevice, c->stream));
        ORBF_CUDA(c, cudaStreamSynchronize(c->stream));
        c->cellRegionGrid = grid;
    }
    if (nVideos > c->regionVideos) {
        ORBF_CUDA(c, cudaStreamSynchronize(c->stream));
        if (c->d_regionTh) cudaFree(c->d_regionTh);
        if (c->d_regionState) cudaFree(c->d_regionState);
        c->d_regionTh = nullptr; c->d_regionState = nullptr; c->regionVideos = 0;
        ORBF_CUDA(c, cudaMalloc((void**)&c->d_regionTh, (size_t)nVideos * AD_MAX_CELLS * sizeof(int)));
        ORBF_CUDA(c, cudaMalloc((void**)&c->d_regionState, (size_t)nVideos * AD_MAX_CELLS * sizeof(double)));
        c->regionVideos = nVideos;
    }
    if (nFrames * 2 * g2 > c->regionLogCap) {
        ORBF_CUDA(c, cudaStreamSynchronize(c->stream));
        if (c->d_regionLog) cudaFree(c->d_regionLog);
        c->d_regionLog = nullptr; c->regionLogCap = 0;
        ORBF_CUDA(c, cudaMalloc((void**)&c->d_regionLog, (size_t)nFrames * 2 * g2 * sizeof(int)));
        c->regionLogCap = nFrames * 2 * g2;
    }
    return ORBF_OK;
}

extern "C" void orbf_default_adaptive_config(orbf_adaptive_config* c)
{
    if (!c) return;
    c->grid = 3; c->edge = 31; c->max_iters = 5;                               // extractor.cpp:67-70
    c->min_features = 67; c->max_features = 113; c->max_per_cell = 113;         // round(600/9), round(1020/9), 1020/9
    c->init_th = 20; c->min_th = 2; c->max_th = 10000; c->inc = 1.3; c->dec = 0.7;
    c->retain_best = 1000;                                                      // Extract(): retainBest(nFeatures), Utils/common.h:77
}

extern "C" int orbf_adaptive_detect(orbf_context* c, const orbf_adaptive_config* cfg, const uint8_t* gray, int32_t n, int32_t stride,
    int64_t frame_stride, double* thresh, orbf_keypoint* out, int32_t* counts, int32_t cap, int32_t* cell_thresh, int32_t* cell_found)
{
    CTX_ENTER(c);
    if (!cfg || !gray || !thresh || !counts || n < 1 || stride < c->cfg.width) return ORBF_ERR_ARG;
    const int g = cfg->grid, nCells = g * g, w = c->cfg.width, h = c->cfg.height;
    if (g < 1 || nCells > AD_MAX_CELLS || cfg->max_per_cell < 1 || w > 4095 || h > 4095) return ORBF_ERR_ARG;
    AdParams P;
    P.w = w; P.h = h; P.nCells = nCells; P.maxPerCell = cfg->max_per_cell;
    P.minFeat = cfg->min_features; P.maxFeat = cfg->max_features; P.maxIters = cfg->max_iters;
    P.minTh = cfg->min_th; P.maxTh = cfg->max_th; P.inc = cfg->inc; P.dec = cfg->dec;
    int candCap = 0;
    for (int i = 0; i < g; ++i)
        for (int j = 0; j < g; ++j) {
            AdGeom& q = P.cell[i * g + j];
            q.y0 = std::max((i * h) / g - cfg->edge, 0); q.y1 = std::min(h, ((i + 1) * h) / g + cfg->edge);
            q.x0 = std::max((j * w) / g - cfg->edge, 0); q.x1 = std::min(w, ((j + 1) * w) / g + cfg->edge);
            if (q.y1 - q.y0 > CC_MAX_ROWS) return ORBF_ERR_GEOMETRY;
            candCap = std::max(candCap, ((q.x1 - q.x0 + 1) / 2) * ((q.y1 - q.y0 + 1) / 2));
        }
    P.candCap = candCap;
    const int B = std::min<int>(n, AD_BATCH);
    const int pitch = align_up(w, 128);
    // scratch of one sub-batch: pieces of the context's persistent scratch (no allocation per call once it has grown)
    Scratch sc(c);
    const size_t oIn = sc.take((size_t)B * pitch * h), oScore = sc.take((size_t)B * pitch * h + 256), oCand = sc.take((size_t)B * nCells * candCap * sizeof(uint32_t));
    const size_t oCandCount = sc.take((size_t)B * nCells * sizeof(int)), oGe = sc.take((size_t)B * nCells * 257 * sizeof(int)), oFinal = sc.take((size_t)B * nCells * sizeof(int));
    const size_t oFound = sc.take((size_t)B * nCells * sizeof(int)), oUnder = sc.take(sizeof(int)), oOutCount = sc.take((size_t)B * nCells * sizeof(int));
    const size_t oThresh = sc.take((size_t)nCells * sizeof(double)), oOut = sc.take((size_t)B * nCells * cfg->max_per_cell * sizeof(orbf_keypoint));
#define AD_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return orbf_cuda_fail(c, e__, #call, __FILE__, __LINE__); } while (0)
    AD_CUDA(sc.alloc());
    uint8_t *dIn = sc.at<uint8_t>(oIn), *dScore = sc.at<uint8_t>(oScore); uint32_t* dCand = sc.at<uint32_t>(oCand);
    int *dCandCount = sc.at<int>(oCandCount), *dGe = sc.at<int>(oGe), *dFinal = sc.at<int>(oFinal), *dFound = sc.at<int>(oFound), *dUnder = sc.at<int>(oUnder),
        *dOutCount = sc.at<int>(oOutCount);
    double* dThresh = sc.at<double>(oThresh); orbf_keypoint* dOut = sc.at<orbf_keypoint>(oOut);
    P.score = dScore; P.scoreFrameStride = (long long)pitch * h; P.scorePitch = pitch;
    P.cand = dCand; P.candCount = dCandCount; P.ge = dGe; P.finalTh = dFinal; P.found = dFound; P.under = dUnder; P.thresh = dThresh;
    P.outKp = dOut; P.outCount = dOutCount;
    std::vector<double> st(thresh, thresh + nCells);
    for (double& t : st) if (!(t > 0)) t = cfg->init_th;
    std::vector<orbf_keypoint> hOut((size_t)B * nCells * cfg->max_per_cell);
    std::vector<int> hCount((size_t)B * nCells), hFinal((size_t)B * nCells), hFound((size_t)B * nCells);
    int rc = ORBF_OK;
    for (int f0 = 0; f0 < n && rc == ORBF_OK; f0 += B) {
        const int nb = std::min(B, n - f0);
        if (frame_stride == (int64_t)stride * h)
            AD_CUDA(cudaMemcpy2DAsync(dIn, pitch, gray + (size_t)f0 * frame_stride, stride, w, (size_t)h * nb, cudaMemcpyHostToDevice, c->stream));
        else
            for (int i = 0; i < nb; ++i)
                AD_CUDA(cudaMemcpy2DAsync(dIn + (size_t)i * pitch * h, pitch, gray + (size_t)(f0 + i) * frame_stride, stride, w, h, cudaMemcpyHostToDevice, c->stream));
        rc = orbf_tma_encode_u8(c, &P.map, dIn, w, h, nb, pitch, (long long)pitch * h, SP_BW, SP_BH);
        if (rc != ORBF_OK) break;
        P.nFrames = nb;
        // Floor of the response plane: one "too few" step below the lowest cell threshold covers almost every frame; a sub-batch
        // whose controllers go lower is redone three steps down, then at the minimum threshold (each redo costs more: the
        // lower the floor, the more pixels get a full corner score).
        const double minState = *std::min_element(st.begin(), st.end());
        const int floors[3] = { std::max((int)cfg->min_th, (int)(minState * cfg->dec)),
            std::max((int)cfg->min_th, (int)(minState * cfg->dec * cfg->dec * cfg->dec)), (int)cfg->min_th };
        for (int attempt = 0; attempt < 3; ++attempt) {
            if (attempt > 0 && floors[attempt] == floors[attempt - 1]) { if (attempt == 2) rc = ORBF_ERR_STATE; continue; }
            P.floorTh = std::max(floors[attempt], 1);
            const int zero = 0;
            AD_CUDA(cudaMemcpyAsync(dUnder, &zero, sizeof(int), cudaMemcpyHostToDevice, c->stream));
            AD_CUDA(cudaMemcpyAsync(dThresh, st.data(), nCells * sizeof(double), cudaMemcpyHostToDevice, c->stream));
            score_plane_kernel<<<dim3((w + SP_W - 1) / SP_W, (h + SP_H - 1) / SP_H, nb), SP_THREADS, 0, c->stream>>>(P);
            c->launches++;
            cell_collect_kernel<<<dim3(nCells, nb), CC_THREADS, 0, c->stream>>>(P);
            c->launches++;
            control_kernel<<<1, 32, 0, c->stream>>>(P);
            c->launches++;
            cell_select_kernel<<<dim3(nCells, nb), CS_THREADS, 0, c->stream>>>(P);
            c->launches++;
            int under = 0;
            AD_CUDA(cudaMemcpyAsync(&under, dUnder, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
            AD_CUDA(cudaStreamSynchronize(c->stream));
            AD_CUDA(cudaGetLastError());
            if (!under) { rc = ORBF_OK; break; }
            rc = ORBF_ERR_STATE;                              // stays only if even the minimum-threshold plane was not enough
        }
        if (rc != ORBF_OK) break;
        std::vector<double> stNew(nCells);
        AD_CUDA(cudaMemcpyAsync(stNew.data(), dThresh, nCells * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        AD_CUDA(cudaMemcpyAsync(hOut.data(), dOut, (size_t)nb * nCells * cfg->max_per_cell * sizeof(orbf_keypoint), cudaMemcpyDeviceToHost, c->stream));
        AD_CUDA(cudaMemcpyAsync(hCount.data(), dOutCount, (size_t)nb * nCells * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        AD_CUDA(cudaMemcpyAsync(hFinal.data(), dFinal, (size_t)nb * nCells * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        AD_CUDA(cudaMemcpyAsync(hFound.data(), dFound, (size_t)nb * nCells * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        AD_CUDA(cudaStreamSynchronize(c->stream));
        st = stNew;
        // aggregate in cell order (videogridadaptedfeaturedetector.cpp:33-50), then KeyPointsFilter::retainBest (extractor.cpp:45-46)
        for (int i = 0; i < nb; ++i) {
            std::vector<orbf_keypoint> all;
            for (int cl = 0; cl < nCells; ++cl) {
                const orbf_keypoint* src = hOut.data() + ((size_t)i * nCells + cl) * cfg->max_per_cell;
                all.insert(all.end(), src, src + hCount[(size_t)i * nCells + cl]);
                if (cell_thresh) cell_thresh[(size_t)(f0 + i) * nCells + cl] = hFinal[(size_t)i * nCells + cl];
                if (cell_found) cell_found[(size_t)(f0 + i) * nCells + cl] = hFound[(size_t)i * nCells + cl];
            }
            if (cfg->retain_best > 0 && (int)all.size() > cfg->retain_best) {
                std::vector<float> r(all.size());
                for (size_t k = 0; k < all.size(); ++k) r[k] = all[k].response;
                std::nth_element(r.begin(), r.begin() + (cfg->retain_best - 1), r.end(), std::greater<float>());
                const float cutv = r[cfg->retain_best - 1];
                std::vector<orbf_keypoint> keep;
                for (const orbf_keypoint& k : all) if (k.response >= cutv) keep.push_back(k);
                all.swap(keep);
            }
            counts[f0 + i] = (int)all.size();
            if ((int)all.size() > cap) { rc = ORBF_ERR_CAPACITY; continue; }
            if (out) std::copy(all.begin(), all.end(), out + (size_t)(f0 + i) * cap);
        }
    }
#undef AD_CUDA
    if (rc == ORBF_OK) std::copy(st.begin(), st.end(), thresh);
    return rc;
}
