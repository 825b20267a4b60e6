// csrc/describe.cu — per-keypoint stages, one warp per keypoint:
//   * intensity-centroid orientation (reference IC_Angle, Features/orbextractor.cpp:14-39) on the un-blurred
//     level: lanes cover u = -15..15 of each patch row, int32 moments, shuffle reduction, then cv::fastAtan2
//     restated with explicit round-to-nearest f32 ops (no FMA);
//   * steered rBRIEF-256 (computeOrbDescriptor, orbextractor.cpp:43-85) on the blurred level: lane k builds
//     descriptor byte k from 16 rotated test points (x*b + y*a, x*a - y*b with separate mul/add, cvRound =
//     round-half-even);
//   * output assembly (orbextractor.cpp:731-740, 805-811): pt += 16 (already in level coords), octave, size,
//     pt *= scale[level]; rows ordered level-major then quadtree list order;
//   * Frame::ExtractFeatures' depth gather + unprojection (Core/frame.cpp:148-164) into SoA x/y/z.
#include "orbf_internal.h"
#include "undistort_device.h"
#include "glibc_sincosf.h"

namespace {

// Device tables, filled once per device by orbf_launch_describe:
//   g_patF  : rBRIEF test t = 8 * byte + bit as floats (x0, y0, x1, y1) at index bit * 32 + byte (lane-major: conflict-free)
//   g_icCoef: umax[0..15], the half-width of the circular patch per |row| (= half-height per |column|: the disc is symmetric)
__device__ float4 g_patF[256];
__device__ uint32_t g_icCoef[16];
const int8_t h_pattern[1024] = {
#include "rbrief_pattern.inc"
};

struct DescParams {
    CUtensorMap mapRaw[ORBF_MAX_LEVELS], mapBlur[ORBF_MAX_LEVELS];    // 64 x 39 windows (raw level 0 = caller's plane: z = slot - z0)
    int z0;
    const uint32_t* lkp; const int* lkpCount;
    int kpStageTotal, K, slot0, L;
    int kpOff[ORBF_MAX_LEVELS]; float scale[ORBF_MAX_LEVELS]; int scaledPatch[ORBF_MAX_LEVELS];
    float *kpx, *kpy, *kpsize, *kpangle, *kpresp, *ptx, *pty, *ptz, *uright, *kpux, *kpuy;
    int distorted; UndistortParams und;                                  // k1 != 0: mvKeysUn = cv::undistortPoints(mvKeys) (Core/frame.cpp:286-313)
    int* kpoct; uint32_t* kplxy; uint8_t* desc; int* count;
    const uint16_t* depth; long long depthFrameStride; int depthPitch;   // elements
    int width, height;
    float cx, cy, invfx, invfy, mbf, depthFactor;
};

__device__ __forceinline__ float fast_atan2_deg(float y, float x)
{
    const float scale = (float)(180 / 3.1415926535897932384626433832795);
    const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale;
    const float p5 = 0.1555786518463281f * scale, p7 = -0.04432655554792128f * scale;
    const float eps = (float)2.2204460492503131e-16;
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, eps));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, eps));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
    return a;
}

constexpr int DS_WARPS = 4, DS_KPW = 16;     // keypoints per warp: window of keypoint k+1 is in flight while keypoint k is computed
constexpr int DS_WIN_BYTES = (ORBF_PATCH_BW * ORBF_PATCH_BH + 127) / 128 * 128;     // TMA destinations: 128-byte aligned

static_assert(DS_WIN_BYTES % 512 == 0 && ORBF_PATCH_BW == 64, "the descriptor window is loaded with CU_TENSOR_MAP_SWIZZLE_64B");

struct KpLoc { int level, x, y, score; };

// keypoint i of the frame (level-major, quadtree list order inside a level): its level and integer level coordinates
__device__ __forceinline__ KpLoc locate(const DescParams& P, const int* lc, int slot, int i, int& total)
{
    KpLoc k; k.level = -1; k.x = k.y = k.score = 0;
    int before = 0; total = 0;
    for (int l = 0; l < P.L; ++l) {
        const int c = lc[l];
        if (k.level < 0 && i < total + c) { k.level = l; before = total; }
        total += c;
    }
    if (k.level >= 0) {
        const uint32_t key = P.lkp[(long long)slot * P.kpStageTotal + P.kpOff[k.level] + (i - before)];
        k.x = (int)(key & 0x7FF) + ORBF_MINB; k.y = (int)((key >> 11) & 0x7FF) + ORBF_MINB; k.score = (int)(key >> 22);
    }
    return k;
}

constexpr int DS_RAW_BYTES = (ORBF_RAW_BW * ORBF_RAW_BH + 127) / 128 * 128;

__device__ __forceinline__ int dp4a_us(uint32_t a, uint32_t b, int c)          // sum of u8(a) * s8(b) + c
{
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// KPW keypoints per warp: 16 for batches (the fetch of keypoint k+1 hides behind keypoint k sixteen times per set-up), 4 for a
// call on one or two frames, where the length of a warp's chain is the kernel's duration
template <int KPW> __global__ void __launch_bounds__(DS_WARPS * 32) describe_kernel(const __grid_constant__ DescParams P)
{
    __shared__ float4 sPat[256];
    __shared__ uint32_t sCoef[16];
    // per warp, double buffered: the keypoint's 31 x 48 window of the raw level (orientation disc) and its 39 x 64 window of the
    // blurred level (descriptor), fetched by two TMA box loads — ~70 sectors instead of ~450 scattered byte gathers through L1
    __shared__ __align__(128) uint8_t sRaw[DS_WARPS][2][DS_RAW_BYTES];
    __shared__ __align__(512) uint8_t sBlur[DS_WARPS][2][DS_WIN_BYTES];    // 64B-swizzled by the TMA (512-byte swizzle atoms)
    __shared__ __align__(8) uint64_t sBar[DS_WARPS][2];
    for (int i = threadIdx.x; i < 256; i += DS_WARPS * 32) sPat[i] = g_patF[i];
    if (threadIdx.x < 16) sCoef[threadIdx.x] = g_icCoef[threadIdx.x];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { mbar_init(&sBar[warp][0], 1); mbar_init(&sBar[warp][1], 1); }
    __syncthreads();
    const int slot = P.slot0 + blockIdx.y;
    const int* lc = P.lkpCount + slot * ORBF_MAX_LEVELS;
    const int base = (blockIdx.x * DS_WARPS + warp) * KPW;
    // lanes 0..KPW-1 locate the warp's keypoints once; each iteration takes its keypoint by shuffle
    int total;
    const KpLoc mine = locate(P, lc, slot, base + (lane & (KPW - 1)), total);
    if (base == 0 && lane == 0) P.count[slot] = total;
    auto take = [&](int it) {
        KpLoc k;
        k.level = __shfl_sync(0xffffffffu, mine.level, it); k.x = __shfl_sync(0xffffffffu, mine.x, it);
        k.y = __shfl_sync(0xffffffffu, mine.y, it); k.score = __shfl_sync(0xffffffffu, mine.score, it);
        return k;
    };
    // orientation moments, lane = disc row v = lane - 15 (lane 31 idles): the row's 31 bytes against two constant coefficient
    // words per 4 columns — u (signed) and 1, both zero outside |u| <= umax[|v|] — so a row costs 16 DP4A
    uint32_t coefU[8], coefM[8];
    {
        const int v = lane - ORBF_HALF_PATCH, um = lane < 31 ? (int)sCoef[abs(v)] : -1;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            uint32_t cu = 0, cm = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int u = 4 * k + j - ORBF_HALF_PATCH;
                if (4 * k + j <= 2 * ORBF_HALF_PATCH && abs(u) <= um) { cu |= ((uint32_t)u & 0xFFu) << (8 * j); cm |= 1u << (8 * j); }
            }
            coefU[k] = cu; coefM[k] = cm;
        }
    }
    auto fetch = [&](const KpLoc& k, int buf) {       // lane 0 only
        // TMA boxes of bytes start on 16-byte boundaries
        mbar_expect_tx(&sBar[warp][buf], ORBF_RAW_BW * ORBF_RAW_BH + ORBF_PATCH_BW * ORBF_PATCH_BH);
        tma_load_3d(sRaw[warp][buf], &P.mapRaw[k.level], (k.x - ORBF_HALF_PATCH) & ~15, k.y - ORBF_HALF_PATCH, k.level == 0 ? slot - P.z0 : slot, &sBar[warp][buf]);
        tma_load_3d(sBlur[warp][buf], &P.mapBlur[k.level], (k.x - ORBF_EDGE) & ~15, k.y - ORBF_EDGE, slot, &sBar[warp][buf]);
    };
    KpLoc cur = take(0);
    if (cur.level >= 0 && lane == 0) fetch(cur, 0);
    const float kMagic = 12582912.f;
    const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
    // depth sample of the keypoint (Core/frame.cpp:155): lanes 0..KPW-1 request the samples of the warp's keypoints up front and
    // unproject them after the loop, so the read — an HBM access, or a ~2 us PCIe round trip when the plane lives in pinned host
    // memory — has the whole warp's work to hide behind
    float myX = (float)mine.x, myY = (float)mine.y;
    if (mine.level > 0) { myX = __fmul_rn(myX, P.scale[mine.level]); myY = __fmul_rn(myY, P.scale[mine.level]); }
    unsigned short myRaw = 0; bool myHave = false;
    if (lane < KPW && mine.level >= 0 && P.depth) {
        const int ui = (int)myX, vi = (int)myY;         // float -> int truncation of the (distorted) keypoint
        if (ui >= 0 && vi >= 0 && ui < P.width && vi < P.height) {
            myRaw = __ldg(P.depth + (long long)slot * P.depthFrameStride + (long long)vi * P.depthPitch + ui);
            myHave = true;
        }
    }
#pragma unroll 1
    for (int it = 0; it < KPW && cur.level >= 0; ++it) {
        const int i = base + it, buf = it & 1;
        KpLoc nxt; nxt.level = -1;
        if (it + 1 < KPW) nxt = take(it + 1);
        if (nxt.level >= 0 && lane == 0) fetch(nxt, buf ^ 1);
        __syncwarp();
        const int level = cur.level, x = cur.x, y = cur.y;
        float kfx = (float)x, kfy = (float)y;
        if (level != 0) { kfx = __fmul_rn(kfx, P.scale[level]); kfy = __fmul_rn(kfy, P.scale[level]); }
        const long long o = (long long)slot * P.K + i;
        mbar_wait(&sBar[warp][buf], (it >> 1) & 1);
        // ---- orientation -----------------------------------------------------------------------------------------------------
        int m10, m01;
        {
            const int off = (x - ORBF_HALF_PATCH) & 15;                   // column of u = -15 inside the raw window
            const uint32_t* row = reinterpret_cast<const uint32_t*>(sRaw[warp][buf] + min(lane, ORBF_RAW_BH - 1) * ORBF_RAW_BW) + (off >> 2);
            const int sh = (off & 3) * 8;
            uint32_t w[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) w[k] = row[k];                    // (off >> 2) + 9 <= 12 words: inside the row
            int su = 0, sm = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint32_t bytes = __funnelshift_r(w[k], w[k + 1], sh);
                su = dp4a_us(bytes, coefU[k], su);
                sm = (int)__dp4a(bytes, coefM[k], (uint32_t)sm);
            }
            m10 = su; m01 = (lane - ORBF_HALF_PATCH) * sm;
        }
#pragma unroll
        for (int o2 = 16; o2 > 0; o2 >>= 1) { m10 += __shfl_xor_sync(0xffffffffu, m10, o2); m01 += __shfl_xor_sync(0xffffffffu, m01, o2); }
        const float angle = fast_atan2_deg((float)m01, (float)m10);
        // ---- steered BRIEF: (x*b + y*a, x*a - y*b) with separate mul / add, cvRound = round-half-even via the 1.5 * 2^23 trick
        // (FADD on the FMA pipe + IADD instead of F2I on the quarter-rate XU pipe; |v| <= 19 << 2^22) ------------------------
        const float ar = __fmul_rn(angle, factorPI);
        const float a = replay::glibc_cosf(ar), b = replay::glibc_sinf(ar);     // std::cos(float) / std::sin(float) of the reference: libm's cosf / sinf
        const int cx0 = x - ((x - ORBF_EDGE) & ~15);                      // window column of the keypoint
        // The window is stored 64B-swizzled (chunk of 16 bytes ^= (row >> 1) & 3): with the plain 64-byte pitch all rows of one parity
        // share their banks and a warp-wide gather took ~4.8 wavefronts; swizzled, a column's rows spread over 8 bank groups.
        const uint8_t* wb = sBlur[warp][buf];
        auto at = [&](int r, int cc) { const int row = r + ORBF_EDGE, col = cc + cx0; return (int)wb[row * ORBF_PATCH_BW + (col ^ ((row & 6) << 3))]; };
        int val = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float4 w = sPat[k * 32 + lane];
            const int r0 = __float_as_int(__fadd_rn(__fadd_rn(__fmul_rn(w.x, b), __fmul_rn(w.y, a)), kMagic)) - 0x4B400000;
            const int c0 = __float_as_int(__fadd_rn(__fsub_rn(__fmul_rn(w.x, a), __fmul_rn(w.y, b)), kMagic)) - 0x4B400000;
            const int r1 = __float_as_int(__fadd_rn(__fadd_rn(__fmul_rn(w.z, b), __fmul_rn(w.w, a)), kMagic)) - 0x4B400000;
            const int c1 = __float_as_int(__fadd_rn(__fsub_rn(__fmul_rn(w.z, a), __fmul_rn(w.w, b)), kMagic)) - 0x4B400000;
            const int t0 = at(r0, c0), t1 = at(r1, c1);
            val |= (t0 < t1) << k;
        }
        P.desc[o * 32 + lane] = (uint8_t)val;
        // ---- keypoint record (the 3D points follow after the loop) -------------------------------------------------------
        if (lane == 0) {
            P.kpx[o] = kfx; P.kpy[o] = kfy; P.kpsize[o] = (float)P.scaledPatch[level]; P.kpangle[o] = angle;
            P.kpresp[o] = (float)cur.score; P.kpoct[o] = level; P.kplxy[o] = (uint32_t)x | ((uint32_t)y << 16);
        }
        __syncwarp();                                   // every lane is done with this buffer before it is refilled
        cur = nxt;
    }
    if (lane < KPW && mine.level >= 0) {
        // mvKeysUn (frame.cpp:286-313): the depth was looked up at the distorted keypoint above, mvuRight and mvKeys3Dc use the undistorted one
        float uX = myX, uY = myY;
        if (P.distorted) undistort_point(P.und, myX, myY, &uX, &uY);
        float X = 0.f, Y = 0.f, Z = 0.f, ur = -1.f;
        if (myHave) {
            const float z = __fmul_rn((float)myRaw, P.depthFactor);
            if (z > 0) {
                ur = __fsub_rn(uX, __fdiv_rn(P.mbf, z));
                X = __fmul_rn(__fmul_rn(__fsub_rn(uX, P.cx), z), P.invfx);
                Y = __fmul_rn(__fmul_rn(__fsub_rn(uY, P.cy), z), P.invfy);
                Z = z;
            }
        }
        const long long o = (long long)slot * P.K + base + lane;
        P.ptx[o] = X; P.pty[o] = Y; P.ptz[o] = Z; P.uright[o] = ur; P.kpux[o] = uX; P.kpuy[o] = uY;
    }
}

__global__ void pack_aos_kernel(const float* kpx, const float* kpy, const float* kpsize, const float* kpangle, const float* kpresp,
    const int* kpoct, const int* count, orbf_keypoint* out, int K, int slot0)
{
    const int slot = slot0 + blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count[slot]) return;
    const long long o = (long long)slot * K + i;
    orbf_keypoint k;
    k.x = kpx[o]; k.y = kpy[o]; k.size = kpsize[o]; k.angle = kpangle[o]; k.response = kpresp[o]; k.octave = kpoct[o]; k.class_id = -1;
    out[o] = k;
}

bool g_constReady[64] = { false };

}  // namespace

int orbf_launch_describe(orbf_context* c, int slot0, int n)
{
    const int dev = c->cfg.device;
    if (dev < 64 && !g_constReady[dev]) {
        float4 patF[256];
        for (int t = 0; t < 256; ++t)
            patF[(t & 7) * 32 + (t >> 3)] = make_float4((float)h_pattern[4 * t], (float)h_pattern[4 * t + 1], (float)h_pattern[4 * t + 2], (float)h_pattern[4 * t + 3]);
        uint32_t coef[16];
        for (int v = 0; v < 16; ++v) coef[v] = (uint32_t)c->umax[v];
        ORBF_CUDA(c, cudaMemcpyToSymbol(g_patF, patF, sizeof(patF)));
        ORBF_CUDA(c, cudaMemcpyToSymbol(g_icCoef, coef, sizeof(coef)));
        g_constReady[dev] = true;
    }
    {
        const int r = orbf_refresh_maps(c);
        if (r != ORBF_OK) return r;
    }
    DescParams P;
    for (int l = 0; l < c->L; ++l) { P.mapRaw[l] = c->tmPatchRaw[l]; P.mapBlur[l] = c->tmPatchBlur[l]; }
    P.z0 = c->cur_slot0;
    P.lkp = c->d_lkp; P.lkpCount = c->d_lkpCount; P.kpStageTotal = c->kpStageTotal; P.K = c->K; P.slot0 = slot0; P.L = c->L;
    for (int l = 0; l < c->L; ++l) { P.kpOff[l] = c->lg[l].kpOff; P.scale[l] = c->scale[l]; P.scaledPatch[l] = c->lg[l].scaledPatch; }
    P.kpx = c->d_kpx; P.kpy = c->d_kpy; P.kpsize = c->d_kpsize; P.kpangle = c->d_kpangle; P.kpresp = c->d_kpresp;
    P.ptx = c->d_ptx; P.pty = c->d_pty; P.ptz = c->d_ptz; P.uright = c->d_uright; P.kpux = c->d_kpux; P.kpuy = c->d_kpuy;
    P.distorted = c->cfg.k1 != 0.0f ? 1 : 0;
    P.und = UndistortParams{ (double)c->cfg.fx, (double)c->cfg.fy, (double)c->cfg.cx, (double)c->cfg.cy, (double)c->cfg.k1, (double)c->cfg.k2,
        (double)c->cfg.p1, (double)c->cfg.p2, (double)c->cfg.k3 };
    P.kpoct = c->d_kpoct; P.kplxy = c->d_kplxy; P.desc = c->d_desc; P.count = c->d_count;
    if (c->cur_depth) {
        P.depth = c->cur_depth - (long long)c->cur_slot0 * c->cur_depthFrameStride;
        P.depthFrameStride = c->cur_depthFrameStride; P.depthPitch = c->cur_depthPitch;
    } else { P.depth = nullptr; P.depthFrameStride = 0; P.depthPitch = 0; }
    P.width = c->cfg.width; P.height = c->cfg.height;
    P.cx = c->cfg.cx; P.cy = c->cfg.cy; P.invfx = 1.0f / c->cfg.fx; P.invfy = 1.0f / c->cfg.fy;
    P.mbf = c->cfg.mbf; P.depthFactor = c->cfg.depth_factor;
    if (n <= 2) {
        constexpr int KPW = 4;
        describe_kernel<KPW><<<dim3((c->K + DS_WARPS * KPW - 1) / (DS_WARPS * KPW), n), DS_WARPS * 32, 0, c->stream>>>(P);
    } else {
        describe_kernel<DS_KPW><<<dim3((c->K + DS_WARPS * DS_KPW - 1) / (DS_WARPS * DS_KPW), n), DS_WARPS * 32, 0, c->stream>>>(P);
    }
    ORBF_LAUNCH_CHECK(c);
    return ORBF_OK;
}

int orbf_launch_pack_aos(orbf_context* c, int slot0, int n)
{
    dim3 grid((c->K + 127) / 128, n);
    pack_aos_kernel<<<grid, 128, 0, c->stream>>>(c->d_kpx, c->d_kpy, c->d_kpsize, c->d_kpangle, c->d_kpresp, c->d_kpoct, c->d_count,
        c->d_kpAos, c->K, slot0);
    ORBF_LAUNCH_CHECK(c);
    return ORBF_OK;
}
