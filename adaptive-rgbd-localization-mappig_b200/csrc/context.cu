// csrc/context.cu — context lifecycle, extractor tables and device-memory layout.
// Tables follow ORBextractor::ORBextractor (reference Features/orbextractor.cpp:346-404), the level
// sizes ComputePyramid (:833-838), the FAST cell grid ComputeKeyPointsOctTree (:665-703), the quadtree
// roots DistributeOctTree (:470-489) and the resize coefficients cv::resize(INTER_LINEAR, 8U) as
// called at :846 (OpenCV fixed-point bilinear, 11-bit coefficients).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>

#include "orbf_internal.h"

int orbf_cuda_fail(orbf_context* ctx, cudaError_t e, const char* what, const char* file, int line)
{
    char buf[512];
    snprintf(buf, sizeof(buf), "%s: %s (%s:%d)", what, cudaGetErrorString(e), file, line);
    if (ctx) ctx->lastError = buf;
    return ORBF_ERR_CUDA;
}

static inline int cv_round_f(float v) { return (int)nearbyintf(v); }
static inline int cv_round_d(double v) { return (int)nearbyint(v); }

// ---- TMA descriptor encoding (driver entry point fetched through the runtime: no link-time libcuda dependency) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int orbf_tma_encode_u8(orbf_context* ctx, CUtensorMap* out, const void* base, int w, int h, int frames, long long pitch,
    long long frameStride, int boxW, int boxH, bool swizzle64)
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
        if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p)
            return orbf_cuda_fail(ctx, e != cudaSuccess ? e : cudaErrorNotSupported, "cuTensorMapEncodeTiled lookup", __FILE__, __LINE__);
        fn = (EncodeTiledFn)p;
    }
    if (((uintptr_t)base & 15) || (pitch & 15) || (frameStride & 15) || (boxW & 15) || boxW > 256 || boxH > 256) return ORBF_ERR_ALIGNMENT;
    if (swizzle64 && boxW != 64) return ORBF_ERR_ALIGNMENT;
    const cuuint64_t gdim[3] = { (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)frames };
    const cuuint64_t gstride[2] = { (cuuint64_t)pitch, (cuuint64_t)frameStride };
    const cuuint32_t box[3] = { (cuuint32_t)boxW, (cuuint32_t)boxH, 1u };
    const cuuint32_t estr[3] = { 1u, 1u, 1u };
    const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(base), gdim, gstride, box, estr,
        CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char buf[96];
        snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
        if (ctx) ctx->lastError = buf;
        return ORBF_ERR_CUDA;
    }
    return ORBF_OK;
}

static void build_resize_tab(int src, int dst, ResizeCoef* out)
{
    const double inv_scale = (double)dst / src;
    const double scale = 1. / inv_scale;
    for (int d = 0; d < dst; ++d) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = (int)floorf(f);
        f -= s;
        if (s < 0) { f = 0; s = 0; }
        if (s >= src - 1) { f = 0; s = src - 1; }
        out[d].ofs = (short)s;
        out[d].a0 = (short)cv_round_f((1.f - f) * 2048.f);
        out[d].a1 = (short)cv_round_f(f * 2048.f);
        out[d].pad = 0;
    }
}

template <typename T>
static int dalloc(orbf_context* ctx, T** p, size_t count)
{
    *p = nullptr;
    if (count == 0) count = 1;
    ORBF_CUDA(ctx, cudaMalloc((void**)p, count * sizeof(T)));
    return ORBF_OK;
}

extern "C" int orbf_abi_version(void) { return ORBF_ABI_VERSION; }

extern "C" void orbf_default_config(orbf_config* c)
{
    if (!c) return;
    memset(c, 0, sizeof(*c));
    c->width = 640; c->height = 480;
    c->nfeatures = 1000; c->nlevels = 8; c->scale_factor = 1.2f;   // Utils/common.h:77, Features/extractor.cpp:86
    c->ini_th_fast = 20; c->min_th_fast = 7;
    c->max_frames = 1; c->max_pairs = 0; c->device = 0;
    c->fx = 517.3f; c->fy = 516.5f; c->cx = 318.6f; c->cy = 255.3f;   // Utils/common.h:35-38 (FR1)
    c->mbf = 40.0f;
    c->depth_factor = 1.0f / 5000.0f;
    c->pipeline_chunk = 0; c->pipeline_streams = 0; c->depth_zero_copy = 0; c->pipeline_overlap = 0;
    c->k1 = c->k2 = c->p1 = c->p2 = c->k3 = 0.0f;               // the synthetic configurations zero the distortion (frame.cpp:288-291)
}

extern "C" void orbf_default_ransac_config(orbf_ransac_config* c)
{
    if (!c) return;
    memset(c, 0, sizeof(*c));
    c->iterations = 200; c->min_inlier_th = 20; c->max_mahal = 3.0f; c->sample_size = 4;   // Odometry/ransac.cpp:9-12
    c->check_depth = 1; c->sort_mode = 0; c->depth_cov = -1.0; c->seed = 42;
}

extern "C" const char* orbf_status_string(int s)
{
    switch (s) {
    case ORBF_OK: return "ok";
    case ORBF_ERR_ARG: return "bad argument";
    case ORBF_ERR_CAPACITY: return "output capacity too small";
    case ORBF_ERR_GEOMETRY: return "image too small for pyramid / patch borders";
    case ORBF_ERR_CUDA: return "CUDA error (see orbf_last_error)";
    case ORBF_ERR_ALIGNMENT: return "device pointer or pitch not 16-byte aligned";
    case ORBF_ERR_STATE: return "call order violated";
    default: return "unknown status";
    }
}

extern "C" const char* orbf_last_error(const orbf_context* ctx) { return ctx ? ctx->lastError.c_str() : "null context"; }

static int build_geometry(orbf_context* c, std::vector<ResizeCoef>& tab, std::vector<CellDesc>& cells, std::vector<StripDesc>& strips,
    std::vector<TileDesc>& blTiles, std::vector<TileDesc>& rsTiles)
{
    int tileW, blurH, blurBW, blurBH, resizeH;
    orbf_stage_tile_geometry(&tileW, &blurH, &blurBW, &blurBH, &resizeH);
    const orbf_config& g = c->cfg;
    const int L = g.nlevels;
    c->L = L;
    const double scaleFactor = (double)g.scale_factor;
    c->scale[0] = 1.0f; c->sigma2[0] = 1.0f;
    for (int i = 1; i < L; ++i) {
        c->scale[i] = (float)((double)c->scale[i - 1] * scaleFactor);
        c->sigma2[i] = c->scale[i] * c->scale[i];
    }
    for (int i = 0; i < L; ++i) { c->invScale[i] = 1.0f / c->scale[i]; c->invSigma2[i] = 1.0f / c->sigma2[i]; }
    const float factor = (float)(1.0 / scaleFactor);
    float desired = (float)g.nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)L));
    int sum = 0;
    for (int l = 0; l < L - 1; ++l) {
        c->lg[l].nfeat = cv_round_f(desired);
        sum += c->lg[l].nfeat;
        desired *= factor;
    }
    c->lg[L - 1].nfeat = std::max(g.nfeatures - sum, 0);
    {   // umax (orbextractor.cpp:389-403)
        int v, v0;
        const int vmax = (int)floor(ORBF_HALF_PATCH * sqrtf(2.f) / 2 + 1);
        const int vmin = (int)ceil(ORBF_HALF_PATCH * sqrtf(2.f) / 2);
        const double hp2 = ORBF_HALF_PATCH * ORBF_HALF_PATCH;
        for (v = 0; v < 16; ++v) c->umax[v] = 0;
        for (v = 0; v <= vmax; ++v) c->umax[v] = cv_round_d(sqrt(hp2 - v * v));
        for (v = ORBF_HALF_PATCH, v0 = 0; v >= vmin; --v) {
            while (c->umax[v0] == c->umax[v0 + 1]) ++v0;
            c->umax[v] = v0;
            ++v0;
        }
    }
    int cellSlot = 0, candOff = 0, kpOff = 0;
    c->maxCellW = c->maxCellH = 0;
    for (int l = 0; l < L; ++l) {
        LevelGeom& q = c->lg[l];
        q.w = cv_round_f((float)g.width * c->invScale[l]);
        q.h = cv_round_f((float)g.height * c->invScale[l]);
        if (q.w < 2 * ORBF_EDGE + 8 || q.h < 2 * ORBF_EDGE + 8 || q.w > 2047 + 2 * ORBF_MINB || q.h > 2047 + 2 * ORBF_MINB)
            return ORBF_ERR_GEOMETRY;
        q.pitch = align_up(q.w, 128);
        q.plane = (size_t)q.pitch * q.h;
        q.scale = c->scale[l];
        q.scaledPatch = (int)(31 * c->scale[l]);
        if (l > 0) {
            q.tabX = (int)tab.size(); tab.resize(tab.size() + q.w);
            build_resize_tab(c->lg[l - 1].w, q.w, &tab[q.tabX]);
            q.tabY = (int)tab.size(); tab.resize(tab.size() + q.h);
            build_resize_tab(c->lg[l - 1].h, q.h, &tab[q.tabY]);
        } else q.tabX = q.tabY = 0;
        // destination tiles of the blur (every level) and resize (levels >= 1) stages, and the resize source box
        for (int y0 = 0; y0 < q.h; y0 += blurH)
            for (int x0 = 0; x0 < q.w; x0 += tileW) { TileDesc t = { (short)l, (short)x0, (short)y0, 0 }; blTiles.push_back(t); }
        c->rsTile0[l] = (int)rsTiles.size(); c->rsBW[l] = c->rsBH[l] = 16;
        if (l > 0) {
            const ResizeCoef* tx = &tab[q.tabX]; const ResizeCoef* ty = &tab[q.tabY];
            for (int y0 = 0; y0 < q.h; y0 += resizeH)
                for (int x0 = 0; x0 < q.w; x0 += tileW) {
                    TileDesc t = { (short)l, (short)x0, (short)y0, 0 };
                    rsTiles.push_back(t);
                    const int xs = tx[x0].ofs & ~15, xe = tx[std::min(x0 + tileW, q.w) - 1].ofs + 2;
                    const int ys = ty[y0].ofs, ye = ty[std::min(y0 + resizeH, q.h) - 1].ofs + 2;
                    c->rsBW[l] = std::max(c->rsBW[l], align_up(xe - xs, 16)); c->rsBH[l] = std::max(c->rsBH[l], ye - ys);
                }
            if (c->rsBW[l] > 256 || c->rsBH[l] > 256) return ORBF_ERR_GEOMETRY;
        }
        c->rsTileN[l] = (int)rsTiles.size() - c->rsTile0[l];
        // FAST cell grid
        const int minB = ORBF_MINB, maxBX = q.w - ORBF_EDGE + 3, maxBY = q.h - ORBF_EDGE + 3;
        const float W = 30;
        const float width = (float)(maxBX - minB), height = (float)(maxBY - minB);
        q.cellsX = (int)(width / W); q.cellsY = (int)(height / W);
        if (q.cellsX < 1 || q.cellsY < 1) return ORBF_ERR_GEOMETRY;
        q.wCell = (int)ceilf(width / q.cellsX); q.hCell = (int)ceilf(height / q.cellsY);
        q.cell0 = (int)cells.size();
        q.candOff = candOff;
        int cap = 0, levelMaxW = 0, levelMaxH = 0;
        for (int i = 0; i < q.cellsY; ++i) {
            int inRow = 0;
            const float iniY = (float)(minB + i * q.hCell);
            float maxY = iniY + q.hCell + 6;
            if (iniY >= maxBY - 3) continue;
            if (maxY > maxBY) maxY = (float)maxBY;
            for (int j = 0; j < q.cellsX; ++j) {
                const float iniX = (float)(minB + j * q.wCell);
                float maxX = iniX + q.wCell + 6;
                if (iniX >= maxBX - 6) continue;
                if (maxX > maxBX) maxX = (float)maxBX;
                const int cw = (int)maxX - (int)iniX - 6, ch = (int)maxY - (int)iniY - 6;
                if (cw <= 0 || ch <= 0) continue;   // cv::FAST scores nothing on a ROI thinner than 7
                CellDesc d;
                d.level = (short)l; d.x0 = (short)((int)iniX + 3); d.y0 = (short)((int)iniY + 3);
                d.w = (short)cw; d.h = (short)ch;
                d.relx = (short)(-minB); d.rely = (short)(-minB); d.pad = 0;
                d.slotOff = cellSlot;
                d.cap = ((cw + 1) / 2) * ((ch + 1) / 2);   // strict 8-neighbour maxima: <= 1 per 2x2 block
                cellSlot += d.cap; cap += d.cap;
                c->maxCellW = std::max(c->maxCellW, cw); c->maxCellH = std::max(c->maxCellH, ch);
                // strips of up to ORBF_STRIP_CELLS adjacent cells (their scored interiors tile the row without gaps), no wider than
                // the fixed tile pitch allows
                if (cw > ORBF_STRIP_MAX_W) return ORBF_ERR_GEOMETRY;
                if (inRow == 0 || strips.back().nCells >= ORBF_STRIP_CELLS || d.x0 + d.w - strips.back().x0 > ORBF_STRIP_MAX_W) {
                    StripDesc sd;
                    sd.level = (short)l; sd.nCells = 0; sd.x0 = d.x0; sd.y0 = d.y0; sd.w = 0; sd.h = d.h; sd.firstCell = (int)cells.size();
                    strips.push_back(sd);
                }
                StripDesc& sd = strips.back();
                sd.nCells++; sd.w = (short)(d.x0 + d.w - sd.x0);
                levelMaxW = std::max(levelMaxW, (int)sd.w); levelMaxH = std::max(levelMaxH, ch);
                ++inRow;
                cells.push_back(d);
            }
        }
        q.nCells = (int)cells.size() - q.cell0;
        c->fastBW[l] = ORBF_FAST_BW;                             // box starts on a 16-byte boundary <= x0 - 3, ends >= 3 px past the strip
        (void)levelMaxW;
        c->fastBH[l] = levelMaxH + 6;
        if (c->fastBW[l] > 256 || c->fastBH[l] > 256) return ORBF_ERR_GEOMETRY;
        q.candCap = cap;
        candOff += cap;
        // quadtree roots (orbextractor.cpp:470-472)
        q.nIni = (int)roundf((float)(maxBX - minB) / (float)(maxBY - minB));
        if (q.nIni < 1) return ORBF_ERR_GEOMETRY;
        q.hX = (float)(maxBX - minB) / q.nIni;
        q.kpCap = std::max(4 * q.nIni, q.nfeat + 3);
        q.kpOff = kpOff;
        kpOff += q.kpCap;
    }
    c->nCellsTotal = (int)cells.size();
    c->cellSlotTotal = cellSlot;
    c->candTotal = candOff;
    c->kpStageTotal = kpOff;
    c->K = align_up(kpOff, 32);
    return ORBF_OK;
}

extern "C" int orbf_create(const orbf_config* cfg, orbf_context** out)
{
    if (!cfg || !out) return ORBF_ERR_ARG;
    *out = nullptr;
    if (cfg->nlevels < 1 || cfg->nlevels > ORBF_MAX_LEVELS || cfg->width < 1 || cfg->height < 1 || cfg->nfeatures < 1
        || cfg->max_frames < 1 || !(cfg->scale_factor > 1.0f) || cfg->ini_th_fast < 1 || cfg->min_th_fast < 1
        || cfg->ini_th_fast > 254 || cfg->min_th_fast > cfg->ini_th_fast)
        return ORBF_ERR_ARG;
    orbf_context* c = new (std::nothrow) orbf_context();
    if (!c) return ORBF_ERR_ARG;
    c->cfg = *cfg;
    c->B = cfg->max_frames;
    c->P = cfg->max_pairs > 0 ? cfg->max_pairs : cfg->max_frames;
    c->launches = 0; c->stream = nullptr; c->ownStream = false; c->profiling = false;
    c->nWork = 0; c->evFork = nullptr; c->hi = nullptr; c->evHiA = c->evHiB = nullptr; c->evRansacIn = c->evRansac = nullptr; c->hiPending = false; c->hiSlot0 = 0; c->hiN = 0; c->hiPair0 = 0; c->hiNPairs = 0;
    for (int i = 0; i < 8; ++i) c->evHiGroup[i] = nullptr;
    for (int i = 0; i < ORBF_MAX_WORKERS; ++i) { c->work[i] = nullptr; c->evDone[i] = nullptr; c->evExtract[i] = nullptr; }
    c->copy = nullptr;
    for (int i = 0; i < ORBF_MAX_CHUNKS; ++i) c->evCopy[i] = nullptr;
    for (int i = 0; i < ORBF_MARKERS; ++i) c->evMarker[i] = nullptr;
    c->chunkFrames = cfg->pipeline_chunk == 0 ? (cfg->pipeline_overlap ? 256 : 64) : (cfg->pipeline_chunk < 0 ? 0 : std::max(cfg->pipeline_chunk, 2));
    for (int i = 0; i < ST_COUNT; ++i) { c->evA[i] = c->evB[i] = nullptr; c->evPending[i] = false; c->stageMs[i] = 0; c->stageCalls[i] = 0; }
    c->hypCap = 0; c->descStageRows = 0; c->xyzStageRows = 0; c->kfCap = 0; c->lastNPairs = 0; c->pairsFromSlots = false;
    c->cur_gray = nullptr; c->cur_depth = nullptr; c->cur_slot0 = 0; c->cur_n = 0;
    std::vector<ResizeCoef> tab; std::vector<CellDesc> cells; std::vector<StripDesc> strips; std::vector<TileDesc> blTiles, rsTiles;
    c->tmStaticReady = false; c->tm0Base = nullptr; c->tm0Pitch = c->tm0FrameStride = 0; c->tm0Frames = 0;
    int rc = build_geometry(c, tab, cells, strips, blTiles, rsTiles);
    c->nStrips = (int)strips.size(); c->nBlTiles = (int)blTiles.size();
    c->h_cells = cells;
    c->d_cellRegion = nullptr; c->cellRegionGrid = 0; c->d_regionTh = nullptr; c->d_regionState = nullptr; c->d_regionLog = nullptr; c->regionLogCap = 0; c->regionVideos = 0;
    if (rc != ORBF_OK) { delete c; return rc; }

    cudaError_t e = cudaSetDevice(cfg->device);
    if (e != cudaSuccess) { delete c; return ORBF_ERR_CUDA; }   // no CPU fallback: fail loudly
    auto fail = [&](int code) { *out = c; return code; };       // caller may read orbf_last_error, then destroy
    {
        cudaError_t e2 = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
        if (e2 != cudaSuccess) { orbf_cuda_fail(c, e2, "cudaStreamCreate", __FILE__, __LINE__); return fail(ORBF_ERR_CUDA); }
        c->ownStream = true;
        const int nw = cfg->pipeline_streams <= 0 ? 4 : std::min(cfg->pipeline_streams, ORBF_MAX_WORKERS);
        auto ev = [&](cudaEvent_t* e) { return cudaEventCreateWithFlags(e, cudaEventDisableTiming) == cudaSuccess; };
        bool ok = ev(&c->evFork) && cudaStreamCreateWithFlags(&c->copy, cudaStreamNonBlocking) == cudaSuccess;
        for (int i = 0; ok && i < ORBF_MAX_CHUNKS; ++i) ok = ev(&c->evCopy[i]);
        for (int i = 0; ok && i < ORBF_MARKERS; ++i) ok = ev(&c->evMarker[i]);
        for (int i = 0; ok && i < nw; ++i) {
            ok = cudaStreamCreateWithFlags(&c->work[i], cudaStreamNonBlocking) == cudaSuccess && ev(&c->evDone[i]) && ev(&c->evExtract[i]);
            if (ok) c->nWork = i + 1;
        }
        int prLo = 0, prHi = 0;
        ok = ok && cudaDeviceGetStreamPriorityRange(&prLo, &prHi) == cudaSuccess
            && cudaStreamCreateWithPriority(&c->hi, cudaStreamNonBlocking, prHi) == cudaSuccess && ev(&c->evHiA) && ev(&c->evHiB)
            && ev(&c->evRansacIn) && ev(&c->evRansac);
        for (int i = 0; ok && i < 8; ++i) ok = ev(&c->evHiGroup[i]);
        if (!ok) { orbf_cuda_fail(c, cudaGetLastError(), "worker streams", __FILE__, __LINE__); return fail(ORBF_ERR_CUDA); }
    }
    const size_t B = c->B, K = c->K, P = c->P;
    c->inPitch = align_up(cfg->width, 128);
    c->inPlane = (size_t)c->inPitch * cfg->height;
#define TRY(x) do { int r__ = (x); if (r__ != ORBF_OK) return fail(r__); } while (0)
    TRY(dalloc(c, &c->d_in, B * c->inPlane));
    TRY(dalloc(c, &c->d_depthIn, B * (size_t)cfg->width * cfg->height));
    c->d_pyr[0] = nullptr;
    for (int l = 0; l < c->L; ++l) {
        if (l > 0) TRY(dalloc(c, &c->d_pyr[l], B * c->lg[l].plane));
        TRY(dalloc(c, &c->d_blur[l], B * c->lg[l].plane));
    }
    TRY(dalloc(c, &c->d_resizeTab, tab.size()));
    TRY(dalloc(c, &c->d_cells, cells.size()));
    TRY(dalloc(c, &c->d_strips, strips.size()));
    TRY(dalloc(c, &c->d_blTiles, blTiles.size()));
    TRY(dalloc(c, &c->d_rsTiles, rsTiles.size()));
    TRY(dalloc(c, &c->d_lg, (size_t)ORBF_MAX_LEVELS));
    TRY(dalloc(c, &c->d_cellCand, B * c->cellSlotTotal));
    TRY(dalloc(c, &c->d_cellCount, B * c->nCellsTotal));
    TRY(dalloc(c, &c->d_cand, B * c->candTotal));
    TRY(dalloc(c, &c->d_candCount, B * ORBF_MAX_LEVELS));
    TRY(dalloc(c, &c->d_nodeScratch, B * c->candTotal));
    TRY(dalloc(c, &c->d_lkp, B * c->kpStageTotal));
    TRY(dalloc(c, &c->d_lkpCount, B * ORBF_MAX_LEVELS));
    TRY(dalloc(c, &c->d_kpx, B * K)); TRY(dalloc(c, &c->d_kpy, B * K)); TRY(dalloc(c, &c->d_kpsize, B * K));
    TRY(dalloc(c, &c->d_kpangle, B * K)); TRY(dalloc(c, &c->d_kpresp, B * K));
    TRY(dalloc(c, &c->d_ptx, B * K)); TRY(dalloc(c, &c->d_pty, B * K)); TRY(dalloc(c, &c->d_ptz, B * K));
    TRY(dalloc(c, &c->d_uright, B * K)); TRY(dalloc(c, &c->d_kpux, B * K)); TRY(dalloc(c, &c->d_kpuy, B * K));
    TRY(dalloc(c, &c->d_kpoct, B * K)); TRY(dalloc(c, &c->d_kplxy, B * K));
    TRY(dalloc(c, &c->d_desc, B * K * 32));
    TRY(dalloc(c, &c->d_count, B));
    TRY(dalloc(c, &c->d_kpAos, B * K));
    TRY(dalloc(c, &c->d_pairs, P * 2));
    TRY(dalloc(c, &c->d_knn, P * K * 2));
    TRY(dalloc(c, &c->d_rev, P * K));
    TRY(dalloc(c, &c->d_matches, P * K));
    TRY(dalloc(c, &c->d_matchCount, P));
    TRY(dalloc(c, &c->d_good, P * K));
    TRY(dalloc(c, &c->d_goodCount, P));
    TRY(dalloc(c, &c->d_rres, P));
    { uint8_t* t = nullptr; TRY(dalloc(c, &t, P * 32)); c->d_rstate = t; }
    TRY(dalloc(c, &c->d_inliers, P * K));
    TRY(dalloc(c, &c->d_depthCov, 2));
    c->d_samples = nullptr; c->d_hyp = nullptr;
    c->d_qdesc = c->d_tdesc = nullptr; c->d_sxyz = c->d_txyz = nullptr;
    c->d_bgr = nullptr; c->bgrSlots = 0;
    c->d_kfDesc = nullptr; c->d_kfCount = nullptr; c->d_kfExtDesc = nullptr; c->d_kfExtCount = nullptr; c->kfExtN = 0;
    c->d_pts = nullptr; c->ptsCap = 0; c->d_userSamples = nullptr; c->userSamplesCap = 0; c->d_cloudSrc = c->d_cloudTgt = nullptr; c->d_cloudCount = nullptr; c->cloudCap = 0;
    memset(&c->lastRs, 0, sizeof(c->lastRs)); memset(&c->lastRansacCfg, 0, sizeof(c->lastRansacCfg)); c->d_kabsch = nullptr; c->kabschCap = 0;
    c->d_kfKnn = nullptr; c->d_kfSurv = nullptr; c->d_kfPairs = nullptr; c->d_kfQCount = nullptr; c->kfOutCap = 0;
    c->ncclComm = nullptr; c->commRanks = 1; c->commRank = 0; c->d_kfGather = nullptr; c->d_kfGatherCount = nullptr; c->kfGatherCap = 0;
    c->d_peerDesc = nullptr; c->d_peerCount = nullptr; c->nPeers = 0; c->peerKf = 0;
    for (void*& q : c->peerOpened) q = nullptr;
    c->h_kp = nullptr; c->h_desc = nullptr; c->h_xyz = nullptr; c->h_counts = nullptr;
    c->h_arena = nullptr; c->arenaCap = 0; c->arenaUsed = 0; c->evArena = nullptr; c->arenaBusy = false;
    c->d_scratch = nullptr; c->scratchCap = 0; c->pendDepthSrc = nullptr; c->pendDepthDst = nullptr; c->pendDepthStride = 0;
    auto cu = [&](cudaError_t e3, const char* w) { if (e3 != cudaSuccess) { orbf_cuda_fail(c, e3, w, __FILE__, __LINE__); return false; } return true; };
    if (!cu(cudaMemcpy(c->d_resizeTab, tab.data(), tab.size() * sizeof(ResizeCoef), cudaMemcpyHostToDevice), "tab")) return fail(ORBF_ERR_CUDA);
    if (!cu(cudaMemcpy(c->d_cells, cells.data(), cells.size() * sizeof(CellDesc), cudaMemcpyHostToDevice), "cells")) return fail(ORBF_ERR_CUDA);
    if (!cu(cudaMemcpy(c->d_strips, strips.data(), strips.size() * sizeof(StripDesc), cudaMemcpyHostToDevice), "strips")) return fail(ORBF_ERR_CUDA);
    if (!cu(cudaMemcpy(c->d_blTiles, blTiles.data(), blTiles.size() * sizeof(TileDesc), cudaMemcpyHostToDevice), "blur tiles")) return fail(ORBF_ERR_CUDA);
    if (!rsTiles.empty() && !cu(cudaMemcpy(c->d_rsTiles, rsTiles.data(), rsTiles.size() * sizeof(TileDesc), cudaMemcpyHostToDevice), "resize tiles"))
        return fail(ORBF_ERR_CUDA);
    if (!cu(cudaMemcpy(c->d_lg, c->lg, sizeof(LevelGeom) * ORBF_MAX_LEVELS, cudaMemcpyHostToDevice), "lg")) return fail(ORBF_ERR_CUDA);
    { const double neg[2] = { -1.0, -1.0 }; if (!cu(cudaMemcpy(c->d_depthCov, neg, sizeof(neg), cudaMemcpyHostToDevice), "depthCov")) return fail(ORBF_ERR_CUDA); }
    if (!cu(cudaMemset(c->d_count, 0, B * sizeof(int)), "memset")) return fail(ORBF_ERR_CUDA);
    if (!cu(cudaMemset(c->d_matchCount, 0, P * sizeof(int)), "memset")) return fail(ORBF_ERR_CUDA);
    if (!cu(cudaMemset(c->d_lkpCount, 0, B * ORBF_MAX_LEVELS * sizeof(int)), "memset")) return fail(ORBF_ERR_CUDA);
    if (!cu(cudaMemset(c->d_candCount, 0, B * ORBF_MAX_LEVELS * sizeof(int)), "memset")) return fail(ORBF_ERR_CUDA);
    if (!cu(cudaMallocHost((void**)&c->h_kp, K * sizeof(orbf_keypoint)), "pinned")) return fail(ORBF_ERR_CUDA);
    if (!cu(cudaMallocHost((void**)&c->h_desc, K * 32), "pinned")) return fail(ORBF_ERR_CUDA);
    if (!cu(cudaMallocHost((void**)&c->h_xyz, K * 3 * sizeof(float)), "pinned")) return fail(ORBF_ERR_CUDA);
    if (!cu(cudaMallocHost((void**)&c->h_counts, (std::max(B, P) + 16) * sizeof(int)), "pinned")) return fail(ORBF_ERR_CUDA);
#undef TRY
    *out = c;
    return ORBF_OK;
}

extern "C" int orbf_destroy(orbf_context* c)
{
    if (!c) return ORBF_ERR_ARG;
    cudaSetDevice(c->cfg.device);
    if (c->hi) cudaStreamSynchronize(c->hi);               // a RANSAC left on the side stream (pipeline_overlap) still reads the buffers freed below
    if (c->copy) cudaStreamSynchronize(c->copy);
    for (int i = 0; i < c->nWork; ++i) if (c->work[i]) cudaStreamSynchronize(c->work[i]);
    if (c->stream) cudaStreamSynchronize(c->stream);
    orbf_comm_release(c);
    void* ptrs[] = { c->d_in, c->d_depthIn, c->d_resizeTab, c->d_cells, c->d_strips, c->d_blTiles, c->d_rsTiles, c->d_bgr, c->d_lg, c->d_cellCand, c->d_cellCount, c->d_cand,
        c->d_candCount, c->d_nodeScratch, c->d_lkp, c->d_lkpCount, c->d_kpx, c->d_kpy, c->d_kpsize, c->d_kpangle, c->d_kpresp,
        c->d_ptx, c->d_pty, c->d_ptz, c->d_uright, c->d_kpux, c->d_kpuy, c->d_kpoct, c->d_kplxy, c->d_desc, c->d_count, c->d_kpAos, c->d_pairs,
        c->d_knn, c->d_rev, c->d_matches, c->d_matchCount, c->d_good, c->d_goodCount, c->d_rres, c->d_rstate, c->d_inliers, c->d_depthCov,
        c->d_samples, c->d_hyp, c->d_qdesc, c->d_tdesc, c->d_sxyz, c->d_txyz, c->d_kfDesc, c->d_kfCount, c->d_pts, c->d_cloudSrc, c->d_cloudTgt, c->d_cloudCount, c->d_scratch,
        c->d_userSamples, c->d_kabsch, c->d_kfKnn, c->d_kfSurv, c->d_kfPairs, c->d_kfQCount, c->d_cellRegion, c->d_regionTh, c->d_regionState, c->d_regionLog };
    for (void* p : ptrs) if (p) cudaFree(p);
    for (int l = 0; l < c->L; ++l) { if (c->d_pyr[l]) cudaFree(c->d_pyr[l]); if (c->d_blur[l]) cudaFree(c->d_blur[l]); }
    if (c->h_arena) cudaFreeHost(c->h_arena);
    if (c->evArena) cudaEventDestroy(c->evArena);
    if (c->h_kp) cudaFreeHost(c->h_kp);
    if (c->h_desc) cudaFreeHost(c->h_desc);
    if (c->h_xyz) cudaFreeHost(c->h_xyz);
    if (c->h_counts) cudaFreeHost(c->h_counts);
    for (int i = 0; i < ST_COUNT; ++i) { if (c->evA[i]) cudaEventDestroy(c->evA[i]); if (c->evB[i]) cudaEventDestroy(c->evB[i]); }
    for (int i = 0; i < ORBF_MAX_WORKERS; ++i) {
        if (c->work[i]) { cudaStreamSynchronize(c->work[i]); cudaStreamDestroy(c->work[i]); }
        if (c->evDone[i]) cudaEventDestroy(c->evDone[i]);
        if (c->evExtract[i]) cudaEventDestroy(c->evExtract[i]);
    }
    if (c->hi) { cudaStreamSynchronize(c->hi); cudaStreamDestroy(c->hi); }
    if (c->evHiA) cudaEventDestroy(c->evHiA);
    if (c->evHiB) cudaEventDestroy(c->evHiB);
    if (c->evRansacIn) cudaEventDestroy(c->evRansacIn);
    if (c->evRansac) cudaEventDestroy(c->evRansac);
    for (int i = 0; i < 8; ++i) if (c->evHiGroup[i]) cudaEventDestroy(c->evHiGroup[i]);
    if (c->evFork) cudaEventDestroy(c->evFork);
    if (c->copy) { cudaStreamSynchronize(c->copy); cudaStreamDestroy(c->copy); }
    for (int i = 0; i < ORBF_MAX_CHUNKS; ++i) if (c->evCopy[i]) cudaEventDestroy(c->evCopy[i]);
    for (int i = 0; i < ORBF_MARKERS; ++i) if (c->evMarker[i]) cudaEventDestroy(c->evMarker[i]);
    if (c->ownStream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return ORBF_OK;
}

extern "C" int orbf_set_stream(orbf_context* c, void* s)
{
    if (!c) return ORBF_ERR_ARG;
    if (c->hiPending) { ORBF_CUDA(c, cudaStreamSynchronize(c->hi)); c->hiPending = false; }   // nothing of the old stream's epoch stays in flight
    ORBF_CUDA(c, cudaStreamSynchronize(c->stream));
    if (c->ownStream && c->stream) cudaStreamDestroy(c->stream);
    c->stream = (cudaStream_t)s;
    c->ownStream = false;
    return ORBF_OK;
}

int orbf_join_side(orbf_context* c)
{
    if (c->hiPending) {
        ORBF_CUDA(c, cudaStreamWaitEvent(c->stream, c->evRansac, 0));
        c->hiPending = false;
    }
    return ORBF_OK;
}

extern "C" int orbf_join(orbf_context* c)
{
    if (!c) return ORBF_ERR_ARG;
    return orbf_join_side(c);
}

extern "C" int orbf_synchronize(orbf_context* c)
{
    if (!c) return ORBF_ERR_ARG;
    { const int r = orbf_join_side(c); if (r != ORBF_OK) return r; }
    ORBF_CUDA(c, cudaStreamSynchronize(c->stream));
    return ORBF_OK;
}

extern "C" int orbf_launch_count(const orbf_context* c, int64_t* n)
{
    if (!c || !n) return ORBF_ERR_ARG;
    *n = c->launches;
    return ORBF_OK;
}

extern "C" int orbf_get_tables(const orbf_context* c, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2,
    int32_t* nfeat, int32_t* lw, int32_t* lh)
{
    if (!c) return ORBF_ERR_ARG;
    for (int l = 0; l < c->L; ++l) {
        if (scale) scale[l] = c->scale[l];
        if (inv_scale) inv_scale[l] = c->invScale[l];
        if (sigma2) sigma2[l] = c->sigma2[l];
        if (inv_sigma2) inv_sigma2[l] = c->invSigma2[l];
        if (nfeat) nfeat[l] = c->lg[l].nfeat;
        if (lw) lw[l] = c->lg[l].w;
        if (lh) lh[l] = c->lg[l].h;
    }
    return ORBF_OK;
}

extern "C" int orbf_keypoint_capacity(const orbf_context* c, int32_t* cap)
{
    if (!c || !cap) return ORBF_ERR_ARG;
    *cap = c->K;
    return ORBF_OK;
}

// (Re-)encodes the TMA descriptors: the ones over the context's own pyramid planes once, the ones over the caller's
// input plane whenever that plane changes.
int orbf_refresh_maps(orbf_context* c)
{
    int tileW, blurH, blurBW, blurBH, resizeH;
    orbf_stage_tile_geometry(&tileW, &blurH, &blurBW, &blurBH, &resizeH);
#define ENC(map, base, lv, frames, pitch, stride, bw, bh)                                                               \
    do {                                                                                                               \
        const int r__ = orbf_tma_encode_u8(c, &(map), (base), c->lg[lv].w, c->lg[lv].h, (frames), (pitch), (stride), (bw), (bh)); \
        if (r__ != ORBF_OK) return r__;                                                                                \
    } while (0)
    if (!c->tmStaticReady) {
        for (int l = 1; l < c->L; ++l) {
            ENC(c->tmFast[l], c->d_pyr[l], l, c->B, c->lg[l].pitch, (long long)c->lg[l].plane, c->fastBW[l], c->fastBH[l]);
            ENC(c->tmBlur[l], c->d_pyr[l], l, c->B, c->lg[l].pitch, (long long)c->lg[l].plane, blurBW, blurBH);
            ENC(c->tmPatchRaw[l], c->d_pyr[l], l, c->B, c->lg[l].pitch, (long long)c->lg[l].plane, ORBF_RAW_BW, ORBF_RAW_BH);
            if (l + 1 < c->L)
                ENC(c->tmResize[l + 1], c->d_pyr[l], l, c->B, c->lg[l].pitch, (long long)c->lg[l].plane, c->rsBW[l + 1], c->rsBH[l + 1]);
        }
        for (int l = 0; l < c->L; ++l) {     // the descriptor windows arrive 64B-swizzled (describe.cu: bank spreading of the rBRIEF gathers)
            const int r__ = orbf_tma_encode_u8(c, &c->tmPatchBlur[l], c->d_blur[l], c->lg[l].w, c->lg[l].h, c->B, c->lg[l].pitch, (long long)c->lg[l].plane,
                ORBF_PATCH_BW, ORBF_PATCH_BH, true);
            if (r__ != ORBF_OK) return r__;
        }
        c->tmStaticReady = true;
    }
    if (c->tm0Base != c->cur_gray || c->tm0Pitch != c->cur_grayPitch || c->tm0FrameStride != c->cur_grayFrameStride || c->tm0Frames != c->cur_n) {
        if (!c->cur_gray) return ORBF_ERR_STATE;
        ENC(c->tmFast[0], c->cur_gray, 0, c->cur_n, c->cur_grayPitch, c->cur_grayFrameStride, c->fastBW[0], c->fastBH[0]);
        ENC(c->tmBlur[0], c->cur_gray, 0, c->cur_n, c->cur_grayPitch, c->cur_grayFrameStride, blurBW, blurBH);
        ENC(c->tmPatchRaw[0], c->cur_gray, 0, c->cur_n, c->cur_grayPitch, c->cur_grayFrameStride, ORBF_RAW_BW, ORBF_RAW_BH);
        if (c->L > 1) ENC(c->tmResize[1], c->cur_gray, 0, c->cur_n, c->cur_grayPitch, c->cur_grayFrameStride, c->rsBW[1], c->rsBH[1]);
        c->tm0Base = c->cur_gray; c->tm0Pitch = c->cur_grayPitch; c->tm0FrameStride = c->cur_grayFrameStride; c->tm0Frames = c->cur_n;
    }
#undef ENC
    return ORBF_OK;
}

PyrView orbf_pyr_view(const orbf_context* c, bool blurred)
{
    PyrView v;
    memset(&v, 0, sizeof(v));
    v.nlevels = c->L;
    for (int l = 0; l < c->L; ++l) {
        LevelView& q = v.lv[l];
        q.w = c->lg[l].w; q.h = c->lg[l].h;
        if (blurred) { q.base = c->d_blur[l]; q.pitch = c->lg[l].pitch; q.frameStride = (long long)c->lg[l].plane; }
        else if (l == 0) {
            // level 0 is the caller's input plane (mvImagePyramid[0] is a copy of the image, orbextractor.cpp:855)
            q.base = c->cur_gray - (long long)c->cur_slot0 * c->cur_grayFrameStride;
            q.pitch = c->cur_grayPitch; q.frameStride = c->cur_grayFrameStride;
        } else { q.base = c->d_pyr[l]; q.pitch = c->lg[l].pitch; q.frameStride = (long long)c->lg[l].plane; }
    }
    return v;
}

// ---- per-stage device timing (CUDA events on the context stream; bench.py's roofline numbers) ----------------
void orbf_prof_begin(orbf_context* c, int st) { if (c->profiling && c->evA[st]) cudaEventRecord(c->evA[st], c->stream); }
void orbf_prof_end(orbf_context* c, int st) { if (c->profiling && c->evB[st]) { cudaEventRecord(c->evB[st], c->stream); c->evPending[st] = true; } }

extern "C" int orbf_profile_enable(orbf_context* c, int32_t on)
{
    if (!c) return ORBF_ERR_ARG;
    if (on) for (int i = 0; i < ST_COUNT; ++i) {
        if (!c->evA[i]) ORBF_CUDA(c, cudaEventCreate(&c->evA[i]));
        if (!c->evB[i]) ORBF_CUDA(c, cudaEventCreate(&c->evB[i]));
    }
    c->profiling = on != 0;
    for (int i = 0; i < ST_COUNT; ++i) { c->evPending[i] = false; c->stageMs[i] = 0; c->stageCalls[i] = 0; }
    return ORBF_OK;
}

// Synchronises the stream and folds the last recorded interval of every stage into the totals.
extern "C" int orbf_profile_collect(orbf_context* c)
{
    if (!c) return ORBF_ERR_ARG;
    ORBF_CUDA(c, cudaStreamSynchronize(c->stream));
    for (int i = 0; i < ST_COUNT; ++i)
        if (c->evPending[i]) {
            float ms = 0.f;
            ORBF_CUDA(c, cudaEventElapsedTime(&ms, c->evA[i], c->evB[i]));
            c->stageMs[i] += ms; c->stageCalls[i]++; c->evPending[i] = false;
        }
    return ORBF_OK;
}

extern "C" int orbf_profile_read(const orbf_context* c, double* total_ms, int64_t* calls, int32_t cap)
{
    if (!c || cap < ST_COUNT) return ORBF_ERR_ARG;
    for (int i = 0; i < ST_COUNT; ++i) { if (total_ms) total_ms[i] = c->stageMs[i]; if (calls) calls[i] = c->stageCalls[i]; }
    return ORBF_OK;
}

extern "C" const char* orbf_profile_stage_name(int32_t i)
{
    static const char* n[ST_COUNT] = { "pyr_resize", "fast_cell", "quadtree", "blur7", "describe", "hamming_knn2", "match_select",
        "ransac_prepare", "ransac_hyp", "ransac_select" };
    return (i >= 0 && i < ST_COUNT) ? n[i] : "";
}
