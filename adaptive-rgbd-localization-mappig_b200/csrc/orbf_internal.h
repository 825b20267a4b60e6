// csrc/orbf_internal.h — context layout and kernel launch prototypes shared by the .cu files.
// Device memory layout (all per frame slot s in [0, max_frames)), see DESIGN.md §3:
//   level images   pyr[l]   : [B][h_l][pitch_l] u8, pitch_l = align128(w_l); level 0 = caller's input plane
//   blurred levels blur[l]  : same geometry
//   cell candidates          : [B][cellSlotTotal] u32 (x|y<<11|score<<22, relative to minBorder) + [B][nCells] counts
//   level candidates         : [B][candTotal] u32 in reference order + [B][L] counts
//   level keypoints          : [B][kpStageTotal] u32 (quadtree output, list order) + [B][L] counts
//   frame SoA                : kp_x,kp_y,kp_size,kp_angle,kp_resp [B][K] f32; kp_oct [B][K] i32; kp_lxy [B][K] u32;
//                              desc [B][K][32] u8; pt_x,pt_y,pt_z,u_right,kpu_x,kpu_y [B][K] f32; count [B]
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/orbfront.h"
#include "tma.h"

#define ORBF_EDGE 19
#define ORBF_MINB 16          // EDGE_THRESHOLD - 3 (orbextractor.cpp:672)
#define ORBF_HALF_PATCH 15
#define ORBF_MAX_SAMPLE 8
#define ORBF_PATCH_BW 64      // describe.cu: TMA box of a keypoint window, (2 * 19 + 1) rows x 64 bytes
#define ORBF_PATCH_BH 39
#define ORBF_RAW_BW 48        // describe.cu: TMA box of the orientation disc, (2 * 15 + 1) rows x 48 bytes of the raw level
#define ORBF_RAW_BH 31
#define ORBF_MAX_WORKERS 4    // internal worker streams of the chunked pipeline (c_abi.cu)
#define ORBF_MAX_CHUNKS 64    // chunks of one pipelined call (one copy-done event each)
#define ORBF_MARKERS 8        // caller-visible progress markers (orbf_read_results_async / orbf_wait_marker)

struct LevelView {
    const uint8_t* base;      // address of slot 0's plane
    long long frameStride;    // bytes between consecutive slots
    int pitch, w, h;
};
struct PyrView { LevelView lv[ORBF_MAX_LEVELS]; int nlevels; };

struct ResizeCoef { short ofs, a0, a1, pad; };

struct CellDesc {             // one FAST cell (orbextractor.cpp:688-723): scored interior in level coordinates
    short level, x0, y0, w, h;
    short relx, rely;         // offset added to cell-local coords to get minBorder-relative coords
    short pad;
    int slotOff;              // offset of this cell's slots inside a frame's cell-candidate buffer
    int cap;
};

struct StripDesc {            // FAST work unit: up to ORBF_STRIP_CELLS adjacent cells of one cell row (fast.cu)
    short level, nCells;
    short x0, y0;             // first scored pixel (level coordinates) = interior origin of its first cell
    short w, h;               // scored width over all its cells, scored rows
    int firstCell;            // index of its first cell in the CellDesc table (the others follow)
};
#define ORBF_STRIP_CELLS 4
#define ORBF_FAST_BW 176       // fast.cu: fixed row pitch of the strip tile in shared memory (= TMA box width): 18 + 155 + 3
#define ORBF_STRIP_MAX_W (ORBF_FAST_BW - 21)
struct TileDesc { short level, x0, y0, pad; };    // destination tile of the resize / blur stages (pyramid.cu)

struct LevelGeom {
    int w, h, pitch;
    size_t plane;             // bytes per slot plane
    int nfeat;
    int cellsX, cellsY, wCell, hCell, cell0, nCells;
    int candOff, candCap;     // contiguous candidate list (per frame offsets, entries)
    int kpOff, kpCap;         // quadtree output staging
    int nIni; float hX;
    float scale;
    int scaledPatch;
    int tabX, tabY;           // offsets into d_resizeTab of the x / y coefficient tables (level l from l-1)
};

enum OrbfStage { ST_PYRAMID = 0, ST_FAST, ST_QUADTREE, ST_BLUR, ST_DESCRIBE, ST_KNN2, ST_MATCH_SELECT, ST_RANSAC_PREPARE,
    ST_RANSAC_HYP, ST_RANSAC_SELECT, ST_COUNT };

// inputs of a RANSAC call: 3D points of the frames (slot SoA or standalone arrays) and the matches of each pair
struct RansacSet {
    const float *sx, *sy, *sz, *tx, *ty, *tz;     // SoA bases of slot 0 (or standalone arrays)
    long long slotStride;                          // elements per slot (K) or 0
    const int* pairs;                              // [npairs][2] or NULL
    const orbf_dmatch* matches; const int* matchCount;   // [P][K] input matches
    int nsrc, ndst;
};

struct orbf_context {
    orbf_config cfg;
    int L;
    LevelGeom lg[ORBF_MAX_LEVELS];
    float scale[ORBF_MAX_LEVELS], invScale[ORBF_MAX_LEVELS], sigma2[ORBF_MAX_LEVELS], invSigma2[ORBF_MAX_LEVELS];
    int umax[16];
    int B, P;                 // frame slots, pair slots
    int K;                    // keypoint capacity per frame (multiple of 32)
    int nCellsTotal, cellSlotTotal, candTotal, kpStageTotal, maxCellW, maxCellH;
    cudaStream_t stream;      // stream every launcher enqueues on: the caller-visible stream, or (inside a pipelined
                              // batch call) the worker stream of the chunk being enqueued
    bool ownStream;
    // chunked multi-stream pipeline: a batch call forks the caller-visible stream into nWork worker streams, enqueues
    // chunk k (H2D copies + all stages of its frames and frame pairs) on worker k % nWork, and joins them again
    cudaStream_t work[ORBF_MAX_WORKERS]; int nWork, chunkFrames;
    cudaEvent_t evFork, evDone[ORBF_MAX_WORKERS], evExtract[ORBF_MAX_WORKERS];
    // all H2D copies of a pipelined call go back to back on their own stream, so the link never waits for a kernel; chunk k's
    // kernels wait for evCopy[k]
    cudaStream_t copy; cudaEvent_t evCopy[ORBF_MAX_CHUNKS]; cudaEvent_t evMarker[ORBF_MARKERS];
    // high-priority side stream for the latency-bound stages (quadtree, RANSAC): their few, long-running CTAs are placed as
    // soon as SM resources free up and overlap the throughput-bound kernels (blur, Hamming) still running on the main stream
    cudaStream_t hi; cudaEvent_t evHiA, evHiB, evHiGroup[8];
    // device-input calls under pipeline_overlap: RANSAC of a call runs on `hi` and is NOT joined into the context stream until someone
    // needs it (orbf_join / any other entry point): evRansac marks its end, hiSlot0 / hiN the frame slots it reads
    cudaEvent_t evRansacIn, evRansac; bool hiPending; int hiSlot0, hiN, hiPair0, hiNPairs;
    int64_t launches;
    bool profiling; cudaEvent_t evA[ST_COUNT], evB[ST_COUNT]; bool evPending[ST_COUNT]; double stageMs[ST_COUNT]; int64_t stageCalls[ST_COUNT];
    std::string lastError;

    // inputs
    uint8_t* d_in; int inPitch; size_t inPlane;
    uint8_t* d_bgr; int bgrSlots;                  // staging of interleaved BGR frames (orbf_extract_batch_bgr), allocated on first use
    uint16_t* d_depthIn;
    const uint8_t* cur_gray; long long cur_grayFrameStride; int cur_grayPitch; int cur_slot0, cur_n;
    const uint16_t* cur_depth; long long cur_depthFrameStride; int cur_depthPitch;
    // pyramid
    uint8_t* d_pyr[ORBF_MAX_LEVELS];
    uint8_t* d_blur[ORBF_MAX_LEVELS];
    ResizeCoef* d_resizeTab;
    CellDesc* d_cells; std::vector<CellDesc> h_cells;
    // region-adapted extraction (BASELINE config 4, 8-level variant): region of every FAST cell for the current grid, integer thresholds
    // of the regions for the frame being detected, controller state, per-frame logs
    uint8_t* d_cellRegion; int cellRegionGrid; int* d_regionTh; double* d_regionState; int* d_regionLog; int regionLogCap; int regionVideos;   // controller state of regionVideos videos, 25 entries each
    StripDesc* d_strips; int nStrips; int fastBW[ORBF_MAX_LEVELS], fastBH[ORBF_MAX_LEVELS];   // TMA box per level (fast.cu)
    TileDesc* d_blTiles; int nBlTiles;                                                       // blur tiles of all levels
    TileDesc* d_rsTiles; int rsTile0[ORBF_MAX_LEVELS], rsTileN[ORBF_MAX_LEVELS];             // resize tiles per destination level
    int rsBW[ORBF_MAX_LEVELS], rsBH[ORBF_MAX_LEVELS];                                        // resize source box per destination level
    // TMA descriptors: FAST / blur read level l, resize reads level l-1; those over the caller's input plane (FAST 0, blur 0,
    // resize 1) are re-encoded whenever the input pointer / pitch / frame count changes
    CUtensorMap tmFast[ORBF_MAX_LEVELS], tmBlur[ORBF_MAX_LEVELS], tmResize[ORBF_MAX_LEVELS]; bool tmStaticReady;
    CUtensorMap tmPatchRaw[ORBF_MAX_LEVELS], tmPatchBlur[ORBF_MAX_LEVELS];   // 64 x 39 keypoint windows of the raw / blurred levels (describe.cu)
    const void* tm0Base; long long tm0Pitch, tm0FrameStride; int tm0Frames;                  // what the input-plane maps were encoded for
    LevelGeom* d_lg;
    uint32_t* d_cellCand; int* d_cellCount;
    uint32_t* d_cand; int* d_candCount;
    uint16_t* d_nodeScratch;
    uint32_t* d_lkp; int* d_lkpCount;
    float *d_kpx, *d_kpy, *d_kpsize, *d_kpangle, *d_kpresp, *d_ptx, *d_pty, *d_ptz, *d_uright, *d_kpux, *d_kpuy;
    int* d_kpoct; uint32_t* d_kplxy; uint8_t* d_desc; int* d_count;
    orbf_keypoint* d_kpAos;   // staging for D2H in cv::KeyPoint layout
    // host staging (pinned)
    orbf_keypoint* h_kp; uint8_t* h_desc; float* h_xyz; int* h_counts;
    // page-locked staging arena of the one-frame-at-a-time calls (orbf_extract*, orbf_knn_match, orbf_ransac_iterate): pageable caller
    // buffers are copied through it so that every transfer of a call is asynchronous and the call synchronises once
    uint8_t* h_arena; size_t arenaCap, arenaUsed; cudaEvent_t evArena; bool arenaBusy;
    // one-frame call: depth plane still to be copied into the staging arena, done on the host under the pyramid / FAST / quadtree kernels
    const uint16_t* pendDepthSrc; uint16_t* pendDepthDst; int64_t pendDepthStride;
    uint8_t* d_scratch; size_t scratchCap;       // device scratch of the host-in / host-out helper calls (grows, never freed per call)

    // matching (pair slots)
    int* d_pairs;             // [P][2]
    uint32_t* d_knn;          // [P][K][2] packed (dist<<16 | idx) best, second
    uint32_t* d_rev;          // [P][K] packed best query per train (cross-check)
    orbf_dmatch* d_matches; int* d_matchCount;   // [P][K]
    int lastNPairs; bool pairsFromSlots;
    RansacSet lastRs; orbf_ransac_config lastRansacCfg;            // what orbf_launch_ransac last ran on
    float4* d_cloudSrc; float4* d_cloudTgt; int* d_cloudCount; size_t cloudCap;   // Ransac::mpSourceCloud / mpTargetCloud of the pairs last solved (lazy)
    // standalone matching staging
    uint8_t* d_qdesc; uint8_t* d_tdesc; int descStageRows;

    // RANSAC (pair slots)
    orbf_dmatch* d_good; int* d_goodCount;        // [P][K] filtered + sorted
    int* d_samples;                               // [P][iters][S]
    orbf_hyp_trace* d_hyp; int hypCap;            // [P][iters]
    orbf_ransac_result* d_rres;                   // [P]
    void* d_rstate;                               // [P] sequential accept-rule state (ransac.cu: RState, 28 B)
    orbf_dmatch* d_inliers;                       // [P][K]
    double* d_depthCov;                           // [2]: covariance latched on the context (quirk Q7), per-call value
    float* d_sxyz; float* d_txyz;                 // standalone staging: SoA x|y|z, grown on demand
    int xyzStageRows;
    void* d_pts; size_t ptsCap;                   // packed sorted correspondences [P][K] (ransac.cu: Pt6)
    int* d_userSamples; int userSamplesCap;
    float* d_kabsch;                              // staging for orbf_kabsch
    int kabschCap;

    // keyframe store
    uint8_t* d_kfDesc; int* d_kfCount; int kfCap;
    const uint8_t* d_kfExtDesc; const int* d_kfExtCount; int kfExtN;   // caller-owned store (orbf_kfdb_attach_device)
    uint32_t* d_kfKnn; int* d_kfSurv; int* d_kfPairs; int* d_kfQCount; int kfOutCap;
    // multi-GPU keyframe store (comm.cu): NCCL communicator (loaded with dlopen), all-gathered copy of every rank's store, and the
    // peer stores opened through CUDA IPC for the gather-free path
    void* ncclComm; int commRanks, commRank;
    uint8_t* d_kfGather; int* d_kfGatherCount; int kfGatherCap;
    const uint8_t** d_peerDesc; const int** d_peerCount; void* peerOpened[2 * 64]; int nPeers, peerKf;
};

// ---- error helpers --------------------------------------------------------------------------------
int orbf_cuda_fail(orbf_context* ctx, cudaError_t e, const char* what, const char* file, int line);
#define ORBF_CUDA(ctx, call)                                                              \
    do {                                                                                  \
        cudaError_t e__ = (call);                                                         \
        if (e__ != cudaSuccess) return orbf_cuda_fail((ctx), e__, #call, __FILE__, __LINE__); \
    } while (0)
#define ORBF_LAUNCH_CHECK(ctx)                                                            \
    do {                                                                                  \
        (ctx)->launches++;                                                                \
        cudaError_t e__ = cudaGetLastError();                                             \
        if (e__ != cudaSuccess) return orbf_cuda_fail((ctx), e__, "kernel launch", __FILE__, __LINE__); \
    } while (0)

__host__ __device__ static inline int align_up(int v, int a) { return (v + a - 1) / a * a; }

void orbf_prof_begin(orbf_context* c, int stage);
void orbf_prof_end(orbf_context* c, int stage);

void orbf_comm_release(orbf_context* c);      // comm.cu: communicator, gathered store, peer mappings

// ---- stage launchers (each enqueues on ctx->stream for slots [slot0, slot0+n)) ---------------------
PyrView orbf_pyr_view(const orbf_context* ctx, bool blurred);
int orbf_refresh_maps(orbf_context* ctx);
void orbf_stage_tile_geometry(int* tileW, int* blurH, int* blurBW, int* blurBH, int* resizeH);
int orbf_launch_pyramid(orbf_context* ctx, int slot0, int n);
int orbf_launch_blur(orbf_context* ctx, int slot0, int n);
// adapted: per-region thresholds (d_cellRegion / d_regionTh); perVideo: frame i of the launch reads the thresholds of video i (25 entries each)
int orbf_launch_fast(orbf_context* ctx, int slot0, int n, bool adapted = false, bool perVideo = false);
int orbf_launch_region_control(orbf_context* ctx, int slot, int frameIdx, const orbf_adaptive_config& cfg, int nVideos = 1);
int orbf_region_tables(orbf_context* ctx, int grid, int nFrames, int nVideos = 1);
int orbf_launch_quadtree(orbf_context* ctx, int slot0, int n);
int orbf_launch_describe(orbf_context* ctx, int slot0, int n);
int orbf_launch_pack_aos(orbf_context* ctx, int slot0, int n);
int orbf_launch_bgr2gray(orbf_context* ctx, const uint8_t* d_bgr, int bgrPitch, long long bgrFrameStride, int slot0, int n);
// matching: query/train descriptor matrices addressed per pair
struct MatchSet {
    const uint8_t* qdesc; const uint8_t* tdesc;   // base of slot 0 (row stride 32 B)
    long long qStride, tStride;                    // bytes per slot (K*32), 0 when every pair uses slot 0
    const int* qCounts; const int* tCounts;        // per-slot row counts (NULL => nq / nt)
    // sharded train side (keyframe stores of several GPUs read in place over NVLink): train slot ts lives on shard ts / shardKf at
    // local index ts % shardKf; tShards / tShardCounts are device arrays of (peer) pointers.  NULL => tdesc / tCounts above
    const uint8_t* const* tShards = nullptr; const int* const* tShardCounts = nullptr; int shardKf = 0;
    const int* pairs;                              // [npairs][2] (query slot, train slot); NULL => slots (0, 0)
    int pair0;                                     // first pair slot of this launch (pair = pair0 + blockIdx)
    int nq, nt;
    uint32_t* knn;                                 // out [npairs][K][2] packed (dist<<16 | trainIdx): best, second
    uint32_t* rev;                                 // out [npairs][K]    packed (dist<<16 | queryIdx): best query per train row
    orbf_dmatch* matches; int* matchCount;         // out [npairs][K], [npairs]   (matches may be NULL: count only)
};
int orbf_launch_knn2(orbf_context* ctx, const MatchSet& ms, int npairs, bool cross);
int orbf_launch_distinctive(orbf_context* ctx, const uint8_t* d_desc, const int* d_offsets, int nLandmarks, int* d_best, int* d_median);
int orbf_launch_undistort(orbf_context* ctx, const float* d_xy, int n, float fx, float fy, float cx, float cy, const float* dist5, float* d_out);
int orbf_launch_unproject(orbf_context* ctx, const float* d_xy, const uint16_t* d_raw, int n, float* d_xyz, float* d_uright, float* d_xyUn);
int orbf_launch_compose(orbf_context* ctx, int npairs, const float* d_pose0, float* d_poses, uint8_t* d_outlier);
int orbf_launch_fuse_search(orbf_context* ctx, const float* Rcw, const float* tcw, const float* camera, const float* d_kpx, const float* d_kpy, const float* d_uright,
    const uint8_t* d_desc, int nFeat, const float* d_lmPos, const uint8_t* d_lmDesc, const uint8_t* d_lmValid, int nLm, float radius, int thLow, int* d_bestIdx,
    int* d_bestDist);
int orbf_launch_bow_match(orbf_context* ctx, const int* d_words1, const int* d_off1, const int* d_idx1, int nw1, const uint8_t* d_desc1, const int* d_words2,
    const int* d_off2, const int* d_idx2, int nw2, const uint8_t* d_desc2, int nTrain, float nnRatio, int thLow, int nEntries, int* d_entryTrain, int* d_entryDist,
    int* d_firstUser, orbf_dmatch* d_out, int* d_nOut);
int orbf_launch_projection_match(orbf_context* ctx, const float* d_kpx, const float* d_kpy, const int* d_kpoct, const uint8_t* d_desc, int nFeat,
    const uint8_t* d_lmDesc, const float* d_projX, const float* d_projY, const uint8_t* d_lmFlags, int nLm, const uint8_t* d_featTaken, float radius, float nnRatio,
    int thHigh, uint32_t* d_cand, int* d_candOct /* [nLm][8] */, int* d_candCount, int* d_bestIdx, int* d_nMatches);
int orbf_launch_match_select(orbf_context* ctx, const MatchSet& ms, int npairs, float ratio, bool cross);
// RANSAC
int orbf_launch_kabsch(orbf_context* ctx, const float* dA, const float* dB, int n, float* dT);
// makes the context stream wait for a RANSAC still running on the side stream (no host wait); no-op when nothing is pending
int orbf_join_side(orbf_context* ctx);
// pairs [pair0, pair0 + npairs).  The depth covariance (quirk Q7) is latched by the first pair, in enqueue order, that reaches
// scoring; standalone = the call neither sees nor replaces the value latched on the context (orbf_ransac_iterate).
int orbf_ransac_reserve(orbf_context* ctx, const orbf_ransac_config& cfg);
// mpSourceCloud / mpTargetCloud (Odometry/ransac.cpp:171-189) of pairs [pair0, pair0 + npairs) of the set last handed to orbf_launch_ransac
int orbf_launch_ransac_clouds(orbf_context* ctx, int pair0, int npairs);
int orbf_launch_ransac(orbf_context* ctx, const RansacSet& rs, int pair0, int npairs, const orbf_ransac_config& cfg,
    const int* d_userSamples, bool standalone, bool fullTable = false, bool probeOnly = false, int firstWave = 0, int lastWave = 5,
    const float* d_composeIn = nullptr, float* d_composeOut = nullptr);
// the done flag of a pair's sequential RANSAC loop (device pointer to one int), valid after orbf_launch_ransac's replay kernels
const int* orbf_ransac_done_flag(orbf_context* ctx, int pair);

// Carve-up of the context's persistent device scratch (host-in / host-out helper calls: the section-8f matcher entry points, the adaptive
// detector): 256-byte aligned pieces of ONE allocation that grows and is never freed per call.  Every user synchronises the context
// stream before it returns, so calls never overlap on it.
struct Scratch {
    orbf_context* ctx; uint8_t* base = nullptr; size_t size = 0;
    explicit Scratch(orbf_context* c) : ctx(c) {}
    size_t take(size_t bytes) { const size_t o = size; size = (size + (bytes ? bytes : 1) + 255) & ~(size_t)255; return o; }
    cudaError_t alloc()
    {
        if (size > ctx->scratchCap) {
            cudaError_t e = cudaStreamSynchronize(ctx->stream);
            if (e != cudaSuccess) return e;
            if (ctx->d_scratch) cudaFree(ctx->d_scratch);
            ctx->d_scratch = nullptr; ctx->scratchCap = 0;
            const size_t cap = (size + (1u << 20)) & ~(size_t)((1u << 20) - 1);
            e = cudaMalloc((void**)&ctx->d_scratch, cap);
            if (e != cudaSuccess) return e;
            ctx->scratchCap = cap;
        }
        base = ctx->d_scratch;
        return cudaSuccess;
    }
    template <typename T> T* at(size_t off) const { return reinterpret_cast<T*>(base + off); }
};
