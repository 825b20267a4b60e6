// csrc/c_abi.cu — extern "C" entry points of include/orbfront.h (the drop-in boundary).
// Host <-> device staging lives here; every stage itself is a CUDA kernel (no CPU fallback anywhere).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cstdio>
#include <cstdlib>
#include "orbf_internal.h"
#include "replay.h"
#include "glibc_sincosf.h"

#define CTX_ENTER_NOJOIN(c)                                                            \
    do {                                                                               \
        if (!(c)) return ORBF_ERR_ARG;                                                 \
        cudaError_t e_ = cudaSetDevice((c)->cfg.device);                               \
        if (e_ != cudaSuccess) return orbf_cuda_fail((c), e_, "cudaSetDevice", __FILE__, __LINE__); \
    } while (0)
// every entry point but orbf_track_sequence_device_at first orders the context stream behind a RANSAC still running on the side stream
#define CTX_ENTER(c)                                                                   \
    do {                                                                               \
        CTX_ENTER_NOJOIN(c);                                                           \
        const int j_ = orbf_join_side(c);                                              \
        if (j_ != ORBF_OK) return j_;                                                  \
    } while (0)
#define TRY(x) do { int r__ = (x); if (r__ != ORBF_OK) return r__; } while (0)

// ---- page-locked staging arena (one-frame-at-a-time calls) ---------------------------------------------------------------
// arena_begin: start of a call that stages through the arena; waits for the transfers of the previous such call that may still read it.
static int arena_begin(orbf_context* c, size_t need)
{
    if (c->arenaBusy) { ORBF_CUDA(c, cudaEventSynchronize(c->evArena)); c->arenaBusy = false; }
    if (need > c->arenaCap) {
        if (c->h_arena) cudaFreeHost(c->h_arena);
        c->h_arena = nullptr; c->arenaCap = 0;
        const size_t cap = (need + (1u << 20)) & ~(size_t)((1u << 20) - 1);
        ORBF_CUDA(c, cudaMallocHost((void**)&c->h_arena, cap));
        c->arenaCap = cap;
    }
    if (!c->evArena) ORBF_CUDA(c, cudaEventCreateWithFlags(&c->evArena, cudaEventDisableTiming));
    c->arenaUsed = 0;
    return ORBF_OK;
}
static uint8_t* arena_take(orbf_context* c, size_t bytes)
{
    uint8_t* p = c->h_arena + c->arenaUsed;
    const size_t used = (c->arenaUsed + bytes + 255) & ~(size_t)255;
    if (used > c->arenaCap) {       // the callers size the arena from upper bounds of what they take: never reached; fail loudly, not silently
        c->lastError = "staging arena overflow (internal sizing error)";
        fprintf(stderr, "orbfront: %s\n", c->lastError.c_str());
        abort();
    }
    c->arenaUsed = used;
    return p;
}
// marks the arena as read by work queued on `stream` (a later arena_begin waits for it); calls that end with a stream
// synchronisation do not need it
static int arena_fence(orbf_context* c, cudaStream_t stream)
{
    ORBF_CUDA(c, cudaEventRecord(c->evArena, stream));
    c->arenaBusy = true;
    return ORBF_OK;
}
static bool is_pageable(const void* p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return at.type == cudaMemoryTypeUnregistered;
}

struct DistanceLess { ORBF_HD bool operator()(const orbf_dmatch& a, const orbf_dmatch& b) const { return a.distance < b.distance; } };

struct StreamSwap {   // runs the launches of a scope on another stream
    orbf_context* c; cudaStream_t saved;
    StreamSwap(orbf_context* ctx, cudaStream_t s) : c(ctx), saved(ctx->stream) { ctx->stream = s; }
    ~StreamSwap() { c->stream = saved; }
};

// sideStream: the quadtree (latency-bound, ~20 % issue utilisation) runs on the high-priority side stream while the blur
// (independent of it) keeps the SMs busy on the main stream; describe waits for both (4.84 -> 4.79 ms per 512 frames).
// The same trick for RANSAC behind the matcher was measured and dropped: RANSAC is a chain of ten latency-bound launches
// whose duration does not shrink with the number of pairs, so splitting the pairs into groups only repeats it (5.2-7.9 ms).
static int run_extract(orbf_context* c, int slot0, int n, bool sideStream)
{
    sideStream = sideStream && c->hi && !c->profiling;
    orbf_prof_begin(c, ST_PYRAMID); TRY(orbf_launch_pyramid(c, slot0, n)); orbf_prof_end(c, ST_PYRAMID);
    orbf_prof_begin(c, ST_FAST); TRY(orbf_launch_fast(c, slot0, n)); orbf_prof_end(c, ST_FAST);
    if (sideStream) {
        ORBF_CUDA(c, cudaEventRecord(c->evHiA, c->stream));
        ORBF_CUDA(c, cudaStreamWaitEvent(c->hi, c->evHiA, 0));
        { StreamSwap sw(c, c->hi); TRY(orbf_launch_quadtree(c, slot0, n)); }
        ORBF_CUDA(c, cudaEventRecord(c->evHiB, c->hi));
        TRY(orbf_launch_blur(c, slot0, n));
        ORBF_CUDA(c, cudaStreamWaitEvent(c->stream, c->evHiB, 0));
    } else {
        orbf_prof_begin(c, ST_QUADTREE); TRY(orbf_launch_quadtree(c, slot0, n)); orbf_prof_end(c, ST_QUADTREE);
        orbf_prof_begin(c, ST_BLUR); TRY(orbf_launch_blur(c, slot0, n)); orbf_prof_end(c, ST_BLUR);
    }
    if (c->pendDepthSrc) {
        // one-frame call: the depth plane goes into the page-locked arena now, on the host, while the kernels queued above run; the
        // describe kernel (launched after this copy has finished) samples it in place
        const int w = c->cfg.width, h = c->cfg.height;
        if (c->pendDepthStride == w) memcpy(c->pendDepthDst, c->pendDepthSrc, (size_t)w * h * 2);
        else for (int y = 0; y < h; ++y) memcpy(c->pendDepthDst + (size_t)y * w, c->pendDepthSrc + (size_t)y * c->pendDepthStride, (size_t)w * 2);
        c->pendDepthSrc = nullptr;
    }
    orbf_prof_begin(c, ST_DESCRIBE); TRY(orbf_launch_describe(c, slot0, n)); orbf_prof_end(c, ST_DESCRIBE);
    return ORBF_OK;
}

// ------------------------------------------------------------------------------------------------------
// chunked multi-stream pipeline
// ------------------------------------------------------------------------------------------------------
// A batched host-input call cuts its frames into chunks.  ALL H2D copies go back to back on the context's copy stream (the PCIe link,
// which bounds this path, never waits for a kernel), one event per chunk; chunk k's kernels — every extraction stage of its frames and,
// for the sequence calls, matching of its consecutive frame pairs — run on worker stream k % nWork behind that event, so the
// latency-bound stages of one chunk (quadtree, descriptor gathers) overlap the throughput-bound stages (FAST, blur, Hamming) of
// another.  The caller-visible stream is forked into copy + worker streams at the start and joins them before the call returns
// (everything stays asynchronous with respect to the host); with orbf_config.pipeline_overlap the fork is dropped, so the copies
// of the next call start while this call's tail (last chunks, RANSAC, result read-back) is still running.
// With per-stage profiling on, or pipeline_chunk < 0, everything runs as one chunk on the caller-visible stream.
struct Pipeline {
    orbf_context* c;
    cudaStream_t main;
    bool used[ORBF_MAX_WORKERS];
    bool active, forked;
    explicit Pipeline(orbf_context* ctx) : c(ctx), main(ctx->stream), active(false), forked(false) { for (bool& u : used) u = false; }
    ~Pipeline() { c->stream = main; }
    int begin(bool wanted)
    {
        active = wanted && c->nWork > 0 && c->chunkFrames > 0 && !c->profiling && c->copy;
        forked = active && !c->cfg.pipeline_overlap;
        if (forked) {
            ORBF_CUDA(c, cudaEventRecord(c->evFork, main));
            ORBF_CUDA(c, cudaStreamWaitEvent(c->copy, c->evFork, 0));
        }
        return ORBF_OK;
    }
    int enter(int k)   // make worker k % nWork the current stream
    {
        if (!active) return ORBF_OK;
        const int s = k % c->nWork;
        if (!used[s]) { if (forked) ORBF_CUDA(c, cudaStreamWaitEvent(c->work[s], c->evFork, 0)); used[s] = true; }
        c->stream = c->work[s];
        return ORBF_OK;
    }
    int end()
    {
        c->stream = main;
        if (!active) return ORBF_OK;
        for (int s = 0; s < c->nWork; ++s)
            if (used[s]) {
                ORBF_CUDA(c, cudaEventRecord(c->evDone[s], c->work[s]));
                ORBF_CUDA(c, cudaStreamWaitEvent(main, c->evDone[s], 0));
            }
        return ORBF_OK;
    }
};

struct HostFrames {    // host-side input of a batched call (NULL gray => inputs are already on the device)
    const uint8_t* gray; int64_t grayStride, grayFrameStride;
    const uint16_t* depth; int64_t depthStride, depthFrameStride;
    bool depthInPlace;     // depth stays in pinned host memory and is sampled over PCIe (orbf_config.depth_zero_copy)
};

static int upload_chunk(orbf_context* c, const HostFrames& hf, int slotA, int first, int n)
{
    const int w = c->cfg.width, h = c->cfg.height;
    uint8_t* dIn = c->d_in + (size_t)slotA * c->inPlane;
    const uint8_t* g = hf.gray + (size_t)first * hf.grayFrameStride;
    if (hf.grayFrameStride == hf.grayStride * h) {
        ORBF_CUDA(c, cudaMemcpy2DAsync(dIn, c->inPitch, g, hf.grayStride, w, (size_t)h * n, cudaMemcpyHostToDevice, c->stream));
    } else {
        for (int i = 0; i < n; ++i)
            ORBF_CUDA(c, cudaMemcpy2DAsync(dIn + (size_t)i * c->inPlane, c->inPitch, g + (size_t)i * hf.grayFrameStride, hf.grayStride, w, h,
                cudaMemcpyHostToDevice, c->stream));
    }
    if (hf.depth && !hf.depthInPlace) {
        uint16_t* dDepth = c->d_depthIn + (size_t)slotA * w * h;
        const uint16_t* d = hf.depth + (size_t)first * hf.depthFrameStride;
        if (hf.depthStride == w && hf.depthFrameStride == (int64_t)w * h) {
            ORBF_CUDA(c, cudaMemcpyAsync(dDepth, d, (size_t)n * w * h * sizeof(uint16_t), cudaMemcpyHostToDevice, c->stream));
        } else {
            for (int i = 0; i < n; ++i)
                ORBF_CUDA(c, cudaMemcpy2DAsync(dDepth + (size_t)i * w * h, (size_t)w * 2, d + (size_t)i * hf.depthFrameStride,
                    (size_t)hf.depthStride * 2, (size_t)w * 2, h, cudaMemcpyHostToDevice, c->stream));
        }
    }
    return ORBF_OK;
}

static MatchSet slot_match_set(orbf_context* c)
{
    MatchSet ms;
    ms.qdesc = c->d_desc; ms.tdesc = c->d_desc; ms.qStride = ms.tStride = (long long)c->K * 32;
    ms.qCounts = ms.tCounts = c->d_count; ms.pairs = c->d_pairs; ms.pair0 = 0; ms.nq = 0; ms.nt = 0;
    ms.knn = c->d_knn; ms.rev = c->d_rev; ms.matches = c->d_matches; ms.matchCount = c->d_matchCount;
    return ms;
}

static RansacSet slot_ransac_set(orbf_context* c)
{
    RansacSet rs;
    rs.sx = c->d_ptx; rs.sy = c->d_pty; rs.sz = c->d_ptz; rs.tx = c->d_ptx; rs.ty = c->d_pty; rs.tz = c->d_ptz;
    rs.slotStride = c->K; rs.pairs = c->d_pairs; rs.matches = c->d_matches; rs.matchCount = c->d_matchCount;
    rs.nsrc = rs.ndst = c->K;
    return rs;
}

static int run_match_group(orbf_context* c, int pair0, int npairs, float ratio, bool cross)
{
    MatchSet ms = slot_match_set(c);
    ms.pair0 = pair0;
    orbf_prof_begin(c, ST_KNN2); TRY(orbf_launch_knn2(c, ms, npairs, cross)); orbf_prof_end(c, ST_KNN2);
    orbf_prof_begin(c, ST_MATCH_SELECT); TRY(orbf_launch_match_select(c, ms, npairs, ratio, cross)); orbf_prof_end(c, ST_MATCH_SELECT);
    return ORBF_OK;
}

// pair slot table entries [pair0, pair0 + np): pair slot p = (frame slot f0 + (p - pair0), the next one)
__global__ void consecutive_pairs_kernel(int* pairs, int pair0, int np, int f0)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < np) { pairs[2 * (pair0 + i)] = f0 + i; pairs[2 * (pair0 + i) + 1] = f0 + i + 1; }
}

// Extraction of frames [slot0, slot0+n) and, when `track` is set, matching + RANSAC of the n-1 consecutive pairs
// (pair slot pairSlot0 + p = (slot0+p, slot0+p+1)), chunk by chunk.  The pair that straddles two chunks runs with the later chunk,
// after an event says the earlier chunk's extraction is complete.
struct TrackArgs { float ratio; bool cross; const orbf_ransac_config* rcfg; int pairSlot0; };

static int run_batch(orbf_context* c, int slot0, int n, const HostFrames* hf, const TrackArgs* track)
{
    const int npairs = track ? n - 1 : 0;
    const int ps0 = track ? track->pairSlot0 : 0;
    if (track && npairs > 0) {
        if (ps0 < 0 || ps0 + npairs > c->P) return ORBF_ERR_ARG;
        if (track->rcfg) TRY(orbf_ransac_reserve(c, *track->rcfg));
        c->lastNPairs = ps0 + npairs; c->pairsFromSlots = true;
    }
    // Measured on B200 (profiles/r1d_pipeline_sweep.txt): the stages are instruction-issue bound, so running chunks on
    // concurrent streams buys nothing when the inputs are already in HBM (7.36 ms/512 frames on one stream vs 7.4-9.5 ms
    // chunked; re-measured after the kernel rewrites: 4.09 ms vs 4.06 ms) — the pipeline exists to hide the PCIe copies of
    // host inputs.
    Pipeline pl(c);
    TRY(pl.begin(hf && hf->gray && n > 1));          // a single frame has nothing to overlap: copy and kernels on the caller's stream
    int chunk = pl.active ? c->chunkFrames : n;
    // chunk boundaries: quarter- and half-size chunks at both ends, so that the first kernels start after a short copy and the
    // last copy is followed by a short tail of compute (the link, not the SMs, bounds this path)
    std::vector<int> bounds;
    for (;; chunk *= 2) {
        bounds.assign(1, 0);
        if (pl.active && chunk >= 64 && n >= 4 * chunk) {
            const int q = chunk / 4, h = chunk / 2;
            bounds.push_back(q); bounds.push_back(q + h);
            const int tailStart = n - q - h;
            for (int a = q + h; a + chunk <= tailStart; a += chunk) bounds.push_back(a + chunk);
            if (bounds.back() < tailStart) bounds.push_back(tailStart);
            bounds.push_back(n - q); bounds.push_back(n);
        } else {
            for (int a = chunk; a < n; a += chunk) bounds.push_back(a);
            bounds.push_back(n);
        }
        if (!pl.active || (int)bounds.size() - 1 <= ORBF_MAX_CHUNKS) break;      // one copy-done event per chunk
    }
    const int nChunks = (int)bounds.size() - 1;
    if (pl.active) {      // every H2D copy of the call, back to back on the copy stream
        StreamSwap sw(c, c->copy);
        for (int k = 0; k < nChunks; ++k) {
            const int a = bounds[k], b = bounds[k + 1];
            if (b <= a) continue;
            TRY(upload_chunk(c, *hf, slot0 + a, a, b - a));
            ORBF_CUDA(c, cudaEventRecord(c->evCopy[k], c->copy));
        }
    }
    int prevWorker = -1;
    for (int k = 0; k < nChunks; ++k) {
        const int a = bounds[k], b = bounds[k + 1];
        if (b <= a) continue;
        TRY(pl.enter(k));
        if (pl.active) ORBF_CUDA(c, cudaStreamWaitEvent(c->stream, c->evCopy[k], 0));
        else if (hf && hf->gray) TRY(upload_chunk(c, *hf, slot0 + a, a, b - a));
        TRY(run_extract(c, slot0 + a, b - a, !pl.active));
        const int wk = pl.active ? k % c->nWork : -1;
        if (pl.active && track) ORBF_CUDA(c, cudaEventRecord(c->evExtract[wk], c->stream));
        if (track && npairs > 0) {
            const int pa = std::max(a - 1, 0), pb = b - 1;          // pairs [pa, pb): the straddling pair a-1 belongs to this chunk
            if (pb > pa) {
                if (pl.active && a > 0 && prevWorker != wk) ORBF_CUDA(c, cudaStreamWaitEvent(c->stream, c->evExtract[prevWorker], 0));
                consecutive_pairs_kernel<<<(pb - pa + 127) / 128, 128, 0, c->stream>>>(c->d_pairs, ps0 + pa, pb - pa, slot0 + pa);
                ORBF_LAUNCH_CHECK(c);
                TRY(run_match_group(c, ps0 + pa, pb - pa, track->ratio, track->cross));
                // RANSAC is a chain of latency-bound launches whose duration barely depends on the number of pairs: per
                // chunk it would be paid once per chunk, so the pipelined path runs it once, after the join, for all pairs
                if (track->rcfg && !pl.active) {
                    if (c->cfg.pipeline_overlap && c->hi && !c->profiling && !(hf && hf->gray)) {
                        // device inputs under pipeline_overlap: RANSAC (latency-bound, a few CTAs) goes to the high-priority side stream and
                        // is not joined back here — the next call's pyramid / FAST on the other slot half run over it.  (The matcher on the
                        // side stream as well was measured: 2.526 -> 2.510 ms per step, 0.6 % — not kept.)
                        ORBF_CUDA(c, cudaEventRecord(c->evRansacIn, c->stream));
                        ORBF_CUDA(c, cudaStreamWaitEvent(c->hi, c->evRansacIn, 0));
                        { StreamSwap sw(c, c->hi); TRY(orbf_launch_ransac(c, slot_ransac_set(c), ps0 + pa, pb - pa, *track->rcfg, nullptr, false)); }
                        ORBF_CUDA(c, cudaEventRecord(c->evRansac, c->hi));
                        c->hiPending = true; c->hiSlot0 = slot0; c->hiN = n; c->hiPair0 = ps0; c->hiNPairs = npairs;
                    } else TRY(orbf_launch_ransac(c, slot_ransac_set(c), ps0 + pa, pb - pa, *track->rcfg, nullptr, false));
                }
            }
        }
        prevWorker = wk;
    }
    const bool deferred = pl.active;
    TRY(pl.end());
    if (deferred && track && npairs > 0 && track->rcfg) {
        // RANSAC for all pairs behind the join, on the high-priority stream: its chain of small dependent launches is placed ahead of
        // the queued CTAs of the next call's extraction kernels (pipeline_overlap), instead of waiting behind each of them
        if (c->hi) {
            ORBF_CUDA(c, cudaEventRecord(c->evHiA, c->stream));
            ORBF_CUDA(c, cudaStreamWaitEvent(c->hi, c->evHiA, 0));
            { StreamSwap sw(c, c->hi); TRY(orbf_launch_ransac(c, slot_ransac_set(c), ps0, npairs, *track->rcfg, nullptr, false)); }
            ORBF_CUDA(c, cudaEventRecord(c->evHiB, c->hi));
            ORBF_CUDA(c, cudaStreamWaitEvent(c->stream, c->evHiB, 0));
        } else TRY(orbf_launch_ransac(c, slot_ransac_set(c), ps0, npairs, *track->rcfg, nullptr, false));
    }
    return ORBF_OK;
}

static int set_device_inputs(orbf_context* c, int slot0, int n, const uint8_t* d_gray, int64_t gray_pitch, int64_t gray_frame_stride,
    const uint16_t* d_depth, int64_t depth_pitch_elems, int64_t depth_frame_stride_elems)
{
    if (!d_gray || n < 1 || slot0 < 0 || slot0 + n > c->B || gray_pitch < c->cfg.width) return ORBF_ERR_ARG;
    if (((uintptr_t)d_gray & 15) || (gray_pitch & 15) || (gray_frame_stride & 15)) return ORBF_ERR_ALIGNMENT;
    if (d_depth && depth_pitch_elems < c->cfg.width) return ORBF_ERR_ARG;
    c->cur_gray = d_gray; c->cur_grayPitch = (int)gray_pitch; c->cur_grayFrameStride = gray_frame_stride;
    c->cur_depth = d_depth; c->cur_depthPitch = (int)depth_pitch_elems; c->cur_depthFrameStride = depth_frame_stride_elems;
    c->cur_slot0 = slot0; c->cur_n = n;
    return ORBF_OK;
}

static int set_host_inputs(orbf_context* c, int slot0, int n, const uint8_t* gray, int64_t gray_stride, int64_t gray_frame_stride,
    const uint16_t* depth, int64_t depth_stride_elems, int64_t depth_frame_stride_elems, HostFrames& hf)
{
    if (!gray || n < 1 || slot0 < 0 || slot0 + n > c->B || gray_stride < c->cfg.width) return ORBF_ERR_ARG;
    if (depth && depth_stride_elems < c->cfg.width) return ORBF_ERR_ARG;
    const int w = c->cfg.width, h = c->cfg.height;
    hf.gray = gray; hf.grayStride = gray_stride; hf.grayFrameStride = gray_frame_stride;
    hf.depth = depth; hf.depthStride = depth_stride_elems; hf.depthFrameStride = depth_frame_stride_elems;
    c->cur_gray = c->d_in + (size_t)slot0 * c->inPlane; c->cur_grayPitch = c->inPitch; c->cur_grayFrameStride = (long long)c->inPlane;
    c->cur_depth = depth ? c->d_depthIn + (size_t)slot0 * w * h : nullptr; c->cur_depthPitch = w; c->cur_depthFrameStride = (long long)w * h;
    hf.depthInPlace = false;
    if (depth && c->cfg.depth_zero_copy >= 0) {
        // A frame needs ~1000 of its 307,200 depth samples (Core/frame.cpp:155: one lookup per keypoint).  When the caller's
        // plane is page-locked, the unprojection reads those samples in place through the unified address space instead of
        // staging 614 KB per frame in HBM: the PCIe link carries 1/3 of the bytes.
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, depth) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer) {
            hf.depthInPlace = true;
            c->cur_depth = static_cast<const uint16_t*>(at.devicePointer);
            c->cur_depthPitch = (int)depth_stride_elems; c->cur_depthFrameStride = depth_frame_stride_elems;
        } else cudaGetLastError();      // pageable memory: not an error, stage it
    }
    c->cur_slot0 = slot0; c->cur_n = n;
    return ORBF_OK;
}

extern "C" int orbf_extract_batch_device(orbf_context* c, int32_t slot0, int32_t n, const uint8_t* d_gray, int64_t gray_pitch,
    int64_t gray_frame_stride, const uint16_t* d_depth, int64_t depth_pitch_elems, int64_t depth_frame_stride_elems)
{
    CTX_ENTER(c);
    TRY(set_device_inputs(c, slot0, n, d_gray, gray_pitch, gray_frame_stride, d_depth, depth_pitch_elems, depth_frame_stride_elems));
    return run_batch(c, slot0, n, nullptr, nullptr);
}

extern "C" int orbf_extract_batch(orbf_context* c, int32_t slot0, int32_t n, const uint8_t* gray, int64_t gray_stride,
    int64_t gray_frame_stride, const uint16_t* depth, int64_t depth_stride_elems, int64_t depth_frame_stride_elems)
{
    CTX_ENTER(c);
    HostFrames hf;
    if (n == 1 && gray && gray_stride >= c->cfg.width && is_pageable(gray)) {
        // the per-frame drop-in call (ORBextractor::operator() / Frame::ExtractFeatures): a pageable plane goes through the page-locked
        // arena (one host memcpy, then a truly asynchronous DMA) instead of the driver's own staged, host-blocking copy
        const int w = c->cfg.width, h = c->cfg.height;
        const bool withDepth = depth && depth_stride_elems >= w;
        TRY(arena_begin(c, (size_t)w * h * (withDepth ? 3 : 1) + 1024));
        uint8_t* g = arena_take(c, (size_t)w * h);
        if (gray_stride == w) memcpy(g, gray, (size_t)w * h);                    // tight rows: one copy
        else for (int y = 0; y < h; ++y) memcpy(g + (size_t)y * w, gray + (size_t)y * gray_stride, (size_t)w);
        uint16_t* d = withDepth ? reinterpret_cast<uint16_t*>(arena_take(c, (size_t)w * h * 2)) : nullptr;
        TRY(set_host_inputs(c, slot0, 1, g, w, (int64_t)w * h, d, w, (int64_t)w * h, hf));
        if (withDepth) {
            if (hf.depthInPlace) { c->pendDepthSrc = depth; c->pendDepthDst = d; c->pendDepthStride = depth_stride_elems; }   // copied under the kernels (run_extract)
            else for (int y = 0; y < h; ++y) memcpy(d + (size_t)y * w, depth + (size_t)y * depth_stride_elems, (size_t)w * 2);
        }
        const int r = run_batch(c, slot0, 1, &hf, nullptr);
        c->pendDepthSrc = nullptr;
        if (r != ORBF_OK) return r;
        return arena_fence(c, c->stream);
    }
    TRY(set_host_inputs(c, slot0, n, gray, gray_stride, gray_frame_stride, depth, depth_stride_elems, depth_frame_stride_elems, hf));
    return run_batch(c, slot0, n, &hf, nullptr);
}

// Frame::Frame + Frame::ExtractFeatures from the colour image (Core/frame.cpp:18-45, 135-170): interleaved 8-bit BGR host
// frames are copied to HBM, converted to gray on the device (cv::cvtColor CV_BGR2GRAY, fixed point) into the frame slots,
// and extracted.  orbf_download_gray returns mImGray of a slot.
extern "C" int orbf_extract_batch_bgr(orbf_context* c, int32_t slot0, int32_t n, const uint8_t* bgr, int64_t bgr_stride,
    int64_t bgr_frame_stride, const uint16_t* depth, int64_t depth_stride_elems, int64_t depth_frame_stride_elems)
{
    CTX_ENTER(c);
    const int w = c->cfg.width, h = c->cfg.height;
    if (!bgr || n < 1 || slot0 < 0 || slot0 + n > c->B || bgr_stride < 3 * (int64_t)w) return ORBF_ERR_ARG;
    if (c->bgrSlots < n) {
        ORBF_CUDA(c, cudaStreamSynchronize(c->stream));
        if (c->d_bgr) cudaFree(c->d_bgr);
        c->d_bgr = nullptr; c->bgrSlots = 0;
        ORBF_CUDA(c, cudaMalloc((void**)&c->d_bgr, (size_t)n * h * w * 3));
        c->bgrSlots = n;
    }
    const size_t rowBytes = (size_t)w * 3;
    if (bgr_frame_stride == bgr_stride * h) {
        ORBF_CUDA(c, cudaMemcpy2DAsync(c->d_bgr, rowBytes, bgr, bgr_stride, rowBytes, (size_t)h * n, cudaMemcpyHostToDevice, c->stream));
    } else {
        for (int i = 0; i < n; ++i)
            ORBF_CUDA(c, cudaMemcpy2DAsync(c->d_bgr + (size_t)i * h * rowBytes, rowBytes, bgr + (size_t)i * bgr_frame_stride, bgr_stride, rowBytes, h,
                cudaMemcpyHostToDevice, c->stream));
    }
    TRY(orbf_launch_bgr2gray(c, c->d_bgr, (int)rowBytes, (long long)h * rowBytes, slot0, n));
    HostFrames hf;
    // the gray planes are already in the frame slots: reuse the host-input bookkeeping for depth, skip the gray upload
    uint8_t dummy = 0;
    TRY(set_host_inputs(c, slot0, n, &dummy, w, (int64_t)w * h, depth, depth_stride_elems, depth_frame_stride_elems, hf));
    if (depth && !hf.depthInPlace) {
        hf.gray = nullptr;
        const uint16_t* d = depth;
        uint16_t* dDepth = c->d_depthIn + (size_t)slot0 * w * h;
        for (int i = 0; i < n; ++i)
            ORBF_CUDA(c, cudaMemcpy2DAsync(dDepth + (size_t)i * w * h, (size_t)w * 2, d + (size_t)i * depth_frame_stride_elems,
                (size_t)depth_stride_elems * 2, (size_t)w * 2, h, cudaMemcpyHostToDevice, c->stream));
    }
    return run_batch(c, slot0, n, nullptr, nullptr);
}

// BASELINE config 4, 8-level variant (defined by the oracle's orc_extract_adapted): ORB extraction of host frames,
// in video order, with iniThFAST adapted per region of a grid x grid partition by controllers that carry their state from frame to
// frame.  The pyramid, blur and descriptor stages run batched over all frames; FAST -> quadtree -> controller step form a chain per
// frame (frame t + 1's thresholds depend on the keypoints frame t returned), enqueued back to back without touching the host.  With
// V independent videos (cameras / sequences) the chain has T links of V frames each: frame t of video v lives in slot slot0 + t * V + v.
static int extract_adapted_impl(orbf_context* c, int slot0, int V, int T, const uint8_t* gray, int64_t gray_stride, int64_t gray_frame_stride,
    int64_t gray_video_stride, const uint16_t* depth, int64_t depth_stride_elems, int64_t depth_frame_stride_elems, int64_t depth_video_stride_elems,
    const orbf_adaptive_config* cfg, double* thresh, int32_t* region_th, int32_t* region_found)
{
    if (!cfg || !thresh || cfg->grid < 1 || cfg->grid > 5 || V < 1 || T < 1) return ORBF_ERR_ARG;
    const int g2 = cfg->grid * cfg->grid, n = V * T, w = c->cfg.width, h = c->cfg.height;
    HostFrames hf;
    TRY(set_host_inputs(c, slot0, n, gray, gray_stride, gray_frame_stride, depth, depth_stride_elems, depth_frame_stride_elems, hf));
    TRY(orbf_region_tables(c, cfg->grid, n, V));
    if (V == 1) TRY(upload_chunk(c, hf, slot0, 0, n));
    else {
        hf.depthInPlace = false;                               // frames are re-ordered into step-major slots: depth is staged, not sampled in place
        if (depth) { c->cur_depth = c->d_depthIn + (size_t)slot0 * w * h; c->cur_depthPitch = w; c->cur_depthFrameStride = (long long)w * h; }
        for (int v = 0; v < V; ++v)
            for (int t = 0; t < T; ++t) {
                const size_t slot = (size_t)slot0 + (size_t)t * V + v;
                ORBF_CUDA(c, cudaMemcpy2DAsync(c->d_in + slot * c->inPlane, c->inPitch, gray + (size_t)v * gray_video_stride + (size_t)t * gray_frame_stride,
                    gray_stride, w, h, cudaMemcpyHostToDevice, c->stream));
                if (depth)
                    ORBF_CUDA(c, cudaMemcpy2DAsync(c->d_depthIn + slot * w * h, (size_t)w * 2, depth + (size_t)v * depth_video_stride_elems + (size_t)t * depth_frame_stride_elems,
                        (size_t)depth_stride_elems * 2, (size_t)w * 2, h, cudaMemcpyHostToDevice, c->stream));
            }
    }
    std::vector<double> st((size_t)V * 25, 0.0);
    std::vector<int> th((size_t)V * 25, 0);
    for (int v = 0; v < V; ++v)
        for (int r = 0; r < g2; ++r) {
            double s0 = thresh[(size_t)v * g2 + r];
            if (!(s0 > 0)) s0 = cfg->init_th;
            st[(size_t)v * 25 + r] = s0;
            th[(size_t)v * 25 + r] = std::max(c->cfg.min_th_fast, std::min(254, (int)s0));
        }
    ORBF_CUDA(c, cudaMemcpyAsync(c->d_regionState, st.data(), st.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    ORBF_CUDA(c, cudaMemcpyAsync(c->d_regionTh, th.data(), th.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    ORBF_CUDA(c, cudaStreamSynchronize(c->stream));           // st / th are locals
    TRY(orbf_launch_pyramid(c, slot0, n));
    for (int t = 0; t < T; ++t) {
        TRY(orbf_launch_fast(c, slot0 + t * V, V, /*adapted=*/true, /*perVideo=*/true));
        TRY(orbf_launch_quadtree(c, slot0 + t * V, V));
        TRY(orbf_launch_region_control(c, slot0 + t * V, t * V, *cfg, V));
    }
    TRY(orbf_launch_blur(c, slot0, n));
    TRY(orbf_launch_describe(c, slot0, n));
    std::vector<int> log((size_t)n * 2 * g2);
    ORBF_CUDA(c, cudaMemcpyAsync(st.data(), c->d_regionState, st.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    ORBF_CUDA(c, cudaMemcpyAsync(log.data(), c->d_regionLog, log.size() * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    ORBF_CUDA(c, cudaStreamSynchronize(c->stream));
    for (int v = 0; v < V; ++v)
        for (int r = 0; r < g2; ++r) thresh[(size_t)v * g2 + r] = st[(size_t)v * 25 + r];
    for (int i = 0; i < n; ++i)                                // log row i = slot order (step-major): [t][v]
        for (int r = 0; r < g2; ++r) {
            if (region_th) region_th[(size_t)i * g2 + r] = log[(size_t)i * 2 * g2 + r];
            if (region_found) region_found[(size_t)i * g2 + r] = log[(size_t)i * 2 * g2 + g2 + r];
        }
    return ORBF_OK;
}

extern "C" int orbf_extract_adapted(orbf_context* c, int32_t slot0, int32_t n, const uint8_t* gray, int64_t gray_stride, int64_t gray_frame_stride,
    const uint16_t* depth, int64_t depth_stride_elems, int64_t depth_frame_stride_elems, const orbf_adaptive_config* cfg, double* thresh,
    int32_t* region_th, int32_t* region_found)
{
    CTX_ENTER(c);
    return extract_adapted_impl(c, slot0, 1, n, gray, gray_stride, gray_frame_stride, 0, depth, depth_stride_elems, depth_frame_stride_elems, 0, cfg, thresh,
        region_th, region_found);
}

extern "C" int orbf_extract_adapted_videos(orbf_context* c, int32_t slot0, int32_t n_videos, int32_t frames_per_video, const uint8_t* gray, int64_t gray_stride,
    int64_t gray_frame_stride, int64_t gray_video_stride, const uint16_t* depth, int64_t depth_stride_elems, int64_t depth_frame_stride_elems,
    int64_t depth_video_stride_elems, const orbf_adaptive_config* cfg, double* thresh, int32_t* region_th, int32_t* region_found)
{
    CTX_ENTER(c);
    return extract_adapted_impl(c, slot0, n_videos, frames_per_video, gray, gray_stride, gray_frame_stride, gray_video_stride, depth, depth_stride_elems,
        depth_frame_stride_elems, depth_video_stride_elems, cfg, thresh, region_th, region_found);
}

// Multi-GPU frame sharding, host arithmetic only: the contiguous chunk [start, stop) of rank `rank` among `world` ranks, the
// halo frame (the previous rank's last frame, extracted again so that the pair straddling two chunks has an owner), the first frame
// the rank extracts and the global pair indices [pair0, pair1) it owns.  Every pair 0 .. n_frames - 2 has exactly one owner.
extern "C" int orbf_frame_shard(int32_t n_frames, int32_t world, int32_t rank, int32_t* start, int32_t* stop, int32_t* halo, int32_t* first,
    int32_t* pair0, int32_t* pair1)
{
    if (world < 1 || rank < 0 || rank >= world || n_frames < 0) return ORBF_ERR_ARG;
    const int base = n_frames / world, extra = n_frames % world;
    const int s0 = rank * base + std::min(rank, extra);
    const int s1 = s0 + base + (rank < extra ? 1 : 0);
    const int h = (rank > 0 && s1 > s0 && s0 > 0) ? 1 : 0;
    const int f = s0 - h;
    if (start) *start = s0;
    if (stop) *stop = s1;
    if (halo) *halo = h;
    if (first) *first = f;
    if (pair0) *pair0 = s1 > s0 ? f : 0;
    if (pair1) *pair1 = s1 > s0 ? std::max(s1 - 1, f) : 0;
    return ORBF_OK;
}

extern "C" int orbf_download_gray(orbf_context* c, int32_t slot, uint8_t* out, int32_t out_stride)
{
    CTX_ENTER(c);
    if (!out || slot < 0 || slot >= c->B || out_stride < c->cfg.width) return ORBF_ERR_ARG;
    ORBF_CUDA(c, cudaMemcpy2DAsync(out, out_stride, c->d_in + (size_t)slot * c->inPlane, c->inPitch, c->cfg.width, c->cfg.height,
        cudaMemcpyDeviceToHost, c->stream));
    ORBF_CUDA(c, cudaStreamSynchronize(c->stream));
    return ORBF_OK;
}

// Tracking::Track's per-frame loop over a sequence (System/tracking.cpp:38-46, 193-208): extract every frame, then
// Matcher(ratio).KnnMatch(last, cur) and Ransac::Iterate(last, cur) for each consecutive pair, pipelined by chunk.
extern "C" int orbf_track_sequence(orbf_context* c, int32_t slot0, int32_t n, const uint8_t* gray, int64_t gray_stride,
    int64_t gray_frame_stride, const uint16_t* depth, int64_t depth_stride_elems, int64_t depth_frame_stride_elems, float ratio,
    int32_t cross_check, const orbf_ransac_config* ransac_cfg)
{
    CTX_ENTER(c);
    HostFrames hf;
    TRY(set_host_inputs(c, slot0, n, gray, gray_stride, gray_frame_stride, depth, depth_stride_elems, depth_frame_stride_elems, hf));
    TrackArgs ta = { ratio, cross_check != 0, ransac_cfg, 0 };
    return run_batch(c, slot0, n, &hf, &ta);
}

extern "C" int orbf_track_sequence_at(orbf_context* c, int32_t slot0, int32_t pair_slot0, int32_t n, const uint8_t* gray, int64_t gray_stride,
    int64_t gray_frame_stride, const uint16_t* depth, int64_t depth_stride_elems, int64_t depth_frame_stride_elems, float ratio,
    int32_t cross_check, const orbf_ransac_config* ransac_cfg)
{
    CTX_ENTER(c);
    HostFrames hf;
    TRY(set_host_inputs(c, slot0, n, gray, gray_stride, gray_frame_stride, depth, depth_stride_elems, depth_frame_stride_elems, hf));
    TrackArgs ta = { ratio, cross_check != 0, ransac_cfg, pair_slot0 };
    return run_batch(c, slot0, n, &hf, &ta);
}

extern "C" int orbf_read_results_async(orbf_context* c, int32_t slot0, int32_t pair_slot0, int32_t n, int32_t* frame_counts, int32_t* match_counts,
    orbf_ransac_result* ransac, int32_t marker)
{
    CTX_ENTER(c);
    if (n < 1 || slot0 < 0 || slot0 + n > c->B || marker < 0 || marker >= ORBF_MARKERS) return ORBF_ERR_ARG;
    if ((match_counts || ransac) && (n < 2 || pair_slot0 < 0 || pair_slot0 + n - 1 > c->P)) return ORBF_ERR_ARG;
    if (frame_counts) ORBF_CUDA(c, cudaMemcpyAsync(frame_counts, c->d_count + slot0, n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    if (match_counts) ORBF_CUDA(c, cudaMemcpyAsync(match_counts, c->d_matchCount + pair_slot0, (n - 1) * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    if (ransac) ORBF_CUDA(c, cudaMemcpyAsync(ransac, c->d_rres + pair_slot0, (size_t)(n - 1) * sizeof(orbf_ransac_result), cudaMemcpyDeviceToHost, c->stream));
    ORBF_CUDA(c, cudaEventRecord(c->evMarker[marker], c->stream));
    return ORBF_OK;
}

extern "C" int orbf_read_features_async(orbf_context* c, int32_t slot0, int32_t pair_slot0, int32_t n, orbf_keypoint* kps, uint8_t* desc, float* xyz,
    orbf_dmatch* matches, int32_t marker)
{
    CTX_ENTER(c);
    if (n < 1 || slot0 < 0 || slot0 + n > c->B || marker < 0 || marker >= ORBF_MARKERS) return ORBF_ERR_ARG;
    if (matches && (n < 2 || pair_slot0 < 0 || pair_slot0 + n - 1 > c->P)) return ORBF_ERR_ARG;
    const size_t K = c->K, o = (size_t)slot0 * K, N = (size_t)n * K;
    if (kps) {
        TRY(orbf_launch_pack_aos(c, slot0, n));
        ORBF_CUDA(c, cudaMemcpyAsync(kps, c->d_kpAos + o, N * sizeof(orbf_keypoint), cudaMemcpyDeviceToHost, c->stream));
    }
    if (desc) ORBF_CUDA(c, cudaMemcpyAsync(desc, c->d_desc + o * 32, N * 32, cudaMemcpyDeviceToHost, c->stream));
    if (xyz) {
        ORBF_CUDA(c, cudaMemcpyAsync(xyz, c->d_ptx + o, N * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
        ORBF_CUDA(c, cudaMemcpyAsync(xyz + N, c->d_pty + o, N * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
        ORBF_CUDA(c, cudaMemcpyAsync(xyz + 2 * N, c->d_ptz + o, N * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    }
    if (matches) ORBF_CUDA(c, cudaMemcpyAsync(matches, c->d_matches + (size_t)pair_slot0 * K, (size_t)(n - 1) * K * sizeof(orbf_dmatch), cudaMemcpyDeviceToHost, c->stream));
    ORBF_CUDA(c, cudaEventRecord(c->evMarker[marker], c->stream));
    return ORBF_OK;
}

extern "C" int orbf_wait_marker(orbf_context* c, int32_t marker)
{
    CTX_ENTER(c);
    if (marker < 0 || marker >= ORBF_MARKERS) return ORBF_ERR_ARG;
    ORBF_CUDA(c, cudaEventSynchronize(c->evMarker[marker]));
    return ORBF_OK;
}

extern "C" int orbf_track_sequence_device_at(orbf_context* c, int32_t slot0, int32_t pair_slot0, int32_t n, const uint8_t* d_gray, int64_t gray_pitch,
    int64_t gray_frame_stride, const uint16_t* d_depth, int64_t depth_pitch_elems, int64_t depth_frame_stride_elems, float ratio,
    int32_t cross_check, const orbf_ransac_config* ransac_cfg)
{
    CTX_ENTER_NOJOIN(c);
    // a RANSAC of an earlier call may still run on the side stream: it reads the frame / pair slots of THAT call, so this one only has
    // to wait for it when the slot ranges overlap (or when it would run its own RANSAC on the context stream)
    if (c->hiPending && (!c->cfg.pipeline_overlap || !ransac_cfg || (slot0 < c->hiSlot0 + c->hiN && c->hiSlot0 < slot0 + n)
            || (pair_slot0 < c->hiPair0 + c->hiNPairs && c->hiPair0 < pair_slot0 + n - 1))) TRY(orbf_join_side(c));
    TRY(set_device_inputs(c, slot0, n, d_gray, gray_pitch, gray_frame_stride, d_depth, depth_pitch_elems, depth_frame_stride_elems));
    TrackArgs ta = { ratio, cross_check != 0, ransac_cfg, pair_slot0 };
    return run_batch(c, slot0, n, nullptr, &ta);
}

extern "C" int orbf_track_sequence_device(orbf_context* c, int32_t slot0, int32_t n, const uint8_t* d_gray, int64_t gray_pitch,
    int64_t gray_frame_stride, const uint16_t* d_depth, int64_t depth_pitch_elems, int64_t depth_frame_stride_elems, float ratio,
    int32_t cross_check, const orbf_ransac_config* ransac_cfg)
{
    return orbf_track_sequence_device_at(c, slot0, 0, n, d_gray, gray_pitch, gray_frame_stride, d_depth, depth_pitch_elems, depth_frame_stride_elems, ratio,
        cross_check, ransac_cfg);
}

extern "C" int orbf_frame_counts(orbf_context* c, int32_t slot0, int32_t n, int32_t* counts)
{
    CTX_ENTER(c);
    if (!counts || n < 1 || slot0 < 0 || slot0 + n > c->B) return ORBF_ERR_ARG;
    ORBF_CUDA(c, cudaMemcpyAsync(counts, c->d_count + slot0, n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    ORBF_CUDA(c, cudaStreamSynchronize(c->stream));
    return ORBF_OK;
}

extern "C" int orbf_download_frame(orbf_context* c, int32_t slot, orbf_keypoint* kps, uint8_t* desc, float* xyz, int32_t cap,
    int32_t* n_out)
{
    CTX_ENTER(c);
    if (!n_out || slot < 0 || slot >= c->B) return ORBF_ERR_ARG;
    if (kps) TRY(orbf_launch_pack_aos(c, slot, 1));
    // the count and every array at full capacity (76 KB at K = 1056) into page-locked memory, then ONE synchronisation
    const size_t o = (size_t)slot * c->K, K = (size_t)c->K;
    ORBF_CUDA(c, cudaMemcpyAsync(c->h_counts, c->d_count + slot, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    if (kps) ORBF_CUDA(c, cudaMemcpyAsync(c->h_kp, c->d_kpAos + o, K * sizeof(orbf_keypoint), cudaMemcpyDeviceToHost, c->stream));
    if (desc) ORBF_CUDA(c, cudaMemcpyAsync(c->h_desc, c->d_desc + o * 32, K * 32, cudaMemcpyDeviceToHost, c->stream));
    if (xyz) {
        ORBF_CUDA(c, cudaMemcpyAsync(c->h_xyz, c->d_ptx + o, K * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
        ORBF_CUDA(c, cudaMemcpyAsync(c->h_xyz + K, c->d_pty + o, K * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
        ORBF_CUDA(c, cudaMemcpyAsync(c->h_xyz + 2 * K, c->d_ptz + o, K * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    }
    ORBF_CUDA(c, cudaStreamSynchronize(c->stream));
    const int n = c->h_counts[0];
    *n_out = n;
    if (n > cap) return ORBF_ERR_CAPACITY;
    if (n == 0) return ORBF_OK;
    if (kps) memcpy(kps, c->h_kp, n * sizeof(orbf_keypoint));
    if (desc) memcpy(desc, c->h_desc, (size_t)n * 32);
    if (xyz) for (int i = 0; i < n; ++i) { xyz[3 * i] = c->h_xyz[i]; xyz[3 * i + 1] = c->h_xyz[K + i]; xyz[3 * i + 2] = c->h_xyz[2 * K + i]; }
    return ORBF_OK;
}

extern "C" int orbf_download_keys_un(orbf_context* c, int32_t slot, float* xy, float* u_right, int32_t cap, int32_t* n_out)
{
    CTX_ENTER(c);
    if (!n_out || slot < 0 || slot >= c->B) return ORBF_ERR_ARG;
    ORBF_CUDA(c, cudaMemcpyAsync(c->h_counts, c->d_count + slot, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    ORBF_CUDA(c, cudaStreamSynchronize(c->stream));
    const int n = c->h_counts[0];
    *n_out = n;
    if (n > cap) return ORBF_ERR_CAPACITY;
    if (n == 0) return ORBF_OK;
    const size_t o = (size_t)slot * c->K;
    if (xy) {
        ORBF_CUDA(c, cudaMemcpyAsync(c->h_xyz, c->d_kpux + o, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
        ORBF_CUDA(c, cudaMemcpyAsync(c->h_xyz + c->K, c->d_kpuy + o, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    }
    if (u_right) ORBF_CUDA(c, cudaMemcpyAsync(c->h_xyz + 2 * c->K, c->d_uright + o, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    ORBF_CUDA(c, cudaStreamSynchronize(c->stream));
    if (xy) for (int i = 0; i < n; ++i) { xy[2 * i] = c->h_xyz[i]; xy[2 * i + 1] = c->h_xyz[c->K + i]; }
    if (u_right) memcpy(u_right, c->h_xyz + 2 * c->K, n * sizeof(float));
    return ORBF_OK;
}

extern "C" int orbf_extract(orbf_context* c, const uint8_t* img, int32_t width, int32_t height, int32_t stride, orbf_keypoint* kps,
    uint8_t* desc, int32_t cap, int32_t* n_out)
{
    CTX_ENTER(c);
    if (!n_out) return ORBF_ERR_ARG;
    *n_out = 0;
    if (!img || width == 0 || height == 0) return ORBF_OK;     // _image.empty(): outputs untouched (orbextractor.cpp:758-759)
    if (width != c->cfg.width || height != c->cfg.height || stride < width) return ORBF_ERR_ARG;
    TRY(orbf_extract_batch(c, 0, 1, img, stride, (int64_t)stride * height, nullptr, 0, 0));
    return orbf_download_frame(c, 0, kps, desc, nullptr, cap, n_out);
}

extern "C" int orbf_pyramid_level(orbf_context* c, int32_t slot, int32_t level, int32_t blurred, uint8_t* out, int32_t out_stride)
{
    CTX_ENTER(c);
    if (!out || slot < 0 || slot >= c->B || level < 0 || level >= c->L || out_stride < c->lg[level].w) return ORBF_ERR_ARG;
    if (!blurred && level == 0 && !c->cur_gray) return ORBF_ERR_STATE;
    PyrView pv = orbf_pyr_view(c, blurred != 0);
    const LevelView& lv = pv.lv[level];
    ORBF_CUDA(c, cudaMemcpy2DAsync(out, out_stride, lv.base + (long long)slot * lv.frameStride, lv.pitch, lv.w, lv.h,
        cudaMemcpyDeviceToHost, c->stream));
    ORBF_CUDA(c, cudaStreamSynchronize(c->stream));
    return ORBF_OK;
}

extern "C" int orbf_level_candidates(orbf_context* c, int32_t slot, int32_t level, orbf_cand* out, int32_t cap, int32_t* n_out)
{
    CTX_ENTER(c);
    if (!n_out || slot < 0 || slot >= c->B || level < 0 || level >= c->L) return ORBF_ERR_ARG;
    int n = 0;
    ORBF_CUDA(c, cudaMemcpyAsync(&n, c->d_candCount + slot * ORBF_MAX_LEVELS + level, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    ORBF_CUDA(c, cudaStreamSynchronize(c->stream));
    *n_out = n;
    if (n > cap) return ORBF_ERR_CAPACITY;
    if (n == 0 || !out) return ORBF_OK;
    std::vector<uint32_t> tmp(n);
    ORBF_CUDA(c, cudaMemcpyAsync(tmp.data(), c->d_cand + (size_t)slot * c->candTotal + c->lg[level].candOff, n * sizeof(uint32_t),
        cudaMemcpyDeviceToHost, c->stream));
    ORBF_CUDA(c, cudaStreamSynchronize(c->stream));
    for (int i = 0; i < n; ++i) { out[i].x = tmp[i] & 0x7FF; out[i].y = (tmp[i] >> 11) & 0x7FF; out[i].score = tmp[i] >> 22; }
    return ORBF_OK;
}

extern "C" int orbf_level_keypoint_counts(orbf_context* c, int32_t slot, int32_t* counts)
{
    CTX_ENTER(c);
    if (!counts || slot < 0 || slot >= c->B) return ORBF_ERR_ARG;
    ORBF_CUDA(c, cudaMemcpyAsync(counts, c->d_lkpCount + slot * ORBF_MAX_LEVELS, c->L * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    ORBF_CUDA(c, cudaStreamSynchronize(c->stream));
    return ORBF_OK;
}

// ------------------------------------------------------------------------------------------------------
// matching
// ------------------------------------------------------------------------------------------------------
static int ensure_desc_stage(orbf_context* c, int rows)
{
    if (rows <= c->descStageRows) return ORBF_OK;
    if (c->d_qdesc) cudaFree(c->d_qdesc);
    if (c->d_tdesc) cudaFree(c->d_tdesc);
    c->d_qdesc = c->d_tdesc = nullptr; c->descStageRows = 0;
    ORBF_CUDA(c, cudaMalloc((void**)&c->d_qdesc, (size_t)rows * 32));
    ORBF_CUDA(c, cudaMalloc((void**)&c->d_tdesc, (size_t)rows * 32));
    c->descStageRows = rows;
    return ORBF_OK;
}

static int standalone_knn(orbf_context* c, const uint8_t* q, int nq, const uint8_t* t, int nt, bool cross, MatchSet& ms)
{
    // pair slot 0 holds the result; K bounds both sets because the packed key stores a 16-bit index per slot row
    if (nq < 0 || nt < 0 || nq > c->K || nt > 65535) return ORBF_ERR_ARG;
    TRY(ensure_desc_stage(c, std::max(std::max(nq, nt), 1)));
    // through the page-locked arena: both uploads and the result read-back of the calling function are asynchronous, the call
    // synchronises once (the arena also has room for the matches that come back: see orbf_knn_match)
    TRY(arena_begin(c, ((size_t)nq + nt) * 32 + (size_t)c->K * sizeof(orbf_dmatch) + 4096));
    if (nq) { uint8_t* hq = arena_take(c, (size_t)nq * 32); memcpy(hq, q, (size_t)nq * 32); ORBF_CUDA(c, cudaMemcpyAsync(c->d_qdesc, hq, (size_t)nq * 32, cudaMemcpyHostToDevice, c->stream)); }
    if (nt) { uint8_t* ht = arena_take(c, (size_t)nt * 32); memcpy(ht, t, (size_t)nt * 32); ORBF_CUDA(c, cudaMemcpyAsync(c->d_tdesc, ht, (size_t)nt * 32, cudaMemcpyHostToDevice, c->stream)); }
    ms.qdesc = c->d_qdesc; ms.tdesc = c->d_tdesc; ms.qStride = ms.tStride = 0; ms.qCounts = ms.tCounts = nullptr; ms.pairs = nullptr;
    ms.pair0 = 0;
    ms.nq = nq; ms.nt = nt; ms.knn = c->d_knn; ms.rev = c->d_rev; ms.matches = c->d_matches; ms.matchCount = c->d_matchCount;
    if (cross && (long long)nt > (long long)c->P * c->K) return ORBF_ERR_ARG;   // rev buffer holds P*K rows
    return orbf_launch_knn2(c, ms, 1, cross);
}

static void unpack_knn(const uint32_t* kk, int nq, int32_t* idx1, int32_t* d1, int32_t* idx2, int32_t* d2)
{
    for (int i = 0; i < nq; ++i) {
        const uint32_t a = kk[2 * i], b = kk[2 * i + 1];
        if (idx1) idx1[i] = (a == 0xFFFFFFFFu) ? -1 : (int)(a & 0xFFFF);
        if (d1) d1[i] = (a == 0xFFFFFFFFu) ? -1 : (int)(a >> 16);
        if (idx2) idx2[i] = (b == 0xFFFFFFFFu) ? -1 : (int)(b & 0xFFFF);
        if (d2) d2[i] = (b == 0xFFFFFFFFu) ? -1 : (int)(b >> 16);
    }
}

extern "C" int orbf_knn2(orbf_context* c, const uint8_t* q, int32_t nq, const uint8_t* t, int32_t nt, int32_t* idx1, int32_t* d1,
    int32_t* idx2, int32_t* d2)
{
    CTX_ENTER(c);
    if ((nq > 0 && !q) || (nt > 0 && !t)) return ORBF_ERR_ARG;
    if (nq == 0) return ORBF_OK;
    if (nt == 0) { unpack_knn(std::vector<uint32_t>(2 * (size_t)nq, 0xFFFFFFFFu).data(), nq, idx1, d1, idx2, d2); return ORBF_OK; }
    MatchSet ms;
    TRY(standalone_knn(c, q, nq, t, nt, false, ms));
    std::vector<uint32_t> kk(2 * (size_t)nq);
    ORBF_CUDA(c, cudaMemcpyAsync(kk.data(), c->d_knn, kk.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    ORBF_CUDA(c, cudaStreamSynchronize(c->stream));
    unpack_knn(kk.data(), nq, idx1, d1, idx2, d2);
    return ORBF_OK;
}

extern "C" int orbf_knn_match(orbf_context* c, const uint8_t* q, int32_t nq, const uint8_t* t, int32_t nt, float ratio,
    int32_t cross_check, orbf_dmatch* out, int32_t cap, int32_t* n_out)
{
    CTX_ENTER(c);
    if (!n_out || (nq > 0 && !q) || (nt > 0 && !t)) return ORBF_ERR_ARG;
    *n_out = 0;
    if (nq == 0 || nt < 2) return ORBF_OK;      // the reference indexes matchesKnn[i][1] (matcher.cpp:65): needs >= 2 train rows
    MatchSet ms;
    TRY(standalone_knn(c, q, nq, t, nt, cross_check != 0, ms));
    TRY(orbf_launch_match_select(c, ms, 1, ratio, cross_check != 0));
    c->lastNPairs = 1; c->pairsFromSlots = false;
    // count and matches (at most nq of them) into the arena, one synchronisation
    int* hn = reinterpret_cast<int*>(arena_take(c, sizeof(int)));
    orbf_dmatch* hm = reinterpret_cast<orbf_dmatch*>(arena_take(c, (size_t)nq * sizeof(orbf_dmatch)));
    ORBF_CUDA(c, cudaMemcpyAsync(hn, c->d_matchCount, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    ORBF_CUDA(c, cudaMemcpyAsync(hm, c->d_matches, (size_t)nq * sizeof(orbf_dmatch), cudaMemcpyDeviceToHost, c->stream));
    ORBF_CUDA(c, cudaStreamSynchronize(c->stream));
    *n_out = *hn;
    if (*hn > cap) return ORBF_ERR_CAPACITY;
    if (*hn && out) memcpy(out, hm, (size_t)*hn * sizeof(orbf_dmatch));
    return ORBF_OK;
}

// Landmark::ComputeDistinctiveDescriptors (Core/landmark.cpp:219-273) for a batch of landmarks: desc holds every landmark's
// observed descriptors back to back, offsets [n+1] delimits them; best[l] = row (inside landmark l) with the least median
// Hamming distance to the others (-1 without observations), median[l] optional.  At most 128 observations per landmark count.
#define SC_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return orbf_cuda_fail(c, e__, #call, __FILE__, __LINE__); } while (0)
#define SC_H2D(off, src, bytes) do { if ((bytes) > 0) SC_CUDA(cudaMemcpyAsync(sc.base + (off), (src), (bytes), cudaMemcpyHostToDevice, c->stream)); } while (0)

extern "C" int orbf_distinctive_descriptors(orbf_context* c, const uint8_t* desc, const int32_t* offsets, int32_t n_landmarks, int32_t* best,
    int32_t* median)
{
    CTX_ENTER(c);
    if (!offsets || !best || n_landmarks < 0 || (n_landmarks > 0 && offsets[n_landmarks] > 0 && !desc)) return ORBF_ERR_ARG;
    if (n_landmarks == 0) return ORBF_OK;
    for (int l = 0; l < n_landmarks; ++l) if (offsets[l + 1] < offsets[l] || offsets[l] < 0) return ORBF_ERR_ARG;
    const size_t rows = (size_t)offsets[n_landmarks], L = (size_t)n_landmarks;
    Scratch sc(c);
    const size_t oDesc = sc.take(rows * 32), oOff = sc.take((L + 1) * sizeof(int)), oBest = sc.take(L * 2 * sizeof(int));
    SC_CUDA(sc.alloc());
    SC_H2D(oDesc, desc, rows * 32); SC_H2D(oOff, offsets, (L + 1) * sizeof(int));
    int* dBest = sc.at<int>(oBest);
    TRY(orbf_launch_distinctive(c, sc.at<uint8_t>(oDesc), sc.at<int>(oOff), n_landmarks, dBest, dBest + L));
    SC_CUDA(cudaMemcpyAsync(best, dBest, L * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    if (median) SC_CUDA(cudaMemcpyAsync(median, dBest + L, L * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    SC_CUDA(cudaStreamSynchronize(c->stream));
    return ORBF_OK;
}

extern "C" int orbf_fuse_search(orbf_context* c, int32_t slot, const float* Rcw, const float* tcw, const float* camera, const float* kp_x, const float* kp_y,
    const float* u_right, const uint8_t* desc, int32_t n_feat, const float* lm_pos, const uint8_t* lm_desc, const uint8_t* lm_valid, int32_t n_landmarks,
    float radius, int32_t th_low, int32_t* best_idx, int32_t* best_dist)
{
    CTX_ENTER(c);
    if (!Rcw || !tcw || !camera || !best_idx || n_landmarks < 0 || slot >= c->B) return ORBF_ERR_ARG;
    if (n_landmarks > 0 && (!lm_pos || !lm_desc || !lm_valid)) return ORBF_ERR_ARG;
    if (slot < 0 && (n_feat < 0 || n_feat > 65535 || (n_feat > 0 && (!kp_x || !kp_y || !u_right || !desc)))) return ORBF_ERR_ARG;
    if (n_landmarks == 0) return ORBF_OK;
    int nFeat = n_feat;
    if (slot >= 0) { SC_CUDA(cudaMemcpyAsync(&nFeat, c->d_count + slot, sizeof(int), cudaMemcpyDeviceToHost, c->stream)); SC_CUDA(cudaStreamSynchronize(c->stream)); }
    const size_t L = (size_t)n_landmarks, F = (size_t)std::max(nFeat, 0);
    Scratch sc(c);
    const size_t oPos = sc.take(L * 12), oLd = sc.take(L * 32), oVal = sc.take(L), oOut = sc.take(L * 8);
    const size_t oKx = sc.take(F * 4), oKy = sc.take(F * 4), oUr = sc.take(F * 4), oDesc = sc.take(F * 32);
    SC_CUDA(sc.alloc());
    SC_H2D(oPos, lm_pos, L * 12); SC_H2D(oLd, lm_desc, L * 32); SC_H2D(oVal, lm_valid, L);
    const float *dKx, *dKy, *dUr; const uint8_t* dDesc;
    if (slot >= 0) {
        dKx = c->d_kpux + (size_t)slot * c->K; dKy = c->d_kpuy + (size_t)slot * c->K; dUr = c->d_uright + (size_t)slot * c->K; dDesc = c->d_desc + (size_t)slot * c->K * 32;
    } else {
        SC_H2D(oKx, kp_x, F * 4); SC_H2D(oKy, kp_y, F * 4); SC_H2D(oUr, u_right, F * 4); SC_H2D(oDesc, desc, F * 32);
        dKx = sc.at<float>(oKx); dKy = sc.at<float>(oKy); dUr = sc.at<float>(oUr); dDesc = sc.at<uint8_t>(oDesc);
    }
    int* dOut = sc.at<int>(oOut);
    TRY(orbf_launch_fuse_search(c, Rcw, tcw, camera, dKx, dKy, dUr, dDesc, nFeat, sc.at<float>(oPos), sc.at<uint8_t>(oLd), sc.at<uint8_t>(oVal), n_landmarks, radius,
        th_low, dOut, dOut + L));
    SC_CUDA(cudaMemcpyAsync(best_idx, dOut, L * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    if (best_dist) SC_CUDA(cudaMemcpyAsync(best_dist, dOut + L, L * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    SC_CUDA(cudaStreamSynchronize(c->stream));
    return ORBF_OK;
}

extern "C" int orbf_bow_match(orbf_context* c, const int32_t* words1, const int32_t* off1, const int32_t* idx1, int32_t nw1, const uint8_t* desc1, int32_t n1,
    const int32_t* words2, const int32_t* off2, const int32_t* idx2, int32_t nw2, const uint8_t* desc2, int32_t n2, float nn_ratio, int32_t th_low,
    orbf_dmatch* out, int32_t cap, int32_t* n_out)
{
    CTX_ENTER(c);
    if (!n_out || nw1 < 0 || nw2 < 0 || n1 < 0 || n2 < 0 || n2 > 65535 || cap < 0 || (cap > 0 && !out)) return ORBF_ERR_ARG;
    *n_out = 0;
    if (nw1 == 0 || nw2 == 0) return ORBF_OK;
    if (!words1 || !off1 || !words2 || !off2) return ORBF_ERR_ARG;
    const int e1 = off1[nw1], e2 = off2[nw2];
    if (e1 < 0 || e2 < 0 || (e1 > 0 && (!idx1 || !desc1)) || (e2 > 0 && (!idx2 || !desc2))) return ORBF_ERR_ARG;
    for (int a = 0; a < nw1; ++a) if (off1[a + 1] < off1[a] || (a > 0 && words1[a] <= words1[a - 1])) return ORBF_ERR_ARG;
    for (int b = 0; b < nw2; ++b) if (off2[b + 1] < off2[b] || (b > 0 && words2[b] <= words2[b - 1]) || off2[b + 1] - off2[b] > 65535) return ORBF_ERR_ARG;
    for (int e = 0; e < e1; ++e) if (idx1[e] < 0 || idx1[e] >= n1) return ORBF_ERR_ARG;
    for (int e = 0; e < e2; ++e) if (idx2[e] < 0 || idx2[e] >= n2) return ORBF_ERR_ARG;
    if (e1 == 0 || e2 == 0) return ORBF_OK;
    Scratch sc(c);
    const size_t oW1 = sc.take((size_t)nw1 * 4), oO1 = sc.take(((size_t)nw1 + 1) * 4), oI1 = sc.take((size_t)e1 * 4), oD1 = sc.take((size_t)n1 * 32);
    const size_t oW2 = sc.take((size_t)nw2 * 4), oO2 = sc.take(((size_t)nw2 + 1) * 4), oI2 = sc.take((size_t)e2 * 4), oD2 = sc.take((size_t)n2 * 32);
    const size_t oET = sc.take((size_t)e1 * 4), oED = sc.take((size_t)e1 * 4), oFU = sc.take((size_t)n2 * 4), oOut = sc.take((size_t)e1 * sizeof(orbf_dmatch)), oN = sc.take(4);
    SC_CUDA(sc.alloc());
    SC_H2D(oW1, words1, (size_t)nw1 * 4); SC_H2D(oO1, off1, ((size_t)nw1 + 1) * 4); SC_H2D(oI1, idx1, (size_t)e1 * 4); SC_H2D(oD1, desc1, (size_t)n1 * 32);
    SC_H2D(oW2, words2, (size_t)nw2 * 4); SC_H2D(oO2, off2, ((size_t)nw2 + 1) * 4); SC_H2D(oI2, idx2, (size_t)e2 * 4); SC_H2D(oD2, desc2, (size_t)n2 * 32);
    TRY(orbf_launch_bow_match(c, sc.at<int>(oW1), sc.at<int>(oO1), sc.at<int>(oI1), nw1, sc.at<uint8_t>(oD1), sc.at<int>(oW2), sc.at<int>(oO2), sc.at<int>(oI2), nw2,
        sc.at<uint8_t>(oD2), n2, nn_ratio, th_low, e1, sc.at<int>(oET), sc.at<int>(oED), sc.at<int>(oFU), sc.at<orbf_dmatch>(oOut), sc.at<int>(oN)));
    int n = 0;
    SC_CUDA(cudaMemcpyAsync(&n, sc.at<int>(oN), sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    SC_CUDA(cudaStreamSynchronize(c->stream));
    *n_out = n;
    if (n > cap) return ORBF_ERR_CAPACITY;
    if (n > 0) { SC_CUDA(cudaMemcpyAsync(out, sc.at<orbf_dmatch>(oOut), (size_t)n * sizeof(orbf_dmatch), cudaMemcpyDeviceToHost, c->stream)); SC_CUDA(cudaStreamSynchronize(c->stream)); }
    return ORBF_OK;
}

extern "C" int orbf_compose_trajectory(orbf_context* c, int32_t npairs, const float* pose0, float* poses, uint8_t* outlier)
{
    CTX_ENTER(c);
    if (!poses || npairs < 0 || npairs > c->P || npairs > c->lastNPairs) return ORBF_ERR_ARG;
    static const float eye[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    Scratch sc(c);
    const size_t oP0 = sc.take(64), oPoses = sc.take(((size_t)npairs + 1) * 64), oOut = sc.take(outlier ? ((size_t)npairs + 1) * c->K : 1);
    SC_CUDA(sc.alloc());
    SC_H2D(oP0, pose0 ? pose0 : eye, 64);
    TRY(orbf_launch_compose(c, npairs, sc.at<float>(oP0), sc.at<float>(oPoses), outlier ? sc.at<uint8_t>(oOut) : nullptr));
    SC_CUDA(cudaMemcpyAsync(poses, sc.at<float>(oPoses), ((size_t)npairs + 1) * 64, cudaMemcpyDeviceToHost, c->stream));
    if (outlier) SC_CUDA(cudaMemcpyAsync(outlier, sc.at<uint8_t>(oOut), ((size_t)npairs + 1) * c->K, cudaMemcpyDeviceToHost, c->stream));
    SC_CUDA(cudaStreamSynchronize(c->stream));
    return ORBF_OK;
}

extern "C" int orbf_undistort_points(orbf_context* c, const float* xy, int32_t n, float fx, float fy, float cx, float cy, const float* dist, float* out)
{
    CTX_ENTER(c);
    if (n < 0 || !dist || (n > 0 && (!xy || !out))) return ORBF_ERR_ARG;
    if (n == 0) return ORBF_OK;
    Scratch sc(c);
    const size_t oIn = sc.take((size_t)n * 8), oOut = sc.take((size_t)n * 8);
    SC_CUDA(sc.alloc());
    SC_H2D(oIn, xy, (size_t)n * 8);
    TRY(orbf_launch_undistort(c, sc.at<float>(oIn), n, fx, fy, cx, cy, dist, sc.at<float>(oOut)));
    SC_CUDA(cudaMemcpyAsync(out, sc.at<float>(oOut), (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
    SC_CUDA(cudaStreamSynchronize(c->stream));
    return ORBF_OK;
}

extern "C" int orbf_unproject_keypoints(orbf_context* c, const orbf_keypoint* kps, int32_t n, const uint16_t* depth, int32_t width, int32_t height,
    int64_t depth_stride_elems, float* xyz, float* u_right, float* xy_un)
{
    CTX_ENTER(c);
    if (n < 0 || (n > 0 && (!kps || !xyz || !u_right || !xy_un)) || (depth && depth_stride_elems < width)) return ORBF_ERR_ARG;
    if (n == 0) return ORBF_OK;
    // the caller's plane is indexed here (one sample per keypoint at the truncated, distorted position, Core/frame.cpp:152-155): the
    // n samples cross the link instead of the plane; all arithmetic (undistortion, unprojection) runs on the device
    std::vector<float> xy((size_t)n * 2); std::vector<uint16_t> raw((size_t)n, 0);
    for (int i = 0; i < n; ++i) {
        xy[2 * i] = kps[i].x; xy[2 * i + 1] = kps[i].y;
        const int u = (int)kps[i].x, v = (int)kps[i].y;
        if (depth && u >= 0 && v >= 0 && u < width && v < height) raw[i] = depth[(int64_t)v * depth_stride_elems + u];
    }
    Scratch sc(c);
    const size_t oXy = sc.take((size_t)n * 8), oRaw = sc.take((size_t)n * 2), oXyz = sc.take((size_t)n * 12), oUr = sc.take((size_t)n * 4), oUn = sc.take((size_t)n * 8);
    SC_CUDA(sc.alloc());
    SC_H2D(oXy, xy.data(), (size_t)n * 8); SC_H2D(oRaw, raw.data(), (size_t)n * 2);
    TRY(orbf_launch_unproject(c, sc.at<float>(oXy), sc.at<uint16_t>(oRaw), n, sc.at<float>(oXyz), sc.at<float>(oUr), sc.at<float>(oUn)));
    SC_CUDA(cudaMemcpyAsync(xyz, sc.at<float>(oXyz), (size_t)n * 12, cudaMemcpyDeviceToHost, c->stream));
    SC_CUDA(cudaMemcpyAsync(u_right, sc.at<float>(oUr), (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
    SC_CUDA(cudaMemcpyAsync(xy_un, sc.at<float>(oUn), (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
    SC_CUDA(cudaStreamSynchronize(c->stream));
    return ORBF_OK;
}

extern "C" int orbf_projection_match(orbf_context* c, int32_t slot, const float* kp_x, const float* kp_y, const int32_t* kp_octave, const uint8_t* desc,
    int32_t n_feat, const uint8_t* lm_desc, const float* proj_x, const float* proj_y, const uint8_t* lm_flags, int32_t n_landmarks,
    const uint8_t* feat_taken, float radius, float nn_ratio, int32_t th_high, int32_t* best_idx, int32_t* n_matches)
{
    CTX_ENTER(c);
    if (!best_idx || !n_matches || n_landmarks < 0) return ORBF_ERR_ARG;
    if (n_landmarks > 0 && (!lm_desc || !proj_x || !proj_y || !lm_flags)) return ORBF_ERR_ARG;
    if (slot >= c->B) return ORBF_ERR_ARG;
    if (slot < 0 && (n_feat < 0 || n_feat > 65535 || (n_feat > 0 && (!kp_x || !kp_y || !kp_octave || !desc)))) return ORBF_ERR_ARG;
    *n_matches = 0;
    if (n_landmarks == 0) return ORBF_OK;
    int nFeat = n_feat;
    if (slot >= 0) {
        if (cudaMemcpyAsync(&nFeat, c->d_count + slot, sizeof(int), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess || cudaStreamSynchronize(c->stream) != cudaSuccess)
            return orbf_cuda_fail(c, cudaGetLastError(), "projection_match: frame count", __FILE__, __LINE__);
    }
    if (feat_taken == nullptr && nFeat == 0) { for (int i = 0; i < n_landmarks; ++i) best_idx[i] = -1; return ORBF_OK; }
    const size_t L = (size_t)n_landmarks, F = (size_t)std::max(nFeat, 1);
    // the context's persistent scratch: landmark inputs | frame inputs (host route) | candidate lists | outputs
    Scratch sc(c);
    const size_t oLmDesc = sc.take(L * 32), oPx = sc.take(L * 4), oPy = sc.take(L * 4), oFlags = sc.take(L), oTakenIn = sc.take(F);
    const size_t oKx = sc.take(F * 4), oKy = sc.take(F * 4), oOct = sc.take(F * 4), oDesc = sc.take(F * 32);
    const size_t oCand = sc.take(L * F * 4), oCandOct = sc.take(L * 8 * 4), oCnt = sc.take(L * 4), oBest = sc.take((L + 1) * 4);
    SC_CUDA(sc.alloc());
    SC_H2D(oLmDesc, lm_desc, L * 32); SC_H2D(oPx, proj_x, L * 4); SC_H2D(oPy, proj_y, L * 4); SC_H2D(oFlags, lm_flags, L);
    if (feat_taken && nFeat > 0) SC_H2D(oTakenIn, feat_taken, (size_t)nFeat);
    const float *dKx, *dKy; const int* dOct; const uint8_t* dDesc;
    if (slot >= 0) {
        dKx = c->d_kpux + (size_t)slot * c->K; dKy = c->d_kpuy + (size_t)slot * c->K; dOct = c->d_kpoct + (size_t)slot * c->K; dDesc = c->d_desc + (size_t)slot * c->K * 32;
    } else {
        if (nFeat > 0) { SC_H2D(oKx, kp_x, (size_t)nFeat * 4); SC_H2D(oKy, kp_y, (size_t)nFeat * 4); SC_H2D(oOct, kp_octave, (size_t)nFeat * 4); SC_H2D(oDesc, desc, (size_t)nFeat * 32); }
        dKx = sc.at<float>(oKx); dKy = sc.at<float>(oKy); dOct = sc.at<int>(oOct); dDesc = sc.at<uint8_t>(oDesc);
    }
    int* dBest = sc.at<int>(oBest);
    TRY(orbf_launch_projection_match(c, dKx, dKy, dOct, dDesc, nFeat, sc.at<uint8_t>(oLmDesc), sc.at<float>(oPx), sc.at<float>(oPy), sc.at<uint8_t>(oFlags), n_landmarks,
        (feat_taken && nFeat > 0) ? sc.at<uint8_t>(oTakenIn) : nullptr, radius, nn_ratio, th_high, sc.at<uint32_t>(oCand), sc.at<int>(oCandOct), sc.at<int>(oCnt), dBest, dBest + L));
    SC_CUDA(cudaMemcpyAsync(best_idx, dBest, L * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    SC_CUDA(cudaMemcpyAsync(n_matches, dBest + L, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    SC_CUDA(cudaStreamSynchronize(c->stream));
    return ORBF_OK;
}

extern "C" int orbf_descriptor_distance(const uint8_t* a, const uint8_t* b, int32_t nbytes, int32_t* dist)
{
    if (!a || !b || !dist || nbytes < 0) return ORBF_ERR_ARG;
    int d = 0;
    for (int i = 0; i < nbytes; ++i) d += __builtin_popcount((unsigned)(a[i] ^ b[i]));
    *dist = d;
    return ORBF_OK;
}


extern "C" int orbf_match_pairs(orbf_context* c, const int32_t* pairs, int32_t npairs, float ratio, int32_t cross_check)
{
    CTX_ENTER(c);
    if (!pairs || npairs < 1 || npairs > c->P) return ORBF_ERR_ARG;
    for (int i = 0; i < 2 * npairs; ++i) if (pairs[i] < 0 || pairs[i] >= c->B) return ORBF_ERR_ARG;
    ORBF_CUDA(c, cudaMemcpyAsync(c->d_pairs, pairs, (size_t)npairs * 2 * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    c->lastNPairs = npairs; c->pairsFromSlots = true;
    return run_match_group(c, 0, npairs, ratio, cross_check != 0);
}

extern "C" int orbf_download_matches(orbf_context* c, int32_t pair, orbf_dmatch* out, int32_t cap, int32_t* n_out)
{
    CTX_ENTER(c);
    if (!n_out || pair < 0 || pair >= c->P) return ORBF_ERR_ARG;
    int n = 0;
    ORBF_CUDA(c, cudaMemcpyAsync(&n, c->d_matchCount + pair, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    ORBF_CUDA(c, cudaStreamSynchronize(c->stream));
    *n_out = n;
    if (n > cap) return ORBF_ERR_CAPACITY;
    if (n && out) {
        ORBF_CUDA(c, cudaMemcpyAsync(out, c->d_matches + (size_t)pair * c->K, n * sizeof(orbf_dmatch), cudaMemcpyDeviceToHost, c->stream));
        ORBF_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    return ORBF_OK;
}

extern "C" int orbf_download_knn(orbf_context* c, int32_t pair, int32_t* idx1, int32_t* d1, int32_t* idx2, int32_t* d2, int32_t cap,
    int32_t* nq_out)
{
    CTX_ENTER(c);
    if (!nq_out || pair < 0 || pair >= c->P || !c->pairsFromSlots) return ORBF_ERR_ARG;
    int pr[2], nq = 0;
    ORBF_CUDA(c, cudaMemcpyAsync(pr, c->d_pairs + 2 * pair, 2 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    ORBF_CUDA(c, cudaStreamSynchronize(c->stream));
    ORBF_CUDA(c, cudaMemcpyAsync(&nq, c->d_count + pr[0], sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    ORBF_CUDA(c, cudaStreamSynchronize(c->stream));
    *nq_out = nq;
    if (nq > cap) return ORBF_ERR_CAPACITY;
    std::vector<uint32_t> kk(2 * (size_t)std::max(nq, 1));
    ORBF_CUDA(c, cudaMemcpyAsync(kk.data(), c->d_knn + (size_t)pair * c->K * 2, (size_t)nq * 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    ORBF_CUDA(c, cudaStreamSynchronize(c->stream));
    unpack_knn(kk.data(), nq, idx1, d1, idx2, d2);
    return ORBF_OK;
}

extern "C" int orbf_match_counts(orbf_context* c, int32_t npairs, int32_t* counts)
{
    CTX_ENTER(c);
    if (!counts || npairs < 1 || npairs > c->P) return ORBF_ERR_ARG;
    ORBF_CUDA(c, cudaMemcpyAsync(counts, c->d_matchCount, npairs * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    ORBF_CUDA(c, cudaStreamSynchronize(c->stream));
    return ORBF_OK;
}

// ------------------------------------------------------------------------------------------------------
// RANSAC / Kabsch
// ------------------------------------------------------------------------------------------------------
extern "C" int orbf_ransac_pairs(orbf_context* c, int32_t npairs, const orbf_ransac_config* cfg)
{
    CTX_ENTER(c);
    if (!cfg || npairs < 1 || npairs > c->P) return ORBF_ERR_ARG;
    if (!c->pairsFromSlots || npairs > c->lastNPairs) return ORBF_ERR_STATE;
    TRY(orbf_ransac_reserve(c, *cfg));
    return orbf_launch_ransac(c, slot_ransac_set(c), 0, npairs, *cfg, nullptr, false);
}

// The covariance the pairs last matched would latch (quirk Q7) if nothing had been latched before them: the value of the first pair, in
// order, that reaches scoring, or -1.  Sharded sequences use it to agree on the globally first value before any rank scores (the
// context's own latch is neither read nor written).
extern "C" int orbf_ransac_probe_depth_cov(orbf_context* c, int32_t npairs, const orbf_ransac_config* cfg, double* cov)
{
    CTX_ENTER(c);
    if (!cfg || !cov || npairs < 1 || npairs > c->P) return ORBF_ERR_ARG;
    if (!c->pairsFromSlots || npairs > c->lastNPairs) return ORBF_ERR_STATE;
    orbf_ransac_config cf = *cfg;
    cf.depth_cov = -1.0;
    TRY(orbf_ransac_reserve(c, cf));
    TRY(orbf_launch_ransac(c, slot_ransac_set(c), 0, npairs, cf, nullptr, /*standalone=*/true, false, /*probeOnly=*/true));
    ORBF_CUDA(c, cudaMemcpyAsync(cov, c->d_depthCov + 1, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    ORBF_CUDA(c, cudaStreamSynchronize(c->stream));
    return ORBF_OK;
}

extern "C" int orbf_download_ransac(orbf_context* c, int32_t pair, orbf_ransac_result* out, orbf_dmatch* inliers, int32_t cap)
{
    CTX_ENTER(c);
    if (!out || pair < 0 || pair >= c->P) return ORBF_ERR_ARG;
    ORBF_CUDA(c, cudaMemcpyAsync(out, c->d_rres + pair, sizeof(orbf_ransac_result), cudaMemcpyDeviceToHost, c->stream));
    ORBF_CUDA(c, cudaStreamSynchronize(c->stream));
    if (inliers && out->n_inliers > cap) return ORBF_ERR_CAPACITY;
    if (inliers && out->n_inliers > 0) {
        ORBF_CUDA(c, cudaMemcpyAsync(inliers, c->d_inliers + (size_t)pair * c->K, out->n_inliers * sizeof(orbf_dmatch), cudaMemcpyDeviceToHost,
            c->stream));
        ORBF_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    return ORBF_OK;
}

extern "C" int orbf_download_ransac_summary(orbf_context* c, int32_t npairs, orbf_ransac_result* out)
{
    CTX_ENTER(c);
    if (!out || npairs < 1 || npairs > c->P) return ORBF_ERR_ARG;
    ORBF_CUDA(c, cudaMemcpyAsync(out, c->d_rres, npairs * sizeof(orbf_ransac_result), cudaMemcpyDeviceToHost, c->stream));
    ORBF_CUDA(c, cudaStreamSynchronize(c->stream));
    return ORBF_OK;
}

namespace {
// what Odometry::Compute adds to Ransac::Iterate (orbf_odometry_compute): the GICP clouds and the composed pose, queued behind the
// RANSAC chain and brought back by the same single synchronisation
struct OdometryExtras { float* cloudSrc; float* cloudTgt; int cloudCap; int* nCloud; const float* pose1; float* pose2; };
}  // namespace

static int ransac_iterate_core(orbf_context* c, const orbf_ransac_config* cfg, const float* src_xyz, int32_t nsrc,
    const float* dst_xyz, int32_t ndst, const orbf_dmatch* m12, int32_t nm, const int32_t* sample_table, orbf_dmatch* inliers_out,
    int32_t cap, orbf_ransac_result* out, orbf_hyp_trace* hyp_trace, orbf_dmatch* good_sorted_out, int32_t* sample_table_out, const OdometryExtras* ex)
{
    if (!cfg || !out || nsrc < 0 || ndst < 0 || nm < 0 || (nm > 0 && (!m12 || !src_xyz || !dst_xyz))) return ORBF_ERR_ARG;
    if (nm > c->K) return ORBF_ERR_CAPACITY;
    for (int i = 0; i < nm; ++i)
        if (m12[i].queryIdx < 0 || m12[i].queryIdx >= nsrc || m12[i].trainIdx < 0 || m12[i].trainIdx >= ndst) return ORBF_ERR_ARG;
    const int rows = std::max(std::max(nsrc, ndst), 1);
    if (rows > c->xyzStageRows) {
        if (c->d_sxyz) cudaFree(c->d_sxyz);
        if (c->d_txyz) cudaFree(c->d_txyz);
        c->d_sxyz = c->d_txyz = nullptr; c->xyzStageRows = 0;
        ORBF_CUDA(c, cudaMalloc((void**)&c->d_sxyz, (size_t)rows * 3 * sizeof(float)));
        ORBF_CUDA(c, cudaMalloc((void**)&c->d_txyz, (size_t)rows * 3 * sizeof(float)));
        c->xyzStageRows = rows;
    }
    const int R = c->xyzStageRows;
    const int tabN0 = cfg->iterations * (int)cfg->sample_size;
    // every transfer of the call goes through the page-locked arena: uploads, kernels and read-backs are queued without a host
    // wait in between and the call synchronises once, before the results are handed to the caller
    TRY(arena_begin(c, (size_t)R * 24 + (size_t)nm * 3 * sizeof(orbf_dmatch) + (size_t)cfg->iterations * sizeof(orbf_hyp_trace) + (size_t)tabN0 * 8 + 8192
        + (ex ? (size_t)std::max(ex->cloudCap, 0) * 2 * sizeof(float4) + 1024 : 0)));
    {
        float* hs = reinterpret_cast<float*>(arena_take(c, (size_t)R * 12)); float* ht = reinterpret_cast<float*>(arena_take(c, (size_t)R * 12));
        for (int i = 0; i < nsrc; ++i) { hs[i] = src_xyz[3 * i]; hs[R + i] = src_xyz[3 * i + 1]; hs[2 * (size_t)R + i] = src_xyz[3 * i + 2]; }
        for (int i = 0; i < ndst; ++i) { ht[i] = dst_xyz[3 * i]; ht[R + i] = dst_xyz[3 * i + 1]; ht[2 * (size_t)R + i] = dst_xyz[3 * i + 2]; }
        for (int k = 0; k < 3; ++k) {          // rows past nsrc / ndst are never indexed (the matches were range-checked above)
            if (nsrc) ORBF_CUDA(c, cudaMemcpyAsync(c->d_sxyz + (size_t)k * R, hs + (size_t)k * R, (size_t)nsrc * sizeof(float), cudaMemcpyHostToDevice, c->stream));
            if (ndst) ORBF_CUDA(c, cudaMemcpyAsync(c->d_txyz + (size_t)k * R, ht + (size_t)k * R, (size_t)ndst * sizeof(float), cudaMemcpyHostToDevice, c->stream));
        }
        int* hnm = reinterpret_cast<int*>(arena_take(c, sizeof(int)));
        *hnm = nm;
        if (nm) {
            orbf_dmatch* hm = reinterpret_cast<orbf_dmatch*>(arena_take(c, (size_t)nm * sizeof(orbf_dmatch)));
            memcpy(hm, m12, (size_t)nm * sizeof(orbf_dmatch));
            ORBF_CUDA(c, cudaMemcpyAsync(c->d_matches, hm, (size_t)nm * sizeof(orbf_dmatch), cudaMemcpyHostToDevice, c->stream));
        }
        ORBF_CUDA(c, cudaMemcpyAsync(c->d_matchCount, hnm, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    }
    c->pairsFromSlots = false; c->lastNPairs = 1;
    const int tabN = cfg->iterations * (int)cfg->sample_size;
    int* dTab = nullptr;
    if (sample_table) {
        if (tabN > c->userSamplesCap) {
            if (c->d_userSamples) cudaFree(c->d_userSamples);
            c->d_userSamples = nullptr; c->userSamplesCap = 0;
            ORBF_CUDA(c, cudaMalloc((void**)&c->d_userSamples, (size_t)tabN * sizeof(int)));
            c->userSamplesCap = tabN;
        }
        int* htab = reinterpret_cast<int*>(arena_take(c, (size_t)tabN * sizeof(int)));
        memcpy(htab, sample_table, (size_t)tabN * sizeof(int));
        ORBF_CUDA(c, cudaMemcpyAsync(c->d_userSamples, htab, (size_t)tabN * sizeof(int), cudaMemcpyHostToDevice, c->stream));
        dTab = c->d_userSamples;
    }
    const orbf_ransac_config& cf = *cfg;     // standalone: depth_cov < 0 latches within this call only (one oracle call with depth_cov < 0)
    RansacSet rs;
    rs.sx = c->d_sxyz; rs.sy = c->d_sxyz + R; rs.sz = c->d_sxyz + 2 * (size_t)R;
    rs.tx = c->d_txyz; rs.ty = c->d_txyz + R; rs.tz = c->d_txyz + 2 * (size_t)R;
    rs.slotStride = 0; rs.pairs = nullptr; rs.matches = c->d_matches; rs.matchCount = c->d_matchCount; rs.nsrc = nsrc; rs.ndst = ndst;
    // The loop of a well-matched pair ends within its first hypotheses (> 80 % inliers), and every later wave is then a chain of empty
    // launches: the first two waves (8 hypotheses) are queued with the read-back behind them; only a pair whose done flag is still clear
    // after the synchronisation gets the remaining waves and a second read-back.
    const bool twoPhase = !hyp_trace && !sample_table_out && cfg->iterations > 8;
    // results: everything the caller asked for into the arena (inliers / sorted matches at their upper bound nm), one synchronisation
    orbf_ransac_result* hres = reinterpret_cast<orbf_ransac_result*>(arena_take(c, sizeof(orbf_ransac_result)));
    int* hdone = reinterpret_cast<int*>(arena_take(c, sizeof(int)));
    orbf_dmatch* hinl = inliers_out && nm ? reinterpret_cast<orbf_dmatch*>(arena_take(c, (size_t)nm * sizeof(orbf_dmatch))) : nullptr;
    orbf_dmatch* hgood = good_sorted_out && nm ? reinterpret_cast<orbf_dmatch*>(arena_take(c, (size_t)nm * sizeof(orbf_dmatch))) : nullptr;
    orbf_hyp_trace* hhyp = hyp_trace ? reinterpret_cast<orbf_hyp_trace*>(arena_take(c, (size_t)cfg->iterations * sizeof(orbf_hyp_trace))) : nullptr;
    int* htabOut = sample_table_out ? reinterpret_cast<int*>(arena_take(c, (size_t)tabN * sizeof(int))) : nullptr;
    int* hCloudN = nullptr; float4 *hCloudS = nullptr, *hCloudT = nullptr; float* hPose = nullptr;
    size_t cloudM = 0;
    size_t oP0 = 0, oPoses = 0;
    Scratch sc(c);
    if (ex) {
        cloudM = (size_t)std::min(std::max(ex->cloudCap, 0), c->K);
        hCloudN = reinterpret_cast<int*>(arena_take(c, sizeof(int)));
        if (ex->cloudSrc && cloudM) hCloudS = reinterpret_cast<float4*>(arena_take(c, cloudM * sizeof(float4)));
        if (ex->cloudTgt && cloudM) hCloudT = reinterpret_cast<float4*>(arena_take(c, cloudM * sizeof(float4)));
        if (ex->pose2) {
            static const float eye[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
            oP0 = sc.take(64); oPoses = sc.take(128);
            ORBF_CUDA(c, sc.alloc());
            hPose = reinterpret_cast<float*>(arena_take(c, 128));
            memcpy(hPose, ex->pose1 ? ex->pose1 : eye, 64);
            ORBF_CUDA(c, cudaMemcpyAsync(sc.at<float>(oP0), hPose, 64, cudaMemcpyHostToDevice, c->stream));
        }
    }
    const float* dPoseIn = hPose ? sc.at<float>(oP0) : nullptr;              // the select kernel writes pose2 = T12 * pose1 next to the result
    float* dPoseOut = hPose ? sc.at<float>(oPoses) : nullptr;
    TRY(orbf_launch_ransac(c, rs, 0, 1, cf, dTab, /*standalone=*/true, /*fullTable=*/sample_table_out != nullptr, false, 0, twoPhase ? 2 : 5, dPoseIn, dPoseOut));
    for (int phase = 0; phase < 2; ++phase) {
        if (phase == 1) {
            if (!twoPhase || *hdone) break;
            TRY(orbf_launch_ransac(c, rs, 0, 1, cf, dTab, true, false, false, 2, 5, dPoseIn, dPoseOut));
        }
        ORBF_CUDA(c, cudaMemcpyAsync(hres, c->d_rres, sizeof(orbf_ransac_result), cudaMemcpyDeviceToHost, c->stream));
        ORBF_CUDA(c, cudaMemcpyAsync(hdone, orbf_ransac_done_flag(c, 0), sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        if (hinl) ORBF_CUDA(c, cudaMemcpyAsync(hinl, c->d_inliers, (size_t)nm * sizeof(orbf_dmatch), cudaMemcpyDeviceToHost, c->stream));
        if (hgood) ORBF_CUDA(c, cudaMemcpyAsync(hgood, c->d_good, (size_t)nm * sizeof(orbf_dmatch), cudaMemcpyDeviceToHost, c->stream));
        if (hhyp) ORBF_CUDA(c, cudaMemcpyAsync(hhyp, c->d_hyp, (size_t)cfg->iterations * sizeof(orbf_hyp_trace), cudaMemcpyDeviceToHost, c->stream));
        if (htabOut) ORBF_CUDA(c, cudaMemcpyAsync(htabOut, c->d_samples, (size_t)tabN * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        // Odometry::Compute's tail on the same stream, behind the RANSAC chain: mpSourceCloud / mpTargetCloud (ransac.cpp:163-189; they do
        // not depend on the loop) and pose2 = T12 * pose1 (odometry.cpp:82-84)
        if (ex) {
            if (phase == 0) {
                TRY(orbf_launch_ransac_clouds(c, 0, 1));
                ORBF_CUDA(c, cudaMemcpyAsync(hCloudN, c->d_cloudCount, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
                if (hCloudS) ORBF_CUDA(c, cudaMemcpyAsync(hCloudS, c->d_cloudSrc, cloudM * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
                if (hCloudT) ORBF_CUDA(c, cudaMemcpyAsync(hCloudT, c->d_cloudTgt, cloudM * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
            }
            if (hPose) ORBF_CUDA(c, cudaMemcpyAsync(hPose + 16, dPoseOut + 16, 64, cudaMemcpyDeviceToHost, c->stream));
        }
        ORBF_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    *out = *hres;
    if (ex) {
        if (ex->nCloud) *ex->nCloud = *hCloudN;
        const size_t nc = std::min((size_t)std::max(*hCloudN, 0), cloudM);
        if (hCloudS && nc) memcpy(ex->cloudSrc, hCloudS, nc * sizeof(float4));
        if (hCloudT && nc) memcpy(ex->cloudTgt, hCloudT, nc * sizeof(float4));
        if (hPose) memcpy(ex->pose2, hPose + 16, 64);
    }
    if (inliers_out && out->n_inliers > cap) return ORBF_ERR_CAPACITY;
    if (hinl && out->n_inliers > 0) memcpy(inliers_out, hinl, (size_t)out->n_inliers * sizeof(orbf_dmatch));
    if (hgood && out->n_good > 0) memcpy(good_sorted_out, hgood, (size_t)out->n_good * sizeof(orbf_dmatch));
    if (hhyp) memcpy(hyp_trace, hhyp, (size_t)cfg->iterations * sizeof(orbf_hyp_trace));
    if (htabOut) memcpy(sample_table_out, htabOut, (size_t)tabN * sizeof(int));
    if (ex && *hCloudN > ex->cloudCap && (ex->cloudSrc || ex->cloudTgt)) return ORBF_ERR_CAPACITY;
    return ORBF_OK;
}

extern "C" int orbf_ransac_iterate(orbf_context* c, const orbf_ransac_config* cfg, const float* src_xyz, int32_t nsrc,
    const float* dst_xyz, int32_t ndst, const orbf_dmatch* m12, int32_t nm, const int32_t* sample_table, orbf_dmatch* inliers_out,
    int32_t cap, orbf_ransac_result* out, orbf_hyp_trace* hyp_trace, orbf_dmatch* good_sorted_out, int32_t* sample_table_out)
{
    CTX_ENTER(c);
    return ransac_iterate_core(c, cfg, src_xyz, nsrc, dst_xyz, ndst, m12, nm, sample_table, inliers_out, cap, out, hyp_trace, good_sorted_out, sample_table_out, nullptr);
}

// Odometry::Compute, RANSAC strategy (Odometry/odometry.cpp:44-90) for one frame pair in ONE call with one synchronisation:
// Ransac::Iterate, the clouds it leaves for GICP, and pose2 = T12 * pose1 as cv::Mat evaluates it.
extern "C" int orbf_odometry_compute(orbf_context* c, const orbf_ransac_config* cfg, const float* src_xyz, int32_t nsrc, const float* dst_xyz, int32_t ndst,
    const orbf_dmatch* m12, int32_t nm, orbf_dmatch* inliers_out, int32_t cap, orbf_ransac_result* out, float* cloud_src_xyzw, float* cloud_tgt_xyzw,
    int32_t cloud_cap, int32_t* n_cloud, const float* pose1, float* pose2)
{
    CTX_ENTER(c);
    if (cloud_cap < 0) return ORBF_ERR_ARG;
    OdometryExtras ex = { cloud_src_xyzw, cloud_tgt_xyzw, cloud_cap, n_cloud, pose1, pose2 };
    return ransac_iterate_core(c, cfg, src_xyz, nsrc, dst_xyz, ndst, m12, nm, nullptr, inliers_out, cap, out, nullptr, nullptr, nullptr, &ex);
}

extern "C" int orbf_ransac_clouds(orbf_context* c, int32_t pair0, int32_t npairs, const float** d_src_xyzw, const float** d_tgt_xyzw,
    const int32_t** d_counts, int32_t* points_per_pair)
{
    CTX_ENTER(c);
    if (npairs < 1 || pair0 < 0 || pair0 + npairs > c->P) return ORBF_ERR_ARG;
    if (c->lastNPairs <= 0) return ORBF_ERR_STATE;
    TRY(orbf_launch_ransac_clouds(c, pair0, npairs));
    if (d_src_xyzw) *d_src_xyzw = reinterpret_cast<const float*>(c->d_cloudSrc);
    if (d_tgt_xyzw) *d_tgt_xyzw = reinterpret_cast<const float*>(c->d_cloudTgt);
    if (d_counts) *d_counts = c->d_cloudCount;
    if (points_per_pair) *points_per_pair = c->K;
    return ORBF_OK;
}

extern "C" int orbf_download_ransac_clouds(orbf_context* c, int32_t pair, float* src_xyzw, float* tgt_xyzw, int32_t cap, int32_t* n_out)
{
    CTX_ENTER(c);
    if (!n_out || pair < 0 || pair >= c->P || cap < 0) return ORBF_ERR_ARG;
    if (c->lastNPairs <= 0) return ORBF_ERR_STATE;
    TRY(orbf_launch_ransac_clouds(c, pair, 1));
    // the count and both clouds at the caller's capacity in one go, ONE synchronisation (entries past the count are never looked at)
    const size_t m = (size_t)std::min(cap, c->K);
    ORBF_CUDA(c, cudaMemcpyAsync(c->h_counts, c->d_cloudCount + pair, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    if (src_xyzw && m) ORBF_CUDA(c, cudaMemcpyAsync(src_xyzw, c->d_cloudSrc + (size_t)pair * c->K, m * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
    if (tgt_xyzw && m) ORBF_CUDA(c, cudaMemcpyAsync(tgt_xyzw, c->d_cloudTgt + (size_t)pair * c->K, m * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
    ORBF_CUDA(c, cudaStreamSynchronize(c->stream));
    *n_out = c->h_counts[0];
    return *n_out > cap ? ORBF_ERR_CAPACITY : ORBF_OK;
}

extern "C" int orbf_kabsch(orbf_context* c, const float* A, const float* B, int32_t n, float* T16)
{
    CTX_ENTER(c);
    if (!T16 || n < 0 || (n > 0 && (!A || !B))) return ORBF_ERR_ARG;
    const int need = 6 * std::max(n, 1) + 16;
    if (need > c->kabschCap) {
        if (c->d_kabsch) cudaFree(c->d_kabsch);
        c->d_kabsch = nullptr; c->kabschCap = 0;
        ORBF_CUDA(c, cudaMalloc((void**)&c->d_kabsch, (size_t)need * sizeof(float)));
        c->kabschCap = need;
    }
    float* dA = c->d_kabsch + 16; float* dB = dA + 3 * (size_t)std::max(n, 1);
    if (n) {
        ORBF_CUDA(c, cudaMemcpyAsync(dA, A, (size_t)n * 3 * sizeof(float), cudaMemcpyHostToDevice, c->stream));
        ORBF_CUDA(c, cudaMemcpyAsync(dB, B, (size_t)n * 3 * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    }
    TRY(orbf_launch_kabsch(c, dA, dB, n, c->d_kabsch));
    ORBF_CUDA(c, cudaMemcpyAsync(T16, c->d_kabsch, 16 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    ORBF_CUDA(c, cudaStreamSynchronize(c->stream));
    return ORBF_OK;
}

// ---- test hooks for the host+device restatements in replay.h (pinned against libstdc++ / glibc in tests) ----
extern "C" int orbf_selftest_introsort(orbf_dmatch* m, int32_t n)
{
    if (!m || n < 0) return ORBF_ERR_ARG;
    replay::IntroSort<orbf_dmatch, DistanceLess> s{ m, DistanceLess() };
    s.sort(n);
    return ORBF_OK;
}

extern "C" int orbf_selftest_glibc_rand(uint32_t seed, int32_t n, int32_t* out)
{
    if (!out || n < 0) return ORBF_ERR_ARG;
    replay::GlibcRand g;
    g.seed(seed);
    for (int i = 0; i < n; ++i) out[i] = g.next();
    return ORBF_OK;
}

// host restatement of glibc's sinf / cosf (csrc/glibc_sincosf.h) over the floats with bit patterns [lo, hi]: out (optional) receives
// sin then cos of each; n_diff the number of inputs where either differs from THIS machine's libm sinf / cosf
extern "C" int orbf_selftest_sincosf(uint32_t lo, uint32_t hi, float* out, int64_t* n_diff)
{
    if (hi < lo) return ORBF_ERR_ARG;
    int64_t nd = 0;
    for (uint64_t b = lo; b <= hi; ++b) {
        float x; const uint32_t bb = (uint32_t)b; memcpy(&x, &bb, 4);
        const float s = replay::glibc_sinf(x), c = replay::glibc_cosf(x);
        if (out) { out[2 * (b - lo)] = s; out[2 * (b - lo) + 1] = c; }
        if (n_diff) nd += (s != sinf(x)) || (c != cosf(x));
    }
    if (n_diff) *n_diff = nd;
    return ORBF_OK;
}

// the device restatement over the same range: out = sin then cos of each input (device-computed, copied back)
__global__ void selftest_sincosf_kernel(uint32_t lo, uint32_t count, float* out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const float x = __uint_as_float(lo + i);
    out[2 * (size_t)i] = replay::glibc_sinf(x); out[2 * (size_t)i + 1] = replay::glibc_cosf(x);
}
extern "C" int orbf_selftest_sincosf_device(orbf_context* c, uint32_t lo, uint32_t hi, float* out)
{
    CTX_ENTER(c);
    if (hi < lo || !out || (uint64_t)hi - lo >= (1u << 28)) return ORBF_ERR_ARG;
    const uint32_t count = hi - lo + 1;
    float* d = nullptr;
    ORBF_CUDA(c, cudaMalloc((void**)&d, (size_t)count * 2 * sizeof(float)));
    selftest_sincosf_kernel<<<(count + 255) / 256, 256, 0, c->stream>>>(lo, count, d);
    cudaError_t e = cudaMemcpyAsync(out, d, (size_t)count * 2 * sizeof(float), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d);
    if (e != cudaSuccess) return orbf_cuda_fail(c, e, "selftest_sincosf", __FILE__, __LINE__);
    return ORBF_OK;
}

extern "C" int orbf_selftest_sample_table(uint32_t seed, int32_t M, int32_t iterations, int32_t sample_size, int32_t* table)
{
    if (!table || sample_size < 1 || sample_size > 8) return ORBF_ERR_ARG;
    replay::GlibcRand g;
    g.seed(seed);
    for (int k = 0; k < iterations; ++k) {
        if (M >= sample_size) replay::sample_row(g, M, sample_size, table + (size_t)k * sample_size);
        else for (int s = 0; s < sample_size; ++s) table[(size_t)k * sample_size + s] = -1;
    }
    return ORBF_OK;
}

