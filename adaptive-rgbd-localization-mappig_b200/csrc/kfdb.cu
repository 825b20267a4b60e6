// csrc/kfdb.cu — device-resident keyframe descriptor store and many-to-many brute-force matching
// (BASELINE config 5).  The reference keeps keyframe descriptors in KeyFrame::mDescriptors (copied from the
// Frame, Core/keyframe.cpp:30-40) and indexes keyframes in Core/keyframedatabase.*; its Query is DBoW3
// inverted-file scoring, so brute-force matching against every keyframe has no literal counterpart
// (quirk Q13): per keyframe it is defined as Matcher::KnnMatch's kNN-2 + ratio (matcher.cpp:23-35).
// Layout: desc [kfCap][K][32] u8 + counts [kfCap]; one contiguous buffer so an NCCL all-gather of a rank's
// shard is a single call on orbf_kfdb_device_buffers().
#include <algorithm>
#include <vector>

#include "orbf_internal.h"

#define CTX_ENTER(c)                                                                   \
    do {                                                                               \
        if (!(c)) return ORBF_ERR_ARG;                                                 \
        cudaError_t e_ = cudaSetDevice((c)->cfg.device);                               \
        if (e_ != cudaSuccess) return orbf_cuda_fail((c), e_, "cudaSetDevice", __FILE__, __LINE__); \
    } while (0)
#define TRY(x) do { int r__ = (x); if (r__ != ORBF_OK) return r__; } while (0)

extern "C" int orbf_kfdb_reserve(orbf_context* c, int32_t maxKf)
{
    CTX_ENTER(c);
    if (maxKf < 1) return ORBF_ERR_ARG;
    if (maxKf <= c->kfCap) return ORBF_OK;
    uint8_t* nd = nullptr; int* nc = nullptr;
    ORBF_CUDA(c, cudaMalloc((void**)&nd, (size_t)maxKf * c->K * 32));
    ORBF_CUDA(c, cudaMalloc((void**)&nc, (size_t)maxKf * sizeof(int)));
    ORBF_CUDA(c, cudaMemsetAsync(nc, 0, (size_t)maxKf * sizeof(int), c->stream));
    if (c->kfCap > 0) {
        ORBF_CUDA(c, cudaMemcpyAsync(nd, c->d_kfDesc, (size_t)c->kfCap * c->K * 32, cudaMemcpyDeviceToDevice, c->stream));
        ORBF_CUDA(c, cudaMemcpyAsync(nc, c->d_kfCount, (size_t)c->kfCap * sizeof(int), cudaMemcpyDeviceToDevice, c->stream));
    }
    ORBF_CUDA(c, cudaStreamSynchronize(c->stream));
    if (c->d_kfDesc) cudaFree(c->d_kfDesc);
    if (c->d_kfCount) cudaFree(c->d_kfCount);
    c->d_kfDesc = nd; c->d_kfCount = nc; c->kfCap = maxKf;
    return ORBF_OK;
}

extern "C" int orbf_kfdb_add_from_slot(orbf_context* c, int32_t kf, int32_t slot)
{
    CTX_ENTER(c);
    if (kf < 0 || kf >= c->kfCap || slot < 0 || slot >= c->B) return ORBF_ERR_ARG;
    ORBF_CUDA(c, cudaMemcpyAsync(c->d_kfDesc + (size_t)kf * c->K * 32, c->d_desc + (size_t)slot * c->K * 32, (size_t)c->K * 32,
        cudaMemcpyDeviceToDevice, c->stream));
    ORBF_CUDA(c, cudaMemcpyAsync(c->d_kfCount + kf, c->d_count + slot, sizeof(int), cudaMemcpyDeviceToDevice, c->stream));
    return ORBF_OK;
}

extern "C" int orbf_kfdb_add_host(orbf_context* c, int32_t kf, const uint8_t* desc, int32_t n)
{
    CTX_ENTER(c);
    if (kf < 0 || kf >= c->kfCap || n < 0 || n > c->K || (n > 0 && !desc)) return ORBF_ERR_ARG;
    if (n) ORBF_CUDA(c, cudaMemcpyAsync(c->d_kfDesc + (size_t)kf * c->K * 32, desc, (size_t)n * 32, cudaMemcpyHostToDevice, c->stream));
    ORBF_CUDA(c, cudaMemcpyAsync(c->d_kfCount + kf, &n, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    ORBF_CUDA(c, cudaStreamSynchronize(c->stream));
    return ORBF_OK;
}

// Match against a caller-owned store instead (e.g. the NCCL all-gather of every rank's shard): same [n_kf][K][32] + [n_kf]
// layout.  The context does not take ownership; attach NULL to return to its own store.
extern "C" int orbf_kfdb_attach_device(orbf_context* c, const uint8_t* d_desc, const int32_t* d_counts, int32_t n_kf)
{
    CTX_ENTER(c);
    if ((d_desc == nullptr) != (d_counts == nullptr) || (d_desc && n_kf < 1)) return ORBF_ERR_ARG;
    if (d_desc && ((uintptr_t)d_desc & 15)) return ORBF_ERR_ALIGNMENT;
    c->d_kfExtDesc = d_desc; c->d_kfExtCount = d_counts; c->kfExtN = d_desc ? n_kf : 0;
    return ORBF_OK;
}

extern "C" int orbf_kfdb_device_buffers(orbf_context* c, uint8_t** d_desc, int32_t** d_counts, int32_t* rows_per_kf, int32_t* n_kf)
{
    if (!c) return ORBF_ERR_ARG;
    if (d_desc) *d_desc = c->d_kfDesc;
    if (d_counts) *d_counts = c->d_kfCount;
    if (rows_per_kf) *rows_per_kf = c->K;
    if (n_kf) *n_kf = c->kfCap;
    return ORBF_OK;
}

__global__ void kf_pairs_kernel(int* pairs, int kf0, int nkf, int qslot)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nkf) { pairs[2 * i] = qslot; pairs[2 * i + 1] = kf0 + i; }
}

static int kf_reserve_out(orbf_context* c, int nkf)
{
    if (nkf <= c->kfOutCap) return ORBF_OK;
    ORBF_CUDA(c, cudaStreamSynchronize(c->stream));
    if (c->d_kfKnn) cudaFree(c->d_kfKnn);
    if (c->d_kfSurv) cudaFree(c->d_kfSurv);
    if (c->d_kfPairs) cudaFree(c->d_kfPairs);
    if (c->d_kfQCount) cudaFree(c->d_kfQCount);
    c->d_kfKnn = nullptr; c->d_kfSurv = nullptr; c->d_kfPairs = nullptr; c->d_kfQCount = nullptr; c->kfOutCap = 0;
    ORBF_CUDA(c, cudaMalloc((void**)&c->d_kfKnn, (size_t)nkf * c->K * 2 * sizeof(uint32_t)));
    ORBF_CUDA(c, cudaMalloc((void**)&c->d_kfSurv, (size_t)nkf * sizeof(int)));
    ORBF_CUDA(c, cudaMalloc((void**)&c->d_kfPairs, (size_t)nkf * 2 * sizeof(int)));
    ORBF_CUDA(c, cudaMalloc((void**)&c->d_kfQCount, sizeof(int)));
    c->kfOutCap = nkf;
    return ORBF_OK;
}

// kNN-2 + ratio survivors of one query against keyframes [kf0, kf0 + nkf) of the current store (own / attached / all-gathered / peers)
static int kf_launch(orbf_context* c, const uint8_t* d_q, const int* d_qCount, int qslot, long long qStride, int nq, int kf0, int nkf, float ratio)
{
    const bool ext = c->d_kfExtDesc != nullptr, peers = c->nPeers > 0;
    kf_pairs_kernel<<<(nkf + 127) / 128, 128, 0, c->stream>>>(c->d_kfPairs, kf0, nkf, qslot);
    ORBF_LAUNCH_CHECK(c);
    MatchSet ms;
    ms.qdesc = d_q; ms.tdesc = ext ? c->d_kfExtDesc : c->d_kfDesc; ms.qStride = qStride; ms.tStride = (long long)c->K * 32;
    ms.qCounts = d_qCount; ms.tCounts = ext ? c->d_kfExtCount : c->d_kfCount; ms.pairs = c->d_kfPairs; ms.pair0 = 0; ms.nq = nq; ms.nt = 0;
    ms.knn = c->d_kfKnn; ms.rev = nullptr; ms.matches = nullptr; ms.matchCount = c->d_kfSurv;
    if (peers) { ms.tShards = c->d_peerDesc; ms.tShardCounts = c->d_peerCount; ms.shardKf = c->peerKf; }      // rows read from the owning GPU
    TRY(orbf_launch_knn2(c, ms, nkf, false));
    TRY(orbf_launch_match_select(c, ms, nkf, ratio, false));
    return ORBF_OK;
}

static int kf_available(const orbf_context* c)
{
    return c->nPeers > 0 ? c->nPeers * c->peerKf : (c->d_kfExtDesc ? c->kfExtN : c->kfCap);
}

// Device-resident query: the descriptors of frame slot `slot` against keyframes [kf0, kf0 + nkf); asynchronous, results stay on the
// device until orbf_kfdb_results.
extern "C" int orbf_kfdb_match_slot(orbf_context* c, int32_t slot, int32_t kf0, int32_t nkf, float ratio)
{
    CTX_ENTER(c);
    if (slot < 0 || slot >= c->B || kf0 < 0 || nkf < 1 || kf0 + nkf > kf_available(c)) return ORBF_ERR_ARG;
    TRY(kf_reserve_out(c, nkf));
    return kf_launch(c, c->d_desc, c->d_count, slot, (long long)c->K * 32, 0, kf0, nkf, ratio);
}

// Results of the last orbf_kfdb_match_slot / orbf_kfdb_match for its first nkf keyframes: per keyframe top-2 per query row (tables
// [nkf][nq], any may be NULL) and ratio survivors [nkf].  Synchronises the stream.
extern "C" int orbf_kfdb_results(orbf_context* c, int32_t nkf, int32_t nq, int32_t* idx1, int32_t* d1, int32_t* idx2, int32_t* d2, int32_t* survivors)
{
    CTX_ENTER(c);
    if (nkf < 1 || nkf > c->kfOutCap || nq < 0 || nq > c->K) return ORBF_ERR_ARG;
    const bool wantTables = idx1 || d1 || idx2 || d2;       // survivor counts alone (keyframe ranking) skip the 8 KB-per-keyframe tables
    std::vector<uint32_t> kk(wantTables ? (size_t)nkf * c->K * 2 : 0);
    if (wantTables) ORBF_CUDA(c, cudaMemcpyAsync(kk.data(), c->d_kfKnn, kk.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    if (survivors) ORBF_CUDA(c, cudaMemcpyAsync(survivors, c->d_kfSurv, (size_t)nkf * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    ORBF_CUDA(c, cudaStreamSynchronize(c->stream));
    if (!wantTables) return ORBF_OK;
    for (int k = 0; k < nkf; ++k)
        for (int i = 0; i < nq; ++i) {
            const uint32_t a = kk[((size_t)k * c->K + i) * 2], b = kk[((size_t)k * c->K + i) * 2 + 1];
            const size_t o = (size_t)k * nq + i;
            if (idx1) idx1[o] = (a == 0xFFFFFFFFu) ? -1 : (int)(a & 0xFFFF);
            if (d1) d1[o] = (a == 0xFFFFFFFFu) ? -1 : (int)(a >> 16);
            if (idx2) idx2[o] = (b == 0xFFFFFFFFu) ? -1 : (int)(b & 0xFFFF);
            if (d2) d2[o] = (b == 0xFFFFFFFFu) ? -1 : (int)(b >> 16);
        }
    return ORBF_OK;
}

extern "C" int orbf_kfdb_match(orbf_context* c, const uint8_t* q, int32_t nq, int32_t kf0, int32_t nkf, float ratio, int32_t* idx1,
    int32_t* d1, int32_t* idx2, int32_t* d2, int32_t* survivors)
{
    CTX_ENTER(c);
    if (!q || nq < 1 || nq > c->K || kf0 < 0 || nkf < 1 || kf0 + nkf > kf_available(c)) return ORBF_ERR_ARG;
    TRY(kf_reserve_out(c, nkf));
    if (nq > c->descStageRows) {
        if (c->d_qdesc) cudaFree(c->d_qdesc);
        if (c->d_tdesc) cudaFree(c->d_tdesc);
        c->d_qdesc = c->d_tdesc = nullptr; c->descStageRows = 0;
        ORBF_CUDA(c, cudaMalloc((void**)&c->d_qdesc, (size_t)nq * 32));
        ORBF_CUDA(c, cudaMalloc((void**)&c->d_tdesc, (size_t)nq * 32));
        c->descStageRows = nq;
    }
    ORBF_CUDA(c, cudaMemcpyAsync(c->d_qdesc, q, (size_t)nq * 32, cudaMemcpyHostToDevice, c->stream));
    ORBF_CUDA(c, cudaMemcpyAsync(c->d_kfQCount, &nq, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    ORBF_CUDA(c, cudaStreamSynchronize(c->stream));
    TRY(kf_launch(c, c->d_qdesc, c->d_kfQCount, 0, 0, nq, kf0, nkf, ratio));
    return orbf_kfdb_results(c, nkf, nq, idx1, d1, idx2, d2, survivors);
}
