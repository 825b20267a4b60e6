// csrc/comm.cu — the one exchange step of the path (BASELINE config 5): keyframe descriptors live sharded, kfCap
// keyframes per GPU ([kfCap][K][32] u8 + [kfCap] i32, csrc/kfdb.cu), and a query is matched against the keyframes of EVERY rank.
// Two ways to bring the shards to the matcher, both one process per GPU:
//   * orbf_kfdb_allgather  — ncclAllGather of the stores (uint8, kfCap * K * 32 bytes per rank) over NVLink / NVSwitch into a
//     gathered copy, which the matcher then reads locally.  NCCL is loaded with dlopen("libnccl.so.2") on first use (the copy the
//     host process already holds, e.g. torch's, else the system one): the library has no link-time NCCL dependency.  The host
//     application distributes the 128-byte ncclUniqueId (orbf_comm_unique_id on rank 0 -> any broadcast it likes -> orbf_comm_init).
//   * orbf_kfdb_attach_peers — no collective and no gathered copy: every rank exports its store as CUDA IPC handles, opens the
//     others', and the matcher kernel reads each keyframe's rows from the GPU that owns it with plain loads over NVLink while it
//     multiplies (the train tile of keyframe j streams in under the MMAs of keyframe j - 1's CTAs); the same bytes cross the switch
//     as in the all-gather, but nothing waits for a collective to finish and 8x less HBM is written.
// The reference has no counterpart (its Database::Query is DBoW3 inverted-file scoring, quirk Q13); per keyframe the result is
// Matcher::KnnMatch's kNN-2 + ratio (Features/matcher.cpp:23-35), as in kfdb.cu.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>

#include "orbf_internal.h"

#define CTX_ENTER(c)                                                                   \
    do {                                                                               \
        if (!(c)) return ORBF_ERR_ARG;                                                 \
        cudaError_t e_ = cudaSetDevice((c)->cfg.device);                               \
        if (e_ != cudaSuccess) return orbf_cuda_fail((c), e_, "cudaSetDevice", __FILE__, __LINE__); \
    } while (0)

namespace {
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    int (*GetVersion)(int*) = nullptr;
};
NcclApi g_nccl;

bool nccl_load(std::string& err)
{
    if (g_nccl.lib) return true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);          // the copy this process already holds (e.g. torch's), if any
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW);
    if (!h) { err = std::string("dlopen libnccl.so.2: ") + dlerror(); return false; }
#define SYM(field, name) do { *(void**)&g_nccl.field = dlsym(h, name); if (!g_nccl.field) { err = std::string("dlsym ") + name; dlclose(h); return false; } } while (0)
    SYM(GetUniqueId, "ncclGetUniqueId"); SYM(CommInitRank, "ncclCommInitRank"); SYM(CommDestroy, "ncclCommDestroy");
    SYM(AllGather, "ncclAllGather"); SYM(GroupStart, "ncclGroupStart"); SYM(GroupEnd, "ncclGroupEnd"); SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    g_nccl.lib = h;
    return true;
}

int nccl_fail(orbf_context* c, ncclResult_t r, const char* what)
{
    if (c) c->lastError = std::string(what) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "NCCL error");
    return ORBF_ERR_CUDA;
}
#define ORBF_NCCL(c, call) do { ncclResult_t r__ = (call); if (r__ != ncclSuccess) return nccl_fail((c), r__, #call); } while (0)
}  // namespace

static_assert(sizeof(ncclUniqueId) == 128, "orbf_comm_unique_id hands out 128 bytes");

extern "C" int orbf_comm_unique_id(uint8_t* id128)
{
    if (!id128) return ORBF_ERR_ARG;
    std::string err;
    if (!nccl_load(err)) return ORBF_ERR_CUDA;
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != ncclSuccess) return ORBF_ERR_CUDA;
    std::memcpy(id128, &id, sizeof(id));
    return ORBF_OK;
}

extern "C" int orbf_comm_init(orbf_context* c, const uint8_t* id128, int32_t nranks, int32_t rank)
{
    CTX_ENTER(c);
    if (!id128 || nranks < 1 || rank < 0 || rank >= nranks || nranks > 64) return ORBF_ERR_ARG;
    std::string err;
    if (!nccl_load(err)) { c->lastError = err; return ORBF_ERR_CUDA; }
    if (c->ncclComm) { g_nccl.CommDestroy((ncclComm_t)c->ncclComm); c->ncclComm = nullptr; }
    ncclUniqueId id;
    std::memcpy(&id, id128, sizeof(id));
    ncclComm_t comm = nullptr;
    ORBF_NCCL(c, g_nccl.CommInitRank(&comm, nranks, id, rank));
    c->ncclComm = comm; c->commRanks = nranks; c->commRank = rank;
    return ORBF_OK;
}

// ncclAllGather of every rank's keyframe store (all ranks reserve the same capacity) into this context's gathered copy, which becomes
// the store orbf_kfdb_match reads: keyframe g = rank * kfCap + local index.  Asynchronous on the context stream.
extern "C" int orbf_kfdb_allgather(orbf_context* c, int32_t* n_kf_total)
{
    CTX_ENTER(c);
    if (!c->ncclComm || c->kfCap < 1) return ORBF_ERR_STATE;
    const int total = c->commRanks * c->kfCap;
    if (total > c->kfGatherCap) {
        ORBF_CUDA(c, cudaStreamSynchronize(c->stream));
        if (c->d_kfGather) cudaFree(c->d_kfGather);
        if (c->d_kfGatherCount) cudaFree(c->d_kfGatherCount);
        c->d_kfGather = nullptr; c->d_kfGatherCount = nullptr; c->kfGatherCap = 0;
        ORBF_CUDA(c, cudaMalloc((void**)&c->d_kfGather, (size_t)total * c->K * 32));
        ORBF_CUDA(c, cudaMalloc((void**)&c->d_kfGatherCount, (size_t)total * sizeof(int)));
        c->kfGatherCap = total;
    }
    ORBF_NCCL(c, g_nccl.GroupStart());
    ORBF_NCCL(c, g_nccl.AllGather(c->d_kfDesc, c->d_kfGather, (size_t)c->kfCap * c->K * 32, ncclUint8, (ncclComm_t)c->ncclComm, c->stream));
    ORBF_NCCL(c, g_nccl.AllGather(c->d_kfCount, c->d_kfGatherCount, (size_t)c->kfCap, ncclInt32, (ncclComm_t)c->ncclComm, c->stream));
    ORBF_NCCL(c, g_nccl.GroupEnd());
    c->launches += 2;
    c->d_kfExtDesc = c->d_kfGather; c->d_kfExtCount = c->d_kfGatherCount; c->kfExtN = total;
    c->nPeers = 0;                      // the gathered copy replaces a peer attachment
    if (n_kf_total) *n_kf_total = total;
    return ORBF_OK;
}

// ---- gather-free path: peer stores through CUDA IPC -----------------------------------------------------------------------------
static_assert(sizeof(cudaIpcMemHandle_t) == 64, "orbf_kfdb_ipc_handles hands out 64 bytes per buffer");

extern "C" int orbf_kfdb_ipc_handles(orbf_context* c, uint8_t* desc_handle64, uint8_t* count_handle64)
{
    CTX_ENTER(c);
    if (!desc_handle64 || !count_handle64 || c->kfCap < 1) return ORBF_ERR_ARG;
    cudaIpcMemHandle_t hd, hc;
    ORBF_CUDA(c, cudaIpcGetMemHandle(&hd, c->d_kfDesc));
    ORBF_CUDA(c, cudaIpcGetMemHandle(&hc, c->d_kfCount));
    std::memcpy(desc_handle64, &hd, 64); std::memcpy(count_handle64, &hc, 64);
    return ORBF_OK;
}

static void close_peers(orbf_context* c)
{
    for (void*& p : c->peerOpened) if (p) { cudaIpcCloseMemHandle(p); p = nullptr; }
    if (c->d_peerDesc) cudaFree((void*)c->d_peerDesc);
    if (c->d_peerCount) cudaFree((void*)c->d_peerCount);
    c->d_peerDesc = nullptr; c->d_peerCount = nullptr; c->nPeers = 0; c->peerKf = 0;
}

// handles: every rank's pair from orbf_kfdb_ipc_handles, in rank order ([nranks][64] each); this rank's own entries are ignored (its
// store is used directly).  Afterwards orbf_kfdb_match addresses nranks * kf_per_rank keyframes, each read from its owner's memory.
extern "C" int orbf_kfdb_attach_peers(orbf_context* c, const uint8_t* desc_handles, const uint8_t* count_handles, int32_t nranks, int32_t rank,
    int32_t kf_per_rank)
{
    CTX_ENTER(c);
    if (nranks < 1 || nranks > 64 || rank < 0 || rank >= nranks || kf_per_rank < 1 || kf_per_rank > c->kfCap || (nranks > 1 && (!desc_handles || !count_handles)))
        return ORBF_ERR_ARG;
    ORBF_CUDA(c, cudaStreamSynchronize(c->stream));
    close_peers(c);
    const uint8_t* hd[64]; const int* hc[64];
    for (int r = 0; r < nranks; ++r) {
        if (r == rank) { hd[r] = c->d_kfDesc; hc[r] = c->d_kfCount; continue; }
        cudaIpcMemHandle_t a, b;
        std::memcpy(&a, desc_handles + 64 * (size_t)r, 64); std::memcpy(&b, count_handles + 64 * (size_t)r, 64);
        void *pd = nullptr, *pc = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&pd, a, cudaIpcMemLazyEnablePeerAccess);
        if (e == cudaSuccess) e = cudaIpcOpenMemHandle(&pc, b, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) { if (pd) cudaIpcCloseMemHandle(pd); close_peers(c); return orbf_cuda_fail(c, e, "cudaIpcOpenMemHandle", __FILE__, __LINE__); }
        c->peerOpened[2 * r] = pd; c->peerOpened[2 * r + 1] = pc;
        hd[r] = (const uint8_t*)pd; hc[r] = (const int*)pc;
    }
    ORBF_CUDA(c, cudaMalloc((void**)&c->d_peerDesc, nranks * sizeof(void*)));
    ORBF_CUDA(c, cudaMalloc((void**)&c->d_peerCount, nranks * sizeof(void*)));
    ORBF_CUDA(c, cudaMemcpy((void*)c->d_peerDesc, hd, nranks * sizeof(void*), cudaMemcpyHostToDevice));
    ORBF_CUDA(c, cudaMemcpy((void*)c->d_peerCount, hc, nranks * sizeof(void*), cudaMemcpyHostToDevice));
    c->nPeers = nranks; c->peerKf = kf_per_rank;
    c->d_kfExtDesc = nullptr; c->d_kfExtCount = nullptr; c->kfExtN = 0;
    return ORBF_OK;
}

extern "C" int orbf_kfdb_detach_peers(orbf_context* c)
{
    CTX_ENTER(c);
    ORBF_CUDA(c, cudaStreamSynchronize(c->stream));
    close_peers(c);
    return ORBF_OK;
}

extern "C" int orbf_comm_destroy(orbf_context* c)
{
    CTX_ENTER(c);
    if (c->ncclComm && g_nccl.CommDestroy) { ORBF_CUDA(c, cudaStreamSynchronize(c->stream)); g_nccl.CommDestroy((ncclComm_t)c->ncclComm); }
    c->ncclComm = nullptr; c->commRanks = 1; c->commRank = 0;
    return ORBF_OK;
}

void orbf_comm_release(orbf_context* c)
{
    close_peers(c);
    if (c->ncclComm && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)c->ncclComm);
    c->ncclComm = nullptr;
    if (c->d_kfGather) cudaFree(c->d_kfGather);
    if (c->d_kfGatherCount) cudaFree(c->d_kfGatherCount);
    c->d_kfGather = nullptr; c->d_kfGatherCount = nullptr; c->kfGatherCap = 0;
}
