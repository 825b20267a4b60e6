// Probe for the tcgen05 path the matcher uses: D[128 x N] (s32, TMEM) = A[128 x K] * B[N x K]^T with s8 operands staged in shared
// memory by ordinary stores in the no-swizzle K-major canonical layout (8-row x 16-byte core matrices).  Checks the descriptor
// fields against a CPU product.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o umma_probe tools/umma_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

constexpr int M = 128, N = 256, K = 256, KCH = K / 16;          // KCH 16-byte chunks per row
constexpr uint32_t SBO = KCH * 128, LBO = 128;                  // bytes: next 8-row group / next 16-byte K chunk

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr)
{
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((LBO >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((SBO >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                                      // descriptor version (sm_100)
    return d;                                                    // layout type 0 = no swizzle
}

__global__ void __launch_bounds__(128) umma_kernel(const int8_t* __restrict__ A, const int8_t* __restrict__ B, int32_t* __restrict__ D, int* status)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA = smem;                         // 128 rows
    uint8_t* sB = smem + M * K;                 // 256 rows
    __shared__ uint32_t sTmem;
    __shared__ __align__(8) uint64_t sBar;
    const int tid = threadIdx.x, warp = tid >> 5;
    // stage: (row, chunk) -> (row >> 3) * SBO + chunk * 128 + (row & 7) * 16
    for (int i = tid; i < M * KCH; i += 128) {
        const int row = i % M, c = i / M;
        *reinterpret_cast<uint4*>(sA + (row >> 3) * SBO + c * LBO + (row & 7) * 16) = *reinterpret_cast<const uint4*>(A + row * K + c * 16);
    }
    for (int i = tid; i < N * KCH; i += 128) {
        const int row = i % N, c = i / N;
        *reinterpret_cast<uint4*>(sB + (row >> 3) * SBO + c * LBO + (row & 7) * 16) = *reinterpret_cast<const uint4*>(B + row * K + c * 16);
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sTmem)), "n"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&sBar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    asm volatile("fence.proxy.async.shared::cta;");             // generic-proxy stores -> visible to the tensor core's async proxy
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = sTmem;
    if (tid == 0) {
        // instruction descriptor: c_format S32 (2) @4, a/b format INT8 (1) @7/@10, K-major both, N >> 3 @17, M >> 4 @24
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        for (int ks = 0; ks < K / 32; ++ks) {
            const uint64_t da = make_desc(smem_u32(sA) + ks * 2 * LBO), db = make_desc(smem_u32(sB) + ks * 2 * LBO);
            const uint32_t acc = ks > 0;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc));
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&sBar)));
    }
    // wait (bounded: a wrong descriptor must not hang the box)
    {
        uint32_t done = 0;
        for (int spin = 0; spin < 2000000 && !done; ++spin)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&sBar)), "r"(0u));
        if (!done) { if (tid == 0) *status = 1; }
    }
    asm volatile("tcgen05.fence::after_thread_sync;");
    // each warp reads its 32 lanes, 32 columns at a time
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]),
                       "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]),
                       "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;");
#pragma unroll
        for (int j = 0; j < 32; ++j) D[(size_t)tid * N + c0 + j] = (int32_t)v[j];
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(256));
}

int main()
{
    std::vector<int8_t> hA(M * K), hB(N * K);
    srand(3);
    for (auto& v : hA) v = (rand() & 1) ? 1 : -1;
    for (auto& v : hB) v = (rand() & 1) ? 64 : -64;
    int8_t *dA, *dB; int32_t* dD; int* dS;
    cudaMalloc(&dA, hA.size()); cudaMalloc(&dB, hB.size()); cudaMalloc(&dD, M * N * 4); cudaMalloc(&dS, 4);
    cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice); cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xCD, M * N * 4); cudaMemset(dS, 0, 4);
    const int smemBytes = (M + N) * K + 1024;
    cudaFuncSetAttribute(umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smemBytes);
    umma_kernel<<<1, 128, smemBytes>>>(dA, dB, dD, dS);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<int32_t> hD(M * N); int st = 0;
    cudaMemcpy(hD.data(), dD, M * N * 4, cudaMemcpyDeviceToHost); cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost);
    printf("status (1 = barrier timed out): %d\n", st);
    int bad = 0;
    for (int i = 0; i < M; ++i) for (int j = 0; j < N; ++j) {
        int ref = 0;
        for (int k = 0; k < K; ++k) ref += (int)hA[i * K + k] * (int)hB[j * K + k];
        if (ref != hD[i * N + j]) { if (bad < 8) printf("  D[%d][%d] = %d, expected %d\n", i, j, hD[i * N + j], ref); ++bad; }
    }
    printf("mismatches: %d of %d\n", bad, M * N);
    return bad != 0;
}
