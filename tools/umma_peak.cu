// tools/umma_peak.cu — measured dense int8 tensor peak of this B200 for the instruction the matcher uses: a bare loop of
// tcgen05.mma.cta_group::1.kind::i8 (M = 128, N = 256, K = 32 per instruction, s8 x s8 -> s32 in TMEM, operands in shared memory in
// the no-swizzle K-major canonical layout), one issuing thread per CTA, a commit + mbarrier wait every 64 instructions, no epilogue.
// Prints one JSON line (copied to profiles/int8_tensor_peak.json, the denominator of bench.py's roofline.hamming_knn2).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_peak tools/umma_peak.cu && ./umma_peak
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int M = 128, N = 256, KCH = 16;                         // 16 chunks of 16 bytes = K 256 per row
constexpr uint32_t LBO = 128, SBO = KCH * 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr)
{
    return (uint64_t)((addr >> 4) & 0x3FFFu) | (uint64_t)(LBO >> 4) << 16 | (uint64_t)(SBO >> 4) << 32 | (uint64_t)1 << 46;
}

__global__ void __launch_bounds__(128) peak_kernel(int passes, int* status)
{
    extern __shared__ __align__(1024) uint8_t smemRaw[];
    uint8_t* sA = smemRaw + ((1024u - (smem_u32(smemRaw) & 1023u)) & 1023u);
    uint8_t* sB = sA + M * KCH * 16;
    __shared__ uint32_t sTmem;
    __shared__ __align__(8) uint64_t sBar;
    const int tid = threadIdx.x;
    for (int i = tid; i < (M + N) * KCH * 4; i += 128) reinterpret_cast<uint32_t*>(sA)[i] = 0x01FF0201u * (uint32_t)(i % 7 + 1);
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sTmem)), "n"(N));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&sBar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = sTmem;
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    if (tid == 0) {
        uint32_t phase = 0;
        for (int p = 0; p < passes; p += 8) {
            for (int q = 0; q < 8; ++q)
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    const uint64_t da = make_desc(smem_u32(sA) + ks * 2 * LBO), db = make_desc(smem_u32(sB) + ks * 2 * LBO);
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                                 ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(1u) : "memory");
                }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&sBar)) : "memory");
            uint32_t done = 0, spins = 0;
            while (!done) {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(smem_u32(&sBar)), "r"(phase) : "memory");
                if (!done && ++spins > (1u << 24)) { *status = 1; __trap(); }
            }
            phase ^= 1u;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(N));
}

int main()
{
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    int* dStatus; cudaMalloc(&dStatus, 4); cudaMemset(dStatus, 0, 4);
    const size_t smem = (size_t)(M + N) * KCH * 16 + 1024;
    cudaFuncSetAttribute(peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int passes = 4096;                                       // x 8 instructions of 128 x 256 x 32
    double best[3] = {0, 0, 0};
    for (int perSm = 1; perSm <= 2; ++perSm) {
        const int grid = prop.multiProcessorCount * perSm;
        peak_kernel<<<grid, 128, smem>>>(64, dStatus);
        cudaDeviceSynchronize();
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            peak_kernel<<<grid, 128, smem>>>(passes, dStatus);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
            const double ops = 2.0 * M * N * 32 * 8.0 * passes * grid;
            best[perSm] = ops / (ms * 1e-3) > best[perSm] ? ops / (ms * 1e-3) : best[perSm];
        }
    }
    int st = 0; cudaMemcpy(&st, dStatus, 4, cudaMemcpyDeviceToHost);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const cudaError_t e = cudaGetLastError();
    printf("{\"what\": \"bare tcgen05.mma kind::i8 loop, M=128 N=256 K=32 per instruction, operands in shared memory (no swizzle, K-major), commit every 64 instructions\", "
           "\"sms\": %d, \"clock_khz\": %d, \"int8_ops_per_s_1cta_per_sm\": %.6e, \"int8_ops_per_s_2cta_per_sm\": %.6e, \"int8_ops_per_clk_per_sm\": %.1f, "
           "\"nominal_ops_per_clk_per_sm\": 16384, \"status\": %d, \"cuda\": \"%s\"}\n",
        prop.multiProcessorCount, clk, best[1], best[2], (best[1] > best[2] ? best[1] : best[2]) / prop.multiProcessorCount / (clk * 1e3), st, cudaGetErrorString(e));
    return st || e != cudaSuccess;
}
