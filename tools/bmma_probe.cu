// Probe: does the legacy 1-bit tensor-core path (mma.sync m16n8k256 b1 and.popc / xor.popc) exist on sm_100a, is it
// exact, and how fast is it?  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o bmma_probe tools/bmma_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ void bmma_and(int (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2])
{
    asm volatile("mma.sync.aligned.m16n8k256.row.col.s32.b1.b1.s32.and.popc {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
#ifdef PROBE_XOR
__device__ __forceinline__ void bmma_xor(int (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2])
{
    asm volatile("mma.sync.aligned.m16n8k256.row.col.s32.b1.b1.s32.xor.popc {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
#endif

__device__ __forceinline__ void imma_u8(int (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2])
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int ILP>
__global__ void imma_tput_kernel(int iters, int* sink)
{
    uint32_t a[4] = {threadIdx.x & 0x01010101u, 0x01000101u, 0x00010100u, 0x01010001u};
    uint32_t b[2] = {0x01010100u, (blockIdx.x & 1u) * 0x01010101u};
    int d[ILP][4];
#pragma unroll
    for (int k = 0; k < ILP; ++k) d[k][0] = d[k][1] = d[k][2] = d[k][3] = 0;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < ILP; ++k) imma_u8(d[k], a, b);
    }
    int s = 0;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += d[k][0] + d[k][1] + d[k][2] + d[k][3];
    if (s == 0x7fffffff) sink[0] = s;
}

__device__ __forceinline__ void qmma_e4m3(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2])
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.f32.e4m3.e4m3.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int ILP>
__global__ void qmma_tput_kernel(int iters, int* sink)
{
    uint32_t a[4] = {0x38383838u, 0x38003838u, 0x00383800u, 0x38380038u};     // 1.0 in e4m3 = 0x38
    uint32_t b[2] = {0x38383800u, (blockIdx.x & 1u) * 0x38383838u};
    float d[ILP][4];
#pragma unroll
    for (int k = 0; k < ILP; ++k) d[k][0] = d[k][1] = d[k][2] = d[k][3] = 0.f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < ILP; ++k) qmma_e4m3(d[k], a, b);
    }
    float s = 0;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += d[k][0] + d[k][1] + d[k][2] + d[k][3];
    if (s == 12345.f) sink[0] = (int)s;
}


// does IMMA issue overlap with integer ALU work of the same / other warps?  NALU independent LOP3-class ops per 8 IMMAs
template <int NALU>
__global__ void mix_kernel(int iters, int* sink)
{
    uint32_t a[4] = {threadIdx.x & 0x01010101u, 0x01000101u, 0x00010100u, 0x01010001u};
    uint32_t b[2] = {0x01010100u, (blockIdx.x & 1u) * 0x01010101u};
    int d[8][4];
#pragma unroll
    for (int k = 0; k < 8; ++k) d[k][0] = d[k][1] = d[k][2] = d[k][3] = 0;
    uint32_t x[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = threadIdx.x * 77u + k;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            imma_u8(d[k], a, b);
#pragma unroll
            for (int j = 0; j < NALU / 8; ++j) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[(k + j) & 7]) : "r"(x[(k + j + 3) & 7]), "r"(x[(k + j + 5) & 7]));
        }
    }
    int s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += d[k][0] + d[k][1] + d[k][2] + d[k][3] + (int)x[k];
    if (s == 0x7fffffff) sink[0] = s;
}

template <int NALU>
static void run_mix(int* dO, cudaDeviceProp& p, int clk, cudaEvent_t e0, cudaEvent_t e1)
{
    const int iters = 50000, warps = 16;
    int blocks = p.multiProcessorCount;
    mix_kernel<NALU><<<blocks, warps * 32>>>(100, dO); cudaDeviceSynchronize();
    cudaEventRecord(e0); mix_kernel<NALU><<<blocks, warps * 32>>>(iters, dO); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double clkPerIter = ms * 1e-3 * clk * 1e3 / iters;          // SM clocks per loop iteration (all 16 warps run one iteration each)
    printf("mix: 8 IMMA + %3d ALU per warp-iteration, 16 warps/SM: %.1f clk per iteration-round  (IMMA alone would be %.1f, ALU alone %.1f)\n", NALU, clkPerIter,
           16 * 8 / 0.48, 16.0 * NALU / 2.0);
}

// correctness: one warp, 16 descriptors x 8 descriptors
__global__ void check_kernel(const uint32_t* A, const uint32_t* B, int* out, int useXor)
{
    int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
    uint32_t a[4] = {A[g * 8 + t], A[(g + 8) * 8 + t], A[g * 8 + 4 + t], A[(g + 8) * 8 + 4 + t]};
    uint32_t b[2] = {B[g * 8 + t], B[g * 8 + 4 + t]};
    int d[4] = {0, 0, 0, 0};
#ifdef PROBE_XOR
    if (useXor) bmma_xor(d, a, b); else
#endif
    bmma_and(d, a, b);
    out[g * 8 + t * 2] = d[0]; out[g * 8 + t * 2 + 1] = d[1];
    out[(g + 8) * 8 + t * 2] = d[2]; out[(g + 8) * 8 + t * 2 + 1] = d[3];
}

template <int ILP>
__global__ void tput_kernel(int iters, int* sink, int useXor)
{
    uint32_t a[4] = {threadIdx.x * 2654435761u, blockIdx.x * 40503u + 1, threadIdx.x + 77u, 0x9e3779b9u};
    uint32_t b[2] = {threadIdx.x * 97u + 13, blockIdx.x + 5u};
    int d[ILP][4];
#pragma unroll
    for (int k = 0; k < ILP; ++k) d[k][0] = d[k][1] = d[k][2] = d[k][3] = 0;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < ILP; ++k) {
#ifdef PROBE_XOR
            if (useXor) bmma_xor(d[k], a, b); else
#endif
            bmma_and(d[k], a, b);
        }
    }
    int s = 0;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += d[k][0] + d[k][1] + d[k][2] + d[k][3];
    if (s == 0x7fffffff) sink[0] = s;
}

int main()
{
    uint32_t hA[16 * 8], hB[8 * 8]; int hO[128];
    srand(7);
    for (auto& v : hA) v = (uint32_t)rand() * 2654435761u ^ (uint32_t)rand();
    for (auto& v : hB) v = (uint32_t)rand() * 2246822519u ^ (uint32_t)rand();
    uint32_t *dA, *dB; int* dO;
    cudaMalloc(&dA, sizeof hA); cudaMalloc(&dB, sizeof hB); cudaMalloc(&dO, sizeof hO);
    cudaMemcpy(dA, hA, sizeof hA, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, sizeof hB, cudaMemcpyHostToDevice);
    for (int x = 0; x < 2; ++x) {
#ifndef PROBE_XOR
        if (x) break;
#endif
        check_kernel<<<1, 32>>>(dA, dB, dO, x);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("check(%s): %s\n", x ? "xor" : "and", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(hO, dO, sizeof hO, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int i = 0; i < 16; ++i) for (int j = 0; j < 8; ++j) {
            int ref = 0;
            for (int w = 0; w < 8; ++w) ref += __builtin_popcount(x ? (hA[i * 8 + w] ^ hB[j * 8 + w]) : (hA[i * 8 + w] & hB[j * 8 + w]));
            if (ref != hO[i * 8 + j]) ++bad;
        }
        printf("check %s.popc: %d mismatches of 128\n", x ? "xor" : "and", bad);
    }
    int dev = 0; cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int warps = 4; warps <= 16; warps *= 2) {
        const int iters = 20000, ILP = 8;
        int blocks = p.multiProcessorCount * 2;
        tput_kernel<ILP><<<blocks, warps * 32>>>(100, dO, 0); cudaDeviceSynchronize();
        cudaEventRecord(e0); tput_kernel<ILP><<<blocks, warps * 32>>>(iters, dO, 0); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double n = (double)blocks * warps * iters * ILP;                     // warp-level MMAs
        printf("warps/CTA %2d: %.1f G bmma/s = %.2f bmma/clk/SM (nominal clock %d MHz) = %.1f T pair-distances/s\n", warps, n / ms * 1e-6,
               n / (ms * 1e-3) / p.multiProcessorCount / (clk * 1e3), clk / 1000, n * 128 / ms * 1e-9);
    }
    for (int warps = 4; warps <= 16; warps *= 2) {
        const int iters = 100000, ILP = 8;
        int blocks = p.multiProcessorCount * 2;
        imma_tput_kernel<ILP><<<blocks, warps * 32>>>(100, dO); cudaDeviceSynchronize();
        cudaEventRecord(e0); imma_tput_kernel<ILP><<<blocks, warps * 32>>>(iters, dO); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double n = (double)blocks * warps * iters * ILP;
        printf("IMMA.16832.U8 warps/CTA %2d: %.1f G imma/s = %.3f imma/clk/SM (nominal %d MHz) = %.1f dense int8 TOP/s; 256-bit pair-distances: %.1f T/s\n", warps,
               n / ms * 1e-6, n / (ms * 1e-3) / p.multiProcessorCount / (clk * 1e3), clk / 1000, n * 4096 * 2 / ms * 1e-9, n * 16 / ms * 1e-9);
    }
    for (int warps = 4; warps <= 16; warps *= 2) {
        const int iters = 100000, ILP = 8;
        int blocks = p.multiProcessorCount * 2;
        qmma_tput_kernel<ILP><<<blocks, warps * 32>>>(100, dO); cudaDeviceSynchronize();
        cudaEventRecord(e0); qmma_tput_kernel<ILP><<<blocks, warps * 32>>>(iters, dO); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double n = (double)blocks * warps * iters * ILP;
        printf("QMMA.16832.E4M3 (f32 acc) warps/CTA %2d: %.1f G mma/s = %.3f mma/clk/SM = %.1f dense fp8 TFLOP/s\n", warps,
               n / ms * 1e-6, n / (ms * 1e-3) / p.multiProcessorCount / (clk * 1e3), n * 4096 * 2 / ms * 1e-9);
    }
    run_mix<0>(dO, p, clk, e0, e1); run_mix<16>(dO, p, clk, e0, e1); run_mix<32>(dO, p, clk, e0, e1); run_mix<64>(dO, p, clk, e0, e1); run_mix<128>(dO, p, clk, e0, e1);
    return 0;
}
// (s4 probe, compile-only check) -----------------------------------------------------------------------------------------
__global__ void s4_probe_kernel(int* out)
{
    uint32_t a[4] = {threadIdx.x, 1, 2, 3}, b[2] = {5, threadIdx.x};
    int d[4] = {0, 0, 0, 0};
    asm volatile("mma.sync.aligned.m16n8k64.row.col.s32.s4.s4.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    out[threadIdx.x] = d[0] + d[1] + d[2] + d[3];
}
