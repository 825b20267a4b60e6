// tools/pipe_bench.cu — integer-pipe throughput probe for the matching / FAST rooflines ("the POPC
// peak is not in MEASURED_PEAKS.json: measure it once with a POPC/LOP3 micro-kernel and store it next to the HBM figure").
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pipe_bench tools/pipe_bench.cu && tools/pipe_bench > profiles/int_pipe_peaks.json
// Every test runs ILP independent dependency chains per thread, 8 warps x 4 CTAs per SM on all SMs, and reports
// thread-ops per clock per SM (from clock64 deltas of the slowest CTA) and chip-wide ops/s (from CUDA events).
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <vector>

constexpr int ILP = 8, ITERS = 4096, THREADS = 256, CTAS_PER_SM = 4;

enum Op { OP_POPC, OP_LOP3, OP_IADD3, OP_IMNMX, OP_IMNMX3, OP_VMNMX16X2, OP_VMNMX3_16X2, OP_HMNMX2, OP_DP4A, OP_IMAD, OP_FMNMX,
    OP_POPC_LOP3_MIX, OP_CSA_POPC_MIX, OP_COUNT };
const char* kNames[OP_COUNT] = { "popc_u32", "lop3_xor", "iadd3", "imnmx_s32", "imnmx3_s32", "vimnmx_u16x2", "vimnmx3_u16x2", "hmnmx2",
    "idp4a", "imad", "fmnmx", "mix_8xor_8popc_per_pair", "mix_csa_4popc_per_pair" };

template <int OP>
__device__ __forceinline__ uint32_t step(uint32_t x, uint32_t a, uint32_t b)
{
    if (OP == OP_POPC) return __popc(x) + a;      // popc + dependent add (add is on another pipe; popc is the slow one)
    if (OP == OP_LOP3) return (x ^ a) | (x & b);
    if (OP == OP_IADD3) return x + a + b;
    if (OP == OP_IMNMX) return (uint32_t)max((int)x, (int)a) ^ 0;
    if (OP == OP_IMNMX3) return (uint32_t)__vimax3_s32((int)x, (int)a, (int)b);
    if (OP == OP_VMNMX16X2) return __vmaxu2(x, a);
    if (OP == OP_VMNMX3_16X2) return __vimax3_u16x2(x, a, b);
    if (OP == OP_HMNMX2) { __half2 h = *reinterpret_cast<__half2*>(&x), g = *reinterpret_cast<__half2*>(&a); h = __hmax2(h, g); return *reinterpret_cast<uint32_t*>(&h); }
    if (OP == OP_DP4A) return __dp4a(x, a, b);
    if (OP == OP_IMAD) return x * a + b;
    if (OP == OP_FMNMX) { float f = fmaxf(__uint_as_float(x), __uint_as_float(a)); return __float_as_uint(f); }
    return x;
}

template <int OP>
__global__ void __launch_bounds__(THREADS) bench_kernel(uint32_t* out, long long* cyc, uint32_t seed)
{
    uint32_t v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = seed * (threadIdx.x + 1) + i * 0x9E3779B9u;
    uint32_t a = seed ^ 0x55AA55AAu, b = seed + 12345u + threadIdx.x;
    __syncthreads();
    const long long t0 = clock64();
    if (OP == OP_POPC_LOP3_MIX) {
        // the matcher's inner loop as written today: 8 XOR + 8 POPC + adds per descriptor pair (v[] = query words)
        uint32_t acc = 0;
        for (int it = 0; it < ITERS; ++it) {
            uint32_t d = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) d += __popc(v[i] ^ (a + i * b));
            acc = min(acc ^ d, d + it);
            a += d | 1;
        }
        v[0] = acc;
    } else if (OP == OP_CSA_POPC_MIX) {
        // carry-save variant: 8 XOR + 4 CSA (8 LOP3) + 4 POPC per pair
        uint32_t acc = 0;
        for (int it = 0; it < ITERS; ++it) {
            uint32_t x[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = v[i] ^ (a + i * b);
            const uint32_t s1 = x[0] ^ x[1] ^ x[2], c1 = (x[0] & x[1]) | (x[2] & (x[0] | x[1]));
            const uint32_t s2 = x[3] ^ x[4] ^ x[5], c2 = (x[3] & x[4]) | (x[5] & (x[3] | x[4]));
            const uint32_t s3 = s1 ^ s2 ^ x[6], c3 = (s1 & s2) | (x[6] & (s1 | s2));
            const uint32_t s4 = c1 ^ c2 ^ c3, c4 = (c1 & c2) | (c3 & (c1 | c2));
            const uint32_t d = __popc(s3) + __popc(x[7]) + 2 * __popc(s4) + 4 * __popc(c4);
            acc = min(acc ^ d, d + it);
            a += d | 1;
        }
        v[0] = acc;
    } else {
        // min/max are idempotent against a fixed operand (the compiler would fold the loop), so the second and third
        // operands are the neighbouring chains: still ILP independent results per round, latency-covered by the other warps
        for (int it = 0; it < ITERS; ++it) {
            uint32_t w[ILP];
#pragma unroll
            for (int i = 0; i < ILP; ++i) w[i] = step<OP>(v[i], v[(i + 1) % ILP], v[(i + 3) % ILP]);
#pragma unroll
            for (int i = 0; i < ILP; ++i) v[i] = w[i];
        }
    }
    const long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s ^= v[i];
    out[blockIdx.x * THREADS + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(int nsm, double clockHz, uint32_t* dOut, long long* dCyc, std::vector<long long>& hCyc, bool last)
{
    const int grid = nsm * CTAS_PER_SM;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    bench_kernel<OP><<<grid, THREADS>>>(dOut, dCyc, 1u);      // warm-up
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    bench_kernel<OP><<<grid, THREADS>>>(dOut, dCyc, 7u);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpy(hCyc.data(), dCyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < grid; ++i) mx = hCyc[i] > mx ? hCyc[i] : mx;
    const bool mix = OP == OP_POPC_LOP3_MIX || OP == OP_CSA_POPC_MIX;
    const double opsPerThread = mix ? (double)ITERS : (double)ITERS * ILP;     // mixes count descriptor pairs
    const double perClkSM = opsPerThread * THREADS * CTAS_PER_SM / (double)mx;
    const double perSec = opsPerThread * THREADS * grid / (ms * 1e-3);
    printf("  \"%s\": {\"per_clk_per_sm\": %.2f, \"chip_per_s\": %.4e, \"ms\": %.4f, \"unit\": \"%s\"}%s\n", kNames[OP], perClkSM, perSec, ms,
        mix ? "descriptor pairs (256-bit)" : "thread-ops", last ? "" : ",");
}

int main()
{
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, 0) != cudaSuccess) { fprintf(stderr, "no CUDA device\n"); return 1; }
    const int nsm = p.multiProcessorCount;
    uint32_t* dOut; long long* dCyc;
    cudaMalloc(&dOut, (size_t)nsm * CTAS_PER_SM * THREADS * sizeof(uint32_t));
    cudaMalloc(&dCyc, (size_t)nsm * CTAS_PER_SM * sizeof(long long));
    std::vector<long long> hCyc(nsm * CTAS_PER_SM);
    int clkKHz = 0;
    cudaDeviceGetAttribute(&clkKHz, cudaDevAttrClockRate, 0);
    printf("{\n  \"device\": \"%s\", \"sms\": %d, \"sm_clock_max_khz\": %d,\n", p.name, nsm, clkKHz);
    printf("  \"how\": \"tools/pipe_bench.cu: %d independent chains/thread, %d threads x %d CTAs per SM, %d iterations; per_clk_per_sm from clock64 of the slowest CTA, chip_per_s from CUDA events\",\n",
        ILP, THREADS, CTAS_PER_SM, ITERS);
    const double hz = clkKHz * 1e3;
    run<OP_POPC>(nsm, hz, dOut, dCyc, hCyc, false);
    run<OP_LOP3>(nsm, hz, dOut, dCyc, hCyc, false);
    run<OP_IADD3>(nsm, hz, dOut, dCyc, hCyc, false);
    run<OP_IMNMX>(nsm, hz, dOut, dCyc, hCyc, false);
    run<OP_IMNMX3>(nsm, hz, dOut, dCyc, hCyc, false);
    run<OP_VMNMX16X2>(nsm, hz, dOut, dCyc, hCyc, false);
    run<OP_VMNMX3_16X2>(nsm, hz, dOut, dCyc, hCyc, false);
    run<OP_HMNMX2>(nsm, hz, dOut, dCyc, hCyc, false);
    run<OP_DP4A>(nsm, hz, dOut, dCyc, hCyc, false);
    run<OP_IMAD>(nsm, hz, dOut, dCyc, hCyc, false);
    run<OP_FMNMX>(nsm, hz, dOut, dCyc, hCyc, false);
    run<OP_POPC_LOP3_MIX>(nsm, hz, dOut, dCyc, hCyc, false);
    run<OP_CSA_POPC_MIX>(nsm, hz, dOut, dCyc, hCyc, true);
    printf("}\n");
    return 0;
}
