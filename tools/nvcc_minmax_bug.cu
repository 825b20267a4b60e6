// micro-test: nvcc 12.9 / sm_100a miscompiles max(a, -max(...)) chains (VIMNMX3 fusion drops the negation?)
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#define N 4096
__device__ int v_ref_pattern(const int* d)   // original formulation (buggy codegen suspected)
{
    int best = -255;
    for (int s = 0; s < 16; ++s) {
        int l = 255, h = -255;
        for (int j = 0; j < 9; ++j) { const int e = d[(s + j) & 15]; l = min(l, e); h = max(h, e); }
        best = max(best, max(l, -h));
    }
    return best;
}
__device__ int v_two_arrays(const int* d)    // no unary minus on a min/max result
{
    int nd[16];
    for (int k = 0; k < 16; ++k) nd[k] = 0 - d[k];
    int best = -255;
    for (int s = 0; s < 16; ++s) {
        int l = 255, l2 = 255;
        for (int j = 0; j < 9; ++j) { l = min(l, d[(s + j) & 15]); l2 = min(l2, nd[(s + j) & 15]); }
        best = max(best, max(l, l2));
    }
    return best;
}
__device__ int v_pairtrick(const int* dd)    // production-style pair trick on d and nd
{
    int d[25], n[25];
    for (int k = 0; k < 16; ++k) { d[k] = dd[k]; n[k] = 0 - dd[k]; }
    for (int k = 16; k < 25; ++k) { d[k] = d[k - 16]; n[k] = n[k - 16]; }
    int best = -255;
#pragma unroll
    for (int k = 0; k < 16; k += 2) {
        int lo = min(d[k + 1], d[k + 2]), lo2 = min(n[k + 1], n[k + 2]);
#pragma unroll
        for (int j = 3; j <= 8; ++j) { lo = min(lo, d[k + j]); lo2 = min(lo2, n[k + j]); }
        best = max(best, max(min(lo, d[k]), min(lo, d[k + 9])));
        best = max(best, max(min(lo2, n[k]), min(lo2, n[k + 9])));
    }
    return best;
}
__device__ int v_pairtrick_neg(const int* dd)   // production formulation with -min(max...)
{
    int d[25];
    for (int k = 0; k < 16; ++k) d[k] = dd[k];
    for (int k = 16; k < 25; ++k) d[k] = d[k - 16];
    int best = -255;
#pragma unroll
    for (int k = 0; k < 16; k += 2) {
        int lo = min(d[k + 1], d[k + 2]), hi = max(d[k + 1], d[k + 2]);
#pragma unroll
        for (int j = 3; j <= 8; ++j) { lo = min(lo, d[k + j]); hi = max(hi, d[k + j]); }
        best = max(best, max(min(lo, d[k]), min(lo, d[k + 9])));
        best = max(best, -min(max(hi, d[k]), max(hi, d[k + 9])));
    }
    return best;
}
__global__ void k(const int* d, int* out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    out[i] = v_ref_pattern(d + 16 * i); out[N + i] = v_two_arrays(d + 16 * i);
    out[2 * N + i] = v_pairtrick(d + 16 * i); out[3 * N + i] = v_pairtrick_neg(d + 16 * i);
}
int main()
{
    static int hd[16 * N], ho[4 * N], ref[N];
    srand(1);
    for (int i = 0; i < 16 * N; ++i) hd[i] = rand() % 101 - 50;
    for (int i = 0; i < N; ++i) {
        int best = -255;
        for (int s = 0; s < 16; ++s) { int l = 255, h = -255; for (int j = 0; j < 9; ++j) { int e = hd[16 * i + ((s + j) & 15)]; l = std::min(l, e); h = std::max(h, e); } best = std::max(best, std::max(l, -h)); }
        ref[i] = best;
    }
    int *d, *o; cudaMalloc(&d, sizeof(hd)); cudaMalloc(&o, sizeof(ho));
    cudaMemcpy(d, hd, sizeof(hd), cudaMemcpyHostToDevice);
    k<<<N / 128, 128>>>(d, o);
    cudaMemcpy(ho, o, sizeof(ho), cudaMemcpyDeviceToHost);
    const char* names[4] = { "naive max(l,-h)", "two arrays", "pair trick d/nd", "pair trick -min(max)" };
    for (int v = 0; v < 4; ++v) { int bad = 0; for (int i = 0; i < N; ++i) bad += ho[v * N + i] != ref[i]; printf("%-24s mismatches %d / %d\n", names[v], bad, N); }
    return 0;
}
