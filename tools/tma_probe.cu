// tools/tma_probe.cu — isolates the TMA box load used by fast.cu (authoring aid).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../adaptive-rgbd-localization-mappig_b200/csrc/tma.h"

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct Params { CUtensorMap maps[4]; int level, x, y, z, BW, BH; uint8_t* out; };

__global__ void probe_direct(const __grid_constant__ CUtensorMap map, int x, int y, int z, int BW, int BH, uint8_t* out)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) mbar_init(&bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) { mbar_expect_tx(&bar, BW * BH); tma_load_3d(smem, &map, x, y, z, &bar); }
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < BW * BH; i += blockDim.x) out[i] = smem[i];
}

__global__ void probe_mbar_only(uint8_t* out)
{
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) mbar_init(&bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar)) : "memory");
    mbar_wait(&bar, 0);
    out[threadIdx.x] = 7;
}

__global__ void probe_bulk1d(const uint8_t* src, uint8_t* out, int bytes)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) mbar_init(&bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar, bytes);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem)), "l"(src),
                     "r"(bytes), "r"(smem_u32(&bar)) : "memory");
    }
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = smem[i];
}

__global__ void probe_indexed(const __grid_constant__ Params P)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) mbar_init(&bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) { mbar_expect_tx(&bar, P.BW * P.BH); tma_load_3d(smem, &P.maps[P.level], P.x, P.y, P.z, &bar); }
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < P.BW * P.BH; i += blockDim.x) P.out[i] = smem[i];
}

int main(int argc, char** argv)
{
    const int variant = argc > 1 ? atoi(argv[1]) : 0;
    const int w = 640, h = 480, pitch = 640, frames = 3, BW = 144, BH = 38;
    std::vector<uint8_t> img((size_t)pitch * h * frames);
    for (size_t i = 0; i < img.size(); ++i) img[i] = (uint8_t)((i * 2654435761u) >> 24);
    uint8_t *d, *dout;
    cudaMalloc(&d, img.size()); cudaMalloc(&dout, BW * BH);
    cudaMemcpy(d, img.data(), img.size(), cudaMemcpyHostToDevice);
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    printf("entry point: %d %d %p\n", (int)e, (int)q, p);
    EncodeTiledFn fn = (EncodeTiledFn)p;
    Params P;
    for (int l = 0; l < 4; ++l) {
        const cuuint64_t gdim[3] = { (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)frames };
        const cuuint64_t gstride[2] = { (cuuint64_t)pitch, (cuuint64_t)pitch * h };
        const cuuint32_t box[3] = { (cuuint32_t)BW, (cuuint32_t)BH, 1u };
        const cuuint32_t estr[3] = { 1u, 1u, 1u };
        CUresult r = fn(&P.maps[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode %d -> %d\n", l, (int)r);
    }
    std::vector<uint8_t> out(BW * BH);
    auto check = [&](const char* name, int x, int y, int z) {
        cudaError_t s = cudaDeviceSynchronize();
        cudaMemcpy(out.data(), dout, out.size(), cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int r = 0; r < BH; ++r) for (int c = 0; c < BW; ++c) {
            const int gx = x + c, gy = y + r;
            const uint8_t ref = (gx >= 0 && gx < w && gy >= 0 && gy < h) ? img[((size_t)z * h + gy) * pitch + gx] : 0;
            bad += out[r * BW + c] != ref;
        }
        printf("%s (%d,%d,%d): sync=%s mismatches=%d\n", name, x, y, z, cudaGetErrorString(s), bad);
    };
    if (variant == 0) { probe_mbar_only<<<1, 128>>>(dout); printf("mbar only: %s\n", cudaGetErrorString(cudaDeviceSynchronize())); }
    if (variant == 1) { probe_bulk1d<<<1, 128, 8192>>>(d, dout, 4096); printf("bulk1d: %s\n", cudaGetErrorString(cudaDeviceSynchronize())); }
    if (variant == 2) { probe_direct<<<1, 128, BW * BH + 256>>>(P.maps[0], 15, 16, 1, BW, BH, dout); check("direct", 15, 16, 1); }
    if (variant == 3) { probe_direct<<<1, 128, BW * BH + 256>>>(P.maps[0], 600, 470, 2, BW, BH, dout); check("direct-oob", 600, 470, 2); }
    if (variant == 4) {
        P.level = 2; P.x = 139; P.y = 48; P.z = 0; P.BW = BW; P.BH = BH; P.out = dout;
        probe_indexed<<<1, 128, BW * BH + 256>>>(P); check("indexed", 139, 48, 0);
    }
    if (variant == 5) {   // 128-byte inner box
        const cuuint64_t gdim[3] = { (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)frames };
        const cuuint64_t gstride[2] = { (cuuint64_t)pitch, (cuuint64_t)pitch * h };
        const unsigned bw = argc > 2 ? atoi(argv[2]) : 128, promo = argc > 3 ? atoi(argv[3]) : 0, bh = argc > 4 ? atoi(argv[4]) : 32;
        const cuuint32_t box[3] = { bw, bh, 1u };
        const cuuint32_t estr[3] = { 1u, 1u, 1u };
        CUtensorMap m;
        CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode128 -> %d\n", (int)r);
        cudaFree(dout); cudaMalloc(&dout, bw * bh); out.resize(bw * bh);
        probe_direct<<<1, 128, bw * bh + 256>>>(m, 15, 16, 1, bw, bh, dout);
        cudaError_t s2 = cudaDeviceSynchronize();
        cudaMemcpy(out.data(), dout, out.size(), cudaMemcpyDeviceToHost);
        int bad = 0;
        for (unsigned r2 = 0; r2 < bh; ++r2) for (unsigned c2 = 0; c2 < bw; ++c2) bad += out[r2 * bw + c2] != img[((size_t)1 * h + 16 + r2) * pitch + 15 + c2];
        printf("box %ux%u promo %u: %s mismatches=%d\n", bw, bh, promo, cudaGetErrorString(s2), bad);
    }
    return 0;
}
