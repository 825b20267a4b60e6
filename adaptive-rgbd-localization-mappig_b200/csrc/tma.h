// csrc/tma.h — Tensor Memory Accelerator plumbing shared by the tile kernels (fast.cu, pyramid.cu).
// Host: cuTensorMapEncodeTiled (looked up through the runtime, no -lcuda) for [frames][rows][pitch] u8 planes.
// Device: mbarrier + cp.async.bulk.tensor.3d (SASS: UTMALDG) — one elected thread fetches a whole (tile + halo)
// box into shared memory; out-of-range elements arrive as zeros, so halo handling costs no instructions.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

struct orbf_context;

// Encodes a rank-3 u8 map: dims (w, h, frames), strides (pitch, frameStride) bytes, box (boxW, boxH, 1).
// boxW must be a multiple of 16 and <= 256, boxH <= 256; base 16-byte aligned, strides multiples of 16, and the x
// coordinate of every box load a multiple of 16 (a misaligned start traps as 'illegal instruction', tools/tma_probe.cu).
// swizzle64: CU_TENSOR_MAP_SWIZZLE_64B (box width must be 64 bytes, destination 512-byte aligned): the 16-byte chunk index of a row
// is XORed with bits 7..8 of the shared-memory address, i.e. with (row >> 1) & 3 — rows of one parity no longer share their banks.
int orbf_tma_encode_u8(orbf_context* ctx, CUtensorMap* out, const void* base, int w, int h, int frames, long long pitch,
    long long frameStride, int boxW, int boxH, bool swizzle64 = false);

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// box of `map` whose first element is (x, y, z) -> dst (128-byte aligned shared memory); completes on `bar`
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int x, int y, int z, uint64_t* bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                     smem_u32(dst)),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
                 : "memory");
}
#endif
