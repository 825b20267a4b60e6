// tools/tmem_pack_probe.cu — what tcgen05.ld ... .pack::16b returns: writes v(lane, col) = 100 * lane + col (+ 0x50000 to show that the
// upper halves are dropped) to 64 TMEM columns with tcgen05.st, reads them back with 32x32b.x32.pack::16b and prints lane 3.
// nvcc -gencode arch=compute_100a,code=sm_100a -o tools/tmem_pack_probe.bin tools/tmem_pack_probe.cu && tools/tmem_pack_probe.bin
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
__global__ void probe(uint32_t* out)
{
    __shared__ uint32_t sT;
    const int lane = threadIdx.x;
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"((uint32_t)__cvta_generic_to_shared(&sT)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    __syncwarp();
    const uint32_t t = sT;
    for (int c = 0; c < 64; ++c) {
        const uint32_t v = 100u * lane + c + 0x50000u;
        asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(t + c), "r"(v) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
                   "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),
                   "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(t));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; ++j) out[lane * 32 + j] = r[j];
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(t));
}
int main()
{
    uint32_t* d; cudaMalloc(&d, 32 * 32 * 4);
    probe<<<1, 32>>>(d);
    uint32_t h[32 * 32];
    cudaError_t e = cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("status %s\n", cudaGetErrorString(e));
    for (int j = 0; j < 32; ++j) printf("lane3 r[%d] = lo %u hi %u\n", j, h[3 * 32 + j] & 0xFFFF, h[3 * 32 + j] >> 16);
    return 0;
}
