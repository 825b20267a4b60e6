// csrc/match.cu — brute-force Hamming kNN-2 + Lowe ratio (+ optional mutual-NN cross-check).
// Replaces cv::BFMatcher(NORM_HAMMING)->knnMatch(q, t, k=2) and the ratio test of Matcher::KnnMatch
// (reference Features/matcher.cpp:55-66, :23-35); result order = (distance asc, trainIdx asc), which the
// packed key (dist << 16 | trainIdx) reproduces under unsigned min.
//
// The O(nq * nt * 256) distance matrix is the one compute-bound piece of the path (64 KB of descriptors per pair, 2.6e8 bit
// operations), so it runs on the 5th-generation tensor cores as an exact integer GEMM: every descriptor bit becomes a signed
// byte (1 -> +8, 0 -> -8), dot(a, b) over the 256 bytes = 64 * (256 - 2 * hamming(a, b)), computed by
// tcgen05.mma kind::i8 (s8 x s8 -> s32) with the 128 x 256 accumulator tile in tensor memory.
//   * CTA = 128 query rows (one TMEM lane each) x the pair's train rows in chunks of 256; 8 warps = 4 lane quarters x 2 column
//     halves; two CTAs share an SM (2 x 256 TMEM columns, 2 x 100 KB of shared memory), so one CTA's epilogue runs under the
//     other's MMAs.
//   * Operands are expanded in the kernel, straight into the no-swizzle K-major canonical layout (8-row x 16-byte core
//     matrices, LBO 128 B, SBO 2 KB): one PRMT turns 4 descriptor bits into 4 bytes.  The order of the 256 dimensions is
//     irrelevant as long as both operands use the same one.
//   * The train tile is expanded and multiplied in two halves of 128 rows: one thread issues the 8 K-steps of a half
//     (UTCIMMA 128 x 128 x 32) and commits to that half's mbarrier, so the first half is in the tensor core while the second is
//     still being expanded and the warps of column half 0 start before half 1 is done.  A 9th K step carries constants that
//     make the accumulator itself the comparable value: acc = 128 * (256 - hamming) + code, code = 64 - column-pair index.
//     Every thread then reads its own row with tcgen05.ld 32x32b.x32.pack::16b — two adjacent columns per register, 64 columns
//     per load — and works on two columns per instruction (VIMNMX.U16x2), 3 instructions per register and nothing else:
//       rows:    per 128 columns a packed running top-2, decoded into (distance << 16 | trainIdx) keys once per 128 columns;
//       columns (cross-check, quirk Q10): the low 7 bits take the row code rc = 32 - lane instead; a 31-shuffle halving butterfly
//                leaves lane L with the best (256 - hamming, row) of the warp's 32 rows for column pair L; the 4 lane quarters are
//                folded through shared memory once per chunk.
// History: a POPC/LOP3 kernel (integer-ALU bound, 0.81 ms per 511 pairs), then mma.sync IMMA s8 (0.43 ms; warp-level MMAs block
// the issue port ~6 of every 8.3 clk, tools/bmma_probe.cu), now tcgen05 (tools/umma_probe.cu pins the descriptor fields).
// Round 2 measured two restructurings against this kernel (0.255 ms per 511 pairs) and kept neither: (a) 256 rows per CTA with the
// cross-check as a second, transposed product with a top-1 epilogue instead of the butterfly — 0.263 ms: the transposed role repeats
// the operand expansion, so the ALU work per pair stays the same (~3.0e5 warp instructions) while the tensor work doubles; (b) the
// same with one CTA per SM and double-buffered column tile + accumulators (MMAs of chunk c + 1 issued before the epilogue of chunk c)
// — 0.359 ms: 16 warps per SM do not hide the TMEM-load and barrier latencies.  What ncu shows for all variants: tensor-active % +
// ALU-active % ~ 100 %, i.e. the two co-resident CTAs run in lockstep (the tensor pipe serves both CTAs' MMAs interleaved, then both
// run their epilogues), so the kernel time is T_tensor + T_alu and only less ALU work per pair makes it faster.
#include "orbf_internal.h"

namespace {

constexpr int UM_THREADS = 256, UM_ROWS = 128, UM_CHUNK = 256;
// K = 256 descriptor dimensions + one extra K step of 32 whose only non-zero bytes carry the bias and the column code (below):
// 18 K chunks of 16 bytes per row
constexpr int UM_KCHUNKS = 18, UM_KSTEPS = 9;
constexpr uint32_t UM_LBO = 128, UM_SBO = UM_KCHUNKS * 128;     // bytes: next 16-byte K chunk / next 8-row group
constexpr size_t UM_SMEM_A = (size_t)UM_ROWS * UM_KCHUNKS * 16, UM_SMEM_B = (size_t)UM_CHUNK * UM_KCHUNKS * 16, UM_SMEM_COL = 4 * (UM_CHUNK / 2) * sizeof(uint32_t);
constexpr size_t UM_SMEM = UM_SMEM_A + UM_SMEM_B + UM_SMEM_COL + 1024;   // + slack to align the operand tiles
constexpr uint32_t KEY_NONE = 0xFFFFFFFFu;
constexpr int V_SHIFT = 7;
static_assert(UM_THREADS == UM_CHUNK && UM_THREADS == 2 * UM_ROWS, "operand expansion maps one train row / half a query row to a thread");
constexpr int OPERAND_MAG = 8;                                  // both operands are +-8: a product of two equal bits is +64

__device__ __forceinline__ void top2_insert(uint32_t& m1, uint32_t& m2, uint32_t key)
{
    m2 = min(m2, max(m1, key));
    m1 = min(m1, key);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// 16 descriptor bits -> 16 signed operand bytes with PRMT: the selector nibbles are the bits themselves.  A nibble may carry its
// bit at position 0, 1 or 2 (values 0 / 1, 0 / 2, 0 / 4 all index a +MAG byte of the pool, 0 the -MAG byte), so three of the four
// bit phases need only a mask; the fourth (bit 3 would select PRMT's sign-replicate mode) comes from the word shifted by one.
// (A shared-memory table byte -> 8 bytes costs fewer instructions, but its random 8-byte reads take ~6.5 wavefronts each and
// the L1 / shared-memory data pipe is what this kernel saturates first.)
__device__ __forceinline__ uint4 expand16(uint32_t x, uint32_t xShifted1)
{
    constexpr uint32_t P = (uint32_t)OPERAND_MAG, N = (uint32_t)(256 - OPERAND_MAG);
    constexpr uint32_t poolLo = N | (P << 8) | (P << 16), poolHi = P;       // bytes 0..7: -M +M +M . +M . . .
    return make_uint4(__byte_perm(poolLo, poolHi, x & 0x1111u), __byte_perm(poolLo, poolHi, x & 0x2222u), __byte_perm(poolLo, poolHi, x & 0x4444u),
                      __byte_perm(poolLo, poolHi, xShifted1 & 0x4444u));
}

// Half a descriptor (4 words = 128 bits) -> 128 operand bytes = K chunks 8 * half .. 8 * half + 7 of `row` in the canonical
// layout.  The order of the bits within a word is irrelevant as long as both operands use the same one.
__device__ __forceinline__ void expand_half(uint8_t* tile, int row, int half, uint4 bits, bool valid)
{
    uint8_t* dst = tile + (row >> 3) * UM_SBO + (8 * half) * UM_LBO + (row & 7) * 16;
    const uint32_t w[4] = {bits.x, bits.y, bits.z, bits.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        uint4 a = make_uint4(0u, 0u, 0u, 0u), b = a;
        if (valid) {
            a = expand16(w[k], w[k] >> 1);
            b = expand16(w[k] >> 16, w[k] >> 17);
        }
        *reinterpret_cast<uint4*>(dst + (2 * k) * UM_LBO) = a;
        *reinterpret_cast<uint4*>(dst + (2 * k + 1) * UM_LBO) = b;
    }
}

// The 9th K step: query rows carry (1, 127, 127, 1), train row `col` carries (code, 127, 2, 1) with code = 64 - (col % 128) / 2, so the
// tensor core itself adds 127 * 127 + 127 * 2 + 1 = 64 * 256 (the bias that makes every accumulator positive) and the column-pair code
// that the packed top-2 carries in its low 7 bits: the epilogue neither adds nor masks anything.  Rows / columns past the end stay
// zero and produce accumulator 0, which loses against every real entry (>= 1).
__device__ __forceinline__ void bias_chunks(uint8_t* tile, int row, uint32_t word0)
{
    uint8_t* dst = tile + (row >> 3) * UM_SBO + 16 * UM_LBO + (row & 7) * 16;
    *reinterpret_cast<uint4*>(dst) = make_uint4(word0, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(dst + UM_LBO) = make_uint4(0u, 0u, 0u, 0u);
}

__device__ __forceinline__ uint64_t umma_desc(uint32_t addr)
{
    // start address, leading / stride byte offsets (all >> 4), descriptor version 1 (sm_100), layout type 0 = no swizzle
    return (uint64_t)((addr >> 4) & 0x3FFFu) | (uint64_t)(UM_LBO >> 4) << 16 | (uint64_t)(UM_SBO >> 4) << 32 | (uint64_t)1 << 46;
}

// 64 accumulator columns as 32 registers: low 16 bits of column 2j | low 16 bits of column 2j + 1 << 16 (tools/tmem_pack_probe.cu)
__device__ __forceinline__ void tmem_ld64_packed(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]),
                   "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]),
                   "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Halving butterfly: lanes with bit HALF set keep the upper half of P[0 .. 2 * HALF) and send the lower one, and vice versa.
template <int HALF>
__device__ __forceinline__ void bfly_max(uint32_t (&P)[32], int lane)
{
    const bool up = (lane & HALF) != 0;
#pragma unroll
    for (int i = 0; i < HALF; ++i) {
        const uint32_t send = up ? P[i] : P[i + HALF], keep = up ? P[i + HALF] : P[i];
        P[i] = __vmaxu2(keep, __shfl_xor_sync(0xffffffffu, send, HALF));
    }
}

// Cross-check: fold one chunk's per-warp column maxima over the 4 lane quarters; thread `pairIdx` owns two columns.
__device__ __forceinline__ void fold_columns(const uint32_t* sCol, uint32_t* revChunk, int cn, int qBase, int pairIdx)
{
    const uint32_t both[4] = {sCol[pairIdx], sCol[UM_CHUNK / 2 + pairIdx], sCol[2 * (UM_CHUNK / 2) + pairIdx], sCol[3 * (UM_CHUNK / 2) + pairIdx]};
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
        uint32_t best = 0u;                                    // h << 9 | (3 - quarter) << 7 | rc: highest h, then lowest row
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const uint32_t v = (both[w] >> (hf * 16)) & 0xFFFFu;
            if (v & 127u) best = max(best, ((v >> V_SHIFT) << 9) | ((uint32_t)(3 - w) << V_SHIFT) | (v & 127u));
        }
        const int col = 2 * pairIdx + hf;
        if (best && col < cn) {
            const uint32_t r = (uint32_t)qBase + (3u - ((best >> V_SHIFT) & 3u)) * 32u + 32u - (best & 127u);
            atomicMin(&revChunk[col], (256u - (best >> 9)) * 65536u + r);
        }
    }
}

// train side of pair slot ts: rows + row count, from the single store or from the shard (possibly a peer GPU's memory) that owns it
__device__ __forceinline__ const uint8_t* train_rows(const MatchSet& ms, int ts, int& nt)
{
    if (ms.tShards) {
        const int sh = ts / ms.shardKf, local = ts - sh * ms.shardKf;
        nt = ms.tShardCounts[sh][local];
        return ms.tShards[sh] + (long long)local * ms.tStride;
    }
    nt = ms.tCounts ? ms.tCounts[ts] : ms.nt;
    return ms.tdesc + (long long)ts * ms.tStride;
}

template <bool CROSS>
__global__ void __launch_bounds__(UM_THREADS, 2) knn2_kernel(MatchSet ms, int K)
{
    extern __shared__ uint8_t umSmemRaw[];
    __shared__ uint32_t sTmem;
    __shared__ __align__(8) uint64_t sBar[2];
    uint32_t* __restrict__ knn = ms.knn;
    uint32_t* __restrict__ rev = ms.rev;
    const int pair = ms.pair0 + blockIdx.y;
    int qs = 0, ts = 0;
    if (ms.pairs) { qs = ms.pairs[2 * pair]; ts = ms.pairs[2 * pair + 1]; }
    int nt;
    const uint32_t* T = reinterpret_cast<const uint32_t*>(train_rows(ms, ts, nt));
    const int nq = ms.qCounts ? ms.qCounts[qs] : ms.nq;
    const int qBase = blockIdx.x * UM_ROWS;
    if (qBase >= nq) return;
    const uint32_t* Q = reinterpret_cast<const uint32_t*>(ms.qdesc + (long long)qs * ms.qStride);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int quarter = warp & 3, colHalf = warp >> 2;         // TMEM lanes 32 * quarter .. + 31, chunk columns 128 * colHalf .. + 127
    uint8_t* sA = umSmemRaw + ((1024u - (smem_u32(umSmemRaw) & 1023u)) & 1023u);
    uint8_t* sB = sA + UM_SMEM_A;
    uint32_t* sCol = reinterpret_cast<uint32_t*>(sB + UM_SMEM_B);           // [4 quarters][UM_CHUNK / 2] packed per-warp column maxima

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sTmem)), "n"(UM_CHUNK));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&sBar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&sBar[1])));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    {
        const int r = tid & (UM_ROWS - 1), half = tid >> 7;
        const bool valid = qBase + r < nq;
        expand_half(sA, r, half, valid ? __ldg(reinterpret_cast<const uint4*>(Q + (long long)(qBase + r) * 8) + half) : make_uint4(0u, 0u, 0u, 0u), valid);
        if (half == 0) bias_chunks(sA, r, valid ? (1u | (127u << 8) | (127u << 16) | (1u << 24)) : 0u);
        // the train tile's 9th K step depends on the column only: written once for full chunks (a short last chunk rewrites it)
        bias_chunks(sB, tid, (64u - (uint32_t)((tid & 127) >> 1)) | (127u << 8) | (2u << 16) | (1u << 24));
    }
    __syncthreads();        // a short first chunk rewrites these column codes from OTHER threads (below): order the two writes
    // this thread's query row = TMEM lane 32 * quarter + lane; two threads (colHalf 0 / 1) share a row
    const int row = qBase + quarter * 32 + lane;
    const bool rowValid = row < nq;
    // cross-check: the row code replaces the column code (one LOP3 per register; rows past nq become 0 and never win)
    const uint32_t rowCode = rowValid ? (32u - (uint32_t)lane) * 0x10001u : 0u, rowMask = rowValid ? 0xFF80FF80u : 0u;
    uint32_t k1 = KEY_NONE, k2 = KEY_NONE;
    // instruction descriptor: D s32 (2 @ bit 4), A / B signed 8-bit (1 @ 7, 1 @ 10), both K-major, N >> 3 @ 17, M >> 4 @ 24
    const uint32_t idescHalf = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((UM_CHUNK / 2) >> 3) << 17) | ((uint32_t)(UM_ROWS >> 4) << 24);
    uint32_t phase = 0;

    // Per chunk the train tile is expanded and multiplied in two halves of 128 rows: the MMAs of half 0 run under the expansion of
    // half 1, and the warps of column half 0 start their epilogue while half 1 is still in the tensor core.
    // The train rows are fetched ahead of their expansion — one half-tile ahead with the cross-check (the kernel sits at the 128-register
    // cap), a whole chunk ahead without it: the loads are in flight during the expansion, MMAs and epilogue of the tiles before, which
    // hides the L2 latency of a local store and most of the NVLink latency when the rows live in a peer GPU's HBM
    // (orbf_kfdb_attach_peers: the transfer then overlaps the math tile by tile).
    constexpr int PD = CROSS ? 1 : 2;                                    // prefetch distance in half-tiles (4 was measured: no better than 2)
    constexpr int UNR = PD == 4 ? 2 : 1;                                 // chunks per trip of the loop: the PD register slots are indexed statically
    const int pcol = tid & (UM_CHUNK / 2 - 1), phalf = tid >> 7;
    uint4 nextBits[PD];
#pragma unroll
    for (int d = 0; d < PD; ++d) {
        const int r0 = d * (UM_CHUNK / 2) + pcol;
        nextBits[d] = r0 < nt ? __ldg(reinterpret_cast<const uint4*>(T + (long long)r0 * 8) + phalf) : make_uint4(0u, 0u, 0u, 0u);
    }
    for (int cc = 0; cc < nt; cc += UNR * UM_CHUNK) {
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int c0 = cc + u * UM_CHUNK;
        if (c0 >= nt) break;
        const int cn = min(UM_CHUNK, nt - c0);
        uint32_t tmem = 0;
#pragma unroll
        for (int part = 0; part < 2; ++part) {
            // every thread that gets here has seen the MMAs that read this half of the tile complete (see the syncs below)
            {
                const int col = part * (UM_CHUNK / 2) + pcol, half = phalf;
                const bool valid = col < cn;
                const uint4 bits = nextBits[(2 * u + part) % PD];
                const int rowNext = c0 + (part + PD) * (UM_CHUNK / 2) + pcol;       // the same column PD half-tiles further on
                if (rowNext < nt) nextBits[(2 * u + part) % PD] = __ldg(reinterpret_cast<const uint4*>(T + (long long)rowNext * 8) + half);
                expand_half(sB, col, half, bits, valid);
                if (half == 1 && cn < UM_CHUNK) bias_chunks(sB, col, valid ? ((64u - (uint32_t)((col & 127) >> 1)) | (127u << 8) | (2u << 16) | (1u << 24)) : 0u);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy stores -> the tensor core's async proxy
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");   // orders the previous chunk's tcgen05.ld before the MMAs
            __syncthreads();
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            tmem = sTmem;
            if (tid == 0) {
#pragma unroll
                for (int ks = 0; ks < UM_KSTEPS; ++ks) {
                    const uint64_t da = umma_desc(smem_u32(sA) + ks * 2 * UM_LBO);
                    const uint64_t db = umma_desc(smem_u32(sB) + part * (UM_CHUNK / 2 / 8) * UM_SBO + ks * 2 * UM_LBO);
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                                 ::"r"(tmem + part * (UM_CHUNK / 2)), "l"(da), "l"(db), "r"(idescHalf), "r"((uint32_t)(ks > 0)) : "memory");
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&sBar[part])) : "memory");
            }
            if (CROSS && part == 0 && c0 > 0 && tid < UM_CHUNK / 2) fold_columns(sCol, rev + (long long)pair * K + (c0 - UM_CHUNK), UM_CHUNK, qBase, tid);
        }
        {
            uint32_t done = 0, spins = 0;
            while (!done) {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(smem_u32(&sBar[colHalf])), "r"(phase) : "memory");
                if (!done && ++spins > (1u << 22)) __trap();                   // a lost commit must fail loudly, not hang the device
            }
            phase ^= 1u;
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

        // this warp's 32 rows x 128 columns = 64 column pairs: two batches of 32 pairs, each one packed TMEM load.  A register holds
        // 128 * (256 - hamming) + code for two adjacent columns (bias and code come out of the 9th K step), ready for the packed top-2
        const uint32_t tbase = tmem + ((uint32_t)(quarter * 32) << 16) + colHalf * 128;
        uint32_t b1 = 0u, b2 = 0u;                             // packed running top-2 of the even / odd column streams
#pragma unroll
        for (int bt = 0; bt < 2; ++bt) {
            const int colBase = colHalf * 128 + bt * 64;
            uint32_t P[32];
            tmem_ld64_packed(tbase + bt * 64, P);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                b2 = __vmaxu2(b2, __vminu2(b1, P[j]));
                b1 = __vmaxu2(b1, P[j]);
            }
            if (CROSS) {
                // columns: the low 7 bits take the row code instead (0 for rows past nq, which then never win: see fold_columns)
#pragma unroll
                for (int j = 0; j < 32; ++j) P[j] = (P[j] & rowMask) | rowCode;
                bfly_max<16>(P, lane); bfly_max<8>(P, lane); bfly_max<4>(P, lane); bfly_max<2>(P, lane); bfly_max<1>(P, lane);
                sCol[quarter * (UM_CHUNK / 2) + (colBase >> 1) + lane] = P[0];
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t e = ((k & 2) ? b2 : b1) >> ((k & 1) * 16) & 0xFFFFu;
            const uint32_t col = (uint32_t)(c0 + colHalf * 128) + (64u - (e & 127u)) * 2u + (k & 1);
            if (e && (int)col < nt) top2_insert(k1, k2, ((256u - (e >> V_SHIFT)) << 16) | col);
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CROSS && nt > 0 && tid < UM_CHUNK / 2) {
        const int cLast = ((nt - 1) / UM_CHUNK) * UM_CHUNK;
        fold_columns(sCol, rev + (long long)pair * K + cLast, nt - cLast, qBase, tid);
    }
    __syncthreads();
    // the two threads of a row merge through shared memory (the column staging area is free now)
    uint2* sMerge = reinterpret_cast<uint2*>(sCol);
    if (colHalf == 1) sMerge[quarter * 32 + lane] = make_uint2(k1, k2);
    __syncthreads();
    if (colHalf == 0 && rowValid) {
        const uint2 o = sMerge[quarter * 32 + lane];
        top2_insert(k1, k2, o.x);
        top2_insert(k1, k2, o.y);
        *reinterpret_cast<uint2*>(knn + ((long long)pair * K + row) * 2) = make_uint2(k1, k2);
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(sTmem), "n"(UM_CHUNK));
}

constexpr int MS_THREADS = 256;

__global__ void __launch_bounds__(MS_THREADS) match_select_kernel(MatchSet ms, int K, float ratio, int cross)
{
    const uint32_t* __restrict__ knn = ms.knn;
    const uint32_t* __restrict__ rev = ms.rev;
    orbf_dmatch* __restrict__ matches = ms.matches;
    int* __restrict__ matchCount = ms.matchCount;
    __shared__ int sWarp[MS_THREADS / 32];
    __shared__ int sBase;
    const int pair = ms.pair0 + blockIdx.x;
    int qs = 0, ts = 0;
    if (ms.pairs) { qs = ms.pairs[2 * pair]; ts = ms.pairs[2 * pair + 1]; }
    int nt;
    (void)train_rows(ms, ts, nt);
    const int nq = ms.qCounts ? ms.qCounts[qs] : ms.nq;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) sBase = 0;
    __syncthreads();
    orbf_dmatch* out = matches + (long long)pair * K;
    for (int base = 0; base < nq; base += MS_THREADS) {
        const int i = base + threadIdx.x;
        bool keep = false;
        uint32_t k1 = 0;
        if (i < nq && nt >= 2) {
            const uint2 kk = *reinterpret_cast<const uint2*>(knn + ((long long)pair * K + i) * 2);
            k1 = kk.x;
            if (kk.y != KEY_NONE) {
                const float d1 = (float)(kk.x >> 16), d2 = (float)(kk.y >> 16);
                keep = d1 < __fmul_rn(ratio, d2);                       // m1.distance < mfNNratio * m2.distance (float)
                if (keep && cross) keep = (rev[(long long)pair * K + (kk.x & 0xFFFFu)] & 0xFFFFu) == (uint32_t)i;
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) sWarp[warp] = __popc(m);
        __syncthreads();
        int off = sBase;
        for (int w = 0; w < warp; ++w) off += sWarp[w];
        if (keep && matches) {
            orbf_dmatch dm;
            dm.queryIdx = i; dm.trainIdx = (int)(k1 & 0xFFFFu); dm.imgIdx = 0; dm.distance = (float)(k1 >> 16);
            out[off + __popc(m & ((1u << lane) - 1))] = dm;
        }
        __syncthreads();
        if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < MS_THREADS / 32; ++w) t += sWarp[w]; sBase += t; }
        __syncthreads();
    }
    if (threadIdx.x == 0) matchCount[pair] = sBase;
}

// ---- Landmark::ComputeDistinctiveDescriptors (Core/landmark.cpp:219-273): one warp per landmark ------------------------------
// Lane j holds observation j's descriptor (chunks of 32 observations); for every row i the distances to all observations are
// formed with the same XOR/POPC arithmetic, the row median sorted[(size_t)(0.5 * (N - 1))] is found by rank counting over
// shuffles, and the first row with the strictly smallest median wins.
constexpr int DD_MAX_OBS = 128;

__global__ void __launch_bounds__(128) distinctive_kernel(const uint8_t* __restrict__ desc, const int* __restrict__ offsets, int nLandmarks,
    int* __restrict__ best, int* __restrict__ bestMedian)
{
    __shared__ uint16_t sDist[4][DD_MAX_OBS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int l = blockIdx.x * 4 + warp;
    if (l >= nLandmarks) return;
    const int a = offsets[l], N = min(offsets[l + 1] - a, DD_MAX_OBS);
    if (N <= 0) { if (lane == 0) { best[l] = -1; if (bestMedian) bestMedian[l] = -1; } return; }
    const int m = (int)(0.5 * (double)(N - 1));
    int bestIdx = 0, bestMed = 0x7fffffff;
    uint16_t* dist = sDist[warp];
    for (int i = 0; i < N; ++i) {
        const uint4 ia = __ldg(reinterpret_cast<const uint4*>(desc + (size_t)(a + i) * 32)), ib = __ldg(reinterpret_cast<const uint4*>(desc + (size_t)(a + i) * 32) + 1);
        for (int j = lane; j < N; j += 32) {
            const uint4 ja = __ldg(reinterpret_cast<const uint4*>(desc + (size_t)(a + j) * 32)), jb = __ldg(reinterpret_cast<const uint4*>(desc + (size_t)(a + j) * 32) + 1);
            dist[j] = (uint16_t)(__popc(ia.x ^ ja.x) + __popc(ia.y ^ ja.y) + __popc(ia.z ^ ja.z) + __popc(ia.w ^ ja.w)
                + __popc(ib.x ^ jb.x) + __popc(ib.y ^ jb.y) + __popc(ib.z ^ jb.z) + __popc(ib.w ^ jb.w));
        }
        __syncwarp();
        // the m-th smallest: the value v with #{< v} <= m < #{<= v}
        int med = -1;
        for (int j = lane; j < N; j += 32) {
            const int v = dist[j];
            int less = 0, leq = 0;
            for (int k = 0; k < N; ++k) { const int u = dist[k]; less += u < v; leq += u <= v; }
            if (less <= m && m < leq) med = v;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) med = max(med, __shfl_xor_sync(0xffffffffu, med, o));
        if (med < bestMed) { bestMed = med; bestIdx = i; }
        __syncwarp();
    }
    if (lane == 0) { best[l] = bestIdx; if (bestMedian) bestMedian[l] = bestMed; }
}

}  // namespace

int orbf_launch_distinctive(orbf_context* c, const uint8_t* d_desc, const int* d_offsets, int nLandmarks, int* d_best, int* d_median)
{
    if (nLandmarks <= 0) return ORBF_OK;
    distinctive_kernel<<<(nLandmarks + 3) / 4, 128, 0, c->stream>>>(d_desc, d_offsets, nLandmarks, d_best, d_median);
    ORBF_LAUNCH_CHECK(c);
    return ORBF_OK;
}

int orbf_launch_knn2(orbf_context* c, const MatchSet& ms, int npairs, bool cross)
{
    const int maxNq = ms.qCounts ? c->K : ms.nq;
    if (maxNq <= 0 || npairs <= 0) return ORBF_OK;
    dim3 grid((maxNq + UM_ROWS - 1) / UM_ROWS, npairs);
    if (cross) ORBF_CUDA(c, cudaFuncSetAttribute(knn2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UM_SMEM));
    else ORBF_CUDA(c, cudaFuncSetAttribute(knn2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UM_SMEM));
    if (cross) {
        ORBF_CUDA(c, cudaMemsetAsync(ms.rev + (size_t)ms.pair0 * c->K, 0xFF, (size_t)npairs * c->K * sizeof(uint32_t), c->stream));
        knn2_kernel<true><<<grid, UM_THREADS, UM_SMEM, c->stream>>>(ms, c->K);
    } else knn2_kernel<false><<<grid, UM_THREADS, UM_SMEM, c->stream>>>(ms, c->K);
    ORBF_LAUNCH_CHECK(c);
    return ORBF_OK;
}

int orbf_launch_match_select(orbf_context* c, const MatchSet& ms, int npairs, float ratio, bool cross)
{
    if (npairs <= 0) return ORBF_OK;
    match_select_kernel<<<npairs, MS_THREADS, 0, c->stream>>>(ms, c->K, ratio, cross ? 1 : 0);
    ORBF_LAUNCH_CHECK(c);
    return ORBF_OK;
}
